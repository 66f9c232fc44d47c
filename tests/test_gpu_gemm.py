"""GPU: the tcgen05 split GEMM against a float64 torch reference under both operand policies: fp16 pairs (3 MMAs per
product, fp32-class: 1e-5 of the output scale) and fp16 + fp8 corrections (2 MMA units per product, 2^-14.5 per product:
1e-4 of the output scale).  The parity bar of the path is 1e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k,bias,res,outs", [
    (128, 128, 64, False, False, "f"), (128, 128, 128, True, False, "f"), (256, 256, 320, True, True, "fs"),
    (300, 320, 320, True, False, "fs"), (1000, 1280, 640, False, True, "f"), (4096, 2560, 320, True, False, "s"),
    (77, 640, 1024, False, False, "fs"), (1, 8, 8, True, False, "f"), (28 * 1024, 640, 640, True, True, "f"),
    (513, 5120, 1280, True, False, "f"),
    # CTA-pair form (>= 148 work items) with an ODD number of M tiles and a ragged last tile: the second CTA of the last pair
    # works on a tile that does not exist (zero-filled operands, no stores)
    (9500, 1280, 320, True, True, "fs"), (128 * 75, 640, 640, False, False, "f"),
])
def test_gemm_split_matches_fp64(cuda, operand_mode, m, n, k, bias, res, outs):
    from vidseg_diffusion_b200.linear import gemm_split, split
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).to(cuda)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda)
    b = torch.randn(n, generator=g).to(cuda) if bias else None
    r = torch.randn(m, n, generator=g).to(cuda) if res else None
    want = a.double() @ w.double().T
    if bias:
        want = want + b.double()
    if res:
        want = want + r.double()
    out, sp = gemm_split(split(a), split(w, 256.0, is_weight=True), b, r, want_f32="f" in outs, want_split="s" in outs)
    tol = 1e-4 if (operand_mode == 1 and k % 64 == 0) else 1e-5
    scale = want.abs().max().item()
    if out is not None:
        err = (out.double() - want).abs().max().item() / scale
        assert err < tol, f"fp32 output rel err {err:.3e}"
    if sp is not None:
        err = (sp.float().double() - want).abs().max().item() / scale
        assert err < max(tol, 1e-4 if sp.fmt == "packed8" else 0), f"split output ({sp.fmt}) rel err {err:.3e}"


def test_split_roundtrip(cuda):
    from vidseg_diffusion_b200.linear import split
    x = torch.randn(1000003, device=cuda) * 3
    s = split(x)
    # 22 bits for |x| >= 0.25, absolute error <= 2^-25 below (fp16 subnormal residual)
    assert ((s.float() - x).abs() <= 2.0 ** -21 * x.abs() + 3.1e-8).all()


@pytest.mark.parametrize("is_weight", [False, True])
def test_packed8_operand_layout(cuda, lib, is_weight):
    """The fp8 side tensor, byte for byte: per 64-element block 64 bytes e5m2((x*scale - hi) * sl) then 64 bytes
    e4m3(x*scale * sx), (sx, sl) = (1, 16) for activations and (1/16, 1) for weights."""
    from vidseg_diffusion_b200.linear import split
    lib.vidseg_set_operand_mode(1)
    x = torch.randn(37, 192, generator=torch.Generator().manual_seed(5)).to(cuda) * (0.05 if is_weight else 2.0)
    scale = 256.0 if is_weight else 1.0
    s = split(x, scale, is_weight=is_weight)
    assert s.fmt == "packed8"
    xs = x * scale
    hi = xs.half()
    assert torch.equal(s.hi, hi)
    sx, sl = (1.0 / 16, 1.0) if is_weight else (1.0, 16.0)
    aux = s.lo.view(torch.uint8).reshape(37, 3, 128)
    want_lo = ((xs - hi.float()) * sl).to(torch.float8_e5m2).view(torch.uint8).reshape(37, 3, 64)
    want_x = (xs * sx).to(torch.float8_e4m3fn).view(torch.uint8).reshape(37, 3, 64)
    assert torch.equal(aux[..., :64], want_lo) and torch.equal(aux[..., 64:], want_x)
    assert ((s.float() - x).abs() <= 2.0 ** -13 * x.abs() + 1e-6).all()


@pytest.mark.parametrize("m,d,k", [(300, 1280, 320), (1000, 256, 64), (129, 32, 96), (4096, 2560, 640)])
def test_fused_geglu_projection(cuda, operand_mode, m, d, k):
    """GEGLU.proj + value * gelu(gate) in one kernel (weight rows permuted into 32 value | 32 gate groups) against
    the float64 definition (attention.py:89-96)."""
    import torch.nn as nn
    import torch.nn.functional as F
    from vidseg_diffusion_b200 import kernels as K
    from vidseg_diffusion_b200.linear import split
    g = torch.Generator(device="cpu").manual_seed(m + d + k)
    proj = nn.Linear(k, 2 * d).to(cuda)
    with torch.no_grad():
        proj.weight.copy_(torch.randn(2 * d, k, generator=g) / k ** 0.5)
        proj.bias.copy_(torch.randn(2 * d, generator=g))
    x = torch.randn(m, k, generator=g).to(cuda)
    val, gate = (x.double() @ proj.weight.double().T + proj.bias.double()).chunk(2, dim=-1)
    want = val * F.gelu(gate)
    got = K.linear_geglu(split(x), proj)
    assert got.hi.shape == (m, d)
    err = (got.float().double() - want).abs().max().item() / want.abs().max().item()
    packed_in = operand_mode == 1 and k % 64 == 0
    assert err < (2e-4 if (packed_in or got.fmt == "packed8") else 1e-5), f"rel err {err:.3e} ({got.fmt})"


@pytest.mark.parametrize("m,c,k", [(300, 320, 320), (28 * 64, 1280, 1280), (1000, 640, 640), (77, 128, 64)])
def test_stacked_projections_equal_separate_gemms(cuda, operand_mode, m, c, k):
    """to_q | to_k | to_v as ONE GEMM with three output sets: bit-identical to three separate GEMMs (same tiles, same
    order of accumulation), unwanted outputs skipped."""
    from vidseg_diffusion_b200.linear import gemm_split, gemm_split_seg, split
    g = torch.Generator(device="cpu").manual_seed(m + c + k)
    a = split(torch.randn(m, k, generator=g).to(cuda))
    ws = [(torch.randn(c, k, generator=g) / k ** 0.5).to(cuda) for _ in range(3)]
    stacked = split(torch.cat(ws, 0).contiguous(), 256.0, is_weight=True)
    got = gemm_split_seg(a, stacked, 3, (True, False, True), (True, True, False))
    assert got[1][0] is None and got[2][1] is None
    for s, w in enumerate(ws):
        f, sp = gemm_split(a, split(w, 256.0, is_weight=True), want_f32=True, want_split=True, split_pair16=True)
        if got[s][0] is not None:
            assert torch.equal(got[s][0], f), f"segment {s}: fp32 output differs"
        if got[s][1] is not None:
            assert got[s][1].fmt == "pair16"
            assert torch.equal(got[s][1].hi, sp.hi) and torch.equal(got[s][1].lo, sp.lo), f"segment {s}: operand output differs"
