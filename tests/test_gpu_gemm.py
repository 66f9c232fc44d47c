"""GPU: the tcgen05 split-fp16 GEMM against a float64 torch reference (tolerance: fp32-class, 1e-5
of the output scale; the parity bar of the path is 1e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k,bias,res,outs", [
    (128, 128, 64, False, False, "f"), (128, 128, 128, True, False, "f"), (256, 256, 320, True, True, "fs"),
    (300, 320, 320, True, False, "fs"), (1000, 1280, 640, False, True, "f"), (4096, 2560, 320, True, False, "s"),
    (77, 640, 1024, False, False, "fs"), (1, 8, 8, True, False, "f"), (28 * 1024, 640, 640, True, True, "f"),
    (513, 5120, 1280, True, False, "f"),
])
def test_gemm_split_matches_fp64(cuda, m, n, k, bias, res, outs):
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
    out, sp = gemm_split(split(a), split(w), b, r, want_f32="f" in outs, want_split="s" in outs)
    scale = want.abs().max().item()
    if out is not None:
        err = (out.double() - want).abs().max().item() / scale
        assert err < 1e-5, f"fp32 output rel err {err:.3e}"
    if sp is not None:
        err = (sp.float().double() - want).abs().max().item() / scale
        assert err < 1e-5, f"split output rel err {err:.3e}"


def test_split_roundtrip(cuda):
    from vidseg_diffusion_b200.linear import split
    x = torch.randn(1000003, device=cuda) * 3
    s = split(x)
    # 22 bits for |x| >= 0.25, absolute error <= 2^-25 below (fp16 subnormal residual)
    assert ((s.float() - x).abs() <= 2.0 ** -21 * x.abs() + 3.1e-8).all()
