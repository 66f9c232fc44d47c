"""GPU: the implicit-GEMM convolution and the streaming normalisation kernels against float64 torch
references (tolerance: fp32-class; the parity bar of the path is 1e-3)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(got, want):
    return float((got.double() - want).abs().max() / want.abs().max())


@pytest.mark.parametrize("b,h,w,cin,cout,k,stride,extras", [
    (2, 16, 16, 64, 128, 3, 1, ""), (3, 8, 8, 320, 320, 3, 1, "cr"), (1, 64, 64, 64, 320, 3, 1, "c"),
    (2, 12, 20, 72, 40, 3, 1, "r"), (5, 4, 4, 128, 64, 3, 1, ""), (2, 32, 32, 128, 256, 3, 2, ""),
    (3, 6, 10, 64, 64, 3, 2, ""), (2, 16, 16, 192, 64, 1, 1, "r"), (2, 16, 16, 4, 320, 3, 1, ""),
    (2, 16, 16, 64, 4, 3, 1, ""), (28, 8, 8, 256, 256, 3, 1, "c"), (1, 96, 96, 64, 160, 3, 1, ""),
])
def test_conv2d_matches_fp64(cuda, operand_mode, b, h, w, cin, cout, k, stride, extras):
    from vidseg_diffusion_b200 import kernels as K
    g = torch.Generator(device="cpu").manual_seed(b * 1000 + h * 10 + cin + cout + k + stride)
    conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2).to(cuda)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (cin * k * k) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g))
    x = torch.randn(b, cin, h, w, generator=g).to(cuda)
    cb = torch.randn(b, cout, generator=g).to(cuda) if "c" in extras else None
    res = torch.randn(b, cout, h // stride, w // stride, generator=g).to(cuda) if "r" in extras else None
    want = F.conv2d(x.double(), conv.weight.double(), conv.bias.double(), stride=stride, padding=k // 2)
    if cb is not None:
        want = want + cb.double()[:, :, None, None]
    if res is not None:
        want = want + res.double()
    got = K.conv2d(K.image_split(x), conv, chan_bias=cb, residual=res)
    assert got.shape == want.shape
    err = relerr(got, want)
    assert err < (1e-4 if (operand_mode == 1 and cin % 64 == 0) else 1e-5), f"conv rel err {err:.3e}"


@pytest.mark.parametrize("rows,c", [(1, 320), (1000, 640), (4097, 1280), (77, 64), (5, 2048)])
def test_layernorm_split(cuda, operand_mode, rows, c):
    from vidseg_diffusion_b200 import kernels as K
    g = torch.Generator(device="cpu").manual_seed(rows + c)
    ln = nn.LayerNorm(c).to(cuda)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.1 * torch.randn(c, generator=g))
        ln.bias.copy_(0.1 * torch.randn(c, generator=g))
    x = (torch.randn(rows, c, generator=g) * 3 + 1).to(cuda)
    want = F.layer_norm(x.double(), (c,), ln.weight.double(), ln.bias.double(), ln.eps)
    assert relerr(K.layer_norm_split(x, ln).float(), want) < (3e-6 if operand_mode == 0 else 6e-5)


def test_geglu_split(cuda, operand_mode):
    from vidseg_diffusion_b200 import kernels as K
    x = torch.randn(300, 2 * 1280, generator=torch.Generator().manual_seed(3)).to(cuda) * 2
    val, gate = x.double().chunk(2, dim=-1)
    assert relerr(K.geglu_split(x).float(), val * F.gelu(gate)) < (3e-6 if operand_mode == 0 else 6e-5)


@pytest.mark.parametrize("b,h,w,c1,c2,silu,eps", [(2, 16, 16, 320, 0, True, 1e-5), (3, 8, 8, 1280, 640, True, 1e-5),
                                                    (1, 64, 64, 320, 320, True, 1e-5), (2, 4, 4, 1280, 1280, True, 1e-5),
                                                    (2, 32, 32, 640, 0, False, 1e-6), (2, 5, 7, 64, 0, True, 1e-5)])
def test_groupnorm_split_with_concat(cuda, operand_mode, b, h, w, c1, c2, silu, eps):
    from vidseg_diffusion_b200 import kernels as K
    g = torch.Generator(device="cpu").manual_seed(b + h + c1 + c2)
    c = c1 + c2
    gn = nn.GroupNorm(32, c, eps=eps).to(cuda)
    with torch.no_grad():
        gn.weight.copy_(1 + 0.1 * torch.randn(c, generator=g))
        gn.bias.copy_(0.1 * torch.randn(c, generator=g))
    xa = (torch.randn(b, c1, h, w, generator=g) * 2 + 0.5).to(cuda).contiguous(memory_format=torch.channels_last)
    xb = torch.randn(b, c2, h, w, generator=g).to(cuda) if c2 else None
    src = K.concat_channels(xa, xb) if c2 else xa
    full = torch.cat([xa, xb], 1) if c2 else xa
    want = F.group_norm(full.double(), 32, gn.weight.double(), gn.bias.double(), eps)
    if silu:
        want = F.silu(want)
    out, raw, first = K.group_norm_split(src, gn, silu=silu, want_raw=True)
    loose = operand_mode == 1
    assert relerr(out.float().permute(0, 3, 1, 2), want) < (6e-5 if loose else 4e-6)
    assert relerr(raw.float().permute(0, 3, 1, 2), full.double()) < (6e-5 if loose else 2e-6)
    assert torch.equal(first.permute(0, 3, 1, 2), xa)
    again, _, _ = K.group_norm_split(src, gn, silu=silu)
    assert torch.equal(again.hi, out.hi) and torch.equal(again.lo, out.lo)  # deterministic


def test_upsample2x_split(cuda, operand_mode):
    from vidseg_diffusion_b200 import kernels as K
    x = torch.randn(2, 64, 5, 7, generator=torch.Generator().manual_seed(1)).to(cuda)
    want = F.interpolate(x.double(), scale_factor=2, mode="nearest")
    got = K.upsample_nearest2x_split(x).float().permute(0, 3, 1, 2)
    assert got.shape == want.shape and relerr(got, want) < (2e-6 if operand_mode == 0 else 6e-5)
