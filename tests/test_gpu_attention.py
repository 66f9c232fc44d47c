"""GPU: tcgen05 split-fp16 attention against a float64 softmax reference.  The logits are 22-bit products (errors there
are exponentiated); the probabilities enter the P V product as single fp16 numbers with the row sum taken over the same
rounded values -- a re-weighting of the keys by 1 + eps, |eps| <= 2^-11 -- and V as an fp16 pair.  On these iid inputs
(the worst case: O is a small average of large, uncorrelated v) that is 1-3e-4 of the output scale; tolerance 5e-4, the
parity bar of the path is 1e-3 and is checked end to end in test_gpu_unet.py / test_gpu_fullsize.py."""
TOL = 5e-4
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,h,nq,nk,sharp", [
    (2, 5, 512, 512, 1.0), (1, 2, 256, 77, 1.0), (3, 1, 64, 64, 1.0), (2, 3, 300, 300, 1.0), (1, 20, 64, 1, 1.0),
    (2, 10, 1024, 1024, 4.0), (1, 5, 4096, 4096, 1.0), (2, 2, 130, 65, 8.0), (1, 1, 1, 1, 1.0),
])
def test_attention_matches_fp64(cuda, b, h, nq, nk, sharp):
    from vidseg_diffusion_b200.linear import attention_split, split
    g = torch.Generator(device="cpu").manual_seed(b * 1000 + h * 100 + nq + nk)
    c = h * 64
    q = (torch.randn(b, nq, c, generator=g) * sharp).to(cuda)
    k = torch.randn(b, nk, c, generator=g).to(cuda)
    v = torch.randn(b, nk, c, generator=g).to(cuda)
    qd, kd, vd = [t.double().reshape(b, -1, h, 64).permute(0, 2, 1, 3) for t in (q, k, v)]
    want = torch.softmax(qd @ kd.transpose(-1, -2) / 8.0, dim=-1) @ vd
    want = want.permute(0, 2, 1, 3).reshape(b, nq, c)
    out, sp = attention_split(split(q, pair16=True), split(k, pair16=True), split(v, pair16=True), h, want_f32=True, want_split=True)
    scale = want.abs().max().item()
    err = (out.double() - want).abs().max().item() / scale
    assert err < TOL, f"fp32 output rel err {err:.3e}"
    err = (sp.float().double() - want).abs().max().item() / scale
    assert err < TOL, f"split output ({sp.fmt}) rel err {err:.3e}"


@pytest.mark.parametrize("step", [0.01, 0.2, 3.0])
def test_attention_running_maximum_ramps_up(cuda, step):
    """Scores that keep growing along the key axis: exercises both sides of the lazy rescale (blocks that raise the
    reference maximum and blocks that stay within its 2^3 headroom)."""
    from vidseg_diffusion_b200.linear import attention_split, split
    g = torch.Generator(device="cpu").manual_seed(17)
    b, h, nq, nk = 1, 2, 256, 640
    c = h * 64
    q = torch.randn(b, nq, c, generator=g).abs()
    k = torch.randn(b, nk, c, generator=g) * 0.1 + (torch.arange(nk).float() * step / 64.0)[None, :, None]
    v = torch.randn(b, nk, c, generator=g)
    q, k, v = q.to(cuda), k.to(cuda), v.to(cuda)
    qd, kd, vd = [t.double().reshape(b, -1, h, 64).permute(0, 2, 1, 3) for t in (q, k, v)]
    want = (torch.softmax(qd @ kd.transpose(-1, -2) / 8.0, dim=-1) @ vd).permute(0, 2, 1, 3).reshape(b, nq, c)
    out, _ = attention_split(split(q, pair16=True), split(k, pair16=True), split(v, pair16=True), h, want_f32=True, want_split=False)
    err = (out.double() - want).abs().max().item() / want.abs().max().item()
    assert err < TOL, f"rel err {err:.3e}"
