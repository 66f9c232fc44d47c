"""Launch the hot tensor-core kernels at their config-2 shapes a few times each (target of `ncu --set full`).
Also prints CUDA-event timings when run without a profiler."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vidseg_diffusion_b200.linear import gemm_split, split, attention_split, Split
from vidseg_diffusion_b200 import kernels as K

from vidseg_diffusion_b200 import _lib
_lib.load().vidseg_set_operand_mode(int(os.environ.get("MODE", "1")))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
res = {}

def ev(fn, it=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

which = sys.argv[1:] or ["attn", "gemm", "conv"]
if "attn" in which:
    for (B, H, N, Nk) in [(28, 5, 4096, 4096), (28, 10, 1024, 1024), (28, 5, 4096, 77)]:
        q = split(torch.randn(B, N, H * 64, device=dev), pair16=True); k = split(torch.randn(B, Nk, H * 64, device=dev), pair16=True); v = split(torch.randn(B, Nk, H * 64, device=dev), pair16=True)
        ms = ev(lambda: attention_split(q, k, v, H, 0.125))
        res[f"attn_B{B}_H{H}_N{N}_Nk{Nk}"] = {"ms": ms, "alg_TFLOPs": 4.0 * B * H * N * Nk * 64 / ms / 1e9}
if "gemm" in which:
    for (m, n, kk) in [(28 * 4096, 2560, 320), (28 * 4096, 320, 1280), (28 * 1024, 5120, 640), (28 * 4096, 320, 320)]:
        a = split(torch.randn(m, kk, device=dev)); w = split(torch.randn(n, kk, device=dev) / kk ** 0.5, 256.0, is_weight=True)
        ms = ev(lambda: gemm_split(a, w))
        res[f"gemm_{m}x{n}x{kk}"] = {"ms": ms, "alg_TFLOPs": 2.0 * m * n * kk / ms / 1e9}
if "gemmres" in which:   # the memory-bound projection shape of the first UNet level, with residual (to_out / proj_out)
    m, n, kk = 28 * 4096, 320, 320
    a = split(torch.randn(m, kk, device=dev)); w = split(torch.randn(n, kk, device=dev) / kk ** 0.5, 256.0, is_weight=True)
    res_t = torch.randn(m, n, device=dev); bias = torch.randn(n, device=dev)
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def run():
        big.zero_()   # flush L2
        return gemm_split(a, w, bias, res_t)
    ms = ev(run, it=3) - ev(lambda: big.zero_(), it=3)
    res["gemmres_114688x320x320_fr"] = {"ms": ms, "GBs": (m * kk * 4 + 2 * m * n * 4) / ms / 1e6}
if "proj" in which:   # K = 320 projections: plain fp32 output, then the q / k form (fp32 stash + fp16-pair operand)
    m, n, kk = 28 * 4096, 320, 320
    a = split(torch.randn(m, kk, device=dev)); w = split(torch.randn(n, kk, device=dev) / kk ** 0.5, 256.0, is_weight=True)
    ms = ev(lambda: gemm_split(a, w, want_f32=True))
    res["proj_f32"] = {"ms": ms, "GBs": (m * kk * 4 + m * n * 4) / ms / 1e6}
    ms = ev(lambda: gemm_split(a, w, want_f32=True, want_split=True, split_pair16=True))
    res["proj_f32_split"] = {"ms": ms, "GBs": (m * kk * 4 + 2 * m * n * 4) / ms / 1e6}
if "conv" in which:
    for (B, H, C, Co) in [(28, 64, 320, 320), (28, 32, 640, 640), (28, 64, 640, 320), (28, 8, 1280, 1280), (28, 8, 2560, 1280), (28, 16, 1280, 1280)]:
        conv = torch.nn.Conv2d(C, Co, 3, padding=1).to(dev)
        x = split(torch.randn(B, H, H, C, device=dev))
        ms = ev(lambda: K.conv2d(x, conv))
        res[f"conv3x3_B{B}_{H}x{H}_{C}to{Co}"] = {"ms": ms, "alg_TFLOPs": 2.0 * B * H * H * Co * 9 * C / ms / 1e9}
print(json.dumps(res, indent=1))
