"""Per-shape time of every Linear GEMM of one UNet step (CUDA events around each call, eager mode)."""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from vidseg_diffusion_b200 import configs, kernels as K, linear
wlname = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[wlname]
cfg = {"sd21": configs.SD21_UNET, "svd": configs.SVD_UNET}[wl["cfg"]]
dev = torch.device("cuda", 0)
sd = bench.make_state_dict(cfg)
with torch.device("meta"):
    model = bench.model_class(cfg)(**cfg)
model = model.to_empty(device=dev); model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True); model.eval()
clip = [a.to(dev) for a in bench.make_clip(wl, cfg, 1)]
kw = dict(num_video_frames=wl["frames"], y=clip[3]) if bench.is_video(cfg) else {}
run = lambda: model(clip[0], timesteps=clip[1], context=clip[2], **kw)
run(); run()
rec = collections.defaultdict(list)
orig = linear.gemm_split
def wrapped(a, w, bias=None, residual=None, want_f32=True, want_split=False, **kw2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(a, w, bias, residual, want_f32=want_f32, want_split=want_split, **kw2); e1.record()
    k = a.hi.shape[-1]; m = a.hi.numel() // k; n = w.hi.shape[0]
    rec[(m, n, k, "f" if want_f32 else "", "s" if want_split else "", "r" if residual is not None else "")].append((e0, e1))
    return out
K.gemm_split = wrapped
run()
torch.cuda.synchronize()
rows = []
for key, evs in rec.items():
    ms = sum(a.elapsed_time(b) for a, b in evs)
    m, n, k = key[:3]
    rows.append((ms, key, len(evs), 2.0 * m * n * k * len(evs) / ms / 1e9))
tot = sum(r[0] for r in rows)
print(f"{wlname}: {sum(r[2] for r in rows)} GEMM calls, {tot:.2f} ms")
for ms, key, cnt, tf in sorted(rows, reverse=True)[:30]:
    print(f"  M={key[0]:7d} N={key[1]:5d} K={key[2]:5d} {''.join(key[3:]):4s} n={cnt:3d} total={ms:7.3f} ms  avg={1e3*ms/cnt:7.1f} us  {tf:6.1f} alg TFLOP/s")
