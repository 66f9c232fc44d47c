"""Scratch timing of the individual stages on one GPU (CUDA events). Not the contract bench."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synthetic_clip_features
from vidseg_diffusion_b200.features import aggregate_normalize
from vidseg_diffusion_b200.kmeans import KMeans
from vidseg_diffusion_b200.refine import refine_masks

def ev_time(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

dev = torch.device("cuda", 0)
res = {}
F, h, w, C, K = 14, 32, 32, 640, 20
blocks, _ = synthetic_clip_features(1, F, h, w, C, K)
db = [torch.from_numpy(b).to(dev) for b in blocks]
res["aggregate_ms"] = ev_time(lambda: aggregate_normalize(db, F), it=20)
res["aggregate_GBs"] = (4 * F * h * w * C * 4) / res["aggregate_ms"] / 1e6
x = aggregate_normalize(db, F)
def km():
    np.random.seed(1)
    k = KMeans(n_clusters=K, n_init=10); l = k.fit_predict(x); return k, l
res["kmeans_ms"] = ev_time(lambda: km(), warm=1, it=3)
k, labels = km(); res["kmeans_info"] = {kk: (vv.tolist() if hasattr(vv, "tolist") else vv) for kk, vv in k.info_.items()}
res["kmeans_n_iter"] = k.n_iter_
t0 = time.perf_counter(); km(); torch.cuda.synchronize(); res["kmeans_wall_ms"] = (time.perf_counter() - t0) * 1e3
lab = labels.reshape(F, h, w)
res["refine_ms"] = ev_time(lambda: refine_masks(db[1], lab, F, h, w), warm=1, it=5)
try:
    from vidseg_diffusion_b200.linear import gemm_split, split
    for (m, n, kk) in [(28 * 4096, 320, 320), (28 * 4096, 2560, 320), (28 * 4096, 320, 1280), (28 * 1024, 640, 640), (28 * 1024, 5120, 640), (28 * 256, 1280, 1280), (28 * 256, 10240, 1280)]:
        a = split(torch.randn(m, kk, device=dev)); wt = split(torch.randn(n, kk, device=dev) / kk ** 0.5, 256.0, is_weight=True)
        ms = ev_time(lambda: gemm_split(a, wt), it=10)
        res[f"gemm_{m}x{n}x{kk}_ms"] = ms
        res[f"gemm_{m}x{n}x{kk}_useful_TFLOPs"] = 2.0 * m * n * kk / ms / 1e9
except Exception as e:
    res["gemm_error"] = repr(e)
print(json.dumps(res, indent=1))
