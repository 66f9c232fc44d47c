"""Times the Lloyd E-step launch alone (CUDA events over 50 back-to-back assign calls on a prepared workspace)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synthetic_clip_features
from vidseg_diffusion_b200.features import aggregate_normalize
from vidseg_diffusion_b200.kmeans import draw_kmeanspp_randoms
from vidseg_diffusion_b200 import distributed as D
dev = torch.device("cuda", 0)
blocks, _ = synthetic_clip_features(1, 14, 32, 32, 640, 20, kind="iid")
X = aggregate_normalize([torch.from_numpy(b).to(dev) for b in blocks], 14)
n = X.shape[0]
np.random.seed(1)
first, rand = draw_kmeanspp_randoms(n, 20, 10)
be = D.CudaLloydBackend(20, 10, 300, 1e-4)
be.prepare(X); be.seed(first, rand)
for _ in range(3):
    be.assign(0, n); w = be.partial_words(0, n, "f64"); be.update_words(w, "f64", False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    be.assign(0, n)
b.record(); torch.cuda.synchronize()
print(f"E-step (score GEMM + epilogue + resolver): {a.elapsed_time(b) / 50 * 1e3:.1f} us per call, debug={os.environ.get('VIDSEG_KM_DEBUG', '0')} fused={os.environ.get('VIDSEG_KMEANS_FUSED_E', '1')}")
