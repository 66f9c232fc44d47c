"""One K-means fit at config-2 size (N = 14336, D = 640, K = 20, n_init = 10) on synthetic clip features: target of an
ncu launch list (per-kernel durations of the Lloyd loop) and a CUDA-event timing of the whole fit."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synthetic_clip_features
from vidseg_diffusion_b200.features import aggregate_normalize
from vidseg_diffusion_b200.kmeans import KMeans
kind = sys.argv[1] if len(sys.argv) > 1 else "objects"
n_init = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
blocks, _ = synthetic_clip_features(1, 14, 32, 32, 640, 20, kind=kind)
X = aggregate_normalize([torch.from_numpy(b).to(dev) for b in blocks], 14)
for rep in range(3):
    np.random.seed(1)
    km = KMeans(n_clusters=20, n_init=n_init)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    km.fit_predict(X)
    torch.cuda.synchronize()
    print(f"fit {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms, iterations {km.info_['max_iter_run']}, launches {km.info_['launches']}")
