"""K-means only, config-2 size, for ncu launch lists: python tools/run_kmeans_once.py [reps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synthetic_clip_features
from vidseg_diffusion_b200.features import aggregate_normalize
from vidseg_diffusion_b200.kmeans import KMeans
dev = torch.device("cuda", 0)
F, h, w, C, K = 14, 32, 32, 640, 20
kind = os.environ.get("KIND", "objects")
blocks, _ = synthetic_clip_features(1, F, h, w, C, K, kind=kind)
x = aggregate_normalize([torch.from_numpy(b).to(dev) for b in blocks], F)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    np.random.seed(1)
    km = KMeans(n_clusters=K, n_init=10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); km.fit_predict(x); e1.record(); torch.cuda.synchronize()
    print("fit_predict ms", e0.elapsed_time(e1), km.info_, flush=True)
