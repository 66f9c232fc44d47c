"""torchrun target: ONE clip with its frames sharded over the ranks (NCCL), checked against the unsharded result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/run_sharded.py [--workload c2] [--steps 3]
Prints one JSON line from rank 0: parity with the single-GPU path (label maps identical), per-clip latency (CUDA events,
max over ranks), Lloyd iterations and collectives per clip."""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from vidseg_diffusion_b200 import configs
from vidseg_diffusion_b200.distributed import ShardedClipSegmenter
from vidseg_diffusion_b200.pipeline import ClipSegmenter

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
wl = bench.WORKLOADS[args.workload]
cfg = {"sd21": configs.SD21_UNET, "tiny": configs.TINY_UNET}[wl["cfg"]]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sd = bench.make_state_dict(cfg)
with torch.device("meta"):
    model = bench.model_class(cfg)(**cfg)
model = model.to_empty(device=dev)
model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
model.eval()
F = wl["frames"]
x, t, ctx = (a.to(dev) for a in bench.make_clip(wl, cfg, 1))
sh = ShardedClipSegmenter(model, num_masks=wl["num_masks"], is_aggre_attn=wl["aggre"], is_refine_mask=wl["refine"])
one = ClipSegmenter(model, num_masks=wl["num_masks"], is_aggre_attn=wl["aggre"], is_refine_mask=wl["refine"])

def timed(fn, steps):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out

for _ in range(2):
    sh.segment(x, t, ctx, F, seed=1)
ms_sh, lab_sh = timed(lambda: sh.segment(x, t, ctx, F, seed=1), args.steps)
info = sh.last["kmeans_info"]
one.segment(x, t, ctx, F, seed=1)
ms_one, (lab_one, _) = timed(lambda: one.segment(x, t, ctx, F, seed=1), args.steps)
same = bool(torch.equal(lab_sh, lab_one))
agree = torch.tensor([int(same)], device=dev)
dist.all_reduce(agree, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"workload": args.workload, "world": world, "frames": F,
                      "sharded_ms_per_clip": ms_sh, "sharded_frames_per_s": F / ms_sh * 1e3,
                      "single_gpu_ms_per_clip": ms_one, "single_gpu_frames_per_s": F / ms_one * 1e3,
                      "labels_identical_to_single_gpu_on_all_ranks": bool(agree.item()),
                      "lloyd_iterations": info["iterations"], "allreduces_per_clip": info["allreduces"],
                      "unsharded_fallback": info["unsharded_fallback"]}), flush=True)
dist.destroy_process_group()
