"""GPU parity AT THE BENCHMARKED SIZES (BASELINE.json configs[1] and configs[2]): 14 frames, 512x512, batch 28.

The small-size tests (test_gpu_unet.py / test_gpu_video_unet.py) stop at configs[0] (4 frames, 256x256); the headline
numbers come from batch 28 at 64x64 latents, N = 4096 self-attention and the 160- / 256-wide GEMM tiles, so this file
runs the same checks there:

  * float stage: stashed ``attn1.q`` of output blocks 6 / 7 / 8 and the UNet output against the UNMODIFIED reference
    module (oracle/reference_path.py -> oracle/_ref or /root/reference) run in fp32 on this box's CPU, with the bench's
    own weights and clip; bar 1e-3 (max|delta| / max|ref| per tensor);
  * integer stage: the label maps of ``ClipSegmenter.segment`` (and of the CUDA-graph + two-stream ``segment_many`` the
    bench times) against the reference's clustering arithmetic -- scikit-learn KMeans, oracle/refine.py -- run on the
    CPU on the very feature tensors the GPU clustered; bit-exact.

The CPU side takes about a minute per workload on the GPU box's host cores (slow marker, still part of ``-m gpu``).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
TOL = 1e-3


def relerr(got, want):
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    return float((got - want).abs().max() / want.abs().max())


@pytest.mark.parametrize("workload", ["c2", "c3"])
def test_full_size_features_and_labels(cuda, workload):
    import bench
    from oracle import reference_path as rp
    from vidseg_diffusion_b200 import configs
    from vidseg_diffusion_b200.pipeline import ClipSegmenter, harvest_self_attn_q
    if not rp.available():
        pytest.skip("oracle/_ref missing: run `python oracle/make_ref.py` where /root/reference exists")
    wl = bench.WORKLOADS[workload]
    cfg = {"sd21": configs.SD21_UNET, "svd": configs.SVD_UNET}[wl["cfg"]]
    F = wl["frames"]
    sd = bench.make_state_dict(cfg)
    clip = bench.make_clip(wl, cfg, 1)
    # ---- GPU path
    with torch.device("meta"):
        model = bench.model_class(cfg)(**cfg)
    model = model.to_empty(device=cuda)
    model.load_state_dict({k: v.to(cuda) for k, v in sd.items()}, strict=True)
    model.eval()
    dev = [t.to(cuda) for t in clip]
    kw = dict(num_video_frames=F, y=dev[3]) if bench.is_video(cfg) else {}
    seg_kw = dict(num_masks=wl["num_masks"], is_aggre_attn=wl["aggre"], is_refine_mask=wl["refine"])
    seg = ClipSegmenter(model, **seg_kw)
    labels, out = seg.segment(dev[0], dev[1], dev[2], F, seed=1, **kw)
    labels = labels.cpu().numpy().reshape(-1)
    q_gpu = {i: q.cpu() for i, q in zip((8, 7, 6), harvest_self_attn_q(model, (8, 7, 6)))}
    out_gpu = out.cpu()
    X = seg.last["features"].cpu().numpy()
    # the schedule the bench times: UNet stage as one CUDA graph, clips pipelined over two streams
    seg_g = ClipSegmenter(model, use_cuda_graph=True, **seg_kw)
    many = list(seg_g.segment_many([(dev[0], dev[1], dev[2], kw)] * 2, F, seed=1, to_host=True))
    assert all(np.array_equal(m.numpy().reshape(-1), labels) for m in many), "segment_many differs from segment"
    del model, seg, seg_g, dev
    torch.cuda.empty_cache()
    # ---- integer stage on the GPU's own features
    want, how = bench.label_oracle(X, q_gpu[7].numpy() if wl["refine"] else None, wl, 1)
    mism = int((labels != want).sum())
    print(f"{workload}: label maps vs {how}: {mism} of {want.size} cells differ")
    assert mism == 0
    # ---- float stage against the unmodified reference module on the CPU
    torch.set_num_threads(os.cpu_count() or 1)
    ref_model = rp.build_model(cfg)
    ref_model.load_state_dict(sd, strict=True)
    out_ref = rp.unet_forward(ref_model, clip, F)
    errs = {f"q{i}": relerr(q_gpu[i], rp.stashed_q(ref_model, i)) for i in (6, 7, 8)}
    errs["out"] = relerr(out_gpu, out_ref)
    print(f"{workload}: rel err vs the reference module: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert max(errs.values()) <= TOL, errs
