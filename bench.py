#!/usr/bin/env python
"""Benchmark of the VidSeg per-clip hot path on B200 (contract: see DESIGN.md section "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference on the host cores (oracle/_ref)

One "step" = one pass of the hot path over one synthetic clip per GPU: SD-2.1 UNet forward at batch 2F
with the Q/K stash -> aggregate(out 8,7,6)/normalise -> K-means(K, n_init=10) fit+predict
(BASELINE.json configs[1]: 14 frames, 512x512, num_masks=20, --is_aggre_attn).  Metric: frames/sec.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "frames/sec per clip (14x512x512)"
UNIT = "frames/s"

WORKLOADS = {
    # name: (config key, frames, latent side, context len, num_masks, aggregate, refine)
    "c2": dict(cfg="sd21", frames=14, latent=64, ctx_len=77, num_masks=20, aggre=True, refine=False,
               desc="SD-2.1 UNet 1 step + aggregate(out 8,7,6) + KMeans(20, n_init=10), 14 frames 512x512 (BASELINE configs[1])"),
    "c2r": dict(cfg="sd21", frames=14, latent=64, ctx_len=77, num_masks=20, aggre=True, refine=True,
                desc="configs[1] plus --is_refine_mask"),
    "c3": dict(cfg="svd", frames=14, latent=64, ctx_len=1, num_masks=20, aggre=True, refine=True,
               desc="SVD VideoUNet 1 step + aggregate(out 8,7,6) + KMeans(20, n_init=10) + mask refinement, 14 frames 512x512 (BASELINE configs[2])"),
    "c5": dict(cfg="svd", frames=28, latent=96, ctx_len=1, num_masks=50, aggre=True, refine=True,
               desc="SVD VideoUNet, 28 frames 768x768, num_masks=50, refinement: ONE clip of BASELINE configs[4] on one GPU (size check)"),
    "c1": dict(cfg="sd21", frames=4, latent=32, ctx_len=77, num_masks=5, aggre=False, refine=False,
               desc="SD-2.1, 4 frames 256x256, num_masks=5 (BASELINE configs[0])"),
    "c2mod": dict(cfg="sd21", frames=14, latent=64, ctx_len=77, num_masks=20, aggre=True, refine=True, modulated=True,
                  desc="configs[1] continued to the final maps: source run of sampler steps 17..24 with q / k / x_t kept in HBM, "
                       "K-means + refinement, then the 2 x 20 modulated sampler runs (8 UNet steps each) + first-stage decode + "
                       "seg-map post-process (svd_single_video_inference.py:404-508); opt-in, one step is ~330 UNet evaluations"),
    "tinymod": dict(cfg="tiny", frames=2, latent=16, ctx_len=7, num_masks=3, aggre=True, refine=True, modulated=True,
                    desc="toy-width plumbing check of the modulated-run workload"),
    "tiny": dict(cfg="tiny", frames=2, latent=16, ctx_len=7, num_masks=3, aggre=True, refine=True,
                 desc="toy-width UNet, plumbing check only"),
    "tinyv": dict(cfg="tinyv", frames=3, latent=16, ctx_len=1, num_masks=3, aggre=True, refine=True,
                  desc="toy-width VideoUNet, plumbing check only"),
}
# algorithmic FLOPs of one UNet step (FlopCounterMode over the reference modules)
UNET_TFLOP = {"c2": 22.519, "c2r": 22.519, "c1": 1.449, "c3": 35.542, "c5": 179.076}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return {"hbm_gbs": float(d["hbm_gbs"]), "tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"}


def is_video(cfg):
    return "video_kernel_size" in cfg


def model_class(cfg):
    if is_video(cfg):
        from vidseg_diffusion_b200.sgm.modules.diffusionmodules.video_model import VideoUNet
        return VideoUNet
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
    return UNetModel


def param_shapes(cfg):
    """{state-dict key: shape} of the architecture, from the module tree on the meta device (no allocation)."""
    import torch
    with torch.device("meta"):
        model = model_class(cfg)(**cfg)
    return {k: tuple(v.shape) for k, v in model.state_dict().items()}


def make_state_dict(cfg, seed=0):
    """Random-init weights of the reference architecture: N(0, 1/fan_in) matrices, unit norms, small biases."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in param_shapes(cfg).items():
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[key] = torch.randn(shape, generator=g) * fan_in ** -0.5
        elif key.endswith(".weight"):
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[key] = 0.05 * torch.randn(shape, generator=g)
    return sd


def make_clip(wl, cfg, seed):
    """Host tensors of one step: x [2F,C,h,w] (uncond rows first), timesteps [2F], context [2F,L,D] (+ y [2F,adm] for
    the SVD VideoUNet: noisy latent | conditioning-frame latent, one CLIP image token, vector conditioning)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    F, hw = wl["frames"], wl["latent"]
    if is_video(cfg):
        half = cfg["in_channels"] // 2
        lat = torch.randn(F, half, hw, hw, generator=g)
        cond = torch.randn(1, half, hw, hw, generator=g).expand(F, -1, -1, -1)
        ctx = torch.randn(1, 1, cfg["context_dim"], generator=g).expand(F, -1, -1)
        yv = torch.randn(1, cfg["adm_in_channels"], generator=g).expand(F, -1)
        x = torch.cat([torch.cat([lat, torch.zeros_like(cond)], 1), torch.cat([lat, cond], 1)], 0).contiguous()
        context = torch.cat([torch.zeros_like(ctx), ctx], 0).contiguous()
        t = torch.full((2 * F,), 0.8)   # c_noise = 0.25 ln(sigma) (denoiser_scaling.py:58)
        return x, t, context, torch.cat([yv, yv], 0).contiguous()
    lat = torch.randn(F, cfg["in_channels"], hw, hw, generator=g)
    ctx = torch.randn(F, wl["ctx_len"], cfg["context_dim"], generator=g)
    x = torch.cat([lat, lat], 0)
    context = torch.cat([torch.zeros_like(ctx), ctx], 0)
    t = torch.full((2 * F,), 21.0)  # DiscreteDenoiser index of the last sampler step (feature_timestep 24 of 25)
    return x, t, context


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bench_config(args, wl, world, multi_gpu):
    """The ``config`` object of the JSON line: what is measured, identical for this repo's arm and the reference arm."""
    F = wl["frames"]
    if world == 1:
        mg, clips = "single GPU", 1
    elif multi_gpu == "sharded":
        mg = ("ONE clip per step, frames sharded over the ranks (BASELINE configs[3]): NCCL all-gather of the feature rows, then "
              + ("the n_init K-means initialisations spread over the ranks (no collective inside the Lloyd loop) and one "
                 "all-reduce of the run records" if args.kmeans_split == "runs" else
                 "one all-reduce of the exchange words per Lloyd iteration"))
        clips = 1
    elif multi_gpu == "cfg-split":
        mg = "ONE clip per step, the two guidance halves of the SVD batch on two ranks, feature broadcast + distributed K-means"
        clips = 1
    else:
        mg, clips = "one clip per GPU per step (replicas), no data-path collective", world
    return {"workload": args.workload, "desc": wl["desc"], "frames": F, "clips_per_step": clips, "multi_gpu": mg,
            "l2": "working set (3.5 GB split weights + >4 GB activations per step) exceeds the 126 MB L2; no flush needed",
            "unet_tflop_per_step": UNET_TFLOP.get(args.workload)}


def multi_gpu_mode(args, cfg, world):
    if world == 1:
        return "single"
    if args.multi_gpu != "auto":
        return args.multi_gpu
    if is_video(cfg):
        return "cfg-split" if world == 2 else "replicas"
    return "sharded"


# --------------------------------------------------------------------------------------------------
# the reference's own CPU path (unmodified reference modules, oracle/reference_path.py)
# --------------------------------------------------------------------------------------------------
_REF_MODEL = {}


def reference_model(cfg, sd):
    """The reference's nn.Module for ``cfg`` with the bench's weights (built once per process)."""
    from oracle import reference_path as rp
    key = id(cfg)
    if key not in _REF_MODEL:
        m = rp.build_model(cfg)
        m.load_state_dict(sd, strict=True)
        _REF_MODEL[key] = m
    return _REF_MODEL[key]


def cpu_reference_step(wl, cfg, sd, clip, seed, keep=False):
    """One pass of the UNMODIFIED reference over the whole clip on the host cores: ``UNetModel`` / ``VideoUNet`` forward
    at batch 2F in fp32, ``torch.save`` of the stashed q tensors, ``feature_extraction_main("kmeans_masks")`` (scikit-learn
    KMeans + PNG tree) and, with refinement, ``feature_extraction_main("correct_low_res_mask")``."""
    from oracle import reference_path as rp
    return rp.clip_step(reference_model(cfg, sd), clip, wl["frames"], wl["num_masks"], aggre=wl["aggre"], refine=wl["refine"],
                        seed=seed, keep=keep)


def reference_sample_text(wl, secs=None):
    import torch
    txt = (f"{wl['frames']} of {wl['frames']} frames per step: the unmodified reference (oracle/_ref mirror of sgm + scripts) on the "
           f"CPU, UNet batch {2 * wl['frames']} at {wl['latent'] * 8}x{wl['latent'] * 8} fp32, torch.save of the stashed q, "
           f"feature_extraction_main kmeans_masks{' + correct_low_res_mask' if wl['refine'] else ''} (scikit-learn KMeans, PNG tree), "
           f"torch {torch.__version__}")
    if secs:
        txt += "; stages " + ", ".join(f"{k} {v:.1f} s" for k, v in secs.items())
    return txt


def run_reference(args, wl, cfg):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import reference_path as rp
    if not rp.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists"})
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(cfg)
    clip = make_clip(wl, cfg, 1)
    t_start = time.perf_counter()
    # every step is the whole 14-frame clip; if W + K such steps do not fit the wall budget the step COUNT is cut (and
    # reported), never the workload
    done_w, per_step = 0, None
    for _ in range(args.warmup):
        r = cpu_reference_step(wl, cfg, sd, clip, 1)
        done_w += 1
        per_step = r["seconds"]["total"]
        left = args.ref_budget_s - (time.perf_counter() - t_start)
        if per_step * (args.warmup - done_w + args.steps) > left:
            break
    steps = args.steps
    if per_step is not None:
        left = args.ref_budget_s - (time.perf_counter() - t_start)
        steps = max(1, min(args.steps, int(left / per_step)))
    res = [cpu_reference_step(wl, cfg, sd, clip, 1) for _ in range(steps)]
    total = sum(r["seconds"]["total"] for r in res)
    stage = {k: sum(r["seconds"][k] for r in res) / steps for k in res[0]["seconds"]}
    fps = wl["frames"] * steps / total
    sample = reference_sample_text(wl, stage)
    mode = multi_gpu_mode(args, cfg, world)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": done_w, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "strong" if mode in ("sharded", "cfg-split") else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, wl, world, mode),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "threads": {"os_cpu_count": cores, "torch": torch.get_num_threads()},
    }
    emit(line)


def relerr(got, want):
    import torch
    got, want = torch.as_tensor(got).double().cpu(), torch.as_tensor(want).double().cpu()
    return float((got - want).abs().max() / want.abs().max())


def partition_agreement(a, b, k):
    """Fraction of cells on which two label maps agree after the best one-to-one matching of their label ids."""
    import numpy as np
    conf = np.zeros((k, k), dtype=np.int64)
    np.add.at(conf, (a.astype(np.int64), b.astype(np.int64)), 1)
    try:
        from scipy.optimize import linear_sum_assignment
        r, c = linear_sum_assignment(-conf)
        return float(conf[r, c].sum() / max(a.size, 1))
    except ImportError:
        return float(conf.max(axis=1).sum() / max(a.size, 1))


def label_oracle(X, q7, wl, seed):
    """Label maps the reference's clustering arithmetic gives on the GPU's OWN features: scikit-learn KMeans (the reference's
    third-party call, feature_extraction.py:52-55; the oracle's restatement when sklearn is missing) and the oracle's
    correct_low_res_mask."""
    import numpy as np
    from oracle import kmeans as okm, refine as oref
    F, fh = wl["frames"], wl["latent"] // 2
    np.random.seed(seed)
    try:
        import warnings
        from sklearn.cluster import KMeans
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            km = KMeans(n_clusters=wl["num_masks"], n_init=10).fit(X)
            labels = km.predict(X)
        how = "scikit-learn KMeans"
    except ImportError:
        labels, _ = okm.kmeans_fit_predict(X, wl["num_masks"])
        how = "oracle/kmeans.py"
    if wl["refine"]:
        labels, _, _, _ = oref.correct_low_res_mask(q7, labels.reshape(F, fh, fh), fh, fh, F)
        how += " + oracle/refine.py"
    return np.asarray(labels).reshape(-1), how


# --------------------------------------------------------------------------------------------------
# this repo's CUDA path
# --------------------------------------------------------------------------------------------------
def run_b200(args, wl, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.pipeline import ClipSegmenter, harvest_self_attn_q

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout when NCCL_DEBUG is VERSION / WARN: keep stdout for the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    mode = multi_gpu_mode(args, cfg, world)
    one_clip = mode in ("sharded", "cfg-split")

    sd = make_state_dict(cfg)
    with torch.device("meta"):
        model = model_class(cfg)(**cfg)
    model = model.to_empty(device=dev)
    model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    model.eval()
    seg_kw = dict(num_masks=wl["num_masks"], is_aggre_attn=wl["aggre"], is_refine_mask=wl["refine"])
    seg = ClipSegmenter(model, use_cuda_graph=not args.no_graph, **seg_kw)
    seg_eager = ClipSegmenter(model, **seg_kw)
    F = wl["frames"]
    # sharded: every rank holds the SAME clip (its frames are what is distributed); replicas: every rank its own clip
    host = [t.pin_memory() for t in make_clip(wl, cfg, 1 if one_clip else 1 + rank)]
    devt = [t.to(dev) for t in host]
    seed = 1
    vkw = dict(num_video_frames=F) if is_video(cfg) else {}
    kw_of = lambda ts: (dict(vkw, y=ts[3]) if vkw else {})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph_launches = lambda: seg.graph_kernel_launches + (sh.local.graph_kernel_launches if one_clip else 0)

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            _lib.profile_enable(True)
        launches0 = _lib.launch_count() + graph_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = None
        if profile:
            _lib.profile_enable(False)
            prof = _lib.profile_read()
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, _lib.launch_count() + graph_launches() - launches0, prof, out

    step_dev = lambda: seg.segment(devt[0], devt[1], devt[2], F, seed, **kw_of(devt))
    clip_dev = (devt[0], devt[1], devt[2], kw_of(devt))
    clip_host = (host[0], host[1], host[2], kw_of(host))
    many = lambda clip, n, to_host: [r for r in seg.segment_many([clip] * n, F, seed, to_host=to_host)]
    extra = {}
    parity = {}
    if one_clip:
        from vidseg_diffusion_b200.distributed import ShardedClipSegmenter
        sh = ShardedClipSegmenter(model, use_cuda_graph=not args.no_graph, kmeans_split=args.kmeans_split, **seg_kw)
        sh_eager = ShardedClipSegmenter(model, kmeans_split=args.kmeans_split, **seg_kw)
        sh_dev = lambda: sh.segment(devt[0], devt[1], devt[2], F, seed, **kw_of(devt))

        def sh_e2e():
            d = [t.to(dev, non_blocking=True) for t in host]
            return sh.segment(d[0], d[1], d[2], F, seed, **kw_of(d)).cpu()
        for _ in range(max(args.warmup, 3)):
            sh_dev()
        sh_many = lambda clip, n, to_host: [r for r in sh.segment_many([clip] * n, F, seed, to_host=to_host)]
        if args.pipelined and mode == "sharded":
            # clips streamed through the split: the rank-local UNet stage of clip i+1 overlaps the collectives and the
            # distributed K-means of clip i (same results; one step = one clip)
            sh_many(clip_dev[:3], 2, False)
            sampler = ClockSampler(local_rank)
            sampler.start()
            ms, launches, _, outs = timed(lambda: sh_many(clip_dev[:3], args.steps, False), 1)
            clocks = sampler.finish()
            labels_dev = outs[-1]
            info = dict(sh.last["kmeans_info"])
            sh_many(clip_host[:3], 2, True)
            ms_e2e, _, _, outs = timed(lambda: sh_many(clip_host[:3], args.steps, True), 1)
            labels = outs[-1]
            ms_lat, _, _, _ = timed(sh_dev, min(args.steps, 3))
            ms_lat /= min(args.steps, 3)
        else:
            sampler = ClockSampler(local_rank)
            sampler.start()
            ms, launches, _, labels_dev = timed(sh_dev, args.steps)
            clocks = sampler.finish()
            info = dict(sh.last["kmeans_info"])
            sh_e2e()
            ms_e2e, _, _, labels = timed(sh_e2e, args.steps)
            ms_lat = ms / args.steps
        step_prof = lambda: sh_eager.segment(devt[0], devt[1], devt[2], F, seed, **kw_of(devt))
        # parity of the split: every rank's label maps equal rank 0's, and equal the single-GPU path on the whole clip
        ref0 = labels_dev.clone()
        dist.broadcast(ref0, src=0)
        single = seg_eager.segment(devt[0], devt[1], devt[2], F, seed, **kw_of(devt))[0]
        flags = torch.tensor([int(torch.equal(ref0, labels_dev)), int(torch.equal(single, labels_dev))], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        parity = {"labels_identical_on_all_ranks": bool(flags[0].item()),
                  "labels_identical_to_single_gpu_path": bool(flags[1].item()),
                  "lloyd_iterations": info.get("iterations"), "allreduces_per_clip": info.get("allreduces"),
                  "kmeans_split": args.kmeans_split, "exchange_words": info.get("exchange"),
                  "unsharded_fallback": info.get("unsharded_fallback")}
        # the zero-communication alternative (one clip per GPU), for comparison: a few pipelined steps
        rep_host = [t.pin_memory() for t in make_clip(wl, cfg, 1 + rank)]
        rep_clip = (rep_host[0], rep_host[1], rep_host[2], kw_of(rep_host))
        n_rep = max(2, min(args.steps, 5))
        many(rep_clip, 2, True)
        ms_rep, _, _, _ = timed(lambda: many(rep_clip, n_rep, True), 1)
        extra["replicas"] = {"value": world * F * n_rep / (ms_rep / 1e3), "unit": UNIT, "steps": n_rep, "clips_per_step": world,
                             "note": "one clip per GPU per step through segment_many with host buffers (weak scaling, no collective)"}
        clips_per_step = 1
    else:
        for _ in range(max(args.warmup, 3)):
            step_dev()
        if args.pipelined:
            # the stream API: K clips per call, the UNet stage of clip i+1 overlaps the clustering of clip i (same kernels,
            # same results; one step = one clip, every step copies its inputs in and its label maps out in the e2e form)
            many(clip_dev, 2, False)
            sampler = ClockSampler(local_rank)
            sampler.start()
            ms, launches, _, _ = timed(lambda: many(clip_dev, args.steps, False), 1)
            clocks = sampler.finish()
            many(clip_host, 2, True)
            ms_e2e, _, _, outs = timed(lambda: many(clip_host, args.steps, True), 1)
            labels = outs[-1]
            ms_lat, _, _, _ = timed(step_dev, min(args.steps, 3))
            ms_lat /= min(args.steps, 3)
        else:
            step_e2e = lambda: seg.segment_host(host[0], host[1], host[2], F, seed, **kw_of(host))
            sampler = ClockSampler(local_rank)
            sampler.start()
            ms, launches, _, _ = timed(step_dev, args.steps)
            clocks = sampler.finish()
            step_e2e()
            ms_e2e, _, _, labels = timed(step_e2e, args.steps)
            ms_lat = ms / args.steps
        step_prof = lambda: seg_eager.segment(devt[0], devt[1], devt[2], F, seed, **kw_of(devt))
        clips_per_step = world
    # per-kernel pass: the SAME step, launched eagerly with every library launch bracketed by CUDA events on its own
    # stream (vidseg_profile_*); the event pairs add host work, so this pass feeds the roofline / stage split only
    step_prof()
    ms_prof, _, prof, _ = timed(step_prof, args.steps, profile=True)
    fps = clips_per_step * F * args.steps / (ms / 1e3)
    fps_e2e = clips_per_step * F * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = labels.numel() * labels.element_size()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = read_peaks()
    # dominant KERNEL of the step.  Profile families map onto kernels: "gemm" and "conv" are both launches of
    # gemm_split_kernel (the Linear and the implicit-GEMM convolution form of one tcgen05 kernel), "attention" is
    # attn_split_kernel; "kmeans" / "elementwise" / "refine" are groups of several small kernels and are listed in
    # stage_ms_per_step, not as one kernel.
    kernels = {
        "gemm_split_kernel (tcgen05 split-fp16 GEMM + implicit-GEMM conv)": ("tensor", ["gemm", "conv"]),
        "attn_split_kernel (tcgen05 flash attention)": ("tensor", ["attention"]),
        "aggregate_normalize_kernel": ("hbm", ["aggregate"]),
    }
    agg = {}
    for name, (bound, fams) in kernels.items():
        agg[name] = {"bound": bound, "ms": sum(prof[f]["ms"] for f in fams), "work": sum(prof[f]["work"] for f in fams),
                     "launches": sum(prof[f]["launches"] for f in fams)}
    kname = max(agg, key=lambda k: agg[k]["ms"])
    p = agg[kname]
    traffic, traffic_src = None, None
    for tname in ("r02_dominant_kernel_traffic.json", "r01_dominant_kernel_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
            try:
                tj = json.load(open(tpath))
                if tj.get("workload") == args.workload and tj.get("kernel", "") in kname and world == 1:
                    traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/" + tname
                    break
            except Exception:
                pass
    # MMA units per algorithmic product: 2 with fp16 + fp8-correction operands (the default policy), 3 with fp16 pairs
    mma_units = 2.0 if (_lib.load().vidseg_get_operand_mode() != 0 and "gemm_split" in kname) else 3.0
    if p["bound"] == "tensor":
        achieved = p["work"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0.0
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops"], "traffic": traffic,
                "tensor_pipe_tflops": mma_units * achieved, "tensor_pipe_frac": mma_units * achieved / peaks["tflops"]}
    else:
        achieved = p["work"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0.0
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic}
    roof.update(kernel=kname, launches_per_step=p["launches"] / args.steps,
                avg_launch_us=1e3 * p["ms"] / max(p["launches"], 1), share_of_step=p["ms"] / ms_prof,
                timed_over=f"{args.steps} eagerly launched steps with per-launch CUDA events ({ms_prof / args.steps:.1f} ms/step; the "
                           f"headline value replays the UNet stage as one CUDA graph)",
                peak_source=peaks["source"], traffic_source=traffic_src,
                note="achieved = algorithmic FLOPs (2MNK of the fp32-equivalent product) / CUDA-event time over all "
                     "launches of the kernel in the profiled pass; the parity bar (1e-3 vs the fp32 reference) needs more "
                     f"than one fp16 MMA per product: each product costs {mma_units:.0f} fp16-MMA units (fp16 value + "
                     "two correction products, run as fp8 MMAs at twice the rate under the default operand policy), so the "
                     f"tensor pipe runs at tensor_pipe_tflops = {mma_units:.0f} x achieved and frac can not exceed 1/{mma_units:.0f}; "
                     "traffic = dram bytes per launch from the committed ncu launch list")
    lib_ms = sum(v["ms"] for v in prof.values())
    breakdown = {k: round(v["ms"] / args.steps, 3) for k, v in prof.items() if v["launches"]}
    breakdown["non_library(torch glue + host gaps)"] = round((ms_prof - lib_ms) / args.steps, 3)

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "timed at N=1 only"}
    if args.no_cpu_baseline:
        cpu["sample"] = "skipped (--no-cpu-baseline)"
    elif world == 1:
        from oracle import reference_path as rp
        if not rp.available():
            cpu["sample"] = "oracle/_ref missing (run python oracle/make_ref.py where /root/reference exists)"
        else:
            # ONE full pass of the unmodified reference over the SAME clip: the CPU baseline and the feature-level parity
            ref = cpu_reference_step(wl, cfg, sd, [t.clone() for t in make_clip(wl, cfg, 1)], seed, keep=True)
            cpu.update(value=F / ref["seconds"]["total"], sample="one pass, " + reference_sample_text(wl, ref["seconds"]))
            lab_gpu, out_gpu = seg_eager.segment(devt[0], devt[1], devt[2], F, seed, **kw_of(devt))
            q_gpu = {i: q.cpu() for i, q in zip((8, 7, 6), harvest_self_attn_q(model, (8, 7, 6)))}
            feat = {f"attn1.q out_block_{i}": relerr(q_gpu[i], ref["q"][i]) for i in (6, 7, 8)}
            feat["unet_output"] = relerr(out_gpu, ref["out"])
            parity.update(features_rel_err_vs_reference=feat, features_max_rel_err=max(feat.values()), features_tol=1e-3,
                          features_ok=max(feat.values()) <= 1e-3,
                          partition_agreement_gpu_path_vs_reference_path=partition_agreement(
                              lab_gpu.cpu().numpy().reshape(-1), ref["labels"].reshape(-1), wl["num_masks"]),
                          partition_agreement_note="fraction of cells on which the two end-to-end paths agree after the best "
                          "one-to-one matching of cluster ids; they cluster DIFFERENT feature tensors (equal to "
                          "features_max_rel_err, not bit-equal) and K-means numbers its clusters arbitrarily, so this is "
                          "information -- the bit-exact check is label_mismatches (same features in, same labels out)")
            del ref
    if world == 1 and not args.no_parity:
        # integer stage: the label maps of the TIMED call against the reference's clustering arithmetic run on the CPU
        # on the very features the GPU clustered
        lab_gpu, _ = seg_eager.segment(devt[0], devt[1], devt[2], F, seed, **kw_of(devt))
        timed_labels = labels.cpu().numpy().reshape(-1)
        X = seg_eager.last["features"].cpu().numpy()
        q7 = harvest_self_attn_q(model, (7,))[0].cpu().numpy() if wl["refine"] else None
        t0 = time.perf_counter()
        want, how = label_oracle(X, q7, wl, seed)
        parity.update(label_mismatches=int((timed_labels != want).sum()), label_cells=int(want.size), label_checker=how,
                      label_checker_seconds=round(time.perf_counter() - t0, 1),
                      timed_call_equals_eager_call=bool((timed_labels == lab_gpu.cpu().numpy().reshape(-1)).all()))
    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if one_clip else "weak", "vs_baseline": None,
        "dtype": "f32 (fp16 tensor-core operands with fp8/fp16 correction terms, fp32 accumulate; parity 1e-3 vs the fp32 reference)",
        "data": "synthetic",
        "config": bench_config(args, wl, world, mode),
        "impl_detail": {"unet_stage": "eager launches" if args.no_graph else "one CUDA graph (UNet + harvest + aggregate/normalise)",
                        "schedule": ("ShardedClipSegmenter.segment_many: clips streamed through the split, UNet stage of clip i+1 "
                                     "over the collectives + K-means of clip i" if (one_clip and args.pipelined and mode == "sharded") else
                                     "ShardedClipSegmenter.segment: one clip at a time over all ranks" if one_clip else
                                     "segment_many: clips software-pipelined over two CUDA streams" if args.pipelined
                                     else "segment: one clip at a time, stages back to back")},
        "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "latency_ms_per_clip": ms_lat,
        "roofline": roof,
        "stage_ms_per_step": breakdown,
        "cpu_baseline": cpu,
        "parity": parity,
        "clocks": clocks,
    }
    line.update(extra)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_modulated(args, wl, cfg):
    """Opt-in workload: the whole single-clip pipeline behind the per-clip hot path, everything resident in HBM."""
    import numpy as np
    import torch
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.features import aggregate_normalize
    from vidseg_diffusion_b200.kmeans import KMeans
    from vidseg_diffusion_b200.modulation import modulated_segmentation
    from vidseg_diffusion_b200.refine import refine_masks
    from vidseg_diffusion_b200.sgm.models.autoencoder import AutoencoderKL
    from vidseg_diffusion_b200.sgm.models.diffusion import FirstStage
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    from vidseg_diffusion_b200.sgm.util import instantiate_from_config
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    dev = torch.device("cuda", 0)
    _lib.load()
    F, K = wl["frames"], wl["num_masks"]
    sd = make_state_dict(cfg)
    with torch.device("meta"):
        model = model_class(cfg)(**cfg)
    model = model.to_empty(device=dev)
    model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    model.eval()
    ddconfig = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128 if wl["cfg"] == "sd21" else 64,
                    ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=(), dropout=0.0)   # sd_2_1.yaml:49-59
    with torch.device("meta"):
        vae = AutoencoderKL(embed_dim=4, ddconfig=ddconfig)
    g = torch.Generator().manual_seed(7)
    vsd = {}
    for key, t in vae.state_dict().items():
        shape = tuple(t.shape)
        fan_in = int(np.prod(shape[1:])) if len(shape) >= 2 else 1
        vsd[key] = (torch.randn(shape, generator=g) * fan_in ** -0.5 if len(shape) >= 2 else
                    (1.0 + 0.1 * torch.randn(shape, generator=g) if key.endswith(".weight") else 0.05 * torch.randn(shape, generator=g)))
    vae = vae.to_empty(device=dev)
    vae.load_state_dict({k: v.to(dev) for k, v in vsd.items()}, strict=True)
    engine = FirstStage(vae.eval(), scale_factor=0.18215, en_and_decode_n_samples_a_time=F)
    ddpm = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    steps, t_start = 25, 17
    smp = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler",
        "params": {"discretization_config": ddpm, "num_steps": steps, "s_churn": 0, "s_tmin": 0, "s_tmax": 999, "s_noise": 1,
                   "device": str(dev),
                   "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": 5.0}}}})
    den = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser",
        "params": {"num_idx": 1000, "discretization_config": ddpm,
                   "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}}}).to(dev)
    denoiser = den.bind(OpenAIWrapper(model))
    x, _, ctx = make_clip(wl, cfg, 1)
    latent, ctx = x[:F].to(dev), ctx[F:].to(dev)
    c, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}
    fh = wl["latent"] // 2
    blocks = (8, 7, 6)

    def one_clip():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        store = {}

        def save_cb(xt, i):   # svd_single_video_inference.py:113-130 without the .pt files
            store[f"xt_time_{i}"] = xt.clone()
            if i == steps - 1:
                for b in blocks:
                    store[("q", b)] = model.output_blocks[b][1].transformer_blocks[0].attn1.q.clone()

        ev[0].record()
        smp(denoiser, latent.clone(), cond=c, uc=uc, img_callback=save_cb, t_start=t_start)          # source run
        ev[1].record()
        X = aggregate_normalize([store[("q", b)] for b in blocks], F)
        np.random.seed(1)
        labels = KMeans(n_clusters=K, n_init=10).fit_predict(X).reshape(F, fh, fh)
        labels, _, _ = refine_masks(store[("q", 7)], labels, F, fh, fh)
        ev[2].record()
        res = modulated_segmentation(smp, denoiser, engine, latent, c, uc, labels.to(torch.int32).contiguous(), np.arange(K),
                                     num_steps=steps, t_start=t_start, modulate_block_idx=(8,), modulate_timestep=(17,),
                                     modulate_layer_type=("spatial", "temporal"), modulate_attn_type=("self_attn",),
                                     is_latent_blending=True, features=store)
        ev[3].record()
        torch.cuda.synchronize()
        return res, [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]

    for _ in range(args.warmup):
        one_clip()
    launches0 = _lib.launch_count()
    total, stages = 0.0, [0.0, 0.0, 0.0]
    for _ in range(args.steps):
        res, ms = one_clip()
        total += sum(ms)
        stages = [a + b for a, b in zip(stages, ms)]
    fps = F * args.steps / (total / 1e3)
    seg = res["seg_raw"]
    unet_evals = (steps - t_start) * (1 + 2 * K)
    emit({"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32 (fp16 tensor-core operands with fp8/fp16 correction terms, fp32 accumulate)", "data": "synthetic",
          "config": bench_config(args, wl, 1, "single"),
          "stage_ms_per_step": {"source_run(8 sampler steps)": stages[0] / args.steps,
                                "kmeans+refine": stages[1] / args.steps,
                                f"{2 * K} modulated runs + decode + post-process": stages[2] / args.steps},
          "unet_evaluations_per_step": unet_evals, "gpu_launches": int(_lib.launch_count() - launches0),
          "result": {"seg_raw_shape": list(seg.shape), "labels_used": int(torch.unique(seg).numel())},
          "e2e": None, "roofline": None,
          "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                           "sample": "not timed: the reference needs ~330 CPU UNet evaluations (hours) for this workload"}})


_REAL_STDOUT = None


def guard_stdout():
    """stdout carries exactly one JSON line: everything libraries print at the C level (the NCCL version banner, cuDNN /
    driver notices) or through Python's sys.stdout goes to stderr; ``emit`` writes to the original descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch the UNet stage eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-pipeline", dest="pipelined", action="store_false",
                    help="time ClipSegmenter.segment (one clip at a time) instead of segment_many, where the UNet stage of "
                         "clip i+1 overlaps the clustering of clip i on a second stream")
    ap.add_argument("--kmeans-split", default="runs", choices=["runs", "rows"],
                    help="sharded / cfg-split: 'runs' = the n_init initialisations spread over the ranks, one all-reduce at the "
                         "end; 'rows' = every rank assigns its own rows, one all-reduce per Lloyd iteration")
    ap.add_argument("--multi-gpu", default="auto", choices=["auto", "sharded", "cfg-split", "replicas"],
                    help="N > 1: 'sharded' = ONE clip per step with its frames sharded over the ranks (all-gather + distributed "
                         "K-means; default for SD-2.1), 'cfg-split' = the two guidance halves of an SVD clip on two ranks "
                         "(default for SVD at N = 2), 'replicas' = one clip per GPU, no collective")
    ap.add_argument("--no-parity", action="store_true", help="skip the CPU check of the timed call's label maps")
    ap.add_argument("--ref-budget-s", type=float, default=1500.0,
                    help="wall budget of the reference arm: the step COUNT is cut to fit it, never the workload")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    from vidseg_diffusion_b200 import configs
    cfg = {"sd21": configs.SD21_UNET, "tiny": configs.TINY_UNET, "svd": configs.SVD_UNET,
           "tinyv": configs.TINY_VIDEO_UNET}[wl["cfg"]]
    if wl.get("modulated"):
        run_modulated(args, wl, cfg)
    elif args.impl == "reference":
        run_reference(args, wl, cfg)
    else:
        run_b200(args, wl, cfg)


if __name__ == "__main__":
    main()
