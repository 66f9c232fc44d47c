"""Per-clip hot path as one call: UNet step with Q/K stash -> aggregate/normalise -> K-means -> refine.

This is the in-memory form of what ``scripts/sampling/svd_single_video_inference.py:sample`` does
between the sampler callback (:113-149, ``torch.save`` of every stashed tensor) and the mask stage
(:357-403, ``feature_extraction_main`` reading them back): the stashed ``attn1.q`` tensors stay in
HBM and are handed to the clustering kernels directly.  Flag names follow the reference CLI
(``--num_masks``, ``--is_aggre_attn``, ``--is_refine_mask``, ``--seed``).
"""
import contextlib

import numpy as np
import torch

from . import _lib
from .features import aggregate_normalize
from .kmeans import KMeans
from .refine import refine_masks

AGGRE_BLOCKS = (8, 7, 6)   # svd_single_video_inference.py:362, in this order
SINGLE_BLOCK = (8,)        # :364
REFINE_BLOCK = 7           # :393


def harvest_self_attn_q(model, blocks):
    """What ``save_feature_maps`` (svd_single_video_inference.py:113-125) reads after a UNet call."""
    feats = []
    for i in blocks:
        layer = model.output_blocks[i][1]
        kind = str(type(layer))  # the reference's own test: sd_pipeline_vspw.py:112 / svd_single_video_inference.py:117
        if "SpatialTransformer" not in kind and "SpatialVideoTransformer" not in kind:
            raise _lib.VidsegError(f"output_blocks[{i}][1] is not a Spatial(Video)Transformer")
        q = layer.transformer_blocks[0].attn1.q
        if q is None:
            raise _lib.VidsegError(f"output_blocks[{i}] has no stashed q: run the UNet first")
        feats.append(q)
    return feats


class ClipSegmenter:
    """One object per process / GPU.  ``model`` is a ``UNetModel`` (or ``VideoUNet``) on the GPU."""

    def __init__(self, model, num_masks=10, is_aggre_attn=False, is_refine_mask=False, n_init=10, use_cuda_graph=False,
                 stash="harvested"):
        """``stash``: "harvested" -- during this object's UNet calls only the attention layers whose q it reads (output
        blocks 8 / 7 / 6, ``attn1``) write the fp32 stash; every other layer's ``.q`` / ``.k`` is decoded from the
        fp16-pair operand of the attention kernel on first access (22 significant bits) instead of costing an fp32 write
        per projection.  "all" -- every layer writes its fp32 stash, as the reference's attributes do."""
        if stash not in ("harvested", "all"):
            raise _lib.VidsegError(f"stash must be 'harvested' or 'all', got {stash!r}")
        self.stash = stash
        self.model = model
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graphs = {}
        self.graph_replays = 0
        self.graph_kernel_launches = 0
        self.num_masks = int(num_masks)
        self.blocks = AGGRE_BLOCKS if is_aggre_attn else SINGLE_BLOCK
        self.is_refine_mask = bool(is_refine_mask)
        self.n_init = n_init
        self.last = {}

    @contextlib.contextmanager
    def _stash_scope(self):
        """fp32 stash only where this object harvests it, for the duration of one of its own UNet calls."""
        if self.stash == "all":
            yield
            return
        keep = set(self.blocks) | ({REFINE_BLOCK} if self.is_refine_mask else set())
        keep_ids = {id(self.model.output_blocks[i][1].transformer_blocks[0].attn1) for i in keep}
        touched = []
        for m in self.model.modules():
            if hasattr(m, "stash_f32") and id(m) not in keep_ids and m.stash_f32:
                m.stash_f32 = False
                touched.append(m)
        try:
            yield
        finally:
            for m in touched:
                m.stash_f32 = True

    @torch.no_grad()
    def unet_step(self, x, timesteps, context, **unet_kwargs):
        """x [2F, C, h, w] (uncond rows first, guiders.py:38-42), timesteps [2F], context [2F, L, D]."""
        with self._stash_scope():
            return self.model(x, timesteps=timesteps, context=context, **unet_kwargs)

    def _features(self, num_frames, cond_only=False):
        return aggregate_normalize(harvest_self_attn_q(self.model, self.blocks), num_frames, cond_only=cond_only)

    def _graphed_unet_features(self, x, timesteps, context, num_frames, unet_kwargs, features="cond_half"):
        """UNet step + harvest + aggregate/normalise as ONE CUDA graph per input signature: ~500 kernel launches of
        the step replay without host work in between.  Inputs are copied into the graph's static buffers; the stashed
        q/k tensors, the UNet output and the feature matrix live in the graph's memory pool and are overwritten by
        every replay (clone what must outlive the next call).  ``features``: "cond_half" (the batch holds both guidance
        halves, rows [F, 2F) are clustered), "all_rows" (the batch IS the conditional half: CFG-half split over two GPUs)
        or None (the unconditional rank of that split: no features)."""
        feat = (lambda: None) if features is None else (lambda: self._features(num_frames, cond_only=(features == "all_rows")))
        tens = {k: v for k, v in unet_kwargs.items() if isinstance(v, torch.Tensor)}
        rest = {k: v for k, v in unet_kwargs.items() if not isinstance(v, torch.Tensor)}
        key = (tuple(x.shape), tuple(timesteps.shape), tuple(context.shape), num_frames,
               tuple(sorted((k, tuple(v.shape)) for k, v in tens.items())), tuple(sorted(rest.items())), features)
        ent = self._graphs.get(key)
        if ent is None:
            static = dict(x=x.clone(), t=timesteps.clone(), c=context.clone(), **{"kw_" + k: v.clone() for k, v in tens.items()})
            kw = dict(rest, **{k: static["kw_" + k] for k in tens})
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), self._stash_scope():   # warm-up outside capture: weight operand caches, workspaces
                for _ in range(2):
                    self.model(static["x"], timesteps=static["t"], context=static["c"], **kw)
                    feat()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph), self._stash_scope():
                out = self.model(static["x"], timesteps=static["t"], context=static["c"], **kw)
                feats = feat()
                # the stash the refinement reads: the graph's own buffer, not whatever the module attribute points at
                # after a later eager forward or another capture
                q7 = harvest_self_attn_q(self.model, (REFINE_BLOCK,))[0] if self.is_refine_mask else None
            ent = self._graphs[key] = (graph, static, out, feats, _lib.launch_count() - n0, q7)
        graph, static, out, feats, n_launches, q7 = ent
        self.graph_replays += 1
        self.graph_kernel_launches += n_launches   # library kernels inside the replayed graph (not counted by the host-side counter)
        static["x"].copy_(x, non_blocking=True)
        static["t"].copy_(timesteps, non_blocking=True)
        static["c"].copy_(context, non_blocking=True)
        for k, v in tens.items():
            static["kw_" + k].copy_(v, non_blocking=True)
        graph.replay()
        return out, feats, q7

    def cluster(self, num_frames, feature_height, feature_width, seed=None, features=None):
        """K-means label maps [F, h, w] (int32, device) from the features stashed by the last UNet call."""
        x = self._features(num_frames) if features is None else features
        if seed is not None:
            np.random.seed(seed)  # seed_everything(), svd_single_video_inference.py:590-594
        km = KMeans(n_clusters=self.num_masks, n_init=self.n_init)
        labels = km.fit_predict(x).reshape(num_frames, feature_height, feature_width)
        self.last = {"kmeans": km, "features": x}
        return labels

    def refine(self, labels, num_frames, feature_height, feature_width, features=None):
        fm = harvest_self_attn_q(self.model, (REFINE_BLOCK,))[0] if features is None else features
        refined, traj, keep = refine_masks(fm, labels, num_frames, feature_height, feature_width)
        self.last.update(trajectories=traj, keep=keep)
        return refined

    @torch.no_grad()
    def segment(self, x, timesteps, context, num_frames, seed=None, **unet_kwargs):
        """Whole path on device tensors.  Returns (label maps int32 [F, h, w], UNet output)."""
        feats = q7 = None
        if self.use_cuda_graph:
            out, feats, q7 = self._graphed_unet_features(x, timesteps, context, num_frames, unet_kwargs)
        else:
            out = self.unet_step(x, timesteps, context, **unet_kwargs)
        fh, fw = x.shape[-2] // 2, x.shape[-1] // 2   # H // (8 * 2): feature grid of output blocks 6-8
        labels = self.cluster(num_frames, fh, fw, seed, features=feats)
        if self.is_refine_mask:
            labels = self.refine(labels, num_frames, fh, fw, features=q7)
        return labels, out

    @torch.no_grad()
    def segment_host(self, x_host, timesteps_host, context_host, num_frames, seed=None, **unet_kwargs):
        """End-to-end call with HOST buffers (pinned for async copies): H2D of the step's inputs, the
        whole path, D2H of the label maps.  Returns a CPU int32 tensor [F, h, w]."""
        dev = next(self.model.parameters()).device
        x = x_host.to(dev, non_blocking=True)
        t = timesteps_host.to(dev, non_blocking=True)
        c = context_host.to(dev, non_blocking=True)
        unet_kwargs = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in unet_kwargs.items()}
        labels, _ = self.segment(x, t, c, num_frames, seed, **unet_kwargs)
        return labels.cpu()

    @torch.no_grad()
    def segment_many(self, clips, num_frames, seed=None, to_host=True):
        """A stream of clips, software-pipelined: while the clustering (K-means polls its convergence flags from the host,
        many small latency-bound launches) and the refinement of clip i run on a second CUDA stream, the UNet stage of
        clip i+1 -- one CUDA-graph launch -- already runs on the first.  Results are identical to ``segment`` /
        ``segment_host`` clip by clip; only the schedule changes.

        ``clips``: iterable of ``(x, timesteps, context)`` or ``(x, timesteps, context, unet_kwargs)``; host tensors
        (pinned) are copied to the device inside.  Yields one label map per clip, in order (CPU int32 if ``to_host``)."""
        dev = next(self.model.parameters()).device
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=dev, priority=-1)   # the small K-means kernels go first when SMs free up
        side = self._side_stream
        main = torch.cuda.current_stream(dev)
        to_dev = lambda v: v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v

        def finish(pending):
            feats, fm7, ready, fh, fw = pending
            with torch.cuda.stream(side):
                side.wait_event(ready)
                labels = self.cluster(num_frames, fh, fw, seed, features=feats)
                if self.is_refine_mask:
                    labels = self.refine(labels, num_frames, fh, fw, features=fm7)
                out = labels.cpu() if to_host else labels
                done = torch.cuda.Event()
                done.record(side)
            if not to_host:
                main.wait_event(done)
            return out

        pending = None
        for clip in clips:
            x, t, c = (to_dev(v) for v in clip[:3])
            kw = {k: to_dev(v) for k, v in (clip[3] if len(clip) > 3 else {}).items()}
            q7 = None
            if self.use_cuda_graph:
                _, feats, q7 = self._graphed_unet_features(x, t, c, num_frames, kw)
            else:
                self.unet_step(x, t, c, **kw)
                feats = self._features(num_frames)
                if self.is_refine_mask:
                    q7 = harvest_self_attn_q(self.model, (REFINE_BLOCK,))[0]
            # the graph's buffers (and the modules' stash) are overwritten by the next clip: snapshot what the second
            # stream will read
            feats = feats.clone()
            feats.record_stream(side)
            fm7 = None
            if self.is_refine_mask:
                fm7 = q7.clone()
                fm7.record_stream(side)
            ready = torch.cuda.Event()
            ready.record(main)
            if pending is not None:
                yield finish(pending)
            pending = (feats, fm7, ready, x.shape[-2] // 2, x.shape[-1] // 2)
        if pending is not None:
            yield finish(pending)
