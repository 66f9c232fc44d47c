"""Per-clip hot path as one call: UNet step with Q/K stash -> aggregate/normalise -> K-means -> refine.

This is the in-memory form of what ``scripts/sampling/svd_single_video_inference.py:sample`` does
between the sampler callback (:113-149, ``torch.save`` of every stashed tensor) and the mask stage
(:357-403, ``feature_extraction_main`` reading them back): the stashed ``attn1.q`` tensors stay in
HBM and are handed to the clustering kernels directly.  Flag names follow the reference CLI
(``--num_masks``, ``--is_aggre_attn``, ``--is_refine_mask``, ``--seed``).
"""
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

    def __init__(self, model, num_masks=10, is_aggre_attn=False, is_refine_mask=False, n_init=10):
        self.model = model
        self.num_masks = int(num_masks)
        self.blocks = AGGRE_BLOCKS if is_aggre_attn else SINGLE_BLOCK
        self.is_refine_mask = bool(is_refine_mask)
        self.n_init = n_init
        self.last = {}

    @torch.no_grad()
    def unet_step(self, x, timesteps, context, **unet_kwargs):
        """x [2F, C, h, w] (uncond rows first, guiders.py:38-42), timesteps [2F], context [2F, L, D]."""
        return self.model(x, timesteps=timesteps, context=context, **unet_kwargs)

    def cluster(self, num_frames, feature_height, feature_width, seed=None):
        """K-means label maps [F, h, w] (int32, device) from the features stashed by the last UNet call."""
        feats = harvest_self_attn_q(self.model, self.blocks)
        x = aggregate_normalize(feats, num_frames)
        if seed is not None:
            np.random.seed(seed)  # seed_everything(), svd_single_video_inference.py:590-594
        km = KMeans(n_clusters=self.num_masks, n_init=self.n_init)
        labels = km.fit_predict(x).reshape(num_frames, feature_height, feature_width)
        self.last = {"kmeans": km, "features": x}
        return labels

    def refine(self, labels, num_frames, feature_height, feature_width):
        fm = harvest_self_attn_q(self.model, (REFINE_BLOCK,))[0]
        refined, traj, keep = refine_masks(fm, labels, num_frames, feature_height, feature_width)
        self.last.update(trajectories=traj, keep=keep)
        return refined

    @torch.no_grad()
    def segment(self, x, timesteps, context, num_frames, seed=None, **unet_kwargs):
        """Whole path on device tensors.  Returns (label maps int32 [F, h, w], UNet output)."""
        out = self.unet_step(x, timesteps, context, **unet_kwargs)
        fh, fw = x.shape[-2] // 2, x.shape[-1] // 2   # H // (8 * 2): feature grid of output blocks 6-8
        labels = self.cluster(num_frames, fh, fw, seed)
        if self.is_refine_mask:
            labels = self.refine(labels, num_frames, fh, fw)
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
