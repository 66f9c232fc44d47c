"""TEST INFRASTRUCTURE ONLY -- the UNMODIFIED reference's own CPU path for one clip of the hot path.

Drives the reference code exactly as ``scripts/sampling/svd_single_video_inference.py:sample`` does between the
sampler callback and the mask stage, on the CPU in fp32 (the "reference CPU/PyTorch path" of the north star):

  1. ``UNetModel.forward`` / ``VideoUNet.forward`` (openaimodel.py:831-954 / video_model.py:451-566), the reference's
     own nn.Modules imported through ``oracle/ref_import.py`` from /root/reference or its mirror ``oracle/_ref``;
  2. ``save_feature_map`` (svd_single_video_inference.py:140-149): ``torch.save`` of the stashed ``attn1.q`` tensors the
     mask stage reads, in the reference's ``feature_maps`` layout (tmpfs when /dev/shm exists);
  3. ``feature_extraction_main("kmeans_masks", ...)`` (feature_extraction.py:670-795 -> :30-113, scikit-learn KMeans,
     PNG tree), with ``np.random.seed(seed)`` where ``seed_everything`` puts it (:590-594);
  4. with ``--is_refine_mask``: ``feature_extraction_main("correct_low_res_mask", ...)`` (:367-461), ``device="cuda"``
     defaults of ``dense_tracking`` / ``dense_feature_matching_iterative`` patched to "cpu" (SURVEY.md section 8c).

Users: ``bench.py --impl reference``, ``bench.py``'s ``cpu_baseline`` / ``parity`` legs, ``tests/test_gpu_fullsize.py``.
Nothing in the product package imports this module.
"""
import os
import shutil
import tempfile
import time
import warnings

import numpy as np
import torch

from .ref_import import import_reference, reference_available

AGGRE_BLOCKS = (8, 7, 6)       # svd_single_video_inference.py:362
TIMESTEP = 24                  # feature_timestep default; generate_aggregate_mask hard-codes it (:385)


def available():
    return reference_available()


def build_model(cfg):
    """The reference's own module for ``cfg`` (vidseg_diffusion_b200.configs), default-initialised.  ``use_checkpoint``
    is a no-op without grad; SVD asks for ``softmax-xformers`` (svd.yaml:27), xformers is absent from this image and the
    reference's SDPA class computes the same function."""
    kw = dict(cfg)
    kw["use_checkpoint"] = False
    if "video_kernel_size" in kw:
        kw["spatial_transformer_attn_type"] = "softmax"
        mod = import_reference("sgm.modules.diffusionmodules.video_model")
        return mod.VideoUNet(**kw).eval()
    mod = import_reference("sgm.modules.diffusionmodules.openaimodel")
    return mod.UNetModel(**kw).eval()


def _feature_extraction():
    fe = import_reference("scripts.sampling.feature_extraction")
    for fn in (fe.dense_tracking, fe.dense_feature_matching_iterative):
        fn.__defaults__ = tuple("cpu" if d == "cuda" else d for d in fn.__defaults__)
    return fe


def scratch_root():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix="vidseg_ref_", dir=base)


@torch.no_grad()
def unet_forward(model, clip, num_frames):
    """clip = (x, timesteps, context[, y]) CPU fp32 tensors at batch 2F (uncond rows first).  Returns the UNet output;
    the stashed q / k sit on the modules exactly as the reference's pipelines read them."""
    if len(clip) > 3:
        return model(clip[0], timesteps=clip[1], context=clip[2], y=clip[3], num_video_frames=num_frames,
                     image_only_indicator=torch.zeros(2, num_frames))
    return model(clip[0], timesteps=clip[1], context=clip[2])


def stashed_q(model, block):
    return model.output_blocks[block][1].transformer_blocks[0].attn1.q


def clip_step(model, clip, num_frames, num_masks, aggre=True, refine=False, seed=1, keep=False):
    """One pass of the reference path over one clip.  Returns a dict: ``seconds`` per stage and total, ``labels``
    (int64 [F, h, w], rebuilt from the PNG tree like the reference itself does) and, with ``keep``, the UNet output
    and the stashed q tensors of output blocks 8, 7, 6."""
    fe = _feature_extraction()
    F = num_frames
    fh, fw = clip[0].shape[-2] // 2, clip[0].shape[-1] // 2
    blocks = AGGRE_BLOCKS if aggre else (8,)
    root = scratch_root()
    exp = "clip"
    secs = {}
    try:
        t0 = time.perf_counter()
        out = unet_forward(model, clip, F)
        secs["unet"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        fm_dir = os.path.join(root, exp, "feature_maps")
        os.makedirs(fm_dir)
        for i in sorted(set(blocks) | ({7} if refine else set())):
            torch.save(stashed_q(model, i), os.path.join(fm_dir, f"output_block_{i}_spatial_self_attn_q_time_{TIMESTEP}.pt"))
        secs["save_feature_maps"] = time.perf_counter() - t1
        t2 = time.perf_counter()
        names = ",".join(f"output_block_{i}" for i in blocks)
        np.random.seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            unique_labels, _, _ = fe.feature_extraction_main(
                "kmeans_masks", num_masks, TIMESTEP, names, exp, exp, "spatial_self_attn_q", fh, fw, str(TIMESTEP),
                frame_name_list=None, base_folder=root, num_frames=F)
        secs["kmeans_masks"] = time.perf_counter() - t2
        mask_folder = os.path.join(root, exp, "kmeans_masks", f"{names.replace(',', '_')}_spatial_self_attn_q_masks_{num_masks}")
        if refine:
            t3 = time.perf_counter()
            _, ref_mask, _ = fe.feature_extraction_main(
                "correct_low_res_mask", num_masks, TIMESTEP, "output_block_7", exp, exp, "spatial_self_attn_q", fh, fw,
                str(TIMESTEP), frame_name_list=None, base_folder=root, num_frames=F, mask_folder=mask_folder,
                ref_unique_labels=unique_labels)
            secs["correct_low_res_mask"] = time.perf_counter() - t3
            labels = np.asarray(ref_mask).reshape(F, fh, fw).astype(np.int64)
        else:
            labels = np.stack([fe.generate_aggregate_mask(mask_folder, TIMESTEP, num_masks, i, fh, fw) for i in range(F)]).astype(np.int64)
        secs["total"] = time.perf_counter() - t0
        res = {"seconds": secs, "labels": labels}
        if keep:
            res["out"] = out
            res["q"] = {i: stashed_q(model, i) for i in AGGRE_BLOCKS}
        return res
    finally:
        shutil.rmtree(root, ignore_errors=True)
