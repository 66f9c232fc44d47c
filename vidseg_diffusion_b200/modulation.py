"""The 2 x K modulated sampler runs of a clip and their post-process, in memory.

Mirrors ``scripts/sampling/svd_single_video_inference.py:404-508``: for every K-means label and for +lambda / -lambda the
sampler is run again from the noised latent with that label's masks modulating the attention / feed-forward outputs
(optionally with the source run's q / k injected and the latent blended outside the mask), the result is decoded by the
first stage and turned into uint8 frames; ``get_seg_map_main`` then turns the 2 x K videos into the final label maps.
In the reference every arrow of that chain is a file: ~9 GB of ``.pt`` feature maps, K x F mask PNGs read back per run,
2 x K x F decoded PNGs, K x F difference JPEGs.  Here the source run's q / k / x_t, the label maps, the decoded frames
and the difference images all stay in HBM; the host only sequences the runs.
"""
import numpy as np
import torch

from . import _lib
from .process_output import resized_masks, seg_maps_from_frames


def get_modulate_timestep_frames(start_timestep, end_timestep=None, num_frames=14, schedule="constant"):
    """reference sgm/util.py:313-327."""
    if schedule == "constant":
        return {}
    if schedule != "linear":
        raise ValueError(f"Unknown modulate timestep frames schedule: {schedule}")
    out = {t: [] for t in range(start_timestep, end_timestep - 1, -1)}
    for frame_id in range(num_frames):
        out[int(start_timestep + (end_timestep - start_timestep) * frame_id / (num_frames - 1))].append(frame_id)
    return out


def modulate_grid(modulate_block_idx, base_height, base_width):
    """(height, width) of the token grid of output block ``modulate_block_idx`` as load_feature_masks computes it
    (svd_single_video_inference.py:80-91, including its swapped names for blocks 3-5)."""
    if modulate_block_idx in (0, 1, 2):
        return base_height, base_width
    if modulate_block_idx in (3, 4, 5):
        return base_width * 2, base_height * 2
    if modulate_block_idx in (6, 7, 8):
        return base_height * 4, base_width * 4
    if modulate_block_idx in (9, 10, 11):
        return base_height * 8, base_width * 8
    raise ValueError(f"modulate_block_idx {modulate_block_idx} is not an output block")


def feature_masks_from_labels(label_maps, mask_id, modulate_block_idx, base_height, base_width):
    """load_feature_masks (svd_single_video_inference.py:68-100) without the PNG tree: per frame the 0/255 mask of label
    ``mask_id`` resized to the modulated block's token grid with Pillow's default (BICUBIC) filter, / 255, flattened.
    label_maps: CUDA int32 [F, h, w].  Returns a list of F float64 CUDA tensors [hw']."""
    lab = _lib.require_cuda_tensor(label_maps.contiguous(), torch.int32, "label_maps")
    F, h, w = lab.shape
    gh, gw = modulate_grid(modulate_block_idx, base_height, base_width)
    if (gh, gw) == (h, w):      # the usual case (block 6-8 masks on the block 6-8 grid): Pillow's resize is a copy
        m = (lab == int(mask_id)).to(torch.float64)          # 255 / 255.0 == 1.0 exactly
    elif gh != h and gw != w:
        # v / 255.0 through a table of numpy's own quotients: torch divides by a scalar as a multiplication with its
        # reciprocal on the GPU, which is not the correctly rounded quotient the reference's numpy division gives
        table = torch.from_numpy(np.arange(256, dtype=np.float64) / 255.0).to(lab.device)
        m = table[resized_masks(lab, [int(mask_id)], gh, gw, filter="bicubic")[0].long()]
    else:
        raise _lib.VidsegError("feature_masks_from_labels: a resize of one axis only is not built")
    return [m[f].reshape(-1) for f in range(F)]


def frames_to_uint8(samples_x):
    """svd_single_video_inference.py:164-176: clamp((x + 1) / 2, 0, 1), 't c h w -> t h w c', * 255, astype(uint8)."""
    s = torch.clamp((samples_x + 1.0) / 2.0, min=0.0, max=1.0)
    return (s.permute(0, 2, 3, 1) * 255).to(torch.uint8).contiguous()


@torch.no_grad()
def modulated_runs(sampler, denoiser, first_stage, latent, cond, uc, label_maps, unique_labels, *, num_steps=None, t_start=0,
                   modulate_block_idx=(8,), modulate_timestep=(17,), modulate_schedule="constant",
                   modulate_lambda_start=50.0, modulate_lambda_end=50.0, modulate_layer_type=("spatial", "temporal"),
                   modulate_attn_type=("self_attn",), modulate_timestep_frames_schedule="constant",
                   is_injected_features=False, injected=None, is_latent_blending=False, features=None,
                   feature_folder=None, exp_name=None, base_height=None, base_width=None, is_smooth_latent=False,
                   on_run=None):
    """Step 4 of the reference script (:404-500).  ``latent``: the noised latent the source run started from [F, C, h, w];
    ``label_maps``: the (refined) K-means label maps int32 [F, fh, fw]; ``features``: the source run's q / k / x_t kept in
    HBM (``{"output_block_7_spatial_self_attn_q_time_17": ..., "xt_time_17": ...}``), or None to read the reference's
    ``.pt`` files under ``feature_folder / exp_name``; ``injected``: dict(injected_block_types, injected_feature_types,
    input_block_indices, output_block_indices) (:411-418).  Returns (frames_pos, frames_neg): uint8 [K, F, H, W, 3]."""
    F = latent.shape[0]
    num_steps = sampler.num_steps if num_steps is None else num_steps
    base_height = latent.shape[-2] // 8 if base_height is None else base_height   # H // (F * 8) with F = 8 (:454)
    base_width = latent.shape[-1] // 8 if base_width is None else base_width
    block_idx = [int(b) for b in modulate_block_idx]
    timesteps = [int(t) for t in modulate_timestep]
    mtf = get_modulate_timestep_frames(start_timestep=20, end_timestep=15, num_frames=F,
                                       schedule=modulate_timestep_frames_schedule)
    inj = dict(injected_block_types=None, injected_feature_types=None, input_block_indices=None, output_block_indices=None)
    if is_injected_features:
        inj.update(injected or dict(
            injected_block_types=["output"],
            injected_feature_types=["temporal_cross_attn_k", "temporal_cross_attn_q", "temporal_self_attn_k", "temporal_self_attn_q"],
            input_block_indices=[1, 2, 4, 5, 7, 8, 10, 11], output_block_indices=list(range(1, 12))))
    gh, gw = modulate_grid(block_idx[0], base_height, base_width)
    out = []
    for sign in (1.0, -1.0):
        runs = []
        for mask_id in unique_labels:
            masks = feature_masks_from_labels(label_maps, mask_id, block_idx[0], base_height, base_width)
            params = {
                "feature_masks": masks, "modulate_block_idx": block_idx, "modulate_layer_type": list(modulate_layer_type),
                "modulate_attn_type": list(modulate_attn_type), "modulate_timestep": timesteps,
                "modulate_schedule": modulate_schedule, "modulate_lambda_start": sign * modulate_lambda_start,
                "modulate_lambda_end": sign * modulate_lambda_end, "num_frames": F, "modulate_uc": True,
                "is_injected_features": is_injected_features, **inj, "feature_folder": feature_folder, "exp_name": exp_name,
                "injected_features_group": {}, "modulate_layer_frames": {}, "modulate_block_frames": {},
                "modulate_timestep_frames": mtf, "modulate_lambda_layers": {}, "latent_mask_start": min(timesteps),
                "latent_mask_end": num_steps,
            }
            if features is not None:
                params["features"] = features
            z = sampler(denoiser, latent.clone(), cond=cond, uc=uc, is_modulate=True, modulate_params=params, t_start=t_start,
                        is_latent_blending=is_latent_blending, feature_height=gh, feature_width=gw,
                        is_smooth_latent=is_smooth_latent, model=first_stage)
            frames = frames_to_uint8(first_stage.decode_first_stage(z))
            runs.append(frames)
            if on_run is not None:
                on_run(sign, mask_id, frames)
        out.append(torch.stack(runs, 0))
    return out[0], out[1]


@torch.no_grad()
def modulated_segmentation(sampler, denoiser, first_stage, latent, cond, uc, label_maps, unique_labels, filter_s=0.7, **kw):
    """Steps 4 and 5 of the reference script (:404-508): the 2 x K modulated runs, then ``get_seg_map_main`` twice
    (filter_difference False, then True with ``filter_s``).  Returns dict(frames_pos, frames_neg, seg_raw, seg_raw_filtered)."""
    pos, neg = modulated_runs(sampler, denoiser, first_stage, latent, cond, uc, label_maps, unique_labels, **kw)
    plain = seg_maps_from_frames(pos, neg, unique_labels, filter_difference=False)
    filt = seg_maps_from_frames(pos, neg, unique_labels, label_maps=label_maps, filter_difference=True, filter_s=filter_s)
    return dict(frames_pos=pos, frames_neg=neg, seg_raw=plain["seg_raw"], seg_raw_filtered=filt["seg_raw"],
                seg_index=plain["seg_index"], seg_index_filtered=filt["seg_index"])
