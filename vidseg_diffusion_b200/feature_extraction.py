"""Host mirror of the reference's ``scripts/sampling/feature_extraction.py`` for the hot path.

Same function names, argument meaning, return values and on-disk ``feature_folder`` layout as the
reference (``feature_extraction_main`` :670-795, ``save_inidividual_masks_kmeans`` :30-113,
``correct_low_res_mask`` :367-461, ``generate_aggregate_mask`` :500-521, ``generate_binary_mask``
:523-535, ``load_experiments_features`` :646-668); the arithmetic runs in libvidseg_b200 on the
GPU (aggregate/normalise, K-means, nearest-neighbour tracking + voting).  Differences, all opt-in:

  * ``features=`` lets the caller hand over the stashed Q tensors that are already resident in HBM
    instead of round-tripping them through ``torch.save`` / ``torch.load`` (the reference's disk
    boundary, svd_single_video_inference.py:137-149); without it the ``.pt`` files are read exactly
    like the reference does.
  * ``write_pngs=False`` skips the per-(frame,label) PNG tree when the caller only wants the label
    maps; the default writes the same mode-L 0/255 PNGs in the same folders.

``match_gt_mask`` (:546-643) lives in match_gt.py.
"""
import os

import numpy as np
import torch
from PIL import Image

from . import _lib
from .features import aggregate_normalize
from .kmeans import KMeans
from .refine import refine_masks

# Hand-over of label maps between ``kmeans_masks`` / ``match_gt_mask`` and ``correct_low_res_mask`` for callers that
# asked for ``write_pngs=False``: with no PNG tree on disk the maps wait here, keyed by mask folder and stamped with the
# window they belong to; the consumer checks the stamp and removes the entry.  Whenever the PNG tree IS written the
# refinement re-reads it, exactly like the reference (:380-389).
_LABEL_CACHE = {}


def _stamp(num_frames, frame_name_list, feature_height, feature_width):
    names = None if frame_name_list is None else tuple(str(n) for n in list(frame_name_list)[:num_frames])
    return (int(num_frames), names, int(feature_height), int(feature_width))


def register_label_maps(masks_folder, labels, num_frames, frame_name_list, feature_height, feature_width):
    _LABEL_CACHE[os.path.normpath(masks_folder)] = (labels, _stamp(num_frames, frame_name_list, feature_height, feature_width))


def invalidate_label_maps(masks_folder):
    _LABEL_CACHE.pop(os.path.normpath(masks_folder), None)


def _device():
    if not torch.cuda.is_available():
        raise _lib.VidsegError("feature_extraction needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device_f32(t):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=_device(), dtype=torch.float32).contiguous()


def load_experiments_features(feature_maps_paths, blocks, feature_type, t, frame_id=None, features=None):
    """reference :646-668.  ``features`` (dict keyed by (block, feature_type, t)) short-circuits the disk."""
    out = []
    for path in feature_maps_paths:
        if features is not None and (blocks, feature_type, t) in features:
            fm = features[(blocks, feature_type, t)]
        elif frame_id is None:
            fm = torch.load(os.path.join(path, f"{blocks}_{feature_type}_time_{t}.pt"), map_location="cpu")
            if "attn" not in feature_type and fm.dim() == 4:
                fm = fm.permute(0, 2, 3, 1).reshape(fm.shape[0], -1, fm.shape[1])  # 'b c h w -> b (h w) c'
        else:
            fm = torch.load(os.path.join(path, f"{blocks}_{feature_type}_time_{t}_frame_{frame_id}.pt"), map_location="cpu")
        out.append(fm)
    return out


def _masks_folder(output_folder, num_clusters):
    """reference :58-60 (string replace of the last path component)."""
    name = output_folder.split("/")[-1]
    return output_folder.replace(name, name + f"_masks_{num_clusters}")


def generate_binary_mask(aggregate_mask, output_folder, labels=None):
    """reference :523-535."""
    aggregate_mask = np.asarray(aggregate_mask)
    if labels is None:
        labels = np.unique(aggregate_mask)
    for i in labels:
        mask = np.where(aggregate_mask == i, 255, 0).astype(np.uint8)
        Image.fromarray(mask).convert("L").save(os.path.join(output_folder, f"mask_{int(i)}.png"))


def generate_aggregate_mask(mask_folder, timestep, num_masks, frame_id, resize_height, resize_width, labels=None):
    """reference :500-521 (argmax over the per-label PNGs)."""
    iterator = range(num_masks) if labels is None else labels
    all_masks = []
    for i in iterator:
        path = os.path.join(mask_folder, f"kmeans_time_{timestep}_frame_{frame_id}", f"mask_{i}.png")
        all_masks.append(np.array(Image.open(path).resize((resize_width, resize_height))))
    seg_map = np.argmax(all_masks, axis=0)
    if labels is not None:
        seg_map = np.asarray(labels)[seg_map]
    return seg_map


def save_inidividual_masks_kmeans(feature_blocks, selected_timestep, output_folder, num_frames=14, num_clusters=10,
                                  feature_height=16, feature_width=16, attn_type="spatial", frame_name_list=None,
                                  write_pngs=True, return_labels=False):
    """reference :30-113 (spatial / features attention types).

    ``feature_blocks``: one tensor [2F, hw, C] or the list of per-block tensors to be averaged
    (the reference averages in ``feature_extraction_main`` and passes the mean; here the mean, the
    max-abs normalisation and the cond-half slice are one fused kernel)."""
    if attn_type not in ("spatial", "features"):
        raise _lib.VidsegError(f"attn_type {attn_type!r}: only the spatial path of the reference is on the hot path")
    if isinstance(feature_blocks, torch.Tensor):
        feature_blocks = [feature_blocks]
    blocks = [_to_device_f32(b) for b in feature_blocks]
    x = aggregate_normalize(blocks, num_frames)
    kmeans = KMeans(n_clusters=num_clusters, n_init=10)
    labels = kmeans.fit_predict(x).reshape(num_frames, feature_height, feature_width)
    out_folder = _masks_folder(output_folder, num_clusters)
    if write_pngs:
        invalidate_label_maps(out_folder)   # the PNG tree is the hand-over, as in the reference
    else:
        register_label_maps(out_folder, labels, num_frames, frame_name_list, feature_height, feature_width)
    if write_pngs:
        labels_np = labels.cpu().numpy()
        for i in range(num_frames):
            frame_name = frame_name_list[i] if frame_name_list is not None else i
            folder = os.path.join(out_folder, f"kmeans_time_{selected_timestep}_frame_{frame_name}")
            os.makedirs(folder, exist_ok=True)
            for label in range(num_clusters):
                mask = np.where(labels_np[i] == label, 255, 0).astype(np.uint8)
                Image.fromarray(mask).save(os.path.join(folder, f"mask_{label}.png"))
    unique_labels = np.arange(num_clusters)
    if return_labels:
        return unique_labels, labels, kmeans
    return unique_labels


def correct_low_res_mask(feature_maps, mask_folder, output_folder=None, num_clusters=10, overlay_images_folder=None,
                         feature_height=16, feature_width=16, attn_type="spatial", num_frames=14, timestep=24,
                         top_k=1, anchor_label_method="common", frame_name_list=None, ref_unique_labels=None,
                         spatial_filter=True, label_maps=None, write_pngs=True):
    """reference :367-461 with the defaults actually in force (top_k=1, use_aux=True, no backtracing,
    "common" anchor label, spatial filter on).  Returns (ref_unique_labels, ref_mask, None)."""
    if attn_type != "spatial" or top_k != 1 or anchor_label_method != "common" or not spatial_filter:
        raise _lib.VidsegError("correct_low_res_mask: only the reference's default configuration is built")
    fm = _to_device_f32(feature_maps)
    if label_maps is None:
        cached = _LABEL_CACHE.pop(os.path.normpath(mask_folder), None)
        if cached is not None and cached[1] == _stamp(num_frames, frame_name_list, feature_height, feature_width):
            label_maps = cached[0]
        else:
            # rebuild from the PNG tree exactly like the reference (:380-389; timestep 24 is hard-coded there)
            maps = []
            for i in range(num_frames):
                frame_name = frame_name_list[i] if frame_name_list is not None else i
                maps.append(generate_aggregate_mask(mask_folder, 24, num_clusters, frame_name, feature_height,
                                                    feature_width, labels=ref_unique_labels))
            label_maps = torch.as_tensor(np.array(maps))
    label_maps = label_maps.to(device=fm.device, dtype=torch.int32).contiguous()
    refined, _, _ = refine_masks(fm, label_maps, num_frames, feature_height, feature_width)
    new_seg = refined.cpu().numpy().astype(np.int64)
    if write_pngs:
        name = mask_folder.split("/")[-1]
        out_root = mask_folder.replace(name, name + "_corrected")
        os.makedirs(out_root, exist_ok=True)
        for i in range(num_frames):
            frame_name = frame_name_list[i] if frame_name_list is not None else i
            folder = os.path.join(out_root, f"kmeans_time_{timestep}_frame_{frame_name}")
            os.makedirs(folder, exist_ok=True)
            generate_binary_mask(new_seg[i], folder, ref_unique_labels)
    return ref_unique_labels, new_seg.reshape(-1), None


def feature_extraction_main(mode, num_clusters, t_start, block_name, experiment_name, fit_experiments, feature_types,
                            feature_height, feature_width, selected_timestep, frame_name_list=None, base_folder=None,
                            ref_mask=None, ref_feature_map=None, ref_unique_labels=None, gt_mask_path=None,
                            num_frames=None, mask_folder=None, use_gt_mask=False, features=None, write_pngs=True):
    """reference :670-795.  Same positional signature and return triple."""
    exp_path_root = "features_outputs" if base_folder is None else base_folder
    selected_timestep = [int(ts) for ts in selected_timestep.split(",") if ts]
    fit_experiments = [item for item in fit_experiments.split(",") if item]
    if num_frames is None:
        num_frames = 14
    block_name = block_name.split(",")
    if len(block_name) == 1:
        block_name = block_name[0]
    feature_maps_paths = [os.path.join(exp_path_root, e, "feature_maps") for e in fit_experiments]
    feature_types = [item for item in feature_types.split(",") if item]
    if mode not in ("kmeans_masks", "correct_low_res_mask", "match_gt_mask"):
        raise ValueError(f"mode {mode} not supported")
    out_root = os.path.join(exp_path_root, experiment_name, mode)
    os.makedirs(out_root, exist_ok=True)
    block_str = "_".join(block_name) if isinstance(block_name, list) else block_name
    unique_labels = None
    for t in selected_timestep:
        for feature_type in feature_types:
            if "temporal" in feature_type:
                attn_type = "temporal"
            elif "features" in feature_type:
                attn_type = "features"
            else:
                attn_type = "spatial"
            names = block_name if isinstance(block_name, list) else [block_name]
            blocks = []
            for sub in names:
                fms = load_experiments_features(feature_maps_paths, sub, feature_type, t, features=features)
                blocks.append(torch.cat([_to_device_f32(f) for f in fms], dim=0) if len(fms) > 1 else _to_device_f32(fms[0]))
            out_path = os.path.join(out_root, f"{block_str}_{feature_type}")
            if mode == "kmeans_masks":
                unique_labels = save_inidividual_masks_kmeans(
                    blocks, t, out_path, num_frames=num_frames, num_clusters=num_clusters,
                    feature_height=feature_height, feature_width=feature_width, attn_type=attn_type,
                    frame_name_list=frame_name_list, write_pngs=write_pngs)
            elif mode == "match_gt_mask":
                from .match_gt import match_gt_mask
                unique_labels, ref_mask, ref_feature_map = match_gt_mask(
                    blocks, gt_mask_path=gt_mask_path, feature_height=feature_height, feature_width=feature_width,
                    output_folder=out_path, num_masks=num_clusters, selected_timestep=t, frame_name_list=frame_name_list,
                    ref_mask=ref_mask, ref_feature_map=ref_feature_map, ref_unique_labels=ref_unique_labels,
                    use_gt_mask=use_gt_mask, num_frames=num_frames, write_pngs=write_pngs)
            else:  # correct_low_res_mask
                if len(blocks) != 1:
                    # the reference would average the blocks here too (:739-745)
                    fm = torch.stack(blocks).sum(0) / len(blocks)
                else:
                    fm = blocks[0]
                unique_labels, ref_mask, ref_feature_map = correct_low_res_mask(
                    fm, mask_folder=mask_folder, feature_height=feature_height, feature_width=feature_width,
                    attn_type=attn_type, timestep=t, num_frames=num_frames, num_clusters=num_clusters,
                    frame_name_list=frame_name_list, ref_unique_labels=ref_unique_labels, write_pngs=write_pngs)
    return unique_labels, ref_mask, ref_feature_map
