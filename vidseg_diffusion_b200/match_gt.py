"""``match_gt_mask`` mode on B200 (reference: scripts/sampling/feature_extraction.py:546-643).

The mode both dataset pipelines use (sd_pipeline_vspw.py:365-385): the first window of a video is clustered with
K-means, every cluster of frame 0 takes the most frequent ground-truth label of its cells, and a 4-nearest-neighbour
classifier fitted on frame 0 labels every token of the window; the following windows are labelled by the classifier
fitted on the previous window's tokens and labels.  Device side: csrc/kmeans.cu (majority map, tcgen05 filter +
float64 k-NN).  Host side keeps the reference's signature and PNG tree.
"""
import os
import shutil

import numpy as np
import torch
from PIL import Image

from . import _lib
from .features import aggregate_normalize
from .kmeans import KMeans


def majority_map(fake_labels, gt_labels, num_fake):
    """Every label of ``fake_labels`` -> the most frequent value of ``gt_labels`` among its cells (reference :589-594;
    smallest value on ties).  int32 CUDA tensors [n]; gt values in [0, 1024)."""
    fake = _lib.require_cuda_tensor(fake_labels, torch.int32, "fake_labels")
    gt = _lib.require_cuda_tensor(gt_labels, torch.int32, "gt_labels")
    if fake.shape != gt.shape:
        raise _lib.VidsegError("majority_map: shape mismatch")
    ref = torch.zeros_like(fake)
    err = torch.zeros(1, dtype=torch.int32, device=fake.device)
    lib = _lib.load()
    with torch.cuda.device(fake.device):
        _lib.check(lib.vidseg_majority_map(fake.data_ptr(), gt.data_ptr(), fake.numel(), int(num_fake), ref.data_ptr(),
                                           err.data_ptr(), _lib.stream_ptr()), "majority_map")
    if int(err.item()):
        raise _lib.VidsegError("majority_map: ground-truth labels must lie in [0, 1024)")
    return ref


def knn_predict(ref_features, ref_labels, queries, n_neighbors=4):
    """``KNeighborsClassifier(n_neighbors).fit(ref_features, ref_labels).predict(queries)`` (reference :606-612).
    fp32 CUDA tensors [n_ref, D] / [n_query, D], int32 labels; returns int32 [n_query]."""
    ref = _lib.require_cuda_tensor(ref_features, torch.float32, "ref_features")
    q = _lib.require_cuda_tensor(queries, torch.float32, "queries")
    lab = _lib.require_cuda_tensor(ref_labels, torch.int32, "ref_labels")
    if ref.dim() != 2 or q.dim() != 2 or ref.shape[1] != q.shape[1] or lab.numel() != ref.shape[0]:
        raise _lib.VidsegError(f"knn_predict: bad shapes ref{tuple(ref.shape)} labels{tuple(lab.shape)} q{tuple(q.shape)}")
    n_ref, d = ref.shape
    if n_ref < n_neighbors:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {n_neighbors}, n_samples_fit = {n_ref}")
    out = torch.empty(q.shape[0], dtype=torch.int32, device=q.device)
    err = torch.zeros(1, dtype=torch.int32, device=q.device)
    lib = _lib.load()
    nbytes = lib.vidseg_knn_workspace_bytes(n_ref, max(q.shape[0], 1), d)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.vidseg_knn_predict(ref.data_ptr(), lab.data_ptr(), n_ref, q.data_ptr(), q.shape[0], d, int(n_neighbors),
                                          out.data_ptr(), err.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr()), "knn_predict")
    if int(err.item()):
        raise _lib.VidsegError("knn_predict: more than 64 reference rows tie with the k-th neighbour (duplicated features)")
    return out


def load_gt_mask(gt_mask_path, feature_height, feature_width, device):
    """reference :580-583: PNG resized NEAREST to the feature grid, flattened."""
    mask = np.array(Image.open(gt_mask_path).resize((feature_width, feature_height), Image.NEAREST))
    if mask.ndim != 2:
        raise _lib.VidsegError(f"ground-truth mask {gt_mask_path}: expected a single-channel label image")
    return torch.from_numpy(mask.astype(np.int32).reshape(-1)).to(device)


def match_gt_mask(feature_blocks, gt_mask_path, feature_height, feature_width, output_folder, num_masks,
                  selected_timestep=24, frame_name_list=None, ref_mask=None, ref_feature_map=None, ref_unique_labels=None,
                  use_gt_mask=False, num_frames=None, write_pngs=True):
    """reference :546-643, same argument meaning and return triple (unique_labels, ref_mask, ref_feature_map); the two
    references are CUDA tensors (int32 [F*hw], fp32 [F*hw, C]) that the caller passes back for the next window.

    ``feature_blocks``: one tensor [2F, hw, C] or the list of per-block tensors to average.

    Departures from the reference, all on the host side: ``kmeans_cluster_labels.png`` (a matplotlib debug plot, :596-600)
    is not written; ground-truth labels must lie in [0, 1024) (the majority kernel's histogram; PNG masks are uint8);
    ``majority_map`` / ``knn_predict`` read one error flag back from the device each (a host sync the reference does not
    have); with ``write_pngs=False`` the label maps reach a following ``correct_low_res_mask`` through
    ``feature_extraction.register_label_maps`` instead of the PNG tree."""
    if isinstance(feature_blocks, torch.Tensor):
        feature_blocks = [feature_blocks]
    blocks = [_lib.require_cuda_tensor(b.float().contiguous() if b.is_cuda else b, torch.float32, "feature_maps")
              for b in feature_blocks]
    if num_frames is None:
        num_frames = blocks[0].shape[0] // 2
    h, w = feature_height, feature_width
    tokens = aggregate_normalize(blocks, num_frames)           # [(F*hw), C]: cond half, max-abs normalised
    name = output_folder.split("/")[-1]
    output_folder = output_folder.replace(name, name + f"_masks_{num_masks}")
    if ref_mask is None:
        km = KMeans(n_clusters=num_masks, n_init=10).fit(tokens)
        fake = km.predict(tokens[: h * w].contiguous())
        if write_pngs:
            os.makedirs(output_folder, exist_ok=True)
        gt = fake if gt_mask_path is None else load_gt_mask(gt_mask_path, h, w, tokens.device)
        if use_gt_mask:
            assert gt_mask_path is not None
            ref_mask = gt
        else:
            ref_mask = majority_map(fake, gt, num_masks)
        ref_feature_map = tokens[: h * w].contiguous()
    else:
        ref_mask = _lib.require_cuda_tensor(torch.as_tensor(ref_mask).to(tokens.device, torch.int32).contiguous(), torch.int32, "ref_mask")
        ref_feature_map = torch.as_tensor(ref_feature_map).to(tokens.device, torch.float32).contiguous()
    unique_labels = np.unique(ref_mask.cpu().numpy())
    if ref_unique_labels is None:
        ref_unique_labels = unique_labels
    labels = knn_predict(ref_feature_map, ref_mask, tokens, 4)
    from .feature_extraction import generate_binary_mask, invalidate_label_maps, register_label_maps
    if write_pngs:
        invalidate_label_maps(output_folder)
    else:   # no PNG tree: a following correct_low_res_mask takes the maps from the hand-over cache
        register_label_maps(output_folder, labels.reshape(num_frames, h, w), num_frames, frame_name_list, h, w)
    if write_pngs:
        labels_np = labels.cpu().numpy().reshape(num_frames, h, w)
        for frame_id in range(num_frames):
            frame_name = frame_name_list[frame_id] if frame_name_list is not None else frame_id
            folder = os.path.join(output_folder, f"kmeans_time_{selected_timestep}_frame_{frame_name}")
            if os.path.exists(folder):
                shutil.rmtree(folder)
            os.makedirs(folder)
            generate_binary_mask(labels_np[frame_id], folder, ref_unique_labels)
    return unique_labels, labels, tokens
