"""R3 host mirror: correspondence-based mask refinement.

Reference: scripts/sampling/feature_extraction.py:176-323, :326-364, :367-461.
"""
import torch

from . import _lib


def refine_masks(feature_maps, label_maps, num_frames, feature_height, feature_width):
    """feature_maps: CUDA float32 [2F, hw, C] of ONE block (uncond rows first);
    label_maps: CUDA int32 [F, h, w] (or [F, hw]).
    Returns (refined int32 [F, h, w], trajectories int32 [F, hw] of cell indices, keep int32 [hw])."""
    fm = _lib.require_cuda_tensor(feature_maps, torch.float32, "feature_maps")
    hw = feature_height * feature_width
    if fm.dim() != 3 or fm.shape[0] != 2 * num_frames or fm.shape[1] != hw:
        raise _lib.VidsegError(f"feature_maps must be [2F, hw, C] = [{2 * num_frames}, {hw}, C], got {tuple(fm.shape)}")
    lab = _lib.require_cuda_tensor(label_maps.reshape(num_frames, hw), torch.int32, "label_maps")
    c = fm.shape[2]
    lib = _lib.load()
    nbytes = lib.vidseg_refine_workspace_bytes(num_frames, hw, c)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=fm.device)
    traj = torch.empty((num_frames, hw), dtype=torch.int32, device=fm.device)
    keep = torch.empty(hw, dtype=torch.int32, device=fm.device)
    out = torch.empty((num_frames, hw), dtype=torch.int32, device=fm.device)
    with torch.cuda.device(fm.device):
        _lib.check(lib.vidseg_refine_masks(fm.data_ptr(), lab.data_ptr(), num_frames, feature_height, feature_width, c,
                                           traj.data_ptr(), keep.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes,
                                           _lib.stream_ptr()), "refine_masks")
    return out.reshape(num_frames, feature_height, feature_width), traj, keep
