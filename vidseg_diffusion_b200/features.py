"""R1 host mirror: multi-block aggregation + per-token max-abs normalisation.

Reference: scripts/sampling/feature_extraction.py:739-745 (mean of the stacked blocks in caller
order), :38-39 (max-abs normalise), :45-46 (conditional half, flattened).
"""
import ctypes

import torch

from . import _lib


def aggregate_normalize(blocks, num_frames, cond_only=False):
    """blocks: list of 1..4 CUDA float32 tensors [2F, hw, C] in caller order ([F, hw, C] with ``cond_only``: the
    conditional half alone, as the CFG-half split of the SVD UNet leaves it on one rank).
    Returns the K-means input matrix [(F*hw), C] float32 on the same device."""
    if isinstance(blocks, torch.Tensor):
        blocks = [blocks]
    if not 1 <= len(blocks) <= 4:
        raise _lib.VidsegError("aggregate_normalize: 1..4 blocks supported")
    b0 = _lib.require_cuda_tensor(blocks[0], torch.float32, "blocks[0]")
    if b0.dim() != 3 or b0.shape[0] != (1 if cond_only else 2) * num_frames:
        raise _lib.VidsegError(f"blocks must be [{'' if cond_only else '2'}F, hw, C] with F={num_frames}, got {tuple(b0.shape)}")
    for i, b in enumerate(blocks[1:], 1):
        _lib.require_cuda_tensor(b, torch.float32, f"blocks[{i}]")
        if b.shape != b0.shape:
            raise _lib.VidsegError("aggregate_normalize: all blocks must have the same shape")
    _, hw, c = b0.shape
    out = torch.empty((num_frames * hw, c), dtype=torch.float32, device=b0.device)
    ptrs = (ctypes.c_void_p * len(blocks))(*[b.data_ptr() for b in blocks])
    lib = _lib.load()
    with torch.cuda.device(b0.device):
        if cond_only:
            _lib.check(lib.vidseg_aggregate_normalize_rows(ctypes.cast(ptrs, ctypes.c_void_p), len(blocks), 0, num_frames * hw,
                                                           c, out.data_ptr(), _lib.stream_ptr()), "aggregate_normalize_rows")
        else:
            _lib.check(lib.vidseg_aggregate_normalize(ctypes.cast(ptrs, ctypes.c_void_p), len(blocks), num_frames, hw, c,
                                                      out.data_ptr(), _lib.stream_ptr()), "aggregate_normalize")
    return out
