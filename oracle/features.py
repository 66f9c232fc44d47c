"""Oracle for R1: multi-block aggregation + per-token max-abs normalisation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows scripts/sampling/feature_extraction.py:
  * :739-745  ``torch.mean(torch.stack([b8, b7, b6]), dim=0)`` -- blocks in the
    order the caller lists them (svd_single_video_inference.py:362 passes
    ``output_block_8,output_block_7,output_block_6``).  torch's CPU mean is
    sum-then-divide: ((b8 + b7) + b6) / 3 in fp32.
  * :35-46    ``x / max(|x|, axis=-1)`` (no epsilon, guarded by C > 1), then the
    cond half ``x[num_frames:]`` flattened to [(F*hw), C].
"""
import numpy as np


def aggregate_blocks(blocks):
    """blocks: list of [2F, hw, C] float32 arrays, in caller order."""
    acc = np.array(blocks[0], dtype=np.float32, copy=True)
    for b in blocks[1:]:
        acc = acc + np.asarray(b, dtype=np.float32)
    if len(blocks) > 1:
        acc = acc / np.float32(len(blocks))
    return acc


def normalize_max_abs(feature_maps):
    """feature_extraction.py:38-39."""
    if feature_maps.shape[-1] > 1:
        feature_maps = feature_maps / np.max(np.abs(feature_maps), axis=-1, keepdims=True)
    return feature_maps


def cond_half_matrix(feature_maps, num_frames):
    """feature_extraction.py:45-46: drop the unconditional half, flatten tokens."""
    split = feature_maps[num_frames:]
    return split.reshape(-1, split.shape[-1])


def aggregate_normalize(blocks, num_frames):
    """R1 end to end: list of [2F, hw, C] -> [(F*hw), C] float32."""
    x = aggregate_blocks(blocks)
    x = normalize_max_abs(x)
    return np.ascontiguousarray(cond_half_matrix(x, num_frames))
