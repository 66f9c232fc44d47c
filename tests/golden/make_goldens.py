"""Generate golden vectors by running the UNMODIFIED reference code.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_goldens.py
Writes tests/golden/cluster_<case>.npz.  For every case in tests/synth.py it
drives the reference's own ``feature_extraction_main`` (scripts/sampling/
feature_extraction.py:670-795) in ``kmeans_masks`` mode and then in
``correct_low_res_mask`` mode, exactly as svd_single_video_inference.py:372-403
does, with the stashed features written as ``.pt`` files in the reference's
feature_folder layout.  ``np.random.seed(seed)`` is called where
``seed_everything`` would (svd_single_video_inference.py:590-594).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle.ref_import import import_reference  # noqa: E402
from synth import CLUSTER_CASES, synthetic_clip_features  # noqa: E402

BLOCKS = ["output_block_8", "output_block_7", "output_block_6"]


def run_case(fe, name, seed, F, h, w, C, K, kind):
    blocks, seg = synthetic_clip_features(seed, F, h, w, C, K, kind=kind)
    with tempfile.TemporaryDirectory() as root:
        exp = "exp"
        fm_dir = os.path.join(root, exp, "feature_maps")
        os.makedirs(fm_dir)
        t = 24
        for bname, arr in zip(BLOCKS, blocks):
            torch.save(torch.from_numpy(arr), os.path.join(fm_dir, f"{bname}_spatial_self_attn_q_time_{t}.pt"))
        np.random.seed(seed)
        unique_labels, _, _ = fe.feature_extraction_main(
            "kmeans_masks", K, t, ",".join(BLOCKS), exp, exp, "spatial_self_attn_q",
            h, w, str(t), frame_name_list=None, base_folder=root, num_frames=F)
        block_str = "_".join(BLOCKS)
        mask_folder = os.path.join(root, exp, "kmeans_masks", f"{block_str}_spatial_self_attn_q_masks_{K}")
        labels = np.stack([fe.generate_aggregate_mask(mask_folder, t, K, i, h, w) for i in range(F)])
        png_tree = sorted(os.path.relpath(os.path.join(d, f), root) for d, _, fs in os.walk(os.path.join(root, exp, "kmeans_masks")) for f in fs)
        # refine: features of output_block_7 only (svd_single_video_inference.py:393)
        all_h, all_w = fe.dense_tracking(torch.from_numpy(blocks[1]), feature_height=h, feature_width=w,
                                         num_frames=F, device="cpu")
        fe.dense_tracking.__defaults__ = tuple("cpu" if d == "cuda" else d for d in fe.dense_tracking.__defaults__)
        _, ref_mask, _ = fe.feature_extraction_main(
            "correct_low_res_mask", K, t, "output_block_7", exp, exp, "spatial_self_attn_q",
            h, w, str(t), frame_name_list=None, base_folder=root, num_frames=F,
            mask_folder=mask_folder, ref_unique_labels=unique_labels)
        corrected_tree = sorted(os.path.relpath(os.path.join(d, f), root) for d, _, fs in os.walk(mask_folder + "_corrected") for f in fs)
    out = os.path.join(HERE, f"cluster_{name}.npz")
    np.savez_compressed(out, labels=labels.astype(np.int32), ref_mask=np.asarray(ref_mask).astype(np.int64),
                        all_h=np.asarray(all_h).astype(np.int16), all_w=np.asarray(all_w).astype(np.int16),
                        unique_labels=np.asarray(unique_labels), png_tree=np.array(png_tree),
                        corrected_tree=np.array(corrected_tree), gt_seg=seg.astype(np.int8),
                        meta=np.array([seed, F, h, w, C, K]))
    print(name, "labels hist", np.bincount(labels.reshape(-1), minlength=K).tolist(),
          "changed by refine", int((labels.reshape(-1) != ref_mask).sum()), "->", out)


def main():
    fe = import_reference("scripts.sampling.feature_extraction")
    only = sys.argv[1:]
    for case in CLUSTER_CASES:
        if only and case[0] not in only:
            continue
        run_case(fe, *case)


if __name__ == "__main__":
    main()
