"""Goldens for the ``match_gt_mask`` mode from the UNMODIFIED reference (scripts/sampling/feature_extraction.py:546-643
driven through ``feature_extraction_main``, exactly as sd_pipeline_vspw.py:365-385 does for consecutive windows of a
video: the first window clusters + maps to the ground-truth mask, the following windows propagate labels with the
4-NN classifier fitted on the previous window).

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_match_gt_goldens.py
"""
import os
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import match_gt as omg  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from synth import synthetic_clip_features  # noqa: E402

# (name, seed, F, h, w, C, K, gt mask height/width on disk, use_gt_mask)
CASES = [
    ("small", 21, 3, 12, 10, 64, 4, (48, 40), False),
    ("c1", 22, 4, 16, 16, 640, 5, (64, 64), False),
    ("c1_usegt", 23, 4, 16, 16, 640, 5, (64, 64), True),
    ("mid", 24, 6, 32, 32, 640, 12, (128, 128), False),
]
BLOCK = "output_block_8"


def gt_png(seg0, size, path):
    """A ground-truth mask at image resolution: the true object map of frame 0, nearest-upsampled, with class ids
    spread over 0..250 like a VSPW palette index."""
    hh, ww = size
    big = np.asarray(Image.fromarray(seg0.astype(np.uint8)).resize((ww, hh), Image.NEAREST))
    Image.fromarray((big * 25 + 3).astype(np.uint8)).save(path)


def run_case(fe, name, seed, F, h, w, C, K, size, use_gt):
    out = {}
    with tempfile.TemporaryDirectory() as root:
        gt_path = os.path.join(root, "gt.png")
        ref_mask = ref_fm = ref_unique = None
        for win in range(2):
            blocks, seg = synthetic_clip_features(seed + 100 * win, F, h, w, C, K, n_blocks=1)
            if win == 0:
                gt_png(seg[0], size, gt_path)
            exp = f"win{win}"
            fm_dir = os.path.join(root, exp, "feature_maps")
            os.makedirs(fm_dir)
            torch.save(torch.from_numpy(blocks[0]), os.path.join(fm_dir, f"{BLOCK}_spatial_self_attn_q_time_24.pt"))
            if win == 0:
                np.random.seed(seed)
            unique_labels, ref_mask, ref_fm = fe.feature_extraction_main(
                "match_gt_mask", K, 24, BLOCK, exp, exp, "spatial_self_attn_q", h, w, "24", frame_name_list=None,
                base_folder=root, ref_mask=ref_mask, ref_feature_map=ref_fm, ref_unique_labels=ref_unique,
                gt_mask_path=gt_path, num_frames=F, use_gt_mask=use_gt)
            if win == 0:
                ref_unique = unique_labels
            out[f"labels{win}"] = np.asarray(ref_mask).astype(np.int32)
            out[f"unique{win}"] = np.asarray(unique_labels)
            tree = sorted(os.path.relpath(os.path.join(d, f), os.path.join(root, exp)) for d, _, fs in
                          os.walk(os.path.join(root, exp, "match_gt_mask")) for f in fs)
            out[f"tree{win}"] = np.array(tree)
        gt_small = np.array(Image.open(gt_path).resize((w, h), Image.NEAREST)).reshape(-1)
    # the restatement must reproduce the reference on both windows
    np.random.seed(seed)
    rm = rf = None
    for win in range(2):
        blocks, _ = synthetic_clip_features(seed + 100 * win, F, h, w, C, K, n_blocks=1)
        _, rm, rf = omg.match_gt_mask(blocks[0], F, h, w, K, gt_mask=gt_small, ref_mask=rm, ref_feature_map=rf,
                                      use_gt_mask=use_gt)
        mism = int((rm != out[f"labels{win}"]).sum())
        print(name, "window", win, "oracle vs reference mismatches:", mism, "labels", np.unique(rm).tolist())
        assert mism == 0
    np.savez_compressed(os.path.join(HERE, f"matchgt_{name}.npz"), gt_small=gt_small.astype(np.int32),
                        meta=np.array([seed, F, h, w, C, K, int(use_gt)]), **out)


def main():
    fe = import_reference("scripts.sampling.feature_extraction")
    from oracle.ref_import import REFERENCE_ROOT
    os.chdir(REFERENCE_ROOT)   # convert_label_to_rgb reads scripts/util/color_map_soft.txt relative to the working directory
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        run_case(fe, *case)


if __name__ == "__main__":
    main()
