"""Pin oracle/process_output.py to the UNMODIFIED reference's seg-map post-process and write goldens.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_process_output_goldens.py
For every case of tests/synth.py:SEGMAP_CASES the synthetic decoded frames of the +lambda / -lambda modulated runs are
written as PNGs in the reference's ``modulated_output`` layout (svd_single_video_inference.py:171-190: cv2.imwrite of
the BGR frame), the K-means label maps as the per-label 0/255 PNG tree (feature_extraction.py:79-85), and the
reference's own ``get_seg_map_main`` (scripts/sampling/process_output.py:74-167) is run twice, exactly as
svd_single_video_inference.py:503-508 does (filter_difference False / filter_s 1.0, then True / 0.7).  Stored: the
``segmentation_map_raw*`` PNG contents and the stored difference JPEGs decoded again.  The in-memory oracle must
reproduce all of it bit for bit.
"""
import os
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import process_output as opo  # noqa: E402
from oracle.ref_import import REFERENCE_ROOT, import_reference  # noqa: E402
from synth import SEGMAP_CASES, synthetic_modulated_frames  # noqa: E402


def run_case(po, name, seed, K, F, H, W, fh, fw):
    import cv2
    pos, neg, labels = synthetic_modulated_frames(seed, K, F, H, W, fh, fw)
    unique = np.arange(K)
    lam, exp, t = 4.0, "exp", "24"
    out = {}
    with tempfile.TemporaryDirectory() as root:
        for sign, frames in ((lam, pos), (-lam, neg)):
            for k in range(K):
                folder = os.path.join(root, exp, "modulated_output", f"{0:06d}_l_{sign}_mask_{k}")
                os.makedirs(folder)
                for f in range(F):
                    cv2.imwrite(os.path.join(folder, f"{f}.png"), cv2.cvtColor(frames[k, f], cv2.COLOR_RGB2BGR))
        mask_folder = os.path.join(root, exp, "kmeans_masks", f"blocks_masks_{K}")
        for f in range(F):
            folder = os.path.join(mask_folder, f"kmeans_time_{t}_frame_{f}")
            os.makedirs(folder)
            for k in range(K):
                Image.fromarray(np.where(labels[f] == k, 255, 0).astype(np.uint8)).save(os.path.join(folder, f"mask_{k}.png"))
        cwd = os.getcwd()
        os.chdir(REFERENCE_ROOT)   # the reference opens scripts/util/color_map_soft.txt relative to the working directory
        try:
            for filt, fs in ((False, 1.0), (True, 0.7)):
                po.get_seg_map_main(exp, 0, lam, K, F, filter_difference=filt, filter_s=fs, resize_height=fh, resize_width=fw,
                                    unique_labels=unique, base_folder=root, frame_name_list=None, mask_folder=mask_folder,
                                    feature_timestep=t)
                sub = f"segmentation_map_raw_f_{fs}" if filt else "segmentation_map_raw"
                raw = np.stack([np.array(Image.open(os.path.join(root, exp, sub, f"{0:06d}_l_{lam}", f"{f}.png"))) for f in range(F)])
                out["raw_filtered" if filt else "raw"] = raw
        finally:
            os.chdir(cwd)
        back = np.stack([np.stack([np.array(Image.open(os.path.join(root, exp, "difference_map", "original_map",
                                                                    f"{0:06d}_l_{lam}_mask_{k}", f"{f}.jpg")))
                                   for f in range(F)]) for k in range(K)])
        tree = sorted(os.path.relpath(os.path.join(d, f), root) for d, _, fs_ in os.walk(os.path.join(root, exp)) for f in fs_
                      if "modulated_output" not in d and "kmeans_masks" not in d)
    for filt, key in ((False, "raw"), (True, "raw_filtered")):
        got, _, stored = opo.seg_maps(pos, neg, unique, labels, filter_difference=filt, filter_s=0.7)
        assert np.array_equal(got, out[key]), (name, key, int((got != out[key]).sum()))
    assert np.array_equal(np.stack([[opo.jpeg_roundtrip(stored[k, f]) for f in range(F)] for k in range(K)]), back)
    hist = np.bincount(out["raw"].reshape(-1), minlength=K).tolist()
    agree = float((out["raw"] == out["raw_filtered"]).mean())
    print(name, "raw hist", hist, "| filtered == unfiltered on", f"{agree:.3f}", "of the pixels | files", len(tree))
    np.savez_compressed(os.path.join(HERE, f"segmap_{name}.npz"), raw=out["raw"], raw_filtered=out["raw_filtered"],
                        back=back if name == "small" else back[:, :1], tree=np.array(tree),
                        meta=np.array([seed, K, F, H, W, fh, fw]))


def main():
    po = import_reference("scripts.sampling.process_output")
    only = sys.argv[1:]
    for case in SEGMAP_CASES:
        if only and case[0] not in only:
            continue
        run_case(po, *case)


if __name__ == "__main__":
    main()
