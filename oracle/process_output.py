"""Oracle for the seg-map post-process -- TEST INFRASTRUCTURE ONLY.

Restates ``scripts/sampling/process_output.py`` of the reference in memory (no files), calling the same third-party
routines at the same places the reference does (OpenCV ``GaussianBlur``, Pillow mode conversion, JPEG save / load and
LANCZOS resize; opencv-python and Pillow are unpinned in requirements/pt2.txt, this image has 4.13.0 / 12.2.0):

  compute_difference          :8-28    uint8 wrap-around difference of the +lambda / -lambda frames, squared IN uint8,
                                       summed over colour (uint64), sqrt (float64), 5x5 Gaussian blur sigma 3,
                                       ``Image.fromarray(float64)`` (mode F, float32) ``.convert("L")`` (clip + truncate),
                                       saved as JPEG quality 75 -- and READ BACK from the JPEG by the next stage
  filter_difference_map       :30-38   d * m + filter_s * d * (1 - m), m = LANCZOS-resized 0/255 mask / 255
  get_seg_map_main            :74-167  per (mask, frame): JPEG -> / (max + 1e-5) -> optional filter; arg-max over masks
                                       (first maximum), mapped through unique_labels -> uint8

``oracle/process_output_emul.py`` restates the library routines themselves (what the CUDA kernels implement) and is
checked against this file; ``tests/golden/make_process_output_goldens.py`` pins this file to the unmodified reference.
"""
import io

import numpy as np
from PIL import Image


def difference_map(frame_pos, frame_neg):
    """compute_difference (:8-28) for one frame pair, uint8 [H, W, 3] each.  Returns (blurred float64 [H, W], the mode-L
    image that is saved as ``original_map/.../{i}.jpg``, the mode-L image saved as ``vis_map/.../{i}.jpg``)."""
    import cv2
    a = np.asarray(frame_pos, dtype=np.uint8)
    b = np.asarray(frame_neg, dtype=np.uint8)
    d = np.sqrt(np.sum((a - b) ** 2, axis=2))          # uint8 arithmetic wraps, np.sum promotes to uint64
    d = cv2.GaussianBlur(d, (5, 5), 3)
    img = np.array(Image.fromarray(d).convert("L"))
    with np.errstate(divide="ignore", invalid="ignore"):
        vis = np.array(Image.fromarray(d / d.max() * 255).convert("L"))
    return d, img, vis


def jpeg_roundtrip(img_l):
    """``Image.save(path.jpg)`` followed by ``np.array(Image.open(path.jpg))`` (:19 -> :122) without the file."""
    buf = io.BytesIO()
    Image.fromarray(np.asarray(img_l, dtype=np.uint8)).save(buf, format="JPEG")
    buf.seek(0)
    return np.array(Image.open(buf))


def resized_mask(label_map, label, height, width):
    """The per-label 0/255 PNG of the K-means stage (feature_extraction.py:79-85) resized as filter_difference_map does
    (:34): LANCZOS to the frame size, divided by 255."""
    m = Image.fromarray(np.where(np.asarray(label_map) == label, 255, 0).astype(np.uint8))
    return np.array(m.resize((width, height), Image.LANCZOS)) / 255.0


def seg_maps(frames_pos, frames_neg, unique_labels, label_maps=None, filter_difference=False, filter_s=0.7):
    """get_seg_map_main (:74-167) in memory.

    frames_pos / frames_neg: uint8 [K, F, H, W, 3], the decoded frames of the +lambda / -lambda runs per mask (entry i
    belongs to unique_labels[i]); label_maps int [F, h, w] (needed with filter_difference).
    Returns (seg_raw uint8 [F, H, W] -- the ``segmentation_map_raw`` PNG content --, seg index maps int64 [F, H, W],
    the stored difference images uint8 [K, F, H, W])."""
    unique_labels = np.asarray(unique_labels)
    K, F, H, W, _ = frames_pos.shape
    stored = np.zeros((K, F, H, W), dtype=np.uint8)
    maps = np.zeros((K, F, H, W), dtype=np.float64)
    for i in range(K):
        for f in range(F):
            _, img, _ = difference_map(frames_pos[i, f], frames_neg[i, f])
            stored[i, f] = img
            back = jpeg_roundtrip(img)
            dm = back / (np.max(back) + 1e-5)
            if filter_difference:
                m = resized_mask(label_maps[f], unique_labels[i], H, W)
                dm = dm * m + filter_s * dm * (1 - m)
            maps[i, f] = dm
    seg = np.argmax(maps, axis=0)
    return unique_labels[seg].astype(np.uint8), seg, stored
