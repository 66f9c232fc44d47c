"""Host mirror of the reference's ``scripts/sampling/process_output.py`` (seg-map post-process).

``get_seg_map_main`` keeps the reference's signature (:74-77), folder layout (``difference_map/{original_map,vis_map}``,
``segmentation_map[_raw][_f_{s}]/{basecount:06d}_l_{lambda}``) and file formats; the arithmetic between the decoded
frames and the label maps runs in libvidseg_b200 (csrc/postprocess.cu), bit-exact with the OpenCV / Pillow / libjpeg
routines the reference calls, including the JPEG save / load round trip the reference makes between its two stages
(which therefore needs no file here).  ``seg_maps_from_frames`` is the same computation on tensors that are still in
HBM after the VAE decode of the modulated runs.
"""
import math
import os

import numpy as np
import torch
from PIL import Image

from . import _lib

_PRECISION_BITS = 32 - 8 - 2   # Pillow src/libImaging/Resample.c, 8 bits per channel
_COEFF_CACHE = {}


def _sinc(x):
    return 1.0 if x == 0.0 else math.sin(x * math.pi) / (x * math.pi)


def _lanczos(t):
    return _sinc(t) * _sinc(t / 3) if -3.0 <= t < 3.0 else 0.0


def _bicubic(t, a=-0.5):
    """Resample.c bicubic_filter (Keys, a = -0.5): Pillow's default filter of Image.resize for mode L."""
    t = -t if t < 0.0 else t
    if t < 1.0:
        return ((a + 2.0) * t - (a + 3.0)) * t * t + 1
    if t < 2.0:
        return (((t - 5) * t + 8) * t - 4) * a
    return 0.0


_FILTERS = {"lanczos": (_lanczos, 3.0), "bicubic": (_bicubic, 2.0)}


def lanczos_windows(in_size, out_size, filter="lanczos"):
    """Pillow's resampling coefficient windows for one axis (Resample.c precompute_coeffs + normalize_coeffs_8bpc): data
    independent, so they are computed once per (in, out, filter) on the host with the same libm calls Pillow makes.
    Returns (bounds int32 [out, 2] = (first input index, count), coeffs int32 [out, ksize], ksize)."""
    key = (in_size, out_size, filter)
    if key in _COEFF_CACHE:
        return _COEFF_CACHE[key]
    fn, fsupport = _FILTERS[filter]
    scale = in_size / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    one = float(1 << _PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = lo if lo > 0 else 0
        hi = int(center + support + 0.5)
        hi = hi if hi < in_size else in_size
        n = hi - lo
        w, total = [], 0.0
        for x in range(n):
            v = fn((x + lo - center + 0.5) * inv)
            w.append(v)
            total += v
        for x in range(n):
            v = w[x] / total if total != 0.0 else w[x]
            coeffs[xx, x] = int(-0.5 + v * one) if v < 0 else int(0.5 + v * one)
        bounds[xx] = (lo, n)
    _COEFF_CACHE[key] = (bounds, coeffs, ksize)
    return _COEFF_CACHE[key]


def _u8(t, name, device):
    if not isinstance(t, torch.Tensor):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if t.dtype != torch.uint8:
        raise _lib.VidsegError(f"{name}: expected uint8 frames, got {t.dtype}")
    t = t.to(device).contiguous()
    if not t.is_cuda:
        raise _lib.VidsegError(f"{name}: the post-process needs a CUDA device (no CPU fallback)")
    return t


def resized_masks(label_maps, unique_labels, height, width, filter="lanczos"):
    """uint8 [K, F, H, W]: Pillow ``Image.resize((W, H), LANCZOS)`` of every 0/255 mask ``label_maps[f] == unique_labels[k]``
    (what filter_difference_map :34 builds from the K-means PNG tree); ``filter="bicubic"`` is Pillow's default filter
    (load_feature_masks, svd_single_video_inference.py:93).  label_maps: CUDA int32 [F, h, w]."""
    lab = _lib.require_cuda_tensor(label_maps.contiguous(), torch.int32, "label_maps")
    F, h, w = lab.shape
    if h == height or w == width:
        raise _lib.VidsegError("resized_masks: Pillow skips an unchanged axis; the feature grid never equals the frame size")
    dev = lab.device
    ul = torch.as_tensor(np.asarray(unique_labels), dtype=torch.int32).to(dev)
    K = ul.numel()
    hb, hk, hks = lanczos_windows(w, width, filter)
    vb, vk, vks = lanczos_windows(h, height, filter)
    d = lambda a: torch.from_numpy(a).to(dev)
    hb, hk, vb, vk = d(hb), d(hk), d(vb), d(vk)
    tmp = torch.empty((K, F, h, width), dtype=torch.uint8, device=dev)
    out = torch.empty((K, F, height, width), dtype=torch.uint8, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.vidseg_lanczos_masks(lab.data_ptr(), ul.data_ptr(), K, F, h, w, height, width, hb.data_ptr(),
                                            hk.data_ptr(), hks, vb.data_ptr(), vk.data_ptr(), vks, tmp.data_ptr(),
                                            out.data_ptr(), _lib.stream_ptr()), "lanczos_masks")
    return out


def difference_images(frames_pos, frames_neg, want_vis=False):
    """compute_difference (:8-28) for a stack of frame pairs, uint8 [..., H, W, 3] each, plus the JPEG round trip of the
    stored image.  Returns dict(diff_l, vis_l or None, back_l: uint8 [..., H, W]; back_max: int32 [...])."""
    dev = frames_pos.device if isinstance(frames_pos, torch.Tensor) and frames_pos.is_cuda else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        raise _lib.VidsegError("difference_images needs a CUDA device (no CPU fallback)")
    a, b = _u8(frames_pos, "frames_pos", dev), _u8(frames_neg, "frames_neg", dev)
    if a.shape != b.shape or a.dim() < 3 or a.shape[-1] != 3:
        raise _lib.VidsegError(f"difference_images: expected two uint8 [..., H, W, 3] stacks, got {tuple(a.shape)} / {tuple(b.shape)}")
    lead, (H, W) = a.shape[:-3], a.shape[-3:-1]
    n = int(np.prod(lead)) if len(lead) else 1
    mk = lambda: torch.empty((*lead, H, W), dtype=torch.uint8, device=dev)
    diff_l, back_l = mk(), mk()
    vis_l = mk() if want_vis else None
    back_max = torch.empty(lead if len(lead) else (1,), dtype=torch.int32, device=dev)
    blur_max = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.vidseg_segmap_difference(a.data_ptr(), b.data_ptr(), n, H, W, diff_l.data_ptr(),
                                                vis_l.data_ptr() if want_vis else None, back_l.data_ptr(),
                                                back_max.data_ptr(), blur_max.data_ptr(), _lib.stream_ptr()), "segmap_difference")
    return dict(diff_l=diff_l, vis_l=vis_l, back_l=back_l, back_max=back_max)


def seg_maps_from_frames(frames_pos, frames_neg, unique_labels, label_maps=None, filter_difference=False, filter_s=0.7,
                         want_vis=False):
    """get_seg_map_main (:74-167) on tensors: frames_pos / frames_neg uint8 [K, F, H, W, 3] (entry k belongs to
    unique_labels[k]); label_maps int32 [F, h, w] (only with filter_difference).  Returns dict(seg_raw uint8 [F, H, W],
    seg_index int32 [F, H, W], diff_l, vis_l, back_l)."""
    res = difference_images(frames_pos, frames_neg, want_vis=want_vis)
    back, bmax = res["back_l"], res["back_max"]
    if back.dim() != 4:
        raise _lib.VidsegError("seg_maps_from_frames: frames must be [K, F, H, W, 3]")
    K, F, H, W = back.shape
    dev = back.device
    ul = torch.as_tensor(np.asarray(unique_labels), dtype=torch.int32).to(dev)
    if ul.numel() != K:
        raise _lib.VidsegError(f"seg_maps_from_frames: {K} mask runs but {ul.numel()} labels")
    masks = None
    if filter_difference:
        if label_maps is None:
            raise _lib.VidsegError("filter_difference needs the K-means label maps")
        lm = label_maps if isinstance(label_maps, torch.Tensor) else torch.as_tensor(np.asarray(label_maps))
        masks = resized_masks(lm.to(device=dev, dtype=torch.int32), unique_labels, H, W)
    seg_raw = torch.empty((F, H, W), dtype=torch.uint8, device=dev)
    seg_index = torch.empty((F, H, W), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.vidseg_segmap_argmax(back.data_ptr(), bmax.data_ptr(), K, F, H, W,
                                            masks.data_ptr() if masks is not None else None, float(filter_s), ul.data_ptr(),
                                            seg_raw.data_ptr(), seg_index.data_ptr(), _lib.stream_ptr()), "segmap_argmax")
    res.update(seg_raw=seg_raw, seg_index=seg_index, mask_resized=masks)
    return res


def default_color_map(n=256):
    """Used when the reference's ``scripts/util/color_map_soft.txt`` is not on disk: a fixed pastel palette."""
    idx = np.arange(n)
    return np.stack([(idx * 67 + 80) % 200 + 40, (idx * 131 + 150) % 200 + 40, (idx * 29 + 30) % 200 + 40], 1).astype(np.float64)


def _read_frames(folder, names):
    return np.stack([np.array(Image.open(os.path.join(folder, f"{n}.png"))) for n in names])


def get_seg_map_main(exp_name, basecount, modulate_lambda, num_masks, num_frames, filter_difference, filter_s=0.7,
                     resize_height=28, resize_width=52, unique_labels=None, base_folder=None, mask_folder=None,
                     frame_name_list=None, feature_timestep="24", is_smooth=False, batch_id=None, color_map_path=None,
                     color_map_mapping="order", label_maps=None, frames=None, write_files=True):
    """reference :74-167 (which first calls generate_difference_map :41-70).  Same positional signature and outputs.
    Extras: ``frames=(pos, neg)`` uint8 [K, F, H, W, 3] tensors hand over the decoded modulated frames directly instead
    of the ``modulated_output`` PNG folders; ``label_maps`` [F, h, w] hands over the K-means label maps instead of the
    ``mask_folder`` PNG tree (filter_difference only); ``write_files=False`` skips every file.  Returns the dict of
    ``seg_maps_from_frames``."""
    if base_folder is None:
        base_folder = "outputs"
        modulated = f"outputs/modulate_video_sample/svd/{exp_name}"
    else:
        modulated = os.path.join(base_folder, f"{exp_name}/modulated_output")
    labels = np.asarray(unique_labels) if unique_labels is not None else np.arange(num_masks)
    names = [frame_name_list[i] if frame_name_list is not None else i for i in range(num_frames)]
    tag = lambda lam, i: f"{basecount:06d}_l_{lam}_mask_{i}"
    if frames is None:
        pos = np.stack([_read_frames(os.path.join(modulated, tag(modulate_lambda, i)), names) for i in labels])
        neg = np.stack([_read_frames(os.path.join(modulated, tag(-modulate_lambda, i)), names) for i in labels])
    else:
        pos, neg = frames
    if filter_difference and label_maps is None:
        from .feature_extraction import generate_aggregate_mask
        if mask_folder is None:
            mask_folder = f"features_outputs/kmeans_masks/{exp_name}/output_block_8_spatial_self_attn_q_masks_{num_masks}"
        probe = np.array(Image.open(os.path.join(mask_folder, f"kmeans_time_{feature_timestep}_frame_{names[0]}", f"mask_{labels[0]}.png")))
        fh, fw = probe.shape[:2]
        # the per-label PNGs partition the grid: rebuild the label map they were written from
        label_maps = np.stack([generate_aggregate_mask(mask_folder, feature_timestep, num_masks, n, fh, fw, labels=labels) for n in names])
    res = seg_maps_from_frames(pos, neg, labels, label_maps=label_maps, filter_difference=filter_difference,
                               filter_s=filter_s, want_vis=write_files)
    if not write_files:
        return res
    diff, vis = res["diff_l"].cpu().numpy(), res["vis_l"].cpu().numpy()
    for k, i in enumerate(labels):
        for sub, arr in (("original_map", diff), ("vis_map", vis)):
            folder = os.path.join(base_folder, f"{exp_name}/difference_map/{sub}/", tag(modulate_lambda, i))
            os.makedirs(folder, exist_ok=True)
            for f, n in enumerate(names):
                Image.fromarray(arr[k, f]).save(os.path.join(folder, f"{n}.jpg"))
    suffix = f"_f_{filter_s}" if filter_difference else ""
    seg_folder = os.path.join(base_folder, f"{exp_name}/segmentation_map{suffix}", f"{basecount:06d}_l_{modulate_lambda}")
    raw_folder = os.path.join(base_folder, f"{exp_name}/segmentation_map_raw{suffix}", f"{basecount:06d}_l_{modulate_lambda}")
    os.makedirs(seg_folder, exist_ok=True)
    os.makedirs(raw_folder, exist_ok=True)
    if color_map_path is None:
        color_map_path = "scripts/util/color_map_soft.txt"
    color_map = np.loadtxt(color_map_path, delimiter=",") if os.path.exists(color_map_path) else default_color_map()
    seg_raw, seg_index = res["seg_raw"].cpu().numpy(), res["seg_index"].cpu().numpy()
    for f, n in enumerate(names):
        Image.fromarray(seg_raw[f]).save(os.path.join(raw_folder, f"{n}.png"))
        colour = color_map[seg_index[f]] if color_map_mapping == "order" else color_map[seg_raw[f]]
        Image.fromarray(colour.astype(np.uint8)).save(os.path.join(seg_folder, f"{n}.jpg"))
    return res
