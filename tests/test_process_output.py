"""Seg-map post-process (reference scripts/sampling/process_output.py).

CPU: the in-memory oracle reproduces the goldens written by the UNMODIFIED reference (file round trips included), and
the restatement of the library routines (what the CUDA kernels implement) equals OpenCV / Pillow / libjpeg on this
machine.  GPU: the CUDA path through the C-ABI equals the goldens and the oracle bit for bit, up to the full
14 x 512 x 512 x 20-mask size, odd sizes and the file-writing mirror of ``get_seg_map_main``."""
import os

import numpy as np
import pytest

from oracle import process_output as opo, process_output_emul as emul
from synth import SEGMAP_CASES, synthetic_modulated_frames

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", [c[0] for c in SEGMAP_CASES])
def test_oracle_reproduces_reference_goldens(case):
    name, seed, K, F, H, W, fh, fw = next(c for c in SEGMAP_CASES if c[0] == case)
    g = np.load(os.path.join(GOLDEN, f"segmap_{name}.npz"))
    pos, neg, labels = synthetic_modulated_frames(seed, K, F, H, W, fh, fw)
    for filt, key in ((False, "raw"), (True, "raw_filtered")):
        got, _, stored = opo.seg_maps(pos, neg, np.arange(K), labels, filter_difference=filt, filter_s=0.7)
        assert np.array_equal(got, g[key])
    nf = g["back"].shape[1]
    back = np.stack([[opo.jpeg_roundtrip(stored[k, f]) for f in range(nf)] for k in range(K)])
    assert np.array_equal(back, g["back"])


def test_restated_library_routines_equal_the_libraries():
    import cv2
    from PIL import Image
    rng = np.random.RandomState(0)
    assert np.array_equal(emul.GAUSS_5_SIGMA3, cv2.getGaussianKernel(5, 3, cv2.CV_64F).ravel())
    for H, W in ((12, 13), (9, 16), (7, 5), (10, 18)):      # blur: widths with and without a remainder of the 4-wide body
        d = np.sqrt(rng.randint(0, 766, (H, W)).astype(np.float64))
        assert np.array_equal(emul.gaussian_blur_5x5_f64(d), cv2.GaussianBlur(d, (5, 5), 3)), (H, W)
    d = rng.rand(40, 50) * 300 - 20
    assert np.array_equal(emul.float64_to_L(d), np.array(Image.fromarray(d).convert("L")))
    for trial in range(24):                                  # JPEG: sizes off the 8-grid, flat / noisy / ramp content
        H, W = rng.randint(1, 70), rng.randint(1, 70)
        img = [rng.randint(0, 28, (H, W)), rng.randint(0, 256, (H, W)), np.add.outer(np.arange(H), np.arange(W)) % 256,
               np.clip(rng.randn(H, W) * 5 + 12, 0, 255)][trial % 4].astype(np.uint8)
        assert np.array_equal(emul.jpeg_roundtrip_L(img), opo.jpeg_roundtrip(img)), (H, W, trial % 4)
    for h, w, H, W in ((32, 32, 512, 512), (16, 16, 256, 256), (9, 7, 100, 64), (48, 48, 768, 768), (64, 64, 32, 24)):
        m = (rng.rand(h, w) < 0.3).astype(np.uint8) * 255
        assert np.array_equal(emul.lanczos_resize_L(m, H, W), np.array(Image.fromarray(m).resize((W, H), Image.LANCZOS)))


def test_restated_chain_equals_oracle_on_a_small_case():
    pos, neg, labels = synthetic_modulated_frames(3, 3, 2, 24, 21, 6, 5)
    for filt in (False, True):
        want, _, _ = opo.seg_maps(pos, neg, np.arange(3), labels, filt, 0.7)
        got, _ = emul.seg_maps(pos, neg, np.arange(3), labels, filt, 0.7)
        assert np.array_equal(got, want)


def test_host_lanczos_windows_equal_the_restatement():
    from vidseg_diffusion_b200.process_output import lanczos_windows
    for a, b in ((32, 512), (16, 256), (7, 64), (48, 768), (64, 32)):
        bounds, coeffs, ksize = lanczos_windows(a, b)
        rb, rk = emul.lanczos_coeffs(a, b)
        assert ksize == rk.shape[1] and np.array_equal(bounds, rb) and np.array_equal(coeffs, rk)


# --------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", [c[0] for c in SEGMAP_CASES])
def test_gpu_matches_reference_goldens(cuda, case):
    import torch
    from vidseg_diffusion_b200.process_output import seg_maps_from_frames
    name, seed, K, F, H, W, fh, fw = next(c for c in SEGMAP_CASES if c[0] == case)
    g = np.load(os.path.join(GOLDEN, f"segmap_{name}.npz"))
    pos, neg, labels = synthetic_modulated_frames(seed, K, F, H, W, fh, fw)
    pos, neg = torch.from_numpy(pos).to(cuda), torch.from_numpy(neg).to(cuda)
    for filt, key in ((False, "raw"), (True, "raw_filtered")):
        res = seg_maps_from_frames(pos, neg, np.arange(K), torch.from_numpy(labels).to(cuda), filter_difference=filt, filter_s=0.7)
        assert np.array_equal(res["seg_raw"].cpu().numpy(), g[key])
    nf = g["back"].shape[1]
    assert np.array_equal(res["back_l"][:, :nf].cpu().numpy(), g["back"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(20, 14, 512, 512, 32, 32), (4, 3, 67, 45, 9, 6), (2, 1, 8, 8, 2, 2), (3, 2, 100, 260, 12, 30)])
def test_gpu_matches_oracle_every_stage(cuda, shape):
    """Every intermediate the reference writes to disk (stored difference image, vis image, JPEG round trip) and the final
    maps, against the oracle (OpenCV / Pillow on this box's CPU), at the full benchmark size and at sizes off every grid."""
    import torch
    from PIL import Image
    from vidseg_diffusion_b200.process_output import resized_masks, seg_maps_from_frames
    K, F, H, W, fh, fw = shape
    pos, neg, labels = synthetic_modulated_frames(5, K, F, H, W, fh, fw)
    unique = np.arange(K) * 3 + 1 if K < 20 else np.arange(K)        # labels need not be 0..K-1
    labels_u = (labels * 3 + 1) if K < 20 else labels
    dp, dn = torch.from_numpy(pos).to(cuda), torch.from_numpy(neg).to(cuda)
    for filt in (False, True):
        res = seg_maps_from_frames(dp, dn, unique, torch.from_numpy(labels_u.astype(np.int32)).to(cuda),
                                   filter_difference=filt, filter_s=0.7, want_vis=True)
        want, _, stored = opo.seg_maps(pos, neg, unique, labels_u, filter_difference=filt, filter_s=0.7)
        assert np.array_equal(res["diff_l"].cpu().numpy(), stored)
        assert np.array_equal(res["seg_raw"].cpu().numpy(), want)
    for k, f in ((0, 0), (K - 1, F - 1)):
        _, img, vis = opo.difference_map(pos[k, f], neg[k, f])
        assert np.array_equal(res["vis_l"][k, f].cpu().numpy(), vis)
        assert np.array_equal(res["back_l"][k, f].cpu().numpy(), opo.jpeg_roundtrip(img))
        m = np.array(Image.fromarray(np.where(labels_u[f] == unique[k], 255, 0).astype(np.uint8)).resize((W, H), Image.LANCZOS))
        assert np.array_equal(res["mask_resized"][k, f].cpu().numpy(), m)


@pytest.mark.gpu
def test_gpu_get_seg_map_main_writes_the_reference_tree(cuda, tmp_path):
    """The file-based mirror: same folders and file names as the reference run recorded in the golden, raw PNGs equal."""
    import cv2
    from PIL import Image
    from vidseg_diffusion_b200.process_output import get_seg_map_main
    name, seed, K, F, H, W, fh, fw = SEGMAP_CASES[0]
    g = np.load(os.path.join(GOLDEN, f"segmap_{name}.npz"))
    pos, neg, labels = synthetic_modulated_frames(seed, K, F, H, W, fh, fw)
    root, exp, lam = str(tmp_path), "exp", 4.0
    for sign, frames in ((lam, pos), (-lam, neg)):
        for k in range(K):
            folder = os.path.join(root, exp, "modulated_output", f"{0:06d}_l_{sign}_mask_{k}")
            os.makedirs(folder)
            for f in range(F):
                cv2.imwrite(os.path.join(folder, f"{f}.png"), cv2.cvtColor(frames[k, f], cv2.COLOR_RGB2BGR))
    mask_folder = os.path.join(root, exp, "kmeans_masks", f"blocks_masks_{K}")
    for f in range(F):
        folder = os.path.join(mask_folder, f"kmeans_time_24_frame_{f}")
        os.makedirs(folder)
        for k in range(K):
            Image.fromarray(np.where(labels[f] == k, 255, 0).astype(np.uint8)).save(os.path.join(folder, f"mask_{k}.png"))
    for filt, fs in ((False, 1.0), (True, 0.7)):
        get_seg_map_main(exp, 0, lam, K, F, filter_difference=filt, filter_s=fs, resize_height=fh, resize_width=fw,
                         unique_labels=np.arange(K), base_folder=root, mask_folder=mask_folder, feature_timestep="24")
        sub = f"segmentation_map_raw_f_{fs}" if filt else "segmentation_map_raw"
        raw = np.stack([np.array(Image.open(os.path.join(root, exp, sub, f"{0:06d}_l_{lam}", f"{f}.png"))) for f in range(F)])
        assert np.array_equal(raw, g["raw_filtered" if filt else "raw"])
    tree = sorted(os.path.relpath(os.path.join(d, f), root) for d, _, fs_ in os.walk(os.path.join(root, exp)) for f in fs_
                  if "modulated_output" not in d and "kmeans_masks" not in d)
    assert tree == list(g["tree"])
    # the stored difference JPEGs decode to what the reference's decode to
    back = np.stack([np.stack([np.array(Image.open(os.path.join(root, exp, "difference_map", "original_map",
                                                                f"{0:06d}_l_{lam}_mask_{k}", f"{f}.jpg"))) for f in range(F)])
                     for k in range(K)])
    assert np.array_equal(back, g["back"])
