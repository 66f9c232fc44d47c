"""``match_gt_mask`` mode (reference scripts/sampling/feature_extraction.py:546-643).

CPU: the oracle against goldens produced by the UNMODIFIED reference + scikit-learn (tests/golden/make_match_gt_goldens.py,
two consecutive windows per case).  GPU: the CUDA path (K-means -> majority map -> tcgen05-filtered float64 4-NN) against
the same goldens and against the oracle on seeded inputs, bit-exact labels."""
import os

import numpy as np
import pytest
import torch

from oracle import match_gt as omg
from synth import synthetic_clip_features

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["small", "c1", "c1_usegt", "mid"]


def _windows(g):
    seed, F, h, w, C, K, use_gt = (int(v) for v in g["meta"])
    for win in range(2):
        blocks, _ = synthetic_clip_features(seed + 100 * win, F, h, w, C, K, n_blocks=1)
        yield win, blocks[0], (seed, F, h, w, C, K, bool(use_gt))


@pytest.mark.parametrize("case", CASES)
def test_oracle_match_gt_reproduces_reference(case):
    g = np.load(os.path.join(GOLDEN, f"matchgt_{case}.npz"))
    rm = rf = None
    for win, feats, (seed, F, h, w, C, K, use_gt) in _windows(g):
        if win == 0:
            np.random.seed(seed)
        uniq, rm, rf = omg.match_gt_mask(feats, F, h, w, K, gt_mask=g["gt_small"], ref_mask=rm, ref_feature_map=rf,
                                         use_gt_mask=use_gt)
        assert np.array_equal(rm, g[f"labels{win}"])
        assert np.array_equal(uniq, g[f"unique{win}"])


def test_oracle_knn_ties_and_votes():
    ref = np.array([[0.0, 0], [1, 0], [0, 1], [1, 1], [5, 5]], dtype=np.float32)
    lab = np.array([7, 3, 3, 7, 9])
    # query at the centre of the unit square: four equidistant neighbours -> two votes each -> smallest label
    assert omg.knn_predict(ref, lab, np.array([[0.5, 0.5]], dtype=np.float32), 4)[0] == 3
    assert omg.knn_predict(ref, lab, np.array([[4.0, 4.0]], dtype=np.float32), 1)[0] == 9
    assert omg.majority_map(np.array([0, 0, 1, 1, 1]), np.array([5, 4, 9, 9, 2])).tolist() == [4, 4, 9, 9, 9]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_match_gt_windows_match_reference_golden(cuda, case, tmp_path):
    from PIL import Image
    from vidseg_diffusion_b200.feature_extraction import feature_extraction_main
    g = np.load(os.path.join(GOLDEN, f"matchgt_{case}.npz"))
    rm = rf = ru = None
    for win, feats, (seed, F, h, w, C, K, use_gt) in _windows(g):
        gt_path = str(tmp_path / "gt.png")
        if win == 0:
            Image.fromarray(g["gt_small"].reshape(h, w).astype(np.uint8)).save(gt_path)   # already at the feature grid
            np.random.seed(seed)
        exp = f"win{win}"
        feats_d = {("output_block_8", "spatial_self_attn_q", 24): torch.from_numpy(feats).to(cuda)}
        uniq, rm, rf = feature_extraction_main(
            "match_gt_mask", K, 24, "output_block_8", exp, exp, "spatial_self_attn_q", h, w, "24", frame_name_list=None,
            base_folder=str(tmp_path), ref_mask=rm, ref_feature_map=rf, ref_unique_labels=ru, gt_mask_path=gt_path,
            num_frames=F, use_gt_mask=use_gt, features=feats_d)
        if win == 0:
            ru = uniq
        assert rm.dtype == torch.int32 and rm.is_cuda and rf.is_cuda
        assert np.array_equal(rm.cpu().numpy(), g[f"labels{win}"])
        assert np.array_equal(np.asarray(uniq), g[f"unique{win}"])
        got_tree = sorted(os.path.relpath(os.path.join(d, f), str(tmp_path / exp)) for d, _, fs in
                          os.walk(str(tmp_path / exp / "match_gt_mask")) for f in fs)
        want_tree = [t for t in g[f"tree{win}"].tolist() if not t.endswith("kmeans_cluster_labels.png")]  # colour preview: not written
        assert got_tree == want_tree


@pytest.mark.gpu
@pytest.mark.parametrize("n_ref,n_q,d,k,n_lab", [(256, 1000, 64, 4, 5), (1024, 4100, 640, 4, 20), (5, 33, 8, 4, 3),
                                                  (14336, 2048, 640, 4, 20), (300, 300, 128, 1, 7), (97, 513, 72, 8, 4)])
def test_gpu_knn_matches_oracle(cuda, n_ref, n_q, d, k, n_lab):
    from vidseg_diffusion_b200.match_gt import knn_predict
    r = np.random.RandomState(n_ref + n_q + d)
    ref = r.standard_normal((n_ref, d)).astype(np.float32)
    ref /= np.abs(ref).max(axis=1, keepdims=True)
    q = r.standard_normal((n_q, d)).astype(np.float32)
    q /= np.abs(q).max(axis=1, keepdims=True)
    q[: min(n_q, n_ref) // 2] = ref[: min(n_q, n_ref) // 2]        # queries that ARE reference points (first window)
    lab = (r.randint(0, n_lab, n_ref) * 11 + 2).astype(np.int32)
    want = omg.knn_predict(ref, lab, q, k)
    got = knn_predict(torch.from_numpy(ref).to(cuda), torch.from_numpy(lab).to(cuda), torch.from_numpy(q).to(cuda), k)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
def test_gpu_majority_map_matches_oracle(cuda):
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.match_gt import majority_map
    r = np.random.RandomState(3)
    fake = r.randint(0, 12, 1024).astype(np.int32)
    gt = (r.randint(0, 6, 1024) * 40 + 3).astype(np.int32)
    got = majority_map(torch.from_numpy(fake).to(cuda), torch.from_numpy(gt).to(cuda), 12)
    assert np.array_equal(got.cpu().numpy(), omg.majority_map(fake, gt))
    with pytest.raises(_lib.VidsegError):
        majority_map(torch.from_numpy(fake).to(cuda), torch.from_numpy(gt + 5000).to(cuda), 12)
