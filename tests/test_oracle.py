"""CPU: the oracle restatements against the goldens produced by the UNMODIFIED reference
(tests/golden/make_goldens.py) and against the installed scikit-learn."""
import os
import warnings

import numpy as np
import pytest

from oracle import features as ofeat
from oracle import kmeans as okm
from oracle import refine as oref
from synth import CLUSTER_CASES, synthetic_clip_features

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FAST_CASES = [c for c in CLUSTER_CASES if c[0] != "c2_objects"]


def test_aggregate_matches_torch_mean_bitwise():
    import torch
    blocks, _ = synthetic_clip_features(7, 3, 5, 6, 24, 3)
    want = torch.mean(torch.stack([torch.from_numpy(b) for b in blocks]), dim=0).numpy()
    got = ofeat.aggregate_blocks(blocks)
    assert np.array_equal(want, got)
    x = ofeat.aggregate_normalize(blocks, 3)
    assert x.shape == (3 * 30, 24) and np.all(np.abs(x).max(axis=1) == 1.0)


@pytest.mark.parametrize("case", FAST_CASES, ids=[c[0] for c in FAST_CASES])
def test_oracle_reproduces_reference_goldens(case):
    name, seed, F, h, w, C, K, kind = case
    g = np.load(os.path.join(GOLDEN, f"cluster_{name}.npz"))
    blocks, _ = synthetic_clip_features(seed, F, h, w, C, K, kind=kind)
    x = ofeat.aggregate_normalize(blocks, F)
    np.random.seed(seed)
    labels, _ = okm.kmeans_fit_predict(x, K)
    labels = labels.reshape(F, h, w)
    assert np.array_equal(labels, g["labels"])  # bit-exact cluster-index maps
    ref_mask, all_h, all_w, _ = oref.correct_low_res_mask(blocks[1], labels, h, w, F)
    assert np.array_equal(all_h, g["all_h"]) and np.array_equal(all_w, g["all_w"])
    assert np.array_equal(ref_mask, g["ref_mask"])


@pytest.mark.parametrize("n,d,k,kind", [(300, 17, 3, "iid"), (1024, 64, 5, "iid"), (777, 40, 6, "blobs"), (64, 8, 1, "iid")])
def test_oracle_kmeans_equals_sklearn(n, d, k, kind):
    from sklearn.cluster import KMeans
    r = np.random.RandomState(n + d)
    if kind == "blobs":
        cent = r.standard_normal((k, d)).astype(np.float32)
        X = cent[r.randint(0, k, n)] + 0.1 * r.standard_normal((n, d)).astype(np.float32)
    else:
        X = r.standard_normal((n, d)).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(5)
        km = KMeans(n_clusters=k, n_init=10)
        km.fit(X)
        want = km.predict(X)
        pos_ref = np.random.get_state()[2]
    np.random.seed(5)
    info = {}
    got, centers = okm.kmeans_fit_predict(X, k, info=info)
    assert np.random.get_state()[2] == pos_ref  # same number of draws from the global stream
    assert np.array_equal(want, got)
    assert np.allclose(centers, km.cluster_centers_, atol=1e-5)
    assert abs(info["inertia"] - km.inertia_) <= 1e-4 * max(1.0, km.inertia_)
    assert info["n_iter"] == km.n_iter_


def test_oracle_refine_vote_rules():
    # trajectory 0 stays, trajectory 1 jumps right by 2 (rejected), ties go to the first label seen
    seg = np.array([[[0, 1]], [[1, 1]], [[0, 0]], [[1, 0]]])  # [F=4, 1, 2]
    all_h = np.zeros((4, 2), dtype=int)
    all_w = np.array([[0, 1], [0, 1], [0, 1], [0, 1]])
    new, keep = oref.refine_labels(seg, all_h, all_w)
    assert keep.tolist() == [True, True]
    assert new[:, 0, 0].tolist() == [0, 0, 0, 0]  # labels 0,1,0,1 -> tie -> first seen (0)
    assert new[:, 0, 1].tolist() == [1, 1, 1, 1]  # labels 1,1,0,0 -> tie -> first seen (1)
    all_w2 = np.array([[0, 0], [0, 2], [0, 0], [0, 0]])
    _, keep2 = oref.refine_labels(np.zeros((4, 1, 3), dtype=int), np.zeros((4, 2), dtype=int), all_w2)
    assert keep2.tolist() == [True, False]


def test_vae_oracle_matches_reference_goldens():
    """Next row, oracle first: the SD-2.1 first stage restated in oracle/vae.py against
    the reference Encoder / Decoder (tests/golden/make_vae_goldens.py).  No product code on this row yet."""
    import torch
    from oracle import vae as ovae
    from synth import synthetic_unet_weights
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "vae_tiny.npz"))
    seed, _ = (int(v) for v in g["meta"])
    cfg = ovae.TINY_VAE_CONFIG
    shapes = ovae.param_shapes(cfg)
    assert list(g["keys"]) == sorted(shapes)
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, seed).items()}
    z = ovae.encode_first_stage(sd, cfg, torch.from_numpy(g["x"]), 0.18215, torch.from_numpy(g["noise"]))
    img = ovae.decode_first_stage(sd, cfg, torch.from_numpy(g["z"]), 0.18215)
    rel = lambda a, b: float((a - torch.from_numpy(b)).abs().max() / np.abs(b).max())
    assert rel(z, g["z"]) < 2e-5 and rel(img, g["image"]) < 2e-5
    assert z.shape == (2, 4, 4, 4) and img.shape == (2, 3, 32, 32)


def test_video_decoder_oracle_matches_reference_golden():
    """SVD's first-stage decoder (temporal_ae.VideoDecoder, "conv-only"): oracle/vae.py against the reference module."""
    import torch
    from oracle import vae as ovae
    from synth import synthetic_unet_weights
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "vae_video_tiny.npz"))
    seed, T = (int(v) for v in g["meta"])
    cfg = ovae.TINY_VAE_CONFIG
    shapes = ovae.video_decoder_param_shapes(cfg)
    assert list(g["keys"]) == sorted(shapes)
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, seed).items()}
    img = ovae.decode_first_stage(sd, cfg, torch.from_numpy(g["z"]), 0.18215, timesteps=T)
    assert float((img - torch.from_numpy(g["image"])).abs().max() / np.abs(g["image"]).max()) < 2e-5
    # clips are independent: decoding the second clip alone gives its frames again
    alone = ovae.decode_first_stage(sd, cfg, torch.from_numpy(g["z"][T:]), 0.18215, timesteps=T)
    assert torch.allclose(alone, img[T:], atol=1e-5)
