"""GPU parity of R1/R2/R3 through the C-ABI: against the oracle on seeded inputs and against the
goldens generated from the unmodified reference.  Integer outputs are compared bit-exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import features as ofeat
from oracle import kmeans as okm
from oracle import refine as oref
from synth import CLUSTER_CASES, synthetic_clip_features

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _dev(arrs, cuda):
    return [torch.from_numpy(a).to(cuda) for a in arrs]


@pytest.mark.parametrize("F,hw,C,nb", [(4, 256, 640, 3), (14, 1024, 640, 3), (3, 63, 40, 3), (2, 17, 37, 2),
                                       (1, 5, 1, 1), (2, 9, 1280, 1), (2, 33, 1284, 4), (0, 16, 8, 3)])
def test_aggregate_normalize_bit_exact(cuda, F, hw, C, nb):
    from vidseg_diffusion_b200.features import aggregate_normalize
    r = np.random.RandomState(F * 1000 + hw + C)
    blocks = [r.standard_normal((2 * F, hw, C)).astype(np.float32) for _ in range(nb)]
    got = aggregate_normalize(_dev(blocks, cuda), F).cpu().numpy()
    want = ofeat.aggregate_normalize(blocks, F) if F > 0 else np.zeros((0, C), np.float32)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # bit-exact, fp32


@pytest.mark.parametrize("case", CLUSTER_CASES, ids=[c[0] for c in CLUSTER_CASES])
def test_cluster_and_refine_match_reference_goldens(cuda, case):
    """aggregate -> K-means -> refine on the GPU == the reference's own output (bit-exact ints)."""
    from vidseg_diffusion_b200.features import aggregate_normalize
    from vidseg_diffusion_b200.kmeans import KMeans
    from vidseg_diffusion_b200.refine import refine_masks
    name, seed, F, h, w, C, K, kind = case
    g = np.load(os.path.join(GOLDEN, f"cluster_{name}.npz"))
    blocks, _ = synthetic_clip_features(seed, F, h, w, C, K, kind=kind)
    dblocks = _dev(blocks, cuda)
    x = aggregate_normalize(dblocks, F)
    np.random.seed(seed)
    km = KMeans(n_clusters=K, n_init=10)
    labels = km.fit_predict(x).reshape(F, h, w)
    assert np.array_equal(labels.cpu().numpy(), g["labels"]), f"{(labels.cpu().numpy() != g['labels']).sum()} labels differ"
    refined, traj, keep = refine_masks(dblocks[1], labels, F, h, w)
    traj = traj.cpu().numpy()
    assert np.array_equal(traj // w, g["all_h"]) and np.array_equal(traj % w, g["all_w"])
    assert np.array_equal(refined.cpu().numpy().reshape(-1), g["ref_mask"])


@pytest.mark.parametrize("n,d,k,kind", [(300, 17, 3, "iid"), (1024, 64, 5, "iid"), (777, 40, 6, "blobs"),
                                        (64, 8, 1, "iid"), (2000, 33, 50, "iid"), (70, 5, 70, "iid")])
def test_kmeans_matches_oracle(cuda, n, d, k, kind):
    from vidseg_diffusion_b200.kmeans import KMeans
    r = np.random.RandomState(n + d)
    if kind == "blobs":
        cent = r.standard_normal((k, d)).astype(np.float32)
        X = cent[r.randint(0, k, n)] + 0.1 * r.standard_normal((n, d)).astype(np.float32)
    else:
        X = r.standard_normal((n, d)).astype(np.float32)
    np.random.seed(11)
    info = {}
    want, centers = okm.kmeans_fit_predict(X, k, info=info)
    pos = np.random.get_state()[2]
    np.random.seed(11)
    km = KMeans(n_clusters=k, n_init=10)
    got = km.fit_predict(torch.from_numpy(X).to(cuda)).cpu().numpy()
    assert np.random.get_state()[2] == pos
    mism = int((got != want).sum())
    assert mism == 0, f"{mism}/{n} labels differ (iters gpu {km.n_iter_} oracle {info['n_iter']})"
    assert np.allclose(km.cluster_centers_.cpu().numpy(), centers, atol=2e-5)
    assert abs(km.inertia_ - info["inertia"]) <= 1e-4 * max(1.0, info["inertia"])
    assert km.n_iter_ == info["n_iter"]
    # predict on new data == oracle predict
    Y = r.standard_normal((50, d)).astype(np.float32)
    assert np.array_equal(km.predict(torch.from_numpy(Y).to(cuda)).cpu().numpy(), okm.kmeans_predict(Y, centers))


def test_kmeans_properties_full_size(cuda):
    """BASELINE.json config 2 size (N=14336, D=640, K=20): size-independent properties."""
    from vidseg_diffusion_b200.kmeans import KMeans
    g = torch.Generator(device="cpu").manual_seed(0)
    K, N, D = 20, 14336, 640
    cent = torch.randn(K, D, generator=g)
    lab = torch.randint(0, K, (N,), generator=g)
    X = (cent[lab] + 0.05 * torch.randn(N, D, generator=g)).to(cuda)
    np.random.seed(1)
    km = KMeans(n_clusters=K, n_init=10)
    got = km.fit_predict(X)
    # well separated blobs: the partition must equal the planted one (up to naming)
    pairs = torch.unique(torch.stack([got.cpu().long(), lab]), dim=1)
    assert pairs.shape[1] == K
    # idempotence: predict(X) again gives the same labels; centres are the cluster means
    assert torch.equal(km.predict(X), got)
    means = torch.stack([X[got == j].double().mean(0) for j in range(K)]).float()
    assert torch.allclose(means, km.cluster_centers_, atol=1e-5)
    # determinism for a fixed seed
    np.random.seed(1)
    again = KMeans(n_clusters=K, n_init=10).fit_predict(X)
    assert torch.equal(again, got)


def test_refine_vote_rules_on_device(cuda):
    from vidseg_diffusion_b200.refine import refine_masks
    # F=3, 1x4 grid, features make every cell match itself (orthogonal one-hot rows)
    F, h, w, C = 3, 1, 4, 8
    eye = torch.eye(C)[:h * w]
    fm = torch.cat([torch.randn(F, h * w, C), eye[None].repeat(F, 1, 1)], 0).to(cuda)
    labels = torch.tensor([[0, 1, 2, 2], [1, 1, 0, 3], [0, 2, 0, 3]], dtype=torch.int32, device=cuda).reshape(F, h, w)
    refined, traj, keep = refine_masks(fm, labels, F, h, w)
    assert traj.cpu().tolist() == [[0, 1, 2, 3]] * 3 and keep.cpu().tolist() == [1, 1, 1, 1]
    want, _ = oref.refine_labels(labels.cpu().numpy(), np.zeros((3, 4), int), np.tile(np.arange(4), (3, 1)))
    assert np.array_equal(refined.cpu().numpy(), want)


def test_feature_extraction_main_layout(cuda, tmp_path):
    """The plugin-level mirror writes the reference's folder layout and returns its triple."""
    from vidseg_diffusion_b200 import feature_extraction as fe
    name, seed, F, h, w, C, K, kind = CLUSTER_CASES[0]
    g = np.load(os.path.join(GOLDEN, f"cluster_{name}.npz"))
    blocks, _ = synthetic_clip_features(seed, F, h, w, C, K, kind=kind)
    names = ["output_block_8", "output_block_7", "output_block_6"]
    fm_dir = tmp_path / "exp" / "feature_maps"
    fm_dir.mkdir(parents=True)
    for n, b in zip(names, blocks):
        torch.save(torch.from_numpy(b), fm_dir / f"{n}_spatial_self_attn_q_time_24.pt")
    np.random.seed(seed)
    uniq, _, _ = fe.feature_extraction_main("kmeans_masks", K, 24, ",".join(names), "exp", "exp", "spatial_self_attn_q",
                                            h, w, "24", base_folder=str(tmp_path), num_frames=F)
    assert np.array_equal(uniq, np.arange(K))
    tree = sorted(os.path.relpath(os.path.join(d, f), tmp_path) for d, _, fs in os.walk(tmp_path / "exp" / "kmeans_masks") for f in fs)
    assert tree == list(g["png_tree"])
    mask_folder = str(tmp_path / "exp" / "kmeans_masks" / ("_".join(names) + f"_spatial_self_attn_q_masks_{K}"))
    labels = np.stack([fe.generate_aggregate_mask(mask_folder, 24, K, i, h, w) for i in range(F)])
    assert np.array_equal(labels, g["labels"])
    fe._LABEL_CACHE.clear()  # force the PNG round trip of the reference
    _, ref_mask, _ = fe.feature_extraction_main("correct_low_res_mask", K, 24, "output_block_7", "exp", "exp",
                                                "spatial_self_attn_q", h, w, "24", base_folder=str(tmp_path),
                                                num_frames=F, mask_folder=mask_folder, ref_unique_labels=uniq)
    assert np.array_equal(ref_mask, g["ref_mask"])
    ctree = sorted(os.path.relpath(os.path.join(d, f), tmp_path) for d, _, fs in os.walk(mask_folder + "_corrected") for f in fs)
    assert ctree == list(g["corrected_tree"])


@pytest.mark.parametrize("n,d,k,rows", [(1500, 72, 7, (0, 1500)), (4096, 640, 20, (0, 4096)), (1500, 72, 7, (300, 1111)),
                                        (14336, 640, 20, (0, 14336))])
def test_mstep_tensor_core_sums_match_float64_sums(cuda, n, d, k, rows):
    """The int8 tensor-core M-step (exact fixed-point sums) against the float64 shared-memory M-step on the same
    labels: counts identical, sums equal to float64 rounding, and against a numpy float64 sum of the centred rows."""
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.distributed import CudaLloydBackend
    from vidseg_diffusion_b200.kmeans import draw_kmeanspp_randoms
    lib = _lib.load()
    r = np.random.RandomState(n + d + k)
    scale = np.exp(r.uniform(-6, 1, size=d)).astype(np.float32)        # columns of very different magnitude
    X = torch.from_numpy((r.standard_normal((n, d)).astype(np.float32) * scale + 0.3).astype(np.float32)).to(cuda)
    np.random.seed(5)
    first, rand = draw_kmeanspp_randoms(n, k, 10)
    out = {}
    keep = lib.vidseg_get_kmeans_mstep()
    try:
        for mode in (0, 1):
            _lib.check(lib.vidseg_set_kmeans_mstep(mode), "set_kmeans_mstep")
            be = CudaLloydBackend(k, 10, 300, 1e-4)
            be.prepare(X)
            be.seed(first, rand)
            be.assign(0, n)
            partial, _ = be.partial(*rows)
            out[mode] = partial.cpu().numpy()
            be.assign(0, n)   # labels unchanged; exercises a second pass over the same buffers
            again, _ = be.partial(*rows)
            assert np.array_equal(again.cpu().numpy(), out[mode])
            be.release()
    finally:
        lib.vidseg_set_kmeans_mstep(keep)
    assert np.array_equal(out[0][..., d], out[1][..., d])
    assert out[1][..., d].sum() == 10 * (rows[1] - rows[0])
    ref_mag = np.abs(out[0][..., :d]).max(axis=(0, 1), keepdims=True) + 1e-30
    # fixed point: every element is rounded to 2^-46 of the largest centred magnitude (most are exact); the float64
    # chain rounds at 2^-53 of its running sum
    xc = X.double().cpu().numpy()
    xmax = float(np.abs(xc - xc.mean(0, keepdims=True)).max())
    err = np.abs(out[0][..., :d] - out[1][..., :d])
    assert float(err.max()) <= (rows[1] - rows[0]) * xmax * 2.0 ** -46, float(err.max())
    assert float((err / ref_mag).max()) < 1e-6


def test_kmeans_fit_identical_under_both_msteps(cuda):
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.kmeans import KMeans
    lib = _lib.load()
    r = np.random.RandomState(3)
    X = torch.from_numpy(r.standard_normal((3000, 64)).astype(np.float32)).to(cuda)   # no structure: many iterations
    res = {}
    keep = lib.vidseg_get_kmeans_mstep()
    try:
        for mode in (0, 1):
            _lib.check(lib.vidseg_set_kmeans_mstep(mode), "set_kmeans_mstep")
            np.random.seed(2)
            km = KMeans(n_clusters=9, n_init=10)
            res[mode] = (km.fit_predict(X).cpu().numpy(), km.cluster_centers_.cpu().numpy(), km.n_iter_)
    finally:
        lib.vidseg_set_kmeans_mstep(keep)
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][2] == res[1][2]
    assert np.allclose(res[0][1], res[1][1], atol=1e-6)
