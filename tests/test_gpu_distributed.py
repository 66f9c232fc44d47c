"""GPU: the frame-sharded path's kernels and plumbing on one device.

(1) Two "virtual ranks" (two workspaces on the same GPU, each owning half of the rows) are driven in lock-step through
    the split E-step / M-step C-ABI with their partial buffers summed by hand -- exactly what the NCCL all-reduce does
    between vidseg_kmeans_partial and vidseg_kmeans_update -- and must reproduce the reference labels (golden from the
    unmodified reference + sklearn) bit for bit.
(2) ShardedClipSegmenter on a 1-rank NCCL group must equal ClipSegmenter.
The true multi-process run (2 GPUs, NCCL) is tools/run_sharded.py under torchrun; its output is kept in profiles/."""
import os

import numpy as np
import pytest
import torch

from synth import CLUSTER_CASES, synthetic_clip_features

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", ["c1_objects", "mid_objects", "c2_objects"])
def test_two_virtual_ranks_reproduce_reference_labels(cuda, case):
    from vidseg_diffusion_b200 import distributed as D
    from vidseg_diffusion_b200.features import aggregate_normalize
    from vidseg_diffusion_b200.kmeans import draw_kmeanspp_randoms
    name, seed, F, h, w, C, K, kind = next(c for c in CLUSTER_CASES if c[0] == case)
    g = np.load(os.path.join(GOLDEN, f"cluster_{name}.npz"))
    blocks, _ = synthetic_clip_features(seed, F, h, w, C, K, kind=kind)
    X = aggregate_normalize([torch.from_numpy(b).to(cuda) for b in blocks], F)
    n = X.shape[0]
    parts = D.frame_partition(F, 2)
    ranges = [(a * h * w, b * h * w) for a, b in parts]
    np.random.seed(seed)
    first, rand = draw_kmeanspp_randoms(n, K, 10)
    bes = [D.CudaLloydBackend(K, 10, 300, 1e-4) for _ in ranges]
    for be in bes:
        be.prepare(X)
        be.seed(first, rand)
    it = 0
    while it < 300:
        for _ in range(4):
            outs = []
            for be, (r0, r1) in zip(bes, ranges):
                be.assign(r0, r1)
                outs.append(be.partial(r0, r1))
            partial = outs[0][0] + outs[1][0]           # the all-reduce
            changed = outs[0][1] + outs[1][1]
            for be in bes:
                be.update(partial.clone(), changed.clone(), local_rows_only=True)
        it += 4
        st = [be.status() for be in bes]
        assert st[0] == st[1]                           # ranks stay in lock-step
        assert st[0][1] == 0, "empty cluster in a well-separated case"
        if st[0][0] == 0:
            break
    inertia = sum(be.inertia(r0, r1) for be, (r0, r1) in zip(bes, ranges))
    same = torch.minimum(*[be.same_matrix(r0, r1) for be, (r0, r1) in zip(bes, ranges)])
    best = D.pick_best(inertia.float().cpu().numpy(), same.cpu().numpy())
    centers = [be.finish(best) for be in bes]
    assert torch.equal(centers[0], centers[1])
    labels = torch.cat([be.predict(X[r0:r1].contiguous(), centers[0]) for be, (r0, r1) in zip(bes, ranges)])
    for be in bes:
        be.release()
    assert np.array_equal(labels.cpu().numpy(), g["labels"].reshape(-1))


@pytest.mark.parametrize("case", ["c1_objects", "c2_objects"])
def test_two_virtual_ranks_integer_exchange_words_are_bit_identical_to_one_gpu(cuda, case):
    """The exchange-word form (one fused array per iteration): with integer words the hand-made all-reduce (an int64 add)
    is exact, so the sharded fit must give the single-GPU fit's centres BIT FOR BIT, not just its labels."""
    from vidseg_diffusion_b200 import distributed as D
    from vidseg_diffusion_b200.features import aggregate_normalize
    from vidseg_diffusion_b200.kmeans import KMeans, draw_kmeanspp_randoms
    name, seed, F, h, w, C, K, kind = next(c for c in CLUSTER_CASES if c[0] == case)
    g = np.load(os.path.join(GOLDEN, f"cluster_{name}.npz"))
    blocks, _ = synthetic_clip_features(seed, F, h, w, C, K, kind=kind)
    X = aggregate_normalize([torch.from_numpy(b).to(cuda) for b in blocks], F)
    n = X.shape[0]
    np.random.seed(seed)
    km = KMeans(n_clusters=K, n_init=10)
    want_labels = km.fit_predict(X)
    parts = D.frame_partition(F, 2)
    ranges = [(a * h * w, b * h * w) for a, b in parts]
    np.random.seed(seed)
    first, rand = draw_kmeanspp_randoms(n, K, 10)
    bes = [D.CudaLloydBackend(K, 10, 300, 1e-4) for _ in ranges]
    for be in bes:
        be.prepare(X)
        be.seed(first, rand)
    modes = {be.exchange_mode(ranges) for be in bes}
    assert modes == {"i64"}
    state = None
    for it in range(0, 300, D.BURST):
        for _ in range(D.BURST):
            ws = []
            for be, (r0, r1) in zip(bes, ranges):
                be.assign(r0, r1)
                ws.append(be.partial_words(r0, r1, "i64"))
            assert ws[0].dtype == torch.int64
            total = ws[0] + ws[1]                        # the all-reduce: exact
            for be in bes:
                be.update_words(total.clone(), "i64", local_rows_only=True)
        st = [be.flags_wait(be.flags_async()) for be in bes]
        assert st[0] == st[1]
        assert st[0][1] == 0
        if st[0][0] == 0:
            state = st[0]
            break
    assert state is not None and state[2] == km.info_["max_iter_run"]
    inertia = sum(be.inertia(r0, r1) for be, (r0, r1) in zip(bes, ranges))
    same = torch.minimum(*[be.same_matrix(r0, r1) for be, (r0, r1) in zip(bes, ranges)])
    best = D.pick_best(inertia.float().cpu().numpy(), same.cpu().numpy())
    assert best == km.info_["best_run"]
    centers = [be.finish(best) for be in bes]
    assert torch.equal(centers[0], centers[1]) and torch.equal(centers[0], km.cluster_centers_)
    labels = torch.cat([be.predict(X[r0:r1].contiguous(), centers[0]) for be, (r0, r1) in zip(bes, ranges)])
    for be in bes:
        be.release()
    assert torch.equal(labels, want_labels) and np.array_equal(labels.cpu().numpy(), g["labels"].reshape(-1))


def test_sharded_segmenter_on_one_rank_equals_clip_segmenter(cuda):
    import torch.distributed as dist
    from oracle import unet as ounet
    from synth import synthetic_unet_inputs, synthetic_unet_weights
    from vidseg_diffusion_b200 import configs
    from vidseg_diffusion_b200.distributed import ShardedClipSegmenter
    from vidseg_diffusion_b200.pipeline import ClipSegmenter
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
    cfg = configs.TINY_UNET
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ounet.param_shapes(cfg), 2).items()}
    model = UNetModel(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    F = 3
    x, t, ctx = (torch.from_numpy(a).to(cuda) for a in synthetic_unet_inputs(2, F, 16, 4, 7, cfg["context_dim"]))
    want, _ = ClipSegmenter(model, num_masks=3, is_aggre_attn=True, is_refine_mask=True).segment(x, t, ctx, F, seed=2)
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda)
        created = True
    try:
        got = ShardedClipSegmenter(model, num_masks=3, is_aggre_attn=True, is_refine_mask=True).segment(x, t, ctx, F, seed=2)
        sh_graph = ShardedClipSegmenter(model, num_masks=3, is_aggre_attn=True, is_refine_mask=True, use_cuda_graph=True)
        got_graph = sh_graph.segment(x, t, ctx, F, seed=2)
        got_many = list(sh_graph.segment_many([(x, t, ctx)] * 3, F, seed=2, to_host=True))
    finally:
        if created:
            dist.destroy_process_group()
    assert torch.equal(got, want) and torch.equal(got_graph, want)
    assert len(got_many) == 3 and all(torch.equal(g, want.cpu()) for g in got_many)


@pytest.mark.parametrize("case,world", [("c1_objects", 4), ("c2_objects", 8), ("c1_iid", 3), ("small_odd", 16)])
def test_run_sharded_fit_is_bit_identical_to_one_gpu(cuda, case, world):
    """The n_init initialisations spread over `world` virtual ranks (each iterating its own runs on all rows, then the
    run records summed as the all-reduce does): winner, centres and labels of the single-GPU fit, bit for bit, and the
    reference's labels (golden)."""
    from vidseg_diffusion_b200 import distributed as D
    from vidseg_diffusion_b200.features import aggregate_normalize
    from vidseg_diffusion_b200.kmeans import KMeans, draw_kmeanspp_randoms
    name, seed, F, h, w, C, K, kind = next(c for c in CLUSTER_CASES if c[0] == case)
    g = np.load(os.path.join(GOLDEN, f"cluster_{name}.npz"))
    blocks, _ = synthetic_clip_features(seed, F, h, w, C, K, kind=kind)
    X = aggregate_normalize([torch.from_numpy(b).to(cuda) for b in blocks], F)
    n = X.shape[0]
    np.random.seed(seed)
    km = KMeans(n_clusters=K, n_init=10)
    want_labels = km.fit_predict(X)
    np.random.seed(seed)
    first, rand = draw_kmeanspp_randoms(n, K, 10)
    parts = D.run_partition(10, world)
    assert parts[0][0] == 0 and parts[-1][1] == 10 and sum(b - a for a, b in parts) == 10
    total, iters = None, 0
    for a, b in parts:
        st = {}
        words = D.fit_run_records(X, K, first, rand, a, b, 300, 1e-4, D.CudaLloydBackend, st)
        assert words.dtype == torch.int32 and bool((words[:a] == 0).all()) and bool((words[b:] == 0).all())
        iters = max(iters, st.get("iterations", 0))
        total = words if total is None else total + words      # the all-reduce: every word has exactly one non-zero owner
    assert iters == km.info_["max_iter_run"]
    labels, centers, best = D.select_from_run_records(X, total, K)
    assert best == km.info_["best_run"]
    assert torch.equal(centers, km.cluster_centers_) and torch.equal(labels, want_labels)
    assert np.array_equal(labels.cpu().numpy(), g["labels"].reshape(-1))
