"""CPU, world_size 2 over gloo: the host logic of the frame-sharded path (vidseg_diffusion_b200/distributed.py) --
frame partition, ragged row all-gather, and the distributed Lloyd driver (ONE all-reduce per iteration of the fused
[n_init*K*(D+1) + n_init] exchange words: sums | counts | change counters; convergence flags polled one burst late;
redundant seeding / best-of selection with inertia and the same-clustering matrix in one collective; final label
all-gather), and the run-sharded form of the same fit (``run_sharded_kmeans_fit_predict``: each rank iterates its own
subset of the n_init initialisations on all rows, ONE all-reduce of the run records, best-of-n_init on every rank).  The
CUDA kernels cannot run here, so the driver is exercised with a numpy stand-in that implements the same split E-step / M-step
contract as include/vidseg_b200.h (R2) on top of the oracle's primitives; the GPU form of the same driver is covered by
tests/test_gpu_distributed.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kmeans as okm
from synth import synthetic_clip_features
from oracle import features as ofeat


class _ReplayRandom:
    """Feeds pre-drawn k-means++ decisions back through the RandomState calls the oracle makes."""

    def __init__(self, first, rand):
        self.first, self.rand, self.c = int(first), rand, 0

    def choice(self, n, p=None):
        return self.first

    def uniform(self, size=None):
        v = self.rand[self.c]
        self.c += 1
        return v


class NumpyLloydBackend:
    """Same contract as distributed.CudaLloydBackend (prepare / seed / assign / partial / update / status / inertia /
    same_matrix / finish / predict) in numpy: float64 cluster sums, fp32 centres, sklearn's convergence rules."""

    def __init__(self, k, r, max_iter=300, tol=1e-4):
        self.k, self.r, self.max_iter, self.tol_rel = k, r, max_iter, tol

    def prepare(self, X):
        X = X.numpy().astype(np.float32)
        self.tol = np.mean(np.var(X, axis=0)) * self.tol_rel
        self.mean = X.mean(axis=0)
        self.xc = X - self.mean
        self.n, self.d = X.shape
        self.labels = np.full((self.r, self.n), -1, dtype=np.int32)
        self.changed = np.zeros(self.r, dtype=np.int32)
        self.done = np.zeros(self.r, dtype=bool)
        self.strict = np.zeros(self.r, dtype=bool)
        self.empty_seen = np.zeros(self.r, dtype=bool)
        self.n_iter = np.zeros(self.r, dtype=np.int32)

    def seed(self, first, rand):
        self.centers = np.stack([okm.kmeans_plusplus(self.xc, self.k, _ReplayRandom(first[r], rand[r]))[0]
                                 for r in range(self.r)])

    def _estep(self, r, r0, r1):
        c = self.centers[r].astype(np.float64)
        x = self.xc[r0:r1].astype(np.float64)
        return np.argmin((c * c).sum(1)[None, :] - 2.0 * x @ c.T, axis=1).astype(np.int32)

    def assign(self, r0, r1):
        for r in range(self.r):
            if self.done[r]:
                continue
            lab = self._estep(r, r0, r1)
            self.changed[r] += int((lab != self.labels[r, r0:r1]).sum())
            self.labels[r, r0:r1] = lab

    def partial(self, r0, r1):
        part = np.zeros((self.r, self.k, self.d + 1))
        for r in range(self.r):
            if self.done[r]:
                continue
            lab = self.labels[r, r0:r1]
            np.add.at(part[r, :, : self.d], lab, self.xc[r0:r1].astype(np.float64))
            part[r, :, self.d] = np.bincount(lab, minlength=self.k)
        ch = self.changed.copy()
        return torch.from_numpy(part), torch.from_numpy(ch)

    def update(self, partial, changed, local_rows_only):
        part, ch = partial.numpy(), changed.numpy()
        for r in range(self.r):
            if self.done[r]:
                continue
            cnt = part[r, :, self.d]
            if (cnt == 0).any():
                assert local_rows_only, "stand-in has no relocation"
                self.empty_seen[r] = True
            new = self.centers[r].copy()
            am = int(np.argmax(cnt))
            for j in range(self.k):
                new[j] = (part[r, j, : self.d] / cnt[j]).astype(np.float32) if cnt[j] > 0 else new[am]
            shift = np.sqrt(((new.astype(np.float64) - self.centers[r]) ** 2).sum(1)).astype(np.float32)
            self.centers[r] = new
            self.n_iter[r] += 1
            if ch[r] == 0:
                self.strict[r] = self.done[r] = True
            elif np.float32((shift * shift).sum()) <= self.tol or self.n_iter[r] >= self.max_iter:
                self.done[r] = True
            self.changed[r] = 0

    def status(self):
        return int((~self.done).sum()), int(self.empty_seen.sum())

    # exchange-word form of the same contract (one fused float64 array per iteration; this stand-in has no int8 M-step)
    def exchange_mode(self, ranges):
        return "f64"

    def partial_words(self, r0, r1, mode):
        assert mode == "f64"
        part, ch = self.partial(r0, r1)
        return torch.cat([part.reshape(-1), ch.double()])

    def update_words(self, words, mode, local_rows_only):
        nk = self.r * self.k * (self.d + 1)
        self.update(words[:nk].reshape(self.r, self.k, self.d + 1), words[nk:].round().to(torch.int32), local_rows_only)

    def flags_async(self):
        return self.status() + (int(self.n_iter.max()),)

    def flags_wait(self, ticket):
        return ticket

    def inertia(self, r0, r1):
        out = np.zeros(self.r)
        for r in range(self.r):
            if not self.strict[r]:
                self.labels[r, r0:r1] = self._estep(r, r0, r1)
            diff = self.xc[r0:r1].astype(np.float64) - self.centers[r][self.labels[r, r0:r1]]
            out[r] = (diff * diff).sum()
        return torch.from_numpy(out)

    def same_matrix(self, r0, r1):
        out = np.zeros((self.r, self.r), dtype=np.int32)
        for a in range(self.r):
            for b in range(self.r):
                out[a, b] = okm.is_same_clustering(self.labels[a, r0:r1], self.labels[b, r0:r1], self.k)
        return torch.from_numpy(out)

    def finish(self, best, want_labels=False):
        cen = torch.from_numpy((self.centers[best] + self.mean).astype(np.float32))
        return (cen, torch.from_numpy(self.labels[best].copy())) if want_labels else cen

    def lloyd(self, iterations):
        for _ in range(iterations):
            self.assign(0, self.n)
            part, ch = self.partial(0, self.n)
            self.update(part, ch, local_rows_only=False)

    def predict(self, rows, centers):
        return torch.from_numpy(okm.kmeans_predict(rows.numpy(), centers.numpy()).astype(np.int32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


rank_known_ranges = True


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vidseg_diffusion_b200 import distributed as D
        F, h, w, C, K = 5, 8, 8, 32, 4
        blocks, _ = synthetic_clip_features(11, F, h, w, C, K)
        parts = D.frame_partition(F, world)
        f0, f1 = parts[rank]
        # the rows this rank would produce from its own frames (conditional half), then exchange step 1
        x_full = ofeat.aggregate_normalize(blocks, F)
        local = torch.from_numpy(x_full[f0 * h * w: f1 * h * w])
        counts = [(b - a) * h * w for a, b in parts]
        X = D.gather_rows(local, counts)
        assert np.array_equal(X.numpy(), x_full)
        np.random.seed(7)
        info = {}
        row0 = sum(counts[:rank])
        ranges = [(sum(counts[:r]), sum(counts[:r + 1])) for r in range(world)]
        labels = D.sharded_kmeans_fit_predict(X, K, ranges[rank], n_init=3, backend=NumpyLloydBackend(K, 3), info=info,
                                              row_ranges=ranges if rank_known_ranges else None)
        q.put((rank, labels.numpy(), info["iterations_issued"], info["allreduces"], info["unsharded_fallback"]))
        # the same fit with the INITIALISATIONS spread over the ranks (3 runs on 2 ranks: 2 + 1)
        np.random.seed(7)
        info2 = {}
        labels2 = D.run_sharded_kmeans_fit_predict(X, K, n_init=3, backend_factory=NumpyLloydBackend, info=info2)
        q.put((rank, labels2.numpy(), info2["runs"], info2["allreduces"], info2["best"]))
    finally:
        dist.destroy_process_group()


def test_frame_partition():
    from vidseg_diffusion_b200.distributed import frame_partition
    assert [b - a for a, b in frame_partition(14, 8)] == [2, 2, 2, 2, 2, 2, 1, 1]
    assert frame_partition(14, 1) == [(0, 14)]
    assert frame_partition(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for F in (1, 7, 14, 28):
        for g in (1, 2, 4, 8):
            p = frame_partition(F, g)
            assert p[0][0] == 0 and p[-1][1] == F and all(a[1] == b[0] for a, b in zip(p, p[1:]))


def test_sharded_kmeans_two_ranks_gloo_matches_oracle():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = sorted((t for t in got if isinstance(t[4], bool)), key=lambda t: t[0])
    res_runs = sorted((t for t in got if not isinstance(t[4], bool)), key=lambda t: t[0])
    (_, lab0, it0, ar0, fb0), (_, lab1, it1, ar1, fb1) = res
    assert np.array_equal(lab0, lab1) and it0 == it1 and ar0 == ar1 > 0 and not fb0 and not fb1
    # the reference's answer on the same rows, same seed
    F, h, w, C, K = 5, 8, 8, 32, 4
    blocks, _ = synthetic_clip_features(11, F, h, w, C, K)
    x_full = ofeat.aggregate_normalize(blocks, F)
    np.random.seed(7)
    want, _ = okm.kmeans_fit_predict(x_full, K, n_init=3)
    assert np.array_equal(lab0, want)
    # ONE all-reduce per issued Lloyd iteration (sums | counts | change counters) + one for inertia | same-clustering
    assert ar0 == it0 + 1
    # run-sharded form: rank 0 iterated runs [0, 2), rank 1 run [2, 3); ONE collective; same labels, same winner everywhere
    (_, rl0, runs0, rar0, best0), (_, rl1, runs1, rar1, best1) = res_runs
    assert tuple(runs0) == (0, 2) and tuple(runs1) == (2, 3) and rar0 == rar1 == 1 and best0 == best1
    assert np.array_equal(rl0, want) and np.array_equal(rl1, want)


def _worker_runs(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vidseg_diffusion_b200 import distributed as D
        F, h, w, C, K = 5, 8, 8, 32, 4
        blocks, _ = synthetic_clip_features(11, F, h, w, C, K)
        X = torch.from_numpy(ofeat.aggregate_normalize(blocks, F))
        np.random.seed(7)
        info = {}
        labels = D.run_sharded_kmeans_fit_predict(X, K, n_init=3, backend_factory=NumpyLloydBackend, info=info)
        q.put((rank, labels.numpy(), tuple(info["runs"]), info["allreduces"], info["best"]))
    finally:
        dist.destroy_process_group()


def test_run_sharded_kmeans_four_ranks_gloo_with_an_idle_rank():
    """3 initialisations on 4 ranks (1, 1, 1, 0): the rank without a run contributes zeros to the one all-reduce and still
    applies the best-of-n_init rule and predicts; every rank returns the reference's labels."""
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_runs, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    F, h, w, C, K = 5, 8, 8, 32, 4
    blocks, _ = synthetic_clip_features(11, F, h, w, C, K)
    np.random.seed(7)
    want, _ = okm.kmeans_fit_predict(ofeat.aggregate_normalize(blocks, F), K, n_init=3)
    assert [t[2] for t in res] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert len({t[4] for t in res}) == 1 and all(t[3] == 1 for t in res)
    for t in res:
        assert np.array_equal(t[1], want)
