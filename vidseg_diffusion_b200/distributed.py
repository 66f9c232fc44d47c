"""Frame-sharded form of the per-clip path (BASELINE.json configs[3]).

One process per GPU.  The SD-2.1 UNet has no op that mixes batch entries, so rank r runs the UNet on its own
frames only (both classifier-free-guidance halves of each frame) with replicated weights -- no collective in the
UNet stage; that stage (UNet + harvest + aggregate/normalise of the local frames) replays as one CUDA graph per rank.
The path has exactly two exchange steps:

  1. one all-gather of the aggregated, normalised, conditional-half feature rows (F*hw*C fp32 = 36.7 MB for a
     14-frame 512x512 clip), so that the k-means++ seeding sees every point in the reference's row order;
  2. ONE all-reduce per Lloyd iteration of the exchange words of all runs -- per-cluster sums | counts | label-change
     counters, [n_init*K*(D+1) + n_init] 8-byte words (1.0 MB at K=20).  The words are the INTEGER fixed-point sums of
     the int8 tensor-core M-step (int64), so the sum over ranks is exact and order independent: every rank performs the
     same M-step on the same integers and the sharded centres are bit-identical to the single-GPU fit -- no broadcast,
     no ulp drift (tiny shards fall back to float64 words).

Convergence is polled without stalling the launch queue: the flags of burst b are copied to pinned memory behind an
event and read while burst b+1 is already queued (finished runs make every kernel return early, so the extra burst is
cheap).  Labels stay sharded and are gathered once at the end (N*4 bytes); the inertia and the same-clustering matrix
of the best-of-n_init rule share one more all-reduce.  k-means++ runs redundantly on every rank (deterministic, same
inputs).  sklearn's relocation of empty clusters needs every label of a run; the sharded M-step skips it and raises a
flag instead, and the fit is then repeated unsharded on every rank (every rank holds the full X after step 1), so the
result is always the reference's.

The SVD VideoUNet does NOT shard by frame (temporal attention and the (3,1,1) convolutions mix frames).  Its two
classifier-free-guidance halves never interact inside the network (guiders.py:78-86 combines them afterwards), so on
two GPUs rank 0 runs the unconditional and rank 1 the conditional half (option 1); the
conditional rank broadcasts its feature rows and both ranks share the Lloyd iterations as above.  More ranks: one clip
per GPU (replicas).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .kmeans import draw_kmeanspp_randoms, n_local_trials


def frame_partition(num_frames, world):
    """Contiguous frame ranges [(begin, end)] per rank, sizes differing by at most one (14 frames on 8 ranks:
    2,2,2,2,2,2,1,1)."""
    base, extra = divmod(num_frames, world)
    out, f = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((f, f + n))
        f += n
    return out


def gather_rows(local, counts, group=None):
    """all-gather of row blocks of unequal length: ``local`` [counts[rank], ...] -> [sum(counts), ...] in rank order.
    One collective on a buffer padded to the longest block (NCCL all-gather needs equal sizes)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    mx = max(counts)
    tail = tuple(local.shape[1:])
    pad = torch.zeros((mx, *tail), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world, mx, *tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(world * mx, *tail), pad, group=group)
    return torch.cat([out[r, : counts[r]] for r in range(world)], 0)


class CudaLloydBackend:
    """The split E-step / M-step entry points of libvidseg_b200 (include/vidseg_b200.h, R2) on one GPU."""

    def __init__(self, n_clusters, n_init, max_iter, tol):
        self.k, self.r, self.max_iter, self.tol = n_clusters, n_init, max_iter, tol
        self.lib = _lib.load()
        self.ws = None

    def _call(self, fn, *args):
        with torch.cuda.device(self.dev):
            _lib.check(getattr(self.lib, fn)(*args), fn)

    def prepare(self, X):
        X = _lib.require_cuda_tensor(X, torch.float32, "X")
        self.release()
        self.X, self.dev = X, X.device
        self.n, self.d = X.shape
        self.t = n_local_trials(self.k)
        self.nbytes = self.lib.vidseg_kmeans_workspace_bytes(self.n, self.d, self.k, self.r, self.t)
        if self.nbytes == 0:
            raise _lib.VidsegError(f"k-means: unsupported shape n={self.n} d={self.d} k={self.k} n_init={self.r}")
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev)
        self._call("vidseg_kmeans_prepare", X.data_ptr(), self.n, self.d, self.k, self.r, self.t, self.tol, self.max_iter,
                   self.ws.data_ptr(), self.nbytes, _lib.stream_ptr())

    def seed(self, first, rand):
        self._first = torch.from_numpy(first).to(self.dev)
        self._rand = torch.from_numpy(np.ascontiguousarray(rand)).to(self.dev)
        self._call("vidseg_kmeans_seed", self._first.data_ptr(), self._rand.data_ptr() if self._rand.numel() else None,
                   self.ws.data_ptr(), self.nbytes, _lib.stream_ptr())

    def assign(self, r0, r1):
        self._call("vidseg_kmeans_assign", self.ws.data_ptr(), self.nbytes, r0, r1, _lib.stream_ptr())

    def partial(self, r0, r1):
        partial = torch.empty((self.r, self.k, self.d + 1), dtype=torch.float64, device=self.dev)
        changed = torch.empty(self.r, dtype=torch.int32, device=self.dev)
        self._call("vidseg_kmeans_partial", self.ws.data_ptr(), self.nbytes, r0, r1, partial.data_ptr(), changed.data_ptr(),
                   _lib.stream_ptr())
        return partial, changed

    def update(self, partial, changed, local_rows_only):
        self._call("vidseg_kmeans_update", self.ws.data_ptr(), self.nbytes, partial.data_ptr(), changed.data_ptr(),
                   1 if local_rows_only else 0, _lib.stream_ptr())

    # ---- exchange-word form: one fused array per iteration, non-blocking convergence polling ----
    def exchange_mode(self, ranges):
        """"i64" when every rank's row range can produce the exact integer words, else "f64" (all ranks must agree)."""
        ok = True
        for a, b in ranges:
            m = self.lib.vidseg_kmeans_exchange_mode(self.ws.data_ptr(), self.nbytes, int(a), int(b))
            if m < 0:
                raise _lib.VidsegError(f"kmeans_exchange_mode failed for rows [{a}, {b})")
            ok = ok and m == 1
        return "i64" if ok else "f64"

    def partial_words(self, r0, r1, mode):
        nwords = self.lib.vidseg_kmeans_exchange_words(self.ws.data_ptr(), self.nbytes)
        words = torch.empty(nwords, dtype=torch.int64 if mode == "i64" else torch.float64, device=self.dev)
        self._call("vidseg_kmeans_partial_words", self.ws.data_ptr(), self.nbytes, r0, r1, 1 if mode == "i64" else 0,
                   words.data_ptr(), _lib.stream_ptr())
        return words

    def update_words(self, words, mode, local_rows_only):
        self._call("vidseg_kmeans_update_words", self.ws.data_ptr(), self.nbytes, 1 if mode == "i64" else 0, words.data_ptr(),
                   1 if local_rows_only else 0, _lib.stream_ptr())

    def flags_async(self):
        """Enqueue a copy of the convergence flags to pinned memory; returns a ticket for ``flags_wait``."""
        host = torch.empty((self.r, 4), dtype=torch.int32, pin_memory=True)
        self._call("vidseg_kmeans_flags_async", self.ws.data_ptr(), self.nbytes, host.data_ptr(), _lib.stream_ptr())
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        return host, ev

    def flags_wait(self, ticket):
        """(runs still iterating, runs that met an empty cluster, max n_iter) of the ticket's moment."""
        host, ev = ticket
        ev.synchronize()
        f = host.numpy()
        return int((f[:, 0] == 0).sum()), int((f[:, 3] != 0).sum()), int(f[:, 2].max())

    def status(self):
        active, empty = ctypes.c_int(), ctypes.c_int()
        self._call("vidseg_kmeans_status", self.ws.data_ptr(), self.nbytes, ctypes.byref(active), ctypes.byref(empty),
                   _lib.stream_ptr())
        return active.value, empty.value

    def inertia(self, r0, r1):
        out = torch.empty(self.r, dtype=torch.float64, device=self.dev)
        self._call("vidseg_kmeans_inertia", self.ws.data_ptr(), self.nbytes, r0, r1, out.data_ptr(), _lib.stream_ptr())
        return out

    def same_matrix(self, r0, r1):
        out = torch.empty((self.r, self.r), dtype=torch.int32, device=self.dev)
        self._call("vidseg_kmeans_same_matrix", self.ws.data_ptr(), self.nbytes, r0, r1, out.data_ptr(), _lib.stream_ptr())
        return out

    def lloyd(self, iterations):
        """``iterations`` Lloyd iterations over all rows, queued without synchronising (single-GPU form)."""
        self._call("vidseg_kmeans_lloyd", self.ws.data_ptr(), self.nbytes, int(iterations), _lib.stream_ptr())

    def finish(self, best, want_labels=False):
        centers = torch.empty((self.k, self.d), dtype=torch.float32, device=self.dev)
        labels = torch.empty(self.n, dtype=torch.int32, device=self.dev) if want_labels else None
        self._call("vidseg_kmeans_finish", self.ws.data_ptr(), self.nbytes, int(best), centers.data_ptr(),
                   labels.data_ptr() if want_labels else None, _lib.stream_ptr())
        return (centers, labels) if want_labels else centers

    def predict(self, rows, centers):
        n = rows.shape[0]
        labels = torch.empty(n, dtype=torch.int32, device=self.dev)
        scratch = torch.empty(self.k, dtype=torch.float64, device=self.dev)
        self._call("vidseg_kmeans_predict", rows.data_ptr(), n, self.d, centers.data_ptr(), self.k, labels.data_ptr(),
                   scratch.data_ptr(), _lib.stream_ptr())
        return labels

    def release(self):
        if self.ws is not None:
            self.lib.vidseg_kmeans_release(self.ws.data_ptr())
            self.ws = None


def pick_best(inertia32, same):
    """sklearn/_kmeans.py:1529-1541: the first run wins ties; a lower inertia only counts if the clustering differs."""
    best = 0
    for i in range(1, len(inertia32)):
        if inertia32[i] < inertia32[best] and not same[i][best]:
            best = i
    return best


BURST = 8   # Lloyd iterations between two looks at the convergence flags


def sharded_kmeans_fit_predict(X, n_clusters, row_range, group=None, n_init=10, max_iter=300, tol=1e-4,
                               random_state=None, backend=None, info=None, row_ranges=None):
    """``KMeans(n_clusters, n_init).fit(X).predict(X)`` with the rows of X sharded over the ranks of ``group``.

    X: the FULL matrix [N, D] (identical on every rank, i.e. after the feature all-gather); row_range: this rank's
    (begin, end); row_ranges: every rank's (begin, end) in rank order when the caller knows them (saves one collective).
    Every rank must have numpy's global RandomState in the same state (the pipelines seed it with the clip seed).
    Returns the labels of ALL rows (int32 [N], identical on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = X.shape[0]
    if n < n_clusters:
        raise ValueError(f"n_samples={n} should be >= n_clusters={n_clusters}.")
    r0, r1 = row_range
    be = backend if backend is not None else CudaLloydBackend(n_clusters, n_init, max_iter, tol)
    first, rand = draw_kmeanspp_randoms(n, n_clusters, n_init, random_state)
    stats = {"iterations": 0, "allreduces": 0, "unsharded_fallback": False, "exchange": None}
    sharded = world > 1
    if sharded and row_ranges is None:
        mine = torch.tensor([r0, r1], dtype=torch.int64, device=X.device)
        allr = torch.empty(2 * world, dtype=torch.int64, device=X.device)
        dist.all_gather_into_tensor(allr, mine, group=group)
        row_ranges = [(int(a), int(b)) for a, b in allr.cpu().view(world, 2)]

    def reduce_(t, op=dist.ReduceOp.SUM):
        if world > 1:
            dist.all_reduce(t, op=op, group=group)
            stats["allreduces"] += 1
        return t

    def run(lo, hi, shard):
        be.prepare(X)
        be.seed(first, rand)
        mode = be.exchange_mode(row_ranges) if shard else "f64"
        stats["exchange"] = mode
        it, pending, state = 0, None, None
        while it < max_iter:
            burst = min(BURST, max_iter - it)
            for _ in range(burst):
                be.assign(lo, hi)
                words = be.partial_words(lo, hi, mode)
                if shard:
                    reduce_(words)                                  # exchange step 2: the only collective of an iteration
                be.update_words(words, mode, local_rows_only=shard)
            it += burst
            ticket = be.flags_async()
            if pending is not None:          # flags of the PREVIOUS burst: this one is already queued behind them
                state = be.flags_wait(pending)
                if state[1] or state[0] == 0:
                    break
            pending = ticket
        else:
            state = be.flags_wait(ticket)    # max_iter reached without an early exit: the last burst's own flags
        if state[1]:
            return None
        stats["iterations"] = state[2]
        stats["iterations_issued"] = it
        inertia = be.inertia(lo, hi)
        same = be.same_matrix(lo, hi)
        if shard:   # one collective for both: a pair of runs is the same clustering iff no rank saw a violation
            r = inertia.numel()
            buf = torch.cat([inertia.double().reshape(-1), (1 - same).double().reshape(-1)])
            reduce_(buf)
            inertia, same = buf[:r], (buf[r:] == 0).reshape(r, r)
        best = pick_best(inertia.float().cpu().numpy(), same.cpu().numpy())
        return be.finish(best)

    try:
        centers = run(r0, r1, sharded)
        if centers is None:   # an empty cluster needed sklearn's relocation: repeat the fit unsharded, on every rank
            stats["unsharded_fallback"] = True
            centers = run(0, n, False)
            if centers is None:
                raise _lib.VidsegError("k-means: unsharded fit reported an unrelocated empty cluster")
        local = be.predict(X[r0:r1].contiguous(), centers)
        labels = gather_rows(local, [b - a for a, b in row_ranges], group) if sharded else local
    finally:
        if hasattr(be, "release"):
            be.release()
    if info is not None:
        info.update(stats, centers=centers)
    return labels


def run_partition(n_init, world):
    """Contiguous ranges of the n_init initialisations per rank, sizes differing by at most one (10 runs on 4 ranks:
    3,3,2,2; on 8 ranks: 2,2,1,1,1,1,1,1; ranks beyond n_init get none)."""
    return frame_partition(n_init, world)


def _is_same_clustering(l1, l2, k):
    """sklearn/cluster/_k_means_common.pyx:_is_same_clustering: labels1 -> labels2 is a function (device tensors)."""
    pairs = torch.unique(l1.long() * k + l2.long())
    return bool(pairs.numel() == torch.unique(l1).numel())


def run_sharded_kmeans_fit_predict(X, n_clusters, group=None, n_init=10, max_iter=300, tol=1e-4, random_state=None,
                                   info=None, backend_factory=None):
    """``KMeans(n_clusters, n_init).fit(X).predict(X)`` with the n_init INITIALISATIONS spread over the ranks of ``group``.

    The runs of ``KMeans(n_init=10)`` are independent until the best-of-n_init rule, so rank r seeds and iterates its own
    subset of the runs on ALL rows of X (the FULL matrix, identical on every rank after the feature all-gather) with the
    single-GPU kernels -- no collective inside the Lloyd loop, empty-cluster relocation included -- and ONE all-reduce at
    the end hands every rank the inertia, the fit labels and the centres of every run (each rank contributes its own
    runs' words and zeros elsewhere, so the integer sum is a bit-exact gather).  Every rank then applies sklearn's
    best-of-n_init rule and predicts all rows.  Every run is computed exactly as in the single-GPU fit, so centres and
    labels are bit-identical to it.  All ranks must have numpy's global RandomState in the same state.
    ``backend_factory(k, runs, max_iter, tol)``: stand-in for ``CudaLloydBackend`` (host tests).
    Returns the labels of ALL rows (int32 [N], identical on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    make = backend_factory if backend_factory is not None else CudaLloydBackend
    n, d = X.shape
    if n < n_clusters:
        raise ValueError(f"n_samples={n} should be >= n_clusters={n_clusters}.")
    first, rand = draw_kmeanspp_randoms(n, n_clusters, n_init, random_state)   # same draws, same order, on every rank
    a, b = run_partition(n_init, world)[rank]
    stats = {"allreduces": 0, "exchange": "runs", "runs": (a, b)}
    words = fit_run_records(X, n_clusters, first, rand, a, b, max_iter, tol, make, stats)
    if world > 1:
        dist.all_reduce(words, group=group)                                      # the only collective of the fit
        stats["allreduces"] += 1
    labels, centers, best = select_from_run_records(X, words, n_clusters, make, max_iter, tol)
    if info is not None:
        info.update(stats, centers=centers, best=best)
    return labels


def fit_run_records(X, k, first, rand, a, b, max_iter, tol, make, stats=None):
    """Runs [a, b) of the n_init initialisations on all rows of X -> int32 words [n_init, 2 + N + K*D]: per run the
    inertia (one float64), the fit labels and the un-centred centres; rows of other runs are zero."""
    n, d = X.shape
    n_init = len(first)
    words = torch.zeros((n_init, 2 + n + k * d), dtype=torch.int32, device=X.device)
    if b <= a:
        return words
    be = make(k, b - a, max_iter, tol)
    try:
        be.prepare(X)
        be.seed(first[a:b], rand[a:b])
        it, pending, state = 0, None, None
        while it < max_iter:
            burst = min(BURST, max_iter - it)
            be.lloyd(burst)
            it += burst
            ticket = be.flags_async()
            if pending is not None:          # flags of the PREVIOUS burst: this one is already queued behind them
                state = be.flags_wait(pending)
                if state[0] == 0:
                    break
            pending = ticket
        else:
            state = be.flags_wait(ticket)
        if stats is not None:
            stats["iterations"] = state[2]
            stats["iterations_issued"] = it
        inertia = be.inertia(0, n)                                               # float64 [runs]
        words[a:b, :2] = inertia.contiguous().view(torch.int32).reshape(b - a, 2)
        for j in range(b - a):
            cen, lab = be.finish(j, want_labels=True)
            words[a + j, 2:2 + n] = lab
            words[a + j, 2 + n:] = cen.contiguous().reshape(-1).view(torch.int32)
    finally:
        if hasattr(be, "release"):
            be.release()
    return words


def select_from_run_records(X, words, k, make=None, max_iter=300, tol=1e-4):
    """Best-of-n_init (sklearn/_kmeans.py:1529-1541) over the gathered run records, then KMeans.predict on all rows.
    Returns (labels int32 [N], centres fp32 [K, D], index of the winning run)."""
    n, d = X.shape
    n_init = words.shape[0]
    inertia32 = words[:, :2].contiguous().view(torch.float64).reshape(-1).float().cpu().numpy()
    labels_fit = words[:, 2:2 + n]
    best = 0
    for i in range(1, n_init):   # the clustering test only where the inertia is lower
        if inertia32[i] < inertia32[best] and not _is_same_clustering(labels_fit[i], labels_fit[best], k):
            best = i
    centers = words[best, 2 + n:].contiguous().view(torch.float32).reshape(k, d)
    pred = (make if make is not None else CudaLloydBackend)(k, 1, max_iter, tol)
    pred.dev, pred.d = X.device, d
    return pred.predict(X, centers), centers, best


class ShardedClipSegmenter:
    """``pipeline.ClipSegmenter`` with ONE clip spread over the ranks of a process group: SD-2.1 by frame (any world
    size), SVD by classifier-free-guidance half (two ranks).  Same call, same label maps, on every rank."""

    def __init__(self, model, num_masks=10, is_aggre_attn=False, is_refine_mask=False, n_init=10, group=None,
                 use_cuda_graph=False, kmeans_split="runs"):
        """``kmeans_split``: how the ranks share the K-means fit after the feature all-gather.  "runs" (default): the n_init
        initialisations are spread over the ranks, no collective inside the Lloyd loop, one all-reduce at the end
        (``run_sharded_kmeans_fit_predict``).  "rows": every rank assigns its own rows and ONE all-reduce per Lloyd
        iteration sums the exchange words (``sharded_kmeans_fit_predict``).  Both give the single-GPU labels bit for bit."""
        from .pipeline import ClipSegmenter
        if kmeans_split not in ("runs", "rows"):
            raise _lib.VidsegError(f"kmeans_split must be 'runs' or 'rows', got {kmeans_split!r}")
        self.kmeans_split = kmeans_split
        self.video = "VideoUNet" in str(type(model))
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.video and world > 2:
            raise NotImplementedError("the SVD VideoUNet does not shard by frame (temporal attention / (3,1,1) convolutions "
                                      "mix frames, SURVEY.md section 8e): two ranks split the guidance halves, more ranks "
                                      "run one clip per GPU")
        self.model = model
        self.num_masks = int(num_masks)
        self.is_refine_mask = bool(is_refine_mask)
        self.n_init = n_init
        # the rank-local UNet stage (eager or one CUDA graph per input signature) is the single-GPU segmenter's
        self.local = ClipSegmenter(model, num_masks=num_masks, is_aggre_attn=is_aggre_attn, is_refine_mask=is_refine_mask,
                                   n_init=n_init, use_cuda_graph=use_cuda_graph)
        self.blocks = self.local.blocks
        self.last = {}

    def _local_stage(self, x, t, c, frames, kw, features):
        """UNet on this rank's rows -> (output, feature rows or None, stashed q of the refinement block or None)."""
        from .pipeline import REFINE_BLOCK, harvest_self_attn_q
        if self.local.use_cuda_graph:
            return self.local._graphed_unet_features(x, t, c, frames, kw, features=features)
        out = self.local.unet_step(x, t, c, **kw)
        feats = None if features is None else self.local._features(frames, cond_only=(features == "all_rows"))
        q7 = harvest_self_attn_q(self.model, (REFINE_BLOCK,))[0] if self.is_refine_mask else None
        return out, feats, q7

    @torch.no_grad()
    def segment(self, x, timesteps, context, num_frames, seed=None, **unet_kwargs):
        """x [2F, C, h, w], timesteps [2F], context [2F, L, D]: the WHOLE clip batch on every rank (uncond rows first).
        Returns the label maps of all frames, int32 [F, h/2, w/2], identical on every rank."""
        if self.video:
            return self._segment_cfg_split(x, timesteps, context, num_frames, seed, unet_kwargs)
        return self._cluster_stage(self._unet_stage(x, timesteps, context, num_frames), seed)

    def _unet_stage(self, x, timesteps, context, num_frames):
        """Rank-local part: the UNet on this rank's frames, their feature rows and (for the refinement) stashed q."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        F = num_frames
        parts = frame_partition(F, world)
        f0, f1 = parts[rank]
        fl = f1 - f0
        fh, fw = x.shape[-2] // 2, x.shape[-1] // 2
        hw = fh * fw
        dev = x.device
        q7 = None
        if fl > 0:
            idx = torch.cat([torch.arange(f0, f1, device=dev), torch.arange(F + f0, F + f1, device=dev)])
            _, x_local, q7 = self._local_stage(x[idx], timesteps[idx], context[idx], fl, {}, "cond_half")
            c = x_local.shape[1]
            if q7 is not None:
                q7 = q7[fl:]                                                # conditional half of the local frames
        else:
            c = self.model.output_blocks[self.blocks[0]][1].in_channels
            x_local = torch.empty((0, c), dtype=torch.float32, device=dev)
        if self.is_refine_mask and q7 is None:
            q7 = torch.empty((0, hw, c), dtype=torch.float32, device=dev)
        return dict(x_local=x_local, q7=q7, parts=parts, rank=rank, world=world, F=F, fh=fh, fw=fw, hw=hw)

    def _cluster_stage(self, st, seed):
        """The two exchange steps and everything behind them: feature all-gather, distributed K-means, refinement."""
        from .refine import refine_masks
        parts, rank, world, F, fh, fw, hw = (st[k] for k in ("parts", "rank", "world", "F", "fh", "fw", "hw"))
        counts = [(b - a) * hw for a, b in parts]
        X = gather_rows(st["x_local"], counts, self.group)                 # exchange step 1
        if seed is not None:
            np.random.seed(seed)
        info = {}
        ranges = [(sum(counts[:r]), sum(counts[:r + 1])) for r in range(world)]
        if self.kmeans_split == "runs":
            labels = run_sharded_kmeans_fit_predict(X, self.num_masks, self.group, n_init=self.n_init, info=info)   # exchange step 2 (once)
        else:
            labels = sharded_kmeans_fit_predict(X, self.num_masks, ranges[rank], self.group, n_init=self.n_init, info=info,
                                                row_ranges=ranges)             # exchange step 2 (per iteration)
        labels = labels.reshape(F, fh, fw)
        self.last = {"features": X, "kmeans_info": info}
        if self.is_refine_mask:
            cond = gather_rows(st["q7"].contiguous(), [b - a for a, b in parts], self.group)
            feats7 = torch.cat([torch.zeros_like(cond), cond], 0)   # refine reads rows [F, 2F) only (feature_extraction.py:221)
            labels, traj, keep = refine_masks(feats7, labels, F, fh, fw)
            self.last.update(trajectories=traj, keep=keep)
        return labels

    @torch.no_grad()
    def segment_many(self, clips, num_frames, seed=None, to_host=True):
        """A stream of clips, software-pipelined like ``ClipSegmenter.segment_many``: the rank-local UNet stage of clip i+1
        (one CUDA graph launch on the current stream) overlaps the exchange steps and the distributed K-means of clip i
        (second stream; every collective of the path is issued there, in clip order on every rank).  Same label maps as
        ``segment`` clip by clip."""
        if self.video:
            for clip in clips:
                kw = clip[3] if len(clip) > 3 else {}
                lab = self.segment(clip[0], clip[1], clip[2], num_frames, seed, **kw)
                yield lab.cpu() if to_host else lab
            return
        dev = next(self.model.parameters()).device
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=dev, priority=-1)
        side = self._side_stream
        main = torch.cuda.current_stream(dev)
        to_dev = lambda v: v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v

        def finish(pending):
            st, ready = pending
            with torch.cuda.stream(side):
                side.wait_event(ready)
                labels = self._cluster_stage(st, seed)
                out = labels.cpu() if to_host else labels
                done = torch.cuda.Event()
                done.record(side)
            if not to_host:
                main.wait_event(done)
            return out

        pending = None
        for clip in clips:
            x, t, c = (to_dev(v) for v in clip[:3])
            st = self._unet_stage(x, t, c, num_frames)
            # the graph's buffers are overwritten by the next clip: snapshot what the second stream will read
            for key in ("x_local", "q7"):
                if st[key] is not None:
                    st[key] = st[key].clone()
                    st[key].record_stream(side)
            ready = torch.cuda.Event()
            ready.record(main)
            if pending is not None:
                yield finish(pending)
            pending = (st, ready)
        if pending is not None:
            yield finish(pending)

    def _segment_cfg_split(self, x, timesteps, context, num_frames, seed, unet_kwargs):
        """SVD on two ranks: rank 0 runs the unconditional rows [0, F) of the batch, rank 1 the conditional rows [F, 2F)
        (one video each; the halves only meet in the guider).  Clustering needs the conditional features: rank 1
        broadcasts its [F*hw, C] rows (36.7 MB at 14 x 512 x 512), both ranks share the Lloyd iterations by row range,
        rank 1 refines and broadcasts the label maps."""
        from .refine import refine_masks
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        F = num_frames
        fh, fw = x.shape[-2] // 2, x.shape[-1] // 2
        hw = fh * fw
        dev = x.device
        if world == 1:
            return self.local.segment(x, timesteps, context, F, seed, **unet_kwargs)[0]
        cond_rank = 1
        half = slice(rank * F, (rank + 1) * F)
        kw = {}
        for k, v in unet_kwargs.items():
            if isinstance(v, torch.Tensor) and v.shape[0] == 2 * F:
                kw[k] = v[half]
            elif k == "image_only_indicator" and isinstance(v, torch.Tensor):
                kw[k] = v[rank:rank + 1]
            else:
                kw[k] = v
        out, x_rows, q7 = self._local_stage(x[half], timesteps[half], context[half], F, kw,
                                            "all_rows" if rank == cond_rank else None)
        c = self.model.output_blocks[self.blocks[0]][1].in_channels
        X = x_rows if rank == cond_rank else torch.empty((F * hw, c), dtype=torch.float32, device=dev)
        src = dist.get_global_rank(self.group, cond_rank) if self.group is not None else cond_rank
        dist.broadcast(X, src=src, group=self.group)                       # exchange step 1
        if seed is not None:
            np.random.seed(seed)
        n = F * hw
        cut = (n // 2 + 127) // 128 * 128
        ranges = [(0, min(cut, n)), (min(cut, n), n)]
        info = {}
        if self.kmeans_split == "runs":
            labels = run_sharded_kmeans_fit_predict(X, self.num_masks, self.group, n_init=self.n_init, info=info).reshape(F, fh, fw)
        else:
            labels = sharded_kmeans_fit_predict(X, self.num_masks, ranges[rank], self.group, n_init=self.n_init, info=info,
                                                row_ranges=ranges).reshape(F, fh, fw)
        self.last = {"features": X, "kmeans_info": info, "unet_out_half": out}
        if self.is_refine_mask:
            if rank == cond_rank:
                feats7 = torch.cat([torch.zeros_like(q7), q7], 0)
                labels, traj, keep = refine_masks(feats7, labels, F, fh, fw)
                self.last.update(trajectories=traj, keep=keep)
                labels = labels.contiguous()
            dist.broadcast(labels, src=src, group=self.group)
        return labels
