"""Frame-sharded form of the per-clip path (BASELINE.json configs[3]).

One process per GPU.  The SD-2.1 UNet has no op that mixes batch entries, so rank r runs the UNet on its own
frames only (both classifier-free-guidance halves of each frame) with replicated weights -- no collective in the
UNet stage.  The path has exactly two exchange steps:

  1. one all-gather of the aggregated, normalised, conditional-half feature rows (F*hw*C fp32 = 36.7 MB for a
     14-frame 512x512 clip), so that the k-means++ seeding sees every point in the reference's row order;
  2. one all-reduce per Lloyd iteration of the fused [n_init, K, D+1] float64 buffer (per-cluster sums | counts,
     1.0 MB at K=20) plus the [n_init] label-change counters; every rank then performs the same M-step, so the
     centres stay bit-identical across ranks without broadcasting them.

Labels stay sharded and are gathered once at the end (N*4 bytes).  k-means++ and the best-of-n_init selection run
redundantly on every rank (deterministic, same inputs), costing no communication.  sklearn's relocation of empty
clusters needs every label of a run; the sharded M-step skips it and raises a flag instead, and the fit is then
repeated unsharded on every rank (every rank holds the full X after step 1), so the result is always the
reference's.  Cross-rank summation order differs from the single-GPU order, so centres may differ by an ulp; label
maps are what is compared.

The SVD VideoUNet does NOT shard by frame (temporal attention and the (3,1,1) convolutions mix frames); its
multi-GPU form is one clip per GPU (replicas).
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

    def finish(self, best):
        centers = torch.empty((self.k, self.d), dtype=torch.float32, device=self.dev)
        self._call("vidseg_kmeans_finish", self.ws.data_ptr(), self.nbytes, int(best), centers.data_ptr(), None,
                   _lib.stream_ptr())
        return centers

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


def sharded_kmeans_fit_predict(X, n_clusters, row_range, group=None, n_init=10, max_iter=300, tol=1e-4,
                               random_state=None, backend=None, info=None):
    """``KMeans(n_clusters, n_init).fit(X).predict(X)`` with the rows of X sharded over the ranks of ``group``.

    X: the FULL matrix [N, D] (identical on every rank, i.e. after the feature all-gather); row_range: this rank's
    (begin, end).  Every rank must have numpy's global RandomState in the same state (the pipelines seed it with the
    clip seed).  Returns the labels of ALL rows (int32 [N], identical on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = X.shape[0]
    if n < n_clusters:
        raise ValueError(f"n_samples={n} should be >= n_clusters={n_clusters}.")
    r0, r1 = row_range
    be = backend if backend is not None else CudaLloydBackend(n_clusters, n_init, max_iter, tol)
    first, rand = draw_kmeanspp_randoms(n, n_clusters, n_init, random_state)
    stats = {"iterations": 0, "allreduces": 0, "unsharded_fallback": False}

    def reduce_(t, op=dist.ReduceOp.SUM):
        if world > 1:
            dist.all_reduce(t, op=op, group=group)
            stats["allreduces"] += 1
        return t

    def run(lo, hi, sharded):
        be.prepare(X)
        be.seed(first, rand)
        it, poll = 0, 4
        while it < max_iter:
            burst = min(poll, max_iter - it)
            for _ in range(burst):
                be.assign(lo, hi)
                partial, changed = be.partial(lo, hi)
                if sharded:
                    reduce_(partial)
                    reduce_(changed)
                be.update(partial, changed, local_rows_only=sharded)
            it += burst
            active, empty = be.status()
            if empty:
                return None
            if active == 0:
                break
            poll = min(poll * 2, 16)
        stats["iterations"] = it
        inertia = be.inertia(lo, hi)
        same = be.same_matrix(lo, hi)
        if sharded:
            reduce_(inertia)
            reduce_(same, dist.ReduceOp.MIN)
        best = pick_best(inertia.float().cpu().numpy(), same.cpu().numpy())
        return be.finish(best)

    try:
        sharded = world > 1
        centers = run(r0, r1, sharded)
        if centers is None:   # an empty cluster needed sklearn's relocation: repeat the fit unsharded, on every rank
            stats["unsharded_fallback"] = True
            centers = run(0, n, False)
            if centers is None:
                raise _lib.VidsegError("k-means: unsharded fit reported an unrelocated empty cluster")
        counts = None
        if sharded:
            cnt = torch.tensor([r1 - r0], dtype=torch.int64, device=X.device)
            allc = torch.empty(world, dtype=torch.int64, device=X.device)
            dist.all_gather_into_tensor(allc, cnt, group=group)
            counts = [int(v) for v in allc.cpu()]
        local = be.predict(X[r0:r1].contiguous(), centers)
        labels = gather_rows(local, counts, group) if sharded else local
    finally:
        if hasattr(be, "release"):
            be.release()
    if info is not None:
        info.update(stats, centers=centers)
    return labels


class ShardedClipSegmenter:
    """``pipeline.ClipSegmenter`` with the frames of ONE clip sharded over the ranks of a process group."""

    def __init__(self, model, num_masks=10, is_aggre_attn=False, is_refine_mask=False, n_init=10, group=None):
        from .pipeline import AGGRE_BLOCKS, SINGLE_BLOCK
        if "VideoUNet" in str(type(model)):
            raise NotImplementedError("the SVD VideoUNet does not shard by frame (temporal attention / (3,1,1) convolutions "
                                      "mix frames, SURVEY.md section 8e): run one clip per GPU instead")
        self.model = model
        self.num_masks = int(num_masks)
        self.blocks = AGGRE_BLOCKS if is_aggre_attn else SINGLE_BLOCK
        self.is_refine_mask = bool(is_refine_mask)
        self.n_init = n_init
        self.group = group
        self.last = {}

    @torch.no_grad()
    def segment(self, x, timesteps, context, num_frames, seed=None):
        """x [2F, C, h, w], timesteps [2F], context [2F, L, D]: the WHOLE clip batch on every rank (uncond rows first).
        Returns the label maps of all frames, int32 [F, h/2, w/2], identical on every rank."""
        from .features import aggregate_normalize
        from .pipeline import REFINE_BLOCK, harvest_self_attn_q
        from .refine import refine_masks
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        F = num_frames
        parts = frame_partition(F, world)
        f0, f1 = parts[rank]
        fl = f1 - f0
        fh, fw = x.shape[-2] // 2, x.shape[-1] // 2
        hw = fh * fw
        counts = [(b - a) * hw for a, b in parts]
        dev = x.device
        if fl > 0:
            idx = torch.cat([torch.arange(f0, f1, device=dev), torch.arange(F + f0, F + f1, device=dev)])
            self.model(x[idx], timesteps=timesteps[idx], context=context[idx])
            feats = harvest_self_attn_q(self.model, self.blocks)
            x_local = aggregate_normalize(feats, fl)
            c = x_local.shape[1]
        else:
            c = self.model.output_blocks[self.blocks[0]][1].in_channels
            x_local = torch.empty((0, c), dtype=torch.float32, device=dev)
        X = gather_rows(x_local, counts, self.group)                       # exchange step 1
        if seed is not None:
            np.random.seed(seed)
        info = {}
        row0 = sum(counts[:rank])
        labels = sharded_kmeans_fit_predict(X, self.num_masks, (row0, row0 + counts[rank]), self.group, n_init=self.n_init,
                                            info=info)                     # exchange step 2 (per iteration)
        labels = labels.reshape(F, fh, fw)
        self.last = {"features": X, "kmeans_info": info}
        if self.is_refine_mask:
            if fl > 0:
                q7 = harvest_self_attn_q(self.model, (REFINE_BLOCK,))[0][fl:]          # conditional half of the local frames
            else:
                q7 = torch.empty((0, hw, c), dtype=torch.float32, device=dev)
            cond = gather_rows(q7.contiguous(), [b - a for a, b in parts], self.group)
            feats7 = torch.cat([torch.zeros_like(cond), cond], 0)   # refine reads rows [F, 2F) only (feature_extraction.py:221)
            labels, traj, keep = refine_masks(feats7, labels, F, fh, fw)
            self.last.update(trajectories=traj, keep=keep)
        return labels
