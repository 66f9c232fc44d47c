"""R2 host mirror: ``sklearn.cluster.KMeans(n_clusters=K, n_init=10)`` fit + predict on B200.

Reference call site: scripts/sampling/feature_extraction.py:52-55.  The class keeps sklearn's
constructor / ``fit`` / ``predict`` / ``cluster_centers_`` / ``inertia_`` / ``n_iter_`` surface so
the mirror of ``save_inidividual_masks_kmeans`` reads like the reference.  Host side: the
data-independent random draws of k-means++ are taken from numpy's RandomState with the very calls
sklearn makes (sklearn/cluster/_kmeans.py:231 ``choice(n, p=w/w.sum())`` and :249
``uniform(size=n_local_trials)``), in the same order, so the global numpy stream advances exactly
as it does under the reference.  Device side: libvidseg_b200 (csrc/kmeans.cu).
"""
import ctypes

import numpy as np
import torch

from . import _lib


def n_local_trials(n_clusters):
    return 2 + int(np.log(n_clusters))  # sklearn/_kmeans.py:228


def draw_kmeanspp_randoms(n_samples, n_clusters, n_init, random_state=None):
    """All random decisions of ``n_init`` k-means++ seedings, drawn in sklearn's order.

    Returns (first_idx int32 [R], rand float64 [R, K-1, T])."""
    rs = np.random.mtrand._rand if random_state is None else random_state
    t = n_local_trials(n_clusters)
    sample_weight = np.ones(n_samples, dtype=np.float32)
    p = sample_weight / sample_weight.sum()
    first = np.empty(n_init, dtype=np.int32)
    rand = np.empty((n_init, max(n_clusters - 1, 0), t), dtype=np.float64)
    for r in range(n_init):
        first[r] = rs.choice(n_samples, p=p)
        for c in range(n_clusters - 1):
            rand[r, c] = rs.uniform(size=t)
    return first, rand


class KMeans:
    """B200 K-means with scikit-learn's ``KMeans`` semantics (lloyd, k-means++, float32)."""

    def __init__(self, n_clusters=8, *, n_init=10, max_iter=300, tol=1e-4, random_state=None):
        self.n_clusters = int(n_clusters)
        self.n_init = int(n_init)
        self.max_iter = int(max_iter)
        self.tol = float(tol)
        if isinstance(random_state, (int, np.integer)):
            random_state = np.random.RandomState(random_state)
        self.random_state = random_state
        self.cluster_centers_ = None
        self.labels_ = None  # labels of predict(X_fit), filled by fit_predict
        self.inertia_ = None
        self.n_iter_ = None
        self.info_ = {}

    def _workspace(self, n, d, device):
        lib = _lib.load()
        t = n_local_trials(self.n_clusters)
        nbytes = lib.vidseg_kmeans_workspace_bytes(n, d, self.n_clusters, self.n_init, t)
        if nbytes == 0:
            raise _lib.VidsegError(f"k-means: unsupported shape n={n} d={d} k={self.n_clusters} n_init={self.n_init}")
        return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes, t

    def fit_predict(self, X):
        """``fit(X)`` followed by ``predict(X)`` (what feature_extraction.py:54-55 does)."""
        X = _lib.require_cuda_tensor(X, torch.float32, "X")
        if X.dim() != 2:
            raise _lib.VidsegError("X must be [n_samples, n_features]")
        n, d = X.shape
        if n < self.n_clusters:
            raise ValueError(f"n_samples={n} should be >= n_clusters={self.n_clusters}.")  # sklearn's message
        lib = _lib.load()
        first, rand = draw_kmeanspp_randoms(n, self.n_clusters, self.n_init, self.random_state)
        ws, nbytes, t = self._workspace(n, d, X.device)
        labels = torch.empty(n, dtype=torch.int32, device=X.device)
        centers = torch.empty((self.n_clusters, d), dtype=torch.float32, device=X.device)
        info = np.zeros(4, dtype=np.int32)
        inertia = np.zeros(self.n_init, dtype=np.float32)
        rand = np.ascontiguousarray(rand)
        with torch.cuda.device(X.device):
            code = lib.vidseg_kmeans_fit_predict(
                X.data_ptr(), n, d, self.n_clusters, self.n_init, t, self.max_iter, self.tol,
                first.ctypes.data, rand.ctypes.data, labels.data_ptr(), centers.data_ptr(),
                info.ctypes.data, inertia.ctypes.data, ws.data_ptr(), nbytes, _lib.stream_ptr())
            lib.vidseg_kmeans_release(ws.data_ptr())
        _lib.check(code, "kmeans_fit_predict")
        self.cluster_centers_ = centers
        self.labels_ = labels
        self.inertia_ = float(inertia[info[0]])
        self.n_iter_ = int(info[1])
        self.info_ = {"best_run": int(info[0]), "max_iter_run": int(info[2]), "launches": int(info[3]),
                      "inertia_per_run": inertia.copy()}
        return labels

    def fit(self, X):
        self.fit_predict(X)
        return self

    def predict(self, X):
        if self.cluster_centers_ is None:
            raise _lib.VidsegError("KMeans.predict called before fit")
        X = _lib.require_cuda_tensor(X, torch.float32, "X")
        n, d = X.shape
        lib = _lib.load()
        labels = torch.empty(n, dtype=torch.int32, device=X.device)
        scratch = torch.empty(self.n_clusters, dtype=torch.float64, device=X.device)
        with torch.cuda.device(X.device):
            _lib.check(lib.vidseg_kmeans_predict(X.data_ptr(), n, d, self.cluster_centers_.data_ptr(), self.n_clusters,
                                                 labels.data_ptr(), scratch.data_ptr(), _lib.stream_ptr()), "kmeans_predict")
        return labels
