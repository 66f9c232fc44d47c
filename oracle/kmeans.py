"""Oracle for R2: ``sklearn.cluster.KMeans(n_clusters=K, n_init=10).fit(X).predict(X)``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Call site in the reference: scripts/sampling/feature_extraction.py:52-55 (no
``random_state`` -> numpy's *global* RandomState, seeded by ``seed_everything``
svd_single_video_inference.py:590-594).  The arithmetic lives in scikit-learn
(pinned scikit_learn==1.5.0 in requirements/pt2.txt:18; 1.9.0 is installed in
this image; the code restated here is identical in both):

  sklearn/cluster/_kmeans.py
    KMeans.fit              :1464-1556  (mean-centring :1488-1490, n_init loop,
                                         best-of by inertia and not-same-clustering)
    _tolerance              :285-294    tol = mean(var(X, axis=0)) * 1e-4
    _kmeans_plusplus        :180-277    first centre ``choice(n, p=w/sum w)``,
                                         then ``uniform(size=trials) * pot`` +
                                         ``searchsorted(cumsum(w * d2))``
    _kmeans_single_lloyd    :630-758
    KMeans.predict          :1075-1107  one more E-step on the *un-centred* X
  sklearn/cluster/_k_means_lloyd.pyx   lloyd_iter_chunked_dense (256-row chunks,
                                         ||c||^2 - 2 x.c via BLAS gemm, strict '<' argmin)
  sklearn/cluster/_k_means_common.pyx  _inertia_dense, _relocate_empty_clusters_dense,
                                         _average_centers, _center_shift, _is_same_clustering
  sklearn/metrics/pairwise.py          _euclidean_distances :376-427 ->
                                         _euclidean_distances_upcast (float64 chunks,
                                         cast to float32, clamp at 0)

Everything is restated with the same numpy primitives sklearn itself calls
(np.cumsum, np.searchsorted, ``@``), so on one machine the restatement and
sklearn agree bit-for-bit except where sklearn goes through its Cython/OpenMP
reductions (centre sums, inertia) -- those are reproduced in the documented
order for one thread and agree to rounding.
"""
import numpy as np

CHUNK_SIZE = 256  # _k_means_common.pyx


def euclidean_sq_upcast(Y, X):
    """_euclidean_distances(Y, X, squared=True) for float32 inputs.

    pairwise.py:403-406 -> _euclidean_distances_upcast: float64 expansion
    ||y||^2 + ||x||^2 - 2 y.x on upcast data, stored as float32, clamped at 0.
    """
    Y64 = Y.astype(np.float64)
    X64 = X.astype(np.float64)
    yy = np.einsum("ij,ij->i", Y64, Y64)[:, None]
    xx = np.einsum("ij,ij->i", X64, X64)[None, :]
    d = -2.0 * (Y64 @ X64.T)
    d += yy
    d += xx
    d32 = d.astype(np.float32)
    np.maximum(d32, 0, out=d32)
    return d32


def kmeans_plusplus(X, n_clusters, random_state, trace=None):
    """_kmeans_plusplus (_kmeans.py:180-277) with unit sample weights."""
    n_samples, n_features = X.shape
    sample_weight = np.ones(n_samples, dtype=X.dtype)
    centers = np.empty((n_clusters, n_features), dtype=X.dtype)
    n_local_trials = 2 + int(np.log(n_clusters))

    center_id = random_state.choice(n_samples, p=sample_weight / sample_weight.sum())
    indices = np.full(n_clusters, -1, dtype=int)
    centers[0] = X[center_id]
    indices[0] = center_id

    closest_dist_sq = euclidean_sq_upcast(centers[0, np.newaxis], X)
    current_pot = closest_dist_sq @ sample_weight

    for c in range(1, n_clusters):
        rand_vals = random_state.uniform(size=n_local_trials) * current_pot
        candidate_ids = np.searchsorted(np.cumsum(sample_weight * closest_dist_sq), rand_vals)
        np.clip(candidate_ids, None, closest_dist_sq.size - 1, out=candidate_ids)
        distance_to_candidates = euclidean_sq_upcast(X[candidate_ids], X)
        np.minimum(closest_dist_sq, distance_to_candidates, out=distance_to_candidates)
        candidates_pot = distance_to_candidates @ sample_weight.reshape(-1, 1)
        best_candidate = np.argmin(candidates_pot)
        current_pot = candidates_pot[best_candidate]
        closest_dist_sq = distance_to_candidates[best_candidate]
        best_candidate = candidate_ids[best_candidate]
        centers[c] = X[best_candidate]
        indices[c] = best_candidate
        if trace is not None:
            trace.append((c, candidate_ids.copy(), int(best_candidate), float(current_pot[0] if np.ndim(current_pot) else current_pot)))
    return centers, indices


def _euclid_dense_dense_sq(a, b):
    """_k_means_common.pyx:_euclidean_dense_dense(squared=True), float32, 4-way groups."""
    d = (a - b).astype(np.float32)
    sq = d * d
    n4 = (sq.shape[-1] // 4) * 4
    res = np.float32(0)
    # grouped (a0+a1+a2+a3) sums, then sequential accumulation -- vectorised over rows
    groups = sq[..., :n4].reshape(*sq.shape[:-1], -1, 4)
    g = ((groups[..., 0] + groups[..., 1]) + groups[..., 2]) + groups[..., 3]
    res = np.zeros(sq.shape[:-1], dtype=np.float32)
    for i in range(g.shape[-1]):
        res = res + g[..., i]
    for i in range(n4, sq.shape[-1]):
        res = res + sq[..., i]
    return res


def lloyd_iter(X, centers_old, update_centers=True):
    """lloyd_iter_chunked_dense + _update_chunk_dense, one thread, unit weights.

    Returns labels, and if update_centers: centers_new, weight_in_clusters, center_shift.
    """
    n_samples, n_features = X.shape
    n_clusters = centers_old.shape[0]
    centers_sq = np.einsum("ij,ij->i", centers_old, centers_old)  # row_norms(squared=True)
    labels = np.empty(n_samples, dtype=np.int32)
    centers_new = np.zeros_like(centers_old)
    weight = np.zeros(n_clusters, dtype=X.dtype)
    for start in range(0, n_samples, CHUNK_SIZE):
        chunk = X[start:start + CHUNK_SIZE]
        pd = np.empty((chunk.shape[0], n_clusters), dtype=X.dtype)
        pd[:] = centers_sq[None, :]
        pd += X.dtype.type(-2.0) * (chunk @ centers_old.T)
        lab = np.argmin(pd, axis=1).astype(np.int32)  # first minimum == strict '<' scan
        labels[start:start + CHUNK_SIZE] = lab
        if update_centers:
            np.add.at(weight, lab, X.dtype.type(1))
            np.add.at(centers_new, lab, chunk)  # sequential in row order, like the C loop
    if not update_centers:
        return labels
    _relocate_empty_clusters(X, centers_old, centers_new, weight, labels)
    # _average_centers
    argmax_w = int(np.argmax(weight))
    for j in range(n_clusters):
        if weight[j] > 0:
            centers_new[j] *= X.dtype.type(1.0) / weight[j]
        else:
            centers_new[j] = centers_new[argmax_w]
    shift = np.sqrt(_euclid_dense_dense_sq(centers_new, centers_old))
    return labels, centers_new, weight, shift


def _relocate_empty_clusters(X, centers_old, centers_new, weight, labels):
    """_k_means_common.pyx:_relocate_empty_clusters_dense (unit weights)."""
    empty = np.where(np.equal(weight, 0))[0].astype(np.int32)
    n_empty = empty.shape[0]
    if n_empty == 0:
        return
    distances = ((np.asarray(X) - np.asarray(centers_old)[labels]) ** 2).sum(axis=1)
    far = np.argpartition(distances, -n_empty)[:-n_empty - 1:-1].astype(np.int32)
    if np.max(distances) == 0:
        return
    for idx in range(n_empty):
        new_id = empty[idx]
        far_idx = far[idx]
        old_id = labels[far_idx]
        centers_new[old_id] -= X[far_idx]
        centers_new[new_id] = X[far_idx]
        weight[new_id] = 1
        weight[old_id] -= 1


def inertia_dense(X, centers, labels):
    """_inertia_dense, one thread: sequential float32 accumulation over rows."""
    sq = _euclid_dense_dense_sq(X, centers[labels])
    # a sequential fp32 sum == the last element of a fp32 cumsum
    return np.cumsum(sq, dtype=np.float32)[-1] if sq.size else np.float32(0)


def is_same_clustering(labels1, labels2, n_clusters):
    """_k_means_common.pyx:_is_same_clustering."""
    mapping = np.full(n_clusters, -1, dtype=np.int64)
    first = np.full(n_clusters, -1, dtype=np.int64)
    # vectorised restatement: labels1 -> labels2 must be a function
    for l1 in np.unique(labels1):
        vals = np.unique(labels2[labels1 == l1])
        if vals.size != 1:
            return False
    return True


def kmeans_single_lloyd(X, centers_init, max_iter=300, tol=1e-4):
    """_kmeans_single_lloyd (_kmeans.py:630-758)."""
    centers = centers_init
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    labels = labels_old
    for i in range(max_iter):
        labels, centers_new, weight, shift = lloyd_iter(X, centers, update_centers=True)
        centers = centers_new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        center_shift_tot = (shift ** 2).sum()
        if center_shift_tot <= tol:
            break
        labels_old = labels
    if not strict:
        labels = lloyd_iter(X, centers, update_centers=False)
    inertia = inertia_dense(X, centers, labels)
    return labels, inertia, centers, i + 1


def kmeans_fit(X, n_clusters, n_init=10, max_iter=300, tol=1e-4, random_state=None, info=None):
    """KMeans.fit on float32 X.  random_state None -> numpy's global RandomState."""
    if random_state is None:
        random_state = np.random.mtrand._rand
    X = np.array(X, dtype=np.float32, order="C", copy=True)
    tol_abs = np.mean(np.var(X, axis=0)) * tol  # _tolerance on the un-centred copy (fit :1478 via _check_params_vs_input)
    X_mean = X.mean(axis=0)
    X -= X_mean
    best = None
    runs = []
    for _ in range(n_init):
        centers_init, idx = kmeans_plusplus(X, n_clusters, random_state)
        labels, inertia, centers, n_iter = kmeans_single_lloyd(X, centers_init, max_iter, tol_abs)
        runs.append((idx, float(inertia), n_iter))
        if best is None or (inertia < best[1] and not is_same_clustering(labels, best[0], n_clusters)):
            best = (labels, inertia, centers, n_iter)
    labels, inertia, centers, n_iter = best
    centers = centers + X_mean
    if info is not None:
        info.update(runs=runs, inertia=float(inertia), n_iter=n_iter, x_mean=X_mean, tol=float(tol_abs))
    return centers, labels, float(inertia)


def kmeans_predict(X, centers):
    """KMeans.predict: E-step on un-centred float32 X against cluster_centers_."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    return lloyd_iter(X, np.ascontiguousarray(centers, dtype=np.float32), update_centers=False)


def kmeans_fit_predict(X, n_clusters, n_init=10, random_state=None, info=None):
    """feature_extraction.py:52-55: ``kmeans.fit(X); kmeans.predict(X)``."""
    centers, _, _ = kmeans_fit(X, n_clusters, n_init=n_init, random_state=random_state, info=info)
    return kmeans_predict(X, centers), centers
