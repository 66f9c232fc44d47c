"""Oracle for the ``match_gt_mask`` mode.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows scripts/sampling/feature_extraction.py:546-643 (``match_gt_mask``):
  * :549-555  conditional half, per-token max-abs normalisation;
  * :566-574  first window only (``ref_mask is None``): ``KMeans(n_clusters, n_init=10).fit`` on all frames' tokens,
              ``predict`` on frame 0 -> the "fake" mask [hw];
  * :577-596  ground-truth mask (PNG resized NEAREST to the feature grid, or the fake mask itself when no path);
              every fake label takes the most frequent ground-truth label among its cells
              (``np.unique(..., return_counts=True)`` + ``np.argmax``: smallest value on ties); with ``use_gt_mask``
              the ground truth is used as is; reference features = frame 0;
  * :606-612  ``KNeighborsClassifier(n_neighbors=4).fit(ref_feature_map, ref_mask).predict(all tokens)``;
  * :639-640  the returned reference for the NEXT window is all frames' predicted labels / all frames' tokens.
Third-party arithmetic: scikit-learn ``KNeighborsClassifier`` (pinned scikit_learn==1.5.0, requirements/pt2.txt:18;
1.9.0 installed): ``algorithm='auto'`` resolves to brute force for 640-d data (sklearn/neighbors/_base.py), squared
euclidean distances ||q||^2 - 2 q.r + ||r||^2 evaluated in float64 on the upcast float32 rows
(metrics/_pairwise_distances_reduction, ArgKmin32 + MiddleTermComputer32), the k smallest per query sorted by
(distance, index), uniform weights -> ``scipy.stats.mode`` of the neighbours' labels (smallest label on ties,
neighbors/_classification.py:predict).
"""
import numpy as np

from . import kmeans as okm


def majority_map(fake_mask, gt_mask):
    """feature_extraction.py:589-594."""
    ref = np.zeros(fake_mask.shape[0], dtype=np.int64)
    for fake_label in np.unique(fake_mask):
        sub = gt_mask[fake_mask == fake_label]
        values, counts = np.unique(sub, return_counts=True)
        ref[fake_mask == fake_label] = values[np.argmax(counts)]
    return ref


def knn_predict(ref_features, ref_labels, queries, k=4, chunk=2048):
    """KNeighborsClassifier(n_neighbors=k).fit(ref_features, ref_labels).predict(queries), brute force, float64
    distances; ties in distance go to the lower reference index, ties in the vote to the smaller label."""
    r = np.asarray(ref_features, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    classes = np.unique(ref_labels)
    y = np.searchsorted(classes, ref_labels)
    rn = np.einsum("ij,ij->i", r, r)
    out = np.empty(q.shape[0], dtype=classes.dtype)
    for s in range(0, q.shape[0], chunk):
        qc = q[s:s + chunk]
        d = rn[None, :] - 2.0 * (qc @ r.T)          # + ||q||^2 is constant per row
        idx = np.argsort(d, axis=1, kind="stable")[:, :k]
        votes = y[idx]                                # [chunk, k]
        counts = np.zeros((qc.shape[0], classes.size), dtype=np.int32)
        np.add.at(counts, (np.arange(qc.shape[0])[:, None], votes), 1)
        out[s:s + chunk] = classes[np.argmax(counts, axis=1)]   # first maximum = smallest label
    return out


def match_gt_mask(feature_maps, num_frames, h, w, num_masks, gt_mask=None, ref_mask=None, ref_feature_map=None,
                  use_gt_mask=False, info=None):
    """feature_maps [2F, hw, C] float32 (uncond rows first).  gt_mask: int array [hw] already resized to the feature
    grid, or None.  Returns (unique_labels, labels of all frames [F*hw], tokens of all frames [F*hw, C]) -- the last
    two are what the reference hands to the next window as ref_mask / ref_feature_map."""
    x = np.asarray(feature_maps, dtype=np.float32)[num_frames:]
    if x.shape[-1] > 1:
        x = x / np.max(np.abs(x), axis=-1, keepdims=True)
    tokens = x.reshape(-1, x.shape[-1])
    if ref_mask is None:
        centers, _, _ = okm.kmeans_fit(tokens, num_masks)
        fake = okm.kmeans_predict(x[0], centers).astype(np.int64)
        gt = fake if gt_mask is None else np.asarray(gt_mask).reshape(-1)
        ref_mask = gt if use_gt_mask else majority_map(fake, gt)
        ref_feature_map = x[0]
        if info is not None:
            info.update(fake_mask=fake, first_ref_mask=np.asarray(ref_mask).copy())
    unique_labels = np.unique(ref_mask)
    labels = knn_predict(ref_feature_map, np.asarray(ref_mask), tokens, 4)
    return unique_labels, labels, tokens
