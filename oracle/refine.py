"""Oracle for R3: correspondence-based mask refinement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows scripts/sampling/feature_extraction.py:
  dense_tracking                    :326-364  one query point per feature cell,
                                              coordinates round-trip h/H -> int(.*H)
  dense_feature_matching_iterative  :176-323  frame t -> t+1 nearest neighbour by
                                              cosine, blended with the cosine to
                                              frame 0 ("aux"): t/(t+1)*cos + 1/(t+1)*cos_aux
                                              (:290-291), 500-point batches with the
                                              target/aux re-normalised inside the
                                              batch loop (:272-274), row arg-max (:293)
  correct_low_res_mask              :367-461  signed-jump filter (:392-409), majority
                                              label per trajectory via Counter.most_common
                                              (:411-421, ties -> first seen), writes in
                                              ascending point order (last writer wins)

Defaults actually in force: top_k=1, use_aux=True,
is_backtracing=False, spatial_filter=True, anchor_label_method="common".
The same torch CPU primitives as the reference are used for the cosine maps
(torch.norm, a [i,c]x[c,s] matmul) so decisions agree on one machine.
"""
from collections import Counter

import numpy as np
import torch

BATCH_SIZE = 500  # feature_extraction.py:196


def dense_tracking(feature_maps, feature_height, feature_width, num_frames):
    """feature_maps: [2F, hw, C] float32 (uncond rows first).  Returns (all_h, all_w)
    as int arrays [F, hw] -- row t holds every trajectory's cell in frame t."""
    fm = torch.as_tensor(np.asarray(feature_maps))
    hw = feature_height * feature_width
    hh, ww = np.meshgrid(np.arange(feature_height), np.arange(feature_width), indexing="ij")
    # :337-343 + :185-190: h / H then int(. * H) on float32 0-dim tensors
    h_list = [int((torch.tensor(h) / feature_height) * feature_height) if h / feature_height < 1 else h for h in hh.reshape(-1)]
    w_list = [int((torch.tensor(w) / feature_width) * feature_width) if True else w for w in ww.reshape(-1)]
    # the reference tests ``input_coordinate[0] < 1`` only (on the h fraction): always true here
    num_points = len(h_list)
    num_batches = num_points // BATCH_SIZE + 1
    all_h = [list(h_list)]
    all_w = [list(w_list)]
    for t in range(num_frames - 1):
        src = fm[num_frames + t].reshape(feature_height, feature_width, -1)
        trg = fm[num_frames + t + 1][None].permute(0, 2, 1)  # [1, c, hw]
        aux = fm[num_frames][None].permute(0, 2, 1)
        out_h, out_w = [], []
        for b in range(num_batches):
            bh = h_list[b * BATCH_SIZE:(b + 1) * BATCH_SIZE]
            bw = w_list[b * BATCH_SIZE:(b + 1) * BATCH_SIZE]
            src_b = src[bh, bw]
            src_b = src_b / torch.norm(src_b, dim=1, keepdim=True)
            trg = trg / torch.norm(trg, dim=1, keepdim=True)
            aux = aux / torch.norm(aux, dim=1, keepdim=True)
            if len(bh) == 0:
                continue
            cos = torch.einsum("ic,fcs->ifs", src_b, trg).reshape(len(bh), -1).numpy()
            cos_aux = torch.einsum("ic,fcs->ifs", src_b, aux).reshape(len(bh), -1).numpy()
            cos = t / (t + 1) * cos + 1 / (t + 1) * cos_aux
            top = np.argmax(cos, axis=1)  # np.argpartition(.., -1)[-1:] == an arg-max
            out_h.extend((top // feature_width).tolist())
            out_w.extend((top % feature_width).tolist())
        h_list, w_list = out_h, out_w
        all_h.append(list(out_h))
        all_w.append(list(out_w))
    return np.array(all_h), np.array(all_w)


def refine_labels(seg_maps, all_h, all_w, spatial_filter=True, spatial_threshold=1):
    """correct_low_res_mask :389-421 on label maps [F, h, w] (int).  Returns the
    refined maps [F, h, w] and the kept-trajectory mask [hw]."""
    ori = np.array(seg_maps)
    new = ori.copy()
    all_h = np.array(all_h)
    all_w = np.array(all_w)
    num_frames, num_points = all_h.shape
    keep = np.ones(num_points, dtype=bool)
    if spatial_filter:
        dh = all_h[1:] - all_h[:-1]
        dw = all_w[1:] - all_w[:-1]
        keep = ~np.any((dh > spatial_threshold) | (dw > spatial_threshold), axis=0)
    for p in np.nonzero(keep)[0]:
        th, tw = all_h[:, p], all_w[:, p]
        labels = [ori[f, th[f], tw[f]] for f in range(num_frames)]
        common = Counter(labels).most_common(1)[0][0]
        for f in range(num_frames):
            new[f, th[f], tw[f]] = common
    return new, keep


def correct_low_res_mask(feature_maps, seg_maps, feature_height, feature_width, num_frames):
    """R3 end to end.  Returns ref_mask = refined maps flattened (:460)."""
    all_h, all_w = dense_tracking(feature_maps, feature_height, feature_width, num_frames)
    new, keep = refine_labels(seg_maps, all_h, all_w)
    return new.reshape(-1), all_h, all_w, keep
