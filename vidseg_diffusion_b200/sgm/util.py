"""Helpers of the reference's ``sgm/util.py`` that sit on the hot path."""
import os

import torch


def load_target_features(feature_maps_folder, exp_name, timestep, injected_block_type, injected_feature_types, block_idx,
                         device, features=None):
    """reference sgm/util.py:277-296: the stashed q / k (/ v) tensors of one UNet block, keyed
    ``{input|output}_block_{i}_{feature_type}_time_{t}``, to be injected into the same block of another pass.

    ``features`` (not in the reference): a dict with those keys holding tensors that are still in HBM from the pass
    that produced them -- the ``.pt`` round trip through ``feature_maps_folder`` (about 9 GB per clip) is then skipped.
    Without it the files are read exactly as the reference does."""
    current = {}
    folder = os.path.join(feature_maps_folder, exp_name, "feature_maps") if feature_maps_folder is not None else None
    feature_type = None
    for feature_type in injected_feature_types:
        key = f"{injected_block_type}_block_{block_idx}_{feature_type}_time_{timestep}"
        if features is not None:
            if key in features:
                current[key] = features[key].detach().to(device)
            continue
        path = os.path.join(folder, key + ".pt")
        if os.path.exists(path):
            current[key] = torch.load(path).detach().to(device)
    if len(current) == 0:
        raise ValueError(f"No feature maps found for block {injected_block_type}_block_{block_idx} feature type {feature_type} "
                         f"at timestep {timestep} in {folder if features is None else 'the in-memory feature store'}")
    return current
