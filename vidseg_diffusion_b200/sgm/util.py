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


def load_xt(feature_maps_folder, exp_name, timestep, device, features=None):
    """reference sgm/util.py:298-311: the source run's latent after sampler step ``timestep`` (``xt_time_{t}.pt``), the
    background of the latent blending.  ``features``: the same in-HBM store as ``load_target_features``."""
    key = f"xt_time_{timestep}"
    if features is not None:
        if key not in features:
            raise ValueError(f"No feature maps found for xt at timestep {timestep} in the in-memory feature store")
        return features[key].detach().to(device)
    path = os.path.join(feature_maps_folder, exp_name, "feature_maps", key + ".pt")
    if os.path.exists(path):
        return torch.load(path).detach().to(device)
    raise ValueError(f"No feature maps found for xt at timestep {timestep} in {os.path.dirname(path)}")


def default(val, d):
    """reference sgm/util.py: ``val`` unless it is None."""
    from inspect import isfunction
    if val is not None:
        return val
    return d() if isfunction(d) else d


def append_zero(x):
    return torch.cat([x, x.new_zeros([1])])


def append_dims(x, target_dims):
    """reference sgm/util.py:192-199"""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def get_obj_from_str(string):
    """reference sgm/util.py:178-185.  The reference's configs name classes as ``sgm.modules...``; those resolve to this
    package's mirror of the same module path, so the reference's yaml files instantiate the B200 classes unchanged."""
    import importlib
    module, cls = string.rsplit(".", 1)
    if module == "sgm" or module.startswith("sgm."):
        module = __name__.rsplit(".", 1)[0] + module[3:]
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    """reference sgm/util.py:168-175"""
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))
