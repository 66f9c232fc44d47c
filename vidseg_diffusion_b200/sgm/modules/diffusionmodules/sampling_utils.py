"""reference sgm/modules/diffusionmodules/sampling_utils.py (the part the Euler EDM sampler uses)."""
from ...util import append_dims


def to_d(x, sigma, denoised):
    return (x - denoised) / append_dims(sigma, x.ndim)
