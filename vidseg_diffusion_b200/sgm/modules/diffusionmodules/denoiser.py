"""Denoiser wrappers (reference sgm/modules/diffusionmodules/denoiser.py): pre-conditioning around the UNet.

``forward`` is the reference's call (returns the denoised sample).  ``raw`` returns the network output together with
c_skip / c_out instead: the fused sampler step (csrc/sampler.cu, ``EulerEDMSampler``) applies
``net * c_out + input * c_skip`` inside its single pass over the latent.  ``bind`` builds the callable the reference's
scripts hand to the sampler (svd_single_video_inference.py:322-330) with that second entry attached."""
import torch
import torch.nn as nn

from ...util import append_dims, instantiate_from_config


class BoundDenoiser:
    """``lambda input, sigma, c, **kw: denoiser(network, input, sigma, c, **kw, **additional_model_inputs)`` as an
    object, so that the sampler can also reach ``raw``."""

    def __init__(self, denoiser, network, **additional_model_inputs):
        self.denoiser, self.network, self.extra = denoiser, network, additional_model_inputs

    def __call__(self, input, sigma, c, **kwargs):
        return self.denoiser(self.network, input, sigma, c, **kwargs, **self.extra)

    def raw(self, input, sigma, c, **kwargs):
        return self.denoiser.raw(self.network, input, sigma, c, **kwargs, **self.extra)


class Denoiser(nn.Module):
    def __init__(self, scaling_config):
        super().__init__()
        self.scaling = instantiate_from_config(scaling_config)

    def possibly_quantize_sigma(self, sigma):
        return sigma

    def possibly_quantize_c_noise(self, c_noise):
        return c_noise

    def raw(self, network, input, sigma, cond, is_modulate_step=False, is_injected_step=False, modulate_params=None,
            **additional_model_inputs):
        """(network output, c_skip [B], c_out [B]) of ``forward`` before they are combined."""
        sigma = self.possibly_quantize_sigma(sigma)
        sigma_shape = sigma.shape
        sigma = append_dims(sigma, input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(sigma_shape))
        net = network(input * c_in, c_noise, cond, is_modulate_step=is_modulate_step, is_injected_step=is_injected_step,
                      modulate_params=modulate_params, **additional_model_inputs)
        return net, c_skip.reshape(sigma_shape), c_out.reshape(sigma_shape)

    def forward(self, network, input, sigma, cond, is_modulate_step=False, is_injected_step=False, modulate_params=None,
                **additional_model_inputs):
        net, c_skip, c_out = self.raw(network, input, sigma, cond, is_modulate_step=is_modulate_step,
                                      is_injected_step=is_injected_step, modulate_params=modulate_params,
                                      **additional_model_inputs)
        return net * append_dims(c_out, input.ndim) + input * append_dims(c_skip, input.ndim)

    def bind(self, network, **additional_model_inputs):
        return BoundDenoiser(self, network, **additional_model_inputs)


class DiscreteDenoiser(Denoiser):
    """SD-2.1: sigma snapped to the 1000-entry training schedule, c_noise = its index (the UNet's timestep)."""

    def __init__(self, scaling_config, num_idx, discretization_config, do_append_zero=False, quantize_c_noise=True, flip=True):
        super().__init__(scaling_config)
        self.discretization = instantiate_from_config(discretization_config)
        sigmas = self.discretization(num_idx, do_append_zero=do_append_zero, flip=flip)
        self.register_buffer("sigmas", sigmas)
        self.quantize_c_noise = quantize_c_noise
        self.num_idx = num_idx

    def sigma_to_idx(self, sigma):
        dists = sigma - self.sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def possibly_quantize_sigma(self, sigma):
        return self.idx_to_sigma(self.sigma_to_idx(sigma))

    def possibly_quantize_c_noise(self, c_noise):
        return self.sigma_to_idx(c_noise) if self.quantize_c_noise else c_noise
