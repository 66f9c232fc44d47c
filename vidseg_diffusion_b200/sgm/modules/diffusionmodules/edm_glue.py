"""Scalar glue of the EDM-style sampler, in one place: noise schedules, denoiser pre-conditioning, guidance and the
conditioning adapter in front of the UNet.

The reference spreads these over ``discretizer.py``, ``denoiser_scaling.py``, ``denoiser.py``, ``guiders.py`` and
``wrappers.py`` and addresses them by ``target:`` strings in its yaml files (configs/inference/sd_2_1.yaml:7-16, 63-79;
svd.yaml); modules of those names re-export the classes defined here, so the configs resolve unchanged.  Everything in
this file is a handful of scalars per sampling step, computed with torch in the reference's operation order (the results
feed ``vidseg_sampler_step`` and must match the reference bit for bit, tests/test_sampler_host.py); the per-element work
on the latent is in csrc/sampler.cu.

Design differences from the reference (same numbers, different structure):
  * one parametrised ``Preconditioner`` instead of four scaling classes: the EDM / v / eps parametrisations differ only
    in sigma_data, the sign of c_out, whether c_skip is 1 and how the noise level is presented to the network;
  * guidance is a per-sample scale vector (``sample_scales``): classifier-free guidance with one scale and the per-frame
    ramp of the video model are the same combination x_u + s * (x_c - x_u), which the fused step kernel applies;
  * the denoiser exposes ``raw`` (network output + c_skip + c_out) next to the reference's ``forward`` so that the
    combination can happen inside that kernel, and ``bind`` to build the callable handed to the sampler.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from ...util import append_dims, instantiate_from_config

# ------------------------------------------------------------------------------------------------------------------
# noise schedules (reference discretizer.py)
# ------------------------------------------------------------------------------------------------------------------


class SigmaSchedule:
    """n noise levels, descending, plus the final 0 (``do_append_zero``); ``flip`` returns them ascending."""

    def levels(self, n, device):
        raise NotImplementedError

    def __call__(self, n, do_append_zero=True, device="cpu", flip=False):
        sig = self.levels(n, device)
        if do_append_zero:
            sig = torch.cat([sig, sig.new_zeros([1])])
        return torch.flip(sig, (0,)) if flip else sig

    get_sigmas = lambda self, n, device="cpu": self.levels(n, device)   # the reference's name for ``levels``


class KarrasSchedule(SigmaSchedule):
    """EDMDiscretization (discretizer.py:29-41): rho-warped interpolation between sigma_max and sigma_min."""

    def __init__(self, sigma_min=0.002, sigma_max=80.0, rho=7.0):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho

    def levels(self, n, device="cpu"):
        hi, lo = self.sigma_max ** (1 / self.rho), self.sigma_min ** (1 / self.rho)
        t = torch.linspace(0, 1, n, device=device)
        return (hi + t * (lo - hi)) ** self.rho


class DDPMSubsampledSchedule(SigmaSchedule):
    """LegacyDDPMDiscretization (discretizer.py:44-70): sigma_t = sqrt((1 - abar_t) / abar_t) of the 1000-step
    linear-in-sqrt(beta) training schedule, at n roughly equally spaced training steps."""

    def __init__(self, linear_start=0.00085, linear_end=0.0120, num_timesteps=1000):
        self.num_timesteps = num_timesteps
        sqrt_beta = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64)
        self.alphas_cumprod = np.cumprod(1.0 - (sqrt_beta ** 2).numpy(), axis=0)

    def levels(self, n, device="cpu"):
        if n > self.num_timesteps:
            raise ValueError(f"{n} sampling steps from a {self.num_timesteps}-step training schedule")
        abar = self.alphas_cumprod
        if n < self.num_timesteps:
            picks = np.linspace(self.num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
            abar = abar[picks]
        ascending = torch.tensor((1 - abar) / abar, dtype=torch.float32, device=device) ** 0.5
        return torch.flip(ascending, (0,))


# ------------------------------------------------------------------------------------------------------------------
# pre-conditioning (reference denoiser_scaling.py, denoiser.py)
# ------------------------------------------------------------------------------------------------------------------


class Preconditioner:
    """D(x; sigma) = c_skip x + c_out F(c_in x; c_noise).

    ``sigma_data``    data standard deviation of the parametrisation (1 for the v / eps forms);
    ``out_sign``      +1: F predicts the clean-signal residual (EDM); -1: F predicts v or the noise;
    ``unit_skip``     eps form: c_skip = 1 and c_out = -sigma (no normalisation of the output);
    ``log_noise``     the network sees 0.25 ln(sigma) (EDM convention) instead of sigma itself."""

    def __init__(self, sigma_data=1.0, out_sign=1.0, unit_skip=False, log_noise=False):
        self.sigma_data, self.out_sign, self.unit_skip, self.log_noise = sigma_data, out_sign, unit_skip, log_noise

    def __call__(self, sigma):
        total_var = sigma ** 2 + self.sigma_data ** 2
        c_in = 1 / total_var ** 0.5
        if self.unit_skip:
            c_skip, c_out = torch.ones_like(sigma, device=sigma.device), self.out_sign * sigma
        else:
            c_skip = self.sigma_data ** 2 / total_var
            c_out = self.out_sign * sigma * self.sigma_data / total_var ** 0.5
        c_noise = 0.25 * sigma.log() if self.log_noise else sigma.clone()
        return c_skip, c_out, c_in, c_noise


class EDMScaling(Preconditioner):
    def __init__(self, sigma_data=0.5):
        super().__init__(sigma_data=sigma_data, log_noise=True)


class EpsScaling(Preconditioner):            # SD-2.1
    def __init__(self):
        super().__init__(out_sign=-1.0, unit_skip=True)


class VScaling(Preconditioner):
    def __init__(self):
        super().__init__(out_sign=-1.0)


class VScalingWithEDMcNoise(Preconditioner):  # SVD
    def __init__(self):
        super().__init__(out_sign=-1.0, log_noise=True)


class BoundDenoiser:
    """The callable the sampler receives (the reference scripts build a closure, svd_single_video_inference.py:322-330):
    ``denoiser(network, input, sigma, c, **flags, **additional_model_inputs)`` with network and extras fixed."""

    def __init__(self, denoiser, network, **additional_model_inputs):
        self.denoiser, self.network, self.extra = denoiser, network, additional_model_inputs

    def _args(self, kwargs):
        merged = dict(kwargs)
        merged.update(self.extra)
        return merged

    def __call__(self, input, sigma, c, **kwargs):
        return self.denoiser(self.network, input, sigma, c, **self._args(kwargs))

    def raw(self, input, sigma, c, **kwargs):
        return self.denoiser.raw(self.network, input, sigma, c, **self._args(kwargs))


class Denoiser(nn.Module):
    """reference denoiser.py:11-48.  ``raw`` stops before the combination with the input."""

    def __init__(self, scaling_config):
        super().__init__()
        self.scaling = instantiate_from_config(scaling_config)

    def possibly_quantize_sigma(self, sigma):
        return sigma

    def possibly_quantize_c_noise(self, c_noise):
        return c_noise

    def raw(self, network, input, sigma, cond, is_modulate_step=False, is_injected_step=False, modulate_params=None,
            **additional_model_inputs):
        """-> (network output, c_skip [B], c_out [B])"""
        per_sample = sigma.shape
        sigma = append_dims(self.possibly_quantize_sigma(sigma), input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        level = self.possibly_quantize_c_noise(c_noise.reshape(per_sample))
        net = network(input * c_in, level, cond, is_modulate_step=is_modulate_step, is_injected_step=is_injected_step,
                      modulate_params=modulate_params, **additional_model_inputs)
        return net, c_skip.reshape(per_sample), c_out.reshape(per_sample)

    def forward(self, network, input, sigma, cond, **kwargs):
        net, c_skip, c_out = self.raw(network, input, sigma, cond, **kwargs)
        return net * append_dims(c_out, input.ndim) + input * append_dims(c_skip, input.ndim)

    def bind(self, network, **additional_model_inputs):
        return BoundDenoiser(self, network, **additional_model_inputs)


class DiscreteDenoiser(Denoiser):
    """reference denoiser.py:51-82 (SD-2.1): sigma snaps to the nearest entry of the training schedule and the network
    is conditioned on that entry's index (its integer timestep)."""

    def __init__(self, scaling_config, num_idx, discretization_config, do_append_zero=False, quantize_c_noise=True, flip=True):
        super().__init__(scaling_config)
        self.discretization = instantiate_from_config(discretization_config)
        self.register_buffer("sigmas", self.discretization(num_idx, do_append_zero=do_append_zero, flip=flip))
        self.quantize_c_noise, self.num_idx = quantize_c_noise, num_idx

    def sigma_to_idx(self, sigma):
        return (sigma - self.sigmas[:, None]).abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def possibly_quantize_sigma(self, sigma):
        return self.sigmas[self.sigma_to_idx(sigma)]

    def possibly_quantize_c_noise(self, c_noise):
        return self.sigma_to_idx(c_noise) if self.quantize_c_noise else c_noise


# ------------------------------------------------------------------------------------------------------------------
# guidance (reference guiders.py)
# ------------------------------------------------------------------------------------------------------------------

_DOUBLED_KEYS = ("vector", "crossattn", "concat")


class Guidance:
    """Classifier-free guidance as data: which conditioning entries are doubled, and the scale of every sample."""

    doubles_batch = True
    extra_keys = ()

    def sample_scales(self, batch, device):
        """fp32 [batch] guidance scales of the un-doubled batch (None: no guidance, the batch is not doubled)."""
        raise NotImplementedError

    def prepare_inputs(self, x, s, c, uc):
        if not self.doubles_batch:
            return x, s, dict(c)
        merged = {}
        for key, value in c.items():
            if key in _DOUBLED_KEYS or key in self.extra_keys:
                merged[key] = torch.cat((uc[key], value), 0)
            else:
                assert value == uc[key], f"conditioning entry {key!r} differs between the two branches"
                merged[key] = value
        return torch.cat([x, x]), torch.cat([s, s]), merged

    def __call__(self, x, sigma):
        """The combination on a doubled batch in torch (the sampler uses the fused kernel instead)."""
        if not self.doubles_batch:
            return x
        x_u, x_c = x.chunk(2)
        return x_u + append_dims(self.sample_scales(x_u.shape[0], x.device), x.ndim) * (x_c - x_u)


class IdentityGuider(Guidance):
    doubles_batch = False

    def sample_scales(self, batch, device):
        return None


class VanillaCFG(Guidance):
    def __init__(self, scale):
        self.scale = scale

    def sample_scales(self, batch, device):
        return torch.full((batch,), float(self.scale), dtype=torch.float32, device=device)


class LinearPredictionGuider(Guidance):
    """SVD (guiders.py:60-100): the scale ramps linearly from min_scale to max_scale over the frames of each clip."""

    def __init__(self, max_scale, num_frames, min_scale=1.0, additional_cond_keys=None):
        self.min_scale, self.max_scale, self.num_frames = min_scale, max_scale, num_frames
        self.scale = torch.linspace(min_scale, max_scale, num_frames).unsqueeze(0)
        keys = additional_cond_keys if additional_cond_keys is not None else []
        self.additional_cond_keys = [keys] if isinstance(keys, str) else list(keys)
        self.extra_keys = tuple(self.additional_cond_keys)

    def sample_scales(self, batch, device):
        if batch % self.num_frames:
            raise ValueError(f"batch {batch} is not a whole number of {self.num_frames}-frame clips")
        return self.scale.reshape(-1).repeat(batch // self.num_frames).to(device=device, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------------------------
# conditioning adapter (reference wrappers.py)
# ------------------------------------------------------------------------------------------------------------------


class IdentityWrapper(nn.Module):
    def __init__(self, diffusion_model, compile_model=False):
        super().__init__()
        if compile_model:
            raise NotImplementedError("torch.compile is not used on this path: the UNet runs on hand-written kernels")
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)


class OpenAIWrapper(IdentityWrapper):
    """Conditioning dict -> UNet signature: "concat" joins the latent along channels (SVD's conditioning frame),
    "crossattn" is the context, "vector" the class embedding input."""

    def forward(self, x, t, c, **kwargs):
        extra = c.get("concat")
        if extra is not None and extra.numel():
            x = torch.cat((x, extra), dim=1)
        return self.diffusion_model(x, timesteps=t, context=c.get("crossattn"), y=c.get("vector"), **kwargs)
