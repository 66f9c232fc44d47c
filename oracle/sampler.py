"""Oracle for the sampler row: Euler EDM sampling loop with classifier-free guidance,
mask modulation / feature injection switches and latent blending.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain fp32 torch-CPU, functional; the network is a callable the
caller supplies (the reference UNet in the golden generator, oracle/unet.py in the tests).

Follows, in the reference tree:
  sgm/modules/diffusionmodules/sampling.py
    BaseDiffusionSampler.prepare_sampling_loop :45-59   sigmas, x *= sqrt(1 + sigma_0^2), s_in
    EDMSampler.sampler_step :102-132                    sigma_hat, denoise, to_d, Euler update
    EDMSampler.__call__ :146-262                        step range [t_start, t_end], modulate / inject switches,
                                                        modulate_params bookkeeping, latent blending, callbacks
  sgm/modules/diffusionmodules/discretizer.py :29-70    EDM and legacy-DDPM schedules
  sgm/modules/diffusionmodules/denoiser.py :25-82       pre-conditioning, sigma / c_noise quantisation
  sgm/modules/diffusionmodules/denoiser_scaling.py      EpsScaling (SD-2.1), VScalingWithEDMcNoise (SVD)
  sgm/modules/diffusionmodules/guiders.py :23-100       VanillaCFG, LinearPredictionGuider
Pinned by tests/golden/make_sampler_goldens.py: with the reference UNet as the network this loop reproduces the
reference sampler bit for bit (torch.equal) on the plain and on the modulated + injected + blended run.
"""
import numpy as np
import torch


def legacy_ddpm_sigmas(n, linear_start=0.00085, linear_end=0.0120, num_timesteps=1000, append_zero=True, flip=False):
    """LegacyDDPMDiscretization (discretizer.py:44-70) through Discretization.__call__ (:17-20)."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64) ** 2).numpy()
    acp = np.cumprod(1.0 - betas, axis=0)
    if n < num_timesteps:
        steps = np.linspace(num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
        acp = acp[steps]
    elif n != num_timesteps:
        raise ValueError
    sig = torch.flip(torch.tensor((1 - acp) / acp, dtype=torch.float32) ** 0.5, (0,))
    if append_zero:
        sig = torch.cat([sig, sig.new_zeros([1])])
    return torch.flip(sig, (0,)) if flip else sig


def edm_sigmas(n, sigma_min=0.002, sigma_max=80.0, rho=7.0):
    """EDMDiscretization (discretizer.py:29-41) + appended zero."""
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (hi + ramp * (lo - hi)) ** rho
    return torch.cat([sig, sig.new_zeros([1])])


def eps_scaling(sigma):
    return torch.ones_like(sigma), -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()


def v_scaling_edm_cnoise(sigma):
    return 1.0 / (sigma ** 2 + 1.0), -sigma / (sigma ** 2 + 1.0) ** 0.5, 1.0 / (sigma ** 2 + 1.0) ** 0.5, 0.25 * sigma.log()


def make_discrete_quantizer(num_idx=1000):
    """DiscreteDenoiser (denoiser.py:51-82) with flip=True, do_append_zero=False: returns (quantize_sigma, c_noise_to_idx)."""
    table = legacy_ddpm_sigmas(num_idx, append_zero=False, flip=True)

    def to_idx(s):
        return (s - table[:, None]).abs().argmin(dim=0).view(s.shape)

    return (lambda s: table[to_idx(s)]), to_idx


def denoise(network, x, sigma, cond, scaling, quantizer=None, **flags):
    """Denoiser.forward (denoiser.py:25-48); network(x_in, c_noise, cond, **flags)."""
    if quantizer is not None:
        sigma = quantizer[0](sigma)
    shape = sigma.shape
    s = sigma[(...,) + (None,) * (x.ndim - sigma.ndim)]
    c_skip, c_out, c_in, c_noise = scaling(s)
    c_noise = c_noise.reshape(shape)
    if quantizer is not None:
        c_noise = quantizer[1](c_noise)
    return network(x * c_in, c_noise, cond, **flags) * c_out + x * c_skip


def euler_edm_sample(network, x, cond, uc, sigmas, scaling, guidance_scale, quantizer=None, t_start=None, t_end=None,
                     is_modulate=False, modulate_params=None, is_latent_blending=False, feature_height=None,
                     feature_width=None, xt_store=None, img_callback=None, frame_scales=None, is_smooth_latent=False,
                     first_stage=None):
    """EDMSampler.__call__ with s_churn = 0 (both configs).  ``guidance_scale``: VanillaCFG scale, or None for the
    identity guider; ``frame_scales`` [T]: LinearPredictionGuider.  ``cond`` / ``uc``: dicts whose "crossattn" /
    "vector" / "concat" entries are concatenated (uc first).  ``xt_store``: {f"xt_time_{i}": latent} for the blending.
    ``is_smooth_latent`` (sampler_step :116-124, switch :199-210): on sampler steps 23 / 24 the denoised latent is decoded
    by ``first_stage = (decode, encode)``, every third frame (offset 1 / 2) is replaced by the mean of its neighbours and
    the clip is encoded again."""
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    num_sigmas = len(sigmas)
    if is_modulate:
        mts = modulate_params["modulate_timestep_frames"]
        modulate_timestep = modulate_params["modulate_timestep"] if len(mts) == 0 else mts.keys()
        is_injected_features = modulate_params["is_injected_features"]
    else:
        is_injected_features = False
    t_start = 0 if t_start is None else t_start
    t_end = num_sigmas if t_end is None else t_end
    guided = guidance_scale is not None or frame_scales is not None
    for i in list(range(num_sigmas - 1))[t_start:(t_end + 1)]:
        is_modulate_step = bool(is_modulate and i in modulate_timestep)
        is_injected_step = bool(is_modulate and is_injected_features and i >= min(modulate_timestep))
        if modulate_params is not None:
            modulate_params["timestep"] = i
        if is_modulate and i in modulate_timestep:
            mts = modulate_params["modulate_timestep_frames"]
            modulate_params["modulate_timestep_frames_group"] = mts[i] if len(mts) > 0 else list(range(modulate_params["num_frames"]))
        sigma, next_sigma = s_in * sigmas[i], s_in * sigmas[i + 1]
        sigma_hat = sigma * (0.0 + 1.0)
        flags = dict(is_modulate_step=is_modulate_step, is_injected_step=is_injected_step, modulate_params=modulate_params)
        if sigma_hat.mean() < 1e-6:
            denoised = x
        elif guided:
            c2 = {k: (torch.cat((uc[k], cond[k]), 0) if k in ("vector", "crossattn", "concat") else cond[k]) for k in cond}
            out = denoise(network, torch.cat([x] * 2), torch.cat([sigma_hat] * 2), c2, scaling, quantizer, **flags)
            x_u, x_c = out.chunk(2)
            if frame_scales is not None:
                t = frame_scales.numel()
                sc = frame_scales.reshape(1, t).repeat(x_u.shape[0] // t, 1).reshape(-1)[(...,) + (None,) * (x.ndim - 1)]
                denoised = x_u + sc * (x_c - x_u)
            else:
                denoised = x_u + guidance_scale * (x_c - x_u)
        else:
            denoised = denoise(network, x, sigma_hat, cond, scaling, quantizer, **flags)
        if is_smooth_latent and i in (23, 24):
            step = 1 if i == 23 else 2
            frames = first_stage[0](denoised)
            for f in range(1, frames.shape[0] - 1):
                if (f - step) % 3 == 0:
                    frames[f] = 0.5 * (frames[f - 1] + frames[f + 1])
            denoised = first_stage[1](frames)
        d = (x - denoised) / sigma_hat[(...,) + (None,) * (x.ndim - 1)]
        dt = (next_sigma - sigma_hat)[(...,) + (None,) * (x.ndim - 1)]
        x = x + dt * d
        if is_latent_blending and modulate_params["latent_mask_start"] <= i <= modulate_params["latent_mask_end"]:
            ori = xt_store[f"xt_time_{i}"].to(x.dtype)
            fm = torch.stack([torch.as_tensor(m) for m in modulate_params["feature_masks"]], dim=0)
            fm = fm.reshape(fm.shape[0], 28 if feature_height is None else feature_height,
                            52 if feature_width is None else feature_width).unsqueeze(1)
            fm = torch.nn.functional.interpolate(fm, size=(x.shape[-2], x.shape[-1]), mode="nearest")
            x = (x * fm + ori * (1 - fm)).float()
        if img_callback is not None and (not is_modulate or i >= min(modulate_timestep)):
            img_callback(x, i)
    return x
