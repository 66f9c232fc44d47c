"""Euler EDM sampler with mask modulation, feature injection and latent blending
(reference sgm/modules/diffusionmodules/sampling.py:25-262, 495-499).

The step schedule, the modulation / injection switches per step and the callbacks are the reference's control flow,
restated.  What changes is the arithmetic between two UNet evaluations: the reference runs ~14 elementwise torch
kernels over the latent per step (denoiser pre-conditioning, guidance, to_d, Euler update, blending); here that is ONE
launch of ``vidseg_sampler_step`` (csrc/sampler.cu), bit-identical to the eager chain.  The source run's q / k / x_t can
stay in HBM (``modulate_params["features"]``, see ``sgm/util.load_target_features`` / ``load_xt``) instead of going
through ``.pt`` files.
"""
import torch

from .... import _lib
from ...util import append_dims, default, instantiate_from_config, load_xt

DEFAULT_GUIDER = {"target": "sgm.modules.diffusionmodules.guiders.IdentityGuider"}


def fused_step(x, net, c_skip, c_out, scales, sigma_hat, sigma_next, mask=None, ori_xt=None):
    """x' = blend(euler(x, guide(net * c_out + x * c_skip))) in one pass.  ``net`` is [B, ...] (scales None) or
    [2B, ...]; ``mask`` [B, fh, fw] fp32 / fp64 with ``ori_xt`` [B, ...], or both None."""
    lib = _lib.load()
    x = _lib.require_cuda_tensor(x.contiguous(), torch.float32, "x")
    if x.ndim != 4:
        raise _lib.VidsegError("the sampler step works on [B, C, H, W] latents")
    b, ch, h, w = x.shape
    guided = scales is not None
    f32 = lambda t, n, name: _check(t, n, name, x.device)
    net = _lib.require_cuda_tensor(net.contiguous(), torch.float32, "net")
    if tuple(net.shape) != ((2 if guided else 1) * b, ch, h, w):
        raise _lib.VidsegError(f"network output {tuple(net.shape)} does not match the latent {tuple(x.shape)} (guided={guided})")
    g = 2 if guided else 1
    c_skip, c_out = f32(c_skip, g * b, "c_skip"), f32(c_out, g * b, "c_out")
    sigma_hat, sigma_next = f32(sigma_hat, b, "sigma_hat"), f32(sigma_next, b, "sigma_next")
    scales = f32(scales, b, "scales") if guided else None
    mh = mw = 0
    is64 = 0
    if mask is not None:
        if mask.dtype not in (torch.float32, torch.float64) or mask.ndim != 3 or mask.shape[0] != b:
            raise _lib.VidsegError("mask must be [B, fh, fw] float32 or float64")
        mask = mask.to(x.device).contiguous()
        ori_xt = _lib.require_cuda_tensor(ori_xt.contiguous(), torch.float32, "ori_xt")
        if ori_xt.shape != x.shape:
            raise _lib.VidsegError("ori_xt must have the latent's shape")
        mh, mw = mask.shape[1], mask.shape[2]
        is64 = 1 if mask.dtype == torch.float64 else 0
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.vidseg_sampler_step(
            x.data_ptr(), net.data_ptr(), c_skip.data_ptr(), c_out.data_ptr(), scales.data_ptr() if guided else None,
            sigma_hat.data_ptr(), sigma_next.data_ptr(), mask.data_ptr() if mask is not None else None, is64,
            ori_xt.data_ptr() if mask is not None else None, out.data_ptr(), b, ch, h, w, mh, mw, 1 if guided else 0,
            _lib.stream_ptr()), "sampler_step")
    return out


def _check(t, n, name, device):
    t = t.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    if t.numel() != n:
        raise _lib.VidsegError(f"{name}: expected {n} values, got {t.numel()}")
    return t


class _ModulationSchedule:
    """Which sampler steps modulate, and from which step on features are injected (reference sampling.py:153-196).

    ``modulate_timestep_frames`` ({step: [frames]}; empty -> every frame on the steps of ``modulate_timestep``) decides
    the steps; injection, when ``is_injected_features`` is set, starts at the first of them and lasts to the end, and the
    image callback of a modulated run only fires from that step on."""

    def __init__(self, params):
        self.params = params
        self.per_step_frames = {}
        self.steps = ()
        self.inject = False
        if params is not None:
            self.per_step_frames = params["modulate_timestep_frames"]
            self.steps = tuple(self.per_step_frames.keys()) if len(self.per_step_frames) else tuple(params["modulate_timestep"])
            self.inject = bool(params["is_injected_features"])

    def _first(self):
        return min(self.steps)   # ValueError on an empty list, as in the reference

    def enter(self, i):
        """-> (modulate on step i, inject on step i); records the frames this step modulates."""
        if self.params is None:
            return False, False
        modulate = i in self.steps
        if modulate:
            frames = self.per_step_frames[i] if len(self.per_step_frames) else list(range(self.params["num_frames"]))
            self.params["modulate_timestep_frames_group"] = frames
        return modulate, bool(self.inject and i >= self._first())

    def reports(self, i):
        return self.params is None or i >= self._first()


class BaseDiffusionSampler:
    def __init__(self, discretization_config, num_steps=None, guider_config=None, verbose=False, device="cuda"):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(default(guider_config, DEFAULT_GUIDER))
        self.verbose = verbose
        self.device = device

    def prepare_sampling_loop(self, x, cond, uc=None, num_steps=None, inversion=False):
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device=self.device)
        if inversion:
            sigmas = sigmas.flip(0)
            sigmas[0] += 1e-8
        uc = default(uc, cond)
        x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
        num_sigmas = len(sigmas)
        s_in = x.new_ones([x.shape[0]])
        return x, s_in, sigmas, num_sigmas, cond, uc

    def denoise(self, x, denoiser, sigma, cond, uc, is_modulate_step=False, is_injected_step=False, modulate_params=None):
        """reference sampling.py:61-67 (kept for callers that want the denoised sample itself)."""
        denoised = denoiser(*self.guider.prepare_inputs(x, sigma, cond, uc), is_modulate_step=is_modulate_step,
                            is_injected_step=is_injected_step, modulate_params=modulate_params)
        return self.guider(denoised, sigma)

    def get_sigma_gen(self, num_sigmas):
        return range(num_sigmas - 1)


class SingleStepDiffusionSampler(BaseDiffusionSampler):
    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, uc, *args, **kwargs):
        raise NotImplementedError

    def euler_step(self, x, d, dt):
        return x + dt * d


class EDMSampler(SingleStepDiffusionSampler):
    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise

    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, uc=None, gamma=0.0, is_modulate_step=False,
                     is_injected_step=False, modulate_params=None, is_smooth_latent=False, model=None, smooth_step_size=None,
                     blend_mask=None, blend_xt=None):
        """reference sampling.py:102-132; ``blend_mask`` / ``blend_xt`` (not in the reference's signature) fold the latent
        blending of ``__call__`` (:229-250) into the same launch."""
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            eps = torch.randn_like(x) * self.s_noise
            x = x + eps * append_dims(sigma_hat**2 - sigma**2, x.ndim) ** 0.5
        b = x.shape[0]
        if is_smooth_latent:
            # reference :116-124: the denoised latent goes through the first stage, every third frame (offset
            # smooth_step_size) becomes the mean of its neighbours, and the clip is encoded again.  The denoised sample is
            # needed by itself here, so the denoiser / guider run unfused; the Euler update (+ blending) stays one launch.
            if model is None:
                raise ValueError("is_smooth_latent needs model= (the first stage that decodes / encodes the latent)")
            if sigma_hat.mean() < 1e-6:
                denoised = x
            else:
                denoised = self.denoise(x, denoiser, sigma_hat, cond, uc, is_modulate_step=is_modulate_step,
                                        is_injected_step=is_injected_step, modulate_params=modulate_params)
            frames = model.decode_first_stage(denoised)
            if not frames.is_contiguous():
                frames = frames.contiguous()
            for frame_id in range(1, frames.shape[0] - 1):
                if (frame_id - smooth_step_size) % 3 == 0:
                    frames[frame_id] = 0.5 * (frames[frame_id - 1] + frames[frame_id + 1])
            denoised = model.encode_first_stage(frames).to(x.dtype)
            one, zero = torch.ones(b, device=x.device), torch.zeros(b, device=x.device)
            x_next = fused_step(x, denoised.contiguous(), zero, one, None, sigma_hat, next_sigma, blend_mask, blend_xt)
            return self.possible_correction_step(x_next, x, None, None, next_sigma, denoiser, cond, uc)
        if sigma_hat.mean() < 1e-6:
            # denoised = x: no network call
            one, zero = torch.ones(b, device=x.device), torch.zeros(b, device=x.device)
            return fused_step(x, x, zero, one, None, sigma_hat, next_sigma, blend_mask, blend_xt)
        inp, sig, c_in = self.guider.prepare_inputs(x, sigma_hat, cond, uc)
        kw = dict(is_modulate_step=is_modulate_step, is_injected_step=is_injected_step, modulate_params=modulate_params)
        scales = self.guider.sample_scales(b, x.device)
        if hasattr(denoiser, "raw"):
            net, c_skip, c_out = denoiser.raw(inp, sig, c_in, **kw)
        else:
            # an opaque callable returns the denoised sample: net * 1 + x * 0 leaves it unchanged bit for bit
            net = denoiser(inp, sig, c_in, **kw)
            c_skip, c_out = torch.zeros(inp.shape[0], device=x.device), torch.ones(inp.shape[0], device=x.device)
        x_next = fused_step(x, net, c_skip, c_out, scales, sigma_hat, next_sigma, blend_mask, blend_xt)
        d = None  # the Euler sampler's correction step is the identity (sampling.py:495-499)
        return self.possible_correction_step(x_next, x, d, None, next_sigma, denoiser, cond, uc)

    def possible_correction_step(self, euler_step, x, d, dt, next_sigma, denoiser, cond, uc):
        raise NotImplementedError

    def add_noise(self, x, cond, uc=None, num_steps=None, noise_level=0):
        """reference sampling.py:134-144"""
        _, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        x = x + torch.randn_like(x) * sigmas[noise_level]
        x /= torch.sqrt(1.0 + sigmas[0] ** 2.0)
        return x

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, callback=None, img_callback=None, is_modulate=False,
                 modulate_params=None, uc_list=None, t_start=None, t_end=None, is_latent_blending=False,
                 feature_height=None, feature_width=None, is_smooth_latent=False, model=None):
        """reference sampling.py:146-262: steps t_start..t_end of the schedule; on the steps named by ``modulate_params``
        the UNet modulates its attention / feed-forward outputs with the feature masks, from the first of them on it
        takes the source run's q / k, and inside [latent_mask_start, latent_mask_end] the latent is blended with the
        source run's x_t outside the masks.  The per-step bookkeeping written into ``modulate_params`` ("timestep",
        "modulate_timestep_frames_group") is what the UNet mirrors read, as in the reference."""
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        schedule = _ModulationSchedule(modulate_params if is_modulate else None)
        first, last = (0 if t_start is None else t_start), (num_sigmas if t_end is None else t_end)
        mask_hw = (28 if feature_height is None else feature_height, 52 if feature_width is None else feature_width)
        for i in range(num_sigmas - 1)[first:last + 1]:
            churn = self.s_tmin <= sigmas[i] <= self.s_tmax
            gamma = min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1) if churn else 0.0
            if modulate_params is not None:
                modulate_params["timestep"] = i
            modulate_now, inject_now = schedule.enter(i)
            if uc_list is not None:
                uc = uc_list[i]
            # reference :199-210: sampler steps 23 and 24 are hard-coded there
            smooth_now = bool(is_smooth_latent and i in (23, 24))
            smooth_step_size = {23: 1, 24: 2}.get(i) if smooth_now else None
            blend_mask = blend_xt = None
            if is_latent_blending and modulate_params["latent_mask_start"] <= i <= modulate_params["latent_mask_end"]:
                blend_xt = load_xt(modulate_params.get("feature_folder"), modulate_params.get("exp_name"), i, x.device,
                                   features=modulate_params.get("features")).to(x.dtype)
                masks = torch.stack(list(modulate_params["feature_masks"]), dim=0)
                blend_mask = masks.reshape(masks.shape[0], *mask_hw)
            x = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], denoiser, x, cond, uc, gamma,
                                  is_modulate_step=modulate_now, is_injected_step=inject_now,
                                  modulate_params=modulate_params, is_smooth_latent=smooth_now, model=model,
                                  smooth_step_size=smooth_step_size, blend_mask=blend_mask, blend_xt=blend_xt)
            if callback:
                callback(i)
            if img_callback and schedule.reports(i):
                img_callback(x, i)
        return x

    def inversion(self, denoiser, x, cond, uc=None, num_steps=None):
        """reference sampling.py:264-297: the same steps over the flipped schedule."""
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps, inversion=True)
        latents_list = [x]
        for i in self.get_sigma_gen(num_sigmas):
            gamma = min(self.s_churn / (num_sigmas - 1), 2**0.5 - 1) if self.s_tmin <= sigmas[i] <= self.s_tmax else 0.0
            x = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], denoiser, x, cond, uc, gamma)
            latents_list.append(x)
        x = x / torch.sqrt(1.0 + sigmas[-1] ** 2.0)
        return x, latents_list

    def null_text_optimization(self, *args, **kwargs):
        raise NotImplementedError("null-text optimisation back-propagates through the UNet; this build is inference only")


class EulerEDMSampler(EDMSampler):
    def possible_correction_step(self, euler_step, x, d, dt, next_sigma, denoiser, cond, uc):
        return euler_step
