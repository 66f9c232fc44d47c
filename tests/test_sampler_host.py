"""Sampler row without a GPU: the oracle loop with the oracle UNet against the goldens the
reference sampler produced (tests/golden/make_sampler_goldens.py), and the host mirror's schedules / configuration."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler as osamp
from oracle import unet as ounet
from synth import synthetic_unet_weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sampler_tiny.npz")
BLOCKS = (7, 8)


def modulate_params_for(seed, frames, tokens, **extra):
    from golden.make_sampler_goldens import sampler_modulate_params
    return sampler_modulate_params(seed, frames, tokens, **extra)


def relerr(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max())


def test_oracle_loop_with_oracle_unet_matches_reference_goldens():
    g = np.load(GOLDEN)
    seed, F, hw, L, steps, t_start = (int(v) for v in g["meta"])
    cfg = ounet.TINY_CONFIG
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ounet.param_shapes(cfg), seed).items()}
    latent, ctx, uctx = (torch.from_numpy(g[k]) for k in ("latent", "ctx", "uctx"))
    c, uc = {"crossattn": ctx}, {"crossattn": uctx}
    store = {}
    last = {}

    def network(x_in, c_noise, cond, is_modulate_step=False, is_injected_step=False, modulate_params=None):
        inj = None
        if is_injected_step:
            mp = modulate_params
            inj = dict(block_types=mp["injected_block_types"], input_block_indices=mp["input_block_indices"],
                       output_block_indices=mp["output_block_indices"], feature_types=mp["injected_feature_types"],
                       timestep=mp["timestep"], features=store)
        last.clear()
        return ounet.unet_forward(sd, cfg, x_in, c_noise, cond["crossattn"], last,
                                  modulate_params=modulate_params if is_modulate_step else None, injection=inj)

    def save_cb(xt, i):
        for b in BLOCKS:
            for n in ("q", "k"):
                store[f"output_block_{b}_spatial_self_attn_{n}_time_{i}"] = last[(f"output_block_{b}", f"spatial_self_attn_{n}")].clone()
        store[f"xt_time_{i}"] = xt.clone()

    sig = osamp.legacy_ddpm_sigmas(steps)
    assert np.array_equal(sig.numpy(), g["sigmas"])
    quant = osamp.make_discrete_quantizer(1000)
    out_a = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.eps_scaling, 5.0, quant, t_start=t_start,
                                   img_callback=save_cb)
    assert relerr(out_a, g["out_a"]) < 2e-5
    for n, i in enumerate(range(t_start, steps)):
        assert relerr(store[f"xt_time_{i}"], g["steps_a"][n]) < 2e-5
    mp = modulate_params_for(seed, F, (hw // 2) ** 2)
    for key, masks in (("out_b", [torch.from_numpy(m) for m in mp["feature_masks"]]),
                       ("out_b64", [torch.from_numpy(m.astype(np.float64) * 0.75) for m in mp["feature_masks"]])):
        out = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.eps_scaling, 5.0, quant, t_start=t_start,
                                     is_modulate=True, modulate_params=dict(mp, feature_masks=masks), is_latent_blending=True,
                                     feature_height=hw // 2, feature_width=hw // 2, xt_store=store)
        assert relerr(out, g[key]) < 5e-5, key
        assert relerr(out, g["out_a"]) > 1e-2


def test_host_mirror_builds_from_the_reference_config_and_matches_the_schedules():
    from vidseg_diffusion_b200.sgm.util import instantiate_from_config
    ddpm = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    smp = instantiate_from_config({   # configs/inference/sd_2_1.yaml:63-79
        "target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler",
        "params": {"discretization_config": ddpm, "num_steps": 6, "s_churn": 0, "s_tmin": 0, "s_tmax": 999, "s_noise": 1,
                   "device": "cpu",
                   "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": 5}}}})
    assert type(smp).__module__.startswith("vidseg_diffusion_b200.")
    g = np.load(GOLDEN)
    assert np.array_equal(smp.discretization(6, device="cpu").numpy(), g["sigmas"])
    den = instantiate_from_config({   # sd_2_1.yaml:7-16
        "target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser",
        "params": {"num_idx": 1000, "discretization_config": ddpm,
                   "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}}})
    assert torch.equal(den.sigmas, osamp.legacy_ddpm_sigmas(1000, append_zero=False, flip=True))
    s = torch.tensor([14.0, 0.7, 0.03])
    assert torch.equal(den.possibly_quantize_sigma(s), osamp.make_discrete_quantizer(1000)[0](s))
    # SVD: EDM schedule with sigma_max 700, linear guidance over the frames (configs/inference/svd.yaml)
    edm = instantiate_from_config({"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization",
                                   "params": {"sigma_max": 700.0}})
    assert torch.equal(edm(25), osamp.edm_sigmas(25, sigma_max=700.0))
    lin = instantiate_from_config({"target": "sgm.modules.diffusionmodules.guiders.LinearPredictionGuider",
                                   "params": {"max_scale": 2.5, "min_scale": 1.0, "num_frames": 14}})
    sc = lin.sample_scales(28, "cpu")
    assert sc.shape == (28,) and torch.equal(sc[:14], torch.linspace(1.0, 2.5, 14)) and torch.equal(sc[14:], sc[:14])
    x = torch.randn(56, 4, 3, 3)
    want = lin(x, None)
    x_u, x_c = x.chunk(2)
    assert torch.equal(want, x_u + sc[:, None, None, None] * (x_c - x_u))


def test_sampler_step_fails_loudly_without_cuda():
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules.sampling import fused_step
    x = torch.zeros(2, 4, 8, 8)
    one = torch.ones(2)
    with pytest.raises(_lib.VidsegError):
        fused_step(x, x, one, one, None, one, one)


VGOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sampler_video_tiny.npz")
V_TYPES = ["spatial_self_attn_q", "spatial_self_attn_k", "temporal_self_attn_q", "temporal_self_attn_k"]


def video_case():
    from golden import make_sampler_goldens as mk
    from oracle import video_unet as ov
    return mk, ov, ov.TINY_VIDEO_CONFIG


def test_oracle_loop_with_oracle_video_unet_matches_reference_goldens():
    """SVD flavour: VScalingWithEDMcNoise, EDM schedule (sigma_max 700), LinearPredictionGuider, VideoUNet inputs."""
    mk, ov, cfg = video_case()
    g = np.load(VGOLDEN)
    seed, F, hw, steps, t_start = (int(v) for v in g["meta"])
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ov.param_shapes(cfg), seed).items()}
    latent, c, uc = mk.video_inputs(cfg)
    ind = torch.zeros(2, F)
    store, last = {}, {}

    def network(x_in, c_noise, cond, is_modulate_step=False, is_injected_step=False, modulate_params=None):
        inj = None
        if is_injected_step:
            mp = modulate_params
            inj = dict(block_types=mp["injected_block_types"], input_block_indices=mp["input_block_indices"],
                       output_block_indices=mp["output_block_indices"], feature_types=mp["injected_feature_types"],
                       timestep=mp["timestep"], features=store)
        last.clear()
        x_cat = torch.cat((x_in, cond["concat"]), dim=1)   # OpenAIWrapper (wrappers.py:24-34)
        return ov.video_unet_forward(sd, cfg, x_cat, c_noise, cond["crossattn"], cond["vector"], F, ind, last,
                                     modulate_params=modulate_params if is_modulate_step else None, injection=inj)

    def save_cb(xt, i):
        for ft in V_TYPES:
            store[f"output_block_7_{ft}_time_{i}"] = last[("output_block_7", ft)].clone()
        store[f"xt_time_{i}"] = xt.clone()

    sig = osamp.edm_sigmas(steps, sigma_max=700.0)
    assert np.array_equal(sig.numpy(), g["sigmas"])
    fs = torch.linspace(1.0, 2.5, F)
    out_a = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.v_scaling_edm_cnoise, None, None,
                                   t_start=t_start, frame_scales=fs, img_callback=save_cb)
    assert relerr(out_a, g["out_a"]) < 5e-5
    mp = mk.video_modulate_params(seed, F, (hw // 2) ** 2)
    mp["feature_masks"] = [torch.from_numpy(m) for m in mp["feature_masks"]]
    out_b = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.v_scaling_edm_cnoise, None, None,
                                   t_start=t_start, frame_scales=fs, is_modulate=True, modulate_params=mp,
                                   is_latent_blending=True, feature_height=hw // 2, feature_width=hw // 2, xt_store=store)
    assert relerr(out_b, g["out_b"]) < 5e-5 and relerr(out_b, g["out_a"]) > 1e-2


@pytest.mark.parametrize("flavour", ["sd", "svd"])
def test_denoiser_mirror_equals_the_oracle_on_cpu(flavour):
    """Denoiser.forward / .raw are plain torch around the network (the kernels sit inside the network and the fused
    step): with a toy network they run on the CPU and must equal oracle.denoise bit for bit, quantisation included."""
    from vidseg_diffusion_b200.sgm.util import instantiate_from_config
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 6, 6, generator=g)
    sigma = torch.tensor([14.1, 3.3, 0.9, 0.05])
    seen = {}

    def network(x_in, c_noise, cond, **kw):
        seen["c_noise"], seen["kw"] = c_noise, kw
        return torch.tanh(x_in) * cond["w"] + 0.01 * c_noise.float()[:, None, None, None]

    cond = {"w": torch.randn(4, 1, 1, 1, generator=g)}
    if flavour == "sd":
        ddpm = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
        den = instantiate_from_config({"target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser",
                                       "params": {"num_idx": 1000, "discretization_config": ddpm, "scaling_config": {
                                           "target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}}})
        want = osamp.denoise(network, x, sigma, cond, osamp.eps_scaling, osamp.make_discrete_quantizer(1000))
    else:
        den = instantiate_from_config({"target": "sgm.modules.diffusionmodules.denoiser.Denoiser", "params": {
            "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise"}}})
        want = osamp.denoise(network, x, sigma, cond, osamp.v_scaling_edm_cnoise)
    got = den(network, x, sigma, cond, is_modulate_step=True)
    assert torch.equal(got, want)
    assert seen["kw"]["is_modulate_step"] is True and seen["kw"]["is_injected_step"] is False
    if flavour == "sd":
        assert seen["c_noise"].dtype == torch.int64       # the UNet's timestep index
    bound = den.bind(network)
    net, c_skip, c_out = bound.raw(x, sigma, cond)
    assert c_skip.shape == c_out.shape == (4,)
    assert torch.equal(net * c_out[:, None, None, None] + x * c_skip[:, None, None, None], want)
    assert torch.equal(bound(x, sigma, cond), want)


def test_load_xt_and_target_features_read_the_reference_layout_or_the_hbm_store(tmp_path):
    """sgm/util.py:277-311: ``{folder}/{exp}/feature_maps/{key}.pt``; the ``features`` dict replaces the files."""
    from vidseg_diffusion_b200.sgm.util import load_target_features, load_xt
    fm = tmp_path / "exp" / "feature_maps"
    fm.mkdir(parents=True)
    xt, q = torch.randn(2, 4, 8, 8), torch.randn(4, 64, 32)
    torch.save(xt, fm / "xt_time_7.pt")
    torch.save(q, fm / "output_block_8_spatial_self_attn_q_time_7.pt")
    assert torch.equal(load_xt(str(tmp_path), "exp", 7, "cpu"), xt)
    got = load_target_features(str(tmp_path), "exp", 7, "output", ["spatial_self_attn_q", "spatial_self_attn_k"], 8, "cpu")
    assert list(got) == ["output_block_8_spatial_self_attn_q_time_7"] and torch.equal(got[list(got)[0]], q)
    store = {"xt_time_7": xt + 1, "output_block_8_spatial_self_attn_q_time_7": q + 1}
    assert torch.equal(load_xt(None, None, 7, "cpu", features=store), xt + 1)
    assert torch.equal(load_target_features(None, None, 7, "output", ["spatial_self_attn_q"], 8, "cpu", features=store)
                       ["output_block_8_spatial_self_attn_q_time_7"], q + 1)
    for call in (lambda: load_xt(str(tmp_path), "exp", 8, "cpu"), lambda: load_xt(None, None, 8, "cpu", features=store),
                 lambda: load_target_features(str(tmp_path), "exp", 7, "input", ["spatial_self_attn_q"], 8, "cpu"),
                 lambda: load_target_features(None, None, 9, "output", ["spatial_self_attn_q"], 8, "cpu", features=store)):
        with pytest.raises(ValueError):
            call()


def test_sampler_options_outside_the_scope_fail_loudly():
    from vidseg_diffusion_b200.sgm.util import instantiate_from_config
    smp = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler",
        "params": {"num_steps": 4, "device": "cpu", "discretization_config": {
            "target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization"}}})
    x = torch.zeros(2, 4, 8, 8)
    with pytest.raises(ValueError):   # the smoothing pass needs the first stage
        smp.sampler_step(torch.ones(2), torch.ones(2), None, x, {}, {}, is_smooth_latent=True)
    with pytest.raises(NotImplementedError):
        smp.null_text_optimization()
    with pytest.raises(KeyError):
        instantiate_from_config({"params": {}})
    with pytest.raises(AttributeError):
        instantiate_from_config({"target": "sgm.modules.diffusionmodules.sampling.HeunEDMSampler", "params": {}})


def _eager_step(x, net, c_skip, c_out, scales, sigma_hat, sigma_next, mask=None, ori_xt=None):
    """The definition of vidseg_sampler_step in torch (the chain the kernel replaces), for host-logic tests on the CPU."""
    ap = lambda t: t[(...,) + (None,) * 3]
    inp = torch.cat([x] * 2) if scales is not None else x
    den = net * ap(c_out) + inp * ap(c_skip)
    if scales is not None:
        x_u, x_c = den.chunk(2)
        den = x_u + ap(scales) * (x_c - x_u)
    out = x + ap(sigma_next - sigma_hat) * ((x - den) / ap(sigma_hat))
    if mask is not None:
        fm = torch.nn.functional.interpolate(mask.unsqueeze(1), size=x.shape[-2:], mode="nearest")
        out = (out * fm + ori_xt.to(out.dtype) * (1 - fm)).float()
    return out


def test_sampler_host_logic_on_cpu_with_the_step_kernel_replaced_by_its_definition(monkeypatch):
    """The mirror's control flow (step range, modulation / injection switches, feature store, blending arguments, bound
    and opaque denoisers) end to end on the CPU: the oracle UNet is the network and the fused step is monkeypatched with
    its eager definition.  Must land on the reference sampler's goldens."""
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules import sampling
    from vidseg_diffusion_b200.sgm.util import instantiate_from_config
    monkeypatch.setattr(sampling, "fused_step", _eager_step)
    g = np.load(GOLDEN)
    seed, F, hw, L, steps, t_start = (int(v) for v in g["meta"])
    cfg = ounet.TINY_CONFIG
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ounet.param_shapes(cfg), seed).items()}
    ddpm = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    smp = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler",
        "params": {"discretization_config": ddpm, "num_steps": steps, "s_churn": 0, "s_tmin": 0, "s_tmax": 999, "s_noise": 1,
                   "device": "cpu",
                   "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": 5.0}}}})
    den = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser",
        "params": {"num_idx": 1000, "discretization_config": ddpm,
                   "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}}})
    store, last = {}, {}

    class Net(torch.nn.Module):   # stands in for OpenAIWrapper(UNetModel): same call signature
        def forward(self, x_in, c_noise, cond, is_modulate_step=False, is_injected_step=False, modulate_params=None):
            inj = None
            if is_injected_step:
                mp = modulate_params
                inj = dict(block_types=mp["injected_block_types"], input_block_indices=mp["input_block_indices"],
                           output_block_indices=mp["output_block_indices"], feature_types=mp["injected_feature_types"],
                           timestep=mp["timestep"], features=mp["features"])
            last.clear()
            return ounet.unet_forward(sd, cfg, x_in, c_noise, cond["crossattn"], last,
                                      modulate_params=modulate_params if is_modulate_step else None, injection=inj)

    def save_cb(xt, i):
        for b in BLOCKS:
            for n in ("q", "k"):
                store[f"output_block_{b}_spatial_self_attn_{n}_time_{i}"] = last[(f"output_block_{b}", f"spatial_self_attn_{n}")].clone()
        store[f"xt_time_{i}"] = xt.clone()

    latent, ctx, uctx = (torch.from_numpy(g[k]) for k in ("latent", "ctx", "uctx"))
    c, uc = {"crossattn": ctx}, {"crossattn": uctx}
    net = Net()
    bound = den.bind(net)
    out_a = smp(bound, latent.clone(), cond=c, uc=uc, img_callback=save_cb, t_start=t_start)
    assert relerr(out_a, g["out_a"]) < 2e-5
    opaque = smp(lambda inp, sigma, cc, **kw: den(net, inp, sigma, cc, **kw), latent.clone(), cond=c, uc=uc, t_start=t_start)
    assert torch.equal(opaque, out_a)
    mp = modulate_params_for(seed, F, (hw // 2) ** 2, features=store)
    for key, masks in (("out_b", [torch.from_numpy(m) for m in mp["feature_masks"]]),
                       ("out_b64", [torch.from_numpy(m.astype(np.float64) * 0.75) for m in mp["feature_masks"]])):
        out = smp(bound, latent.clone(), cond=c, uc=uc, is_modulate=True, modulate_params=dict(mp, feature_masks=masks),
                  t_start=t_start, is_latent_blending=True, feature_height=hw // 2, feature_width=hw // 2)
        assert relerr(out, g[key]) < 5e-5, key


def test_sampler_host_logic_on_cpu_svd_flavour(monkeypatch):
    """Same as above for the SVD configuration: Denoiser + VScalingWithEDMcNoise, Karras schedule, per-frame guidance ramp,
    the conditioning adapter (OpenAIWrapper: "concat" / "vector" entries) and the extra model inputs fixed by ``bind``."""
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules import sampling
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    from vidseg_diffusion_b200.sgm.util import instantiate_from_config
    monkeypatch.setattr(sampling, "fused_step", _eager_step)
    mk, ov, cfg = video_case()
    g = np.load(VGOLDEN)
    seed, F, hw, steps, t_start = (int(v) for v in g["meta"])
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ov.param_shapes(cfg), seed).items()}
    store, last = {}, {}

    class VideoNet(torch.nn.Module):   # stands in for VideoUNet: the signature OpenAIWrapper calls
        def forward(self, x, timesteps=None, context=None, y=None, num_video_frames=None, image_only_indicator=None,
                    is_modulate_step=False, is_injected_step=False, modulate_params=None):
            inj = None
            if is_injected_step:
                mp = modulate_params
                inj = dict(block_types=mp["injected_block_types"], input_block_indices=mp["input_block_indices"],
                           output_block_indices=mp["output_block_indices"], feature_types=mp["injected_feature_types"],
                           timestep=mp["timestep"], features=mp["features"])
            last.clear()
            return ov.video_unet_forward(sd, cfg, x, timesteps, context, y, num_video_frames, image_only_indicator, last,
                                         modulate_params=modulate_params if is_modulate_step else None, injection=inj)

    smp = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler",
        "params": {"num_steps": steps, "device": "cpu",
                   "discretization_config": {"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization",
                                             "params": {"sigma_max": 700.0}},
                   "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.LinearPredictionGuider",
                                     "params": {"max_scale": 2.5, "min_scale": 1.0, "num_frames": F}}}})
    den = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.denoiser.Denoiser",
        "params": {"scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise"}}})
    bound = den.bind(OpenAIWrapper(VideoNet()), image_only_indicator=torch.zeros(2, F), num_video_frames=F)
    latent, c, uc = mk.video_inputs(cfg)

    def save_cb(xt, i):
        for ft in V_TYPES:
            store[f"output_block_7_{ft}_time_{i}"] = last[("output_block_7", ft)].clone()
        store[f"xt_time_{i}"] = xt.clone()

    out_a = smp(bound, latent.clone(), cond=c, uc=uc, img_callback=save_cb, t_start=t_start)
    assert relerr(out_a, g["out_a"]) < 5e-5
    mp = mk.video_modulate_params(seed, F, (hw // 2) ** 2, features=store)
    mp["feature_masks"] = [torch.from_numpy(m) for m in mp["feature_masks"]]
    out_b = smp(bound, latent.clone(), cond=c, uc=uc, is_modulate=True, modulate_params=mp, t_start=t_start,
                is_latent_blending=True, feature_height=hw // 2, feature_width=hw // 2)
    assert relerr(out_b, g["out_b"]) < 5e-5
