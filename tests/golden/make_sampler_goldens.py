"""Pin oracle/sampler.py to the UNMODIFIED reference sampler and write the sampler goldens.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_sampler_goldens.py
The reference ``EulerEDMSampler`` + ``DiscreteDenoiser(EpsScaling)`` + ``VanillaCFG`` (configs/inference/sd_2_1.yaml:7-16,
63-79) drive the reference ``UNetModel`` (tiny width, seeded synthetic weights) wrapped in the reference
``OpenAIWrapper``, fp32 on the CPU, exactly as scripts/sampling/svd_single_video_inference.py:146-160, 322-330 does:
  run A  plain: 6-step schedule, t_start = 2 (steps 2..5); the img_callback writes the stashed q / k of output blocks 7
         and 8 and x_t as ``.pt`` files in the reference's feature_maps layout (svd_single_video_inference.py:113-130);
  run B  mask modulation at steps 3 and 4, feature injection from step 3 on, latent blending on steps 3..4, reading those files.
What is stored comes from the REFERENCE runs.  The oracle loop (oracle/sampler.py) with the same reference UNet as its
network must reproduce both runs bit for bit.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import sampler as osamp  # noqa: E402
from oracle import unet as ounet  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from synth import synthetic_modulate_params, synthetic_unet_inputs, synthetic_unet_weights  # noqa: E402

SEED, F, HW, L, STEPS, T_START = 3, 2, 16, 7, 6, 2
BLOCKS = (7, 8)


def sampler_modulate_params(seed, frames, tokens, **extra):
    """The request svd_single_video_inference.py:438-500 builds for one modulated run (values seeded)."""
    mp = synthetic_modulate_params(seed, frames, tokens)
    mp.update(modulate_timestep=[3, 4], modulate_timestep_frames={}, is_injected_features=True, modulate_lambda_start=60.0,
              modulate_lambda_end=25.0,
              injected_block_types=["output"], output_block_indices=list(BLOCKS), input_block_indices=[],
              injected_feature_types=["spatial_self_attn_q", "spatial_self_attn_k"], latent_mask_start=3, latent_mask_end=4)
    mp.update(extra)
    return mp


def main():
    om = import_reference("sgm.modules.diffusionmodules.openaimodel")
    rs = import_reference("sgm.modules.diffusionmodules.sampling")
    rd = import_reference("sgm.modules.diffusionmodules.denoiser")
    rw = import_reference("sgm.modules.diffusionmodules.wrappers")
    cfg = ounet.TINY_CONFIG
    model = om.UNetModel(use_checkpoint=False, use_linear_in_transformer=True, transformer_depth=1, **cfg).eval()
    shapes = ounet.param_shapes(cfg)
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, SEED).items()}
    model.load_state_dict(sd, strict=True)
    net = rw.OpenAIWrapper(model)
    ddpm = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    den = rd.DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                              num_idx=1000, discretization_config=ddpm)
    smp = rs.EulerEDMSampler(discretization_config=ddpm, num_steps=STEPS, device="cpu", s_churn=0.0, s_tmin=0.0, s_tmax=999.0,
                             s_noise=1.0, guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                                         "params": {"scale": 5.0}})
    x, _, ctx = synthetic_unet_inputs(SEED, F, HW, cfg["in_channels"], L, cfg["context_dim"])
    latent = torch.from_numpy(x)[:F].contiguous()   # the synthetic UNet batch is already CFG-doubled: one half here
    ctx = torch.from_numpy(ctx)[:F].contiguous()
    c, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx) + 0.05 * ctx.flip(0)}

    def denoiser(inp, sigma, cc, **kw):
        return den(net, inp, sigma, cc, **kw)

    with tempfile.TemporaryDirectory() as root, torch.no_grad():
        fm = os.path.join(root, "src", "feature_maps")
        os.makedirs(fm)
        per_step = {}

        def save_cb(xt, i):   # svd_single_video_inference.py:110-130
            for b in BLOCKS:
                tb = model.output_blocks[b][1].transformer_blocks[0]
                torch.save(tb.attn1.q.clone(), os.path.join(fm, f"output_block_{b}_spatial_self_attn_q_time_{i}.pt"))
                torch.save(tb.attn1.k.clone(), os.path.join(fm, f"output_block_{b}_spatial_self_attn_k_time_{i}.pt"))
            torch.save(xt.clone(), os.path.join(fm, f"xt_time_{i}.pt"))
            per_step[i] = xt.clone()

        out_a = smp(denoiser, latent.clone(), cond=c, uc=uc, img_callback=save_cb, t_start=T_START)
        mp = sampler_modulate_params(SEED, F, (HW // 2) ** 2, feature_folder=root, exp_name="src")
        mp_t = dict(mp, feature_masks=[torch.from_numpy(m) for m in mp["feature_masks"]])
        out_b = smp(denoiser, latent.clone(), cond=c, uc=uc, is_modulate=True, modulate_params=mp_t, t_start=T_START,
                    is_latent_blending=True, feature_height=HW // 2, feature_width=HW // 2)
        # float64 masks (what the reference script builds: numpy / 255.0): the blend is promoted to float64
        mp_64 = dict(mp, feature_masks=[torch.from_numpy(m.astype(np.float64) * 0.75) for m in mp["feature_masks"]])
        out_b64 = smp(denoiser, latent.clone(), cond=c, uc=uc, is_modulate=True, modulate_params=mp_64, t_start=T_START,
                      is_latent_blending=True, feature_height=HW // 2, feature_width=HW // 2)

        # ---- the oracle loop on the same network must agree bit for bit
        def network(x_in, c_noise, cond, **flags):
            return net(x_in, c_noise, cond, **flags)

        sig = osamp.legacy_ddpm_sigmas(STEPS)
        assert torch.equal(sig, smp.discretization(STEPS, device="cpu"))
        quant = osamp.make_discrete_quantizer(1000)
        o_a = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.eps_scaling, 5.0, quant, t_start=T_START)
        xt_store = {f"xt_time_{i}": v for i, v in per_step.items()}
        o_b = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.eps_scaling, 5.0, quant, t_start=T_START,
                                     is_modulate=True, modulate_params=dict(mp_t), is_latent_blending=True,
                                     feature_height=HW // 2, feature_width=HW // 2, xt_store=xt_store)
        o_b64 = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.eps_scaling, 5.0, quant, t_start=T_START,
                                       is_modulate=True, modulate_params=dict(mp_64), is_latent_blending=True,
                                       feature_height=HW // 2, feature_width=HW // 2, xt_store=xt_store)
        assert torch.equal(o_a, out_a) and torch.equal(o_b, out_b) and torch.equal(o_b64, out_b64)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        print("oracle == reference on runs A, B, B64; modulation+injection+blending changed the result by",
              f"{rel(out_b, out_a):.2e}; float64 masks by {rel(out_b64, out_b):.2e}; |x| max {float(out_a.abs().max()):.3f}")
        assert rel(out_b, out_a) > 1e-2
    np.savez_compressed(os.path.join(HERE, "sampler_tiny.npz"), latent=latent.numpy(), ctx=ctx.numpy(),
                        uctx=uc["crossattn"].numpy(), out_a=out_a.numpy(), out_b=out_b.numpy(), out_b64=out_b64.numpy(),
                        steps_a=np.stack([per_step[i].numpy() for i in sorted(per_step)]), sigmas=sig.numpy(),
                        meta=np.array([SEED, F, HW, L, STEPS, T_START]))


V_SEED, V_F, V_STEPS, V_T_START = 4, 3, 5, 2
V_TYPES = ["spatial_self_attn_q", "spatial_self_attn_k", "temporal_self_attn_q", "temporal_self_attn_k"]


def video_modulate_params(seed, frames, tokens, **extra):
    mp = synthetic_modulate_params(seed, frames, tokens)
    mp.update(modulate_layer_type=["spatial", "temporal"], modulate_attn_type=["self_attn", "cross_attn", "ff_out"],
              modulate_layer_frames={"temporal": [0, 2]}, modulate_lambda_start=40.0, modulate_lambda_end=15.0,
              modulate_timestep=[3], modulate_timestep_frames={}, is_injected_features=True,
              injected_block_types=["output"], output_block_indices=[7], input_block_indices=[],
              injected_feature_types=list(V_TYPES), latent_mask_start=3, latent_mask_end=3)
    mp.update(extra)
    return mp


def video_inputs(cfg):
    """Conditioning of one SVD clip as svd_single_video_inference.py builds it: CLIP image token (crossattn), the
    conditioning-frame latent (concat), fps / motion / cond_aug embeddings (vector); zeros in the unconditional branch."""
    from synth import synthetic_video_unet_inputs
    x, _, ctx, y = synthetic_video_unet_inputs(V_SEED, V_F, HW, cfg["in_channels"], cfg["context_dim"], cfg["adm_in_channels"])
    half = cfg["in_channels"] // 2
    latent = torch.from_numpy(x[V_F:, :half]).contiguous()
    c = {"crossattn": torch.from_numpy(ctx[V_F:]).contiguous(), "concat": torch.from_numpy(x[V_F:, half:]).contiguous(),
         "vector": torch.from_numpy(y[V_F:]).contiguous()}
    uc = {"crossattn": torch.zeros_like(c["crossattn"]), "concat": torch.zeros_like(c["concat"]), "vector": c["vector"].clone()}
    return latent, c, uc


def main_video():
    """SVD flavour (configs/inference/svd.yaml): Denoiser(VScalingWithEDMcNoise), EDM schedule with sigma_max = 700,
    LinearPredictionGuider over the frames, VideoUNet with its extra inputs (svd_single_video_inference.py:316-330)."""
    from make_video_unet_goldens import build_reference
    from oracle import video_unet as ov
    rs = import_reference("sgm.modules.diffusionmodules.sampling")
    rd = import_reference("sgm.modules.diffusionmodules.denoiser")
    rw = import_reference("sgm.modules.diffusionmodules.wrappers")
    cfg = ov.TINY_VIDEO_CONFIG
    model = build_reference(cfg)
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ov.param_shapes(cfg), V_SEED).items()}
    model.load_state_dict(sd, strict=True)
    net = rw.OpenAIWrapper(model)
    den = rd.Denoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise"})
    smp = rs.EulerEDMSampler(
        discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization", "params": {"sigma_max": 700.0}},
        num_steps=V_STEPS, device="cpu",
        guider_config={"target": "sgm.modules.diffusionmodules.guiders.LinearPredictionGuider",
                       "params": {"max_scale": 2.5, "min_scale": 1.0, "num_frames": V_F}})
    latent, c, uc = video_inputs(cfg)
    extra = dict(image_only_indicator=torch.zeros(2, V_F), num_video_frames=V_F)

    def denoiser(inp, sigma, cc, **kw):
        return den(net, inp, sigma, cc, **kw, **extra)

    with tempfile.TemporaryDirectory() as root, torch.no_grad():
        fm = os.path.join(root, "src", "feature_maps")
        os.makedirs(fm)
        per_step = {}

        def save_cb(xt, i):
            layer = model.output_blocks[7][1]
            for ft in V_TYPES:
                blk = layer.transformer_blocks[0] if ft.startswith("spatial") else layer.time_stack[0]
                torch.save(getattr(blk.attn1, ft[-1]).clone(), os.path.join(fm, f"output_block_7_{ft}_time_{i}.pt"))
            torch.save(xt.clone(), os.path.join(fm, f"xt_time_{i}.pt"))
            per_step[i] = xt.clone()

        out_a = smp(denoiser, latent.clone(), cond=c, uc=uc, img_callback=save_cb, t_start=V_T_START)
        mp = video_modulate_params(V_SEED, V_F, (HW // 2) ** 2, feature_folder=root, exp_name="src")
        mp_t = dict(mp, feature_masks=[torch.from_numpy(m) for m in mp["feature_masks"]])
        out_b = smp(denoiser, latent.clone(), cond=c, uc=uc, is_modulate=True, modulate_params=mp_t, t_start=V_T_START,
                    is_latent_blending=True, feature_height=HW // 2, feature_width=HW // 2)

        def network(x_in, c_noise, cond, **flags):
            return net(x_in, c_noise, cond, **flags, **extra)

        sig = osamp.edm_sigmas(V_STEPS, sigma_max=700.0)
        assert torch.equal(sig, smp.discretization(V_STEPS, device="cpu"))
        fs = torch.linspace(1.0, 2.5, V_F)
        o_a = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.v_scaling_edm_cnoise, None, None,
                                     t_start=V_T_START, frame_scales=fs)
        o_b = osamp.euler_edm_sample(network, latent.clone(), c, uc, sig, osamp.v_scaling_edm_cnoise, None, None,
                                     t_start=V_T_START, frame_scales=fs, is_modulate=True, modulate_params=dict(mp_t),
                                     is_latent_blending=True, feature_height=HW // 2, feature_width=HW // 2,
                                     xt_store={f"xt_time_{i}": v for i, v in per_step.items()})
        assert torch.equal(o_a, out_a) and torch.equal(o_b, out_b)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        print("video: oracle == reference on runs A, B; modulation+injection+blending changed the result by",
              f"{rel(out_b, out_a):.2e}; |x| max {float(out_a.abs().max()):.3f}")
        assert rel(out_b, out_a) > 1e-2
    np.savez_compressed(os.path.join(HERE, "sampler_video_tiny.npz"), out_a=out_a.numpy(), out_b=out_b.numpy(),
                        steps_a=np.stack([per_step[i].numpy() for i in sorted(per_step)]), sigmas=sig.numpy(),
                        meta=np.array([V_SEED, V_F, HW, V_STEPS, V_T_START]))


S_SEED, S_F, S_STEPS, S_T_START, S_SCALE, S_NOISE_SEED = 8, 5, 25, 22, 0.18215, 123


def main_smooth():
    """``is_smooth_latent`` (sampling.py:116-124, 199-210): the reference sampler on the last three steps of a 25-step
    schedule with a ``model`` whose ``decode_first_stage`` / ``encode_first_stage`` are the reference ``DiffusionEngine``
    methods (diffusion.py:117-151) over the reference ``AutoencoderKL`` (tiny width, seeded weights).  The posterior noise
    of the re-encoding comes from torch's global CPU generator, seeded right before the run."""
    from oracle import vae as ovae
    om = import_reference("sgm.modules.diffusionmodules.openaimodel")
    rs = import_reference("sgm.modules.diffusionmodules.sampling")
    rd = import_reference("sgm.modules.diffusionmodules.denoiser")
    rw = import_reference("sgm.modules.diffusionmodules.wrappers")
    ra = import_reference("sgm.models.autoencoder")
    rdiff = import_reference("sgm.models.diffusion")
    cfg = ounet.TINY_CONFIG
    model = om.UNetModel(use_checkpoint=False, use_linear_in_transformer=True, transformer_depth=1, **cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ounet.param_shapes(cfg), S_SEED).items()}
    model.load_state_dict(sd, strict=True)
    net = rw.OpenAIWrapper(model)
    vcfg = ovae.TINY_VAE_CONFIG
    vae = ra.AutoencoderKL(embed_dim=4, monitor="val/rec_loss", ddconfig=dict(vcfg, attn_type="vanilla"),
                           lossconfig={"target": "torch.nn.Identity"}).eval()
    vsd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ovae.param_shapes(vcfg), S_SEED + 1).items()}
    vae.load_state_dict(vsd, strict=True)

    class Engine:   # the attributes the two DiffusionEngine methods read
        scale_factor, en_and_decode_n_samples_a_time, disable_first_stage_autocast, first_stage_model = S_SCALE, None, True, vae
        decode_first_stage = rdiff.DiffusionEngine.decode_first_stage
        encode_first_stage = rdiff.DiffusionEngine.encode_first_stage

    ddpm = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    den = rd.DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                              num_idx=1000, discretization_config=ddpm)
    smp = rs.EulerEDMSampler(discretization_config=ddpm, num_steps=S_STEPS, device="cpu", s_churn=0.0, s_tmin=0.0, s_tmax=999.0,
                             s_noise=1.0, guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                                         "params": {"scale": 5.0}})
    x, _, ctx = synthetic_unet_inputs(S_SEED, S_F, HW, cfg["in_channels"], L, cfg["context_dim"])
    latent = torch.from_numpy(x)[:S_F].contiguous()
    ctx = torch.from_numpy(ctx)[:S_F].contiguous()
    c, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}

    def denoiser(inp, sigma, cc, **kw):
        return den(net, inp, sigma, cc, **kw)

    with torch.no_grad():
        out_plain = smp(denoiser, latent.clone(), cond=c, uc=uc, t_start=S_T_START)
        torch.manual_seed(S_NOISE_SEED)
        out_smooth = smp(denoiser, latent.clone(), cond=c, uc=uc, t_start=S_T_START, is_smooth_latent=True, model=Engine())
        sig = osamp.legacy_ddpm_sigmas(S_STEPS)
        quant = osamp.make_discrete_quantizer(1000)
        first_stage = (lambda z: ovae.decode_first_stage(vsd, vcfg, z, S_SCALE),
                       lambda im: ovae.encode_first_stage(vsd, vcfg, im, S_SCALE, torch.randn(im.shape[0], 4, im.shape[2] // 8, im.shape[3] // 8)))
        torch.manual_seed(S_NOISE_SEED)
        o_smooth = osamp.euler_edm_sample(lambda x_in, cn, cond, **fl: net(x_in, cn, cond, **fl), latent.clone(), c, uc, sig,
                                          osamp.eps_scaling, 5.0, quant, t_start=S_T_START, is_smooth_latent=True,
                                          first_stage=first_stage)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print("smooth: oracle vs reference", f"{rel(o_smooth, out_smooth):.2e}", "| smoothing changed the result by",
          f"{rel(out_smooth, out_plain):.2e}; |x| max {float(out_smooth.abs().max()):.3f}")
    assert rel(o_smooth, out_smooth) < 2e-5 and rel(out_smooth, out_plain) > 1e-2
    np.savez_compressed(os.path.join(HERE, "sampler_smooth_tiny.npz"), out_plain=out_plain.numpy(), out_smooth=out_smooth.numpy(),
                        meta=np.array([S_SEED, S_F, HW, L, S_STEPS, S_T_START, S_NOISE_SEED]))


if __name__ == "__main__":
    which = sys.argv[1:] or ["sd", "video", "smooth"]
    if "smooth" in which:
        main_smooth()
    if "sd" in which:
        main()
    if "video" in which:
        main_video()
