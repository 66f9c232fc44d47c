"""GPU parity of the first stage against goldens written by the UNMODIFIED reference
modules (tests/golden/make_vae_goldens.py, make_sampler_goldens.py smooth) and against the oracle (oracle/vae.py):
``AutoencoderKL`` encode / decode, SVD's ``VideoDecoder``, the chunked ``encode_first_stage`` / ``decode_first_stage``
of the engine, and the sampler's ``is_smooth_latent`` steps that run through them.  Bar: 1e-3 (max|delta| / max|ref|)."""
import os

import numpy as np
import pytest
import torch

from oracle import unet as ounet, vae as ovae
from synth import synthetic_unet_inputs, synthetic_unet_weights

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL, SCALE = 1e-3, 0.18215


def relerr(got, want):
    got, want = torch.as_tensor(got).double().cpu(), torch.as_tensor(want).double().cpu()
    return float((got - want).abs().max() / want.abs().max())


def build_kl(seed, cuda, cfg=ovae.TINY_VAE_CONFIG):
    from vidseg_diffusion_b200.sgm.models.autoencoder import AutoencoderKL
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ovae.param_shapes(cfg), seed).items()}
    m = AutoencoderKL(embed_dim=4, monitor="val/rec_loss", ddconfig=dict(cfg, attn_type="vanilla"),
                      lossconfig={"target": "torch.nn.Identity"})
    m.load_state_dict(sd, strict=True)
    return m.to(cuda).eval(), sd


def test_autoencoder_kl_matches_reference_golden(cuda, operand_mode):
    from vidseg_diffusion_b200 import kernels as K
    from vidseg_diffusion_b200.sgm.models.autoencoder import DiagonalGaussianDistribution
    g = np.load(os.path.join(GOLDEN, "vae_tiny.npz"))
    seed, _ = (int(v) for v in g["meta"])
    model, sd = build_kl(seed, cuda)
    assert sorted(model.state_dict()) == list(g["keys"])
    x, noise = torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["noise"])
    with torch.no_grad():
        moments = K.conv2d(K.image_split(model.encoder(x)), model.quant_conv)
        z = SCALE * DiagonalGaussianDistribution(moments).sample(noise)
        img = model.decode(1.0 / SCALE * torch.from_numpy(g["z"]).to(cuda))
    errs = {"z": relerr(z, g["z"]), "image": relerr(img, g["image"])}
    # the oracle on this box, full tensors incl. the encoder moments
    m_or = torch.nn.functional.conv2d(ovae.encoder_forward(sd, ovae.TINY_VAE_CONFIG, torch.from_numpy(g["x"])),
                                      sd["quant_conv.weight"], sd["quant_conv.bias"])
    errs["moments_oracle"] = relerr(moments, m_or)
    print("vae:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) <= TOL, errs


def test_video_decoder_matches_reference_golden(cuda, operand_mode):
    from vidseg_diffusion_b200.sgm.modules.autoencoding.temporal_ae import VideoDecoder
    g = np.load(os.path.join(GOLDEN, "vae_video_tiny.npz"))
    seed, T = (int(v) for v in g["meta"])
    cfg = ovae.TINY_VAE_CONFIG
    shapes = ovae.video_decoder_param_shapes(cfg)
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, seed).items()}
    dec = VideoDecoder(**cfg, video_kernel_size=[3, 1, 1])
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items()}, strict=True)
    dec = dec.to(cuda).eval()
    z = torch.from_numpy(g["z"]).to(cuda)
    with torch.no_grad():
        img = dec(1.0 / SCALE * z, timesteps=T)
        img_1 = dec(1.0 / SCALE * z[:T], timesteps=T)                      # one clip alone: clips do not mix
        img_skip = dec(1.0 / SCALE * z, timesteps=T, skip_video=True) if False else None
    err = relerr(img, g["image"])
    print(f"video decoder: {err:.2e}")
    assert err <= TOL
    assert relerr(img_1, g["image"][:T]) <= TOL


def test_first_stage_chunking_matches_oracle(cuda):
    """decode_first_stage / encode_first_stage (diffusion.py:117-151) with en_and_decode_n_samples_a_time: for the
    VideoDecoder the chunk length is the clip length its temporal layers see, so chunked != unchunked and both must
    match the oracle run the same way."""
    from vidseg_diffusion_b200.sgm.models.autoencoder import AutoencodingEngine, DiagonalGaussianRegularizer
    from vidseg_diffusion_b200.sgm.models.diffusion import FirstStage
    from vidseg_diffusion_b200.sgm.modules.autoencoding.temporal_ae import VideoDecoder
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules.model import Encoder
    cfg = ovae.TINY_VAE_CONFIG
    shapes = dict(ovae.video_decoder_param_shapes(cfg))
    shapes.update({k: v for k, v in ovae.param_shapes(cfg).items() if k.startswith("encoder.")})
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, 21).items()}
    eng = AutoencodingEngine(encoder=Encoder(**cfg), decoder=VideoDecoder(**cfg, video_kernel_size=[3, 1, 1]),
                             regularization=DiagonalGaussianRegularizer(sample=False))
    eng.load_state_dict(sd, strict=True)
    eng = eng.to(cuda).eval()
    g = torch.Generator().manual_seed(21)
    z = torch.randn(5, 4, 4, 4, generator=g)
    for n in (None, 2):
        fs = FirstStage(eng, scale_factor=SCALE, en_and_decode_n_samples_a_time=n)
        img = fs.decode_first_stage(z.to(cuda))
        chunks = [z] if n is None else [z[i:i + n] for i in range(0, 5, n)]
        want = torch.cat([ovae.decode_first_stage(sd, cfg, c, SCALE, timesteps=c.shape[0]) for c in chunks], 0)
        assert relerr(img, want) <= TOL
        back = fs.encode_first_stage(img)
        want_z = SCALE * torch.chunk(ovae.encoder_forward(sd, cfg, img.cpu().float()), 2, dim=1)[0]
        assert relerr(back, want_z) <= TOL


def test_sampler_smooth_latent_matches_reference_golden(cuda):
    """EulerEDMSampler(is_smooth_latent=True): steps 23 and 24 decode the denoised latent, average every third frame with
    its neighbours and encode again (sampling.py:116-124) -- UNet, VAE and sampler all on the CUDA path, against the run of
    the unmodified reference (posterior noise from torch's CPU generator, seeded like the golden run)."""
    from test_gpu_unet import build
    from vidseg_diffusion_b200.sgm.models.diffusion import FirstStage
    from vidseg_diffusion_b200.sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    from vidseg_diffusion_b200.sgm.util import instantiate_from_config
    g = np.load(os.path.join(GOLDEN, "sampler_smooth_tiny.npz"))
    seed, F, hw, L, steps, t_start, noise_seed = (int(v) for v in g["meta"])
    cfg = ounet.TINY_CONFIG
    model, _ = build(cfg, seed, cuda)
    vae, _ = build_kl(seed + 1, cuda)
    engine = FirstStage(vae, scale_factor=SCALE)
    ddpm = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    smp = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler",
        "params": {"discretization_config": ddpm, "num_steps": steps, "s_churn": 0, "s_tmin": 0, "s_tmax": 999, "s_noise": 1,
                   "device": str(cuda),
                   "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": 5.0}}}})
    den = instantiate_from_config({
        "target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser",
        "params": {"num_idx": 1000, "discretization_config": ddpm,
                   "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}}}).to(cuda)
    denoiser = den.bind(OpenAIWrapper(model))
    x, _, ctx = synthetic_unet_inputs(seed, F, hw, cfg["in_channels"], L, cfg["context_dim"])
    latent = torch.from_numpy(x)[:F].contiguous().to(cuda)
    ctx = torch.from_numpy(ctx)[:F].contiguous().to(cuda)
    c, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}
    out_plain = smp(denoiser, latent.clone(), cond=c, uc=uc, t_start=t_start)
    torch.manual_seed(noise_seed)
    out_smooth = smp(denoiser, latent.clone(), cond=c, uc=uc, t_start=t_start, is_smooth_latent=True, model=engine)
    e1, e2 = relerr(out_plain, g["out_plain"]), relerr(out_smooth, g["out_smooth"])
    print(f"smooth sampler: plain {e1:.2e} smooth {e2:.2e}")
    assert e1 <= TOL and e2 <= TOL
