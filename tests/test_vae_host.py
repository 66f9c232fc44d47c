"""CPU: host-side checks of the first-stage row -- module trees / state-dict keys of the mirrors equal the reference's
(recorded in the goldens), the oracle sampler with ``is_smooth_latent`` reproduces the reference golden on the oracle
networks, and the mirrors refuse to run without a GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler as osamp, unet as ounet, vae as ovae
from synth import synthetic_unet_inputs, synthetic_unet_weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_mirror_state_dict_keys_equal_the_reference():
    from vidseg_diffusion_b200.sgm.models.autoencoder import AutoencoderKL
    from vidseg_diffusion_b200.sgm.modules.autoencoding.temporal_ae import VideoDecoder
    cfg = ovae.TINY_VAE_CONFIG
    kl = AutoencoderKL(embed_dim=4, ddconfig=dict(cfg), lossconfig={"target": "torch.nn.Identity"})
    g = np.load(os.path.join(GOLDEN, "vae_tiny.npz"))
    assert sorted(kl.state_dict()) == list(g["keys"])
    assert {k: tuple(v.shape) for k, v in kl.state_dict().items()} == ovae.param_shapes(cfg)
    vd = VideoDecoder(**cfg, video_kernel_size=[3, 1, 1])
    gv = np.load(os.path.join(GOLDEN, "vae_video_tiny.npz"))
    assert sorted("decoder." + k for k in vd.state_dict()) == list(gv["keys"])
    full = ovae.SD_VAE_CONFIG
    with torch.device("meta"):
        big = AutoencoderKL(embed_dim=4, ddconfig=dict(full))
    assert {k: tuple(v.shape) for k, v in big.state_dict().items()} == ovae.param_shapes(full)


def test_first_stage_fails_loudly_without_cuda():
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.sgm.models.autoencoder import AutoencoderKL
    kl = AutoencoderKL(embed_dim=4, ddconfig=dict(ovae.TINY_VAE_CONFIG)).eval()
    with pytest.raises(_lib.VidsegError):
        kl.decode(torch.zeros(1, 4, 4, 4))


def test_oracle_smooth_latent_reproduces_reference_golden():
    g = np.load(os.path.join(GOLDEN, "sampler_smooth_tiny.npz"))
    seed, F, hw, L, steps, t_start, noise_seed = (int(v) for v in g["meta"])
    cfg, vcfg = ounet.TINY_CONFIG, ovae.TINY_VAE_CONFIG
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ounet.param_shapes(cfg), seed).items()}
    vsd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(ovae.param_shapes(vcfg), seed + 1).items()}
    x, _, ctx = synthetic_unet_inputs(seed, F, hw, cfg["in_channels"], L, cfg["context_dim"])
    latent, ctx = torch.from_numpy(x)[:F].contiguous(), torch.from_numpy(ctx)[:F].contiguous()
    c, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}
    network = lambda x_in, cn, cond, **fl: ounet.unet_forward(sd, cfg, x_in, cn, cond["crossattn"])
    first_stage = (lambda z: ovae.decode_first_stage(vsd, vcfg, z, 0.18215),
                   lambda im: ovae.encode_first_stage(vsd, vcfg, im, 0.18215, torch.randn(im.shape[0], 4, im.shape[2] // 8, im.shape[3] // 8)))
    torch.manual_seed(noise_seed)
    out = osamp.euler_edm_sample(network, latent.clone(), c, uc, osamp.legacy_ddpm_sigmas(steps), osamp.eps_scaling, 5.0,
                                 osamp.make_discrete_quantizer(1000), t_start=t_start, is_smooth_latent=True, first_stage=first_stage)
    err = float((out - torch.from_numpy(g["out_smooth"])).abs().max() / np.abs(g["out_smooth"]).max())
    assert err < 1e-4, err
