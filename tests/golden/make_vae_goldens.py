"""Pin oracle/vae.py to the UNMODIFIED reference Encoder / Decoder and write the first-stage goldens
(oracle first, the CUDA path follows).

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_vae_goldens.py
The reference ``sgm.modules.diffusionmodules.model.Encoder`` / ``Decoder`` are built with the sd_2_1.yaml ddconfig at toy
width, loaded with seeded synthetic weights and run in fp32 on the CPU; quant_conv / post_quant_conv, the Gaussian
posterior and the scale factor follow sgm/models/autoencoder.py:440-506 and sgm/models/diffusion.py:117-151 with the
same torch calls.  Stored from the REFERENCE modules: encoder moments, decoded image; the oracle must agree to 2e-5.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import vae as ovae  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from synth import synthetic_unet_weights  # noqa: E402

SEED, B, SCALE = 6, 2, 0.18215


def main():
    rm = import_reference("sgm.modules.diffusionmodules.model")
    cfg = ovae.TINY_VAE_CONFIG
    enc, dec = rm.Encoder(**cfg).eval(), rm.Decoder(**cfg).eval()
    shapes = ovae.param_shapes(cfg)
    ref_shapes = {f"encoder.{k}": tuple(v.shape) for k, v in enc.state_dict().items()}
    ref_shapes.update({f"decoder.{k}": tuple(v.shape) for k, v in dec.state_dict().items()})
    ours = {k: v for k, v in shapes.items() if not k.startswith(("quant_conv", "post_quant_conv"))}
    assert ref_shapes == ours, sorted(set(ref_shapes) ^ set(ours))[:10]
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, SEED).items()}
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(B, 3, cfg["resolution"], cfg["resolution"], generator=g)
    noise = torch.randn(B, 4, cfg["resolution"] // 8, cfg["resolution"] // 8, generator=g)
    F = torch.nn.functional
    with torch.no_grad():
        moments = F.conv2d(enc(x), sd["quant_conv.weight"], sd["quant_conv.bias"])
        mean, logvar = torch.chunk(moments, 2, dim=1)
        z = SCALE * (mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise)
        img = dec(F.conv2d(1.0 / SCALE * z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
        z_or = ovae.encode_first_stage(sd, cfg, x, SCALE, noise)
        img_or = ovae.decode_first_stage(sd, cfg, z, SCALE)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print("oracle vs reference: z", f"{rel(z_or, z):.2e}", "image", f"{rel(img_or, img):.2e}", "| z absmax", float(z.abs().max()),
          "image absmax", float(img.abs().max()))
    assert rel(z_or, z) < 2e-5 and rel(img_or, img) < 2e-5
    np.savez_compressed(os.path.join(HERE, "vae_tiny.npz"), x=x.numpy(), noise=noise.numpy(), z=z.numpy(), image=img.numpy(),
                        keys=np.array(sorted(shapes)), meta=np.array([SEED, B]))


def main_video():
    """SVD's VideoDecoder (svd.yaml: temporal_ae.VideoDecoder, video_kernel_size [3,1,1], time_mode "conv-only")."""
    ta = import_reference("sgm.modules.autoencoding.temporal_ae")
    cfg = ovae.TINY_VAE_CONFIG
    T = 3
    dec = ta.VideoDecoder(**cfg, video_kernel_size=[3, 1, 1]).eval()
    shapes = ovae.video_decoder_param_shapes(cfg)
    ref_shapes = {f"decoder.{k}": tuple(v.shape) for k, v in dec.state_dict().items()}
    assert ref_shapes == shapes, sorted(set(ref_shapes) ^ set(shapes))[:10]
    sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, SEED + 1).items()}
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items()}, strict=True)
    g = torch.Generator().manual_seed(SEED + 1)
    z = torch.randn(2 * T, 4, cfg["resolution"] // 8, cfg["resolution"] // 8, generator=g)
    with torch.no_grad():
        img = dec(1.0 / SCALE * z, timesteps=T)
        img_or = ovae.decode_first_stage(sd, cfg, z, SCALE, timesteps=T)
        img_flat = ovae.decode_first_stage({k: v for k, v in sd.items()}, cfg, z, SCALE, timesteps=1)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print("video decoder: oracle vs reference", f"{rel(img_or, img):.2e}", "| frames mix: changed the image by",
          f"{rel(img_flat, img):.2e}", "| image absmax", float(img.abs().max()))
    assert rel(img_or, img) < 2e-5 and rel(img_flat, img) > 1e-3
    np.savez_compressed(os.path.join(HERE, "vae_video_tiny.npz"), z=z.numpy(), image=img.numpy(), keys=np.array(sorted(shapes)),
                        meta=np.array([SEED + 1, T]))


if __name__ == "__main__":
    which = sys.argv[1:] or ["image", "video"]
    if "image" in which:
        main()
    if "video" in which:
        main_video()
