"""The first-stage side of the reference's ``DiffusionEngine`` (sgm/models/diffusion.py:117-151).

Only what the per-clip path touches: ``decode_first_stage`` / ``encode_first_stage`` with ``scale_factor`` and the
``en_and_decode_n_samples_a_time`` chunking (for the ``VideoDecoder`` the chunk length IS the number of frames its
temporal convolutions see, :127-131, so it is part of the result, not a memory knob).  The sampler reaches them through
``model.decode_first_stage`` / ``model.encode_first_stage`` when ``is_smooth_latent`` is on (sampling.py:116-124).
"""
import math

import torch
import torch.nn as nn

from ..modules.autoencoding.temporal_ae import VideoDecoder


class FirstStage(nn.Module):
    """``first_stage_model`` + ``scale_factor`` + chunking, with the reference's two method names."""

    def __init__(self, first_stage_model, scale_factor=1.0, en_and_decode_n_samples_a_time=None,
                 disable_first_stage_autocast=False):
        super().__init__()
        self.first_stage_model = first_stage_model.eval()
        for p in self.first_stage_model.parameters():
            p.requires_grad = False
        self.scale_factor = scale_factor
        self.en_and_decode_n_samples_a_time = en_and_decode_n_samples_a_time
        self.disable_first_stage_autocast = disable_first_stage_autocast   # the B200 path has one precision policy

    @torch.no_grad()
    def decode_first_stage(self, z):
        z = 1.0 / self.scale_factor * z
        n_samples = z.shape[0] if self.en_and_decode_n_samples_a_time is None else self.en_and_decode_n_samples_a_time
        all_out = []
        for n in range(math.ceil(z.shape[0] / n_samples)):
            chunk = z[n * n_samples: (n + 1) * n_samples]
            kwargs = {"timesteps": len(chunk)} if isinstance(self.first_stage_model.decoder, VideoDecoder) else {}
            all_out.append(self.first_stage_model.decode(chunk, **kwargs))
        return torch.cat(all_out, dim=0)

    @torch.no_grad()
    def encode_first_stage(self, x):
        n_samples = x.shape[0] if self.en_and_decode_n_samples_a_time is None else self.en_and_decode_n_samples_a_time
        all_out = [self.first_stage_model.encode(x[n * n_samples: (n + 1) * n_samples])
                   for n in range(math.ceil(x.shape[0] / n_samples))]
        return self.scale_factor * torch.cat(all_out, dim=0)
