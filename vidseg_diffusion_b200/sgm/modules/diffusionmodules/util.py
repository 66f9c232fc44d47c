"""Shared primitives of the UNet (reference: sgm/modules/diffusionmodules/util.py)."""
import math

import torch
import torch.nn as nn


_FREQS = {}


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """reference :209-233: [N] -> [N, dim] = (cos | sin) of t * max_period^(-i/half).  B x dim values:
    host-latency bound, evaluated with the same fp32 expression order as the reference."""
    if repeat_only:
        return timesteps[:, None].expand(-1, dim)
    half = dim // 2
    # the frequency table is evaluated on the host with the reference's expression (bit-identical to its CPU run) and
    # kept on the device, so that a forward issues no host-to-device copy (CUDA-graph capturable)
    key = (half, float(max_period), timesteps.device)
    freqs = _FREQS.get(key)
    if freqs is None:
        freqs = _FREQS[key] = torch.exp(
            -math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    """reference :276-278 (parameter holder; the arithmetic is the fused GroupNorm+SiLU kernel)."""


def normalization(channels):
    return GroupNorm32(32, channels)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class AlphaBlender(nn.Module):
    """reference :314-391.  Holds ``mix_factor`` under the reference's name; the blend itself
    (alpha * x_spatial + (1 - alpha) * x_temporal) runs in the epilogue of the GEMM / convolution that produces
    x_temporal, so this module only supplies alpha per frame."""
    strategies = ["learned", "fixed", "learned_with_images"]

    def __init__(self, alpha, merge_strategy="learned_with_images", rearrange_pattern="b t -> (b t) 1 1"):
        super().__init__()
        assert merge_strategy in self.strategies, f"merge_strategy needs to be in {self.strategies}"
        self.merge_strategy = merge_strategy
        self.rearrange_pattern = rearrange_pattern
        if merge_strategy == "fixed":
            self.register_buffer("mix_factor", torch.Tensor([alpha]))
        else:
            self.register_parameter("mix_factor", nn.Parameter(torch.Tensor([alpha])))

    def frame_alpha(self, image_only_indicator, videos, frames):
        """alpha of every frame, fp32 [(b t)] on the device (reference get_alpha :357-366; both rearrange patterns
        index the same (b, t) entry)."""
        mf = self.mix_factor.detach().float()
        if self.merge_strategy == "fixed":
            a = mf.expand(videos, frames)
        elif self.merge_strategy == "learned":
            a = torch.sigmoid(mf).expand(videos, frames)
        else:
            assert image_only_indicator is not None, "need image_only_indicator ..."
            if tuple(image_only_indicator.shape) != (videos, frames):
                raise ValueError(f"image_only_indicator must be [{videos}, {frames}], got {tuple(image_only_indicator.shape)}")
            a = torch.where(image_only_indicator.to(mf.device).bool(), torch.ones(1, 1, device=mf.device),
                            torch.sigmoid(mf)[..., None])
        return a.reshape(-1).contiguous()
