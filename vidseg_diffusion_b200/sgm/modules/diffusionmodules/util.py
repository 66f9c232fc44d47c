"""Shared primitives of the UNet (reference: sgm/modules/diffusionmodules/util.py)."""
import math

import torch
import torch.nn as nn


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """reference :209-233: [N] -> [N, dim] = (cos | sin) of t * max_period^(-i/half).  B x dim values:
    host-latency bound, evaluated with the same fp32 expression order as the reference."""
    if repeat_only:
        return timesteps[:, None].expand(-1, dim)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
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
