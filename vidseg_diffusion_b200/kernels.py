"""Functional layer between the ``sgm`` operator mirrors and libvidseg_b200.

Every function takes / returns CUDA tensors and launches on the caller's current stream.  Weights
are nn.Parameters owned by the modules; their tensor-core operand form (fp16 hi/lo split, conv
kernels re-laid as [Cout, taps*Cin]) is derived once and cached until the parameter changes.
"""
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from .linear import Split, attention_split, gemm_split, split

_WEIGHT_CACHE = {}


def _cached(param, tag, make):
    key = (id(param), tag)
    hit = _WEIGHT_CACHE.get(key)
    stamp = (param.data_ptr(), param._version, param.device)
    if hit is not None and hit[0] == stamp and hit[1]() is param:
        return hit[2]
    val = make(param.detach())
    _WEIGHT_CACHE[key] = (stamp, weakref.ref(param), val)
    return val


def clear_weight_cache():
    _WEIGHT_CACHE.clear()


def weight_split(param):
    """[N, K] nn.Linear weight -> cached Split."""
    return _cached(param, "lin", lambda w: split(w.float().contiguous()))


def linear(xs, weight, bias=None, residual=None, want_f32=True, want_split=False):
    """nn.Linear on a Split activation: xs [.., K] x weight[N, K]^T (+ bias) (+ residual)."""
    b = None if bias is None else _cached(bias, "bias", lambda t: t.float().contiguous())
    return gemm_split(xs, weight_split(weight), b, residual, want_f32=want_f32, want_split=want_split)


def attention(qs, ks, vs, heads, scale):
    return attention_split(qs, ks, vs, heads, scale, want_f32=False, want_split=True)


def _require(x, name):
    return _lib.require_cuda_tensor(x, torch.float32, name)


# ------------------------------------------------------------------------------------------------
# normalisation / gating kernels
# ------------------------------------------------------------------------------------------------
def layer_norm_split(x, ln):
    """LayerNorm over the last dim (eps from the module) -> Split operand of the next GEMM."""
    _require(x, "x")
    y = F.layer_norm(x, (x.shape[-1],), ln.weight, ln.bias, ln.eps)
    return split(y)


def geglu_split(h):
    """h [.., 2*D] = (value | gate) -> Split(value * gelu(gate)) (erf form, attention.py:95-96)."""
    _require(h, "h")
    val, gate = h.chunk(2, dim=-1)
    return split((val * F.gelu(gate)).contiguous())


def group_norm_tokens_split(x, gn, silu=False):
    """x [B, C, H, W] -> (x as tokens [B, H*W, C] fp32, Split(GroupNorm(x)) as tokens)."""
    _require(x.contiguous(), "x")
    b, c, h, w = x.shape
    y = F.group_norm(x, gn.num_groups, gn.weight, gn.bias, gn.eps)
    if silu:
        y = F.silu(y)
    tok = lambda t: t.permute(0, 2, 3, 1).reshape(b, h * w, c).contiguous()
    return tok(x), split(tok(y))


def tokens_to_nchw(t, b, c, h, w):
    return t.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()


def group_norm_silu(x, gn):
    _require(x.contiguous(), "x")
    return F.silu(F.group_norm(x, gn.num_groups, gn.weight, gn.bias, gn.eps))


def conv2d(x, conv, stride=1, padding=1, channel_bias=None, residual=None):
    """nn.Conv2d (+ per-(sample, channel) bias [B, Cout]) (+ residual [B, Cout, Ho, Wo])."""
    _require(x.contiguous(), "x")
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = F.conv2d(x, conv.weight, conv.bias, stride=stride, padding=padding)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    if channel_bias is not None:
        y = y + channel_bias[:, :, None, None]
    if residual is not None:
        y = y + residual
    return y


def upsample_nearest2x(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


def concat_channels(a, b):
    return torch.cat([a, b], dim=1)


def dense(x, lin, act_silu_in=False):
    """Small host-latency-bound Linear on [B, K] fp32 (time embedding MLP, ResBlock emb_layers)."""
    _require(x, "x")
    if act_silu_in:
        x = F.silu(x)
    out, _ = linear(split(x.contiguous()), lin.weight, lin.bias, want_f32=True)
    return out
