"""Functional layer between the ``sgm`` operator mirrors and libvidseg_b200.

Every function takes / returns CUDA tensors and launches on the caller's current stream.  Image-shaped
activations are ``[B, C, H, W]`` tensors in ``torch.channels_last`` memory format, i.e. the memory the
kernels see is ``[B, H, W, C]``: tokens of the transformer blocks and pixels of the convolutions are the
same rows, no transposes exist anywhere in the forward.  Weights are nn.Parameters owned by the modules;
their tensor-core operand form (hi | lo fp16 split, conv kernels re-laid tap-major as
[Cout, ky*kx*Cin]) is derived once and cached until the parameter changes.
"""
import weakref

import torch

from . import _lib
from .linear import WEIGHT_SCALE, Split, attention_split, gemm_split, gemm_split_seg, packed8, split

_WEIGHT_CACHE = {}


def _cached(param, tag, make):
    key = (id(param), tag, _lib.load().vidseg_get_operand_mode())
    hit = _WEIGHT_CACHE.get(key)
    stamp = (param.data_ptr(), param._version, param.device)
    if hit is not None and hit[0] == stamp and hit[1]() is param:
        return hit[2]
    val = make(param.detach())
    _WEIGHT_CACHE[key] = (stamp, weakref.ref(param), val)
    return val


def clear_weight_cache():
    _WEIGHT_CACHE.clear()


def _f32(param, tag="f32"):
    return _cached(param, tag, lambda t: t.float().contiguous())


def weight_split(param):
    """[N, K] nn.Linear weight -> cached Split."""
    return _cached(param, "lin", lambda w: split(w.float().contiguous(), WEIGHT_SCALE, is_weight=True))


def conv_weight_split(param, cin_pad=None, cout_pad=None):
    """[Cout, Cin, kh, kw] nn.Conv2d weight -> cached Split of [Cout', kh*kw*Cin'] (tap-major, Cin / Cout zero-padded)."""
    def make(w):
        w = w.float().permute(0, 2, 3, 1)  # [Cout, kh, kw, Cin]
        if cin_pad is not None and cin_pad != w.shape[-1]:
            w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[-1]))
        if cout_pad is not None and cout_pad != w.shape[0]:
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, cout_pad - w.shape[0]))
        if packed8(w.shape[-1]):   # same policy as the activation it meets: decided by the channels of one tap
            return split(w.reshape(-1, w.shape[-1]).contiguous(), WEIGHT_SCALE, is_weight=True).reshape(w.shape[0], -1)
        return split(w.reshape(w.shape[0], -1).contiguous(), WEIGHT_SCALE, is_weight=True, pair16=True)
    return _cached(param, f"conv{cin_pad}_{cout_pad}", make)


def _bias_padded(bias, n):
    """fp32 bias zero-padded to n entries (convolutions whose Cout is not a multiple of 4: the RGB output of the VAE)."""
    if bias is None:
        return None
    if bias.shape[0] == n:
        return _f32(bias)
    return _cached(bias, f"bias_pad{n}", lambda b: torch.nn.functional.pad(b.float(), (0, n - b.shape[0])).contiguous())


def _workspace(device, nbytes):
    """Scratch of ONE call.  Allocated per call from torch's caching allocator (stream-ordered, so concurrent streams
    never share it; during CUDA-graph capture it lands in the graph's private pool and lives as long as the graph)."""
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)


def _require(x, name):
    return _lib.require_cuda_tensor(x, torch.float32, name)


def nhwc(x, name="x"):
    """[B, C, H, W] fp32 CUDA tensor (any memory format) -> contiguous [B, H, W, C] view / copy."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise _lib.VidsegError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if x.dtype != torch.float32 or x.dim() != 4:
        raise _lib.VidsegError(f"{name}: expected a 4-d float32 tensor, got {x.dtype} {tuple(x.shape)}")
    return x.permute(0, 2, 3, 1).contiguous()


def as_nchw(t):
    """contiguous [B, H, W, C] -> [B, C, H, W] view (channels_last strides)."""
    return t.permute(0, 3, 1, 2)


class ChannelCat:
    """``torch.cat([a, b], dim=1)`` that is never materialised: the consuming GroupNorm kernel reads both sources."""

    def __init__(self, a, b):
        if a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
            raise _lib.VidsegError(f"concat_channels: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
        self.a, self.b = a, b
        self.shape = (a.shape[0], a.shape[1] + b.shape[1], a.shape[2], a.shape[3])
        self.device = a.device

    def materialize(self):
        return torch.cat([self.a, self.b], dim=1)


def concat_channels(a, b):
    return ChannelCat(a, b)


# ------------------------------------------------------------------------------------------------
# tensor-core ops
# ------------------------------------------------------------------------------------------------
def linear(xs, weight, bias=None, residual=None, want_f32=True, want_split=False, **epilogue):
    """nn.Linear on a Split activation: xs [.., K] x weight[N, K]^T (+ bias) (+ residual); ``epilogue`` = the
    row_bias / blend terms of linear.gemm_split (temporal layers of the SVD UNet)."""
    b = None if bias is None else _f32(bias)
    return gemm_split(xs, weight_split(weight), b, residual, want_f32=want_f32, want_split=want_split, **epilogue)


def _stacked_weight_split(params):
    """Several [N_i, K] nn.Linear weights stacked along N -> one cached Split (keyed on every parameter's storage and
    version, so that replacing or updating any of them rebuilds it)."""
    key = (tuple(id(p) for p in params), "stack", _lib.load().vidseg_get_operand_mode())
    stamp = tuple((p.data_ptr(), p._version, p.device) for p in params)
    hit = _WEIGHT_CACHE.get(key)
    if hit is not None and hit[0] == stamp and all(r() is p for r, p in zip(hit[1], params)):
        return hit[2]
    val = split(torch.cat([p.detach().float() for p in params], dim=0).contiguous(), WEIGHT_SCALE, is_weight=True)
    _WEIGHT_CACHE[key] = (stamp, [weakref.ref(p) for p in params], val)
    return val


def linear_stacked(xs, weights, want_f32, want_split):
    """Bias-free projections of one activation by several equally shaped weights (to_q | to_k | to_v) as ONE GEMM with one
    output set per weight; the split outputs are fp16 pairs (operands of the attention kernel).  Returns a list of
    (fp32 or None, Split or None).  Falls back to None when the shapes do not allow the stacked form."""
    n = weights[0].shape[0]
    if any(tuple(w.shape) != tuple(weights[0].shape) for w in weights) or not (n % 128 == 0 or n % 160 == 0):
        return None
    return gemm_split_seg(xs, _stacked_weight_split(list(weights)), len(weights), want_f32, want_split, split_pair16=True)


def _geglu_perm(d, device):
    """Row order of the fused GEGLU projection: 64-row groups of 32 value rows + the 32 gate rows of the same features."""
    j = torch.arange(d // 32, device=device)[:, None] * 32 + torch.arange(32, device=device)[None, :]   # [d/32, 32]
    return torch.cat([j, j + d], dim=1).reshape(-1)


def linear_geglu(xs, proj):
    """GEGLU.proj followed by value * gelu(gate) (attention.py:89-96) as ONE GEMM whose epilogue applies the gate.
    xs: Split [.., K]; proj: nn.Linear(K, 2*D).  Returns the Split operand [.., D] of FeedForward.net[2]."""
    k = xs.hi.shape[-1]
    d = proj.out_features // 2
    if d % 32 or proj.in_features != k:
        raise _lib.VidsegError(f"linear_geglu: unsupported shape K={k} D={d}")
    ws = _cached(proj.weight, "geglu", lambda w: split(w.float()[_geglu_perm(d, w.device)].contiguous(), WEIGHT_SCALE, is_weight=True))
    bias = None if proj.bias is None else _cached(proj.bias, "geglu_b", lambda b: b.float()[_geglu_perm(d, b.device)].contiguous())
    want = "packed8" if packed8(k) else "pair16"
    if xs.fmt != want or ws.fmt != want:
        raise _lib.VidsegError(f"linear_geglu: operands are {xs.fmt} / {ws.fmt}, the policy expects {want} for K={k}")
    m = xs.hi.numel() // k
    out = _empty_split((*xs.hi.shape[:-1], d), xs.hi.device)
    lib = _lib.load()
    with torch.cuda.device(xs.hi.device):
        _lib.check(lib.vidseg_gemm_geglu_split(xs.hi.data_ptr(), xs.lo.data_ptr(), ws.hi.data_ptr(), ws.lo.data_ptr(),
                                               bias.data_ptr() if bias is not None else None, out.hi.data_ptr(),
                                               out.lo.data_ptr(), m, d, k, 1.0 / (xs.scale * ws.scale), _lib.stream_ptr()),
                   "gemm_geglu_split")
    return out


def attention(qs, ks, vs, heads, scale):
    return attention_split(qs, ks, vs, heads, scale, want_f32=False, want_split=True)


def conv2d(xs, conv, chan_bias=None, residual=None):
    """nn.Conv2d on a Split activation xs [B, H, W, Cin'] -> fp32 [B, Cout, Ho, Wo] (channels_last memory).
    chan_bias [B, Cout] is added per sample and channel; residual is an image-shaped fp32 tensor."""
    k, stride = conv.kernel_size[0], conv.stride[0]
    if conv.kernel_size[0] != conv.kernel_size[1] or k not in (1, 3) or conv.padding[0] != k // 2 or stride not in (1, 2):
        raise _lib.VidsegError(f"conv2d: unsupported geometry {conv}")
    b, h, w, cin = xs.hi.shape
    cout_true = conv.out_channels
    cout = -(-cout_true // 4) * 4     # zero output channels up to a multiple of 4, sliced off below
    ws = conv_weight_split(conv.weight, cin, cout)
    if ws.hi.shape[1] != k * k * cin:
        raise _lib.VidsegError(f"conv2d: activation has {cin} channels, weight expects {conv.in_channels}")
    ho, wo = h // stride, w // stride
    out = torch.empty((b, ho, wo, cout), dtype=torch.float32, device=xs.hi.device)
    res = None
    if residual is not None:
        res = nhwc(residual, "residual")
        if res.shape != out.shape:
            raise _lib.VidsegError("conv2d: residual shape mismatch")
    if chan_bias is not None:
        _require(chan_bias, "chan_bias")
        if tuple(chan_bias.shape) != (b, cout):
            raise _lib.VidsegError("conv2d: chan_bias must be [B, Cout]")
    if cout != cout_true and (res is not None or chan_bias is not None):
        raise _lib.VidsegError("conv2d: residual / chan_bias with a padded Cout is not supported")
    bias = _bias_padded(conv.bias, cout)
    lib = _lib.load()
    with torch.cuda.device(out.device):
        _lib.check(lib.vidseg_conv2d_split(
            xs.hi.data_ptr(), xs.lo.data_ptr(), ws.hi.data_ptr(), ws.lo.data_ptr(),
            bias.data_ptr() if bias is not None else None,
            chan_bias.data_ptr() if chan_bias is not None else None,
            res.data_ptr() if res is not None else None,
            out.data_ptr(), None, None, b, h, w, cin, cout, k, stride, 1.0 / (xs.scale * ws.scale), _lib.stream_ptr()), "conv2d_split")
    return as_nchw(out if cout == cout_true else out[..., :cout_true])


def conv2d_pad_after(xs, conv):
    """``Downsample`` of the first-stage encoder (model.py:77-94): F.pad(x, (0, 1, 0, 1)) + 3x3 stride-2 convolution with
    padding 0, on a Split activation [B, H, W, Cin] -> fp32 [B, Cout, H/2, W/2] (channels_last memory)."""
    if tuple(conv.kernel_size) != (3, 3) or tuple(conv.stride) != (2, 2) or tuple(conv.padding) != (0, 0):
        raise _lib.VidsegError(f"conv2d_pad_after: unsupported geometry {conv}")
    b, h, w, cin = xs.hi.shape
    cout = conv.out_channels
    if cin != conv.in_channels or h % 2 or w % 2 or cin % 64 or cout % 4:
        raise _lib.VidsegError(f"conv2d_pad_after: needs even H, W, Cin % 64 == 0, Cout % 4 == 0; got {tuple(xs.hi.shape)} -> {cout}")
    ws = conv_weight_split(conv.weight, cin)
    out = torch.empty((b, h // 2, w // 2, cout), dtype=torch.float32, device=xs.hi.device)
    bias = None if conv.bias is None else _f32(conv.bias)
    lib = _lib.load()
    with torch.cuda.device(out.device):
        _lib.check(lib.vidseg_conv2d_down_pad_after_split(
            xs.hi.data_ptr(), xs.lo.data_ptr(), ws.hi.data_ptr(), ws.lo.data_ptr(), bias.data_ptr() if bias is not None else None,
            out.data_ptr(), None, None, b, h, w, cin, cout, 1.0 / (xs.scale * ws.scale), _lib.stream_ptr()),
            "conv2d_down_pad_after_split")
    return as_nchw(out)


def softmax_rows_split(logits, scale):
    """softmax(logits * scale) over the last dim of an fp32 [rows, cols] tensor -> Split operand [rows, cols]."""
    _require(logits, "logits")
    rows, cols = logits.shape
    out = _empty_split((rows, cols), logits.device)
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        _lib.check(lib.vidseg_softmax_rows_split(logits.data_ptr(), float(scale), out.hi.data_ptr(), out.lo.data_ptr(), rows, cols,
                                                 _lib.stream_ptr()), "softmax_rows_split")
    return out


def single_head_attention(q, k, v, scale):
    """softmax(q k^T * scale) v for ONE head of arbitrary dimension (the first-stage AttnBlock, model.py:160-204: head
    dimension = channels, 512 at full width, which the 64-wide flash kernel of the UNet does not take).  q, k, v: fp32
    [n, c].  Two tensor-core GEMMs around the row-softmax kernel; k and v^T take the weight side of their GEMM (the
    2^8-scaled operand format, so |k|, |v| < 255: activations behind a GroupNorm are far below)."""
    for t, name in ((q, "q"), (k, "k"), (v, "v")):
        _require(t, name)
    n, c = q.shape
    s, _ = gemm_split(split(q), split(k, WEIGHT_SCALE, is_weight=True), want_f32=True)                      # [n, n] logits
    p = softmax_rows_split(s, scale)
    o, _ = gemm_split(p, split(v.t().contiguous(), WEIGHT_SCALE, is_weight=True), want_f32=True)            # [n, c]
    return o


def conv_temporal(xs, conv, videos, frames, frame_bias=None, residual=None, blend=None, blend_alpha=None):
    """nn.Conv3d with a (3,1,1) kernel over the frame axis.  xs: Split [(V T), H, W, Cin] (the 2-D layers' layout);
    frame_bias [(V T), Cout]; residual / blend: image-shaped fp32 [(V T), Cout, H, W]; blend_alpha [(V T)].
    Returns fp32 [(V T), Cout, H, W] (channels_last memory)."""
    if tuple(conv.kernel_size) != (3, 1, 1) or tuple(conv.padding) != (1, 0, 0) or tuple(conv.stride) != (1, 1, 1):
        raise _lib.VidsegError(f"conv_temporal: unsupported geometry {conv}")
    bt, h, w, cin = xs.hi.shape
    if bt != videos * frames or cin < conv.in_channels or (cin != conv.in_channels and cin != -(-conv.in_channels // 8) * 8):
        raise _lib.VidsegError("conv_temporal: shape mismatch")
    cout_true = conv.out_channels
    cout = -(-cout_true // 4) * 4
    def make_t(t):
        w = t.float()[:, :, :, 0, 0].permute(0, 2, 1)   # [Cout, 3, Cin], tap-major
        w = torch.nn.functional.pad(w, (0, cin - w.shape[2], 0, 0, 0, cout - w.shape[0])).contiguous()
        if packed8(cin):
            return split(w.reshape(-1, cin), WEIGHT_SCALE, is_weight=True).reshape(cout, 3 * cin)
        return split(w.reshape(cout, 3 * cin), WEIGHT_SCALE, is_weight=True, pair16=True)
    ws = _cached(conv.weight, f"conv_t{cin}_{cout}", make_t)
    out = torch.empty((bt, h, w, cout), dtype=torch.float32, device=xs.hi.device)
    res = None if residual is None else nhwc(residual, "residual")
    bl = None if blend is None else nhwc(blend, "blend")
    for t, name in ((res, "residual"), (bl, "blend")):
        if t is not None and t.shape != out.shape:
            raise _lib.VidsegError(f"conv_temporal: {name} shape mismatch")
    if frame_bias is not None:
        _require(frame_bias, "frame_bias")
        if tuple(frame_bias.shape) != (bt, cout):
            raise _lib.VidsegError("conv_temporal: frame_bias must be [(V T), Cout]")
    if (bl is None) != (blend_alpha is None):
        raise _lib.VidsegError("conv_temporal: blend and blend_alpha go together")
    if blend_alpha is not None:
        _require(blend_alpha, "blend_alpha")
        if blend_alpha.numel() != bt:
            raise _lib.VidsegError("conv_temporal: blend_alpha must have one entry per frame")
    if cout != cout_true and (res is not None or bl is not None or frame_bias is not None):
        raise _lib.VidsegError("conv_temporal: epilogue terms with a padded Cout are not supported")
    bias = _bias_padded(conv.bias, cout)
    ptr = lambda t: t.data_ptr() if t is not None else None
    lib = _lib.load()
    with torch.cuda.device(out.device):
        _lib.check(lib.vidseg_conv_temporal_split(
            xs.hi.data_ptr(), xs.lo.data_ptr(), ws.hi.data_ptr(), ws.lo.data_ptr(), ptr(bias), ptr(frame_bias), ptr(res),
            ptr(bl), ptr(blend_alpha), out.data_ptr(), None, None, videos, frames, h * w, cin, cout,
            1.0 / (xs.scale * ws.scale), _lib.stream_ptr()), "conv_temporal_split")
    return as_nchw(out if cout == cout_true else out[..., :cout_true])


def temporal_attention(q, k, v, videos, frames, heads, scale):
    """Self-attention over the frames of every spatial site.  q, k, v: fp32 [(V T), S, heads*64] in the frame-major
    layout of the spatial layers.  Returns the Split operand [(V T), S, heads*64] of the to_out GEMM."""
    for t, name in ((q, "q"), (k, "k"), (v, "v")):
        _require(t, name)
    bt, s, c = q.shape
    if bt != videos * frames or c != heads * 64 or k.shape != q.shape or v.shape != q.shape:
        raise _lib.VidsegError(f"temporal_attention: bad shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)}")
    out = _empty_split(q.shape, q.device)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.vidseg_temporal_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), None, out.hi.data_ptr(),
                                                 out.lo.data_ptr(), videos, frames, s, heads, float(scale),
                                                 _lib.stream_ptr()), "temporal_attention")
    return out


# ------------------------------------------------------------------------------------------------
# normalisation / gating / resampling kernels (all emit the split operand of the next GEMM)
# ------------------------------------------------------------------------------------------------
def _empty_split(shape, device):
    """Output buffers of a producer kernel; the kernel fills them in the policy's format for rows of shape[-1]."""
    return Split(torch.empty(shape, dtype=torch.float16, device=device), torch.empty(shape, dtype=torch.float16, device=device))


def layer_norm_split(x, ln, row_bias=None, rows_per_bias=1):
    """LayerNorm over the last dim of (x + row_bias[row // rows_per_bias]) -> Split operand of the next GEMM."""
    _require(x, "x")
    c = x.shape[-1]
    rows = x.numel() // c
    out = _empty_split(x.shape, x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        if row_bias is None:
            _lib.check(lib.vidseg_layernorm_split(x.data_ptr(), _f32(ln.weight).data_ptr(), _f32(ln.bias).data_ptr(),
                                                  float(ln.eps), out.hi.data_ptr(), out.lo.data_ptr(), rows, c,
                                                  _lib.stream_ptr()), "layernorm_split")
        else:
            _require(row_bias, "row_bias")
            if rows_per_bias < 1 or row_bias.numel() != -(-rows // rows_per_bias) * c:
                raise _lib.VidsegError(f"layer_norm: row_bias {tuple(row_bias.shape)} does not cover {rows} rows in groups of {rows_per_bias}")
            _lib.check(lib.vidseg_layernorm_bias_split(x.data_ptr(), row_bias.data_ptr(), int(rows_per_bias),
                                                       _f32(ln.weight).data_ptr(), _f32(ln.bias).data_ptr(), float(ln.eps),
                                                       out.hi.data_ptr(), out.lo.data_ptr(), rows, c, _lib.stream_ptr()),
                       "layernorm_bias_split")
    return out


def geglu_split(h):
    """h [.., 2*D] = (value | gate) -> Split(value * gelu(gate)) (erf form, attention.py:95-96)."""
    _require(h, "h")
    d = h.shape[-1] // 2
    out = _empty_split((*h.shape[:-1], d), h.device)
    lib = _lib.load()
    with torch.cuda.device(h.device):
        _lib.check(lib.vidseg_geglu_split(h.data_ptr(), out.hi.data_ptr(), out.lo.data_ptr(), h.numel() // (2 * d), d,
                                          _lib.stream_ptr()), "geglu_split")
    return out


def group_norm_split(x, gn, silu, want_raw=False, samples=None):
    """GroupNorm (+ SiLU) of an image-shaped fp32 tensor or a ChannelCat of two.

    Returns (Split [B, H, W, C] of the normalised activation, Split of the raw input or None, [B, H, W, Ca] fp32
    view of the first source)."""
    if isinstance(x, ChannelCat):
        a, bsrc = nhwc(x.a, "x.a"), nhwc(x.b, "x.b")
    else:
        a, bsrc = nhwc(x), None
    b, h, w, c1 = a.shape
    c2 = 0 if bsrc is None else bsrc.shape[3]
    c = c1 + c2
    if c != gn.num_channels:
        raise _lib.VidsegError(f"group_norm: {c} channels, module expects {gn.num_channels}")
    out = _empty_split((b, h, w, c), a.device)
    raw = _empty_split((b, h, w, c), a.device) if want_raw else None
    lib = _lib.load()
    hw = h * w
    if samples is not None:  # statistics over groups of b // samples consecutive images (GroupNorm on 'b c t h w')
        if b % samples:
            raise _lib.VidsegError(f"group_norm: {b} images do not split into {samples} clips")
        hw, b = hw * (b // samples), samples
    nbytes = lib.vidseg_groupnorm_workspace_bytes(b, gn.num_groups)
    ws = _workspace(a.device, nbytes)
    with torch.cuda.device(a.device):
        _lib.check(lib.vidseg_groupnorm_split(
            a.data_ptr(), c1, bsrc.data_ptr() if bsrc is not None else None, c2,
            _f32(gn.weight).data_ptr(), _f32(gn.bias).data_ptr(), float(gn.eps), gn.num_groups, 1 if silu else 0,
            out.hi.data_ptr(), out.lo.data_ptr(), raw.hi.data_ptr() if raw else None, raw.lo.data_ptr() if raw else None,
            b, hw, ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "groupnorm_split")
    return out, raw, a


def upsample_nearest2x_split(x):
    """[B, C, H, W] fp32 -> Split [B, 2H, 2W, C] of the nearest-neighbour x2 upsampling."""
    a = nhwc(x)
    b, h, w, c = a.shape
    out = _empty_split((b, 2 * h, 2 * w, c), a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.vidseg_upsample2x_split(a.data_ptr(), out.hi.data_ptr(), out.lo.data_ptr(), b, h, w, c,
                                               _lib.stream_ptr()), "upsample2x_split")
    return out


def image_split(x, pad_to=8):
    """[B, C, H, W] fp32 -> Split [B, H, W, C'] with C padded up to a multiple of ``pad_to`` (input conv: 4 -> 8)."""
    a = nhwc(x)
    c = a.shape[3]
    if c % pad_to:
        a = torch.nn.functional.pad(a, (0, pad_to - c % pad_to))
    return split(a)


def dense(x, lin, act_silu_in=False):
    """Small host-latency-bound Linear on [B, K] fp32 (time embedding MLP, ResBlock emb_layers)."""
    _require(x, "x")
    if act_silu_in:
        x = torch.nn.functional.silu(x)
    out, _ = linear(split(x.contiguous()), lin.weight, lin.bias, want_f32=True)
    return out
