"""Host side of the tcgen05 linear layers (csrc/gemm_tc.cu): split-fp16 operands and the
3-MMA GEMM.  Reference call sites: the nn.Linear layers of sgm/modules/attention.py."""
import torch

from . import _lib


def packed8(channels):
    """The library's operand policy (include/vidseg_b200.h, vidseg_set_operand_mode): rows of ``channels`` elements are
    carried as fp16 + fp8 corrections when the mode is on and the row length is a multiple of 64."""
    return _lib.load().vidseg_get_operand_mode() != 0 and channels % 64 == 0


class Split:
    """A tensor-core operand: an fp32 tensor x carried as hi = fp16(x * scale) plus a same-sized side tensor ``lo``:
    either lo = fp16(x * scale - hi) (``fmt == "pair16"``) or, per 64-element block of a row, 64 bytes of
    e5m2((x * scale - hi) * sl) followed by 64 bytes of e4m3(x * scale * sx) (``fmt == "packed8"``).  ``scale`` is a
    power of two (1 for activations, 2^8 for weights).  The kernels pick the format from the library policy and the row
    length; this object only records it so that ``float()`` can decode either."""
    __slots__ = ("hi", "lo", "scale", "fmt", "sl")

    def __init__(self, hi, lo, scale=1.0, fmt=None, sl=16.0):
        self.hi = hi
        self.lo = lo
        self.scale = scale
        self.fmt = fmt if fmt is not None else ("packed8" if packed8(hi.shape[-1]) else "pair16")
        self.sl = sl

    @property
    def shape(self):
        return self.hi.shape

    def reshape(self, *shape):
        out = Split(self.hi.reshape(*shape), self.lo.reshape(*shape), self.scale, self.fmt, self.sl)
        if self.fmt == "packed8" and out.hi.shape[-1] % 64:
            raise _lib.VidsegError("Split.reshape: a packed8 operand needs rows that are multiples of 64 elements")
        return out

    def float(self):
        """Decode to fp32 (tests / debugging; the e4m3 copy of x is redundant here)."""
        if self.fmt == "pair16":
            return (self.hi.float() + self.lo.float()) / self.scale
        c = self.hi.shape[-1]
        aux = self.lo.contiguous().view(torch.uint8).reshape(*self.hi.shape[:-1], c // 64, 128)
        lo8 = aux[..., :64].contiguous().view(torch.float8_e5m2).float().reshape(self.hi.shape)
        return (self.hi.float() + lo8 / self.sl) / self.scale

    def __getitem__(self, idx):
        return Split(self.hi[idx].contiguous(), self.lo[idx].contiguous(), self.scale, self.fmt, self.sl)


WEIGHT_SCALE = 256.0  # tc::kWeightScale


def split(x, scale=1.0, is_weight=False, pair16=False):
    """fp32 CUDA tensor [.., C] -> Split of x * scale (one streaming kernel).  ``pair16`` forces the fp16-pair format
    (attention operands); otherwise the library policy decides from C.  ``is_weight`` selects the weight scales."""
    x = _lib.require_cuda_tensor(x, torch.float32, "x")
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lib = _lib.load()
    cols = x.shape[-1] if x.dim() else 1
    with torch.cuda.device(x.device):
        if pair16 or x.numel() == 0:
            _lib.check(lib.vidseg_split_f16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), float(scale), _lib.stream_ptr()), "split_f16")
            return Split(hi, lo, float(scale), "pair16")
        _lib.check(lib.vidseg_split_rows(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel() // cols, cols, float(scale),
                                         1 if is_weight else 0, _lib.stream_ptr()), "split_rows")
    return Split(hi, lo, float(scale), None, 1.0 if is_weight else 16.0)


def gemm_split(a, w, bias=None, residual=None, want_f32=True, want_split=False, row_bias=None, rows_per_bias=1,
               blend=None, blend_alpha=None, rows_per_alpha=1, split_pair16=False, row_scalar=None):
    """out = a @ w.T (+ bias) (+ row_bias[row // rows_per_bias]) (+ residual), then optionally
    out = alpha * blend + (1 - alpha) * out with alpha = blend_alpha[row // rows_per_alpha].  ``row_scalar`` [M]: one value
    per output row added to all of its columns (mask modulation).
    a: Split [.., K], w: Split [N, K] (nn.Linear layout), both in the format the policy gives rows of K elements.
    ``split_pair16``: the split output feeds the attention kernel (fp16 pair) instead of another GEMM.
    Returns (out_f32 or None, Split or None)."""
    k = a.hi.shape[-1]
    lead = a.hi.shape[:-1]
    m = a.hi.numel() // k
    n = w.hi.shape[0]
    if w.hi.shape[1] != k:
        raise _lib.VidsegError(f"gemm_split: K mismatch {k} vs {w.hi.shape[1]}")
    for t, name in ((a.hi, "a.hi"), (w.hi, "w.hi")):
        _lib.require_cuda_tensor(t, torch.float16, name)
    for t, name in ((a.lo, "a.lo"), (w.lo, "w.lo")):
        _lib.require_cuda_tensor(t, torch.float16, name)
    want_fmt = "packed8" if packed8(k) else "pair16"
    if a.fmt != want_fmt or w.fmt != want_fmt:
        raise _lib.VidsegError(f"gemm_split: operands are {a.fmt} / {w.fmt}, the policy expects {want_fmt} for K={k}")
    dev = a.hi.device
    out = torch.empty((*lead, n), dtype=torch.float32, device=dev) if want_f32 else None
    oh = torch.empty((*lead, n), dtype=torch.float16, device=dev) if want_split else None
    ol = torch.empty((*lead, n), dtype=torch.float16, device=dev) if want_split else None
    if bias is not None:
        _lib.require_cuda_tensor(bias, torch.float32, "bias")
    if residual is not None:
        _lib.require_cuda_tensor(residual, torch.float32, "residual")
        if residual.numel() != m * n:
            raise _lib.VidsegError("gemm_split: residual shape mismatch")
    if row_bias is not None:
        _lib.require_cuda_tensor(row_bias, torch.float32, "row_bias")
        if rows_per_bias < 1 or row_bias.numel() != -(-m // rows_per_bias) * n:
            raise _lib.VidsegError(f"gemm_split: row_bias {tuple(row_bias.shape)} does not cover {m} rows in groups of {rows_per_bias}")
    if (blend is None) != (blend_alpha is None):
        raise _lib.VidsegError("gemm_split: blend and blend_alpha go together")
    if blend is not None:
        _lib.require_cuda_tensor(blend, torch.float32, "blend")
        _lib.require_cuda_tensor(blend_alpha, torch.float32, "blend_alpha")
        if blend.numel() != m * n or rows_per_alpha < 1 or blend_alpha.numel() != -(-m // rows_per_alpha):
            raise _lib.VidsegError("gemm_split: blend / blend_alpha shape mismatch")
    if row_scalar is not None:
        _lib.require_cuda_tensor(row_scalar, torch.float32, "row_scalar")
        if row_scalar.numel() != m:
            raise _lib.VidsegError(f"gemm_split: row_scalar must have one entry per output row ({m})")
    lib = _lib.load()
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        if row_bias is None and blend is None and row_scalar is None and not split_pair16:
            _lib.check(lib.vidseg_gemm_split(
                a.hi.data_ptr(), a.lo.data_ptr(), w.hi.data_ptr(), w.lo.data_ptr(), ptr(bias), ptr(residual),
                ptr(out), ptr(oh), ptr(ol), m, n, k, 1.0 / (a.scale * w.scale), _lib.stream_ptr()), "gemm_split")
        else:
            _lib.check(lib.vidseg_gemm_split_ex(
                a.hi.data_ptr(), a.lo.data_ptr(), w.hi.data_ptr(), w.lo.data_ptr(), ptr(bias), ptr(residual),
                ptr(row_bias), int(rows_per_bias), ptr(blend), ptr(blend_alpha), int(rows_per_alpha), ptr(row_scalar),
                ptr(out), ptr(oh), ptr(ol), 1 if split_pair16 else 0, m, n, k, 1.0 / (a.scale * w.scale),
                _lib.stream_ptr()), "gemm_split_ex")
    return out, (Split(oh, ol, 1.0, "pair16" if split_pair16 else None) if want_split else None)


def gemm_split_seg(a, w, nseg, want_f32, want_split, split_pair16=True):
    """``nseg`` projections of the same activation as one GEMM: w = the weights stacked along N, Split [nseg * seg_n, K].
    ``want_f32`` / ``want_split``: one flag per segment.  Returns a list of (out_f32 or None, Split or None), each
    [.., seg_n]."""
    import ctypes
    k = a.hi.shape[-1]
    lead = a.hi.shape[:-1]
    m = a.hi.numel() // k
    if w.hi.shape[1] != k or w.hi.shape[0] % nseg:
        raise _lib.VidsegError(f"gemm_split_seg: weight {tuple(w.hi.shape)} does not stack {nseg} projections of K={k}")
    seg_n = w.hi.shape[0] // nseg
    want_fmt = "packed8" if packed8(k) else "pair16"
    if a.fmt != want_fmt or w.fmt != want_fmt:
        raise _lib.VidsegError(f"gemm_split_seg: operands are {a.fmt} / {w.fmt}, the policy expects {want_fmt} for K={k}")
    for t, name in ((a.hi, "a.hi"), (a.lo, "a.lo"), (w.hi, "w.hi"), (w.lo, "w.lo")):
        _lib.require_cuda_tensor(t, torch.float16, name)
    dev = a.hi.device
    outs = []
    for s in range(nseg):
        if not (want_f32[s] or want_split[s]):
            raise _lib.VidsegError("gemm_split_seg: every segment needs an output")
        o = torch.empty((*lead, seg_n), dtype=torch.float32, device=dev) if want_f32[s] else None
        oh = torch.empty((*lead, seg_n), dtype=torch.float16, device=dev) if want_split[s] else None
        ol = torch.empty((*lead, seg_n), dtype=torch.float16, device=dev) if want_split[s] else None
        outs.append((o, oh, ol))
    arr = lambda i: (ctypes.c_void_p * nseg)(*[(t[i].data_ptr() if t[i] is not None else None) for t in outs])
    f32s, his, los = arr(0), arr(1), arr(2)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().vidseg_gemm_split_seg(
            a.hi.data_ptr(), a.lo.data_ptr(), w.hi.data_ptr(), w.lo.data_ptr(), nseg, seg_n, f32s, his, los,
            1 if split_pair16 else 0, m, k, 1.0 / (a.scale * w.scale), _lib.stream_ptr()), "gemm_split_seg")
    return [(o, Split(oh, ol, 1.0, "pair16" if split_pair16 else None) if oh is not None else None) for o, oh, ol in outs]


def attention_split(q, k, v, heads, scale=None, want_f32=False, want_split=True):
    """softmax(q k^T * scale) v per head.  q: Split [B, Nq, heads*64]; k, v: Split [B, Nk, heads*64].
    Returns (out_f32 or None, Split or None), both [B, Nq, heads*64]."""
    b, nq, c = q.hi.shape
    nk = k.hi.shape[1]
    if c != heads * 64 or k.hi.shape[2] != c or v.hi.shape != k.hi.shape or k.hi.shape[0] != b:
        raise _lib.VidsegError(f"attention_split: bad shapes q{tuple(q.hi.shape)} k{tuple(k.hi.shape)} v{tuple(v.hi.shape)} heads={heads}")
    for t, name in ((q.hi, "q.hi"), (k.hi, "k.hi"), (v.hi, "v.hi")):
        _lib.require_cuda_tensor(t, torch.float16, name)
    for t, name in ((q.lo, "q.lo"), (k.lo, "k.lo"), (v.lo, "v.lo")):
        _lib.require_cuda_tensor(t, torch.float16, name)
    if q.fmt != "pair16" or k.fmt != "pair16" or v.fmt != "pair16":
        raise _lib.VidsegError("attention_split: q / k / v must be fp16-pair operands (split(..., pair16=True))")
    dev = q.hi.device
    out = torch.empty((b, nq, c), dtype=torch.float32, device=dev) if want_f32 else None
    oh = torch.empty((b, nq, c), dtype=torch.float16, device=dev) if want_split else None
    ol = torch.empty((b, nq, c), dtype=torch.float16, device=dev) if want_split else None
    if scale is None:
        scale = 64 ** -0.5
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.vidseg_attention_split(
            q.hi.data_ptr(), q.lo.data_ptr(), k.hi.data_ptr(), k.lo.data_ptr(), v.hi.data_ptr(), v.lo.data_ptr(),
            out.data_ptr() if out is not None else None, oh.data_ptr() if oh is not None else None,
            ol.data_ptr() if ol is not None else None, b, heads, nq, nk, float(scale), _lib.stream_ptr()), "attention_split")
    return out, (Split(oh, ol) if want_split else None)
