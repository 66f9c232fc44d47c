"""TEST INFRASTRUCTURE ONLY -- the third-party arithmetic inside the seg-map post-process, restated.

``oracle/process_output.py`` follows the reference and CALLS OpenCV / Pillow where the reference does.  Those library
routines are not under /root/reference; what they compute is restated here from their published algorithms, in plain
numpy / Python integers, and this restatement is what ``csrc/postprocess.cu`` implements.  Each piece is checked against
the installed library on the CPU (tests/test_process_output.py), so the chain is
reference == oracle/process_output.py (goldens) == this file (library equivalence) == CUDA kernel (GPU tests).

  gaussian_blur_5x5_f64   OpenCV 4.x ``GaussianBlur(src64f, (5, 5), 3)``: separable filter, BORDER_REFLECT_101, kernel
                          ``getGaussianKernel(5, 3, CV_64F)``; the row filter is the scalar ``RowFilter<double,double>``
                          (sum_k kx[k] * S[k], k ascending; the compiler contracts the 4-wide unrolled body into fused
                          multiply-adds, the remainder columns are plain multiply + add), the column filter is
                          ``SymmColumnFilter``: ky[0] * S0, then += ky[k] * (S[+k] + S[-k]), no contraction
                          (modules/imgproc/src/filter.simd.hpp).
  float64_to_L            Pillow ``Image.fromarray(float64)`` (mode F: C cast to float32) + ``convert("L")``
                          (src/libImaging/Convert.c f2l: v <= 0 -> 0, v >= 255 -> 255, else truncate).
  jpeg_roundtrip_L        libjpeg(-turbo) baseline grayscale, quality 75, ISLOW DCT both ways: level shift, jfdctint.c
                          forward DCT (CONST_BITS 13, PASS1_BITS 2), quantisation by (q << 3) with round-half-away, then
                          the decoder's dequantise + jidctint.c inverse DCT + range limit.  Entropy coding is lossless
                          and drops out of the round trip.  Partial edge blocks are padded by edge replication
                          (jcprepct.c expand_bottom_edge / jccoefct.c).
  lanczos_resize_L        Pillow ``Image.resize(size, LANCZOS)`` for 8-bit images (src/libImaging/Resample.c): float64
                          coefficient windows, fixed point with 22 fractional bits, horizontal pass then vertical pass,
                          each rounded to uint8.
"""
import math
from fractions import Fraction

import numpy as np

# cv2.getGaussianKernel(5, 3, cv2.CV_64F) (bit-exact kernel of OpenCV >= 4.0); tests compare it with the installed cv2
GAUSS_5_SIGMA3 = np.array([float.fromhex(h) for h in (
    "0x1.6cf5d45c5fe17p-3", "0x1.af264d4f67a34p-3", "0x1.c7c7bca870f66p-3", "0x1.af264d4f67a34p-3", "0x1.6cf5d45c5fe17p-3")])


def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * n - 2 - i
    return i


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def squared_difference_sqrt(frame_pos, frame_neg):
    """process_output.py:13: uint8 wrap-around difference, squared in uint8 (mod 256), summed over colour, sqrt."""
    a = frame_pos.astype(np.int64)
    b = frame_neg.astype(np.int64)
    d = (a - b) & 255
    sq = (d * d) & 255
    return np.sqrt(sq.sum(axis=2).astype(np.float64))


def gaussian_blur_5x5_f64(d):
    """Small images only (exact fused multiply-adds through Python fractions)."""
    k = GAUSS_5_SIGMA3
    H, W = d.shape
    tmp = np.zeros((H, W))
    body = (W // 4) * 4
    for y in range(H):
        for x in range(W):
            v = [d[y, _reflect101(x + j - 2, W)] for j in range(5)]
            acc = k[0] * v[0]
            for j in range(1, 5):
                acc = _fma(k[j], v[j], acc) if x < body else acc + k[j] * v[j]
            tmp[y, x] = acc
    out = np.zeros((H, W))
    for y in range(H):
        rows = [tmp[_reflect101(y + j - 2, H)] for j in range(5)]
        acc = k[2] * rows[2]
        acc = acc + k[3] * (rows[3] + rows[1])
        acc = acc + k[4] * (rows[4] + rows[0])
        out[y] = acc
    return out


def float64_to_L(d):
    v = d.astype(np.float32)
    out = np.where(v <= 0, 0, np.where(v >= 255, 255, np.trunc(np.clip(v, 0, 255)))).astype(np.uint8)
    return out


# ---------------------------------------------------------------------------------------------------
# JPEG round trip
# ---------------------------------------------------------------------------------------------------
STD_LUMA_QUANT = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64).reshape(8, 8)


def quant_table(quality=75):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table (force_baseline)."""
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((STD_LUMA_QUANT * scale + 50) // 100, 1, 255)


_F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069,
          f2053=16819, f2562=20995, f3072=25172)
CONST_BITS, PASS1_BITS = 13, 2


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, first):
    """One pass of jfdctint.c over the LAST axis of an int64 array [..., 8]."""
    t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    out = np.empty_like(d)
    sh = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS
    if first:
        out[..., 0] = (t10 + t11) << PASS1_BITS
        out[..., 4] = (t10 - t11) << PASS1_BITS
    else:
        out[..., 0] = _descale(t10 + t11, PASS1_BITS)
        out[..., 4] = _descale(t10 - t11, PASS1_BITS)
    z1 = (t12 + t13) * _F["f0541"]
    out[..., 2] = _descale(z1 + t13 * _F["f0765"], sh)
    out[..., 6] = _descale(z1 + t12 * (-_F["f1847"]), sh)
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * _F["f1175"]
    t4, t5, t6, t7 = t4 * _F["f0298"], t5 * _F["f2053"], t6 * _F["f3072"], t7 * _F["f1501"]
    z1, z2, z3, z4 = z1 * (-_F["f0899"]), z2 * (-_F["f2562"]), z3 * (-_F["f1961"]), z4 * (-_F["f0390"])
    z3, z4 = z3 + z5, z4 + z5
    out[..., 7] = _descale(t4 + z1 + z3, sh)
    out[..., 5] = _descale(t5 + z2 + z4, sh)
    out[..., 3] = _descale(t6 + z2 + z3, sh)
    out[..., 1] = _descale(t7 + z1 + z4, sh)
    return out


def _idct_1d(c, first):
    """One pass of jidctint.c over the LAST axis of an int64 array [..., 8] (dequantised coefficients in)."""
    z2, z3 = c[..., 2], c[..., 6]
    z1 = (z2 + z3) * _F["f0541"]
    t2 = z1 + z3 * (-_F["f1847"])
    t3 = z1 + z2 * _F["f0765"]
    z2, z3 = c[..., 0], c[..., 4]
    t0, t1 = (z2 + z3) << CONST_BITS, (z2 - z3) << CONST_BITS
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = c[..., 7], c[..., 5], c[..., 3], c[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * _F["f1175"]
    t0, t1, t2, t3 = t0 * _F["f0298"], t1 * _F["f2053"], t2 * _F["f3072"], t3 * _F["f1501"]
    z1, z2, z3, z4 = z1 * (-_F["f0899"]), z2 * (-_F["f2562"]), z3 * (-_F["f1961"]), z4 * (-_F["f0390"])
    z3, z4 = z3 + z5, z4 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    sh = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS + 3
    out = np.empty_like(c)
    out[..., 0], out[..., 7] = _descale(t10 + t3, sh), _descale(t10 - t3, sh)
    out[..., 1], out[..., 6] = _descale(t11 + t2, sh), _descale(t11 - t2, sh)
    out[..., 2], out[..., 5] = _descale(t12 + t1, sh), _descale(t12 - t1, sh)
    out[..., 3], out[..., 4] = _descale(t13 + t0, sh), _descale(t13 - t0, sh)
    return out


def jpeg_roundtrip_L(img, quality=75):
    img = np.asarray(img, dtype=np.uint8)
    H, W = img.shape
    Hp, Wp = (H + 7) // 8 * 8, (W + 7) // 8 * 8
    pad = np.pad(img, ((0, Hp - H), (0, Wp - W)), mode="edge").astype(np.int64)
    blocks = pad.reshape(Hp // 8, 8, Wp // 8, 8).transpose(0, 2, 1, 3) - 128          # [by, bx, row, col]
    q = quant_table(quality)
    c = _fdct_1d(blocks, True)                                                          # rows
    c = _fdct_1d(c.transpose(0, 1, 3, 2), False).transpose(0, 1, 3, 2)                  # columns
    qv = q << 3
    mag = (np.abs(c) + (qv >> 1)) // qv
    coef = np.where(c < 0, -mag, mag) * q                                               # quantise, dequantise
    w = _idct_1d(coef.transpose(0, 1, 3, 2), True).transpose(0, 1, 3, 2)                # columns first (jidctint.c pass 1)
    o = _idct_1d(w, False)                                                              # rows
    o = np.clip(o + 128, 0, 255)
    return o.transpose(0, 2, 1, 3).reshape(Hp, Wp)[:H, :W].astype(np.uint8)


# ---------------------------------------------------------------------------------------------------
# Pillow LANCZOS resize, 8 bits per channel
# ---------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _sinc(x):
    return 1.0 if x == 0.0 else math.sin(x * math.pi) / (x * math.pi)


def _lanczos(x):
    return _sinc(x) * _sinc(x / 3) if -3.0 <= x < 3.0 else 0.0


def lanczos_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc.  Returns (bounds int [out, 2] = (xmin, count), int coefficient
    windows [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis_last(img, out_size):
    bounds, kk = lanczos_coeffs(img.shape[-1], out_size)
    out = np.zeros(img.shape[:-1] + (out_size,), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (src[..., xmin:xmin + n] * kk[xx, :n]).sum(axis=-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def lanczos_resize_L(img, height, width):
    img = np.asarray(img, dtype=np.uint8)
    if img.shape[1] != width:
        img = _resample_axis_last(img, width)                       # horizontal pass first
    if img.shape[0] != height:
        img = _resample_axis_last(img.T, height).T
    return img


def seg_maps(frames_pos, frames_neg, unique_labels, label_maps=None, filter_difference=False, filter_s=0.7):
    """The whole post-process built from the restated pieces only (small sizes: the blur runs in Python)."""
    unique_labels = np.asarray(unique_labels)
    K, F, H, W, _ = frames_pos.shape
    maps = np.zeros((K, F, H, W))
    for i in range(K):
        for f in range(F):
            img = float64_to_L(gaussian_blur_5x5_f64(squared_difference_sqrt(frames_pos[i, f], frames_neg[i, f])))
            back = jpeg_roundtrip_L(img)
            dm = back / (np.max(back) + 1e-5)
            if filter_difference:
                m = lanczos_resize_L(np.where(label_maps[f] == unique_labels[i], 255, 0).astype(np.uint8), H, W) / 255.0
                dm = dm * m + filter_s * dm * (1 - m)
            maps[i, f] = dm
    seg = np.argmax(maps, axis=0)
    return unique_labels[seg].astype(np.uint8), seg
