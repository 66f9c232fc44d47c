// F4: segmentation-map post-process (scripts/sampling/process_output.py:8-28, 30-38, 74-167) on the device.
//
// The reference turns the decoded frames of the +lambda / -lambda modulated runs of every mask into the final label
// maps through a chain of library calls with file round trips in between; every one of them is deterministic integer or
// IEEE arithmetic and is reproduced here bit for bit (restated in oracle/process_output_emul.py, pinned to OpenCV /
// Pillow / libjpeg on the CPU):
//   difference   uint8 wrap-around (a - b), squared in uint8, summed over colour, sqrt in float64            (:13)
//   blur         cv2.GaussianBlur(float64, (5, 5), 3): separable, BORDER_REFLECT_101; row filter sum_k kx[k] S[k] with
//                the library's fused multiply-adds in the 4-wide body and plain mul + add in the remainder columns,
//                column filter ky[0] S0 + ky[1] (S1 + S-1) + ky[2] (S2 + S-2) without contraction                  (:15)
//   to "L"       float64 -> float32 (Image.fromarray) -> clip / truncate (convert("L"))                            (:18)
//   JPEG         quality-75 grayscale baseline round trip: forward ISLOW DCT, quantise, dequantise, inverse ISLOW DCT
//                (what .save(.jpg) followed by Image.open(.jpg) does to the pixels)                          (:19 -> :122)
//   normalise    / (max + 1e-5), optional mask filter d m + s d (1 - m) with the LANCZOS-resized 0/255 mask (:124-136)
//   arg-max      over masks, first maximum, mapped through unique_labels                                     (:150-160)
// All kernels are HBM-bound streaming passes: 6 B per pixel and mask in (two RGB frames), 1-2 B out.
#define VS_FAMILY vidseg::kFamOther
#include "common.cuh"

namespace vidseg {

// cv2.getGaussianKernel(5, 3, CV_64F)
__constant__ double kGauss[5] = {0x1.6cf5d45c5fe17p-3, 0x1.af264d4f67a34p-3, 0x1.c7c7bca870f66p-3, 0x1.af264d4f67a34p-3,
                                 0x1.6cf5d45c5fe17p-3};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ double sq_diff_sqrt(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t pix) {
  int s = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int d = ((int)a[pix * 3 + c] - (int)b[pix * 3 + c]) & 255;   // uint8 subtraction wraps
    s += (d * d) & 255;                                                // ... and so does the square
  }
  return sqrt((double)s);
}

__device__ __forceinline__ unsigned char f64_to_l(double d) {
  const float v = (float)d;
  if (!(v > 0.0f)) return 0;       // v <= 0 (and NaN, which the x86 cast turns into 0)
  if (v >= 255.0f) return 255;
  return (unsigned char)v;
}

constexpr int kBlurTile = 32;

// one block = one 32x32 tile of one image: row filter of the 36 rows the tile needs into shared memory, column filter,
// conversion to "L".  mode 0: write L and fold the blurred maximum into dmax; mode 1: write the "vis" image d / dmax * 255.
__global__ void __launch_bounds__(256)
segmap_blur_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ neg, int height, int width,
                   uint8_t* __restrict__ out_l, unsigned long long* __restrict__ dmax_bits, int mode) {
  __shared__ double rowf[kBlurTile + 4][kBlurTile];
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kBlurTile, y0 = blockIdx.y * kBlurTile;
  const size_t base = (size_t)img * height * width;
  const uint8_t* a = pos + base * 3;
  const uint8_t* b = neg + base * 3;
  const int body = (width / 4) * 4;
  for (int e = threadIdx.x; e < (kBlurTile + 4) * kBlurTile; e += blockDim.x) {
    const int ry = e / kBlurTile, rx = e % kBlurTile;
    const int x = x0 + rx;
    const int y = reflect101(y0 + ry - 2, height);
    double acc = 0.0;
    if (x < width && y0 + ry - 2 < height + 2) {
      const size_t rowp = (size_t)y * width;
      double v[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) v[j] = sq_diff_sqrt(a, b, rowp + reflect101(x + j - 2, width));
      acc = __dmul_rn(kGauss[0], v[0]);
      if (x < body) {
#pragma unroll
        for (int j = 1; j < 5; ++j) acc = __fma_rn(kGauss[j], v[j], acc);
      } else {
#pragma unroll
        for (int j = 1; j < 5; ++j) acc = __dadd_rn(acc, __dmul_rn(kGauss[j], v[j]));
      }
    }
    rowf[ry][rx] = acc;
  }
  __syncthreads();
  double local_max = 0.0;
  const double scale_max = (mode == 1) ? __longlong_as_double((long long)dmax_bits[img]) : 0.0;
  for (int e = threadIdx.x; e < kBlurTile * kBlurTile; e += blockDim.x) {
    const int ty = e / kBlurTile, tx = e % kBlurTile;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= width || y >= height) continue;
    double acc = __dmul_rn(kGauss[2], rowf[ty + 2][tx]);
    acc = __dadd_rn(acc, __dmul_rn(kGauss[3], __dadd_rn(rowf[ty + 3][tx], rowf[ty + 1][tx])));
    acc = __dadd_rn(acc, __dmul_rn(kGauss[4], __dadd_rn(rowf[ty + 4][tx], rowf[ty][tx])));
    if (mode == 0) {
      out_l[base + (size_t)y * width + x] = f64_to_l(acc);
      local_max = fmax(local_max, acc);
    } else {
      const double vis = (scale_max > 0.0) ? __dmul_rn(__ddiv_rn(acc, scale_max), 255.0) : 0.0;
      out_l[base + (size_t)y * width + x] = f64_to_l(vis);
    }
  }
  if (mode == 0) {
    // blurred values are non-negative: their IEEE bit patterns order like unsigned integers
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmax(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0 && local_max > 0.0)
      atomicMax(&dmax_bits[img], (unsigned long long)__double_as_longlong(local_max));
  }
}

// ---------------------------------------------------------------------------------------------------
// JPEG quality-75 grayscale round trip of one 8x8 block per thread (jfdctint.c / jcdctmgr.c / jidctint.c)
// ---------------------------------------------------------------------------------------------------
__constant__ int kQuant75[64] = {8,  6,  5,  8,  12, 20, 26, 31, 6,  6,  7,  10, 13, 29, 30, 28, 7,  7,  8,  12, 20, 29,
                                 35, 28, 7,  9,  11, 15, 26, 44, 40, 31, 9,  11, 19, 28, 34, 55, 52, 39, 12, 18, 28, 32,
                                 41, 52, 57, 46, 25, 32, 39, 44, 52, 61, 60, 51, 36, 46, 48, 49, 56, 50, 52, 50};
constexpr int kConstBits = 13, kPass1Bits = 2;
constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299,
              F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

template <bool FIRST>
__device__ __forceinline__ void fdct8(int* d, int stride) {
  const int t0 = d[0] + d[7 * stride], t7 = d[0] - d[7 * stride], t1 = d[stride] + d[6 * stride], t6 = d[stride] - d[6 * stride];
  const int t2 = d[2 * stride] + d[5 * stride], t5 = d[2 * stride] - d[5 * stride], t3 = d[3 * stride] + d[4 * stride],
            t4 = d[3 * stride] - d[4 * stride];
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  constexpr int sh = FIRST ? kConstBits - kPass1Bits : kConstBits + kPass1Bits;
  if (FIRST) {
    d[0] = (t10 + t11) << kPass1Bits;
    d[4 * stride] = (t10 - t11) << kPass1Bits;
  } else {
    d[0] = descale(t10 + t11, kPass1Bits);
    d[4 * stride] = descale(t10 - t11, kPass1Bits);
  }
  int z1 = (t12 + t13) * F0541;
  d[2 * stride] = descale(z1 + t13 * F0765, sh);
  d[6 * stride] = descale(z1 + t12 * (-F1847), sh);
  z1 = t4 + t7;
  int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * F1175;
  const int u4 = t4 * F0298, u5 = t5 * F2053, u6 = t6 * F3072, u7 = t7 * F1501;
  z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
  z3 += z5; z4 += z5;
  d[7 * stride] = descale(u4 + z1 + z3, sh);
  d[5 * stride] = descale(u5 + z2 + z4, sh);
  d[3 * stride] = descale(u6 + z2 + z3, sh);
  d[stride] = descale(u7 + z1 + z4, sh);
}

template <bool FIRST>
__device__ __forceinline__ void idct8(int* c, int stride) {
  int z2 = c[2 * stride], z3 = c[6 * stride];
  int z1 = (z2 + z3) * F0541;
  const int e2 = z1 + z3 * (-F1847), e3 = z1 + z2 * F0765;
  z2 = c[0]; z3 = c[4 * stride];
  const int e0 = (z2 + z3) << kConstBits, e1 = (z2 - z3) << kConstBits;
  const int t10 = e0 + e3, t13 = e0 - e3, t11 = e1 + e2, t12 = e1 - e2;
  int t0 = c[7 * stride], t1 = c[5 * stride], t2 = c[3 * stride], t3 = c[stride];
  z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * F1175;
  t0 *= F0298; t1 *= F2053; t2 *= F3072; t3 *= F1501;
  z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
  z3 += z5; z4 += z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  constexpr int sh = FIRST ? kConstBits - kPass1Bits : kConstBits + kPass1Bits + 3;
  c[0] = descale(t10 + t3, sh);          c[7 * stride] = descale(t10 - t3, sh);
  c[stride] = descale(t11 + t2, sh);     c[6 * stride] = descale(t11 - t2, sh);
  c[2 * stride] = descale(t12 + t1, sh); c[5 * stride] = descale(t12 - t1, sh);
  c[3 * stride] = descale(t13 + t0, sh); c[4 * stride] = descale(t13 - t0, sh);
}

__global__ void __launch_bounds__(128)
segmap_jpeg_kernel(const uint8_t* __restrict__ in_l, int images, int height, int width, uint8_t* __restrict__ back,
                   int* __restrict__ back_max) {
  const int bw = (width + 7) / 8, bh = (height + 7) / 8;
  const long long total = (long long)images * bh * bw;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int vmax = 0, img = -1;
  if (idx < total) {
    img = (int)(idx / ((long long)bh * bw));
    const int rem = (int)(idx % ((long long)bh * bw));
    const int by = rem / bw, bx = rem % bw;
    const uint8_t* src = in_l + (size_t)img * height * width;
    int d[64];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int y = min(by * 8 + r, height - 1);   // partial edge blocks: edge replication
#pragma unroll
      for (int c = 0; c < 8; ++c) d[r * 8 + c] = (int)src[(size_t)y * width + min(bx * 8 + c, width - 1)] - 128;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) fdct8<true>(d + r * 8, 1);
#pragma unroll
    for (int c = 0; c < 8; ++c) fdct8<false>(d + c, 8);
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const int q = kQuant75[i], qv = q << 3;
      const int mag = (abs(d[i]) + (qv >> 1)) / qv;
      d[i] = (d[i] < 0 ? -mag : mag) * q;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) idct8<true>(d + c, 8);
#pragma unroll
    for (int r = 0; r < 8; ++r) idct8<false>(d + r * 8, 1);
    uint8_t* dst = back + (size_t)img * height * width;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int y = by * 8 + r;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int x = bx * 8 + c;
        if (y < height && x < width) {
          const int v = min(255, max(0, d[r * 8 + c] + 128));
          dst[(size_t)y * width + x] = (uint8_t)v;
          vmax = max(vmax, v);
        }
      }
    }
  }
  // all threads of a warp usually belong to one image: one atomic per warp in that case
  const int img0 = __shfl_sync(0xffffffffu, img, 0);
  const bool uniform = __all_sync(0xffffffffu, img == img0);
  if (uniform) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0 && img >= 0 && vmax > 0) atomicMax(&back_max[img], vmax);
  } else if (img >= 0 && vmax > 0) {
    atomicMax(&back_max[img], vmax);
  }
}

// ---------------------------------------------------------------------------------------------------
// Pillow LANCZOS resize of the per-label 0/255 masks (Resample.c, 8 bits per channel): horizontal, then vertical
// ---------------------------------------------------------------------------------------------------
constexpr int kLzBits = 22;
__device__ __forceinline__ uint8_t clip8(int v) { return (uint8_t)min(255, max(0, v >> kLzBits)); }

// out[m, f, y, X] over the label maps [F, h, w]
__global__ void __launch_bounds__(256)
lanczos_h_kernel(const int* __restrict__ labels, const int* __restrict__ unique_labels, int masks, int frames, int h, int w,
                 int wout, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, uint8_t* __restrict__ out) {
  const long long total = (long long)masks * frames * h * wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % wout);
    const int y = (int)((i / wout) % h);
    const int f = (int)((i / ((long long)wout * h)) % frames);
    const int m = (int)(i / ((long long)wout * h * frames));
    const int lab = unique_labels[m];
    const int xmin = bounds[2 * X], n = bounds[2 * X + 1];
    const int* row = labels + ((size_t)f * h + y) * w + xmin;
    const int* k = kk + (size_t)X * ksize;
    int acc = 1 << (kLzBits - 1);
    for (int x = 0; x < n; ++x) acc += (row[x] == lab ? 255 : 0) * k[x];
    out[i] = clip8(acc);
  }
}

// out[m, f, Y, X] from tmp[m, f, y, X]
__global__ void __launch_bounds__(256)
lanczos_v_kernel(const uint8_t* __restrict__ tmp, long long images, int h, int hout, int wout, const int* __restrict__ bounds,
                 const int* __restrict__ kk, int ksize, uint8_t* __restrict__ out) {
  const long long total = images * hout * wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % wout);
    const int Y = (int)((i / wout) % hout);
    const long long im = i / ((long long)wout * hout);
    const int ymin = bounds[2 * Y], n = bounds[2 * Y + 1];
    const int* k = kk + (size_t)Y * ksize;
    const uint8_t* col = tmp + ((size_t)im * h + ymin) * wout + X;
    int acc = 1 << (kLzBits - 1);
    for (int y = 0; y < n; ++y) acc += (int)col[(size_t)y * wout] * k[y];
    out[i] = clip8(acc);
  }
}

// ---------------------------------------------------------------------------------------------------
// normalise + optional filter + arg-max over masks
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
segmap_argmax_kernel(const uint8_t* __restrict__ back, const int* __restrict__ back_max, int masks, int frames,
                     long long hw, const uint8_t* __restrict__ mask_resized, double filter_s,
                     const int* __restrict__ unique_labels, uint8_t* __restrict__ seg_raw, int* __restrict__ seg_index) {
  const long long total = (long long)frames * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / hw);
    double best = -1.0;
    int best_i = 0;
    for (int m = 0; m < masks; ++m) {
      const size_t off = ((size_t)m * frames + f) * hw + (size_t)(i - (long long)f * hw);
      const double denom = __dadd_rn((double)back_max[m * frames + f], 1e-5);
      double v = __ddiv_rn((double)back[off], denom);
      if (mask_resized) {
        const double mk = __ddiv_rn((double)mask_resized[off], 255.0);
        // difference_map * mask + filter_s * difference_map * (1 - mask), numpy's evaluation order
        v = __dadd_rn(__dmul_rn(v, mk), __dmul_rn(__dmul_rn(filter_s, v), __dsub_rn(1.0, mk)));
      }
      if (v > best) { best = v; best_i = m; }   // np.argmax: the first maximum
    }
    seg_raw[i] = (uint8_t)unique_labels[best_i];
    if (seg_index) seg_index[i] = best_i;
  }
}

static int grid_1d(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vidseg

using namespace vidseg;

VS_API int vidseg_segmap_difference(const uint8_t* frames_pos, const uint8_t* frames_neg, int images, int height, int width,
                                    uint8_t* diff_l, uint8_t* vis_l, uint8_t* back_l, int32_t* back_max, double* blur_max,
                                    void* stream) {
  VS_REQUIRE(images >= 0 && height >= 1 && width >= 1 && images <= 65535, "bad shape");
  if (images == 0) return 0;
  VS_REQUIRE(frames_pos && frames_neg && diff_l && back_l && back_max && blur_max, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  VS_CHECK_CUDA(cudaMemsetAsync(blur_max, 0, (size_t)images * 8, st));
  VS_CHECK_CUDA(cudaMemsetAsync(back_max, 0, (size_t)images * 4, st));
  dim3 grid((width + kBlurTile - 1) / kBlurTile, (height + kBlurTile - 1) / kBlurTile, images);
  const double px = (double)images * height * width;
  VS_LAUNCH_W(7.0 * px, segmap_blur_kernel, grid, 256, 0, st, frames_pos, frames_neg, height, width, diff_l,
              reinterpret_cast<unsigned long long*>(blur_max), 0);
  VS_POST_LAUNCH();
  if (vis_l) {
    VS_LAUNCH_W(7.0 * px, segmap_blur_kernel, grid, 256, 0, st, frames_pos, frames_neg, height, width, vis_l,
                reinterpret_cast<unsigned long long*>(blur_max), 1);
    VS_POST_LAUNCH();
  }
  const long long blocks = (long long)images * ((height + 7) / 8) * ((width + 7) / 8);
  VS_REQUIRE((blocks + 127) / 128 <= 0x7fffffffLL, "too many blocks");
  VS_LAUNCH_W(2.0 * px, segmap_jpeg_kernel, (int)((blocks + 127) / 128), 128, 0, st, diff_l, images, height, width, back_l,
              back_max);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_lanczos_masks(const int32_t* label_maps, const int32_t* unique_labels, int masks, int frames, int h, int w,
                                int height, int width, const int32_t* h_bounds, const int32_t* h_coeffs, int h_ksize,
                                const int32_t* v_bounds, const int32_t* v_coeffs, int v_ksize, uint8_t* tmp, uint8_t* out,
                                void* stream) {
  VS_REQUIRE(masks >= 0 && frames >= 0 && h >= 1 && w >= 1 && height >= 1 && width >= 1 && h_ksize >= 1 && v_ksize >= 1, "bad shape");
  VS_REQUIRE(width != w && height != h, "Pillow skips a pass whose size is unchanged: resize both axes");
  if (masks == 0 || frames == 0) return 0;
  VS_REQUIRE(label_maps && unique_labels && h_bounds && h_coeffs && v_bounds && v_coeffs && tmp && out, "null pointer");
  const long long n1 = (long long)masks * frames * h * width;
  VS_LAUNCH(lanczos_h_kernel, grid_1d(n1, 256), 256, 0, stream, label_maps, unique_labels, masks, frames, h, w, width,
            h_bounds, h_coeffs, h_ksize, tmp);
  VS_POST_LAUNCH();
  const long long n2 = (long long)masks * frames * height * width;
  VS_LAUNCH(lanczos_v_kernel, grid_1d(n2, 256), 256, 0, stream, tmp, (long long)masks * frames, h, height, width, v_bounds,
            v_coeffs, v_ksize, out);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_segmap_argmax(const uint8_t* back_l, const int32_t* back_max, int masks, int frames, int height, int width,
                                const uint8_t* mask_resized, double filter_s, const int32_t* unique_labels, uint8_t* seg_raw,
                                int32_t* seg_index, void* stream) {
  VS_REQUIRE(masks >= 1 && frames >= 0 && height >= 1 && width >= 1, "bad shape");
  if (frames == 0) return 0;
  VS_REQUIRE(back_l && back_max && unique_labels && seg_raw, "null pointer");
  const long long hw = (long long)height * width;
  VS_LAUNCH_W((double)masks * frames * hw * (mask_resized ? 2.0 : 1.0), segmap_argmax_kernel, grid_1d(frames * hw, 256), 256, 0,
              stream, back_l, back_max, masks, frames, hw, mask_resized, filter_s, unique_labels, seg_raw, seg_index);
  VS_POST_LAUNCH();
  return 0;
}
