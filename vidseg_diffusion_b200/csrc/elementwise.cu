// Streaming (HBM-bound) kernels between the tensor-core GEMMs of the UNet: every one of them reads fp32
// channels-last activations once and writes the split operand pair (hi | lo fp16) the next GEMM /
// convolution consumes through TMA, so normalisation, activation, concatenation and operand conversion
// cost one pass instead of four.
//
//   layernorm_split   nn.LayerNorm (sgm/modules/attention.py:567-569, norm1/2/3)                 one warp per token
//   geglu_split       GEGLU: value * gelu(gate), erf form (attention.py:95-96)
//   groupnorm_split   GroupNorm32 + SiLU of the ResBlocks (openaimodel.py:267-271, 300-303, util.py:276-278),
//                     Normalize of SpatialTransformer (attention.py:127-130), over the channel concatenation of two
//                     sources (the th.cat([h, hs.pop()], 1) of openaimodel.py:911) without materialising it
//   upsample2x_split  nearest x2 (openaimodel.py:153) fused with the operand conversion of the following conv
#define VS_FAMILY vidseg::kFamElementwise
#include "common.cuh"
#include <cstdlib>

#include "tc_common.cuh"

namespace vidseg {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------------
// LayerNorm -> split.  One warp per row; the row lives in registers (C <= 32 * 4 * kLnMaxQuads).
// ---------------------------------------------------------------------------------------------
constexpr int kLnMaxQuads = 16;  // C <= 2048

// Q = float4 per lane (C <= 128 * Q): the row lives in 4*Q registers, so the common C = 320 / 640 instantiations keep
// enough warps resident to cover the HBM latency (the single 16-quad version ran at 1.2 TB/s)
template <int Q>
__global__ void __launch_bounds__(256)
layernorm_split_kernel(const float* __restrict__ x, const float* __restrict__ row_bias, long long rows_per_bias,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       __half* __restrict__ hi, __half* __restrict__ lo, long long rows, int c, int packed8) {
  // TWO rows per warp, both loaded before either is reduced: twice the bytes in flight per warp (the one-row form ran
  // at half of the HBM roof on the C = 320 / 640 layers)
  constexpr int R = 2;
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  const int nq = c >> 2;  // float4 per row
  float4 v[R][Q];
  float s[R];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const long long row = row0 + rr;
    s[rr] = 0.f;
    if (row < rows) {
      const float4* xr = reinterpret_cast<const float4*>(x + row * c);
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        const int q = lane + 32 * i;
        if (q < nq) v[rr][i] = ld_stream_f4(xr + q);
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const long long row = row0 + rr;
    if (row >= rows) continue;
    // optional bias shared by groups of rows_per_bias consecutive rows, added BEFORE the normalisation: the
    // frame-position embedding / single-token cross-attention output of the temporal layers
    const float4* br = row_bias ? reinterpret_cast<const float4*>(row_bias + (row / rows_per_bias) * c) : nullptr;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const int q = lane + 32 * i;
      if (q < nq) {
        if (br) {
          const float4 bb = __ldg(br + q);
          v[rr][i].x += bb.x; v[rr][i].y += bb.y; v[rr][i].z += bb.z; v[rr][i].w += bb.w;
        }
        s[rr] += (v[rr][i].x + v[rr][i].y) + (v[rr][i].z + v[rr][i].w);
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const long long row = row0 + rr;
    if (row >= rows) continue;   // warp-uniform
    const float mean = warp_sum(s[rr]) / (float)c;
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const int q = lane + 32 * i;
      if (q < nq) {
        const float a = v[rr][i].x - mean, b = v[rr][i].y - mean, cc = v[rr][i].z - mean, d = v[rr][i].w - mean;
        s2 += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(s2) / (float)c + eps);
    __half* hr = hi + row * c;
    __half* lr = lo + row * c;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const int q = lane + 32 * i;
      if (q < nq) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + q);
        tc::store_split4(hr, lr, q * 4, (v[rr][i].x - mean) * rstd * g.x + bt.x, (v[rr][i].y - mean) * rstd * g.y + bt.y,
                         (v[rr][i].z - mean) * rstd * g.z + bt.z, (v[rr][i].w - mean) * rstd * g.w + bt.w, packed8 != 0,
                         tc::kAct8Sx, tc::kAct8Sl);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GEGLU -> split.  h [rows, 2*d] = (value | gate); out [rows, d].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
geglu_split_kernel(const float* __restrict__ h, __half* __restrict__ hi, __half* __restrict__ lo, long long rows,
                   int d, int packed8) {
  const int dq = d >> 2;
  const long long total = rows * dq;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / dq;
    const int q = (int)(i - row * dq);
    const float4* hr = reinterpret_cast<const float4*>(h + row * 2 * d);
    const float4 val = ld_stream_f4(hr + q);
    const float4 gate = ld_stream_f4(hr + dq + q);
    tc::store_split4(hi + row * d, lo + row * d, q * 4, val.x * gelu_erf_f(gate.x), val.y * gelu_erf_f(gate.y),
                     val.z * gelu_erf_f(gate.z), val.w * gelu_erf_f(gate.w), packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) -> split over the channel concatenation [x1 (c1) | x2 (c2)], channels-last.
// Pass 1: per (sample, pixel chunk) per-group sum and sum of squares, fixed summation order (deterministic).
// Pass 2: normalise, affine, activation, operand split.
// ---------------------------------------------------------------------------------------------
constexpr int kGnMaxChunks = 256; // pixel chunks per sample in pass 1 (chosen per call so that the grid fills the GPU)
constexpr int kGnThreads = 256;
constexpr int kGnMaxGroups = 32;
constexpr int kGnInFlight = 4;    // 16-byte loads in flight per thread in the apply pass

__device__ __forceinline__ float4 gn_load(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2,
                                          long long pix, int ch) {
  if (ch < c1) return ld_stream_f4(reinterpret_cast<const float4*>(x1 + pix * c1 + ch));
  return ld_stream_f4(reinterpret_cast<const float4*>(x2 + pix * c2 + (ch - c1)));
}

__global__ void __launch_bounds__(kGnThreads)
groupnorm_stats_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2, int hw, int groups,
                       double* __restrict__ partial /* [B, chunks, groups, 2] */) {
  const int kGnChunks = gridDim.x;
  extern __shared__ float sm[];  // [lanes][C][2]
  const int c = c1 + c2;
  const int nq = c >> 2;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int per = (hw + kGnChunks - 1) / kGnChunks;
  const int p0 = chunk * per, p1 = min(hw, p0 + per);
  // thread -> (pixel lane, first channel quad); with nq >= 256 every thread strides over the quads of one pixel
  const int lanes = (nq >= kGnThreads) ? 1 : kGnThreads / nq;
  const int qpl = (nq >= kGnThreads) ? kGnThreads : nq;  // threads per pixel lane
  const int pl = threadIdx.x / qpl, q0 = threadIdx.x % qpl;
  const bool active = pl < lanes;
  const long long base = (long long)b * hw;
  for (int q = q0; q < nq && active; q += qpl) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), s2 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pp = p0 + pl; pp < p1; pp += lanes) {
      const float4 v = gn_load(x1, c1, x2, c2, base + pp, q * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
    }
    float* dst = sm + ((size_t)pl * c + q * 4) * 2;
    dst[0] = s.x; dst[1] = s2.x; dst[2] = s.y; dst[3] = s2.y; dst[4] = s.z; dst[5] = s2.z; dst[6] = s.w; dst[7] = s2.w;
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x, cpg = c / groups;
    double a = 0.0, a2 = 0.0;
    for (int l = 0; l < lanes; ++l)
      for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
        a += (double)sm[((size_t)l * c + ch) * 2];
        a2 += (double)sm[((size_t)l * c + ch) * 2 + 1];
      }
    double* out = partial + (((size_t)b * kGnChunks + chunk) * groups + g) * 2;
    out[0] = a;
    out[1] = a2;
  }
}

__global__ void __launch_bounds__(kGnThreads)
groupnorm_apply_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2, int hw, int groups,
                       const double* __restrict__ partial, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, int silu, __half* __restrict__ out_hi,
                       __half* __restrict__ out_lo, __half* __restrict__ raw_hi, __half* __restrict__ raw_lo,
                       int pix_per_block, int kGnChunks, int packed8) {
  __shared__ float s_mean[kGnMaxGroups], s_rstd[kGnMaxGroups];
  const int c = c1 + c2;
  const int nq = c >> 2;
  const int b = blockIdx.y;
  if (threadIdx.x < groups) {
    double a = 0.0, a2 = 0.0;
    for (int ch = 0; ch < kGnChunks; ++ch) {
      const double* pp = partial + (((size_t)b * kGnChunks + ch) * groups + threadIdx.x) * 2;
      a += pp[0];
      a2 += pp[1];
    }
    const double cnt = (double)hw * (double)(c / groups);
    const double mean = a / cnt;
    double var = a2 / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  // per-channel affine form in shared memory: y = v * scale[c] + shift[c] with scale = rstd_g * gamma, shift = beta -
  // mean_g * scale (one FMA per element and no per-element group lookup; the first version spent four integer
  // divisions per quad on it)
  extern __shared__ float s_aff[];   // [c] scale, [c] shift
  const int cpg = c / groups;
  for (int ch = threadIdx.x; ch < c; ch += kGnThreads) {
    const int g = ch / cpg;
    const float sc = s_rstd[g] * __ldg(gamma + ch);
    s_aff[ch] = sc;
    s_aff[c + ch] = __ldg(beta + ch) - s_mean[g] * sc;
  }
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block, p1 = min(hw, p0 + pix_per_block);
  const int total = (p1 - p0) * nq;
  // (pixel, quad) of this thread advance incrementally: no division in the loop.  kGnInFlight items per trip, all loads
  // issued before any is processed: one 16-byte load in flight per thread left the kernel at ~3.5 TB/s, two at ~4.5.
  int pp = p0 + threadIdx.x / nq, q = threadIdx.x % nq;
  const int dp = kGnThreads / nq, dq = kGnThreads % nq;
  auto process = [&](int ppx, int qx, const float4& v) {
    const int ch = qx * 4;
    const long long pix = (long long)b * hw + ppx;
    const float4 sc = *reinterpret_cast<const float4*>(s_aff + ch);
    const float4 sh = *reinterpret_cast<const float4*>(s_aff + c + ch);
    float y0 = fmaf(v.x, sc.x, sh.x), y1 = fmaf(v.y, sc.y, sh.y), y2 = fmaf(v.z, sc.z, sh.z), y3 = fmaf(v.w, sc.w, sh.w);
    if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
    tc::store_split4(out_hi + pix * c, out_lo + pix * c, ch, y0, y1, y2, y3, packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
    if (raw_hi)
      tc::store_split4(raw_hi + pix * c, raw_lo + pix * c, ch, v.x, v.y, v.z, v.w, packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
  };
  for (int i = threadIdx.x; i < total; i += kGnInFlight * kGnThreads) {
    int ppi[kGnInFlight], qi[kGnInFlight];
    float4 v[kGnInFlight];
#pragma unroll
    for (int u = 0; u < kGnInFlight; ++u) {
      ppi[u] = pp; qi[u] = q;
      pp += dp; q += dq;
      if (q >= nq) { q -= nq; ++pp; }
    }
#pragma unroll
    for (int u = 0; u < kGnInFlight; ++u)
      if (i + u * kGnThreads < total) v[u] = gn_load(x1, c1, x2, c2, (long long)b * hw + ppi[u], qi[u] * 4);
#pragma unroll
    for (int u = 0; u < kGnInFlight; ++u)
      if (i + u * kGnThreads < total) process(ppi[u], qi[u], v[u]);
  }
}

// ---------------------------------------------------------------------------------------------
// nearest-neighbour x2 upsampling -> split.  x [B, H, W, C] fp32 -> out [B, 2H, 2W, C].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int batch,
                        int h, int w, int c, int packed8) {
  const int nq = c >> 2;
  const long long total = (long long)batch * h * w * nq;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int q = (int)(i % nq);
    const long long pix = i / nq;
    const int xw = (int)(pix % w);
    const int yh = (int)((pix / w) % h);
    const int b = (int)(pix / ((long long)w * h));
    const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(x + pix * c) + q);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long op = ((long long)b * 2 * h + 2 * yh + dy) * (2 * w) + 2 * xw + dx;
        tc::store_split4(hi + op * c, lo + op * c, q * 4, v.x, v.y, v.z, v.w, packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
      }
  }
}

// ---------------------------------------------------------------------------------------------
// row softmax -> split.  s [rows, cols] fp32 logits; out = softmax(s * scale) as the operand of the p.v GEMM of the
// first-stage AttnBlock (model.py:183-195).  One block per row; the row is read twice from L2 (max, then exp + sum
// kept in registers is not possible for 4096+ columns), written once.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_split_kernel(const float* __restrict__ s, float scale, __half* __restrict__ hi, __half* __restrict__ lo,
                          int cols, int packed8) {
  __shared__ float red[8];
  __shared__ float bcast;
  const float* row = s + (size_t)blockIdx.x * cols;
  const int nq = cols >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + q);
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]);
    bcast = t * scale;
  }
  __syncthreads();
  const float mx = bcast;
  float sum = 0.f;
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + q);
    sum += (expf(v.x * scale - mx) + expf(v.y * scale - mx)) + (expf(v.z * scale - mx) + expf(v.w * scale - mx));
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = bcast;
  __half* hr = hi + (size_t)blockIdx.x * cols;
  __half* lr = lo + (size_t)blockIdx.x * cols;
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + q);
    tc::store_split4(hr, lr, q * 4, expf(v.x * scale - mx) * inv, expf(v.y * scale - mx) * inv, expf(v.z * scale - mx) * inv,
                     expf(v.w * scale - mx) * inv, packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
  }
}

static int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace vidseg

using namespace vidseg;

static int layernorm_entry(const float* x, const float* row_bias, long long rows_per_bias, const float* gamma,
                           const float* beta, float eps, void* out_hi, void* out_lo, long long rows, int channels,
                           void* stream);

VS_API int vidseg_layernorm_split(const float* x, const float* gamma, const float* beta, float eps, void* out_hi,
                                  void* out_lo, long long rows, int channels, void* stream) {
  return layernorm_entry(x, nullptr, 1, gamma, beta, eps, out_hi, out_lo, rows, channels, stream);
}

VS_API int vidseg_layernorm_bias_split(const float* x, const float* row_bias, long long rows_per_bias, const float* gamma,
                                       const float* beta, float eps, void* out_hi, void* out_lo, long long rows,
                                       int channels, void* stream) {
  VS_REQUIRE(row_bias == nullptr || rows_per_bias >= 1, "rows_per_bias must be positive");
  return layernorm_entry(x, row_bias, rows_per_bias, gamma, beta, eps, out_hi, out_lo, rows, channels, stream);
}

static int layernorm_entry(const float* x, const float* row_bias, long long rows_per_bias, const float* gamma,
                           const float* beta, float eps, void* out_hi, void* out_lo, long long rows, int channels,
                           void* stream) {
  VS_REQUIRE(x && gamma && beta && out_hi && out_lo, "null pointer");
  VS_REQUIRE(rows >= 0 && channels >= 4 && channels % 4 == 0 && channels <= 128 * kLnMaxQuads, "C must be a multiple of 4, <= 2048");
  if (rows == 0) return 0;
  const int warps = 8;
  const long long grid = (rows + 2 * warps - 1) / (2 * warps);   // two rows per warp
  VS_REQUIRE(grid <= 0x7fffffffLL, "too many rows");
  const int quads = (channels / 4 + 31) / 32;
  const int p8 = operand_packed8(channels) ? 1 : 0;
#define VS_LN_LAUNCH(QQ)                                                                                                  \
  VS_LAUNCH_W(8.0 * rows * channels, layernorm_split_kernel<QQ>, (int)grid, warps * 32, 0, stream, x, row_bias, rows_per_bias, \
              gamma, beta, eps, (__half*)out_hi, (__half*)out_lo, rows, channels, p8)
  if (quads <= 1) VS_LN_LAUNCH(1);
  else if (quads <= 3) VS_LN_LAUNCH(3);
  else if (quads <= 5) VS_LN_LAUNCH(5);
  else if (quads <= 10) VS_LN_LAUNCH(10);
  else VS_LN_LAUNCH(16);
#undef VS_LN_LAUNCH
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_geglu_split(const float* h, void* out_hi, void* out_lo, long long rows, int d, void* stream) {
  VS_REQUIRE(h && out_hi && out_lo, "null pointer");
  VS_REQUIRE(rows >= 0 && d >= 4 && d % 4 == 0, "D must be a multiple of 4");
  if (rows == 0) return 0;
  VS_LAUNCH_W(12.0 * rows * d, geglu_split_kernel, grid_for(rows * (d / 4), 256), 256, 0, stream, h, (__half*)out_hi,
              (__half*)out_lo, rows, d, operand_packed8(d) ? 1 : 0);
  VS_POST_LAUNCH();
  return 0;
}

VS_API size_t vidseg_groupnorm_workspace_bytes(int batch, int groups) {
  if (batch <= 0 || groups <= 0) return 0;
  return (size_t)batch * kGnMaxChunks * groups * 2 * sizeof(double);
}

VS_API int vidseg_groupnorm_split(const float* x1, int c1, const float* x2, int c2, const float* gamma, const float* beta,
                                  float eps, int groups, int silu, void* out_hi, void* out_lo, void* raw_hi, void* raw_lo,
                                  int batch, int hw, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(x1 && gamma && beta && out_hi && out_lo && workspace, "null pointer");
  VS_REQUIRE((x2 == nullptr) == (c2 == 0), "x2 and c2 go together");
  VS_REQUIRE((raw_hi == nullptr) == (raw_lo == nullptr), "raw_hi and raw_lo go together");
  const int c = c1 + c2;
  VS_REQUIRE(batch >= 0 && hw >= 1 && c1 >= 4 && c1 % 4 == 0 && c2 % 4 == 0, "channel counts must be multiples of 4");
  VS_REQUIRE(groups >= 1 && groups <= kGnMaxGroups && c % groups == 0, "groups must divide C, <= 32");
  VS_REQUIRE(batch <= 65535, "batch exceeds the grid limit");
  if (batch == 0) return 0;
  VS_REQUIRE(workspace_bytes >= vidseg_groupnorm_workspace_bytes(batch, groups), "group-norm workspace too small");
  const int nq = c / 4;
  const int lanes = (nq >= kGnThreads) ? 1 : kGnThreads / nq;
  const size_t smem = (size_t)lanes * c * 2 * sizeof(float);
  VS_REQUIRE(smem <= 48 * 1024, "C too large for the group-norm statistics kernel");
  double* partial = reinterpret_cast<double*>(workspace);
  const double bytes = 4.0 * batch * hw * c;
  // more chunks for long samples (the video ResBlocks normalise over whole clips: batch 2, hw = T*h*w); a function
  // of hw only, so that a sample's statistics do not depend on which other samples share the batch
  // (a clip-wide sample of the video ResBlocks at batch 2 had 32 chunks = 64 blocks on 148 SMs: 279 us for a 73 MB tensor;
  // samples of 8192 pixels and more are therefore cut down to 32 pixels per chunk)
  int chunks = 32;
  const int min_pix = hw >= 8192 ? 32 : 256;
  while (chunks < kGnMaxChunks && hw / (chunks * 2) >= min_pix) chunks *= 2;
  VS_LAUNCH_W(bytes, groupnorm_stats_kernel, dim3(chunks, batch), kGnThreads, smem, stream, x1, c1, x2, c2, hw, groups,
              partial);
  VS_POST_LAUNCH();
  // pass 2: ~8K quads per block (twice the blocks of the first version: 3-4 resident blocks per SM did not cover the
  // HBM latency)
  static const int gn_quads = [] { const char* e = getenv("VIDSEG_GN_QUADS"); return e ? atoi(e) : 8192; }();
  int pix_per_block = (gn_quads + nq - 1) / nq;
  if (pix_per_block > hw) pix_per_block = hw;
  const int blocks = (hw + pix_per_block - 1) / pix_per_block;
  VS_REQUIRE((size_t)c * 8 <= 48 * 1024, "C too large for the group-norm apply kernel");
  VS_LAUNCH_W(bytes * (raw_hi ? 3.0 : 2.0), groupnorm_apply_kernel, dim3(blocks, batch), kGnThreads, (size_t)c * 8, stream, x1, c1, x2,
              c2, hw, groups, partial, gamma, beta, eps, silu, (__half*)out_hi, (__half*)out_lo, (__half*)raw_hi,
              (__half*)raw_lo, pix_per_block, chunks, operand_packed8(c) ? 1 : 0);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_upsample2x_split(const float* x, void* out_hi, void* out_lo, int batch, int h, int w, int c,
                                   void* stream) {
  VS_REQUIRE(x && out_hi && out_lo, "null pointer");
  VS_REQUIRE(batch >= 0 && h >= 1 && w >= 1 && c >= 4 && c % 4 == 0, "C must be a multiple of 4");
  if (batch == 0) return 0;
  const long long items = (long long)batch * h * w * (c / 4);
  VS_LAUNCH_W(4.0 * batch * h * w * c * 5.0, upsample2x_split_kernel, grid_for(items, 256), 256, 0, stream, x,
              (__half*)out_hi, (__half*)out_lo, batch, h, w, c, operand_packed8(c) ? 1 : 0);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_softmax_rows_split(const float* logits, float scale, void* out_hi, void* out_lo, long long rows, int cols,
                                     void* stream) {
  VS_REQUIRE(logits && out_hi && out_lo, "null pointer");
  VS_REQUIRE(rows >= 0 && rows <= 0x7fffffffLL && cols >= 4 && cols % 4 == 0, "cols must be a multiple of 4");
  if (rows == 0) return 0;
  VS_LAUNCH_W(12.0 * rows * cols, softmax_rows_split_kernel, (int)rows, 256, 0, stream, logits, scale, (__half*)out_hi,
              (__half*)out_lo, cols, operand_packed8(cols) ? 1 : 0);
  VS_POST_LAUNCH();
  return 0;
}
