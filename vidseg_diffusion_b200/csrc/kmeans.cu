// R2: K-means (sklearn KMeans(n_clusters=K, n_init=R).fit + .predict) on the device.
//
// Replaces scripts/sampling/feature_extraction.py:52-55.  The algorithm is sklearn's
// (sklearn/cluster/_kmeans.py, _k_means_lloyd.pyx, _k_means_common.pyx; see oracle/kmeans.py for
// the line-by-line restatement).  Design for B200:
//   * all R = n_init runs advance in lock-step: one launch does step c of the k-means++ seeding
//     of every run, one launch does the E-step of every unfinished run (X is read once for all
//     runs, the R*K centres are the "N" dimension of a skinny GEMM), so the whole fit is a few
//     hundred launches instead of a few thousand and each launch fills the 148 SMs;
//   * the random decisions are data-independent numpy draws made by the host mirror; the device
//     reproduces the decision arithmetic that can be reproduced exactly: sequential fp32 column
//     sums for the mean (numpy's add.reduce over axis 0), float64 distances rounded to fp32
//     (_euclidean_distances_upcast), a *sequential* fp32 cumsum for the searchsorted sampling;
//   * reductions that sklearn leaves to BLAS/OpenMP (potentials, centre sums, inertia) cannot be
//     reproduced bit-for-bit on any other machine; they are computed in float64 in a fixed order
//     (deterministic run-to-run) and rounded to fp32, i.e. at the centre of sklearn's error band.
//   * HBM-bound streaming: per Lloyd iteration X (N*D*4 bytes) is read once by the E-step and once
//     by the M-step; at the clip sizes of BASELINE.json X is L2-resident (36.7 MB < 126 MB).
#include <mutex>
#include <atomic>
#include <unordered_map>
#include <vector>

#define VS_FAMILY vidseg::kFamKMeans
#include "common.cuh"
#include "tc_common.cuh"

namespace vidseg {

constexpr int kMaxTrials = 8;
constexpr int kPotBlocks = 148;      // row blocks of the seeding distance kernels (one per SM)
constexpr int kInertiaBlocks = 64;   // row blocks per run of the inertia kernel
constexpr int kPotWarps = 8;         // warps per block there
constexpr int kPotParts = kPotBlocks * kPotWarps;
constexpr int kSlabs = 32;           // row slabs of the M-step partial sums
constexpr int kColTile = 128;        // columns per M-step block
constexpr int kScanChunk = 4096;
constexpr int kMqPlanes = 6;         // signed base-256 digits of the 48-bit fixed-point rows
constexpr int kMqFracBits = 35;      // q = rint(x * s * 2^35) with |x s| < 2^11  ->  |q| < 2^46
constexpr int kMqStages = 6;         // 32 KB per stage: one-hot tile + digit tile
constexpr int kMqSplit = 2;          // CTAs sharing one output tile's reduction range
constexpr int kMqBK = 128;           // rows (reduction index) per stage: 128 int8 = one 128-byte swizzle row
constexpr int kMqThreads = 192;      // TMA warp, MMA warp, 4 epilogue warps
constexpr int kOhRows = 16;          // points per thread in the one-hot kernel (one 16-byte store per cluster)
constexpr int kOhThreads = 128;

struct KmLayout {
  int n, d, k, r, t;
  float tol_rel;
  int max_iter;
  size_t xc, mean, var, xx, closest, newdist, potpart, cand, pot, centers, cnorm, center_idx, labels, part, partcnt,
      partial, changed, flags, tol, inertia_part, inertia, same, rand, first_idx, xs_hi, xs_lo, cs_hi, cs_lo, sdot,
      absmax, amb_list, upd_cnt, upd_argmax, upd_shift, upd_ticket, mq_planes, mq_onehot, mq_cnt, mq_part, km_cnt, reloc, total;
  int rk_pad;   // R*K rounded up to a multiple of 8 (row length of the tensor-core score matrix)
  int use_tc;   // E-step on the tensor cores (needs D % 8 == 0)
  int use_mq;   // M-step sums as an exact int8 tensor-core product (fixed-point digit planes x one-hot labels)
  int n_pad;    // rows rounded up to the 128-row reduction block of that product
  int mq_blocks;  // one-hot blocks per run (count partials)
  CUtensorMap tm_onehot, tm_planes;
};

// M-step variant: 1 = exact int8 tensor-core sums (default), 0 = float64 shared-memory sums (km_partial_kernel).
// vidseg_set_kmeans_mstep() or VIDSEG_KMEANS_MSTEP=0/1; read when a workspace is sized / prepared.
static std::atomic<int> g_km_mstep{-1};
static int km_mstep_mode() {
  int m = g_km_mstep.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("VIDSEG_KMEANS_MSTEP");
    m = (e && e[0] == '0') ? 0 : 1;
    g_km_mstep.store(m, std::memory_order_relaxed);
  }
  return m;
}

static KmLayout km_layout(int n, int d, int k, int r, int t) {
  KmLayout L{};
  L.n = n; L.d = d; L.k = k; L.r = r; L.t = t;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.xc = take((size_t)n * d * 4);
  L.mean = take((size_t)d * 4);
  L.var = take((size_t)d * 4);
  L.xx = take((size_t)n * 8);
  L.closest = take((size_t)r * n * 4);
  L.newdist = take((size_t)r * t * n * 4);
  L.potpart = take((size_t)r * t * kPotParts * 8);
  L.cand = take((size_t)r * kMaxTrials * 4);
  L.pot = take((size_t)r * 4);
  L.centers = take((size_t)r * k * d * 4);
  L.cnorm = take((size_t)r * k * 8);
  L.center_idx = take((size_t)r * k * 4);
  L.labels = take((size_t)r * n * 4);
  L.part = take((size_t)r * kSlabs * k * d * 8);
  L.partcnt = take((size_t)r * kSlabs * k * 4);
  L.partial = take((size_t)r * k * (d + 1) * 8);
  L.changed = take((size_t)r * 4);
  L.flags = take((size_t)r * 4 * 4);  // done, strict, n_iter, reserved
  L.tol = take(16);
  L.inertia_part = take((size_t)r * kInertiaBlocks * 8);
  L.inertia = take((size_t)r * 8);
  L.same = take((size_t)r * r * 4);
  L.rand = take((size_t)r * (k > 1 ? k - 1 : 1) * t * 8);
  L.first_idx = take((size_t)r * 4);
  L.rk_pad = (r * k + 7) / 8 * 8;
  L.use_tc = (d % 8 == 0 && n >= 512 && (size_t)r * k * 16 <= 48 * 1024) ? 1 : 0;
  L.absmax = take(16);
  L.upd_cnt = take((size_t)r * k * 8);
  L.upd_argmax = take((size_t)r * 4);
  L.upd_shift = take((size_t)r * k * 4);
  L.upd_ticket = take((size_t)r * 4);
  if (L.use_tc) {
    L.xs_hi = take((size_t)n * d * 2);
    L.xs_lo = take((size_t)n * d * 2);
    L.cs_hi = take((size_t)L.rk_pad * d * 2);
    L.cs_lo = take((size_t)L.rk_pad * d * 2);
    L.sdot = take((size_t)n * L.rk_pad * 4);
    L.amb_list = take((size_t)n * r * 16);   // int4 entries (fused E-step), int2 in the unfused form
  }
  L.n_pad = (n + kMqBK - 1) / kMqBK * kMqBK;
  L.mq_blocks = (L.n_pad / kOhRows + kOhThreads - 1) / kOhThreads;
  L.use_mq = (L.use_tc && km_mstep_mode() != 0 && n <= (1 << 17) && d % 4 == 0) ? 1 : 0;
  if (L.use_mq) {
    L.mq_planes = take((size_t)kMqPlanes * d * L.n_pad);
    L.mq_onehot = take((size_t)r * k * L.n_pad);
    L.mq_cnt = take((size_t)r * L.mq_blocks * k * 4);
    L.mq_part = take((size_t)kMqSplit * kMqPlanes * r * k * d * 4);
  }
  L.km_cnt = take((size_t)r * k * 4);
  L.reloc = take((size_t)r * 2 * 4);   // relocation ticket | done flag per run (fused update kernel)
  // (Tried and removed: the arg-min as the EPILOGUE of the score GEMM.  Measured on B200, N = 14336, R*K = 200: same labels,
  // 47.6 us against 49.4 us per E-step for GEMM + assign + resolve as separate full-occupancy kernels -- one tile per CTA
  // leaves the dependent compare / select chains to a handful of warps -- and it kept 1 400 lines of epilogue in the
  // instruction stream of every other GEMM.)
  L.total = off;
  return L;
}

static std::mutex g_km_mu;
static std::unordered_map<void*, KmLayout> g_km_registry;

static bool km_lookup(void* ws, KmLayout* out) {
  std::lock_guard<std::mutex> lk(g_km_mu);
  auto it = g_km_registry.find(ws);
  if (it == g_km_registry.end()) return false;
  *out = it->second;
  return true;
}

template <typename T>
__host__ __device__ inline T* at(void* ws, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off);
}

// ------------------------------------------------------------------------------------------
// prepare: column mean / variance exactly as numpy computes them for a C-contiguous float32
// [N, D] array reduced over axis 0 (row-by-row sequential fp32 accumulation, then a float64
// true_divide by the count cast back to fp32: numpy/_core/_methods.py _mean/_var).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) km_colstats_kernel(const float* __restrict__ x, int n, int d,
                                                         float* __restrict__ mean, float* __restrict__ var) {
  const int col = blockIdx.x * 32 + threadIdx.x;
  if (col >= d) return;
  const float* p = x + col;
  float s = 0.f;
  int i = 0;
  for (; i + 32 <= n; i += 32) {
    float v[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) v[u] = p[(size_t)(i + u) * d];
#pragma unroll
    for (int u = 0; u < 32; ++u) s = __fadd_rn(s, v[u]);
  }
  for (; i < n; ++i) s = __fadd_rn(s, p[(size_t)i * d]);
  const float m = (float)((double)s / (double)n);
  mean[col] = m;
  float s2 = 0.f;
  i = 0;
  for (; i + 32 <= n; i += 32) {
    float v[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) v[u] = p[(size_t)(i + u) * d];
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const float t = __fsub_rn(v[u], m);
      s2 = __fadd_rn(s2, __fmul_rn(t, t));
    }
  }
  for (; i < n; ++i) {
    const float t = __fsub_rn(p[(size_t)i * d], m);
    s2 = __fadd_rn(s2, __fmul_rn(t, t));
  }
  var[col] = (float)((double)s2 / (double)n);
}

// Same arithmetic (one sequential fp32 chain per column, rows in order), fed through an 8-deep cp.async ring so that
// ~900 rows of an 8-column strip (80 blocks at D = 640) are in flight instead of 32: the chain of adds, not the load latency, sets the time.
constexpr int kCsCols = 8, kCsRows = 128, kCsStages = 8, kCsThreads = 256;
__global__ void __launch_bounds__(kCsThreads) km_colstats_ring_kernel(const float* __restrict__ x, int n, int d,
                                                                      float* __restrict__ mean, float* __restrict__ var) {
  extern __shared__ __align__(16) float ring[];   // [stage][row][kCsCols]
  const int c0 = blockIdx.x * kCsCols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_chunks = (n + kCsRows - 1) / kCsRows;
  const int cols = min(kCsCols, d - c0);          // multiple of 4
  constexpr int kSegs = kCsCols / 4;
  auto issue = [&](int chunk) {
    if (chunk < n_chunks) {
      float* dst = ring + (size_t)(chunk % kCsStages) * kCsRows * kCsCols;
      for (int e = threadIdx.x; e < kCsRows * kSegs; e += kCsThreads) {
        const int row = e / kSegs, seg = e % kSegs;
        const int gi = chunk * kCsRows + row;
        if (gi < n && seg * 4 < cols) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + row * kCsCols + seg * 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(x + (size_t)gi * d + c0 + seg * 4) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float m = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    float s = 0.f;
    for (int c = 0; c < kCsStages - 1; ++c) issue(c);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kCsStages - 2) : "memory");
      __syncthreads();                        // chunk landed for everyone; the slot of chunk-1 is free again
      issue(chunk + kCsStages - 1);
      if (warp == 0 && lane < kCsCols) {
        const float* src = ring + (size_t)(chunk % kCsStages) * kCsRows * kCsCols + lane;
        const int rows = min(kCsRows, n - chunk * kCsRows);
        if (rows == kCsRows) {
          float v[kCsRows];
#pragma unroll
          for (int u = 0; u < kCsRows; ++u) v[u] = src[u * kCsCols];
          if (pass == 0) {
#pragma unroll
            for (int u = 0; u < kCsRows; ++u) s = __fadd_rn(s, v[u]);
          } else {
#pragma unroll
            for (int u = 0; u < kCsRows; ++u) { const float t = __fsub_rn(v[u], m); s = __fadd_rn(s, __fmul_rn(t, t)); }
          }
        } else {
          for (int u = 0; u < rows; ++u) {
            const float v = src[u * kCsCols];
            if (pass == 0) s = __fadd_rn(s, v);
            else { const float t = __fsub_rn(v, m); s = __fadd_rn(s, __fmul_rn(t, t)); }
          }
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (pass == 0) m = (float)((double)s / (double)n);   // only warp 0's value is used
    else if (warp == 0 && lane < cols) { mean[c0 + lane] = m; var[c0 + lane] = (float)((double)s / (double)n); }
  }
}

// tol = mean(var) * tol_rel (sklearn/_kmeans.py:285-294); also resets the per-run state.
__global__ void km_tol_reset_kernel(const float* __restrict__ var, int d, float tol_rel, float* __restrict__ tol,
                                    unsigned* __restrict__ absmax,
                                    int* __restrict__ flags, int* __restrict__ changed, int r) {
  __shared__ double red[32];
  double s = 0.0;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s += (double)var[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float m = (float)(tot / (double)d);
    tol[0] = __fmul_rn(m, tol_rel);
  }
  for (int i = threadIdx.x; i < r * 4; i += blockDim.x) flags[i] = 0;
  for (int i = threadIdx.x; i < r; i += blockDim.x) changed[i] = 0;
  if (threadIdx.x == 0) { absmax[0] = 0u; absmax[1] = 0u; absmax[2] = 0u; }   // |x| max, ambiguity count, its ticket
}

// xc = x - mean (fp32), xx = float64 squared norm of the centred fp32 row.  Warp per row.
__global__ void __launch_bounds__(256) km_center_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                        int n, int d, float* __restrict__ xc,
                                                        double* __restrict__ xx, unsigned* __restrict__ absmax) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + (size_t)row * d;
  float* o = xc + (size_t)row * d;
  double s = 0.0;
  float mx = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = __fsub_rn(xr[c], mean[c]);
    o[c] = v;
    s += (double)v * (double)v;
    mx = fmaxf(mx, fabsf(v));
  }
  s = warp_sum(s);
  mx = warp_max(mx);
  if (lane == 0) {
    xx[row] = s;
    atomicMax(absmax, __float_as_uint(mx));  // non-negative floats order like their bit patterns; max is order-free
  }
}

// ------------------------------------------------------------------------------------------
// k-means++ seeding
// ------------------------------------------------------------------------------------------
// distances of every point to the T candidates of each run, in float64 as
// _euclidean_distances_upcast (-2 x.y + ||y||^2 + ||x||^2), rounded to fp32, clamped at 0,
// min-ed with the current closest distances; per-warp float64 partial potentials.
// Register-tiled: a warp works on kPotRows rows at a time so that every candidate value read from shared memory feeds
// kPotRows FMAs (the first version issued one shared load per FMA and ran at 110 M warp instructions per launch);
// the candidate count is a template parameter so that no predicated-off work is issued.  Per (row, candidate) the
// lane partial sums run over ascending c and are combined by the same butterfly as before: results are unchanged.
constexpr int kPotRows = 4;
template <int T>
__global__ void __launch_bounds__(kPotWarps * 32)
km_kpp_dist_kernel(const float* __restrict__ xc, const double* __restrict__ xx, int n, int d,
                   const int* __restrict__ cand, const float* __restrict__ closest, int use_min,
                   float* __restrict__ newdist, double* __restrict__ potpart, int t_stride) {
  extern __shared__ double cs[];  // [T][d]
  const int r = blockIdx.y;
  const int* my_cand = cand + r * kMaxTrials;
  for (int e = threadIdx.x; e < T * d; e += blockDim.x) {
    const int t = e / d, c = e - t * d;
    cs[e] = (double)xc[(size_t)my_cand[t] * d + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kPotWarps + warp;
  double cand_xx = 0.0;
  if (lane < T) cand_xx = xx[my_cand[lane]];
  double potacc = 0.0;
  // rows gwarp, gwarp + kPotParts, ... as before (the per-warp potential partials keep their composition); four at a time
  for (int row0 = gwarp; row0 < n; row0 += kPotParts * kPotRows) {
    double acc[kPotRows][T];
#pragma unroll
    for (int q = 0; q < kPotRows; ++q)
#pragma unroll
      for (int t = 0; t < T; ++t) acc[q][t] = 0.0;
    const float* xr[kPotRows];
#pragma unroll
    for (int q = 0; q < kPotRows; ++q) {
      const int row = row0 + q * kPotParts;
      xr[q] = xc + (size_t)(row < n ? row : row0) * d;   // out-of-range rows recompute row0 and are discarded
    }
    for (int c = lane; c < d; c += 32) {
      double xv[kPotRows];
#pragma unroll
      for (int q = 0; q < kPotRows; ++q) xv[q] = (double)xr[q][c];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const double cv = cs[t * d + c];
#pragma unroll
        for (int q = 0; q < kPotRows; ++q) acc[q][t] = fma(xv[q], cv, acc[q][t]);
      }
    }
#pragma unroll
    for (int q = 0; q < kPotRows; ++q) {
      const int row = row0 + q * kPotParts;
      double mine = 0.0;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const double s = warp_sum(acc[q][t]);
        if (lane == t) mine = s;
      }
      if (lane < T && row < n) {
        double dd = -2.0 * mine;
        dd += cand_xx;
        dd += xx[row];
        float f = (float)dd;
        f = fmaxf(f, 0.f);
        const size_t rn = (size_t)r * n + row;
        if (use_min) f = fminf(closest[rn], f);
        newdist[((size_t)r * t_stride + lane) * n + row] = f;
        potacc += (double)f;
      }
    }
  }
  if (lane < T) potpart[((size_t)r * t_stride + lane) * kPotParts + gwarp] = potacc;
}

// The same distances for ALL runs from one pass over X: the kernel above re-reads X once per run (R x 36.7 MB from L2
// per seeding step, which is what bounded it).  Here the candidates of every run sit in shared memory as doubles
// ([R][T][d], 205 KB at R = 10, T = 4, d = 640), a warp keeps four rows of X in registers as doubles and walks the
// runs; per (row, candidate) the lane chains and the butterfly are those of the per-run kernel, so the distances are
// bit-identical to it.
// Sum NV values across the warp with the halving butterfly: after the step with lane mask o a lane keeps only the half
// of the values selected by its bit o, so 16 values cost 15 exchanges instead of 80.  Every value is still combined
// pairwise over lane masks 16, 8, 4, 2, 1 in that order -- the adds (and so the bits) of warp_sum(double).
// Result: value index (lane >> 1) for NV = 16, lane for NV = 32.
template <int NV>
__device__ __forceinline__ double warp_sum_halving(double (&v)[NV], int lane) {
  static_assert(NV == 16 || NV == 32, "16 or 32 values");
  double cur[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) cur[i] = v[i];
  int o = 16;
#pragma unroll
  for (int cnt = NV / 2; cnt >= 1; cnt >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const double lo = cur[i], hi = cur[i + cnt];
      const double send = up ? lo : hi, keep = up ? hi : lo;
      cur[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
    o >>= 1;
  }
  if (NV == 16) cur[0] += __shfl_xor_sync(0xffffffffu, cur[0], 1);
  return cur[0];
}

template <int T, int CH>   // CH = 32-column chunks held in registers per row (d <= 32 * CH)
__global__ void __launch_bounds__(kPotWarps * 32, 1)
km_kpp_dist_all_kernel(const float* __restrict__ xc, const double* __restrict__ xx, int n, int d, int runs,
                       const int* __restrict__ cand, const float* __restrict__ closest, int use_min,
                       float* __restrict__ newdist, double* __restrict__ potpart, int t_stride) {
  constexpr int NV = (kPotRows * T <= 16) ? 16 : 32;   // (row, candidate) pairs reduced together
  extern __shared__ double cs[];  // [runs][T][d], then candidate norms [runs][T], then potentials [warps][runs][NV]
  double* cxx = cs + (size_t)runs * T * d;
  double* pot_sm = cxx + runs * T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int rt = warp; rt < runs * T; rt += kPotWarps) {   // a warp per candidate row: all its loads in flight at once
    const int src = cand[(rt / T) * kMaxTrials + (rt % T)];
    const float* xr = xc + (size_t)src * d;
    float tmp[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) tmp[i] = (lane + 32 * i < d) ? xr[lane + 32 * i] : 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (lane + 32 * i < d) cs[(size_t)rt * d + lane + 32 * i] = (double)tmp[i];
    if (lane == 0) cxx[rt] = xx[src];
  }
  for (int e = threadIdx.x; e < kPotWarps * runs * NV; e += blockDim.x) pot_sm[e] = 0.0;
  __syncthreads();
  const int gwarp = blockIdx.x * kPotWarps + warp;
  double* my_pot = pot_sm + (size_t)warp * runs * NV;
  const int kk = (NV == 16) ? (lane >> 1) : lane;        // the pair this lane finalises
  const int my_q = kk / T, my_t = kk - my_q * T;
  const bool owner = (NV == 32 || (lane & 1) == 0) && kk < kPotRows * T;
  for (int row0 = gwarp; row0 < n; row0 += kPotParts * kPotRows) {
    double xv[kPotRows][CH];
#pragma unroll
    for (int q = 0; q < kPotRows; ++q) {
      const int row = row0 + q * kPotParts;
      const float* xr = xc + (size_t)(row < n ? row : row0) * d;   // out-of-range rows recompute row0 and are discarded
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c = lane + 32 * i;
        xv[q][i] = (c < d) ? (double)xr[c] : 0.0;
      }
    }
    const int my_row = row0 + my_q * kPotParts;
    const bool live = owner && my_row < n;
    const double my_xx = live ? xx[my_row] : 0.0;
    for (int r = 0; r < runs; ++r) {
      double acc[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) acc[e] = 0.0;
      const double* cr = cs + (size_t)r * T * d;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c = lane + 32 * i;
        if (32 * i < d) {                       // uniform
          const int cc = min(c, d - 1);         // a lane past the row end multiplies by its zero x value
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const double cv = cr[t * d + cc];
#pragma unroll
            for (int q = 0; q < kPotRows; ++q) acc[q * T + t] = fma(xv[q][i], cv, acc[q * T + t]);
          }
        }
      }
      const double mine = warp_sum_halving<NV>(acc, lane);
      if (live) {
        double dd = -2.0 * mine;
        dd += cxx[r * T + my_t];
        dd += my_xx;
        float f = (float)dd;
        f = fmaxf(f, 0.f);
        if (use_min) f = fminf(closest[(size_t)r * n + my_row], f);
        newdist[((size_t)r * t_stride + my_t) * n + my_row] = f;
        my_pot[r * NV + kk] += (double)f;
      }
    }
  }
  __syncwarp();
  for (int e = lane; e < runs * T; e += 32) {
    const int r = e / T, t = e - r * T;
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < kPotRows; ++q) tot += my_pot[r * NV + q * T + t];   // fixed order
    potpart[((size_t)r * t_stride + t) * kPotParts + gwarp] = tot;
  }
}

// The same distances on the FP64 tensor cores (mma.sync m8n8k4: 256 multiply-adds per warp instruction instead of 32).
// The kernel above is bound by instruction issue and latency, not by the FP64 pipe (ncu: 8 warps per SM at 255 registers,
// 25 % issue utilisation, 20 % FP64 pipe): 5 400 instructions per four rows.  Here a warp owns EIGHT rows and all R*T
// candidates (NT tiles of eight): per 16 columns one 16-byte load of its row, four conversions and, per candidate tile, two
// 16-byte shared-memory loads and four DMMAs -- 1 400 instructions per eight rows.  Within a 16-column chunk the K index of
// the MMA is permuted (sub-step s takes columns kc + 4 q + s of quad lane q) so that both fragments come from contiguous
// vector loads; A and B use the same permutation, the sum runs over the same products.  The order of the float64
// additions differs from the shuffle kernel's, i.e. a distance can differ from it in the last float64 bit before it is
// rounded to fp32 (as both differ from any BLAS); the potentials are float64 sums cast to fp32 by the selection kernel.
// Candidates sit in shared memory as doubles, rows padded by two: a 16-byte load is served per quarter warp = two fragment
// rows, each touching every other 16 bytes of a 128-byte line (lane q reads doubles 4q, 4q+1, then 4q+2, 4q+3), so the two
// rows must be 16 bytes apart modulo 128.
// Measured (ncu, R*T = 40): 64.8 us against 137 us for the shuffle kernel, 7.5 M instead of 31 M warp instructions; what
// bounds it now is the DMMA rate itself (1.43 M m8n8k4 per launch, ~12 cycles each per SM on this part).
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NT>   // candidate tiles of eight: NT * 8 >= runs * T
__global__ void __launch_bounds__(kPotWarps * 32, 1)
km_kpp_dist_mma_kernel(const float* __restrict__ xc, const double* __restrict__ xx, int n, int d, int runs, int T,
                       const int* __restrict__ cand, const float* __restrict__ closest, int use_min,
                       float* __restrict__ newdist, double* __restrict__ potpart, int t_stride) {
  extern __shared__ double cs[];   // [NT * 8][d + 2], then the candidate norms [NT * 8]
  const int dpad = d + 2, nc = runs * T;
  double* cxx = cs + (size_t)NT * 8 * dpad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int rt = warp; rt < NT * 8; rt += kPotWarps) {   // a warp per candidate row; padding candidates are zero rows
    const int src = rt < nc ? cand[(rt / T) * kMaxTrials + (rt % T)] : -1;
    for (int c = lane; c < d; c += 32) cs[(size_t)rt * dpad + c] = src >= 0 ? (double)xc[(size_t)src * d + c] : 0.0;
    if (lane == 0) cxx[rt] = src >= 0 ? xx[src] : 0.0;
  }
  __syncthreads();
  const int r8 = lane >> 2, q = lane & 3;
  const int gwarp = blockIdx.x * kPotWarps + warp, nwarps = gridDim.x * kPotWarps;
  double pot[NT][2];
#pragma unroll
  for (int t = 0; t < NT; ++t) pot[t][0] = pot[t][1] = 0.0;
  for (int g = gwarp; g * 8 < n; g += nwarps) {
    const int row = g * 8 + r8;
    const bool valid = row < n;
    const float* xr = xc + (size_t)(valid ? row : n - 1) * d + 4 * q;
    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;
    for (int kc = 0; kc < d; kc += 64) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)   // four chunks of the row in flight
        v[u] = (kc + 16 * u < d) ? __ldg(reinterpret_cast<const float4*>(xr + kc + 16 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kc + 16 * u < d) {   // uniform
          const double a0 = (double)v[u].x, a1 = (double)v[u].y, a2 = (double)v[u].z, a3 = (double)v[u].w;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const double2* bp = reinterpret_cast<const double2*>(cs + (size_t)(t * 8 + r8) * dpad + kc + 16 * u + 4 * q);
            const double2 b01 = bp[0], b23 = bp[1];
            dmma_m8n8k4(acc[t][0], acc[t][1], a0, b01.x);
            dmma_m8n8k4(acc[t][0], acc[t][1], a1, b01.y);
            dmma_m8n8k4(acc[t][0], acc[t][1], a2, b23.x);
            dmma_m8n8k4(acc[t][0], acc[t][1], a3, b23.y);
          }
        }
      }
    }
    const double xxr = valid ? xx[row] : 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ci = t * 8 + 2 * q + j;   // the accumulator fragment: row lane / 4, columns 2 (lane % 4) + {0, 1}
        if (valid && ci < nc) {
          const int rr = ci / T, tt = ci - rr * T;
          double dd = -2.0 * acc[t][j];
          dd += cxx[ci];
          dd += xxr;
          float f = (float)dd;
          f = fmaxf(f, 0.f);
          if (use_min) f = fminf(closest[(size_t)rr * n + row], f);
          newdist[((size_t)rr * t_stride + tt) * n + row] = f;
          pot[t][j] += (double)f;
        }
      }
  }
  // potential partial of this warp per candidate: over its eight row lanes in a fixed order
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = pot[t][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      const int ci = t * 8 + 2 * q + j;
      if (r8 == 0 && ci < nc) {
        const int rr = ci / T, tt = ci - rr * T;
        potpart[((size_t)rr * t_stride + tt) * kPotParts + gwarp] = v;
      }
    }
}

typedef void (*KppDistMmaFn)(const float*, const double*, int, int, int, int, const int*, const float*, int, float*, double*, int);
static KppDistMmaFn kpp_dist_mma_fn(int nt) {
  switch (nt) {
    case 1: return km_kpp_dist_mma_kernel<1>;
    case 2: return km_kpp_dist_mma_kernel<2>;
    case 3: return km_kpp_dist_mma_kernel<3>;
    case 4: return km_kpp_dist_mma_kernel<4>;
    case 5: return km_kpp_dist_mma_kernel<5>;
    default: return nullptr;
  }
}

typedef void (*KppDistAllFn)(const float*, const double*, int, int, int, const int*, const float*, int, float*, double*, int);
template <int CH>
static KppDistAllFn kpp_dist_all_fn_ch(int t) {
  switch (t) {
    case 1: return km_kpp_dist_all_kernel<1, CH>;
    case 2: return km_kpp_dist_all_kernel<2, CH>;
    case 3: return km_kpp_dist_all_kernel<3, CH>;
    case 4: return km_kpp_dist_all_kernel<4, CH>;
    case 5: return km_kpp_dist_all_kernel<5, CH>;
    case 6: return km_kpp_dist_all_kernel<6, CH>;
    default: return nullptr;
  }
}
static KppDistAllFn kpp_dist_all_fn(int t, int d) {
  if (d <= 128) return kpp_dist_all_fn_ch<4>(t);
  if (d <= 640) return kpp_dist_all_fn_ch<20>(t);
  return nullptr;
}

typedef void (*KppDistFn)(const float*, const double*, int, int, const int*, const float*, int, float*, double*, int);
static KppDistFn kpp_dist_fn(int t) {
  switch (t) {
    case 1: return km_kpp_dist_kernel<1>;
    case 2: return km_kpp_dist_kernel<2>;
    case 3: return km_kpp_dist_kernel<3>;
    case 4: return km_kpp_dist_kernel<4>;
    case 5: return km_kpp_dist_kernel<5>;
    case 6: return km_kpp_dist_kernel<6>;
    case 7: return km_kpp_dist_kernel<7>;
    default: return km_kpp_dist_kernel<8>;
  }
}

// one block per run: pick the best candidate of step c (np.argmin of the potentials), commit it
// (closest distances, centre row), then draw the candidates of step c+1:
//   rand_vals = u * pot;  ids = searchsorted(cumsum_fp32(closest), rand_vals)  (side='left')
// The cumsum is a strictly sequential fp32 accumulation (np.cumsum), done by one thread over
// shared-memory chunks; the searchsorted of a non-decreasing array is a count of elements < v.
__global__ void __launch_bounds__(1024)
km_kpp_select_scan_kernel(const float* __restrict__ xc, int n, int d, int k, int t_count, int t_stride, int c,
                          int prev_t, const double* __restrict__ potpart, const float* __restrict__ newdist,
                          float* __restrict__ closest, int* __restrict__ cand, float* __restrict__ pot,
                          float* __restrict__ centers, int* __restrict__ center_idx, const double* __restrict__ rand) {
  __shared__ __align__(16) float buf[kScanChunk];
  __shared__ float s_pot[kMaxTrials];
  __shared__ double s_vals[kMaxTrials];
  __shared__ int s_cnt[kMaxTrials];
  __shared__ int s_best;
  __shared__ float s_run[2];   // running sum entering chunk i lives in slot i & 1
  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  // potentials: warp t adds the kPotParts partials of candidate t -- lane-strided chains, then the fixed butterfly
  // (one thread walking the partials was 512 dependent global loads, 40 us)
  if ((tid >> 5) < prev_t) {
    const int t = tid >> 5;
    const double* pp = potpart + ((size_t)r * t_stride + t) * kPotParts;
    double sacc = 0.0;
    for (int i = tid & 31; i < kPotParts; i += 32) sacc += pp[i];
    sacc = warp_sum(sacc);
    if ((tid & 31) == 0) s_pot[t] = (float)sacc;
  }
  if (tid < kMaxTrials) s_cnt[tid] = 0;
  __syncthreads();
  if (tid == 0) {
    int best = 0;
    for (int t = 1; t < prev_t; ++t)
      if (s_pot[t] < s_pot[best]) best = t;
    s_best = best;
    pot[r] = s_pot[best];
    center_idx[r * k + c] = cand[r * kMaxTrials + best];
    s_run[0] = 0.f;
  }
  __syncthreads();
  const int best = s_best;
  const int best_row = cand[r * kMaxTrials + best];
  const bool more = (c + 1 < k);
  if (tid < t_count && more) s_vals[tid] = rand[((size_t)r * (k - 1) + c) * t_count + tid] * (double)s_pot[best];
  for (int cc = tid; cc < d; cc += blockDim.x)
    centers[((size_t)r * k + c) * d + cc] = xc[(size_t)best_row * d + cc];
  __syncthreads();  // cand[] of this step fully consumed before it is overwritten below
  const float* src = newdist + ((size_t)r * t_stride + best) * n;
  float* dst = closest + (size_t)r * n;
  // searchsorted(cumsum, v) = index of the first running sum >= v (the sums are non-decreasing).  One thread walks the
  // chain -- 4 cycles per dependent fp32 add is the floor -- and records the running sum only at 16-element group ends;
  // the crossing group of every v is then found in parallel and replayed from its recorded start value (same adds in
  // the same order: the same floats as np.cumsum).
  constexpr int kGrp = 16;
  __shared__ float bound[kScanChunk / kGrp];
  __shared__ int s_found[kMaxTrials];
  if (tid < kMaxTrials) s_found[tid] = 0;
  for (int base = 0; base < n; base += kScanChunk) {
    const int len = min(kScanChunk, n - base);
    for (int j = tid; j < len; j += blockDim.x) {
      const float v = src[base + j];
      buf[j] = v;
      dst[base + j] = v;
    }
    __syncthreads();
    if (!more) { __syncthreads(); continue; }
    const int ngroups = (len + kGrp - 1) / kGrp;
    const int slot = (base / kScanChunk) & 1;
    const float run_in = s_run[slot];
    if (tid == 0) {
      float run = run_in;
      int j = 0;
      // two groups per trip, no branch around the loads: group j+16 is fetched before the adds of group j and
      // group j+32 (clamped to the last full group -- a harmless re-read) before the adds of group j+16
      const int last_full = (len / kGrp - 1) * kGrp;
      float v[kGrp], w[kGrp];
      if (len >= kGrp) {
#pragma unroll
        for (int u = 0; u < kGrp; u += 4) *reinterpret_cast<float4*>(&v[u]) = *reinterpret_cast<const float4*>(&buf[u]);
      }
      for (; j + 2 * kGrp <= len; j += 2 * kGrp) {
#pragma unroll
        for (int u = 0; u < kGrp; u += 4)
          *reinterpret_cast<float4*>(&w[u]) = *reinterpret_cast<const float4*>(&buf[j + kGrp + u]);
#pragma unroll
        for (int u = 0; u < kGrp; ++u) run = __fadd_rn(run, v[u]);
        bound[j / kGrp] = run;
        const int nj = min(j + 2 * kGrp, last_full);
#pragma unroll
        for (int u = 0; u < kGrp; u += 4)
          *reinterpret_cast<float4*>(&v[u]) = *reinterpret_cast<const float4*>(&buf[nj + u]);
#pragma unroll
        for (int u = 0; u < kGrp; ++u) run = __fadd_rn(run, w[u]);
        bound[j / kGrp + 1] = run;
      }
      if (j + kGrp <= len) {   // odd number of full groups: v already holds group j
#pragma unroll
        for (int u = 0; u < kGrp; ++u) run = __fadd_rn(run, v[u]);
        bound[j / kGrp] = run;
        j += kGrp;
      }
      if (j < len) {
        for (; j < len; ++j) run = __fadd_rn(run, buf[j]);
        bound[ngroups - 1] = run;
      }
      s_run[slot ^ 1] = run;
    }
    __syncthreads();
    // warp t looks for the crossing group of value t
    const int wid = tid >> 5, lane = tid & 31;
    if (wid < t_count && !s_found[wid]) {
      const double v = s_vals[wid];
      int below = 0;
      for (int g = lane; g < ngroups; g += 32) below += ((double)bound[g] < v) ? 1 : 0;
      below = warp_sum(below);
      if (below < ngroups && lane == 0) {
        float run = (below == 0) ? run_in : bound[below - 1];
        const int j0 = below * kGrp, j1 = min(len, j0 + kGrp);
        int idx = j0;
        for (int j = j0; j < j1; ++j) {
          run = __fadd_rn(run, buf[j]);
          if ((double)run < v) ++idx; else break;
        }
        s_cnt[wid] = base + idx;
        s_found[wid] = 1;
      }
    }
    __syncthreads();
  }
  if (!more) return;
  if (tid < t_count) cand[r * kMaxTrials + tid] = s_found[tid] ? min(s_cnt[tid], n - 1) : n - 1;  // np.clip(.., None, n-1)
}

// ------------------------------------------------------------------------------------------
// Lloyd E-step: labels = argmin_j ( ||c_j||^2 - 2 x.c_j ), float64 accumulation, lowest j on ties.
// Block tile 64 rows x (all K centres of one run, 64 at a time), k-chunks of 32 staged in shared
// memory as doubles.  256 threads: tx = centre lane (16), ty = row group (16 x 4 rows).
// ------------------------------------------------------------------------------------------
constexpr int kBM = 64, kBK = 32, kBJ = 64;

__global__ void __launch_bounds__(256)
km_assign_kernel(const float* __restrict__ x, int n, int d, int k, int row_begin, int row_end,
                 const float* __restrict__ centers, const double* __restrict__ cnorm, int* __restrict__ labels,
                 int labels_stride, int* __restrict__ changed, const int* __restrict__ flags, int count_changes,
                 int only_nonstrict) {
  __shared__ double xs[kBM][kBK + 1];
  __shared__ double cs[kBJ][kBK + 1];
  const int r = blockIdx.y;
  if (flags) {
    const int done = flags[r * 4 + 0], strict = flags[r * 4 + 1];
    if (only_nonstrict) { if (strict) return; }
    else if (done) return;
  }
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = row_begin + blockIdx.x * kBM;
  const float* cr = centers + (size_t)r * k * d;
  const double* cn = cnorm + (size_t)r * k;
  double best_s[4];
  int best_j[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { best_s[q] = 1e300; best_j[q] = 0x7fffffff; }
  for (int j0 = 0; j0 < k; j0 += kBJ) {
    double acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int jt = 0; jt < 4; ++jt) acc[q][jt] = 0.0;
    for (int k0 = 0; k0 < d; k0 += kBK) {
      __syncthreads();
#pragma unroll
      for (int q = 0; q < (kBM * kBK) / 256; ++q) {
        const int e = tid + 256 * q;
        const int rr = e >> 5, cc = e & 31;
        const int grow = row0 + rr, gcol = k0 + cc;
        xs[rr][cc] = (grow < row_end && gcol < d) ? (double)x[(size_t)grow * d + gcol] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < (kBJ * kBK) / 256; ++q) {
        const int e = tid + 256 * q;
        const int jj = e >> 5, cc = e & 31;
        const int gj = j0 + jj, gcol = k0 + cc;
        cs[jj][cc] = (gj < k && gcol < d) ? (double)cr[(size_t)gj * d + gcol] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < kBK; ++kk) {
        double xv[4], cv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = xs[ty * 4 + q][kk];
#pragma unroll
        for (int jt = 0; jt < 4; ++jt) cv[jt] = cs[tx + 16 * jt][kk];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int jt = 0; jt < 4; ++jt) acc[q][jt] = fma(xv[q], cv[jt], acc[q][jt]);
      }
    }
#pragma unroll
    for (int jt = 0; jt < 4; ++jt) {
      const int gj = j0 + tx + 16 * jt;
      if (gj < k) {
        const double cnj = cn[gj];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double s = cnj - 2.0 * acc[q][jt];
          if (s < best_s[q] || (s == best_s[q] && gj < best_j[q])) { best_s[q] = s; best_j[q] = gj; }
        }
      }
    }
  }
  int nchanged = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double s = best_s[q];
    int j = best_j[q];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const double so = __shfl_xor_sync(0xffffffffu, s, o);
      const int jo = __shfl_xor_sync(0xffffffffu, j, o);
      if (so < s || (so == s && jo < j)) { s = so; j = jo; }
    }
    const int grow = row0 + ty * 4 + q;
    if (tx == 0 && grow < row_end) {
      int* lp = labels + (size_t)r * labels_stride + grow;
      if (count_changes && *lp != j) ++nchanged;
      *lp = j;
    }
  }
  if (count_changes) {
    nchanged = warp_sum(nchanged);
    if ((tid & 31) == 0 && nchanged) atomicAdd(&changed[r], nchanged);
  }
}

// ------------------------------------------------------------------------------------------
// E-step on the tensor cores, exact by construction.
//
// The dot products x_i . c_j of all runs are one skinny GEMM [N, D] x [D, R*K] (what sklearn hands to BLAS,
// _k_means_lloyd.pyx:_update_chunk_dense).  It runs as a split-fp16 tcgen05 GEMM whose result is only used as a
// FILTER: with |S_ij - x_i.c_j| <= (D 2^-22 + 2^-20) |x_i| |c_j| (operand split + worst-case fp32 accumulation of D
// products inside the tensor core) every centre whose approximate score is within the error band of the best one is a candidate,
// and only candidates are re-evaluated with the float64 FMA chain of km_assign_kernel (same order over d, same
// score cn_j - 2 acc, same lowest-index tie rule).  Almost every point has a single candidate, so the labels are
// bit-identical to the float64 kernel at a fraction of its cost.
// Operands are carried scaled by s = 2^(11 - exponent(max |xc|)) so that the fp16 pairs keep 22 bits; centres are
// means of rows, hence bounded by the same maximum.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
km_split_scaled_kernel(const float* __restrict__ x, size_t n_valid, size_t n_total, const unsigned* __restrict__ absmax,
                       __half* __restrict__ hi, __half* __restrict__ lo) {
  const float s = km_operand_scale(absmax[0]);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += stride) {
    __half h = __float2half_rn(0.f), l = h;
    if (i < n_valid) tc::split_f16(x[i] * s, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// One thread per (point, run).  The filter arithmetic is fp32 (half the issue cost and latency of float64 on this part,
// FP32 : FP64 = 2 : 1) and single pass: the two float64 passes over the K scores of every pair had made this kernel
// almost as expensive as the GEMM that feeds it (16.6 us against 22.7 us; now 10.8 us).
// The error radii are widened by 1 % and `slack` bounds the fp32 rounding of two scores, so the fp32 test can only
// flag MORE pairs / candidates than the float64 test would -- every flagged pair is settled exactly by the resolver,
// every unflagged label is the float64 arg-min.  Single pass: `best` is the running minimum, `minlow` the smallest
// lower bound sc_j - tau_j |x| among the other centres; the pair is ambiguous iff minlow <= best + tau_best |x| + slack.
// An ambiguous pair re-reads its K scores once and hands the resolver the set of centres inside the band as a bit mask.
__global__ void __launch_bounds__(256)
km_assign_tc_kernel(int k, int runs, int row_begin, int row_end, const double* __restrict__ cnorm,
                    const double* __restrict__ xx, const float* __restrict__ sdot, int ld,
                    const unsigned* __restrict__ absmax, int* __restrict__ labels, int labels_stride,
                    int* __restrict__ changed, const int* __restrict__ flags, int count_changes, int only_nonstrict,
                    double band, int* __restrict__ amb_count, int4* __restrict__ amb_list) {
  extern __shared__ float sm_f[];  // [runs*k] squared centre norms, [runs*k] error radii per unit |x|, [8] partial maxima
  pdl_wait();
  pdl_launch_dependents();
  float* sm_cn = sm_f;
  float* sm_tau = sm_f + runs * k;
  float* sm_red = sm_tau + runs * k;
  float mc = 0.f;
  for (int i = threadIdx.x; i < runs * k; i += blockDim.x) {
    const double c = cnorm[i];
    sm_cn[i] = (float)c;
    sm_tau[i] = (float)(1.01 * band * sqrt(c));
    mc = fmaxf(mc, (float)c);
  }
  mc = warp_max(mc);
  if ((threadIdx.x & 31) == 0) sm_red[threadIdx.x >> 5] = mc;
  __syncthreads();
  float cn_max = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) cn_max = fmaxf(cn_max, sm_red[w]);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(row_end - row_begin) * runs;
  if (idx >= total) return;
  const int row = row_begin + (int)(idx / runs);
  const int r = (int)(idx % runs);
  if (flags) {
    const int done = flags[r * 4 + 0], strict = flags[r * 4 + 1];
    if (only_nonstrict ? (strict != 0) : (done != 0)) return;
  }
  const float s = km_operand_scale(absmax[0]);
  const float m2inv = -2.0f / (s * s);   // a power of two: exact
  const float* sr = sdot + (size_t)row * ld + (size_t)r * k;
  const float* cn = sm_cn + r * k;
  const float* tau = sm_tau + r * k;
  const float xn = (float)sqrt(xx[row]) * 1.0000002f;   // rounded up: radii only grow
  // |fp32 score - exact score| <= 2^-23 (cn + |x||c|) per centre; two scores meet in every comparison
  const float slack = 0x1p-21f * (cn_max + xn * sqrtf(cn_max));
  float best = 3.0e38f, best_tau = 0.f, minlow = 3.0e38f;
  int best_j = 0;
  for (int j = 0; j < k; ++j) {
    const float sc = fmaf(sr[j], m2inv, cn[j]);
    const float low = sc - tau[j] * xn;
    const bool nb = sc < best;
    minlow = fminf(minlow, nb ? best - best_tau * xn : low);
    best_tau = nb ? tau[j] : best_tau;
    best_j = nb ? j : best_j;
    best = nb ? sc : best;
  }
  const float thr = best + best_tau * xn + slack;
  if (minlow <= thr) {  // not separated: defer to the resolver with the set of centres that can still win
    unsigned long long cand = ~0ull;
    if (k <= 64) {
      cand = 0ull;
      for (int j = 0; j < k; ++j)
        if (fmaf(sr[j], m2inv, cn[j]) - tau[j] * xn <= thr) cand |= 1ull << j;
    }
    const int slot = atomicAdd(amb_count, 1);
    amb_list[slot] = make_int4(row, r, (int)(unsigned)cand, (int)(unsigned)(cand >> 32));
    return;
  }
  int* lp = labels + (size_t)r * labels_stride + row;
  if (count_changes && *lp != best_j) atomicAdd(&changed[r], 1);
  *lp = best_j;
}

// float64 evaluation of the candidates of every ambiguous pair: score cn_j - 2 x.c_j with the dot product accumulated
// in float64 (lane-strided partial sums + butterfly; km_assign_kernel uses one sequential chain -- the two agree to
// ~1e-16 relative, far below any gap the fp32 data can produce), lowest index wins ties.  One warp per pair.
// `sdot` (the score matrix of the unfused E-step) narrows the evaluation to the centres inside the filter's error band;
// the fused E-step (gemm_tc.cu, KmEpilogue) keeps no scores and hands over that set as a bit mask per pair instead (a
// superset: the arg-min over it is the same).
constexpr int kRsMaxD = 768;   // the row is held in registers (24 floats per lane) by the fp32 stage of the resolver

__global__ void __launch_bounds__(256)
km_assign_resolve_kernel(const float* __restrict__ x, int d, int k, const float* __restrict__ centers,
                         const double* __restrict__ cnorm, const double* __restrict__ xx, const float* __restrict__ sdot,
                         int ld, const unsigned* __restrict__ absmax, int* __restrict__ labels, int labels_stride,
                         int* __restrict__ changed, int count_changes, double band, int* __restrict__ amb_count,
                         const int* __restrict__ amb_list, int entry_ints) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  pdl_wait();
  pdl_launch_dependents();
  const int count = *amb_count;
  const float s = km_operand_scale(absmax[0]);
  const double inv = 1.0 / ((double)s * (double)s);
  // fp32 stage: a lane-strided fp32 FMA chain of at most 24 terms plus a 5-level butterfly has
  // |dot32 - dot| <= 30 * 2^-24 * sum |x_i c_i| <= 30 * 2^-24 |x||c| (standard gamma_n bound + Cauchy-Schwarz): two
  // orders of magnitude tighter than the tensor-core filter, at the fp32 issue rate.  Only centres that this second
  // band cannot separate either go to the float64 chain.
  const double band2 = 2.0 * 30.0 * 0x1p-24 * 1.01;
  for (int e = warp; e < count; e += nwarps) {
    const int* ent = amb_list + (size_t)e * entry_ints;
    const int row = ent[0], r = ent[1];
    // fused E-step: the set of centres that can still win, as a bit mask (all of them when k > 64)
    unsigned long long cmask = (entry_ints == 4) ? ((unsigned long long)(unsigned)ent[2] | ((unsigned long long)(unsigned)ent[3] << 32))
                                                 : ~0ull;
    const float* sr = sdot ? sdot + (size_t)row * ld + (size_t)r * k : nullptr;
    const double* cn = cnorm + (size_t)r * k;
    const double xn = sqrt(xx[row]);
    const float* xr = x + (size_t)row * d;
    if (sr) {   // unfused E-step: the candidate set from the stored scores (k <= 64: as a mask, like the fused form)
      double best = 1e300, best_tau = 0.0;
      for (int j = 0; j < k; ++j) {
        const double sc = cn[j] - 2.0 * ((double)sr[j] * inv);
        if (sc < best) { best = sc; best_tau = band * sqrt(cn[j]); }
      }
      if (k <= 64) {
        cmask = 0ull;
        for (int j0 = 0; j0 < k; j0 += 32) {
          const int jl = j0 + lane;
          bool need = false;
          if (jl < k) {
            const double sc = cn[jl] - 2.0 * ((double)sr[jl] * inv);
            need = sc <= best + (best_tau + band * sqrt(cn[jl])) * xn;
          }
          cmask |= (unsigned long long)__ballot_sync(0xffffffffu, need) << j0;
        }
      }
    }
    double bs = 1e300;
    int bj = 0x7fffffff;
    if (k <= 64 && d <= kRsMaxD) {
      if (k < 64) cmask &= (1ull << k) - 1ull;
      // ---- stage A: fp32 scores of every candidate; lane (t & 31) keeps candidate t's result in slot t >> 5
      float xa[kRsMaxD / 32];
#pragma unroll
      for (int u = 0; u < kRsMaxD / 32; ++u) { const int c = lane + 32 * u; xa[u] = (c < d) ? xr[c] : 0.f; }
      double es0 = 1e300, es1 = 1e300, rho0 = 0.0, rho1 = 0.0;
      int js0 = -1, js1 = -1;
      double thr = 1e300;   // min over candidates of es + rho: an upper bound of the true minimum score
      int t = 0;
      for (unsigned long long m = cmask; m; m &= m - 1, ++t) {
        const int j = __ffsll((long long)m) - 1;
        const float* cr = centers + ((size_t)r * k + j) * d;
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < kRsMaxD / 32; ++u) { const int c = lane + 32 * u; acc = fmaf(xa[u], (c < d) ? cr[c] : 0.f, acc); }
        acc = warp_sum(acc);
        const double es = cn[j] - 2.0 * (double)acc;
        const double rho = band2 * xn * sqrt(cn[j]);
        thr = fmin(thr, es + rho);
        if (lane == (t & 31)) {
          if (t < 32) { es0 = es; rho0 = rho; js0 = j; } else { es1 = es; rho1 = rho; js1 = j; }
        }
      }
      // ---- stage B: survivors of the fp32 band; one survivor is the arg-min, several go to the float64 chain
      unsigned long long surv = 0ull;
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const int js = sl ? js1 : js0;
        const bool in = js >= 0 && (sl ? es1 - rho1 : es0 - rho0) <= thr;
        unsigned b = __ballot_sync(0xffffffffu, in);
        while (b) {   // translate lane positions back to centre indices
          const int src = __ffs(b) - 1;
          b &= b - 1;
          const int j = __shfl_sync(0xffffffffu, js, src);
          surv |= 1ull << j;
        }
      }
      if (__popcll(surv) == 1) {
        bj = __ffsll((long long)surv) - 1;
      } else {
        for (unsigned long long m = surv; m; m &= m - 1) {   // ascending j: the first minimum is kept
          const int j = __ffsll((long long)m) - 1;
          const float* cr = centers + ((size_t)r * k + j) * d;
          double acc = 0.0;
          for (int c0 = lane; c0 < d; c0 += 8 * 32) {
            float ca[8], xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int c = c0 + 32 * u;
              xv[u] = (c < d) ? xr[c] : 0.f;
              ca[u] = (c < d) ? cr[c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fma((double)xv[u], (double)ca[u], acc);
          }
          acc = warp_sum(acc);
          const double es = cn[j] - 2.0 * acc;
          if (es < bs) { bs = es; bj = j; }
        }
      }
    } else {
      // many clusters / very long rows: float64 evaluation of every centre inside the filter's band (or of all of them)
      double best = 1e300, best_tau = 0.0;
      if (sr)
        for (int j = 0; j < k; ++j) {
          const double sc = cn[j] - 2.0 * ((double)sr[j] * inv);
          if (sc < best) { best = sc; best_tau = band * sqrt(cn[j]); }
        }
      for (int j0 = 0; j0 < k; j0 += 32) {
        const int jl = j0 + lane;
        bool need = false;
        if (jl < k) {
          if (sr) {
            const double sc = cn[jl] - 2.0 * ((double)sr[jl] * inv);
            need = sc <= best + (best_tau + band * sqrt(cn[jl])) * xn;
          } else {
            need = (k > 64) || ((cmask >> jl) & 1ull);
          }
        }
        unsigned mask = __ballot_sync(0xffffffffu, need);
        while (mask) {
          const int j = j0 + __ffs(mask) - 1;
          mask &= mask - 1;
          const float* cr = centers + ((size_t)r * k + j) * d;
          double acc = 0.0;
          for (int c0 = lane; c0 < d; c0 += 8 * 32) {
            float xv[8], ca[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int c = c0 + 32 * u;
              xv[u] = (c < d) ? xr[c] : 0.f;
              ca[u] = (c < d) ? cr[c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fma((double)xv[u], (double)ca[u], acc);
          }
          acc = warp_sum(acc);
          const double es = cn[j] - 2.0 * acc;
          if (es < bs) { bs = es; bj = j; }   // ascending j: the first minimum is kept
        }
      }
    }
    if (lane == 0) {
      int* lp = labels + (size_t)r * labels_stride + row;
      if (count_changes && *lp != bj) atomicAdd(&changed[r], 1);
      *lp = bj;
    }
  }
  // the last block to finish clears the list for the next E-step (every block has read the count by then);
  // amb_count[1] is the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&amb_count[1], 1) == (int)gridDim.x - 1) { amb_count[0] = 0; amb_count[1] = 0; }
  }
}

// ------------------------------------------------------------------------------------------
// Lloyd M-step, part 1: per (run, row slab, 128-column tile) float64 cluster sums in shared memory.
// ------------------------------------------------------------------------------------------
template <int RPB>   // runs per block: the X tile is read once for RPB runs (the kernel is bound by re-reading X from L2)
__global__ void __launch_bounds__(kColTile)
km_partial_kernel(const float* __restrict__ x, int n, int d, int k, int runs, int row_begin, int row_end,
                  const int* __restrict__ labels, const int* __restrict__ flags, double* __restrict__ part,
                  int* __restrict__ partcnt) {
  extern __shared__ double acc[];  // [RPB][k][kColTile]
  const int rb = blockIdx.z * RPB;
  bool live[RPB];
  bool any = false;
#pragma unroll
  for (int q = 0; q < RPB; ++q) { live[q] = (rb + q < runs) && !flags[(rb + q) * 4 + 0]; any |= live[q]; }
  if (!any) return;
  const int slab = blockIdx.x, tile = blockIdx.y;
  const int tid = threadIdx.x;
  const int col = tile * kColTile + tid;
  const int rows = row_end - row_begin;
  const int per = (rows + kSlabs - 1) / kSlabs;
  const int r0 = row_begin + slab * per;
  const int r1 = min(row_end, r0 + per);
  for (int j = 0; j < RPB * k; ++j) acc[j * kColTile + tid] = 0.0;
  const int* lab[RPB];
#pragma unroll
  for (int q = 0; q < RPB; ++q) lab[q] = labels + (size_t)(live[q] ? rb + q : rb) * n;
  if (col < d) {
    int i = r0;
    // 16 rows in flight per thread: the loop is bound by L2 latency, not by the shared-memory adds
    for (; i + 16 <= r1; i += 16) {
      int l[RPB][16];
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = x[(size_t)(i + u) * d + col];
#pragma unroll
      for (int q = 0; q < RPB; ++q)
#pragma unroll
        for (int u = 0; u < 16; ++u) l[q][u] = lab[q][i + u];
#pragma unroll
      for (int q = 0; q < RPB; ++q)
        if (live[q]) {
#pragma unroll
          for (int u = 0; u < 16; ++u) acc[(q * k + l[q][u]) * kColTile + tid] += (double)v[u];
        }
    }
    for (; i < r1; ++i) {
      const double v = (double)x[(size_t)i * d + col];
#pragma unroll
      for (int q = 0; q < RPB; ++q)
        if (live[q]) acc[(q * k + lab[q][i]) * kColTile + tid] += v;
    }
#pragma unroll
    for (int q = 0; q < RPB; ++q)
      if (live[q]) {
        double* out = part + (((size_t)(rb + q) * kSlabs + slab) * k) * d + col;
        for (int j = 0; j < k; ++j) out[(size_t)j * d] = acc[(q * k + j) * kColTile + tid];
      }
  }
  if (tile == 0) {
#pragma unroll
    for (int q = 0; q < RPB; ++q)
      if (live[q])
        for (int j = tid; j < k; j += kColTile) {
          int cnt = 0;
          for (int i = r0; i < r1; ++i) cnt += (lab[q][i] == j);
          partcnt[((size_t)(rb + q) * kSlabs + slab) * k + j] = cnt;
        }
  }
}


// ------------------------------------------------------------------------------------------
// Lloyd M-step on the tensor cores, exact by construction.
//
// The cluster sums of all runs are ONE product  sums[r*K + j][c] = sum_i onehot[r*K + j][i] * x[i][c]  with a 0/1
// left operand.  Floating-point tensor-core accumulation would make the result depend on the (unspecified) order of
// the adds; integers do not.  Each centred value is therefore carried as a 48-bit fixed-point number
// q = rint(x * s * 2^35) (s = the power-of-two operand scale of the E-step, |x s| < 2^11), written once per fit as six
// SIGNED base-256 digit planes [6][D][N_pad] int8 (row index contiguous: the reduction dimension of the MMA).  Per
// iteration the labels become an int8 one-hot matrix [R*K][N_pad] and tcgen05.mma.kind::i8 accumulates
// digit-plane x one-hot in int32 (|sum| <= N * 128 < 2^31; N <= 2^17 keeps the 48-bit total inside int64): every product and every add is exact, so
// sum_p 256^p * acc_p is THE integer sum of the q_i, independent of tiling and order.  Values with magnitude below
// 2^-22 of the data's maximum lose bits below 2^-47 of that maximum in the conversion (the float64 chain of
// km_partial_kernel rounds at 2^-53 of the running sum); everything else is exact, which the float64 chain is not.
// Cost per iteration at N = 14336, D = 640, R*K = 200: 22 G int8 MACs on the tensor cores instead of 9 M float64
// read-modify-writes per SM in shared memory.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
km_quantize_kernel(const float* __restrict__ x, int n, int n_pad, int d, const unsigned* __restrict__ absmax,
                   int8_t* __restrict__ planes) {
  __shared__ __align__(16) int8_t dig[kMqPlanes][32][kMqBK + 4];   // [plane][column][row]
  const int i0 = blockIdx.x * kMqBK, c0 = blockIdx.y * 32;
  const double scale = (double)km_operand_scale(absmax[0]) * (double)(1ll << kMqFracBits);
  for (int e = threadIdx.x; e < kMqBK * 32; e += blockDim.x) {
    const int i = e >> 5, c = e & 31;
    long long q = 0;
    if (i0 + i < n && c0 + c < d) q = __double2ll_rn((double)x[(size_t)(i0 + i) * d + c0 + c] * scale);
#pragma unroll
    for (int p = 0; p < kMqPlanes; ++p) {
      const int dgt = (int)((q + 128) & 255) - 128;   // balanced digit in [-128, 127]
      q = (q - dgt) >> 8;
      dig[p][c][i] = (int8_t)dgt;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kMqPlanes * 32 * (kMqBK / 4); e += blockDim.x) {
    const int iw = e & 31, c = (e >> 5) & 31, p = e >> 10;
    if (c0 + c < d)
      *reinterpret_cast<uint32_t*>(planes + ((size_t)p * d + c0 + c) * n_pad + i0 + iw * 4) =
          *reinterpret_cast<const uint32_t*>(&dig[p][c][iw * 4]);
  }
}

// labels -> one-hot rows (every byte of an active run's rows is rewritten, so no clearing pass) + per-block counts
__global__ void __launch_bounds__(kOhThreads)
km_onehot_kernel(const int* __restrict__ labels, int n, int n_pad, int k, int row_begin, int row_end,
                 const int* __restrict__ flags, int8_t* __restrict__ onehot, int* __restrict__ cntpart) {
  extern __shared__ int hist[];  // [k]
  const int r = blockIdx.y;
  pdl_wait();
  pdl_launch_dependents();
  if (flags[r * 4 + 0]) return;
  for (int j = threadIdx.x; j < k; j += blockDim.x) hist[j] = 0;
  __syncthreads();
  const int i0 = (blockIdx.x * kOhThreads + threadIdx.x) * kOhRows;
  if (i0 < n_pad) {
    int lab[kOhRows];
    const int* lr = labels + (size_t)r * n;
#pragma unroll
    for (int u = 0; u < kOhRows; ++u) {
      const int i = i0 + u;
      lab[u] = (i >= row_begin && i < row_end) ? lr[i] : -1;
      if (lab[u] >= 0) atomicAdd(&hist[lab[u]], 1);
    }
    int8_t* orow = onehot + (size_t)r * k * n_pad + i0;
    for (int j = 0; j < k; ++j) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        w[q] = (lab[4 * q] == j ? 1u : 0u) | (lab[4 * q + 1] == j ? 0x100u : 0u) | (lab[4 * q + 2] == j ? 0x10000u : 0u) |
               (lab[4 * q + 3] == j ? 0x1000000u : 0u);
      *reinterpret_cast<uint4*>(orow + (size_t)j * n_pad) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) cntpart[((size_t)r * gridDim.x + blockIdx.x) * k + j] = hist[j];
}

struct MqParams {
  int kb_lo, kb_hi, kb_per_split, m_tiles, n_tiles, rk, d, k, runs;
  int* out;          // [split][plane][rk][d] int32
  const int* flags;
};

__global__ void __launch_bounds__(kMqThreads, 1)
km_mstep_mma_kernel(const __grid_constant__ CUtensorMap tm_onehot, const __grid_constant__ CUtensorMap tm_planes,
                    const MqParams p) {
  extern __shared__ __align__(1024) uint8_t mq_smem_raw[];
  constexpr int kTile = 128 * kMqBK;              // bytes of one operand tile (128 rows x 128 int8)
  constexpr int kStageBytes = 2 * kTile;
  int item = blockIdx.x;
  const int nt = item % p.n_tiles; item /= p.n_tiles;
  const int mt = item % p.m_tiles; item /= p.m_tiles;
  const int pl = item % kMqPlanes;
  const int sp = item / kMqPlanes;
  const int m0 = mt * 128, n0 = nt * 128;
  pdl_wait_async_proxy();
  pdl_launch_dependents();
  {  // a tile whose runs have all converged has nothing to add (uniform: decided before any barrier exists)
    const int r_lo = m0 / p.k, r_hi = min(p.runs - 1, (m0 + 127) / p.k);
    bool live = false;
    for (int r = r_lo; r <= r_hi; ++r) live |= (p.flags[r * 4 + 0] == 0);
    if (!live) return;
  }
  const int kb0 = p.kb_lo + sp * p.kb_per_split, kb1 = min(p.kb_hi, kb0 + p.kb_per_split);
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)mq_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kMqStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kMqStages;
  uint64_t* acc_bar = empty_bar + kMqStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tm_onehot);
    tc::prefetch_tmap(&tm_planes);
    for (int s = 0; s < kMqStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(acc_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<128>(tmem_ptr);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * kStageBytes;
        tc::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
        // the maps are declared over 16-bit elements: 64 of them per 128-byte row of int8
        tc::tma_load_2d(st, &tm_onehot, &full_bar[stage], kb * (kMqBK / 2), m0);
        tc::tma_load_3d(st + kTile, &tm_planes, &full_bar[stage], kb * (kMqBK / 2), n0, pl);
        if (++stage == kMqStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_i8(128, 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + stage * kStageBytes);
        const uint64_t a = tc::make_sw128_desc(sa), b = tc::make_sw128_desc(sa + kTile);
#pragma unroll
        for (int ks = 0; ks < kMqBK / 32; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 32 >> 4);
          tc::umma_i8(tmem_base, a + adv, b + adv, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
        }
        tc::umma_commit(&empty_bar[stage]);
        if (++stage == kMqStages) { stage = 0; phase ^= 1; }
      }
      tc::umma_commit(acc_bar);
    }
  } else {
    // epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (its hardware quadrant); lane = one output row
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    tc::mbar_wait(acc_bar, 0);
    tc::tc_fence_after();
    int* orow = p.out + (((size_t)sp * kMqPlanes + pl) * p.rk + row) * p.d + n0;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t v[32];
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cc * 32), v);
      tc::tmem_wait_ld();
      if (row < p.rk) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = n0 + cc * 32 + q * 4;
          if (c < p.d)   // d % 4 == 0
            *reinterpret_cast<int4*>(orow + cc * 32 + q * 4) =
                make_int4((int)v[4 * q], (int)v[4 * q + 1], (int)v[4 * q + 2], (int)v[4 * q + 3]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<128>(tmem_base);
}

// sum over splits and digit planes (exact in int64), back to float64, plus the counts -> partial [R, K, D+1].
// I64: the multi-GPU exchange form -- the integer sums themselves (and integer counts) as 8-byte words, so that an
// all-reduce(SUM) over ranks is exact and order independent; `tail` (R words behind the sums) takes the change counters.
template <bool I64>
__global__ void __launch_bounds__(256)
km_mstep_combine_kernel(const int* __restrict__ part, const int* __restrict__ cntpart, int cnt_blocks, int d, int k,
                        int rk, const int* __restrict__ flags, const unsigned* __restrict__ absmax,
                        void* __restrict__ partial_v, const int* __restrict__ changed_ws, int* __restrict__ changed_out,
                        void* __restrict__ tail, int* __restrict__ cnt_direct) {
  const int r = blockIdx.y, j = blockIdx.x;
  if (j == 0 && threadIdx.x == 0) {
    if (changed_out != nullptr && changed_out != changed_ws) changed_out[r] = changed_ws[r];
    if (tail != nullptr) {
      if (I64) reinterpret_cast<long long*>(tail)[r] = (long long)changed_ws[r];
      else reinterpret_cast<double*>(tail)[r] = (double)changed_ws[r];
    }
  }
  if (flags[r * 4 + 0]) return;
  const double inv = 1.0 / ((double)km_operand_scale(absmax[0]) * (double)(1ll << kMqFracBits));
  const size_t obase = ((size_t)r * k + j) * (d + 1);
  double* o = reinterpret_cast<double*>(partial_v) + obase;
  long long* oi = reinterpret_cast<long long*>(partial_v) + obase;
  const size_t row = (size_t)r * k + j;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    int v[kMqSplit * kMqPlanes];
#pragma unroll
    for (int e = 0; e < kMqSplit * kMqPlanes; ++e) v[e] = part[((size_t)e * rk + row) * d + c];   // all in flight
    long long tot = 0;
#pragma unroll
    for (int p = kMqPlanes - 1; p >= 0; --p) {
      long long dsum = 0;
#pragma unroll
      for (int s = 0; s < kMqSplit; ++s) dsum += v[s * kMqPlanes + p];
      tot = tot * 256 + dsum;
    }
    if (I64) oi[c] = tot;
    else o[c] = (double)tot * inv;   // |tot| < 2^24 * 2^46: the conversion may round once at 2^-53 relative
  }
  if (threadIdx.x == 0) {
    long long cnt = 0;
    if (cnt_direct) {   // fused E-step: one counter per (run, cluster), handed back cleared for the next iteration
      cnt = cnt_direct[(size_t)r * k + j];
      cnt_direct[(size_t)r * k + j] = 0;
    } else {
      for (int b = 0; b < cnt_blocks; ++b) cnt += cntpart[((size_t)r * cnt_blocks + b) * k + j];
    }
    if (I64) oi[d] = cnt;
    else o[d] = (double)cnt;
  }
}

// fixed-order reduction over slabs -> partial [R, K, D+1] (column D holds the count)
__global__ void __launch_bounds__(256)
km_reduce_kernel(const double* __restrict__ part, const int* __restrict__ partcnt, int d, int k,
                 const int* __restrict__ flags, double* __restrict__ partial, const int* __restrict__ changed_ws,
                 int* __restrict__ changed_out, double* __restrict__ tail) {
  const int r = blockIdx.y, j = blockIdx.x;
  if (j == 0 && threadIdx.x == 0) {
    if (changed_out != nullptr && changed_out != changed_ws) changed_out[r] = changed_ws[r];
    if (tail != nullptr) tail[r] = (double)changed_ws[r];
  }
  if (flags[r * 4 + 0]) return;
  double* o = partial + ((size_t)r * k + j) * (d + 1);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    double v[kSlabs];
#pragma unroll
    for (int sl = 0; sl < kSlabs; ++sl) v[sl] = part[(((size_t)r * kSlabs + sl) * k + j) * d + c];  // all in flight
    double s = 0.0;
#pragma unroll
    for (int sl = 0; sl < kSlabs; ++sl) s += v[sl];   // fixed order
    o[c] = s;
  }
  if (threadIdx.x == 0) {
    long long cnt = 0;
    for (int sl = 0; sl < kSlabs; ++sl) cnt += partcnt[((size_t)r * kSlabs + sl) * k + j];
    o[d] = (double)cnt;
  }
}

// ------------------------------------------------------------------------------------------
// Lloyd M-step, part 2: empty-cluster relocation, averaging, centre shift, convergence flags -- ONE launch.
//   One WARP per (run, cluster).  Every warp reads the K counts of its run (so the number of empty clusters and the
//   arg-max count need no pass of their own), takes its cluster's sums straight from where the M-step left them -- the
//   int32 digit-plane partial sums of the tensor-core product (SRC 2: no combine pass), the all-reduced exchange words
//   of the multi-GPU form (SRC 1) or a float64 [R, K, D+1] array (SRC 0) -- and performs _average_centers +
//   _center_shift for that centre; the last warp of a run to finish (ticket counter) adds the K squared shifts in
//   cluster order and sets the convergence flags.
//   Empty clusters (rare) with relocation allowed: the warps of the run first materialise their float64 rows in
//   `scratch`, the last one to arrive relocates (_relocate_empty_clusters_dense) while the others wait for its flag,
//   then all of them average from `scratch`.  All R*K warps are resident (R*K/4 blocks), so the wait cannot deadlock.
// ------------------------------------------------------------------------------------------
struct UpdArgs {
  int d, k, runs, n, max_iter, can_relocate, cnt_blocks, rk;
  const void* sums;            // SRC 0: double [R,K,D+1]; SRC 1: long long [R,K,D+1]; SRC 2: int [split][plane][rk][d]
  const int* cntpart;          // SRC 2: [R][cnt_blocks][K]
  const void* tail;            // SRC 0 / 1: change counters of the exchange words (double / long long [R]) or null
  const int* changed_in;       // used when tail is null
  int* changed_ws;
  double* scratch;             // float64 [R,K,D+1] rows of the relocation path (SRC 0: the sums array itself)
  float* centers; double* cnorm; int* flags; const float* tol; float* shiftsq; int* ticket;
  const unsigned* absmax; __half* cs_hi; __half* cs_lo;
  const float* x; const int* labels; float* dist;
  int* reloc_ticket; int* reloc_done;
};

constexpr int kUpdThreads = 128;   // one BLOCK per (run, cluster): 200 single warps left the loads of the sums latency-bound

template <int SRC>
__device__ __forceinline__ double upd_count(const UpdArgs& a, int r, int q) {
  if (SRC == 0) return reinterpret_cast<const double*>(a.sums)[((size_t)r * a.k + q) * (a.d + 1) + a.d];
  if (SRC == 1) return (double)reinterpret_cast<const long long*>(a.sums)[((size_t)r * a.k + q) * (a.d + 1) + a.d];
  long long cnt = 0;
  for (int b = 0; b < a.cnt_blocks; ++b) cnt += a.cntpart[((size_t)r * a.cnt_blocks + b) * a.k + q];
  return (double)cnt;
}

template <int SRC>
__device__ __forceinline__ double upd_sum(const UpdArgs& a, int r, int q, int c, double winv) {
  if (SRC == 0) return reinterpret_cast<const double*>(a.sums)[((size_t)r * a.k + q) * (a.d + 1) + c];
  if (SRC == 1) return (double)reinterpret_cast<const long long*>(a.sums)[((size_t)r * a.k + q) * (a.d + 1) + c] * winv;
  const int* part = reinterpret_cast<const int*>(a.sums);
  const size_t row = (size_t)r * a.k + q;
  int v[kMqSplit * kMqPlanes];
#pragma unroll
  for (int e = 0; e < kMqSplit * kMqPlanes; ++e) v[e] = part[((size_t)e * a.rk + row) * a.d + c];   // all in flight
  long long tot = 0;
#pragma unroll
  for (int p = kMqPlanes - 1; p >= 0; --p) {
    long long dsum = 0;
#pragma unroll
    for (int sp = 0; sp < kMqSplit; ++sp) dsum += v[sp * kMqPlanes + p];
    tot = tot * 256 + dsum;
  }
  return (double)tot * winv;   // |tot| < 2^24 * 2^46: the conversion may round once at 2^-53 relative
}

// counts of the run -> {this cluster's count, number of empty clusters, first arg-max, its count} by warp 0 into shared memory
template <typename F>
__device__ __forceinline__ void upd_run_stats(F count_of, int k, int j, double* s_stat) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    double bc = -1.0, mine = 0.0;
    int bi = 0x7fffffff, ne = 0;
    for (int q = lane; q < k; q += 32) {
      const double c = count_of(q);
      ne += (c == 0.0);
      if (c > bc) { bc = c; bi = q; }   // ascending q per lane: first maximum
      if (q == j) mine = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
      ne += __shfl_xor_sync(0xffffffffu, ne, o);
      mine += __shfl_xor_sync(0xffffffffu, mine, o);   // exactly one lane holds the value, the rest 0
    }
    if (lane == 0) { s_stat[0] = mine; s_stat[1] = (double)ne; s_stat[2] = (double)bi; s_stat[3] = bc; }
  }
  __syncthreads();
}

template <int SRC>
__global__ void __launch_bounds__(kUpdThreads)
km_update_fused_kernel(const UpdArgs a) {
  __shared__ double s_stat[4];
  __shared__ double s_red_v[kUpdThreads];
  __shared__ int s_red_i[kUpdThreads];
  __shared__ int s_ticket;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = a.d, k = a.k;
  const int r = blockIdx.x / k, j = blockIdx.x - r * k;
  pdl_wait();
  pdl_launch_dependents();
  if (a.flags[r * 4 + 0]) return;   // set only by the finalising block of an EARLIER launch (uniform over the block)
  const double winv = (SRC == 0) ? 1.0 : 1.0 / ((double)km_operand_scale(a.absmax[0]) * (double)(1ll << kMqFracBits));
  upd_run_stats([&](int q) { return upd_count<SRC>(a, r, q); }, k, j, s_stat);
  double my_cnt = s_stat[0];
  int nempty = (int)s_stat[1], am = (int)s_stat[2];
  double am_cnt = s_stat[3];
  if (nempty > 0 && !a.can_relocate && j == 0 && tid == 0) a.flags[r * 4 + 3] = 1;  // sharded caller must redo the fit unsharded
  const bool reloc = nempty > 0 && a.can_relocate && SRC != 1;
  double* srow_base = a.scratch + (size_t)r * k * (d + 1);
  if (reloc) {
    // ---- rare path: sklearn's _relocate_empty_clusters_dense needs the whole run
    if (SRC == 2) {
      double* mine = srow_base + (size_t)j * (d + 1);
      for (int c = tid; c < d; c += kUpdThreads) mine[c] = upd_sum<SRC>(a, r, j, c, winv);
      if (tid == 0) mine[d] = my_cnt;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&a.reloc_ticket[r], 1);
    __syncthreads();
    if (s_ticket == k - 1) {
      __threadfence();
      const int* lab = a.labels + (size_t)r * a.n;
      const float* cr = a.centers + (size_t)r * k * d;
      float* dist = a.dist + (size_t)r * a.n;
      // distances of every point to its (old) centre, the n_empty farthest points (descending, lowest index on ties)
      // seed the empty clusters (ascending cluster id)
      for (int i = tid; i < a.n; i += kUpdThreads) {
        const float* xr = a.x + (size_t)i * d;
        const float* cc = cr + (size_t)lab[i] * d;
        double sdist = 0.0;
        for (int c = 0; c < d; ++c) { const double tt = (double)xr[c] - (double)cc[c]; sdist = fma(tt, tt, sdist); }
        dist[i] = (float)sdist;
      }
      __syncthreads();
      int next_empty = 0;
      for (int e = 0; e < nempty; ++e) {
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int i = tid; i < a.n; i += kUpdThreads) {
          const double v = (double)dist[i];
          if (v > bv) { bv = v; bi = i; }
        }
        s_red_v[tid] = bv; s_red_i[tid] = bi;
        __syncthreads();
        for (int o = kUpdThreads / 2; o > 0; o >>= 1) {
          if (tid < o) {
            const double v2 = s_red_v[tid + o]; const int i2 = s_red_i[tid + o];
            if (v2 > s_red_v[tid] || (v2 == s_red_v[tid] && i2 < s_red_i[tid])) { s_red_v[tid] = v2; s_red_i[tid] = i2; }
          }
          __syncthreads();
        }
        const double maxv = s_red_v[0];
        const int far = s_red_i[0];
        __syncthreads();
        if (e == 0 && maxv == 0.0) break;  // np.max(distances) == 0 -> return
        while (srow_base[(size_t)next_empty * (d + 1) + d] != 0.0) ++next_empty;  // uniform across the block
        const int new_id = next_empty++;
        const int old_id = lab[far];
        for (int c = tid; c < d; c += kUpdThreads) {
          const double xv = (double)a.x[(size_t)far * d + c];
          srow_base[(size_t)old_id * (d + 1) + c] -= xv;
          srow_base[(size_t)new_id * (d + 1) + c] = xv;
        }
        __syncthreads();
        if (tid == 0) {
          srow_base[(size_t)new_id * (d + 1) + d] = 1.0;
          srow_base[(size_t)old_id * (d + 1) + d] -= 1.0;
          dist[far] = -1.f;
        }
        __syncthreads();
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicExch(&a.reloc_done[r], 1);
    } else {
      if (tid == 0)
        while (atomicAdd(&a.reloc_done[r], 0) == 0) __nanosleep(200);
    }
    __syncthreads();
    __threadfence();
    upd_run_stats([&](int q) { return __ldcg(&srow_base[(size_t)q * (d + 1) + d]); }, k, j, s_stat);
    my_cnt = s_stat[0]; nempty = (int)s_stat[1]; am = (int)s_stat[2]; am_cnt = s_stat[3];
  }
  float* cr = a.centers + (size_t)r * k * d;
  // _average_centers + _center_shift
  double ss = 0.0, nn = 0.0;
  const bool has = my_cnt > 0.0;
  // centers[j] = centers[argmax_weight] for an empty cluster: averaged already if argmax < j, raw sum otherwise
  const int src = has ? j : am;
  const double div = has ? my_cnt : ((am < j && am_cnt > 0.0) ? am_cnt : 1.0);
  float* crow = cr + (size_t)j * d;
  // the tensor-core E-step reads the centres as scaled fp16 pairs: written here instead of by a pass of their own
  const float op_scale = a.cs_hi ? km_operand_scale(a.absmax[0]) : 1.f;
  __half* hrow = a.cs_hi ? a.cs_hi + ((size_t)r * k + j) * d : nullptr;
  __half* lrow = a.cs_hi ? a.cs_lo + ((size_t)r * k + j) * d : nullptr;
  for (int c0 = tid; c0 < d; c0 += 3 * kUpdThreads) {   // three independent elements in flight per thread
    double raw[3];
    float ov[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int c = c0 + kUpdThreads * u;
      raw[u] = 0.0; ov[u] = 0.f;
      if (c < d) {
        raw[u] = reloc ? __ldcg(&srow_base[(size_t)src * (d + 1) + c]) : upd_sum<SRC>(a, r, src, c, winv);
        ov[u] = crow[c];
      }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int c = c0 + kUpdThreads * u;
      if (c < d) {
        const float nv = (div == 1.0 && !has) ? (float)raw[u] : (float)(raw[u] / div);
        const double df = (double)nv - (double)ov[u];
        ss = fma(df, df, ss);
        nn = fma((double)nv, (double)nv, nn);
        crow[c] = nv;
        if (hrow) {
          __half h, l;
          tc::split_f16(nv * op_scale, h, l);
          hrow[c] = h;
          lrow[c] = l;
        }
      }
    }
  }
  ss = warp_sum(ss);
  nn = warp_sum(nn);
  if (lane == 0) { s_red_v[warp] = ss; s_red_v[8 + warp] = nn; }
  __syncthreads();
  if (tid == 0) {
    ss = 0.0; nn = 0.0;
    for (int w2 = 0; w2 < kUpdThreads / 32; ++w2) { ss += s_red_v[w2]; nn += s_red_v[8 + w2]; }   // fixed order
    const float sh = (float)sqrt(ss);
    a.shiftsq[(size_t)r * k + j] = __fmul_rn(sh, sh);
    a.cnorm[(size_t)r * k + j] = nn;
    __threadfence();
    if (atomicAdd(&a.ticket[r], 1) == k - 1) {
      __threadfence();
      float tot = 0.f;
      for (int jj = 0; jj < k; ++jj) tot = __fadd_rn(tot, __ldcg(&a.shiftsq[(size_t)r * k + jj]));
      const int it = a.flags[r * 4 + 2] + 1;
      a.flags[r * 4 + 2] = it;
      int changed;
      if (a.tail == nullptr) changed = a.changed_in[r];
      else if (SRC == 1) changed = (int)reinterpret_cast<const long long*>(a.tail)[r];
      else changed = (int)reinterpret_cast<const double*>(a.tail)[r];
      if (changed == 0) { a.flags[r * 4 + 1] = 1; a.flags[r * 4 + 0] = 1; }
      else if (tot <= a.tol[0]) { a.flags[r * 4 + 0] = 1; }
      else if (it >= a.max_iter) { a.flags[r * 4 + 0] = 1; }
      a.changed_ws[r] = 0;
      a.ticket[r] = 0;
      a.reloc_ticket[r] = 0;
      a.reloc_done[r] = 0;
    }
  }
}

// squared norms of the initial centres (after seeding)
__global__ void km_cnorm_kernel(const float* __restrict__ centers, int d, int total, double* __restrict__ cnorm) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= total) return;
  double s = 0.0;
  for (int c = lane; c < d; c += 32) { const double v = (double)centers[(size_t)j * d + c]; s = fma(v, v, s); }
  s = warp_sum(s);
  if (lane == 0) cnorm[j] = s;
}

// inertia: sum over rows of ||x - c_label||^2 in float64, per-block partials in a fixed order
__global__ void __launch_bounds__(256)
km_inertia_kernel(const float* __restrict__ x, int n, int d, int k, int row_begin, int row_end,
                  const float* __restrict__ centers, const int* __restrict__ labels, double* __restrict__ part) {
  __shared__ double red[8];
  const int r = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * 8 + warp;
  double acc = 0.0;
  for (int row = row_begin + gwarp; row < row_end; row += kInertiaBlocks * 8) {
    const float* xr = x + (size_t)row * d;
    const float* cc = centers + ((size_t)r * k + labels[(size_t)r * n + row]) * d;
    double s = 0.0;
    for (int c = lane; c < d; c += 32) { const double t = (double)xr[c] - (double)cc[c]; s = fma(t, t, s); }
    acc += warp_sum(s);
  }
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    part[(size_t)r * kInertiaBlocks + blockIdx.x] = t;
  }
}
__global__ void km_inertia_reduce_kernel(const double* __restrict__ part, double* __restrict__ out) {
  const int r = blockIdx.x;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < kInertiaBlocks; ++b) t += part[(size_t)r * kInertiaBlocks + b];
    out[r] = t;
  }
}

// same[a][b] = 1 iff labels_a -> labels_b is a function on rows [row_begin,row_end)
// (_k_means_common.pyx:_is_same_clustering(labels_a, labels_b)).  grid (R, R).
__global__ void __launch_bounds__(256)
km_same_kernel(const int* __restrict__ labels, int n, int k, int row_begin, int row_end, int* __restrict__ same) {
  extern __shared__ int mapping[];  // [k]
  __shared__ int ok;
  const int a = blockIdx.x, b = blockIdx.y, r = gridDim.x;
  for (int j = threadIdx.x; j < k; j += blockDim.x) mapping[j] = -1;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  const int* la = labels + (size_t)a * n;
  const int* lb = labels + (size_t)b * n;
  for (int i = row_begin + threadIdx.x; i < row_end; i += blockDim.x) {
    const int l1 = la[i], l2 = lb[i];
    const int old = atomicCAS(&mapping[l1], -1, l2);
    if (old != -1 && old != l2) ok = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) same[a * r + b] = ok;
}

// un-centre the winning run: centers_out = centers[best] + mean (fp32 add, sklearn :1546)
__global__ void km_finish_kernel(const float* __restrict__ centers, const float* __restrict__ mean, int d, int k,
                                 int best, float* __restrict__ centers_out, const int* __restrict__ labels, int n,
                                 int* __restrict__ labels_out) {
  const int total = k * d;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x)
    centers_out[e] = __fadd_rn(centers[(size_t)best * total + e], mean[e % d]);
  if (labels_out)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      labels_out[i] = labels[(size_t)best * n + i];
}

static int km_check_ws(void* ws, size_t bytes, KmLayout* L) {
  if (ws == nullptr) return set_error(VIDSEG_E_INVALID, "%s", "null workspace");
  if (!km_lookup(ws, L)) return set_error(VIDSEG_E_INVALID, "%s", "workspace not prepared (call vidseg_kmeans_prepare)");
  if (bytes < L->total) return set_error(VIDSEG_E_WORKSPACE, "%s: need %lld bytes, got %lld", "k-means workspace", (long long)L->total, (long long)bytes);
  return 0;
}

static int km_launch_assign(const float* x, const KmLayout& L, int runs, int row_begin, int row_end,
                            const float* centers, const double* cnorm, int* labels, int labels_stride, int* changed,
                            const int* flags, int count_changes, int only_nonstrict, void* stream) {
  const int rows = row_end - row_begin;
  if (rows <= 0) return 0;
  dim3 grid((rows + kBM - 1) / kBM, runs);
  VS_LAUNCH(km_assign_kernel, grid, 256, 0, stream, x, L.n, L.d, L.k, row_begin, row_end, centers, cnorm, labels,
            labels_stride, changed, flags, count_changes, only_nonstrict);
  VS_POST_LAUNCH();
  return 0;
}


// defined in gemm_tc.cu: out[M,N] = acc_scale * A[M,K] . W[N,K]^T on the split operands, tagged with `family`
int gemm_split_run(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, float* out_f32, int m, int n,
                   int k, float acc_scale, int family, void* stream);

// E-step of every unfinished run over rows [row_begin, row_end) of the centred data held in the workspace
static int km_assign_runs(void* ws, const KmLayout& L, int row_begin, int row_end, int count_changes, int only_nonstrict,
                          void* stream) {
  const int rows = row_end - row_begin;
  if (rows <= 0) return 0;
  if (!L.use_tc)
    return km_launch_assign(at<float>(ws, L.xc), L, L.r, row_begin, row_end, at<float>(ws, L.centers),
                            at<double>(ws, L.cnorm), at<int>(ws, L.labels), L.n, at<int>(ws, L.changed),
                            at<int>(ws, L.flags), count_changes, only_nonstrict, stream);
  // error radius of the filter per unit |x||c|: the score is cn - 2 S; S carries the operand split (2^-21) and the
  // fp32 accumulation of d products inside the tensor core (<= d * 2^-22 with truncating alignment)
  // cs_hi / cs_lo: the centres as scaled fp16 pairs, kept current by vidseg_kmeans_seed and km_update_avg_kernel
  if (int e = gemm_split_run(at<__half>(ws, L.xs_hi) + (size_t)row_begin * L.d, at<__half>(ws, L.xs_lo) + (size_t)row_begin * L.d,
                             at<__half>(ws, L.cs_hi), at<__half>(ws, L.cs_lo),
                             at<float>(ws, L.sdot) + (size_t)row_begin * L.rk_pad, rows, L.rk_pad, L.d, 1.0f, kFamKMeans, stream))
    return e;
  // error radius of the filter per unit |x||c|: the score is cn - 2 S; S carries the operand split (2^-21) and the
  // fp32 accumulation of d products inside the tensor core (<= d * 2^-22 with truncating alignment)
  const double band = 2.0 * ((double)L.d * 0x1p-22 + 0x1p-20);
  const long long total = (long long)rows * L.r;
  int* amb_count = reinterpret_cast<int*>(at<unsigned>(ws, L.absmax) + 1);
  const size_t smem = (size_t)L.r * L.k * 8 + 64;   // amb_count {count, ticket} is cleared by prepare and by every resolve
  VS_REQUIRE(smem <= 48 * 1024, "n_init * k too large for the tensor-core E-step");
  VS_LAUNCH_PDL_L(1, km_assign_tc_kernel, (int)((total + 255) / 256), 256, smem, stream, L.k, L.r, row_begin, row_end,
            at<double>(ws, L.cnorm), at<double>(ws, L.xx), at<float>(ws, L.sdot), L.rk_pad, at<unsigned>(ws, L.absmax),
            at<int>(ws, L.labels), L.n, at<int>(ws, L.changed), at<int>(ws, L.flags), count_changes, only_nonstrict, band,
            amb_count, at<int4>(ws, L.amb_list));
  VS_POST_LAUNCH();
  VS_LAUNCH_PDL_L(2, km_assign_resolve_kernel, kNumSMs * 2, 256, 0, stream, at<float>(ws, L.xc), L.d, L.k, at<float>(ws, L.centers),
            at<double>(ws, L.cnorm), at<double>(ws, L.xx), L.k <= 64 ? (const float*)nullptr : at<float>(ws, L.sdot), L.rk_pad,
            at<unsigned>(ws, L.absmax), at<int>(ws, L.labels), L.n, at<int>(ws, L.changed), count_changes, band, amb_count,
            at<int>(ws, L.amb_list), 4);
  VS_POST_LAUNCH();
  return 0;
}


// ------------------------------------------------------------------------------------------
// match_gt_mask mode: majority map and brute-force 4-NN label propagation
//   scripts/sampling/feature_extraction.py:589-594 (every K-means label takes the most frequent ground-truth label
//   of its cells) and :606-612 (KNeighborsClassifier(n_neighbors=4).fit(ref).predict(all tokens)).
// The k-NN search is the K-means E-step design again: one pair16 tcgen05 GEMM gives the approximate scores
// ||r||^2 - 2 q.r of a chunk of queries against every reference point; per query one warp finds the k-th smallest
// score, collects every reference whose score lies inside the filter's error band of it, re-evaluates those few in
// float64 and keeps the k nearest by (distance, index) -- what sklearn's float64 brute force returns.
// ------------------------------------------------------------------------------------------
constexpr int kGtRange = 1024;   // ground-truth labels must lie in [0, kGtRange)
__global__ void __launch_bounds__(1024)
mg_majority_kernel(const int* __restrict__ fake, const int* __restrict__ gt, int n, int num_fake, int* __restrict__ ref,
                   int* __restrict__ err) {
  __shared__ int hist[kGtRange];
  __shared__ int s_best;
  for (int f = 0; f < num_fake; ++f) {
    for (int i = threadIdx.x; i < kGtRange; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (fake[i] == f) {
        const int g = gt[i];
        if (g < 0 || g >= kGtRange) atomicExch(err, 1); else atomicAdd(&hist[g], 1);
      }
    __syncthreads();
    if (threadIdx.x == 0) {
      int best = -1, bc = 0;
      for (int g = 0; g < kGtRange; ++g) if (hist[g] > bc) { bc = hist[g]; best = g; }   // first maximum: smallest label
      s_best = best;
    }
    __syncthreads();
    if (s_best >= 0)
      for (int i = threadIdx.x; i < n; i += blockDim.x) if (fake[i] == f) ref[i] = s_best;
    __syncthreads();
  }
}

// float64 squared norms of fp32 rows + running |max| (operand scale); one warp per row
__global__ void __launch_bounds__(256)
knn_norm_kernel(const float* __restrict__ x, int n, int d, double* __restrict__ nrm, unsigned* __restrict__ absmax) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* xr = x + (size_t)row * d;
  double s = 0.0;
  float m = 0.f;
  for (int c = lane; c < d; c += 32) { const float v = xr[c]; s = fma((double)v, (double)v, s); m = fmaxf(m, fabsf(v)); }
  s = warp_sum(s);
  m = warp_max(m);
  if (lane == 0) { nrm[row] = s; atomicMax(absmax, __float_as_uint(m)); }
}

constexpr int kKnnMaxK = 8;
constexpr int kKnnMaxCand = 64;
constexpr int kKnnWarps = 8;
__global__ void __launch_bounds__(kKnnWarps * 32)
knn_select_kernel(const float* __restrict__ q, const float* __restrict__ ref, const int* __restrict__ ref_labels, int nq,
                  int nr, int d, int k, const float* __restrict__ sdot, int ld, const double* __restrict__ rnorm,
                  const double* __restrict__ qnorm, const unsigned* __restrict__ absmax, double band, double rn_max_sqrt,
                  int* __restrict__ out, int* __restrict__ err) {
  __shared__ int s_cand[kKnnWarps][kKnnMaxCand];
  __shared__ double s_dist[kKnnWarps][kKnnMaxCand];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * kKnnWarps + warp;
  if (row >= nq) return;
  const float sc_op = km_operand_scale(absmax[0]);
  const double inv = 1.0 / ((double)sc_op * (double)sc_op);
  const float* sr = sdot + (size_t)row * ld;
  // pass 1: the k smallest approximate scores (lane-local sorted lists, then k rounds of warp minimum)
  double loc[kKnnMaxK];
#pragma unroll
  for (int t = 0; t < kKnnMaxK; ++t) loc[t] = 1e300;
  for (int j = lane; j < nr; j += 32) {
    double v = rnorm[j] - 2.0 * ((double)sr[j] * inv);
#pragma unroll
    for (int t = 0; t < kKnnMaxK; ++t)
      if (t < k && v < loc[t]) { const double tmp = loc[t]; loc[t] = v; v = tmp; }
  }
  double vk = 1e300;
  for (int round = 0; round < k; ++round) {
    double m = loc[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    vk = m;
    // the lane holding it pops its head (lowest lane on ties)
    const unsigned who = __ballot_sync(0xffffffffu, loc[0] == m);
    if (lane == __ffs(who) - 1) {
#pragma unroll
      for (int t = 0; t + 1 < kKnnMaxK; ++t) loc[t] = loc[t + 1];
      loc[kKnnMaxK - 1] = 1e300;
    }
  }
  // pass 2: every reference inside the error band of the k-th score (ascending index order)
  const double thresh = vk + 2.0 * band * sqrt(qnorm[row]) * rn_max_sqrt;
  int count = 0;
  for (int j0 = 0; j0 < nr; j0 += 32) {
    const int j = j0 + lane;
    const bool in = (j < nr) && (rnorm[j] - 2.0 * ((double)sr[j] * inv) <= thresh);
    const unsigned mask = __ballot_sync(0xffffffffu, in);
    if (in) {
      const int pos = count + __popc(mask & ((1u << lane) - 1));
      if (pos < kKnnMaxCand) s_cand[warp][pos] = j;
    }
    count += __popc(mask);
  }
  if (count > kKnnMaxCand) { if (lane == 0) atomicExch(err, 2); count = kKnnMaxCand; }
  __syncwarp();
  // exact float64 distances of the candidates (minus the constant ||q||^2), warp-cooperative dot products
  const float* qr = q + (size_t)row * d;
  for (int e = 0; e < count; ++e) {
    const int j = s_cand[warp][e];
    const float* rr = ref + (size_t)j * d;
    double acc = 0.0;
    for (int c0 = lane; c0 < d; c0 += 8 * 32) {
      float xa[8], ca[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + 32 * u;
        xa[u] = (c < d) ? qr[c] : 0.f;
        ca[u] = (c < d) ? rr[c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fma((double)xa[u], (double)ca[u], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_dist[warp][e] = rnorm[j] - 2.0 * acc;
  }
  __syncwarp();
  if (lane == 0) {
    // k nearest by (distance, index); candidates are in ascending index order, so a strict '<' keeps the lower index
    int lab[kKnnMaxK];
    for (int t = 0; t < k; ++t) {
      int bi = -1;
      double bd = 1e300;
      for (int e = 0; e < count; ++e)
        if (s_cand[warp][e] >= 0 && s_dist[warp][e] < bd) { bd = s_dist[warp][e]; bi = e; }
      lab[t] = (bi >= 0) ? ref_labels[s_cand[warp][bi]] : 0x7fffffff;
      if (bi >= 0) s_cand[warp][bi] = -1;
    }
    // scipy.stats.mode: most frequent label, smallest on ties
    int best = 0x7fffffff, bc = 0;
    for (int t = 0; t < k; ++t) {
      int c = 0;
      for (int u = 0; u < k; ++u) c += (lab[u] == lab[t]);
      if (c > bc || (c == bc && lab[t] < best)) { bc = c; best = lab[t]; }
    }
    out[row] = best;
  }
}

__global__ void knn_max_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, v[i]);
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 16)); m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 8));
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 4)); m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
  __shared__ double sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, sm[w]);
    out[0] = t;
  }
}

constexpr int kKnnChunk = 4096;   // queries per filter GEMM
struct KnnLayout { size_t rn, qn, absmax, err, rmax, rs_hi, rs_lo, qs_hi, qs_lo, sdot, total; int nr_pad; };
static KnnLayout knn_layout(int nr, int nq, int d) {
  KnnLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.nr_pad = (nr + 7) / 8 * 8;
  const int qc = nq < kKnnChunk ? nq : kKnnChunk;
  L.rn = take((size_t)nr * 8);
  L.qn = take((size_t)nq * 8);
  L.absmax = take(16);
  L.err = take(16);
  L.rmax = take(16);
  L.rs_hi = take((size_t)L.nr_pad * d * 2);
  L.rs_lo = take((size_t)L.nr_pad * d * 2);
  L.qs_hi = take((size_t)qc * d * 2);
  L.qs_lo = take((size_t)qc * d * 2);
  L.sdot = take((size_t)qc * L.nr_pad * 4);
  L.total = off;
  return L;
}
}  // namespace vidseg

using namespace vidseg;

VS_API int vidseg_set_kmeans_mstep(int mode) {
  VS_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (float64 sums) or 1 (int8 tensor-core sums)");
  g_km_mstep.store(mode, std::memory_order_relaxed);
  return 0;
}
VS_API int vidseg_get_kmeans_mstep(void) { return km_mstep_mode(); }

VS_API size_t vidseg_kmeans_workspace_bytes(int n, int d, int k, int n_init, int n_trials) {
  if (n <= 0 || d <= 0 || k <= 0 || n_init <= 0 || n_trials <= 0 || n_trials > kMaxTrials) return 0;
  return km_layout(n, d, k, n_init, n_trials).total;
}

VS_API int vidseg_kmeans_prepare(const float* x, int n, int d, int k, int n_init, int n_trials, float tol_rel,
                                     int max_iter, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(x != nullptr && workspace != nullptr, "null pointer");
  VS_REQUIRE(n >= 1 && d >= 1 && k >= 1 && k <= n, "need 1 <= k <= n, d >= 1");
  VS_REQUIRE(n_init >= 1 && n_init <= 64, "n_init must be 1..64");
  VS_REQUIRE(n_trials >= 1 && n_trials <= kMaxTrials, "n_trials must be 1..8");
  KmLayout L = km_layout(n, d, k, n_init, n_trials);
  L.tol_rel = tol_rel;
  L.max_iter = max_iter;
  if (workspace_bytes < L.total)
    return set_error(VIDSEG_E_WORKSPACE, "%s: need %lld bytes, got %lld", "k-means workspace", (long long)L.total, (long long)workspace_bytes);
  void* ws = workspace;
  if (L.use_mq) {
    // int8 tiles described as 16-bit elements (two per element): 64 per 128-byte swizzle row
    const uint64_t d_oh[2] = {(uint64_t)L.n_pad / 2, (uint64_t)n_init * k}, s_oh[1] = {(uint64_t)L.n_pad};
    const uint32_t b_oh[2] = {kMqBK / 2, 128};
    if (int e = encode_tmap_16bit(&L.tm_onehot, at<int8_t>(ws, L.mq_onehot), 2, d_oh, s_oh, b_oh)) return e;
    const uint64_t d_pl[3] = {(uint64_t)L.n_pad / 2, (uint64_t)d, (uint64_t)kMqPlanes};
    const uint64_t s_pl[2] = {(uint64_t)L.n_pad, (uint64_t)d * L.n_pad};
    const uint32_t b_pl[3] = {kMqBK / 2, 128, 1};
    if (int e = encode_tmap_16bit(&L.tm_planes, at<int8_t>(ws, L.mq_planes), 3, d_pl, s_pl, b_pl)) return e;
  }
  {
    std::lock_guard<std::mutex> lk(g_km_mu);
    g_km_registry[workspace] = L;
  }
  if (d % 4 == 0 && ((uintptr_t)x % 16) == 0) {
    constexpr int kCsSmem = kCsStages * kCsRows * kCsCols * 4;
    static cudaError_t attr_cs = cudaFuncSetAttribute(km_colstats_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCsSmem);
    VS_CHECK_CUDA(attr_cs);
    VS_LAUNCH(km_colstats_ring_kernel, (d + kCsCols - 1) / kCsCols, kCsThreads, kCsSmem, stream, x, n, d, at<float>(ws, L.mean), at<float>(ws, L.var));
  } else {
    VS_LAUNCH(km_colstats_kernel, (d + 31) / 32, 32, 0, stream, x, n, d, at<float>(ws, L.mean), at<float>(ws, L.var));
  }
  VS_POST_LAUNCH();
  VS_LAUNCH(km_tol_reset_kernel, 1, 256, 0, stream, at<float>(ws, L.var), d, tol_rel, at<float>(ws, L.tol),
            at<unsigned>(ws, L.absmax), at<int>(ws, L.flags), at<int>(ws, L.changed), n_init);
  VS_POST_LAUNCH();
  VS_LAUNCH(km_center_kernel, (n * 32 + 255) / 256, 256, 0, stream, x, at<float>(ws, L.mean), n, d, at<float>(ws, L.xc),
            at<double>(ws, L.xx), at<unsigned>(ws, L.absmax));
  VS_POST_LAUNCH();
  VS_CHECK_CUDA(cudaMemsetAsync(at<int>(ws, L.labels), 0xFF, (size_t)n_init * n * 4, (cudaStream_t)stream));
  VS_CHECK_CUDA(cudaMemsetAsync(at<int>(ws, L.upd_ticket), 0, (size_t)n_init * 4, (cudaStream_t)stream));
  VS_CHECK_CUDA(cudaMemsetAsync(at<int>(ws, L.reloc), 0, (size_t)n_init * 2 * 4, (cudaStream_t)stream));
  if (L.use_mq) {
    VS_LAUNCH(km_quantize_kernel, dim3(L.n_pad / kMqBK, (d + 31) / 32), 256, 0, stream, at<float>(ws, L.xc), n, L.n_pad, d,
              at<unsigned>(ws, L.absmax), at<int8_t>(ws, L.mq_planes));
    VS_POST_LAUNCH();
  }
  if (L.use_tc) {
    const size_t tot = (size_t)n * d;
    VS_LAUNCH(km_split_scaled_kernel, (int)std::min<size_t>((tot + 255) / 256, (size_t)kNumSMs * 16), 256, 0, stream,
              at<float>(ws, L.xc), tot, tot, at<unsigned>(ws, L.absmax), at<__half>(ws, L.xs_hi), at<__half>(ws, L.xs_lo));
    VS_POST_LAUNCH();
  }
  return 0;
}

VS_API int vidseg_kmeans_seed(const int32_t* first_idx, const double* rand, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(first_idx != nullptr && (rand != nullptr || L.k == 1), "null pointer");
  void* ws = workspace;
  cudaStream_t st = (cudaStream_t)stream;
  // cand[r][0] = first_idx[r]
  VS_CHECK_CUDA(cudaMemcpy2DAsync(at<int>(ws, L.cand), kMaxTrials * 4, first_idx, 4, 4, L.r, cudaMemcpyDeviceToDevice, st));
  const size_t smem = (size_t)L.t * L.d * 8;
  VS_REQUIRE(smem <= 200 * 1024, "n_trials * d too large for the seeding kernel");
  dim3 grid(kPotBlocks, L.r);
  for (int c = 0; c < L.k; ++c) {
    const int tc = (c == 0) ? 1 : L.t;
    // all runs from one pass over X when their candidates fit in shared memory
    const size_t smem_all = ((size_t)L.r * tc * L.d + (size_t)L.r * tc + (size_t)kPotWarps * L.r * (kPotRows * tc <= 16 ? 16 : 32)) * 8;
    // FP64 tensor-core form: all runs' candidates as tiles of eight, rows of X in chunks of 16 columns
    static const int kpp_mma = [] { const char* e = getenv("VIDSEG_KPP_MMA"); return e ? atoi(e) : 1; }();
    const int nt = (L.r * tc + 7) / 8;
    const size_t smem_mma = ((size_t)nt * 8 * (L.d + 2) + (size_t)nt * 8) * 8;
    KppDistMmaFn fm = (kpp_mma && L.d % 16 == 0 && nt <= 5 && smem_mma <= 220 * 1024) ? kpp_dist_mma_fn(nt) : nullptr;
    if (fm) {
      VS_CHECK_CUDA(cudaFuncSetAttribute(fm, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      VS_LAUNCH(fm, kPotBlocks, kPotWarps * 32, smem_mma, st, at<float>(ws, L.xc), at<double>(ws, L.xx), L.n, L.d, L.r, tc,
                at<int>(ws, L.cand), at<float>(ws, L.closest), c > 0 ? 1 : 0, at<float>(ws, L.newdist),
                at<double>(ws, L.potpart), L.t);
      VS_POST_LAUNCH();
      VS_LAUNCH(km_kpp_select_scan_kernel, L.r, 1024, 0, st, at<float>(ws, L.xc), L.n, L.d, L.k, L.t, L.t, c, tc,
                at<double>(ws, L.potpart), at<float>(ws, L.newdist), at<float>(ws, L.closest), at<int>(ws, L.cand),
                at<float>(ws, L.pot), at<float>(ws, L.centers), at<int>(ws, L.center_idx), rand);
      VS_POST_LAUNCH();
      continue;
    }
    KppDistAllFn fa = (smem_all <= 220 * 1024) ? kpp_dist_all_fn(tc, L.d) : nullptr;
    if (fa) {
      VS_CHECK_CUDA(cudaFuncSetAttribute(fa, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      VS_LAUNCH(fa, kPotBlocks, kPotWarps * 32, smem_all, st, at<float>(ws, L.xc), at<double>(ws, L.xx), L.n, L.d, L.r,
                at<int>(ws, L.cand), at<float>(ws, L.closest), c > 0 ? 1 : 0, at<float>(ws, L.newdist),
                at<double>(ws, L.potpart), L.t);
      VS_POST_LAUNCH();
      VS_LAUNCH(km_kpp_select_scan_kernel, L.r, 1024, 0, st, at<float>(ws, L.xc), L.n, L.d, L.k, L.t, L.t, c, tc,
                at<double>(ws, L.potpart), at<float>(ws, L.newdist), at<float>(ws, L.closest), at<int>(ws, L.cand),
                at<float>(ws, L.pot), at<float>(ws, L.centers), at<int>(ws, L.center_idx), rand);
      VS_POST_LAUNCH();
      continue;
    }
    KppDistFn fn = kpp_dist_fn(tc);
    if (smem > 48 * 1024) VS_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH(fn, grid, kPotWarps * 32, (size_t)tc * L.d * 8, st, at<float>(ws, L.xc), at<double>(ws, L.xx), L.n, L.d,
              at<int>(ws, L.cand), at<float>(ws, L.closest), c > 0 ? 1 : 0, at<float>(ws, L.newdist),
              at<double>(ws, L.potpart), L.t);
    VS_POST_LAUNCH();
    VS_LAUNCH(km_kpp_select_scan_kernel, L.r, 1024, 0, st, at<float>(ws, L.xc), L.n, L.d, L.k, L.t, L.t, c, tc,
              at<double>(ws, L.potpart), at<float>(ws, L.newdist), at<float>(ws, L.closest), at<int>(ws, L.cand),
              at<float>(ws, L.pot), at<float>(ws, L.centers), at<int>(ws, L.center_idx), rand);
    VS_POST_LAUNCH();
  }
  VS_LAUNCH(km_cnorm_kernel, (L.r * L.k * 32 + 255) / 256, 256, 0, st, at<float>(ws, L.centers), L.d, L.r * L.k,
            at<double>(ws, L.cnorm));
  VS_POST_LAUNCH();
  if (L.use_tc) {
    const size_t c_valid = (size_t)L.r * L.k * L.d, c_total = (size_t)L.rk_pad * L.d;   // padding rows are zero
    VS_LAUNCH(km_split_scaled_kernel, (int)((c_total + 255) / 256), 256, 0, st, at<float>(ws, L.centers), c_valid, c_total,
              at<unsigned>(ws, L.absmax), at<__half>(ws, L.cs_hi), at<__half>(ws, L.cs_lo));
    VS_POST_LAUNCH();
  }
  return 0;
}

VS_API int vidseg_kmeans_assign(void* workspace, size_t workspace_bytes, int row_begin, int row_end, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= L.n, "bad row range");
  void* ws = workspace;
  return km_assign_runs(ws, L, row_begin, row_end, 1, 0, stream);
}

static bool km_mq_range_ok(const KmLayout& L, int row_begin, int row_end) {
  // the int8 product splits its row range over kMqSplit CTAs per tile: ranges shorter than that many 128-row blocks
  // (tiny shards) take the float64 kernels, which produce the same sums
  return L.use_mq && ((row_end + kMqBK - 1) / kMqBK - row_begin / kMqBK >= kMqSplit || row_end == row_begin);
}

// M-step part 1.  words_mode: 0 = `partial` float64 [R,K,D+1] + `changed` int32 [R] (the original pair of buffers);
// 1 = ONE exchange array of float64 words [R*K*(D+1) + R]; 2 = the same array as int64 words (integer sums, exact).
// labels -> one-hot rows + count partials -> digit-plane x one-hot tensor-core product (partial sums in the workspace)
static int km_mstep_product(void* ws, const KmLayout& L, int row_begin, int row_end, void* stream) {
  {
    VS_LAUNCH_PDL_L(3, km_onehot_kernel, dim3(L.mq_blocks, L.r), kOhThreads, (size_t)L.k * 4, stream, at<int>(ws, L.labels), L.n, L.n_pad,
              L.k, row_begin, row_end, at<int>(ws, L.flags), at<int8_t>(ws, L.mq_onehot), at<int>(ws, L.mq_cnt));
    VS_POST_LAUNCH();
    MqParams mp{};
    mp.kb_lo = (row_end > row_begin) ? row_begin / kMqBK : 0;
    mp.kb_hi = (row_end > row_begin) ? (row_end + kMqBK - 1) / kMqBK : 0;   // empty range: zero sums (memset below)
    mp.kb_per_split = (mp.kb_hi - mp.kb_lo + kMqSplit - 1) / kMqSplit;
    mp.rk = L.r * L.k;
    mp.d = L.d;
    mp.k = L.k;
    mp.runs = L.r;
    mp.m_tiles = (mp.rk + 127) / 128;
    mp.n_tiles = (L.d + 127) / 128;
    mp.out = at<int>(ws, L.mq_part);
    mp.flags = at<int>(ws, L.flags);
    constexpr size_t kMqSmem = (size_t)kMqStages * 2 * 128 * kMqBK + 1024 + 256;
    static cudaError_t attr_mq = cudaFuncSetAttribute(km_mstep_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMqSmem);
    VS_CHECK_CUDA(attr_mq);
    if (mp.kb_hi > mp.kb_lo) {
      const double macs = (double)mp.m_tiles * 128 * mp.n_tiles * 128 * kMqPlanes * (double)(mp.kb_hi - mp.kb_lo) * kMqBK;
      (void)macs;
      VS_LAUNCH_PDL_L(4, km_mstep_mma_kernel, mp.m_tiles * mp.n_tiles * kMqPlanes * kMqSplit, kMqThreads, kMqSmem, stream,
                    L.tm_onehot, L.tm_planes, mp);
      VS_POST_LAUNCH();
    } else {
      VS_CHECK_CUDA(cudaMemsetAsync(mp.out, 0, (size_t)kMqSplit * kMqPlanes * mp.rk * L.d * 4, (cudaStream_t)stream));
    }
  }
  return 0;
}

static int km_partial_impl(void* ws, const KmLayout& L, int row_begin, int row_end, void* partial, int32_t* changed,
                           int words_mode, void* stream) {
  void* tail = words_mode ? static_cast<void*>(reinterpret_cast<char*>(partial) + (size_t)L.r * L.k * (L.d + 1) * 8) : nullptr;
  if (km_mq_range_ok(L, row_begin, row_end)) {
    int* cnt_direct = nullptr;
    if (int e = km_mstep_product(ws, L, row_begin, row_end, stream)) return e;
    const int rk = L.r * L.k;
    if (words_mode == 2)
      VS_LAUNCH(km_mstep_combine_kernel<true>, dim3(L.k, L.r), 256, 0, stream, at<int>(ws, L.mq_part), at<int>(ws, L.mq_cnt),
                L.mq_blocks, L.d, L.k, rk, at<int>(ws, L.flags), at<unsigned>(ws, L.absmax), partial, at<int>(ws, L.changed),
                changed, tail, cnt_direct);
    else
      VS_LAUNCH(km_mstep_combine_kernel<false>, dim3(L.k, L.r), 256, 0, stream, at<int>(ws, L.mq_part), at<int>(ws, L.mq_cnt),
                L.mq_blocks, L.d, L.k, rk, at<int>(ws, L.flags), at<unsigned>(ws, L.absmax), partial, at<int>(ws, L.changed),
                changed, tail, cnt_direct);
    VS_POST_LAUNCH();
    return 0;
  }
  VS_REQUIRE(words_mode != 2, "integer exchange words need the int8 tensor-core M-step on this row range");
  const size_t smem1 = (size_t)L.k * kColTile * 8;
  VS_REQUIRE(smem1 <= 200 * 1024, "k too large for the M-step kernel");
  // (two runs per pass over X were measured slower, 83 vs 72 us: the shared-memory adds bound the kernel, not the reads)
  static cudaError_t attr1 = cudaFuncSetAttribute(km_partial_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  VS_CHECK_CUDA(attr1);
  dim3 grid(kSlabs, (L.d + kColTile - 1) / kColTile, L.r);
  VS_LAUNCH(km_partial_kernel<1>, grid, kColTile, smem1, stream, at<float>(ws, L.xc), L.n, L.d, L.k, L.r, row_begin, row_end,
            at<int>(ws, L.labels), at<int>(ws, L.flags), at<double>(ws, L.part), at<int>(ws, L.partcnt));
  VS_POST_LAUNCH();
  VS_LAUNCH(km_reduce_kernel, dim3(L.k, L.r), 256, 0, stream, at<double>(ws, L.part), at<int>(ws, L.partcnt), L.d, L.k,
            at<int>(ws, L.flags), reinterpret_cast<double*>(partial), at<int>(ws, L.changed), changed,
            reinterpret_cast<double*>(tail));
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_kmeans_partial(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                                     double* partial, int32_t* changed, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= L.n, "bad row range");
  void* ws = workspace;
  if (partial == nullptr) partial = at<double>(ws, L.partial);
  if (changed == nullptr) changed = at<int>(ws, L.changed);
  return km_partial_impl(ws, L, row_begin, row_end, partial, changed, 0, stream);
}

// M-step part 2 (one launch).  src 0: `sums` = float64 [R,K,D+1] (relocation happens in place); src 1: the all-reduced
// integer exchange words; src 2: the tensor-core product's own partial sums in the workspace (`sums` ignored).
// `tail`: change counters stored behind the sums (exchange words), else `changed`.
static int km_update_fused(void* ws, const KmLayout& L, int src, void* sums, const void* tail, const int32_t* changed,
                           int local_rows_only, void* stream) {
  UpdArgs a{};
  a.d = L.d; a.k = L.k; a.runs = L.r; a.n = L.n; a.max_iter = L.max_iter;
  a.can_relocate = local_rows_only ? 0 : 1;
  a.cnt_blocks = L.mq_blocks; a.rk = L.r * L.k;
  a.sums = (src == 2) ? static_cast<const void*>(at<int>(ws, L.mq_part)) : sums;
  a.cntpart = at<int>(ws, L.mq_cnt);
  a.tail = tail;
  a.changed_in = changed ? changed : at<int>(ws, L.changed);
  a.changed_ws = at<int>(ws, L.changed);
  a.scratch = (src == 0) ? reinterpret_cast<double*>(sums) : at<double>(ws, L.partial);
  a.centers = at<float>(ws, L.centers); a.cnorm = at<double>(ws, L.cnorm); a.flags = at<int>(ws, L.flags);
  a.tol = at<float>(ws, L.tol); a.shiftsq = at<float>(ws, L.upd_shift); a.ticket = at<int>(ws, L.upd_ticket);
  a.absmax = at<unsigned>(ws, L.absmax);
  a.cs_hi = L.use_tc ? at<__half>(ws, L.cs_hi) : nullptr;
  a.cs_lo = L.use_tc ? at<__half>(ws, L.cs_lo) : nullptr;
  a.x = at<float>(ws, L.xc); a.labels = at<int>(ws, L.labels); a.dist = at<float>(ws, L.newdist);
  a.reloc_ticket = at<int>(ws, L.reloc); a.reloc_done = at<int>(ws, L.reloc) + L.r;
  const int grid = L.r * L.k;   // all blocks are resident (the relocation path waits across the blocks of a run)
  VS_REQUIRE(grid <= kNumSMs * 16, "n_init * k too large for the fused update kernel");
  if (src == 0) VS_LAUNCH_PDL_L(5, km_update_fused_kernel<0>, grid, kUpdThreads, 0, stream, a);
  else if (src == 1) VS_LAUNCH_PDL_L(5, km_update_fused_kernel<1>, grid, kUpdThreads, 0, stream, a);
  else VS_LAUNCH_PDL_L(5, km_update_fused_kernel<2>, grid, kUpdThreads, 0, stream, a);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_kmeans_update(void* workspace, size_t workspace_bytes, double* partial, const int32_t* changed,
                                    int local_rows_only, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  void* ws = workspace;
  if (partial == nullptr) partial = at<double>(ws, L.partial);
  return km_update_fused(ws, L, 0, partial, nullptr, changed, local_rows_only, stream);
}

// ---- multi-GPU exchange form: ONE array of 8-byte words per Lloyd iteration (sums | counts | change counters) ----
VS_API size_t vidseg_kmeans_exchange_words(void* workspace, size_t workspace_bytes) {
  KmLayout L;
  if (km_check_ws(workspace, workspace_bytes, &L)) return 0;
  return (size_t)L.r * L.k * (L.d + 1) + (size_t)L.r;
}

VS_API int vidseg_kmeans_exchange_mode(void* workspace, size_t workspace_bytes, int row_begin, int row_end) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e < 0 ? e : -e;
  if (!(0 <= row_begin && row_begin <= row_end && row_end <= L.n)) return VIDSEG_E_INVALID;
  return km_mq_range_ok(L, row_begin, row_end) ? 1 : 0;
}

VS_API int vidseg_kmeans_partial_words(void* workspace, size_t workspace_bytes, int row_begin, int row_end, int words_i64,
                                       void* words, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= L.n && words != nullptr, "bad row range / null words");
  return km_partial_impl(workspace, L, row_begin, row_end, words, nullptr, words_i64 ? 2 : 1, stream);
}

VS_API int vidseg_kmeans_update_words(void* workspace, size_t workspace_bytes, int words_i64, void* words,
                                      int local_rows_only, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(words != nullptr, "null words");
  VS_REQUIRE(!words_i64 || local_rows_only, "integer exchange words carry no relocation: local_rows_only must be set");
  void* ws = workspace;
  // the all-reduced change counters sit behind the sums: the averaging kernel tests them in place
  const void* tail = reinterpret_cast<const char*>(words) + (size_t)L.r * L.k * (L.d + 1) * 8;
  return km_update_fused(ws, L, words_i64 ? 1 : 0, words, tail, nullptr, local_rows_only, stream);
}

// the convergence flags [R][4] = {done, strict, n_iter, empty cluster seen} copied to (pinned) host memory WITHOUT
// synchronising: the caller records an event behind it and reads the flags when that event has passed, so polling
// never stalls the launch queue
VS_API int vidseg_kmeans_flags_async(void* workspace, size_t workspace_bytes, int32_t* flags_host, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(flags_host != nullptr, "null pointer");
  VS_CHECK_CUDA(cudaMemcpyAsync(flags_host, at<int>(workspace, L.flags), (size_t)L.r * 16, cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
  return 0;
}

VS_API int vidseg_kmeans_active_runs(void* workspace, size_t workspace_bytes, int* n_active_host, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  std::vector<int> flags((size_t)L.r * 4);
  VS_CHECK_CUDA(cudaMemcpyAsync(flags.data(), at<int>(workspace, L.flags), flags.size() * 4, cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
  VS_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  int active = 0;
  for (int r = 0; r < L.r; ++r) active += flags[(size_t)r * 4] == 0;
  *n_active_host = active;
  return 0;
}

VS_API int vidseg_kmeans_status(void* workspace, size_t workspace_bytes, int* n_active_host, int* empty_seen_host,
                                void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  std::vector<int> flags((size_t)L.r * 4);
  VS_CHECK_CUDA(cudaMemcpyAsync(flags.data(), at<int>(workspace, L.flags), flags.size() * 4, cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
  VS_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  int active = 0, empty = 0;
  for (int r = 0; r < L.r; ++r) { active += flags[(size_t)r * 4] == 0; empty += flags[(size_t)r * 4 + 3] != 0; }
  if (n_active_host) *n_active_host = active;
  if (empty_seen_host) *empty_seen_host = empty;
  return 0;
}

VS_API int vidseg_kmeans_inertia(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                                     double* inertia_partial, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= L.n, "bad row range");
  void* ws = workspace;
  if (inertia_partial == nullptr) inertia_partial = at<double>(ws, L.inertia);
  // rerun the E-step of runs that stopped on tolerance / max_iter (sklearn/_kmeans.py:741-753)
  if (int e = km_assign_runs(ws, L, row_begin, row_end, 0, 1, stream)) return e;
  VS_LAUNCH(km_inertia_kernel, dim3(kInertiaBlocks, L.r), 256, 0, stream, at<float>(ws, L.xc), L.n, L.d, L.k, row_begin,
            row_end, at<float>(ws, L.centers), at<int>(ws, L.labels), at<double>(ws, L.inertia_part));
  VS_POST_LAUNCH();
  VS_LAUNCH(km_inertia_reduce_kernel, L.r, 32, 0, stream, at<double>(ws, L.inertia_part), inertia_partial);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_kmeans_same_matrix(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                                         int32_t* same, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= L.n, "bad row range");
  if (same == nullptr) same = at<int>(workspace, L.same);
  VS_LAUNCH(km_same_kernel, dim3(L.r, L.r), 256, (size_t)L.k * 4, stream, at<int>(workspace, L.labels), L.n, L.k,
            row_begin, row_end, same);
  VS_POST_LAUNCH();
  return 0;
}

// host-side best-of-R rule (sklearn/_kmeans.py:1529-1541)
VS_API int vidseg_kmeans_pick_best_host(const float* inertia_host, const int32_t* same_host, int n_init) {
  int best = 0;
  for (int i = 1; i < n_init; ++i)
    if (inertia_host[i] < inertia_host[best] && !same_host[i * n_init + best]) best = i;
  return best;
}

VS_API int vidseg_kmeans_finish(void* workspace, size_t workspace_bytes, int best, float* centers_out,
                                    int32_t* labels_fit_out, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(best >= 0 && best < L.r && centers_out != nullptr, "bad best run / null output");
  void* ws = workspace;
  VS_LAUNCH(km_finish_kernel, 64, 256, 0, stream, at<float>(ws, L.centers), at<float>(ws, L.mean), L.d, L.k, best,
            centers_out, at<int>(ws, L.labels), L.n, labels_fit_out);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_kmeans_select(void* workspace, size_t workspace_bytes, const double* inertia, float* centers_out,
                                    int32_t* labels_fit_out, int32_t* info_host, float* inertia_host, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  void* ws = workspace;
  cudaStream_t st = (cudaStream_t)stream;
  if (inertia == nullptr) inertia = at<double>(ws, L.inertia);
  if (int e = vidseg_kmeans_same_matrix(workspace, workspace_bytes, 0, L.n, nullptr, stream)) return e;
  std::vector<double> in64(L.r);
  std::vector<int> same((size_t)L.r * L.r), flags((size_t)L.r * 4);
  VS_CHECK_CUDA(cudaMemcpyAsync(in64.data(), inertia, (size_t)L.r * 8, cudaMemcpyDeviceToHost, st));
  VS_CHECK_CUDA(cudaMemcpyAsync(same.data(), at<int>(ws, L.same), same.size() * 4, cudaMemcpyDeviceToHost, st));
  VS_CHECK_CUDA(cudaMemcpyAsync(flags.data(), at<int>(ws, L.flags), flags.size() * 4, cudaMemcpyDeviceToHost, st));
  VS_CHECK_CUDA(cudaStreamSynchronize(st));
  std::vector<float> in32(L.r);
  for (int r = 0; r < L.r; ++r) in32[r] = (float)in64[r];
  const int best = vidseg_kmeans_pick_best_host(in32.data(), same.data(), L.r);
  if (inertia_host) for (int r = 0; r < L.r; ++r) inertia_host[r] = in32[r];
  if (info_host) {
    info_host[0] = best;
    info_host[1] = flags[(size_t)best * 4 + 2];
    int tot = 0;
    for (int r = 0; r < L.r; ++r) tot = tot > flags[(size_t)r * 4 + 2] ? tot : flags[(size_t)r * 4 + 2];
    info_host[2] = tot;
    info_host[3] = 0;
  }
  return vidseg_kmeans_finish(workspace, workspace_bytes, best, centers_out, labels_fit_out, stream);
}

VS_API int vidseg_kmeans_predict(const float* x, int n, int d, const float* centers, int k, int32_t* labels_out,
                                     double* cnorm_scratch, void* stream) {
  VS_REQUIRE(x != nullptr && centers != nullptr && labels_out != nullptr && cnorm_scratch != nullptr, "null pointer");
  VS_REQUIRE(n >= 0 && d >= 1 && k >= 1, "bad shape");
  if (n == 0) return 0;
  KmLayout L{};
  L.n = n; L.d = d; L.k = k; L.r = 1;
  VS_LAUNCH(km_cnorm_kernel, (k * 32 + 255) / 256, 256, 0, stream, centers, d, k, cnorm_scratch);
  VS_POST_LAUNCH();
  return km_launch_assign(x, L, 1, 0, n, centers, cnorm_scratch, labels_out, n, nullptr, nullptr, 0, 0, stream);
}

// `iterations` Lloyd iterations of every unfinished run over ALL rows, queued without synchronising (finished runs make
// every kernel return early): E-step (tensor-core filter + resolver), one-hot labels -> exact int8 tensor-core sums ->
// averaging / relocation / convergence flags straight from the product's partial sums.  The loop body of
// vidseg_kmeans_fit_predict; callers that poll the flags themselves (vidseg_kmeans_flags_async) drive it directly -- the
// run-sharded multi-GPU fit, where every rank iterates its own subset of the n_init runs.
VS_API int vidseg_kmeans_lloyd(void* workspace, size_t workspace_bytes, int iterations, void* stream) {
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  VS_REQUIRE(iterations >= 0, "negative iteration count");
  void* ws = workspace;
  for (int b = 0; b < iterations; ++b) {
    if (int e = km_assign_runs(ws, L, 0, L.n, 1, 0, stream)) return e;
    if (km_mq_range_ok(L, 0, L.n)) {
      if (int e = km_mstep_product(ws, L, 0, L.n, stream)) return e;
      if (int e = km_update_fused(ws, L, 2, nullptr, nullptr, nullptr, 0, stream)) return e;
    } else {
      if (int e = vidseg_kmeans_partial(workspace, workspace_bytes, 0, L.n, nullptr, nullptr, stream)) return e;
      if (int e = vidseg_kmeans_update(workspace, workspace_bytes, nullptr, nullptr, 0, stream)) return e;
    }
  }
  return 0;
}

VS_API int vidseg_kmeans_fit_predict(const float* x, int n, int d, int k, int n_init, int n_trials, int max_iter,
                                         float tol_rel, const int32_t* first_idx_host, const double* rand_host,
                                         int32_t* labels_out, float* centers_out, int32_t* info_host,
                                         float* inertia_host, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(first_idx_host != nullptr && labels_out != nullptr && centers_out != nullptr, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long launches0 = g_launch_count.load();
  if (int e = vidseg_kmeans_prepare(x, n, d, k, n_init, n_trials, tol_rel, max_iter, workspace, workspace_bytes, stream)) return e;
  KmLayout L;
  if (int e = km_check_ws(workspace, workspace_bytes, &L)) return e;
  void* ws = workspace;
  VS_CHECK_CUDA(cudaMemcpyAsync(at<int>(ws, L.first_idx), first_idx_host, (size_t)n_init * 4, cudaMemcpyHostToDevice, st));
  if (k > 1) {
    VS_REQUIRE(rand_host != nullptr, "null rand_host");
    VS_CHECK_CUDA(cudaMemcpyAsync(at<double>(ws, L.rand), rand_host, (size_t)n_init * (k - 1) * n_trials * 8,
                                  cudaMemcpyHostToDevice, st));
  }
  if (int e = vidseg_kmeans_seed(at<int>(ws, L.first_idx), at<double>(ws, L.rand), workspace, workspace_bytes, stream)) return e;
  int it = 0;
  int poll = 4;
  while (it < max_iter) {
    const int burst = (max_iter - it) < poll ? (max_iter - it) : poll;
    if (int e = vidseg_kmeans_lloyd(workspace, workspace_bytes, burst, stream)) return e;
    it += burst;
    int active = 0;
    if (int e = vidseg_kmeans_active_runs(workspace, workspace_bytes, &active, stream)) return e;
    if (active == 0) break;
    if (poll < 16) poll *= 2;
  }
  if (int e = vidseg_kmeans_inertia(workspace, workspace_bytes, 0, n, nullptr, stream)) return e;
  if (int e = vidseg_kmeans_select(workspace, workspace_bytes, nullptr, centers_out, nullptr, info_host, inertia_host, stream)) return e;
  // KMeans.predict on the un-centred data (sklearn/_kmeans.py:1075-1107)
  if (int e = vidseg_kmeans_predict(x, n, d, centers_out, k, labels_out, at<double>(ws, L.cnorm), stream)) return e;
  VS_CHECK_CUDA(cudaStreamSynchronize(st));
  if (info_host) info_host[3] = (int)(g_launch_count.load() - launches0);
  return 0;
}

VS_API int vidseg_kmeans_release(void* workspace) {
  std::lock_guard<std::mutex> lk(g_km_mu);
  g_km_registry.erase(workspace);
  return 0;
}


// ------------------------------------------------------------------------------------------
// match_gt_mask entry points
// ------------------------------------------------------------------------------------------
VS_API int vidseg_majority_map(const int32_t* fake_labels, const int32_t* gt_labels, int n, int num_fake,
                               int32_t* ref_labels, int32_t* err_flag, void* stream) {
  VS_REQUIRE(fake_labels && gt_labels && ref_labels && err_flag, "null pointer");
  VS_REQUIRE(n >= 0 && num_fake >= 1, "bad shape");
  if (n == 0) return 0;
  VS_CHECK_CUDA(cudaMemsetAsync(err_flag, 0, 4, (cudaStream_t)stream));
  VS_LAUNCH(mg_majority_kernel, 1, 1024, 0, stream, fake_labels, gt_labels, n, num_fake, ref_labels, err_flag);
  VS_POST_LAUNCH();
  return 0;
}

VS_API size_t vidseg_knn_workspace_bytes(int n_ref, int n_query, int d) {
  if (n_ref <= 0 || n_query <= 0 || d <= 0) return 0;
  return knn_layout(n_ref, n_query, d).total;
}

VS_API int vidseg_knn_predict(const float* ref, const int32_t* ref_labels, int n_ref, const float* query, int n_query, int d,
                              int k, int32_t* labels_out, int32_t* err_flag, void* workspace, size_t workspace_bytes,
                              void* stream) {
  VS_REQUIRE(ref && ref_labels && query && labels_out && err_flag && workspace, "null pointer");
  VS_REQUIRE(n_ref >= 1 && n_query >= 0 && d >= 8 && d % 8 == 0, "need D % 8 == 0");
  VS_REQUIRE(k >= 1 && k <= kKnnMaxK && k <= n_ref, "k must be 1..8 and <= n_ref");
  if (n_query == 0) return 0;
  const KnnLayout L = knn_layout(n_ref, n_query, d);
  if (workspace_bytes < L.total)
    return set_error(VIDSEG_E_WORKSPACE, "%s: need %lld bytes, got %lld", "k-NN workspace", (long long)L.total, (long long)workspace_bytes);
  void* ws = workspace;
  cudaStream_t st = (cudaStream_t)stream;
  VS_CHECK_CUDA(cudaMemsetAsync(at<unsigned>(ws, L.absmax), 0, 16, st));
  VS_CHECK_CUDA(cudaMemsetAsync(err_flag, 0, 4, st));
  VS_LAUNCH(knn_norm_kernel, (n_ref * 32 + 255) / 256, 256, 0, st, ref, n_ref, d, at<double>(ws, L.rn), at<unsigned>(ws, L.absmax));
  VS_POST_LAUNCH();
  VS_LAUNCH(knn_norm_kernel, (n_query * 32 + 255) / 256, 256, 0, st, query, n_query, d, at<double>(ws, L.qn), at<unsigned>(ws, L.absmax));
  VS_POST_LAUNCH();
  VS_LAUNCH(knn_max_kernel, 1, 1024, 0, st, at<double>(ws, L.rn), n_ref, at<double>(ws, L.rmax));
  VS_POST_LAUNCH();
  const size_t r_valid = (size_t)n_ref * d, r_total = (size_t)L.nr_pad * d;
  VS_LAUNCH(km_split_scaled_kernel, (int)std::min<size_t>((r_total + 255) / 256, (size_t)kNumSMs * 16), 256, 0, st, ref, r_valid,
            r_total, at<unsigned>(ws, L.absmax), at<__half>(ws, L.rs_hi), at<__half>(ws, L.rs_lo));
  VS_POST_LAUNCH();
  double rmax = 0.0;   // the band needs max ||r||: a scalar read-back (this mode runs once per window, not per step)
  VS_CHECK_CUDA(cudaMemcpyAsync(&rmax, at<double>(ws, L.rmax), 8, cudaMemcpyDeviceToHost, st));
  VS_CHECK_CUDA(cudaStreamSynchronize(st));
  const double band = 2.0 * ((double)d * 0x1p-22 + 0x1p-20);
  for (int q0 = 0; q0 < n_query; q0 += kKnnChunk) {
    const int qc = std::min(kKnnChunk, n_query - q0);
    const size_t q_total = (size_t)qc * d;
    VS_LAUNCH(km_split_scaled_kernel, (int)std::min<size_t>((q_total + 255) / 256, (size_t)kNumSMs * 16), 256, 0, st,
              query + (size_t)q0 * d, q_total, q_total, at<unsigned>(ws, L.absmax), at<__half>(ws, L.qs_hi), at<__half>(ws, L.qs_lo));
    VS_POST_LAUNCH();
    if (int e = gemm_split_run(at<__half>(ws, L.qs_hi), at<__half>(ws, L.qs_lo), at<__half>(ws, L.rs_hi), at<__half>(ws, L.rs_lo),
                               at<float>(ws, L.sdot), qc, L.nr_pad, d, 1.0f, kFamKMeans, stream))
      return e;
    VS_LAUNCH(knn_select_kernel, (qc + kKnnWarps - 1) / kKnnWarps, kKnnWarps * 32, 0, st, query + (size_t)q0 * d, ref, ref_labels,
              qc, n_ref, d, k, at<float>(ws, L.sdot), L.nr_pad, at<double>(ws, L.rn), at<double>(ws, L.qn) + q0,
              at<unsigned>(ws, L.absmax), band, sqrt(rmax), labels_out + q0, err_flag);
    VS_POST_LAUNCH();
  }
  return 0;
}
