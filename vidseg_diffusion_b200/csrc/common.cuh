// Shared helpers for libvidseg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/vidseg_b200.h"

#define VS_API extern "C" __attribute__((visibility("default")))

namespace vidseg {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launch_count;

inline int set_error(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b, c);
  return code;
}

#define VS_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      snprintf(vidseg::g_last_error, sizeof(vidseg::g_last_error), "%s:%d %s -> %s", __FILE__, \
               __LINE__, #expr, cudaGetErrorString(_e));                                      \
      return (int)_e;                                                                         \
    }                                                                                         \
  } while (0)

#define VS_REQUIRE(cond, msg)                                                              \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      snprintf(vidseg::g_last_error, sizeof(vidseg::g_last_error), "%s:%d %s (%s)", __FILE__, \
               __LINE__, msg, #cond);                                                      \
      return VIDSEG_E_INVALID;                                                             \
    }                                                                                      \
  } while (0)

// Kernel families for the live per-kernel timing bench.py asks for (vidseg_profile_*).
enum KernelFamily {
  kFamOther = 0, kFamGemm = 1, kFamAttention = 2, kFamConv = 3, kFamAggregate = 4, kFamKMeans = 5, kFamRefine = 6,
  kFamElementwise = 7, kNumFamilies = 8
};
extern std::atomic<int> g_profile_on;
void profile_before(int family, double work, cudaStream_t stream);
void profile_after(cudaStream_t stream);

#ifndef VS_FAMILY
#define VS_FAMILY vidseg::kFamOther
#endif

// every kernel launch in the library goes through this so that launches are counted; when profiling
// is enabled the launch is bracketed by CUDA events on its own stream (work = algorithmic FLOPs or bytes)
#define VS_LAUNCH_FW(family, work, kernel, grid, block, smem, stream, ...)                         \
  do {                                                                                             \
    const bool _prof = vidseg::g_profile_on.load(std::memory_order_relaxed) != 0;                  \
    if (_prof) vidseg::profile_before((family), (double)(work), (cudaStream_t)(stream));           \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                      \
    if (_prof) vidseg::profile_after((cudaStream_t)(stream));                                      \
    vidseg::g_launch_count.fetch_add(1, std::memory_order_relaxed);                                \
  } while (0)
#define VS_LAUNCH_W(work, kernel, grid, block, smem, stream, ...) \
  VS_LAUNCH_FW(VS_FAMILY, work, kernel, grid, block, smem, stream, __VA_ARGS__)
#define VS_LAUNCH(kernel, grid, block, smem, stream, ...) \
  VS_LAUNCH_FW(VS_FAMILY, 0.0, kernel, grid, block, smem, stream, __VA_ARGS__)

// Programmatic dependent launch for chains of short, strictly dependent kernels (the Lloyd iteration: six launches of
// 6-23 us each): the kernel may be scheduled while its predecessor in the stream is still draining, runs its prologue and
// blocks in pdl_wait() until the predecessor has completed and flushed -- the launch latency and the prologue leave the
// critical path.  ONLY for kernels that call pdl_wait() before their first access to global memory (and
// pdl_launch_dependents() to let their own successor in).  Falls back to a plain launch under the per-launch profiler.
bool pdl_enabled(int link = 0);   // link: bit index of the chain link (debug mask VIDSEG_KM_PDL)
#define VS_LAUNCH_PDL_L(link, kernel, grid, block, smem, stream, ...)                                      \
  do {                                                                                               \
    if (vidseg::g_profile_on.load(std::memory_order_relaxed) != 0 || !vidseg::pdl_enabled(link)) {    \
      VS_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__);                                     \
    } else {                                                                                         \
      cudaLaunchConfig_t _cfg = {};                                                                  \
      _cfg.gridDim = dim3(grid);                                                                     \
      _cfg.blockDim = dim3(block);                                                                   \
      _cfg.dynamicSmemBytes = (smem);                                                                \
      _cfg.stream = (cudaStream_t)(stream);                                                          \
      cudaLaunchAttribute _at[1];                                                                    \
      _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                \
      _at[0].val.programmaticStreamSerializationAllowed = 1;                                         \
      _cfg.attrs = _at;                                                                              \
      _cfg.numAttrs = 1;                                                                             \
      VS_CHECK_CUDA(cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__));                                 \
      vidseg::g_launch_count.fetch_add(1, std::memory_order_relaxed);                                \
    }                                                                                                \
  } while (0)

#define VS_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) VS_LAUNCH_PDL_L(0, kernel, grid, block, smem, stream, __VA_ARGS__)

#define VS_POST_LAUNCH() VS_CHECK_CUDA(cudaGetLastError())

constexpr int kNumSMs = 148;  // B200

// see VS_LAUNCH_PDL; both are no-ops in a kernel launched without the attribute
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// for kernels whose first reads are TMA loads (async proxy) of data the predecessor wrote with ordinary stores
__device__ __forceinline__ void pdl_wait_async_proxy() {
  asm volatile("griddepcontrol.wait;\n\tfence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// x * Phi(x) with Phi through erf(t) = 1 - exp(-q(t)), t = |x| / sqrt(2), q a degree-6 polynomial without constant term
// fitted (minimax, scipy) to -ln(erfc(t)) on [0, 4.3]: |erf error| <= 2.8e-7 in fp32, |gelu error| <= 5.8e-7 absolute and
// 2e-7 |x| (erff() itself is 1 ulp of 1).  One range, no selects: 11 instructions against ~32 for 0.5 x (1 + erff(x / sqrt 2)),
// which is what bounded the fused GEGLU epilogue (50 instructions per output on eight warps against 5 120 MMA cycles per tile).
// The coefficients carry -log2(e), so the exponential is one ex2.approx; t is clamped at 6 (erf = 1 in fp32 from 3.9 on; the
// polynomial turns around near t = 17).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 6.0f);
  float p = 1.420341199e-04f;
  p = fmaf(p, t, -3.664243535e-03f);
  p = fmaf(p, t, 3.089613741e-02f);
  p = fmaf(p, t, -1.496994018e-01f);
  p = fmaf(p, t, -9.181654774e-01f);
  p = fmaf(p, t, -1.627925070e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * t));
  const float hx = 0.5f * x, a = fabsf(hx);
  return fmaf(-a, e, hx + a);   // 0.5 x + 0.5 |x| erf(t)
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load that does not pollute L1 (inputs are read once)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// K-means operands are carried scaled by s = 2^(11 - exponent(max |x centred|)) so that the fp16 pairs keep 22 bits
__device__ __forceinline__ float km_operand_scale(unsigned absmax_bits) {
  const float m = __uint_as_float(absmax_bits);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);               // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(11-e) < 2048
  return ldexpf(1.f, 11 - e);
}

}  // namespace vidseg
