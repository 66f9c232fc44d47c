// Linear layers of the attention blocks on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the nn.Linear call sites of sgm/modules/attention.py: to_q/to_k/to_v (:308,315,317),
// to_out (:364), GEGLU.proj (:95), FeedForward.net[2] (:110-112), SpatialTransformer.proj_in/proj_out
// (:903,923).  out[M,N] = A[M,K] . W[N,K]^T (+ bias[N]) (+ residual[M,N]).
//
// Precision: the parity bar of the path is 1e-3 against the reference's fp32 CPU run; single-pass
// fp16/tf32 operands (10-bit mantissa) measure 1.4-1.8e-3 on the stashed q features, so operands are
// carried as fp16 pairs x ~= hi + lo/2048 (22 significant bits) and every product is three tensor-core
// MMAs with fp32 accumulation in TMEM:  acc0 += A_hi.B_hi ;  acc1 += A_hi.B_lo + A_lo.B_hi ;
// out = acc0 + acc1/2048.  Weights are split once at load time, activations by the producing kernel.
//
// Kernel: persistent, one CTA per SM, warp-specialised.  warp 0 = TMA producer (4 tiles per stage:
// A_hi, A_lo [128x64], B_hi, B_lo [BNx64], 128-byte swizzle), warp 1 = MMA issuer (one elected
// thread, 12 tcgen05.mma per 64-wide k-block), warp 2 = TMEM allocator, warps 4-7 = epilogue
// (tcgen05.ld 32x32b -> registers -> bias/residual -> fp32 and/or fp16 hi/lo stores).  3-stage smem
// ring (mbarrier full/empty), 2-stage TMEM accumulator ring (tmem_full/tmem_empty) so the epilogue of
// tile i overlaps the MMAs of tile i+1.
#include <mutex>

#define VS_FAMILY vidseg::kFamGemm
#include "common.cuh"
#include "tc_common.cuh"

namespace vidseg {

// ---------------------------------------------------------------------------------------------
// tensor-map encoding via the driver entry point
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(VIDSEG_E_UNSUPPORTED, "%s", "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VIDSEG_E_INVALID, "%s: CUresult %lld (rank %lld)", "cuTensorMapEncodeTiled failed", (long long)r, (long long)rank);
  return 0;
}

int encode_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t strides[1] = {outer_stride_bytes};
  const uint32_t box[2] = {box_inner, box_outer};
  return encode_tmap_f16(out, base, 2, dims, strides, box);
}
int encode_tmap_3d_f16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  const uint64_t dims[3] = {d0, d1, d2};
  const uint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const uint32_t box[3] = {b0, b1, b2};
  return encode_tmap_f16(out, base, 3, dims, strides, box);
}

// ---------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) fp16 split, elementwise, 128-bit loads / 64-bit stores
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, size_t n4, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    __half h[4], l[4];
    tc::split_f16(v.x, h[0], l[0]);
    tc::split_f16(v.y, h[1], l[1]);
    tc::split_f16(v.z, h[2], l[2]);
    tc::split_f16(v.w, h[3], l[3]);
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
  }
  // tail (n not a multiple of 4)
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    __half h, l;
    tc::split_f16(x[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// ---------------------------------------------------------------------------------------------
// split GEMM
// ---------------------------------------------------------------------------------------------
constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 64, kGemmStages = 3, kGemmAccStages = 2;
constexpr int kTileABytes = kGemmBM * kGemmBK * 2;  // 16 KB
constexpr int kTileBBytes = kGemmBN * kGemmBK * 2;  // 16 KB
constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;
constexpr int kGemmSmemBytes = kGemmStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  int m, n, k;
  const float* bias;      // [N] or null
  const float* residual;  // [M, N] or null
  float* out_f32;         // [M, N] or null
  __half* out_hi;         // [M, N] or null
  __half* out_lo;
};

__global__ void __launch_bounds__(256, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
                  const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
                  const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGemmStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kGemmStages;
  uint64_t* tmem_full_bar = empty_bar + kGemmStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + kGemmAccStages;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + kGemmAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.m + kGemmBM - 1) / kGemmBM;
  const int n_tiles = (p.n + kGemmBN - 1) / kGemmBN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.k + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a_hi);
    tc::prefetch_tmap(&tmap_a_lo);
    tc::prefetch_tmap(&tmap_b_hi);
    tc::prefetch_tmap(&tmap_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kGemmStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < kGemmAccStages; ++s) { tc::mbar_init(&tmem_full_bar[s], 1); tc::mbar_init(&tmem_empty_bar[s], 128); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_base_ptr);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kGemmBM;
        const int n0 = (tile % n_tiles) * kGemmBN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          tc::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          const int k0 = kb * kGemmBK;
          tc::tma_load_2d(st, &tmap_a_hi, &full_bar[stage], k0, m0);
          tc::tma_load_2d(st + kTileABytes, &tmap_a_lo, &full_bar[stage], k0, m0);
          tc::tma_load_2d(st + 2 * kTileABytes, &tmap_b_hi, &full_bar[stage], k0, n0);
          tc::tma_load_2d(st + 2 * kTileABytes + kTileBBytes, &tmap_b_lo, &full_bar[stage], k0, n0);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(kGemmBM, kGemmBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        tc::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * kGemmBN);
        const uint32_t d_cross = d_main + kGemmBN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * kStageBytes);
          const uint64_t a_hi = tc::make_sw128_desc(sa);
          const uint64_t a_lo = tc::make_sw128_desc(sa + kTileABytes);
          const uint64_t b_hi = tc::make_sw128_desc(sa + 2 * kTileABytes);
          const uint64_t b_lo = tc::make_sw128_desc(sa + 2 * kTileABytes + kTileBBytes);
#pragma unroll
          for (int ks = 0; ks < kGemmBK / 16; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 32 >> 4);  // 16 fp16 = 32 bytes along K inside the swizzle atom
            const uint32_t accum = (kb > 0 || ks > 0) ? 1u : 0u;
            tc::umma_f16(d_main, a_hi + adv, b_hi + adv, idesc, accum);
            tc::umma_f16(d_cross, a_hi + adv, b_lo + adv, idesc, accum);
            tc::umma_f16(d_cross, a_lo + adv, b_hi + adv, idesc, 1u);
          }
          tc::umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          if (kb == k_blocks - 1) tc::umma_commit(&tmem_full_bar[acc]);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;  // TMEM lanes [32*ew, 32*ew+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * kGemmBM;
      const int n0 = (tile % n_tiles) * kGemmBN;
      tc::mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc::tc_fence_after();
      const int row = m0 + ew * 32 + lane;
      const uint32_t t_main = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * 2 * kGemmBN);
#pragma unroll 1
      for (int c = 0; c < kGemmBN; c += 32) {
        uint32_t r0[32], r1[32];
        tc::tmem_ld_32x32(t_main + c, r0);
        tc::tmem_ld_32x32(t_main + kGemmBN + c, r1);
        tc::tmem_wait_ld();
        const int col0 = n0 + c;
        if (row < p.m && col0 < p.n) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r1[j]), tc::kLoInv, __uint_as_float(r0[j]));
          const int ncols = min(32, p.n - col0);  // N % 8 == 0 is required by the host wrapper
          const size_t off = (size_t)row * p.n + col0;
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
          }
          if (p.residual) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncols) {
                const float4 rr = *reinterpret_cast<const float4*>(p.residual + off + j);
                v[j] += rr.x; v[j + 1] += rr.y; v[j + 2] += rr.z; v[j + 3] += rr.w;
              }
          }
          if (p.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncols) *reinterpret_cast<float4*>(p.out_f32 + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (p.out_hi) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (j < ncols) {
                uint4 hv, lv;
                tc::split8_f16(v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7], hv, lv);
                *reinterpret_cast<uint4*>(p.out_hi + off + j) = hv;
                *reinterpret_cast<uint4*>(p.out_lo + off + j) = lv;
              }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc<512>(tmem_base);
}

}  // namespace vidseg

using namespace vidseg;

#undef VS_FAMILY
#define VS_FAMILY vidseg::kFamElementwise
VS_API int vidseg_split_f16(const float* x, void* hi, void* lo, long long n, void* stream) {
  VS_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  VS_REQUIRE(x && hi && lo, "null pointer");
  VS_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)hi % 8 == 0) && ((uintptr_t)lo % 8 == 0), "unaligned pointer");
  const size_t n4 = (size_t)n / 4;
  int grid = (int)std::min<size_t>((n4 + 255) / 256 + 1, (size_t)kNumSMs * 8);
  VS_LAUNCH_W(8.0 * n, split_f16_kernel, grid, 256, 0, stream, x, (__half*)hi, (__half*)lo, n4, (size_t)n);
  VS_POST_LAUNCH();
  return 0;
}

#undef VS_FAMILY
#define VS_FAMILY vidseg::kFamGemm
VS_API int vidseg_gemm_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                             const float* residual, float* out_f32, void* out_hi, void* out_lo, int m, int n, int k,
                             void* stream) {
  VS_REQUIRE(a_hi && a_lo && w_hi && w_lo, "null operand pointer");
  VS_REQUIRE(out_f32 != nullptr || (out_hi != nullptr && out_lo != nullptr), "no output requested");
  VS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "out_hi and out_lo go together");
  VS_REQUIRE(m >= 0 && n >= 1 && k >= 1, "bad shape");
  VS_REQUIRE(n % 8 == 0 && k % 8 == 0, "N and K must be multiples of 8 (16-byte rows for TMA and vector stores)");
  if (m == 0) return 0;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if (int e = encode_tmap_2d_f16(&ta_hi, a_hi, k, m, (uint64_t)k * 2, kGemmBK, kGemmBM)) return e;
  if (int e = encode_tmap_2d_f16(&ta_lo, a_lo, k, m, (uint64_t)k * 2, kGemmBK, kGemmBM)) return e;
  if (int e = encode_tmap_2d_f16(&tb_hi, w_hi, k, n, (uint64_t)k * 2, kGemmBK, kGemmBN)) return e;
  if (int e = encode_tmap_2d_f16(&tb_lo, w_lo, k, n, (uint64_t)k * 2, kGemmBK, kGemmBN)) return e;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  });
  VS_CHECK_CUDA(attr_err);
  GemmParams p{m, n, k, bias, residual, out_f32, (__half*)out_hi, (__half*)out_lo};
  const int m_tiles = (m + kGemmBM - 1) / kGemmBM, n_tiles = (n + kGemmBN - 1) / kGemmBN;
  const int grid = std::min(m_tiles * n_tiles, kNumSMs);
  VS_LAUNCH_W(2.0 * m * n * k, gemm_split_kernel, grid, 256, kGemmSmemBytes, stream, ta_hi, ta_lo, tb_hi, tb_lo, p);
  VS_POST_LAUNCH();
  return 0;
}
