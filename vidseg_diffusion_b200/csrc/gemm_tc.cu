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
// thread: 12 kind::f16 MMAs per 64-wide k-block with fp16-pair operands, 4 kind::f16 + 4 kind::f8f6f4 with
// the fp16 + fp8-correction operands), warp 2 = TMEM allocator, warps 4-11 = epilogue (tcgen05.ld 32x32b ->
// registers -> 1 KB shared-memory transpose per warp -> scale / bias / residual / blend / GEGLU gate in a
// layout whose global accesses cover full 128-byte lines -> fp32 and/or operand stores).  2-4 stage smem ring
// (mbarrier full/empty), 2-stage TMEM accumulator ring (tmem_full/tmem_empty) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  Large grids run as CTA PAIRS (tcgen05.mma.cta_group::2, M = 256 over the
// two SMs of a TPC, each CTA half of the B tile): see gemm_split_kernel.
#include <cstdlib>
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

int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(VIDSEG_E_UNSUPPORTED, "%s", "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16 /* bf16 tiles move identically */, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
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
  return encode_tmap_16bit(out, base, 2, dims, strides, box);
}
int encode_tmap_3d_f16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  const uint64_t dims[3] = {d0, d1, d2};
  const uint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const uint32_t box[3] = {b0, b1, b2};
  return encode_tmap_16bit(out, base, 3, dims, strides, box);
}

// ---------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) fp16 split of x * scale, elementwise, 128-bit loads / 64-bit stores
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, size_t n4, size_t n, float scale) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 h, l;
    tc::split4_f16(v.x * scale, v.y * scale, v.z * scale, v.w * scale, h, l);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
  // tail (n not a multiple of 4)
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    __half h;
    __half l;
    tc::split_f16(x[i] * scale, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// row-aware form: x [rows, cols] -> operand in the format the library policy gives rows of `cols` channels
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                                         __half* __restrict__ lo, long long rows, int cols, float scale,
                                                         int packed8, float sx, float sl) {
  const int nq = cols >> 2;
  const long long total = rows * nq;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / nq;
    const int q = (int)(i - row * nq);
    const float4 v = reinterpret_cast<const float4*>(x + row * cols)[q];
    tc::store_split4(hi + row * cols, lo + row * cols, q * 4, v.x * scale, v.y * scale, v.z * scale, v.w * scale,
                     packed8 != 0, sx, sl);
  }
}

// ---------------------------------------------------------------------------------------------
// split GEMM / implicit-GEMM convolution
//
// out[p, n] = sum_{tap, c} A[pixel p shifted by tap, c] * W[n, tap*Cin + c]  (+ bias[n]) (+ chan_bias[b(p), n])
//             (+ residual[p, n])
// The A operand is a channels-last activation tensor seen through a rank-5 TMA map
// (C', W', P, H', B): a 128-pixel tile is a bw x bh x bb patch of the OUTPUT grid, and the tile of tap (dy, dx) is
// the same patch shifted by the tap offset -- out-of-range pixels (the zero padding of the convolution) are
// filled by the TMA unit, so no im2col matrix ever exists.  A plain Linear is the 1-tap case with W' = M.
// Stride-2 convolutions view [B, H, W, C] as [B, H/2, 2, W/2, 2C] so that the row/column parity of a tap becomes
// a coordinate (P) / channel offset.
// ---------------------------------------------------------------------------------------------
constexpr int kGemmBM = 128, kGemmBK = 64, kGemmAccStages = 2;
constexpr int kTileABytes = kGemmBM * kGemmBK * 2;  // 16 KB
constexpr int kMaxTaps = 9;
// warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = spare, 4..11 = epilogue.  Eight epilogue warps: a
// TMEM lane quarter (warp id % 4) is shared by two warps that take alternate 32-column chunks of the tile, so that every
// scheduler has two epilogue warps to interleave (with one warp per scheduler the epilogue ran at one instruction per ~7
// cycles -- dependent issue, instruction-cache misses on the long unrolled body -- and bounded every GEMM with a short K
// loop: ncu on the K = 320 projections showed 28 % of the DRAM roof with 24 % of the tensor pipe).
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = (kGemmThreads - 128) / 32;
// Epilogue staging: TMEM hands every thread one ROW of the tile (32 consecutive columns per load), so stores issued from
// that layout touch 32 different cache lines per instruction with 16 bytes each.  Each epilogue warp therefore turns its
// 32 x 32 block through a private 8-row x 128-byte shared-memory buffer (four passes of 8 rows, XOR-swizzled float4
// slots, conflict-free both ways) and reads / writes global memory with 8 lanes per row: every request covers four full
// 128-byte lines of the fp32 output and residual, full 32-byte sectors of the operand outputs.  All element-wise terms
// (scale, bias, residual, blend) are applied in that layout, where a lane keeps the same four columns for all rows.
constexpr int kEpiWarpBytes = 8 * 128;
constexpr int kEpiStageBytes = kEpiWarps * kEpiWarpBytes;

template <int BN, int CL = 1>
struct GemmCfg {
  static constexpr int kTileBBytes = (BN / CL) * kGemmBK * 2;   // a CTA of a pair holds half of the B rows
  static constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;
  static constexpr int kStages = (CL == 2) ? (BN <= 160 ? 4 : 3) : ((BN <= 128) ? 3 : (BN <= 160 ? 3 : 2));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiStageBytes;
  static constexpr int kAccStride = (BN <= 128) ? 128 : 256;  // TMEM columns per accumulator stage
  static constexpr int kTmemCols = (BN <= 128) ? 256 : 512;
};

struct GemmParams {
  int n, k;                 // GEMM N and K (K = taps * cin)
  int taps, cin, kc_per_tap;
  int bw, bh, bb;           // tile patch (bw * bh * bb == 128)
  int wo, ho, nb;           // output pixel grid; M = nb * ho * wo
  int tiles_w, tiles_h, tiles_b;
  int c_off[kMaxTaps], w_off[kMaxTaps], p_idx[kMaxTaps], h_off[kMaxTaps];
  int geglu;                // fused GEGLU epilogue: columns come in 64-wide groups [32 value | 32 gate], out[.., N/2] = value * gelu(gate)
  int in_packed8;           // operands in the packed8 format (tc_common.cuh): 4 fp16 + 4 fp8 MMAs per k-block instead of 12
  int out_packed8;          // format of the split output (out_hi / out_lo)
  float acc_scale;          // exact power-of-two inverse of the operand pre-scaling (weights carry 2^8)
  const float* bias;        // [N] or null
  const float* chan_bias;   // [M / cb_div, N] or null: bias per group of cb_div consecutive output rows (the ResBlock
                            // time embedding per sample; per-frame / per-video vectors of the temporal layers)
  long long cb_div;
  const float* residual;    // [M, N] or null
  const float* row_scalar;  // [M] or null: one value per output row added to all of its columns (mask modulation)
  const float* blend;       // [M, N] or null: out = a * blend + (1 - a) * out, a = blend_alpha[row / ba_div] (AlphaBlender)
  const float* blend_alpha;
  long long ba_div;
  float* out_f32;           // [M, N] or null
  __half* out_hi;           // [M, N] or null
  __half* out_lo;
  int mma_n;                // N of the MMA instruction (multiple of 16, <= BN); 0 = BN
  // segmented output (the fused q | k | v projection): columns [s * seg_n, (s + 1) * seg_n) go to the s-th set of
  // [M, seg_n] tensors instead of out_f32 / out_hi / out_lo; seg_n is a multiple of the N tile.  0 = off
  int seg_n;
  float* seg_f32[3];
  __half* seg_hi[3];
  __half* seg_lo[3];
};

// CL = 2: a CTA PAIR (two CTAs of a cluster on the two SMs of a TPC) works on two adjacent M tiles of the SAME N tile
// with ONE tcgen05.mma.cta_group::2 of M = 256 per k-slice, issued by the leader (cluster rank 0).  Each CTA loads its own
// A tile and HALF of the B tile (BN / 2 weight rows) into its own shared memory -- per CTA the operand bytes per k-block
// drop from A + B to A + B/2, both the L2 -> SM traffic (what bounds the short-K GEMMs) and the shared-memory reads of
// the tensor core (what bounds the long ones) -- and reads its own 128 accumulator rows from its own TMEM.  The TMA loads
// of both CTAs complete on the LEADER's full barrier; the leader's commits arrive on both CTAs' empty / tmem_full
// barriers; the epilogue warps of both CTAs arrive on the leader's tmem_empty barrier.
template <int BN, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
                  const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
                  const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, CL>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + kGemmAccStages;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + kGemmAccStages);
  // same bytes as smem + ..., derived from the __shared__ array itself so that the compiler keeps the shared address
  // space (LDS / STS instead of generic accesses) for the epilogue staging buffers
  uint8_t* epi_smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u) + kStages * Cfg::kStageBytes + 256;

  // launched as a programmatic dependent (the K-means score GEMM, see VS_LAUNCH_PDL): wait for the predecessor first.
  // (With the barrier initialisation / TMEM allocation ahead of the wait the fit's iteration count changed from run to run,
  // measured; only the launch latency is hidden, not the prologue.)
  pdl_wait_async_proxy();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int n_tiles = (p.n + BN - 1) / BN;
  const int k_blocks = p.taps * p.kc_per_tap;
  // work items: (group of CL adjacent M tiles, N tile); CTA `crank` of the cluster takes M tile CL * group + crank
  // (beyond the last M tile the TMA out-of-bounds fill and the epilogue's row check make the CTA a no-op that still
  // keeps the pair's barriers in step)
  const int crank = (CL > 1) ? (int)tc::cluster_ctarank() : 0;
  const int num_tiles = ((m_tiles + CL - 1) / CL) * n_tiles;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;
  auto m_tile_of = [&](int item) { return (item / n_tiles) * CL + crank; };

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a_hi);
    tc::prefetch_tmap(&tmap_a_lo);
    tc::prefetch_tmap(&tmap_b_hi);
    tc::prefetch_tmap(&tmap_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    // tmem_empty: one arrival per epilogue warp, of both CTAs of a pair (only the leader's barrier is used then)
    for (int s = 0; s < kGemmAccStages; ++s) { tc::mbar_init(&tmem_full_bar[s], 1); tc::mbar_init(&tmem_empty_bar[s], kEpiWarps * CL); }
    tc::fence_barrier_init();
  }
  if (warp == 2) { if (CL == 2) tc::tmem_alloc_pair<Cfg::kTmemCols>(tmem_base_ptr); else tc::tmem_alloc<Cfg::kTmemCols>(tmem_base_ptr); }
  tc::tc_fence_before();
  __syncthreads();
  if (CL > 1) tc::cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        const int mt = m_tile_of(tile);
        const int n0 = (tile % n_tiles) * BN;
        const int w0 = (mt % p.tiles_w) * p.bw;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
        const int b0 = (mt / (p.tiles_w * p.tiles_h)) * p.bb;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int cw = w0 + p.w_off[tap], ch = h0 + p.h_off[tap], cp = p.p_idx[tap];
          for (int kc = 0; kc < p.kc_per_tap; ++kc) {
            tc::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * Cfg::kStageBytes;
            const int ca = p.c_off[tap] + kc * kGemmBK;
            const int kb0 = tap * p.cin + kc * kGemmBK;
            if (CL == 2) {
              // both CTAs' bytes are counted by the leader's barrier (armed by the leader alone); this CTA's half of the
              // B rows goes to its own shared memory
              if (crank == 0) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              const uint32_t lead_bar = tc::mapa_u32(tc::smem_u32(&full_bar[stage]), 0);
              constexpr int kHalfRows = BN / 2;
              tc::tma_load_5d_pair(st, &tmap_a_hi, lead_bar, ca, cw, cp, ch, b0);
              tc::tma_load_5d_pair(st + kTileABytes, &tmap_a_lo, lead_bar, ca, cw, cp, ch, b0);
              tc::tma_load_2d_pair(st + 2 * kTileABytes, &tmap_b_hi, lead_bar, kb0, n0 + crank * kHalfRows);
              tc::tma_load_2d_pair(st + 2 * kTileABytes + Cfg::kTileBBytes, &tmap_b_lo, lead_bar, kb0, n0 + crank * kHalfRows);
            } else {
              tc::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              tc::tma_load_5d(st, &tmap_a_hi, &full_bar[stage], ca, cw, cp, ch, b0);
              tc::tma_load_5d(st + kTileABytes, &tmap_a_lo, &full_bar[stage], ca, cw, cp, ch, b0);
              tc::tma_load_2d(st + 2 * kTileABytes, &tmap_b_hi, &full_bar[stage], kb0, n0);
              tc::tma_load_2d(st + 2 * kTileABytes + Cfg::kTileBBytes, &tmap_b_lo, &full_bar[stage], kb0, n0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader's alone in a pair) =====================
    if (lane == 0 && crank == 0) {
      constexpr int kMmaM = kGemmBM * CL;
      const uint32_t idesc = tc::make_idesc_f16(kMmaM, p.mma_n > 0 ? p.mma_n : BN, 0, 0);
      auto mma16 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t accum) {
        if (CL == 2) tc::umma_f16_pair(d, a, b, id, accum); else tc::umma_f16(d, a, b, id, accum);
      };
      auto mma8 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t accum) {
        if (CL == 2) tc::umma_f8_pair(d, a, b, id, accum); else tc::umma_f8(d, a, b, id, accum);
      };
      auto commit = [](uint64_t* bar) {   // in a pair: the same barrier of both CTAs
        if (CL == 2) tc::umma_commit_pair(bar, (uint16_t)0x3); else tc::umma_commit(bar);
      };
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        if (CL == 2) tc::mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1); else tc::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_acc = tmem_base + (uint32_t)(acc * Cfg::kAccStride);
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t a_hi = tc::make_sw128_desc(sa);
          const uint64_t a_lo = tc::make_sw128_desc(sa + kTileABytes);
          const uint64_t b_hi = tc::make_sw128_desc(sa + 2 * kTileABytes);
          const uint64_t b_lo = tc::make_sw128_desc(sa + 2 * kTileABytes + Cfg::kTileBBytes);
          if (p.in_packed8) {
            // the a_lo / b_lo tiles hold [lo8 | x8] rows: corrections as fp8 MMAs (K = 32 each), small terms first
            constexpr uint32_t idesc_lx = tc::make_idesc_f8(kMmaM, BN, 1, 0);  // A e5m2 residual x B e4m3 value
            constexpr uint32_t idesc_xl = tc::make_idesc_f8(kMmaM, BN, 0, 1);  // A e4m3 value    x B e5m2 residual
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t adv = (uint64_t)(h * 32 >> 4), xoff = (uint64_t)(64 >> 4);
              mma8(d_acc, a_lo + adv, b_lo + xoff + adv, idesc_lx, (kb > 0 || h > 0) ? 1u : 0u);
              mma8(d_acc, a_lo + xoff + adv, b_lo + adv, idesc_xl, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < kGemmBK / 16; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 32 >> 4);
              mma16(d_acc, a_hi + adv, b_hi + adv, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < kGemmBK / 16; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 32 >> 4);  // 16 elements = 32 bytes along K inside the swizzle atom
              // small terms first, so that they are not absorbed one by one into a large partial sum
              mma16(d_acc, a_lo + adv, b_hi + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
              mma16(d_acc, a_hi + adv, b_lo + adv, idesc, 1u);
              mma16(d_acc, a_hi + adv, b_hi + adv, idesc, 1u);
            }
          }
          commit(&empty_bar[stage]);   // frees the smem stage (of both CTAs of a pair) once these MMAs have read it
          if (kb == k_blocks - 1) commit(&tmem_full_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp & 3;           // TMEM lanes [32*ew, 32*ew+32) (a warp may only touch the quarter warp id % 4)
    const int chalf = (warp - 4) >> 2; // which of the alternating column chunks
    const int r = ew * 32 + lane;      // row of the tile = pixel of the patch (row layout: TMEM loads, prefetches)
    const int pw = r % p.bw, ph = (r / p.bw) % p.bh, pb = r / (p.bw * p.bh);
    // staged layout (see kEpiWarpBytes): lane = (row quarter qrow, 4 columns at qcol); slot u of a chunk is row
    // 4 * u + qrow of this warp's 32 rows
    const int qrow = lane >> 3, qx = lane & 7, qcol = qx * 4;
    uint4* stage = reinterpret_cast<uint4*>(epi_smem + (warp - 4) * kEpiWarpBytes);
    const int st_row = (lane & 7) * 8;   // row written by this lane in its pass (rows are 8 float4 slots)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_item; tile < num_tiles; tile += item_stride) {
      const int mt = m_tile_of(tile);
      const int n0 = (tile % n_tiles) * BN;
      const int w = (mt % p.tiles_w) * p.bw + pw;
      const int h = ((mt / p.tiles_w) % p.tiles_h) * p.bh + ph;
      const int b = (mt / (p.tiles_w * p.tiles_h)) * p.bb + pb;
      const bool row_ok = (w < p.wo) && (h < p.ho) && (b < p.nb);
      const int my_pix = row_ok ? (b * p.ho + h) * p.wo + w : -1;   // M < 2^31 (checked by the host wrappers)
      // The residual (and blend) rows of this tile come from HBM: start them towards L2 now, while the MMAs of the tile
      // are still running
      if (row_ok && (p.residual || p.blend)) {
        const int tile_cols = min(BN, p.n - n0);
        const size_t o = (size_t)my_pix * p.n + n0;
        for (int j = chalf * 32; j < tile_cols; j += 64) {
          if (p.residual) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + o + j));
          if (p.blend) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.blend + o + j));
        }
      }
      int prow[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) prow[u] = __shfl_sync(0xffffffffu, my_pix, 4 * u + qrow);
      float* of32 = p.out_f32;
      __half* ohi = p.out_hi;
      __half* olo = p.out_lo;
      int opitch = p.n, oshift = 0;   // row pitch of the output tensors, first column of the tile's segment
      if (p.seg_n) {
        const int sg = n0 / p.seg_n;
        of32 = p.seg_f32[sg]; ohi = p.seg_hi[sg]; olo = p.seg_lo[sg];
        opitch = p.seg_n; oshift = sg * p.seg_n;
      }
      // rr[32] = this thread's row (32 columns, raw accumulator bits) -> fn(u, four columns at qcol of row 4u + qrow)
      auto staged = [&](const uint32_t (&rr)[32], auto&& fn) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (qrow == q) {
#pragma unroll
            for (int j = 0; j < 8; ++j) stage[st_row + (j ^ qx)] = make_uint4(rr[4 * j], rr[4 * j + 1], rr[4 * j + 2], rr[4 * j + 3]);
          }
          __syncwarp();
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int rl = 4 * s + qrow;
            fn(2 * q + s, stage[rl * 8 + (qx ^ rl)]);
          }
          __syncwarp();
        }
      };
      tc::mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * Cfg::kAccStride);
      if (p.geglu) {
        // FeedForward's GEGLU (attention.py:89-96) fused into the projection GEMM: the weight rows were permuted on the
        // host so that each 64-column group holds 32 value columns followed by their 32 gate columns; the fp32
        // [M, 2D] intermediate (the largest tensor of the UNet) is never written.
        const int dn = p.n >> 1;
#pragma unroll 1
        for (int c = chalf * 64; c < BN; c += 128) {
          const int col0 = n0 + c;
          if (col0 >= p.n) break;  // warp-uniform
          uint32_t rv[32], rg[32];
          tc::tmem_ld_32x32(t_acc + c, rv);
          tc::tmem_ld_32x32(t_acc + c + 32, rg);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
            if (p.bias) {
              bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              bg = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 32 + j));
            }
            const float vb[4] = {bv.x, bv.y, bv.z, bv.w}, gb[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float val = fmaf(__uint_as_float(rv[j + u]), p.acc_scale, vb[u]);
              const float gate = fmaf(__uint_as_float(rg[j + u]), p.acc_scale, gb[u]);
              rv[j + u] = __float_as_uint(val * gelu_erf_fast(gate));
            }
          }
          staged(rv, [&](int u, uint4 x) {
            if (prow[u] >= 0) {
              const size_t ro = (size_t)prow[u] * dn;
              tc::store_split4(p.out_hi + ro, p.out_lo + ro, (col0 >> 1) + qcol, __uint_as_float(x.x), __uint_as_float(x.y),
                               __uint_as_float(x.z), __uint_as_float(x.w), p.out_packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
            }
          });
        }
      } else {
        int crow[8];   // chan_bias row of every slot
        if (p.chan_bias) {
          const int cbd = (int)p.cb_div;
#pragma unroll
          for (int u = 0; u < 8; ++u) crow[u] = prow[u] >= 0 ? prow[u] / cbd : 0;
        }
#pragma unroll 1
        for (int c = chalf * 32; c < BN; c += 64) {
          const int col0 = n0 + c;
          if (col0 >= p.n) break;  // warp-uniform
          uint32_t rr[32];
          tc::tmem_ld_32x32(t_acc + c, rr);
          const bool col_ok = col0 + qcol < p.n;   // N % 4 == 0 is required by the host wrapper
          float4 res[8];
          if (p.residual) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              res[u] = (prow[u] >= 0 && col_ok) ? *reinterpret_cast<const float4*>(p.residual + (size_t)prow[u] * p.n + col0 + qcol)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const float4 bias4 = (p.bias && col_ok) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + qcol))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          tc::tmem_wait_ld();
          staged(rr, [&](int u, uint4 raw) {
            if (prow[u] < 0 || !col_ok) return;
            const size_t off = (size_t)prow[u] * p.n + col0 + qcol;             // residual / blend: [M, N]
            const size_t oro = (size_t)prow[u] * opitch;                         // outputs: [M, N] or the segment's [M, seg_n]
            const int ocol = col0 - oshift + qcol;
            // same order of operations as the unfused chain: (acc * scale + bias) + chan_bias + row_scalar + residual
            // (the scale is a power of two, so the fused multiply-add rounds exactly like multiply, then add)
            float4 x;
            if (p.bias) {
              x = make_float4(fmaf(__uint_as_float(raw.x), p.acc_scale, bias4.x), fmaf(__uint_as_float(raw.y), p.acc_scale, bias4.y),
                              fmaf(__uint_as_float(raw.z), p.acc_scale, bias4.z), fmaf(__uint_as_float(raw.w), p.acc_scale, bias4.w));
            } else {
              x = make_float4(__uint_as_float(raw.x) * p.acc_scale, __uint_as_float(raw.y) * p.acc_scale,
                              __uint_as_float(raw.z) * p.acc_scale, __uint_as_float(raw.w) * p.acc_scale);
            }
            if (p.chan_bias) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(p.chan_bias + (size_t)crow[u] * p.n + col0 + qcol));
              x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
            }
            if (p.row_scalar) {
              const float ra = __ldg(p.row_scalar + prow[u]);
              x.x += ra; x.y += ra; x.z += ra; x.w += ra;
            }
            if (p.residual) {
              const float4 q = res[u];
              x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
            }
            if (p.blend) {
              const float a = __ldg(p.blend_alpha + (long long)prow[u] / p.ba_div), na = 1.0f - a;
              const float4 q = *reinterpret_cast<const float4*>(p.blend + off);
              x.x = a * q.x + na * x.x; x.y = a * q.y + na * x.y; x.z = a * q.z + na * x.z; x.w = a * q.w + na * x.w;
            }
            if (of32) *reinterpret_cast<float4*>(of32 + oro + ocol) = x;
            if (ohi) tc::store_split4(ohi + oro, olo + oro, ocol, x.x, x.y, x.z, x.w, p.out_packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
          });
        }
      }
      // this warp has read its share of the accumulator: one arrival per warp on the (leader's) tmem_empty barrier
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 2) tc::mbar_arrive_cluster(tc::mapa_u32(tc::smem_u32(&tmem_empty_bar[acc]), 0));
        else tc::mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL > 1) tc::cluster_sync_all();   // no CTA leaves (or frees its TMEM) while its peer may still signal / compute into it
  if (warp == 2) { if (CL == 2) tc::tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tc::tmem_dealloc<Cfg::kTmemCols>(tmem_base); }
}

// 128-pixel tile = bw x bh x bb patch (powers of two) of the output grid that pads the grid the least;
// ties go to the widest, then tallest patch (longest contiguous runs for TMA and for the epilogue stores)
static void pick_patch(int wo, int ho, int nb, int* bw_out, int* bh_out, int* bb_out) {
  long best = -1;
  for (int bw = 128; bw >= 1; bw /= 2)
    for (int bh = 128 / bw; bh >= 1; bh /= 2) {
      const int bb = 128 / (bw * bh);
      const long vol = (long)((wo + bw - 1) / bw) * bw * ((ho + bh - 1) / bh) * bh * ((nb + bb - 1) / bb) * bb;
      if (best < 0 || vol < best) { best = vol; *bw_out = bw; *bh_out = bh; *bb_out = bb; }
    }
}

template <int BN, int CL>
static int launch_gemm_cl(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const void* w_hi, const void* w_lo,
                          const GemmParams& p, double flops, int family, void* stream) {
  using Cfg = GemmCfg<BN, CL>;
  CUtensorMap tb_hi, tb_lo;
  if (int e = encode_tmap_2d_f16(&tb_hi, w_hi, p.k, p.n, (uint64_t)p.k * 2, kGemmBK, BN / CL)) return e;
  if (int e = encode_tmap_2d_f16(&tb_lo, w_lo, p.k, p.n, (uint64_t)p.k * 2, kGemmBK, BN / CL)) return e;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_split_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  });
  VS_CHECK_CUDA(attr_err);
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b, n_tiles = (p.n + BN - 1) / BN;
  const int items = ((m_tiles + CL - 1) / CL) * n_tiles;
  const int grid = std::min(items, kNumSMs / CL) * CL;
  const bool prof = g_profile_on.load(std::memory_order_relaxed) != 0;
  if (prof) profile_before(family, flops, (cudaStream_t)stream);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CL > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (family == kFamKMeans && !prof && pdl_enabled(0)) {   // a link of the Lloyd chain (see VS_LAUNCH_PDL)
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    VS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_split_kernel<BN, CL>, ta_hi, ta_lo, tb_hi, tb_lo, p));
  }
  if (prof) profile_after((cudaStream_t)stream);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  VS_POST_LAUNCH();
  return 0;
}

// The CTA-pair form (cta_group::2) for grids with enough work for 74 pairs; VIDSEG_GEMM_PAIR=0 turns it off.  (The
// earlier 2-CTA form that only MULTICAST the B halves into both CTAs was measured 0-3 % faster at best: multicast does not
// reduce the bytes that must land in each SM's shared memory.)
static bool use_cluster(const GemmParams& p, int bn) {
  static const int env = [] { const char* e = getenv("VIDSEG_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (!env || p.mma_n != 0) return false;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b, n_tiles = (p.n + bn - 1) / bn;
  return m_tiles >= 2 && (long long)((m_tiles + 1) / 2) * n_tiles >= kNumSMs;
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const void* w_hi, const void* w_lo,
                       const GemmParams& p, double flops, int family, void* stream) {
  if (use_cluster(p, BN)) return launch_gemm_cl<BN, 2>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
  return launch_gemm_cl<BN, 1>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
}

// shared host path of the Linear and convolution entry points
static int run_gemm(const void* a_hi, const void* a_lo, const uint64_t* adims, const uint64_t* astrides, GemmParams p,
                    const void* w_hi, const void* w_lo, int family, void* stream) {
  pick_patch(p.wo, p.ho, p.nb, &p.bw, &p.bh, &p.bb);
  p.tiles_w = (p.wo + p.bw - 1) / p.bw;
  p.tiles_h = (p.ho + p.bh - 1) / p.bh;
  p.tiles_b = (p.nb + p.bb - 1) / p.bb;
  p.kc_per_tap = (p.cin + kGemmBK - 1) / kGemmBK;
  if (p.in_packed8 < 0) p.in_packed8 = operand_packed8(p.cin) ? 1 : 0;       // library policy unless the caller fixed it
  if (p.out_packed8 < 0) p.out_packed8 = operand_packed8(p.geglu ? p.n / 2 : p.n) ? 1 : 0;
  if (p.cb_div <= 0) p.cb_div = (long long)p.ho * p.wo;  // default: one bias row per sample
  if (p.ba_div <= 0) p.ba_div = 1;
  const uint32_t box[5] = {(uint32_t)kGemmBK, (uint32_t)p.bw, 1u, (uint32_t)p.bh, (uint32_t)p.bb};
  CUtensorMap ta_hi, ta_lo;
  if (int e = encode_tmap_16bit(&ta_hi, a_hi, 5, adims, astrides, box)) return e;
  if (int e = encode_tmap_16bit(&ta_lo, a_lo, 5, adims, astrides, box)) return e;
  const double flops = 2.0 * (double)p.nb * p.ho * p.wo * (double)p.n * (double)p.k;
  // small grids (the 8x8 / 16x16 levels of the UNet): 128-wide N tiles double the number of CTAs when 256-wide ones would
  // leave SMs idle
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
  if (p.seg_n) {   // an N tile must not straddle two segments
    if (p.seg_n % 256 == 0) return launch_gemm<256>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
    if (p.seg_n % 160 == 0) return launch_gemm<160>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
    VS_REQUIRE(p.seg_n % 128 == 0, "segment width must be a multiple of 128 or 160");
    return launch_gemm<128>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
  }
  if (!p.in_packed8 && !p.geglu && p.n <= 128 && p.mma_n == 0) {
    // one N tile whose MMAs are trimmed to the columns that exist (fp16-pair operands: the K-means score GEMM
    // [N, D] x [D, R*K] of a rank of the run-sharded fit, R*K = 20..60).  (Measured: for 128 < N <= 256 a single 256-wide
    // tile is SLOWER than two 128-wide ones, 32.7 vs 23.0 us at [14336, 640] x [640, 200] -- two pipeline stages of 96 KB
    // instead of three of 64 KB -- so those keep the generic choice.)
    p.mma_n = (p.n + 15) / 16 * 16;
    return launch_gemm<128>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
  }
  const bool underfilled = p.n % 128 == 0 && p.n > 128 && m_tiles * ((p.n + 255) / 256) < (3 * kNumSMs) / 4;
  if (!underfilled && (p.n % 256 == 0 || (p.geglu && p.n > 128)))
    return launch_gemm<256>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
  if (underfilled) return launch_gemm<128>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
  if (p.n % 160 == 0 && !p.geglu) return launch_gemm<160>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
  return launch_gemm<128>(ta_hi, ta_lo, w_hi, w_lo, p, flops, family, stream);
}

// internal entry for other translation units (the K-means E-step): plain split GEMM with an explicit family tag
int gemm_split_run(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, float* out_f32, int m, int n,
                   int k, float acc_scale, int family, void* stream) {
  VS_REQUIRE(a_hi && a_lo && w_hi && w_lo && out_f32, "null pointer");
  VS_REQUIRE(m >= 1 && n >= 4 && n % 4 == 0 && k >= 8 && k % 8 == 0, "bad shape");
  GemmParams p{};
  p.n = n; p.k = k; p.taps = 1; p.cin = k;
  p.wo = m; p.ho = 1; p.nb = 1;
  p.acc_scale = acc_scale;
  p.out_f32 = out_f32;
  p.in_packed8 = 0;   // the K-means filter keeps the fp16 pair (its error band is derived for it)
  p.out_packed8 = 0;
  const uint64_t adims[5] = {(uint64_t)k, (uint64_t)m, 1, 1, 1};
  const uint64_t row = (uint64_t)k * 2;
  const uint64_t astrides[4] = {row, row * m, row * m, row * m};
  return run_gemm(a_hi, a_lo, adims, astrides, p, w_hi, w_lo, family, stream);
}

}  // namespace vidseg

using namespace vidseg;

#undef VS_FAMILY
#define VS_FAMILY vidseg::kFamElementwise
VS_API int vidseg_split_f16(const float* x, void* hi, void* lo, long long n, float scale, void* stream) {
  VS_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  VS_REQUIRE(x && hi && lo, "null pointer");
  VS_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)hi % 8 == 0) && ((uintptr_t)lo % 8 == 0), "unaligned pointer");
  const size_t n4 = (size_t)n / 4;
  int grid = (int)std::min<size_t>((n4 + 255) / 256 + 1, (size_t)kNumSMs * 8);
  VS_LAUNCH_W(8.0 * n, split_f16_kernel, grid, 256, 0, stream, x, (__half*)hi, (__half*)lo, n4, (size_t)n, scale);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_split_rows(const float* x, void* hi, void* lo, long long rows, int cols, float scale, int is_weight,
                             void* stream) {
  VS_REQUIRE(rows >= 0 && cols >= 1, "bad shape");
  if (rows == 0) return 0;
  VS_REQUIRE(x && hi && lo, "null pointer");
  if (cols % 4 != 0 || (uintptr_t)x % 16 != 0) return vidseg_split_f16(x, hi, lo, rows * cols, scale, stream);
  const int packed8 = operand_packed8(cols) ? 1 : 0;
  const long long items = rows * (cols / 4);
  int grid = (int)std::min<long long>((items + 255) / 256, (long long)kNumSMs * 8);
  VS_LAUNCH_W(8.0 * rows * cols, split_rows_kernel, grid, 256, 0, stream, x, (__half*)hi, (__half*)lo, rows, cols, scale,
              packed8, is_weight ? tc::kWgt8Sx : tc::kAct8Sx, is_weight ? tc::kWgt8Sl : tc::kAct8Sl);
  VS_POST_LAUNCH();
  return 0;
}

#undef VS_FAMILY
#define VS_FAMILY vidseg::kFamGemm
static int gemm_entry(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                      const float* residual, const float* row_bias, long long rows_per_bias, const float* blend,
                      const float* blend_alpha, long long rows_per_alpha, const float* row_scalar, float* out_f32,
                      void* out_hi, void* out_lo, int out_pair16, int m, int n, int k, float acc_scale, void* stream) {
  VS_REQUIRE(a_hi && a_lo && w_hi && w_lo, "null operand pointer");
  VS_REQUIRE(out_f32 != nullptr || (out_hi != nullptr && out_lo != nullptr), "no output requested");
  VS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "out_hi and out_lo go together");
  VS_REQUIRE(m >= 0 && n >= 1 && k >= 1, "bad shape");
  VS_REQUIRE(n % 4 == 0 && k % 8 == 0, "N must be a multiple of 4 and K of 8 (16-byte rows for TMA and vector stores)");
  VS_REQUIRE(out_hi == nullptr || n % 8 == 0, "split output needs N % 8 == 0");
  VS_REQUIRE(row_bias == nullptr || rows_per_bias >= 1, "rows_per_bias must be positive");
  VS_REQUIRE((blend == nullptr) == (blend_alpha == nullptr), "blend and blend_alpha go together");
  VS_REQUIRE(blend == nullptr || rows_per_alpha >= 1, "rows_per_alpha must be positive");
  if (m == 0) return 0;
  GemmParams p{};
  p.n = n; p.k = k; p.taps = 1; p.cin = k;
  p.wo = m; p.ho = 1; p.nb = 1;
  p.acc_scale = acc_scale;
  p.bias = bias; p.residual = residual; p.out_f32 = out_f32;
  p.chan_bias = row_bias; p.cb_div = rows_per_bias;
  p.blend = blend; p.blend_alpha = blend_alpha; p.ba_div = rows_per_alpha;
  p.row_scalar = row_scalar;
  p.out_hi = (__half*)out_hi; p.out_lo = (__half*)out_lo;
  p.in_packed8 = -1;
  p.out_packed8 = out_pair16 ? 0 : -1;
  const uint64_t adims[5] = {(uint64_t)k, (uint64_t)m, 1, 1, 1};
  const uint64_t row = (uint64_t)k * 2;
  const uint64_t astrides[4] = {row, row * m, row * m, row * m};
  return run_gemm(a_hi, a_lo, adims, astrides, p, w_hi, w_lo, kFamGemm, stream);
}

VS_API int vidseg_gemm_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                             const float* residual, float* out_f32, void* out_hi, void* out_lo, int m, int n, int k,
                             float acc_scale, void* stream) {
  return gemm_entry(a_hi, a_lo, w_hi, w_lo, bias, residual, nullptr, 1, nullptr, nullptr, 1, nullptr, out_f32, out_hi, out_lo,
                    0, m, n, k, acc_scale, stream);
}

VS_API int vidseg_gemm_geglu_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                                   void* out_hi, void* out_lo, int m, int d, int k, float acc_scale, void* stream) {
  VS_REQUIRE(a_hi && a_lo && w_hi && w_lo && out_hi && out_lo, "null pointer");
  VS_REQUIRE(m >= 0 && d >= 32 && d % 32 == 0 && k >= 8 && k % 8 == 0, "D must be a multiple of 32, K of 8");
  if (m == 0) return 0;
  GemmParams p{};
  p.n = 2 * d; p.k = k; p.taps = 1; p.cin = k;
  p.wo = m; p.ho = 1; p.nb = 1;
  p.acc_scale = acc_scale;
  p.bias = bias;
  p.geglu = 1;
  p.out_hi = (__half*)out_hi; p.out_lo = (__half*)out_lo;
  p.in_packed8 = -1; p.out_packed8 = -1;
  const uint64_t adims[5] = {(uint64_t)k, (uint64_t)m, 1, 1, 1};
  const uint64_t row = (uint64_t)k * 2;
  const uint64_t astrides[4] = {row, row * m, row * m, row * m};
  return run_gemm(a_hi, a_lo, adims, astrides, p, w_hi, w_lo, kFamGemm, stream);
}

VS_API int vidseg_gemm_split_ex(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                                const float* residual, const float* row_bias, long long rows_per_bias, const float* blend,
                                const float* blend_alpha, long long rows_per_alpha, const float* row_scalar,
                                float* out_f32, void* out_hi, void* out_lo, int out_pair16, int m, int n, int k,
                                float acc_scale, void* stream) {
  return gemm_entry(a_hi, a_lo, w_hi, w_lo, bias, residual, row_bias, rows_per_bias, blend, blend_alpha, rows_per_alpha,
                    row_scalar, out_f32, out_hi, out_lo, out_pair16, m, n, k, acc_scale, stream);
}

// Several projections of the SAME activation as one GEMM (to_q | to_k | to_v of a self-attention layer, attention.py:308-317):
// w = the weights stacked along N, [nseg * seg_n, K]; segment s is written to out_f32[s] (fp32 [M, seg_n], may be null) and
// / or out_hi[s] / out_lo[s] (operand [M, seg_n]).  The A operand is read once instead of nseg times.
VS_API int vidseg_gemm_split_seg(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int nseg, int seg_n,
                                 float* const* out_f32, void* const* out_hi, void* const* out_lo, int out_pair16, int m, int k,
                                 float acc_scale, void* stream) {
  VS_REQUIRE(a_hi && a_lo && w_hi && w_lo && out_f32 && out_hi && out_lo, "null pointer");
  VS_REQUIRE(nseg >= 1 && nseg <= 3, "1 to 3 segments");
  VS_REQUIRE(m >= 0 && k >= 8 && k % 8 == 0 && seg_n >= 128 && (seg_n % 128 == 0 || seg_n % 160 == 0),
             "K must be a multiple of 8, the segment width of 128 or 160");
  if (m == 0) return 0;
  GemmParams p{};
  p.n = nseg * seg_n; p.k = k; p.taps = 1; p.cin = k;
  p.wo = m; p.ho = 1; p.nb = 1;
  p.acc_scale = acc_scale;
  p.seg_n = seg_n;
  for (int s = 0; s < nseg; ++s) {
    VS_REQUIRE((out_hi[s] == nullptr) == (out_lo[s] == nullptr), "out_hi and out_lo go together");
    VS_REQUIRE(out_f32[s] != nullptr || out_hi[s] != nullptr, "a segment without output");
    p.seg_f32[s] = out_f32[s]; p.seg_hi[s] = (__half*)out_hi[s]; p.seg_lo[s] = (__half*)out_lo[s];
  }
  p.in_packed8 = -1;
  p.out_packed8 = out_pair16 ? 0 : (operand_packed8(seg_n) ? 1 : 0);
  const uint64_t adims[5] = {(uint64_t)k, (uint64_t)m, 1, 1, 1};
  const uint64_t row = (uint64_t)k * 2;
  const uint64_t astrides[4] = {row, row * m, row * m, row * m};
  return run_gemm(a_hi, a_lo, adims, astrides, p, w_hi, w_lo, kFamGemm, stream);
}

#undef VS_FAMILY
#define VS_FAMILY vidseg::kFamConv
static int conv2d_entry(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                        const float* chan_bias, const float* residual, float* out_f32, void* out_hi, void* out_lo,
                        int batch, int height, int width, int cin, int cout, int ksize, int stride, int pad_before,
                        float acc_scale, void* stream) {
  VS_REQUIRE(x_hi && x_lo && w_hi && w_lo, "null operand pointer");
  VS_REQUIRE(out_f32 != nullptr || (out_hi != nullptr && out_lo != nullptr), "no output requested");
  VS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "out_hi and out_lo go together");
  VS_REQUIRE(batch >= 0 && height >= 1 && width >= 1 && cin >= 1 && cout >= 1, "bad shape");
  VS_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "kernel 1 or 3, stride 1 or 2");
  VS_REQUIRE(ksize == 3 || stride == 1, "1x1 convolutions are stride 1");
  VS_REQUIRE(cin % 8 == 0 && cout % 4 == 0, "Cin must be a multiple of 8 and Cout of 4");
  VS_REQUIRE(out_hi == nullptr || cout % 8 == 0, "split output needs Cout % 8 == 0");
  if (batch == 0) return 0;
  GemmParams p{};
  p.taps = ksize * ksize;
  p.cin = cin;
  p.n = cout; p.k = p.taps * cin;
  p.bias = bias; p.chan_bias = chan_bias; p.residual = residual; p.out_f32 = out_f32;
  p.out_hi = (__half*)out_hi; p.out_lo = (__half*)out_lo;
  p.nb = batch;
  p.acc_scale = acc_scale;
  p.in_packed8 = -1; p.out_packed8 = -1;
  uint64_t adims[5], astrides[4];
  const uint64_t px = (uint64_t)cin * 2;  // bytes per pixel
  if (stride == 1) {
    p.wo = width; p.ho = height;
    adims[0] = cin; adims[1] = width; adims[2] = 1; adims[3] = height; adims[4] = batch;
    astrides[0] = px; astrides[1] = px * width; astrides[2] = px * width; astrides[3] = px * width * height;
    for (int t = 0; t < p.taps; ++t) {
      const int dy = (ksize == 3) ? t / 3 - 1 : 0, dx = (ksize == 3) ? t % 3 - 1 : 0;
      p.c_off[t] = 0; p.w_off[t] = dx; p.p_idx[t] = 0; p.h_off[t] = dy;
    }
  } else {
    VS_REQUIRE(height % 2 == 0 && width % 2 == 0, "stride-2 convolution needs even H and W");
    VS_REQUIRE(cin % 64 == 0, "stride-2 convolution needs Cin % 64 == 0 (a 64-channel chunk must not straddle pixels)");
    p.wo = width / 2; p.ho = height / 2;
    // [B, H, W, C] viewed as (2C, W/2, 2, H/2, B): input row 2*ho + dy - 1 -> (parity, index) = dy==1 ? (0, ho) : (1, ho + (dy-1)/2 ...)
    adims[0] = 2 * (uint64_t)cin; adims[1] = width / 2; adims[2] = 2; adims[3] = height / 2; adims[4] = batch;
    astrides[0] = 2 * px; astrides[1] = px * width; astrides[2] = 2 * px * width; astrides[3] = px * width * height;
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3, dx = t % 3;
      if (pad_before) {   // input row = 2*ho + dy - 1, column = 2*wo + dx - 1 (padding 1 on every side)
        p.p_idx[t] = (dy == 1) ? 0 : 1;
        p.h_off[t] = (dy == 0) ? -1 : 0;
        p.c_off[t] = (dx == 1) ? 0 : cin;
        p.w_off[t] = (dx == 0) ? -1 : 0;
      } else {            // input row = 2*ho + dy, column = 2*wo + dx (zero row / column AFTER the image only)
        p.p_idx[t] = (dy == 1) ? 1 : 0;
        p.h_off[t] = (dy == 2) ? 1 : 0;
        p.c_off[t] = (dx == 1) ? cin : 0;
        p.w_off[t] = (dx == 2) ? 1 : 0;
      }
    }
  }
  return run_gemm(x_hi, x_lo, adims, astrides, p, w_hi, w_lo, kFamConv, stream);
}

VS_API int vidseg_conv2d_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                               const float* chan_bias, const float* residual, float* out_f32, void* out_hi, void* out_lo,
                               int batch, int height, int width, int cin, int cout, int ksize, int stride, float acc_scale,
                               void* stream) {
  return conv2d_entry(x_hi, x_lo, w_hi, w_lo, bias, chan_bias, residual, out_f32, out_hi, out_lo, batch, height, width, cin,
                      cout, ksize, stride, 1, acc_scale, stream);
}

// 3x3 stride-2 convolution of F.pad(x, (0, 1, 0, 1)) with padding 0: the Downsample of the first-stage encoder
// (sgm/modules/diffusionmodules/model.py:77-94)
VS_API int vidseg_conv2d_down_pad_after_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                                              const float* bias, float* out_f32, void* out_hi, void* out_lo, int batch,
                                              int height, int width, int cin, int cout, float acc_scale, void* stream) {
  return conv2d_entry(x_hi, x_lo, w_hi, w_lo, bias, nullptr, nullptr, out_f32, out_hi, out_lo, batch, height, width, cin,
                      cout, 3, 2, 0, acc_scale, stream);
}


// (3,1,1) convolution over the frame axis of a video tensor [V, T, HW, C] (channels-last '(b t) h w c' memory):
// the three taps are the same pixel patch shifted by -1 / 0 / +1 frames, out-of-range frames are the TMA zero fill.
VS_API int vidseg_conv_temporal_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                                      const float* bias, const float* frame_bias, const float* residual,
                                      const float* blend, const float* blend_alpha, float* out_f32, void* out_hi,
                                      void* out_lo, int videos, int frames, int hw, int cin, int cout, float acc_scale,
                                      void* stream) {
  VS_REQUIRE(x_hi && x_lo && w_hi && w_lo, "null operand pointer");
  VS_REQUIRE(out_f32 != nullptr || (out_hi != nullptr && out_lo != nullptr), "no output requested");
  VS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "out_hi and out_lo go together");
  VS_REQUIRE(videos >= 0 && frames >= 1 && hw >= 1 && cin >= 1 && cout >= 1, "bad shape");
  VS_REQUIRE(cin % 8 == 0 && cout % 4 == 0, "Cin must be a multiple of 8 and Cout of 4");
  VS_REQUIRE(out_hi == nullptr || cout % 8 == 0, "split output needs Cout % 8 == 0");
  VS_REQUIRE((blend == nullptr) == (blend_alpha == nullptr), "blend and blend_alpha go together");
  if (videos == 0) return 0;
  GemmParams p{};
  p.taps = 3;
  p.cin = cin;
  p.n = cout; p.k = 3 * cin;
  p.bias = bias; p.residual = residual; p.out_f32 = out_f32;
  p.chan_bias = frame_bias; p.cb_div = hw;           // [V*T, Cout]: one row per frame
  p.blend = blend; p.blend_alpha = blend_alpha; p.ba_div = hw;  // alpha[V*T]
  p.out_hi = (__half*)out_hi; p.out_lo = (__half*)out_lo;
  p.nb = videos; p.ho = frames; p.wo = hw;
  p.acc_scale = acc_scale;
  p.in_packed8 = -1; p.out_packed8 = -1;
  const uint64_t px = (uint64_t)cin * 2;
  const uint64_t adims[5] = {(uint64_t)cin, (uint64_t)hw, 1, (uint64_t)frames, (uint64_t)videos};
  const uint64_t astrides[4] = {px, px * hw, px * hw, px * hw * frames};
  for (int t = 0; t < 3; ++t) { p.c_off[t] = 0; p.w_off[t] = 0; p.p_idx[t] = 0; p.h_off[t] = t - 1; }
  return run_gemm(x_hi, x_lo, adims, astrides, p, w_hi, w_lo, kFamConv, stream);
}
