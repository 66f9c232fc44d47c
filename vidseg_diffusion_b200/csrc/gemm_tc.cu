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
constexpr int kKmMaxCols = 256;                                // R*K of the fused K-means E-step epilogue (one N tile)
constexpr int kKmSmemBytes = kKmMaxCols * 16 + 64 * 4 + 32;   // per-column {cn, tau, meta}, changed[R], per-warp maxima

template <int BN>
struct GemmCfg {
  static constexpr int kTileBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;
  static constexpr int kStages = (BN <= 128) ? 3 : (BN <= 160 ? 3 : 2);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kKmSmemBytes;
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
  KmEpilogue km;            // K-means E-step epilogue (km.on)
};

constexpr int kGemmThreads = 256;   // TMA warp, MMA warp, TMEM-alloc warp, spare, 4 epilogue warps (8 measured no faster)
// CL = 2: two CTAs of a thread-block cluster work on two adjacent M tiles of the SAME N tile; each loads half of the
// weight (B) tile and multicasts it into both CTAs' shared memory, so the L2 -> SM operand traffic per k-block drops
// from A + B to A + B/2 (the packed8 GEMMs are bound by that traffic, not by the MMA pipe: ncu 69 % tensor-active).
// A stage may only be refilled when BOTH CTAs have consumed it: the MMA issuer's commit arrives on both empty barriers.
// State of one epilogue thread (= one point) while it scans the score columns of one K-means run.
struct KmScan {
  float best, best_tau, minlow;
  int best_j, old_label;
};

// second pass of the K-means epilogue over one run's k score columns (warp-collective TMEM loads, 8 columns at a time):
// the centres whose lower bound lies inside the final best's band.  Not inlined; compact on purpose -- the epilogue is
// executed once per tile by one warp per scheduler, so it runs at the speed its code can be fetched.
__device__ __noinline__ unsigned long long km_candidates(uint32_t taddr, int k, const float4* __restrict__ tab, float m2inv,
                                                         float xn, float thr) {
  unsigned long long cand = 0ull;
#pragma unroll 1
  for (int q0 = 0; q0 < k; q0 += 8) {
    uint32_t r2[8];
    tc::tmem_ld_32x8(taddr + q0, r2);   // may run past the run's last column: masked below
    tc::tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 e2 = tab[q0 + q];
      const float sc2 = fmaf(__uint_as_float(r2[q]), m2inv, e2.x);
      if (q0 + q < k && sc2 - e2.y * xn <= thr) cand |= 1ull << ((q0 + q) & 63);
    }
  }
  return cand;
}

// end of a run for one point: take the label, or hand the pair to the resolver with its candidate set
__device__ __noinline__ void km_run_end(const KmEpilogue& km, const KmScan st, int meta, int col, bool row_ok, int row,
                                        float xn, float slack, float m2inv, uint32_t t_acc,
                                        const float4* __restrict__ s_tab, int* __restrict__ s_changed, int lane) {
  const int run = meta >> 16;
  const bool take = row_ok && (meta & 0x200);
  const float thr = st.best + st.best_tau * xn + slack;
  const bool amb = take && (st.minlow <= thr);
  bool moved = false;
  if (take && !amb) {
    moved = km.count_changes && st.old_label != st.best_j;
    km.labels[(size_t)run * km.labels_stride + row] = st.best_j;
  }
  const unsigned mv = __ballot_sync(0xffffffffu, moved);   // one shared atomic per warp and run
  if (lane == 0 && mv) atomicAdd(&s_changed[run], __popc(mv));
  const unsigned am = __ballot_sync(0xffffffffu, amb);
  if (am) {
    const int col0 = col - (km.k - 1);
    unsigned long long cand = km_candidates(t_acc + col0, km.k, s_tab + col0, m2inv, xn, thr);
    int base_slot = 0;   // one global atomic per warp and run: the counter is a single address for the whole grid
    if (lane == 0) base_slot = atomicAdd(km.amb_count, __popc(am));
    base_slot = __shfl_sync(0xffffffffu, base_slot, 0);
    if (amb) {
      if (km.k > 64) cand = ~0ull;
      const int slot = base_slot + __popc(am & ((1u << lane) - 1u));
      km.amb_list[slot] = make_int4(row, run, (int)(unsigned)cand, (int)(unsigned)(cand >> 32));
    }
  }
}

template <int BN, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
                  const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
                  const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + kGemmAccStages;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + kGemmAccStages);
  // same bytes as smem + ..., derived from the __shared__ array itself so that the compiler keeps the shared address
  // space (LDS / ATOMS instead of generic loads and atomics) for the K-means epilogue tables
  uint8_t* km_smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u) + kStages * Cfg::kStageBytes + 256;

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
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], CL); }
    for (int s = 0; s < kGemmAccStages; ++s) { tc::mbar_init(&tmem_full_bar[s], 1); tc::mbar_init(&tmem_empty_bar[s], kGemmThreads - 128); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<Cfg::kTmemCols>(tmem_base_ptr);
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
            tc::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            const int ca = p.c_off[tap] + kc * kGemmBK;
            const int kb0 = tap * p.cin + kc * kGemmBK;
            tc::tma_load_5d(st, &tmap_a_hi, &full_bar[stage], ca, cw, cp, ch, b0);
            tc::tma_load_5d(st + kTileABytes, &tmap_a_lo, &full_bar[stage], ca, cw, cp, ch, b0);
            if (CL > 1) {   // this CTA's half of the B rows, delivered to both CTAs of the pair
              constexpr int kHalfRows = BN / 2, kHalfBytes = Cfg::kTileBBytes / 2;
              tc::tma_load_2d_mc(st + 2 * kTileABytes + crank * kHalfBytes, &tmap_b_hi, &full_bar[stage], kb0,
                                 n0 + crank * kHalfRows, (uint16_t)0x3);
              tc::tma_load_2d_mc(st + 2 * kTileABytes + Cfg::kTileBBytes + crank * kHalfBytes, &tmap_b_lo, &full_bar[stage],
                                 kb0, n0 + crank * kHalfRows, (uint16_t)0x3);
            } else {
              tc::tma_load_2d(st + 2 * kTileABytes, &tmap_b_hi, &full_bar[stage], kb0, n0);
              tc::tma_load_2d(st + 2 * kTileABytes + Cfg::kTileBBytes, &tmap_b_lo, &full_bar[stage], kb0, n0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_f16(kGemmBM, p.mma_n > 0 ? p.mma_n : BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        tc::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
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
            constexpr uint32_t idesc_lx = tc::make_idesc_f8(kGemmBM, BN, 1, 0);  // A e5m2 residual x B e4m3 value
            constexpr uint32_t idesc_xl = tc::make_idesc_f8(kGemmBM, BN, 0, 1);  // A e4m3 value    x B e5m2 residual
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t adv = (uint64_t)(h * 32 >> 4), xoff = (uint64_t)(64 >> 4);
              tc::umma_f8(d_acc, a_lo + adv, b_lo + xoff + adv, idesc_lx, (kb > 0 || h > 0) ? 1u : 0u);
              tc::umma_f8(d_acc, a_lo + xoff + adv, b_lo + adv, idesc_xl, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < kGemmBK / 16; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 32 >> 4);
              tc::umma_f16(d_acc, a_hi + adv, b_hi + adv, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < kGemmBK / 16; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 32 >> 4);  // 16 elements = 32 bytes along K inside the swizzle atom
              // small terms first, so that they are not absorbed one by one into a large partial sum
              tc::umma_f16(d_acc, a_lo + adv, b_hi + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
              tc::umma_f16(d_acc, a_hi + adv, b_lo + adv, idesc, 1u);
              tc::umma_f16(d_acc, a_hi + adv, b_hi + adv, idesc, 1u);
            }
          }
          // frees the smem stage once these MMAs have read it (in both CTAs of a pair: either may overwrite it)
          if (CL > 1) tc::umma_commit_mc(&empty_bar[stage], (uint16_t)0x3); else tc::umma_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) tc::umma_commit(&tmem_full_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // epilogue warps: a TMEM lane quarter (warp id % 4) may be shared by two warps that split the tile's columns
    // (kGemmThreads = 384); with four warps each one walks all columns
    const int ew = (warp - 4) & 3;     // TMEM lanes [32*ew, 32*ew+32)
    const int chalf = (warp - 4) >> 2; // which half of the 32-column chunks
    constexpr int kChunks = (BN + 31) / 32, kChunks0 = (kGemmThreads > 256) ? (kChunks + 1) / 2 : kChunks;
    const int c_begin = chalf ? kChunks0 * 32 : 0, c_end = chalf ? BN : kChunks0 * 32;
    const int r = ew * 32 + lane;  // row of the tile = pixel of the patch
    const int pw = r % p.bw, ph = (r / p.bw) % p.bh, pb = r / (p.bw * p.bh);
    int acc = 0;
    uint32_t acc_phase = 0;
    if (BN == 256 && p.km.on) {
      // ===================== K-means E-step epilogue =====================
      // labels = argmin_j ( ||c_j||^2 - 2 x.c_j ) per run, decided from the tensor-core scores when the best centre is
      // separated from every other by more than the filter's error band, deferred to the resolver otherwise.
      // Pass 1, per run, branch-free per column: `best` is the running minimum, `minlow` the smallest lower bound
      // sc_j - tau_j |x| among the OTHER centres; the point is ambiguous iff minlow <= best + tau_best |x| (+ slack), i.e.
      // iff some other centre satisfies sc_j <= best + (tau_best + tau_j) |x|.  Pass 2, only for runs in which some
      // lane of the warp is ambiguous: the run's columns are loaded from TMEM once more and every centre inside the band
      // of the final best goes into the pair's candidate mask for the resolver.
      // The test runs in fp32 (the scan is one dependent compare / select chain per point, issued by a single warp per
      // scheduler: fp32 halves its latency): the radii are widened by 1 % and `slack` bounds the fp32 rounding of two scores, so the
      // fp32 test can only flag MORE pairs / candidates than the float64 one -- every flagged pair is settled exactly by
      // the resolver, every unflagged label is the float64 arg-min.
      const KmEpilogue& km = p.km;
      float4* s_tab = reinterpret_cast<float4*>(km_smem);                  // per column {cn, tau, meta, -}
      int* s_changed = reinterpret_cast<int*>(s_tab + kKmMaxCols);
      float* s_max = reinterpret_cast<float*>(s_changed + 64);             // per epilogue warp {max cn, max tau}
      const int et = threadIdx.x - 128;          // 0..127 among the epilogue threads
      const int rk = km.runs * km.k;
      float mc = 0.f, mtau = 0.f;
      for (int i = et; i < kKmMaxCols; i += 128) {
        float4 e = make_float4(3.0e38f, 0.f, 0.f, 0.f);   // padding columns: never the best, never a candidate
        if (i < rk) {
          const double c = km.cnorm[i];
          const int run = i / km.k, j = i - run * km.k;
          const int done = km.flags[run * 4 + 0], strict = km.flags[run * 4 + 1];
          const int live = (km.only_nonstrict ? (strict == 0) : (done == 0)) ? 1 : 0;
          e.x = (float)c;
          e.y = (float)(1.01 * km.band * sqrt(c));
          e.z = __int_as_float(j | ((j == km.k - 1) ? 0x100 : 0) | (live << 9) | ((j == 0) ? 0x400 : 0) | (run << 16));
          mc = fmaxf(mc, e.x); mtau = fmaxf(mtau, e.y);
        }
        s_tab[i] = e;
      }
      for (int i = et; i < 64; i += 128) s_changed[i] = 0;
      mc = warp_max(mc); mtau = warp_max(mtau);
      if (lane == 0) { s_max[2 * ew] = mc; s_max[2 * ew + 1] = mtau; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float cn_max = fmaxf(fmaxf(s_max[0], s_max[2]), fmaxf(s_max[4], s_max[6]));
      const float sc_f = km_operand_scale(km.absmax[0]);
      const float m2inv = -2.0f / (sc_f * sc_f);   // a power of two: exact
      for (int tile = first_item; tile < num_tiles; tile += item_stride) {
        const int mt = m_tile_of(tile);
        const int rel = mt * 128 + r;            // Linear form: the tile is 128 consecutive rows
        const bool row_ok = rel < p.wo;
        const int row = km.row_begin + rel;
        const float xn = row_ok ? (float)sqrt(km.xx[row]) * 1.0000002f : 0.f;   // rounded up: radii only grow
        // |fp32 score - exact score| <= 2^-23 (cn + |x||c|) per centre; two scores meet in every comparison
        const float slack = 0x1p-21f * (cn_max + xn * sqrtf(cn_max));
        tc::mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc::tc_fence_after();
        const uint32_t t_acc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * Cfg::kAccStride);
        KmScan st{3.0e38f, 0.f, 3.0e38f, 0, -1};
        // 8 score columns per TMEM load, the next load in flight while the current eight are scanned (rolled loop:
        // ~100 instructions of body instead of 32 unrolled columns)
        uint32_t cur[8], nxt[8];
        {
          tc::tmem_ld_32x8(t_acc, cur);
          tc::tmem_wait_ld();
#pragma unroll 1
          for (int c = 0; c < rk; c += 8) {
            if (c + 8 < rk) tc::tmem_ld_32x8(t_acc + c + 8, nxt);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float4 e = s_tab[c + u];
              const int meta = __float_as_int(e.z);
              if ((meta & 0x600) == 0x600 && row_ok && km.count_changes)   // first column of a live run: its old label is
                st.old_label = km.labels[(size_t)(meta >> 16) * km.labels_stride + row];   // on its way during the scan
              const float sc = fmaf(__uint_as_float(cur[u]), m2inv, e.x);
              const float low = sc - e.y * xn;
              const bool nb = sc < st.best;
              st.minlow = fminf(st.minlow, nb ? st.best - st.best_tau * xn : low);   // 3e38 on a run's first column
              st.best_tau = nb ? e.y : st.best_tau;
              st.best_j = nb ? (meta & 0xff) : st.best_j;
              st.best = nb ? sc : st.best;
              if (meta & 0x100) {   // last column of a run (warp-uniform): decide, then start the next run
                km_run_end(km, st, meta, c + u, row_ok, row, xn, slack, m2inv, t_acc, s_tab, s_changed, lane);
                st.best = 3.0e38f; st.best_tau = 0.f; st.minlow = 3.0e38f; st.best_j = 0;
              }
            }
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < km.runs; i += 128)
        if (s_changed[i]) atomicAdd(&km.changed[i], s_changed[i]);
    } else
    for (int tile = first_item; tile < num_tiles; tile += item_stride) {
      const int mt = m_tile_of(tile);
      const int n0 = (tile % n_tiles) * BN;
      const int w = (mt % p.tiles_w) * p.bw + pw;
      const int h = ((mt / p.tiles_w) % p.tiles_h) * p.bh + ph;
      const int b = (mt / (p.tiles_w * p.tiles_h)) * p.bb + pb;
      const bool row_ok = (w < p.wo) && (h < p.ho) && (b < p.nb);
      const size_t pix = ((size_t)b * p.ho + h) * p.wo + w;
      const size_t cb_row = p.chan_bias ? (size_t)((long long)pix / p.cb_div) : 0;
      const float blend_a = (p.blend && row_ok) ? __ldg(p.blend_alpha + (long long)pix / p.ba_div) : 0.f;
      const float row_add = (p.row_scalar && row_ok) ? __ldg(p.row_scalar + pix) : 0.f;
      // The residual (and blend) rows of this tile come from HBM: start them towards L2 now, while the MMAs of the tile
      // are still running, and keep the loads of chunk c+1 in flight while chunk c is processed (measured before: the
      // epilogue of the K=320 projections sat on these loads, 29 % of DRAM bandwidth)
      const int tile_cols = min(c_end, p.n - n0);
      if (row_ok) {
        if (p.residual)
          for (int j = c_begin; j < tile_cols; j += 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + pix * p.n + n0 + j));
        if (p.blend)
          for (int j = c_begin; j < tile_cols; j += 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.blend + pix * p.n + n0 + j));
      }
      tc::mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * Cfg::kAccStride);
      if (p.geglu) {
        // FeedForward's GEGLU (attention.py:89-96) fused into the projection GEMM: the weight rows were permuted on the
        // host so that each 64-column group holds 32 value columns followed by their 32 gate columns; the fp32
        // [M, 2D] intermediate (the largest tensor of the UNet) is never written.
        const int dn = p.n >> 1;
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          const int col0 = n0 + c;
          if (col0 >= p.n) break;  // warp-uniform
          uint32_t rv[32], rg[32];
          tc::tmem_ld_32x32(t_acc + c, rv);
          tc::tmem_ld_32x32(t_acc + c + 32, rg);
          tc::tmem_wait_ld();
          if (row_ok) {
            float o[32];
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
                o[j + u] = val * (0.5f * gate * (1.0f + erff(gate * 0.70710678118654752440f)));
              }
            }
            __half* hrow = p.out_hi + pix * dn;
            __half* lrow = p.out_lo + pix * dn;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              tc::store_split8(hrow, lrow, (col0 >> 1) + j, &o[j], p.out_packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      float4 res_cur[8], res_nxt[8];
      auto load_res = [&](float4 (&dst)[8], int c0) {
        const int nc = min(32, p.n - c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = (p.residual && row_ok && 4 * j < nc) ? *reinterpret_cast<const float4*>(p.residual + pix * p.n + c0 + 4 * j)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      load_res(res_cur, n0 + c_begin);
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        const int col0 = n0 + c;
        if (col0 >= p.n) break;  // warp-uniform
        uint32_t rr[32];
        tc::tmem_ld_32x32(t_acc + c, rr);
        if (c + 32 < c_end && col0 + 32 < p.n) load_res(res_nxt, col0 + 32);
        tc::tmem_wait_ld();
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) * p.acc_scale;
          const int ncols = min(32, p.n - col0);  // N % 4 == 0 is required by the host wrapper
          const size_t off = pix * p.n + col0;
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncols) {
                const float4 bb4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                v[j] += bb4.x; v[j + 1] += bb4.y; v[j + 2] += bb4.z; v[j + 3] += bb4.w;
              }
          }
          if (p.chan_bias) {
            const float* cb = p.chan_bias + cb_row * p.n + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncols) {
                const float4 bb4 = __ldg(reinterpret_cast<const float4*>(cb + j));
                v[j] += bb4.x; v[j + 1] += bb4.y; v[j + 2] += bb4.z; v[j + 3] += bb4.w;
              }
          }
          if (p.row_scalar) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += row_add;
          }
          if (p.residual) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 q = res_cur[j >> 2];   // zero beyond ncols
              v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
            }
          }
          if (p.blend) {
            const float a = blend_a, na = 1.0f - a;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncols) {
                const float4 q = *reinterpret_cast<const float4*>(p.blend + off + j);
                v[j] = a * q.x + na * v[j]; v[j + 1] = a * q.y + na * v[j + 1];
                v[j + 2] = a * q.z + na * v[j + 2]; v[j + 3] = a * q.w + na * v[j + 3];
              }
          }
          if (p.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncols) *reinterpret_cast<float4*>(p.out_f32 + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (p.out_hi) {
            __half* hrow = p.out_hi + pix * p.n;
            __half* lrow = p.out_lo + pix * p.n;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (j < ncols) tc::store_split8(hrow, lrow, col0 + j, &v[j], p.out_packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) res_cur[j] = res_nxt[j];
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == kGemmAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL > 1) tc::cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
  if (warp == 2) tc::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
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
  using Cfg = GemmCfg<BN>;
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
  if (CL == 1) {
    gemm_split_kernel<BN, CL><<<grid, kGemmThreads, Cfg::kSmemBytes, (cudaStream_t)stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_split_kernel<BN, CL>, ta_hi, ta_lo, tb_hi, tb_lo, p));
  }
  if (prof) profile_after((cudaStream_t)stream);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  VS_POST_LAUNCH();
  return 0;
}

// Measured on B200 (tools/prof_kernels.py, packed8): the 2-CTA multicast form is correct but only 0-3 % faster (3x3 conv
// 562 vs 556, [28672,5120,640] 455 vs 442 algorithmic TFLOP/s) -- the packed8 GEMMs are bound by the bytes that must be
// in flight INTO shared memory per k-block (72-96 KB every ~0.34 us against ~1 us of TMA latency with 216 KB of smem),
// which multicast does not reduce.  It therefore stays opt-in (VIDSEG_GEMM_CLUSTER=1); the cure is the 2-SM MMA
// (cta_group::2, half of B per SM), a next-round item.
static bool use_cluster(const GemmParams& p, int bn) {
  static const int env = [] { const char* e = getenv("VIDSEG_GEMM_CLUSTER"); return e ? atoi(e) : 0; }();
  if (!env) return false;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b, n_tiles = (p.n + bn - 1) / bn;
  return m_tiles >= 2 && (long long)m_tiles * n_tiles >= 2LL * kNumSMs;
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

// internal entry for kmeans.cu: the score GEMM [m, d] x [d, R*K] of the Lloyd E-step with the arg-min / ambiguity test as
// its epilogue (KmEpilogue); one 256-wide N tile holds the scores of every run, so R*K <= 256.
int gemm_km_estep_run(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int m, int rk_pad, int k,
                      const KmEpilogue& km, void* stream) {
  VS_REQUIRE(a_hi && a_lo && w_hi && w_lo, "null pointer");
  VS_REQUIRE(m >= 1 && rk_pad >= 8 && rk_pad <= kKmMaxCols && rk_pad % 8 == 0 && k >= 8 && k % 8 == 0, "bad shape");
  VS_REQUIRE(km.runs * km.k <= rk_pad && km.runs <= 64, "runs * k must fit the padded column count");
  GemmParams p{};
  p.n = rk_pad; p.k = k; p.taps = 1; p.cin = k;
  p.wo = m; p.ho = 1; p.nb = 1;
  p.acc_scale = 1.0f;
  p.in_packed8 = 0;   // fp16 pairs: the filter's error band is derived for them
  p.out_packed8 = 0;
  p.mma_n = (rk_pad + 15) / 16 * 16;
  p.km = km;
  p.km.on = 1;
  const uint64_t adims[5] = {(uint64_t)k, (uint64_t)m, 1, 1, 1};
  const uint64_t row = (uint64_t)k * 2;
  const uint64_t astrides[4] = {row, row * m, row * m, row * m};
  p.bw = 128; p.bh = 1; p.bb = 1;   // 128 consecutive rows per tile (what the epilogue assumes)
  p.tiles_w = (m + 127) / 128; p.tiles_h = 1; p.tiles_b = 1;
  p.kc_per_tap = (k + kGemmBK - 1) / kGemmBK;
  p.cb_div = 1; p.ba_div = 1;
  const uint32_t box[5] = {(uint32_t)kGemmBK, 128u, 1u, 1u, 1u};
  CUtensorMap ta_hi, ta_lo;
  if (int e = encode_tmap_16bit(&ta_hi, a_hi, 5, adims, astrides, box)) return e;
  if (int e = encode_tmap_16bit(&ta_lo, a_lo, 5, adims, astrides, box)) return e;
  const double flops = 2.0 * (double)m * rk_pad * (double)k;
  // pairs of CTAs share the centre tile through TMA multicast: every CTA would otherwise pull all R*K centre rows from
  // L2 for each of its k-blocks, and that traffic (not the MMAs) bounds this skinny GEMM
  static const int cl2 = [] { const char* e = getenv("VIDSEG_KM_CLUSTER"); return e ? atoi(e) : 0; }();
  if (cl2 && p.tiles_w >= 2) return launch_gemm_cl<256, 2>(ta_hi, ta_lo, w_hi, w_lo, p, flops, kFamKMeans, stream);
  return launch_gemm_cl<256, 1>(ta_hi, ta_lo, w_hi, w_lo, p, flops, kFamKMeans, stream);
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
