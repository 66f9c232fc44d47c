// Scaled-dot-product attention of the UNet transformer blocks on tcgen05 tensor cores.
//
// Replaces F.scaled_dot_product_attention at sgm/modules/attention.py:352-356 (and the xformers
// call :473-485): out[b, i, h*64:(h+1)*64] = softmax(q_i . k_j / 8) v_j over the keys of (b, h), head dim 64,
// no mask.  q/k/v arrive pre-head-split exactly as the reference stashes them ([B, N, heads*64]), as
// split-fp16 pairs (see gemm_tc.cu); both GEMMs of the flash loop are 3-MMA split products with fp32
// accumulation in TMEM, the softmax runs in fp32 registers.
//
// One CTA = 256 queries of one (batch, head): two 128-row query tiles, each owned by a softmax
// warpgroup (128 threads = 128 TMEM lanes).  Per 64-key block and tile:
//   MMA warp : S = Q K^T   (12 x tcgen05.mma 128x128x16: lo.hi + hi.lo + hi.hi into one fp32 accumulator) for a block of
//              128 keys; S(j+1) of a tile is issued right behind PV(j) of the same tile, the other tile's softmax runs meanwhile
//   softmax  : tcgen05.ld S, row max / exp2 / row sum in registers, P -> fp16 back into the TMEM columns of S
//   MMA warp : O += P V    (2 x 8 tcgen05.mma 128x64x16, P from TMEM, V the MN-major B operand as an fp16 pair); O stays in TMEM across the whole
//              key loop, so the softmax group does not wait for this product either
//   rescale  : the running maximum is only raised when a block exceeds it by more than 2^3 (the probabilities are
//              carried with 2^12 headroom inside fp16); only then -- typically the first one or two blocks of a row --
//              the group loads O from TMEM, multiplies by 2^(m_old - m_new) and stores it back
// TMA (warp 0) streams K/V blocks through a 3-stage ring.  TMEM: 2 tiles x (S 2x64 + O 64) = 384 of 512 columns.
#include <mutex>

#define VS_FAMILY vidseg::kFamAttention
#include "common.cuh"
#include "tc_common.cuh"

namespace vidseg {

constexpr int kAtBQ = 128;        // queries per tile
constexpr int kAtBK = 128;        // keys per block (128-wide S tiles: the 64-wide ones were bound by operand fetch, 6 KB of smem per 32 math cycles)
constexpr int kAtD = 64;          // head dim
constexpr int kAtQTileBytes = kAtBQ * kAtD * 2;  // 16 KB (one of hi / lo)
constexpr int kAtKTileBytes = kAtBK * kAtD * 2;  // 16 KB
constexpr int kAtTileCols = 192;       // per tile: S (128 columns; P overwrites it in place) | O (64 columns)
// Two shapes of CTA.  <2, 2>: 256 queries (two tiles whose softmax / MMA phases interleave), K/V through a two-stage ring,
// all 512 TMEM columns, one CTA per SM -- the long key loops of self-attention.  <1, 1>: 128 queries, one K/V stage,
// 256 TMEM columns, 96 KB: TWO CTAs per SM -- cross-attention to the 77 context tokens is a single key block per CTA, i.e.
// a chain of latencies (Q / K / V load, S, softmax, P V, store) with nothing to overlap it inside the CTA; a second resident
// CTA does (the 4096 x 77 layer ran at 1.3 TB/s of its 294 MB, one CTA per SM).
template <int TILES, int STAGES>
struct AtCfg {
  static constexpr int kSmemQ = TILES * 2 * kAtQTileBytes;
  static constexpr int kSmemKV = STAGES * 4 * kAtKTileBytes;
  static constexpr int kSmemBytes = kSmemQ + kSmemKV + 1024 + 256;
  static constexpr int kThreads = 64 + TILES * 128;    // producer warp, MMA warp, one softmax warpgroup per tile
  static constexpr int kTmemCols = TILES == 1 ? 256 : 512;
};
constexpr float kAtTau = 3.0f;         // lazy-rescale threshold (log2 units): p <= 2^(12+3) stays inside fp16  // producer warp, MMA warp, 2 softmax warpgroups

// 2^x on the SFU, flush-to-zero: one MUFU.EX2 (exp2f() adds a denormal-range rescale: two FMUL and a compare per call)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  int batch, heads, nq, nk;
  float scale_log2;  // softmax scale * log2(e)
  float* out_f32;    // [B, Nq, heads*64] or null
  __half* out_hi;    // split output or null
  __half* out_lo;
  int out_packed8;   // format of the split output (it feeds the to_out GEMM)
};

template <int kAtTiles, int kAtStages>
__global__ void __launch_bounds__(AtCfg<kAtTiles, kAtStages>::kThreads, kAtTiles == 1 ? 2 : 1)
attn_split_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                  const __grid_constant__ CUtensorMap tm_k_hi, const __grid_constant__ CUtensorMap tm_k_lo,
                  const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                  const AttnParams p) {
  using Cfg = AtCfg<kAtTiles, kAtStages>;
  constexpr int kAtSmemQ = Cfg::kSmemQ, kAtSmemKV = Cfg::kSmemKV, kAtTmemCols = Cfg::kTmemCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sm_q = smem;                       // [tile][hi|lo][128x64]
  uint8_t* sm_kv = sm_q + kAtSmemQ;           // [stage][k_hi|k_lo|v_hi|v_lo][128x64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_kv + kAtSmemKV);
  uint64_t* q_full = bars;                                // 1
  uint64_t* kv_full = bars + 1;                           // [kAtStages]
  uint64_t* kv_empty = kv_full + kAtStages;               // [kAtStages]
  uint64_t* s_full = kv_empty + kAtStages;                // [tile][S buffer]
  uint64_t* p_full = s_full + 2 * kAtTiles;               // [tile]
  uint64_t* pv_done = p_full + kAtTiles;                  // [tile]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(pv_done + kAtTiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (kAtBQ * kAtTiles);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int nkb = (p.nk + kAtBK - 1) / kAtBK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tm_q_hi); tc::prefetch_tmap(&tm_q_lo);
    tc::prefetch_tmap(&tm_k_hi); tc::prefetch_tmap(&tm_k_lo);
    tc::prefetch_tmap(&tm_v_hi); tc::prefetch_tmap(&tm_v_lo);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < kAtStages; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    for (int g = 0; g < kAtTiles; ++g) {
      tc::mbar_init(&s_full[2 * g], 1); tc::mbar_init(&s_full[2 * g + 1], 1);
      tc::mbar_init(&p_full[g], 128); tc::mbar_init(&pv_done[g], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<kAtTmemCols>(tmem_base_ptr);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(q_full, kAtSmemQ);
      for (int g = 0; g < kAtTiles; ++g) {
        tc::tma_load_3d(sm_q + (g * 2 + 0) * kAtQTileBytes, &tm_q_hi, q_full, head * kAtD, q0 + g * kAtBQ, b);
        tc::tma_load_3d(sm_q + (g * 2 + 1) * kAtQTileBytes, &tm_q_lo, q_full, head * kAtD, q0 + g * kAtBQ, b);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkb; ++j) {
        tc::mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* st = sm_kv + stage * 4 * kAtKTileBytes;
        tc::mbar_arrive_expect_tx(&kv_full[stage], 4 * kAtKTileBytes);
        tc::tma_load_3d(st + 0 * kAtKTileBytes, &tm_k_hi, &kv_full[stage], head * kAtD, j * kAtBK, b);
        tc::tma_load_3d(st + 1 * kAtKTileBytes, &tm_k_lo, &kv_full[stage], head * kAtD, j * kAtBK, b);
        tc::tma_load_3d(st + 2 * kAtKTileBytes, &tm_v_hi, &kv_full[stage], head * kAtD, j * kAtBK, b);
        tc::tma_load_3d(st + 3 * kAtKTileBytes, &tm_v_lo, &kv_full[stage], head * kAtD, j * kAtBK, b);
        if (++stage == kAtStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = tc::make_idesc_f16(kAtBQ, kAtBK, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = tc::make_idesc_f16(kAtBQ, kAtD, 0, 1);    // P (K-major) x V (MN-major)
      auto issue_s = [&](int g, int stage) {
        const uint32_t qa = tc::smem_u32(sm_q + g * 2 * kAtQTileBytes);
        const uint32_t ka = tc::smem_u32(sm_kv + stage * 4 * kAtKTileBytes);
        const uint64_t q_hi = tc::make_sw128_desc(qa), q_lo = tc::make_sw128_desc(qa + kAtQTileBytes);
        const uint64_t k_hi = tc::make_sw128_desc(ka), k_lo = tc::make_sw128_desc(ka + kAtKTileBytes);
        const uint32_t d_s = tmem_base + (uint32_t)(g * kAtTileCols);
#pragma unroll
        for (int ks = 0; ks < kAtD / 16; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 32 >> 4);
          tc::umma_f16(d_s, q_lo + adv, k_hi + adv, idesc_s, ks > 0);
          tc::umma_f16(d_s, q_hi + adv, k_lo + adv, idesc_s, 1u);
          tc::umma_f16(d_s, q_hi + adv, k_hi + adv, idesc_s, 1u);
        }
        tc::umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int stage, bool first) {
        const uint32_t va = tc::smem_u32(sm_kv + stage * 4 * kAtKTileBytes + 2 * kAtKTileBytes);
        const uint64_t v_hi = tc::make_sw128_desc(va), v_lo = tc::make_sw128_desc(va + kAtKTileBytes);
        const uint32_t d_o = tmem_base + (uint32_t)(g * kAtTileCols + 128);
        // P(j) overwrote S(j) in place: fp16 values of the 128 keys in the first 64 columns that held their scores, two
        // keys per column.  P is carried as ONE fp16 number per probability: the row sum is taken over the same rounded
        // values, so the rounding is a re-weighting of the keys by 1 + eps, |eps| <= 2^-11, whose first-order effect on
        // O = sum p v / sum p is sum p eps (v - O) / sum p -- it vanishes for a peaked row and averages out over a diffuse
        // one (measured against fp64 in tests/test_gpu_attention.py) -- while V keeps its 22 bits: two MMAs per k-slice
        // instead of three.  (The logits keep all three products: their error is exponentiated.)
        const uint32_t p_base = tmem_base + (uint32_t)(g * kAtTileCols);
#pragma unroll
        for (int ks = 0; ks < kAtBK / 16; ++ks) {
          const uint32_t p_hi = p_base + (uint32_t)(ks * 8);
          const uint64_t adv_b = (uint64_t)(ks * 16 * 128 >> 4);    // 16 key rows of 128 B in the V tile
          tc::umma_f16_ts(d_o, p_hi, v_lo + adv_b, idesc_o, (first && ks == 0) ? 0u : 1u);
          tc::umma_f16_ts(d_o, p_hi, v_hi + adv_b, idesc_o, 1u);
        }
        tc::umma_commit(&pv_done[g]);
      };
      // K/V of key block b sit in ring stage b % kAtStages; that stage's kv_full completes for the (b / kAtStages)-th time
      auto wait_kv = [&](int blk) {
        tc::mbar_wait(&kv_full[blk % kAtStages], (uint32_t)((blk / kAtStages) & 1));
        tc::tc_fence_after();
      };
      tc::mbar_wait(q_full, 0);
      wait_kv(0);
      for (int g = 0; g < kAtTiles; ++g) issue_s(g, 0);
      for (int j = 0; j < nkb; ++j) {
        const int stage = j % kAtStages;
        for (int g = 0; g < kAtTiles; ++g) {
          tc::mbar_wait(&p_full[g], (uint32_t)(j & 1));   // P(j) written over S(j), O rescaled if needed
          tc::tc_fence_after();
          issue_pv(g, stage, j == 0);
          if (j + 1 < nkb) {   // S(j+1) of this tile goes into the region PV(j) has just read (same pipe, in order)
            if (g == 0) wait_kv(j + 1);
            issue_s(g, (j + 1) % kAtStages);
          }
        }
        tc::umma_commit(&kv_empty[stage]);
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int g = (warp - 2) >> 2;       // query tile of this warpgroup
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access (warp id % 4)
    const int r = quarter * 32 + lane;   // row inside the tile
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * kAtTileCols);
    const uint32_t t_o = t_row + 128;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkb; ++j) {
      const int valid = min(kAtBK, p.nk - j * kAtBK);
      tc::mbar_wait(&s_full[g], (uint32_t)(j & 1));
      tc::tc_fence_after();
      // the 128 scores of this row stay in registers between the max pass and the exp pass
      uint32_t sc[kAtBK];
#pragma unroll
      for (int c = 0; c < kAtBK; c += 32) tc::tmem_ld_32x32(t_row + c, *reinterpret_cast<uint32_t(*)[32]>(&sc[c]));
      tc::tmem_wait_ld();
      if (valid < kAtBK) {  // warp-uniform: only the last key block of a ragged Nk (e.g. the 77 context tokens)
#pragma unroll
        for (int i = 0; i < kAtBK; ++i)
          if (i >= valid) sc[i] = 0xff800000u;  // -inf: exp2 -> 0, max unaffected
      }
      // four independent chains for the row maximum and the row sum (one thread owns the whole row, so a single
      // chain would serialise dependent operations at the ALU latency)
      float mx0 = __uint_as_float(sc[0]), mx1 = __uint_as_float(sc[1]), mx2 = __uint_as_float(sc[2]),
            mx3 = __uint_as_float(sc[3]);
#pragma unroll
      for (int i = 4; i < kAtBK; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sc[i]));
        mx1 = fmaxf(mx1, __uint_as_float(sc[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sc[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sc[i + 3]));
      }
      const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
      // lazy rescale: keep the old reference maximum unless this block exceeds it by more than 2^kAtTau
      const bool grow = m_blk > m_run + kAtTau;          // always true on the first block (m_run = -inf)
      const float m_use = grow ? m_blk : m_run;
      const float alpha = grow ? ex2_approx(m_run - m_blk) : 1.0f;  // 0 on the first block
      if (j > 0) {
        // PV(j-1) has landed in O (it preceded S(j) in the tensor pipe, so this wait returns at once; it is still made
        // on EVERY block because an mbarrier parity wait is only meaningful one phase behind)
        tc::mbar_wait(&pv_done[g], (uint32_t)((j - 1) & 1));
        tc::tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
          // rare: raise the reference maximum of O; tcgen05.ld / .st are warp-wide, rows that keep their maximum
          // multiply by alpha = 1
#pragma unroll
          for (int c = 0; c < kAtD; c += 32) {
            uint32_t a0[32];
            tc::tmem_ld_32x32(t_o + c, a0);
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) a0[i] = __float_as_uint(__uint_as_float(a0[i]) * alpha);
            tc::tmem_st_32x32(t_o + c, a0);
          }
        }
      }
      // probabilities are carried scaled by 2^12 (<= 2^15 with the lazy maximum, far above fp16's subnormals); the row
      // sum is taken over the ROUNDED values and carries the same factor, which cancels in the final O / l.  P replaces S
      // in place, 32 keys at a time: 16 columns of fp16 pairs (the A operand of the PV product is read from tensor memory)
      const float bias = 12.0f - m_use;
      float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
#pragma unroll
      for (int c = 0; c < kAtBK; c += 32) {
        uint32_t hv[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float pv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) pv[u] = ex2_approx(fmaf(__uint_as_float(sc[c + i + u]), p.scale_log2, bias));
          const __half2 h01 = __floats2half2_rn(pv[0], pv[1]), h23 = __floats2half2_rn(pv[2], pv[3]);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          ls0 += f01.x; ls1 += f01.y; ls2 += f23.x; ls3 += f23.y;
          hv[i >> 1] = tc::h2_bits(h01);
          hv[(i >> 1) + 1] = tc::h2_bits(h23);
        }
        tc::tmem_st_32x16(t_row + (uint32_t)((c >> 5) * 16), hv);
      }
      tc::tmem_wait_st();
      l_run = fmaf(l_run, alpha, (ls0 + ls1) + (ls2 + ls3));
      m_run = m_use;
      tc::tc_fence_before();
      tc::mbar_arrive(&p_full[g]);
    }
    // all PV products accumulated: O / l
    tc::mbar_wait(&pv_done[g], (uint32_t)((nkb - 1) & 1));
    tc::tc_fence_after();
    float o[kAtD];
    {
      uint32_t (&o0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[0]);
      uint32_t (&o1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[32]);
      tc::tmem_ld_32x32(t_o, o0);
      tc::tmem_ld_32x32(t_o + 32, o1);
      tc::tmem_wait_ld();
    }
    const int q = q0 + g * kAtBQ + r;
    if (q < p.nq) {
      const float inv = 1.0f / l_run;
      const size_t off = ((size_t)b * p.nq + q) * ((size_t)p.heads * kAtD) + (size_t)head * kAtD;
#pragma unroll
      for (int d = 0; d < kAtD; ++d) o[d] *= inv;
      if (p.out_f32) {
#pragma unroll
        for (int d = 0; d < kAtD; d += 4)
          *reinterpret_cast<float4*>(p.out_f32 + off + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
      }
      if (p.out_hi) {
        const size_t row = ((size_t)b * p.nq + q) * ((size_t)p.heads * kAtD);
#pragma unroll
        for (int d = 0; d < kAtD; d += 8)
          tc::store_split8(p.out_hi + row, p.out_lo + row, head * kAtD + d, &o[d], p.out_packed8 != 0, tc::kAct8Sx, tc::kAct8Sl);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kAtTmemCols>(tmem_base);
}

}  // namespace vidseg

using namespace vidseg;

VS_API int vidseg_attention_split(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo,
                                  const void* v_hi, const void* v_lo, float* out_f32, void* out_hi, void* out_lo,
                                  int batch, int heads, int nq, int nk, float scale, void* stream) {
  VS_REQUIRE(q_hi && q_lo && k_hi && k_lo && v_hi && v_lo, "null operand pointer");
  VS_REQUIRE(out_f32 != nullptr || (out_hi != nullptr && out_lo != nullptr), "no output requested");
  VS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "out_hi and out_lo go together");
  VS_REQUIRE(batch >= 0 && heads >= 1 && nq >= 0 && nk >= 1, "bad shape");
  VS_REQUIRE(batch <= 65535 && heads <= 65535, "batch/heads exceed the grid limits");
  if (batch == 0 || nq == 0) return 0;
  const uint64_t c = (uint64_t)heads * kAtD;
  CUtensorMap tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo;
  if (int e = encode_tmap_3d_f16(&tq_hi, q_hi, c, nq, batch, c * 2, c * 2 * nq, kAtD, kAtBQ, 1)) return e;
  if (int e = encode_tmap_3d_f16(&tq_lo, q_lo, c, nq, batch, c * 2, c * 2 * nq, kAtD, kAtBQ, 1)) return e;
  if (int e = encode_tmap_3d_f16(&tk_hi, k_hi, c, nk, batch, c * 2, c * 2 * nk, kAtD, kAtBK, 1)) return e;
  if (int e = encode_tmap_3d_f16(&tk_lo, k_lo, c, nk, batch, c * 2, c * 2 * nk, kAtD, kAtBK, 1)) return e;
  if (int e = encode_tmap_3d_f16(&tv_hi, v_hi, c, nk, batch, c * 2, c * 2 * nk, kAtD, kAtBK, 1)) return e;
  if (int e = encode_tmap_3d_f16(&tv_lo, v_lo, c, nk, batch, c * 2, c * 2 * nk, kAtD, kAtBK, 1)) return e;
  AttnParams p{batch, heads, nq, nk, scale * 1.4426950408889634f, out_f32, (__half*)out_hi, (__half*)out_lo,
               operand_packed8((long long)heads * kAtD) ? 1 : 0};
  const double flops = 4.0 * batch * heads * (double)nq * nk * kAtD;
  if (nk <= kAtBK) {   // one key block: the light CTA, two per SM
    using Cfg = AtCfg<1, 1>;
    static cudaError_t attr1 = cudaFuncSetAttribute(attn_split_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    VS_CHECK_CUDA(attr1);
    dim3 grid((nq + kAtBQ - 1) / kAtBQ, heads, batch);
    VS_LAUNCH_W(flops, (attn_split_kernel<1, 1>), grid, Cfg::kThreads, Cfg::kSmemBytes, stream, tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, p);
  } else {
    using Cfg = AtCfg<2, 2>;
    static cudaError_t attr2 = cudaFuncSetAttribute(attn_split_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    VS_CHECK_CUDA(attr2);
    dim3 grid((nq + kAtBQ * 2 - 1) / (kAtBQ * 2), heads, batch);
    VS_LAUNCH_W(flops, (attn_split_kernel<2, 2>), grid, Cfg::kThreads, Cfg::kSmemBytes, stream, tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, p);
  }
  VS_POST_LAUNCH();
  return 0;
}
