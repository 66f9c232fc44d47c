// Scaled-dot-product attention of the UNet transformer blocks on tcgen05 tensor cores.
//
// Replaces F.scaled_dot_product_attention at sgm/modules/attention.py:352-356 (and the xformers
// call :473-485): out[b, i, h*64:(h+1)*64] = softmax(q_i . k_j / 8) v_j over the keys of (b, h), head dim 64,
// no mask.  q/k/v arrive pre-head-split exactly as the reference stashes them ([B, N, heads*64]), as
// split-fp16 pairs (see gemm_tc.cu); both GEMMs of the flash loop are 3-MMA split products with fp32
// accumulation in TMEM, the softmax runs in fp32 registers.
//
// One CTA = 256 queries of one (batch, head): two 128-row query tiles, each owned by a softmax
// warpgroup (128 threads = 128 TMEM lanes); the two groups ping-pong so that the tensor pipe works on
// one tile while the CUDA cores exponentiate the other.  Per 64-key block and tile:
//   MMA warp : S = Q K^T   (12 x tcgen05.mma 128x64x16: lo.hi + hi.lo + hi.hi into one fp32 accumulator)
//   softmax  : tcgen05.ld S, online max / exp2 / row sum, P -> fp16 hi/lo into swizzled smem
//   MMA warp : PV = P V    (12 x tcgen05.mma 128x64x16, V is the MN-major B operand)
//   softmax  : tcgen05.ld PV, O = O * alpha + PV   (O lives in registers, 64 fp32 per thread)
// TMA (warp 0) streams K/V blocks through a 2-stage ring.  TMEM: 2 tiles x (S 64 + PV 64) = 256 columns.
#include <mutex>

#define VS_FAMILY vidseg::kFamAttention
#include "common.cuh"
#include "tc_common.cuh"

namespace vidseg {

constexpr int kAtBQ = 128;        // queries per tile
constexpr int kAtTiles = 2;       // query tiles per CTA
constexpr int kAtBK = 64;         // keys per block
constexpr int kAtD = 64;          // head dim
constexpr int kAtStages = 2;
constexpr int kAtQTileBytes = kAtBQ * kAtD * 2;  // 16 KB (one of hi / lo)
constexpr int kAtKTileBytes = kAtBK * kAtD * 2;  // 8 KB
constexpr int kAtPTileBytes = kAtBQ * kAtBK * 2; // 16 KB
constexpr int kAtSmemQ = kAtTiles * 2 * kAtQTileBytes;          // 64 KB
constexpr int kAtSmemKV = kAtStages * 4 * kAtKTileBytes;        // 64 KB
constexpr int kAtSmemP = kAtTiles * 2 * kAtPTileBytes;          // 64 KB
constexpr int kAtSmemBytes = kAtSmemQ + kAtSmemKV + kAtSmemP + 1024 + 256;
constexpr int kAtThreads = 64 + kAtTiles * 128;  // producer warp, MMA warp, 2 softmax warpgroups

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
};

__global__ void __launch_bounds__(kAtThreads, 1)
attn_split_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                  const __grid_constant__ CUtensorMap tm_k_hi, const __grid_constant__ CUtensorMap tm_k_lo,
                  const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                  const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sm_q = smem;                       // [tile][hi|lo][128x64]
  uint8_t* sm_kv = sm_q + kAtSmemQ;           // [stage][k_hi|k_lo|v_hi|v_lo][64x64]
  uint8_t* sm_p = sm_kv + kAtSmemKV;          // [tile][hi|lo][128x64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + kAtSmemP);
  uint64_t* q_full = bars;                    // 1
  uint64_t* kv_full = bars + 1;               // [2]
  uint64_t* kv_empty = bars + 3;              // [2]
  uint64_t* s_full = bars + 5;                // [2]
  uint64_t* p_full = bars + 7;                // [2]
  uint64_t* o_full = bars + 9;                // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 11);

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
    for (int g = 0; g < kAtTiles; ++g) { tc::mbar_init(&s_full[g], 1); tc::mbar_init(&p_full[g], 128); tc::mbar_init(&o_full[g], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<256>(tmem_base_ptr);
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
        const uint32_t d_s = tmem_base + (uint32_t)(g * 128);
#pragma unroll
        for (int ks = 0; ks < kAtD / 16; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 32 >> 4);
          tc::umma_f16(d_s, q_lo + adv, k_hi + adv, idesc_s, ks > 0);
          tc::umma_f16(d_s, q_hi + adv, k_lo + adv, idesc_s, 1u);
          tc::umma_f16(d_s, q_hi + adv, k_hi + adv, idesc_s, 1u);
        }
        tc::umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int stage) {
        const uint32_t pa = tc::smem_u32(sm_p + g * 2 * kAtPTileBytes);
        const uint32_t va = tc::smem_u32(sm_kv + stage * 4 * kAtKTileBytes + 2 * kAtKTileBytes);
        const uint64_t p_hi = tc::make_sw128_desc(pa), p_lo = tc::make_sw128_desc(pa + kAtPTileBytes);
        const uint64_t v_hi = tc::make_sw128_desc(va), v_lo = tc::make_sw128_desc(va + kAtKTileBytes);
        const uint32_t d_o = tmem_base + (uint32_t)(g * 128 + 64);
#pragma unroll
        for (int ks = 0; ks < kAtBK / 16; ++ks) {
          const uint64_t adv_a = (uint64_t)(ks * 32 >> 4);          // 16 keys along K inside P's swizzle atom
          const uint64_t adv_b = (uint64_t)(ks * 16 * 128 >> 4);    // 16 key rows of 128 B in the V tile
          tc::umma_f16(d_o, p_lo + adv_a, v_hi + adv_b, idesc_o, ks > 0);
          tc::umma_f16(d_o, p_hi + adv_a, v_lo + adv_b, idesc_o, 1u);
          tc::umma_f16(d_o, p_hi + adv_a, v_hi + adv_b, idesc_o, 1u);
        }
        tc::umma_commit(&o_full[g]);
      };
      tc::mbar_wait(q_full, 0);
      tc::mbar_wait(&kv_full[0], 0);
      tc::tc_fence_after();
      for (int g = 0; g < kAtTiles; ++g) issue_s(g, 0);
      int stage = 0;
      uint32_t phase = 0;  // phase of kv_full[stage] for block j
      for (int j = 0; j < nkb; ++j) {
        const int nstage = (stage + 1 == kAtStages) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == kAtStages) ? (phase ^ 1) : phase;
        const bool more = (j + 1 < nkb);
        if (more) { tc::mbar_wait(&kv_full[nstage], nphase); tc::tc_fence_after(); }
        for (int g = 0; g < kAtTiles; ++g) {
          tc::mbar_wait(&p_full[g], (uint32_t)(j & 1));
          tc::tc_fence_after();
          issue_pv(g, stage);
          if (more) issue_s(g, nstage);
        }
        tc::umma_commit(&kv_empty[stage]);
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int g = (warp - 2) >> 2;       // query tile of this warpgroup
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access (warp id % 4)
    const int r = quarter * 32 + lane;   // row inside the tile
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 128);
    uint8_t* p_hi_tile = sm_p + (g * 2 + 0) * kAtPTileBytes;
    uint8_t* p_lo_tile = sm_p + (g * 2 + 1) * kAtPTileBytes;
    float o[kAtD];
#pragma unroll
    for (int d = 0; d < kAtD; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkb; ++j) {
      const int valid = min(kAtBK, p.nk - j * kAtBK);
      tc::mbar_wait(&s_full[g], (uint32_t)(j & 1));
      tc::tc_fence_after();
      // the 64 scores of this row stay in registers between the max pass and the exp pass
      uint32_t sc[kAtBK];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sc[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sc[32]);
        tc::tmem_ld_32x32(t_row, s0);
        tc::tmem_ld_32x32(t_row + 32, s1);
        tc::tmem_wait_ld();
      }
      if (valid < kAtBK) {  // warp-uniform: only the last key block of a ragged Nk (e.g. the 77 context tokens)
#pragma unroll
        for (int i = 0; i < kAtBK; ++i)
          if (i >= valid) sc[i] = 0xff800000u;  // -inf: exp2 -> 0, max unaffected
      }
      // four independent chains for the row maximum and the row sum (one thread owns the whole row, so a single
      // chain would serialise 64 dependent operations at the ALU latency)
      float mx0 = __uint_as_float(sc[0]), mx1 = __uint_as_float(sc[1]), mx2 = __uint_as_float(sc[2]),
            mx3 = __uint_as_float(sc[3]);
#pragma unroll
      for (int i = 4; i < kAtBK; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sc[i]));
        mx1 = fmaxf(mx1, __uint_as_float(sc[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sc[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sc[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = ex2_approx(m_run - m_new);  // 0 on the first block (m_run = -inf)
      // probabilities are carried scaled by 2^12 (<= 4096) so that their fp16 residuals stay normal numbers;
      // the row sum carries the same factor, which cancels in the final O / l
      const float bias = 12.0f - m_new;
      float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
      // probabilities -> fp16 hi / lo -> swizzled smem (A operand of the PV product)
#pragma unroll
      for (int i = 0; i < kAtBK; i += 8) {
        float pv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) pv[u] = ex2_approx(fmaf(__uint_as_float(sc[i + u]), p.scale_log2, bias));
        ls0 += pv[0] + pv[4]; ls1 += pv[1] + pv[5]; ls2 += pv[2] + pv[6]; ls3 += pv[3] + pv[7];
        uint4 hv, lv;
        tc::split8_f16(pv[0], pv[1], pv[2], pv[3], pv[4], pv[5], pv[6], pv[7], hv, lv);
        const int chunk = i >> 3;  // 16-byte chunk index inside the 128-byte row
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(p_hi_tile + off) = hv;
        *reinterpret_cast<uint4*>(p_lo_tile + off) = lv;
      }
      const float lsum = (ls0 + ls1) + (ls2 + ls3);
      l_run = fmaf(l_run, alpha, lsum);
      m_run = m_new;
      // make the generic-proxy smem writes visible to the tensor core (async proxy), then signal
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc::tc_fence_before();
      tc::mbar_arrive(&p_full[g]);
      // PV of this block
      tc::mbar_wait(&o_full[g], (uint32_t)(j & 1));
      tc::tc_fence_after();
#pragma unroll
      for (int c = 0; c < kAtD; c += 32) {
        uint32_t a0[32];
        tc::tmem_ld_32x32(t_row + 64 + c, a0);
        tc::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha, __uint_as_float(a0[i]));
      }
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
#pragma unroll
        for (int d = 0; d < kAtD; d += 8) {
          uint4 hv, lv;
          tc::split8_f16(o[d], o[d + 1], o[d + 2], o[d + 3], o[d + 4], o[d + 5], o[d + 6], o[d + 7], hv, lv);
          *reinterpret_cast<uint4*>(p.out_hi + off + d) = hv;
          *reinterpret_cast<uint4*>(p.out_lo + off + d) = lv;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<256>(tmem_base);
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
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes);
  });
  VS_CHECK_CUDA(attr_err);
  AttnParams p{batch, heads, nq, nk, scale * 1.4426950408889634f, out_f32, (__half*)out_hi, (__half*)out_lo};
  dim3 grid((nq + kAtBQ * kAtTiles - 1) / (kAtBQ * kAtTiles), heads, batch);
  VS_LAUNCH_W(4.0 * batch * heads * (double)nq * nk * kAtD, attn_split_kernel, grid, kAtThreads, kAtSmemBytes, stream, tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, p);
  VS_POST_LAUNCH();
  return 0;
}
