// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 MMA / TMEM, shared-memory matrix descriptors.  Raw PTX, no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

namespace vidseg {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// same, acquiring at cluster scope: the arrivals come from the peer CTA of a pair
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "WAIT_DONE_C:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA ----
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of one TPC issue ONE tcgen05.mma of M = 256.  CTA
// rank r holds rows [128 r, 128 r + 128) of the A tile and of the accumulator (its own TMEM) and HALF of the B tile (N/2
// rows) in its own shared memory; the tensor cores of both SMs read both halves.  Only the leader (rank 0) issues the
// MMAs and waits for the operands, so the TMA loads of BOTH CTAs complete on the leader's mbarrier.
// shared::cluster address of the same shared-memory offset in CTA `cta_rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// loads into THIS CTA's shared memory, completion bytes to the mbarrier at shared::cluster address `bar_addr`
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in every CTA of cta_mask once all previously issued pair MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- TMEM ----
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit: thread t of the warp gets lane (base_lane + t), 32 consecutive columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 8 columns (compact loops over columns: the unrolled x32 form costs ~1 KB of code per use)
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// 32 lanes x 32 columns of 32-bit back into TMEM (same lane/column mapping as tmem_ld_32x32)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA ----
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle,
// tile rows of 128 bytes, 8-row groups of 1024 bytes.  Works for K-major operands (rows = M/N index,
// 64 fp16 of K per row) and MN-major operands (rows = K index, 64 fp16 of M/N per row).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr_bytes & 0x3FFFFu) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                                     // leading byte offset (unused with swizzle) = 1
  d |= (uint64_t)(1024u >> 4) << 32;                          // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                                     // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                                     // layout type SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 with fp16 operands and fp32 accumulation
// (cute::UMMA::InstrDescriptor): c_format[4,6)=1 (F32), a/b_format=0 (F16), a_major bit15, b_major bit16,
// n_dim[17,23)=N>>3, m_dim[24,29)=M>>4.
// a_bf16 / b_bf16 select the operand element type (a_format bits [7,10), b_format bits [10,13): 0 = F16, 1 = BF16);
// the hardware requires both operands of one MMA to have the same format.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, int a_mn_major, int b_mn_major, int a_bf16 = 0,
                                                      int b_bf16 = 0) {
  return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from TENSOR MEMORY (row i of A = TMEM lane i, two consecutive 16-bit K elements per
// 32-bit column): the P V product of the attention kernel, whose P tile is written by the softmax warps straight into
// TMEM and never touches shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Split an fp32 value into hi = fp16(x) and lo = fp16(x - hi): x ~= hi + lo.  All three products
// lo.hi + hi.lo + hi.hi of a split GEMM accumulate in ONE fp32 TMEM accumulator.  (tcgen05 kind::f16 rejects an
// fp16 x bf16 operand mix -- illegal instruction, measured -- so lo cannot borrow bf16's exponent range.)
// |lo| <= 2^-12 |x| is a normal fp16 number for |x| >= ~0.25 (22 significant bits); below that the residual falls
// into fp16's subnormals and the absolute error of the pair is bounded by 2^-25 ~ 3e-8.  Operands whose values
// are systematically tiny are therefore carried pre-scaled by a power of two and the exact inverse is applied to
// the fp32 accumulator: weights x 2^8 (kWeightScale), softmax probabilities x 2^12 (cancels in the normalisation).
// Activations are unscaled, so the overflow threshold is fp16's 65504 -- the same as the reference's own
// fp16-autocast GPU path.
constexpr float kWeightScale = 256.0f;
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// two fp32 -> packed (hi, hi) and (lo, lo) fp16 pairs, all in registers
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  hi = h2_bits(h);
  lo = h2_bits(__floats2half2_rn(a - hf.x, b - hf.y));
}
// eight fp32 -> one 16-byte vector of fp16 hi and one of fp16 lo
__device__ __forceinline__ void split8_f16(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7,
                                           uint4& hi, uint4& lo) {
  split2_f16(v0, v1, hi.x, lo.x);
  split2_f16(v2, v3, hi.y, lo.y);
  split2_f16(v4, v5, hi.z, lo.z);
  split2_f16(v6, v7, hi.w, lo.w);
}
__device__ __forceinline__ void split4_f16(float v0, float v1, float v2, float v3, uint2& hi, uint2& lo) {
  split2_f16(v0, v1, hi.x, lo.x);
  split2_f16(v2, v3, hi.y, lo.y);
}

// ---------------------------------------------------------------------------------------------------------------
// Operand formats of the tensor-core GEMMs.
//   pair16  : x ~= hi + lo, two fp16 tensors -> three kind::f16 MMAs per product (lo.hi + hi.lo + hi.hi), 22 bits.
//   packed8 : hi = fp16(x) plus an fp8 side tensor of the SAME size as lo: every 64-element block of a row is 128
//             bytes = [lo8: e5m2((x - hi) * sl), 64 B | x8: e4m3(x * sx), 64 B].  The two correction products
//             (x - hi).w and x.(w - w_hi) only need ~4 significant bits each (they are 2^-11 of the result), so they
//             run as kind::f8f6f4 MMAs at twice the fp16 rate: one product costs 1 + 1/2 + 1/2 = 2 fp16-MMA units
//             instead of 3, at 2^-14.5 relative precision per product (measured 1.6e-5 rms per GEMM against fp64; the
//             parity bar of the path is 1e-3).  Scales: activations (sx, sl) = (1, 16), weights -- already carrying
//             2^8 -- (1/16, 1), so that lo8_a.x8_w and x8_a.lo8_w both land on the 2^8 scale of hi_a.hi_w and all
//             three accumulate in ONE fp32 TMEM accumulator.  A 64-element block is one 128-byte swizzle row, i.e.
//             the fp8 tensor moves through the very same TMA maps as the fp16 lo tensor did.
// Needs the row length (channels) to be a multiple of 64; other operands stay pair16.
// ---------------------------------------------------------------------------------------------------------------
constexpr float kAct8Sx = 1.0f, kAct8Sl = 16.0f;
constexpr float kWgt8Sx = 0.0625f, kWgt8Sl = 1.0f;

__device__ __forceinline__ uint32_t pack4_fp8(float a, float b, float c, float d, __nv_fp8_interpretation_t kind) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, kind);   // a in the low byte
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, kind);
  return lo | (hi << 16);
}
// four consecutive elements c..c+3 (c % 4 == 0) of one operand row
__device__ __forceinline__ void store_split4(__half* hi_row, __half* lo_row, int c, float v0, float v1, float v2,
                                             float v3, bool packed8, float sx, float sl) {
  const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  *reinterpret_cast<uint2*>(hi_row + c) = make_uint2(h2_bits(h01), h2_bits(h23));
  const float r0 = v0 - f01.x, r1 = v1 - f01.y, r2 = v2 - f23.x, r3 = v3 - f23.y;
  if (!packed8) {
    *reinterpret_cast<uint2*>(lo_row + c) = make_uint2(h2_bits(__floats2half2_rn(r0, r1)), h2_bits(__floats2half2_rn(r2, r3)));
  } else {
    uint8_t* aux = reinterpret_cast<uint8_t*>(lo_row) + ((c >> 6) << 7) + (c & 63);
    *reinterpret_cast<uint32_t*>(aux) = pack4_fp8(r0 * sl, r1 * sl, r2 * sl, r3 * sl, __NV_E5M2);
    *reinterpret_cast<uint32_t*>(aux + 64) = pack4_fp8(v0 * sx, v1 * sx, v2 * sx, v3 * sx, __NV_E4M3);
  }
}
// eight consecutive elements (c % 8 == 0)
__device__ __forceinline__ void store_split8(__half* hi_row, __half* lo_row, int c, const float* v, bool packed8, float sx,
                                             float sl) {
  if (!packed8) {
    uint4 hv, lv;
    split8_f16(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], hv, lv);
    *reinterpret_cast<uint4*>(hi_row + c) = hv;
    *reinterpret_cast<uint4*>(lo_row + c) = lv;
  } else {
    uint4 hv;
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const __half2 h = __floats2half2_rn(v[i], v[i + 1]);
      const float2 f = __half22float2(h);
      (&hv.x)[i >> 1] = h2_bits(h);
      r[i] = (v[i] - f.x) * sl;
      r[i + 1] = (v[i + 1] - f.y) * sl;
    }
    *reinterpret_cast<uint4*>(hi_row + c) = hv;
    uint8_t* aux = reinterpret_cast<uint8_t*>(lo_row) + ((c >> 6) << 7) + (c & 63);
    *reinterpret_cast<uint2*>(aux) = make_uint2(pack4_fp8(r[0], r[1], r[2], r[3], __NV_E5M2), pack4_fp8(r[4], r[5], r[6], r[7], __NV_E5M2));
    *reinterpret_cast<uint2*>(aux + 64) = make_uint2(pack4_fp8(v[0] * sx, v[1] * sx, v[2] * sx, v[3] * sx, __NV_E4M3),
                                                     pack4_fp8(v[4] * sx, v[5] * sx, v[6] * sx, v[7] * sx, __NV_E4M3));
  }
}

// Instruction descriptor for kind::f8f6f4 (same bit layout as make_idesc_f16; a/b_format: 0 = E4M3, 1 = E5M2), fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_f8(int m, int n, int a_e5m2, int b_e5m2) {
  return (1u << 4) | ((uint32_t)a_e5m2 << 7) | ((uint32_t)b_e5m2 << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem], 8-bit operands, K = 32 per instruction
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor for kind::i8: signed 8-bit operands (a/b_format = 1), 32-bit integer accumulation (c_format = 2)
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem], int8 operands, int32 accumulators, K = 32 per instruction
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace tc

// library-wide operand policy (vidseg_set_operand_mode): 0 = pair16 everywhere, 1 = packed8 wherever the channel count allows
extern std::atomic<int> g_operand_mode;
inline bool operand_packed8(long long channels) { return g_operand_mode.load(std::memory_order_relaxed) != 0 && channels % 64 == 0; }

// host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
int encode_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer);
int encode_tmap_3d_f16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2);
// generic rank-1..5 map of 16-bit elements (fp16 and bf16 tiles move identically), 128-byte swizzle, zero OOB fill
int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);

}  // namespace vidseg
