// Temporal self-attention of the SVD VideoTransformerBlock (reference: sgm/modules/video_attention.py:152,195 ->
// CrossAttention.forward, sgm/modules/attention.py:286-364 on the '(b s) t c' layout).
//
// The sequence is the T <= 32 frames of one spatial site; there are V*S*heads such sequences (40960 at the first
// UNet level of a 14-frame 512x512 clip), each a 14x14 score matrix over 64 channels.  That is 50 kFLOP per
// sequence -- far too small for a tensor-core tile -- so the op is bound by moving q, k, v once through HBM.
// The kernel therefore works directly on the frame-major activation layout the rest of the UNet uses
// ([V, T, S, C], i.e. the reference's '(b t) s c'): the '(b t) s c -> (b s) t c' rearrange and its inverse
// (video_attention.py:152, 282-284) are index arithmetic inside the loads, not passes over memory.
// One warp per (video, site, head): q, k, v rows staged in shared memory with 128-bit loads, scores and softmax in
// fp32 registers (lane j owns key j), PV with lane d owning two output channels, result written as the split
// operand (hi | lo fp16) of the to_out GEMM.
#define VS_FAMILY vidseg::kFamAttention
#include "common.cuh"
#include "tc_common.cuh"

namespace vidseg {

constexpr int kTaD = 64;          // head dim
constexpr int kTaWarps = 4;       // warps (sequences) per block

template <int kMaxT>
__global__ void __launch_bounds__(kTaWarps * 32)
temporal_attn_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     float* __restrict__ out_f32, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                     int videos, int frames, int sites, int heads, float scale, int packed8) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long seq = (long long)blockIdx.x * kTaWarps + warp;
  const long long total = (long long)videos * sites * heads;
  if (seq >= total) return;  // no block-wide barrier below
  const int head = (int)(seq % heads);
  const int site = (int)((seq / heads) % sites);
  const int vid = (int)(seq / ((long long)heads * sites));
  const int T = frames;
  const int c = heads * kTaD;
  constexpr int kKStride = kTaD + 1;  // lane j reads row j: odd stride -> conflict-free
  float* qs = sm + (size_t)warp * (kMaxT * (2 * kTaD + kKStride));
  float* ks = qs + kMaxT * kTaD;
  float* vs = ks + kMaxT * kKStride;
  // row of frame j: ((vid*T + j)*S + site)*C + head*64
  const size_t row0 = ((size_t)vid * T * sites + site) * c + (size_t)head * kTaD;
  const size_t frame_stride = (size_t)sites * c;
  for (int idx = lane; idx < T * (kTaD / 4); idx += 32) {
    const int j = idx >> 4, c4 = idx & 15;
    const size_t off = row0 + (size_t)j * frame_stride + c4 * 4;
    const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(q + off));
    const float4 b = ld_stream_f4(reinterpret_cast<const float4*>(k + off));
    const float4 d = ld_stream_f4(reinterpret_cast<const float4*>(v + off));
    *reinterpret_cast<float4*>(qs + j * kTaD + c4 * 4) = a;
    float* kd = ks + j * kKStride + c4 * 4;
    kd[0] = b.x; kd[1] = b.y; kd[2] = b.z; kd[3] = b.w;
    *reinterpret_cast<float4*>(vs + j * kTaD + c4 * 4) = d;
  }
  __syncwarp();
  // scores: lane j holds key j in registers
  const int jrow = min(lane, T - 1);
  float kreg[kTaD];
#pragma unroll
  for (int d = 0; d < kTaD; ++d) kreg[d] = ks[jrow * kKStride + d];
  float prob[kMaxT];
#pragma unroll
  for (int i = 0; i < kMaxT; ++i) {
    prob[i] = 0.f;
    if (i < T) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < kTaD; d += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(qs + i * kTaD + d);  // broadcast
        acc = fmaf(qq.x, kreg[d], acc); acc = fmaf(qq.y, kreg[d + 1], acc);
        acc = fmaf(qq.z, kreg[d + 2], acc); acc = fmaf(qq.w, kreg[d + 3], acc);
      }
      const float sc = (lane < T) ? acc * scale : -INFINITY;
      const float m = warp_max(sc);
      const float e = (lane < T) ? expf(sc - m) : 0.f;
      prob[i] = e / warp_sum(e);
    }
  }
  // PV: lane owns channels 2*lane, 2*lane+1
  float2 vreg[kMaxT];
#pragma unroll
  for (int j = 0; j < kMaxT; ++j)
    vreg[j] = (j < T) ? *reinterpret_cast<const float2*>(vs + j * kTaD + 2 * lane) : make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kMaxT; ++i) {
    if (i < T) {
      float ox = 0.f, oy = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxT; ++j) {
        if (j < T) {
          const float pj = __shfl_sync(0xffffffffu, prob[i], j);
          ox = fmaf(pj, vreg[j].x, ox);
          oy = fmaf(pj, vreg[j].y, oy);
        }
      }
      const size_t off = row0 + (size_t)i * frame_stride + 2 * lane;
      if (out_f32) *reinterpret_cast<float2*>(out_f32 + off) = make_float2(ox, oy);
      if (out_hi) {
        // a lane pair shares one 4-element group: even lanes store (the packed8 format is written 4 elements at a time)
        const float px = __shfl_down_sync(0xffffffffu, ox, 1), py = __shfl_down_sync(0xffffffffu, oy, 1);
        if ((lane & 1) == 0) {
          const size_t rowoff = row0 - (size_t)head * kTaD + (size_t)i * frame_stride;   // start of the token's row
          tc::store_split4(out_hi + rowoff, out_lo + rowoff, head * kTaD + 2 * lane, ox, oy, px, py, packed8 != 0,
                           tc::kAct8Sx, tc::kAct8Sl);
        }
      }
    }
  }
}

}  // namespace vidseg

using namespace vidseg;

template <int kMaxT>
static int launch_temporal(const float* q, const float* k, const float* v, float* out_f32, void* out_hi, void* out_lo,
                           int videos, int frames, int sites, int heads, float scale, void* stream) {
  const long long total = (long long)videos * sites * heads;
  const long long blocks = (total + kTaWarps - 1) / kTaWarps;
  VS_REQUIRE(blocks <= 0x7fffffffLL, "too many sequences");
  const size_t smem = (size_t)kTaWarps * kMaxT * (2 * kTaD + kTaD + 1) * sizeof(float);
  static cudaError_t attr_err = cudaFuncSetAttribute(temporal_attn_kernel<kMaxT>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  VS_CHECK_CUDA(attr_err);
  const double bytes = 16.0 * (double)videos * frames * sites * heads * kTaD;  // q, k, v fp32 in; hi | lo fp16 out
  VS_LAUNCH_W(bytes, temporal_attn_kernel<kMaxT>, (int)blocks, kTaWarps * 32, smem, stream, q, k, v, out_f32,
              (__half*)out_hi, (__half*)out_lo, videos, frames, sites, heads, scale,
              operand_packed8((long long)heads * kTaD) ? 1 : 0);
  VS_POST_LAUNCH();
  return 0;
}

VS_API int vidseg_temporal_attention(const float* q, const float* k, const float* v, float* out_f32, void* out_hi,
                                     void* out_lo, int videos, int frames, int sites, int heads, float scale,
                                     void* stream) {
  VS_REQUIRE(q && k && v, "null pointer");
  VS_REQUIRE(out_f32 != nullptr || (out_hi != nullptr && out_lo != nullptr), "no output requested");
  VS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "out_hi and out_lo go together");
  VS_REQUIRE(videos >= 0 && sites >= 0 && heads >= 1, "bad shape");
  VS_REQUIRE(frames >= 1 && frames <= 32, "temporal attention supports 1..32 frames");
  VS_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0), "unaligned pointer");
  if (videos == 0 || sites == 0) return 0;
  if (frames <= 16) return launch_temporal<16>(q, k, v, out_f32, out_hi, out_lo, videos, frames, sites, heads, scale, stream);
  return launch_temporal<32>(q, k, v, out_f32, out_hi, out_lo, videos, frames, sites, heads, scale, stream);
}
