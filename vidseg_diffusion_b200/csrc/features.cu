// R1: fused multi-block mean -> cond-half slice -> per-token max-abs normalise.
//
// Replaces scripts/sampling/feature_extraction.py:739-745 (torch.mean(torch.stack(blocks), 0),
// i.e. ((b0 + b1) + b2) / n in fp32 on the CPU path), :38-39 (x / max|x| over channels, IEEE
// division, no epsilon) and :45-46 (rows [F, 2F) only).  HBM-bound streaming kernel: one warp
// per token row, 128-bit loads, the row stays in registers between the max reduction and the
// divide, so every input byte is read once and every output byte written once.
#define VS_FAMILY vidseg::kFamAggregate
#include "common.cuh"

namespace vidseg {

struct BlockPtrs {
  const float* p[4];
};

template <int NB>
__device__ __forceinline__ float agg1(const float (&v)[4]) {
  float s = v[0];
#pragma unroll
  for (int b = 1; b < NB; ++b) s = __fadd_rn(s, v[b]);
  if (NB > 1) s = __fdiv_rn(s, (float)NB);
  return s;
}

// VEC4 = number of float4 per lane kept in registers (C <= 128 * VEC4, C % 4 == 0)
template <int NB, int VEC4>
__global__ void __launch_bounds__(256) aggregate_normalize_vec_kernel(BlockPtrs blocks, float* __restrict__ out,
                                                                       int rows, int hw_frames_offset_rows,
                                                                       int channels) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int c4 = channels >> 2;
  const size_t in_row = (size_t)(warp + hw_frames_offset_rows) * channels;
  float4 acc[VEC4];
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < VEC4; ++i) {
    const int idx = lane + 32 * i;
    if (idx < c4) {
      float4 t[4];
#pragma unroll
      for (int b = 0; b < NB; ++b) t[b] = ld_stream_f4(reinterpret_cast<const float4*>(blocks.p[b] + in_row) + idx);
      float vx[4], vy[4], vz[4], vw[4];
#pragma unroll
      for (int b = 0; b < NB; ++b) { vx[b] = t[b].x; vy[b] = t[b].y; vz[b] = t[b].z; vw[b] = t[b].w; }
      acc[i].x = agg1<NB>(vx); acc[i].y = agg1<NB>(vy); acc[i].z = agg1<NB>(vz); acc[i].w = agg1<NB>(vw);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(acc[i].x), fabsf(acc[i].y)), fmaxf(fabsf(acc[i].z), fabsf(acc[i].w))));
    }
  }
  m = warp_max(m);
  float4* o = reinterpret_cast<float4*>(out + (size_t)warp * channels);
#pragma unroll
  for (int i = 0; i < VEC4; ++i) {
    const int idx = lane + 32 * i;
    if (idx < c4) {
      float4 r = acc[i];
      if (channels > 1) {
        r.x = __fdiv_rn(r.x, m); r.y = __fdiv_rn(r.y, m); r.z = __fdiv_rn(r.z, m); r.w = __fdiv_rn(r.w, m);
      }
      o[idx] = r;
    }
  }
}

// generic fallback: any C, two passes over the inputs (second pass hits L2)
template <int NB>
__global__ void __launch_bounds__(256) aggregate_normalize_generic_kernel(BlockPtrs blocks, float* __restrict__ out,
                                                                           int rows, int hw_frames_offset_rows,
                                                                           int channels) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const size_t in_row = (size_t)(warp + hw_frames_offset_rows) * channels;
  float m = 0.f;
  for (int c = lane; c < channels; c += 32) {
    float v[4];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = blocks.p[b][in_row + c];
    m = fmaxf(m, fabsf(agg1<NB>(v)));
  }
  m = warp_max(m);
  for (int c = lane; c < channels; c += 32) {
    float v[4];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = blocks.p[b][in_row + c];
    float s = agg1<NB>(v);
    if (channels > 1) s = __fdiv_rn(s, m);
    out[(size_t)warp * channels + c] = s;
  }
}

template <int NB>
static int launch_aggregate(const BlockPtrs& bp, float* out, int rows, int offset_rows, int channels, void* stream) {
  const int threads = 256;
  const int grid = (rows * 32 + threads - 1) / threads;
  bool aligned = (channels % 4 == 0) && ((uintptr_t)out % 16 == 0);
  for (int b = 0; b < NB; ++b) aligned = aligned && ((uintptr_t)bp.p[b] % 16 == 0);
  if (aligned && channels <= 128 * 2) {
    VS_LAUNCH((aggregate_normalize_vec_kernel<NB, 2>), grid, threads, 0, stream, bp, out, rows, offset_rows, channels);
  } else if (aligned && channels <= 128 * 5) {
    VS_LAUNCH((aggregate_normalize_vec_kernel<NB, 5>), grid, threads, 0, stream, bp, out, rows, offset_rows, channels);
  } else if (aligned && channels <= 128 * 10) {
    VS_LAUNCH((aggregate_normalize_vec_kernel<NB, 10>), grid, threads, 0, stream, bp, out, rows, offset_rows, channels);
  } else {
    VS_LAUNCH((aggregate_normalize_generic_kernel<NB>), grid, threads, 0, stream, bp, out, rows, offset_rows, channels);
  }
  VS_POST_LAUNCH();
  return 0;
}

}  // namespace vidseg

using namespace vidseg;

VS_API int vidseg_aggregate_normalize(const float* const* blocks_host, int n_blocks, int num_frames, int hw,
                                          int channels, float* out, void* stream) {
  VS_REQUIRE(n_blocks >= 1 && n_blocks <= 4, "n_blocks must be 1..4");
  VS_REQUIRE(num_frames >= 0 && hw >= 0 && channels >= 1, "bad shape");
  const long long rows_ll = (long long)num_frames * hw;
  VS_REQUIRE(rows_ll < (1ll << 26), "too many rows");
  const int rows = (int)rows_ll;
  if (rows == 0) return 0;  // empty clip: nothing to do (pointers may be null)
  VS_REQUIRE(blocks_host != nullptr && out != nullptr, "null pointer");
  BlockPtrs bp{};
  for (int b = 0; b < n_blocks; ++b) {
    VS_REQUIRE(blocks_host[b] != nullptr, "null block pointer");
    bp.p[b] = blocks_host[b];
  }
  switch (n_blocks) {
    case 1: return launch_aggregate<1>(bp, out, rows, rows, channels, stream);
    case 2: return launch_aggregate<2>(bp, out, rows, rows, channels, stream);
    case 3: return launch_aggregate<3>(bp, out, rows, rows, channels, stream);
    default: return launch_aggregate<4>(bp, out, rows, rows, channels, stream);
  }
}

// The same arithmetic on an arbitrary row window [first_row, first_row + rows) of every block: the CFG-half split of the
// SVD UNet over two GPUs leaves the conditional half alone on one rank, i.e. first_row = 0.
VS_API int vidseg_aggregate_normalize_rows(const float* const* blocks_host, int n_blocks, long long first_row,
                                           long long rows_ll, int channels, float* out, void* stream) {
  VS_REQUIRE(n_blocks >= 1 && n_blocks <= 4, "n_blocks must be 1..4");
  VS_REQUIRE(first_row >= 0 && first_row < (1ll << 26) && rows_ll >= 0 && rows_ll < (1ll << 26) && channels >= 1, "bad shape");
  const int rows = (int)rows_ll;
  if (rows == 0) return 0;
  VS_REQUIRE(blocks_host != nullptr && out != nullptr, "null pointer");
  BlockPtrs bp{};
  for (int b = 0; b < n_blocks; ++b) {
    VS_REQUIRE(blocks_host[b] != nullptr, "null block pointer");
    bp.p[b] = blocks_host[b];
  }
  switch (n_blocks) {
    case 1: return launch_aggregate<1>(bp, out, rows, (int)first_row, channels, stream);
    case 2: return launch_aggregate<2>(bp, out, rows, (int)first_row, channels, stream);
    case 3: return launch_aggregate<3>(bp, out, rows, (int)first_row, channels, stream);
    default: return launch_aggregate<4>(bp, out, rows, (int)first_row, channels, stream);
  }
}
