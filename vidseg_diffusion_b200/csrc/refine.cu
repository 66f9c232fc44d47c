// R3: correspondence-based mask refinement on the device.
//
// Replaces scripts/sampling/feature_extraction.py:176-323 (dense_feature_matching_iterative),
// :326-364 (dense_tracking) and :367-461 (correct_low_res_mask).  Reference behaviour kept:
//   * one trajectory per feature cell, point p starts at cell p (meshgrid 'ij', :337-343);
//   * hop t -> t+1: cosine of the (once-normalised) source feature against every cell of frame
//     t+1 and of frame 0 ("aux"), blended t/(t+1)*cos + 1/(t+1)*cos_aux in fp32 (:290-291), row
//     arg-max (:293, any maximum on ties; here the lowest cell index);
//   * the target/aux maps are re-normalised once per 500-point batch (:272-274), so batch b sees
//     them normalised b+1 times: kept via `levels` of normalised maps;
//   * signed-jump filter (> 1 cell down/right only, :392-409), majority label per trajectory with
//     ties to the first label seen in frame order (Counter.most_common, :411-421), write-back in
//     ascending point order, i.e. the highest point index wins a contested cell.
// The reference round-trips every cosine map to the host for np.argpartition; here the arg-max is
// fused into the similarity kernel and positions never leave HBM.
#define VS_FAMILY vidseg::kFamRefine
#include "common.cuh"

namespace vidseg {

constexpr int kRfBatch = 500;  // feature_extraction.py:196
constexpr int kRfBM = 32;      // points per block
constexpr int kRfBN = 128;     // target cells per block
constexpr int kRfBK = 32;
constexpr int kRfMaxFrames = 64;

struct RfLayout {
  int levels;
  size_t lvl, best, winner, common, total;
};

static RfLayout rf_layout(int f, int hw, int c) {
  RfLayout L{};
  L.levels = hw / kRfBatch + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.lvl = take((size_t)L.levels * f * hw * c * 4);
  L.best = take((size_t)(f > 1 ? f - 1 : 1) * hw * 8);
  L.winner = take((size_t)f * hw * 4);
  L.common = take((size_t)hw * 4);
  L.total = off;
  return L;
}

// level 0 = x / ||x||, level l = level(l-1) / ||level(l-1)||  (fp32 IEEE division, norm via
// float64 accumulation rounded to fp32).  One warp per (frame, cell).
template <int MAXV>
__global__ void __launch_bounds__(256)
rf_normalize_levels_kernel(const float* __restrict__ feats, int num_frames, int hw, int channels, int levels,
                           float* __restrict__ lvl) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // f * hw + cell
  const int lane = threadIdx.x & 31;
  if (row >= num_frames * hw) return;
  const float* src = feats + ((size_t)num_frames * hw + row) * channels;  // conditional half
  float v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < channels) ? src[c] : 0.f;
  }
  for (int l = 0; l < levels; ++l) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) s = fma((double)v[i], (double)v[i], s);
    s = warp_sum(s);
    const float nrm = (float)sqrt(s);
    float* dst = lvl + ((size_t)l * num_frames * hw + row) * channels;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      v[i] = __fdiv_rn(v[i], nrm);
      if (c < channels) dst[c] = v[i];
    }
  }
}

__device__ __forceinline__ unsigned long long rf_pack(float score, int cell) {
  unsigned int b = __float_as_uint(score);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)cell);
}
__device__ __forceinline__ int rf_unpack_cell(unsigned long long key) {
  return (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull));
}

// one hop t -> t+1 for all trajectories.  grid = (point tiles, target splits).
__global__ void __launch_bounds__(256)
rf_match_kernel(const float* __restrict__ lvl, int num_frames, int hw, int channels, int t,
                const unsigned long long* __restrict__ best_prev, unsigned long long* __restrict__ best_out,
                float coef_t, float coef_aux) {
  __shared__ __align__(16) float qs[kRfBM][kRfBK + 4];
  __shared__ __align__(16) float ts[kRfBN][kRfBK + 4];
  __shared__ __align__(16) float as[kRfBN][kRfBK + 4];
  __shared__ int s_src[kRfBM];
  const int tiles_per_batch = (kRfBatch + kRfBM - 1) / kRfBM;
  const int batch = blockIdx.x / tiles_per_batch;
  const int tile = blockIdx.x - batch * tiles_per_batch;
  const int p0 = batch * kRfBatch + tile * kRfBM;
  const int p_end = min(min((batch + 1) * kRfBatch, hw), p0 + kRfBM);
  if (p0 >= p_end) return;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  if (tid < kRfBM) {
    const int p = p0 + tid;
    int cell = 0;
    if (p < p_end) cell = (t == 0) ? p : rf_unpack_cell(best_prev[p]);
    s_src[tid] = cell;
  }
  __syncthreads();
  const size_t frame_stride = (size_t)hw * channels;
  const float* q_base = lvl + (size_t)t * frame_stride;                                          // level 0, frame t
  const float* t_base = lvl + ((size_t)batch * num_frames + (t + 1)) * frame_stride;             // level `batch`
  const float* a_base = lvl + ((size_t)batch * num_frames + 0) * frame_stride;
  const int c0 = blockIdx.y * kRfBN;
  float acc_t[2][8], acc_a[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_t[i][j] = 0.f; acc_a[i][j] = 0.f; }
  for (int k0 = 0; k0 < channels; k0 += kRfBK) {
    __syncthreads();
    {
      // q tile: 32 points x 32 channels = 1024 elements, 4 per thread
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = tid + 256 * q;
        const int rr = e >> 5, cc = e & 31;
        const int gc = k0 + cc;
        qs[rr][cc] = (p0 + rr < p_end && gc < channels) ? q_base[(size_t)s_src[rr] * channels + gc] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int e = tid + 256 * q;
        const int rr = e >> 5, cc = e & 31;
        const int cell = c0 + rr, gc = k0 + cc;
        const bool ok = (cell < hw && gc < channels);
        ts[rr][cc] = ok ? t_base[(size_t)cell * channels + gc] : 0.f;
        as[rr][cc] = ok ? a_base[(size_t)cell * channels + gc] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kRfBK; kk += 4) {
      float4 qv[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) qv[i] = *reinterpret_cast<const float4*>(&qs[ty * 2 + i][kk]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 tv = *reinterpret_cast<const float4*>(&ts[tx + 16 * j][kk]);
        const float4 av = *reinterpret_cast<const float4*>(&as[tx + 16 * j][kk]);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          acc_t[i][j] = fmaf(qv[i].x, tv.x, acc_t[i][j]);
          acc_t[i][j] = fmaf(qv[i].y, tv.y, acc_t[i][j]);
          acc_t[i][j] = fmaf(qv[i].z, tv.z, acc_t[i][j]);
          acc_t[i][j] = fmaf(qv[i].w, tv.w, acc_t[i][j]);
          acc_a[i][j] = fmaf(qv[i].x, av.x, acc_a[i][j]);
          acc_a[i][j] = fmaf(qv[i].y, av.y, acc_a[i][j]);
          acc_a[i][j] = fmaf(qv[i].z, av.z, acc_a[i][j]);
          acc_a[i][j] = fmaf(qv[i].w, av.w, acc_a[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    unsigned long long key = 0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cell = c0 + tx + 16 * j;
      if (cell < hw) {
        // cos_map = t/(t+1) * cos_map + 1/(t+1) * cos_map_aux   (fp32, three roundings)
        const float s = __fadd_rn(__fmul_rn(coef_t, acc_t[i][j]), __fmul_rn(coef_aux, acc_a[i][j]));
        const unsigned long long kj = rf_pack(s, cell);
        key = kj > key ? kj : key;
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    const int p = p0 + ty * 2 + i;
    if (tx == 0 && p < p_end) atomicMax(&best_out[p], key);
  }
}

// one thread per trajectory: decode positions, spatial filter, majority label, claim cells.
__global__ void __launch_bounds__(256)
rf_vote_kernel(const unsigned long long* __restrict__ best, const int* __restrict__ labels_in, int num_frames, int hw,
               int width, int* __restrict__ traj_out, int* __restrict__ keep_out, int* __restrict__ common,
               int* __restrict__ winner) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  int lab[kRfMaxFrames];
  int pos_prev = p;
  bool keep = true;
  traj_out[p] = p;
  lab[0] = labels_in[p];
  for (int f = 1; f < num_frames; ++f) {
    const int pos = rf_unpack_cell(best[(size_t)(f - 1) * hw + p]);
    traj_out[(size_t)f * hw + p] = pos;
    const int dh = pos / width - pos_prev / width;
    const int dw = pos % width - pos_prev % width;
    if (dh > 1 || dw > 1) keep = false;  // signed: only down/right jumps are rejected (:399)
    lab[f] = labels_in[(size_t)f * hw + pos];
    pos_prev = pos;
  }
  keep_out[p] = keep ? 1 : 0;
  int best_label = lab[0], best_count = 0;
  for (int f = 0; f < num_frames; ++f) {
    bool seen = false;
    for (int g = 0; g < f; ++g) seen = seen || (lab[g] == lab[f]);
    if (seen) continue;
    int cnt = 0;
    for (int g = f; g < num_frames; ++g) cnt += (lab[g] == lab[f]);
    if (cnt > best_count) { best_count = cnt; best_label = lab[f]; }  // strict: first seen wins ties
  }
  common[p] = best_label;
  if (keep) {
    atomicMax(&winner[p], p);  // frame 0
    for (int f = 1; f < num_frames; ++f) atomicMax(&winner[(size_t)f * hw + traj_out[(size_t)f * hw + p]], p);
  }
}

__global__ void rf_apply_kernel(const int* __restrict__ labels_in, const int* __restrict__ winner,
                                const int* __restrict__ common, int total, int* __restrict__ labels_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = winner[i];
  labels_out[i] = (w >= 0) ? common[w] : labels_in[i];
}

}  // namespace vidseg

using namespace vidseg;

VS_API size_t vidseg_refine_workspace_bytes(int num_frames, int hw, int channels) {
  if (num_frames <= 0 || hw <= 0 || channels <= 0) return 0;
  return rf_layout(num_frames, hw, channels).total;
}

VS_API int vidseg_refine_masks(const float* feats, const int32_t* labels_in, int num_frames, int height, int width,
                                   int channels, int32_t* traj_out, int32_t* keep_out, int32_t* labels_out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(feats && labels_in && traj_out && keep_out && labels_out && workspace, "null pointer");
  VS_REQUIRE(num_frames >= 1 && num_frames <= kRfMaxFrames, "num_frames must be 1..64");
  VS_REQUIRE(height >= 1 && width >= 1 && channels >= 1 && channels <= 2048, "bad shape");
  const int hw = height * width;
  const RfLayout L = rf_layout(num_frames, hw, channels);
  if (workspace_bytes < L.total)
    return set_error(VIDSEG_E_WORKSPACE, "%s: need %lld bytes, got %lld", "refine workspace", (long long)L.total, (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  float* lvl = reinterpret_cast<float*>(ws + L.lvl);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(ws + L.best);
  int* winner = reinterpret_cast<int*>(ws + L.winner);
  int* common = reinterpret_cast<int*>(ws + L.common);
  VS_CHECK_CUDA(cudaMemsetAsync(best, 0, (size_t)(num_frames > 1 ? num_frames - 1 : 1) * hw * 8, st));
  VS_CHECK_CUDA(cudaMemsetAsync(winner, 0xFF, (size_t)num_frames * hw * 4, st));
  const int rows = num_frames * hw;
  const int grid_n = (rows * 32 + 255) / 256;
  if (channels <= 32 * 4) {
    VS_LAUNCH((rf_normalize_levels_kernel<4>), grid_n, 256, 0, st, feats, num_frames, hw, channels, L.levels, lvl);
  } else if (channels <= 32 * 20) {
    VS_LAUNCH((rf_normalize_levels_kernel<20>), grid_n, 256, 0, st, feats, num_frames, hw, channels, L.levels, lvl);
  } else {
    VS_LAUNCH((rf_normalize_levels_kernel<64>), grid_n, 256, 0, st, feats, num_frames, hw, channels, L.levels, lvl);
  }
  VS_POST_LAUNCH();
  const int tiles_per_batch = (kRfBatch + kRfBM - 1) / kRfBM;
  dim3 grid(L.levels * tiles_per_batch, (hw + kRfBN - 1) / kRfBN);
  for (int t = 0; t + 1 < num_frames; ++t) {
    // python: t / (t + 1) and 1 / (t + 1) are doubles, cast to fp32 when they meet the fp32 maps
    const float coef_t = (float)((double)t / (double)(t + 1));
    const float coef_aux = (float)(1.0 / (double)(t + 1));
    VS_LAUNCH(rf_match_kernel, grid, 256, 0, st, lvl, num_frames, hw, channels, t,
              t == 0 ? (const unsigned long long*)nullptr : best + (size_t)(t - 1) * hw, best + (size_t)t * hw, coef_t,
              coef_aux);
    VS_POST_LAUNCH();
  }
  VS_LAUNCH(rf_vote_kernel, (hw + 255) / 256, 256, 0, st, best, labels_in, num_frames, hw, width, traj_out, keep_out,
            common, winner);
  VS_POST_LAUNCH();
  VS_LAUNCH(rf_apply_kernel, (rows + 255) / 256, 256, 0, st, labels_in, winner, common, rows, labels_out);
  VS_POST_LAUNCH();
  return 0;
}
