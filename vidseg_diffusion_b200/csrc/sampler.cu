// One Euler step of the EDM sampler on the latent, fused.
//
// Replaces, for one call of EDMSampler.sampler_step + the latent blending of EulerEDMSampler.__call__
// (reference sgm/modules/diffusionmodules/sampling.py:102-132 and :229-250), the chain of ~14 elementwise torch
// kernels that runs between two UNet evaluations:
//   Denoiser.forward        net * c_out + input * c_skip                       denoiser.py:41-48
//   VanillaCFG / LinearPredictionGuider     x_u + scale * (x_c - x_u)          guiders.py:28-31, 81-90
//   to_d                    (x - denoised) / sigma_hat                         sampling_utils.py:34-35
//   euler_step              x + dt * d,   dt = next_sigma - sigma_hat          sampling.py:92-93, 127-128
//   latent blending         x * mask + ori_xt * (1 - mask), mask = nearest-neighbour upsampling of the
//                           per-frame feature mask                             sampling.py:231-250
// Every operation is the same IEEE fp32 operation torch issues (round-to-nearest mul / add / sub / div, no
// contraction into FMAs), in the same order, so the step is bit-identical to the reference's eager chain given the
// same network output.  HBM-bound: reads x, the (doubled) network output and, when blending, ori_xt once; writes x once.
#define VS_FAMILY vidseg::kFamElementwise
#include "common.cuh"

namespace vidseg {

struct SamplerStepParams {
  const float* x;         // [B, C, H, W] latent entering the step (already noised: the network saw x * c_in)
  const float* net;       // [G*B, C, H, W] network output, unconditional half first when G == 2
  const float* c_skip;    // [G*B]
  const float* c_out;     // [G*B]
  const float* scale;     // [B] guidance scale per sample (G == 2), else unused
  const float* sigma_hat; // [B]
  const float* sigma_next;// [B]
  const void* mask;       // [B, fh, fw] fp32 or fp64 (mask_f64), or null: latent blending
  const float* ori;       // [B, C, H, W] the source run's latent of this step (with mask)
  float* out;             // [B, C, H, W]
  int b, chw, hw, w, h, fh, fw, guided, mask_f64;
};

__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerStepParams p) {
  const long long total = (long long)p.b * p.chw;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float sh = (float)p.fh / (float)p.h, sw = (float)p.fw / (float)p.w;   // nearest: src = min(int(dst * in/out), in-1)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int s = (int)(i / p.chw);
    const float xin = p.x[i];
    float den = __fadd_rn(__fmul_rn(p.net[i], p.c_out[s]), __fmul_rn(xin, p.c_skip[s]));
    if (p.guided) {
      const int s2 = s + p.b;
      const float dc = __fadd_rn(__fmul_rn(p.net[i + total], p.c_out[s2]), __fmul_rn(xin, p.c_skip[s2]));
      den = __fadd_rn(den, __fmul_rn(p.scale[s], __fsub_rn(dc, den)));
    }
    const float sg = p.sigma_hat[s];
    const float d = __fdiv_rn(__fsub_rn(xin, den), sg);
    const float dt = __fsub_rn(p.sigma_next[s], sg);
    float xn = __fadd_rn(xin, __fmul_rn(dt, d));
    if (p.mask) {
      const int pix = (int)(i % p.hw);
      const int yy = pix / p.w, xx = pix - yy * p.w;
      const int my = min((int)((float)yy * sh), p.fh - 1), mx = min((int)((float)xx * sw), p.fw - 1);
      const size_t mi = ((size_t)s * p.fh + my) * p.fw + mx;
      if (p.mask_f64) {
        // the reference's masks are float64 (numpy / 255.0): torch promotes the blend to float64, `.float()` rounds once
        const double m = static_cast<const double*>(p.mask)[mi];
        xn = (float)__dadd_rn(__dmul_rn((double)xn, m), __dmul_rn((double)p.ori[i], __dsub_rn(1.0, m)));
      } else {
        const float m = static_cast<const float*>(p.mask)[mi];
        xn = __fadd_rn(__fmul_rn(xn, m), __fmul_rn(p.ori[i], __fsub_rn(1.0f, m)));
      }
    }
    p.out[i] = xn;
  }
}

}  // namespace vidseg

using namespace vidseg;

VS_API int vidseg_sampler_step(const float* x, const float* net, const float* c_skip, const float* c_out,
                               const float* scale, const float* sigma_hat, const float* sigma_next, const void* mask,
                               int mask_is_f64, const float* ori_xt, float* out, int batch, int channels, int height,
                               int width, int mask_h, int mask_w, int guided, void* stream) {
  VS_REQUIRE(x && net && c_skip && c_out && sigma_hat && sigma_next && out, "null pointer");
  VS_REQUIRE(batch >= 0 && channels >= 1 && height >= 1 && width >= 1, "bad shape");
  VS_REQUIRE(guided == 0 || guided == 1, "guided must be 0 or 1");
  VS_REQUIRE(!guided || scale != nullptr, "a guided step needs the per-sample scale");
  VS_REQUIRE((mask == nullptr) == (ori_xt == nullptr), "mask and ori_xt go together");
  VS_REQUIRE(mask == nullptr || (mask_h >= 1 && mask_w >= 1), "bad mask shape");
  if (batch == 0) return 0;
  SamplerStepParams p{};
  p.x = x; p.net = net; p.c_skip = c_skip; p.c_out = c_out; p.scale = scale; p.sigma_hat = sigma_hat;
  p.sigma_next = sigma_next; p.mask = mask; p.ori = ori_xt; p.out = out;
  p.b = batch; p.hw = height * width; p.chw = channels * p.hw; p.w = width; p.h = height;
  p.fh = mask ? mask_h : 1; p.fw = mask ? mask_w : 1; p.guided = guided; p.mask_f64 = mask_is_f64 ? 1 : 0;
  const long long total = (long long)batch * p.chw;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 8);
  const double bytes = (double)total * 4.0 * (2.0 + (guided ? 2.0 : 1.0) + (mask ? 1.0 : 0.0));
  VS_LAUNCH_W(bytes, sampler_step_kernel, blocks, 256, 0, stream, p);
  VS_POST_LAUNCH();
  return 0;
}
