// Fused optimiser step on the flat parameter buffer: replaces `gradient:div(cls_count)` (objective.lua:200) followed by
// optim.rmsprop(eval_objective_grad, weights, rmsprop_state) (main.lua:122,133) -- in the reference four full passes
// of TH vector ops over the 26.8 M-float `weights` / `gradient` / state buffers, here ONE pass: 12 bytes read and
// 8 bytes written per parameter, + 4 for the divided gradient the caller still sees (HBM-bound; 643 MB per step for
// vgg_small's 26.8 M parameters).
//
// optim.rmsprop is an un-vendored dependency; restated from its published algorithm (optim/rmsprop.lua,
// 2015): [dfdx += wd * x]; m = alpha * m + (1 - alpha) * dfdx^2; tmp = sqrt(m) + epsilon; x += -lr * dfdx / tmp, every
// statement a separate fp32 TH vector op.  The arithmetic below reproduces exactly that sequence of individually
// rounded fp32 operations (no FMA contraction), so the result is bit-identical to the CPU restatement in
// oracle/optim.py.
#include "common.h"

namespace frcnn {

__global__ void __launch_bounds__(256) rmsprop_step_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m, long n,
                                                           float grad_div, float neg_lr, float alpha, float one_minus_alpha, float eps, float wd) {
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 wv = *reinterpret_cast<const float4*>(w + i);
      float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<const float4*>(m + i);
      float* wp = reinterpret_cast<float*>(&wv);
      float* gp = reinterpret_cast<float*>(&gv);
      float* mp = reinterpret_cast<float*>(&mv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d = grad_div != 1.0f ? __fdiv_rn(gp[k], grad_div) : gp[k];           // gradient:div(cls_count)
        if (wd != 0.0f) d = __fadd_rn(d, __fmul_rn(wd, wp[k]));                    // dfdx:add(wd, x)
        const float mm = __fadd_rn(__fmul_rn(mp[k], alpha), __fmul_rn(__fmul_rn(one_minus_alpha, d), d));  // m:mul(alpha):addcmul(1-alpha, dfdx, dfdx)
        const float t = __fadd_rn(__fsqrt_rn(mm), eps);                           // tmp:sqrt(m):add(epsilon)
        wp[k] = __fadd_rn(wp[k], __fdiv_rn(__fmul_rn(neg_lr, d), t));              // x:addcdiv(-lr, dfdx, tmp)
        mp[k] = mm;
        gp[k] = d;
      }
      *reinterpret_cast<float4*>(w + i) = wv;
      *reinterpret_cast<float4*>(m + i) = mv;
      if (grad_div != 1.0f || wd != 0.0f) *reinterpret_cast<float4*>(g + i) = gv;   // the caller sees gradient:div / dfdx:add, as in Lua
    } else {
      for (long j = i; j < n; ++j) {
        float d = grad_div != 1.0f ? __fdiv_rn(g[j], grad_div) : g[j];
        if (wd != 0.0f) d = __fadd_rn(d, __fmul_rn(wd, w[j]));
        const float mm = __fadd_rn(__fmul_rn(m[j], alpha), __fmul_rn(__fmul_rn(one_minus_alpha, d), d));
        const float t = __fadd_rn(__fsqrt_rn(mm), eps);
        w[j] = __fadd_rn(w[j], __fdiv_rn(__fmul_rn(neg_lr, d), t));
        m[j] = mm;
        if (grad_div != 1.0f || wd != 0.0f) g[j] = d;
      }
    }
  }
}

// The scalar arguments are Lua doubles in the reference and reach the TH vector ops as `real` (float) values: -lr,
// alpha, (1.0 - alpha) evaluated in double and THEN rounded, epsilon, wd, cls_count.
void launch_rmsprop_step(float* w, float* g, float* m, long n, double grad_div, double lr, double alpha, double eps, double wd, int num_sms,
                         cudaStream_t st) {
  FRCNN_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(m) & 15) == 0,
                FRCNN_E_INVALID, "rmsprop: the flat buffers must be 16-byte aligned");
  const long vec = (n + 3) / 4;
  const int blocks = (int)std::min<long>((vec + 255) / 256, (long)num_sms * 8);   // grid = a multiple of the SM count
  rmsprop_step_kernel<<<std::max(blocks, 1), 256, 0, st>>>(w, g, m, n, (float)grad_div, (float)(-lr), (float)alpha, (float)(1.0 - alpha),
                                                           (float)eps, (float)wd);
  FRCNN_CUDA_TRY(cudaGetLastError());
}

}  // namespace frcnn
