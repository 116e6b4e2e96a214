// Internal interface of the training-side kernels (train_kernels.cu).
#pragma once
#include <algorithm>

#include "common.h"

namespace frcnn {

void launch_dropout_mask(float* mask, int n, float p, uint64_t seed, uint32_t layer, cudaStream_t st);
void launch_unpool_prelu_bwd(const float* g, const uint8_t* arg, const bf16* yp, const float* slope, const float* mask, bf16* dpre,
                             float* dbias, float* dslope, int N, int H, int W, int C, int num_sms, cudaStream_t st);
void launch_prelu_bwd(bf16* d, const bf16* y, const float* slope, const float* mask, float* dbias, float* dslope, int N, int H, int W,
                      int C, int num_sms, cudaStream_t st);

struct HeadTailBwd {
  const float* d_out;   // [N][18][HW] fp32 (Torch layout)
  const float* ws;      // [splits][npix][256] forward split-K slices of the k x k conv
  int splits;
  long slice_stride, npix;
  int HW;
  const float *bias, *prelu, *w2;
  bf16* dpre;           // [npix][256] gradient wrt the k x k conv's pre-activation output
  float *dw2, *db2, *db1, *dslope;
};
void launch_head_tail_bwd(const HeadTailBwd& H, int num_sms, cudaStream_t st);
void launch_first_wgrad(const bf16* dpre, const float* img, float* dw, int N, int H, int W, int pad, int num_sms, cudaStream_t st);
void launch_add_chw_to_nhwc(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st);
void launch_check_slopes(const float* const* slopes_dev, int n, int* flag, cudaStream_t st);

}  // namespace frcnn
