// Anchor labelling kernels (label_kernels.cu): Anchors:findPositive / sampleNegative (Anchors.lua:86-235).
#pragma once
#include "common.h"

namespace frcnn {

static constexpr int LUT_CELLS = 200;      // Anchors.lua:15: the LUTs hold 200 cells per (scale, aspect)
static constexpr int MAX_LABEL_IJ = 12;    // 4 scales x 3 aspects (Anchors.lua:98-99)

struct FindPositiveParams {
  const float *w_lut, *h_lut;   // [n_scales][3][200][2] fp32 (Anchors.lua:18-19)
  int n_scales;
  const double* rois;           // [n_rois][4] {minX, minY, maxX, maxY}
  double clip[4];
  int has_clip;
  double pos_threshold, neg_threshold;
  int include_best;
  frcnn_anchor_ref* out;        // [n_rois][cap_per_roi]
  frcnn_anchor_ref* best_scratch;
  int cap_per_roi;
  int* n_out;                   // [n_rois]
  int* status;                  // set to 1 when a ROI has more matches than cap_per_roi
};
void launch_find_positive(const FindPositiveParams& p, int n_rois, cudaStream_t st);

struct SampleNegativeParams {
  const float *w_lut, *h_lut;
  int n_scales;
  double image_rect[4];
  const double* rois;
  int n_rois;
  double neg_threshold;
  int count;
  const uint32_t* rnd;          // 3 * n_trials values of torch.random()
  int n_trials;
  frcnn_anchor_ref* out;
  int cap;
  int retry_in;                 // consecutive rejections carried over from a previous call of the same loop
  int* result;                  // {n_out, trials consumed, stopping rule fired, #ranges, retry at the end}
};
void launch_sample_negative(const SampleNegativeParams& p, cudaStream_t st);

// Anchors:findNearby + the nearby-aversion filter of BatchIterator.lua:206-217
struct FindNearbyParams {
  const float *w_lut, *h_lut;   // [n_scales][3][200][2]
  const double *cen_x, *cen_y;  // [n_scales][200] cell centres (Anchors.lua:40-41,49-50), doubles
  int n_scales;
  const frcnn_anchor_ref* pos;  // [n_pos] the positive anchors p[1]
  int n_pos;
  double neg_threshold;
  frcnn_anchor_ref* out;        // [cap]
  int* out_pos;                 // [cap] 0-based index of the positive each entry belongs to
  int cap;
  int* result;                  // {n_out (may exceed cap: only cap are written)}
};
void launch_find_nearby(const FindNearbyParams& p, cudaStream_t st);

}  // namespace frcnn
