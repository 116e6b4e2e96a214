// Internal interface of the GPU NMS (nms.cu).
#pragma once
#include "common.h"

namespace frcnn {

static constexpr int NMS_CTA_MAX_SEG = 8192;  // largest segment of the CTA-level path (the only one that takes device-side counts)

struct NmsState {
  int* seg_beg;    // [n_seg] first row of segment s
  int* seg_len;    // [n_seg] number of rows of segment s
  int* order;      // [n] priority position -> segment-local box index
  uint8_t* alive;  // [n] per priority position
  int* cursor;     // [n_seg] first undecided priority position
  int* counts;     // [n_seg] number of picks so far
  int* sel_cnt;    // [n_seg] size of this round's selection
  int* newk_cnt;   // [n_seg] keepers found this round
  int* sel_pos;    // [n_seg][B]
  float4* sel_box;
  float* sel_area;
  float4* newk_box;
  float* newk_area;
  uint32_t* mask;  // [n_seg][B][B/32]
  int* pick;       // [n] segment-local picks in pick order, written at seg_beg[s] + k
  int* remaining;  // != 0 while some candidate is still alive behind a cursor
};

struct NmsWorkspace {
  NmsState st;
  uint32_t* rkeys[2];
  uint32_t* rvals[2];
  uint8_t* rseg[2];
  uint32_t* rhist;
  int cap_total, cap_seg;
  const int* seg_counts = nullptr;  // consumed by the next nms_run (see nms_set_segments_from_counts)
  int seg_stride = 0;
  bool fused = false;               // next nms_run: all segments are small -> one fused launch
};

size_t nms_workspace_bytes(int cap_total, int cap_seg);
void nms_workspace_init(NmsWorkspace* ws, void* mem, size_t bytes, int cap_total, int cap_seg);
int nms_run(NmsWorkspace* ws, const float* boxes_dev, int row_stride, int n_seg, int n_total_cap, int max_seg_len, float thr,
            int order_mode, int order_col, cudaStream_t st, int* h_remaining_pinned);
void nms_export(NmsWorkspace* ws, int n_seg, int max_seg_len, int64_t* pick64, int64_t* counts64, cudaStream_t st);
void nms_set_segments_from_counts(NmsWorkspace* ws, const int* counts_dev, int n_seg, int stride, cudaStream_t st);

}  // namespace frcnn
