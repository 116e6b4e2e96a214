/* C restatement of nms.lua:23-102 (the CPU path of the reference: TH FloatTensor ops).
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never linked into the product library.
 *
 * Follows the reference pick by pick: ascending sort of the key, pop the last element, recompute the
 * +1-pixel IoU of every remaining box against it in fp32 with the reference's operation order
 * (nms.lua:35,85-94), keep IoU <= overlap (nms.lua:96).  Tie order of TH's unstable quicksort is PARITY
 * UNPINNED; defined here as stable ascending by (key, index).
 * Build: see oracle/Makefile (gcc -O3 -pthread -ffp-contract=off).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float key; int64_t idx; } kv_t;

/* number of pair-IoU evaluations the reference algorithm performed (every pick tests all remaining boxes, nms.lua:72-96)
 * since the last reset: the unit of SURVEY 8d's "pair-IoU/s" */
static int64_t g_pairs = 0;
int64_t oracle_nms_pairs(int reset) {
  int64_t v = __atomic_load_n(&g_pairs, __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&g_pairs, 0, __ATOMIC_RELAXED);
  return v;
}

static int cmp_kv(const void *a, const void *b) {
  const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
  if (x->key < y->key) return -1;
  if (x->key > y->key) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* order_mode: 0 = y2 (nms.lua:41-42), 1 = area (nms.lua:39-40), 2 = column order_col (0-based; nms.lua:37-38) */
int64_t oracle_nms(const float *boxes, int64_t n, int64_t row_stride, float overlap, int order_mode, int order_col,
                   int64_t *pick) {
  if (n <= 0) return 0;
  float *area = (float *)malloc(sizeof(float) * (size_t)n);
  kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)n);
  int64_t *I = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  for (int64_t j = 0; j < n; ++j) {
    const float *b = boxes + j * row_stride;
    float w = (b[2] - b[0]) + 1.0f, h = (b[3] - b[1]) + 1.0f; /* nms.lua:35 */
    area[j] = w * h;
    kv[j].key = order_mode == 1 ? area[j] : (order_mode == 2 ? b[order_col] : b[3]);
    kv[j].idx = j;
  }
  qsort(kv, (size_t)n, sizeof(kv_t), cmp_kv); /* nms.lua:45 */
  for (int64_t j = 0; j < n; ++j) I[j] = kv[j].idx;
  int64_t m = n, count = 0, pairs = 0;
  while (m > 0) { /* nms.lua:58-97 */
    int64_t i = I[m - 1];
    pick[count++] = i;
    if (m == 1) break;
    --m;
    const float *bi = boxes + i * row_stride;
    const float x1i = bi[0], y1i = bi[1], x2i = bi[2], y2i = bi[3], ai = area[i];
    int64_t out = 0;
    pairs += m;
    for (int64_t t = 0; t < m; ++t) {
      int64_t j = I[t];
      const float *b = boxes + j * row_stride;
      float xx1 = b[0] > x1i ? b[0] : x1i, yy1 = b[1] > y1i ? b[1] : y1i;
      float xx2 = b[2] < x2i ? b[2] : x2i, yy2 = b[3] < y2i ? b[3] : y2i;
      float w = (xx2 - xx1) + 1.0f, h = (yy2 - yy1) + 1.0f; /* nms.lua:85-86 */
      w = w > 0.0f ? w : 0.0f;
      h = h > 0.0f ? h : 0.0f;
      float inter = w * h;
      float iou = inter / ((area[j] + ai) - inter); /* nms.lua:93-94 */
      if (iou <= overlap) I[out++] = j;              /* nms.lua:96 */
    }
    m = out;
  }
  free(area); free(kv); free(I);
  __atomic_fetch_add(&g_pairs, pairs, __ATOMIC_RELAXED);
  return count;
}

/* Per-class NMS as Detector.lua:125-136 runs it; segments are independent, so all host cores can be used
 * (pthreads; this image's gcc has no libgomp).  pick receives segment-local indices at pick[seg_offsets[s] ...];
 * counts[s] = number of picks of segment s. */
#include <pthread.h>
typedef struct {
  const float *boxes; int64_t row_stride; const int64_t *seg_offsets; int n_seg; float overlap;
  int order_mode, order_col; int64_t *pick, *counts; int next;
} seg_job_t;

static void *seg_worker(void *arg) {
  seg_job_t *j = (seg_job_t *)arg;
  for (;;) {
    int s = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (s >= j->n_seg) break;
    int64_t a = j->seg_offsets[s], b = j->seg_offsets[s + 1];
    j->counts[s] = oracle_nms(j->boxes + a * j->row_stride, b - a, j->row_stride, j->overlap, j->order_mode,
                              j->order_col, j->pick + a);
  }
  return 0;
}

void oracle_nms_segmented(const float *boxes, int64_t row_stride, const int64_t *seg_offsets, int n_seg,
                          float overlap, int order_mode, int order_col, int64_t *pick, int64_t *counts, int threads) {
  seg_job_t job = {boxes, row_stride, seg_offsets, n_seg, overlap, order_mode, order_col, pick, counts, 0};
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256];
  for (int t = 1; t < threads; ++t) pthread_create(&th[t], 0, seg_worker, &job);
  seg_worker(&job);
  for (int t = 1; t < threads; ++t) pthread_join(th[t], 0);
}
