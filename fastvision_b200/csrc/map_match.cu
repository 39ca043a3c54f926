// K5: mAP per-image matcher for a whole batch of images in one launch (one CTA per image).
// Replaces CalculateMAP.process_one (metrics/map.py:16-83): pairwise IoU on device, D2H, numpy
// argsort/unique per image.  Deterministic rule equivalent to the reference's dedupe:
//   t*(p)  = arg-max-IoU target among (IoU > thr[0] in fp32  and  same class), lowest t on ties;
//   winner = lowest-index p with t*(p) = t;
//   correct[p,k] = winner and (double)iou > thr[k].
#include "common.cuh"

namespace fvb {

constexpr int kMapThreads = 128;
constexpr int kMaxThr = 16;

struct MapParams {
  const float* dets;
  const int* det_off;
  const float* gts;
  const int* gt_off;
  double thr[kMaxThr];
  int n_thr;
  unsigned char* correct;
  int* ws_best;  // [sum M] scratch: best target of each detection (global, sized like dets)
};

constexpr int kMapGtCap = 1024;  // targets of one image staged in shared memory (20 KB); beyond: the global-memory path

// One CTA per image.  The image's targets are staged in shared memory once (every detection walks all of them); the dedupe
// "lowest-index p with t*(p) = t wins" is an atomicMin per target instead of a serial scan over the earlier detections.
__global__ void __launch_bounds__(kMapThreads) map_match_kernel(const MapParams p, float* ws_iou) {
  __shared__ float4 s_box[kMapGtCap];
  __shared__ float s_cls[kMapGtCap];
  __shared__ int s_win[kMapGtCap];
  const int img = blockIdx.x;
  const int d0 = p.det_off[img], M = p.det_off[img + 1] - d0;
  const int g0 = p.gt_off[img], N = p.gt_off[img + 1] - g0;
  if (M <= 0) return;
  const float thr0 = (float)p.thr[0];  // torch compares the fp32 IoU tensor with the scalar folded to fp32 (map.py:51)
  const bool staged = N <= kMapGtCap;
  if (staged) {
    for (int t = threadIdx.x; t < N; t += kMapThreads) {
      const float* g = p.gts + (size_t)(g0 + t) * 5;
      s_cls[t] = g[0];
      s_box[t] = make_float4(g[1], g[2], g[3], g[4]);
      s_win[t] = 0x7fffffff;
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < M; q += kMapThreads) {
    const float* d = p.dets + (size_t)(d0 + q) * 6;
    const float dcls = d[0];
    Box pb;
    pb.x1 = d[2]; pb.y1 = d[3]; pb.x2 = d[4]; pb.y2 = d[5];
    float best = -1.0f;
    int bt = -1;
    for (int t = 0; t < N; ++t) {
      float gc;
      Box tb;
      if (staged) {
        gc = s_cls[t];
        const float4 b4 = s_box[t];
        tb.x1 = b4.x; tb.y1 = b4.y; tb.x2 = b4.z; tb.y2 = b4.w;
      } else {
        const float* g = p.gts + (size_t)(g0 + t) * 5;
        gc = g[0];
        tb.x1 = g[1]; tb.y1 = g[2]; tb.x2 = g[3]; tb.y2 = g[4];
      }
      if (gc != dcls) continue;  // map.py:54 class equality on floats
      float iou = iou_plain<false>(tb, pb, 1e-7f);  // cal_iou_batch(target, predict), map.py:50
      if (iou > thr0 && iou > best) {
        best = iou;
        bt = t;
      }
    }
    p.ws_best[d0 + q] = bt;
    ws_iou[d0 + q] = best;
    if (staged && bt >= 0) atomicMin(&s_win[bt], q);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < M; q += kMapThreads) {
    const int bt = p.ws_best[d0 + q];
    bool win = bt >= 0;
    if (staged) {
      win = win && s_win[bt] == q;
    } else {
      for (int q2 = 0; win && q2 < q; ++q2)
        if (p.ws_best[d0 + q2] == bt) win = false;
    }
    const double iou = (double)ws_iou[d0 + q];
    unsigned char* c = p.correct + (size_t)(d0 + q) * p.n_thr;
    for (int k = 0; k < p.n_thr; ++k) c[k] = (win && iou > p.thr[k]) ? 1 : 0;  // map.py:81 (numpy: fp32 vs f64)
  }
}

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_map_match_workspace_bytes(int64_t total_dets) { return (size_t)total_dets * 8 + 512; }

extern "C" int fvb_map_match_f32(const float* d_dets, const int32_t* d_det_off, const float* d_gts, const int32_t* d_gt_off,
                                 int images, int64_t total_dets, const double* thresholds, int n_thr, uint8_t* d_correct,
                                 void* d_ws, void* stream) {
  FVB_REQUIRE(images >= 0 && n_thr >= 1 && n_thr <= kMaxThr, "map_match: images=%d n_thr=%d (max %d)", images, n_thr, kMaxThr);
  FVB_REQUIRE(thresholds != nullptr, "map_match: thresholds NULL");
  if (images == 0 || total_dets == 0) return FVB_OK;
  FVB_REQUIRE(d_dets && d_det_off && d_gt_off && d_correct && d_ws, "map_match: NULL pointer");
  FVB_REQUIRE(((uintptr_t)d_ws & 255) == 0, "map_match: workspace must be 256-byte aligned");
  MapParams p;
  p.dets = d_dets;
  p.det_off = d_det_off;
  p.gts = d_gts;
  p.gt_off = d_gt_off;
  for (int k = 0; k < n_thr; ++k) p.thr[k] = thresholds[k];
  p.n_thr = n_thr;
  p.correct = d_correct;
  p.ws_best = (int*)d_ws;
  float* ws_iou = (float*)((unsigned char*)d_ws + (((size_t)total_dets * 4 + 255) / 256) * 256);
  map_match_kernel<<<images, kMapThreads, 0, (cudaStream_t)stream>>>(p, ws_iou);
  count_launch();
  return check_launch("map_match_kernel");
}
