// The box post-processing between decode and NMS of the demos' `postProcess` (second half):
// demos/yolov3_u/inference.py:92-109 == demos/yolov3_huaweiShip/inference.py:112-129 -- undo the letterbox, clamp to the
// original image, drop boxes not larger than 5 px, xywh -> xyxy, clamp again.  One fused pass over the 5 head channels of
// every decoded row instead of ~20 ATen launches (each a strided read-modify-write of a column) and a boolean-mask copy.
// Dropped rows stay in place with objectness -1, which no confidence threshold >= 0 passes: the relative order of the
// surviving rows -- the tie-break order of the NMS that follows -- is the reference's.
#include "common.cuh"

namespace fvb {

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

__global__ void demo_boxes_kernel(float* rows, long long n, int K, float pad_left, float pad_top, float ratio, float ori_w,
                                  float ori_h, float min_wh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* r = rows + i * K;
  float x = (r[0] - pad_left) / ratio, y = (r[1] - pad_top) / ratio, w = r[2] / ratio, h = r[3] / ratio;  // :92-95
  x = clampf(x, 0.0f, ori_w - 1.0f);                                                                       // :97-100
  y = clampf(y, 0.0f, ori_h - 1.0f);
  w = clampf(w, 0.0f, ori_w);
  h = clampf(h, 0.0f, ori_h);
  const bool keep = w > min_wh && h > min_wh;                                                              // :102-103
  const Box b = xywh_to_xyxy(x, y, w, h);                                                                  // :105
  r[0] = clampf(b.x1, 0.0f, ori_w - 1.0f);                                                                 // :106-109
  r[1] = clampf(b.y1, 0.0f, ori_h - 1.0f);
  r[2] = clampf(b.x2, 0.0f, ori_w - 1.0f);
  r[3] = clampf(b.y2, 0.0f, ori_h - 1.0f);
  if (!keep) r[4] = -1.0f;
}

}  // namespace fvb

using namespace fvb;

extern "C" int fvb_demo_boxes_postprocess_f32(float* d_rows, int64_t n_rows, int channels, float pad_left, float pad_top,
                                              float resize_ratio, float ori_width, float ori_height, float min_wh,
                                              void* stream) {
  FVB_REQUIRE(n_rows >= 0 && channels >= 5, "demo_boxes_postprocess: n_rows=%lld channels=%d", (long long)n_rows, channels);
  if (n_rows == 0) return FVB_OK;
  FVB_REQUIRE(d_rows != nullptr, "demo_boxes_postprocess: NULL pointer");
  demo_boxes_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_rows, n_rows, channels, pad_left, pad_top,
                                                                                          resize_ratio, ori_width, ori_height, min_wh);
  count_launch();
  return check_launch("demo_boxes_kernel");
}
