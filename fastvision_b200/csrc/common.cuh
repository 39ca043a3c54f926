// Shared device/host helpers for libfvb200 (sm_100a).  Compiled with -fmad=false: the parity spec
// is the reference's fp32 op order, so no multiply-add contraction anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/fvb200.h"

namespace fvb {

// ---- host-side error plumbing --------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaPeekAtLastError -> FVB_OK / FVB_E_CUDA
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define FVB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      fvb::set_error(__VA_ARGS__);    \
      return FVB_E_INVALID;           \
    }                                 \
  } while (0)

// ---- geometry passed to kernels by value -----------------------------------------------------------
struct Geom {
  int L, B, A, K;
  int H[FVB_MAX_LEVELS], W[FVB_MAX_LEVELS];
  int HW[FVB_MAX_LEVELS];
  int row_off[FVB_MAX_LEVELS + 1];  // first row of each level inside one image; row_off[L] = N
  float stride[FVB_MAX_LEVELS];
  float aw[FVB_MAX_LEVELS][FVB_MAX_ANCHORS];  // pixels
  float ah[FVB_MAX_LEVELS][FVB_MAX_ANCHORS];
  const float* head[FVB_MAX_LEVELS];
  int nchw;  // heads are [B,A*K,H,W] (decode only)
};

int make_geom(const fvb_yolo_geom* g, const float* const* d_heads, Geom* out);

static const int kDecodeThreads = 640;
// Shared memory the persistent decode kernel may take per SM: leaves room for one NMS CTA (~112 KB) beside it.
static const size_t kDecodeSmemBudget = 112 * 1024;

// Launch shape of the decode kernel (decode.cu); also fixes the layout of the fused objectness-BCE partials:
// one double per tile, level-major ([l][b][tile of the level]); tile = decode_tile_rows(K) rows of one segment.
struct DecodeShape {
  int tile_rows, tile_floats;
  int tiles_level_end[FVB_MAX_LEVELS];
  int tiles_per_image;
  long long total_tiles;
  int warps_per_cta, grid, stages;
  size_t smem_bytes;
};
int decode_launch_shape(const Geom& g, DecodeShape* s);
long long* debug_trace_ptr();  // nms.cu: the buffer of fvb_debug_set_nms_trace (NULL unless a tool set it)
int decode_tile_rows(int K);

// host+device for the pure arithmetic: tests/host_check.cu runs the same source on the CPU (no GPU in the build container)
#define FVB_HD __host__ __device__ __forceinline__

// ---- device math, written to mirror torch's fp32 op order ------------------------------------------
FVB_HD float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

// MUFU forms used by the streaming kernels: <= 2.5e-6 relative to torch's sigmoid for |x| <= 30
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(x * -kLog2e)); }

// -t*log(p+1e-8) - (1-t)*log(1-p+1e-8), loss/classification_loss.py:55, evaluated exactly in that form.
FVB_HD float bce_term(float p, float t) {
  float a = (-t) * logf(p + 1e-8f);
  float b = (1.0f - t) * logf((1.0f - p) + 1e-8f);
  return a - b;
}

// The same expression with t = 0: the first product is (-0)*finite = +-0, so the result is exactly 0 - 1*log(...).
FVB_HD float bce_term_zero(float p) { return 0.0f - logf((1.0f - p) + 1e-8f); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum (fixed tree); result valid in thread 0.  `scratch` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? scratch[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// ---- IoU family, element form.  Boxes are xyxy here; callers convert first. -----------
struct Box {
  float x1, y1, x2, y2;
};

FVB_HD Box xywh_to_xyxy(float x, float y, float w, float h) {
  // detection/tools/BOX.py:4-10: divide by 2, then subtract/add
  float hw = w / 2.0f, hh = h / 2.0f;
  Box b;
  b.x1 = x - hw;
  b.y1 = y - hh;
  b.x2 = x + hw;
  b.y2 = y + hh;
  return b;
}

FVB_HD float clamp0(float z) { return fmaxf(z, 0.0f); }

FVB_HD float inter_area(const Box& a, const Box& b) {
  float iw = clamp0(fminf(a.x2, b.x2) - fmaxf(a.x1, b.x1));
  float ih = clamp0(fminf(a.y2, b.y2) - fmaxf(a.y1, b.y1));
  return iw * ih;
}

// torch.minimum/maximum/clamp propagate NaN; fminf/fmaxf do not.  The reference never feeds NaN boxes
// on this path (decoded boxes are finite), so the IEEE-754 minNum/maxNum behaviour is acceptable.

// inner_eps: element-wise xyxy_iou puts eps inside the height factor (IOU.py:74-75); pairwise does not (:143-144)
template <bool INNER_EPS>
FVB_HD float iou_plain(const Box& a, const Box& b, float eps, float* union_out = nullptr) {
  float area_a, area_b;
  if (INNER_EPS) {
    area_a = (a.x2 - a.x1) * ((a.y2 - a.y1) + eps);
    area_b = (b.x2 - b.x1) * ((b.y2 - b.y1) + eps);
  } else {
    area_a = (a.x2 - a.x1) * (a.y2 - a.y1);
    area_b = (b.x2 - b.x1) * (b.y2 - b.y1);
  }
  float inter = inter_area(a, b);
  float uni = ((area_a + area_b) - inter) + eps;
  if (union_out) *union_out = uni;
  return inter / uni;
}

// kind in {IOU,GIOU,DIOU,CIOU}; PAIRWISE selects the *_batch arithmetic of the reference.
template <bool PAIRWISE>
FVB_HD float iou_family(const Box& a, const Box& b, int kind, int variant, float eps) {
  if (kind == FVB_IOU) return iou_plain<!PAIRWISE>(a, b, eps);
  float cw = fmaxf(a.x2, b.x2) - fminf(a.x1, b.x1);
  float ch = fmaxf(a.y2, b.y2) - fminf(a.y1, b.y1);
  if (kind == FVB_GIOU) {
    // IOU.py:220-239 (element: iou - (C-U)/C) and :270-290 (batch: iou + (C-U)/C); areas without inner eps
    float uni;
    float iou = iou_plain<false>(a, b, eps, &uni);
    float convex = cw * ch + eps;
    float pen = (convex - uni) / convex;
    return PAIRWISE ? iou + pen : iou - pen;
  }
  // DIOU: IOU.py:307,324-341 / :358,375-393.  iou is xyxy_iou (inner eps) element-wise, xyxy_iou_batch pairwise.
  float iou = iou_plain<!PAIRWISE>(a, b, eps);
  float c2 = (cw * cw + ch * ch) + eps;
  float rho2;
  if (variant == FVB_VARIANT_DEMO) {  // demos/yolov3_u/utils/iou.py:334-341: centre sums not halved, minus sign
    float dx = (a.x1 + a.x2) - (b.x1 + b.x2);
    float dy = (a.y1 + a.y2) - (b.y1 + b.y2);
    rho2 = dx * dx + dy * dy;
  } else {
    float dx = (a.x1 + a.x2) * 0.5f - (b.x1 + b.x2) * 0.5f;
    float dy = (a.y1 + a.y2) * 0.5f - (b.y1 + b.y2) * 0.5f;
    rho2 = dx * dx + dy * dy;
  }
  float pen = rho2 / c2;
  float diou = (variant == FVB_VARIANT_DEMO) ? iou - pen : iou + pen;
  if (kind == FVB_DIOU) return diou;
  // CIOU: IOU.py:410-438 / :455-480
  float w1 = a.x2 - a.x1, h1 = a.y2 - a.y1;
  float w2 = b.x2 - b.x1, h2 = b.y2 - b.y1;
  float d = PAIRWISE ? atanf(w1 / (h1 + eps)) - atanf(w2 / (h2 + eps)) : atanf(w2 / (h2 + eps)) - atanf(w1 / (h1 + eps));
  const float four_over_pi2 = 0.4052847345693511f;  // (4 / math.pi ** 2) folded to fp32
  float v = four_over_pi2 * (d * d);
  float alpha = v / ((v - iou) + (1.0f + eps));  // 1 + eps is folded in double, then to fp32 (IOU.py:437)
  return diou - alpha * v;
}

FVB_HD Box load_box(const float* p, int box_mode) {
  if (box_mode == FVB_BOX_XYWH) return xywh_to_xyxy(p[0], p[1], p[2], p[3]);
  Box b;
  b.x1 = p[0]; b.y1 = p[1]; b.x2 = p[2]; b.y2 = p[3];
  return b;
}

// wh_iou: detection/tools/IOU.py:108-120 / :177-189
FVB_HD float wh_iou(float w1, float h1, float w2, float h2, float eps) {
  float inter = fminf(w1, w2) * fminf(h1, h2);
  float uni = ((w1 * h1 + w2 * h2) - inter) + eps;
  return inter / uni;
}

}  // namespace fvb
