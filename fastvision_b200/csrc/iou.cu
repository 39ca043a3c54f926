// K2: the IoU family (element-wise and pairwise, IoU/GIoU/DIoU/CIoU, lib + demo variants), the IoU
// losses, the binary cross-entropy loss and the box conversions -- one fused kernel each instead of
// the reference's 15-40 ATen launches per call (detection/tools/IOU.py, detection/tools/BOX.py,
// loss/iou_loss.py, loss/classification_loss.py).  The arithmetic lives in common.cuh (iou_family).
#include "common.cuh"

namespace fvb {

__global__ void iou_elementwise_kernel(const float* a, const float* b, long long n, int box_mode, int kind, int variant,
                                       float eps, float* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (box_mode == FVB_BOX_WH) {
    out[i] = wh_iou(a[i * 2], a[i * 2 + 1], b[i * 2], b[i * 2 + 1], eps);
    return;
  }
  Box ba = load_box(a + i * 4, box_mode), bb = load_box(b + i * 4, box_mode);
  out[i] = iou_family<false>(ba, bb, kind, variant, eps);
}

constexpr int kPairRows = 64, kPairCols = 128;

// Pairwise tile kernel.  A CTA owns a strip of 128 "b" boxes (one per thread column, in registers) and walks its share of the
// "a" boxes in tiles of 64 staged in shared memory (double-buffered: the next tile is fetched while the current one is
// computed, one barrier per tile); thread (col, half) computes 32 pairs per tile; every store instruction of a warp writes
// 128 contiguous bytes of one output row.  KIND is a template parameter: no per-pair dispatch.  (The first version gave every
// CTA 32 x 128 pairs: 16 pairs per thread behind a global load + barrier -- latency-bound at 24 % of the HBM write peak.)
template <int KIND, bool WH>
__global__ void __launch_bounds__(256) iou_pairwise_kernel(const float* __restrict__ a, long long n, const float* __restrict__ b,
                                                           long long m, int box_mode, int variant, float eps,
                                                           float* __restrict__ out, long long rows_per_cta) {
  __shared__ Box sa[2][kPairRows];
  const long long c0 = (long long)blockIdx.x * kPairCols;
  const long long row_beg = (long long)blockIdx.y * rows_per_cta, row_end = min(n, row_beg + rows_per_cta);
  const int tid = threadIdx.x;
  const int col = tid & (kPairCols - 1), half = tid / kPairCols;  // 2 row phases
  const bool col_ok = c0 + col < m;
  auto load_a = [&](long long r) -> Box {
    if (WH) {
      Box t; t.x1 = a[r * 2]; t.y1 = a[r * 2 + 1]; t.x2 = 0; t.y2 = 0;
      return t;
    }
    return load_box(a + r * 4, box_mode);
  };
  Box bb;
  bb.x1 = bb.y1 = bb.x2 = bb.y2 = 0.0f;
  if (col_ok) {
    if (WH) { bb.x1 = b[(c0 + col) * 2]; bb.y1 = b[(c0 + col) * 2 + 1]; }
    else bb = load_box(b + (c0 + col) * 4, box_mode);
  }
  if (row_beg >= row_end) return;
  if (tid < kPairRows && row_beg + tid < row_end) sa[0][tid] = load_a(row_beg + tid);
  __syncthreads();
  int buf = 0;
  for (long long r0 = row_beg; r0 < row_end; r0 += kPairRows, buf ^= 1) {
    // prefetch the next tile into registers; it lands in the other buffer after this tile's pairs
    const long long rn = r0 + kPairRows + tid;
    Box nxt;
    const bool have_next = tid < kPairRows && rn < row_end;
    if (have_next) nxt = load_a(rn);
    const int rows = (int)min((long long)kPairRows, row_end - r0);
    if (col_ok) {
      float* o = out + (r0 + half) * m + c0 + col;
      const long long step = (long long)(256 / kPairCols) * m;
      if (rows == kPairRows) {
        // full tile: fixed trip count, no bounds checks in the pair loop
#pragma unroll 8
        for (int i = 0; i < kPairRows / (256 / kPairCols); ++i) {
          const Box ba = sa[buf][half + i * (256 / kPairCols)];
          const float v = WH ? wh_iou(ba.x1, ba.y1, bb.x1, bb.y1, eps) : iou_family<true>(ba, bb, KIND, variant, eps);
          __stcs(o, v);
          o += step;
        }
      } else {
        for (int r = half; r < rows; r += 256 / kPairCols) {
          const Box ba = sa[buf][r];
          const float v = WH ? wh_iou(ba.x1, ba.y1, bb.x1, bb.y1, eps) : iou_family<true>(ba, bb, KIND, variant, eps);
          __stcs(o, v);
          o += step;
        }
      }
    }
    if (have_next) sa[buf ^ 1][tid] = nxt;
    __syncthreads();
  }
}

// ---- reductions ---------------------------------------------------------------------------------------------
constexpr int kRedThreads = 256;
constexpr int kRedPerBlock = 4096;

// partial[block] = {sum of (1 - kind_i) * w_i (or unweighted), sum of w_i}
__global__ void __launch_bounds__(kRedThreads) iou_loss_partial_kernel(const float* a, const float* b, const float* w,
                                                                       long long n, int box_mode, int kind, int variant,
                                                                       float eps, int outer_weights, double* partial) {
  __shared__ double scratch[32];
  double s = 0.0, sw = 0.0;
  long long base = (long long)blockIdx.x * kRedPerBlock;
  for (int k = threadIdx.x; k < kRedPerBlock; k += kRedThreads) {
    long long i = base + k;
    if (i >= n) break;
    float v;
    if (box_mode == FVB_BOX_WH) v = wh_iou(a[i * 2], a[i * 2 + 1], b[i * 2], b[i * 2 + 1], eps);
    else v = iou_family<false>(load_box(a + i * 4, box_mode), load_box(b + i * 4, box_mode), kind, variant, eps);
    float loss = 1.0f - v;                  // loss/iou_loss.py:20,46,72,98
    if (w != nullptr) {
      if (outer_weights) sw += (double)w[i];
      else loss = loss * w[i];
    }
    s += (double)loss;
  }
  s = block_sum(s, scratch);
  sw = block_sum(sw, scratch);
  if (threadIdx.x == 0) {
    partial[(size_t)blockIdx.x * 2] = s;
    partial[(size_t)blockIdx.x * 2 + 1] = sw;
  }
}

// out = sum(partials) [* sum(w) when the reference broadcasts [n]x[n,1]] / denom
__global__ void __launch_bounds__(1024) reduce_finish_kernel(const double* partial, int blocks, int outer_weights,
                                                             double denom, float* out) {
  __shared__ double scratch[32];
  double s = 0.0, sw = 0.0;
  for (int i = threadIdx.x; i < blocks; i += blockDim.x) {
    s += partial[(size_t)i * 2];
    sw += partial[(size_t)i * 2 + 1];
  }
  s = block_sum(s, scratch);
  sw = block_sum(sw, scratch);
  if (threadIdx.x == 0) {
    double r = outer_weights ? s * sw : s;
    out[0] = (float)(r / denom);
  }
}

__global__ void __launch_bounds__(kRedThreads) bce_partial_kernel(const float* pre, long long rows, int classes,
                                                                  const long long* tidx, const float* tval,
                                                                  int already_sigmoid, const float* w, double* partial) {
  __shared__ double scratch[32];
  double s = 0.0;
  long long total = rows * classes;
  long long base = (long long)blockIdx.x * kRedPerBlock;
  for (int k = threadIdx.x; k < kRedPerBlock; k += kRedThreads) {
    long long i = base + k;
    if (i >= total) break;
    float t;
    if (classes > 1) {
      long long r = i / classes;
      int c = (int)(i - r * classes);
      t = (tidx[r] == (long long)c) ? 1.0f : 0.0f;  // one_hot, datasets/common/id_2_onehot.py:10-15
    } else {
      t = tval[i];                                    // classification_loss.py:47-48
    }
    float p = already_sigmoid ? pre[i] : sigmoid_precise(pre[i]);
    float loss = bce_term(p, t);
    if (w != nullptr) loss = loss * w[i];
    s += (double)loss;
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    partial[(size_t)blockIdx.x * 2] = s;
    partial[(size_t)blockIdx.x * 2 + 1] = 0.0;
  }
}

__global__ void box_convert_kernel(const float* in, long long n, int op, float height, float width, float* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p0 = in[i * 4], p1 = in[i * 4 + 1], p2 = in[i * 4 + 2], p3 = in[i * 4 + 3];
  float o0, o1, o2, o3;
  if (op == 0) {  // xywh2xyxy, BOX.py:4-10
    Box b = xywh_to_xyxy(p0, p1, p2, p3);
    o0 = b.x1; o1 = b.y1; o2 = b.x2; o3 = b.y2;
  } else if (op == 1) {  // xyxy2xywh, BOX.py:12-18
    o0 = (p0 + p2) / 2.0f; o1 = (p1 + p3) / 2.0f; o2 = p2 - p0; o3 = p3 - p1;
  } else {  // xyxy2xywhn, BOX.py:20-26
    o0 = ((p0 + p2) / 2.0f) / width; o1 = ((p1 + p3) / 2.0f) / height;
    o2 = (p2 - p0) / width; o3 = (p3 - p1) / height;
  }
  out[i * 4] = o0; out[i * 4 + 1] = o1; out[i * 4 + 2] = o2; out[i * 4 + 3] = o3;
}

static int check_iou_args(int box_mode, int kind, int variant, const char* who) {
  FVB_REQUIRE(box_mode >= FVB_BOX_XYXY && box_mode <= FVB_BOX_WH, "%s: mode must be xyxy or xywh or wh", who);
  FVB_REQUIRE(kind >= FVB_IOU && kind <= FVB_CIOU, "%s: unknown IoU kind %d", who, kind);
  FVB_REQUIRE(variant == FVB_VARIANT_LIB || variant == FVB_VARIANT_DEMO, "%s: unknown variant %d", who, variant);
  FVB_REQUIRE(!(box_mode == FVB_BOX_WH && kind != FVB_IOU), "%s: wh mode supports plain IoU only", who);
  return FVB_OK;
}

}  // namespace fvb

using namespace fvb;

extern "C" int fvb_box_convert_f32(const float* d_in, int64_t n, int op, float height, float width, float* d_out,
                                   void* stream) {
  FVB_REQUIRE(n >= 0 && op >= 0 && op <= 2, "box_convert: bad arguments");
  if (n == 0) return FVB_OK;
  FVB_REQUIRE(d_in && d_out, "box_convert: NULL pointer");
  box_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, n, op, height, width, d_out);
  count_launch();
  return check_launch("box_convert_kernel");
}

extern "C" int fvb_iou_elementwise_f32(const float* d_a, const float* d_b, int64_t n, int box_mode, int kind, int variant,
                                       float eps, float* d_out, void* stream) {
  int rc = check_iou_args(box_mode, kind, variant, "iou_elementwise");
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(n >= 0, "iou_elementwise: n=%lld", (long long)n);
  if (n == 0) return FVB_OK;
  FVB_REQUIRE(d_a && d_b && d_out, "iou_elementwise: NULL pointer");
  iou_elementwise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, box_mode, kind, variant, eps, d_out);
  count_launch();
  return check_launch("iou_elementwise_kernel");
}

extern "C" int fvb_iou_pairwise_f32(const float* d_a, int64_t n, const float* d_b, int64_t m, int box_mode, int kind,
                                    int variant, float eps, float* d_out, void* stream) {
  int rc = check_iou_args(box_mode, kind, variant, "iou_pairwise");
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(n >= 0 && m >= 0, "iou_pairwise: negative size");
  if (n == 0 || m == 0) return FVB_OK;
  FVB_REQUIRE(d_a && d_b && d_out, "iou_pairwise: NULL pointer");
  const long long strips = (m + kPairCols - 1) / kPairCols;
  FVB_REQUIRE(strips < (1ll << 31), "iou_pairwise: M=%lld too large for one launch", (long long)m);
  // enough CTAs to fill the machine (~8 per SM), each walking a contiguous share of the rows in tiles of kPairRows
  const long long tiles = (n + kPairRows - 1) / kPairRows;
  long long splits = (1184 + strips - 1) / strips;
  if (splits > tiles) splits = tiles;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  const long long rows_per_cta = ((tiles + splits - 1) / splits) * kPairRows;
  dim3 grid((unsigned)strips, (unsigned)((n + rows_per_cta - 1) / rows_per_cta));
  cudaStream_t st = (cudaStream_t)stream;
#define FVB_PAIR(KIND_, WH_) iou_pairwise_kernel<KIND_, WH_><<<grid, 256, 0, st>>>(d_a, n, d_b, m, box_mode, variant, eps, d_out, rows_per_cta)
  if (box_mode == FVB_BOX_WH) FVB_PAIR(FVB_IOU, true);
  else if (kind == FVB_IOU) FVB_PAIR(FVB_IOU, false);
  else if (kind == FVB_GIOU) FVB_PAIR(FVB_GIOU, false);
  else if (kind == FVB_DIOU) FVB_PAIR(FVB_DIOU, false);
  else FVB_PAIR(FVB_CIOU, false);
#undef FVB_PAIR
  count_launch();
  return check_launch("iou_pairwise_kernel");
}

extern "C" size_t fvb_reduce_workspace_bytes(int64_t n) {
  size_t blocks = (size_t)((n + kRedPerBlock - 1) / kRedPerBlock) + 1;
  return blocks * 16 + 256;
}

extern "C" int fvb_iou_loss_f32(const float* d_pre, const float* d_true, const float* d_weights, int64_t n, int box_mode,
                                int kind, int variant, float eps, int reduction, float* d_out, void* d_ws, void* stream) {
  int rc = check_iou_args(box_mode, kind, variant, "iou_loss");
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(n >= 1, "iou_loss: n=%lld (the reference's mean of an empty tensor is NaN)", (long long)n);
  FVB_REQUIRE(d_pre && d_true && d_out && d_ws, "iou_loss: NULL pointer");
  FVB_REQUIRE(reduction == FVB_REDUCE_MEAN || reduction == FVB_REDUCE_SUM, "iou_loss: reduction");
  int blocks = (int)((n + kRedPerBlock - 1) / kRedPerBlock);
  // GIOU returns [n]; times [n,1] weights the reference broadcasts to [n,n] (loss/iou_loss.py:51)
  int outer = (kind == FVB_GIOU && d_weights != nullptr) ? 1 : 0;
  double denom = 1.0;
  if (reduction == FVB_REDUCE_MEAN) denom = outer ? (double)n * (double)n : (double)n;
  cudaStream_t s = (cudaStream_t)stream;
  iou_loss_partial_kernel<<<blocks, kRedThreads, 0, s>>>(d_pre, d_true, d_weights, n, box_mode, kind, variant, eps, outer, (double*)d_ws);
  reduce_finish_kernel<<<1, 1024, 0, s>>>((const double*)d_ws, blocks, outer, denom, d_out);
  count_launch(2);
  return check_launch("iou_loss");
}

extern "C" int fvb_bce_loss_f32(const float* d_pre, int64_t rows, int classes, const int64_t* d_target_idx,
                                const float* d_target_val, int already_sigmoid, const float* d_weights, int reduction,
                                float* d_out, void* d_ws, void* stream) {
  FVB_REQUIRE(rows >= 1 && classes >= 1, "bce_loss: rows=%lld classes=%d", (long long)rows, classes);
  FVB_REQUIRE(d_pre && d_out && d_ws, "bce_loss: NULL pointer");
  FVB_REQUIRE(classes > 1 ? d_target_idx != nullptr : d_target_val != nullptr, "bce_loss: target pointer for C=%d missing", classes);
  FVB_REQUIRE(reduction == FVB_REDUCE_MEAN || reduction == FVB_REDUCE_SUM, "bce_loss: reduction");
  long long total = rows * classes;
  int blocks = (int)((total + kRedPerBlock - 1) / kRedPerBlock);
  double denom = reduction == FVB_REDUCE_MEAN ? (double)total : 1.0;  // classification_loss.py:63-65
  cudaStream_t s = (cudaStream_t)stream;
  bce_partial_kernel<<<blocks, kRedThreads, 0, s>>>(d_pre, rows, classes, (const long long*)d_target_idx, d_target_val,
                                                     already_sigmoid, d_weights, (double*)d_ws);
  reduce_finish_kernel<<<1, 1024, 0, s>>>((const double*)d_ws, blocks, 0, denom, d_out);
  count_launch(2);
  return check_launch("bce_loss");
}
