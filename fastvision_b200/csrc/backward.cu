// Backward of the loss side of the path: what `loss.backward()` in Fit._train (utils/fit.py:57-63)
// asks autograd to do for Yolov3Loss (loss/yolov3_loss.py:29-72), the IoU losses (loss/iou_loss.py:5-107) and
// BiCrossEntropyLoss (loss/classification_loss.py:36-65) -- as three hand-written kernels instead of the ~400 ATen
// backward launches of the reference graph.
//
//   yolo_grad_dense  : HBM-bound stream.  Writes the whole gradient of every level tensor [B,A,H,W,K]: zero
//                      everywhere except channel 4, which gets the objectness-BCE gradient for target 0 (the dense term
//                      of yolov3_loss.py:63-64).  One 16-byte store per 4 floats; the objectness logits come from the compact
//                      copy the forward saved (4 bytes per row) -- read strided from the heads they cost a 128-byte DRAM
//                      fetch per row (ncu: 349 MB of reads next to 869 MB of writes at B=256).
//   yolo_grad_box / yolo_grad_rows : the matched rows.  The LAST match of a cell (the one whose IoU the reference's
//                      index_put keeps, :61) owns the cell's row: it adds, in target order, the class-BCE, CIoU and
//                      IoU-target gradients of every match that hit the cell (autograd's index backward scatter-adds
//                      duplicates; index_put's backward hands the cell's objectness-target gradient to every duplicate)
//                      and overwrites channel 4 with the gradient for target = IoU.  Box math: one thread per (level,
//                      target, anchor); class channels + row store: one warp per owned row.  Single writer per row: no
//                      atomics on the gradients, bit-reproducible.
//   iou_loss_grad / bce_grad : element-wise.
#include "iou_grad.cuh"
#include "loss_common.cuh"

namespace fvb {

struct GradParams {
  Geom g;
  float* grad[FVB_MAX_LEVELS];
  const float* grad_out;   // [1] device scalar or NULL (= 1)
  const double* partials;  // [L][4], M_l at +3 (all-reduced under data parallelism)
  const float* labels;
  int T;
  const int* flags;        // [0] labels grouped by image
  const float* saved_conf; // compact objectness logits from the forward ([l][b][row]) or NULL (strided loads from the heads)
  long long saved_off[FVB_MAX_LEVELS];
  float r_box, r_conf, r_cls;
  long long batch_global;
  // dense pass
  long long lvl_floats[FVB_MAX_LEVELS];   // B*A*H*W*K
  int cta_begin[FVB_MAX_LEVELS + 1];      // first CTA of each level
};

constexpr int kDenseThreads = 256;
constexpr int kDenseIters = 16;                                   // float4 per thread
constexpr int kDenseChunk = kDenseThreads * kDenseIters * 4;      // floats per CTA

__device__ __forceinline__ float upstream(const GradParams& p) { return p.grad_out ? p.grad_out[0] : 1.0f; }

// coefficient of the objectness sum of level l in the scalar: loss = B_g * r_conf * S_conf / (B_g*A*H*W)
__device__ __forceinline__ float conf_coef(const GradParams& p, int l, float gout) {
  const double cells = (double)p.batch_global * p.g.A * p.g.HW[l];
  return (float)((double)gout * (double)p.batch_global * (double)p.r_conf / cells);
}

__device__ __forceinline__ float dense_conf_grad(float logit, float coef) {
  const float pr = sigmoid_precise(logit);
  return coef * (bce_dp(pr, 0.0f) * ((1.0f - pr) * pr));  // sigmoid backward: grad * (1 - y) * y
}

// Two phases per CTA (one 64 KB chunk of one level tensor): (1) the chunk's <= kDenseChunk/K + 1 objectness logits are
// loaded with one thread per row -- every strided load of the CTA is in flight at once (a load is a 32-byte sector per
// K*4-byte row; issued lazily from the store loop they were latency-bound: 41 % of the HBM peak) -- turned into gradients
// and parked in shared memory; (2) the chunk is written with 16-byte stores, channel-4 slots picked from shared memory.
constexpr int kDenseMaxRows = kDenseChunk / 6 + 2;  // K >= 6

template <bool VEC>
__global__ void __launch_bounds__(kDenseThreads) yolo_grad_dense_kernel(const GradParams p) {
  __shared__ float s_grad[kDenseMaxRows];
  int l = 0;
#pragma unroll
  for (int i = 1; i < FVB_MAX_LEVELS; ++i)
    if (i < p.g.L && (int)blockIdx.x >= p.cta_begin[i]) l = i;
  const int K = p.g.K;
  const long long n = p.lvl_floats[l];
  const long long base = (long long)((int)blockIdx.x - p.cta_begin[l]) * kDenseChunk;
  const long long stop = min(base + (long long)kDenseChunk, n);
  const float* __restrict__ head = p.g.head[l];
  float* __restrict__ out = p.grad[l];
  const float coef = conf_coef(p, l, upstream(p));
  // rows whose channel-4 element (float index r*K + 4) lies in [base, stop)
  const long long first_row = base <= 4 ? 0 : (base - 4 + K - 1) / K;
  const int n_rows = (int)max(0ll, (stop - 4 + K - 1) / K - first_row);
  if (p.saved_conf != nullptr) {
    const float* __restrict__ sc = p.saved_conf + p.saved_off[l] + first_row;  // coalesced: 4 bytes per row instead of a sector
    for (int r = threadIdx.x; r < n_rows; r += kDenseThreads) s_grad[r] = dense_conf_grad(__ldg(sc + r), coef);
  } else {
    for (int r = threadIdx.x; r < n_rows; r += kDenseThreads)
      s_grad[r] = dense_conf_grad(__ldg(head + (first_row + r) * K + 4), coef);
  }
  __syncthreads();
  if (VEC) {
    // thread's first float index, its row and channel; consecutive iterations advance by 4*kDenseThreads floats
    long long i = base + (long long)threadIdx.x * 4;
    const long long row0 = i / K;
    int c = (int)(i - row0 * K);
    int rl = (int)(row0 - first_row);  // row of float i, relative to the first staged row (may be -1)
    const int step_c = (4 * kDenseThreads) % K, step_r = (4 * kDenseThreads) / K;
#pragma unroll 4
    for (int it = 0; it < kDenseIters; ++it) {
      if (i + 3 < n) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        // which of the four floats (if any) is channel 4: e = (4 - c) mod K; K >= 6 > 4: at most one per float4
        int e = 4 - c, rr = rl;
        if (e < 0) {
          e += K;
          rr += 1;
        }
        if (e < 4) {
          const float gq = s_grad[rr];
          if (e == 0) v.x = gq; else if (e == 1) v.y = gq; else if (e == 2) v.z = gq; else v.w = gq;
        }
        __stcs(reinterpret_cast<float4*>(out + i), v);
      } else {
        for (int e = 0; e < 4 && i + e < n; ++e) {
          int ce = c + e, rr = rl;
          if (ce >= K) {
            ce -= K;
            rr += 1;
          }
          out[i + e] = ce == 4 ? s_grad[rr] : 0.0f;
        }
      }
      i += 4 * kDenseThreads;
      c += step_c;
      rl += step_r;
      if (c >= K) {
        c -= K;
        rl += 1;
      }
    }
  } else {
    for (int it = 0; it < kDenseIters * 4; ++it) {
      const long long i = base + (long long)it * kDenseThreads + threadIdx.x;
      if (i < n) {
        const long long row = i / K;
        out[i] = (int)(i - row * K) == 4 ? s_grad[(int)(row - first_row)] : 0.0f;
      }
    }
  }
}

constexpr int kMatchThreads = 256;

// flags[0] = 1 iff the labels are grouped by image (collate_fn order): lets the duplicate scans stop early
__global__ void labels_grouped_kernel(const float* labels, int T, int* flags) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int t = threadIdx.x; t + 1 < T; t += blockDim.x)
    if ((int)labels[(size_t)t * 6] > (int)labels[(size_t)(t + 1) * 6]) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    flags[0] = bad ? 0 : 1;
    flags[1] = 0;  // record counter of the matched-row pass
  }
}

// Matched rows, two phases.  The box / objectness gradient of a match is scalar work (reverse mode of CIoU + IoU): a warp
// per match made 32 lanes repeat it (99 registers, 2 CTAs per SM, 45 us at B=256).  Phase 1 gives every (level, target,
// anchor) ONE THREAD: ratio test, duplicate-cell scan, the four box-logit gradients summed over the cell's matches in target
// order, the objectness gradient -- and appends a record for each cell it owns.  Phase 2 gives every record one WARP: the
// class-BCE gradients of the row's channels (lanes over channels) and the coalesced row store.
struct RowRecord {
  unsigned long long row;  // offset of the row in floats inside its level tensor
  int level, t, lo, dups;  // owner target, first target index of its image range to rescan, matches on the cell
  int a, cls;
  float g[4];
  float conf_grad;
  float pad;
};

struct MatchWs {
  int* flags;            // [0] grouped, [1] record count
  RowRecord* records;    // [L*T*A]
};

__device__ __forceinline__ bool same_cell(const TargetCell& x, const TargetCell& y) {
  return y.match && y.gx == x.gx && y.gy == x.gy;
}

__global__ void __launch_bounds__(kMatchThreads) yolo_grad_box_kernel(const GradParams p, MatchWs ws) {
  const int TA = p.T * p.g.A;
  const int i = blockIdx.x * kMatchThreads + threadIdx.x;
  if (i >= TA * p.g.L) return;
  const int l = i / TA, ta = i - l * TA;
  const int t = ta / p.g.A, a = ta - t * p.g.A;
  const TargetCell c = target_cell(p.g, l, p.labels + (size_t)t * 6, a);
  if (!(c.match && c.b >= 0 && c.b < p.g.B)) return;
  const bool grouped = ws.flags[0] != 0;
  // loser: a later target of the same image hits the same cell with the same anchor -> that thread owns the row
  for (int t2 = t + 1; t2 < p.T; ++t2) {
    const float* lab2 = p.labels + (size_t)t2 * 6;
    const int b2 = (int)lab2[0];
    if (b2 == c.b) {
      if (same_cell(c, target_cell(p.g, l, lab2, a))) return;
    } else if (grouped && b2 > c.b) {
      break;
    }
  }
  const size_t rofs = cell_row(p.g, l, c.b, a, c.gy, c.gx) * p.g.K;
  const float* row = p.g.head[l] + rofs;
  const float r0 = __ldg(row), r1 = __ldg(row + 1), r2 = __ldg(row + 2), r3 = __ldg(row + 3), r4 = __ldg(row + 4);
  const float gout = upstream(p);
  const double M = p.partials[l * 4 + 3];
  // (loss_box + loss_conf + loss_cls) * bs, ratios applied to the per-level means (yolov3_loss.py:52,58,66-72)
  const float w_box = (float)((double)gout * (double)p.batch_global * (double)p.r_box / M);
  const float w_conf = conf_coef(p, l, gout);
  const float p4 = sigmoid_precise(r4);
  const float g_tgt = w_conf * bce_dt(p4);  // gradient reaching targets_conf[cell]; index_put backward gives it to every duplicate
  // first target of this image (grouped labels) -- earlier matches of the cell are added in ascending target order
  int lo = 0;
  if (grouped) {
    lo = t;
    while (lo > 0 && (int)p.labels[(size_t)(lo - 1) * 6] == c.b) --lo;
  }
  RowRecord r;
  r.g[0] = r.g[1] = r.g[2] = r.g[3] = 0.0f;
  r.dups = 0;
  for (int t2 = lo; t2 < t; ++t2) {
    const float* lab2 = p.labels + (size_t)t2 * 6;
    if ((int)lab2[0] != c.b) continue;
    const TargetCell m = target_cell(p.g, l, lab2, a);
    if (!same_cell(c, m)) continue;
    const MatchRowGrad mg = match_row_grad(r0, r1, r2, r3, m, w_box, g_tgt);
#pragma unroll
    for (int k = 0; k < 4; ++k) r.g[k] += mg.g[k];
    r.dups += 1;
  }
  const MatchRowGrad mg = match_row_grad(r0, r1, r2, r3, c, w_box, g_tgt);
#pragma unroll
  for (int k = 0; k < 4; ++k) r.g[k] += mg.g[k];
  r.dups += 1;
  r.conf_grad = w_conf * (bce_dp(p4, mg.iou) * ((1.0f - p4) * p4));
  r.row = (unsigned long long)rofs;
  r.level = l;
  r.t = t;
  r.lo = lo;
  r.a = a;
  r.cls = c.cls;
  r.pad = 0.0f;
  ws.records[atomicAdd(&ws.flags[1], 1)] = r;  // record order is arbitrary; every record owns a distinct row
}

__global__ void __launch_bounds__(kMatchThreads) yolo_grad_rows_kernel(const GradParams p, MatchWs ws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int count = ws.flags[1];
  const int K = p.g.K, C = K - 5;
  for (int ri = blockIdx.x * (kMatchThreads / 32) + warp; ri < count; ri += gridDim.x * (kMatchThreads / 32)) {
    const RowRecord r = ws.records[ri];
    const int l = r.level;
    const float* row = p.g.head[l] + r.row;
    float* grow = p.grad[l] + r.row;
    const double M = p.partials[l * 4 + 3];
    const float w_cls = (float)((double)upstream(p) * (double)p.batch_global * (double)p.r_cls / (M * (double)C));
    TargetCell own;
    if (r.dups > 1) own = target_cell(p.g, l, p.labels + (size_t)r.t * 6, r.a);
    for (int ch = lane; ch < K; ch += 32) {
      float gv;
      if (ch < 4) gv = ch == 0 ? r.g[0] : (ch == 1 ? r.g[1] : (ch == 2 ? r.g[2] : r.g[3]));
      else if (ch == 4) gv = r.conf_grad;
      else {
        const float pr = sigmoid_precise(__ldg(row + ch));
        const float dsig = (1.0f - pr) * pr;
        const float g_neg = w_cls * (bce_dp(pr, 0.0f) * dsig), g_pos = w_cls * (bce_dp(pr, 1.0f) * dsig);
        if (r.dups == 1) {
          gv = (ch - 5 == r.cls) ? g_pos : g_neg;
        } else {
          // several matches share the cell: their class terms add up in target order (torch's index backward)
          gv = 0.0f;
          const int b = (int)p.labels[(size_t)r.t * 6];
          for (int t2 = r.lo; t2 <= r.t; ++t2) {
            const float* lab2 = p.labels + (size_t)t2 * 6;
            if ((int)lab2[0] != b) continue;
            const TargetCell m = target_cell(p.g, l, lab2, r.a);
            if (t2 != r.t && !same_cell(own, m)) continue;
            gv += (ch - 5 == m.cls) ? g_pos : g_neg;
          }
        }
      }
      grow[ch] = gv;
    }
  }
}

// ---- IoU losses and BCE -------------------------------------------------------------------------------------------
struct IouGradParams {
  const float *a, *b, *w;
  long long n;
  int box_mode, kind, variant, reduction, outer;
  float eps;
  const float* grad_out;
  const double* wsum;  // [1] sum of weights (GIoU's [n] x [n,1] broadcast, loss/iou_loss.py:46-51)
  float *ga, *gb;
};

__global__ void weight_sum_kernel(const float* w, long long n, double* out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)w[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void iou_loss_grad_kernel(const IouGradParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const float gout = p.grad_out ? p.grad_out[0] : 1.0f;
  // d loss / d kind_i = -(w_i or sum_j w_j) / denom
  double coef = -(double)gout;
  if (p.outer) coef *= p.wsum[0];
  else if (p.w) coef *= (double)p.w[i];
  if (p.reduction == FVB_REDUCE_MEAN) coef /= p.outer ? (double)p.n * (double)p.n : (double)p.n;
  const float g = (float)coef;
  if (p.box_mode == FVB_BOX_WH) {
    const float w1 = p.a[i * 2], h1 = p.a[i * 2 + 1], w2 = p.b[i * 2], h2 = p.b[i * 2 + 1];
    const float mw = fminf(w1, w2), mh = fminf(h1, h2);
    const float inter = mw * mh;
    const float uni = ((w1 * h1 + w2 * h2) - inter) + p.eps;
    const float iou = inter / uni;
    const float g_uni = 0.0f - g * iou / uni, g_inter = g / uni - g_uni;
    const float sw = min_share(w1, w2), sh = min_share(h1, h2);
    if (p.ga) {
      p.ga[i * 2] = g_inter * mh * sw + g_uni * h1;
      p.ga[i * 2 + 1] = g_inter * mw * sh + g_uni * w1;
    }
    if (p.gb) {
      p.gb[i * 2] = g_inter * mh * (1.0f - sw) + g_uni * h2;
      p.gb[i * 2 + 1] = g_inter * mw * (1.0f - sh) + g_uni * w2;
    }
    return;
  }
  const Box ba = load_box(p.a + i * 4, p.box_mode), bb = load_box(p.b + i * 4, p.box_mode);
  BoxGrad ga = zero_grad(), gb = zero_grad();
  iou_family_grad(ba, bb, p.kind, p.variant, p.eps, g, ga, gb);
  for (int s = 0; s < 2; ++s) {
    float* o = s == 0 ? p.ga : p.gb;
    if (!o) continue;
    const BoxGrad& q = s == 0 ? ga : gb;
    if (p.box_mode == FVB_BOX_XYWH) {
      float gx, gy, gw, gh;
      xyxy_grad_to_xywh(q, &gx, &gy, &gw, &gh);
      o[i * 4] = gx; o[i * 4 + 1] = gy; o[i * 4 + 2] = gw; o[i * 4 + 3] = gh;
    } else {
      o[i * 4] = q.x1; o[i * 4 + 1] = q.y1; o[i * 4 + 2] = q.x2; o[i * 4 + 3] = q.y2;
    }
  }
}

__global__ void bce_grad_kernel(const float* pre, long long rows, int classes, const long long* tidx, const float* tval,
                                int already_sigmoid, const float* w, int reduction, const float* grad_out, float* gpre) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * classes;
  if (i >= total) return;
  float t;
  if (classes > 1) {
    const long long r = i / classes;
    t = (tidx[r] == (i - r * classes)) ? 1.0f : 0.0f;
  } else {
    t = tval[i];
  }
  const float gout = grad_out ? grad_out[0] : 1.0f;
  double coef = (double)gout;
  if (w) coef *= (double)w[i];
  if (reduction == FVB_REDUCE_MEAN) coef /= (double)total;
  float g;
  if (already_sigmoid) {
    g = bce_dp(pre[i], t);
  } else {
    const float pr = sigmoid_precise(pre[i]);
    g = bce_dp(pr, t) * ((1.0f - pr) * pr);
  }
  gpre[i] = (float)coef * g;
}

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_yolov3_loss_backward_workspace_bytes(const fvb_yolo_geom* geom, int64_t num_labels) {
  if (geom == nullptr || num_labels < 0) return 0;
  return 256 + (size_t)geom->levels * (size_t)num_labels * (size_t)geom->anchors * sizeof(RowRecord) + 256;
}

extern "C" int fvb_yolov3_loss_backward_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                            int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                                            int64_t batch_global, const double* d_partials, const float* d_saved_conf,
                                            const float* d_grad_out, float* const* d_grad_heads, void* d_ws, void* stream) {
  FVB_REQUIRE(d_heads && d_grad_heads && d_partials && d_ws, "yolov3_loss_backward: NULL pointer");
  FVB_REQUIRE(num_labels >= 0 && num_labels < (1ll << 24), "yolov3_loss_backward: num_labels=%lld", (long long)num_labels);
  FVB_REQUIRE(num_labels == 0 || d_labels, "yolov3_loss_backward: labels NULL");
  FVB_REQUIRE(batch_global >= 1, "yolov3_loss_backward: batch_global=%lld", (long long)batch_global);
  FVB_REQUIRE(((uintptr_t)d_ws & 255) == 0, "yolov3_loss_backward: workspace must be 256-byte aligned");
  GradParams p;
  int rc = make_geom(geom, d_heads, &p.g);
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(!p.g.nchw, "yolov3_loss_backward: heads must be [B,A,H,W,K] (FVB_HEAD_BAHWK)");
  FVB_REQUIRE(p.g.B >= 1, "yolov3_loss_backward: empty batch");
  const Geom& g = p.g;
  bool vec = true;
  long long ctas = 0;
  for (int l = 0; l < FVB_MAX_LEVELS; ++l) {
    p.grad[l] = nullptr;
    p.lvl_floats[l] = 0;
  }
  for (int l = 0; l < g.L; ++l) {
    FVB_REQUIRE(d_heads[l] && d_grad_heads[l], "yolov3_loss_backward: level %d pointer is NULL", l);
    p.grad[l] = d_grad_heads[l];
    p.lvl_floats[l] = (long long)g.B * g.A * g.HW[l] * g.K;
    p.cta_begin[l] = (int)ctas;
    p.saved_off[l] = (long long)g.B * g.row_off[l];
    ctas += (p.lvl_floats[l] + kDenseChunk - 1) / kDenseChunk;
    if (((uintptr_t)d_grad_heads[l] & 15) != 0) vec = false;
  }
  for (int l = g.L; l <= FVB_MAX_LEVELS; ++l) p.cta_begin[l] = (int)ctas;
  FVB_REQUIRE(ctas < (1ll << 31), "yolov3_loss_backward: tensor too large for one launch");
  p.grad_out = d_grad_out;
  p.saved_conf = d_saved_conf;
  p.partials = d_partials;
  p.labels = d_labels;
  p.T = (int)num_labels;
  p.flags = (int*)d_ws;
  p.r_box = ratio_box;
  p.r_conf = ratio_conf;
  p.r_cls = ratio_cls;
  p.batch_global = batch_global;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec) yolo_grad_dense_kernel<true><<<(unsigned)ctas, kDenseThreads, 0, s>>>(p);
  else yolo_grad_dense_kernel<false><<<(unsigned)ctas, kDenseThreads, 0, s>>>(p);
  count_launch();
  if (p.T > 0) {
    MatchWs mw;
    mw.flags = (int*)d_ws;
    mw.records = (RowRecord*)((unsigned char*)d_ws + 256);
    labels_grouped_kernel<<<1, 1024, 0, s>>>(d_labels, p.T, mw.flags);   // also zeroes the record counter
    const long long pairs = (long long)p.T * g.A * g.L;
    yolo_grad_box_kernel<<<(unsigned)((pairs + kMatchThreads - 1) / kMatchThreads), kMatchThreads, 0, s>>>(p, mw);
    const long long row_ctas = (pairs + kMatchThreads / 32 - 1) / (kMatchThreads / 32);
    yolo_grad_rows_kernel<<<(unsigned)(row_ctas < 1184 ? row_ctas : 1184), kMatchThreads, 0, s>>>(p, mw);
    count_launch(3);
  }
  return check_launch("yolov3_loss_backward");
}

extern "C" int fvb_iou_loss_backward_f32(const float* d_pre, const float* d_true, const float* d_weights, int64_t n,
                                         int box_mode, int kind, int variant, float eps, int reduction,
                                         const float* d_grad_out, float* d_grad_pre, float* d_grad_true, void* d_ws,
                                         void* stream) {
  FVB_REQUIRE(box_mode >= FVB_BOX_XYXY && box_mode <= FVB_BOX_WH, "iou_loss_backward: mode must be xyxy or xywh or wh");
  FVB_REQUIRE(kind >= FVB_IOU && kind <= FVB_CIOU, "iou_loss_backward: unknown IoU kind %d", kind);
  FVB_REQUIRE(variant == FVB_VARIANT_LIB || variant == FVB_VARIANT_DEMO, "iou_loss_backward: unknown variant %d", variant);
  FVB_REQUIRE(!(box_mode == FVB_BOX_WH && kind != FVB_IOU), "iou_loss_backward: wh mode supports plain IoU only");
  FVB_REQUIRE(n >= 1, "iou_loss_backward: n=%lld", (long long)n);
  FVB_REQUIRE(d_pre && d_true && d_ws && (d_grad_pre || d_grad_true), "iou_loss_backward: NULL pointer");
  FVB_REQUIRE(reduction == FVB_REDUCE_MEAN || reduction == FVB_REDUCE_SUM, "iou_loss_backward: reduction");
  IouGradParams p;
  p.a = d_pre; p.b = d_true; p.w = d_weights; p.n = n;
  p.box_mode = box_mode; p.kind = kind; p.variant = variant; p.reduction = reduction;
  p.outer = (kind == FVB_GIOU && d_weights != nullptr) ? 1 : 0;
  p.eps = eps; p.grad_out = d_grad_out; p.wsum = (const double*)d_ws; p.ga = d_grad_pre; p.gb = d_grad_true;
  cudaStream_t s = (cudaStream_t)stream;
  if (p.outer) {
    weight_sum_kernel<<<1, 1024, 0, s>>>(d_weights, n, (double*)d_ws);
    count_launch();
  }
  iou_loss_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
  count_launch();
  return check_launch("iou_loss_backward");
}

extern "C" int fvb_bce_loss_backward_f32(const float* d_pre, int64_t rows, int classes, const int64_t* d_target_idx,
                                         const float* d_target_val, int already_sigmoid, const float* d_weights,
                                         int reduction, const float* d_grad_out, float* d_grad_pre, void* stream) {
  FVB_REQUIRE(rows >= 1 && classes >= 1, "bce_loss_backward: rows=%lld classes=%d", (long long)rows, classes);
  FVB_REQUIRE(d_pre && d_grad_pre, "bce_loss_backward: NULL pointer");
  FVB_REQUIRE(classes > 1 ? d_target_idx != nullptr : d_target_val != nullptr, "bce_loss_backward: target pointer for C=%d missing", classes);
  FVB_REQUIRE(reduction == FVB_REDUCE_MEAN || reduction == FVB_REDUCE_SUM, "bce_loss_backward: reduction");
  const long long total = rows * classes;
  bce_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_pre, rows, classes, (const long long*)d_target_idx, d_target_val, already_sigmoid, d_weights, reduction, d_grad_out, d_grad_pre);
  count_launch();
  return check_launch("bce_grad_kernel");
}
