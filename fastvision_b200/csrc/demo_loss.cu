// The demos' training loss `ComputeLoss`, forward and backward, on the raw conv outputs [B, A*K, H, W]:
//   flavour SHIP  demos/yolov3_huaweiShip/utils/lossv3.py:19-125  -> (loss_box [CIoU, demo variant], loss_cls, loss_conf)
//   flavour U     demos/yolov3_u/utils/lossv3.py:17-119           -> 2*loss_xy + loss_wh + loss_cls + loss_conf
// Per level every target is assigned to its best anchor by wh-IoU (first maximum) and to the cell floor(xy); the
// objectness target is 1 at those cells, "ignore" where the predicted box of a cell overlaps any target of its image by
// IoU > 0.5 (the reference's per-image Python loop over xywh_iou_batch([A*H*W,4],[T_i,4]), lossv3.py:102-113) and 0
// elsewhere; every term is a mean (BCE-with-logits / MSE / 1-CIoU).
//
//   demo_prep     : per-image target lists (identity + boundaries for grouped labels, else counting sort, one CTA);
//   demo_keys     : keys (cell ids) of every (level, target)
//   demo_targets  : one warp per (level, target): gather the matched row from the NCHW planes, box / xy-wh and class terms
//   demo_cells    : HBM-bound stream over the 5 head planes of every anchor: decode the box, max pairwise IoU against the
//                   image's targets staged in shared memory, ignore / positive / negative, objectness BCE; optionally
//                   saves the int8 mask for the backward.  Reads 5 of K planes: 20 bytes per predicted box.
//   demo_finalize : fixed-order fp64 reduction -> per-level partials {S_a, S_b, S_cls, S_conf, n_valid, T} + the outputs
//   demo_grad_dense / demo_grad_targets : backward (whole gradient written, single writer per row, no atomics)
#include "iou_grad.cuh"

namespace fvb {

constexpr int kDemoParts = 6;  // per level: S_a (box | xy), S_b (wh, flavour U), S_cls, S_conf, n_valid, T
constexpr int kCellThreads = 256;
constexpr int kCellBoxes = 4;  // boxes per thread
constexpr int kCellChunk = kCellThreads * kCellBoxes;
constexpr int kStage = 64;     // targets staged per round
constexpr int kTgtThreads = 256;

// F.binary_cross_entropy_with_logits element: (1 - t) * x - log_sigmoid(x), log_sigmoid(x) = min(x,0) - log1p(exp(-|x|))
FVB_HD float bce_logits(float x, float t) { return (1.0f - t) * x - (fminf(x, 0.0f) - log1pf(expf(-fabsf(x)))); }

struct DemoTarget {
  bool ok;
  int b, cls, gx, gy, a;
  float x, y, w, h;  // feature units
  float offx, offy, aw, ah;
};

// lossv3.py:46-62: scale to the feature map, best anchor by wh_iou_batch (first maximum), cell = floor(xy)
FVB_HD DemoTarget demo_target(const Geom& g, int l, const float* lab) {
  DemoTarget t;
  const float fw = (float)g.W[l], fh = (float)g.H[l];
  t.b = (int)lab[0];
  t.cls = (int)lab[1];
  t.x = lab[2] * fw;
  t.y = lab[3] * fh;
  t.w = lab[4] * fw;
  t.h = lab[5] * fh;
  float best = -1.0f;
  t.a = 0;
  t.aw = t.ah = 1.0f;
  for (int a = 0; a < g.A; ++a) {
    const float aw = g.aw[l][a] / g.stride[l], ah = g.ah[l][a] / g.stride[l];
    const float v = wh_iou(t.w, t.h, aw, ah, 1e-7f);
    if (a == 0 || v > best) {  // torch.max(dim=1): first maximum
      best = v;
      t.a = a;
      t.aw = aw;
      t.ah = ah;
    }
  }
  const float fx = floorf(t.x), fy = floorf(t.y);
  t.offx = t.x - fx;
  t.offy = t.y - fy;
  // the reference indexes with the raw floor (IndexError / negative wrap when a centre leaves the map); clamped here
  t.gx = (int)fminf(fmaxf(fx, 0.0f), (float)(g.W[l] - 1));
  t.gy = (int)fminf(fmaxf(fy, 0.0f), (float)(g.H[l] - 1));
  t.ok = t.b >= 0 && t.b < g.B && t.cls >= 0 && t.cls < g.K - 5;
  return t;
}

struct DemoParams {
  Geom g;
  const float* labels;
  int T, flavour;
  // workspace
  int* key;        // [L][T]  (b*A + a)*HW + cell, or -1
  int* img_off;    // [B+1]
  int* img_cur;    // [B]
  int* img_list;   // [T] target ids grouped by image
  double* tgt_ws;  // [L][tgt_blocks][4]
  double* cell_ws; // [cell_ctas][2]
  int tgt_blocks;
  int cell_cta_begin[FVB_MAX_LEVELS + 1];  // first demo_cells CTA of each level (B CTAs per level)
  int cell_chunks[FVB_MAX_LEVELS];
  long long mask_off[FVB_MAX_LEVELS];
  signed char* mask;  // [l][b][a][cell]: the caller's buffer or a workspace region
  double* partials;   // [L][kDemoParts]
  float* out;         // [3]
};

// Per-image target lists.  Labels grouped by image (collate_fn order, the normal case): the list is the identity and the
// offsets are the positions where the image index changes -- one pass, no atomics.  Otherwise: counting sort with atomic
// cursors, then every (short) list is sorted ascending so that the backward adds duplicates in target order.
__global__ void __launch_bounds__(1024) demo_prep_kernel(const DemoParams p) {
  __shared__ int warp_tot[33];
  __shared__ int carry;
  __shared__ int bad;
  const int B = p.g.B, T = p.T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    carry = 0;
    bad = 0;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int b = (int)p.labels[(size_t)t * 6];
    const int bp = t > 0 ? (int)p.labels[(size_t)(t - 1) * 6] : -1;
    if (b < bp || b < 0 || b >= B) bad = 1;  // out of order, or an image index outside the batch
  }
  __syncthreads();
  if (!bad) {
    for (int t = threadIdx.x; t <= T; t += blockDim.x) {
      const int b = t < T ? (int)p.labels[(size_t)t * 6] : B;
      const int bp = t > 0 ? (int)p.labels[(size_t)(t - 1) * 6] : -1;
      for (int i = bp + 1; i <= b; ++i) p.img_off[i] = t;  // images bp+1 .. b start at t (empty ones too)
      if (t < T) p.img_list[t] = t;
    }
    return;
  }
  for (int i = threadIdx.x; i <= B; i += blockDim.x) p.img_off[i] = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) p.img_cur[i] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int b = (int)p.labels[(size_t)t * 6];
    if (b >= 0 && b < B) atomicAdd(&p.img_off[b + 1], 1);
  }
  __syncthreads();
  // inclusive scan of the counts img_off[1..B], 1024 entries per round
  for (int base = 1; base <= B; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i <= B ? p.img_off[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int tv = warp_tot[lane];
      int ti = tv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - tv;
      if (lane == 31) warp_tot[32] = ti;
    }
    __syncthreads();
    if (i <= B) p.img_off[i] = carry + warp_tot[warp] + inc;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[32];
    __syncthreads();
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int b = (int)p.labels[(size_t)t * 6];
    if (b >= 0 && b < B) p.img_list[p.img_off[b] + atomicAdd(&p.img_cur[b], 1)] = t;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int* lst = p.img_list + p.img_off[b];
    const int n = p.img_off[b + 1] - p.img_off[b];
    for (int i = 1; i < n; ++i) {
      const int v = lst[i];
      int j = i - 1;
      while (j >= 0 && lst[j] > v) {
        lst[j + 1] = lst[j];
        --j;
      }
      lst[j + 1] = v;
    }
  }
}

// keys (b*A + a)*HW + cell of every (level, target), -1 for rows the loss skips: one thread each
__global__ void demo_keys_kernel(const DemoParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.T * p.g.L) return;
  const int l = i / p.T, t = i - l * p.T;
  const DemoTarget d = demo_target(p.g, l, p.labels + (size_t)t * 6);
  p.key[(size_t)l * p.T + t] = d.ok ? (d.b * p.g.A + d.a) * p.g.HW[l] + d.gy * p.g.W[l] + d.gx : -1;
}

// address of channel k of the matched row in the NCHW tensor
__device__ __forceinline__ size_t nchw_at(const Geom& g, int l, int b, int a, int k, int cell) {
  return ((size_t)(b * g.A + a) * g.K + k) * g.HW[l] + cell;
}

__global__ void __launch_bounds__(kTgtThreads) demo_targets_kernel(const DemoParams p) {
  __shared__ double part[kTgtThreads / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = blockIdx.y;
  const int t = blockIdx.x * (kTgtThreads / 32) + warp;
  double s_a = 0.0, s_b = 0.0, s_cls = 0.0, cnt = 0.0;
  if (t < p.T) {
    const DemoTarget d = demo_target(p.g, l, p.labels + (size_t)t * 6);
    if (d.ok) {
      const int K = p.g.K, cell = d.gy * p.g.W[l] + d.gx;
      const float* head = p.g.head[l];
      const float first = lane < K ? head[nchw_at(p.g, l, d.b, d.a, lane, cell)] : 0.0f;
      double sc = 0.0;
      for (int ch = lane; ch < K; ch += 32) {
        if (ch < 5) continue;
        const float v = ch < 32 ? first : head[nchw_at(p.g, l, d.b, d.a, ch, cell)];
        sc += (double)bce_logits(v, (ch - 5 == d.cls) ? 1.0f : 0.0f);   // lossv3.py:92-95
      }
      s_cls = warp_sum(sc);
      const float r0 = __shfl_sync(0xffffffffu, first, 0), r1 = __shfl_sync(0xffffffffu, first, 1);
      const float r2 = __shfl_sync(0xffffffffu, first, 2), r3 = __shfl_sync(0xffffffffu, first, 3);
      if (p.flavour == FVB_DEMO_LOSS_SHIP) {
        // predict_xywh = [sigmoid + grid, exp * anchor] (lossv3.py:65-69) vs the target in feature units (:85-88)
        const float px = sigmoid_precise(r0) + (float)d.gx, py = sigmoid_precise(r1) + (float)d.gy;
        const float pw = expf(r2) * d.aw, ph = expf(r3) * d.ah;
        const float ciou = iou_family<false>(xywh_to_xyxy(px, py, pw, ph), xywh_to_xyxy(d.x, d.y, d.w, d.h), FVB_CIOU,
                                             FVB_VARIANT_DEMO, 1e-7f);
        s_a = (double)(1.0f - ciou);
      } else {
        // yolov3_u/utils/lossv3.py:71-78: BCE-with-logits on the xy logits, MSE on the wh logits
        s_a = (double)bce_logits(r0, d.offx) + (double)bce_logits(r1, d.offy);
        const float tw = logf(d.w / d.aw + 1e-14f), th = logf(d.h / d.ah + 1e-14f);
        const float dw = r2 - tw, dh = r3 - th;
        s_b = (double)(dw * dw) + (double)(dh * dh);
      }
      cnt = 1.0;
    }
  }
  if (lane == 0) {
    part[warp][0] = s_a;
    part[warp][1] = s_b;
    part[warp][2] = s_cls;
    part[warp][3] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kTgtThreads / 32; ++w) s += part[w][threadIdx.x];
    p.tgt_ws[((size_t)l * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = s;
  }
}

// objectness BCE-with-logits of the dense stream: MUFU ex2 for exp(-|x|) (2^-22 relative), exact log1pf
__device__ __forceinline__ float bce_logits_stream(float x, float t) {
  const float e = ex2_approx(-fabsf(x) * kLog2e);
  return (1.0f - t) * x - (fminf(x, 0.0f) - log1pf(e));
}

// exact cell -> (row, column) without an integer division (cell < 2^22, W <= 2^11 by the geometry limits)
__device__ __forceinline__ void cell_yx(int cell, int W, float inv_w, int* gy, int* gx) {
  int y = __float2int_rz((float)cell * inv_w);
  int x = cell - y * W;
  if (x >= W) {
    x -= W;
    ++y;
  } else if (x < 0) {
    x += W;
    --y;
  }
  *gy = y;
  *gx = x;
}

// One CTA = one (level, image, anchor): it stages the image's targets ONCE and then walks the anchor's plane in chunks of
// kCellChunk cells (per-CTA fixed costs -- the dependent img_off -> img_list -> labels -> key loads, the barriers, the block
// reduction -- were paid per 1024 boxes before and capped the kernel at ~35 % active warps).  Inside a plane a = const and
// cells are consecutive: no per-box divisions.
__global__ void __launch_bounds__(kCellThreads) demo_cells_kernel(const DemoParams p) {
  __shared__ Box s_box[kStage];
  __shared__ int s_key[kStage];
  __shared__ double scratch[32];
  __shared__ int s_cnt[kCellThreads / 32];
  // big levels first: CTA x -> level L-1-(x / (B*A)) keeps the long CTAs at the front of the schedule
  const int HW0 = p.g.A * p.g.B;
  const int l = p.g.L - 1 - (int)blockIdx.x / HW0;
  const int rem = (int)blockIdx.x % HW0;
  const int b = rem / p.g.A, a = rem - b * p.g.A;
  const int HW = p.g.HW[l], W = p.g.W[l], K = p.g.K, A = p.g.A;
  const float inv_w = 1.0f / (float)W;
  const float* __restrict__ img = p.g.head[l] + (size_t)b * A * K * HW;
  const int chunks = (HW + kCellChunk - 1) / kCellChunk;
  const int beg = p.img_off[b], end = p.img_off[b + 1];
  const int key_base = b * A * HW;
  signed char* mimg = p.mask + p.mask_off[l] + (long long)b * A * HW;  // caller's mask, or a workspace copy
  float sum = 0.0f;  // a few dozen terms per thread at most: fp32 here, fp64 across threads
  int cnt = 0;

  for (int s0 = beg; s0 < max(end, beg + 1); s0 += kStage) {  // one round for <= kStage targets (also for none)
    const int ns = max(0, min(kStage, end - s0));
    const bool first_round = s0 == beg, last_round = s0 + kStage >= end;
    __syncthreads();
    if ((int)threadIdx.x < ns) {
      const int t = p.img_list[s0 + threadIdx.x];
      const DemoTarget d = demo_target(p.g, l, p.labels + (size_t)t * 6);
      s_box[threadIdx.x] = xywh_to_xyxy(d.x, d.y, d.w, d.h);
      const int key = p.key[(size_t)l * p.T + t];
      s_key[threadIdx.x] = key >= 0 ? key - key_base : -1;  // a*HW + cell inside this image
    }
    __syncthreads();
    {
      const float* __restrict__ plane = img + (size_t)a * K * HW;  // channel 0 of this anchor
      const float aw = p.g.aw[l][a] / p.g.stride[l], ah = p.g.ah[l][a] / p.g.stride[l];
      for (int chunk = 0; chunk < chunks; ++chunk) {
        const int cell0 = chunk * kCellChunk;
        // rows touched by this chunk: a target whose box misses them cannot make one of its cells a candidate
        const float row_lo = (float)(cell0 / W), row_hi = (float)(min(cell0 + kCellChunk, HW) - 1) / (float)W + 1.0f;
#pragma unroll
        for (int q = 0; q < kCellBoxes; ++q) {
          const int cell = cell0 + q * kCellThreads + threadIdx.x;
          if (cell >= HW) continue;
          const float conf = __ldg(plane + 4 * (size_t)HW + cell);
          int gy, gx;
          cell_yx(cell, W, inv_w, &gy, &gx);
          const float cgx = (float)gx, cgy = (float)gy;
          // mask state carried across target rounds (only when an image has more than kStage targets)
          signed char m = first_round ? (signed char)0 : mimg[a * HW + cell];
          bool pos = m > 0, ign = m < 0, c = false;
          const int jkey = a * HW + cell;
          // IoU > 0.5 needs the intersection to cover more than half of each box, hence more than half of the predicted
          // box's width and height: its centre (sigmoid + cell, inside [g, g+1]) must lie inside the target box.  Cells
          // that no target box touches skip the four box planes and all transcendental math (exact: conservative test).
          for (int i = 0; i < ns; ++i) {
            pos = pos || (s_key[i] == jkey);
            const Box tb = s_box[i];
            if (tb.y2 < row_lo || tb.y1 > row_hi) continue;  // uniform over the CTA
            c = c || (cgx <= tb.x2 && cgx + 1.0f >= tb.x1 && cgy <= tb.y2 && cgy + 1.0f >= tb.y1);
          }
          if (c && !ign) {
            const float* q0 = plane + cell;
            const float t0 = __ldg(q0), t1 = __ldg(q0 + HW), t2 = __ldg(q0 + 2 * (size_t)HW), t3 = __ldg(q0 + 3 * (size_t)HW);
            // lossv3.py:65-69
            const Box pb = xywh_to_xyxy(sigmoid_precise(t0) + cgx, sigmoid_precise(t1) + cgy, expf(t2) * aw, expf(t3) * ah);
            const float area_p = (pb.x2 - pb.x1) * (pb.y2 - pb.y1);
            for (int i = 0; i < ns && !ign; ++i) {
              const Box tb = s_box[i];
              const float inter = inter_area(pb, tb);
              if (inter > 0.0f) {
                // xywh_iou_batch (lossv3.py:106): inter / (area_p + area_t - inter + eps) > 0.5 (:110); the division is
                // only evaluated inside a guard band around the threshold
                const float uni = ((area_p + (tb.x2 - tb.x1) * (tb.y2 - tb.y1)) - inter) + 1e-7f;
                if (inter > 0.5000005f * uni) ign = true;
                else if (inter >= 0.4999995f * uni) ign = inter / uni > 0.5f;
              }
            }
          }
          // mask: -1 ignore (max IoU > 0.5, lossv3.py:110), then positives overwrite with 1 (:115)
          m = pos ? 1 : (ign ? -1 : 0);
          mimg[a * HW + cell] = m;
          if (last_round && m >= 0) {
            sum += bce_logits_stream(conf, (float)m);  // :118-120
            cnt += 1;
          }
        }
      }
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  const double total = block_sum((double)sum, scratch);  // syncs: s_cnt is complete too
  if (threadIdx.x == 0) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < kCellThreads / 32; ++w) c += s_cnt[w];
    p.cell_ws[(size_t)blockIdx.x * 2] = total;
    p.cell_ws[(size_t)blockIdx.x * 2 + 1] = (double)c;
  }
}

__device__ __forceinline__ void demo_combine(const Geom& g, const double* parts, int flavour, float* out) {
  double a = 0.0, b2 = 0.0, c = 0.0, d = 0.0;
  const int C = g.K - 5;
  for (int l = 0; l < g.L; ++l) {
    const double* q = parts + l * kDemoParts;
    const double T = q[5];
    if (flavour == FVB_DEMO_LOSS_SHIP) a += q[0] / T;  // (1 - ciou).mean()
    else {
      a += q[0] / (2.0 * T);                           // BCE-with-logits over [T,2]
      b2 += q[1] / (2.0 * T);                          // MSE over [T,2]
    }
    c += q[2] / (T * C);
    d += q[3] / q[4];
  }
  if (flavour == FVB_DEMO_LOSS_SHIP) {
    out[0] = (float)a;
    out[1] = (float)c;
    out[2] = (float)d;
  } else {
    out[0] = (float)(2.0 * a + b2 + c + d);  // loss_xy *= 2.0; loss_xy + loss_wh + loss_cls + loss_conf
    out[1] = 0.0f;
    out[2] = 0.0f;
  }
}

// One warp per (level, component): walks its column of per-CTA partials in a fixed order (lane-strided, two loads in flight,
// shuffle tree) -- no block barrier until the single one before the combine.
__global__ void __launch_bounds__(1024) demo_finalize_kernel(const DemoParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = warp / kDemoParts, c = warp - l * kDemoParts;
  if (l < p.g.L) {
    // partial slot c of level l: 0..2 <- tgt_ws comps 0..2, 3 <- cell sums, 4 <- cell counts, 5 <- tgt_ws comp 3 (T)
    const double* col;
    int n, stride;
    if (c == 3 || c == 4) {
      col = p.cell_ws + (size_t)p.cell_cta_begin[l] * 2 + (c - 3);
      n = p.g.B * p.g.A;
      stride = 2;
    } else {
      col = p.tgt_ws + (size_t)l * p.tgt_blocks * 4 + (c == 5 ? 3 : c);
      n = p.tgt_blocks;
      stride = 4;
    }
    double s0 = 0.0, s1 = 0.0;
    int i = lane;
    for (; i + 32 < n; i += 64) {
      s0 += col[(size_t)i * stride];
      s1 += col[(size_t)(i + 32) * stride];
    }
    if (i < n) s0 += col[(size_t)i * stride];
    const double tot = warp_sum(s0 + s1);
    if (lane == 0) p.partials[l * kDemoParts + c] = tot;
  }
  __threadfence_block();
  __syncthreads();
  if (threadIdx.x == 0 && p.out != nullptr) demo_combine(p.g, p.partials, p.flavour, p.out);
}

struct DemoCombineParams {
  Geom g;
  const double* partials;
  int flavour;
  float* out;
};
__global__ void demo_combine_kernel(const DemoCombineParams p) {
  if (threadIdx.x == 0) demo_combine(p.g, p.partials, p.flavour, p.out);
}

// ---- backward ---------------------------------------------------------------------------------------------------------
struct DemoGradParams {
  Geom g;
  const float* labels;
  int T, flavour;
  const int* key;             // [L][T] (demo_prep)
  const int* img_off;         // [B+1] per-image target lists, ascending target ids (demo_prep)
  const int* img_list;
  const signed char* mask;    // forward's mask
  long long mask_off[FVB_MAX_LEVELS];
  const double* partials;     // [L][kDemoParts] (all-reduced under data parallelism)
  const float* grad_out;      // [3] (ship) / [1] (u), or NULL for ones
  float* grad[FVB_MAX_LEVELS];
  long long lvl_floats[FVB_MAX_LEVELS];
  int cta_begin[FVB_MAX_LEVELS + 1];
};

constexpr int kDgThreads = 256;
constexpr int kDgIters = 16;
constexpr int kDgChunk = kDgThreads * kDgIters * 4;

__device__ __forceinline__ float demo_up(const DemoGradParams& p, int which) {
  if (!p.grad_out) return 1.0f;
  return p.flavour == FVB_DEMO_LOSS_SHIP ? p.grad_out[which] : p.grad_out[0];
}

// whole gradient of one level tensor [B, A*K, H, W]: zero except the objectness planes, (sigmoid(x) - t) / n_valid there
template <bool VEC>
__global__ void __launch_bounds__(kDgThreads) demo_grad_dense_kernel(const DemoGradParams p) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < FVB_MAX_LEVELS; ++i)
    if (i < p.g.L && (int)blockIdx.x >= p.cta_begin[i]) l = i;
  const int HW = p.g.HW[l], K = p.g.K;
  const long long n = p.lvl_floats[l];
  const long long base = (long long)((int)blockIdx.x - p.cta_begin[l]) * kDgChunk;
  const float* __restrict__ head = p.g.head[l];
  float* __restrict__ out = p.grad[l];
  const signed char* __restrict__ mask = p.mask + p.mask_off[l];
  const float coef = (float)((double)demo_up(p, 2) / p.partials[l * kDemoParts + 4]);
  auto elem = [&](long long i, int plane, int cell) -> float {
    const int ba = plane / K;
    if (plane - ba * K != 4) return 0.0f;
    const signed char m = mask[(long long)ba * HW + cell];
    if (m < 0) return 0.0f;
    return coef * (sigmoid_precise(__ldg(head + i)) - (float)m);
  };
  const int step = VEC ? 4 * kDgThreads : kDgThreads;
  long long i = base + (long long)threadIdx.x * (VEC ? 4 : 1);
  int plane = (int)(i / HW);  // < B*A*K
  int cell = (int)(i - (long long)plane * HW);
  const int step_p = step / HW, step_c = step % HW;
  for (int it = 0; it < (VEC ? kDgIters : kDgIters * 4); ++it) {
    if (VEC) {
      if (i + 3 < n) {
        float4 v;
        if (cell + 3 < HW) {
          if (plane % K != 4) v = make_float4(0.f, 0.f, 0.f, 0.f);
          else v = make_float4(elem(i, plane, cell), elem(i + 1, plane, cell + 1), elem(i + 2, plane, cell + 2), elem(i + 3, plane, cell + 3));
        } else {
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool wrap = cell + q >= HW;
            e[q] = elem(i + q, wrap ? plane + 1 : plane, wrap ? cell + q - HW : cell + q);
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
        __stcs(reinterpret_cast<float4*>(out + i), v);
      } else {
        for (int q = 0; q < 4 && i + q < n; ++q) {
          const bool wrap = cell + q >= HW;
          out[i + q] = elem(i + q, wrap ? plane + 1 : plane, wrap ? cell + q - HW : cell + q);
        }
      }
    } else if (i < n) {
      out[i] = elem(i, plane, cell);
    }
    i += step;
    plane += step_p;
    cell += step_c;
    if (cell >= HW) {
      cell -= HW;
      plane += 1;
    }
  }
}

__global__ void __launch_bounds__(kTgtThreads) demo_grad_targets_kernel(const DemoGradParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = blockIdx.y;
  const int t = blockIdx.x * (kTgtThreads / 32) + warp;
  if (t >= p.T) return;
  const int* keys = p.key + (size_t)l * p.T;
  const int key = keys[t];
  if (key < 0) return;
  const DemoTarget self = demo_target(p.g, l, p.labels + (size_t)t * 6);
  const int K = p.g.K, C = K - 5, HW = p.g.HW[l];
  const int cell = self.gy * p.g.W[l] + self.gx;
  const float* head = p.g.head[l];
  float* grad = p.grad[l];
  const double Tn = p.partials[l * kDemoParts + 5];
  const float w_cls = (float)((double)demo_up(p, 1) / (Tn * C));
  const float w_box = (float)((double)demo_up(p, 0) / Tn);
  constexpr int kRegs = 4;  // channels lane + 32*j of the row kept in registers (K <= 128); wider rows read-modify-write
  const bool in_regs = K <= 32 * kRegs;
  float rv[kRegs], sg[kRegs], acc[kRegs];
#pragma unroll
  for (int j = 0; j < kRegs; ++j) rv[j] = (lane + 32 * j < K) ? __ldg(head + nchw_at(p.g, l, self.b, self.a, lane + 32 * j, cell)) : 0.0f;
  // (the row is requested before the duplicate scan so that its DRAM latency overlaps it)
  // Duplicates of a (cell, anchor) live in the same image: only that image's (sorted) target list is scanned.  The last
  // target of the key owns the row; earlier duplicates are added by it in target order.
  const int img = (int)p.labels[(size_t)t * 6];
  const int lb = p.img_off[img], le = p.img_off[img + 1];
  for (int b0 = lb; b0 < le; b0 += 32) {
    const int i = b0 + lane;
    const int t2 = i < le ? p.img_list[i] : -1;
    if (__any_sync(0xffffffffu, t2 > t && keys[t2] == key)) return;
  }
#pragma unroll
  for (int j = 0; j < kRegs; ++j) {
    sg[j] = sigmoid_precise(rv[j]);
    acc[j] = 0.0f;
  }
  const float first = rv[0];
  const float r0 = __shfl_sync(0xffffffffu, first, 0), r1 = __shfl_sync(0xffffffffu, first, 1);
  const float r2 = __shfl_sync(0xffffffffu, first, 2), r3 = __shfl_sync(0xffffffffu, first, 3);

  auto add_target = [&](int t2) {
    const DemoTarget d = t2 == t ? self : demo_target(p.g, l, p.labels + (size_t)t2 * 6);
    float g4[4];
    if (p.flavour == FVB_DEMO_LOSS_SHIP) {
      const float sx = sigmoid_precise(r0), sy = sigmoid_precise(r1);
      const float pw = expf(r2) * d.aw, ph = expf(r3) * d.ah;
      BoxGrad ga = zero_grad(), gb = zero_grad();
      iou_family_grad(xywh_to_xyxy(sx + (float)d.gx, sy + (float)d.gy, pw, ph), xywh_to_xyxy(d.x, d.y, d.w, d.h), FVB_CIOU,
                      FVB_VARIANT_DEMO, 1e-7f, 0.0f - w_box, ga, gb);
      float gx, gy, gw, gh;
      xyxy_grad_to_xywh(ga, &gx, &gy, &gw, &gh);
      g4[0] = gx * ((1.0f - sx) * sx);
      g4[1] = gy * ((1.0f - sy) * sy);
      g4[2] = gw * pw;
      g4[3] = gh * ph;
    } else {
      // total = 2 * mean_{[T,2]} bce_logits(xy) + mean_{[T,2]} (wh - log(w/a))^2 + ...
      const float w_xy = (float)((double)demo_up(p, 0) * 2.0 / (2.0 * Tn)), w_wh = (float)((double)demo_up(p, 0) / (2.0 * Tn));
      g4[0] = w_xy * (sigmoid_precise(r0) - d.offx);
      g4[1] = w_xy * (sigmoid_precise(r1) - d.offy);
      g4[2] = w_wh * (2.0f * (r2 - logf(d.w / d.aw + 1e-14f)));
      g4[3] = w_wh * (2.0f * (r3 - logf(d.h / d.ah + 1e-14f)));
    }
    if (in_regs) {
      // this warp is the only writer of the row after the dense pass (zeros there): sum in registers, in target order
#pragma unroll
      for (int j = 0; j < kRegs; ++j) {
        const int ch = lane + 32 * j;
        float add;
        if (j == 0 && lane < 4) add = lane == 0 ? g4[0] : (lane == 1 ? g4[1] : (lane == 2 ? g4[2] : g4[3]));
        else add = w_cls * (sg[j] - ((ch - 5 == d.cls) ? 1.0f : 0.0f));
        acc[j] += add;
      }
    } else {
      for (int ch = lane; ch < K; ch += 32) {
        if (ch == 4) continue;  // objectness: written by the dense pass (positive mask)
        float add;
        if (ch < 4) {
          add = ch == 0 ? g4[0] : (ch == 1 ? g4[1] : (ch == 2 ? g4[2] : g4[3]));
        } else {
          add = w_cls * (sigmoid_precise(head[nchw_at(p.g, l, self.b, self.a, ch, cell)]) - ((ch - 5 == d.cls) ? 1.0f : 0.0f));
        }
        grad[nchw_at(p.g, l, self.b, self.a, ch, cell)] += add;
      }
    }
  };
  for (int b0 = lb; b0 < le; b0 += 32) {
    const int i = b0 + lane;
    const int t2 = i < le ? p.img_list[i] : -1;
    unsigned mk = __ballot_sync(0xffffffffu, t2 >= 0 && t2 < t && keys[t2] == key);
    while (mk) {
      const int src = __ffs(mk) - 1;
      mk &= mk - 1;
      add_target(__shfl_sync(0xffffffffu, t2, src));
    }
  }
  add_target(t);
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
      const int ch = lane + 32 * j;
      if (ch < K && ch != 4) grad[nchw_at(p.g, l, self.b, self.a, ch, cell)] = acc[j];
    }
  }
  (void)HW;
}

static size_t align_up_(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DemoLayout {
  size_t key, img_off, img_cur, img_list, tgt_ws, cell_ws, mask, total;
  int tgt_blocks, cell_ctas;
};

static DemoLayout demo_layout(const Geom& g, long long T, DemoParams* p) {
  DemoLayout L;
  size_t o = 0;
  L.key = o;      o = align_up_(o + (size_t)g.L * (size_t)(T > 0 ? T : 1) * 4, 256);
  L.img_off = o;  o = align_up_(o + (size_t)(g.B + 1) * 4, 256);
  L.img_cur = o;  o = align_up_(o + (size_t)(g.B > 0 ? g.B : 1) * 4, 256);
  L.img_list = o; o = align_up_(o + (size_t)(T > 0 ? T : 1) * 4, 256);
  L.tgt_blocks = (int)((T + kTgtThreads / 32 - 1) / (kTgtThreads / 32));
  L.tgt_ws = o;   o = align_up_(o + (size_t)g.L * (size_t)(L.tgt_blocks > 0 ? L.tgt_blocks : 1) * 4 * 8, 256);
  // one demo_cells CTA per (level, image, anchor); CTA x handles level L-1-x/(B*A): level l owns B*A consecutive CTAs
  const int ctas = g.L * g.B * g.A;
  if (p)
    for (int l = 0; l < FVB_MAX_LEVELS; ++l) {
      p->cell_cta_begin[l] = l < g.L ? (g.L - 1 - l) * g.B * g.A : 0;
      p->cell_chunks[l] = 1;
    }
  L.cell_ctas = ctas;
  L.cell_ws = o;  o = align_up_(o + (size_t)(ctas > 0 ? ctas : 1) * 16, 256);
  L.mask = o;     o = align_up_(o + (size_t)g.B * g.row_off[g.L], 256);  // used when the caller does not ask for the mask
  L.total = o + 256;
  return L;
}

static int demo_common_checks(const Geom& g, const float* const* d_heads, int flavour, const char* who) {
  FVB_REQUIRE(g.nchw, "%s: heads must be the conv outputs [B,A*K,H,W] (FVB_HEAD_NCHW)", who);
  FVB_REQUIRE(flavour == FVB_DEMO_LOSS_SHIP || flavour == FVB_DEMO_LOSS_U, "%s: unknown flavour %d", who, flavour);
  FVB_REQUIRE(g.B >= 1, "%s: empty batch", who);
  for (int l = 0; l < g.L; ++l) {
    FVB_REQUIRE(d_heads[l] != nullptr, "%s: head %d is NULL", who, l);
    FVB_REQUIRE((long long)g.B * g.A * g.HW[l] < (1ll << 31), "%s: level %d has too many boxes for 32-bit keys", who, l);
  }
  return FVB_OK;
}

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_demo_loss_workspace_bytes(const fvb_yolo_geom* geom, int64_t num_labels) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return 0;
  return demo_layout(g, num_labels, nullptr).total;
}

extern "C" int64_t fvb_demo_loss_mask_bytes(const fvb_yolo_geom* geom) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return 0;
  return (int64_t)g.B * g.row_off[g.L];
}

extern "C" int fvb_demo_loss_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                 int64_t num_labels, int flavour, double* d_partials, float* d_out, int8_t* d_mask,
                                 void* d_ws, void* stream) {
  FVB_REQUIRE(d_heads && d_partials && d_ws, "demo_loss: NULL pointer");
  FVB_REQUIRE(num_labels >= 0 && num_labels < (1ll << 24), "demo_loss: num_labels=%lld", (long long)num_labels);
  FVB_REQUIRE(num_labels == 0 || d_labels, "demo_loss: labels NULL");
  FVB_REQUIRE(((uintptr_t)d_ws & 255) == 0, "demo_loss: workspace must be 256-byte aligned");
  DemoParams p;
  int rc = make_geom(geom, d_heads, &p.g);
  if (rc != FVB_OK) return rc;
  rc = demo_common_checks(p.g, d_heads, flavour, "demo_loss");
  if (rc != FVB_OK) return rc;
  const DemoLayout L = demo_layout(p.g, num_labels, &p);
  unsigned char* w = (unsigned char*)d_ws;
  p.labels = d_labels;
  p.T = (int)num_labels;
  p.flavour = flavour;
  p.key = (int*)(w + L.key);
  p.img_off = (int*)(w + L.img_off);
  p.img_cur = (int*)(w + L.img_cur);
  p.img_list = (int*)(w + L.img_list);
  p.tgt_ws = (double*)(w + L.tgt_ws);
  p.cell_ws = (double*)(w + L.cell_ws);
  p.tgt_blocks = L.tgt_blocks;
  p.mask = d_mask ? (signed char*)d_mask : (signed char*)(w + L.mask);
  for (int l = 0; l < FVB_MAX_LEVELS; ++l) p.mask_off[l] = l < p.g.L ? (long long)p.g.B * p.g.row_off[l] : 0;
  p.partials = d_partials;
  p.out = d_out;
  cudaStream_t s = (cudaStream_t)stream;
  demo_prep_kernel<<<1, 1024, 0, s>>>(p);
  count_launch();
  if (p.T > 0) {
    demo_keys_kernel<<<(unsigned)(((long long)p.T * p.g.L + 255) / 256), 256, 0, s>>>(p);
    count_launch();
    dim3 grid((unsigned)L.tgt_blocks, (unsigned)p.g.L);
    demo_targets_kernel<<<grid, kTgtThreads, 0, s>>>(p);
    count_launch();
  }
  demo_cells_kernel<<<(unsigned)L.cell_ctas, kCellThreads, 0, s>>>(p);
  demo_finalize_kernel<<<1, 1024, 0, s>>>(p);
  count_launch(2);
  return check_launch("demo_loss");
}

extern "C" int fvb_demo_loss_combine_f32(const fvb_yolo_geom* geom, const double* d_partials, int flavour, float* d_out,
                                         void* stream) {
  FVB_REQUIRE(d_partials && d_out, "demo_loss_combine: NULL pointer");
  DemoCombineParams p;
  int rc = make_geom(geom, nullptr, &p.g);
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(flavour == FVB_DEMO_LOSS_SHIP || flavour == FVB_DEMO_LOSS_U, "demo_loss_combine: unknown flavour %d", flavour);
  p.partials = d_partials;
  p.flavour = flavour;
  p.out = d_out;
  demo_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  return check_launch("demo_combine_kernel");
}

extern "C" int fvb_demo_loss_backward_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                          int64_t num_labels, int flavour, const double* d_partials, const int8_t* d_mask,
                                          const float* d_grad_out, float* const* d_grad_heads, void* d_ws, void* stream) {
  FVB_REQUIRE(d_heads && d_partials && d_mask && d_grad_heads && d_ws, "demo_loss_backward: NULL pointer");
  FVB_REQUIRE(num_labels >= 0 && num_labels < (1ll << 24), "demo_loss_backward: num_labels=%lld", (long long)num_labels);
  FVB_REQUIRE(num_labels == 0 || d_labels, "demo_loss_backward: labels NULL");
  FVB_REQUIRE(((uintptr_t)d_ws & 255) == 0, "demo_loss_backward: workspace must be 256-byte aligned");
  DemoParams fp;
  int rc = make_geom(geom, d_heads, &fp.g);
  if (rc != FVB_OK) return rc;
  rc = demo_common_checks(fp.g, d_heads, flavour, "demo_loss_backward");
  if (rc != FVB_OK) return rc;
  const DemoLayout L = demo_layout(fp.g, num_labels, &fp);
  unsigned char* w = (unsigned char*)d_ws;
  fp.labels = d_labels;
  fp.T = (int)num_labels;
  fp.flavour = flavour;
  fp.key = (int*)(w + L.key);
  fp.img_off = (int*)(w + L.img_off);
  fp.img_cur = (int*)(w + L.img_cur);
  fp.img_list = (int*)(w + L.img_list);
  DemoGradParams p;
  p.g = fp.g;
  p.labels = d_labels;
  p.T = fp.T;
  p.flavour = flavour;
  p.key = fp.key;
  p.img_off = fp.img_off;
  p.img_list = fp.img_list;
  p.mask = (const signed char*)d_mask;
  p.partials = d_partials;
  p.grad_out = d_grad_out;
  bool vec = true;
  long long ctas = 0;
  for (int l = 0; l < FVB_MAX_LEVELS; ++l) {
    p.grad[l] = nullptr;
    p.lvl_floats[l] = 0;
    p.mask_off[l] = 0;
  }
  for (int l = 0; l < p.g.L; ++l) {
    FVB_REQUIRE(d_grad_heads[l] != nullptr, "demo_loss_backward: grad %d is NULL", l);
    p.grad[l] = d_grad_heads[l];
    p.lvl_floats[l] = (long long)p.g.B * p.g.A * p.g.K * p.g.HW[l];
    p.mask_off[l] = (long long)p.g.B * p.g.row_off[l];
    p.cta_begin[l] = (int)ctas;
    ctas += (p.lvl_floats[l] + kDgChunk - 1) / kDgChunk;
    if (((uintptr_t)d_grad_heads[l] & 15) != 0) vec = false;
  }
  for (int l = p.g.L; l <= FVB_MAX_LEVELS; ++l) p.cta_begin[l] = (int)ctas;
  FVB_REQUIRE(ctas < (1ll << 31), "demo_loss_backward: tensor too large for one launch");
  cudaStream_t s = (cudaStream_t)stream;
  if (vec) demo_grad_dense_kernel<true><<<(unsigned)ctas, kDgThreads, 0, s>>>(p);
  else demo_grad_dense_kernel<false><<<(unsigned)ctas, kDgThreads, 0, s>>>(p);
  count_launch();
  if (p.T > 0) {
    demo_prep_kernel<<<1, 1024, 0, s>>>(fp);
    demo_keys_kernel<<<(unsigned)(((long long)fp.T * fp.g.L + 255) / 256), 256, 0, s>>>(fp);
    dim3 grid((unsigned)L.tgt_blocks, (unsigned)p.g.L);
    demo_grad_targets_kernel<<<grid, kTgtThreads, 0, s>>>(p);
    count_launch(3);
  }
  return check_launch("demo_loss_backward");
}
