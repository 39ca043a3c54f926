// K4: YOLOv3 target assignment + loss, fused (replaces ~200 ATen launches and >=6 host syncs of
// loss/yolov3_loss.py:29-124).
//
//   loss_prep      : are the labels grouped by image (collate_fn order)?  -> lets the duplicate-cell
//                    scan stop early; any order stays correct.
//   loss_match     : one warp per (level, target, anchor): ratio test, cell, coalesced gather of the
//                    matched raw row, class BCE (lanes over classes, shuffle reduction), CIoU box
//                    term, IoU objectness target with the reference's "last write wins" rule for
//                    duplicate cells resolved deterministically.
//   conf_stream    : (only when the decode kernel did not already fuse it) zero-target objectness
//                    BCE over every cell -- the one HBM-streaming part of the loss.
//   loss_finalize  : fixed-order fp64 reduction of all partials -> {S_cls,S_box,S_conf,M} per level
//                    (what data-parallel ranks all-reduce) and, optionally, the scalar.
// The objectness sum is  sum_cells bce(p,0) + sum_winners (bce(p,iou) - bce(p,0)) : the first part
// is a pure stream, the second touches only matched rows.
#include "common.cuh"
#include "loss_common.cuh"

namespace fvb {

constexpr int kLossThreads = 256;
constexpr int kStreamRows = 1024;  // rows per conf_stream CTA

struct LossParams {
  Geom g;
  const float* labels;
  int T;
  double* match_ws;  // [L][blocks][4] = per-CTA sums of {S_cls, box term, conf correction, matched}
  int* flags;        // [0] labels grouped by image
  const double* conf0;               // zero-target objectness partials (decode tiles or conf_stream chunks), level-major
  int conf_begin[FVB_MAX_LEVELS];    // each level's contiguous range; CTA x of level l folds slice x of it into its sums
  int conf_end[FVB_MAX_LEVELS];
  int dense_precise;                 // the dense zero-target partials were formed with the PRECISE sigmoid (decode precise=1)
};

__global__ void loss_prep_kernel(const float* labels, int T, int* flags) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int t = threadIdx.x; t + 1 < T; t += blockDim.x)
    if ((int)labels[(size_t)t * 6] > (int)labels[(size_t)(t + 1) * 6]) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    flags[0] = bad ? 0 : 1;
    flags[1] = 0;  // arrival counter of loss_finish_kernel
  }
}

__global__ void __launch_bounds__(kLossThreads) loss_match_kernel(const LossParams p) {
  // grid.x: blocks of 8 (target, anchor) pairs, grid.y: level; one warp per (level, target, anchor)
  __shared__ double part[kLossThreads / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int TA = p.T * p.g.A;
  const int l = blockIdx.y;
  const int ta = blockIdx.x * (kLossThreads / 32) + warp;
  double r_cls = 0.0, r_box = 0.0, r_conf = 0.0, r_m = 0.0;

  if (ta < TA) {
    const int t = ta / p.g.A, a = ta - t * p.g.A;
    const TargetCell c = target_cell(p.g, l, p.labels + (size_t)t * 6, a);
    if (c.match && c.b >= 0 && c.b < p.g.B) {  // (an out-of-range batch index raises in the reference)
      const int K = p.g.K;
      const float* row = p.g.head[l] + ((((size_t)c.b * p.g.A + a) * p.g.H[l] + c.gy) * p.g.W[l] + c.gx) * K;

      // class BCE over the matched row (:50-52), lanes over channels
      const float first = lane < K ? row[lane] : 0.0f;  // channels 0..31 (K >= 6)
      double s_cls = 0.0;
      for (int ch = lane; ch < K; ch += 32) {
        const float v = ch < 32 ? first : row[ch];
        if (ch >= 5) {
          const float prob = sigmoid_precise(v);
          const float tgt = (ch - 5 == c.cls) ? 1.0f : 0.0f;
          s_cls += (double)bce_term(prob, tgt);
        }
      }
      s_cls = warp_sum(s_cls);

      const float r0 = __shfl_sync(0xffffffffu, first, 0), r1 = __shfl_sync(0xffffffffu, first, 1);
      const float r2 = __shfl_sync(0xffffffffu, first, 2), r3 = __shfl_sync(0xffffffffu, first, 3);
      const float r4 = __shfl_sync(0xffffffffu, first, 4);

      // duplicate cells: targets_conf[...] = iou (:61) keeps the LAST match in (t,a) order on CPU.
      // This match loses iff a later target of the same image hits the same cell with the same anchor.
      const bool grouped = p.flags[0] != 0;
      bool loser = false;
      for (int t2b = t + 1; t2b < p.T; t2b += 32) {
        const int t2 = t2b + lane;
        bool hit = false;
        int b2 = 0x7fffffff;
        if (t2 < p.T) {
          const float* lab2 = p.labels + (size_t)t2 * 6;
          b2 = (int)lab2[0];
          if (b2 == c.b) {
            const TargetCell c2 = target_cell(p.g, l, lab2, a);
            hit = c2.match && c2.gx == c.gx && c2.gy == c.gy;
          }
        }
        if (__any_sync(0xffffffffu, hit)) {
          loser = true;
          break;
        }
        // grouped labels: once a whole row of 32 is past image b nothing later can collide
        if (grouped && __all_sync(0xffffffffu, b2 > c.b)) break;
      }

      // predicted box in cell units (:54-56) vs [offset, wh] (:57), both xywh (:58,:60)
      const float px = sigmoid_precise(r0), py = sigmoid_precise(r1);
      const float pw = expf(r2) * c.aw, ph = expf(r3) * c.ah;
      const Box pb = xywh_to_xyxy(px, py, pw, ph);
      const Box tb = xywh_to_xyxy(c.offx, c.offy, c.tw, c.th);
      const float eps = 1e-7f;
      const float ciou = iou_family<false>(pb, tb, FVB_CIOU, FVB_VARIANT_LIB, eps);
      const float iou = iou_plain<true>(pb, tb, eps);
      if (!loser) {
        // the dense zero-target sum (decode tiles / conf_stream) holds bce(sigmoid_fast(t4), 0) for this cell: take
        // exactly that back out and put the true term in
        // (the same sigmoid form the producer of the dense sum used -- for large objectness logits 1 ulp of p is ~20 % of the term)
        const float p_dense = p.dense_precise ? sigmoid_precise(r4) : sigmoid_fast(r4);
        r_conf = (double)bce_term(sigmoid_precise(r4), iou) - (double)bce_term_zero(p_dense);
      }
      r_cls = s_cls;
      r_box = (double)(1.0f - ciou);
      r_m = 1.0;
    }
  }
  if (lane == 0) {
    part[warp][0] = r_cls;
    part[warp][1] = r_box;
    part[warp][2] = r_conf;
    part[warp][3] = r_m;
  }
  // this CTA's slice of the level's dense objectness partials (fixed order: thread-strided, then the block tree)
  __shared__ double scratch[32];
  double cs = 0.0;
  {
    const int n = p.conf_end[l] - p.conf_begin[l];
    const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
    const int beg = p.conf_begin[l] + (int)blockIdx.x * per;
    const int end = min(beg + per, p.conf_end[l]);
    for (int i = beg + (int)threadIdx.x; i < end; i += kLossThreads) cs += p.conf0[i];
  }
  cs = block_sum(cs, scratch);  // syncs: part[] is complete too
  if (threadIdx.x < 4) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += part[w][threadIdx.x];  // fixed order
    if (threadIdx.x == 2) s += cs;  // block_sum leaves the total in every lane of warp 0
    p.match_ws[((size_t)l * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = s;
  }
}

// zero-target objectness BCE, one thread per cell row; partials [L][B][chunks_l]
struct StreamParams {
  Geom g;
  int chunk_off[FVB_MAX_LEVELS + 1];  // first partial of each level
  int chunks[FVB_MAX_LEVELS];         // chunks per image on each level
  double* partials;
  float* save;                        // optional compact copy of the objectness logits, level-major [l][b][row] (for the backward)
  long long save_off[FVB_MAX_LEVELS];
};

__global__ void __launch_bounds__(256) conf_stream_kernel(const StreamParams p) {
  // blockIdx.x enumerates (level, image, chunk)
  int id = blockIdx.x, l = 0;
#pragma unroll
  for (int i = 0; i < FVB_MAX_LEVELS - 1; ++i)
    if (i < p.g.L - 1 && id >= p.chunk_off[i + 1]) l = i + 1;
  id -= p.chunk_off[l];
  const int b = id / p.chunks[l], ch = id - b * p.chunks[l];
  const int rows = p.g.A * p.g.HW[l];
  const float* base = p.g.head[l] + (size_t)b * rows * p.g.K + 4;
  double acc = 0.0;
  const int r0 = ch * kStreamRows;
#pragma unroll
  for (int i = 0; i < kStreamRows / 256; ++i) {
    int r = r0 + i * 256 + threadIdx.x;
    // same arithmetic as the fused partials of the decode kernel: objectness as decode stores it, target 0
    if (r < rows) {
      const float logit = __ldg(base + (size_t)r * p.g.K);
      acc += (double)bce_term_zero(sigmoid_fast(logit));
      if (p.save != nullptr) p.save[p.save_off[l] + (long long)b * rows + r] = logit;
    }
  }
  __shared__ double scratch[32];
  double s = block_sum(acc, scratch);
  if (threadIdx.x == 0) p.partials[blockIdx.x] = s;
}

struct FinalizeParams {
  Geom g;
  int match_blocks;  // per-level CTA count of loss_match (0 when there are no labels)
  const double* match_ws;
  const double* conf0;              // zero-target objectness partials, level-major (decode or conf_stream layout)
  int level_begin[FVB_MAX_LEVELS];  // contiguous partial range of each level
  int level_end[FVB_MAX_LEVELS];
  double* partials;  // [L][4]
  float* out_loss;
  float r_box, r_conf, r_cls;
  long long batch_global;
};

__device__ __forceinline__ float combine_loss(const Geom& g, const double* partials, long long batch_global,
                                              float r_box, float r_conf, float r_cls) {
  double tot = 0.0;
  const int C = g.K - 5;
  for (int l = 0; l < g.L; ++l) {
    double s_cls = partials[l * 4 + 0], s_box = partials[l * 4 + 1], s_conf = partials[l * 4 + 2], m = partials[l * 4 + 3];
    if (m > 0.0) tot += (double)r_cls * s_cls / (m * C) + (double)r_box * s_box / m;  // yolov3_loss.py:49-58
    double cells = (double)batch_global * g.A * g.HW[l];
    tot += (double)r_conf * s_conf / cells;                                            // :63-64
  }
  return (float)(tot * (double)batch_global);                                          // :66-72
}

// sum of a contiguous double range with 4 independent loads in flight per thread (single CTA)
__device__ __forceinline__ double strided_sum(const double* v, int n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const int step = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * step < n; i += 4 * step) {
    double a = v[i], b = v[i + step], c = v[i + 2 * step], d = v[i + 3 * step];
    s0 += a; s1 += b; s2 += c; s3 += d;
  }
  for (; i < n; i += step) s0 += v[i];
  return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(1024) loss_finalize_kernel(const FinalizeParams p) {
  __shared__ double scratch[32];
  const int TA = p.match_blocks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (TA > 0) {
    // loss_match already folded the dense objectness partials into its per-CTA sums: one warp per (level, component)
    // walks that column in a fixed order -- no block barrier on the way
    const int l = warp >> 2, c = warp & 3;
    if (l < p.g.L) {
      const double* ws = p.match_ws + (size_t)l * TA * 4 + c;
      double s0 = 0.0, s1 = 0.0;
      int i = lane;
      for (; i + 32 < TA; i += 64) {
        s0 += ws[(size_t)i * 4];
        s1 += ws[(size_t)(i + 32) * 4];
      }
      if (i < TA) s0 += ws[(size_t)i * 4];
      const double tot = warp_sum(s0 + s1);
      if (lane == 0) p.partials[l * 4 + c] = tot;
    }
  } else {
    // no labels: only the objectness term exists (yolov3_loss.py:63-64)
    for (int l = 0; l < p.g.L; ++l) {
      double c0 = strided_sum(p.conf0 + p.level_begin[l], p.level_end[l] - p.level_begin[l]);
      c0 = block_sum(c0, scratch);
      if (threadIdx.x == 0) {
        p.partials[l * 4 + 0] = 0.0;
        p.partials[l * 4 + 1] = 0.0;
        p.partials[l * 4 + 2] = c0;
        p.partials[l * 4 + 3] = 0.0;
      }
    }
  }
  __threadfence_block();
  __syncthreads();
  if (threadIdx.x == 0 && p.out_loss != nullptr)
    p.out_loss[0] = combine_loss(p.g, p.partials, p.batch_global, p.r_box, p.r_conf, p.r_cls);
}

// The second half of the two-part loss (fvb_yolov3_loss_match_f32 / fvb_yolov3_loss_finish_f32): loss_match ran EARLY (it reads
// only raw heads and labels) without folding the decode's objectness partials, so this kernel does that sum -- grid (kFinishCtas,
// L): CTA (x, l) sums slice x of level l's partials in a fixed order -- and the last CTA to arrive reduces the slices and
// loss_match's per-CTA columns (again in index order: the result does not depend on which CTA is last) and forms the scalar.
constexpr int kFinishCtas = 48;
struct FinishParams {
  FinalizeParams f;
  double* slices;  // [L][kFinishCtas]
  int* counter;    // zero on entry (loss_prep), left zero
};

__global__ void __launch_bounds__(256) loss_finish_kernel(const FinishParams p) {
  __shared__ double scratch[32];
  __shared__ int last;
  const int l = blockIdx.y, x = blockIdx.x;
  {
    const int n = p.f.level_end[l] - p.f.level_begin[l];
    const int per = (n + kFinishCtas - 1) / kFinishCtas;
    const int beg = p.f.level_begin[l] + x * per;
    const int cnt = max(0, min(per, p.f.level_end[l] - beg));
    double cs = strided_sum(p.f.conf0 + beg, cnt);
    cs = block_sum(cs, scratch);
    if (threadIdx.x == 0) p.slices[l * kFinishCtas + x] = cs;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(p.counter, 1) == (int)(gridDim.x * gridDim.y) - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int TA = p.f.match_blocks;
  for (int wc = warp; wc < p.f.g.L * 4; wc += 8) {  // one warp per (level, component)
    const int ll = wc >> 2, c = wc & 3;
    double s0 = 0.0, s1 = 0.0;
    if (TA > 0) {
      const double* ws = p.f.match_ws + (size_t)ll * TA * 4 + c;
      int i = lane;
      for (; i + 32 < TA; i += 64) {
        s0 += ws[(size_t)i * 4];
        s1 += ws[(size_t)(i + 32) * 4];
      }
      if (i < TA) s0 += ws[(size_t)i * 4];
    }
    double tot = warp_sum(s0 + s1);
    if (c == 2) {
      double d = 0.0;
      for (int i = lane; i < kFinishCtas; i += 32) d += __ldcg(&p.slices[ll * kFinishCtas + i]);  // written by other CTAs: L2
      tot += warp_sum(d);
    }
    if (lane == 0) p.f.partials[wc] = tot;
  }
  __threadfence_block();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p.f.out_loss != nullptr)
      p.f.out_loss[0] = combine_loss(p.f.g, p.f.partials, p.f.batch_global, p.f.r_box, p.f.r_conf, p.f.r_cls);
    p.counter[0] = 0;
  }
}

struct CombineParams {
  Geom g;
  const double* partials;
  long long batch_global;
  float r_box, r_conf, r_cls;
  float* out;
};
__global__ void loss_combine_kernel(const CombineParams p) {
  if (threadIdx.x == 0) p.out[0] = combine_loss(p.g, p.partials, p.batch_global, p.r_box, p.r_conf, p.r_cls);
}

// ---- build_target as an API of its own (padded, optionally compacted in (t,a) order) ---------------------
struct BuildTargetParams {
  Geom g;
  int level;
  const float* labels;
  int T;
  int compact;
  long long *b, *gxy, *a, *cls;
  float *xywh, *anchor;
  unsigned char* match;
  int* count;
};

__global__ void __launch_bounds__(1024) build_target_kernel(const BuildTargetParams p) {
  __shared__ uint32_t warp_tot[33];
  __shared__ uint32_t carry;
  const int TA = p.T * p.g.A;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < TA; base += blockDim.x) {
    int i = base + threadIdx.x;
    TargetCell c;
    c.match = false;
    int t = 0, a = 0;
    if (i < TA) {
      t = i / p.g.A;
      a = i - t * p.g.A;
      c = target_cell(p.g, p.level, p.labels + (size_t)t * 6, a);
    }
    // block exclusive scan of the match flags (1024 threads = 32 warps)
    uint32_t v = c.match ? 1u : 0u, inc = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t tv = warp_tot[lane], ti = tv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - tv;
      if (lane == 31) warp_tot[32] = ti;
    }
    __syncthreads();
    uint32_t pos = carry + warp_tot[warp] + inc - v;
    if (i < TA) {
      p.match[i] = c.match ? 1 : 0;
      int o = p.compact ? (c.match ? (int)pos : -1) : i;
      if (o >= 0) {
        p.b[o] = c.b;
        p.gxy[(size_t)o * 2 + 0] = c.gx;
        p.gxy[(size_t)o * 2 + 1] = c.gy;
        p.a[o] = a;
        p.cls[o] = c.cls;
        p.xywh[(size_t)o * 4 + 0] = c.offx;
        p.xywh[(size_t)o * 4 + 1] = c.offy;
        p.xywh[(size_t)o * 4 + 2] = c.tw;
        p.xywh[(size_t)o * 4 + 3] = c.th;
        p.anchor[(size_t)o * 2 + 0] = c.aw;
        p.anchor[(size_t)o * 2 + 1] = c.ah;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) p.count[0] = (int)carry;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int stream_chunks(const Geom& g, int l) { return (g.A * g.HW[l] + kStreamRows - 1) / kStreamRows; }

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_yolov3_loss_workspace_bytes(const fvb_yolo_geom* geom, int64_t num_labels) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return 0;
  size_t o = 256;  // flags
  o = align_up(o + (size_t)g.L * (size_t)num_labels * g.A * 4 * 8, 256);
  size_t parts = 0;
  for (int l = 0; l < g.L; ++l) parts += (size_t)stream_chunks(g, l) * g.B;
  o = align_up(o + parts * 8, 256);
  o = align_up(o + (size_t)g.L * kFinishCtas * 8, 256);  // slice sums of the two-part form
  return o + 256;
}

extern "C" int fvb_yolov3_loss_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                   int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                                   const double* d_conf_bce0, double* d_partials, float* d_out_loss, void* d_ws,
                                   void* stream) {
  return fvb_yolov3_loss_train_f32(geom, d_heads, d_labels, num_labels, ratio_box, ratio_conf, ratio_cls, d_conf_bce0,
                                   d_partials, d_out_loss, nullptr, d_ws, stream);
}

extern "C" int64_t fvb_yolov3_saved_conf_floats(const fvb_yolo_geom* geom) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return 0;
  return (int64_t)g.B * g.row_off[g.L];
}

enum { kLossAll = 0, kLossMatchOnly = 1, kLossFinishOnly = 2 };

static int run_yolov3_loss(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                           int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                           const double* d_conf_bce0, double* d_partials, float* d_out_loss,
                           float* d_saved_conf, void* d_ws, void* stream, int mode, int dense_precise = 0) {
  FVB_REQUIRE(d_ws != nullptr, "yolov3_loss: NULL workspace");
  FVB_REQUIRE(mode == kLossMatchOnly || d_partials != nullptr, "yolov3_loss: NULL partials");
  FVB_REQUIRE(mode == kLossFinishOnly || d_heads != nullptr, "yolov3_loss: NULL heads");
  FVB_REQUIRE(mode != kLossFinishOnly || d_conf_bce0 != nullptr, "yolov3_loss_finish: needs the decode's objectness partials");
  FVB_REQUIRE(num_labels >= 0 && num_labels < (1ll << 24), "yolov3_loss: num_labels=%lld", (long long)num_labels);
  FVB_REQUIRE(num_labels == 0 || d_labels || mode == kLossFinishOnly, "yolov3_loss: labels NULL");
  FVB_REQUIRE(((uintptr_t)d_ws & 255) == 0, "yolov3_loss: workspace must be 256-byte aligned");
  LossParams lp;
  int rc = make_geom(geom, mode == kLossFinishOnly ? nullptr : d_heads, &lp.g);
  if (rc != FVB_OK) return rc;
  if (mode != kLossFinishOnly)
    for (int l = 0; l < lp.g.L; ++l) FVB_REQUIRE(d_heads[l] != nullptr, "yolov3_loss: head %d is NULL", l);
  FVB_REQUIRE(lp.g.B >= 1, "yolov3_loss: empty batch");
  FVB_REQUIRE(!lp.g.nchw, "yolov3_loss: heads must be [B,A,H,W,K] (FVB_HEAD_BAHWK)");
  const Geom& g = lp.g;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)d_ws;
  lp.flags = (int*)w;
  size_t o = 256;
  lp.match_ws = (double*)(w + o);
  o = align_up(o + (size_t)g.L * (size_t)num_labels * g.A * 4 * 8, 256);
  double* stream_parts = (double*)(w + o);
  size_t parts = 0;
  for (int l = 0; l < g.L; ++l) parts += (size_t)stream_chunks(g, l) * g.B;
  o = align_up(o + parts * 8, 256);
  double* finish_slices = (double*)(w + o);
  lp.labels = d_labels;
  lp.T = (int)num_labels;
  // only partials that came from fvb_yolo_decode_f32(precise=1) are "precise"; conf_stream_kernel uses the MUFU sigmoid
  lp.dense_precise = (dense_precise != 0 && (mode == kLossMatchOnly || (d_conf_bce0 != nullptr && d_saved_conf == nullptr))) ? 1 : 0;

  FinalizeParams fp;
  fp.g = g;
  fp.partials = d_partials;
  fp.out_loss = d_out_loss;
  fp.r_box = ratio_box;
  fp.r_conf = ratio_conf;
  fp.r_cls = ratio_cls;
  fp.batch_global = g.B;
  if (mode == kLossMatchOnly) {
    fp.conf0 = nullptr;  // loss_match folds nothing: the objectness partials do not exist yet
    for (int l = 0; l < FVB_MAX_LEVELS; ++l) fp.level_begin[l] = fp.level_end[l] = 0;
  } else if (d_conf_bce0 && d_saved_conf == nullptr) {
    fp.conf0 = d_conf_bce0;
    const int tr = decode_tile_rows(g.K);  // the decode kernel wrote one partial per tile, level-major
    int t = 0;
    for (int l = 0; l < g.L; ++l) {
      fp.level_begin[l] = t;
      t += ((g.A * g.HW[l] + tr - 1) / tr) * g.B;
      fp.level_end[l] = t;
    }
  } else {
    StreamParams sp;
    sp.g = g;
    int t = 0;
    for (int l = 0; l < g.L; ++l) {
      sp.chunk_off[l] = t;
      sp.chunks[l] = stream_chunks(g, l);
      sp.save_off[l] = (long long)g.B * g.row_off[l];
      fp.level_begin[l] = t;
      t += sp.chunks[l] * g.B;
      fp.level_end[l] = t;
    }
    for (int l = g.L; l <= FVB_MAX_LEVELS; ++l) sp.chunk_off[l] = t;
    sp.partials = stream_parts;
    sp.save = d_saved_conf;
    conf_stream_kernel<<<t, 256, 0, s>>>(sp);
    count_launch();
    fp.conf0 = stream_parts;
  }

  const int wpb = kLossThreads / 32;
  const int match_blocks = (int)(((long long)lp.T * g.A + wpb - 1) / wpb);
  fp.match_blocks = match_blocks;
  fp.match_ws = lp.match_ws;
  if (mode == kLossFinishOnly) {
    FinishParams q;
    q.f = fp;
    q.slices = finish_slices;
    q.counter = lp.flags + 1;
    loss_finish_kernel<<<dim3(kFinishCtas, (unsigned)g.L), 256, 0, s>>>(q);
    count_launch();
    return check_launch("yolov3_loss_finish");
  }
  if (lp.T > 0 || mode == kLossMatchOnly) {
    loss_prep_kernel<<<1, 1024, 0, s>>>(d_labels, lp.T, lp.flags);  // (also zeroes the finish kernel's arrival counter)
    count_launch();
  }
  if (lp.T > 0) {
    lp.conf0 = fp.conf0;
    for (int l = 0; l < FVB_MAX_LEVELS; ++l) {
      lp.conf_begin[l] = l < g.L ? fp.level_begin[l] : 0;
      lp.conf_end[l] = l < g.L ? fp.level_end[l] : 0;
    }
    dim3 grid((unsigned)match_blocks, (unsigned)g.L);
    loss_match_kernel<<<grid, kLossThreads, 0, s>>>(lp);
    count_launch();
  }
  if (mode == kLossMatchOnly) return check_launch("yolov3_loss_match");
  loss_finalize_kernel<<<1, 1024, 0, s>>>(fp);
  count_launch();
  return check_launch("yolov3_loss");
}

extern "C" int fvb_yolov3_loss_train_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                         int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                                         const double* d_conf_bce0, double* d_partials, float* d_out_loss,
                                         float* d_saved_conf, void* d_ws, void* stream) {
  return run_yolov3_loss(geom, d_heads, d_labels, num_labels, ratio_box, ratio_conf, ratio_cls, d_conf_bce0, d_partials,
                         d_out_loss, d_saved_conf, d_ws, stream, kLossAll);
}

// The same loss in two parts.  Target assignment and the matched-row terms read only the RAW heads and the labels, so
// fvb_yolov3_loss_match_f32 can be enqueued beside the decode kernel and hide under it; what needs the decode is only the sum of
// its objectness partials: fvb_yolov3_loss_finish_f32 (same d_ws, same num_labels, after both) does that sum, reduces the
// matched terms and writes {S_cls, S_box, S_conf, M} per level and the scalar.  match + finish == fvb_yolov3_loss_f32 up to
// the association of the fp64 sums (both orders are fixed, so each form is bit-reproducible).
extern "C" int fvb_yolov3_loss_match_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                         int64_t num_labels, void* d_ws, void* stream) {
  return run_yolov3_loss(geom, d_heads, d_labels, num_labels, 0.f, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, d_ws, stream,
                         kLossMatchOnly);
}

/* v2: the same two entry points for objectness partials produced by fvb_yolo_decode_f32(..., precise = 1) */
extern "C" int fvb_yolov3_loss_dense_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                         int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                                         const double* d_conf_bce0, int conf_bce0_precise, double* d_partials,
                                         float* d_out_loss, void* d_ws, void* stream) {
  return run_yolov3_loss(geom, d_heads, d_labels, num_labels, ratio_box, ratio_conf, ratio_cls, d_conf_bce0, d_partials,
                         d_out_loss, nullptr, d_ws, stream, kLossAll, conf_bce0_precise);
}

extern "C" int fvb_yolov3_loss_match_dense_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                               int64_t num_labels, int conf_bce0_precise, void* d_ws, void* stream) {
  return run_yolov3_loss(geom, d_heads, d_labels, num_labels, 0.f, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, d_ws, stream,
                         kLossMatchOnly, conf_bce0_precise);
}

extern "C" int fvb_yolov3_loss_finish_f32(const fvb_yolo_geom* geom, int64_t num_labels, float ratio_box, float ratio_conf,
                                          float ratio_cls, const double* d_conf_bce0, double* d_partials, float* d_out_loss,
                                          void* d_ws, void* stream) {
  return run_yolov3_loss(geom, nullptr, nullptr, num_labels, ratio_box, ratio_conf, ratio_cls, d_conf_bce0, d_partials,
                         d_out_loss, nullptr, d_ws, stream, kLossFinishOnly);
}

extern "C" int fvb_yolov3_loss_combine_f32(const fvb_yolo_geom* geom, int64_t batch_global, const double* d_partials,
                                           float ratio_box, float ratio_conf, float ratio_cls, float* d_out_loss,
                                           void* stream) {
  FVB_REQUIRE(d_partials && d_out_loss, "loss_combine: NULL pointer");
  FVB_REQUIRE(batch_global >= 1, "loss_combine: batch_global=%lld", (long long)batch_global);
  CombineParams p;
  int rc = make_geom(geom, nullptr, &p.g);
  if (rc != FVB_OK) return rc;
  p.partials = d_partials;
  p.batch_global = batch_global;
  p.r_box = ratio_box;
  p.r_conf = ratio_conf;
  p.r_cls = ratio_cls;
  p.out = d_out_loss;
  loss_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  return check_launch("loss_combine_kernel");
}

extern "C" int fvb_yolov3_build_target_f32(const fvb_yolo_geom* geom, int level, const float* d_labels,
                                           int64_t num_labels, int compact, int64_t* d_b, int64_t* d_gxy, int64_t* d_a,
                                           int64_t* d_cls, float* d_xywh, float* d_anchor, uint8_t* d_match,
                                           int32_t* d_count, void* stream) {
  BuildTargetParams p;
  int rc = make_geom(geom, nullptr, &p.g);
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(level >= 0 && level < p.g.L, "build_target: level %d", level);
  FVB_REQUIRE(num_labels >= 0 && num_labels < (1ll << 24), "build_target: num_labels");
  FVB_REQUIRE(d_count && (num_labels == 0 || (d_labels && d_b && d_gxy && d_a && d_cls && d_xywh && d_anchor && d_match)),
              "build_target: NULL pointer");
  p.level = level;
  p.labels = d_labels;
  p.T = (int)num_labels;
  p.compact = compact;
  p.b = (long long*)d_b;
  p.gxy = (long long*)d_gxy;
  p.a = (long long*)d_a;
  p.cls = (long long*)d_cls;
  p.xywh = d_xywh;
  p.anchor = d_anchor;
  p.match = d_match;
  p.count = d_count;
  build_target_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  return check_launch("build_target_kernel");
}
