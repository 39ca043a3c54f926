// K1: YOLOv3 head decode, one pass over the raw head tensors (HBM-bound).
//
// Replaces detection/models/yolov3.py:33-53 of the reference (~30 ATen launches, ~6 full passes).
// Layout fact the kernel is built on: inside one (image, level) segment the raw head [A,H,W,K] and
// the decoded rows [a*H*W + y*W + x, K] are the SAME flat order, so decode is a contiguous ->
// contiguous map.  Work unit = one warp x 32 consecutive rows of one segment:
//   * lanes <-> channels (c = lane + 32 j), so a warp instruction reads/writes 128 contiguous bytes,
//     the channel of every register is known statically (no per-element index arithmetic, the
//     xy / wh / objectness special cases cost a few selects on iteration j = 0 only) and the cell
//     coordinates (x, y, anchor) are per-row values advanced incrementally;
//   * rows narrower than 17 floats are packed 2 or 4 per warp iteration so lanes stay busy;
//   * 4 rows (up to 12 independent 128-byte loads per warp) are in flight before the first use.
// Fused side outputs fall out of channel 4 while it is in registers: every lane keeps the objectness
// of "its" row of the group, then ONE ballot gives 32 bits of the NMS candidate bitmap and ONE
// 32-lane pass computes the zero-target objectness BCE of Yolov3Loss for the whole group.
#include "common.cuh"

namespace fvb {

struct DecodeParams {
  Geom g;
  int blocks_level_end[FVB_MAX_LEVELS];  // cumulative CTAs (8 groups of 32 rows each) per image
  int blocks_per_image;
  float* out;
  float conf_thr;
  uint32_t* bitmap;
  int bitmap_words;
  float* cand_rec;   // [B][N][8] = {row[0..3], conf, max_c(cls*conf), argmax as int bits, -}, written for candidates only
  double* bce0;      // [blocks_per_image][B]: one zero-target objectness BCE partial per CTA
};

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

// One 32-row group.  J = ceil(K/32) register columns per row (RPI == 1), or J = 1 with RPI rows packed
// per warp iteration.  FULL: all 32 rows exist (every group but the last of a segment).
// side: per-warp shared scratch [2][32] (raw objectness logit and decoded objectness of each row).
template <int J, int RPI, int FORM, bool PRECISE, bool FULL>
__device__ __forceinline__ void decode_group(const DecodeParams& p, const int l, const int b, const int grp,
                                             float* side, double* block_acc) {
  constexpr int SLOT = 32 / RPI;   // lanes per row
  constexpr int ITERS = 32 / RPI;  // warp iterations per 32-row group
  constexpr int UN = 4;            // iterations in flight
  const int lane = threadIdx.x & 31;
  const int K = p.g.K, W = p.g.W[l], H = p.g.H[l], HW = p.g.HW[l];
  const int rows_l = p.g.A * HW;
  const int row0 = grp * 32;
  const int nrows = FULL ? 32 : rows_l - row0;
  const int sub = lane / SLOT;       // row inside one iteration
  const int c0 = lane - sub * SLOT;  // channel of register column 0
  const int lane_off = (RPI == 1) ? lane : sub * K + c0;
  const int K1 = RPI * K;            // floats per warp iteration
  const size_t out_row0 = (size_t)b * p.g.row_off[p.g.L] + p.g.row_off[l] + row0;
  const float* __restrict__ rp = p.g.head[l] + ((size_t)b * rows_l + row0) * K + lane_off;
  float* __restrict__ wp = p.out + out_row0 * K + lane_off;

  // RPI == 1: columns j < J-1 are always inside the row (J = ceil(K/32)); only the last one is ragged
  const bool last_live = (RPI == 1) ? (lane + 32 * (J - 1) < K) : (c0 < K);

  // cell coordinates of this lane's row at iteration 0, then advanced incrementally
  int a, y, x;
  {
    int rl = row0 + sub;
    a = min(rl / HW, p.g.A - 1);
    int yx = rl - a * HW;
    y = yx / W;
    x = yx - y * W;
  }
  const float stride = p.g.stride[l];
  const bool is_wh = (c0 == 2) | (c0 == 3);
  const bool is_xy = c0 < 2;
  const float scale0 = (FORM == FVB_DECODE_V3 && is_wh) ? kLog2e : -kLog2e;
  float anc = (c0 == 2) ? p.g.aw[l][a] : p.g.ah[l][a];
  const bool fused = (p.bitmap != nullptr) | (p.bce0 != nullptr);

#pragma unroll 1
  for (int it0 = 0; it0 < ITERS; it0 += UN) {
    float v[UN][J];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const bool row_ok = FULL || ((it0 + u) * RPI + sub < nrows);
      const float* r = rp + u * K1;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const bool ok = row_ok && (j < J - 1 || last_live);
        v[u][j] = ok ? r[32 * j] : 0.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int it = it0 + u;
      const bool row_ok = FULL || (it * RPI + sub < nrows);
      float* w = wp + u * K1;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float t = v[u][j];
        float o;
        if (j == 0) {
          float ex, sg;
          if (PRECISE) {
            ex = expf((FORM == FVB_DECODE_V3 && is_wh) ? t : -t);
            sg = 1.0f / (1.0f + ex);
          } else {
            ex = ex2_approx(t * scale0);
            sg = rcp_approx(1.0f + ex);
          }
          const float gxy = (float)((c0 == 0) ? x : y);
          float oxy, owh;
          if (FORM == FVB_DECODE_V3) {
            oxy = (sg + gxy) * stride;  // yolov3.py:47
            owh = ex * anc;             // yolov3.py:48
          } else {                      // demos/yolov3_u/inference.py:86-89
            const float s2 = sg * 2.0f;
            oxy = ((s2 - 0.5f) + gxy) * stride;
            owh = (s2 * s2) * anc;
          }
          o = is_xy ? oxy : (is_wh ? owh : sg);
          if (fused && c0 == 4 && row_ok) {  // objectness of row it*RPI+sub: raw logit and decoded value
            side[it * RPI + sub] = t;
            side[32 + it * RPI + sub] = o;
          }
        } else {
          if (PRECISE) o = 1.0f / (1.0f + expf(-t));
          else o = rcp_approx(1.0f + ex2_approx(t * -kLog2e));
        }
        if (row_ok && (j < J - 1 || last_live)) w[32 * j] = o;
      }
      // advance this lane's row by RPI
      x += RPI;
      if (x >= W) {
        do {
          x -= W;
          y += 1;
        } while (x >= W);
        if (y >= H) {
          y -= H;
          a = min(a + 1, p.g.A - 1);
          anc = (c0 == 2) ? p.g.aw[l][a] : p.g.ah[l][a];
        }
      }
    }
    rp += UN * K1;
    wp += UN * K1;
  }

  if (!fused) return;
  __syncwarp();  // side[] and this warp's decoded rows are now visible to all of its lanes
  const bool valid = lane < nrows;
  const float my_t4 = valid ? side[lane] : 0.0f;
  const float my_conf = valid ? side[32 + lane] : 0.0f;
  if (p.bitmap != nullptr) {
    unsigned m = __ballot_sync(0xffffffffu, valid && my_conf > p.conf_thr);  // NMS.py:7 on the stored value
    const unsigned gr = (unsigned)(p.g.row_off[l] + row0);
    const unsigned sh = gr & 31u;
    uint32_t* wptr = p.bitmap + (size_t)b * p.bitmap_words + (gr >> 5);
    if (lane == 0) {
      const unsigned lo = m << sh;
      if (lo) atomicOr(wptr, lo);
    } else if (lane == 1 && sh) {
      const unsigned hi = m >> (32u - sh);
      if (hi) atomicOr(wptr + 1, hi);
    }
    // candidate records (~7% of rows): re-read the decoded row (L1/L2 hit), score = max_c(cls_c*conf) on the
    // STORED fp32 values (NMS.py:13,16: first maximum on ties)
    if (p.cand_rec != nullptr) {
      while (m) {
        const int rr = __ffs(m) - 1;
        m &= m - 1;
        const float* row = p.out + (out_row0 + rr) * K;
        constexpr int JR = (RPI == 1) ? J : 1;  // K <= 32 when rows are packed
        float o[JR];
#pragma unroll
        for (int j = 0; j < JR; ++j) o[j] = (lane + 32 * j < K) ? row[lane + 32 * j] : 0.0f;
        const float conf = __shfl_sync(0xffffffffu, o[0], 4);
        // products of two sigmoids are >= +0, so their bit patterns order like unsigned integers
        unsigned best = 0u;
        int bidx = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < JR; ++j) {
          const int ch = lane + 32 * j;
          if (ch >= 5 && ch < K) {
            const unsigned pr = __float_as_uint(o[j] * conf);
            if (bidx == 0x7fffffff || pr > best) {  // strict >: the first maximum wins inside the lane
              best = pr;
              bidx = ch - 5;
            }
          }
        }
        const unsigned wbest = __reduce_max_sync(0xffffffffu, best);
        const int widx = __reduce_min_sync(0xffffffffu, (best == wbest) ? bidx : 0x7fffffff);
        if (lane < 7) {
          const float val = lane < 5 ? o[0] : (lane == 5 ? __uint_as_float(wbest) : __int_as_float(widx));
          p.cand_rec[(out_row0 + rr) * 8 + lane] = val;
        }
      }
    }
  }
  if (p.bce0 != nullptr) {
    const float term = valid ? bce_term(sigmoid_precise(my_t4), 0.0f) : 0.0f;
    const double s = warp_sum((double)term);
    if (lane == 0) *block_acc = s;
  }
}

// grid.x enumerates 8-group blocks level by level (a block never straddles two levels), grid.y = image.
template <int J, int RPI, int FORM, bool PRECISE>
__global__ void __launch_bounds__(kDecodeThreads) decode_kernel(const DecodeParams p) {
  constexpr int WPB = kDecodeThreads / 32;
  __shared__ float side[WPB][64];
  __shared__ double acc[WPB];
  __shared__ unsigned arrived;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5;
  int blk = blockIdx.x, l = 0;
#pragma unroll
  for (int i = 0; i < FVB_MAX_LEVELS - 1; ++i)
    if (i < p.g.L - 1 && blk >= p.blocks_level_end[i]) l = i + 1;
  if (l > 0) blk -= p.blocks_level_end[l - 1];
  const int grp = blk * WPB + warp;
  const int rows_l = p.g.A * p.g.HW[l];
  if ((threadIdx.x & 31) == 0) acc[warp] = 0.0;
  if (p.bce0 != nullptr) {
    if (threadIdx.x == 0) arrived = 0u;
    __syncthreads();  // before any work: cheap, nobody waits on a slow warp here
  }
  if (grp * 32 < rows_l) {
    if (rows_l - grp * 32 >= 32) decode_group<J, RPI, FORM, PRECISE, true>(p, l, b, grp, side[warp], &acc[warp]);
    else decode_group<J, RPI, FORM, PRECISE, false>(p, l, b, grp, side[warp], &acc[warp]);
  }
  if (p.bce0 != nullptr && (threadIdx.x & 31) == 0) {
    // one partial per block, summed in a fixed order by whichever warp finishes last (no barrier: warps
    // that are done must not hold back the CTA's slots while a slow warp still streams)
    __threadfence_block();
    if (atomicAdd(&arrived, 1u) == WPB - 1) {
      __threadfence_block();
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < WPB; ++i) s += ((volatile double*)acc)[i];
      p.bce0[(size_t)blockIdx.x * p.g.B + b] = s;  // [blocks_per_image][B]: a level's partials are contiguous
    }
  }
}

int make_geom(const fvb_yolo_geom* g, const float* const* d_heads, Geom* out) {
  FVB_REQUIRE(g != nullptr, "geom is NULL");
  FVB_REQUIRE(g->levels >= 1 && g->levels <= FVB_MAX_LEVELS, "levels=%d out of range [1,%d]", g->levels, FVB_MAX_LEVELS);
  FVB_REQUIRE(g->anchors >= 1 && g->anchors <= FVB_MAX_ANCHORS, "anchors=%d out of range [1,%d]", g->anchors, FVB_MAX_ANCHORS);
  FVB_REQUIRE(g->channels >= 6, "channels=%d: need 5 + at least one class", g->channels);
  FVB_REQUIRE(g->batch >= 0, "batch=%d", g->batch);
  out->L = g->levels;
  out->B = g->batch;
  out->A = g->anchors;
  out->K = g->channels;
  long long rows = 0;
  for (int l = 0; l < g->levels; ++l) {
    FVB_REQUIRE(g->height[l] >= 1 && g->width[l] >= 1, "level %d has empty feature map", l);
    out->H[l] = g->height[l];
    out->W[l] = g->width[l];
    out->HW[l] = g->height[l] * g->width[l];
    out->stride[l] = g->stride[l];
    out->row_off[l] = (int)rows;
    rows += (long long)g->anchors * out->HW[l];
    long long seg = (long long)g->anchors * out->HW[l] * g->channels;
    if (seg >= (1ll << 31)) {
      set_error("level %d too large for 32-bit segment indexing", l);
      return FVB_E_LIMIT;
    }
    for (int a = 0; a < g->anchors; ++a) {
      out->aw[l][a] = g->anchor_w[l][a];
      out->ah[l][a] = g->anchor_h[l][a];
    }
    out->head[l] = d_heads ? d_heads[l] : nullptr;
  }
  if (rows >= (1ll << 24)) {
    set_error("rows per image %lld >= 2^24", rows);
    return FVB_E_LIMIT;
  }
  out->row_off[g->levels] = (int)rows;
  return FVB_OK;
}

template <int FORM, bool PRECISE>
static void launch_decode(const DecodeParams& p, dim3 grid, cudaStream_t s) {
  const int K = p.g.K;
  if (K <= 8) decode_kernel<1, 4, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
  else if (K <= 16) decode_kernel<1, 2, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
  else if (K <= 32) decode_kernel<1, 1, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
  else if (K <= 64) decode_kernel<2, 1, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
  else if (K <= 96) decode_kernel<3, 1, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
  else if (K <= 128) decode_kernel<4, 1, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
  else if (K <= 192) decode_kernel<6, 1, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
  else decode_kernel<8, 1, FORM, PRECISE><<<grid, kDecodeThreads, 0, s>>>(p);
}

}  // namespace fvb

using namespace fvb;

extern "C" int fvb_yolo_rows_per_image(const fvb_yolo_geom* geom) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return -1;
  return g.row_off[g.L];
}

extern "C" int fvb_yolo_bitmap_words(const fvb_yolo_geom* geom) {
  int n = fvb_yolo_rows_per_image(geom);
  return n < 0 ? -1 : (n + 31) / 32;
}

extern "C" int fvb_yolo_decode_tiles(const fvb_yolo_geom* geom) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return -1;
  int t = 0;
  for (int l = 0; l < g.L; ++l) t += decode_blocks_level(g, l);
  return t;
}

extern "C" int fvb_yolo_decode_f32(const fvb_yolo_geom* geom, const float* const* d_heads, int form, int precise,
                                   float* d_results, float conf_thr, uint32_t* d_cand_bitmap, float* d_cand_rec,
                                   double* d_conf_bce0, void* stream) {
  DecodeParams p;
  FVB_REQUIRE(d_heads != nullptr && d_results != nullptr, "decode: NULL head/result pointer");
  int rc = make_geom(geom, d_heads, &p.g);
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(form == FVB_DECODE_V3 || form == FVB_DECODE_V5, "decode: unknown form %d", form);
  FVB_REQUIRE(p.g.B <= 65535, "decode: batch %d > 65535 (grid.y)", p.g.B);
  FVB_REQUIRE(p.g.K <= 256, "decode: channels %d > 256", p.g.K);
  for (int l = 0; l < p.g.L; ++l) FVB_REQUIRE(d_heads[l] != nullptr, "decode: head %d is NULL", l);
  if (p.g.B == 0) return FVB_OK;
  int t = 0;
  for (int l = 0; l < p.g.L; ++l) {
    t += decode_blocks_level(p.g, l);
    p.blocks_level_end[l] = t;
  }
  for (int l = p.g.L; l < FVB_MAX_LEVELS; ++l) p.blocks_level_end[l] = t;
  p.blocks_per_image = t;
  p.out = d_results;
  p.conf_thr = conf_thr;
  p.bitmap = d_cand_bitmap;
  p.bitmap_words = (p.g.row_off[p.g.L] + 31) / 32;
  p.bce0 = d_conf_bce0;
  p.cand_rec = d_cand_rec;
  FVB_REQUIRE(d_cand_rec == nullptr || d_cand_bitmap != nullptr, "decode: candidate records need the candidate bitmap too");
  dim3 grid((unsigned)t, (unsigned)p.g.B);
  cudaStream_t s = (cudaStream_t)stream;
  if (form == FVB_DECODE_V3) {
    if (precise) launch_decode<FVB_DECODE_V3, true>(p, grid, s);
    else launch_decode<FVB_DECODE_V3, false>(p, grid, s);
  } else {
    if (precise) launch_decode<FVB_DECODE_V5, true>(p, grid, s);
    else launch_decode<FVB_DECODE_V5, false>(p, grid, s);
  }
  count_launch();
  return check_launch("decode_kernel");
}
