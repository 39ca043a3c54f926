// K1: YOLOv3 head decode, one pass over the raw head tensors (HBM-bound).
//
// Replaces detection/models/yolov3.py:33-53 of the reference (~30 ATen launches, ~6 full passes).
// Layout fact the kernel is built on: inside one (image, level) segment the raw head
// [A,H,W,K] and the decoded rows [a*H*W + y*W + x, K] are the SAME flat order, so decode is a
// contiguous -> contiguous element-wise map with channel = e mod K and cell = e div K.
// Each CTA owns one 4096-float tile of one segment; lanes touch consecutive floats (fully
// coalesced 128 B per warp instruction on both sides), 16 independent loads in flight per thread.
// Channel / cell are tracked incrementally (no per-element division); the grid coordinates
// (x, y, anchor) are only derived for channels 0..3 via multiply-shift division.
// Fused side outputs: the NMS candidate bitmap (conf > thr) and the zero-target objectness BCE
// partial sums for Yolov3Loss -- both fall out of channel 4 while it is in registers.
#include "common.cuh"

namespace fvb {

struct DecodeParams {
  Geom g;
  int tiles_level_end[FVB_MAX_LEVELS];  // cumulative tile count per image
  int tiles_per_image;
  unsigned long long magic_hw[FVB_MAX_LEVELS];  // floor(2^40 / HW) + 1
  unsigned long long magic_w[FVB_MAX_LEVELS];   // floor(2^40 / W) + 1
  int dc, dr;                                   // 256 mod K, 256 div K
  float* out;
  float conf_thr;
  uint32_t* bitmap;
  int bitmap_words;
  double* bce0;
};

__device__ __forceinline__ uint32_t div_magic(uint32_t n, unsigned long long m) {
  return (uint32_t)(((unsigned long long)n * m) >> 40);
}

template <int FORM, bool PRECISE>
__global__ void __launch_bounds__(kDecodeThreads) decode_kernel(const DecodeParams p) {
  constexpr int U = kDecodeTile / kDecodeThreads;  // 16
  const int b = blockIdx.y;
  int tile = blockIdx.x;
  int l = 0;
#pragma unroll
  for (int i = 0; i < FVB_MAX_LEVELS - 1; ++i)
    if (i < p.g.L - 1 && tile >= p.tiles_level_end[i]) l = i + 1;
  if (l > 0) tile -= p.tiles_level_end[l - 1];

  const int K = p.g.K;
  const uint32_t seg = (uint32_t)p.g.A * p.g.HW[l] * K;  // floats in this (image, level) segment
  const float* __restrict__ in = p.g.head[l] + (size_t)b * seg;
  float* __restrict__ out = p.out + ((size_t)b * p.g.row_off[p.g.L] + p.g.row_off[l]) * K;

  const uint32_t e0 = (uint32_t)tile * kDecodeTile + threadIdx.x;
  float v[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    uint32_t e = e0 + i * kDecodeThreads;
    v[i] = e < seg ? __ldcs(in + e) : 0.0f;
  }

  uint32_t r = e0 / (uint32_t)K;  // cell (row inside the level)
  int c = (int)(e0 - r * K);      // channel
  const float stride = p.g.stride[l];
  float acc = 0.0f;

#pragma unroll
  for (int i = 0; i < U; ++i) {
    uint32_t e = e0 + i * kDecodeThreads;
    float t = v[i];
    bool is_wh = (c == 2) | (c == 3);
    float o;
    if (FORM == FVB_DECODE_V3) {
      float ex, sg;
      if (PRECISE) {
        ex = expf(is_wh ? t : -t);
        sg = 1.0f / (1.0f + ex);
      } else {
        ex = __expf(is_wh ? t : -t);
        sg = __fdividef(1.0f, 1.0f + ex);
      }
      o = sg;
      if (c < 4) {
        uint32_t a = div_magic(r, p.magic_hw[l]);
        uint32_t yx = r - a * p.g.HW[l];
        uint32_t y = div_magic(yx, p.magic_w[l]);
        uint32_t x = yx - y * p.g.W[l];
        if (c == 0) o = (sg + (float)x) * stride;
        else if (c == 1) o = (sg + (float)y) * stride;
        else if (c == 2) o = ex * p.g.aw[l][a];
        else o = ex * p.g.ah[l][a];
      }
    } else {  // FVB_DECODE_V5: demos/yolov3_u/inference.py:86-89
      float ex = PRECISE ? expf(-t) : __expf(-t);
      float sg = PRECISE ? 1.0f / (1.0f + ex) : __fdividef(1.0f, 1.0f + ex);
      o = sg;
      if (c < 4) {
        uint32_t a = div_magic(r, p.magic_hw[l]);
        uint32_t yx = r - a * p.g.HW[l];
        uint32_t y = div_magic(yx, p.magic_w[l]);
        uint32_t x = yx - y * p.g.W[l];
        float s2 = sg * 2.0f;
        if (c == 0) o = ((s2 - 0.5f) + (float)x) * stride;
        else if (c == 1) o = ((s2 - 0.5f) + (float)y) * stride;
        else if (c == 2) o = (s2 * s2) * p.g.aw[l][a];
        else o = (s2 * s2) * p.g.ah[l][a];
      }
    }
    if (c == 4 && e < seg) {
      if (p.bitmap != nullptr && o > p.conf_thr) {
        uint32_t row = (uint32_t)p.g.row_off[l] + r;
        atomicOr(p.bitmap + (size_t)b * p.bitmap_words + (row >> 5), 1u << (row & 31));
      }
      if (p.bce0 != nullptr) acc += bce_term(sigmoid_precise(t), 0.0f);
    }
    if (e < seg) out[e] = o;
    c += p.dc;
    r += p.dr;
    if (c >= K) {
      c -= K;
      r += 1;
    }
  }

  if (p.bce0 != nullptr) {
    __shared__ double scratch[32];
    double s = block_sum((double)acc, scratch);
    if (threadIdx.x == 0) p.bce0[(size_t)b * p.tiles_per_image + blockIdx.x] = s;
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
    if (seg >= (1ll << 31) || (long long)g->anchors * out->HW[l] * (long long)out->HW[l] >= (1ll << 40)) {
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
  for (int l = 0; l < g.L; ++l) t += decode_tiles_level(g, l);
  return t;
}

extern "C" int fvb_yolo_decode_f32(const fvb_yolo_geom* geom, const float* const* d_heads, int form, int precise,
                                   float* d_results, float conf_thr, uint32_t* d_cand_bitmap, double* d_conf_bce0,
                                   void* stream) {
  DecodeParams p;
  FVB_REQUIRE(d_heads != nullptr && d_results != nullptr, "decode: NULL head/result pointer");
  int rc = make_geom(geom, d_heads, &p.g);
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(form == FVB_DECODE_V3 || form == FVB_DECODE_V5, "decode: unknown form %d", form);
  FVB_REQUIRE(p.g.B <= 65535, "decode: batch %d > 65535 (grid.y)", p.g.B);
  for (int l = 0; l < p.g.L; ++l) FVB_REQUIRE(d_heads[l] != nullptr, "decode: head %d is NULL", l);
  if (p.g.B == 0) return FVB_OK;
  int t = 0;
  for (int l = 0; l < p.g.L; ++l) {
    t += decode_tiles_level(p.g, l);
    p.tiles_level_end[l] = t;
    p.magic_hw[l] = (1ull << 40) / (unsigned long long)p.g.HW[l] + 1;
    p.magic_w[l] = (1ull << 40) / (unsigned long long)p.g.W[l] + 1;
  }
  for (int l = p.g.L; l < FVB_MAX_LEVELS; ++l) p.tiles_level_end[l] = t;
  p.tiles_per_image = t;
  p.dc = kDecodeThreads % p.g.K;
  p.dr = kDecodeThreads / p.g.K;
  p.out = d_results;
  p.conf_thr = conf_thr;
  p.bitmap = d_cand_bitmap;
  p.bitmap_words = (p.g.row_off[p.g.L] + 31) / 32;
  p.bce0 = d_conf_bce0;
  dim3 grid((unsigned)t, (unsigned)p.g.B);
  cudaStream_t s = (cudaStream_t)stream;
  if (form == FVB_DECODE_V3) {
    if (precise) decode_kernel<FVB_DECODE_V3, true><<<grid, kDecodeThreads, 0, s>>>(p);
    else decode_kernel<FVB_DECODE_V3, false><<<grid, kDecodeThreads, 0, s>>>(p);
  } else {
    if (precise) decode_kernel<FVB_DECODE_V5, true><<<grid, kDecodeThreads, 0, s>>>(p);
    else decode_kernel<FVB_DECODE_V5, false><<<grid, kDecodeThreads, 0, s>>>(p);
  }
  count_launch();
  return check_launch("decode_kernel");
}
