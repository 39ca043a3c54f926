// K1: YOLOv3 head decode, one pass over the raw head tensors (HBM-bound).
//
// Replaces detection/models/yolov3.py:33-53 of the reference (~30 ATen launches, ~6 full passes).
// Layout fact the kernel is built on: inside one (image, level) segment the raw head [A,H,W,K] and the decoded
// rows [a*H*W + y*W + x, K] are the SAME flat order, so decode is a contiguous -> contiguous map and a TILE
// (T consecutive rows of one segment, ~5 KB) is one contiguous run of floats on both sides.
//
// Persistent kernel, one CTA per SM, every warp an independent pipeline over tiles gw, gw + NW, gw + 2 NW, ...:
//   * the tile after the current one is always in flight: cp.async (16-byte when the source run is 16-byte
//     aligned, else 4-byte) into the warp's second shared-memory buffer, no registers held, no warp waiting;
//   * B0  lane <-> row: the five head channels (x, y, w, h, objectness) of "its" row are read from the tile and
//         decoded (cell coordinates from one pair of integer divisions per lane per tile); the zero-target
//         objectness BCE of Yolov3Loss is accumulated per warp and level in fp64;
//   * A   the whole tile is a flat in-place sigmoid, 128-bit shared-memory accesses, no index arithmetic;
//   * B1  the decoded head channels overwrite their slots; ONE ballot gives the rows' bits of the NMS candidate
//         bitmap; candidate rows (~7 %) get their 32-byte NMS record from the class scores sitting in shared memory;
//   * the finished tile is copied out flat: every store instruction writes 128 contiguous bytes (512 when the
//     destination run is 16-byte aligned).
// The kernel keeps <= 113 KB of shared memory per SM so that one NMS CTA can be co-resident (pipeline.py).
#include "common.cuh"

#include <stdlib.h>

namespace fvb {

constexpr int kDecodeWarps = kDecodeThreads / 32;  // default warps per CTA
constexpr int kDecodeMaxThreads = 768;

struct DecodeParams {
  Geom g;
  int tile_rows;                        // T: rows per tile (16, 32, 64 or 128; wider rows -> fewer rows)
  int tiles_level_end[FVB_MAX_LEVELS];  // cumulative tiles per image, level by level
  int tiles_per_image;
  long long total_tiles;
  int tile_floats;  // floats per shared-memory buffer (T*K rounded up to a multiple of 4)
  int warps_per_cta;
  int stages;  // shared-memory buffers per warp: 1 (load, then process) or 2 (next tile in flight while processing)
  float* out;
  float conf_thr;
  uint32_t* bitmap;
  int bitmap_words;
  float* cand_rec;  // [B][N][8] = {row[0..3], conf, max_c(cls*conf), argmax as int bits, -}, written for candidates only
  double* bce0;     // [L][NW]: zero-target objectness BCE partial of every warp of the grid, per level
};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool PRECISE>
__device__ __forceinline__ float sigmoid_dec(float t) {
  if (PRECISE) return 1.0f / (1.0f + expf(-t));
  return sigmoid_fast(t);
}
template <bool PRECISE>
__device__ __forceinline__ float exp_dec(float t) {
  if (PRECISE) return expf(t);
  return ex2_approx(t * kLog2e);
}

struct Tile {
  const float* src;  // first float of the run in the raw head
  float* dst;        // first float of the run in results
  size_t out_row;    // global row index (b*N + row_off[l] + row0)
  int n;             // floats in the run (nrows*K)
  int nrows, l, b, row0;
};

__device__ __forceinline__ Tile describe_tile(const DecodeParams& p, int u) {
  Tile t;
  const int b = u / p.tiles_per_image;
  int r = u - b * p.tiles_per_image;
  int l = 0;
#pragma unroll
  for (int i = 0; i < FVB_MAX_LEVELS - 1; ++i)
    if (i < p.g.L - 1 && r >= p.tiles_level_end[i]) l = i + 1;
  if (l > 0) r -= p.tiles_level_end[l - 1];
  const int rows_l = p.g.A * p.g.HW[l];
  t.row0 = r * p.tile_rows;
  t.nrows = min(p.tile_rows, rows_l - t.row0);
  t.n = t.nrows * p.g.K;
  t.l = l;
  t.b = b;
  t.out_row = (size_t)b * p.g.row_off[p.g.L] + p.g.row_off[l] + t.row0;
  t.src = p.g.head[l] + ((size_t)b * rows_l + t.row0) * p.g.K;
  t.dst = p.out + t.out_row * p.g.K;
  return t;
}

__device__ __forceinline__ void issue_tile(const Tile& t, float* buf, int lane) {
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(t.src) & 15) == 0) {
    const int n4 = t.n >> 2;
    for (int i = lane; i < n4; i += 32) cp_async16(buf + 4 * i, t.src + 4 * i);
    done = n4 << 2;
  }
  for (int i = done + lane; i < t.n; i += 32) cp_async4(buf + i, t.src + i);
}

// NSB = 32-row sub-blocks per tile (lane <-> row passes)
template <int NSB, int FORM, bool PRECISE>
__device__ __forceinline__ void process_tile(const DecodeParams& p, const Tile& t, float* buf, double (&acc)[FVB_MAX_LEVELS]) {
  const int lane = threadIdx.x & 31;
  const int K = p.g.K, l = t.l;
  const int W = p.g.W[l], HW = p.g.HW[l];
  const float stride = p.g.stride[l];

  // ---- B0: lane <-> row, head channels from the raw tile ----------------------------------------------------------
  float ox[NSB], oy[NSB], ow[NSB], oh[NSB], conf[NSB];
  double bce = 0.0;
#pragma unroll
  for (int sb = 0; sb < NSB; ++sb) {
    const int r = sb * 32 + lane;
    const bool valid = r < t.nrows;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
    if (valid) {
      const float* sr = buf + r * K;
      t0 = sr[0]; t1 = sr[1]; t2 = sr[2]; t3 = sr[3]; t4 = sr[4];
    }
    const int rl = t.row0 + (valid ? r : 0);
    const int a = rl / HW;
    const int yx = rl - a * HW;
    const int y = yx / W;
    const int x = yx - y * W;
    const float aw = p.g.aw[l][a], ah = p.g.ah[l][a];
    const float s0 = sigmoid_dec<PRECISE>(t0), s1 = sigmoid_dec<PRECISE>(t1);
    conf[sb] = sigmoid_dec<PRECISE>(t4);
    if (FORM == FVB_DECODE_V3) {
      ox[sb] = (s0 + (float)x) * stride;  // yolov3.py:47
      oy[sb] = (s1 + (float)y) * stride;
      ow[sb] = exp_dec<PRECISE>(t2) * aw;  // yolov3.py:48
      oh[sb] = exp_dec<PRECISE>(t3) * ah;
    } else {  // demos/yolov3_u/inference.py:86-89
      const float s2 = sigmoid_dec<PRECISE>(t2) * 2.0f, s3 = sigmoid_dec<PRECISE>(t3) * 2.0f;
      ox[sb] = ((s0 * 2.0f - 0.5f) + (float)x) * stride;
      oy[sb] = ((s1 * 2.0f - 0.5f) + (float)y) * stride;
      ow[sb] = (s2 * s2) * aw;
      oh[sb] = (s3 * s3) * ah;
    }
    // yolov3_loss.py:63-64 with target 0, on the objectness this kernel stores (<= 2.5e-6 relative to torch's sigmoid)
    if (p.bce0 != nullptr && valid) bce += (double)bce_term_zero(conf[sb]);
  }
  if (p.bce0 != nullptr) {
#pragma unroll
    for (int i = 0; i < FVB_MAX_LEVELS; ++i) acc[i] += (i == l) ? bce : 0.0;  // per-lane fp64, reduced once at the end
  }
  __syncwarp();

  // ---- A: flat in-place sigmoid ------------------------------------------------------------------------------------
  {
    float4* b4 = reinterpret_cast<float4*>(buf);
    const int n4 = t.n >> 2;
#pragma unroll 4
    for (int i = lane; i < n4; i += 32) {
      float4 v = b4[i];
      v.x = sigmoid_dec<PRECISE>(v.x);
      v.y = sigmoid_dec<PRECISE>(v.y);
      v.z = sigmoid_dec<PRECISE>(v.z);
      v.w = sigmoid_dec<PRECISE>(v.w);
      b4[i] = v;
    }
    const int i = (n4 << 2) + lane;
    if (i < t.n) buf[i] = sigmoid_dec<PRECISE>(buf[i]);
  }
  __syncwarp();

  // ---- B1: head channels back into the tile, candidate bitmap + records ---------------------------------------------
#pragma unroll
  for (int sb = 0; sb < NSB; ++sb) {
    const int r = sb * 32 + lane;
    const bool valid = r < t.nrows;
    if (valid) {
      float* sr = buf + r * K;
      sr[0] = ox[sb]; sr[1] = oy[sb]; sr[2] = ow[sb]; sr[3] = oh[sb]; sr[4] = conf[sb];
    }
    if (p.bitmap != nullptr && sb * 32 < t.nrows) {
      unsigned m = __ballot_sync(0xffffffffu, valid && conf[sb] > p.conf_thr);  // NMS.py:7 on the stored value
      const unsigned gr = (unsigned)(p.g.row_off[l] + t.row0 + sb * 32);
      const unsigned sh = gr & 31u;
      uint32_t* wptr = p.bitmap + (size_t)t.b * p.bitmap_words + (gr >> 5);
      if (lane == 0) {
        const unsigned lo = m << sh;
        if (lo) atomicOr(wptr, lo);
      } else if (lane == 1 && sh) {
        const unsigned hi = m >> (32u - sh);
        if (hi) atomicOr(wptr + 1, hi);
      }
      // candidate records: score = max_c(cls_c*conf) on the STORED fp32 values (NMS.py:13,16: first maximum on ties)
      if (p.cand_rec != nullptr) {
        while (m) {
          const int rr = __ffs(m) - 1;
          m &= m - 1;
          const float* row = buf + (sb * 32 + rr) * K;
          const float rconf = __shfl_sync(0xffffffffu, conf[sb], rr);
          const float r0 = __shfl_sync(0xffffffffu, ox[sb], rr), r1 = __shfl_sync(0xffffffffu, oy[sb], rr);
          const float r2 = __shfl_sync(0xffffffffu, ow[sb], rr), r3 = __shfl_sync(0xffffffffu, oh[sb], rr);
          // products of two sigmoids are >= +0, so their bit patterns order like unsigned integers
          unsigned best = 0u;
          int bidx = 0x7fffffff;
          for (int ch = 5 + lane; ch < K; ch += 32) {
            const unsigned pr = __float_as_uint(row[ch] * rconf);
            if (bidx == 0x7fffffff || pr > best) {  // strict >: the first maximum wins inside the lane
              best = pr;
              bidx = ch - 5;
            }
          }
          const unsigned wbest = __reduce_max_sync(0xffffffffu, best);
          const int widx = __reduce_min_sync(0xffffffffu, (best == wbest) ? bidx : 0x7fffffff);
          if (lane < 7) {
            float val = r0;
            val = lane == 1 ? r1 : val;
            val = lane == 2 ? r2 : val;
            val = lane == 3 ? r3 : val;
            val = lane == 4 ? rconf : val;
            val = lane == 5 ? __uint_as_float(wbest) : val;
            val = lane == 6 ? __int_as_float(widx) : val;
            p.cand_rec[(t.out_row + sb * 32 + rr) * 8 + lane] = val;
          }
        }
      }
    }
  }
  __syncwarp();

  // ---- copy-out: flat, coalesced ------------------------------------------------------------------------------------
  {
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(t.dst) & 15) == 0) {
      const float4* b4 = reinterpret_cast<const float4*>(buf);
      float4* d4 = reinterpret_cast<float4*>(t.dst);
      const int n4 = t.n >> 2;
#pragma unroll 4
      for (int i = lane; i < n4; i += 32) d4[i] = b4[i];
      done = n4 << 2;
    }
#pragma unroll 8
    for (int i = done + lane; i < t.n; i += 32) t.dst[i] = buf[i];
  }
}

template <int NSB, int FORM, bool PRECISE>
__global__ void __launch_bounds__(kDecodeMaxThreads, 1) decode_kernel(const DecodeParams p) {
  extern __shared__ __align__(16) float dec_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* buf = dec_smem + (size_t)warp * p.stages * p.tile_floats;
  const int nw = (int)gridDim.x * p.warps_per_cta;
  const int gw = (int)blockIdx.x * p.warps_per_cta + warp;
  const int total = (int)p.total_tiles;
  double acc[FVB_MAX_LEVELS];
#pragma unroll
  for (int i = 0; i < FVB_MAX_LEVELS; ++i) acc[i] = 0.0;
  if (p.stages == 1) {
    for (int u = gw; u < total; u += nw) {
      const Tile cur = describe_tile(p, u);
      issue_tile(cur, buf, lane);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();  // every lane's copies have landed
      process_tile<NSB, FORM, PRECISE>(p, cur, buf, acc);
      __syncwarp();  // the buffer may be refilled
    }
  } else {
    int u = gw;
    Tile cur, nxt;
    cur.n = 0;
    if (u < total) {
      cur = describe_tile(p, u);
      issue_tile(cur, buf, lane);
    }
    cp_async_commit();
    int stage = 0;
    while (u < total) {
      const int un = u + nw;
      if (un < total) {
        nxt = describe_tile(p, un);
        issue_tile(nxt, buf + (stage ^ 1) * p.tile_floats, lane);
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
      process_tile<NSB, FORM, PRECISE>(p, cur, buf + stage * p.tile_floats, acc);
      __syncwarp();
      cur = nxt;
      u = un;
      stage ^= 1;
    }
    cp_async_wait<0>();
  }
  if (p.bce0 != nullptr) {
#pragma unroll
    for (int i = 0; i < FVB_MAX_LEVELS; ++i) {
      const double sum = warp_sum(acc[i]);
      if (lane == 0 && i < p.g.L) p.bce0[(size_t)i * nw + gw] = sum;  // fixed tile -> warp map: reproducible
    }
  }
}

// Tuning knobs (environment, read once): warps per CTA, shared-memory stages per warp, rows per tile.
static int knob(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  const int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}

// Launch shape shared by the decode entry point and by the loss (which sums the per-warp partials).
int decode_tile_rows(int K) { return K > 42 ? 16 : (K > 21 ? 32 : (K > 10 ? 64 : 128)); }

int decode_launch_shape(const Geom& g, DecodeShape* s) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) {
    set_error("decode: cannot query the CUDA device (no CPU fallback)");
    (void)cudaGetLastError();
    return FVB_E_CUDA;
  }
  s->tile_rows = decode_tile_rows(g.K);
  {
    const int tr = knob("FVB_DECODE_TILE_ROWS", 0, 0, 128);
    if (tr == 8 || tr == 16 || tr == 32 || tr == 64 || tr == 128) s->tile_rows = tr;
  }
  s->tile_floats = (s->tile_rows * g.K + 3) & ~3;
  int t = 0;
  for (int l = 0; l < g.L; ++l) {
    t += (g.A * g.HW[l] + s->tile_rows - 1) / s->tile_rows;
    s->tiles_level_end[l] = t;
  }
  for (int l = g.L; l < FVB_MAX_LEVELS; ++l) s->tiles_level_end[l] = t;
  s->tiles_per_image = t;
  s->total_tiles = (long long)t * g.B;
  if (s->total_tiles >= (1ll << 31)) {
    set_error("decode: %lld tiles", s->total_tiles);
    return FVB_E_LIMIT;
  }
  s->stages = knob("FVB_DECODE_STAGES", 1, 1, 2);
  const size_t per_warp = (size_t)s->stages * s->tile_floats * sizeof(float);
  int wpc = (int)(kDecodeSmemBudget / per_warp);
  const int want = knob("FVB_DECODE_WARPS", kDecodeWarps, 1, kDecodeMaxThreads / 32);
  if (wpc > want) wpc = want;
  if (wpc < 1) {
    set_error("decode: K=%d rows do not fit the shared-memory tile", g.K);
    return FVB_E_LIMIT;
  }
  s->warps_per_cta = wpc;
  s->smem_bytes = per_warp * wpc;
  long long ctas = (s->total_tiles + wpc - 1) / wpc;
  if (ctas > sms) ctas = sms;
  if (ctas < 1) ctas = 1;
  s->grid = (int)ctas;
  return FVB_OK;
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
static int launch_decode(const DecodeParams& p, const DecodeShape& sh, cudaStream_t s) {
  const int nsb = (p.tile_rows + 31) / 32;
  const void* fn = nullptr;
  if (nsb == 1) fn = (const void*)decode_kernel<1, FORM, PRECISE>;
  else if (nsb == 2) fn = (const void*)decode_kernel<2, FORM, PRECISE>;
  else fn = (const void*)decode_kernel<4, FORM, PRECISE>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem_bytes);
  if (e != cudaSuccess) {
    set_error("decode: cudaFuncSetAttribute(%zu): %s", sh.smem_bytes, cudaGetErrorString(e));
    return FVB_E_CUDA;
  }
  if (nsb == 1) decode_kernel<1, FORM, PRECISE><<<sh.grid, 32 * sh.warps_per_cta, sh.smem_bytes, s>>>(p);
  else if (nsb == 2) decode_kernel<2, FORM, PRECISE><<<sh.grid, 32 * sh.warps_per_cta, sh.smem_bytes, s>>>(p);
  else decode_kernel<4, FORM, PRECISE><<<sh.grid, 32 * sh.warps_per_cta, sh.smem_bytes, s>>>(p);
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

extern "C" int fvb_yolo_decode_partials(const fvb_yolo_geom* geom) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return -1;
  DecodeShape sh;
  if (decode_launch_shape(g, &sh) != FVB_OK) return -1;
  return g.L * sh.grid * sh.warps_per_cta;
}

extern "C" int fvb_yolo_decode_f32(const fvb_yolo_geom* geom, const float* const* d_heads, int form, int precise,
                                   float* d_results, float conf_thr, uint32_t* d_cand_bitmap, float* d_cand_rec,
                                   double* d_conf_bce0, void* stream) {
  DecodeParams p;
  FVB_REQUIRE(d_heads != nullptr && d_results != nullptr, "decode: NULL head/result pointer");
  int rc = make_geom(geom, d_heads, &p.g);
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(form == FVB_DECODE_V3 || form == FVB_DECODE_V5, "decode: unknown form %d", form);
  FVB_REQUIRE(p.g.K <= 1024, "decode: channels %d > 1024", p.g.K);
  for (int l = 0; l < p.g.L; ++l) {
    FVB_REQUIRE(d_heads[l] != nullptr, "decode: head %d is NULL", l);
    FVB_REQUIRE((reinterpret_cast<uintptr_t>(d_heads[l]) & 3) == 0, "decode: head %d is not 4-byte aligned", l);
  }
  FVB_REQUIRE((reinterpret_cast<uintptr_t>(d_results) & 3) == 0, "decode: results not 4-byte aligned");
  FVB_REQUIRE(d_cand_rec == nullptr || d_cand_bitmap != nullptr, "decode: candidate records need the candidate bitmap too");
  if (p.g.B == 0) return FVB_OK;
  DecodeShape sh;
  rc = decode_launch_shape(p.g, &sh);
  if (rc != FVB_OK) return rc;
  p.tile_rows = sh.tile_rows;
  for (int l = 0; l < FVB_MAX_LEVELS; ++l) p.tiles_level_end[l] = sh.tiles_level_end[l];
  p.tiles_per_image = sh.tiles_per_image;
  p.total_tiles = sh.total_tiles;
  p.tile_floats = sh.tile_floats;
  p.warps_per_cta = sh.warps_per_cta;
  p.stages = sh.stages;
  p.out = d_results;
  p.conf_thr = conf_thr;
  p.bitmap = d_cand_bitmap;
  p.bitmap_words = (p.g.row_off[p.g.L] + 31) / 32;
  p.bce0 = d_conf_bce0;
  p.cand_rec = d_cand_rec;
  cudaStream_t s = (cudaStream_t)stream;
  if (form == FVB_DECODE_V3) rc = precise ? launch_decode<FVB_DECODE_V3, true>(p, sh, s) : launch_decode<FVB_DECODE_V3, false>(p, sh, s);
  else rc = precise ? launch_decode<FVB_DECODE_V5, true>(p, sh, s) : launch_decode<FVB_DECODE_V5, false>(p, sh, s);
  if (rc != FVB_OK) return rc;
  count_launch();
  return check_launch("decode_kernel");
}
