// K1: YOLOv3 head decode, one pass over the raw head tensors (HBM-bound).
//
// Replaces detection/models/yolov3.py:33-53 of the reference (~30 ATen launches, ~6 full passes).
// Layout fact the kernel is built on: inside one (image, level) segment the raw head [A,H,W,K] and the decoded
// rows [a*H*W + y*W + x, K] are the SAME flat order, so decode is a contiguous -> contiguous map and a TILE
// (T consecutive rows of one segment, ~5 KB) is one contiguous run of floats on both sides.
//
// Persistent kernel, one CTA per SM, every warp an independent pipeline over tiles gw, gw + NW, gw + 2 NW, ...:
//   * a tile is fetched with cp.async (16-byte when the source run is 16-byte aligned, else 4-byte) into the warp's
//     shared-memory buffer, no registers held; the ~20 warps of a CTA sit at different phases (fetch / math / store),
//     which hides the latency better than two buffers per warp did (measured: 0.320 vs 0.306 ms at B=256);
//   * B0  lane <-> row: the five head channels (x, y, w, h, objectness) of "its" row are read from the tile and
//         decoded (cell coordinates from one pair of integer divisions per lane per tile); the zero-target
//         objectness BCE of Yolov3Loss is accumulated per warp and level in fp64;
//   * A   the whole tile is a flat in-place sigmoid, 128-bit shared-memory accesses, no index arithmetic;
//   * B1  the decoded head channels overwrite their slots; ONE ballot gives the rows' bits of the NMS candidate
//         bitmap; candidate rows (~7 %) get their 32-byte NMS record from the class scores sitting in shared memory;
//   * the finished tile is copied out flat: every store instruction writes 128 contiguous bytes (512 when the
//     destination run is 16-byte aligned).
// The kernel keeps <= 113 KB of shared memory per SM and, through its launch bound, <= 64 registers per thread (58 used, no
// spills; 640 threads -> 40 960 registers) so that one NMS CTA (512 threads x 40 registers, 82 KB) can be co-resident
// (pipeline.py: the NMS tail of batch i under the decode of batch i+1).
#include "common.cuh"

#include <stdlib.h>

namespace fvb {

constexpr int kDecodeWarps = kDecodeThreads / 32;  // default warps per CTA
constexpr int kLaneRowClasses = 32;   // up to this many classes a candidate's record is built by its own lane (process_tile B1)
constexpr int kDecodeMaxThreads = 1024;  // launch bound: caps the kernel at 64 registers per thread

struct DecodeParams {
  Geom g;
  int tile_rows;                        // T: rows per tile (16, 32, 64 or 128; wider rows -> fewer rows)
  int tiles_level_end[FVB_MAX_LEVELS];  // cumulative tiles per image, level by level
  int tiles_per_image;
  int total_tiles;
  float inv_tpi;                        // 1/tiles_per_image, 1/HW, 1/W: seeds of the exact float divmod
  float inv_hw[FVB_MAX_LEVELS], inv_w[FVB_MAX_LEVELS];
  int tile_floats;  // floats per shared-memory buffer (T*K rounded up to a multiple of 4)
  int warps_per_cta;
  float* out;
  float conf_thr;
  uint32_t* bitmap;
  int bitmap_words;
  float* cand_rec;  // [B][N][8] = {row[0..3], conf, max_c(cls*conf), argmax as int bits, -}, written for candidates only
  double* bce0;     // one zero-target objectness BCE partial per tile, level-major: [l][b][tile of the level]
  unsigned* sched;  // [0] next tile, [1] warps that have drained; both zero between launches
  int batch_max;    // tiles drawn per atomic while the queue is long
  long long* trace;     // debug (fvb_debug_set_nms_trace): [9 B] earliest CTA start, [9 B + 1] latest CTA end; NULL in production
  unsigned* tile_done;  // optional [B]: finished tiles per image, published with release semantics (consumer: yolo_nms_kernel
                        // launched as a programmatic dependent, which starts an image's NMS while later images still decode)
};

// Publish `n` finished tiles of image b: every lane's stores of those tiles (results, candidate bits, records, objectness
// partials) were ordered before this point by __syncwarp(); the gpu-scope fence + relaxed atomic by lane 0 is the release.
__device__ __forceinline__ void publish_tiles(const DecodeParams& p, int b, int n, int lane) {
  if (lane == 0 && n > 0) {
    __threadfence();
    atomicAdd(&p.tile_done[b], (unsigned)n);
  }
}

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

// n / d and n % d for 0 <= n < 2^23, 1 <= d: float estimate (inv = 1.0f/d) corrected by at most one
__device__ __forceinline__ void divmod_f(int n, int d, float inv, int& q, int& r) {
  q = __float2int_rz(__int2float_rn(n) * inv);
  r = n - q * d;
  if (r < 0) {
    q -= 1;
    r += d;
  } else if (r >= d) {
    q += 1;
    r -= d;
  }
}

struct Tile {
  const float* src;  // first float of the run in the raw head
  float* dst;        // first float of the run in results
  size_t out_row;    // global row index (b*N + row_off[l] + row0)
  int n;             // floats in the run (nrows*K)
  int nrows, l, b, row0;
  int part;          // index of this tile's objectness partial
};

__device__ __forceinline__ Tile describe_tile(const DecodeParams& p, int u) {
  Tile t;
  int b, r;
  divmod_f(u, p.tiles_per_image, p.inv_tpi, b, r);
  int l = 0;
#pragma unroll
  for (int i = 0; i < FVB_MAX_LEVELS - 1; ++i)
    if (i < p.g.L - 1 && r >= p.tiles_level_end[i]) l = i + 1;
  const int lvl_first = l > 0 ? p.tiles_level_end[l - 1] : 0;
  const int tiles_l = p.tiles_level_end[l] - lvl_first;
  r -= lvl_first;
  const int rows_l = p.g.A * p.g.HW[l];
  t.row0 = r * p.tile_rows;
  t.nrows = min(p.tile_rows, rows_l - t.row0);
  t.n = t.nrows * p.g.K;
  t.l = l;
  t.b = b;
  t.part = lvl_first * p.g.B + b * tiles_l + r;
  t.out_row = (size_t)b * p.g.row_off[p.g.L] + p.g.row_off[l] + t.row0;
  t.src = p.g.head[l] + ((size_t)b * rows_l + t.row0) * p.g.K;
  t.dst = p.out + t.out_row * p.g.K;
  return t;
}

// All flat loops below run `full` unguarded warp-wide steps (warp-uniform trip count: no divergence bookkeeping) and
// one guarded remainder.
__device__ __forceinline__ void issue_tile(const Tile& t, float* buf, int lane) {
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(t.src) & 15) == 0) {
    const int n4 = t.n >> 2;
    const float* s = t.src + 4 * lane;
    float* d = buf + 4 * lane;
    int k = n4 >> 5;
#pragma unroll 1
    for (; k >= 4; k -= 4, s += 512, d += 512) {
      cp_async16(d, s);
      cp_async16(d + 128, s + 128);
      cp_async16(d + 256, s + 256);
      cp_async16(d + 384, s + 384);
    }
#pragma unroll 1
    for (; k > 0; --k, s += 128, d += 128) cp_async16(d, s);
    if (lane < (n4 & 31)) cp_async16(d, s);
    done = n4 << 2;
    if (done + lane < t.n) cp_async4(buf + done + lane, t.src + done + lane);
    return;
  }
  const float* s = t.src + lane;
  float* d = buf + lane;
  int k = t.n >> 5;
#pragma unroll 1
  for (; k >= 8; k -= 8, s += 256, d += 256) {
#pragma unroll
    for (int j = 0; j < 8; ++j) cp_async4(d + 32 * j, s + 32 * j);
  }
#pragma unroll 1
  for (; k > 0; --k, s += 32, d += 32) cp_async4(d, s);
  if (lane < (t.n & 31)) cp_async4(d, s);
}

// NCHW heads ([B, A*K, H, W], the conv output): the tile's rows are T consecutive cells of one anchor plane stack,
// channel k of row r lives HW floats after channel k-1.  Lanes <-> (row, channel phase): every instruction reads
// runs of min(T,32) consecutive cells (64-128 contiguous bytes) of 32/min(T,32) channels and drops them transposed
// into the row-major tile, which the rest of the pipeline then treats exactly like the [B,A,H,W,K] case.
template <int NSB>
__device__ __forceinline__ void issue_tile_nchw(const DecodeParams& p, const Tile& t, float* buf, int lane) {
  const int K = p.g.K, l = t.l, HW = p.g.HW[l];
  const int ri = p.tile_rows < 32 ? p.tile_rows : 32;  // rows per instruction (power of two)
  const int kl = 32 / ri;                              // channels per instruction
  const int r_in = lane & (ri - 1), kk = lane / ri;
#pragma unroll
  for (int sb = 0; sb < NSB; ++sb) {
    const int r = sb * 32 + r_in;
    if (r < t.nrows) {
      int a, pos;
      divmod_f(t.row0 + r, HW, p.inv_hw[l], a, pos);
      const float* s = p.g.head[l] + ((size_t)(t.b * p.g.A + a) * K + kk) * HW + pos;
      float* d = buf + r * K + kk;
      const size_t step = (size_t)kl * HW;
      int k = kk;
#pragma unroll 1
      for (; k + 3 * kl < K; k += 4 * kl, s += 4 * step, d += 4 * kl) {
        cp_async4(d, s);
        cp_async4(d + kl, s + step);
        cp_async4(d + 2 * kl, s + 2 * step);
        cp_async4(d + 3 * kl, s + 3 * step);
      }
#pragma unroll 1
      for (; k < K; k += kl, s += step, d += kl) cp_async4(d, s);
    }
  }
}

// NSB = 32-row sub-blocks per tile (lane <-> row passes)
template <int NSB, int FORM, bool PRECISE>
__device__ __forceinline__ void process_tile(const DecodeParams& p, const Tile& t, float* buf) {
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
    int a, yx, y, x;
    divmod_f(t.row0 + (valid ? r : 0), HW, p.inv_hw[l], a, yx);
    divmod_f(yx, W, p.inv_w[l], y, x);
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
    bce = warp_sum(bce);  // fixed shuffle tree, one partial per tile: reproducible whichever warp runs the tile
    if (lane == 0) p.bce0[t.part] = bce;
  }
  __syncwarp();

  // ---- A: flat in-place sigmoid ------------------------------------------------------------------------------------
  {
    const int n4 = t.n >> 2;
    float4* q = reinterpret_cast<float4*>(buf) + lane;
    int k = n4 >> 5;
#pragma unroll 1
    for (; k >= 4; k -= 4, q += 128) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = q[32 * j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j].x = sigmoid_dec<PRECISE>(v[j].x);
        v[j].y = sigmoid_dec<PRECISE>(v[j].y);
        v[j].z = sigmoid_dec<PRECISE>(v[j].z);
        v[j].w = sigmoid_dec<PRECISE>(v[j].w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) q[32 * j] = v[j];
    }
    // <= 3 full steps + the ragged one, fused: at most 4 guarded vectors
    {
      const int rem = (k << 5) + (n4 & 31);  // vectors left, < 128
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane + 32 * j < rem) v[j] = q[32 * j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j].x = sigmoid_dec<PRECISE>(v[j].x);
        v[j].y = sigmoid_dec<PRECISE>(v[j].y);
        v[j].z = sigmoid_dec<PRECISE>(v[j].z);
        v[j].w = sigmoid_dec<PRECISE>(v[j].w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane + 32 * j < rem) q[32 * j] = v[j];
    }
    const int i = (n4 << 2) + lane;
    if (i < t.n) buf[i] = sigmoid_dec<PRECISE>(buf[i]);
  }
  __syncwarp();

  // ---- B1: head channels back into the tile, candidate bitmap + records ---------------------------------------------
  // Records: the warp-serial walk costs ~100 instructions per candidate, the lane-per-row walk ~(3 + 3 NSB) per class
  // whatever the number of candidates: few classes, or a candidate-dense tile (low thresholds), take the latter.
  bool lane_rows = K - 5 <= kLaneRowClasses;
  if (p.cand_rec != nullptr && !lane_rows) {
    int nc = 0;
#pragma unroll
    for (int sb = 0; sb < NSB; ++sb) nc += __popc(__ballot_sync(0xffffffffu, sb * 32 + lane < t.nrows && conf[sb] > p.conf_thr));
    lane_rows = nc * 100 > (K - 5) * (3 + 3 * NSB);
  }
#pragma unroll
  for (int sb = 0; sb < NSB; ++sb) {
    const int r = sb * 32 + lane;
    const bool valid = r < t.nrows;
    if (valid) {
      float* sr = buf + r * K;
      sr[0] = ox[sb]; sr[1] = oy[sb]; sr[2] = ow[sb]; sr[3] = oh[sb]; sr[4] = conf[sb];
    }
    __syncwarp();
    if (p.bitmap != nullptr && sb * 32 < t.nrows) {
      unsigned m = __ballot_sync(0xffffffffu, valid && conf[sb] > p.conf_thr);  // NMS.py:7 on the stored value
      if (m) {
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
      }
      // candidate records: score = max_c(cls_c*conf) on the STORED fp32 values (NMS.py:13,16: first maximum on ties)
      if (p.cand_rec != nullptr && !lane_rows) {
        while (m) {
          const int rr = __ffs(m) - 1;
          m &= m - 1;
          const float* row = buf + (sb * 32 + rr) * K;
          const float rconf = __shfl_sync(0xffffffffu, conf[sb], rr);
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
            // lanes 0..4: box + objectness, which B1 just stored into the tile row itself
            const float val = lane < 5 ? row[lane] : (lane == 5 ? __uint_as_float(wbest) : __int_as_float(widx));
            p.cand_rec[(t.out_row + sb * 32 + rr) * 8 + lane] = val;
          }
        }
      }
    }
  }
  // few classes (K = 15: a candidate every 16 rows, several per 32-row group) or a dense tile: every candidate lane walks its OWN rows' class
  // scores -- all candidates of the tile in parallel, the NSB rows of a lane as independent chains, one 32-byte record
  // store per candidate -- instead of the warp taking the candidates one after the other (measured at 608 / C=10 / B=1024:
  // 740 cycles per candidate, decode 0.46 -> 0.65 ms).  Same arithmetic and tie rule (first maximum) as the loop above.
  if (p.cand_rec != nullptr && lane_rows) {
    bool cand[NSB];
    bool any = false;
#pragma unroll
    for (int sb = 0; sb < NSB; ++sb) {
      cand[sb] = (sb * 32 + lane < t.nrows) && conf[sb] > p.conf_thr;
      any |= cand[sb];
    }
    if (any) {
      unsigned best[NSB];
      int bidx[NSB];
      const float* cls[NSB];  // a lane's row of sub-block sb; rows past the tile's end (never stored) fall back to row 0
#pragma unroll
      for (int sb = 0; sb < NSB; ++sb) {
        cls[sb] = buf + (sb * 32 + lane < t.nrows ? sb * 32 + lane : 0) * K + 5;
        best[sb] = __float_as_uint(cls[sb][0] * conf[sb]);
        bidx[sb] = 0;
      }
      for (int c = 1; c < K - 5; ++c) {
#pragma unroll
        for (int sb = 0; sb < NSB; ++sb) {
          const unsigned pr = __float_as_uint(cls[sb][c] * conf[sb]);
          if (pr > best[sb]) {  // strict >: the first maximum wins
            best[sb] = pr;
            bidx[sb] = c;
          }
        }
      }
#pragma unroll
      for (int sb = 0; sb < NSB; ++sb) {
        if (cand[sb]) {
          float4* rec = reinterpret_cast<float4*>(p.cand_rec + (t.out_row + sb * 32 + lane) * 8);
          rec[0] = make_float4(ox[sb], oy[sb], ow[sb], oh[sb]);
          rec[1] = make_float4(conf[sb], __uint_as_float(best[sb]), __int_as_float(bidx[sb]), 0.0f);
        }
      }
    }
  }
  __syncwarp();

  // ---- copy-out: flat, coalesced ------------------------------------------------------------------------------------
  if ((reinterpret_cast<uintptr_t>(t.dst) & 15) == 0) {
    const int n4 = t.n >> 2;
    const float4* q = reinterpret_cast<const float4*>(buf) + lane;
    float4* d = reinterpret_cast<float4*>(t.dst) + lane;
    int k = n4 >> 5;
#pragma unroll 1
    for (; k >= 4; k -= 4, q += 128, d += 128) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = q[32 * j];
#pragma unroll
      for (int j = 0; j < 4; ++j) d[32 * j] = v[j];
    }
    const int rem = (k << 5) + (n4 & 31);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane + 32 * j < rem) d[32 * j] = q[32 * j];
    const int i = (n4 << 2) + lane;
    if (i < t.n) t.dst[i] = buf[i];
  } else {
    const float* q = buf + lane;
    float* d = t.dst + lane;
    int k = t.n >> 5;
#pragma unroll 1
    for (; k >= 8; k -= 8, q += 256, d += 256) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = q[32 * j];
#pragma unroll
      for (int j = 0; j < 8; ++j) d[32 * j] = v[j];
    }
    const int rem = (k << 5) + (t.n & 31);  // < 256
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (lane + 32 * j < rem) d[32 * j] = q[32 * j];
  }
}

// PUB: publish per-image progress (p.tile_done) for the overlapped NMS.  A template parameter, not a run-time test: the three
// counters it needs pushed the 64-row-tile instantiation (narrow rows, which never publish) into register spills.
template <int NSB, int FORM, bool PRECISE, bool PUB>
__global__ void __launch_bounds__(kDecodeMaxThreads, 1) decode_kernel(const DecodeParams p) {
  extern __shared__ __align__(16) float dec_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* buf = dec_smem + (size_t)warp * p.tile_floats;
  const int total = p.total_tiles;
  // Programmatic dependent launch: the NMS kernel enqueued right behind this one may be scheduled as soon as every CTA of this
  // grid has got here (i.e. IS RESIDENT -- which is what makes its spinning on tile_done[] deadlock-free); it does not wait for
  // this grid to finish.  Without a dependent this is a no-op.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.trace != nullptr && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMin(p.trace + 9 * p.g.B, t);
  }
  int pend_b = -1, pend_n = 0;  // finished tiles of image pend_b not yet published
  int pub_n = 0;                // ... of which this many are due at the next opportunity
  // dynamic tile queue: tiles are handed out in memory order to whichever warp is free, so a CTA that starts late or
  // shares its SM with NMS CTAs of the previous batch simply takes fewer tiles.  A warp draws a BATCH of consecutive
  // tiles per atomic (same-address atomics serialise in one L2 slice: one per tile costs more than the tile), and the
  // batch shrinks towards the end of the queue (guided self-scheduling) so the tail stays one tile long.
  const int nw = (int)(gridDim.x * (blockDim.x >> 5));
  const int gmax = p.batch_max;
  int u = 0, uend = 0;   // current batch [u, uend)
  int nu = 0, nend = 0;  // prefetched next batch
  int seen = 0;          // last ticket value seen: estimate of the queue position
  {
    if (lane == 0) nu = (int)atomicAdd(&p.sched[0], (unsigned)gmax);
    nu = __shfl_sync(0xffffffffu, nu, 0);
    nend = min(nu + gmax, total);
    seen = nu;
  }
  bool have_next = true;
  while (true) {
    if (u >= uend) {
      if (!have_next || nu >= total) break;
      u = nu;
      uend = nend;
      have_next = false;
    }
    const Tile cur = describe_tile(p, u);
    if (p.g.nchw) issue_tile_nchw<NSB>(p, cur, buf, lane);
    else issue_tile(cur, buf, lane);
    cp_async_commit();
    if (PUB && pub_n) {
      // the previous batch's publication, deferred to here: its fence waits for that batch's stores to be acknowledged --
      // with this tile's loads already in flight behind it the warp loses nothing (at the batch's end it cost ~1 us per batch)
      publish_tiles(p, pend_b, pub_n, lane);
      pub_n = 0;
    }
    int t0 = 0, g = 1;
    const bool fetch = !have_next && (u + 1 >= uend);  // on the last tile of the batch: draw the next batch now
    if (fetch) {
      const int rem = total - seen;
      g = min(gmax, max(1, rem / (2 * nw)));
      if (lane == 0) t0 = (int)atomicAdd(&p.sched[0], (unsigned)g);  // its latency hides behind this tile
    }
    cp_async_wait<0>();
    __syncwarp();  // every lane's copies have landed
    process_tile<NSB, FORM, PRECISE>(p, cur, buf);
    __syncwarp();  // the buffer may be refilled
    if (PUB) {
      // one publication per image per drawn batch (a batch is <= batch_max consecutive tiles, mostly of one image)
      if (cur.b != pend_b) {  // crossed into another image (rare): the old image's count goes out now
        publish_tiles(p, pend_b, pend_n, lane);
        pend_b = cur.b;
        pend_n = 0;
      }
      ++pend_n;
      if (u + 1 >= uend) {  // batch finished: publish after the next tile's loads have been issued
        pub_n = pend_n;
        pend_n = 0;
      }
    }
    if (fetch) {
      nu = __shfl_sync(0xffffffffu, t0, 0);
      nend = min(nu + g, total);
      seen = nu;
      have_next = true;
    }
    ++u;
  }
  if (PUB) publish_tiles(p, pend_b, pub_n + pend_n, lane);
  if (p.trace != nullptr && lane == 0) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMax(p.trace + 9 * p.g.B + 1, t);
  }
  // the last warp to drain re-arms the queue for the next launch
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(&p.sched[1], 1u) == (unsigned)nw - 1u) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

// Tuning knobs (environment; each is read ONCE per process, at the first launch): warps per CTA, tiles drawn per ticket.
static int read_knob(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  const int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}
static int g_knob_warps = -1, g_knob_batch = -1;  // -1 = not read yet (benign race: every reader computes the same value)
static int knob_warps() {
  if (g_knob_warps < 0) g_knob_warps = read_knob("FVB_DECODE_WARPS", 0, 0, kDecodeMaxThreads / 32);  // 0 = built-in defaults
  return g_knob_warps;
}
static int knob_batch() {
  if (g_knob_batch < 0) g_knob_batch = read_knob("FVB_DECODE_BATCH", 0, 0, 64);  // 0 = by queue length (below)
  return g_knob_batch;
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
  s->stages = 1;
  const size_t per_warp = (size_t)s->tile_floats * sizeof(float);
  // Narrow rows (K <= 21: 64-128 rows, < 8 KB per tile) are bound by the bytes a CTA keeps in flight (one tile per warp) and by
  // their denser per-row work: they take the whole SM -- 32 warps x 64 registers, up to 200 KB of tiles -- since more than 20
  // warps leave no registers for an NMS CTA anyway.  608 / C=10, fused decode / whole step: 24 warps 0.540 / 0.721 ms at B=1024
  // and 0.103 / 0.119 ms at B=128; 32 warps 0.523 / 0.705 and 0.101 / 0.114 (128-row tiles: slower at small batches).
  // Wide rows are best at 20 warps (0.304 vs 0.306 ms) and keep the 112 KB budget that leaves room for an NMS CTA.
  const bool narrow = s->tile_rows >= 64;
  const size_t budget = narrow ? (size_t)200 * 1024 : kDecodeSmemBudget;
  int wpc = (int)(budget / per_warp);
  const int want = knob_warps() ? knob_warps() : (narrow ? 32 : kDecodeWarps);
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
  FVB_REQUIRE(g->head_layout == FVB_HEAD_BAHWK || g->head_layout == FVB_HEAD_NCHW, "unknown head_layout %d", g->head_layout);
  out->nchw = g->head_layout == FVB_HEAD_NCHW;
  return FVB_OK;
}

template <int NSB, int FORM, bool PRECISE, bool PUB>
static int launch_decode_inst(const DecodeParams& p, const DecodeShape& sh, cudaStream_t s) {
  const void* fn = (const void*)decode_kernel<NSB, FORM, PRECISE, PUB>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem_bytes);
  // Ask for the largest shared-memory carve-out when an NMS CTA can join this one (programmatic dependent, or the previous
  // batch's tail): the SM partition is fixed while a CTA is resident, and the NMS CTA needs its 82 KB next to our ~109 KB.
  // Not otherwise: narrow rows (24 warps, no room for an NMS CTA anyway) lose 4.5 % to the smaller L1 (608 / C=10 / B=1024:
  // fused decode 0.550 -> 0.585 ms).
  const bool room = sh.warps_per_cta * 32 * 64 + 512 * 40 <= 65536;
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, room ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
  if (e != cudaSuccess) {
    set_error("decode: cudaFuncSetAttribute(%zu): %s", sh.smem_bytes, cudaGetErrorString(e));
    return FVB_E_CUDA;
  }
  decode_kernel<NSB, FORM, PRECISE, PUB><<<sh.grid, 32 * sh.warps_per_cta, sh.smem_bytes, s>>>(p);
  return FVB_OK;
}

template <int FORM, bool PRECISE>
static int launch_decode(const DecodeParams& p, const DecodeShape& sh, cudaStream_t s) {
  const int nsb = (p.tile_rows + 31) / 32;
  const bool pub = p.tile_done != nullptr;
  if (nsb == 1) return pub ? launch_decode_inst<1, FORM, PRECISE, true>(p, sh, s) : launch_decode_inst<1, FORM, PRECISE, false>(p, sh, s);
  if (nsb == 2) return pub ? launch_decode_inst<2, FORM, PRECISE, true>(p, sh, s) : launch_decode_inst<2, FORM, PRECISE, false>(p, sh, s);
  return pub ? launch_decode_inst<4, FORM, PRECISE, true>(p, sh, s) : launch_decode_inst<4, FORM, PRECISE, false>(p, sh, s);
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
  long long t = 0;
  const int tr = decode_tile_rows(g.K);
  for (int l = 0; l < g.L; ++l) t += (g.A * g.HW[l] + tr - 1) / tr;
  t *= g.B;
  if (t >= (1ll << 23)) {
    set_error("decode: %lld tiles (limit 2^23)", t);
    return -1;
  }
  return (int)t;
}

extern "C" size_t fvb_yolo_decode_workspace_bytes(void) { return 256; }

/* debug hook (tools/decode_sweep.py): forget the cached FVB_DECODE_* knobs so the next launch re-reads the environment */
extern "C" void fvb_debug_reload_knobs(void) { g_knob_warps = g_knob_batch = -1; }

extern "C" int fvb_yolo_decode_tiles_per_image(const fvb_yolo_geom* geom) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return -1;
  int t = 0;
  const int tr = decode_tile_rows(g.K);
  for (int l = 0; l < g.L; ++l) t += (g.A * g.HW[l] + tr - 1) / tr;
  return t;
}

// Can one NMS CTA (512 threads x 40 registers, nms.cu) be resident beside a decode CTA of this geometry?  Only then does the
// programmatic-dependent NMS really run under the decode; otherwise its CTAs just queue until decode CTAs leave.
extern "C" int fvb_yolo_decode_leaves_room_for_nms(const fvb_yolo_geom* geom) {
  Geom g;
  if (make_geom(geom, nullptr, &g) != FVB_OK) return -1;
  if (g.B == 0) return 0;
  DecodeShape sh;
  if (decode_launch_shape(g, &sh) != FVB_OK) return -1;
  const int regs = sh.warps_per_cta * 32 * 64 + 512 * 40;
  const size_t smem = sh.smem_bytes + 84 * 1024 + 2 * 1024;
  return (regs <= 65536 && smem <= 228 * 1024 && sh.warps_per_cta * 32 + 512 <= 2048) ? 1 : 0;
}

extern "C" int fvb_yolo_decode_f32(const fvb_yolo_geom* geom, const float* const* d_heads, int form, int precise,
                                   float* d_results, float conf_thr, uint32_t* d_cand_bitmap, float* d_cand_rec,
                                   double* d_conf_bce0, void* d_ws, void* stream) {
  return fvb_yolo_decode_sync_f32(geom, d_heads, form, precise, d_results, conf_thr, d_cand_bitmap, d_cand_rec, d_conf_bce0,
                                  nullptr, d_ws, stream);
}

extern "C" int fvb_yolo_decode_sync_f32(const fvb_yolo_geom* geom, const float* const* d_heads, int form, int precise,
                                        float* d_results, float conf_thr, uint32_t* d_cand_bitmap, float* d_cand_rec,
                                        double* d_conf_bce0, uint32_t* d_tile_sync, void* d_ws, void* stream) {
  DecodeParams p;
  FVB_REQUIRE(d_heads != nullptr && d_results != nullptr && d_ws != nullptr, "decode: NULL head/result/workspace pointer");
  int rc = make_geom(geom, d_heads, &p.g);
  if (rc != FVB_OK) return rc;
  FVB_REQUIRE(form == FVB_DECODE_V3 || form == FVB_DECODE_V5, "decode: unknown form %d", form);
  FVB_REQUIRE(p.g.K <= 1024, "decode: channels %d > 1024", p.g.K);
  for (int l = 0; l < p.g.L; ++l) {
    FVB_REQUIRE(d_heads[l] != nullptr, "decode: head %d is NULL", l);
    FVB_REQUIRE((reinterpret_cast<uintptr_t>(d_heads[l]) & 3) == 0, "decode: head %d is not 4-byte aligned", l);
  }
  FVB_REQUIRE((reinterpret_cast<uintptr_t>(d_results) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_ws) & 7) == 0, "decode: results / workspace misaligned");
  FVB_REQUIRE(d_cand_rec == nullptr || d_cand_bitmap != nullptr, "decode: candidate records need the candidate bitmap too");
  FVB_REQUIRE(((uintptr_t)d_cand_rec & 15) == 0, "decode: candidate records must be 16-byte aligned");
  if (p.g.B == 0) return FVB_OK;
  DecodeShape sh;
  rc = decode_launch_shape(p.g, &sh);
  if (rc != FVB_OK) return rc;
  if (sh.total_tiles >= (1ll << 23)) {
    set_error("decode: %lld tiles (limit 2^23)", sh.total_tiles);
    return FVB_E_LIMIT;
  }
  p.tile_rows = sh.tile_rows;
  for (int l = 0; l < FVB_MAX_LEVELS; ++l) {
    p.tiles_level_end[l] = sh.tiles_level_end[l];
    p.inv_hw[l] = l < p.g.L ? 1.0f / (float)p.g.HW[l] : 0.0f;
    p.inv_w[l] = l < p.g.L ? 1.0f / (float)p.g.W[l] : 0.0f;
  }
  p.tiles_per_image = sh.tiles_per_image;
  p.total_tiles = (int)sh.total_tiles;
  p.inv_tpi = 1.0f / (float)sh.tiles_per_image;
  p.tile_floats = sh.tile_floats;
  p.warps_per_cta = sh.warps_per_cta;
  p.out = d_results;
  p.conf_thr = conf_thr;
  p.bitmap = d_cand_bitmap;
  p.bitmap_words = (p.g.row_off[p.g.L] + 31) / 32;
  p.bce0 = d_conf_bce0;
  p.cand_rec = d_cand_rec;
  p.sched = (unsigned*)d_ws;
  p.tile_done = d_tile_sync;
  p.trace = debug_trace_ptr();
  FVB_REQUIRE(((uintptr_t)d_tile_sync & 3) == 0, "decode: tile_sync misaligned");
  // tiles drawn per ticket: 8 while the queue is long; a short queue (a data-parallel shard: 608 / C=10 / 128 images is ~10 tiles
  // per warp) starts with smaller batches so that the first round of draws does not already decide the load balance
  // (B=128: step 0.1103 ms at 8, 0.1066 at 4)
  {
    const long long nw = (long long)sh.grid * sh.warps_per_cta;
    const long long by_len = sh.total_tiles / (2 * nw);
    p.batch_max = knob_batch() ? knob_batch() : (int)(by_len < 1 ? 1 : (by_len > 8 ? 8 : by_len));
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (form == FVB_DECODE_V3) rc = precise ? launch_decode<FVB_DECODE_V3, true>(p, sh, s) : launch_decode<FVB_DECODE_V3, false>(p, sh, s);
  else rc = precise ? launch_decode<FVB_DECODE_V5, true>(p, sh, s) : launch_decode<FVB_DECODE_V5, false>(p, sh, s);
  if (rc != FVB_OK) return rc;
  count_launch();
  return check_launch("decode_kernel");
}
