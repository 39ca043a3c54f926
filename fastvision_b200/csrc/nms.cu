// K3: confidence filter + NMS for every image of a batch in ONE launch (one CTA per image).
//
// Replaces the per-image Python loop of the reference (utils/fit.py:94-95) around
// non_max_suppression (detection/tools/NMS.py:5-23; demo flavours demos/yolov3_u/utils/nms.py) and
// the third-party torchvision.ops.nms it calls.  Phases, all inside one CTA:
//   0  candidates: popcount/scan of the image's candidate bitmap -> rows in ascending order = the
//      reference's boolean-mask order, so "slot" order is the tie-break order of its stable sort;
//   1  one thread per candidate: load its 32-byte record {x,y,w,h,conf,max_c(cls*conf),argmax}
//      (written by the decode kernel while the row was in registers, or by yolo_score_kernel when
//      NMS is called on its own), xywh->xyxy, optional class gap (box + cat*max_wh in fp32);
//   2  block sort of (rank desc, slot asc) keys: bitonic network in registers/shared memory up to kCapS keys, LSD radix beyond;
//   3  chunked greedy suppression with a kept list (nms.cuh);
//   4  padded outputs + count; consumed bitmap words are cleared for the next step.
// Candidates live in shared memory up to kCapS per image (keys, boxes, rows: 28 B each -- 82 KB per CTA so
// that TWO CTAs fit an SM and 256 images are one wave; sorted there with the in-register bitonic network of nms.cuh); beyond that the same code runs on a global workspace
// slice (slower, still exact) so there is no overflow case.
#include "nms.cuh"

#include <math.h>

namespace fvb {

__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr int kCapS = 2048;  // candidates per image held in shared memory

struct YoloNmsParams {
  const float* results;
  int B, N, K;
  float conf_thr, iou_thr;
  int max_det, flavour;
  float max_wh;
  int max_nms;
  uint32_t* bitmap;  // [B, words]; either caller-provided (decode) or a workspace slice
  const float* rec;  // [B, N, 8] candidate records
  int words;
  int clear_bitmap;
  float* out_boxes;
  float* out_scores;
  long long* out_cls;
  int* out_rows;
  int* out_cnt;
  // global fallback, per image strides of N entries
  unsigned long long* ws_keys;  // [B][2][N]
  float4* ws_box;               // [B][N]
  int* ws_row;                  // [B][N]
  long long* trace;             // debug: [B][8] clock64 stamps per phase (NULL in production)
  // Programmatic-dependent form (fvb_yolo_nms_after_decode_f32): the grid starts while the decode kernel still runs.
  unsigned* tile_sync;          // [B] finished tiles per image (published by decode_kernel), [B] = this grid's exit counter; or NULL
  unsigned tiles_per_image;
};

__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct NmsSmemLayout {
  size_t keys0, box, row, cnt, warp_tot, kbox, karea, kslot, gs, misc, total;
};

__host__ __device__ inline NmsSmemLayout nms_layout(int cap, int max_keep, int kNmsWarps = fvb::kNmsWarps) {
  NmsSmemLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes, size_t align) {
    o = (o + align - 1) / align * align;
    size_t r = o;
    o += bytes;
    return r;
  };
  L.keys0 = take((size_t)cap * 8, 16);  // sorted in place (bitonic); the radix sort's second buffer exists only in the global fallback
  L.box = take((size_t)cap * 16, 16);
  L.row = take((size_t)cap * 4, 16);
  L.cnt = take((size_t)kNmsWarps * 256 * 4, 16);
  L.warp_tot = take((size_t)(kNmsWarps + 1) * 4, 16);
  L.kbox = take((size_t)max_keep * 16, 16);
  L.karea = take((size_t)max_keep * 4, 16);
  L.kslot = take((size_t)max_keep * 4, 16);
  L.gs = take(sizeof(GreedyShared), 16);
  L.misc = take(64, 16);
  L.total = o;
  return L;
}

template <int NT>
__device__ __forceinline__ void yolo_nms_image(const YoloNmsParams& p, const int b, unsigned char* smem) {
  constexpr int kNmsWarps = NT / 32, kNmsThreads = NT;  // shadow the namespace constants inside this function
  const NmsSmemLayout L = nms_layout(kCapS, p.max_det, kNmsWarps);
  const int lane = threadIdx.x & 31;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + L.cnt);
  uint32_t* warp_tot = reinterpret_cast<uint32_t*>(smem + L.warp_tot);
  float4* kbox = reinterpret_cast<float4*>(smem + L.kbox);
  float* karea = reinterpret_cast<float*>(smem + L.karea);
  int* kslot = reinterpret_cast<int*>(smem + L.kslot);
  GreedyShared* gs = reinterpret_cast<GreedyShared*>(smem + L.gs);
  int* misc = reinterpret_cast<int*>(smem + L.misc);  // [0] n_valid

  uint32_t* bm = p.bitmap + (size_t)b * p.words;
  if (p.trace && threadIdx.x == 0) p.trace[b * 8 + 0] = gtime_ns();

  // ---- phase 0: ordered candidate rows from the bitmap -------------------------------------------------
  const int wpt = (p.words + kNmsThreads - 1) / kNmsThreads;
  const int wbeg = min(p.words, (int)threadIdx.x * wpt), wend = min(p.words, wbeg + wpt);
  uint32_t my = 0;
  for (int w = wbeg; w < wend; ++w) my += __popc(bm[w]);
  uint32_t base = block_exclusive_scan<NT>(my, warp_tot);
  const int n = (int)warp_tot[kNmsWarps];
  __syncthreads();  // warp_tot is reused by the sort

  unsigned long long *keys0, *keys1;
  float4* sbox;
  int* srow;
  if (n <= kCapS) {
    keys0 = reinterpret_cast<unsigned long long*>(smem + L.keys0);
    keys1 = nullptr;
    sbox = reinterpret_cast<float4*>(smem + L.box);
    srow = reinterpret_cast<int*>(smem + L.row);
  } else {
    keys0 = p.ws_keys + (size_t)b * 2 * p.N;
    keys1 = keys0 + p.N;
    sbox = p.ws_box + (size_t)b * p.N;
    srow = p.ws_row + (size_t)b * p.N;
  }
  {
    uint32_t pos = base;
    for (int w = wbeg; w < wend; ++w) {
      uint32_t bits = bm[w];
      while (bits) {
        int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        srow[pos++] = w * 32 + bit;
      }
    }
  }
  if (threadIdx.x == 0) misc[0] = 0;
  __syncthreads();
  if (p.clear_bitmap)
    for (int w = threadIdx.x; w < p.words; w += kNmsThreads) bm[w] = 0u;

  if (n == 0) {
    if (threadIdx.x == 0) p.out_cnt[b] = 0;
    return;
  }

  if (p.trace && threadIdx.x == 0) p.trace[b * 8 + 1] = gtime_ns();
  // ---- phase 1: one thread per candidate, from its 32-byte record -----------------------------------------
  const float4* rec4 = reinterpret_cast<const float4*>(p.rec + (size_t)b * p.N * 8);
  int valid_local = 0;
  for (int i = threadIdx.x; i < n; i += kNmsThreads) {
    const int r = srow[i];
    const float4 q0 = rec4[(size_t)r * 2], q1 = rec4[(size_t)r * 2 + 1];
    const float conf = q1.x, best = q1.y;
    const int bidx = __float_as_int(q1.z);
    Box bx;
    if (p.flavour == FVB_NMS_DEMO) {  // boxes arrive as xyxy (demos/yolov3_u/utils/nms.py:8)
      bx.x1 = q0.x; bx.y1 = q0.y; bx.x2 = q0.z; bx.y2 = q0.w;
    } else {
      bx = xywh_to_xyxy(q0.x, q0.y, q0.z, q0.w);
    }
    const float rank = (p.flavour == FVB_NMS_DEMO) ? conf : best;
    bool ok = true;
    if (p.flavour == FVB_NMS_DEMO_BATCH) ok = best > p.conf_thr;  // nms.py:80 second filter on the score
    if (p.flavour != FVB_NMS_LIB) {
      const float gap = (float)bidx * p.max_wh;  // nms.py:44-45, fp32
      bx.x1 += gap; bx.y1 += gap; bx.x2 += gap; bx.y2 += gap;
    }
    sbox[i] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
    const uint32_t hi = ok ? desc_key(rank) : 0xffffffffu;
    keys0[i] = ((unsigned long long)hi << 32) | (uint32_t)i;
    valid_local += ok ? 1 : 0;
  }
  valid_local = (int)warp_sum((float)valid_local);
  if (lane == 0 && valid_local) atomicAdd(&misc[0], valid_local);
  __syncthreads();
  const int n_use = min(misc[0], p.max_nms);

  // ---- phase 2 + 3 ---------------------------------------------------------------------------------------
  if (p.trace && threadIdx.x == 0) p.trace[b * 8 + 2] = gtime_ns();
  unsigned long long* sorted = keys0;
  if (n <= kCapS)
    block_bitonic_sort64<NT>(keys0, n);
  else
    sorted = block_radix_sort_hi32<NT>(keys0, keys1, n, cnt, warp_tot);
  if (p.trace && threadIdx.x == 0) p.trace[b * 8 + 3] = gtime_ns();
  int kept = block_greedy_nms<NT>(sorted, n_use, sbox, p.iou_thr, p.max_det, kbox, karea, kslot, gs);
  if (p.trace && threadIdx.x == 0) { p.trace[b * 8 + 4] = gtime_ns(); p.trace[b * 8 + 6] = n; p.trace[b * 8 + 7] = kept; }

  // ---- phase 4: padded outputs -------------------------------------------------------------------------------
  for (int i = threadIdx.x; i < kept; i += kNmsThreads) {
    int slot = kslot[i];
    int r = srow[slot];
    const float4 q0 = rec4[(size_t)r * 2], q1 = rec4[(size_t)r * 2 + 1];  // score / class come back from the record
    Box bx;
    if (p.flavour == FVB_NMS_DEMO) {
      bx.x1 = q0.x; bx.y1 = q0.y; bx.x2 = q0.z; bx.y2 = q0.w;
    } else {
      bx = xywh_to_xyxy(q0.x, q0.y, q0.z, q0.w);
    }
    size_t o = (size_t)b * p.max_det + i;
    reinterpret_cast<float4*>(p.out_boxes)[o] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
    p.out_scores[o] = (p.flavour == FVB_NMS_DEMO) ? q1.x : q1.y;
    p.out_cls[o] = (long long)__float_as_int(q1.z);
    if (p.out_rows) p.out_rows[o] = r;
  }
  if (threadIdx.x == 0) p.out_cnt[b] = kept;
  if (p.trace && threadIdx.x == 0) p.trace[b * 8 + 5] = gtime_ns();
}

// One CTA per image.  PDL = launched as a programmatic dependent of decode_kernel (which has executed
// griddepcontrol.launch_dependents in every CTA, i.e. all of its CTAs are resident and make progress on their own): the CTA
// waits until its image's tiles are all published -- one thread polls with an acquire load, the others sleep in the
// barrier -- and then runs the image's NMS while the decode kernel is still streaming the later images.  Only the images
// decoded last are left when the decode kernel exits, so the step's tail is one image's NMS latency instead of a whole wave's.
// NT = 512: 40 registers, so that a CTA fits beside a decode CTA (and two fit an SM).  NT = 1024 (plain launches of at most one
// image per SM -- small batches, e.g. a data-parallel shard): twice the warps on the SM the image has to itself anyway; the
// greedy loop is latency-bound (~0.4 instructions per cycle per scheduler with 4 warps each), so its chunks go ~1.6x faster.
template <bool PDL, int NT>
__global__ void __maxnreg__(NT == 512 ? 40 : 64) yolo_nms_kernel(const YoloNmsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x;
  bool ok = true;
  if (p.trace && threadIdx.x == 0) p.trace[p.B * 8 + b] = gtime_ns();  // CTA entry (before the wait for the image)
  if (PDL) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
      const long long t0 = gtime_ns();
      int good = 1;
      // relaxed polls (an acquire load per poll would also invalidate this SM's L1 every time), one fence when the count is there
      while (ld_relaxed_gpu(&p.tile_sync[b]) < p.tiles_per_image) {
        __nanosleep(512);
        if (gtime_ns() - t0 > 4000000000ll) {  // 4 s: the producer is not running (misuse) -- report instead of hanging the GPU
          good = 0;
          break;
        }
      }
      __threadfence();      // acquire side of decode_kernel's publish_tiles()
      p.tile_sync[b] = 0u;  // re-armed for the next decode launch (stream-ordered after this grid)
      s_ok = good;
    }
    __syncthreads();  // (cumulativity: thread 0's acquire + this barrier order every thread's reads after the decode's writes)
    ok = s_ok != 0;
  }
  if (ok) yolo_nms_image<NT>(p, b, smem);
  else if (threadIdx.x == 0) p.out_cnt[b] = -1;
  if (PDL) {
    // The grid after this one in the stream must not start before the DECODE grid has fully exited (it re-arms its tile queue
    // last): the last CTA to leave waits for the prerequisite grid's completion, so this grid's completion implies it.
    __syncthreads();
    if (threadIdx.x == 0) {
      if (atomicAdd(&p.tile_sync[p.B], 1u) == (unsigned)p.B - 1u) {
        p.tile_sync[p.B] = 0u;
        asm volatile("griddepcontrol.wait;" ::: "memory");
      }
    }
  }
}

// ---- stand-alone scoring: candidate bitmap + records straight from a decoded [B,N,K] tensor ------------------
// (used when NMS is called without the fused decode outputs).  One warp per 32 consecutive rows: lanes
// read the objectness of "their" row, one ballot gives the bitmap word (plain store: row groups are
// word-aligned here), then the warp walks the set bits and reads each candidate row coalesced.
struct ScoreParams {
  const float* results;
  int N, K;
  float conf_thr;
  uint32_t* bitmap;
  int words;
  float* rec;
};

__global__ void __launch_bounds__(256) yolo_score_kernel(const ScoreParams p) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int w0 = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w0 >= p.words) return;
  const float* res = p.results + (size_t)b * p.N * p.K;
  const int r = w0 * 32 + lane;
  const bool c = (r < p.N) && (res[(size_t)r * p.K + 4] > p.conf_thr);  // NMS.py:7
  unsigned m = __ballot_sync(0xffffffffu, c);
  if (lane == 0) p.bitmap[(size_t)b * p.words + w0] = m;
  while (m) {
    const int bit = __ffs(m) - 1;
    m &= m - 1;
    const int row_i = w0 * 32 + bit;
    const float* row = res + (size_t)row_i * p.K;
    const float first = lane < p.K ? row[lane] : 0.0f;
    const float conf = __shfl_sync(0xffffffffu, first, 4);
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int ch = lane; ch < p.K; ch += 32) {
      if (ch >= 5) {
        const float v = (ch < 32 ? first : row[ch]) * conf;  // NMS.py:13
        if (v > best) {
          best = v;
          bidx = ch - 5;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) {  // torch.max keeps the first maximum (NMS.py:16)
        best = ob;
        bidx = oi;
      }
    }
    if (lane < 7) {
      const float val = lane < 5 ? first : (lane == 5 ? best : __int_as_float(bidx));
      p.rec[((size_t)b * p.N + row_i) * 8 + lane] = val;
    }
  }
}

// ---- segmented NMS: the torchvision.ops.nms equivalent, one CTA per segment ---------------------------------
struct SegNmsParams {
  const float* boxes;
  const float* scores;
  const int* seg_off;
  float iou_thr;
  int max_keep;
  int* keep_idx;
  int* keep_cnt;
  unsigned long long* ws_keys;  // [2 * total]
  float4* ws_box;               // [total]
  // kept list of long keeps (max_keep > kKeepSmem): global memory, segment s owns entries [seg_off[s], seg_off[s+1])
  float4* ws_kbox;              // [total]
  float* ws_karea;              // [total]
  int* ws_kslot;                // [total]
};

constexpr int kKeepSmem = 4096;  // kept entries (24 B each) held in shared memory; longer keeps live in the workspace

__global__ void __launch_bounds__(kNmsThreads) seg_nms_kernel(const SegNmsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const bool keep_in_smem = p.max_keep <= kKeepSmem;
  const NmsSmemLayout L = nms_layout(kCapS, keep_in_smem ? p.max_keep : 0);
  const int s = blockIdx.x;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + L.cnt);
  uint32_t* warp_tot = reinterpret_cast<uint32_t*>(smem + L.warp_tot);
  GreedyShared* gs = reinterpret_cast<GreedyShared*>(smem + L.gs);
  const int beg = p.seg_off[s], n = p.seg_off[s + 1] - beg;
  // a segment never keeps more than its own n boxes, so its slice [beg, beg + n) of the workspace arrays always suffices
  float4* kbox = keep_in_smem ? reinterpret_cast<float4*>(smem + L.kbox) : p.ws_kbox + beg;
  float* karea = keep_in_smem ? reinterpret_cast<float*>(smem + L.karea) : p.ws_karea + beg;
  int* kslot = keep_in_smem ? reinterpret_cast<int*>(smem + L.kslot) : p.ws_kslot + beg;
  if (n <= 0) {
    if (threadIdx.x == 0) p.keep_cnt[s] = 0;
    return;
  }
  unsigned long long *keys0, *keys1;
  float4* sbox;
  if (n <= kCapS) {
    keys0 = reinterpret_cast<unsigned long long*>(smem + L.keys0);
    keys1 = nullptr;
    sbox = reinterpret_cast<float4*>(smem + L.box);
  } else {
    keys0 = p.ws_keys + (size_t)2 * beg;
    keys1 = keys0 + n;
    sbox = p.ws_box + beg;
  }
  for (int i = threadIdx.x; i < n; i += kNmsThreads) {
    const float* bp = p.boxes + (size_t)(beg + i) * 4;
    sbox[i] = make_float4(bp[0], bp[1], bp[2], bp[3]);
    keys0[i] = ((unsigned long long)desc_key(p.scores[beg + i]) << 32) | (uint32_t)i;
  }
  __syncthreads();
  unsigned long long* sorted = keys0;
  if (n <= kCapS)
    block_bitonic_sort64(keys0, n);
  else
    sorted = block_radix_sort_hi32(keys0, keys1, n, cnt, warp_tot);
  int kept = block_greedy_nms(sorted, n, sbox, p.iou_thr, p.max_keep, kbox, karea, kslot, gs);
  for (int i = threadIdx.x; i < kept; i += kNmsThreads) p.keep_idx[(size_t)s * p.max_keep + i] = kslot[i];
  if (threadIdx.x == 0) p.keep_cnt[s] = kept;
}

// fp32 threshold t such that (x > t) in fp32  <=>  ((double)x > thr) -- torchvision's CPU op compares
// the fp32 ratio against the double threshold.
static float thr_round_down(double thr) {
  float t = (float)thr;
  if ((double)t > thr) t = nextafterf(t, -INFINITY);
  return t;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int ensure_smem(const void* fn, size_t bytes, const char* what) {
  if (bytes > 227 * 1024) {
    set_error("%s: %zu bytes of shared memory needed (max_det/max_keep too large)", what, bytes);
    return FVB_E_LIMIT;
  }
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu): %s", what, bytes, cudaGetErrorString(e));
    return FVB_E_CUDA;
  }
  return FVB_OK;
}

}  // namespace fvb

using namespace fvb;

// Debug hook (tools/nms_trace.py, tools/overlap_probe.py; declared in the "debug hooks" section of fvb200.h): device buffer of
// 9*B + 2 int64 -- [B][8] globaltimer stamps per phase written by yolo_nms_kernel, [B] CTA entry stamps, then the decode
// kernel's earliest CTA start / latest CTA end (atomicMin / atomicMax: preset to INT64_MAX / 0).  Process-global and NOT thread-safe by design -- the one piece of mutable state outside the
// thread-local error string; NULL (the default) disables it and production code never sets it.
static long long* g_nms_trace = nullptr;
extern "C" void fvb_debug_set_nms_trace(void* d_buf) { g_nms_trace = (long long*)d_buf; }
namespace fvb {
long long* debug_trace_ptr() { return g_nms_trace; }
}

extern "C" size_t fvb_yolo_nms_workspace_bytes(int batch, int rows_per_image) {
  size_t bn = (size_t)batch * (size_t)rows_per_image;
  size_t words = ((size_t)rows_per_image + 31) / 32;
  size_t o = 0;
  o = align_up(o + bn * 2 * 8, 256);            // keys
  o = align_up(o + bn * 16, 256);               // boxes
  o = align_up(o + bn * 4, 256);                // row
  o = align_up(o + (size_t)batch * words * 4, 256);  // private bitmap (stand-alone use)
  o = align_up(o + bn * 32, 256);                    // private candidate records (stand-alone use)
  return o + 256;
}

extern "C" int fvb_yolo_nms_f32(const float* d_results, int batch, int rows_per_image, int channels, float conf_thr,
                                double iou_thr, int max_det, int flavour, float max_wh, uint32_t* d_cand_bitmap,
                                const float* d_cand_rec, int clear_bitmap, float* d_out_boxes, float* d_out_scores, int64_t* d_out_cls,
                                int32_t* d_out_rows, int32_t* d_out_cnt, void* d_ws, void* stream) {
  return fvb_yolo_nms_after_decode_f32(d_results, batch, rows_per_image, channels, conf_thr, iou_thr, max_det, flavour, max_wh,
                                       d_cand_bitmap, d_cand_rec, clear_bitmap, d_out_boxes, d_out_scores, d_out_cls, d_out_rows,
                                       d_out_cnt, nullptr, 0, d_ws, stream);
}

extern "C" int fvb_yolo_nms_after_decode_f32(const float* d_results, int batch, int rows_per_image, int channels, float conf_thr,
                                             double iou_thr, int max_det, int flavour, float max_wh, uint32_t* d_cand_bitmap,
                                             const float* d_cand_rec, int clear_bitmap, float* d_out_boxes, float* d_out_scores,
                                             int64_t* d_out_cls, int32_t* d_out_rows, int32_t* d_out_cnt, uint32_t* d_tile_sync,
                                             int tiles_per_image, void* d_ws, void* stream) {
  FVB_REQUIRE(batch >= 0 && batch <= 65535 && rows_per_image >= 1 && channels >= 6, "yolo_nms: bad shape B=%d N=%d K=%d", batch, rows_per_image, channels);
  FVB_REQUIRE(max_det >= 1, "yolo_nms: max_det=%d", max_det);
  FVB_REQUIRE(flavour >= FVB_NMS_LIB && flavour <= FVB_NMS_DEMO_BATCH, "yolo_nms: unknown flavour %d", flavour);
  FVB_REQUIRE(d_results && d_out_boxes && d_out_scores && d_out_cls && d_out_cnt && d_ws, "yolo_nms: NULL pointer");
  FVB_REQUIRE(((uintptr_t)d_out_boxes & 15) == 0 && ((uintptr_t)d_ws & 255) == 0, "yolo_nms: out_boxes must be 16-byte and workspace 256-byte aligned");
  if (batch == 0) return FVB_OK;
  YoloNmsParams p;
  p.results = d_results;
  p.B = batch;
  p.N = rows_per_image;
  p.K = channels;
  p.conf_thr = conf_thr;
  p.iou_thr = thr_round_down(iou_thr);
  p.max_det = max_det;
  p.flavour = flavour;
  p.max_wh = max_wh;
  p.max_nms = 30000;  // demos/yolov3_u/utils/nms.py:16
  p.words = (rows_per_image + 31) / 32;
  size_t bn = (size_t)batch * (size_t)rows_per_image;
  unsigned char* w = (unsigned char*)d_ws;
  size_t o = 0;
  p.ws_keys = (unsigned long long*)(w + o); o = align_up(o + bn * 2 * 8, 256);
  p.ws_box = (float4*)(w + o);              o = align_up(o + bn * 16, 256);
  p.ws_row = (int*)(w + o);                 o = align_up(o + bn * 4, 256);
  FVB_REQUIRE((d_cand_bitmap == nullptr) == (d_cand_rec == nullptr), "yolo_nms: pass both the candidate bitmap and the records, or neither");
  FVB_REQUIRE(d_tile_sync == nullptr || (d_cand_bitmap != nullptr && tiles_per_image >= 1 && ((uintptr_t)d_tile_sync & 3) == 0),
              "yolo_nms_after_decode: needs the decode's candidate bitmap/records and tiles_per_image >= 1");
  cudaStream_t cs = (cudaStream_t)stream;
  if (d_cand_bitmap) {
    p.bitmap = d_cand_bitmap;
    p.rec = d_cand_rec;
    p.clear_bitmap = clear_bitmap & FVB_NMS_CLEAR_BITMAP;
    FVB_REQUIRE(((uintptr_t)d_cand_rec & 15) == 0, "yolo_nms: candidate records must be 16-byte aligned");
  } else {
    p.bitmap = (uint32_t*)(w + o);
    o = align_up(o + (size_t)batch * p.words * 4, 256);
    float* rec = (float*)(w + o);
    p.rec = rec;
    p.clear_bitmap = 0;
    ScoreParams sp;
    sp.results = d_results;
    sp.N = rows_per_image;
    sp.K = channels;
    sp.conf_thr = conf_thr;
    sp.bitmap = p.bitmap;
    sp.words = p.words;
    sp.rec = rec;
    dim3 grid((unsigned)((p.words + 7) / 8), (unsigned)batch);
    yolo_score_kernel<<<grid, 256, 0, cs>>>(sp);
    count_launch();
  }
  p.out_boxes = d_out_boxes;
  p.out_scores = d_out_scores;
  p.out_cls = (long long*)d_out_cls;
  p.out_rows = d_out_rows;
  p.out_cnt = d_out_cnt;
  p.trace = g_nms_trace;
  p.tile_sync = d_tile_sync;
  p.tiles_per_image = (unsigned)tiles_per_image;
  NmsSmemLayout L = nms_layout(kCapS, max_det);
  if (d_tile_sync == nullptr) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
    if (batch <= sms) {  // every image has an SM to itself: 1024 threads per image
      const NmsSmemLayout L2 = nms_layout(kCapS, max_det, 32);
      int rc = ensure_smem((const void*)yolo_nms_kernel<false, 1024>, L2.total, "yolo_nms");
      if (rc != FVB_OK) return rc;
      yolo_nms_kernel<false, 1024><<<batch, 1024, L2.total, cs>>>(p);
      count_launch();
      return check_launch("yolo_nms_kernel<1024>");
    }
    int rc = ensure_smem((const void*)yolo_nms_kernel<false, 512>, L.total, "yolo_nms");
    if (rc != FVB_OK) return rc;
    yolo_nms_kernel<false, 512><<<batch, kNmsThreads, L.total, cs>>>(p);
    count_launch();
    return check_launch("yolo_nms_kernel");
  }
  // FVB_NMS_WIDE_CTA: the caller says no NMS CTA fits beside a decode CTA anyway (narrow rows) -- with at most one image per SM the
  // CTAs then take 1024 threads and simply start, SM by SM, as the decode CTAs leave (no launch gap after the decode grid)
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
  const bool big = (clear_bitmap & FVB_NMS_WIDE_CTA) != 0 && batch <= sms;
  const NmsSmemLayout L2 = nms_layout(kCapS, max_det, 32);
  int rc = big ? ensure_smem((const void*)yolo_nms_kernel<true, 1024>, L2.total, "yolo_nms_after_decode")
               : ensure_smem((const void*)yolo_nms_kernel<true, 512>, L.total, "yolo_nms_after_decode");
  if (rc != FVB_OK) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)batch);
  cfg.blockDim = dim3(big ? 1024 : kNmsThreads);
  cfg.dynamicSmemBytes = big ? L2.total : L.total;
  cfg.stream = cs;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = big ? cudaLaunchKernelEx(&cfg, yolo_nms_kernel<true, 1024>, p) : cudaLaunchKernelEx(&cfg, yolo_nms_kernel<true, 512>, p);
  if (e != cudaSuccess) {
    set_error("yolo_nms_after_decode: launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return FVB_E_CUDA;
  }
  count_launch();
  return check_launch("yolo_nms_kernel<pdl>");
}

extern "C" size_t fvb_nms_segmented_workspace_bytes(int64_t total_boxes, int segments) {
  (void)segments;
  const size_t t = (size_t)total_boxes;
  return align_up(t * 16, 256) /* keys x2 */ + align_up(t * 16, 256) /* boxes */ + align_up(t * 16, 256) /* kept boxes */ +
         2 * align_up(t * 4, 256) /* kept areas, slots */ + 256;
}

extern "C" int fvb_nms_segmented_f32(const float* d_boxes, const float* d_scores, const int32_t* d_seg_offsets,
                                     int segments, int64_t total_boxes, double iou_thr, int max_keep,
                                     int32_t* d_keep_idx, int32_t* d_keep_cnt, void* d_ws, void* stream) {
  FVB_REQUIRE(segments >= 0 && total_boxes >= 0 && max_keep >= 1, "nms_segmented: bad sizes");
  FVB_REQUIRE(d_seg_offsets && d_keep_idx && d_keep_cnt && d_ws, "nms_segmented: NULL pointer");
  FVB_REQUIRE(total_boxes == 0 || (d_boxes && d_scores), "nms_segmented: NULL boxes/scores");
  FVB_REQUIRE(((uintptr_t)d_ws & 255) == 0, "nms_segmented: workspace must be 256-byte aligned");
  if (segments == 0) return FVB_OK;
  SegNmsParams p;
  p.boxes = d_boxes;
  p.scores = d_scores;
  p.seg_off = d_seg_offsets;
  p.iou_thr = thr_round_down(iou_thr);
  p.max_keep = max_keep;
  p.keep_idx = d_keep_idx;
  p.keep_cnt = d_keep_cnt;
  {
    unsigned char* w = (unsigned char*)d_ws;
    const size_t t = (size_t)total_boxes;
    p.ws_keys = (unsigned long long*)w;  w += align_up(t * 16, 256);
    p.ws_box = (float4*)w;               w += align_up(t * 16, 256);
    p.ws_kbox = (float4*)w;              w += align_up(t * 16, 256);
    p.ws_karea = (float*)w;              w += align_up(t * 4, 256);
    p.ws_kslot = (int*)w;
  }
  NmsSmemLayout L = nms_layout(kCapS, max_keep <= kKeepSmem ? max_keep : 0);
  int rc = ensure_smem((const void*)seg_nms_kernel, L.total, "nms_segmented");
  if (rc != FVB_OK) return rc;
  seg_nms_kernel<<<segments, kNmsThreads, L.total, (cudaStream_t)stream>>>(p);
  count_launch();
  return check_launch("seg_nms_kernel");
}
