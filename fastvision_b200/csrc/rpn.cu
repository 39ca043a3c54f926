// Faster R-CNN RPN proposal filter for a whole batch: 2 launches, no host sync.
//
// Replaces RPN.filter_proposals (demos/faster_rcnn/models/rpn.py:168-208) and what it calls:
// make_anchors_xywh :160-166, dxdydwdh2xywh :111-119, xywh2xyxy :131-137, the per-image Python loop of
// topk -> torchvision.ops.nms -> first post_n -> xyxy2xywh :139-145.
//   rpn_decode_kernel   one thread per anchor: reads its 16-byte regression and 8-byte class pair once,
//                       writes the clamped xyxy box and a 64-bit sort key (descending score, anchor index);
//   rpn_nms_kernel      one CTA per image: block radix sort of all H*W*A keys (the reference's topk is the
//                       first pre_n of that order), chunked greedy suppression over the first pre_n with the
//                       kept list (<= post_n boxes) in shared memory, xywh outputs + count.
// Ties: equal scores rank by lower anchor index (torch.topk leaves tie order unspecified).
#include "nms.cuh"

#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace fvb {

struct RpnParams {
  const float* cls;  // [B,H,W,A,2]
  const float* reg;  // [B,H,W,A,4]
  int B, H, W, A, n;  // n = H*W*A
  float aw[FVB_MAX_ANCHORS], ah[FVB_MAX_ANCHORS];
  int pre_n, post_n;  // post_n already clipped to n
  int out_pitch;      // rows per image of the padded outputs (the caller's post_n)
  float iou_thr;
  unsigned long long* keys;  // [B][2][n]
  float4* box;               // [B][n] clamped xyxy, feature units
  float* out_xywh;           // [B][post_n][4]
  int* out_idx;              // [B][post_n] anchor index (h,w,a flat) of every proposal, or NULL
  int* out_cnt;              // [B]
};

__global__ void __launch_bounds__(256) rpn_decode_kernel(const RpnParams p) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= p.n) return;
  const int cell = i / p.A, a = i - cell * p.A;
  const int y = cell / p.W, x = cell - y * p.W;
  const size_t g = (size_t)b * p.n + i;
  const float4 d = reinterpret_cast<const float4*>(p.reg)[g];
  const float2 c = reinterpret_cast<const float2*>(p.cls)[g];
  const float aw = p.aw[a], ah = p.ah[a];
  // rpn.py:114-117 (h uses exp(dw), sic)
  const float bx = d.x * aw + (float)x;
  const float by = d.y * ah + (float)y;
  const float e = expf(d.z);
  const float bw = e * aw, bh = e * ah;
  // softmax over the pair, torch op order: subtract the max, exp, normalise (rpn.py:173-175)
  const float m = fmaxf(c.x, c.y);
  const float e0 = expf(c.x - m), e1 = expf(c.y - m);
  const float score = e1 / (e0 + e1);
  // rpn.py:181-185: xyxy, clamped to the feature map
  const float hw = bw / 2.0f, hh = bh / 2.0f;
  const float xmax = (float)(p.W - 1), ymax = (float)(p.H - 1);
  float4 o;
  o.x = fminf(fmaxf(bx - hw, 0.0f), xmax);
  o.y = fminf(fmaxf(by - hh, 0.0f), ymax);
  o.z = fminf(fmaxf(bx + hw, 0.0f), xmax);
  o.w = fminf(fmaxf(by + hh, 0.0f), ymax);
  p.box[g] = o;
  p.keys[(size_t)b * 2 * p.n + i] = ((unsigned long long)desc_key(score) << 32) | (uint32_t)i;
}

struct RpnSmemLayout {
  size_t cnt, warp_tot, kbox, karea, kslot, gs, total;
};

__host__ __device__ inline RpnSmemLayout rpn_layout(int max_keep) {
  RpnSmemLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    o = (o + 15) / 16 * 16;
    size_t r = o;
    o += bytes;
    return r;
  };
  L.cnt = take((size_t)kNmsWarps * 256 * 4);
  L.warp_tot = take((size_t)(kNmsWarps + 1) * 4);
  L.kbox = take((size_t)max_keep * 16);
  L.karea = take((size_t)max_keep * 4);
  L.kslot = take((size_t)max_keep * 4);
  L.gs = take(sizeof(GreedyShared));
  L.total = o;
  return L;
}

__global__ void __launch_bounds__(kNmsThreads) rpn_nms_kernel(const RpnParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const RpnSmemLayout L = rpn_layout(p.post_n);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + L.cnt);
  uint32_t* warp_tot = reinterpret_cast<uint32_t*>(smem + L.warp_tot);
  float4* kbox = reinterpret_cast<float4*>(smem + L.kbox);
  float* karea = reinterpret_cast<float*>(smem + L.karea);
  int* kslot = reinterpret_cast<int*>(smem + L.kslot);
  GreedyShared* gs = reinterpret_cast<GreedyShared*>(smem + L.gs);
  const int b = blockIdx.x;
  unsigned long long* keys0 = p.keys + (size_t)b * 2 * p.n;
  const float4* box = p.box + (size_t)b * p.n;
  unsigned long long* sorted = block_radix_sort_hi32(keys0, keys0 + p.n, p.n, cnt, warp_tot);
  const int n_use = min(p.pre_n, p.n);  // rpn.py:193-195 topk
  const int kept = block_greedy_nms(sorted, n_use, box, p.iou_thr, p.post_n, kbox, karea, kslot, gs);
  for (int i = threadIdx.x; i < kept; i += kNmsThreads) {
    const float4 q = kbox[i];
    // rpn.py:139-145 xyxy2xywh
    const float4 o = make_float4((q.x + q.z) / 2.0f, (q.y + q.w) / 2.0f, q.z - q.x, q.w - q.y);
    reinterpret_cast<float4*>(p.out_xywh)[(size_t)b * p.out_pitch + i] = o;
    if (p.out_idx) p.out_idx[(size_t)b * p.out_pitch + i] = kslot[i];
  }
  if (threadIdx.x == 0) p.out_cnt[b] = kept;
}

// ---- thread-block-cluster variant: two CTAs (two SMs) per image ---------------------------------------------------------------
// With pre_n = 12000 / post_n = 2000 the greedy pass is ~12 M pair tests per image, fp32-issue-bound on the ONE SM the image's
// CTA runs on, and a batch of 64 images leaves 84 of the 148 SMs idle.  Here a cluster of kRpnCluster CTAs shares an image: the
// kept list is dealt round-robin over the CTAs (kept i lives in CTA i % C), every CTA tests the 64-candidate chunk against
// ITS slice only, the partial "still alive" masks are exchanged through distributed shared memory (each CTA stores its 64-bit
// partial into every peer's exchange slot, one hardware cluster barrier per chunk) and AND-ed; the cheap intra-chunk mask and
// the serial resolve are replicated, so all CTAs take identical decisions and append identical kept indices (each keeps its
// own share).  The keep set is exactly the sequential algorithm's.
namespace cg = cooperative_groups;
constexpr int kRpnCluster = 2;

struct RpnClusterSmem {
  size_t cnt, warp_tot, kbox, karea, kslot, gs, xalive, xtot, dig_base, total;
};
__host__ __device__ inline RpnClusterSmem rpn_cluster_layout(int max_keep) {
  RpnClusterSmem L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    o = (o + 15) / 16 * 16;
    size_t r = o;
    o += bytes;
    return r;
  };
  const int local = (max_keep + kRpnCluster - 1) / kRpnCluster + 1;
  L.cnt = take((size_t)kNmsWarps * 256 * 4);
  L.warp_tot = take((size_t)(kNmsWarps + 1) * 4);
  L.kbox = take((size_t)local * 16);
  L.karea = take((size_t)local * 4);
  L.kslot = take((size_t)local * 4);
  L.gs = take(sizeof(GreedyShared));
  L.xalive = take((size_t)2 * kRpnCluster * 2 * 4);  // [chunk parity][source rank][word]
  L.xtot = take((size_t)kRpnCluster * 256 * 4);       // per-digit totals of every CTA (cluster sort)
  L.dig_base = take((size_t)256 * 4);
  L.total = o;
  return L;
}

// Stable LSD radix sort of the image's n keys by their high 32 bits, split over the CTAs of the cluster: CTA c counts and
// scatters the c-th contiguous share of the keys (so, for equal digits, CTA 0's keys stay in front of CTA 1's), the
// per-digit totals travel through distributed shared memory, one cluster barrier after the histograms and one after the
// scatter.  xtot: [kRpnCluster][256] u32 in every CTA (slot r written by CTA r).  Returns the buffer holding the result.
__device__ inline unsigned long long* cluster_radix_sort_hi32(cg::cluster_group& cluster, int rank, unsigned long long* src,
                                                              unsigned long long* dst, int n, uint32_t* cnt, uint32_t* warp_tot,
                                                              uint32_t* xtot, uint32_t* dig_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int share = (((n + kRpnCluster - 1) / kRpnCluster) + 31) & ~31;
  const int cbeg = min(n, rank * share), cend = min(n, cbeg + share);
  const int chunk = ((((cend - cbeg) + kNmsWarps - 1) / kNmsWarps) + 31) & ~31;
  const int beg = min(cend, cbeg + warp * chunk), end = min(cend, beg + chunk);
  uint32_t* peer_tot[kRpnCluster];
#pragma unroll
  for (int r = 0; r < kRpnCluster; ++r) peer_tot[r] = cluster.map_shared_rank(xtot, r);
  for (int shift = 32; shift < 64; shift += 8) {
    for (int i = threadIdx.x; i < kNmsWarps * 256; i += kNmsThreads) cnt[i] = 0;
    __syncthreads();
    for (int base = beg; base < end; base += 128) {
      unsigned long long k4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 32 * u + lane;
        k4[u] = i < end ? src[i] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 32 * u + lane;
        const bool act = i < end;
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (act) {
          const uint32_t d = (uint32_t)(k4[u] >> shift) & 255u;
          const unsigned peers = __match_any_sync(m, d);
          if (lane == __ffs(peers) - 1) cnt[d * kNmsWarps + warp] += __popc(peers);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    {
      constexpr int PER = kNmsWarps * 256 / kNmsThreads;  // 8
      uint32_t loc[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        loc[k] = cnt[threadIdx.x * PER + k];
        sum += loc[k];
      }
      uint32_t run = block_exclusive_scan(sum, warp_tot);
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        cnt[threadIdx.x * PER + k] = run;
        run += loc[k];
      }
    }
    __syncthreads();
    // cnt[d*W + w] = my keys with a smaller digit, or digit d in an earlier warp.  My total of digit d:
    if (threadIdx.x < 256) {
      const int d = threadIdx.x;
      const uint32_t nxt = d < 255 ? cnt[(d + 1) * kNmsWarps] : (uint32_t)(cend - cbeg);
      const uint32_t tot = nxt - cnt[d * kNmsWarps];
#pragma unroll
      for (int r = 0; r < kRpnCluster; ++r) peer_tot[r][rank * 256 + d] = tot;
    }
    cluster.sync();
    // global start of digit d = keys of all CTAs with a smaller digit (+ digit d of the CTAs before me)
    {
      uint32_t all = 0, before = 0;
      if (threadIdx.x < 256) {
#pragma unroll
        for (int r = 0; r < kRpnCluster; ++r) {
          const uint32_t t = xtot[r * 256 + threadIdx.x];
          all += t;
          if (r < rank) before += t;
        }
      }
      const uint32_t excl = block_exclusive_scan(all, warp_tot);
      if (threadIdx.x < 256) dig_base[threadIdx.x] = excl + before - cnt[threadIdx.x * kNmsWarps];
    }
    __syncthreads();
    for (int base = beg; base < end; base += 128) {
      unsigned long long k4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 32 * u + lane;
        k4[u] = i < end ? src[i] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 32 * u + lane;
        const bool act = i < end;
        const unsigned m = __ballot_sync(0xffffffffu, act);
        uint32_t d = 0;
        unsigned peers = 0;
        if (act) {
          const unsigned long long key = k4[u];
          d = (uint32_t)(key >> shift) & 255u;
          peers = __match_any_sync(m, d);
          const uint32_t pos = dig_base[d] + cnt[d * kNmsWarps + warp] + __popc(peers & lt_mask);
          dst[pos] = key;
        }
        __syncwarp();
        if (act && lane == __ffs(peers) - 1) cnt[d * kNmsWarps + warp] += __popc(peers);
        __syncwarp();
      }
    }
    __threadfence();
    cluster.sync();  // dst is complete (both shares) before anybody reads it as the next pass's source
    unsigned long long* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

__global__ void __cluster_dims__(kRpnCluster, 1, 1) __launch_bounds__(kNmsThreads) rpn_nms_cluster_kernel(const RpnParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const RpnClusterSmem L = rpn_cluster_layout(p.post_n);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + L.cnt);
  uint32_t* warp_tot = reinterpret_cast<uint32_t*>(smem + L.warp_tot);
  float4* kbox = reinterpret_cast<float4*>(smem + L.kbox);      // my slice: kept i with i % C == rank, at i / C
  float* karea = reinterpret_cast<float*>(smem + L.karea);
  int* kslot = reinterpret_cast<int*>(smem + L.kslot);
  GreedyShared* gs = reinterpret_cast<GreedyShared*>(smem + L.gs);
  unsigned* xalive = reinterpret_cast<unsigned*>(smem + L.xalive);
  const int b = blockIdx.x / kRpnCluster;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  unsigned long long* keys0 = p.keys + (size_t)b * 2 * p.n;
  const float4* box = p.box + (size_t)b * p.n;
  uint32_t* xtot = reinterpret_cast<uint32_t*>(smem + L.xtot);
  uint32_t* dig_base = reinterpret_cast<uint32_t*>(smem + L.dig_base);
  const unsigned long long* sorted = cluster_radix_sort_hi32(cluster, rank, keys0, keys0 + p.n, p.n, cnt, warp_tot, xtot, dig_base);
  const int n = min(p.pre_n, p.n);  // rpn.py:193-195 topk
  const int max_keep = p.post_n;
  const float thr = p.iou_thr;
  const NmsFast fast = make_nms_fast(thr);
  unsigned* peer_x[kRpnCluster];
#pragma unroll
  for (int r = 0; r < kRpnCluster; ++r) peer_x[r] = cluster.map_shared_rank(xalive, r);

  if (tid == 0) {
    gs->kcount = 0;
    const int cn0 = min(kNmsChunk, n);
    gs->alive32[0] = cn0 >= 32 ? 0xffffffffu : ((1u << cn0) - 1u);
    gs->alive32[1] = cn0 >= 64 ? 0xffffffffu : (cn0 > 32 ? ((1u << (cn0 - 32)) - 1u) : 0u);
    gs->nz32[0] = gs->nz32[1] = 0u;
  }
  if (tid < kNmsChunk) stage_chunk(gs, 0, sorted, 0, n, box, tid);
  __syncthreads();
  int buf = 0;
  for (int c0 = 0; c0 < n; c0 += kNmsChunk, buf ^= 1) {
    const int kc = gs->kcount;  // replicated: identical in every CTA of the cluster
    if (kc >= max_keep) break;
    const int cn = min(kNmsChunk, n - c0);
    const int kl = (kc - rank + kRpnCluster - 1) / kRpnCluster;  // kept boxes in my slice
    // (a) the chunk against MY slice of the kept list
    {
      const int r = tid & (kNmsChunk - 1), s = tid / kNmsChunk;
      constexpr int S = kNmsThreads / kNmsChunk;
      if (r < cn && s < kl) {
        const float4 bj = gs->cbox[buf][r];
        const float aj = gs->carea[buf][r];
        bool hit = false;
        if (fast.ok) {
          bool amb = false;
          int i = s;
          for (; i + S < kl; i += 2 * S) {
            const float4 k0 = kbox[i], k1 = kbox[i + S];
            const float a0 = karea[i], a1 = karea[i + S];
            nms_overlap_fast(k0, a0, bj, aj, fast, hit, amb);
            nms_overlap_fast(k1, a1, bj, aj, fast, hit, amb);
          }
          if (i < kl) nms_overlap_fast(kbox[i], karea[i], bj, aj, fast, hit, amb);
          if (amb && !hit) {
            for (i = s; i < kl; i += S)
              if (nms_overlap(kbox[i], karea[i], bj, aj, thr)) {
                hit = true;
                break;
              }
          }
        } else {
          for (int i = s; i < kl; i += S)
            if (nms_overlap(kbox[i], karea[i], bj, aj, thr)) {
              hit = true;
              break;
            }
        }
        if (hit) atomicAnd(&gs->alive32[r >> 5], ~(1u << (r & 31)));
      }
    }
    // (b) intra-chunk mask (replicated in every CTA: 64 x 64 tests, small next to (a))
    for (int r = warp; r < cn; r += kNmsWarps) {
      const float4 br = gs->cbox[buf][r];
      const float ar = gs->carea[buf][r];
      unsigned long long row = 0ull;
#pragma unroll
      for (int h = 0; h < kNmsChunk / 32; ++h) {
        const int q = h * 32 + lane;
        bool hit = false;
        if (q > r && q < cn) {
          const float4 bq = gs->cbox[buf][q];
          const float aq = gs->carea[buf][q];
          if (fast.ok) {
            bool amb = false;
            nms_overlap_fast(br, ar, bq, aq, fast, hit, amb);
            if (amb && !hit) hit = nms_overlap(br, ar, bq, aq, thr);
          } else {
            hit = nms_overlap(br, ar, bq, aq, thr);
          }
        }
        const unsigned w = __ballot_sync(0xffffffffu, hit);
        row |= (unsigned long long)w << (32 * h);
      }
      if (lane == 0) {
        gs->mask[r] = row;
        if (row) atomicOr(&gs->nz32[r >> 5], 1u << (r & 31));
      }
    }
    __syncthreads();
    // exchange the partial alive masks: my two words go into slot [buf][rank] of every CTA of the cluster (DSMEM stores)
    if (tid < 2 * kRpnCluster) {
      const int dst = tid >> 1, w = tid & 1;
      peer_x[dst][(buf * kRpnCluster + rank) * 2 + w] = gs->alive32[w];
    }
    cluster.sync();
    // (c) resolve on warp 0 with the AND of all partials; warps 1-2 stage the next chunk
    if (warp == 0) {
      unsigned a0 = 0xffffffffu, a1 = 0xffffffffu;
#pragma unroll
      for (int r = 0; r < kRpnCluster; ++r) {
        a0 &= xalive[(buf * kRpnCluster + r) * 2 + 0];
        a1 &= xalive[(buf * kRpnCluster + r) * 2 + 1];
      }
      unsigned long long remaining = ((unsigned long long)a1 << 32) | a0;
      const unsigned long long nz = ((unsigned long long)gs->nz32[1] << 32) | gs->nz32[0];
      unsigned long long todo = remaining & nz;
      while (todo) {  // warp-uniform
        const int j = __ffsll((long long)todo) - 1;
        remaining &= ~gs->mask[j];
        todo = remaining & nz & ~((2ull << j) - 1ull);
      }
      const unsigned long long kept = remaining;
      const int room = max_keep - kc;
#pragma unroll
      for (int h = 0; h < kNmsChunk / 32; ++h) {
        const int q = h * 32 + lane;
        if ((kept >> q) & 1ull) {
          const int rk = __popcll(kept & ((1ull << q) - 1ull));
          const int gidx = kc + rk;  // position in the (virtual) global kept list
          if (rk < room && gidx % kRpnCluster == rank) {
            const int li = gidx / kRpnCluster;
            kbox[li] = gs->cbox[buf][q];
            karea[li] = gs->carea[buf][q];
            kslot[li] = gs->cslot[buf][q];
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        gs->kcount = kc + min(room, __popcll(kept));
        const int cn1 = min(kNmsChunk, max(0, n - c0 - kNmsChunk));
        gs->alive32[0] = cn1 >= 32 ? 0xffffffffu : ((1u << cn1) - 1u);
        gs->alive32[1] = cn1 >= 64 ? 0xffffffffu : (cn1 > 32 ? ((1u << (cn1 - 32)) - 1u) : 0u);
        gs->nz32[0] = gs->nz32[1] = 0u;
      }
    } else if (warp <= kNmsChunk / 32) {
      stage_chunk(gs, buf ^ 1, sorted, c0 + kNmsChunk, n, box, tid - 32);
    }
    __syncthreads();
  }
  const int kept = gs->kcount;
  for (int i = rank + kRpnCluster * tid; i < kept; i += kRpnCluster * kNmsThreads) {
    const int li = i / kRpnCluster;
    const float4 q = kbox[li];
    const float4 o = make_float4((q.x + q.z) / 2.0f, (q.y + q.w) / 2.0f, q.z - q.x, q.w - q.y);  // rpn.py:139-145
    reinterpret_cast<float4*>(p.out_xywh)[(size_t)b * p.out_pitch + i] = o;
    if (p.out_idx) p.out_idx[(size_t)b * p.out_pitch + i] = kslot[li];
  }
  if (rank == 0 && tid == 0) p.out_cnt[b] = kept;
  cluster.sync();  // no CTA may exit while a peer can still address its shared memory
}

static size_t rpn_align(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_rpn_workspace_bytes(int batch, int height, int width, int anchors) {
  if (batch < 0 || height < 0 || width < 0 || anchors < 0) return 0;
  const size_t bn = (size_t)batch * height * width * anchors;
  return rpn_align(bn * 2 * 8) + rpn_align(bn * 16) + 256;
}

extern "C" int fvb_rpn_proposals_f32(const float* d_cls, const float* d_reg, const float* base_anchors, int batch,
                                     int height, int width, int anchors, int pre_n, int post_n, double iou_thr,
                                     float* d_out_xywh, int32_t* d_out_idx, int32_t* d_out_cnt, void* d_ws,
                                     void* stream) {
  FVB_REQUIRE(batch >= 0 && batch <= 65535 && height >= 1 && width >= 1, "rpn_proposals: bad shape B=%d H=%d W=%d", batch, height, width);
  FVB_REQUIRE(anchors >= 1 && anchors <= FVB_MAX_ANCHORS, "rpn_proposals: anchors=%d out of range [1,%d]", anchors, FVB_MAX_ANCHORS);
  FVB_REQUIRE(pre_n >= 1 && post_n >= 1, "rpn_proposals: pre_n=%d post_n=%d", pre_n, post_n);
  FVB_REQUIRE(d_cls && d_reg && base_anchors && d_out_xywh && d_out_cnt && d_ws, "rpn_proposals: NULL pointer");
  FVB_REQUIRE(((uintptr_t)d_reg & 15) == 0 && ((uintptr_t)d_cls & 7) == 0 && ((uintptr_t)d_out_xywh & 15) == 0 && ((uintptr_t)d_ws & 255) == 0,
              "rpn_proposals: reg/out_xywh must be 16-byte, cls 8-byte and the workspace 256-byte aligned");
  const long long n = (long long)height * width * anchors;
  if (n >= (1ll << 31)) {
    set_error("rpn_proposals: %lld anchors per image", n);
    return FVB_E_LIMIT;
  }
  if (batch == 0) return FVB_OK;
  RpnParams p;
  p.cls = d_cls;
  p.reg = d_reg;
  p.B = batch;
  p.H = height;
  p.W = width;
  p.A = anchors;
  p.n = (int)n;
  for (int a = 0; a < anchors; ++a) {
    p.aw[a] = base_anchors[2 * a];
    p.ah[a] = base_anchors[2 * a + 1];
  }
  p.pre_n = pre_n;
  p.post_n = (int)(post_n < n ? post_n : n);  // rpn.py:201: never more than the survivors
  const float t = (float)iou_thr;
  p.iou_thr = ((double)t > iou_thr) ? nextafterf(t, -INFINITY) : t;  // fp32 ratio vs double threshold (nms.cu)
  const size_t bn = (size_t)batch * (size_t)n;
  p.keys = (unsigned long long*)d_ws;
  p.box = (float4*)((unsigned char*)d_ws + rpn_align(bn * 2 * 8));
  p.out_xywh = d_out_xywh;
  p.out_idx = d_out_idx;
  p.out_cnt = d_out_cnt;
  p.out_pitch = post_n;
  cudaStream_t cs = (cudaStream_t)stream;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)batch);
  rpn_decode_kernel<<<grid, 256, 0, cs>>>(p);
  count_launch();
  // images with many anchors (sort + greedy both long) run on a 2-CTA cluster per image; FVB_RPN_CLUSTER=0/1 overrides
  bool use_cluster = n >= 4096;
  if (const char* e = getenv("FVB_RPN_CLUSTER")) use_cluster = e[0] == '1';
  if (use_cluster) {
    RpnClusterSmem LC = rpn_cluster_layout(p.post_n);
    if (LC.total > 227 * 1024) {
      set_error("rpn_proposals: post_n=%d needs %zu bytes of shared memory for the kept list", post_n, LC.total);
      return FVB_E_LIMIT;
    }
    cudaError_t e = cudaFuncSetAttribute((const void*)rpn_nms_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LC.total);
    if (e != cudaSuccess) {
      set_error("rpn_proposals: cudaFuncSetAttribute(%zu): %s", LC.total, cudaGetErrorString(e));
      return FVB_E_CUDA;
    }
    rpn_nms_cluster_kernel<<<batch * kRpnCluster, kNmsThreads, LC.total, cs>>>(p);
    count_launch();
    return check_launch("rpn_nms_cluster_kernel");
  }
  RpnSmemLayout L = rpn_layout(p.post_n);
  if (L.total > 227 * 1024) {
    set_error("rpn_proposals: post_n=%d needs %zu bytes of shared memory for the kept list", post_n, L.total);
    return FVB_E_LIMIT;
  }
  cudaError_t e = cudaFuncSetAttribute((const void*)rpn_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) {
    set_error("rpn_proposals: cudaFuncSetAttribute(%zu): %s", L.total, cudaGetErrorString(e));
    return FVB_E_CUDA;
  }
  rpn_nms_kernel<<<batch, kNmsThreads, L.total, cs>>>(p);
  count_launch();
  return check_launch("rpn_nms_kernel");
}
