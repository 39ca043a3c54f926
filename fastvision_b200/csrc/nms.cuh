// Building blocks of the batched NMS kernels: block radix sort (LSD, 8-bit digits, warp-level
// multisplit with match.any) and the chunked greedy suppression with a kept list and warp-ballot
// bitmasks.  One CTA owns one image / segment; everything lives in shared memory when the
// candidate count fits, otherwise in a global workspace through the same (generic) pointers.
#pragma once
#include "common.cuh"

namespace fvb {

constexpr int kNmsThreads = 512;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kNmsChunk = 64;  // candidates resolved per greedy step

// float -> uint32 whose ascending order is the float's DESCENDING order (NaN-agnostic).
__device__ __forceinline__ uint32_t desc_key(float s) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
  return ~u;
}

// torchvision.ops.nms arithmetic: strict >, IEEE divide, NaN never suppresses.
// The decision is exactly `inter / ((area_a + area_b) - inter) > thr` in fp32, evaluated lazily:
//   * inter == 0 (the common, disjoint case): the ratio is +-0 or NaN, never > thr for thr >= 0 -- no divide
//     (a zero numerator also sends the IEEE-division sequence down its slow path);
//   * otherwise an approximate ratio (rcp.approx, a few ulp) decides unless it lies within 1e-5 (relative)
//     of the threshold, and only then is the exact IEEE division evaluated.
__device__ __forceinline__ bool nms_overlap(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  const float w = fmaxf(0.0f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
  const float h = fmaxf(0.0f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
  const float inter = w * h;
  if (thr >= 0.0f && !(inter > 0.0f)) return false;
  const float uni = (area_a + area_b) - inter;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(uni));
  const float q = inter * r;
  if (fabsf(q - thr) > 1e-5f * fmaxf(fabsf(thr), 1e-30f) && fabsf(q) < 1e30f && uni > 1e-30f && uni < 1e30f) return q > thr;
  return inter / uni > thr;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_tot /*[kNmsWarps+1]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = lane < kNmsWarps ? warp_tot[lane] : 0u;
    uint32_t ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    if (lane < kNmsWarps) warp_tot[lane] = ti - t;  // exclusive
    if (lane == kNmsWarps - 1) warp_tot[kNmsWarps] = ti;
  }
  __syncthreads();
  uint32_t r = warp_tot[warp] + inc - v;
  return r;
}

// Stable LSD radix sort of n 64-bit keys by their HIGH 32 bits, ascending.  Keys must enter in the
// order that should break ties (low word = slot index, already ascending).  Returns the buffer that
// holds the sorted keys.  cnt: kNmsWarps*256 u32, warp_tot: kNmsWarps+1 u32 (shared memory).
__device__ unsigned long long* block_radix_sort_hi32(unsigned long long* src, unsigned long long* dst, int n,
                                                     uint32_t* cnt, uint32_t* warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  int chunk = (((n + kNmsWarps - 1) / kNmsWarps) + 31) & ~31;
  int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  for (int shift = 32; shift < 64; shift += 8) {
    for (int i = threadIdx.x; i < kNmsWarps * 256; i += kNmsThreads) cnt[i] = 0;
    __syncthreads();
    for (int base = beg; base < end; base += 32) {
      int i = base + lane;
      bool act = i < end;
      unsigned m = __ballot_sync(0xffffffffu, act);
      if (act) {
        uint32_t d = (uint32_t)(src[i] >> shift) & 255u;
        unsigned peers = __match_any_sync(m, d);
        if (lane == __ffs(peers) - 1) cnt[d * kNmsWarps + warp] += __popc(peers);
      }
      __syncwarp();
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
    for (int base = beg; base < end; base += 32) {
      int i = base + lane;
      bool act = i < end;
      unsigned m = __ballot_sync(0xffffffffu, act);
      uint32_t d = 0;
      unsigned peers = 0;
      if (act) {
        unsigned long long key = src[i];
        d = (uint32_t)(key >> shift) & 255u;
        peers = __match_any_sync(m, d);
        uint32_t pos = cnt[d * kNmsWarps + warp] + __popc(peers & lt_mask);
        dst[pos] = key;
      }
      __syncwarp();
      if (act && lane == __ffs(peers) - 1) cnt[d * kNmsWarps + warp] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    unsigned long long* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

struct GreedyShared {
  unsigned long long mask[kNmsChunk];  // intra-chunk suppression rows (bit q: row suppresses column q)
  unsigned long long alive;            // columns of the current chunk not suppressed by the kept list
  int kcount;
};

// Greedy NMS over candidates already sorted by rank: sorted[p] low word = slot into box[] (xyxy, the
// boxes the IoU is taken on).  Walks the order in chunks of 64: (a) each candidate is tested against
// the kept list (warp per candidate, lanes over kept boxes, __any_sync), (b) the 64x64 intra-chunk
// mask is built with __ballot_sync, (c) warp 0 resolves the chunk serially on register-held rows.
// Identical keep set to the sequential algorithm: j is dropped iff an earlier KEPT i has IoU > thr.
// kbox/karea/kslot: kept list (max_keep entries, shared memory).  Returns the kept count (<= max_keep).
__device__ int block_greedy_nms(const unsigned long long* sorted, int n, const float4* box, float thr, int max_keep,
                                float4* kbox, float* karea, int* kslot, GreedyShared* gs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) gs->kcount = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += kNmsChunk) {
    const int kc = gs->kcount;
    if (kc >= max_keep) break;
    const int cn = min(kNmsChunk, n - c0);
    if (threadIdx.x == 0) gs->alive = 0ull;
    __syncthreads();
    // (a) against the kept list
    for (int r = warp; r < cn; r += kNmsWarps) {
      int slot = (int)(uint32_t)sorted[c0 + r];
      float4 bj = box[slot];
      float aj = (bj.z - bj.x) * (bj.w - bj.y);
      bool dead = false;
      for (int i0 = 0; i0 < kc && !dead; i0 += 32) {
        int i = i0 + lane;
        bool hit = (i < kc) && nms_overlap(kbox[i], karea[i], bj, aj, thr);
        dead = __any_sync(0xffffffffu, hit);
      }
      if (!dead && lane == 0) atomicOr(&gs->alive, 1ull << r);
    }
    // (b) intra-chunk mask: row r suppresses column q > r
    for (int r = warp; r < cn; r += kNmsWarps) {
      int slot = (int)(uint32_t)sorted[c0 + r];
      float4 br = box[slot];
      float ar = (br.z - br.x) * (br.w - br.y);
      unsigned long long row = 0ull;
#pragma unroll
      for (int h = 0; h < kNmsChunk / 32; ++h) {
        int q = h * 32 + lane;
        bool hit = false;
        if (q > r && q < cn) {
          float4 bq = box[(int)(uint32_t)sorted[c0 + q]];
          float aq = (bq.z - bq.x) * (bq.w - bq.y);
          hit = nms_overlap(br, ar, bq, aq, thr);
        }
        unsigned w = __ballot_sync(0xffffffffu, hit);
        row |= (unsigned long long)w << (32 * h);
      }
      if (lane == 0) gs->mask[r] = row;
    }
    __syncthreads();
    // (c) serial resolve on warp 0
    if (warp == 0) {
      unsigned long long m0 = lane < cn ? gs->mask[lane] : 0ull;
      unsigned long long m1 = lane + 32 < cn ? gs->mask[lane + 32] : 0ull;
      unsigned long long remaining = gs->alive, kept = 0ull;
      while (remaining) {
        int j = __ffsll((long long)remaining) - 1;
        kept |= 1ull << j;
        unsigned long long mine = j < 32 ? m0 : m1;
        unsigned long long row = __shfl_sync(0xffffffffu, mine, j & 31);
        remaining &= ~row;
        remaining &= ~(1ull << j);
      }
      int room = max_keep - kc;
#pragma unroll
      for (int h = 0; h < kNmsChunk / 32; ++h) {
        int q = h * 32 + lane;
        if ((kept >> q) & 1ull) {
          int rank = __popcll(kept & ((1ull << q) - 1ull));
          if (rank < room) {
            int slot = (int)(uint32_t)sorted[c0 + q];
            float4 bq = box[slot];
            kbox[kc + rank] = bq;
            karea[kc + rank] = (bq.z - bq.x) * (bq.w - bq.y);
            kslot[kc + rank] = slot;
          }
        }
      }
      if (lane == 0) gs->kcount = kc + min(room, __popcll(kept));
    }
    __syncthreads();
  }
  return gs->kcount;
}

}  // namespace fvb
