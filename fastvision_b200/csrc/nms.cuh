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
//   * otherwise compare inter against thr*union: unless the two are within 1e-5 (relative) of each
//     other that product test has the same outcome as the rounded quotient test, and only in the
//     narrow band is the exact IEEE division evaluated.
__device__ __forceinline__ bool nms_overlap(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  const float w = fmaxf(0.0f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
  const float h = fmaxf(0.0f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
  const float inter = w * h;
  if (thr >= 0.0f && !(inter > 0.0f)) return false;
  const float uni = (area_a + area_b) - inter;
  const float tu = thr * uni;
  if (uni > 1e-30f && uni < 1e30f && inter < 1e30f && thr > 1e-6f && thr < 1e6f) {
    if (inter > tu * 1.00001f) return true;
    if (inter < tu * 0.99999f) return false;
  }
  return inter / uni > thr;
}

// The same decision without the division for all but a sliver of pairs:  inter/(s - inter) > thr  <=>  inter > t*s
// with t = thr/(1+thr), s = area_a + area_b (inter > 0 implies both areas and the union are positive).  `hit` is set
// when inter clears the bound by 1e-5 (far beyond fp32 rounding of either form), `amb` when it is at least within
// 1e-5 below it; a pair that is only `amb` must be re-decided with nms_overlap().  NaN boxes set neither.
struct NmsFast {
  float thi, tlo;
  bool ok;  // thresholds outside (1e-6, 1e6) take the exact path only
};
__device__ __forceinline__ NmsFast make_nms_fast(float thr) {
  NmsFast f;
  f.ok = thr > 1e-6f && thr < 1e6f;
  const float t = thr / (1.0f + thr);
  f.thi = t * 1.00001f;
  f.tlo = t * 0.99999f;
  return f;
}
__device__ __forceinline__ void nms_overlap_fast(const float4& a, float area_a, const float4& b, float area_b,
                                                 const NmsFast& f, bool& hit, bool& amb) {
  const float w = fminf(a.z, b.z) - fmaxf(a.x, b.x);
  const float h = fminf(a.w, b.w) - fmaxf(a.y, b.y);
  const float inter = fmaxf(w, 0.0f) * fmaxf(h, 0.0f);
  const float s = area_a + area_b;
  hit |= inter > fmaxf(f.thi * s, 0.0f);
  amb |= inter >= f.tlo * s;
}

// (NT = threads of the calling CTA: 512 everywhere except the 1024-thread small-batch variant of yolo_nms_kernel)
template <int NT = kNmsThreads>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_tot /*[NT/32+1]*/) {
  constexpr int kNmsWarps = NT / 32;  // shadows the namespace constant inside this function
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
template <int NT = kNmsThreads>
__device__ inline unsigned long long* block_radix_sort_hi32(unsigned long long* src, unsigned long long* dst, int n,
                                                     uint32_t* cnt, uint32_t* warp_tot) {
  constexpr int kNmsWarps = NT / 32, kNmsThreads = NT;  // shadow the namespace constants inside this function
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  int chunk = (((n + kNmsWarps - 1) / kNmsWarps) + 31) & ~31;
  int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  for (int shift = 32; shift < 64; shift += 8) {
    for (int i = threadIdx.x; i < kNmsWarps * 256; i += kNmsThreads) cnt[i] = 0;
    __syncthreads();
    // keys may live in global memory (more candidates than the shared-memory capacity, e.g. the RPN's 22 500 anchors): four
    // batches of 32 keys are loaded before any of them is counted, so the loop is not one L2 round trip per 32 keys
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
      uint32_t run = block_exclusive_scan<NT>(sum, warp_tot);
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        cnt[threadIdx.x * PER + k] = run;
        run += loc[k];
      }
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
          const uint32_t pos = cnt[d * kNmsWarps + warp] + __popc(peers & lt_mask);
          dst[pos] = key;
        }
        __syncwarp();
        if (act && lane == __ffs(peers) - 1) cnt[d * kNmsWarps + warp] += __popc(peers);
        __syncwarp();
      }
    }
    __syncthreads();
    unsigned long long* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

// ---- block bitonic sort of up to 2048 distinct 64-bit keys held in shared memory ------------------------------
// The per-image candidate lists of the YOLO path (hundreds to ~2000 keys) are latency-bound, not throughput-bound:
// the 4-pass LSD radix sort above costs ~24 block barriers (16 us for 778 keys).  A bitonic network over the full
// 64-bit key (rank in the high word, slot in the low word: all keys distinct, so ascending order == the stable
// order the radix sort produces) keeps E consecutive keys per thread in registers; every compare-exchange whose
// partner lies inside the warp's 32*E keys is a register swap or one shuffle, only partners in other warps go
// through shared memory (10 barriers for 1024 keys).
template <int E>
__device__ __forceinline__ void bitonic_warp_steps(unsigned long long (&v)[E], int base, int k, int jstart) {
  const int lane = threadIdx.x & 31;
  // partners in other lanes: j = jstart .. E
  for (int j = jstart; j >= E; j >>= 1) {
    const int lm = j / E;
    const bool lower = (lane & lm) == 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool asc = ((base + e) & k) == 0;
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[e], lm);
      const bool take_min = lower == asc;
      const bool o_less = o < v[e];
      v[e] = (o_less == take_min) ? o : v[e];
    }
  }
  // partners in the same thread: j = min(jstart, E/2) .. 1
#pragma unroll
  for (int j = E / 2; j >= 1; j >>= 1) {
    if (j <= jstart) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & j) == 0) {
          const bool asc = ((base + e) & k) == 0;
          const unsigned long long a = v[e], b = v[e | j];
          const bool sw = (a > b) == asc;
          v[e] = sw ? b : a;
          v[e | j] = sw ? a : b;
        }
      }
    }
  }
}

template <int E, int NT>
__device__ __forceinline__ void block_bitonic_run(unsigned long long* keys, int np) {
  constexpr int kNmsThreads = NT;
  const int tid = threadIdx.x;
  const int base = tid * E;
  const bool active = base < np;  // warp-uniform: np is a multiple of 32*E or smaller than it only for np = 64 < 128 (E = 4 never sees that)
  unsigned long long v[E];
  constexpr int W = 32 * E;  // keys per warp
  if (active) {
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(keys + base + e);
      v[e] = t.x;
      v[e + 1] = t.y;
    }
    for (int k = 2; k <= min(np, W); k <<= 1) bitonic_warp_steps<E>(v, base, k, k >> 1);
  }
  for (int k = 2 * W; k <= np; k <<= 1) {
    if (active) {
#pragma unroll
      for (int e = 0; e < E; e += 2) *reinterpret_cast<ulonglong2*>(keys + base + e) = make_ulonglong2(v[e], v[e + 1]);
    }
    __syncthreads();
    for (int j = k >> 1; j >= W; j >>= 1) {
      for (int q = tid; q < (np >> 1); q += kNmsThreads) {
        const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        const unsigned long long a = keys[i], b = keys[i | j];
        const bool asc = (i & k) == 0;
        if ((a > b) == asc) {
          keys[i] = b;
          keys[i | j] = a;
        }
      }
      __syncthreads();
    }
    if (active) {
#pragma unroll
      for (int e = 0; e < E; e += 2) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(keys + base + e);
        v[e] = t.x;
        v[e + 1] = t.y;
      }
      bitonic_warp_steps<E>(v, base, k, W >> 1);
    }
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < E; e += 2) *reinterpret_cast<ulonglong2*>(keys + base + e) = make_ulonglong2(v[e], v[e + 1]);
  }
  __syncthreads();
}

// Sorts keys[0..n) ascending in place.  keys must be a 16-byte aligned SHARED-memory array with room for the next power of two
// >= max(n, 64) (<= 2048) entries; entries [n, np) are overwritten with the all-ones key.  Call with all NT threads of the CTA.
template <int NT = kNmsThreads>
__device__ inline void block_bitonic_sort64(unsigned long long* keys, int n) {
  int np = 64;
  while (np < n) np <<= 1;
  for (int i = n + threadIdx.x; i < np; i += NT) keys[i] = ~0ull;
  __syncthreads();
  if (np <= 1024)  // (also with 1024 threads: 4 keys per thread keep more of a 2048-key network inside a warp: 8.6 vs 9.4 us)
    block_bitonic_run<2, NT>(keys, np);
  else
    block_bitonic_run<4, NT>(keys, np);
}

struct GreedyShared {
  float4 cbox[2][kNmsChunk];           // staged boxes of the current / next chunk (double buffer)
  float carea[2][kNmsChunk];
  int cslot[2][kNmsChunk];
  unsigned long long mask[kNmsChunk];  // intra-chunk suppression rows (bit q: row suppresses column q)
  unsigned alive32[2];                 // columns of the current chunk not suppressed by the kept list
  unsigned nz32[2];                    // rows whose intra-chunk mask is non-zero
  int kcount;
};

__device__ __forceinline__ void stage_chunk(GreedyShared* gs, int buf, const unsigned long long* sorted, int c0, int n,
                                            const float4* box, int t) {
  if (c0 + t < n) {
    const int slot = (int)(uint32_t)sorted[c0 + t];
    const float4 bq = box[slot];
    gs->cbox[buf][t] = bq;
    gs->carea[buf][t] = (bq.z - bq.x) * (bq.w - bq.y);
    gs->cslot[buf][t] = slot;
  }
}

// Greedy NMS over candidates already sorted by rank: sorted[p] low word = slot into box[] (xyxy, the
// boxes the IoU is taken on).  Walks the order in chunks of 64 staged in shared memory:
//   (a) every candidate of the chunk against the kept list -- thread (r, s) tests candidate r against
//       kept boxes s, s+8, ... (independent iterations, the kept box is a warp-wide broadcast);
//   (b) the 64x64 intra-chunk mask, one __ballot_sync per 32 columns (warp per row);
//   (c) warp 0 resolves the chunk: only alive rows with a non-zero mask need a serial step; the other
//       warps meanwhile stage the next chunk.
// Identical keep set to the sequential algorithm: j is dropped iff an earlier KEPT i has IoU > thr.
// kbox/karea/kslot: kept list (max_keep entries, shared memory).  Returns the kept count (<= max_keep).
template <int NT = kNmsThreads>
__device__ inline int block_greedy_nms(const unsigned long long* sorted, int n, const float4* box, float thr, int max_keep,
                                float4* kbox, float* karea, int* kslot, GreedyShared* gs) {
  constexpr int kNmsWarps = NT / 32, kNmsThreads = NT;  // shadow the namespace constants inside this function
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  const NmsFast fast = make_nms_fast(thr);
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
    const int kc = gs->kcount;
    if (kc >= max_keep) break;
    const int cn = min(kNmsChunk, n - c0);
    // (a) against the kept list: branch-free fast decisions (two kept boxes in flight per step); a candidate that
    // met a pair inside the guard band and no clear hit re-walks its slice with the exact test (rare)
    {
      const int r = tid & (kNmsChunk - 1), s = tid / kNmsChunk;
      constexpr int S = kNmsThreads / kNmsChunk;
      if (r < cn && s < kc) {
        const float4 bj = gs->cbox[buf][r];
        const float aj = gs->carea[buf][r];
        bool hit = false;
        if (fast.ok) {
          bool amb = false;
          int i = s;
          for (; i + S < kc; i += 2 * S) {
            const float4 k0 = kbox[i], k1 = kbox[i + S];
            const float a0 = karea[i], a1 = karea[i + S];
            nms_overlap_fast(k0, a0, bj, aj, fast, hit, amb);
            nms_overlap_fast(k1, a1, bj, aj, fast, hit, amb);
          }
          if (i < kc) nms_overlap_fast(kbox[i], karea[i], bj, aj, fast, hit, amb);
          if (amb && !hit) {
            for (i = s; i < kc; i += S)
              if (nms_overlap(kbox[i], karea[i], bj, aj, thr)) {
                hit = true;
                break;
              }
          }
        } else {
          for (int i = s; i < kc; i += S) {
            if (nms_overlap(kbox[i], karea[i], bj, aj, thr)) {
              hit = true;
              break;
            }
          }
        }
        if (hit) atomicAnd(&gs->alive32[r >> 5], ~(1u << (r & 31)));
      }
    }
    // (b) intra-chunk mask: row r suppresses column q > r
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
    // (c) resolve on warp 0; warps 1-2 stage the next chunk
    if (warp == 0) {
      unsigned long long remaining = ((unsigned long long)gs->alive32[1] << 32) | gs->alive32[0];
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
          const int rank = __popcll(kept & ((1ull << q) - 1ull));
          if (rank < room) {
            kbox[kc + rank] = gs->cbox[buf][q];
            karea[kc + rank] = gs->carea[buf][q];
            kslot[kc + rank] = gs->cslot[buf][q];
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
  return gs->kcount;
}

}  // namespace fvb
