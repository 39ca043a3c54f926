// AP integration on the device: CalculateMAP.fetch / _ap_per_class / compute_ap, metrics/map.py:85-141.
// The reference concatenates every image's `correct` rows on the host and, per seen class, argsorts by confidence, takes
// cumulative TP/FP sums, the precision envelope and a 101-point interpolation -- O(sum M log sum M) of numpy on <= 1.5 M
// rows at BASELINE configs[4].  Here:
//   ap_keys      : key = (class << 32) | descending-orderable(conf), payload = row index
//   radix passes : global stable LSD radix sort, 8-bit digits (per-tile histogram -> per-digit row scans -> stable scatter with the
//                  warp match.any multisplit of nms.cuh); 4 passes for the confidence bits + 1-2 for the class bits
//   ap_segments  : class segment boundaries in the sorted order; ap_targets: positives per class (integer atomics)
//   ap_class     : one CTA per (class, IoU threshold): TP prefix sums, precision, right-to-left envelope, the 101
//                  np.interp queries by binary search over the recall, np.trapz -- all in float64 with numpy's formulas.
// Integer work is exact; equal (class, conf) keys keep their row order (the reference's argsort leaves ties unspecified).
#include "nms.cuh"

namespace fvb {

constexpr int kSortThreads = kNmsThreads;  // 512: block_exclusive_scan of nms.cuh is written for this CTA size
constexpr int kSortWarps = kNmsWarps;
constexpr int kSortPerWarp = 256;
constexpr int kSortTile = kSortWarps * kSortPerWarp;  // 4096 keys per CTA

struct ApKeyParams {
  const float* dets;  // [n,6] = [cls, conf, ...]
  long long n;
  int max_class;
  unsigned long long* key;
  uint32_t* idx;
};

__global__ void ap_keys_kernel(const ApKeyParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const float c = p.dets[i * 6], conf = p.dets[i * 6 + 1];
  // rows whose class is not an integer in [0, max_class] can never equal a seen target class: park them in the last slot
  int ci = (c >= 0.0f && c <= (float)p.max_class && c == floorf(c)) ? (int)c : p.max_class + 1;
  p.key[i] = ((unsigned long long)(uint32_t)ci << 32) | (unsigned long long)desc_key(conf);
  p.idx[i] = (uint32_t)i;
}

struct SortParams {
  const unsigned long long* src_key;
  const uint32_t* src_idx;
  unsigned long long* dst_key;
  uint32_t* dst_idx;
  long long n;
  int shift, nblocks;
  uint32_t* hist;    // [256][nblocks]
  uint32_t* totals;  // [256]
};

__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const SortParams p) {
  __shared__ uint32_t cnt[256];
  if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortTile;
  for (int k = threadIdx.x; k < kSortTile; k += kSortThreads) {
    const long long i = base + k;
    if (i < p.n) atomicAdd(&cnt[(uint32_t)(p.src_key[i] >> p.shift) & 255u], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256) p.hist[(size_t)threadIdx.x * p.nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// One CTA per digit: exclusive scan of the digit's row hist[d][0..nblocks) in place, row total -> totals[d].  The scatter
// kernel adds the exclusive scan of the 256 totals itself (a single-CTA scan of all 256*nblocks entries took 87 us per pass).
__global__ void __launch_bounds__(256) sort_scan_kernel(uint32_t* hist, int nblocks, uint32_t* totals) {
  __shared__ uint32_t warp_tot[9];
  __shared__ uint32_t carry;
  uint32_t* row = hist + (size_t)blockIdx.x * nblocks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 256) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? row[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t tv = lane < 8 ? warp_tot[lane] : 0u;
      uint32_t ti = tv;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      if (lane < 8) warp_tot[lane] = ti - tv;
      if (lane == 7) warp_tot[8] = ti;
    }
    __syncthreads();
    if (i < nblocks) row[i] = carry + warp_tot[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[8];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const SortParams p) {
  __shared__ uint32_t cnt[kSortWarps * 256];
  __shared__ uint32_t warp_tot[kSortWarps + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long tile0 = (long long)blockIdx.x * kSortTile;
  const long long beg = min(p.n, tile0 + (long long)warp * kSortPerWarp), end = min(p.n, beg + kSortPerWarp);
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) cnt[i] = 0;
  __syncthreads();
  for (long long b0 = beg; b0 < end; b0 += 32) {
    const long long i = b0 + lane;
    const bool act = i < end;
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (act) {
      const uint32_t d = (uint32_t)(p.src_key[i] >> p.shift) & 255u;
      const unsigned peers = __match_any_sync(m, d);
      if (lane == __ffs(peers) - 1) cnt[d * kSortWarps + warp] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  {
    constexpr int PER = kSortWarps * 256 / kSortThreads;  // 8
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
  // cnt[d*W + w] = keys of this tile with a smaller digit, or the same digit in an earlier warp.  Turn it into the global
  // position: + (scanned global histogram of (d, tile)) - (keys of this tile with a smaller digit).
  __shared__ uint32_t dig_base[256];
  {
    // exclusive scan of the 256 digit totals (first 8 warps), then: global start of digit d + keys of digit d in earlier
    // tiles - keys of this tile with a smaller digit
    const uint32_t tv = threadIdx.x < 256 ? p.totals[threadIdx.x] : 0u;
    uint32_t inc = tv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (threadIdx.x < 256 && lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp && w < 8; ++w) before += warp_tot[w];
    if (threadIdx.x < 256)
      dig_base[threadIdx.x] = (before + inc - tv) + p.hist[(size_t)threadIdx.x * p.nblocks + blockIdx.x] - cnt[threadIdx.x * kSortWarps];
  }
  __syncthreads();
  for (long long b0 = beg; b0 < end; b0 += 32) {
    const long long i = b0 + lane;
    const bool act = i < end;
    const unsigned m = __ballot_sync(0xffffffffu, act);
    uint32_t d = 0;
    unsigned peers = 0;
    if (act) {
      const unsigned long long key = p.src_key[i];
      d = (uint32_t)(key >> p.shift) & 255u;
      peers = __match_any_sync(m, d);
      const uint32_t pos = dig_base[d] + cnt[d * kSortWarps + warp] + __popc(peers & lt_mask);
      p.dst_key[pos] = key;
      p.dst_idx[pos] = p.src_idx[i];
    }
    __syncwarp();
    if (act && lane == __ffs(peers) - 1) cnt[d * kSortWarps + warp] += __popc(peers);
    __syncwarp();
  }
}

struct ApParams {
  const unsigned long long* key;  // sorted
  const uint32_t* idx;
  long long n;
  const unsigned char* correct;   // [n, n_thr] in ORIGINAL row order
  int n_thr, slots;               // slots = max_class + 2
  int* seg_start;                 // [slots]
  int* seg_end;
  int* pos_count;                 // [slots]
  const float* target_cls;
  long long m;
  uint32_t* tp;                   // [n_thr][n]
  double* env;                    // [n_thr][n]
  double* ap;                     // [slots-1][n_thr]
};

__global__ void ap_init_kernel(const ApParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.slots) {
    p.seg_start[i] = 0;
    p.seg_end[i] = 0;
    p.pos_count[i] = 0;
  }
}

__global__ void ap_segments_kernel(const ApParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int c = (int)(p.key[i] >> 32);
  if (i == 0 || (int)(p.key[i - 1] >> 32) != c) p.seg_start[c] = (int)i;
  if (i == p.n - 1 || (int)(p.key[i + 1] >> 32) != c) p.seg_end[c] = (int)i + 1;
}

__global__ void ap_targets_kernel(const ApParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.m) return;
  const float c = p.target_cls[i];
  const int ci = (c >= 0.0f && c <= (float)(p.slots - 2) && c == floorf(c)) ? (int)c : p.slots - 1;
  atomicAdd(&p.pos_count[ci], 1);
}

constexpr int kApThreads = 512;

__global__ void __launch_bounds__(kApThreads) ap_class_kernel(const ApParams p) {
  __shared__ uint32_t s_u[kApThreads / 32 + 1];
  __shared__ double s_d[kApThreads / 32 + 1];
  __shared__ double s_y[101];
  __shared__ uint32_t carry_u;
  __shared__ double carry_d;
  const int c = blockIdx.x, k = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = p.pos_count[c];
  if (P == 0) {  // class never seen among the targets: not part of the mean (map.py:126-127)
    if (threadIdx.x == 0) p.ap[(size_t)c * p.n_thr + k] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  const int s = p.seg_start[c], n = p.seg_end[c] - s;
  uint32_t* tp = p.tp + (size_t)k * p.n + s;
  double* env = p.env + (size_t)k * p.n + s;
  // forward: TP prefix sums (np.cumsum(correct), map.py:108)
  if (threadIdx.x == 0) carry_u = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kApThreads) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < n ? (uint32_t)p.correct[(size_t)p.idx[s + i] * p.n_thr + k] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_u[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t tv = lane < kApThreads / 32 ? s_u[lane] : 0u;
      uint32_t ti = tv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      if (lane < kApThreads / 32) s_u[lane] = ti - tv;
      if (lane == kApThreads / 32 - 1) s_u[kApThreads / 32] = ti;
    }
    __syncthreads();
    if (i < n) tp[i] = carry_u + s_u[warp] + inc;
    __syncthreads();
    if (threadIdx.x == 0) carry_u += s_u[kApThreads / 32];
    __syncthreads();
  }
  // backward: envelope[i] = max_{j >= i} precision[j], precision = TP / (TP + FP + 1e-16) with TP + FP = i + 1 (:110-113, :86-89)
  if (threadIdx.x == 0) carry_d = 0.0;  // the appended sentinel precision 0.0
  __syncthreads();
  const int chunks = (n + kApThreads - 1) / kApThreads;
  for (int ch = chunks - 1; ch >= 0; --ch) {
    const int i = ch * kApThreads + threadIdx.x;
    double v = 0.0;
    if (i < n) {
      const double t = (double)tp[i];
      v = t / ((t + (double)(i + 1 - (int)tp[i])) + 1e-16);
    }
    double mx = v;  // suffix max inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_down_sync(0xffffffffu, mx, o);
      if (lane + o < 32) mx = fmax(mx, u);
    }
    if (lane == 0) s_d[warp] = mx;
    __syncthreads();
    if (warp == 0) {
      double t = lane < kApThreads / 32 ? s_d[lane] : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_down_sync(0xffffffffu, t, o);
        if (lane + o < 32) t = fmax(t, u);
      }
      // t = max over warps >= lane; store the max over warps > lane
      const double nxt = __shfl_down_sync(0xffffffffu, t, 1);
      if (lane < kApThreads / 32) s_d[lane] = lane + 1 < kApThreads / 32 ? nxt : 0.0;
      if (lane == 0) s_d[kApThreads / 32] = t;
    }
    __syncthreads();
    if (i < n) env[i] = fmax(fmax(mx, s_d[warp]), carry_d);
    __syncthreads();
    if (threadIdx.x == 0) carry_d = fmax(carry_d, s_d[kApThreads / 32]);
    __syncthreads();
  }
  // np.interp(np.linspace(0, 1, 101), m_recall, envelope) with m_recall = [0, recall..., 1], envelope = [1, env..., 0]
  const double denom_fn = (double)P;
  auto xp = [&](int j) -> double {  // j in [0, n+1]
    if (j == 0) return 0.0;
    if (j == n + 1) return 1.0;
    const double t = (double)tp[j - 1];
    return t / ((t + (denom_fn - t)) + 1e-16);  // TP / (TP + FN + 1e-16), FN = total_positive - TP (:107-110)
  };
  auto fp = [&](int j) -> double {
    if (j == 0) return fmax(1.0, n > 0 ? env[0] : 0.0);
    if (j == n + 1) return 0.0;
    return env[j - 1];
  };
  if (threadIdx.x < 101) {
    const int q = threadIdx.x;
    const double x = q == 100 ? 1.0 : (double)q * (1.0 / 100.0);
    int lo = 0, hi = n + 2;  // numpy binary search: largest j with xp[j] <= x
    while (lo < hi) {
      const int mid = lo + ((hi - lo) >> 1);
      if (x >= xp(mid)) lo = mid + 1;
      else hi = mid;
    }
    const int j = lo - 1;
    double y;
    if (j == n + 1) y = fp(j);
    else {
      const double xj = xp(j);
      if (xj == x) y = fp(j);
      else {
        const double slope = (fp(j + 1) - fp(j)) / (xp(j + 1) - xj);
        y = slope * (x - xj) + fp(j);
      }
    }
    s_y[q] = y;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // np.trapz: sum(d * (y[1:] + y[:-1]) / 2.0) with numpy's pairwise order for 100 terms (8 lanes, then the tail)
    double r[8];
    auto term = [&](int i) -> double {
      const double x1 = i + 1 == 100 ? 1.0 : (double)(i + 1) * (1.0 / 100.0), x0 = (double)i * (1.0 / 100.0);
      return (x1 - x0) * (s_y[i + 1] + s_y[i]) / 2.0;
    };
    for (int j = 0; j < 8; ++j) r[j] = term(j);
    for (int i = 8; i < 96; i += 8)
      for (int j = 0; j < 8; ++j) r[j] += term(i + j);
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (int i = 96; i < 100; ++i) res += term(i);
    p.ap[(size_t)c * p.n_thr + k] = res;
  }
}

static size_t ap_align(size_t x) { return (x + 255) / 256 * 256; }

struct ApLayout {
  size_t keyA, keyB, idxA, idxB, hist, seg_start, seg_end, tp, env, total;
  int nblocks;
};

static ApLayout ap_layout(long long n, int n_thr, int max_class) {
  ApLayout L;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  L.nblocks = (int)((nn + kSortTile - 1) / kSortTile);
  size_t o = 0;
  L.keyA = o; o = ap_align(o + nn * 8);
  L.keyB = o; o = ap_align(o + nn * 8);
  L.idxA = o; o = ap_align(o + nn * 4);
  L.idxB = o; o = ap_align(o + nn * 4);
  L.hist = o; o = ap_align(o + (size_t)256 * L.nblocks * 4 + 1024);
  L.seg_start = o; o = ap_align(o + (size_t)(max_class + 2) * 4);
  L.seg_end = o; o = ap_align(o + (size_t)(max_class + 2) * 4);
  L.tp = o; o = ap_align(o + nn * n_thr * 4);
  L.env = o; o = ap_align(o + nn * n_thr * 8);
  L.total = o + 256;
  return L;
}

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_map_ap_workspace_bytes(int64_t n_dets, int n_thr, int max_class) {
  if (n_dets < 0 || n_thr < 1 || max_class < 0) return 0;
  return ap_layout(n_dets, n_thr, max_class).total;
}

extern "C" int fvb_map_ap_f64(const float* d_dets, const uint8_t* d_correct, int64_t n_dets, const float* d_target_cls,
                              int64_t n_targets, int n_thr, int max_class, double* d_ap, int32_t* d_pos_count, void* d_ws,
                              void* stream) {
  FVB_REQUIRE(n_dets >= 0 && n_dets < (1ll << 31) && n_targets >= 0, "map_ap: n_dets=%lld n_targets=%lld", (long long)n_dets, (long long)n_targets);
  FVB_REQUIRE(n_thr >= 1 && n_thr <= 16, "map_ap: n_thr=%d", n_thr);
  FVB_REQUIRE(max_class >= 0 && max_class < 65535, "map_ap: max_class=%d (class ids must be < 65535)", max_class);
  FVB_REQUIRE(d_ap && d_pos_count && d_ws, "map_ap: NULL pointer");
  FVB_REQUIRE(n_dets == 0 || (d_dets && d_correct), "map_ap: dets NULL");
  FVB_REQUIRE(n_targets == 0 || d_target_cls, "map_ap: targets NULL");
  FVB_REQUIRE(((uintptr_t)d_ws & 255) == 0, "map_ap: workspace must be 256-byte aligned");
  const ApLayout L = ap_layout(n_dets, n_thr, max_class);
  unsigned char* w = (unsigned char*)d_ws;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* key[2] = {(unsigned long long*)(w + L.keyA), (unsigned long long*)(w + L.keyB)};
  uint32_t* idx[2] = {(uint32_t*)(w + L.idxA), (uint32_t*)(w + L.idxB)};
  int cur = 0;
  if (n_dets > 0) {
    ApKeyParams kp;
    kp.dets = d_dets; kp.n = n_dets; kp.max_class = max_class; kp.key = key[0]; kp.idx = idx[0];
    ap_keys_kernel<<<(unsigned)((n_dets + 255) / 256), 256, 0, s>>>(kp);
    count_launch();
    const int class_bits = max_class + 1 < 256 ? 8 : 16;
    for (int shift = 0; shift < 32 + class_bits; shift += 8) {
      SortParams sp;
      sp.src_key = key[cur]; sp.src_idx = idx[cur]; sp.dst_key = key[cur ^ 1]; sp.dst_idx = idx[cur ^ 1];
      sp.n = n_dets; sp.shift = shift; sp.nblocks = L.nblocks; sp.hist = (uint32_t*)(w + L.hist);
      sp.totals = sp.hist + (size_t)256 * L.nblocks;
      sort_hist_kernel<<<L.nblocks, kSortThreads, 0, s>>>(sp);
      sort_scan_kernel<<<256, 256, 0, s>>>(sp.hist, L.nblocks, sp.totals);
      sort_scatter_kernel<<<L.nblocks, kSortThreads, 0, s>>>(sp);
      count_launch(3);
      cur ^= 1;
    }
  }
  ApParams p;
  p.key = key[cur]; p.idx = idx[cur]; p.n = n_dets; p.correct = d_correct; p.n_thr = n_thr; p.slots = max_class + 2;
  p.seg_start = (int*)(w + L.seg_start); p.seg_end = (int*)(w + L.seg_end); p.pos_count = d_pos_count;
  p.target_cls = d_target_cls; p.m = n_targets; p.tp = (uint32_t*)(w + L.tp); p.env = (double*)(w + L.env); p.ap = d_ap;
  ap_init_kernel<<<(p.slots + 255) / 256, 256, 0, s>>>(p);
  count_launch();
  if (n_dets > 0) {
    ap_segments_kernel<<<(unsigned)((n_dets + 255) / 256), 256, 0, s>>>(p);
    count_launch();
  }
  if (n_targets > 0) {
    ap_targets_kernel<<<(unsigned)((n_targets + 255) / 256), 256, 0, s>>>(p);
    count_launch();
  }
  dim3 grid((unsigned)(max_class + 1), (unsigned)n_thr);
  ap_class_kernel<<<grid, kApThreads, 0, s>>>(p);
  count_launch();
  return check_launch("map_ap");
}
