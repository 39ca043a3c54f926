// Glue of the validation loop between the fused step and the mAP matcher, on the device and without a host sync.
//
// Replaces, per batch, what utils/fit.py:94-99 does per IMAGE in Python: `cat([cls.float(), conf, xyxy])` of the kept detections
// (:96) and the targets `labels[labels[:,0]==i]` -> `[cls, xyxy * (w,h,w,h)]` (:98-99).  The reference pays a boolean-mask sync
// per image; the first version of this drop-in did the same with ~15 ATen launches and one mask sync per batch.
//   val_offsets_kernel  (one CTA)   CSR offsets of the detections (scan of cnt) and of the targets (counts per image, scan);
//                                   targets converted to pixel xyxy and written grouped by image, input order kept inside an
//                                   image (what the reference's per-image mask selects) -- directly when the labels already
//                                   come grouped (collate_fn order), by a stable per-image pass otherwise;
//   val_compact_kernel  (CTA/image) padded [B,max_det] boxes/scores/classes -> compact rows [cls, conf, x1, y1, x2, y2].
// Outputs have CAPACITY shapes (B*max_det rows / T rows); the offsets say what is valid, the host never needs the totals.
#include "common.cuh"

namespace fvb {

struct ValGlueParams {
  const float* boxes;      // [B, md, 4]
  const float* scores;     // [B, md]
  const long long* cls;    // [B, md]
  const int* cnt;          // [B] (negative = failed image: treated as empty)
  int B, md;
  const float* labels;     // [T, 6] = [image, cls, xc, yc, w, h] normalised
  int T;
  float img_w, img_h;
  float* dets;             // [B*md, 6]
  int* det_off;            // [B+1]
  float* gts;              // [T, 5]
  int* gt_off;             // [B+1]
};

__device__ __forceinline__ int block_scan_excl_1024(int v, int* warp_tot /*[33]*/, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int t = warp_tot[lane];
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;
    if (lane == 31) warp_tot[32] = ti;
  }
  __syncthreads();
  const int r = warp_tot[warp] + inc - v;
  total = warp_tot[32];
  __syncthreads();
  return r;
}

__device__ __forceinline__ void write_gt(const ValGlueParams& p, int t, int pos) {
  const float* l = p.labels + (size_t)t * 6;
  // utils/fit.py:98-99: xywh2xyxy (BOX.py:4-10: x - w/2 ...), then * (w, h, w, h)
  const float hw = l[4] / 2.0f, hh = l[5] / 2.0f;
  float* o = p.gts + (size_t)pos * 5;
  o[0] = l[1];
  o[1] = (l[2] - hw) * p.img_w;
  o[2] = (l[3] - hh) * p.img_h;
  o[3] = (l[2] + hw) * p.img_w;
  o[4] = (l[3] + hh) * p.img_h;
}

__global__ void __launch_bounds__(1024) val_offsets_kernel(const ValGlueParams p) {
  __shared__ int warp_tot[33];
  __shared__ int unsorted;
  const int tid = threadIdx.x;
  if (tid == 0) unsorted = 0;
  // detections: exclusive scan of max(cnt, 0)
  int carry = 0;
  for (int base = 0; base < p.B; base += 1024) {
    const int i = base + tid;
    const int v = i < p.B ? max(p.cnt[i], 0) : 0;
    int total;
    const int ex = block_scan_excl_1024(v, warp_tot, total);
    if (i < p.B) p.det_off[i] = carry + ex;
    carry += total;
  }
  if (tid == 0) p.det_off[p.B] = carry;
  // targets per image
  for (int i = tid; i <= p.B; i += 1024) p.gt_off[i] = 0;
  __syncthreads();
  for (int t = tid; t < p.T; t += 1024) {
    const int b = (int)p.labels[(size_t)t * 6];
    if (b >= 0 && b < p.B) atomicAdd(&p.gt_off[b + 1], 1);
    if (t + 1 < p.T && b > (int)p.labels[(size_t)(t + 1) * 6]) unsorted = 1;
    if (b < 0 || b >= p.B) unsorted = 1;  // rows of no image of this batch are dropped: positions shift
  }
  __syncthreads();
  carry = 0;
  for (int base = 0; base < p.B; base += 1024) {
    const int i = base + tid;
    const int v = i < p.B ? p.gt_off[i + 1] : 0;
    int total;
    const int ex = block_scan_excl_1024(v, warp_tot, total);
    __syncthreads();
    if (i < p.B) p.gt_off[i + 1] = carry + ex + v;  // inclusive -> offset of the NEXT image
    carry += total;
    __syncthreads();
  }
  if (!unsorted) {
    for (int t = tid; t < p.T; t += 1024) write_gt(p, t, t);  // grouped already: same order
    return;
  }
  // general case: one warp per image walks all labels in order (stable)
  const int lane = tid & 31, warp = tid >> 5;
  for (int b = warp; b < p.B; b += 32) {
    int pos = p.gt_off[b];
    for (int t0 = 0; t0 < p.T; t0 += 32) {
      const int t = t0 + lane;
      const bool mine = t < p.T && (int)p.labels[(size_t)t * 6] == b;
      const unsigned m = __ballot_sync(0xffffffffu, mine);
      if (mine) write_gt(p, t, pos + __popc(m & ((1u << lane) - 1u)));
      pos += __popc(m);
    }
  }
}

__global__ void __launch_bounds__(128) val_compact_kernel(const ValGlueParams p) {
  const int b = blockIdx.x;
  const int n = min(max(p.cnt[b], 0), p.md);
  const int o0 = p.det_off[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t s = (size_t)b * p.md + i;
    const float4 bx = reinterpret_cast<const float4*>(p.boxes)[s];
    float* o = p.dets + (size_t)(o0 + i) * 6;
    o[0] = (float)p.cls[s];  // utils/fit.py:96 categories.float()
    o[1] = p.scores[s];
    o[2] = bx.x;
    o[3] = bx.y;
    o[4] = bx.z;
    o[5] = bx.w;
  }
}

}  // namespace fvb

using namespace fvb;

extern "C" int fvb_val_evidence_f32(const float* d_boxes, const float* d_scores, const int64_t* d_cls, const int32_t* d_cnt,
                                    int batch, int max_det, const float* d_labels, int64_t num_labels, float img_w,
                                    float img_h, float* d_out_dets, int32_t* d_out_det_off, float* d_out_gts,
                                    int32_t* d_out_gt_off, void* stream) {
  FVB_REQUIRE(batch >= 1 && batch <= 65535 && max_det >= 1, "val_evidence: bad shape B=%d max_det=%d", batch, max_det);
  FVB_REQUIRE(num_labels >= 0 && num_labels < (1ll << 24), "val_evidence: num_labels=%lld", (long long)num_labels);
  FVB_REQUIRE(d_boxes && d_scores && d_cls && d_cnt && d_out_dets && d_out_det_off && d_out_gt_off, "val_evidence: NULL pointer");
  FVB_REQUIRE(num_labels == 0 || (d_labels && d_out_gts), "val_evidence: NULL labels / gts");
  FVB_REQUIRE(((uintptr_t)d_boxes & 15) == 0, "val_evidence: boxes must be 16-byte aligned");
  ValGlueParams p;
  p.boxes = d_boxes;
  p.scores = d_scores;
  p.cls = (const long long*)d_cls;
  p.cnt = d_cnt;
  p.B = batch;
  p.md = max_det;
  p.labels = d_labels;
  p.T = (int)num_labels;
  p.img_w = img_w;
  p.img_h = img_h;
  p.dets = d_out_dets;
  p.det_off = d_out_det_off;
  p.gts = d_out_gts;
  p.gt_off = d_out_gt_off;
  cudaStream_t s = (cudaStream_t)stream;
  val_offsets_kernel<<<1, 1024, 0, s>>>(p);
  count_launch();
  val_compact_kernel<<<batch, 128, 0, s>>>(p);
  count_launch();
  return check_launch("val_evidence");
}
