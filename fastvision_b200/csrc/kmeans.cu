// Anchor k-means (second half): one Lloyd iteration of KMeans._fit, detection/tools/ANCHOR.py:33-46, with
// the reference's distance 1 - wh_iou_batch(samples, centers) (IOU.py:158-175).  The reference does it in numpy on the host:
// an [n,k] distance matrix, argmin, and k boolean-mask means per iteration.  Here: one thread per sample finds its nearest
// centre (first minimum, as np.argmin) and the per-cluster coordinate sums / counts are reduced in shared memory (fp64), then
// one atomic per cluster per CTA; a second tiny kernel forms the new centres (an empty cluster keeps its centre, :39-40).
#include "common.cuh"

namespace fvb {

constexpr int kKmThreads = 256;
constexpr int kKmMaxK = 64;

struct KmParams {
  const float* samples;  // [n,2]
  long long n;
  const float* centers;  // [k,2]
  int k;
  float eps;
  long long* categories;  // [n], 1-based (ANCHOR.py:35)
  double* acc;            // [k][3] = sum_w, sum_h, count
  float* new_centers;     // [k,2]
};

__global__ void km_zero_kernel(KmParams p) {
  if (threadIdx.x < p.k * 3) p.acc[threadIdx.x] = 0.0;
}

__global__ void __launch_bounds__(kKmThreads) km_assign_kernel(KmParams p) {
  __shared__ float s_c[kKmMaxK * 2];
  __shared__ double s_acc[kKmMaxK * 3];
  for (int i = threadIdx.x; i < p.k * 2; i += kKmThreads) s_c[i] = p.centers[i];
  for (int i = threadIdx.x; i < p.k * 3; i += kKmThreads) s_acc[i] = 0.0;
  __syncthreads();
  const long long i = (long long)blockIdx.x * kKmThreads + threadIdx.x;
  if (i < p.n) {
    const float w = p.samples[i * 2], h = p.samples[i * 2 + 1];
    float best = 0.0f;
    int bi = 0;
    for (int c = 0; c < p.k; ++c) {
      const float d = 1.0f - wh_iou(w, h, s_c[c * 2], s_c[c * 2 + 1], p.eps);  // cal_distance, ANCHOR.py:21-24
      if (c == 0 || d < best) {  // np.argmin: first minimum
        best = d;
        bi = c;
      }
    }
    p.categories[i] = bi + 1;
    atomicAdd(&s_acc[bi * 3 + 0], (double)w);
    atomicAdd(&s_acc[bi * 3 + 1], (double)h);
    atomicAdd(&s_acc[bi * 3 + 2], 1.0);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < p.k * 3; j += kKmThreads)
    if (s_acc[j] != 0.0) atomicAdd(&p.acc[j], s_acc[j]);
}

__global__ void km_update_kernel(KmParams p) {
  const int c = threadIdx.x;
  if (c >= p.k) return;
  const double cnt = p.acc[c * 3 + 2];
  if (cnt > 0.0) {  // np.mean of the cluster's widths / heights (:42-44)
    p.new_centers[c * 2] = (float)(p.acc[c * 3] / cnt);
    p.new_centers[c * 2 + 1] = (float)(p.acc[c * 3 + 1] / cnt);
  } else {          // empty cluster keeps its centre (:39-40)
    p.new_centers[c * 2] = p.centers[c * 2];
    p.new_centers[c * 2 + 1] = p.centers[c * 2 + 1];
  }
}

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_kmeans_workspace_bytes(int k) { return (size_t)(k > 0 ? k : 1) * 3 * 8 + 256; }

extern "C" int fvb_kmeans_step_f32(const float* d_samples, int64_t n, const float* d_centers, int k, float eps,
                                   int64_t* d_categories, float* d_new_centers, void* d_ws, void* stream) {
  FVB_REQUIRE(n >= 1 && k >= 1 && k <= kKmMaxK, "kmeans_step: n=%lld k=%d (k <= %d)", (long long)n, k, kKmMaxK);
  FVB_REQUIRE(d_samples && d_centers && d_categories && d_new_centers && d_ws, "kmeans_step: NULL pointer");
  FVB_REQUIRE(d_centers != d_new_centers, "kmeans_step: centers and new_centers must not alias");
  KmParams p;
  p.samples = d_samples; p.n = n; p.centers = d_centers; p.k = k; p.eps = eps;
  p.categories = (long long*)d_categories; p.acc = (double*)d_ws; p.new_centers = d_new_centers;
  cudaStream_t s = (cudaStream_t)stream;
  km_zero_kernel<<<1, 256, 0, s>>>(p);
  km_assign_kernel<<<(unsigned)((n + kKmThreads - 1) / kKmThreads), kKmThreads, 0, s>>>(p);
  km_update_kernel<<<1, 64, 0, s>>>(p);
  count_launch(3);
  return check_launch("kmeans_step");
}
