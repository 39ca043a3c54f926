// Data-parallel finish of Yolov3Loss over NVLink peer memory: all-reduce of the L*4 fp64 partial sums
// {S_cls, S_box, S_conf, M} FUSED with the combine into the scalar -- one single-CTA kernel instead of
// NCCL all-reduce (96 bytes: pure launch + protocol latency, ~25-40 us at 8 ranks) + loss_combine.
// Every rank owns a small buffer that its peers map (torch symmetric memory hands out the peer pointers): two slots of
// {vals[16], flag} plus a device-resident epoch.  Call e (epoch kept on the device, so the kernel replays inside a CUDA graph):
// publish my partials into slot e&1 of MY buffer, release-store flag = e, acquire-spin on every peer's flag, then every rank
// sums the W vectors in rank order -- bit-identical results on all ranks -- and forms the loss with the global-batch
// normalisers.  Slot reuse is safe: a rank reaches epoch e+2 only after all peers published e+1, i.e. finished reading e.
// A peer that never arrives trips a ~10 s timeout: status := 1, loss := NaN (no hang); a flag from a LATER epoch (a peer that ran
// on after somebody's timeout and re-used the slot) gives status := 2, loss := NaN.  The status word is sticky;
// dist.PeerReducer.raise_if_failed() surfaces it.
#include "common.cuh"

namespace fvb {

constexpr int kPeerVals = 16;
struct PeerSlot {
  double vals[kPeerVals];
  unsigned long long flag;
  unsigned long long pad;
};
struct PeerBuf {
  PeerSlot slot[2];
  unsigned long long epoch;  // local call counter (only its owner touches it)
};

struct PeerParams {
  Geom g;
  long long batch_global;
  const double* partials;
  const unsigned long long* peers;  // [world] base addresses of every rank's PeerBuf (peer-mapped)
  int rank, world, n;
  float r_box, r_conf, r_cls;
  double* out_partials;
  float* out_loss;
  int* status;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// the loss formula of loss.cu's combine (yolov3_loss.py:49-72 on global sums)
__device__ __forceinline__ float peer_combine(const Geom& g, const double* partials, long long batch_global, float r_box,
                                              float r_conf, float r_cls) {
  double tot = 0.0;
  const int C = g.K - 5;
  for (int l = 0; l < g.L; ++l) {
    const double s_cls = partials[l * 4 + 0], s_box = partials[l * 4 + 1], s_conf = partials[l * 4 + 2], m = partials[l * 4 + 3];
    if (m > 0.0) tot += (double)r_cls * s_cls / (m * C) + (double)r_box * s_box / m;
    tot += (double)r_conf * s_conf / ((double)batch_global * g.A * g.HW[l]);
  }
  return (float)(tot * (double)batch_global);
}

__global__ void __launch_bounds__(32) peer_reduce_combine_kernel(const PeerParams p) {
  __shared__ double s_sum[kPeerVals];
  const int tid = threadIdx.x;
  PeerBuf* mine = reinterpret_cast<PeerBuf*>(p.peers[p.rank]);
  unsigned long long e = 0;
  if (tid == 0) {
    e = mine->epoch + 1;
    mine->epoch = e;
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  const int s = (int)(e & 1ull);
  // one thread publishes: plain stores of the values, then ONE release store of the flag orders them system-wide (a
  // separate __threadfence_system() by the whole warp in front of it cost a second system-scope barrier on the critical path)
  const double mv = tid < p.n ? p.partials[tid] : 0.0;
  for (int i = 0; i < p.n; ++i) {
    const double v = __shfl_sync(0xffffffffu, mv, i);
    if (tid == 0) mine->slot[s].vals[i] = v;
  }
  if (tid == 0) st_release_sys(&mine->slot[s].flag, e);
  // wait for every peer's epoch-e publication (lane <-> peer).  The flag must end up EXACTLY e: a larger value means that peer
  // has already re-used the slot for epoch e+2 (only possible after somebody timed out), i.e. the values are not epoch e's.
  int ok_code = 0;  // 0 ok, 1 timeout, 2 desynchronised
  if (tid < p.world && tid != p.rank) {
    const PeerBuf* pb = reinterpret_cast<const PeerBuf*>(p.peers[tid]);
    const long long t0 = clock64();
    unsigned long long f;
    while ((f = ld_acquire_sys(&pb->slot[s].flag)) < e) {
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a rank is missing -- give up instead of hanging the GPU
        ok_code = 1;
        break;
      }
      __nanosleep(100);
    }
    if (ok_code == 0 && f != e) ok_code = 2;
  }
  const bool ok = __all_sync(0xffffffffu, ok_code == 0);
  const bool timed_out = __any_sync(0xffffffffu, ok_code == 1);
  // vote.sync is not a memory barrier: lane T reads values that lane L's acquire made visible, so order them with the
  // warp-level barrier (memory ordering among the participating lanes)
  __syncwarp();
  if (tid < p.n) {
    double acc = 0.0;
    for (int r = 0; r < p.world; ++r)  // rank order: the same bits on every rank
      acc += ld_relaxed_sys(&reinterpret_cast<const PeerBuf*>(p.peers[r])->slot[s].vals[tid]);
    s_sum[tid] = acc;
    p.out_partials[tid] = acc;
  }
  __syncwarp();
  if (tid == 0) {
    if (p.status && !ok) p.status[0] = timed_out ? 1 : 2;  // sticky: once a reduction failed the buffer stays marked
    p.out_loss[0] = ok ? peer_combine(p.g, s_sum, p.batch_global, p.r_box, p.r_conf, p.r_cls)
                       : __int_as_float(0x7fc00000);
  }
}

}  // namespace fvb

using namespace fvb;

extern "C" size_t fvb_peer_buffer_bytes(void) { return sizeof(PeerBuf); }

extern "C" int fvb_yolov3_loss_peer_combine_f32(const fvb_yolo_geom* geom, int64_t batch_global, const double* d_partials,
                                                const uint64_t* d_peer_ptrs, int rank, int world, float ratio_box,
                                                float ratio_conf, float ratio_cls, double* d_partials_out,
                                                float* d_out_loss, int32_t* d_status, void* stream) {
  FVB_REQUIRE(d_partials && d_peer_ptrs && d_partials_out && d_out_loss, "loss_peer_combine: NULL pointer");
  FVB_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "loss_peer_combine: rank %d of %d", rank, world);
  FVB_REQUIRE(batch_global >= 1, "loss_peer_combine: batch_global=%lld", (long long)batch_global);
  PeerParams p;
  int rc = make_geom(geom, nullptr, &p.g);
  if (rc != FVB_OK) return rc;
  p.n = p.g.L * 4;
  FVB_REQUIRE(p.n <= kPeerVals, "loss_peer_combine: %d partials", p.n);
  p.batch_global = batch_global;
  p.partials = d_partials;
  p.peers = (const unsigned long long*)d_peer_ptrs;
  p.rank = rank;
  p.world = world;
  p.r_box = ratio_box;
  p.r_conf = ratio_conf;
  p.r_cls = ratio_cls;
  p.out_partials = d_partials_out;
  p.out_loss = d_out_loss;
  p.status = d_status;
  peer_reduce_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  return check_launch("peer_reduce_combine_kernel");
}
