// TEST INFRASTRUCTURE ONLY: host-side entry points over the FVB_HD arithmetic of fastvision_b200/csrc (the same source the
// kernels compile), so the build container (no GPU) can check the hand-derived gradients against torch autograd.
// Built on demand by tests/test_host_arith.py:  nvcc -shared -Xcompiler -fPIC tests/host_check.cu -o tests/_host_check.so
#include "../fastvision_b200/csrc/iou_grad.cuh"
#include "../fastvision_b200/csrc/loss_common.cuh"

using namespace fvb;

extern "C" void hc_iou_family(const float* a, const float* b, int n, int box_mode, int kind, int variant, float eps,
                              const float* g, float* value, float* ga, float* gb) {
  for (int i = 0; i < n; ++i) {
    const Box ba = load_box(a + i * 4, box_mode), bb = load_box(b + i * 4, box_mode);
    BoxGrad qa = zero_grad(), qb = zero_grad();
    value[i] = iou_family_grad(ba, bb, kind, variant, eps, g[i], qa, qb);
    const float fwd = iou_family<false>(ba, bb, kind, variant, eps);
    if (fwd != value[i]) value[i] = __builtin_nanf("");  // the backward's forward value must be the forward's, bit for bit
    for (int s = 0; s < 2; ++s) {
      float* o = (s == 0 ? ga : gb) + i * 4;
      const BoxGrad& q = s == 0 ? qa : qb;
      if (box_mode == FVB_BOX_XYWH) xyxy_grad_to_xywh(q, o, o + 1, o + 2, o + 3);
      else { o[0] = q.x1; o[1] = q.y1; o[2] = q.x2; o[3] = q.y2; }
    }
  }
}

// one matched row: logits r[0..3], target (offx, offy, tw, th), anchor (aw, ah) in feature units
extern "C" void hc_match_row_grad(const float* r, const float* tgt, const float* anchor, int n, float w_box,
                                  const float* g_tgt, float* grad, float* iou) {
  for (int i = 0; i < n; ++i) {
    TargetCell m;
    m.match = true; m.b = 0; m.cls = 0; m.gx = 0; m.gy = 0;
    m.offx = tgt[i * 4]; m.offy = tgt[i * 4 + 1]; m.tw = tgt[i * 4 + 2]; m.th = tgt[i * 4 + 3];
    m.aw = anchor[i * 2]; m.ah = anchor[i * 2 + 1];
    const MatchRowGrad q = match_row_grad(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3], m, w_box, g_tgt[i]);
    for (int k = 0; k < 4; ++k) grad[i * 4 + k] = q.g[k];
    iou[i] = q.iou;
  }
}

extern "C" void hc_bce(const float* p, const float* t, int n, float* value, float* dp, float* dt) {
  for (int i = 0; i < n; ++i) {
    value[i] = bce_term(p[i], t[i]);
    dp[i] = bce_dp(p[i], t[i]);
    dt[i] = bce_dt(p[i]);
  }
}
