// Target assignment of Yolov3Loss.build_target (loss/yolov3_loss.py:75-124) for one (target, anchor) pair --
// shared by the forward (loss.cu) and backward (loss_bwd.cu) kernels.
#pragma once
#include "common.cuh"
#include "iou_grad.cuh"

namespace fvb {

struct TargetCell {
  bool match;
  int b, cls, gx, gy;
  float offx, offy, tw, th, aw, ah;
};

// build_target for one (target, anchor) on one level: loss/yolov3_loss.py:88-117.
FVB_HD TargetCell target_cell(const Geom& g, int l, const float* lab, int a) {
  TargetCell c;
  const int W = g.W[l], H = g.H[l];
  const float fw = (float)W, fh = (float)H;
  const float tx = lab[2] * fw, ty = lab[3] * fh;  // :94-95  y_true[:, 2:] * [W,H,W,H]
  c.tw = lab[4] * fw;
  c.th = lab[5] * fh;
  c.aw = g.aw[l][a] / g.stride[l];  // :88-89 anchors in feature units
  c.ah = g.ah[l][a] / g.stride[l];
  const float rw = c.tw / c.aw, rh = c.th / c.ah;  // :98
  const float m = fmaxf(fmaxf(rw, 1.0f / rw), fmaxf(rh, 1.0f / rh));
  c.match = m < 4.0f;  // :99
  c.b = (int)lab[0];
  c.cls = (int)lab[1];
  const float fx = floorf(tx), fy = floorf(ty);  // :113
  c.offx = tx - fx;                              // :114 (before the clamp)
  c.offy = ty - fy;
  // clamp in float first so that a huge coordinate cannot overflow the int conversion
  c.gx = (int)fminf(fmaxf(fx, 0.0f), (float)(W - 1));  // :116
  c.gy = (int)fminf(fmaxf(fy, 0.0f), (float)(H - 1));  // :117
  return c;
}

// first flat row of the matched cell inside the level tensor [B,A,H,W,K]
FVB_HD size_t cell_row(const Geom& g, int l, int b, int a, int gy, int gx) {
  return (((size_t)b * g.A + a) * g.H[l] + gy) * g.W[l] + gx;
}

// d/dp of -t*log(p+1e-8) - (1-t)*log(1-p+1e-8) (loss/classification_loss.py:55) and d/dt of the same expression
FVB_HD float bce_dp(float p, float t) { return (1.0f - t) / ((1.0f - p) + 1e-8f) - t / (p + 1e-8f); }
FVB_HD float bce_dt(float p) { return logf((1.0f - p) + 1e-8f) - logf(p + 1e-8f); }

// Gradient of one match w.r.t. the four box logits of its row (yolov3_loss.py:54-61): w_box * d(1 - CIoU)/dt plus
// g_tgt * d IoU/dt (the objectness target is not detached in the reference).  Also returns the IoU (the target).
struct MatchRowGrad {
  float g[4];
  float iou;
};
FVB_HD MatchRowGrad match_row_grad(float r0, float r1, float r2, float r3, const TargetCell& m, float w_box, float g_tgt) {
  const float eps = 1e-7f;
  const float px = sigmoid_precise(r0), py = sigmoid_precise(r1);
  const float pw = expf(r2) * m.aw, ph = expf(r3) * m.ah;
  const Box pb = xywh_to_xyxy(px, py, pw, ph);
  const Box tb = xywh_to_xyxy(m.offx, m.offy, m.tw, m.th);
  BoxGrad ga = zero_grad(), gb = zero_grad();
  iou_family_grad(pb, tb, FVB_CIOU, FVB_VARIANT_LIB, eps, 0.0f - w_box, ga, gb);  // mean(1 - ciou)
  MatchRowGrad out;
  out.iou = iou_plain_grad<true>(pb, tb, eps, g_tgt, ga, gb);                     // targets_conf[...] = iou
  float gx, gy, gw, gh;
  xyxy_grad_to_xywh(ga, &gx, &gy, &gw, &gh);
  out.g[0] = gx * ((1.0f - px) * px);  // sigmoid backward: grad * (1 - y) * y
  out.g[1] = gy * ((1.0f - py) * py);
  out.g[2] = gw * pw;                  // exp backward: grad * result (the anchor factor is folded into pw)
  out.g[3] = gh * ph;
  return out;
}

}  // namespace fvb
