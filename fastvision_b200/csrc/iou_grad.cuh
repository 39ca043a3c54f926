// Reverse mode of the element-wise IoU family (common.cuh: iou_family<false>) -- what torch autograd computes when
// loss.backward() runs through detection/tools/IOU.py (cal_iou :7, GIOU :193, DIOU :294, CIOU :397) from
// loss/iou_loss.py:5-107 and loss/yolov3_loss.py:58-61.  The conventions are torch's: minimum/maximum split the
// gradient 1/2 : 1/2 on ties, clamp(0) passes it where the argument is >= 0, and CIoU's alpha is a constant
// (IOU.py:436-437, no_grad).  Boxes are xyxy here; callers chain through their own xywh -> xyxy conversion.
#pragma once
#include "common.cuh"

namespace fvb {

struct BoxGrad {
  float x1, y1, x2, y2;
};

FVB_HD BoxGrad zero_grad() {
  BoxGrad g;
  g.x1 = g.y1 = g.x2 = g.y2 = 0.0f;
  return g;
}

// d min(a,b): share of the incoming gradient that goes to `a` (torch.minimum backward)
FVB_HD float min_share(float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); }
FVB_HD float max_share(float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); }

// plain IoU (inner_eps as in iou_plain); returns iou, accumulates g * d iou / d box into ga, gb
template <bool INNER_EPS>
FVB_HD float iou_plain_grad(const Box& a, const Box& b, float eps, float g, BoxGrad& ga, BoxGrad& gb,
                                                float* union_out = nullptr, float g_union_extra = 0.0f) {
  const float wa = a.x2 - a.x1, ha = a.y2 - a.y1, wb = b.x2 - b.x1, hb = b.y2 - b.y1;
  const float hae = INNER_EPS ? ha + eps : ha, hbe = INNER_EPS ? hb + eps : hb;
  const float area_a = wa * hae, area_b = wb * hbe;
  const float ix = fminf(a.x2, b.x2) - fmaxf(a.x1, b.x1), iy = fminf(a.y2, b.y2) - fmaxf(a.y1, b.y1);
  const float iw = clamp0(ix), ih = clamp0(iy);
  const float inter = iw * ih;
  const float uni = ((area_a + area_b) - inter) + eps;
  const float iou = inter / uni;
  if (union_out) *union_out = uni;
  // iou = inter / uni
  const float g_uni = (0.0f - g * iou / uni) + g_union_extra;  // extra: GIoU's penalty also reads the union
  const float g_inter = g / uni - g_uni;                       // uni = ... - inter
  const float g_iw = ix >= 0.0f ? g_inter * ih : 0.0f;
  const float g_ih = iy >= 0.0f ? g_inter * iw : 0.0f;
  // ix = min(ax2,bx2) - max(ax1,bx1)
  float s = min_share(a.x2, b.x2);
  ga.x2 += g_iw * s;
  gb.x2 += g_iw * (1.0f - s);
  s = max_share(a.x1, b.x1);
  ga.x1 -= g_iw * s;
  gb.x1 -= g_iw * (1.0f - s);
  s = min_share(a.y2, b.y2);
  ga.y2 += g_ih * s;
  gb.y2 += g_ih * (1.0f - s);
  s = max_share(a.y1, b.y1);
  ga.y1 -= g_ih * s;
  gb.y1 -= g_ih * (1.0f - s);
  // areas
  const float g_wa = g_uni * hae, g_ha = g_uni * wa, g_wb = g_uni * hbe, g_hb = g_uni * wb;
  ga.x2 += g_wa; ga.x1 -= g_wa; ga.y2 += g_ha; ga.y1 -= g_ha;
  gb.x2 += g_wb; gb.x1 -= g_wb; gb.y2 += g_hb; gb.y1 -= g_hb;
  return iou;
}

// Element form of the family: returns the value and ACCUMULATES g * d value / d box into ga, gb.
FVB_HD float iou_family_grad(const Box& a, const Box& b, int kind, int variant, float eps, float g,
                                                 BoxGrad& ga, BoxGrad& gb) {
  if (kind == FVB_IOU) return iou_plain_grad<true>(a, b, eps, g, ga, gb);
  const float cw = fmaxf(a.x2, b.x2) - fminf(a.x1, b.x1);
  const float ch = fmaxf(a.y2, b.y2) - fminf(a.y1, b.y1);
  float g_cw = 0.0f, g_ch = 0.0f, value;
  if (kind == FVB_GIOU) {
    // value = iou - pen, pen = (convex - uni) / convex  (IOU.py:220-239)
    const float convex = cw * ch + eps;
    // pen needs uni before the gradient of iou is formed: evaluate the union first
    float uni;
    {
      const float area_a = (a.x2 - a.x1) * (a.y2 - a.y1), area_b = (b.x2 - b.x1) * (b.y2 - b.y1);
      uni = ((area_a + area_b) - inter_area(a, b)) + eps;
    }
    const float pen = (convex - uni) / convex;
    const float g_pen = 0.0f - g;
    const float g_convex = g_pen * (uni / (convex * convex));  // d pen / d convex = uni / convex^2
    const float g_uni_extra = 0.0f - g_pen / convex;           // d pen / d uni = -1 / convex
    const float iou = iou_plain_grad<false>(a, b, eps, g, ga, gb, nullptr, g_uni_extra);
    g_cw = g_convex * ch;
    g_ch = g_convex * cw;
    value = iou - pen;
  } else {
    const float iou = iou_plain_grad<true>(a, b, eps, g, ga, gb);
    const float c2 = (cw * cw + ch * ch) + eps;
    const bool demo = variant == FVB_VARIANT_DEMO;
    float dx, dy;
    if (demo) {
      dx = (a.x1 + a.x2) - (b.x1 + b.x2);
      dy = (a.y1 + a.y2) - (b.y1 + b.y2);
    } else {
      dx = (a.x1 + a.x2) * 0.5f - (b.x1 + b.x2) * 0.5f;
      dy = (a.y1 + a.y2) * 0.5f - (b.y1 + b.y2) * 0.5f;
    }
    const float rho2 = dx * dx + dy * dy;
    const float pen = rho2 / c2;
    const float g_pen = demo ? 0.0f - g : g;  // lib: iou + pen (IOU.py:341); demo: iou - pen
    const float g_rho2 = g_pen / c2;
    const float g_c2 = 0.0f - g_pen * pen / c2;
    g_cw = g_c2 * (2.0f * cw);
    g_ch = g_c2 * (2.0f * ch);
    const float cs = demo ? 1.0f : 0.5f;
    const float g_dx = g_rho2 * (2.0f * dx) * cs, g_dy = g_rho2 * (2.0f * dy) * cs;
    ga.x1 += g_dx; ga.x2 += g_dx; gb.x1 -= g_dx; gb.x2 -= g_dx;
    ga.y1 += g_dy; ga.y2 += g_dy; gb.y1 -= g_dy; gb.y2 -= g_dy;
    value = demo ? iou - pen : iou + pen;
    if (kind == FVB_CIOU) {
      const float w1 = a.x2 - a.x1, h1 = a.y2 - a.y1, w2 = b.x2 - b.x1, h2 = b.y2 - b.y1;
      const float h1e = h1 + eps, h2e = h2 + eps;
      const float q1 = w1 / h1e, q2 = w2 / h2e;
      const float d = atanf(q2) - atanf(q1);
      const float four_over_pi2 = 0.4052847345693511f;
      const float v = four_over_pi2 * (d * d);
      const float alpha = v / ((v - iou) + (1.0f + eps));  // constant for autograd
      const float g_v = 0.0f - alpha * g;
      const float g_d = g_v * four_over_pi2 * (2.0f * d);
      const float g_q2 = g_d / (1.0f + q2 * q2), g_q1 = (0.0f - g_d) / (1.0f + q1 * q1);
      const float g_w1 = g_q1 / h1e, g_h1 = 0.0f - g_q1 * q1 / h1e;
      const float g_w2 = g_q2 / h2e, g_h2 = 0.0f - g_q2 * q2 / h2e;
      ga.x2 += g_w1; ga.x1 -= g_w1; ga.y2 += g_h1; ga.y1 -= g_h1;
      gb.x2 += g_w2; gb.x1 -= g_w2; gb.y2 += g_h2; gb.y1 -= g_h2;
      value = value - alpha * v;
    }
  }
  // cw = max(ax2,bx2) - min(ax1,bx1)
  float s = max_share(a.x2, b.x2);
  ga.x2 += g_cw * s;
  gb.x2 += g_cw * (1.0f - s);
  s = min_share(a.x1, b.x1);
  ga.x1 -= g_cw * s;
  gb.x1 -= g_cw * (1.0f - s);
  s = max_share(a.y2, b.y2);
  ga.y2 += g_ch * s;
  gb.y2 += g_ch * (1.0f - s);
  s = min_share(a.y1, b.y1);
  ga.y1 -= g_ch * s;
  gb.y1 -= g_ch * (1.0f - s);
  return value;
}

// chain an xyxy gradient back through xywh_to_xyxy (BOX.py:4-10): x1 = x - w/2, x2 = x + w/2
FVB_HD void xyxy_grad_to_xywh(const BoxGrad& g, float* gx, float* gy, float* gw, float* gh) {
  *gx = g.x1 + g.x2;
  *gy = g.y1 + g.y2;
  *gw = (g.x2 - g.x1) / 2.0f;
  *gh = (g.y2 - g.y1) / 2.0f;
}

}  // namespace fvb
