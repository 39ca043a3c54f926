// RPN proposal filter (demos/faster_rcnn/models/rpn.py:168-208) -- implemented in a later milestone.
#include "common.cuh"

extern "C" size_t fvb_rpn_workspace_bytes(int batch, int height, int width, int anchors) {
  (void)batch; (void)height; (void)width; (void)anchors;
  return 256;
}

extern "C" int fvb_rpn_proposals_f32(const float* d_cls, const float* d_reg, const float* base_anchors, int batch,
                                     int height, int width, int anchors, int pre_n, int post_n, double iou_thr,
                                     float* d_out_xywh, int32_t* d_out_cnt, void* d_ws, void* stream) {
  (void)d_cls; (void)d_reg; (void)base_anchors; (void)batch; (void)height; (void)width; (void)anchors;
  (void)pre_n; (void)post_n; (void)iou_thr; (void)d_out_xywh; (void)d_out_cnt; (void)d_ws; (void)stream;
  fvb::set_error("rpn_proposals: not implemented yet");
  return FVB_E_LIMIT;
}
