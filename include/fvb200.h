/*
 * fvb200.h -- C ABI of libfvb200.so, the B200 (sm_100a) detection hot path of fastvision.
 *
 * The reference (ielym/fastvision) is pure Python: it has no FFI or operator registry; its
 * boundary for this path is a set of Python callables.  Each entry point
 * below names the reference callable it replaces (file:line under the reference root); the
 * Python host side in fastvision_b200/ keeps the reference's names and signatures and binds
 * these symbols with ctypes (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No C++ / torch types cross the ABI.
 *   - Pointers named d_* are DEVICE pointers (current device); everything else is host memory
 *     that is read during the call only (small geometry tables are passed to kernels by value).
 *   - The caller owns every buffer (inputs, outputs, workspaces).  The library never allocates,
 *     frees or retains a pointer past the call.  Tensors are contiguous fp32 unless stated.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no host sync inside.
 *   - Return 0 on success, negative FVB_E_* otherwise; fvb_last_error() gives a thread-local
 *     message.  Launch errors are picked up with cudaPeekAtLastError only.
 *   - There is no CPU path: without a CUDA device every compute entry point fails.
 */
#ifndef FVB200_H_
#define FVB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVB_ABI_VERSION 2 /* 2 = 1 + the *_sync / *_after_decode / *_dense entry points (additive: every v1 signature is unchanged) */
#define FVB_MAX_LEVELS 4
#define FVB_MAX_ANCHORS 16

#define FVB_OK 0
#define FVB_E_INVALID (-1) /* bad argument */
#define FVB_E_CUDA (-2)    /* CUDA runtime / launch error */
#define FVB_E_LIMIT (-3)   /* shape beyond a compiled-in limit */

/* box layouts (detection/tools/IOU.py:7-25 `mode`) */
#define FVB_BOX_XYXY 0
#define FVB_BOX_XYWH 1
#define FVB_BOX_WH 2
/* IoU kinds */
#define FVB_IOU 0
#define FVB_GIOU 1
#define FVB_DIOU 2
#define FVB_CIOU 3
/* arithmetic variant: library files vs demos/<x>/utils/iou.py */
#define FVB_VARIANT_LIB 0
#define FVB_VARIANT_DEMO 1
/* decode forms: detection/models/yolov3.py:47-48 vs demos/yolov3_u/inference.py:86-89 */
#define FVB_DECODE_V3 0
#define FVB_DECODE_V5 1
/* NMS front-ends */
#define FVB_NMS_LIB 0        /* detection/tools/NMS.py:5-23, class-agnostic, rank by max_c(cls*obj), boxes xywh */
#define FVB_NMS_DEMO 1       /* demos/yolov3_u/utils/nms.py:5-53, class-aware gap trick, rank by obj, boxes xyxy */
#define FVB_NMS_DEMO_BATCH 2 /* demos/yolov3_u/utils/nms.py:55-98, class-aware, rank by max_c(cls*obj), boxes xywh */
/* flag bits of the `clear_bitmap` argument of the fvb_yolo_nms_* calls (1 = the v1 meaning) */
#define FVB_NMS_CLEAR_BITMAP 1 /* zero the candidate words the kernel consumed (ready for the next decode) */
#define FVB_NMS_WIDE_CTA 2     /* fvb_yolo_nms_after_decode_f32 only: fvb_yolo_decode_leaves_room_for_nms() said 0 */
/* reductions */
#define FVB_REDUCE_MEAN 0
#define FVB_REDUCE_SUM 1

int fvb_abi_version(void);
const char* fvb_last_error(void);
/* Number of kernel launches this process has enqueued through the library (bench.py: gpu_launches). */
uint64_t fvb_launch_count(void);

/* raw head layouts */
#define FVB_HEAD_BAHWK 0 /* [B,A,H,W,K] contiguous, detection/head/yolov3head.py:63 (what Yolov3.forward decodes) */
#define FVB_HEAD_NCHW 1  /* [B,A*K,H,W] contiguous: the conv output itself, channel a*K+k (demos/yolov3_huaweiShip/inference.py:107,
                            customize_service.py:437) -- decode only; the loss entry points need FVB_HEAD_BAHWK */

/* Geometry of a YOLO head stack, host memory.  Level order = concat order (stride 32,16,8). */
typedef struct {
  int32_t levels;   /* L <= FVB_MAX_LEVELS */
  int32_t batch;    /* B */
  int32_t anchors;  /* A per level <= FVB_MAX_ANCHORS */
  int32_t channels; /* K = 5 + C */
  int32_t height[FVB_MAX_LEVELS];
  int32_t width[FVB_MAX_LEVELS];
  float stride[FVB_MAX_LEVELS];
  float anchor_w[FVB_MAX_LEVELS][FVB_MAX_ANCHORS]; /* pixels */
  float anchor_h[FVB_MAX_LEVELS][FVB_MAX_ANCHORS];
  int32_t head_layout; /* FVB_HEAD_* */
} fvb_yolo_geom;

/* ---- K1 decode ---------------------------------------------------------------------------
 * Replaces the decode block of Yolov3.forward, detection/models/yolov3.py:33-53 (and the demo
 * forms, demos/yolov3_u/inference.py:86-89).  d_heads[l] is the contiguous raw head tensor of level l in
 * geom->head_layout ([B,A,H_l,W_l,K], detection/head/yolov3head.py:63, or the conv output [B,A*K,H_l,W_l] whose
 * permute().contiguous() copy the kernel then saves); d_results is [B,N,K] with
 * N = A*sum(H_l*W_l), row a*H*W + y*W + x inside a level.
 * Optional fused side outputs (either may be NULL):
 *   d_cand_bitmap  [B, fvb_yolo_bitmap_words(geom)] u32, must be zero on entry: bit r of image b
 *                  is set iff results[b,r,4] > conf_thr -- the candidate set of
 *                  non_max_suppression (detection/tools/NMS.py:7-8) without a second pass;
 *   d_cand_rec     [B, N, 8] f32, 16-byte aligned (needs d_cand_bitmap): for every candidate row r the record
 *                  {results[b,r,0..3], conf, max_c(cls_c*conf), argmax_c as int bits, unused} -- what
 *                  NMS.py:13-16 computes per candidate, produced while the row is in registers so the
 *                  NMS kernel never re-reads the 4*K-byte rows; non-candidate records are not written;
 *   d_conf_bce0    [fvb_yolo_decode_partials(geom)] f64: one partial sum per decode tile (fixed reduction tree,
 *                  so reproducible) of -log(1 - conf + 1e-8), the zero-target part of the objectness BCE of
 *                  Yolov3Loss (loss/yolov3_loss.py:63-64), consumed by fvb_yolov3_loss_f32.
 * d_ws: fvb_yolo_decode_workspace_bytes() bytes (the tile queue of the persistent kernel), 8-byte aligned, ZERO
 * before the first call; every launch leaves it zero again.  One workspace per concurrently running decode.
 * precise != 0 uses expf + IEEE division instead of ex2.approx/rcp.approx (both meet rtol 1e-5).
 */
int fvb_yolo_rows_per_image(const fvb_yolo_geom* geom);
int fvb_yolo_bitmap_words(const fvb_yolo_geom* geom);
int fvb_yolo_decode_partials(const fvb_yolo_geom* geom); /* doubles in d_conf_bce0 */
size_t fvb_yolo_decode_workspace_bytes(void);
int fvb_yolo_decode_f32(const fvb_yolo_geom* geom, const float* const* d_heads, int form, int precise,
                        float* d_results, float conf_thr, uint32_t* d_cand_bitmap, float* d_cand_rec,
                        double* d_conf_bce0, void* d_ws, void* stream);
/* The same decode, additionally publishing its progress per image so that the per-image body of Fit._val (utils/fit.py:94-95:
 * non_max_suppression of image i) can start while later images are still being decoded.  d_tile_sync: [B + 1] u32, ZERO
 * before the first call; entry b counts the finished tiles of image b (release semantics: all of the image's results, candidate
 * bits, records and objectness partials are visible once it reads fvb_yolo_decode_tiles_per_image(geom)); the consumer,
 * fvb_yolo_nms_after_decode_f32, must be the NEXT launch on `stream` and leaves the array zero again.  NULL = plain decode. */
int fvb_yolo_decode_tiles_per_image(const fvb_yolo_geom* geom);
/* 1 if an NMS CTA fits on an SM beside a decode CTA of this geometry (registers / shared memory), i.e. if the pair below really
 * overlaps; 0 if not (narrow rows run 24 decode warps: the pair is then correct but no faster than the plain calls); < 0 on error. */
int fvb_yolo_decode_leaves_room_for_nms(const fvb_yolo_geom* geom);
int fvb_yolo_decode_sync_f32(const fvb_yolo_geom* geom, const float* const* d_heads, int form, int precise,
                             float* d_results, float conf_thr, uint32_t* d_cand_bitmap, float* d_cand_rec,
                             double* d_conf_bce0, uint32_t* d_tile_sync, void* d_ws, void* stream);

/* ---- box conversion ------------------------------------------------------------------------
 * detection/tools/BOX.py:4-26.  op: 0 xywh2xyxy, 1 xyxy2xywh, 2 xyxy2xywhn (needs height,width).
 */
int fvb_box_convert_f32(const float* d_in, int64_t n, int op, float height, float width, float* d_out,
                        void* stream);

/* ---- K2 IoU family -------------------------------------------------------------------------
 * detection/tools/IOU.py: cal_iou :7 / GIOU :193 / DIOU :294 / CIOU :397 (element-wise, out[n])
 * and cal_iou_batch :17 / GIOU_batch :243 / DIOU_batch :345 / CIOU_batch :442 (pairwise,
 * out[N,M] row-major).  Bug-compatible with the torch branches.  box_mode WH takes
 * [n,2] inputs and supports kind FVB_IOU only.
 */
int fvb_iou_elementwise_f32(const float* d_a, const float* d_b, int64_t n, int box_mode, int kind,
                            int variant, float eps, float* d_out, void* stream);
int fvb_iou_pairwise_f32(const float* d_a, int64_t n, const float* d_b, int64_t m, int box_mode, int kind,
                         int variant, float eps, float* d_out, void* stream);
/* loss/iou_loss.py:5-107: out[0] = reduce((1 - kind(a,b)) * w); d_weights may be NULL.
 * d_ws: workspace of fvb_reduce_workspace_bytes(n) bytes. */
size_t fvb_reduce_workspace_bytes(int64_t n);
int fvb_iou_loss_f32(const float* d_pre, const float* d_true, const float* d_weights, int64_t n, int box_mode,
                     int kind, int variant, float eps, int reduction, float* d_out, void* d_ws, void* stream);
/* loss/classification_loss.py:36-65 BiCrossEntropyLoss.  d_pre is [rows, C]; when C > 1 d_target_idx
 * ([rows] int64 class ids, one-hot expanded on the fly) is used, otherwise d_target_val ([rows] fp32). */
int fvb_bce_loss_f32(const float* d_pre, int64_t rows, int classes, const int64_t* d_target_idx,
                     const float* d_target_val, int already_sigmoid, const float* d_weights, int reduction,
                     float* d_out, void* d_ws, void* stream);

/* ---- demo box post-processing -------------------------------------------------------------------------------------
 * The part of the demos' postProcess between decode and NMS (demos/yolov3_u/inference.py:92-109, identical in
 * demos/yolov3_huaweiShip/inference.py:112-129), in place on decoded rows [n_rows, channels] (xywh in channels 0..3):
 * x,y,w,h -> ((x - pad_left)/ratio, (y - pad_top)/ratio, w/ratio, h/ratio); clamp to the original image; rows with
 * w <= min_wh or h <= min_wh (5 px in the demos) are marked dropped (objectness := -1, they keep their slot so the order
 * of the others is unchanged); xywh -> xyxy; clamp to [0, ori-1].  Follow with fvb_yolo_nms_f32(flavour FVB_NMS_DEMO). */
int fvb_demo_boxes_postprocess_f32(float* d_rows, int64_t n_rows, int channels, float pad_left, float pad_top,
                                   float resize_ratio, float ori_width, float ori_height, float min_wh, void* stream);

/* ---- K3 NMS --------------------------------------------------------------------------------
 * fvb_nms_segmented_f32: the batched equivalent of torchvision.ops.nms (third party; call sites
 * detection/tools/NMS.py:18, demos/yolov3_u/utils/nms.py:47,92, demos/faster_rcnn/models/rpn.py:198).
 * Segment s owns boxes[seg_offsets[s] .. seg_offsets[s+1]) (xyxy) and scores; writes up to max_keep
 * kept indices (relative to the segment start, score-descending, ties by lower index) into
 * d_keep_idx[s*max_keep ..] and the count into d_keep_cnt[s].  iou_thr is a double because the
 * CPU op compares the fp32 ratio against the double threshold; the kernels compare in fp32 against
 * the largest float <= iou_thr, which decides identically.
 */
size_t fvb_nms_segmented_workspace_bytes(int64_t total_boxes, int segments);
int fvb_nms_segmented_f32(const float* d_boxes, const float* d_scores, const int32_t* d_seg_offsets,
                          int segments, int64_t total_boxes, double iou_thr, int max_keep, int32_t* d_keep_idx,
                          int32_t* d_keep_cnt, void* d_ws, void* stream);

/* fvb_yolo_nms_f32: non_max_suppression for every image of a decoded [B,N,K] tensor in one launch.
 * Replaces detection/tools/NMS.py:5-23 (flavour LIB) and demos/yolov3_u/utils/nms.py:5-53 / :55-98.
 * d_cand_bitmap + d_cand_rec: the candidate bitmap [B, words] and records [B,N,8] produced by
 * fvb_yolo_decode_f32 for the same conf_thr, or both NULL to have a scoring kernel build them from
 * results (one extra launch that reads the objectness channel and the candidate rows).  If clear_bitmap != 0 the kernel zeroes the words
 * it consumed (ready for the next decode).  Outputs are padded: d_out_boxes [B,max_det,4] xyxy (not
 * gap-offset), d_out_scores [B,max_det], d_out_cls [B,max_det] int64, d_out_rows [B,max_det] int32
 * (row index inside the image, or NULL), d_out_cnt [B] int32.
 */
size_t fvb_yolo_nms_workspace_bytes(int batch, int rows_per_image);
int fvb_yolo_nms_f32(const float* d_results, int batch, int rows_per_image, int channels, float conf_thr,
                     double iou_thr, int max_det, int flavour, float max_wh, uint32_t* d_cand_bitmap,
                     const float* d_cand_rec, int clear_bitmap, float* d_out_boxes, float* d_out_scores, int64_t* d_out_cls,
                     int32_t* d_out_rows, int32_t* d_out_cnt, void* d_ws, void* stream);
/* The same call as the consumer of fvb_yolo_decode_sync_f32 (same d_tile_sync, tiles_per_image =
 * fvb_yolo_decode_tiles_per_image(geom), same d_cand_bitmap / d_cand_rec, enqueued directly behind it on the same stream): the
 * kernel is launched as a PROGRAMMATIC DEPENDENT of the decode kernel, so image b's NMS starts as soon as image b is decoded
 * and only the images decoded last remain when the decode kernel exits.  Results are identical to fvb_yolo_nms_f32.  When no NMS
 * CTA fits beside a decode CTA (fvb_yolo_decode_leaves_room_for_nms() == 0) the pair still saves the launch gap for batches of at
 * most one image per SM: pass FVB_NMS_WIDE_CTA in clear_bitmap and the CTAs (1024 threads then) start SM by SM as decode CTAs leave.  If the
 * decode launch is missing the kernel gives up after ~4 s and writes d_out_cnt[b] = -1 (it never hangs the device). */
int fvb_yolo_nms_after_decode_f32(const float* d_results, int batch, int rows_per_image, int channels, float conf_thr,
                                  double iou_thr, int max_det, int flavour, float max_wh, uint32_t* d_cand_bitmap,
                                  const float* d_cand_rec, int clear_bitmap, float* d_out_boxes, float* d_out_scores,
                                  int64_t* d_out_cls, int32_t* d_out_rows, int32_t* d_out_cnt, uint32_t* d_tile_sync,
                                  int tiles_per_image, void* d_ws, void* stream);

/* fvb_rpn_proposals_f32: RPN.filter_proposals, demos/faster_rcnn/models/rpn.py:168-208 (+ :111-119,
 * :160-166).  d_cls [B,H,W,A,2], d_reg [B,H,W,A,4], base_anchors (host) [A,2] (w,h) feature units.
 * The reference's per-image topk(pre_n) -> nms(iou_thr) -> first post_n runs for all images in one launch.
 * Outputs (padded, entries past d_out_cnt[b] undefined): d_out_xywh [B,post_n,4] feature units, d_out_idx
 * [B,post_n] int32 = flat (h,w,a) anchor index of each proposal (may be NULL), d_out_cnt [B].
 * Equal scores rank by lower anchor index (torch.topk leaves the order of ties unspecified). */
size_t fvb_rpn_workspace_bytes(int batch, int height, int width, int anchors);
int fvb_rpn_proposals_f32(const float* d_cls, const float* d_reg, const float* base_anchors, int batch,
                          int height, int width, int anchors, int pre_n, int post_n, double iou_thr,
                          float* d_out_xywh, int32_t* d_out_idx, int32_t* d_out_cnt, void* d_ws, void* stream);

/* ---- K4 target assignment + loss ---------------------------------------------------------------
 * fvb_yolov3_loss_f32 replaces Yolov3Loss.forward, loss/yolov3_loss.py:29-72 (with build_target
 * :75-124, CIOULoss loss/iou_loss.py:83-107, BiCrossEntropyLoss loss/classification_loss.py:36-65).
 * d_labels [T,6] = [batch_idx, cls, xc, yc, w, h] normalised.  d_conf_bce0 = per-warp zero-target
 * objectness sums from fvb_yolo_decode_f32 over the same heads, or NULL (the call then streams
 * channel 4 itself).  d_partials [L*4] f64 receives per level {S_cls, S_box, S_conf, M}
 * (what a data-parallel run all-reduces).  If d_out_loss != NULL the scalar
 *   B * sum_l ( r_box*S_box/M + r_cls*S_cls/(M*C) + r_conf*S_conf/(B*A*H*W) )   is written to it.
 */
size_t fvb_yolov3_loss_workspace_bytes(const fvb_yolo_geom* geom, int64_t num_labels);
int fvb_yolov3_loss_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                        int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                        const double* d_conf_bce0, double* d_partials, float* d_out_loss, void* d_ws,
                        void* stream);
/* The same loss in two parts.  Target assignment and the matched-row terms (loss/yolov3_loss.py:44-61, :75-124) read only the raw
 * heads and the labels: fvb_yolov3_loss_match_f32 may be enqueued beside fvb_yolo_decode_f32 and hides under it.  What needs the
 * decode is only the sum of its objectness partials (:63-64): fvb_yolov3_loss_finish_f32 -- same d_ws and num_labels, ordered
 * after both -- does that sum, reduces the matched terms and writes d_partials / d_out_loss as fvb_yolov3_loss_f32 would
 * (equal up to the association of the fp64 sums; each form is bit-reproducible). */
int fvb_yolov3_loss_match_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                              int64_t num_labels, void* d_ws, void* stream);
int fvb_yolov3_loss_finish_f32(const fvb_yolo_geom* geom, int64_t num_labels, float ratio_box, float ratio_conf,
                               float ratio_cls, const double* d_conf_bce0, double* d_partials, float* d_out_loss,
                               void* d_ws, void* stream);
/* v2: forms for objectness partials that came from fvb_yolo_decode_f32(..., precise = 1): the matched cells then take the
 * dense term back out with the same (precise) sigmoid the decode used -- for objectness logits above ~13 one ulp of p is
 * ~20 % of -log(1 - p + 1e-8), so the two forms must agree bit for bit to cancel.  conf_bce0_precise = 0 is the v1 call. */
int fvb_yolov3_loss_dense_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                              int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                              const double* d_conf_bce0, int conf_bce0_precise, double* d_partials, float* d_out_loss,
                              void* d_ws, void* stream);
int fvb_yolov3_loss_match_dense_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                    int64_t num_labels, int conf_bce0_precise, void* d_ws, void* stream);
/* Training form of the same call: additionally writes d_saved_conf ([fvb_yolov3_saved_conf_floats(geom)] f32, may be
 * NULL), a compact level-major copy [l][b][row] of the objectness logits, which fvb_yolov3_loss_backward_f32 then reads
 * instead of striding through the heads again.  With d_saved_conf the call always streams channel 4 itself. */
int64_t fvb_yolov3_saved_conf_floats(const fvb_yolo_geom* geom);
int fvb_yolov3_loss_train_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                              int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                              const double* d_conf_bce0, double* d_partials, float* d_out_loss, float* d_saved_conf,
                              void* d_ws, void* stream);
/* Combine (all-reduced) partials into the scalar; batch_global/cells use the GLOBAL batch. */
int fvb_yolov3_loss_combine_f32(const fvb_yolo_geom* geom, int64_t batch_global, const double* d_partials,
                                float ratio_box, float ratio_conf, float ratio_cls, float* d_out_loss,
                                void* stream);
/* Data-parallel finish over NVLink peer memory: all-reduce of the [L*4] partials over `world` ranks fused with
 * the combine, one single-CTA kernel (replaces ncclAllReduce(96 B) + fvb_yolov3_loss_combine_f32; capturable in a CUDA graph).
 * d_peer_ptrs: device array [world] of the base addresses of every rank's peer buffer (fvb_peer_buffer_bytes() bytes, ZEROED
 * once before first use, mapped into this process -- torch symmetric memory provides them).  Every rank must make the same
 * sequence of calls on one buffer set.  d_partials_out [L*4] receives the global sums (same bits on all ranks), d_out_loss the
 * scalar; d_status (may be NULL) 0, or 1 if a peer did not arrive within ~10 s (loss := NaN, no hang). */
size_t fvb_peer_buffer_bytes(void);
int fvb_yolov3_loss_peer_combine_f32(const fvb_yolo_geom* geom, int64_t batch_global, const double* d_partials,
                                     const uint64_t* d_peer_ptrs, int rank, int world, float ratio_box, float ratio_conf,
                                     float ratio_cls, double* d_partials_out, float* d_out_loss, int32_t* d_status,
                                     void* stream);
/* Yolov3Loss.build_target, loss/yolov3_loss.py:75-124, for one level: padded outputs of T*A rows in
 * (t,a) row-major order with d_match[T*A] u8 flags; d_count[1] int32 = M.  Compacted (matches first,
 * order kept) when compact != 0. */
int fvb_yolov3_build_target_f32(const fvb_yolo_geom* geom, int level, const float* d_labels, int64_t num_labels,
                                int compact, int64_t* d_b, int64_t* d_gxy, int64_t* d_a, int64_t* d_cls,
                                float* d_xywh, float* d_anchor, uint8_t* d_match, int32_t* d_count,
                                void* stream);

/* ---- K4' backward (what loss.backward() in Fit._train, utils/fit.py:57-63, makes autograd do) ---------
 * fvb_yolov3_loss_backward_f32: gradients of Yolov3Loss.forward (loss/yolov3_loss.py:29-72) w.r.t. the raw head
 * tensors.  d_grad_heads[l] has the layout of d_heads[l] ([B,A,H,W,K]) and is written completely (no pre-zeroing):
 * channel 4 of every cell gets the objectness-BCE gradient (:63-64), matched rows additionally get the class-BCE
 * (:50-52), CIoU (:54-58, alpha constant as in detection/tools/IOU.py:436-437) and IoU-target (:60-61, the reference
 * does not detach targets_conf) gradients; duplicate matches of one cell accumulate as torch's index / index_put
 * backward do.  d_partials: the [L*4] f64 partials of the forward (M_l at [l*4+3]; all-reduced under data
 * parallelism, with batch_global the global batch).  d_saved_conf: the compact objectness logits written by
 * fvb_yolov3_loss_train_f32 over the same heads, or NULL.  d_grad_out: device scalar [1] (the upstream gradient) or NULL
 * for 1.  d_ws: fvb_yolov3_loss_backward_workspace_bytes(geom, num_labels) bytes, 256-byte aligned.  Bit-reproducible (one writer per row, no atomics).
 */
size_t fvb_yolov3_loss_backward_workspace_bytes(const fvb_yolo_geom* geom, int64_t num_labels);
int fvb_yolov3_loss_backward_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                                 int64_t num_labels, float ratio_box, float ratio_conf, float ratio_cls,
                                 int64_t batch_global, const double* d_partials, const float* d_saved_conf,
                                 const float* d_grad_out, float* const* d_grad_heads, void* d_ws, void* stream);
/* Gradients of fvb_iou_loss_f32 (loss/iou_loss.py:5-107) w.r.t. y_pre and/or y_true (either output may be NULL);
 * same box_mode layout as the inputs ([n,4], or [n,2] for wh).  torch conventions: minimum/maximum split ties 1/2,
 * clamp(0) passes the gradient at 0.  d_ws: >= 256 bytes. */
int fvb_iou_loss_backward_f32(const float* d_pre, const float* d_true, const float* d_weights, int64_t n, int box_mode,
                              int kind, int variant, float eps, int reduction, const float* d_grad_out,
                              float* d_grad_pre, float* d_grad_true, void* d_ws, void* stream);
/* Gradient of fvb_bce_loss_f32 (loss/classification_loss.py:36-65) w.r.t. y_pre, [rows, C]. */
int fvb_bce_loss_backward_f32(const float* d_pre, int64_t rows, int classes, const int64_t* d_target_idx,
                              const float* d_target_val, int already_sigmoid, const float* d_weights, int reduction,
                              const float* d_grad_out, float* d_grad_pre, void* stream);

/* ---- demo training loss ---------------------------------------------------------------------------
 * ComputeLoss.forward(predict_layers, target_all, model): demos/yolov3_huaweiShip/utils/lossv3.py:19-125 (flavour SHIP,
 * d_out[3] = loss_box, loss_cls, loss_conf) and demos/yolov3_u/utils/lossv3.py:17-119 (flavour U, d_out[0] = the scalar).
 * d_heads[l] = the conv outputs [B, A*K, H_l, W_l] (geom->head_layout must be FVB_HEAD_NCHW); anchors in feature units
 * are geom->anchor_{w,h} / geom->stride (model.anchors with stride 1).  d_labels [T,6] = [batch_idx, cls, xc, yc, w, h].
 * d_partials [L*6] f64 per level {S_box|S_xy, S_wh, S_cls, S_conf, n_valid, T}: what data-parallel ranks all-reduce before
 * fvb_demo_loss_combine_f32.  d_mask (optional, [fvb_demo_loss_mask_bytes] int8, level-major [l][b][a][cell]): the
 * objectness mask (-1 ignore, 0 negative, 1 positive), needed by the backward.  A target whose centre leaves the feature map
 * is clamped to the border cell (the reference indexes out of range); an image without targets has no ignore region
 * (the reference raises IndexError at lossv3.py:107 -- the Python wrapper reproduces that when strict). */
#define FVB_DEMO_LOSS_SHIP 0
#define FVB_DEMO_LOSS_U 1
size_t fvb_demo_loss_workspace_bytes(const fvb_yolo_geom* geom, int64_t num_labels);
int64_t fvb_demo_loss_mask_bytes(const fvb_yolo_geom* geom);
int fvb_demo_loss_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels, int64_t num_labels,
                      int flavour, double* d_partials, float* d_out, int8_t* d_mask, void* d_ws, void* stream);
int fvb_demo_loss_combine_f32(const fvb_yolo_geom* geom, const double* d_partials, int flavour, float* d_out, void* stream);
/* Gradients w.r.t. the conv outputs (d_grad_heads[l] like d_heads[l], written completely).  d_grad_out: [3] (SHIP: upstream
 * gradients of the three outputs) or [1] (U), or NULL for ones.  d_ws: fvb_demo_loss_workspace_bytes. */
int fvb_demo_loss_backward_f32(const fvb_yolo_geom* geom, const float* const* d_heads, const float* d_labels,
                               int64_t num_labels, int flavour, const double* d_partials, const int8_t* d_mask,
                               const float* d_grad_out, float* const* d_grad_heads, void* d_ws, void* stream);

/* ---- K5 mAP matcher ----------------------------------------------------------------------------
 * CalculateMAP.process_one, metrics/map.py:16-83, for I images in one launch.  d_dets [sum M,6] =
 * [cls, conf, x1,y1,x2,y2], d_det_off [I+1]; d_gts [sum N,5] = [cls, x1,y1,x2,y2], d_gt_off [I+1];
 * thresholds (host) f64[n_thr] (<= 16).  d_correct [sum M, n_thr] u8.  d_ws: fvb_map_match_workspace_bytes.
 */
size_t fvb_map_match_workspace_bytes(int64_t total_dets);
int fvb_map_match_f32(const float* d_dets, const int32_t* d_det_off, const float* d_gts, const int32_t* d_gt_off,
                      int images, int64_t total_dets, const double* thresholds, int n_thr, uint8_t* d_correct,
                      void* d_ws, void* stream);

/* ---- AP integration on the device ---------------------------------------------------------------------
 * CalculateMAP.fetch / _ap_per_class / compute_ap, metrics/map.py:85-141, without the host numpy pass.
 * d_dets [n,6] = [cls, conf, ...] (the rows given to fvb_map_match_f32, all images concatenated), d_correct [n, n_thr] u8
 * (its output), d_target_cls [m] f32 (every target's class).  Class ids are integers in [0, max_class] (< 65535).
 * d_ap [(max_class+1), n_thr] f64: AP of class c at threshold k, NaN for classes without targets (the reference leaves
 * them out of the means, map.py:126-127); d_pos_count [max_class+2] i32: targets per class (last slot: other ids).
 * Rows of equal (class, conf) keep their input order (np.argsort(-conf) at map.py:131 leaves ties unspecified).
 * float64 arithmetic with numpy's formulas (np.interp, np.trapz); TP/FP counts are exact integers. */
size_t fvb_map_ap_workspace_bytes(int64_t n_dets, int n_thr, int max_class);
int fvb_map_ap_f64(const float* d_dets, const uint8_t* d_correct, int64_t n_dets, const float* d_target_cls,
                   int64_t n_targets, int n_thr, int max_class, double* d_ap, int32_t* d_pos_count, void* d_ws,
                   void* stream);

/* ---- anchor k-means --------------------------------------------------------------------------------------
 * One iteration of KMeans._fit, detection/tools/ANCHOR.py:33-46: d_categories[i] = 1 + argmin_c (1 - wh_iou(sample_i,
 * centre_c)) (first minimum; wh_iou_batch numpy branch, detection/tools/IOU.py:166-175), d_new_centers[c] = mean of the
 * samples assigned to c, or the old centre for an empty cluster.  d_samples [n,2], d_centers / d_new_centers [k,2] (k <= 64,
 * must not alias).  Means are formed from fp64 sums (the reference: np.mean of float32). */
size_t fvb_kmeans_workspace_bytes(int k);
int fvb_kmeans_step_f32(const float* d_samples, int64_t n, const float* d_centers, int k, float eps, int64_t* d_categories,
                        float* d_new_centers, void* d_ws, void* stream);

/* ---- validation-loop glue ---------------------------------------------------------------------
 * fvb_val_evidence_f32: what utils/fit.py:94-99 builds per image in Python, for a whole batch without a host sync:
 * the padded outputs of fvb_yolo_nms_*_f32 become compact rows d_out_dets [<= B*max_det, 6] = [cls, conf, x1, y1, x2, y2]
 * (:96) with CSR offsets d_out_det_off [B+1]; the labels [T,6] = [image, cls, xc, yc, w, h] (normalised) become
 * d_out_gts [<= T, 5] = [cls, x1, y1, x2, y2] in pixels (:98-99), grouped by image in input order, with d_out_gt_off [B+1].
 * Both outputs have capacity shapes; only the offsets say how many rows are valid (feed them to fvb_map_match_f32 with
 * total_dets = B*max_det).  A negative d_cnt[b] (failed image) counts as 0 detections.
 */
int fvb_val_evidence_f32(const float* d_boxes, const float* d_scores, const int64_t* d_cls, const int32_t* d_cnt, int batch,
                         int max_det, const float* d_labels, int64_t num_labels, float img_w, float img_h,
                         float* d_out_dets, int32_t* d_out_det_off, float* d_out_gts, int32_t* d_out_gt_off, void* stream);

/* ---- debug hooks -------------------------------------------------------------------------------
 * NOT part of the product path and NOT thread-safe: process-global state, used by tools/ only.
 *   fvb_debug_set_nms_trace   device buffer [B][8] int64 of globaltimer stamps per NMS phase written by the
 *                             fvb_yolo_nms_* kernels (tools/nms_trace.py); NULL (default) disables it.
 *   fvb_debug_reload_knobs    the FVB_DECODE_WARPS / FVB_DECODE_BATCH environment knobs are read once per
 *                             process; this makes the next decode launch read them again (tools/decode_sweep.py).
 */
void fvb_debug_set_nms_trace(void* d_buf);
void fvb_debug_reload_knobs(void);

#ifdef __cplusplus
}
#endif
#endif /* FVB200_H_ */
