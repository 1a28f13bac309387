/*
 * boxgeom.h -- C ABI of libboxgeom.so, the B200 (sm_100a) implementation of the
 * box-geometry hot path of ches-001/vision-conglomerate.
 *
 * The reference is pure Python and has no FFI of its own: its "plugin boundary"
 * is five Python call sites.  Each entry point below names the
 * reference routine it replaces (paths relative to the reference root); the
 * Python shim in vision_conglomerate_b200/ binds these with ctypes and re-creates
 * the reference signatures on top.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the comment says "host";
 *  - the library never allocates or frees device memory and keeps no state
 *    between calls: inputs, outputs and the scratch `workspace` are owned by the
 *    caller (query the size with the matching *_workspace_bytes function);
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *    synchronises with the host.  Data-dependent result sizes are written to
 *    device memory (`out_counts`), to be read by the caller after its own sync;
 *  - return value: BG_OK or an error code (see bg_strerror); never throws;
 *  - fp32 arithmetic follows the reference's CPU operation order without FMA
 *    contraction; integer outputs are bit-exact with the reference.
 */
#ifndef BOXGEOM_H_
#define BOXGEOM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_OK 0
#define BG_ERR_INVALID 1   /* bad argument (null pointer, negative size, unsupported shape) */
#define BG_ERR_WORKSPACE 2 /* workspace too small for this call */
#define BG_ERR_LAUNCH 3    /* CUDA reported an error while enqueueing */

/* bits of the device-side status word (out_counts[1]) */
#define BG_STATUS_GROUP_RANGE 1 /* batched_nms: max(idxs)-min(idxs) exceeds max_groups */
#define BG_STATUS_MASK_SPACE 2  /* suppression scratch (mask_bytes) exhausted: neither the overlap-edge list nor the dense
                                 * bit matrix fits; retry with a larger mask_bytes */
#define BG_STATUS_NEED_GENERAL 4 /* bg_detect, per-image NMS path: an image has too many survivors (out_counts[2+B+b]) or overlaps;
                                   * retry with nms_path = 2 (after 5), 4 (up to 8,192 survivors) or 1 */

#define BG_MAX_ANCHORS 8
#define BG_MAX_TRACKED 64

const char *bg_strerror(int code);
int bg_version(void);
/* number of kernels this library has launched in this process (for bench.py's gpu_launches) */
uint64_t bg_launch_count(void);
/* sizeof(bg_detect_params) / sizeof(bg_loss_params) as compiled, so a binding can verify its struct layout */
size_t bg_sizeof_detect_params(void);
size_t bg_sizeof_loss_params(void);

/* ------------------------------------------------------------------ B4
 * torchvision.ops.batched_nms(boxes, scores, idxs, iou_threshold), the call at
 * inference_det.py:77-82 / inference_seg.py:85-90.  Greedy NMS independently per
 * distinct idxs value (the `_batched_nms_vanilla` semantics every BASELINE
 * configuration takes), IoU test bit-identical to torchvision's CPU kernel.
 *   boxes [n,4] f32 xyxy, scores [n] f32, idxs [n] i64.
 *   out_keep [n] i64: kept indices, score-descending, index-ascending inside
 *                     equal scores (torchvision's order inside ties is arbitrary);
 *   out_counts [2] i32: [0] = number kept, [1] = status bits.
 *   max_groups bounds max(idxs)-min(idxs)+1 (sizes the per-group tables).
 *   mask_bytes: size of the suppression scratch inside the workspace.  For 0.05 <= iou_threshold < 1 it
 *               holds overlap edges (8 bytes per overlapping pair, shared out to the groups in proportion to
 *               their size); when a group's share overflows, the same call falls back to the dense bit
 *               matrix (sum over groups of count * ceil(count/64) * 8 bytes) if that fits; other thresholds
 *               use the matrix directly.  BG_STATUS_MASK_SPACE if neither fits.
 */
size_t bg_batched_nms_workspace_bytes(int64_t n, int64_t max_groups, size_t mask_bytes);
int bg_batched_nms(const float *boxes, const float *scores, const int64_t *idxs, int64_t n,
                   double iou_threshold, int64_t max_groups, int64_t *out_keep, int32_t *out_counts,
                   void *workspace, size_t workspace_bytes, size_t mask_bytes, void *stream);

/* ------------------------------------------------------------------ B5
 * Fused replacement of DetectionNet.forward(inference=True) from the three head
 * outputs on (modules/detection.py:69-91,98-190) followed by
 * inference_det.post_process_preds lines 57-97 and the tracked-class filter at
 * :107-109: decode, score, strict score threshold, per-image NMS, row assembly.
 */
typedef struct {
    int32_t B, C, na;           /* batch, classes, anchors per cell */
    int32_t H, W;               /* network input size (x.shape[2:]) */
    int32_t og_H, og_W;         /* original frame size; <= 0 means og_size=None */
    int32_t ny[3], nx[3];       /* feature-map shape per scale (sm, md, lg) */
    float anchors[3][BG_MAX_ANCHORS][2]; /* normalised (w,h) per scale, read from the model at call time */
    float box_allowance;        /* added to w and h before xyxy (inference_det.py:73-74); 0 = None */
    float score_threshold;      /* rows with score > threshold survive (strict, :84) */
    double iou_threshold;
    int32_t n_tracked;          /* 0 = no class filter */
    int32_t tracked[BG_MAX_TRACKED];
    int32_t order;              /* 0: image-major, score-descending inside an image; 1: globally score-descending (reference row order) */
    int32_t variant;            /* decode kernel tile loads: 0 auto, 1 plain loads, 2 TMA bulk pipeline (needs 16-byte aligned inputs) */
    int32_t nms_path;           /* 0 auto (per-image CTAs, up to 4,096 score survivors per image), 1 general segmented engine,
                                 * 2 = 0 but an error instead of the general engine when the threshold rules it out,
                                 * 3 = 2 with one CTA per image (no helper CTA), 4 per-image CTAs for up to 8,192 survivors,
                                 * 5 lean per-image CTAs (512 threads, 46 KB: up to 2,048 survivors per image) that fit on an SM
                                 * next to the decode CTAs of batches in flight on other streams */
    int32_t extra_cols;         /* columns per row after the four box columns that the kernels skip (mask coefficients and
                                 * keypoints of the segmentation / keypoint heads, inference_seg.py:66-68): rows are
                                 * 5 + C + extra_cols floats; the caller gathers them by out_keep */
    int32_t throughput;         /* 0: tuned for the latency of one batch (helper CTA per image while the GPU has room, the NMS
                                 * kernel launched programmatically behind the decode kernel); 1: tuned for several batches in
                                 * flight on different streams (one NMS CTA per image, plain stream-ordered launches) */
    void *nms_stream;           /* optional (NULL = everything on `stream`): a second stream of the caller for the NMS kernels,
                                 * normally one of higher priority, so that the block scheduler places the few NMS CTAs of a
                                 * batch ahead of the pending decode CTAs of the batches behind it.  The library orders the two
                                 * with nms_event: decode on `stream` -> event -> NMS on nms_stream -> event -> `stream` (which
                                 * therefore still sees the finished batch; no host synchronisation) */
    void *nms_event;            /* cudaEvent_t of the caller (timing disabled is fine), required with nms_stream */
    void *host_flag;            /* optional (NULL = off; replaces the per-image `boxes.detach().cpu().numpy()` of the reference's host
                                 * loop, inference_det.py:116-118): a 32-bit word in page-locked HOST memory that the device can address
                                 * (cudaHostAlloc / cudaHostRegister; bg_host_mapped_ptr checks it).  The last kernel of the call
                                 * stores host_flag_value there once every output of the call is visible to the host, so a caller
                                 * that ALSO placed out_boxes / out_img / out_keep / out_counts in such memory (batch-1 video
                                 * frames: a few dozen rows) gets the rows without a device->host copy and without a stream
                                 * synchronisation: it polls the word (BASELINE config 5; the kernels write through PCIe, which is
                                 * only sensible for small outputs) */
    int32_t host_flag_value;    /* value to store (use a fresh value per call, e.g. a sequence number) */
    int32_t reserved0;
} bg_detect_params;

/* The device address of page-locked host memory `host_ptr` for the current device, NULL if the device cannot address it
 * (cudaHostGetDevicePointer); under unified addressing it equals host_ptr. */
void *bg_host_mapped_ptr(void *host_ptr);

size_t bg_detect_workspace_bytes(const bg_detect_params *p /*host*/, size_t mask_bytes);
/*   raw_* [B,ny,nx,na,5+C] f32 contiguous, channels [obj, cls*C, tx,ty,tw,th] (common.py:912-931)
 *   out_boxes [B*N,6] f32 = (score, class, x1,y1,x2,y2)           (inference_det.py:93-97)
 *   out_img   [B*N] i64   = image index of each row               (sample_idxs, :87)
 *   out_keep  [B*N] i64   = flat candidate index b*N + i of each row (keep_idxs after the threshold)
 *   out_counts [2+2*B] i32: [0] rows written, [1] status bits, [2+b] rows of image b,
 *                           [2+B+b] candidates of image b that passed the score threshold
 */
int bg_detect(const float *raw_sm, const float *raw_md, const float *raw_lg, const bg_detect_params *p /*host*/,
              float *out_boxes, int64_t *out_img, int64_t *out_keep, int32_t *out_counts, void *workspace,
              size_t workspace_bytes, size_t mask_bytes, void *stream);

/* The same pipeline from the DECODED tensor the reference hands to its post-processing: inference_det.py
 * post_process_preds lines 57-97 and :107-109 (score = max_c sigmoid(cls_c) * sigmoid(obj), box allowance, xyxy,
 * per-image NMS, strict score threshold, row assembly, tracked-class filter) for preds [B, N, 5+C] f32 =
 * DetectionNet.forward(x, inference=True) (modules/detection.py:69-91), rows [obj, cls*C, x, y, w, h] with the
 * boxes already in pixels.  `p` as for bg_detect (ny/nx/na give the three scale segments of N; H, W, og_*, anchors
 * are ignored); workspace size from bg_detect_workspace_bytes; outputs as bg_detect. */
int bg_post_process(const float *preds, const bg_detect_params *p /*host*/, float *out_boxes, int64_t *out_img,
                    int64_t *out_keep, int32_t *out_counts, void *workspace, size_t workspace_bytes, size_t mask_bytes,
                    void *stream);

/* Profiling hook for bench.py: when both are non-NULL, the next bg_detect call records `start`/`stop`
 * (cudaEvent_t) on its stream immediately around the decode+filter kernel, then clears the hook. */
void bg_profile_events(void *start, void *stop);
/* The same for bg_loss_bwd: events around its dense-gradient fill (loss_bwd_stream_kernel for interleaved rows; the
 * memsets + objectness-plane kernel for the split form), i.e. before the matched-row kernel.  Hooks are per host
 * thread: the thread that arms one is the thread whose next call consumes it. */
void bg_profile_events_loss(void *start, void *stop);
/* Profiling hook: while `dev_buf` ([B, bg_profile_stamps_per_image()] u64, device) is non-NULL, the per-image NMS
 * kernel of bg_detect writes the %globaltimer value (ns) at each of its stage boundaries for every image. */
void bg_profile_stamps(void *dev_buf);
int bg_profile_stamps_per_image(void);
/* Profiling hook: while `dev_buf` ([1024, 8] u64, device) is non-NULL, thread 0 of every CTA of the decode kernel of
 * bg_detect accumulates the SM clock cycles it spends in each of its 8 phases (wait, phase 1, barrier, list,
 * phase 2a, barrier + phase 2b, barrier, refill). */
void bg_profile_decode_cycles(void *dev_buf);

/* DetectionNet._get_scale_pred (modules/detection.py:98-173) for one scale, optionally followed by
 * _bbox_to_size (:175-190): writes the decoded tensor, same shape as raw.  inference = 0 gives the
 * training-mode decode (xy = 2s-0.5, wh = (2s)^2 only). */
int bg_decode_scale(const float *raw, float *out, int32_t B, int32_t ny, int32_t nx, int32_t na, int32_t C,
                    const float *anchors /*host [na,2]*/, int32_t H, int32_t W, int32_t inference, int32_t og_H,
                    int32_t og_W, void *stream);
/* The same for the heads whose rows carry extra_cols more columns behind the box: rows are
 * 5 + C + extra_cols floats; the first tanh_cols of the extra columns are the mask coefficients of the segmentation
 * head, which _get_scale_pred passes through tanh (modules/detection.py:131-134); the rest is copied. */
int bg_decode_scale_ex(const float *raw, float *out, int32_t B, int32_t ny, int32_t nx, int32_t na, int32_t C,
                       int32_t extra_cols, int32_t tanh_cols, const float *anchors /*host [na,2]*/, int32_t H, int32_t W,
                       int32_t inference, int32_t og_H, int32_t og_W, void *stream);

/* Rows of DetectionNet.forward(x, inference=True) (modules/detection.py:69-91) for selected candidates only:
 * out [n, 5+C+extra] = [obj logit, class logits, x, y, w, h, ...] of the flat candidates idx [n] (i64, b*N + i, e.g. the
 * out_keep of bg_detect), decoded and rescaled exactly as bg_detect does internally (before the box allowance and the
 * xyxy step).  `p` as for bg_detect.  Lets the reference's own post-processing / host loop (inference_det.py:57-165)
 * run on the kept candidates only. */
int bg_decode_rows(const float *raw_sm, const float *raw_md, const float *raw_lg, const bg_detect_params *p /*host*/,
                   const int64_t *idx, int64_t n, float *out, void *stream);

/* DetectionNet._bbox_to_size (modules/detection.py:175-190) as called at :79-81, in place on decoded rows of D floats
 * (box columns at C+1..C+4): box = (box / from) * to; from4 / to4 are the DEVICE int64[4] tensors [W,H,W,H] /
 * [W0,H0,W0,H0] the reference builds at :77-78 (read on the device: no host sync). */
int bg_bbox_to_size(float *pred, int64_t rows, int32_t C, int32_t D, const int64_t *from4, const int64_t *to4, void *stream);

/* Backward of the training-mode decode (inference = 0 above; modules/detection.py:122,125,164), rows of 5+C floats:
 * grad_raw = grad_out on the objectness / class columns, grad_out * 2s(1-s) on x, y, grad_out * 8s^2(1-s) on w, h,
 * s = sigmoid(raw).  Makes DetectionNet._get_scale_pred differentiable on the CUDA path when its result is consumed
 * by something other than the fused loss (which takes the logits directly, BG_LOSS_RAW). */
int bg_decode_train_bwd(const float *raw, const float *grad_out, float *grad_raw, int64_t rows, int32_t C, void *stream);
/* The same for rows of 5 + C + extra_cols floats whose first tanh_cols extra columns went through tanh (bg_decode_scale_ex,
 * the segmentation head: modules/detection.py:131-134): grad_out * (1 - tanh(raw)^2) there, grad_out on the rest. */
int bg_decode_train_bwd_ex(const float *raw, const float *grad_out, float *grad_raw, int64_t rows, int32_t C, int32_t extra_cols,
                           int32_t tanh_cols, void *stream);

/* ------------------------------------------------------------------ B1
 * DetectionDataset.build_target_by_scale (dataset/detection_dataset.py:90-246), detection branch.
 *   targets [nt,6] f32 (img, cls, x, y, w, h); anchors host [na,2] normalised.
 *   out_idx4 [4,cap] i64 rows = batch_idx, grid_j, grid_i, anchor_idx; out_cls [cap] i64;
 *   out_anchor [cap,2] f32 (grid units); out_box [cap,4] f32; cap = 5*na*nt.
 *   out_count [1] i32 = M.  Output order: (offset k, anchor a, target t) lexicographic.
 */
size_t bg_assign_workspace_bytes(int64_t nt, int32_t na);
int bg_assign_targets(const float *targets, int64_t nt, int32_t ny, int32_t nx, const float *anchors /*host*/,
                      int32_t na, float anchor_t, float edge_t, int64_t *out_idx4, int64_t *out_cls,
                      float *out_anchor, float *out_box, int64_t cap, int32_t *out_count, void *workspace,
                      size_t workspace_bytes, void *stream);

/* The segmentation / keypoint variants of the same routine (detection_dataset.py:132-172,239-245):
 *   targets [nt, row_stride] with row_stride = 6 + 3*num_keypoints; the extra columns are returned per match in
 *   out_kpts [cap, row_stride-6] (the reference's `keypoints`).
 *   tmask_mode 0: none; 1: overlap_masks=False (mask index = the target's position); 2: overlap_masks=True
 *   (1 + position inside the image's block, blocks laid out by the per-image counts for ids 0..batch_size-1);
 *   out_tmask [cap] i64 (the reference's `tmask_idx`).
 *   out_count [2] i32: [0] = M, [1] = 1 if the per-image counts do not add up to nt (the reference raises).
 */
size_t bg_assign_ex_workspace_bytes(int64_t nt, int32_t na, int32_t batch_size);
int bg_assign_targets_ex(const float *targets, int64_t nt, int32_t row_stride, int32_t ny, int32_t nx,
                         const float *anchors /*host*/, int32_t na, float anchor_t, float edge_t, int32_t tmask_mode,
                         int32_t batch_size, int64_t *out_idx4, int64_t *out_cls, float *out_anchor, float *out_box,
                         int64_t *out_tmask, float *out_kpts, int64_t cap, int32_t *out_count, void *workspace,
                         size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ B2
 * DetectionLoss.compute_ciou (modules/detection_loss.py:229-264), element-wise [M,4] x [M,4] -> [M].
 * bg_ciou_bwd: grad_p[m,:] = grad_out[m] * d ciou / d preds_xywh (alpha held constant, :261-262). */
int bg_ciou_fwd(const float *preds_xywh, const float *targets_xywh, int64_t M, float eps, float *out, void *stream);
int bg_ciou_bwd(const float *preds_xywh, const float *targets_xywh, const float *grad_out, int64_t M, float eps,
                float *grad_p, void *stream);

/* ------------------------------------------------------------------ B3
 * DetectionLoss.forward (modules/detection_loss.py:84-226) for the default configuration
 * (BCEWithLogits, no focal loss, no keypoints): target assignment, matched-row gather, CIoU,
 * "last match wins" objectness targets, dense objectness BCE, class BCE and the per-class
 * confusion counters behind the sklearn metrics, for all three scales in one call -- from the decoded
 * tensors the reference's loss receives, or directly from the head's logits (the training-mode decode of
 * DetectionNet._get_scale_pred, modules/detection.py:98-173, fused in), interleaved or as the head's three
 * conv outputs.
 */
#define BG_LOSS_DECODED 0   /* rows [obj, cls*C, x, y, w, h (+extra)] as DetectionNet._get_scale_pred(inference=False) returns them */
#define BG_LOSS_RAW 1       /* the head's own rows (logits): the training-mode decode xy = 2s-0.5, wh = (2s)^2
                             * (modules/detection.py:122,125) is applied in registers, its derivative in the backward */
#define BG_LOSS_RAW_SPLIT 2 /* the head's three conv outputs before EffiDecHead.forward concatenates them
                             * (modules/common.py:908-919), each permuted to channels-last:
                             * conf [B,ny,nx,na], cls [B,ny,nx,na,C], bbox [B,ny,nx,na,4] */

typedef struct {
    int32_t B, C, na;
    int32_t ny[3], nx[3];
    float anchors[3][BG_MAX_ANCHORS][2];
    float anchor_t, edge_t, label_smoothing;
    float box_w, conf_w, class_w;
    float scale_w[3];
    int64_t nt;
    int32_t input_form;  /* BG_LOSS_DECODED / BG_LOSS_RAW / BG_LOSS_RAW_SPLIT */
    int32_t extra_cols;  /* interleaved forms: columns after the box columns (mask coefficients, keypoints) that the loss
                          * skips; rows are 5 + C + extra_cols floats and their gradient is written as zeros */
} bg_loss_params;

/* One scale of the prediction tensors.  Interleaved forms: `obj` is the base of the [B,ny,nx,na,5+C+extra] tensor,
 * `cls` / `box` are ignored.  BG_LOSS_RAW_SPLIT: the three tensors.  All 16-byte aligned, contiguous. */
typedef struct { const float *obj, *cls, *box; } bg_head_ptrs;
typedef struct { float *obj, *cls, *box; } bg_head_grads;

size_t bg_loss_workspace_bytes(const bg_loss_params *p /*host*/);
/*   in[3] (host array of device pointers): scales sm, md, lg;  targets [nt,6] f32.
 *   out_scalars [3,8] f64 per scale: lbox, lconf, lcls (NaN->0 applied), mean_ciou, avg_pos_conf,
 *                 avg_neg_conf, M, n_neg;   out_hist [3,3,C] i64: tp, n_true, n_pred per class;
 *   out_loss [1] f32: box_w*sum_s(scale_w*lbox) + conf_w*... + class_w*...  (:107-110);
 *   out_status [1] i32 (may be NULL): bit 0 = a target row names an image outside 0..B-1 or a class outside 0..C-1
 *                 (the reference raises IndexError there; here the row is dropped and the bit is set).
 *   The workspace keeps what bg_loss_bwd needs (matches, CIoU gradients, objectness residuals): it must stay
 *   untouched until the matching bg_loss_bwd has run -- use one workspace per forward that is still awaiting
 *   its backward.  Launches: one memset + three kernels (the 2nd and 3rd with programmatic dependent launch).
 */
int bg_loss_fwd(const bg_head_ptrs in[3] /*host*/, const float *targets, const bg_loss_params *p /*host*/,
                double *out_scalars, int64_t *out_hist, float *out_loss, int32_t *out_status, void *workspace,
                size_t workspace_bytes, void *stream);
/*   grads[3]: d(grad_out * loss)/d in, same form and shapes as `in`, every element written (BG_LOSS_RAW*: the
 *   gradient with respect to the logits).  The upstream gradient is read from device memory (grad_out_dev [1] f32)
 *   when non-NULL -- no host sync in loss.backward() -- else grad_out_host is used.
 *   flags: BG_LOSS_BWD_PRECLEARED (split form) = the class / box planes of `grads` are already zero (see
 *   bg_loss_clear_grads); the call then only writes the objectness plane and the matched rows. */
#define BG_LOSS_BWD_PRECLEARED 1
int bg_loss_bwd(const bg_head_ptrs in[3] /*host*/, const bg_loss_params *p /*host*/, const float *grad_out_dev,
                float grad_out_host, const bg_head_grads grads[3] /*host*/, int32_t flags, void *workspace,
                size_t workspace_bytes, void *stream);
/* BG_LOSS_RAW_SPLIT: clears the class / box planes of the gradient tensors (adjacent planes with one cudaMemsetAsync),
 * for callers that allocate the gradients up front and want the 2 GB of zeros -- which do not depend on the forward --
 * written elsewhere in their schedule; bg_loss_bwd then takes BG_LOSS_BWD_PRECLEARED.  (Measured on B200: running the
 * clear on a second stream NEXT TO bg_loss_fwd gains nothing -- a memset fills every thread slot of the machine and
 * the forward kernels queue behind it, a fill kernel small enough to leave them room writes at 1-2 TB/s) */
int bg_loss_clear_grads(const bg_loss_params *p /*host*/, const bg_head_grads grads[3] /*host*/, void *stream);

/* ------------------------------------------------------------------ f2
 * The mask term of SegmentationLoss (modules/segmentation_loss.py:26-75,147-171,208-231) for overlap_masks=True and
 * BCEWithLogits, added to the detection terms that bg_loss_fwd / bg_loss_bwd compute from the same prediction tensors
 * (BG_LOSS_DECODED with extra_cols >= K: the mask coefficients are the first K columns after the box columns):
 *     loss = detection loss + seg_w * sum_s scale_w[s] * seg_loss_s      (:57-58)
 * seg_loss_s / dice_score_s as in :147-171 with segmentation_metrics (:208-231), compute_dice_score and crop_section
 * (utils/utils.py:130-172); the masks are nearest-resized to the protos' size like F.interpolate (:152-153).
 * Call order: bg_loss_fwd, then bg_seg_loss_fwd with the same out_loss (it adds its term); bg_loss_bwd, then
 * bg_seg_loss_bwd with the same gradient tensors (it adds to their coefficient columns) -- all on one stream.
 */
typedef struct {
    int32_t B, C, na, K;        /* K mask coefficients: 8, 16 or 32 */
    int32_t extra_cols;         /* rows are 5 + C + extra_cols floats, extra_cols >= K */
    int32_t ny[3], nx[3];
    float anchors[3][BG_MAX_ANCHORS][2];
    float anchor_t, edge_t;
    int32_t Hp, Wp;             /* protos [B, K, Hp, Wp] */
    int32_t Hm, Wm;             /* target_masks [B, Hm, Wm] f32: 1 + position of the covering object inside its image */
    float scale_w[3];
    float seg_w;
    int64_t nt;
} bg_seg_params;

size_t bg_sizeof_seg_params(void);
size_t bg_seg_loss_workspace_bytes(const bg_seg_params *p /*host*/);
/*   preds[3] (host array of device pointers) [B,ny,nx,na,5+C+extra]; targets [nt,6]; protos; target_masks.
 *   inout_loss [1] f32: the detection loss on entry, the segmentation loss on return.
 *   out_scalars [3,2] f64: (seg_loss, dice_score) per scale.  out_status [1] i32: 1 if the per-image target counts for
 *   image ids 0..B-1 do not add up to nt (the reference raises, detection_dataset.py:151-157).
 *   The workspace keeps what bg_seg_loss_bwd needs.  Nine launches, no host synchronisation. */
int bg_seg_loss_fwd(const float *const preds[3] /*host*/, const float *targets, const float *protos, const float *target_masks,
                    const bg_seg_params *p /*host*/, float *inout_loss, double *out_scalars, int32_t *out_status,
                    void *workspace, size_t workspace_bytes, void *stream);
/*   grad_preds[3]: the dense gradients bg_loss_bwd has written (zeros in the extra columns): the mask term's
 *   d loss / d coefficients is ADDED to the matched rows.  grad_protos [B,K,Hp,Wp]: every element written. */
int bg_seg_loss_bwd(const float *const preds[3] /*host*/, const float *protos, const float *target_masks,
                    const bg_seg_params *p /*host*/, const float *grad_out_dev, float *const grad_preds[3] /*host*/,
                    float *grad_protos, void *workspace, size_t workspace_bytes, void *stream);

/* inference_seg.post_process_preds lines 115-117: for the kept rows of every image,
 *     masks = sigmoid(coefs @ protos_i) -> F.interpolate(size=(H, W), mode="bilinear", align_corners=False) -> > 0.5.
 *   coefs [n, K] f32: the mask coefficients of the kept rows, image by image; row_offsets [B+1] i32 (device): rows of
 *   image i are row_offsets[i] .. row_offsets[i+1]-1; protos [B, K, Hp, Wp]; scratch [n, Hp*Wp] f32 (caller's);
 *   out_masks [n, H, W] u8 (0 / 1, i.e. torch.bool), 4-byte aligned.  K <= 64.  Two launches. */
int bg_seg_masks(const float *coefs, const int32_t *row_offsets, const float *protos, int32_t B, int32_t K, int32_t Hp,
                 int32_t Wp, int64_t n, int32_t H, int32_t W, float *scratch, uint8_t *out_masks, void *stream);

/* Image-sharded training: per-shard sums that add up over the shards, and the big-batch loss from the
 * summed terms.  pack15 [3,5] f64 per scale = {lbox*M, lconf*cells, lcls*M*C, M, cells}; the caller all-reduces (SUM)
 * the 15 doubles between the two calls (NCCL).  cells3: host int64[3], this shard's B*ny*nx*na per scale.
 * bg_loss_combine: out_loss [1] f64 = box_w*sum_s(w_s*lbox_s) + conf_w*... + class_w*... with global means
 * (modules/detection_loss.py:107-110 on the concatenated batch); only p->C, the weights and scale_w are read. */
int bg_loss_pack(const double *scalars, const int64_t *cells3 /*host*/, int32_t C, double *pack15, void *stream);
int bg_loss_combine(const double *pack15, const bg_loss_params *p /*host*/, double *out_loss, void *stream);

/* ------------------------------------------------------------------ a13
 * utils/make_anchors.py:14-39 ratio_metrics / ratio_metrics_w_extras.
 *   wh [n,2] f32, anchors host [k,2]; out3 [3] f64 = sum(v*m), sum(m), n  (score = out[0]/n, bpr = out[1]/n, aat = out[1]).
 */
int bg_ratio_metrics(const float *wh, int64_t n, const float *anchors /*host*/, int32_t k, float threshold,
                     double *out3, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BOXGEOM_H_ */
