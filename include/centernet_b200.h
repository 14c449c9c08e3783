/*
 * centernet_b200 -- C ABI of the B200-native CenterNet heat-map hot path.
 *
 * This is the drop-in boundary: plain device pointers, explicit sizes and a
 * CUDA stream handle; no torch types.  The reference has two bindings on this
 * path and both are replaced here:
 *
 *   (1) the Python ATen-op functions of src/lib/models/decode.py,
 *       src/lib/models/utils.py and src/lib/models/losses.py (no FFI of their
 *       own -- each cnb_* entry below names the function it replaces), and
 *   (2) the cffi module `_ext.dcn_v2` declared in
 *       src/lib/models/networks/DCNv2/src/dcn_v2_cuda.h:9-55
 *       (dcn_v2_cuda_forward / dcn_v2_cuda_backward / dcn_v2_psroi_pooling_*).
 *
 * Conventions
 *   - every tensor is fp32, NCHW, contiguous, resident on the current device;
 *     index outputs use the reference's dtypes (int64 indices, int32 classes);
 *   - the caller owns every buffer, including the scratch `workspace`
 *     (size from the matching cnb_*_workspace_bytes); the library never calls
 *     cudaMalloc and keeps no global mutable state, so it may be called from
 *     one thread per GPU concurrently (torch DataParallel style);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *     nothing synchronises the host;
 *   - return value: CNB_OK or an error code; cnb_last_error() gives a
 *     thread-local description.  The reference's THArgCheck/THError
 *     (dcn_v2_cuda.c:20-38) map to CNB_EINVAL.
 */
#ifndef CENTERNET_B200_H_
#define CENTERNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNB_OK 0
#define CNB_EINVAL 1       /* bad shape / null pointer / misaligned buffer        */
#define CNB_EUNSUPPORTED 2 /* valid request outside the implemented envelope     */
#define CNB_ECUDA 3        /* CUDA runtime / launch error                        */
#define CNB_EWORKSPACE 4   /* workspace_bytes smaller than cnb_*_workspace_bytes */

/* Library version (major*10000 + minor*100 + patch). */
int cnb_version(void);
/* Thread-local text of the last error returned on this thread ("" if none). */
const char *cnb_last_error(void);
/* Number of kernels launched by this library since process start (all threads);
 * bench.py reports it as `gpu_launches`. */
unsigned long long cnb_launch_count(void);

/* ---------------------------------------------------------------- A1: _nms
 * models/decode.py:9-15.  out = heat * (maxpool3x3(heat) == heat).
 * heat,out: [n, c, h, w]. */
int cnb_nms(const float *heat, float *out, int n, int c, int h, int w, void *stream);

/* ------------------------------------------------------------- A2/A3: top-K
 * Workspace for every selection-based entry point below (depends on the
 * largest (n_img, c, h, w, k) the call processes). */
size_t cnb_topk_workspace_bytes(int n_img, int c, int h, int w, int k);

/* models/decode.py:103-119 (_topk).  scores: [b, c, h, w] (already NMS'd by the
 * caller, as in the reference) -> per image the K best over all classes:
 * out_scores[b,k] f32, out_inds[b,k] i64 (y*w+x), out_clses[b,k] i32,
 * out_ys/out_xs[b,k] f32.  Order: score desc, ties by flat index asc.
 * fuse_nms != 0 applies the 3x3 peak test of _nms first (scores = raw heat). */
int cnb_topk(const float *scores, int b, int c, int h, int w, int k, int fuse_nms,
             float *out_scores, int64_t *out_inds, int32_t *out_clses,
             float *out_ys, float *out_xs,
             void *workspace, size_t workspace_bytes, void *stream);

/* models/decode.py:92-101 (_topk_channel): per (b, channel) top-K.
 * Outputs are [b, c, k]. */
int cnb_topk_channel(const float *scores, int b, int c, int h, int w, int k, int fuse_nms,
                     float *out_scores, int64_t *out_inds, float *out_ys, float *out_xs,
                     void *workspace, size_t workspace_bytes, void *stream);

/* --------------------------------------------------------------- A4: gather
 * models/utils.py:22-26 (_transpose_and_gather_feat) without the transpose
 * copy: out[b, m, ch] = feat[b, ch, ind[b, m]];  feat [b, c, h*w], ind [b, m]. */
int cnb_gather_feat(const float *feat, const int64_t *ind, float *out,
                    int b, int c, int hw, int m, void *stream);
/* Adjoint of cnb_gather_feat (what autograd derives for the reference's
 * permute + gather, models/utils.py:12-26; used by every Reg*Loss, L1Loss and
 * BinRotLoss of models/losses.py through `pred = _transpose_and_gather_feat`):
 * grad_feat[b, ch, ind[b, m]] += grad_out[b, m, ch].  grad_feat must arrive
 * zero-filled; repeated indices accumulate. */
int cnb_gather_feat_backward(const float *grad_out, const int64_t *ind, float *grad_feat,
                             int b, int c, int hw, int m, void *stream);

/* ---------------------------------------------------------- A5: ctdet_decode
 * models/decode.py:464-495.  heat [b,c,h,w] (post-sigmoid), wh [b,2,h,w] (or
 * [b,2c,h,w] when cat_spec_wh), reg [b,2,h,w] or NULL -> dets [b,k,6] =
 * x1,y1,x2,y2,score,class.  One fused pass: peak-NMS + top-K + gathers. */
int cnb_ctdet_decode(const float *heat, const float *wh, const float *reg, int cat_spec_wh,
                     int b, int c, int h, int w, int k, float *dets,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------- N1: ctdet_decode with fused sigmoid
 * Replaces `hm = output['hm'].sigmoid_(); ctdet_decode(hm, wh, reg, ...)`
 * (detectors/ctdet.py:30-45): hm_logits [b,c,h,w] is the head's raw output.
 * Scores are bit-identical to torch's CUDA sigmoid of the logits, indices and
 * classes identical to cnb_ctdet_decode on that heat map.  The workspace query
 * takes the logits pointer because its alignment decides the code path. */
size_t cnb_ctdet_logits_workspace_bytes(const float *hm_logits, int b, int c, int h, int w, int k);
int cnb_ctdet_decode_logits(const float *hm_logits, const float *wh, const float *reg, int cat_spec_wh,
                            int b, int c, int h, int w, int k, float *dets,
                            void *workspace, size_t workspace_bytes, void *stream);

/* out[i] = 1 / (1 + exp(-x[i])): the bare `sigmoid_()` of detectors/ctdet.py:31, the very expression the
 * fused path applies to its candidates (exposed so tests can pin it against torch and prove it monotone). */
int cnb_sigmoid(const float *x, float *out, long long n, void *stream);

/* ------------------------------------------- N1: flip-test merge of the heads
 * detectors/ctdet.py:34-37 and detectors/multi_pose.py:43-51.  src [2n,c,h,w]:
 * n images followed by their horizontally flipped copies; dst [n,c,h,w] =
 * (f(src[i,ch]) + sign[ch] * f(flip_w(src[n+i, perm[ch]]))) / 2, f = sigmoid when
 * apply_sigmoid.  perm/sign NULL = flip_tensor (models/utils.py:28-29); perm =
 * joint swap for flip_lr (:33-39); perm + sign for flip_lr_off (:41-50). */
int cnb_flip_merge(const float *src, float *dst, int n, int c, int h, int w, int apply_sigmoid,
                   const int32_t *perm, const float *sign, void *stream);

/* ------------------------------------------------------------ A9: ddd_decode
 * models/decode.py:426-462.  rot [b,8,h,w], depth [b,1,h,w], dim [b,3,h,w],
 * wh/reg [b,2,h,w] or NULL -> dets [b,k,18] (16 when wh == NULL). */
int cnb_ddd_decode(const float *heat, const float *rot, const float *depth, const float *dim,
                   const float *wh, const float *reg, int b, int c, int h, int w, int k,
                   float *dets, void *workspace, size_t workspace_bytes, void *stream);

/* ----------------------------------------------------- A6: multi_pose_decode
 * models/decode.py:497-571.  heat [b,1,h,w]; wh, reg(NULL ok) [b,2,h,w];
 * kps [b,2j,h,w]; hm_hp [b,j,h,w] or NULL; hp_offset [b,2,h,w] or NULL
 * -> dets [b,k,4+1+2j+1]. */
size_t cnb_multi_pose_workspace_bytes(int b, int c, int j, int h, int w, int k);
int cnb_multi_pose_decode(const float *heat, const float *wh, const float *kps, const float *reg,
                          const float *hm_hp, const float *hp_offset,
                          int b, int c, int j, int h, int w, int k, float *dets,
                          void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------ A7: edge aggregation
 * models/decode.py:17-77.  out = heat + weight*(dirA + dirB) with
 * horizontal != 0: _h_aggregate (left+right), else _v_aggregate (top+bottom).
 * Strictly sequential fp32 running sums, as the reference. */
int cnb_edge_aggregate(const float *heat, float *out, int n, int c, int h, int w,
                       float aggr_weight, int horizontal, void *stream);

/* ------------------------------------------- A8: exct_decode / agnex_ct_decode
 * models/decode.py:273-424 and :122-271.  t,l,b,r heat [b,c,h,w] (c==1 for
 * agnostic), ct heat [b,cc,h,w]; regr [b,2,h,w] each or all NULL
 * -> dets [b,num_dets,14]. */
size_t cnb_exct_workspace_bytes(int b, int c, int cc, int h, int w, int k, int num_dets);
int cnb_exct_decode(const float *t_heat, const float *l_heat, const float *b_heat,
                    const float *r_heat, const float *ct_heat,
                    const float *t_regr, const float *l_regr, const float *b_regr,
                    const float *r_regr, int b, int c, int cc, int h, int w, int k,
                    float scores_thresh, float center_thresh, float aggr_weight,
                    int num_dets, int agnostic, float *dets,
                    void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------- A13/A14/A15: focal + splat
 * _sigmoid (models/utils.py:8-10), _neg_loss (models/losses.py:42-67) and the
 * Gaussian target splat draw_umich_gaussian (utils/image.py:126-141) fused.
 *
 * cnb_focal_loss: pred/gt [n] flat (gt given as a dense map, reference form).
 *   logits != 0: `pred` holds raw logits and _sigmoid (clamped) is applied
 *   first.  Writes out[0]=loss, out[1]=num_pos; grad (nullable) receives
 *   d loss / d pred (or / d logit when logits != 0) scaled by grad_scale.
 * cnb_focal_splat_loss: same loss but the target map is never materialised:
 *   it is rebuilt on the fly from per-image object lists
 *   obj_cls[b,m] i32, obj_cx/obj_cy[b,m] i32 (ct_int), obj_radius[b,m] i32,
 *   obj_valid[b,m] u8 -- exactly the draw_umich_gaussian calls of
 *   datasets/sample/ctdet.py:111-117.  m <= 128 (the reference's max_objs,
 *   datasets/dataset/coco.py:38); larger lists return CNB_EUNSUPPORTED. */
size_t cnb_focal_workspace_bytes(long long n);
int cnb_focal_loss(const float *pred, const float *gt, long long n, int logits,
                   float grad_scale, float *out2, float *grad,
                   void *workspace, size_t workspace_bytes, void *stream);
int cnb_splat_gaussian(const int32_t *obj_cls, const int32_t *obj_cx, const int32_t *obj_cy,
                       const int32_t *obj_radius, const uint8_t *obj_valid, int b, int m,
                       int c, int h, int w, float *hm, void *stream);
int cnb_focal_splat_loss(const float *pred, const int32_t *obj_cls, const int32_t *obj_cx,
                         const int32_t *obj_cy, const int32_t *obj_radius,
                         const uint8_t *obj_valid, int b, int m, int c, int h, int w,
                         int logits, float grad_scale, float *out2, float *grad,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------ A16: RegL1Loss
 * models/losses.py:139-149 (mode 0), RegLoss smooth-L1 :123-137 (mode 1),
 * NormRegL1Loss :151-163 (mode 2), RegWeightedL1Loss :165-175 (mode 3; mask is
 * then float [b,m,d]).  output [b,d,h*w], ind [b,m] i64, target [b,m,d].
 * out[0] = loss; grad_output (nullable, [b,d,hw], must be zero-filled) gets
 * d loss/d output * grad_scale. */
int cnb_reg_loss(const float *output, const void *mask, const int64_t *ind, const float *target,
                 int b, int d, int hw, int m, int mode, float grad_scale,
                 float *out1, float *grad_output, void *stream);

/* ---------------------------------------------------------- A10/A11: DCNv2
 * Replaces dcn_v2_cuda_forward / dcn_v2_cuda_backward
 * (DCNv2/src/dcn_v2_cuda.h:9-33, dcn_v2_cuda.c:10-241).  im2col-free: no
 * `columns` / `ones` scratch.  input [b,cin,h,w], offset [b,2*kh*kw*dg,ho,wo]
 * (channel 2t = dy, 2t+1 = dx of tap t), mask [b,kh*kw*dg,ho,wo],
 * weight [cout,cin,kh,kw], bias [cout] -> output [b,cout,ho,wo].
 * Backward ACCUMULATES into grad_weight/grad_bias and into
 * grad_input/grad_offset/grad_mask, all of which must arrive zero-filled as in
 * dcn_v2_func.py:44-48.
 * Forward workspace (cnb_dcnv2_workspace_bytes, 16-byte aligned): the weights
 * re-tiled for the tensor-core path, a channels-last copy of the input and, for
 * maps too small to fill the GPU, split-K partial sums; the query returns 0 for
 * kernels with more than 9 taps.  workspace == NULL (or smaller than the query,
 * or anisotropic stride/pad/dilation) selects the fp32 CUDA-core forward instead.
 * cnb_dcnv2_backward: with a workspace of cnb_dcnv2_backward_workspace_bytes (16-byte aligned; channels-last
 * copies of the input and of dX, TF32 hi/lo tiles of grad_output and of the weights, the column gradient of
 * up to ~1.5 GiB worth of images at a time) the contractions run on the tensor cores; workspace == NULL (or
 * smaller, or anisotropic stride/pad/dilation, or more than 9 taps -- the query then returns 0) selects the
 * fp32 CUDA-core backward. */
size_t cnb_dcnv2_workspace_bytes(int b, int cin, int cout, int h, int w, int kh, int kw,
                                 int stride, int pad, int dil, int dg);
int cnb_dcnv2_forward(const float *input, const float *offset, const float *mask,
                      const float *weight, const float *bias, float *output,
                      int b, int cin, int h, int w, int cout, int kh, int kw,
                      int stride_h, int stride_w, int pad_h, int pad_w,
                      int dil_h, int dil_w, int deformable_groups,
                      void *workspace, size_t workspace_bytes, void *stream);
/* Steady-state form of the forward (what the Python module uses): the weight
 * tiles (TF32 hi/lo split, UMMA layout) are built ONCE per weight version into a
 * caller-owned buffer of cnb_dcnv2_wtiles_bytes and reused by every call;
 * input_channels_last != 0 says `input` already is [b][h*w][cin] (torch
 * channels_last) so no re-layout pass runs (needs cin/dg % 32 == 0, else the flag
 * is ignored and `input` must be NCHW).  workspace: cnb_dcnv2_prepared_workspace_bytes. */
size_t cnb_dcnv2_wtiles_bytes(int cin, int cout, int kh, int kw, int dg);
int cnb_dcnv2_prepare_weights(const float *weight, int cin, int cout, int kh, int kw,
                              int deformable_groups, void *wtiles, size_t wtiles_bytes, void *stream);
size_t cnb_dcnv2_prepared_workspace_bytes(int b, int cin, int cout, int h, int w, int kh, int kw,
                                          int stride, int pad, int dil, int dg);
int cnb_dcnv2_forward_prepared(const float *input, int input_channels_last, const float *offset,
                               const float *mask, const void *wtiles, const float *bias, float *output,
                               int b, int cin, int h, int w, int cout, int kh, int kw,
                               int stride_h, int stride_w, int pad_h, int pad_w,
                               int dil_h, int dil_w, int deformable_groups,
                               void *workspace, size_t workspace_bytes, void *stream);
/* N3 (inference): the DCN module with its prologue and epilogue fused around the tensor-core kernel.
 * offset_mask is the RAW output of DCN.conv_offset_mask, [b, 3*kh*kw*dg, ho, wo]: its first two thirds are
 * read as the offsets and the sigmoid of its last third as the mask inside the sampler (dcn_v2.py:64-70:
 * chunk, cat and sigmoid disappear); bn_scale / bn_shift (nullable, [cout]) are the folded inference
 * BatchNorm that follows the DCN in DeformConv and relu its activation (pose_dla_dcn.py:345-357):
 * y = relu(bn_scale * (dcn + bias) + bn_shift).  Workspace as cnb_dcnv2_forward_prepared. */
int cnb_dcnv2_forward_fused(const float *input, int input_channels_last, const float *offset_mask,
                            const void *wtiles, const float *bias, const float *bn_scale,
                            const float *bn_shift, int relu, float *output,
                            int b, int cin, int h, int w, int cout, int kh, int kw,
                            int stride, int pad, int dil, int deformable_groups,
                            void *workspace, size_t workspace_bytes, void *stream);
size_t cnb_dcnv2_backward_workspace_bytes(int b, int cin, int cout, int h, int w, int kh, int kw,
                                          int stride, int pad, int dil, int dg);
/* The backward for a channels-last pipeline (torch channels_last): input_channels_last / grad_input_channels_last
 * != 0 say the tensor is [b][h*w][cin]; the re-layout pass of the input, the zero-fill and the final NCHW pass of
 * grad_input disappear (grad_input still arrives zero-filled).  Tensor-core path only: needs the workspace,
 * cin / deformable_groups % 32 == 0, cout <= 256, non-deterministic dX; CNB_EUNSUPPORTED otherwise. */
int cnb_dcnv2_backward_ex(const float *input, int input_channels_last, const float *offset,
                          const float *mask, const float *weight, const float *grad_output,
                          float *grad_input, int grad_input_channels_last,
                          float *grad_offset, float *grad_mask, float *grad_weight, float *grad_bias,
                          int b, int cin, int h, int w, int cout, int kh, int kw,
                          int stride, int pad, int dil, int deformable_groups,
                          void *workspace, size_t workspace_bytes, void *stream);
/* grad_offset, grad_mask, grad_weight and grad_bias of the tensor-core backward are bit-identical run to run.
 * grad_input is by default scattered with (vector) atomics like the reference's col2im
 * (dcn_v2_im2col_cuda.cu:182-239), i.e. its last bits depend on the order the adds land.  on != 0 selects a
 * gather instead (stride 1): every input position collects the samples that fall on it in a fixed order, which
 * makes grad_input bit-identical as well as long as no offset moves a sample more than 2 pixels from its tap
 * (those few still go through atomics); it costs about one extra forward.  Process-wide, default off. */
void cnb_dcnv2_set_deterministic(int on);
int cnb_dcnv2_get_deterministic(void);
int cnb_dcnv2_backward(const float *input, const float *offset, const float *mask,
                       const float *weight, const float *grad_output,
                       float *grad_input, float *grad_offset, float *grad_mask,
                       float *grad_weight, float *grad_bias,
                       int b, int cin, int h, int w, int cout, int kh, int kw,
                       int stride_h, int stride_w, int pad_h, int pad_w,
                       int dil_h, int dil_w, int deformable_groups,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------- N4: deformable PSROI pooling
 * Replaces dcn_v2_psroi_pooling_cuda_forward / _backward (src/dcn_v2_cuda.h:35-55,
 * src/cuda/dcn_v2_psroi_pooling_cuda.cu:47-255).  data [b,c,h,w]; rois [num_rois,5]
 * = (batch index, x1, y1, x2, y2); trans [num_rois, channels_trans, part, part] or
 * NULL when no_trans; out / top_count [num_rois, output_dim, pooled, pooled].
 * Backward ACCUMULATES into grad_data / grad_trans, which must arrive zero-filled
 * (dcn_v2_func.py:119-120). */
int cnb_psroi_pooling_forward(const float *data, const float *rois, const float *trans,
                              float *out, float *top_count,
                              int b, int c, int h, int w, int num_rois, int channels_trans,
                              int no_trans, float spatial_scale, int output_dim, int group_size,
                              int pooled_size, int part_size, int sample_per_part, float trans_std,
                              void *stream);
int cnb_psroi_pooling_backward(const float *grad_out, const float *data, const float *rois,
                               const float *trans, const float *top_count,
                               float *grad_data, float *grad_trans,
                               int b, int c, int h, int w, int num_rois, int channels_trans,
                               int no_trans, float spatial_scale, int output_dim, int group_size,
                               int pooled_size, int part_size, int sample_per_part, float trans_std,
                               void *stream);

/* ------------------------------------------- N2: detection post-processing
 * Device-side replacements of utils/post_process.py:83-114 (ctdet_post_process,
 * multi_pose_post_process), external/nms.pyx:77-275 (soft_nms, soft_nms_39) and
 * detectors/ctdet.py:76-92 (merge_outputs): the detections stay on the GPU
 * from decode to the final per-class lists.
 *
 * cnb_post_transform: rows [b,n,d] -> out [b,n,d]; the (x,y) pairs at columns
 *   [a0, a0+2*na) and [b0, b0+2*nb) are mapped through the image's 2x3 float64
 *   matrix trans[b][6] (the inverse affine of utils/image.py:27-60, computed by
 *   the caller on the host), every other column is copied.
 * cnb_group_by_class: dets [b,n,6] (x1,y1,x2,y2,score,class) -> rows [b,n,5]
 *   grouped by class, input order kept inside a class; offsets [b,classes+1].
 * cnb_soft_nms: one list per (image, class): rows image*n_cap +
 *   [offsets[image][c], offsets[image][c+1]) of `rows` ([*,d], d = 5 or 39;
 *   extra columns follow soft_nms_39's swap rule).  The array, rows beyond the
 *   final length included, ends exactly as the reference's in-place routine
 *   leaves it; final_n (nullable, [n_img*lists_per_img]) receives len(keep)
 *   (negative: list longer than 1536 rows, left untouched).  out may alias rows
 *   only when d == 5.  method 0 hard, 1 linear, 2 gaussian.
 * cnb_topk_keep: keep[b,n_cap] = score >= (max_per_image-th largest score of
 *   the image's offsets[b][classes] rows); thresh (nullable) [b]. */
int cnb_post_transform(const float *rows, float *out, const double *trans, int b, int n, int d,
                       int a0, int na, int b0, int nb, void *stream);
int cnb_group_by_class(const float *dets, int b, int n, int num_classes, float *rows,
                       int32_t *offsets, void *stream);
int cnb_soft_nms(const float *rows, float *out, const int32_t *offsets, int n_img, int lists_per_img,
                 int n_cap, int d, float sigma, float nt, float threshold, int method,
                 int32_t *final_n, void *stream);
int cnb_topk_keep(const float *rows, int b, int n_cap, int d, const int32_t *offsets, int num_classes,
                  int max_per_image, uint8_t *keep, float *thresh, void *stream);

/* ------------------------------------------- N4: detector input pre-processing
 * detectors/base_detector.py:37-65 for one image: cv2.warpAffine(INTER_LINEAR,
 * border 0) + (x/255 - mean)/std + HWC->CHW (+ mirrored copy when flip_test).
 * image_hwc: uint8 [h,w,3] on the device; minv6: the INVERTED 2x3 float64 matrix
 * (output pixel -> source coordinates, as cv::warpAffine derives it from its
 * forward matrix); mean3/std3: fp32; out: [1 or 2, 3, out_h, out_w] fp32.
 * Bit-identical to OpenCV's fixed-point bilinear warp. */
/* cv2.resize(image, (out_w, out_h)) with INTER_LINEAR for uint8 [h,w,3] images (the multi-scale test path of
 * base_detector.py:55), bit-identical to OpenCV's fixed-point filter. */
int cnb_resize_image(const uint8_t *image_hwc, int h, int w, uint8_t *out_hwc, int out_h, int out_w,
                     void *stream);
int cnb_preprocess_image(const uint8_t *image_hwc, int h, int w, const double *minv6,
                         const float *mean3, const float *std3, float *out,
                         int out_h, int out_w, int flip_test, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CENTERNET_B200_H_ */
