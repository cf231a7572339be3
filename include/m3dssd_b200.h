/*
 * m3dssd_b200 -- C ABI of the Blackwell (sm_100a) implementation of M3DSSD's
 * dense forward path.  Plain pointers and sizes only; every device pointer is
 * caller-owned; every call is asynchronous on `stream` unless noted.
 *
 * All functions return M3D_OK (0) or a negative error code; the message of the
 * last failure on the calling thread is available from m3d_last_error().
 * Unlike the reference (kernel-launch errors are printf'd and dropped,
 * model/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:331-335, lib/nms/nms_kernel.cu:12-19)
 * errors are always reported to the caller.
 *
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef M3DSSD_B200_H_
#define M3DSSD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* m3d_stream_t; /* cudaStream_t */

enum {
  M3D_OK = 0,
  M3D_ERR_INVALID = -1,     /* bad argument / shape mismatch (reference: THError, dcn_v2_cuda.c:33-38) */
  M3D_ERR_CUDA = -2,        /* CUDA runtime / launch failure */
  M3D_ERR_UNSUPPORTED = -3, /* configuration outside what the kernels implement */
  M3D_ERR_WORKSPACE = -4    /* workspace too small */
};

/* M3D_BF16X3 is a *precision* (fp32 tensors, 3-part bf16 split on the tensor cores), not a storage type */
enum { M3D_BF16 = 0, M3D_F32 = 1, M3D_BF16X3 = 2 };

const char* m3d_last_error(void);
/* Name of the kernel instantiation the last m3d_conv2d_nhwc call on this thread dispatched to (measurement aid). */
const char* m3d_last_kernel(void);
int m3d_version(void);
/* Cap the number of SMs the persistent tensor-core kernels launched (or graph-captured) after this call occupy;
 * 0 = the whole device.  The engine leaves one SM per image free while the detection tail of the previous batch
 * (one CTA per image, minutes of L2 latency each) runs under the trunk of the next: a persistent CTA needs a
 * whole SM's shared memory, so a grid of all SMs would otherwise run as two waves. */
int m3d_set_sm_limit(int sms);
/* Programmatic dependent launch for the kernels launched / captured from now on (default on).  Switch it off while two
 * streams launch persistent kernels concurrently: a successor grid blocked in griddepcontrol.wait must never hold SMs
 * that unscheduled CTAs of its predecessor need. */
int m3d_set_pdl(int on);

/* ------------------------------------------------------------------------
 * Engine-level convolution / DCNv2 on NHWC activations (tcgen05 implicit GEMM).
 *
 * Replaces, per layer, the reference's conv + BN + LeakyReLU (+ residual /
 * concat) module chains (model/pose_dla_dcn.py:93-121, 251-269) and, with
 * `om` set, DCNv2Function.forward -> dcn_v2_cuda_forward
 * (model/DCNv2/dcn_v2_func.py:22-38, model/DCNv2/src/dcn_v2_cuda.c:10-102).
 *
 *   out[n,p,q,co] = lrelu_slope( sum_{i,r,s,c} in_i[n, p*stride-pad+r*dil (+dh), q*stride-pad+s*dil (+dw), c]
 *                                              (* mask) * W[co; i,r,s,c] + bias[co] (+ res[n,p,q,co]) )
 *
 * act_dtype M3D_BF16: bf16 activations, bf16 weights, fp32 accumulate.
 * act_dtype M3D_F32 : fp32 activations, fp32 output.  Two arithmetic modes:
 *    weight_f32 set  -> reference-accuracy mode: CUDA-core implicit GEMM, IEEE fp32 FMA accumulation
 *                       (the arithmetic class of the reference's cuDNN / SGEMM fp32 path);
 *    weight_mid/lo set -> "bf16x3": weights as three bf16 parts (24 mantissa bits), six tcgen05 partial
 *                       products per k-step, fp32 accumulation in TMEM (~3e-6 relative per layer:
 *                       tensor-core accumulators truncate).
 * Weights are packed [rows][K] with K = concat_i (tap-major, channel-minor).
 * ---------------------------------------------------------------------- */
#define M3D_MAX_CONCAT 6

typedef struct m3d_conv_desc {
  int act_dtype; /* M3D_BF16 | M3D_F32 */
  int out_dtype; /* M3D_BF16 | M3D_F32 (M3D_F32 required when act_dtype is M3D_F32) */
  int num_inputs;
  const void* in[M3D_MAX_CONCAT]; /* NHWC buffers */
  int in_c[M3D_MAX_CONCAT];       /* channels consumed from input i (multiple of the k-block) */
  int in_cstride[M3D_MAX_CONCAT]; /* channels per pixel of the buffer */
  int in_coff[M3D_MAX_CONCAT];    /* first channel */
  int in_goff[M3D_MAX_CONCAT];    /* extra channel offset per group */
  int N, H, W;                    /* input geometry */
  int R, S, stride, pad, dil;
  int out_h, out_w; /* 0 = (H + 2 pad - dil (R-1) - 1) / stride + 1; set to crop the output (pad is then top/left only) */
  int Cout;   /* output channels per group */
  int groups; /* independent GEMMs sharing geometry (batched heads); >1 only for plain bf16 convs */
  const void* weight;     /* bf16 [weight_rows][K] (M3D_F32: the high 8 mantissa bits) */
  const void* weight_mid; /* M3D_F32 only: bf16 of the next 8 mantissa bits */
  const void* weight_lo;  /* M3D_F32 only: bf16 of the last 8 mantissa bits */
  const float* weight_f32; /* M3D_F32 reference-accuracy mode: fp32 [weight_rows][K]; `weight` may be NULL */
  int weight_rows;       /* total rows in the packed matrix */
  int weight_goff;       /* row offset per group */
  const float* bias;     /* [groups * bias_goff] or NULL */
  int bias_goff;
  const void* res; /* residual, activation dtype, or NULL */
  int res_cstride, res_coff, res_goff;
  void* out;
  int out_cstride, out_coff, out_goff;
  float slope; /* LeakyReLU slope, 1.0f = none */
  /* deformable part: fp32 NHWC [N,P,Q,om_cstride]; channels [0,2RS) = (dh,dw)
   * per tap in the reference's order (dcn_v2_im2col_cuda.cu:155-156), [2RS,3RS)
   * = modulation mask (logits if sigmoid_mask).  NULL = plain convolution. */
  const float* om;
  int om_cstride;
  int sigmoid_mask;
  int force_gather; /* testing: route a plain conv through the gather producer */
  /* Optional hint (0 = none): bit j set = the j-th 16-element slice of K (packed K index 16 j .. 16 j + 15) is zero in
   * EVERY weight row, so its k-step may be skipped.  The 2x2 space-to-depth rewrite of a 3x3 conv (level0) has 20 such
   * slices out of 36.  A kernel that cannot use the hint ignores it; a wrong hint gives wrong results. */
  unsigned long long k16_zero[2];
} m3d_conv_desc;

int m3d_conv2d_nhwc(const m3d_conv_desc* desc, m3d_stream_t stream);
/* sizeof(m3d_conv_desc) in the library: a binding written in another language compares it with the size of its own
 * mirror of the struct before the first call (fields are only ever appended). */
size_t m3d_conv_desc_size(void);

/* ------------------------------------------------------------------------
 * DCNv2 operator, reference FFI shape (replaces dcn_v2_cuda_forward,
 * model/DCNv2/src/dcn_v2_cuda.h:9-17): NCHW fp32 device tensors,
 * offset [B, 2*kh*kw, Ho, Wo] (dh, dw interleaved per tap), mask [B, kh*kw, Ho, Wo].
 * Differences: `ones`/`columns` scratch tensors are gone (the caller passes one
 * opaque workspace of m3d_dcn_v2_forward_workspace() bytes), shape errors are
 * returned (reference: THError, dcn_v2_cuda.c:33-38), the batch is handled in
 * one launch.  precision: M3D_F32 = reference accuracy (IEEE fp32 FMA), M3D_BF16X3 = 3-part bf16 split on tensor cores,
 * M3D_BF16 = bf16 operands / fp32 accumulate.  deformable_group > 1 (offset [B, dg*2*kh*kw, Ho, Wo], mask
 * [B, dg*kh*kw, Ho, Wo], dcn_v2_im2col_cuda.cu:139-149) is evaluated as the sum of the dg single-group operators on
 * channel slices; C must be a multiple of deformable_group.
 * ---------------------------------------------------------------------- */
size_t m3d_dcn_v2_forward_workspace(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
                                    int deformable_group, int precision);
int m3d_dcn_v2_forward(const float* input, const float* weight, const float* bias, const float* offset,
                       const float* mask, float* output, int B, int C, int H, int W, int Cout, int kh, int kw,
                       int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int deformable_group,
                       int precision, void* workspace, size_t workspace_bytes, m3d_stream_t stream);

/* DCNv2 backward, reference FFI shape (replaces dcn_v2_cuda_backward, dcn_v2_cuda.h:19-29).  All five
 * gradients are OVERWRITTEN (the reference accumulates into buffers its Python wrapper zero-fills,
 * dcn_v2_func.py:44-48).  fp32; grad_input / grad_weight / grad_bias use float atomics, so -- as in the
 * reference -- the summation order is not reproducible bit for bit.  deformable_group >= 1 as in the forward. */
size_t m3d_dcn_v2_backward_workspace(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
                                     int deformable_group);
int m3d_dcn_v2_backward(const float* input, const float* weight, const float* offset, const float* mask,
                        const float* grad_output, float* grad_input, float* grad_weight, float* grad_bias,
                        float* grad_offset, float* grad_mask, int B, int C, int H, int W, int Cout, int kh, int kw,
                        int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int deformable_group,
                        int precision /* M3D_F32: GEMMs on the CUDA cores in IEEE fp32; M3D_BF16X3: W^T dY on the tensor cores (3-part
                                         split); M3D_BF16: W^T dY and dW (m3d_conv2d_wgrad on the sampled columns) on
                                         the tensor cores with bf16 operands; every gradient deterministic (fixed-point col2im
                                         accumulation, ordered split-K and channel sums) */,
                        void* workspace, size_t workspace_bytes, m3d_stream_t stream);

/* ------------------------------------------------------------------------
 * Training path (BASELINE config 4; scripts/train_rpn_3d.py:204-218 runs loss.backward() through cuDNN): weight
 * gradient of a convolution on the tensor cores,
 *     dW[co][ci][r][s] = sum_{n,p,q} gy[n,p,q,co] * x[n, p*stride - pad + r*dil, q*stride - pad + s*dil, ci],
 * x / gy bf16 NHWC (channel strides multiples of 8), dW fp32 in torch's [Cout][Cin][R][S] layout, fp32 accumulation,
 * deterministic (split-K partials reduced in a fixed order).  The input gradient of a convolution is itself a
 * convolution (m3d_conv2d_nhwc on gy with the flipped, transposed weights).
 * ---------------------------------------------------------------------- */
size_t m3d_conv2d_wgrad_workspace(int N, int P, int Q, int Cin, int Cout, int R, int S);
int m3d_conv2d_wgrad(const void* x, int x_cstride, int x_coff, const void* gy, int gy_cstride, int gy_coff, float* dw,
                     int N, int H, int W, int Cin, int P, int Q, int Cout, int R, int S, int stride, int pad, int dil,
                     void* workspace, size_t workspace_bytes, m3d_stream_t stream);

/* Bias gradient: out[c] = sum over the npix pixels of x[pix][coff + c] (bf16 NHWC, fp32 sums, deterministic). */
size_t m3d_channel_sum_workspace(int C);
int m3d_channel_sum(const void* x, long npix, int C, int cstride, int coff, float* out, void* workspace,
                    size_t workspace_bytes, m3d_stream_t stream);

/* ------------------------------------------------------------------------
 * gpu_nms.  m3d_nms is the drop-in for `_nms` (lib/nms/gpu_nms.hpp:1-2): HOST
 * pointers, boxes [n, boxes_dim] already sorted by score, keep_out receives the
 * kept row indices; synchronous.  The +1 pixel IoU convention and the strict
 * `IoU > thresh` test follow lib/nms/nms_kernel.cu:24-32,71 in the reference's
 * compiled operation order.  m3d_nms_batched is the device-resident, batched
 * form used by the detection tail: boxes [batch, max_n, box_stride] (x1,y1,x2,y2
 * first), num[b] valid rows (NULL = max_n), keep [batch, max_n], num_keep [batch].
 * ---------------------------------------------------------------------- */
int m3d_nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
            float nms_overlap_thresh, int device_id);
size_t m3d_nms_workspace_bytes(int batch, int max_n);
int m3d_nms_batched(const float* boxes, int box_stride, const int* num, int batch, int max_n, float thresh,
                    void* workspace, size_t workspace_bytes, int* keep, int* num_keep, m3d_stream_t stream);

/* Detection decode (lib/rpn_util.py:1444-1521): per image, exact top-`topk` of
 * score (descending, ties by lower index), anchor decode of the selected rows
 * into dets [batch, topk, 14] = x1,y1,x2,y2,score,cls,x3d,y3d,z3d,w3d,h3d,l3d,ry3d,anchor. */
size_t m3d_decode_topk_workspace(int batch);
int m3d_decode_topk(const float* score, const unsigned char* cls_pred, const float* bbox_2d, const float* bbox_3d,
                    const float* anchors /*[A,9] device*/, const float* means11 /*host*/, const float* stds11 /*host*/, int batch, int A, int H,
                    int W, float feat_stride, float scale_factor, int topk, float* dets, int* det_idx, int* det_num,
                    void* workspace /*device, 16-byte aligned, m3d_decode_topk_workspace(batch) bytes, caller-owned*/,
                    size_t workspace_bytes, m3d_stream_t stream);
/* Same selection and decode with the regression outputs read where the heads wrote them: the NHWC buffer
 * [batch, H, W, heads_cstride] in which output j (x,y,w,h,x3d,y3d,z3d,w3d,h3d,l3d,rY3d) of anchor a at a pixel is
 * channel slot_of_output[j] * A + a.  Only the <= topk selected rows are read, so the detection path never
 * materialises the reference's flattened bbox_2d / bbox_3d (lib/rpn_util.py:892-901): results are bit-identical to
 * m3d_flatten_heads followed by m3d_decode_topk. */
int m3d_decode_topk_heads(const float* score, const unsigned char* cls_pred, const float* heads, int heads_cstride,
                          const int* slot_of_output /*host [11]*/, const float* anchors, const float* means11,
                          const float* stds11, int batch, int A, int H, int W, float feat_stride, float scale_factor,
                          int topk, float* dets, int* det_idx, int* det_num, void* workspace, size_t workspace_bytes,
                          m3d_stream_t stream);
int m3d_gather_kept(const float* dets, int row_len, int batch, int max_n, const int* keep, const int* num_keep,
                    int max_out, float* out, m3d_stream_t stream);

/* ------------------------------------------------------------------------
 * Bandwidth-bound layers around the GEMMs (NHWC activations).
 * ---------------------------------------------------------------------- */
/* DLA.base_layer (model/pose_dla_dcn.py:336-340): 7x7 conv on the NCHW fp32 image, BN folded, LeakyReLU. */
int m3d_stem_conv7x7(const float* image_nchw, const float* weight /*[16,3,7,7]*/, const float* bias, void* out,
                     int out_dtype, int out_cstride, int N, int H, int W, float slope, m3d_stream_t stream);
/* Input pipeline on the device (the reference's Normalize + BGR->RGB + HWC->CHW, lib/augmentations.py:44-57,
 * lib/dataloader.py:942-950): uint8 HWC images [N,H,W,3] as cv2.imread returns them -> fp32 NCHW
 * out[n][swap_rb ? 2-c : c][h][w] = ((u8 / 255 - mean3[c]) / std3[c]), fp32 IEEE in that order (bit-identical to
 * numpy).  mean3 / std3 are host arrays indexed by the INPUT channel.  4x less host->device traffic than fp32. */
int m3d_preprocess_u8(const unsigned char* image_hwc, float* out_nchw, int N, int H, int W, const float* mean3,
                      const float* std3, int swap_rb, m3d_stream_t stream);
/* The reference's whole test-time transform for a RAGGED batch: Preprocess = ConvertToFloat + Padding(size) + Normalize
 * (lib/augmentations.py:472-492; Padding :136-160 = cv2.copyMakeBorder bottom / right with 0) + BGR->RGB + HWC->CHW
 * (lib/dataloader.py:904,942-950).  Image n is uint8 HWC [heights[n], widths[n], 3] at images + offsets[n] (device
 * memory, any alignment); offsets / heights / widths are HOST arrays.  The 0 padding goes through Normalize like the
 * reference's (padded value = -mean/std).  An image larger than H x W is an error, as in cv2 (M3D_ERR_INVALID).
 * Bit-identical to numpy. */
int m3d_preprocess_u8_pad(const unsigned char* images, const long long* offsets, const int* heights, const int* widths,
                          float* out_nchw, int N, int H, int W, const float* mean3, const float* std3, int swap_rb,
                          m3d_stream_t stream);
/* The same layer on the tensor cores, written in 2x2 space-to-depth form: out [N, H/2, W/2, 64] bf16 with
 * channel (dy*2 + dx)*16 + c = stem output channel c at pixel (2Y+dy, 2X+dx).  The 7x7 conv becomes a
 * K = 3*8*8 implicit GEMM (stride-2 8x8 windows of the fp32 NCHW image gathered straight into the
 * swizzled A tile).  weight: bf16 [64][192] packed by the host (m3dssd_b200.ops.pack_stem_s2d). */
int m3d_stem_conv7x7_s2d(const float* image_nchw, const void* weight_bf16, const float* bias64, void* out, int N,
                         int H, int W, float slope, m3d_stream_t stream);
/* nn.MaxPool2d(2) (model/pose_dla_dcn.py:306). */
int m3d_maxpool2x2_nhwc(const void* in, void* out, int dtype, int N, int H, int W, int C, int in_cstride,
                        int out_cstride, m3d_stream_t stream);
/* IDAUp up_i + skip add (model/pose_dla_dcn.py:536-552): depthwise ConvTranspose2d(2f, stride f, pad f/2). */
int m3d_upsample_add_nhwc(const void* x, const float* weight /*tap-major [(2f)^2, C]*/, const void* skip, void* out, int dtype,
                          int N, int H, int W, int C, int f, int x_cstride, int skip_cstride, int out_cstride,
                          m3d_stream_t stream);
/* Backward of the same up-sampling (training path; bf16 NHWC, dense channels): gx = depthwise strided conv of gy with the
 * weights, gw [(2f)^2, C] fp32 (tap-major) = per-tap correlation of gy with x. */
size_t m3d_upsample_backward_workspace(int C, int f);
int m3d_upsample_backward(const void* gy, const void* x, const float* weight, void* gx, float* gw, int N, int H, int W,
                          int C, int f, void* workspace, size_t workspace_bytes, m3d_stream_t stream);
/* softmax over classes + fg prob + top-1 anchor + score/class (model/M3d_inference_align.py:229-234).
 * cls_out / prob_out (the flattened copies RPN.forward returns) may both be NULL: the detection path does not read them. */
int m3d_cls_softmax(const float* logits, int logits_cstride, int N, int H, int W, int A, int K, float* cls_out,
                    float* prob_out, float* fg_max, int* fg_arg, float* score, unsigned char* cls_pred,
                    m3d_stream_t stream);
/* The same with m3d_shape_align_om folded into the per-pixel tail (one launch less; identical values). */
int m3d_cls_softmax_shape_om(const float* logits, int logits_cstride, int N, int H, int W, int A, int K, float* cls_out,
                             float* prob_out, float* fg_max, int* fg_arg, float* score, unsigned char* cls_pred,
                             const float* anchors, int anchor_ld, float feat_stride, float thresh,
                             float* shape_om /*[N*H*W,27]*/, m3d_stream_t stream);
/* shape_align / center_align offset builders (model/module/feturealign_mgpu.py:119-136,58-77). */
int m3d_shape_align_om(const float* fg_max, const int* fg_arg, const float* anchors, int anchor_ld, float feat_stride,
                       float thresh, float* om /*[npix,27]*/, long npix, m3d_stream_t stream);
int m3d_center_align_om(const float* fg_max, const int* fg_arg, const float* heads, int heads_cstride, int x_coff,
                        int y_coff, const float* anchors, int anchor_ld, float feat_stride, float mean_x, float mean_y,
                        float std_x, float std_y, float thresh, float* om, int om_cstride, long npix,
                        m3d_stream_t stream);
/* Both centre alignments of a pixel in one launch: xy_coff4 = {x_coff, y_coff} of om_a then of om_b, mean4 / std4 =
 * {mean_x, mean_y} / {std_x, std_y} likewise (host arrays).  Same values as two m3d_center_align_om calls. */
int m3d_center_align_om2(const float* fg_max, const int* fg_arg, const float* heads, int heads_cstride,
                         const int* xy_coff4, const float* mean4, const float* std4, const float* anchors, int anchor_ld,
                         float feat_stride, float thresh, float* om_a, float* om_b, int om_cstride, long npix,
                         m3d_stream_t stream);
/* G three-layer 1x1 regression heads that share the input x, fused in one kernel (conv1x1 + BN + LeakyReLU,
 * conv1x1 + BN + LeakyReLU, conv1x1: model/M3d_inference_align.py:66-210, 236-277; BatchNorm folded into w/b).
 * x: bf16 NHWC [N,H,W,x_cstride], channels [x_coff, x_coff+Cx), Cx in {64,128}.  w1: bf16 [G*256][Cx],
 * w2: bf16 [G*256][256], w3: bf16 [G*rows3][256] (rows >= A of each head zero), b1,b2: fp32 [G*256], b3: fp32 [G*rows3].
 * out: fp32 NHWC [N,H,W,out_cstride]; head g -> channels [out_coff + g*A, out_coff + (g+1)*A).  The 256-channel
 * intermediates stay in shared memory / TMEM (rounded to bf16 between layers, like the layer-by-layer bf16 path). */
int m3d_head_mlp(const void* x, int N, int H, int W, int x_cstride, int x_coff, int Cx, const void* w1, const float* b1,
                 const void* w2, const float* b2, const void* w3, const float* b3, int G, int A, int rows3, float* out,
                 int out_cstride, int out_coff, float slope, m3d_stream_t stream);
/* flatten_tensor + cat of the 11 regression heads (model/M3d_inference_align.py:280-295). */
int m3d_flatten_heads(const float* heads, int heads_cstride, int N, int H, int W, int A,
                      const int* slot_of_output /*host, 11 ints: slot of x,y,w,h,x3d,y3d,z3d,w3d,h3d,l3d,rY3d*/,
                      float* bbox_2d, float* bbox_3d, m3d_stream_t stream);
/* Post-NMS 3D refinement: the per-box loop of test_kitti_3d with hill_climb / test_projection / project_3d
 * (lib/rpn_util.py:1801-1852, 652-708, 2015-2050, 921-970; lib/util.py:516-535), float64 like the reference's numpy,
 * one thread per kept box.  kept: fp32 [B, max_out, row_len >= 13] rows from m3d_gather_kept; p2 / p2_inv: float64
 * [B, 16] (row-major 4x4, device).  out: float64 [B, max_out, 14] = (class index, alpha, x1, y1, x2, y2, h3d, w3d,
 * l3d, x3d, y3d, z3d, ry3d, score), the numbers of the KITTI result line; valid[B, max_out] = passed the score cut. */
int m3d_refine_3d(const float* kept, const int* num_keep, int B, int max_out, int row_len, const double* p2,
                  const double* p2_inv, float score_thresh, int hill_climbing, double step_r_init, double r_lim,
                  double* out, int* valid, m3d_stream_t stream);
/* Training targets on the device: compute_targets (lib/rpn_util.py:430-532) + the per-image
 * post-processing of Dataset._targets (lib/dataloader.py:1014-1144) for a whole batch.  Ground truth per image, padded
 * to max_gts / max_ign rows (device, float64 as numpy holds them): gts_val [B, max_gts, 4] (x1,y1,x2,y2), gts_3d
 * [B, max_gts, 7] (cx, cy, z, w, h, l, rotY), box_lbls [B, max_gts] (class index >= 1), n_val [B]; ignore regions gts_ign
 * [B, max_ign, 4], n_ign [B].  The anchors' boxes are rebuilt from (anchor, row, column) as locate_anchors does.
 * Outputs in the reference's `imobjs` layout, M = A*H*W rows per image: labels_fg / labels_bg / labels_ign (0/1 bytes),
 * labels (0 background, class, 3000 ignored), bbox_2d [B,M,4], bbox_3d [B,M,7] ((t - mean) / std, float32), any_val [B].
 * means11 / stds11: host arrays (conf.bbox_means[0], conf.bbox_stds[0]). */
size_t m3d_compute_targets_workspace(int batch, int max_gts);
int m3d_compute_targets(const double* gts_val, const double* gts_3d, const int* box_lbls, const int* n_val, int max_gts,
                        const double* gts_ign, const int* n_ign, int max_ign, const float* anchors /*[A,9] device*/,
                        int batch, int A, int H, int W, float feat_stride, double fg_thresh, double ign_thresh,
                        double bg_thresh_lo, double bg_thresh_hi, double best_thresh, const float* means11,
                        const float* stds11, unsigned char* labels_fg, unsigned char* labels_bg, unsigned char* labels_ign,
                        long long* labels, float* bbox_2d, float* bbox_3d, unsigned char* any_val, void* workspace,
                        size_t workspace_bytes, m3d_stream_t stream);
/* layout conversion for the NCHW-facing operators */
int m3d_nchw_to_nhwc(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                     int out_cstride, int out_coff, m3d_stream_t stream);
int m3d_nhwc_to_nchw(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                     int in_cstride, int in_coff, m3d_stream_t stream);

/* ------------------------------------------------------------------------
 * ANAB (model/module/attention.py:120-216).  kvs = fp32 NHWC output of the
 * concatenated key | value | spatial 1x1 convolutions.  m3d_anab_pool builds
 * the T = sum(size^2) attention-weighted pyramid tokens (PAPAModule :120-147);
 * m3d_anab_attention computes LeakyReLU(BN(softmax(Q K^T) V + x)) with the
 * eval-mode BatchNorm given as per-channel (scale, shift).
 * ---------------------------------------------------------------------- */
size_t m3d_anab_pool_workspace(int N, int H, int nlev, const int* sizes, int ck, int cv);
int m3d_anab_pool(const float* kvs, int kvs_cstride, int N, int H, int W, int ck, int cv, int nlev,
                  const int* sizes /*host*/, void* workspace, size_t workspace_bytes, float* ktok /*[N,T,ck]*/,
                  float* vtok /*[N,T,cv]*/, m3d_stream_t stream);
size_t m3d_anab_attention_workspace(int N, int act_dtype);
int m3d_anab_attention(const void* q, int q_cstride, const float* ktok, const float* vtok, const void* x,
                       int x_cstride, int act_dtype, const float* scale, const float* shift, float slope, void* out,
                       int out_cstride, int N, int HW, int ck, int cv, int T,
                       void* workspace /*device, caller-owned, m3d_anab_attention_workspace(N, act_dtype) bytes*/,
                       size_t workspace_bytes, m3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* M3DSSD_B200_H_ */
