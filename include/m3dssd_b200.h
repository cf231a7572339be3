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

enum { M3D_BF16 = 0, M3D_F32 = 1 };

const char* m3d_last_error(void);
int m3d_version(void);

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
 * act_dtype M3D_F32 : fp32 activations; weights given as bf16 hi + lo parts;
 *                     products formed as hi*hi + lo*hi + hi*lo (fp32-accurate).
 * Weights are packed [rows][K] with K = concat_i (tap-major, channel-minor).
 * ---------------------------------------------------------------------- */
#define M3D_MAX_CONCAT 4

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
  int Cout;   /* output channels per group */
  int groups; /* independent GEMMs sharing geometry (batched heads); >1 only for plain bf16 convs */
  const void* weight;    /* bf16 [weight_rows][K] */
  const void* weight_lo; /* bf16 low parts (M3D_F32 only) */
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
} m3d_conv_desc;

int m3d_conv2d_nhwc(const m3d_conv_desc* desc, m3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* M3DSSD_B200_H_ */
