// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC activations.
//
//   out[n,p,q,co] = act( sum_{r,s,ci} A(n,p,q; r,s,ci) * Wt[co; r,s,ci] + bias[co] (+ res[n,p,q,co]) )
//
// The GEMM view is M = output pixels (tiles of 128 = TH x TW pixels of one
// image), N = Cout tile (BN), K = taps x input channels, walked in k-blocks of
// 64 bf16 (one 128-byte swizzled row per pixel).  Accumulators live in TMEM
// (two stages so the epilogue of tile i overlaps the main loop of tile i+1);
// kernels are persistent (grid = #SMs, static round-robin tile schedule).
//
// Two producers for the A operand:
//   * conv_tma_kernel    - plain convolution: TMA box loads of the (shifted,
//                          strided) NHWC window, zero fill outside the image.
//   * conv_gather_kernel - DCNv2: producer warps bilinear-sample the input at
//                          the learned / computed offsets, modulate by the
//                          mask and write the swizzled bf16 tile themselves
//                          (no im2col buffer in HBM).  Also the fp32-accurate
//                          mode: fp32 activations and weights are split into
//                          three bf16 parts (hi + mid + lo = 24 mantissa bits)
//                          and contracted with the 6 MMAs whose terms are
//                          >= 2^-16 of the product (error ~2^-23: fp32 quality).
//
// Semantics follow the reference's modulated_deformable_im2col_gpu_kernel +
// SGEMM (model/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:118-180,
// model/DCNv2/src/dcn_v2_cuda.c:61-97) and torch's conv2d for plain convs.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace m3d {

constexpr int kMaxConcat = 6;  // dla102: the innermost Root of a 4-level Tree reads 6 tensors
constexpr int kTileM = 128;

enum : int { DT_BF16 = 0, DT_F32 = 1 };

struct alignas(64) ConvTmaParams {
  CUtensorMap tmap_a[kMaxConcat];  // NHWC inputs as 4-D (C, W, H, N) maps, box {BK, TW*stride, TH*stride, 1}
  CUtensorMap tmap_b;              // packed weights as 3-D (BK, rows, K/BK) map, box {BK, BN, KSUB}
  CUtensorMap tmap_out, tmap_res;  // staged epilogue: output / residual as (C, Q, P, N), box {64, TW, TH, 1}
  int num_inputs;
  int chunks[kMaxConcat];  // k-blocks per tap for each concat input (= C_i / BK)
  int a_coff[kMaxConcat];  // first channel of input i inside its buffer
  int a_goff[kMaxConcat];  // + group * a_goff
  int R, S, stride, pad, dil;
  int N, P, Q;  // output geometry
  int TW, TH, tiles_w, tiles_h;
  int Cout;     // valid output channels per group
  int n_tiles;  // ceil(Cout / BN)
  int groups;
  int b_goff;  // weight row offset per group
  void* out;
  int out_cstride, out_coff, out_goff;
  const float* bias;  // may be null; indexed [g * bias_goff + co]
  int bias_goff;
  const void* res;  // may be null; activation dtype
  int res_cstride, res_coff, res_goff;
  float slope;  // LeakyReLU negative slope; 1.0 = identity
  int total_tiles;
  int dbg;     // development probe bits (M3D_DBG): 1 skip A loads, 4 skip the epilogue body, 8 skip MMAs, 16 skip the TMA store
  int a_wide;  // tmap_a are 5-D (BK, W, H, N, C/BK) maps whose box holds the stage's KSUB channel chunks
  // conv_halo resident-weight kernel, one chunk: bit (r * 3 + s) * 4 + k set = that k-step's weights are all zero
  // (m3d_conv_desc::k16_zero) and its MMA is not issued
  unsigned long long kzero;
};

struct alignas(64) ConvGatherParams {
  CUtensorMap tmap_b;      // bf16 weights (hi part in split mode)
  CUtensorMap tmap_b_mid;  // split mode: second 8 mantissa bits
  CUtensorMap tmap_b_lo;   // split mode: third 8 mantissa bits
  CUtensorMap tmap_out, tmap_res;  // staged epilogue (bf16 out, Cout % 64 == 0)
  CUtensorMap tmap_img;            // stem mode: fp32 NCHW image as (W, H, 3, N), box {2TW+8, 2TH+6, 3, 1}
  int num_inputs;
  const void* in[kMaxConcat];
  int in_cstride[kMaxConcat], in_coff[kMaxConcat], chunks[kMaxConcat];
  int H, W;  // input geometry
  int R, S, stride, pad, dil;
  int N, P, Q;
  int TW, TH, tiles_w, tiles_h;
  int Cout, n_tiles;
  // DCN offsets / mask: fp32 NHWC [N,P,Q,om_cstride]; channels [0,2RS) =
  // (dh,dw) per tap, [2RS,3RS) = mask.  null => plain convolution.
  const float* om;
  int om_cstride;
  int sigmoid_mask;
  // stem mode: A rows are stride-2 8x8 windows of this fp32 NCHW 3-channel image (k-block = channel)
  const float* stem_img;
  void* out;
  int out_cstride, out_coff;
  const float* bias;
  const void* res;
  int res_cstride, res_coff;
  float slope;
  int total_tiles;
  // dcn_fused.cu, staged-window mode: tmap_img is then the bf16 NHWC input as (C, W, H, N) with box
  // {64, halo_w, halo_h, 1}; the window of a tile starts halo_x / halo_y pixels left of / above it
  int halo_w, halo_h, halo_x, halo_y;
};

// Host launchers (igemm.cu).  Return 0 or a negative m3d error code.
int launch_conv_tma(const ConvTmaParams& p, int BN, int BK, int ksub, int out_dtype, bool staged,
                    cudaStream_t stream);
int launch_conv_gather(const ConvGatherParams& p, int BN, int in_dtype, int out_dtype, bool staged,
                       cudaStream_t stream);

// Dedicated bf16 DCNv2 kernel (dcn_fused.cu): 16 producer warps for the bilinear blend.
bool dcn_fused_supported(const ConvGatherParams& p, int BN, int in_dtype, int out_dtype);
int launch_dcn_fused(const ConvGatherParams& p, int BN, cudaStream_t stream);

}  // namespace m3d
