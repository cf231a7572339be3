// m3d_dcn_v2_forward: drop-in for the reference FFI entry dcn_v2_cuda_forward
// (model/DCNv2/src/dcn_v2_cuda.h:9-17, dcn_v2_cuda.c:10-102) on raw NCHW fp32
// device pointers.  No per-sample host loop, no `ones`/`columns` scratch in HBM:
// inputs are re-laid out as NHWC in the caller's workspace and the fused
// gather+tcgen05 kernel (igemm.cu) does bias + sampling + contraction.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdint>

#include "common.cuh"

namespace m3d {

// [Cout, Cin, kh, kw] fp32 -> [Cout][tap][Cpad] bf16 (hi, lo), zero channel padding.
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo, int Cout, int Cin,
                                   int KK, int Cpad) {
  const long total = static_cast<long>(Cout) * KK * Cpad;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const int t = static_cast<int>((i / Cpad) % KK);
    const int co = static_cast<int>(i / (static_cast<long>(Cpad) * KK));
    float v = 0.f;
    if (c < Cin) v = w[(static_cast<long>(co) * Cin + c) * KK + t];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (mid != nullptr) {
      const float r1 = v - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      mid[i] = m;
      lo[i] = __float2bfloat16_rn(r1 - __bfloat162float(m));
    }
  }
}

__global__ void pack_weight_f32_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KK,
                                       int Cpad) {
  const long total = static_cast<long>(Cout) * KK * Cpad;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const int t = static_cast<int>((i / Cpad) % KK);
    const int co = static_cast<int>(i / (static_cast<long>(Cpad) * KK));
    out[i] = c < Cin ? w[(static_cast<long>(co) * Cin + c) * KK + t] : 0.f;
  }
}

static inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct DcnLayout {
  int Ho, Wo, Cpad, KK;
  size_t x_bytes, om_bytes, w_bytes, out_bytes, total;
};

static DcnLayout dcn_layout(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
                            int precision) {
  DcnLayout L;
  L.Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
  L.Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
  L.Cpad = (C + 63) / 64 * 64;
  L.KK = kh * kw;
  const size_t esz = precision == M3D_BF16 ? 2 : 4;
  L.x_bytes = align256(static_cast<size_t>(B) * H * W * L.Cpad * esz);
  L.om_bytes = align256(static_cast<size_t>(B) * L.Ho * L.Wo * 3 * L.KK * 4);
  L.w_bytes = align256(static_cast<size_t>(Cout) * L.KK * L.Cpad * 2);  // one bf16 part; fp32 weights use two slots
  L.out_bytes = align256(static_cast<size_t>(B) * L.Ho * L.Wo * Cout * 4);
  L.total = L.x_bytes + L.om_bytes + 3 * L.w_bytes + L.out_bytes;
  return L;
}

}  // namespace m3d

extern "C" int m3d_nchw_to_nhwc(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                                int out_cstride, int out_coff, m3d_stream_t stream);
extern "C" int m3d_nhwc_to_nchw(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                                int in_cstride, int in_coff, m3d_stream_t stream);

namespace m3d {
// Slices for deformable_group > 1 (dcn_v2_im2col_cuda.cu:139-149: channel c belongs to group c / (C / dg), and group g
// of sample b reads offset channels [(b*dg + g) * 2*kh*kw, ...) and mask channels [(b*dg + g) * kh*kw, ...)): the
// operator is the SUM over groups of single-group operators on channel slices, which is how it is evaluated here.
struct GroupSlices {
  size_t x, off, msk, w, out, total;
};
static inline size_t al256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
GroupSlices group_slices(int B, int Cg, int H, int W, int Cout, int KK, int Ho, int Wo) {
  GroupSlices g;
  g.x = align256(static_cast<size_t>(B) * Cg * H * W * 4);
  g.off = align256(static_cast<size_t>(B) * 2 * KK * Ho * Wo * 4);
  g.msk = align256(static_cast<size_t>(B) * KK * Ho * Wo * 4);
  g.w = align256(static_cast<size_t>(Cout) * Cg * KK * 4);
  g.out = align256(static_cast<size_t>(B) * Cout * Ho * Wo * 4);
  g.total = g.x + g.off + g.msk + g.w + g.out;
  return g;
}

__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    y[i] += x[i];
}

}  // namespace m3d
using namespace m3d;

extern "C" size_t m3d_dcn_v2_forward_workspace(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad,
                                               int dil, int deformable_group, int precision) {
  const int dg = deformable_group < 1 ? 1 : deformable_group;
  if (dg == 1) return dcn_layout(B, C, H, W, Cout, kh, kw, stride, pad, dil, precision).total;
  const DcnLayout L = dcn_layout(B, C / dg, H, W, Cout, kh, kw, stride, pad, dil, precision);
  return L.total + group_slices(B, C / dg, H, W, Cout, kh * kw, L.Ho, L.Wo).total;
}

static int dcn_forward_one_group(const float* input, const float* weight, const float* bias, const float* offset,
                                 const float* mask, float* output, int B, int C, int H, int W, int Cout, int kh, int kw,
                                 int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int precision,
                                 void* workspace, size_t workspace_bytes, m3d_stream_t stream_);

extern "C" int m3d_dcn_v2_forward(const float* input, const float* weight, const float* bias, const float* offset,
                                  const float* mask, float* output, int B, int C, int H, int W, int Cout, int kh, int kw,
                                  int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                  int deformable_group, int precision, void* workspace, size_t workspace_bytes,
                                  m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int dg = deformable_group;
  if (dg == 1)
    return dcn_forward_one_group(input, weight, bias, offset, mask, output, B, C, H, W, Cout, kh, kw, stride_h, stride_w,
                                 pad_h, pad_w, dil_h, dil_w, precision, workspace, workspace_bytes, stream_);
  M3D_REQUIRE(input && weight && offset && mask && output, "NULL tensor pointer");
  M3D_REQUIRE(dg >= 1 && C % dg == 0, "channels (%d) must be a multiple of deformable_group (%d)", C, dg);
  const int Cg = C / dg, KK = kh * kw;
  const DcnLayout L = dcn_layout(B, Cg, H, W, Cout, kh, kw, stride_h, pad_h, dil_h, precision);
  M3D_REQUIRE(L.Ho >= 1 && L.Wo >= 1, "empty output");
  const GroupSlices G = group_slices(B, Cg, H, W, Cout, KK, L.Ho, L.Wo);
  if (workspace == nullptr || workspace_bytes < L.total + G.total) {
    set_last_error("DCNv2 workspace too small: %zu < %zu", workspace_bytes, L.total + G.total);
    return M3D_ERR_WORKSPACE;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* xg = reinterpret_cast<float*>(ws);
  float* og = reinterpret_cast<float*>(ws + G.x);
  float* mg = reinterpret_cast<float*>(ws + G.x + G.off);
  float* wg = reinterpret_cast<float*>(ws + G.x + G.off + G.msk);
  float* yg = reinterpret_cast<float*>(ws + G.x + G.off + G.msk + G.w);
  void* inner = ws + G.total;
  const size_t hw = static_cast<size_t>(H) * W * 4, howo = static_cast<size_t>(L.Ho) * L.Wo * 4;
  const long nout = static_cast<long>(B) * Cout * L.Ho * L.Wo;
  for (int g = 0; g < dg; ++g) {
    M3D_CUDA_OK(cudaMemcpy2DAsync(xg, Cg * hw, input + static_cast<size_t>(g) * Cg * H * W, C * hw, Cg * hw, B,
                                  cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(og, 2 * KK * howo, offset + static_cast<size_t>(g) * 2 * KK * L.Ho * L.Wo,
                                  dg * 2 * KK * howo, 2 * KK * howo, B, cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(mg, KK * howo, mask + static_cast<size_t>(g) * KK * L.Ho * L.Wo, dg * KK * howo,
                                  KK * howo, B, cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(wg, static_cast<size_t>(Cg) * KK * 4, weight + static_cast<size_t>(g) * Cg * KK,
                                  static_cast<size_t>(C) * KK * 4, static_cast<size_t>(Cg) * KK * 4, Cout,
                                  cudaMemcpyDeviceToDevice, stream));
    const int rc = dcn_forward_one_group(xg, wg, g == 0 ? bias : nullptr, og, mg, g == 0 ? output : yg, B, Cg, H, W, Cout,
                                         kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, precision, inner,
                                         workspace_bytes - G.total, stream_);
    if (rc) return rc;
    if (g > 0) {
      add_inplace_kernel<<<static_cast<int>(std::min<long>((nout + 255) / 256, 4096)), 256, 0, stream>>>(output, yg, nout);
      M3D_CUDA_OK(cudaGetLastError());
    }
  }
  return M3D_OK;
}

static int dcn_forward_one_group(const float* input, const float* weight, const float* bias, const float* offset,
                                 const float* mask, float* output, int B, int C, int H, int W, int Cout, int kh, int kw,
                                 int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int precision,
                                 void* workspace, size_t workspace_bytes, m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int deformable_group = 1;
  M3D_REQUIRE(input && weight && offset && mask && output, "NULL tensor pointer");
  M3D_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1 && Cout >= 1 && kh >= 1 && kw >= 1, "bad shape");
  if (deformable_group != 1 || stride_h != stride_w || pad_h != pad_w || dil_h != dil_w || kh * kw > 9) {
    set_last_error("DCNv2 configuration not implemented (dg=%d stride=%dx%d pad=%dx%d dil=%dx%d k=%dx%d)",
                   deformable_group, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, kh, kw);
    return M3D_ERR_UNSUPPORTED;
  }
  M3D_REQUIRE(precision == M3D_F32 || precision == M3D_BF16 || precision == M3D_BF16X3,
              "precision must be M3D_F32, M3D_BF16X3 or M3D_BF16");
  const int act = precision == M3D_BF16 ? M3D_BF16 : M3D_F32;
  const DcnLayout L = dcn_layout(B, C, H, W, Cout, kh, kw, stride_h, pad_h, dil_h, precision);
  M3D_REQUIRE(L.Ho >= 1 && L.Wo >= 1, "empty output");
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_last_error("DCNv2 workspace too small: %zu < %zu", workspace_bytes, L.total);
    return M3D_ERR_WORKSPACE;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* x = ws;
  float* om = reinterpret_cast<float*>(ws + L.x_bytes);
  __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(ws + L.x_bytes + L.om_bytes);
  __nv_bfloat16* w_mid = reinterpret_cast<__nv_bfloat16*>(ws + L.x_bytes + L.om_bytes + L.w_bytes);
  __nv_bfloat16* w_lo = reinterpret_cast<__nv_bfloat16*>(ws + L.x_bytes + L.om_bytes + 2 * L.w_bytes);
  float* out_nhwc = reinterpret_cast<float*>(ws + L.x_bytes + L.om_bytes + 3 * L.w_bytes);

  if (L.Cpad != C) M3D_CUDA_OK(cudaMemsetAsync(x, 0, L.x_bytes, stream));
  int rc = m3d_nchw_to_nhwc(input, M3D_F32, x, act, B, C, H, W, L.Cpad, 0, stream_);
  if (rc) return rc;
  rc = m3d_nchw_to_nhwc(offset, M3D_F32, om, M3D_F32, B, 2 * L.KK, L.Ho, L.Wo, 3 * L.KK, 0, stream_);
  if (rc) return rc;
  rc = m3d_nchw_to_nhwc(mask, M3D_F32, om, M3D_F32, B, L.KK, L.Ho, L.Wo, 3 * L.KK, 2 * L.KK, stream_);
  if (rc) return rc;
  {
    const long total = static_cast<long>(Cout) * L.KK * L.Cpad;
    const int grid = static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    if (precision == M3D_F32)
      pack_weight_f32_kernel<<<grid, 256, 0, stream>>>(weight, reinterpret_cast<float*>(w_hi), Cout, C, L.KK, L.Cpad);
    else
      pack_weight_kernel<<<grid, 256, 0, stream>>>(weight, w_hi, precision == M3D_BF16X3 ? w_mid : nullptr, w_lo, Cout,
                                                   C, L.KK, L.Cpad);
    M3D_CUDA_OK(cudaGetLastError());
  }
  m3d_conv_desc d = {};
  d.act_dtype = act;
  d.out_dtype = M3D_F32;
  d.num_inputs = 1;
  d.in[0] = x;
  d.in_c[0] = L.Cpad;
  d.in_cstride[0] = L.Cpad;
  d.N = B, d.H = H, d.W = W;
  d.R = kh, d.S = kw, d.stride = stride_h, d.pad = pad_h, d.dil = dil_h;
  d.Cout = Cout, d.groups = 1;
  if (precision == M3D_F32) {
    d.weight_f32 = reinterpret_cast<const float*>(w_hi);
  } else {
    d.weight = w_hi;
    d.weight_mid = precision == M3D_BF16X3 ? w_mid : nullptr;
    d.weight_lo = precision == M3D_BF16X3 ? w_lo : nullptr;
  }
  d.weight_rows = Cout;
  d.bias = bias;
  d.out = out_nhwc, d.out_cstride = Cout;
  d.slope = 1.0f;
  d.om = om, d.om_cstride = 3 * L.KK, d.sigmoid_mask = 0;
  rc = m3d_conv2d_nhwc(&d, stream_);
  if (rc) return rc;
  return m3d_nhwc_to_nchw(out_nhwc, M3D_F32, output, M3D_F32, B, Cout, L.Ho, L.Wo, Cout, 0, stream_);
}
