// m3d_dcn_v2_backward: replacement for the reference FFI entry dcn_v2_cuda_backward
// (model/DCNv2/src/dcn_v2_cuda.h:19-29, dcn_v2_cuda.c:104-241) on NCHW fp32 device pointers.
//
// Same algorithm as the reference, batched over samples and in NHWC internally:
//   gcol = W^T dY                     (dcn_v2_cuda.c:175-178)    -> fp32 implicit GEMM (conv_simt.cu)
//   d offset, d mask                  (col2im_coord kernel, dcn_v2_im2col_cuda.cu:241-312)
//   d input  (atomic scatter)         (col2im kernel, :182-239)
//   col = im2col(x) ; dW += dY col^T  (:204-220) ; db += dY 1   (:225-230)
// fp32 throughout; d input / dW / db use float atomics, i.e. the summation order is not fixed -- as in
// the reference ("Backward is not reentrant", model/DCNv2/README.md:45).  All grads are OVERWRITTEN
// (the reference accumulates into buffers its Python wrapper zero-fills first, dcn_v2_func.py:44-48).
#include <algorithm>
#include <cstdint>

#include <cuda_bf16.h>

#include "common.cuh"

namespace m3d {

struct BwdGeom {
  int B, C, H, W, Cout, Ho, Wo, kh, kw, stride, pad, dil, KK, K;
};

struct Sample {
  bool valid;
  int h_low, w_low;
  float lh, lw;
  bool ok[4];  // corner (low,low), (low,high), (high,low), (high,high) inside the image
};

__device__ __forceinline__ Sample make_sample(const BwdGeom& g, int p, int q, int tap, float off_h, float off_w) {
  Sample s;
  const int i = tap / g.kw, j = tap % g.kw;
  const float h = static_cast<float>(p * g.stride - g.pad + i * g.dil) + off_h;
  const float w = static_cast<float>(q * g.stride - g.pad + j * g.dil) + off_w;
  s.valid = h > -1.f && w > -1.f && h < static_cast<float>(g.H) && w < static_cast<float>(g.W);
  const float hl = floorf(h), wl = floorf(w);
  s.h_low = static_cast<int>(hl), s.w_low = static_cast<int>(wl);
  s.lh = h - hl, s.lw = w - wl;
  const bool hl_ok = s.h_low >= 0, wl_ok = s.w_low >= 0, hh_ok = s.h_low + 1 <= g.H - 1, wh_ok = s.w_low + 1 <= g.W - 1;
  s.ok[0] = hl_ok && wl_ok, s.ok[1] = hl_ok && wh_ok, s.ok[2] = hh_ok && wl_ok, s.ok[3] = hh_ok && wh_ok;
  return s;
}

// One warp per (pixel, tap); lanes stride over channels.
//   d_off_h = sum_c cw_h(c) * gcol * m ;  d_off_w likewise ;  d_mask = sum_c gcol * bilinear(c)
__global__ void dcn_bwd_coord_kernel(const BwdGeom g, const float* __restrict__ x /*NHWC*/,
                                     const float* __restrict__ om /*NHWC [.,3KK]*/,
                                     const float* __restrict__ gcol /*[pix][tap*C + c]*/,
                                     float* __restrict__ grad_offset /*NCHW*/, float* __restrict__ grad_mask /*NCHW*/) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long total = static_cast<long>(g.B) * g.Ho * g.Wo * g.KK;
  if (wid >= total) return;
  const int tap = static_cast<int>(wid % g.KK);
  const long pix = wid / g.KK;
  const int q = static_cast<int>(pix % g.Wo), p = static_cast<int>((pix / g.Wo) % g.Ho), n = static_cast<int>(pix / (static_cast<long>(g.Wo) * g.Ho));
  const float* omp = om + pix * 3 * g.KK;
  const float m = omp[2 * g.KK + tap];
  const Sample s = make_sample(g, p, q, tap, omp[2 * tap], omp[2 * tap + 1]);
  float dh = 0.f, dw = 0.f, dm = 0.f;
  if (s.valid) {
    const float hh = 1.f - s.lh, hw = 1.f - s.lw;
    const float* x00 = x + ((static_cast<long>(n) * g.H + max(s.h_low, 0)) * g.W + max(s.w_low, 0)) * g.C;
    const float* x01 = x + ((static_cast<long>(n) * g.H + max(s.h_low, 0)) * g.W + min(s.w_low + 1, g.W - 1)) * g.C;
    const float* x10 = x + ((static_cast<long>(n) * g.H + min(s.h_low + 1, g.H - 1)) * g.W + max(s.w_low, 0)) * g.C;
    const float* x11 = x + ((static_cast<long>(n) * g.H + min(s.h_low + 1, g.H - 1)) * g.W + min(s.w_low + 1, g.W - 1)) * g.C;
    const float* gc = gcol + pix * g.K + static_cast<long>(tap) * g.C;
    for (int c = lane; c < g.C; c += 32) {
      const float v1 = s.ok[0] ? x00[c] : 0.f, v2 = s.ok[1] ? x01[c] : 0.f;
      const float v3 = s.ok[2] ? x10[c] : 0.f, v4 = s.ok[3] ? x11[c] : 0.f;
      const float cv = gc[c];
      const float val = hh * hw * v1 + hh * s.lw * v2 + s.lh * hw * v3 + s.lh * s.lw * v4;
      const float cwh = -hw * v1 - s.lw * v2 + hw * v3 + s.lw * v4;   // d val / d h
      const float cww = -hh * v1 + hh * v2 - s.lh * v3 + s.lh * v4;   // d val / d w
      dh += cwh * cv * m;
      dw += cww * cv * m;
      dm += cv * val;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dh += __shfl_xor_sync(0xffffffffu, dh, o);
    dw += __shfl_xor_sync(0xffffffffu, dw, o);
    dm += __shfl_xor_sync(0xffffffffu, dm, o);
  }
  if (lane == 0) {
    const long hw_o = static_cast<long>(g.Ho) * g.Wo;
    const long sp = static_cast<long>(p) * g.Wo + q;
    grad_offset[(static_cast<long>(n) * 2 * g.KK + 2 * tap) * hw_o + sp] = dh;
    grad_offset[(static_cast<long>(n) * 2 * g.KK + 2 * tap + 1) * hw_o + sp] = dw;
    grad_mask[(static_cast<long>(n) * g.KK + tap) * hw_o + sp] = dm;
  }
}

// d input (NHWC, zero-initialised): scatter w_corner * gcol * m to the four corners.
// Also rewrites gcol in place with the forward column value (sampled * m) for the dW GEMM.
// FIXED: the scatter accumulates in 64-bit FIXED POINT (2^-32 resolution): integer addition is associative, so the
// result does not depend on the order in which the atomics land -- a deterministic col2im (the reference's float
// atomicAdd, dcn_v2_im2col_cuda.cu:182-231, is not reproducible run to run).
constexpr float kFixScale = 4294967296.f;  // 2^32
template <bool FIXED>
__global__ void dcn_bwd_input_kernel(const BwdGeom g, const float* __restrict__ x, const float* __restrict__ om,
                                     float* __restrict__ gcol, float* __restrict__ grad_x,
                                     unsigned long long* __restrict__ grad_x_fix) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long total = static_cast<long>(g.B) * g.Ho * g.Wo * g.KK;
  if (wid >= total) return;
  const int tap = static_cast<int>(wid % g.KK);
  const long pix = wid / g.KK;
  const int q = static_cast<int>(pix % g.Wo), p = static_cast<int>((pix / g.Wo) % g.Ho), n = static_cast<int>(pix / (static_cast<long>(g.Wo) * g.Ho));
  const float* omp = om + pix * 3 * g.KK;
  const float m = omp[2 * g.KK + tap];
  const Sample s = make_sample(g, p, q, tap, omp[2 * tap], omp[2 * tap + 1]);
  float* gc = gcol + pix * g.K + static_cast<long>(tap) * g.C;
  if (!s.valid) {
    for (int c = lane; c < g.C; c += 32) gc[c] = 0.f;
    return;
  }
  const float hh = 1.f - s.lh, hw = 1.f - s.lw;
  const float wgt[4] = {hh * hw, hh * s.lw, s.lh * hw, s.lh * s.lw};
  long off[4];
  off[0] = ((static_cast<long>(n) * g.H + max(s.h_low, 0)) * g.W + max(s.w_low, 0)) * g.C;
  off[1] = ((static_cast<long>(n) * g.H + max(s.h_low, 0)) * g.W + min(s.w_low + 1, g.W - 1)) * g.C;
  off[2] = ((static_cast<long>(n) * g.H + min(s.h_low + 1, g.H - 1)) * g.W + max(s.w_low, 0)) * g.C;
  off[3] = ((static_cast<long>(n) * g.H + min(s.h_low + 1, g.H - 1)) * g.W + min(s.w_low + 1, g.W - 1)) * g.C;
  for (int c = lane; c < g.C; c += 32) {
    const float top = gc[c] * m;
    float val = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (s.ok[k]) {
        if (wgt[k] != 0.f) {
          if constexpr (FIXED)
            atomicAdd(grad_x_fix + off[k] + c, static_cast<unsigned long long>(__float2ll_rn(wgt[k] * top * kFixScale)));
          else
            atomicAdd(grad_x + off[k] + c, wgt[k] * top);
        }
        val += wgt[k] * x[off[k] + c];
      }
    }
    gc[c] = val * m;
  }
}

// dW[co][k] += sum_pix gy[pix][co] * col[pix][k];  db[co] += sum_pix gy[pix][co].
// 64 x 64 tile of dW per block, 4 x 4 per thread, the pixel range split over blockIdx.z (atomics).
__global__ void __launch_bounds__(256) dcn_bwd_weight_kernel(const float* __restrict__ gy, int gy_ld,
                                                             const float* __restrict__ col, int K, int Cout,
                                                             long npix, long pix_per_block,
                                                             float* __restrict__ grad_w /*[Cout][K] tap-major*/,
                                                             float* __restrict__ grad_b) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int k0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const long p_begin = blockIdx.z * pix_per_block;
  const long p_end = min(npix, p_begin + pix_per_block);
  float acc[4][4] = {};
  const int lp = tid / 16, lc = (tid % 16) * 4;  // loader: pixel lp of the 16, 4 consecutive columns lc
  for (long pp = p_begin; pp < p_end; pp += 16) {
    const long px = pp + lp;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px < p_end) {
      const float* ga = gy + px * gy_ld + c0 + lc;
      a.x = (c0 + lc + 0 < Cout) ? ga[0] : 0.f, a.y = (c0 + lc + 1 < Cout) ? ga[1] : 0.f;
      a.z = (c0 + lc + 2 < Cout) ? ga[2] : 0.f, a.w = (c0 + lc + 3 < Cout) ? ga[3] : 0.f;
      const float* cb = col + px * K + k0 + lc;
      b.x = (k0 + lc + 0 < K) ? cb[0] : 0.f, b.y = (k0 + lc + 1 < K) ? cb[1] : 0.f;
      b.z = (k0 + lc + 2 < K) ? cb[2] : 0.f, b.w = (k0 + lc + 3 < K) ? cb[3] : 0.f;
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lp][lc]) = a;
    *reinterpret_cast<float4*>(&Bs[lp][lc]) = b;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = c0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(grad_w + static_cast<long>(co) * K + k, acc[i][j]);
    }
  }
}

__global__ void dcn_bwd_bias_kernel(const float* __restrict__ gy, int gy_ld, int Cout, long npix,
                                    float* __restrict__ grad_b) {
  // block = 256 threads: thread -> channel (strided), block -> pixel range
  const long per = (npix + gridDim.x - 1) / gridDim.x;
  const long b = blockIdx.x * per, e = min(npix, b + per);
  for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
    float s = 0.f;
    for (long p = b; p < e; ++p) s += gy[p * gy_ld + c];
    atomicAdd(grad_b + c, s);
  }
}

// tap-major packed [Cout][kh*kw][C]  <->  reference layout [Cout][C][kh][kw]
__global__ void repack_weight_grad_kernel(const float* __restrict__ packed, float* __restrict__ out, int Cout, int C,
                                          int KK) {
  const long total = static_cast<long>(Cout) * C * KK;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % KK);
    const int c = static_cast<int>((i / KK) % C);
    const int co = static_cast<int>(i / (static_cast<long>(KK) * C));
    out[i] = packed[(static_cast<long>(co) * KK + t) * C + c];
  }
}

// W [Cout][C][KK] -> W^T packed for the gcol GEMM: rows k = tap*C + c, columns co (padded to CoutP)
__global__ void pack_wt_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int C, int KK, int CoutP) {
  const long total = static_cast<long>(KK) * C * CoutP;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % CoutP);
    const long k = i / CoutP;
    const int c = static_cast<int>(k % C), t = static_cast<int>(k / C);
    wt[i] = co < Cout ? w[(static_cast<long>(co) * C + c) * KK + t] : 0.f;
  }
}

// the same matrix as three bf16 parts (hi + mid + lo = 24 mantissa bits) for the tensor-core (bf16x3) GEMM
__global__ void pack_wt_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo, int Cout, int C, int KK,
                                     int CoutP) {
  const long total = static_cast<long>(KK) * C * CoutP;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % CoutP);
    const long k = i / CoutP;
    const int c = static_cast<int>(k % C), t = static_cast<int>(k / C);
    const float v = co < Cout ? w[(static_cast<long>(co) * C + c) * KK + t] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    hi[i] = h, mid[i] = m, lo[i] = __float2bfloat16_rn(r1 - __bfloat162float(m));
  }
}

__global__ void fixed_to_f32_kernel(const unsigned long long* __restrict__ in, float* __restrict__ out, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = static_cast<float>(static_cast<double>(static_cast<long long>(in[i])) * (1.0 / 4294967296.0));
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long n) {
  for (long i = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) * 4; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x * 4) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(in + i);
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a), o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(out + i) = o;
    } else {
      for (long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
    }
  }
}

static inline size_t al(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct BwdLayout {
  int Ho, Wo, KK, K, CoutP;
  size_t x, om, gy, gcol, wt, gx, gw, colb, gyb, wg, gx64, total;  // colb / gyb / wg / gx64: M3D_BF16 mode (bf16 copies,
                                                                   // wgrad workspace, fixed-point input-gradient accumulator)
};

static BwdLayout bwd_layout(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil) {
  BwdLayout L;
  L.Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
  L.Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
  L.KK = kh * kw, L.K = C * L.KK;
  L.CoutP = (Cout + 63) / 64 * 64;  // k-block of the tensor-core (bf16x3) W^T dY GEMM
  const size_t npo = static_cast<size_t>(B) * L.Ho * L.Wo;
  L.x = al(static_cast<size_t>(B) * H * W * C * 4);
  L.om = al(npo * 3 * L.KK * 4);
  L.gy = al(npo * L.CoutP * 4);
  L.gcol = al(npo * L.K * 4);
  L.wt = al(static_cast<size_t>(L.K) * L.CoutP * 6);  // fp32, or three bf16 parts
  L.gx = L.x;
  L.gw = al(static_cast<size_t>(Cout) * L.K * 4);
  L.colb = al(npo * L.K * 2);
  L.gyb = al(npo * L.CoutP * 2);
  L.wg = al(m3d_conv2d_wgrad_workspace(B, L.Ho, L.Wo, L.K, Cout, 1, 1)) + al(m3d_channel_sum_workspace(Cout));
  L.gx64 = al(static_cast<size_t>(B) * H * W * C * 8);
  L.total = L.x + L.om + L.gy + L.gcol + L.wt + L.gx + L.gw + L.colb + L.gyb + L.wg + L.gx64;
  return L;
}

}  // namespace m3d

using namespace m3d;

extern "C" int m3d_nchw_to_nhwc(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                                int out_cstride, int out_coff, m3d_stream_t stream);
extern "C" int m3d_nhwc_to_nchw(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                                int in_cstride, int in_coff, m3d_stream_t stream);

// deformable_group > 1: the operator is a sum over channel groups (see dcn_api.cu), so its backward is the single-group
// backward of every group on slices: inputs sliced in, input / offset / mask / weight gradients scattered back;
// grad_bias does not depend on the group.
struct BwdGroupSlices {
  size_t x, off, msk, w, gx, goff, gmsk, gw, total;
};
static BwdGroupSlices bwd_group_slices(int B, int Cg, int H, int W, int Cout, int KK, int Ho, int Wo) {
  BwdGroupSlices g;
  g.x = g.gx = al(static_cast<size_t>(B) * Cg * H * W * 4);
  g.off = g.goff = al(static_cast<size_t>(B) * 2 * KK * Ho * Wo * 4);
  g.msk = g.gmsk = al(static_cast<size_t>(B) * KK * Ho * Wo * 4);
  g.w = g.gw = al(static_cast<size_t>(Cout) * Cg * KK * 4);
  g.total = 2 * (g.x + g.off + g.msk + g.w);
  return g;
}

extern "C" size_t m3d_dcn_v2_backward_workspace(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad,
                                                int dil, int deformable_group) {
  const int dg = deformable_group < 1 ? 1 : deformable_group;
  if (dg == 1) return bwd_layout(B, C, H, W, Cout, kh, kw, stride, pad, dil).total;
  const BwdLayout L = bwd_layout(B, C / dg, H, W, Cout, kh, kw, stride, pad, dil);
  return L.total + bwd_group_slices(B, C / dg, H, W, Cout, kh * kw, L.Ho, L.Wo).total;
}

static int dcn_backward_one_group(const float* input, const float* weight, const float* offset, const float* mask,
                                  const float* grad_output, float* grad_input, float* grad_weight, float* grad_bias,
                                  float* grad_offset, float* grad_mask, int B, int C, int H, int W, int Cout, int kh,
                                  int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                  int precision, void* workspace, size_t workspace_bytes, m3d_stream_t stream_);

extern "C" int m3d_dcn_v2_backward(const float* input, const float* weight, const float* offset, const float* mask,
                                   const float* grad_output, float* grad_input, float* grad_weight, float* grad_bias,
                                   float* grad_offset, float* grad_mask, int B, int C, int H, int W, int Cout, int kh,
                                   int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                   int deformable_group, int precision, void* workspace, size_t workspace_bytes,
                                   m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3D_REQUIRE(precision == M3D_F32 || precision == M3D_BF16X3 || precision == M3D_BF16,
              "backward precision must be M3D_F32, M3D_BF16X3 or M3D_BF16");
  const int dg = deformable_group;
  if (dg == 1)
    return dcn_backward_one_group(input, weight, offset, mask, grad_output, grad_input, grad_weight, grad_bias,
                                  grad_offset, grad_mask, B, C, H, W, Cout, kh, kw, stride_h, stride_w, pad_h, pad_w,
                                  dil_h, dil_w, precision, workspace, workspace_bytes, stream_);
  M3D_REQUIRE(input && weight && offset && mask && grad_output && grad_input && grad_weight && grad_bias && grad_offset &&
                  grad_mask,
              "NULL tensor pointer");
  M3D_REQUIRE(dg >= 1 && C % dg == 0, "channels (%d) must be a multiple of deformable_group (%d)", C, dg);
  const int Cg = C / dg, KK = kh * kw;
  const BwdLayout L = bwd_layout(B, Cg, H, W, Cout, kh, kw, stride_h, pad_h, dil_h);
  M3D_REQUIRE(L.Ho >= 1 && L.Wo >= 1, "empty output");
  const BwdGroupSlices G = bwd_group_slices(B, Cg, H, W, Cout, KK, L.Ho, L.Wo);
  if (workspace == nullptr || workspace_bytes < L.total + G.total) {
    set_last_error("DCNv2 backward workspace too small: %zu < %zu", workspace_bytes, L.total + G.total);
    return M3D_ERR_WORKSPACE;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  size_t o = 0;
  auto take = [&](size_t n) { float* q = reinterpret_cast<float*>(ws + o); o += n; return q; };
  float *xg = take(G.x), *og = take(G.off), *mg = take(G.msk), *wg = take(G.w);
  float *gxg = take(G.gx), *gog = take(G.goff), *gmg = take(G.gmsk), *gwg = take(G.gw);
  void* inner = ws + G.total;
  const size_t hw = static_cast<size_t>(H) * W * 4, howo = static_cast<size_t>(L.Ho) * L.Wo * 4;
  const size_t wrow = static_cast<size_t>(Cg) * KK * 4;
  for (int g = 0; g < dg; ++g) {
    const size_t xo = static_cast<size_t>(g) * Cg * H * W, oo = static_cast<size_t>(g) * 2 * KK * L.Ho * L.Wo;
    const size_t mo = static_cast<size_t>(g) * KK * L.Ho * L.Wo, wo = static_cast<size_t>(g) * Cg * KK;
    M3D_CUDA_OK(cudaMemcpy2DAsync(xg, Cg * hw, input + xo, C * hw, Cg * hw, B, cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(og, 2 * KK * howo, offset + oo, dg * 2 * KK * howo, 2 * KK * howo, B,
                                  cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(mg, KK * howo, mask + mo, dg * KK * howo, KK * howo, B, cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(wg, wrow, weight + wo, static_cast<size_t>(C) * KK * 4, wrow, Cout,
                                  cudaMemcpyDeviceToDevice, stream));
    const int rc = dcn_backward_one_group(xg, wg, og, mg, grad_output, gxg, gwg, grad_bias, gog, gmg, B, Cg, H, W, Cout,
                                          kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, precision, inner,
                                          workspace_bytes - G.total, stream_);
    if (rc) return rc;
    M3D_CUDA_OK(cudaMemcpy2DAsync(grad_input + xo, C * hw, gxg, Cg * hw, Cg * hw, B, cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(grad_offset + oo, dg * 2 * KK * howo, gog, 2 * KK * howo, 2 * KK * howo, B,
                                  cudaMemcpyDeviceToDevice, stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(grad_mask + mo, dg * KK * howo, gmg, KK * howo, KK * howo, B, cudaMemcpyDeviceToDevice,
                                  stream));
    M3D_CUDA_OK(cudaMemcpy2DAsync(grad_weight + wo, static_cast<size_t>(C) * KK * 4, gwg, wrow, wrow, Cout,
                                  cudaMemcpyDeviceToDevice, stream));
  }
  return M3D_OK;
}

static int dcn_backward_one_group(const float* input, const float* weight, const float* offset, const float* mask,
                                  const float* grad_output, float* grad_input, float* grad_weight, float* grad_bias,
                                  float* grad_offset, float* grad_mask, int B, int C, int H, int W, int Cout, int kh,
                                  int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                  int precision, void* workspace, size_t workspace_bytes, m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int deformable_group = 1;
  M3D_REQUIRE(input && weight && offset && mask && grad_output && grad_input && grad_weight && grad_bias && grad_offset &&
                  grad_mask,
              "NULL tensor pointer");
  if (deformable_group != 1 || stride_h != stride_w || pad_h != pad_w || dil_h != dil_w) {
    set_last_error("DCNv2 backward configuration not implemented (dg=%d)", deformable_group);
    return M3D_ERR_UNSUPPORTED;
  }
  const BwdLayout L = bwd_layout(B, C, H, W, Cout, kh, kw, stride_h, pad_h, dil_h);
  M3D_REQUIRE(L.Ho >= 1 && L.Wo >= 1, "empty output");
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_last_error("DCNv2 backward workspace too small: %zu < %zu", workspace_bytes, L.total);
    return M3D_ERR_WORKSPACE;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws);
  float* om = reinterpret_cast<float*>(ws + L.x);
  float* gy = reinterpret_cast<float*>(ws + L.x + L.om);
  float* gcol = reinterpret_cast<float*>(ws + L.x + L.om + L.gy);
  float* wt = reinterpret_cast<float*>(ws + L.x + L.om + L.gy + L.gcol);
  float* gx = reinterpret_cast<float*>(ws + L.x + L.om + L.gy + L.gcol + L.wt);
  float* gw = reinterpret_cast<float*>(ws + L.x + L.om + L.gy + L.gcol + L.wt + L.gx);
  __nv_bfloat16* colb = reinterpret_cast<__nv_bfloat16*>(ws + L.x + L.om + L.gy + L.gcol + L.wt + L.gx + L.gw);
  __nv_bfloat16* gyb = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(colb) + L.colb);
  uint8_t* wg_ws = reinterpret_cast<uint8_t*>(gyb) + L.gyb;
  unsigned long long* gx64 = reinterpret_cast<unsigned long long*>(wg_ws + L.wg);
  const long npo = static_cast<long>(B) * L.Ho * L.Wo;

  int rc = m3d_nchw_to_nhwc(input, M3D_F32, x, M3D_F32, B, C, H, W, C, 0, stream_);
  if (rc) return rc;
  rc = m3d_nchw_to_nhwc(offset, M3D_F32, om, M3D_F32, B, 2 * L.KK, L.Ho, L.Wo, 3 * L.KK, 0, stream_);
  if (rc) return rc;
  rc = m3d_nchw_to_nhwc(mask, M3D_F32, om, M3D_F32, B, L.KK, L.Ho, L.Wo, 3 * L.KK, 2 * L.KK, stream_);
  if (rc) return rc;
  if (L.CoutP != Cout) M3D_CUDA_OK(cudaMemsetAsync(gy, 0, L.gy, stream));
  rc = m3d_nchw_to_nhwc(grad_output, M3D_F32, gy, M3D_F32, B, Cout, L.Ho, L.Wo, L.CoutP, 0, stream_);
  if (rc) return rc;
  M3D_CUDA_OK(cudaMemsetAsync(gx, 0, L.gx, stream));
  M3D_CUDA_OK(cudaMemsetAsync(gw, 0, L.gw, stream));
  M3D_CUDA_OK(cudaMemsetAsync(grad_bias, 0, sizeof(float) * Cout, stream));
  {
    const long total = static_cast<long>(L.K) * L.CoutP;
    const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 4096));
    if (precision == M3D_F32) {
      pack_wt_kernel<<<grid, 256, 0, stream>>>(weight, wt, Cout, C, L.KK, L.CoutP);
    } else {  // (M3D_BF16 uses the high part only)
      __nv_bfloat16* w3 = reinterpret_cast<__nv_bfloat16*>(wt);
      pack_wt_split_kernel<<<grid, 256, 0, stream>>>(weight, w3, w3 + total, w3 + 2 * total, Cout, C, L.KK, L.CoutP);
    }
    M3D_CUDA_OK(cudaGetLastError());
  }
  // gcol[pix][k] = sum_co Wt[k][co] * gy[pix][co]  : a 1x1 "convolution" with K output channels
  m3d_conv_desc d = {};
  d.act_dtype = M3D_F32, d.out_dtype = M3D_F32;
  d.num_inputs = 1;
  d.in[0] = gy, d.in_c[0] = L.CoutP, d.in_cstride[0] = L.CoutP;
  d.N = B, d.H = L.Ho, d.W = L.Wo;
  d.R = 1, d.S = 1, d.stride = 1, d.pad = 0, d.dil = 1;
  d.Cout = L.K, d.groups = 1;
  if (precision == M3D_BF16) {  // bf16 operands (the training throughput mode): gy rounded to bf16 once, used twice
    const long n = npo * L.CoutP;
    f32_to_bf16_kernel<<<static_cast<int>(std::min<long>((n / 4 + 255) / 256, 4096)), 256, 0, stream>>>(gy, gyb, n);
    M3D_CUDA_OK(cudaGetLastError());
    d.act_dtype = M3D_BF16;
    d.in[0] = gyb;
    d.weight = reinterpret_cast<const __nv_bfloat16*>(wt);
  } else if (precision == M3D_F32) {  // IEEE fp32 FMA on the CUDA cores (reference accuracy)
    d.weight_f32 = wt;
  } else {  // 3-part bf16 split on the tensor cores (~3e-6 relative)
    const long total = static_cast<long>(L.K) * L.CoutP;
    const __nv_bfloat16* w3 = reinterpret_cast<const __nv_bfloat16*>(wt);
    d.weight = w3, d.weight_mid = w3 + total, d.weight_lo = w3 + 2 * total;
  }
  d.weight_rows = L.K;
  d.out = gcol, d.out_cstride = L.K;
  d.slope = 1.0f;
  rc = m3d_conv2d_nhwc(&d, stream_);
  if (rc) return rc;

  BwdGeom g;
  g.B = B, g.C = C, g.H = H, g.W = W, g.Cout = Cout, g.Ho = L.Ho, g.Wo = L.Wo, g.kh = kh, g.kw = kw;
  g.stride = stride_h, g.pad = pad_h, g.dil = dil_h, g.KK = L.KK, g.K = L.K;
  const long warps = npo * L.KK;
  const int blocks = static_cast<int>((warps * 32 + 255) / 256);
  dcn_bwd_coord_kernel<<<blocks, 256, 0, stream>>>(g, x, om, gcol, grad_offset, grad_mask);
  M3D_CUDA_OK(cudaGetLastError());
  if (precision == M3D_BF16) {  // deterministic: fixed-point accumulation, then one conversion pass
    const long nx = static_cast<long>(B) * H * W * C;
    M3D_CUDA_OK(cudaMemsetAsync(gx64, 0, static_cast<size_t>(nx) * 8, stream));
    dcn_bwd_input_kernel<true><<<blocks, 256, 0, stream>>>(g, x, om, gcol, gx, gx64);  // gcol becomes col
    M3D_CUDA_OK(cudaGetLastError());
    fixed_to_f32_kernel<<<static_cast<int>(std::min<long>((nx + 255) / 256, 8192)), 256, 0, stream>>>(gx64, gx, nx);
  } else {
    dcn_bwd_input_kernel<false><<<blocks, 256, 0, stream>>>(g, x, om, gcol, gx, nullptr);  // gcol becomes col
  }
  M3D_CUDA_OK(cudaGetLastError());
  if (precision == M3D_BF16 && L.K % 8 == 0) {  // (K = C * kh * kw not a multiple of 8: TMA strides, SIMT path below)
    // dW = gy^T col on the tensor cores (wgrad.cu: the sampled columns are the "input" of a 1x1 convolution with K
    // channels), bias gradient by the deterministic channel sum -- no atomics in this mode except the input scatter
    const long n = npo * L.K;
    f32_to_bf16_kernel<<<static_cast<int>(std::min<long>((n / 4 + 255) / 256, 8192)), 256, 0, stream>>>(gcol, colb, n);
    M3D_CUDA_OK(cudaGetLastError());
    const size_t wsz = m3d_conv2d_wgrad_workspace(B, L.Ho, L.Wo, L.K, Cout, 1, 1);
    rc = m3d_conv2d_wgrad(colb, L.K, 0, gyb, L.CoutP, 0, gw, B, L.Ho, L.Wo, L.K, L.Ho, L.Wo, Cout, 1, 1, 1, 0, 1, wg_ws, wsz,
                          stream_);
    if (rc) return rc;
    rc = m3d_channel_sum(gyb, npo, Cout, L.CoutP, 0, grad_bias, wg_ws + al(wsz), m3d_channel_sum_workspace(Cout), stream_);
    if (rc) return rc;
    const long total = static_cast<long>(Cout) * L.K;
    repack_weight_grad_kernel<<<static_cast<int>(std::min<long>((total + 255) / 256, 4096)), 256, 0, stream>>>(
        gw, grad_weight, Cout, C, L.KK);
    M3D_CUDA_OK(cudaGetLastError());
  } else {
    const int split = static_cast<int>(std::max<long>(1, std::min<long>(64, npo / 2048)));
    const long per = ((npo + split - 1) / split + 15) / 16 * 16;
    dim3 grid((L.K + 63) / 64, (Cout + 63) / 64, split);
    dcn_bwd_weight_kernel<<<grid, 256, 0, stream>>>(gy, L.CoutP, gcol, L.K, Cout, npo, per, gw, grad_bias);
    M3D_CUDA_OK(cudaGetLastError());
    dcn_bwd_bias_kernel<<<static_cast<int>(std::max<long>(1, std::min<long>(256, npo / 256))), 256, 0, stream>>>(
        gy, L.CoutP, Cout, npo, grad_bias);
    M3D_CUDA_OK(cudaGetLastError());
    const long total = static_cast<long>(Cout) * L.K;
    repack_weight_grad_kernel<<<static_cast<int>(std::min<long>((total + 255) / 256, 4096)), 256, 0, stream>>>(
        gw, grad_weight, Cout, C, L.KK);
    M3D_CUDA_OK(cudaGetLastError());
  }
  return m3d_nhwc_to_nchw(gx, M3D_F32, grad_input, M3D_F32, B, C, H, W, C, 0, stream_);
}
