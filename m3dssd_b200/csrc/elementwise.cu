// Bandwidth-bound kernels around the implicit-GEMM layers (NHWC activations):
// stem 7x7 conv (C_in = 3), 2x2 max-pool, depthwise transposed-conv up-sample +
// skip add, class softmax / fg-prob / top-1 anchor, alignment offset builders,
// head flattening, and NCHW <-> NHWC conversion for the drop-in operators.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdint>

#include "common.cuh"

namespace m3d {

template <typename T>
__device__ __forceinline__ float to_f(T v) {
  return static_cast<float>(v);
}
template <typename T>
__device__ __forceinline__ T from_f(float v) {
  return static_cast<T>(v);
}

static inline int cdiv(long a, long b) { return static_cast<int>((a + b - 1) / b); }

// ---------------------------------------------------------------------------
// Stem: 7x7 / pad 3 / stride 1 convolution of the NCHW fp32 image (3 channels)
// with BN folded into (w, bias) and LeakyReLU, written as NHWC with 16 real
// channels.  model/pose_dla_dcn.py:336-340 (DLA.base_layer).
// Each thread produces 2 vertically adjacent pixels x 16 channels from a
// shared-memory input tile; weights are broadcast from shared memory.
// ---------------------------------------------------------------------------
constexpr int STEM_TW = 32, STEM_TH = 16, STEM_CO = 16;

template <typename OutT>
__global__ void __launch_bounds__(256) stem_conv7x7_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                           const float* __restrict__ bias, OutT* __restrict__ out,
                                                           int out_cstride, int N, int H, int W, float slope) {
  __shared__ float s_in[3][STEM_TH + 6][STEM_TW + 6 + 2];
  __shared__ __align__(16) float s_w[147][STEM_CO];  // [c*49 + r*7 + s][co]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads; thread -> rows ty and ty + 8
  const int w0 = blockIdx.x * STEM_TW, h0 = blockIdx.y * STEM_TH, n = blockIdx.z;
  for (int i = threadIdx.x; i < 147 * STEM_CO; i += 256) {
    const int co = i % STEM_CO, k = i / STEM_CO;
    s_w[k][co] = w[co * 147 + k];
  }
  for (int i = threadIdx.x; i < 3 * (STEM_TH + 6) * (STEM_TW + 6); i += 256) {
    const int x = i % (STEM_TW + 6);
    const int y = (i / (STEM_TW + 6)) % (STEM_TH + 6);
    const int c = i / ((STEM_TW + 6) * (STEM_TH + 6));
    const int gy = h0 + y - 3, gx = w0 + x - 3;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(img + ((static_cast<long>(n) * 3 + c) * H + gy) * W + gx);
    s_in[c][y][x] = v;
  }
  __syncthreads();
  float acc0[STEM_CO], acc1[STEM_CO];
#pragma unroll
  for (int co = 0; co < STEM_CO; ++co) acc0[co] = acc1[co] = __ldg(bias + co);
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int r = 0; r < 7; ++r) {
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const float a0 = s_in[c][ty + r][tx + s];
        const float a1 = s_in[c][ty + 8 + r][tx + s];
        const float4* wp = reinterpret_cast<const float4*>(&s_w[c * 49 + r * 7 + s][0]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wp[q];
          acc0[4 * q + 0] = fmaf(a0, wv.x, acc0[4 * q + 0]);
          acc0[4 * q + 1] = fmaf(a0, wv.y, acc0[4 * q + 1]);
          acc0[4 * q + 2] = fmaf(a0, wv.z, acc0[4 * q + 2]);
          acc0[4 * q + 3] = fmaf(a0, wv.w, acc0[4 * q + 3]);
          acc1[4 * q + 0] = fmaf(a1, wv.x, acc1[4 * q + 0]);
          acc1[4 * q + 1] = fmaf(a1, wv.y, acc1[4 * q + 1]);
          acc1[4 * q + 2] = fmaf(a1, wv.z, acc1[4 * q + 2]);
          acc1[4 * q + 3] = fmaf(a1, wv.w, acc1[4 * q + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int gy = h0 + ty + 8 * half, gx = w0 + tx;
    if (gy < H && gx < W) {
      float* acc = half ? acc1 : acc0;
      OutT* o = out + ((static_cast<long>(n) * H + gy) * W + gx) * out_cstride;
      if constexpr (sizeof(OutT) == 2) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = acc[2 * i], b = acc[2 * i + 1];
          a = a > 0.f ? a : a * slope;
          b = b > 0.f ? b : b * slope;
          __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
          pk[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 v;
          v.x = acc[4 * i], v.y = acc[4 * i + 1], v.z = acc[4 * i + 2], v.w = acc[4 * i + 3];
          v.x = v.x > 0.f ? v.x : v.x * slope;
          v.y = v.y > 0.f ? v.y : v.y * slope;
          v.z = v.z > 0.f ? v.z : v.z * slope;
          v.w = v.w > 0.f ? v.w : v.w * slope;
          reinterpret_cast<float4*>(o)[i] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// 2x2 / stride 2 max-pool (nn.MaxPool2d(2), model/pose_dla_dcn.py:306), NHWC.
// One thread per 8 channels (bf16) / 4 channels (fp32): 16-byte vectors.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2x2_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C,
                                  int in_cstride, int out_cstride) {
  grid_dep_sync();  // PDL: launched while the previous kernel drains
  constexpr int V = 16 / sizeof(T);
  const int Ho = H / 2, Wo = W / 2, cv = C / V;
  const long total = static_cast<long>(N) * Ho * Wo * cv;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * V;
    long r = i / cv;
    const int x = static_cast<int>(r % Wo);
    r /= Wo;
    const int y = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    const T* p00 = in + ((static_cast<long>(n) * H + 2 * y) * W + 2 * x) * in_cstride + c;
    const T* p10 = p00 + static_cast<long>(W) * in_cstride;
    T a[V], b[V], cc[V], d[V], o[V];
    *reinterpret_cast<uint4*>(a) = __ldg(reinterpret_cast<const uint4*>(p00));
    *reinterpret_cast<uint4*>(b) = __ldg(reinterpret_cast<const uint4*>(p00 + in_cstride));
    *reinterpret_cast<uint4*>(cc) = __ldg(reinterpret_cast<const uint4*>(p10));
    *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(p10 + in_cstride));
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = from_f<T>(fmaxf(fmaxf(to_f(a[k]), to_f(b[k])), fmaxf(to_f(cc[k]), to_f(d[k]))));
    *reinterpret_cast<uint4*>(out + ((static_cast<long>(n) * Ho + y) * Wo + x) * out_cstride + c) =
        *reinterpret_cast<uint4*>(o);
  }
}

// ---------------------------------------------------------------------------
// IDAUp: depthwise ConvTranspose2d(k = 2f, stride f, pad f/2, groups = C, no
// bias) followed by "+ skip" (model/pose_dla_dcn.py:536-552); weights are passed tap-major,
// [k*k][C], so a thread's 8 channels of one tap are two 16-byte loads.  Output pixel
// (oy, ox) gathers the <= ceil(k/f)^2 = 4 input pixels that reach it:
//   oy = iy * f - pad + ky   <=>   iy = (oy + pad - ky) / f   when divisible.
// ---------------------------------------------------------------------------
// 16-byte vector load / store of V = 16 / sizeof(T) elements as floats, written so that the arrays stay in registers
// (type-punning `*reinterpret_cast<uint4*>(array)` made ptxas keep them in local memory: ncu showed local loads / stores
// and 60 % of the stall cycles on them).
template <typename T>
__device__ __forceinline__ void load_vec_f(const T* __restrict__ p, float (&v)[16 / sizeof(T)]);
template <>
__device__ __forceinline__ void load_vec_f<__nv_bfloat16>(const __nv_bfloat16* __restrict__ p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  v[0] = __uint_as_float(u.x << 16), v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16), v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16), v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16), v[7] = __uint_as_float(u.w & 0xffff0000u);
}
template <>
__device__ __forceinline__ void load_vec_f<float>(const float* __restrict__ p, float (&v)[4]) {
  const float4 u = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = u.x, v[1] = u.y, v[2] = u.z, v[3] = u.w;
}
__device__ __forceinline__ void store_vec_f(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
    o[q] = static_cast<uint32_t>(__bfloat16_as_ushort(h.x)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h.y)) << 16);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void store_vec_f(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

template <typename T>
__global__ void __launch_bounds__(256) upsample_add_kernel(const T* __restrict__ x, const float* __restrict__ wt,
                                                           const T* __restrict__ skip, T* __restrict__ out, int N, int H,
                                                           int W, int C, int f, int x_cstride, int skip_cstride,
                                                           int out_cstride) {
  grid_dep_sync();  // PDL: launched while the previous kernel drains
  // grid = (segments of an output row, row parity class x slice, image): 32-bit index math only.  A thread owns one
  // (output column, 8 channels) and walks the output rows oy = m * f + parity of its slice: the four taps that reach
  // those pixels are the same for all of them, so their weights are loaded ONCE into registers (they were 8 of the 13
  // 16-byte loads per output: the kernel was bound by L1 wavefronts, 2.0 TB/s of algorithmic traffic).
  constexpr int V = 16 / sizeof(T);
  const int k = 2 * f, pad = f / 2;
  const int Ho = H * f, Wo = W * f, cv = C / V;
  const int n = blockIdx.z;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Wo * cv) return;
  const int ox = idx / cv;
  const int c = (idx - ox * cv) * V;
  const int kx0 = (ox + pad) % f;
  const int ix0 = (ox + pad - kx0) / f;
  const int parity = blockIdx.y % f, slice = blockIdx.y / f, nslices = gridDim.y / f;
  const int ky0 = (parity + pad) % f;
  const int diy = (parity + pad - ky0) / f;  // iy0 = m + diy for output row m * f + parity
  float wv[2][2][V];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ky = ky0 + a * f, kx = kx0 + b * f;
#pragma unroll
      for (int q = 0; q < V; q += 4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wt + (ky * k + kx) * C + c + q));
        wv[a][b][q] = w4.x, wv[a][b][q + 1] = w4.y, wv[a][b][q + 2] = w4.z, wv[a][b][q + 3] = w4.w;
      }
    }
  }
  // every load of an output is issued before the first use (addresses clamped into the tensor; taps outside the input
  // are dropped by predicate afterwards, as in the reference): one memory latency per output, not five
  const bool x_ok[2] = {ix0 >= 0 && ix0 < W, ix0 - 1 >= 0 && ix0 - 1 < W};
  const int ixc[2] = {min(max(ix0, 0), W - 1), min(max(ix0 - 1, 0), W - 1)};
#pragma unroll 2
  for (int m = slice; m < H; m += nslices) {
    const int oy = m * f + parity;
    const long opix = (static_cast<long>(n) * Ho + oy) * Wo + ox;
    const int iy0 = m + diy;
    float v[2][2][V];
#pragma unroll
    for (int a = 0; a < 2; ++a) {  // k = 2f: exactly two taps per axis reach an output pixel
      const int iyc = min(max(iy0 - a, 0), H - 1);
#pragma unroll
      for (int b = 0; b < 2; ++b) load_vec_f<T>(x + ((static_cast<long>(n) * H + iyc) * W + ixc[b]) * x_cstride + c, v[a][b]);
    }
    float acc[V];
    if (skip != nullptr) {
      load_vec_f<T>(skip + opix * skip_cstride + c, acc);
    } else {
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] = 0.f;
    }
    float up[V];
#pragma unroll
    for (int q = 0; q < V; ++q) up[q] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const bool y_ok = iy0 - a >= 0 && iy0 - a < H;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const bool ok = y_ok && x_ok[b];
#pragma unroll
        for (int q = 0; q < V; ++q) up[q] = ok ? fmaf(v[a][b][q], wv[a][b][q], up[q]) : up[q];
      }
    }
    float o[V];
#pragma unroll
    for (int q = 0; q < V; ++q) o[q] = up[q] + acc[q];
    store_vec_f(out + opix * out_cstride + c, o);
  }
}

// ---------------------------------------------------------------------------
// Class softmax over the K classes of every (anchor, pixel), fg probability
// 1 - p_bg, its top-1 anchor per pixel, detection score / class, and the
// reference's flattened output layout (model/M3d_inference_align.py:229-234,
// 300-301; lib/rpn_util.py:892-901, 1510-1511):
//   logits NHWC fp32 [N,H,W,K*A], channel = class * A + anchor
//   cls, prob  [N, (a*H + h)*W + w, K]
// One block per (n, h, 32-pixel segment); tile staged in shared memory so both
// the NHWC reads and the anchor-major writes are coalesced.
// ---------------------------------------------------------------------------
constexpr int SM_PIX = 32;

// K == 4 (the model's 3 classes + background): 16-byte loads / stores, everything in registers.  The tile is
// staged class-major in shared memory ([K*A][SM_PIX + 1]) so that the NHWC reads (144 contiguous floats per
// pixel) and the anchor-major writes (16 bytes per (anchor, pixel), pixels contiguous) are both coalesced and
// the column accesses are bank-conflict free.
__device__ __forceinline__ void shape_align_om_pixel(float fg, int a, const float* __restrict__ anchors, int anchor_ld,
                                                     float feat_stride, float thresh, float* __restrict__ om, long i);
__global__ void __launch_bounds__(256) cls_softmax4_kernel(const float* __restrict__ logits, int lc_stride, int N, int H,
                                                           int W, int A, float* __restrict__ cls_out,
                                                           float* __restrict__ prob_out, float* __restrict__ fg_max,
                                                           int* __restrict__ fg_arg, float* __restrict__ score,
                                                           unsigned char* __restrict__ cls_pred,
                                                           const float* __restrict__ anchors, int anchor_ld,
                                                           float feat_stride, float thresh, float* __restrict__ shape_om) {
  grid_dep_sync();
  extern __shared__ float s_tile[];  // [4*A][SM_PIX + 1], then fg [A][SM_PIX + 1]
  constexpr int K = 4, LD = SM_PIX + 1;
  const int KA = K * A;
  float* s_fg = s_tile + KA * LD;
  const int w0 = blockIdx.x * SM_PIX, h = blockIdx.y, n = blockIdx.z;
  const int npix = min(SM_PIX, W - w0);
  const float* src = logits + ((static_cast<long>(n) * H + h) * W + w0) * lc_stride;
  const int c4n = KA >> 2;  // float4 per pixel (host: KA % 4 == 0, lc_stride % 4 == 0)
  for (int px = threadIdx.x >> 5; px < npix; px += 8) {
    for (int c4 = threadIdx.x & 31; c4 < c4n; c4 += 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<long>(px) * lc_stride) + c4);
      float* d = s_tile + (c4 * 4) * LD + px;
      d[0] = v.x, d[LD] = v.y, d[2 * LD] = v.z, d[3 * LD] = v.w;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A * SM_PIX; i += blockDim.x) {
    const int px = i & (SM_PIX - 1), a = i / SM_PIX;
    if (px >= npix) continue;
    const float v0 = s_tile[a * LD + px], v1 = s_tile[(A + a) * LD + px];
    const float v2 = s_tile[(2 * A + a) * LD + px], v3 = s_tile[(3 * A + a) * LD + px];
    const float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
    const float e0 = expf(v0 - mx), e1 = expf(v1 - mx), e2 = expf(v2 - mx), e3 = expf(v3 - mx);
    const float sum = ((e0 + e1) + e2) + e3;
    const float p0 = e0 / sum, p1 = e1 / sum, p2 = e2 / sum, p3 = e3 / sum;
    const long row = (static_cast<long>(n) * A + a) * H * W + static_cast<long>(h) * W + (w0 + px);
    if (cls_out != nullptr) {  // (NULL: the detection stages only need score / class / fg below)
      *reinterpret_cast<float4*>(cls_out + row * 4) = make_float4(v0, v1, v2, v3);
      *reinterpret_cast<float4*>(prob_out + row * 4) = make_float4(p0, p1, p2, p3);
    }
    float best = p1;
    int bestk = 1;
    if (p2 > best) best = p2, bestk = 2;
    if (p3 > best) best = p3, bestk = 3;
    score[row] = best;
    cls_pred[row] = static_cast<unsigned char>(bestk);
    s_fg[a * LD + px] = 1.f - p0;
  }
  __syncthreads();
  if (threadIdx.x < npix) {
    const int px = threadIdx.x;
    float best = -1.f;
    int arg = 0;
    for (int a = 0; a < A; ++a) {
      const float f = s_fg[a * LD + px];
      if (f > best) {  // first maximum wins, like torch.max / topk(k=1)
        best = f;
        arg = a;
      }
    }
    const long pix = (static_cast<long>(n) * H + h) * W + w0 + px;
    fg_max[pix] = best;
    fg_arg[pix] = arg;
    if (shape_om != nullptr) shape_align_om_pixel(best, arg, anchors, anchor_ld, feat_stride, thresh, shape_om, pix);
  }
}


__global__ void __launch_bounds__(256) cls_softmax_kernel(const float* __restrict__ logits, int lc_stride, int N, int H,
                                                          int W, int A, int K, float* __restrict__ cls_out,
                                                          float* __restrict__ prob_out, float* __restrict__ fg_max,
                                                          int* __restrict__ fg_arg, float* __restrict__ score,
                                                          unsigned char* __restrict__ cls_pred) {
  grid_dep_sync();  // PDL: launched while the previous kernel drains
  extern __shared__ float s_tile[];  // [SM_PIX][K*A + 1]
  const int KA = K * A, ld = KA + 1;
  const int w0 = blockIdx.x * SM_PIX, h = blockIdx.y, n = blockIdx.z;
  const int npix = min(SM_PIX, W - w0);
  const float* src = logits + ((static_cast<long>(n) * H + h) * W + w0) * lc_stride;
  for (int i = threadIdx.x; i < npix * KA; i += blockDim.x) {
    const int px = i / KA, c = i - px * KA;
    s_tile[px * ld + c] = __ldg(src + static_cast<long>(px) * lc_stride + c);
  }
  __syncthreads();
  // softmax per (pixel, anchor): thread -> (a, px) with px fastest so writes are contiguous in w
  for (int i = threadIdx.x; i < A * SM_PIX; i += blockDim.x) {
    const int px = i % SM_PIX, a = i / SM_PIX;
    if (px >= npix) continue;
    float v[8];
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
      v[k] = s_tile[px * ld + k * A + a];
      mx = fmaxf(mx, v[k]);
    }
    float sum = 0.f;
    float e[8];
    for (int k = 0; k < K; ++k) {
      e[k] = expf(v[k] - mx);
      sum += e[k];
    }
    const long row = (static_cast<long>(n) * A + a) * H * W + static_cast<long>(h) * W + (w0 + px);
    float best = -1.f;
    int bestk = 1;
    for (int k = 0; k < K; ++k) {
      const float p = e[k] / sum;
      if (cls_out != nullptr) cls_out[row * K + k] = v[k], prob_out[row * K + k] = p;
      if (k >= 1 && p > best) {
        best = p;
        bestk = k;
      }
      if (k == 0) s_tile[px * ld + a] = 1.f - p;  // fg prob overwrites the class-0 logit slot
    }
    score[row] = best;
    cls_pred[row] = static_cast<unsigned char>(bestk);
  }
  __syncthreads();
  if (threadIdx.x < npix) {
    const int px = threadIdx.x;
    float best = -1.f;
    int arg = 0;
    for (int a = 0; a < A; ++a) {
      const float f = s_tile[px * ld + a];
      if (f > best) {  // first maximum wins, like torch.max / topk(k=1)
        best = f;
        arg = a;
      }
    }
    const long pix = (static_cast<long>(n) * H + h) * W + w0 + px;
    fg_max[pix] = best;
    fg_arg[pix] = arg;
  }
}

// ---------------------------------------------------------------------------
// shape_align offsets (model/module/feturealign_mgpu.py:119-136, 160-172):
// tap (i, j) of the 3x3 DCNv2 moves by ((ah/stride/3 - 1)(i - 1), (aw/stride/3 - 1)(j - 1))
// for the top-1 anchor, zeroed where fg <= thresh; modulation mask = fg.
// om layout: [N,H,W,27] = 18 offsets (dh, dw per tap) + 9 masks.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void shape_align_om_pixel(float fg, int a, const float* __restrict__ anchors, int anchor_ld,
                                                     float feat_stride, float thresh, float* __restrict__ om, long i);

__global__ void shape_align_om_kernel(const float* __restrict__ fg_max, const int* __restrict__ fg_arg,
                                      const float* __restrict__ anchors, int anchor_ld, float feat_stride, float thresh,
                                      float* __restrict__ om, long npix) {
  grid_dep_sync();  // PDL: launched while the previous kernel drains
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= npix) return;
  shape_align_om_pixel(fg_max[i], fg_arg[i], anchors, anchor_ld, feat_stride, thresh, om, i);
}

__device__ __forceinline__ void shape_align_om_pixel(float fg, int a, const float* __restrict__ anchors, int anchor_ld,
                                                     float feat_stride, float thresh, float* __restrict__ om, long i) {
  const float hard = fg > thresh ? 1.f : 0.f;
  const float aw = anchors[a * anchor_ld + 2] - anchors[a * anchor_ld + 0];
  const float ah = anchors[a * anchor_ld + 3] - anchors[a * anchor_ld + 1];
  const float hstep = ah / feat_stride / 3.f - 1.f;
  const float wstep = aw / feat_stride / 3.f - 1.f;
  float* o = om + i * 27;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ti = t / 3, tj = t % 3;
    o[2 * t] = hstep * (static_cast<float>(ti) - 1.5f + 0.5f) * hard;
    o[2 * t + 1] = wstep * (static_cast<float>(tj) - 1.5f + 0.5f) * hard;
    o[18 + t] = fg;
  }
}

// center_align offsets (feturealign_mgpu.py:58-77): the 1x1 DCNv2 samples at
// (dy, dx) = ((by*std_y + mean_y) * ah/stride, (bx*std_x + mean_x) * aw/stride)
// of the top-1 anchor, zeroed where fg <= thresh; mask = fg.  om: [N,H,W,3].
struct CenterAlignSet {
  int x_coff, y_coff;
  float mean_x, mean_y, std_x, std_y;
  float* om;
};
__device__ __forceinline__ void center_align_om_pixel(float fg, int a, const float* __restrict__ heads, int heads_cstride,
                                                      int x_coff, int y_coff, const float* __restrict__ anchors,
                                                      int anchor_ld, float feat_stride, float mean_x, float mean_y,
                                                      float std_x, float std_y, float thresh, float* __restrict__ om,
                                                      int om_cstride, long i);

__global__ void center_align_om_kernel(const float* __restrict__ fg_max, const int* __restrict__ fg_arg,
                                       const float* __restrict__ heads, int heads_cstride, int x_coff, int y_coff,
                                       const float* __restrict__ anchors, int anchor_ld, float feat_stride,
                                       float mean_x, float mean_y, float std_x, float std_y, float thresh,
                                       float* __restrict__ om, int om_cstride, long npix) {
  grid_dep_sync();  // PDL: launched while the previous kernel drains
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= npix) return;
  center_align_om_pixel(fg_max[i], fg_arg[i], heads, heads_cstride, x_coff, y_coff, anchors, anchor_ld, feat_stride, mean_x,
                        mean_y, std_x, std_y, thresh, om, om_cstride, i);
}

// both centre alignments (2D and 3D centres) of a pixel in one launch
__global__ void center_align_om2_kernel(const float* __restrict__ fg_max, const int* __restrict__ fg_arg,
                                        const float* __restrict__ heads, int heads_cstride, const CenterAlignSet s0,
                                        const CenterAlignSet s1, const float* __restrict__ anchors, int anchor_ld,
                                        float feat_stride, float thresh, int om_cstride, long npix) {
  grid_dep_sync();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= npix) return;
  const float fg = fg_max[i];
  const int a = fg_arg[i];
  center_align_om_pixel(fg, a, heads, heads_cstride, s0.x_coff, s0.y_coff, anchors, anchor_ld, feat_stride, s0.mean_x,
                        s0.mean_y, s0.std_x, s0.std_y, thresh, s0.om, om_cstride, i);
  center_align_om_pixel(fg, a, heads, heads_cstride, s1.x_coff, s1.y_coff, anchors, anchor_ld, feat_stride, s1.mean_x,
                        s1.mean_y, s1.std_x, s1.std_y, thresh, s1.om, om_cstride, i);
}

__device__ __forceinline__ void center_align_om_pixel(float fg, int a, const float* __restrict__ heads, int heads_cstride,
                                                      int x_coff, int y_coff, const float* __restrict__ anchors,
                                                      int anchor_ld, float feat_stride, float mean_x, float mean_y,
                                                      float std_x, float std_y, float thresh, float* __restrict__ om,
                                                      int om_cstride, long i) {
  const float hard = fg > thresh ? 1.f : 0.f;
  const float aw = (anchors[a * anchor_ld + 2] - anchors[a * anchor_ld + 0]) / feat_stride;
  const float ah = (anchors[a * anchor_ld + 3] - anchors[a * anchor_ld + 1]) / feat_stride;
  const float bx = heads[i * heads_cstride + x_coff + a];
  const float by = heads[i * heads_cstride + y_coff + a];
  om[i * om_cstride + 0] = (by * std_y + mean_y) * ah * hard;
  om[i * om_cstride + 1] = (bx * std_x + mean_x) * aw * hard;
  om[i * om_cstride + 2] = fg;
}

// ---------------------------------------------------------------------------
// flatten_tensor + cat for the regression heads (M3d_inference_align.py:280-295):
//   heads NHWC fp32 [N,H,W,11*A] (head-major: x,y,w,h,x3d,y3d,z3d,w3d,h3d,l3d,rY3d)
//   bbox_2d [N, (a*H + h)*W + w, 4], bbox_3d [N, same, 7]
// ---------------------------------------------------------------------------
struct HeadSlots {
  int s[11];  // buffer slot (36-channel group) of x,y,w,h,x3d,y3d,z3d,w3d,h3d,l3d,rY3d
};

__global__ void __launch_bounds__(256) flatten_heads_kernel(const float* __restrict__ heads, int hc_stride, int N, int H,
                                                            int W, int A, const HeadSlots slots,
                                                            float* __restrict__ bbox_2d, float* __restrict__ bbox_3d) {
  grid_dep_sync();  // PDL: launched while the previous kernel drains
  extern __shared__ float s_tile[];  // [SM_PIX][11*A + 1]
  const int C = 11 * A, ld = C + 1;
  const int w0 = blockIdx.x * SM_PIX, h = blockIdx.y, n = blockIdx.z;
  const int npix = min(SM_PIX, W - w0);
  const float* src = heads + ((static_cast<long>(n) * H + h) * W + w0) * hc_stride;
  if ((C & 3) == 0 && (hc_stride & 3) == 0) {
    const int c4 = C / 4;
    for (int i = threadIdx.x; i < npix * c4; i += blockDim.x) {
      const int px = i / c4, c = (i - px * c4) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<long>(px) * hc_stride + c));
      float* d = s_tile + px * ld + c;
      d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < npix * C; i += blockDim.x) {
      const int px = i / C, c = i - px * C;
      s_tile[px * ld + c] = __ldg(src + static_cast<long>(px) * hc_stride + c);
    }
  }
  __syncthreads();
  // bbox_2d: thread -> (anchor, pixel), pixel fastest: one 16-byte row each, 512 contiguous bytes per warp
  for (int i = threadIdx.x; i < A * SM_PIX; i += blockDim.x) {
    const int px = i % SM_PIX, a = i / SM_PIX;
    if (px >= npix) continue;
    const float* t = s_tile + px * ld + a;
    const long row = (static_cast<long>(n) * A + a) * H * W + static_cast<long>(h) * W + (w0 + px);
    *reinterpret_cast<float4*>(bbox_2d + row * 4) =
        make_float4(t[slots.s[0] * A], t[slots.s[1] * A], t[slots.s[2] * A], t[slots.s[3] * A]);
  }
  // bbox_3d: the npix x 7 floats of one anchor are contiguous in the output: thread -> element, fully coalesced
  // 4-byte stores (the 28-byte rows written one thread each cost 7 partial-sector stores per row)
  const int per_a = npix * 7;  // <= 224 < blockDim.x
  if (threadIdx.x < per_a) {
    const int e = threadIdx.x;
    const int px = e / 7, j = e - px * 7;
    const float* t = s_tile + px * ld + slots.s[4 + j] * A;
    float* o = bbox_3d + ((static_cast<long>(n) * A) * H * W + static_cast<long>(h) * W + w0) * 7 + e;
    const long a_stride = static_cast<long>(H) * W * 7;
#pragma unroll 4
    for (int a = 0; a < A; ++a) o[a * a_stride] = t[a];
  }
}

// ---------------------------------------------------------------------------
// Layout conversion for the NCHW drop-in operators (DCNv2 / ANAB modules).
// 32x32 shared-memory transpose between (C) and (H*W).
// ---------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void nchw_to_nhwc_kernel(const TI* __restrict__ in, TO* __restrict__ out, int C, int HW, int out_cstride,
                                    int out_coff) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = to_f(in[(static_cast<long>(n) * C + c) * HW + p]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) out[(static_cast<long>(n) * HW + p) * out_cstride + out_coff + c] = from_f<TO>(tile[threadIdx.x][i]);
  }
}

template <typename TI, typename TO>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ in, TO* __restrict__ out, int C, int HW, int in_cstride,
                                    int in_coff) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = to_f(in[(static_cast<long>(n) * HW + p) * in_cstride + in_coff + c]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[(static_cast<long>(n) * C + c) * HW + p] = from_f<TO>(tile[threadIdx.x][i]);
  }
}

// ---------------------------------------------------------------------------
// Input pipeline: the reference's Normalize transform + BGR->RGB + HWC->CHW
// (lib/augmentations.py:44-57, lib/dataloader.py:942-950) on the device, from the uint8 HWC image cv2.imread
// returns.  Arithmetic in the reference's order and type (fp32: x / 255, - mean[c], / std[c], with mean/std
// indexed by the HWC channel BEFORE the swap -- quirk 8 of SURVEY appendix B is preserved), IEEE division, so the
// result is bit-identical to numpy's.  4 pixels (12 bytes) per thread, one float4 store per plane.
// ---------------------------------------------------------------------------
struct Norm3 {
  float mean[3], stdv[3];
};

__global__ void __launch_bounds__(256) preprocess_u8_kernel(const unsigned char* __restrict__ img, float* __restrict__ out,
                                                            long npix4, long HW, Norm3 nm, int swap_rb) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;  // group of 4 pixels inside the batch
  if (i >= npix4) return;
  const long pix = i * 4;
  const long n = pix / HW, p = pix - n * HW;
  const uint3 raw = *reinterpret_cast<const uint3*>(img + pix * 3);  // 12 bytes: 4 pixels x 3 channels
  const unsigned int w[3] = {raw.x, raw.y, raw.z};
  float v[3][4];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const unsigned int byte = (w[k >> 2] >> ((k & 3) * 8)) & 0xffu;
    const int c = k % 3, px = k / 3;
    float x = __fdiv_rn(static_cast<float>(byte), 255.0f);
    x = __fsub_rn(x, nm.mean[c]);
    v[c][px] = __fdiv_rn(x, nm.stdv[c]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int oc = swap_rb ? 2 - c : c;
    *reinterpret_cast<float4*>(out + (n * 3 + oc) * HW + p) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

// ---------------------------------------------------------------------------
// The reference's whole test-time transform, Preprocess = ConvertToFloat + Padding(size) + Normalize
// (lib/augmentations.py:472-492, Padding :136-160), + BGR->RGB + HWC->CHW, for a RAGGED batch: image n is uint8 HWC
// [h[n], w[n], 3] (KITTI frames are 370-376 x 1224-1242) and is padded on the bottom / right to H x W.  cv2's
// copyMakeBorder pads with 0 BEFORE Normalize, so a padded pixel becomes (0/255 - mean[c]) / std[c], not 0 -- kept.
// Same arithmetic as preprocess_u8_kernel (bit-identical to numpy).  4 output pixels per thread, byte loads (rows of
// 3*w bytes have no alignment), one float4 store per plane.
// ---------------------------------------------------------------------------
constexpr int kPadBatch = 64;
struct RaggedImages {
  long long off[kPadBatch];
  int h[kPadBatch], w[kPadBatch];
};

__global__ void __launch_bounds__(256) preprocess_u8_pad_kernel(const unsigned char* __restrict__ img, float* __restrict__ out,
                                                                RaggedImages ri, int H, int W, Norm3 nm, int swap_rb) {
  const int n = blockIdx.y;
  const int W4 = W >> 2;
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(H) * W4) return;
  const int y = static_cast<int>(i / W4), x0 = static_cast<int>(i - static_cast<long>(y) * W4) * 4;
  const int h = ri.h[n], w = ri.w[n];
  const unsigned char* row = img + ri.off[n] + static_cast<long>(y) * w * 3;
  float v[3][4];
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const bool in = y < h && x0 + px < w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned int byte = in ? __ldg(row + (x0 + px) * 3 + c) : 0u;
      float x = __fdiv_rn(static_cast<float>(byte), 255.0f);
      x = __fsub_rn(x, nm.mean[c]);
      v[c][px] = __fdiv_rn(x, nm.stdv[c]);
    }
  }
  const long HW = static_cast<long>(H) * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int oc = swap_rb ? 2 - c : c;
    *reinterpret_cast<float4*>(out + (static_cast<long>(n) * 3 + oc) * HW + static_cast<long>(y) * W + x0) =
        make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

// ---------------------------------------------------------------------------
// Backward of the depthwise ConvTranspose2d(2f, stride f, pad f/2) up-sampling (training path), bf16 NHWC:
//   gx[n,iy,ix,c]   = sum_{ky,kx} gy[n, iy*f - pad + ky, ix*f - pad + kx, c] * w[ky,kx,c]
//   gw[ky,kx,c]     = sum_{n,iy,ix} gy[n, iy*f - pad + ky, ix*f - pad + kx, c] * x[n,iy,ix,c]
// One thread = one input pixel x 8 channels; the weight gradient is reduced per block in shared memory in 64-bit FIXED
// POINT (integer atomics: the order of arrival does not change the sum) and written as per-block partials
// [blocks][k*k][C], added in block order by a second kernel: deterministic end to end.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const __nv_bfloat16* __restrict__ gy,
                                                           const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                                                           __nv_bfloat16* __restrict__ gx, float* __restrict__ gw_partial,
                                                           int N, int H, int W, int C, int f) {
  extern __shared__ unsigned long long s_gw[];  // [k*k][C], fixed point 2^-32
  const int k = 2 * f, pad = f / 2, kk = k * k;
  const int Ho = H * f, Wo = W * f, cv = C / 8;
  for (int i = threadIdx.x; i < kk * C; i += blockDim.x) s_gw[i] = 0ull;
  __syncthreads();
  const long total = static_cast<long>(N) * H * W * cv;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % cv) * 8;
    long pix = idx / cv;
    const int ix = static_cast<int>(pix % W);
    pix /= W;
    const int iy = static_cast<int>(pix % H), n = static_cast<int>(pix / H);
    float xv[8], acc[8];
    {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long>(n) * H + iy) * W + ix) * C + c));
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) xv[2 * q] = __uint_as_float(w4[q] << 16), xv[2 * q + 1] = __uint_as_float(w4[q] & 0xffff0000u);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int oy = iy * f - pad + ky;
      if (oy < 0 || oy >= Ho) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ox = ix * f - pad + kx;
        if (ox < 0 || ox >= Wo) continue;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(gy + ((static_cast<long>(n) * Ho + oy) * Wo + ox) * C + c));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
        const float* wp = wt + (ky * k + kx) * C + c;
        unsigned long long* gp = s_gw + (ky * k + kx) * C + c;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float g0 = __uint_as_float(w4[q] << 16), g1 = __uint_as_float(w4[q] & 0xffff0000u);
          acc[2 * q] += g0 * __ldg(wp + 2 * q);
          acc[2 * q + 1] += g1 * __ldg(wp + 2 * q + 1);
          atomicAdd(gp + 2 * q, static_cast<unsigned long long>(__float2ll_rn(g0 * xv[2 * q] * 4294967296.f)));
          atomicAdd(gp + 2 * q + 1, static_cast<unsigned long long>(__float2ll_rn(g1 * xv[2 * q + 1] * 4294967296.f)));
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat162 b2 = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
      o[q] = *reinterpret_cast<uint32_t*>(&b2);
    }
    *reinterpret_cast<uint4*>(gx + ((static_cast<long>(n) * H + iy) * W + ix) * C + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kk * C; i += blockDim.x)
    gw_partial[static_cast<long>(blockIdx.x) * kk * C + i] =
        static_cast<float>(static_cast<double>(static_cast<long long>(s_gw[i])) * (1.0 / 4294967296.0));
}

__global__ void upsample_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int n, float* __restrict__ gw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[static_cast<long>(b) * n + i];
  gw[i] = s;
}

}  // namespace m3d

using namespace m3d;

static inline cudaStream_t S(m3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kUpBwdBlocks = 296;
extern "C" size_t m3d_upsample_backward_workspace(int C, int f) {
  return static_cast<size_t>(kUpBwdBlocks) * 4 * f * f * C * sizeof(float);
}

extern "C" int m3d_upsample_backward(const void* gy, const void* x, const float* weight, void* gx, float* gw, int N, int H,
                                     int W, int C, int f, void* workspace, size_t workspace_bytes, m3d_stream_t stream) {
  M3D_REQUIRE(gy && x && weight && gx && gw && workspace, "NULL pointer");
  M3D_REQUIRE(C % 8 == 0 && f >= 1 && f <= 4, "C must be a multiple of 8, f in 1..4");
  const int kk = 4 * f * f;
  const size_t smem = static_cast<size_t>(kk) * C * sizeof(unsigned long long);
  M3D_REQUIRE(smem <= 48 * 1024, "k*k*C too large for the shared weight-gradient tile");
  if (workspace_bytes < m3d_upsample_backward_workspace(C, f)) {
    set_last_error("upsample backward workspace too small");
    return M3D_ERR_WORKSPACE;
  }
  float* partial = static_cast<float*>(workspace);
  upsample_bwd_kernel<<<kUpBwdBlocks, 256, smem, S(stream)>>>(static_cast<const __nv_bfloat16*>(gy),
                                                               static_cast<const __nv_bfloat16*>(x), weight,
                                                               static_cast<__nv_bfloat16*>(gx), partial, N, H, W, C, f);
  M3D_CUDA_OK(cudaGetLastError());
  upsample_bwd_reduce_kernel<<<(kk * C + 255) / 256, 256, 0, S(stream)>>>(partial, kUpBwdBlocks, kk * C, gw);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_preprocess_u8(const unsigned char* image_hwc, float* out_nchw, int N, int H, int W, const float* mean3,
                                 const float* std3, int swap_rb, m3d_stream_t stream) {
  M3D_REQUIRE(image_hwc && out_nchw && mean3 && std3, "NULL pointer");
  const long HW = static_cast<long>(H) * W;
  M3D_REQUIRE(HW % 4 == 0, "H*W must be a multiple of 4 (got %dx%d)", H, W);
  M3D_REQUIRE((reinterpret_cast<uintptr_t>(image_hwc) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_nchw) & 15) == 0,
              "image must be 4-byte aligned, output 16-byte aligned");
  Norm3 nm;
  for (int c = 0; c < 3; ++c) nm.mean[c] = mean3[c], nm.stdv[c] = std3[c];
  const long npix4 = static_cast<long>(N) * HW / 4;
  preprocess_u8_kernel<<<cdiv(npix4, 256), 256, 0, S(stream)>>>(image_hwc, out_nchw, npix4, HW, nm, swap_rb);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_preprocess_u8_pad(const unsigned char* images, const long long* offsets, const int* heights,
                                     const int* widths, float* out_nchw, int N, int H, int W, const float* mean3,
                                     const float* std3, int swap_rb, m3d_stream_t stream) {
  M3D_REQUIRE(images && offsets && heights && widths && out_nchw && mean3 && std3, "NULL pointer");
  M3D_REQUIRE(N >= 0 && H > 0 && W > 0 && W % 4 == 0, "W must be a multiple of 4 (got %dx%d)", H, W);
  M3D_REQUIRE((reinterpret_cast<uintptr_t>(out_nchw) & 15) == 0, "output must be 16-byte aligned");
  for (int n = 0; n < N; ++n)  // cv2.copyMakeBorder raises on a negative border: an image larger than the size is an error
    M3D_REQUIRE(heights[n] >= 0 && widths[n] >= 0 && heights[n] <= H && widths[n] <= W && offsets[n] >= 0,
                "image %d is %dx%d, larger than the padded size %dx%d", n, heights[n], widths[n], H, W);
  Norm3 nm;
  for (int c = 0; c < 3; ++c) nm.mean[c] = mean3[c], nm.stdv[c] = std3[c];
  const long per_image = static_cast<long>(H) * (W / 4);
  for (int n0 = 0; n0 < N; n0 += kPadBatch) {
    const int nb = N - n0 < kPadBatch ? N - n0 : kPadBatch;
    RaggedImages ri;
    for (int k = 0; k < kPadBatch; ++k) {
      ri.off[k] = k < nb ? offsets[n0 + k] : 0;
      ri.h[k] = k < nb ? heights[n0 + k] : 0;
      ri.w[k] = k < nb ? widths[n0 + k] : 0;
    }
    preprocess_u8_pad_kernel<<<dim3(static_cast<unsigned>(cdiv(per_image, 256)), nb), 256, 0, S(stream)>>>(
        images, out_nchw + static_cast<long>(n0) * 3 * H * W, ri, H, W, nm, swap_rb);
    M3D_CUDA_OK(cudaGetLastError());
  }
  return M3D_OK;
}

extern "C" int m3d_stem_conv7x7(const float* image_nchw, const float* weight, const float* bias, void* out, int out_dtype,
                                int out_cstride, int N, int H, int W, float slope, m3d_stream_t stream) {
  M3D_REQUIRE(image_nchw && weight && bias && out, "NULL pointer");
  M3D_REQUIRE(out_cstride >= 16 && out_cstride % 8 == 0, "out_cstride=%d", out_cstride);
  dim3 grid(cdiv(W, STEM_TW), cdiv(H, STEM_TH), N);
  if (out_dtype == M3D_BF16)
    stem_conv7x7_kernel<__nv_bfloat16><<<grid, 256, 0, S(stream)>>>(image_nchw, weight, bias,
                                                                    static_cast<__nv_bfloat16*>(out), out_cstride, N, H, W, slope);
  else
    stem_conv7x7_kernel<float><<<grid, 256, 0, S(stream)>>>(image_nchw, weight, bias, static_cast<float*>(out),
                                                            out_cstride, N, H, W, slope);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_maxpool2x2_nhwc(const void* in, void* out, int dtype, int N, int H, int W, int C, int in_cstride,
                                   int out_cstride, m3d_stream_t stream) {
  M3D_REQUIRE(in && out, "NULL pointer");
  M3D_REQUIRE(H % 2 == 0 && W % 2 == 0, "max-pool needs even H, W (got %dx%d)", H, W);
  const int V = dtype == M3D_BF16 ? 8 : 4;
  M3D_REQUIRE(C % V == 0 && in_cstride % V == 0 && out_cstride % V == 0, "channels must keep 16-byte vectors");
  const long total = static_cast<long>(N) * (H / 2) * (W / 2) * (C / V);
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16));
  if (dtype == M3D_BF16)
    M3D_CUDA_OK(launch_pdl(maxpool2x2_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, S(stream), static_cast<const __nv_bfloat16*>(in),
                                                                  static_cast<__nv_bfloat16*>(out), N, H, W, C, in_cstride, out_cstride));
  else
    M3D_CUDA_OK(launch_pdl(maxpool2x2_kernel<float>, dim3(grid), dim3(256), 0, S(stream), static_cast<const float*>(in), static_cast<float*>(out), N, H,
                                                          W, C, in_cstride, out_cstride));
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_upsample_add_nhwc(const void* x, const float* weight, const void* skip, void* out, int dtype, int N,
                                     int H, int W, int C, int f, int x_cstride, int skip_cstride, int out_cstride,
                                     m3d_stream_t stream) {
  M3D_REQUIRE(x && weight && out, "NULL pointer");
  M3D_REQUIRE(f >= 1 && f <= 8, "up-sampling factor %d", f);
  const int V = dtype == M3D_BF16 ? 8 : 4;
  M3D_REQUIRE(C % V == 0 && x_cstride % V == 0 && out_cstride % V == 0 && (skip == nullptr || skip_cstride % V == 0),
              "channels must keep 16-byte vectors");
  M3D_REQUIRE(H * f <= 65535 && N <= 65535, "upsample: output too tall for the launch grid");
  const unsigned gx = static_cast<unsigned>((static_cast<long>(W) * f * (C / V) + 255) / 256);
  // ~8 blocks per SM in all; grid.y = f row-parity classes x slices of the input rows
  long slices = (148L * 8 + static_cast<long>(gx) * N * f - 1) / (static_cast<long>(gx) * N * f);
  if (slices < 1) slices = 1;
  if (slices > H) slices = H;
  const unsigned gy = static_cast<unsigned>(slices * f);
  const dim3 grid(gx, gy, static_cast<unsigned>(N));
  if (dtype == M3D_BF16)
    M3D_CUDA_OK(launch_pdl(upsample_add_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, S(stream), 
        static_cast<const __nv_bfloat16*>(x), weight, static_cast<const __nv_bfloat16*>(skip),
        static_cast<__nv_bfloat16*>(out), N, H, W, C, f, x_cstride, skip_cstride, out_cstride));
  else
    M3D_CUDA_OK(launch_pdl(upsample_add_kernel<float>, dim3(grid), dim3(256), 0, S(stream), static_cast<const float*>(x), weight,
                                                            static_cast<const float*>(skip), static_cast<float*>(out), N,
                                                            H, W, C, f, x_cstride, skip_cstride, out_cstride));
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

static int cls_softmax_impl(const float* logits, int logits_cstride, int N, int H, int W, int A, int K, float* cls_out,
                            float* prob_out, float* fg_max, int* fg_arg, float* score, unsigned char* cls_pred,
                            const float* anchors, int anchor_ld, float feat_stride, float thresh, float* shape_om,
                            m3d_stream_t stream) {
  M3D_REQUIRE(logits && fg_max && fg_arg && score && cls_pred, "NULL pointer");
  M3D_REQUIRE((cls_out == nullptr) == (prob_out == nullptr), "cls_out and prob_out: both or neither");
  M3D_REQUIRE(K >= 2 && K <= 8 && A >= 1, "K=%d A=%d unsupported", K, A);
  dim3 grid(cdiv(W, SM_PIX), H, N);
  if (K == 4 && logits_cstride % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    const size_t smem4 = static_cast<size_t>(5 * A) * (SM_PIX + 1) * sizeof(float);
    if (smem4 <= 48 * 1024) {
      M3D_CUDA_OK(launch_pdl(cls_softmax4_kernel, grid, dim3(256), smem4, S(stream), logits, logits_cstride, N, H, W, A,
                             cls_out, prob_out, fg_max, fg_arg, score, cls_pred, anchors, anchor_ld, feat_stride, thresh,
                             shape_om));
      return M3D_OK;
    }
  }
  const size_t smem = static_cast<size_t>(SM_PIX) * (K * A + 1) * sizeof(float);
  M3D_REQUIRE(smem <= 48 * 1024, "K*A too large");
  M3D_CUDA_OK(launch_pdl(cls_softmax_kernel, dim3(grid), dim3(256), smem, S(stream), logits, logits_cstride, N, H, W, A, K, cls_out, prob_out, fg_max,
                                                     fg_arg, score, cls_pred));
  M3D_CUDA_OK(cudaGetLastError());
  if (shape_om != nullptr)
    return m3d_shape_align_om(fg_max, fg_arg, anchors, anchor_ld, feat_stride, thresh, shape_om,
                              static_cast<long>(N) * H * W, stream);
  return M3D_OK;
}

extern "C" int m3d_cls_softmax(const float* logits, int logits_cstride, int N, int H, int W, int A, int K, float* cls_out,
                               float* prob_out, float* fg_max, int* fg_arg, float* score, unsigned char* cls_pred,
                               m3d_stream_t stream) {
  return cls_softmax_impl(logits, logits_cstride, N, H, W, A, K, cls_out, prob_out, fg_max, fg_arg, score, cls_pred, nullptr,
                          0, 0.f, 0.f, nullptr, stream);
}

extern "C" int m3d_cls_softmax_shape_om(const float* logits, int logits_cstride, int N, int H, int W, int A, int K,
                                        float* cls_out, float* prob_out, float* fg_max, int* fg_arg, float* score,
                                        unsigned char* cls_pred, const float* anchors, int anchor_ld, float feat_stride,
                                        float thresh, float* shape_om, m3d_stream_t stream) {
  M3D_REQUIRE(anchors && shape_om, "NULL pointer");
  return cls_softmax_impl(logits, logits_cstride, N, H, W, A, K, cls_out, prob_out, fg_max, fg_arg, score, cls_pred, anchors,
                          anchor_ld, feat_stride, thresh, shape_om, stream);
}

extern "C" int m3d_shape_align_om(const float* fg_max, const int* fg_arg, const float* anchors, int anchor_ld,
                                  float feat_stride, float thresh, float* om, long npix, m3d_stream_t stream) {
  M3D_REQUIRE(fg_max && fg_arg && anchors && om, "NULL pointer");
  M3D_CUDA_OK(launch_pdl(shape_align_om_kernel, dim3(cdiv(npix, 256)), dim3(256), 0, S(stream), fg_max, fg_arg, anchors, anchor_ld, feat_stride, thresh,
                                                                om, npix));
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_center_align_om(const float* fg_max, const int* fg_arg, const float* heads, int heads_cstride,
                                   int x_coff, int y_coff, const float* anchors, int anchor_ld, float feat_stride,
                                   float mean_x, float mean_y, float std_x, float std_y, float thresh, float* om,
                                   int om_cstride, long npix, m3d_stream_t stream) {
  M3D_REQUIRE(fg_max && fg_arg && heads && anchors && om, "NULL pointer");
  M3D_REQUIRE(om_cstride >= 3, "om_cstride=%d", om_cstride);
  M3D_CUDA_OK(launch_pdl(center_align_om_kernel, dim3(cdiv(npix, 256)), dim3(256), 0, S(stream), fg_max, fg_arg, heads, heads_cstride, x_coff, y_coff,
                                                                 anchors, anchor_ld, feat_stride, mean_x, mean_y, std_x,
                                                                 std_y, thresh, om, om_cstride, npix));
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_center_align_om2(const float* fg_max, const int* fg_arg, const float* heads, int heads_cstride,
                                    const int* xy_coff4, const float* mean4, const float* std4, const float* anchors,
                                    int anchor_ld, float feat_stride, float thresh, float* om_a, float* om_b,
                                    int om_cstride, long npix, m3d_stream_t stream) {
  M3D_REQUIRE(fg_max && fg_arg && heads && anchors && om_a && om_b && xy_coff4 && mean4 && std4, "NULL pointer");
  M3D_REQUIRE(om_cstride >= 3, "om_cstride=%d", om_cstride);
  CenterAlignSet s0{xy_coff4[0], xy_coff4[1], mean4[0], mean4[1], std4[0], std4[1], om_a};
  CenterAlignSet s1{xy_coff4[2], xy_coff4[3], mean4[2], mean4[3], std4[2], std4[3], om_b};
  M3D_CUDA_OK(launch_pdl(center_align_om2_kernel, dim3(cdiv(npix, 256)), dim3(256), 0, S(stream), fg_max, fg_arg, heads,
                         heads_cstride, s0, s1, anchors, anchor_ld, feat_stride, thresh, om_cstride, npix));
  return M3D_OK;
}

extern "C" int m3d_flatten_heads(const float* heads, int heads_cstride, int N, int H, int W, int A,
                                 const int* slot_of_output, float* bbox_2d, float* bbox_3d, m3d_stream_t stream) {
  M3D_REQUIRE(heads && bbox_2d && bbox_3d && slot_of_output, "NULL pointer");
  HeadSlots slots;
  for (int i = 0; i < 11; ++i) {
    M3D_REQUIRE(slot_of_output[i] >= 0 && slot_of_output[i] < 11, "bad head slot");
    slots.s[i] = slot_of_output[i];
  }
  const size_t smem = static_cast<size_t>(SM_PIX) * (11 * A + 1) * sizeof(float);
  M3D_REQUIRE(smem <= 96 * 1024, "A too large");
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(flatten_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  M3D_ONCE_PER_DEVICE_END
  dim3 grid(cdiv(W, SM_PIX), H, N);
  M3D_CUDA_OK(launch_pdl(flatten_heads_kernel, dim3(grid), dim3(256), smem, S(stream), heads, heads_cstride, N, H, W, A, slots, bbox_2d, bbox_3d));
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_nchw_to_nhwc(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                                int out_cstride, int out_coff, m3d_stream_t stream) {
  M3D_REQUIRE(in && out, "NULL pointer");
  M3D_REQUIRE(in_dtype == M3D_F32, "NCHW side must be fp32");
  const int HW = H * W;
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), N), block(32, 8);
  if (out_dtype == M3D_BF16)
    nchw_to_nhwc_kernel<float, __nv_bfloat16><<<grid, block, 0, S(stream)>>>(
        static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), C, HW, out_cstride, out_coff);
  else
    nchw_to_nhwc_kernel<float, float><<<grid, block, 0, S(stream)>>>(static_cast<const float*>(in),
                                                                     static_cast<float*>(out), C, HW, out_cstride, out_coff);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_nhwc_to_nchw(const void* in, int in_dtype, void* out, int out_dtype, int N, int C, int H, int W,
                                int in_cstride, int in_coff, m3d_stream_t stream) {
  M3D_REQUIRE(in && out, "NULL pointer");
  M3D_REQUIRE(out_dtype == M3D_F32, "NCHW side must be fp32");
  const int HW = H * W;
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), N), block(32, 8);
  if (in_dtype == M3D_BF16)
    nhwc_to_nchw_kernel<__nv_bfloat16, float><<<grid, block, 0, S(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<float*>(out), C, HW, in_cstride, in_coff);
  else
    nhwc_to_nchw_kernel<float, float><<<grid, block, 0, S(stream)>>>(static_cast<const float*>(in),
                                                                     static_cast<float*>(out), C, HW, in_cstride, in_coff);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}
