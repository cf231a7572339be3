// Weight gradient of a convolution on the tensor cores (training path, BASELINE config 4):
//
//   dW[co][ci][r][s] = sum_{n,p,q} gy[n, p, q, co] * x[n, p*stride - pad + r*dil, q*stride - pad + s*dil, ci]
//
// (what torch / cuDNN compute for nn.Conv2d in the reference's training loop, scripts/train_rpn_3d.py:204-218).
// Per kernel tap (r, s) this is a GEMM  dW_rs[Cout][Cin] = GY^T [Cout][pixels] * X_rs [pixels][Cin]  whose reduction
// dimension is the PIXEL index.  In NHWC both operands have their channels contiguous and the pixels strided, i.e.
// they are MN-major tiles -- which tcgen05 reads directly (instruction-descriptor bits a_major / b_major = 1, shared
// memory descriptor of the MN-major 128-byte-swizzle atom: 64 channels x 8 pixels), so nothing is transposed:
//
//   * k-block = 64 output pixels (a TW x TH patch of one image).  TMA boxes {64 channels, TW, TH} of gy (2 boxes =
//     128 output channels) and of the input window shifted by the tap (BN / 64 boxes; element strides = the conv
//     stride; zero fill outside the image = the padding, and outside the channel range = channel padding for free);
//   * one CTA = (tap, 128-row Cout tile, BN-column Cin tile, K slice): the pixel range is split over enough slices to
//     fill the device (split-K), accumulators in TMEM, 4-stage TMA ring, 4 k16 MMAs per k-block;
//   * partial tiles go to a workspace [slice][tap][Cout_pad][Cin_pad] (plain stores) and a second kernel adds the
//     slices in a FIXED order into dW in torch's [Cout][Cin][R][S] layout: deterministic, no float atomics.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdint>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace m3d {

namespace {

constexpr int kStages = 4;
constexpr int kKB = 64;  // pixels per k-block

struct WgradParams {
  CUtensorMap tmap_gy;  // bf16 NHWC gy as (C, Q, P, N), box {64, TW, TH, 1}
  CUtensorMap tmap_x;   // bf16 NHWC x  as (C, W, H, N), box {64, TW*stride, TH*stride, 1}, element strides {1, s, s, 1}
  int gy_coff, x_coff;
  int TW, TH, tiles_w, tiles_h, total_kb;  // k-blocks = N * tiles_h * tiles_w
  int R, S, stride, pad, dil;
  int co_tiles, ci_tiles;  // of 128 rows / BN columns
  int kslices, kb_per_slice;
  float* partial;  // [kslices][R*S][co_tiles*128][ci_tiles*BN]
};

// MN-major operand tile: atoms of 64 channels (128 bytes) x 8 pixels, 128-byte swizzle; atoms along K (pixels) are
// 1024 bytes apart (SBO), atoms along M/N (the next 64 channels = the next TMA box) `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int n) {
  return umma_idesc_bf16(n) | (1u << 15) | (1u << 16);  // A and B MN-major
}

template <int BN>
__global__ void __launch_bounds__(192, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int A_BYTES = 2 * kKB * 128;          // two 64-channel boxes
  constexpr int B_BYTES = (BN / 64) * kKB * 128;  // BN / 64 boxes
  constexpr int STAGE = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item of this CTA
  int id = blockIdx.x;
  const int ci_t = id % p.ci_tiles;
  id /= p.ci_tiles;
  const int co_t = id % p.co_tiles;
  id /= p.co_tiles;
  const int taps = p.R * p.S;
  const int tap = id % taps;
  const int slice = id / taps;
  const int kb0 = slice * p.kb_per_slice;
  const int kb1 = min(p.total_kb, kb0 + p.kb_per_slice);
  const int nkb = max(0, kb1 - kb0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
    prefetch_tmap(&p.tmap_gy);
    prefetch_tmap(&p.tmap_x);
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const int r = tap / p.S, s = tap - r * p.S;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      int t = kb;
      const int tw = t % p.tiles_w;
      t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int n = t / p.tiles_h;
      const int q0 = tw * p.TW, p0 = th * p.TH;
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + stage * STAGE;
        mbar_arrive_expect_tx(&full[stage], STAGE);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_4d(sa + h * (kKB * 128), &p.tmap_gy, &full[stage], p.gy_coff + co_t * 128 + h * 64, q0, p0, n);
#pragma unroll
        for (int h = 0; h < BN / 64; ++h)
          tma_load_4d(sa + A_BYTES + h * (kKB * 128), &p.tmap_x, &full[stage], p.x_coff + ci_t * BN + h * 64,
                      q0 * p.stride - p.pad + s * p.dil, p0 * p.stride - p.pad + r * p.dil, n);
      }
      __syncwarp();
      if (++stage == kStages) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_mn(BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * STAGE);
#pragma unroll
        for (int k = 0; k < kKB / 16; ++k) {  // 16 pixels = two 8-pixel atoms = 2048 bytes per step
          const uint64_t da = umma_smem_desc_mn(sa + k * 2048, kKB * 128);
          const uint64_t db = umma_smem_desc_mn(sa + A_BYTES + k * 2048, kKB * 128);
          umma_f16(tmem_base, da, db, idesc, (i | k) != 0);
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == kStages) stage = 0, phase ^= 1;
    }
    if (elect_one()) umma_commit(tfull);
    __syncwarp();
  } else {
    // --------------------------------------------------------------- epilogue: TMEM -> partial tile (fp32)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // output channel inside the tile
    const long ld = static_cast<long>(p.ci_tiles) * BN;
    float* dst = p.partial + ((static_cast<long>(slice) * taps + tap) * (p.co_tiles * 128) + co_t * 128 + row) * ld +
                 ci_t * BN;
    if (nkb > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t acc[16];
      if (nkb > 0) {
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

// dW[co][ci][r][s] = sum over slices, in slice order
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int kslices, int taps,
                                    int co_pad, int ci_pad, int Cout, int Cin) {
  const long total = static_cast<long>(Cout) * Cin * taps;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % taps);
    const int ci = static_cast<int>((i / taps) % Cin);
    const int co = static_cast<int>(i / (static_cast<long>(taps) * Cin));
    const long tile = static_cast<long>(co_pad) * ci_pad;
    const float* src = partial + (static_cast<long>(tap) * co_pad + co) * ci_pad + ci;
    float acc = 0.f;
    for (int s = 0; s < kslices; ++s) acc += src[static_cast<long>(s) * taps * tile];
    dw[i] = acc;
  }
}

int make_tmap(CUtensorMap* map, const void* base, int C_extent, int cstride, int W, int H, int N, int box_w, int box_h,
              int estride) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  M3D_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  M3D_REQUIRE(enc != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C_extent), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(cstride) * 2, static_cast<cuuint64_t>(W) * cstride * 2,
                           static_cast<cuuint64_t>(H) * W * cstride * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_w * estride), static_cast<cuuint32_t>(box_h * estride), 1};
  cuuint32_t es[4] = {1, static_cast<cuuint32_t>(estride), static_cast<cuuint32_t>(estride), 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad C=%d/%d W=%d H=%d N=%d box=%dx%d stride=%d) failed: %d",
              C_extent, cstride, W, H, N, box_w, box_h, estride, static_cast<int>(r));
  return M3D_OK;
}

struct WgradPlan {
  int BN, co_tiles, ci_tiles, TW, TH, tiles_w, tiles_h, total_kb, kslices, kb_per_slice;
  size_t partial_bytes;
};

WgradPlan plan(int N, int P, int Q, int Cin, int Cout, int R, int S) {
  WgradPlan w;
  w.BN = Cin > 64 ? 128 : 64;
  w.co_tiles = (Cout + 127) / 128;
  w.ci_tiles = (Cin + w.BN - 1) / w.BN;
  // 64-pixel patch wasting the fewest padded pixels
  const int cands[4][2] = {{16, 4}, {8, 8}, {32, 2}, {64, 1}};
  long best = -1;
  w.TW = 16, w.TH = 4;
  for (int i = 0; i < 4; ++i) {
    const long tiles = static_cast<long>((Q + cands[i][0] - 1) / cands[i][0]) * ((P + cands[i][1] - 1) / cands[i][1]);
    if (best < 0 || tiles < best) best = tiles, w.TW = cands[i][0], w.TH = cands[i][1];
  }
  w.tiles_w = (Q + w.TW - 1) / w.TW, w.tiles_h = (P + w.TH - 1) / w.TH;
  w.total_kb = N * w.tiles_w * w.tiles_h;
  const int items = R * S * w.co_tiles * w.ci_tiles;
  int sms = persistent_sms();
  if (sms <= 0) sms = 148;
  int slices = std::max(1, (2 * sms) / items);  // up to two waves of small CTAs
  slices = std::min(slices, std::max(1, w.total_kb / 4));  // at least 4 k-blocks per slice
  w.kb_per_slice = (w.total_kb + slices - 1) / slices;
  w.kslices = (w.total_kb + w.kb_per_slice - 1) / w.kb_per_slice;
  w.partial_bytes = static_cast<size_t>(w.kslices) * R * S * w.co_tiles * 128 * w.ci_tiles * w.BN * 4;
  return w;
}

}  // namespace
}  // namespace m3d

namespace m3d {
namespace {
// Bias gradient of a convolution = per-channel sum of the output gradient (bf16 NHWC -> fp32[C]).  Two passes, fixed
// order (deterministic): blocks of 256 threads sum 8 channels x their slab of pixels, then the slabs are added.
constexpr int kSumSlabs = 64;
__global__ void __launch_bounds__(256) channel_sum_partial_kernel(const __nv_bfloat16* __restrict__ x, long npix, int C,
                                                                  int cstride, int coff, float* __restrict__ partial) {
  __shared__ float red[32][65];
  const int cgrp = blockIdx.x, slab = blockIdx.y;     // 64 channels per block column
  const int lane8 = threadIdx.x & 7, rlane = threadIdx.x >> 3;  // 8 threads x 8 channels, 32 row lanes
  const int c0 = cgrp * 64 + lane8 * 8;
  const long per = (npix + gridDim.y - 1) / gridDim.y;
  const long r0 = slab * per, r1 = min(npix, r0 + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool vec = c0 + 8 <= C && ((cstride | coff) & 7) == 0;
  for (long r = r0 + rlane; r < r1; r += 32) {
    const __nv_bfloat16* px = x + r * cstride + coff + c0;
    if (vec) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(px));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + i < C) acc[i] += __bfloat162float(px[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rlane][lane8 * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    const int c = cgrp * 64 + threadIdx.x;
    if (c < C) partial[static_cast<long>(slab) * C + c] = s;
  }
}
__global__ void channel_sum_final_kernel(const float* __restrict__ partial, int nslab, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int i = 0; i < nslab; ++i) s += partial[static_cast<long>(i) * C + c];
  out[c] = s;
}
}  // namespace
}  // namespace m3d

using namespace m3d;

extern "C" size_t m3d_channel_sum_workspace(int C) { return static_cast<size_t>(kSumSlabs) * C * sizeof(float); }

extern "C" int m3d_channel_sum(const void* x, long npix, int C, int cstride, int coff, float* out, void* workspace,
                               size_t workspace_bytes, m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3D_REQUIRE(x && out && workspace && npix >= 1 && C >= 1 && cstride >= C, "bad arguments");
  if (workspace_bytes < m3d_channel_sum_workspace(C)) {
    set_last_error("channel-sum workspace too small");
    return M3D_ERR_WORKSPACE;
  }
  float* partial = static_cast<float*>(workspace);
  const int slabs = static_cast<int>(std::min<long>(kSumSlabs, (npix + 255) / 256));
  dim3 grid((C + 63) / 64, slabs);
  channel_sum_partial_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), npix, C, cstride, coff, partial);
  M3D_CUDA_OK(cudaGetLastError());
  channel_sum_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(partial, slabs, C, out);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" size_t m3d_conv2d_wgrad_workspace(int N, int P, int Q, int Cin, int Cout, int R, int S) {
  return plan(N, P, Q, Cin, Cout, R, S).partial_bytes + 256;
}

extern "C" int m3d_conv2d_wgrad(const void* x, int x_cstride, int x_coff, const void* gy, int gy_cstride, int gy_coff,
                                float* dw, int N, int H, int W, int Cin, int P, int Q, int Cout, int R, int S, int stride,
                                int pad, int dil, void* workspace, size_t workspace_bytes, m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3D_REQUIRE(x && gy && dw && workspace, "NULL pointer");
  M3D_REQUIRE(N >= 1 && H >= 1 && W >= 1 && Cin >= 1 && P >= 1 && Q >= 1 && Cout >= 1 && R >= 1 && S >= 1 && stride >= 1,
              "bad geometry");
  M3D_REQUIRE(x_cstride % 8 == 0 && gy_cstride % 8 == 0, "channel strides must be multiples of 8 (16-byte TMA strides)");
  M3D_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0, "16-byte alignment");
  const WgradPlan w = plan(N, P, Q, Cin, Cout, R, S);
  M3D_REQUIRE(w.TW * stride <= 256 && w.TH * stride <= 256, "stride too large for a TMA box");
  if (workspace_bytes < w.partial_bytes) {
    set_last_error("wgrad workspace too small: %zu < %zu", workspace_bytes, w.partial_bytes);
    return M3D_ERR_WORKSPACE;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap(&p.tmap_gy, gy, gy_coff + Cout, gy_cstride, Q, P, N, w.TW, w.TH, 1);
  if (rc != M3D_OK) return rc;
  rc = make_tmap(&p.tmap_x, x, x_coff + Cin, x_cstride, W, H, N, w.TW, w.TH, stride);
  if (rc != M3D_OK) return rc;
  p.gy_coff = gy_coff, p.x_coff = x_coff;
  p.TW = w.TW, p.TH = w.TH, p.tiles_w = w.tiles_w, p.tiles_h = w.tiles_h, p.total_kb = w.total_kb;
  p.R = R, p.S = S, p.stride = stride, p.pad = pad, p.dil = dil;
  p.co_tiles = w.co_tiles, p.ci_tiles = w.ci_tiles;
  p.kslices = w.kslices, p.kb_per_slice = w.kb_per_slice;
  uintptr_t a = (reinterpret_cast<uintptr_t>(workspace) + 15) & ~static_cast<uintptr_t>(15);
  p.partial = reinterpret_cast<float*>(a);
  const int grid = w.kslices * R * S * w.co_tiles * w.ci_tiles;
  const int stage = (2 + w.BN / 64) * kKB * 128;
  const int smem = kStages * stage + 1024 + 256;
  if (w.BN == 128) {
    M3D_ONCE_PER_DEVICE_BEGIN
      M3D_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    M3D_ONCE_PER_DEVICE_END
    set_last_kernel("conv_wgrad_kernel<128>");
    conv_wgrad_kernel<128><<<grid, 192, smem, stream>>>(p);
  } else {
    M3D_ONCE_PER_DEVICE_BEGIN
      M3D_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    M3D_ONCE_PER_DEVICE_END
    set_last_kernel("conv_wgrad_kernel<64>");
    conv_wgrad_kernel<64><<<grid, 192, smem, stream>>>(p);
  }
  M3D_CUDA_OK(cudaGetLastError());
  const long total = static_cast<long>(Cout) * Cin * R * S;
  wgrad_reduce_kernel<<<static_cast<int>(std::min<long>((total + 255) / 256, 2048)), 256, 0, stream>>>(
      p.partial, dw, w.kslices, R * S, w.co_tiles * 128, w.ci_tiles * w.BN, Cout, Cin);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}
