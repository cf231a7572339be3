// ANAB attention on tensor cores (bf16 throughput mode), model/module/attention.py:183-216 + the BN / LeakyReLU
// that follow it (model/M3d_inference_align.py:168-173):
//
//   out = LeakyReLU(BN(softmax(Q K_tok^T) V_tok + x))        Q [HW, ck<=192], K_tok [T<=352, ck], V_tok [T, 128]
//
// One CTA per 128-query tile, nothing but the tile's Q, x and the image's tokens is read, the [HW, T] attention
// matrix never exists in memory (the reference materialises it: 10.4 MB / image):
//
//   1. TMA: Q tile (3 k-blocks of 64 channels, zero-filled past ck) + the image's K tokens (bf16, [352][192])
//   2. GEMM1 (tcgen05): S[128 x 352] = Q K^T into TMEM columns 0-351 (two N = 176 MMAs per k16)
//   3. TMA: V tokens (bf16, transposed [128][384]) into the shared memory GEMM1 has finished with
//   4. softmax: thread = query row; two passes over its 352 TMEM columns (max, then exp / sum); the unnormalised
//      probabilities go to shared memory as the bf16 A operand of GEMM2 (swizzled k-blocks); 1/sum stays in a register
//   5. GEMM2: O[128 x 128] = P V into TMEM columns 384-511
//   6. epilogue: O / sum + x -> BN scale / shift -> LeakyReLU -> bf16 NHWC
//
// Warp roles: warp 0 TMA, warp 1 MMA + TMEM allocation, warps 2-5 softmax / epilogue (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <cstring>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace m3d {

int make_tmap_b3d(CUtensorMap* map, const void* base, long rows, long cols, int bk, int box_rows, int ksub);

namespace {

constexpr int kTP = 352;    // token rows of K (N of GEMM1), two MMAs of 176
constexpr int kCKP = 192;   // padded key channels (K of GEMM1)
constexpr int kTK = 384;    // padded token columns of V^T / P (K of GEMM2, 6 k-blocks; 22 k16 steps are used)
constexpr int kCV = 128;
constexpr int kQBytes = 3 * 16384;
constexpr int kKBytes = 3 * kTP * 128;
constexpr int kPBytes = 6 * 16384;
constexpr int kVBytes = 6 * 16384;
constexpr int kAnabSmem = kPBytes + kVBytes + 1024 + 128;  // phase 1 (Q + K = 180 KB) fits inside phase 2 (192 KB)
static_assert(kQBytes + kKBytes <= kPBytes + kVBytes, "phase-1 operands must fit the phase-2 footprint");

struct alignas(64) AnabParams {
  CUtensorMap tmap_q;  // (C, HW, N) bf16, box {64, 128, 1}
  CUtensorMap tmap_k;  // (64, N*352, 3) bf16, box {64, 176, 1}
  CUtensorMap tmap_v;  // (64, N*128, 6) bf16, box {64, 128, 1}
  const __nv_bfloat16* x;
  int x_cs;
  const float *scale, *shift;
  float slope;
  __nv_bfloat16* out;
  int out_cs;
  int HW, T, tiles_per_image;
};

// fp32 tokens -> bf16 operands: Kb[n][t][c] (zero past T / ck), Vb[n][c][t] (V transposed, zero past T)
__global__ void anab_tokens_bf16_kernel(const float* __restrict__ ktok, const float* __restrict__ vtok, int T, int ck,
                                        int cv, __nv_bfloat16* __restrict__ kb, __nv_bfloat16* __restrict__ vb) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kTP * kCKP) {
    const int t = i / kCKP, c = i - t * kCKP;
    const float v = (t < T && c < ck) ? ktok[(static_cast<long>(n) * T + t) * ck + c] : 0.f;
    kb[static_cast<long>(n) * kTP * kCKP + i] = __float2bfloat16_rn(v);
  }
  if (i < kCV * kTK) {
    const int c = i / kTK, t = i - c * kTK;
    const float v = (t < T && c < cv) ? vtok[(static_cast<long>(n) * T + t) * cv + c] : 0.f;
    vb[static_cast<long>(n) * kCV * kTK + i] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(192, 1) anab_attention_tc_kernel(const __grid_constant__ AnabParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_q = smem;             // phase 1
  uint8_t* s_k = smem + kQBytes;   // phase 1: [3 k-blocks][352 rows][128 B]
  uint8_t* s_p = smem;             // phase 2: [6 k-blocks][128 rows][128 B]
  uint8_t* s_v = smem + kPBytes;   // phase 2: [6 k-blocks][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPBytes + kVBytes);
  uint64_t* b_qk = bars;       // Q + K landed
  uint64_t* b_s = bars + 1;    // GEMM1 complete (S in TMEM, phase-1 shared memory free)
  uint64_t* b_v = bars + 2;    // V landed
  uint64_t* b_p = bars + 3;    // P written (one arrival per softmax warp)
  uint64_t* b_o = bars + 4;    // GEMM2 complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(b_qk, 1);
    mbar_init(b_s, 1);
    mbar_init(b_v, 1);
    mbar_init(b_p, 4);
    mbar_init(b_o, 1);
    fence_barrier_init();
    prefetch_tmap(&p.tmap_q);
    prefetch_tmap(&p.tmap_k);
    prefetch_tmap(&p.tmap_v);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  const int n = blockIdx.x / p.tiles_per_image;
  const int q0 = (blockIdx.x - n * p.tiles_per_image) * 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(b_qk, kQBytes + kKBytes);
      for (int kb = 0; kb < 3; ++kb) tma_load_3d(s_q + kb * 16384, &p.tmap_q, b_qk, kb * 64, q0, n);
      for (int kb = 0; kb < 3; ++kb)
        for (int h = 0; h < 2; ++h)
          tma_load_3d(s_k + kb * (kTP * 128) + h * (176 * 128), &p.tmap_k, b_qk, 0, n * kTP + h * 176, kb);
    }
    __syncwarp();
    mbar_wait(b_s, 0);  // GEMM1 has consumed Q and K: their shared memory becomes P / V
    if (elect_one()) {
      mbar_arrive_expect_tx(b_v, kVBytes);
      for (int kb = 0; kb < 6; ++kb) tma_load_3d(s_v + kb * 16384, &p.tmap_v, b_v, 0, n * kCV, kb);
    }
    __syncwarp();
  } else if (warp == 1) {
    mbar_wait(b_qk, 0);
    tc_fence_after();
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(176);
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        const uint64_t da = umma_smem_desc<128>(smem_u32(s_q + kb * 16384));
        const uint64_t db = umma_smem_desc<128>(smem_u32(s_k + kb * (kTP * 128)));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc_s, (kb | k) != 0);
          umma_f16(tmem_base + 176, da + 2 * k, db + ((176 * 128) >> 4) + 2 * k, idesc_s, (kb | k) != 0);
        }
      }
      umma_commit(b_s);
    }
    __syncwarp();
    mbar_wait(b_p, 0);
    mbar_wait(b_v, 0);
    tc_fence_after();
    if (elect_one()) {
      constexpr uint32_t idesc_o = umma_idesc_bf16(kCV);
#pragma unroll 1
      for (int ks = 0; ks < kTP / 16; ++ks) {  // 22 k16 steps over the 352 token columns
        const int kb = ks >> 2, k = ks & 3;
        const uint64_t da = umma_smem_desc<128>(smem_u32(s_p + kb * 16384));
        const uint64_t db = umma_smem_desc<128>(smem_u32(s_v + kb * 16384));
        umma_f16(tmem_base + 384, da + 2 * k, db + 2 * k, idesc_o, ks != 0);
      }
      umma_commit(b_o);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    mbar_wait(b_s, 0);
    tc_fence_after();
    // pass 1: row maximum over the T valid columns
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < kTP; c0 += 32) {
      uint32_t a[32];
      tmem_ld32(tmem_base + lane_off + c0, a);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (c0 + e < p.T) mx = fmaxf(mx, __uint_as_float(a[e]));
    }
    // pass 2: e = exp(s - max) (bf16, as the tensor cores will see it), running sum of the rounded values
    float sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kTP; c0 += 32) {
      uint32_t a[32];
      tmem_ld32(tmem_base + lane_off + c0, a);
      tmem_ld_wait();
      uint8_t* slab = s_p + (c0 >> 6) * 16384;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + j * 8 + 2 * e;
          const float e0 = c < p.T ? __expf(__uint_as_float(a[j * 8 + 2 * e]) - mx) : 0.f;
          const float e1 = c + 1 < p.T ? __expf(__uint_as_float(a[j * 8 + 2 * e + 1]) - mx) : 0.f;
          const __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
          sum += __low2float(t) + __high2float(t);
          w[e] = *reinterpret_cast<const uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(slab + swizzled_offset<128>(row, ((c0 & 63) >> 3) + j)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(b_p);
    const float inv = 1.f / sum;
    // epilogue
    mbar_wait(b_o, 0);
    tc_fence_after();
    const int qi = q0 + row;
    const bool ok = qi < p.HW;
    const long pix = static_cast<long>(n) * p.HW + qi;
    const __nv_bfloat16* xr = p.x + pix * p.x_cs;
    __nv_bfloat16* orow = p.out + pix * p.out_cs;
#pragma unroll 1
    for (int c0 = 0; c0 < kCV; c0 += 32) {
      uint32_t a[32];
      tmem_ld32(tmem_base + 384 + lane_off + c0, a);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xr + c0 + j * 8);
          const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + j * 8 + 2 * e;
            float v0 = __uint_as_float(a[j * 8 + 2 * e]) * inv + __uint_as_float(xw[e] << 16);
            float v1 = __uint_as_float(a[j * 8 + 2 * e + 1]) * inv + __uint_as_float(xw[e] & 0xffff0000u);
            v0 = lrelu(v0 * __ldg(p.scale + c) + __ldg(p.shift + c), p.slope);
            v1 = lrelu(v1 * __ldg(p.scale + c + 1) + __ldg(p.shift + c + 1), p.slope);
            const __nv_bfloat162 t = __floats2bfloat162_rn(v0, v1);
            w[e] = *reinterpret_cast<const uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(orow + c0 + j * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// bf16 token operands (K padded [N][kTP][kCKP], V^T padded [N][kCV][kTK]) live in a CALLER-owned workspace: a
// process-global buffer would be baked into captured CUDA graphs and freed under them when another engine grows it
size_t anab_tc_workspace_bytes(int N) {
  return static_cast<size_t>(N) * (static_cast<size_t>(kTP) * kCKP + static_cast<size_t>(kCV) * kTK) * 2 + 256;
}

bool anab_tc_supported(int ck, int cv, int T, int q_cs, int x_cs, int out_cs) {
  return ck <= kCKP && cv == kCV && T <= kTP && q_cs % 8 == 0 && x_cs % 8 == 0 && out_cs % 8 == 0;
}

int launch_anab_attention_tc(const void* q, int q_cs, const float* ktok, const float* vtok, const void* x, int x_cs,
                             const float* scale, const float* shift, float slope, void* out, int out_cs, int N, int HW,
                             int ck, int cv, int T, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < anab_tc_workspace_bytes(N)) {
    set_last_error("m3d_anab_attention: workspace of %zu bytes needed (m3d_anab_attention_workspace), got %zu",
                   anab_tc_workspace_bytes(N), workspace_bytes);
    return M3D_ERR_WORKSPACE;
  }
  struct { __nv_bfloat16 *kb, *vb; } g_tok;
  {
    uintptr_t a = (reinterpret_cast<uintptr_t>(workspace) + 127) & ~static_cast<uintptr_t>(127);
    g_tok.kb = reinterpret_cast<__nv_bfloat16*>(a);
    g_tok.vb = g_tok.kb + static_cast<size_t>(N) * kTP * kCKP;  // kTP * kCKP * 2 bytes is a multiple of 128
  }
  {
    const int work = kTP * kCKP > kCV * kTK ? kTP * kCKP : kCV * kTK;
    anab_tokens_bf16_kernel<<<dim3((work + 255) / 256, N), 256, 0, stream>>>(ktok, vtok, T, ck, cv, g_tok.kb, g_tok.vb);
    M3D_CUDA_OK(cudaGetLastError());
  }
  AnabParams p;
  memset(&p, 0, sizeof(p));
  // Q as (C, HW, N): box {64, 128, 1}; channels past ck and queries past HW are zero-filled by TMA
  {
    PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    M3D_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    M3D_REQUIRE(enc != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(ck), static_cast<cuuint64_t>(HW), static_cast<cuuint64_t>(N)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(q_cs) * 2, static_cast<cuuint64_t>(HW) * q_cs * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.tmap_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(q), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(anab q) failed: %d", static_cast<int>(r));
  }
  int rc = make_tmap_b3d(&p.tmap_k, g_tok.kb, static_cast<long>(N) * kTP, kCKP, 64, 176, 1);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_b3d(&p.tmap_v, g_tok.vb, static_cast<long>(N) * kCV, kTK, 64, 128, 1);
  if (rc != M3D_OK) return rc;
  p.x = static_cast<const __nv_bfloat16*>(x), p.x_cs = x_cs;
  p.scale = scale, p.shift = shift, p.slope = slope;
  p.out = static_cast<__nv_bfloat16*>(out), p.out_cs = out_cs;
  p.HW = HW, p.T = T, p.tiles_per_image = (HW + 127) / 128;
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(anab_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAnabSmem));
  M3D_ONCE_PER_DEVICE_END
  M3D_CUDA_OK(launch_pdl(anab_attention_tc_kernel, dim3(p.tiles_per_image * N), dim3(192), kAnabSmem, stream, p));
  return M3D_OK;
}

}  // namespace m3d
