// Does tcgen05.mma accept a 128B-swizzled K-major A operand whose 8-row groups start at 128-byte (not 1024-byte)
// aligned addresses with a stride-between-groups (SBO) that is not a multiple of 1024 bytes?  That is what a 3x3 conv
// needs to read all nine taps out of ONE (TH+2) x (TW+2) pixel window with TW = 8: tile row g of tap (r, s) starts at
// window row (g + r) * (TW + 2) + s.  The window is written the way TMA writes it (16-byte chunk index XOR the
// absolute 128-byte row index mod 8).  Development aid:  nvcc -arch=sm_100a -I m3dssd_b200/csrc tools/microbench_desc.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace m3d;

constexpr int WROWS = 18 * 10;  // window pixels
constexpr int N = 64;

__global__ void __launch_bounds__(128, 1) k(const float* a_val /*[WROWS][64]*/, const float* b_val /*[N][64]*/, float* d_out,
                                            int r, int s, int sbo_bytes, int use_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;               // WROWS x 128 B (24 KB region)
  uint8_t* sb = smem + 24 * 1024;   // N x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < WROWS * 64; i += 128) {
    const int row = i / 64, c = i % 64;
    const uint32_t off = row * 128 + (((c / 8) ^ (row & 7)) * 16) + (c % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16(a_val[i]);
  }
  for (int i = threadIdx.x; i < N * 64; i += 128) {
    const int row = i / 64, c = i % 64;
    const uint32_t off = row * 128 + (((c / 8) ^ (row & 7)) * 16) + (c % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sb + off) = __float2bfloat16(b_val[i]);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t a_addr = smem_u32(sa) + (r * 10 + s) * 128;
    uint64_t da = static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(1) << 16) |
                  (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
                  (static_cast<uint64_t>(2) << 61);
    if (use_base_offset) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t db = umma_smem_desc<128>(smem_u32(sb));
    for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, da + 2 * kk, db + 2 * kk, umma_idesc_bf16(N), kk != 0);
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[64];
  tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), *reinterpret_cast<uint32_t(*)[32]>(v));
  tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 32, *reinterpret_cast<uint32_t(*)[32]>(v + 32));
  tmem_ld_wait();
  for (int j = 0; j < 64; ++j) d_out[(warp * 32 + lane) * 64 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  float *a, *b, *d;
  cudaMallocManaged(&a, WROWS * 64 * 4);
  cudaMallocManaged(&b, N * 64 * 4);
  cudaMallocManaged(&d, 128 * 64 * 4);
  srand(1);
  for (int i = 0; i < WROWS * 64; ++i) a[i] = static_cast<float>(rand() % 7 - 3);
  for (int i = 0; i < N * 64; ++i) b[i] = static_cast<float>(rand() % 5 - 2);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const int cases[][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}, {2, 2}, {0, 2}};
  for (int ubo = 0; ubo < 2; ++ubo)
    for (auto& c : cases) {
      k<<<1, 128, 40 * 1024>>>(a, b, d, c[0], c[1], 1280, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("tap (%d,%d) base_offset=%d: CUDA error %s\n", c[0], c[1], ubo, cudaGetErrorString(e));
        return 1;
      }
      int bad = 0;
      for (int m = 0; m < 128; ++m) {
        const int wrow = (m / 8 + c[0]) * 10 + c[1] + m % 8;
        for (int n = 0; n < N; ++n) {
          float ref = 0.f;
          for (int q = 0; q < 64; ++q) ref += a[wrow * 64 + q] * b[n * 64 + q];
          if (ref != d[m * 64 + n]) ++bad;
        }
      }
      printf("tap (%d,%d) SBO=1280 base_offset_field=%d: %d / %d wrong\n", c[0], c[1], ubo, bad, 128 * N);
    }
  return 0;
}
