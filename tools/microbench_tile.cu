// Skeleton of a persistent tcgen05 conv kernel: producer warp, MMA warp, NW epilogue warps, S-stage operand ring,
// 2 accumulator stages; no loads, no MMAs.  Clocks per tile for KPT stages per tile (development aid).
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace m3d;

constexpr int TILES = 200;

// VAR: 0 baseline; 1 epilogue arrives per thread (32*NW arrivals); 2 no epilogue hand-shake at all
template <int STAGES, int KPT, int NW, int VAR>
__global__ void __launch_bounds__(64 + 32 * NW, 1) skel(long long* out) {
  __shared__ __align__(8) uint64_t full[8], empty[8], tfull[2], tempty[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&tfull[i], 1), mbar_init(&tempty[i], VAR == 1 ? 32 * NW : NW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = clock64();
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < TILES; ++t)
      for (int k = 0; k < KPT; ++k) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) mbar_arrive_expect_tx(&full[stage], 0);
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < TILES; ++t) {
      const int as = t & 1;
      if (VAR != 2) mbar_wait(&tempty[as], ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int k = 0; k < KPT; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) umma_commit(&empty[stage]);
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      if (VAR != 2) {
        if (elect_one()) umma_commit(&tfull[as]);
        __syncwarp();
      }
    }
  } else if (VAR != 2) {
    for (int t = 0; t < TILES; ++t) {
      const int as = t & 1;
      mbar_wait(&tfull[as], (t >> 1) & 1);
      tc_fence_after();
      tc_fence_before();
      if (VAR == 1) {
        mbar_arrive(&tempty[as]);
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
      }
    }
  }
  long long t1 = clock64();
  if (lane == 0 && warp < 3) out[warp] = t1 - t0;
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(slot);
}

template <int S, int K, int NW, int V>
void run(const char* name, long long* d) {
  skel<S, K, NW, V><<<1, 64 + 32 * NW>>>(d);
  cudaDeviceSynchronize();
  long long h[3];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s producer %7.1f  mma %7.1f  epilogue %7.1f clk/tile\n", name, double(h[0]) / TILES, double(h[1]) / TILES,
         double(h[2]) / TILES);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<6, 3, 8, 0>("6 stages, 3/tile, 8 epi warps", d);
  run<6, 3, 8, 1>("  per-thread tempty arrivals", d);
  run<6, 3, 8, 2>("  no accumulator hand-shake", d);
  run<6, 3, 4, 0>("6 stages, 3/tile, 4 epi warps", d);
  run<6, 1, 8, 0>("6 stages, 1/tile, 8 epi warps", d);
  run<6, 9, 8, 0>("6 stages, 9/tile, 8 epi warps", d);
  run<4, 18, 4, 0>("4 stages, 18/tile, 4 epi warps", d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
