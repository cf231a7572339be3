// Producer/consumer mbarrier ring between two warps (the conv pipelines' skeleton), clocks per iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I m3dssd_b200/csrc tools/microbench_ring.cu -o build/microbench_ring
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace m3d;

constexpr int ITERS = 512;

// mode 0: consumer frees the slot with mbarrier.arrive; 1: with tcgen05.commit; 2: mode 0 + bounded spin wait (no clock)
template <int STAGES, int MODE>
__global__ void __launch_bounds__(96, 1) ring(long long* out) {
  __shared__ __align__(8) uint64_t full[8], empty[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<32>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  auto wait = [&](uint64_t* b, uint32_t par) {
    if (MODE == 2) {
      while (!mbar_try_wait(b, par)) {
      }
    } else {
      mbar_wait(b, par);
    }
  };
  long long t0 = clock64();
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < ITERS; ++i) {
      wait(&empty[stage], phase ^ 1);
      if (elect_one()) mbar_arrive_expect_tx(&full[stage], 0);
      __syncwarp();
      if (++stage == STAGES) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < ITERS; ++i) {
      wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        if (MODE == 1) umma_commit(&empty[stage]);
        else mbar_arrive(&empty[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) stage = 0, phase ^= 1;
    }
  }
  long long t1 = clock64();
  if (lane == 0 && warp < 2) out[warp] = t1 - t0;
  __syncthreads();
  if (warp == 2) tmem_dealloc<32>(slot);
}

template <int S, int M>
void run(const char* name, long long* d) {
  ring<S, M><<<1, 96>>>(d);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-40s producer %7.1f  consumer %7.1f clk/iter\n", name, double(h[0]) / ITERS, double(h[1]) / ITERS);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<1, 0>("1 stage, arrive", d);
  run<2, 0>("2 stages, arrive", d);
  run<4, 0>("4 stages, arrive", d);
  run<6, 0>("6 stages, arrive", d);
  run<6, 1>("6 stages, tcgen05.commit", d);
  run<1, 1>("1 stage, tcgen05.commit", d);
  run<6, 2>("6 stages, arrive, plain spin", d);
  run<1, 2>("1 stage, arrive, plain spin", d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
