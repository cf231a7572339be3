// Single-warp latency of the synchronisation primitives the conv pipelines are built from (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I m3dssd_b200/csrc tools/microbench.cu -o build/microbench
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace m3d;

constexpr int ITERS = 256;

__global__ void __launch_bounds__(64, 1) bench(long long* out) {
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<32>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp != 0) return;
  long long t0, t1;
  int k = 0;

  // 0: empty loop with clock reads
  t0 = clock64();
  for (int i = 0; i < ITERS; ++i) asm volatile("" ::: "memory");
  t1 = clock64();
  if (lane == 0) out[k] = t1 - t0;
  ++k;

  // 1: all lanes: arrive (count 1 -> phase completes each time) + try_wait on the completed phase
  {
    uint32_t ph = 0;
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      if (elect_one()) mbar_arrive(&bars[0]);
      __syncwarp();
      mbar_wait(&bars[0], ph);
      ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0;
    ++k;
  }
  // 2: elect + syncwarp only
  {
    int acc = 0;
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      if (elect_one()) acc += i;
      __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0 + (acc == 12345);
    ++k;
  }
  // 3: arrive.expect_tx(0) + wait
  {
    uint32_t ph = 0;
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      if (elect_one()) mbar_arrive_expect_tx(&bars[1], 0);
      __syncwarp();
      mbar_wait(&bars[1], ph);
      ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0;
    ++k;
  }
  // 4: tcgen05.commit (no MMAs outstanding) + wait
  {
    uint32_t ph = 0;
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      if (elect_one()) umma_commit(&bars[2]);
      __syncwarp();
      mbar_wait(&bars[2], ph);
      ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0;
    ++k;
  }
  // 5: try_wait on an already completed phase only (no arrive): wait for the previous phase parity
  {
    if (elect_one()) mbar_arrive(&bars[3]);  // completes phase 0
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) mbar_wait(&bars[3], 0);
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0;
    ++k;
  }
  // 6: arrive only, pipelined (no wait): barrier with count 1 flips phases freely
  {
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      if (elect_one()) mbar_arrive(&bars[4]);
      __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0;
    ++k;
  }
  // 7: tcgen05 fence after + before
  {
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      tc_fence_after();
      tc_fence_before();
    }
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0;
    ++k;
  }
  // 8: commit only, pipelined
  {
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      if (elect_one()) umma_commit(&bars[5]);
      __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) out[k] = t1 - t0;
    ++k;
  }
  // 9: lane-0-only (divergent) arrive + wait, the old code style
  if (lane == 0) {
    uint32_t ph = 0;
    // drain whatever phase bars[6] is in: fresh barrier, phase 0
    t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
      mbar_arrive(&bars[6]);
      mbar_wait(&bars[6], ph);
      ph ^= 1;
    }
    t1 = clock64();
    out[k] = t1 - t0;
  }
  ++k;
  __syncwarp();
  __nanosleep(2000);
  tmem_dealloc<32>(tmem_slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  cudaMemset(d, 0, 64 * sizeof(long long));
  bench<<<1, 64>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  long long h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"empty loop", "arrive+wait", "elect+syncwarp", "arrive.expect_tx(0)+wait", "tcgen05.commit+wait",
                         "wait on completed phase", "arrive only", "tc fences", "commit only", "lane0-only arrive+wait"};
  for (int i = 0; i < 10; ++i) printf("%-28s %8.1f clk/iter\n", names[i], double(h[i]) / ITERS);
  return 0;
}
