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

// Cost of the generic->async proxy hand-off that every CUDA-core A producer pays per slab / k-block:
// st.shared.v4 + fence.proxy.async (+ mbarrier arrive), with 1 / 4 / 16 warps running it concurrently.
__global__ void __launch_bounds__(512, 1) bench_fence(long long* out, int mode) {
  __shared__ __align__(16) uint4 buf[512 * 2];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1 << 19);
    fence_barrier_init();
  }
  __syncthreads();
  uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    v.x += i;
    buf[threadIdx.x * 2 + (i & 1)] = v;
    if (mode >= 1) fence_proxy_async_smem();
    if (mode >= 2) {
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&bar);
    }
    if (mode == 3) {  // 64 independent FMAs after the hand-off (does the fence overlap with math?)
      float a = __uint_as_float(v.y);
#pragma unroll
      for (int k = 0; k < 64; ++k) a = fmaf(a, 1.0001f, 0.5f);
      v.y = __float_as_uint(a);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  const uint4 r = buf[(threadIdx.x * 2 + 3) % (blockDim.x * 2)];  // keep the stores observable
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0 + (v.y + r.x == 77);
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
  const char* modes[] = {"st.shared.v4 only", "+ fence.proxy.async", "+ syncwarp + arrive", "+ 64 dependent FMAs"};
  for (int warps = 1; warps <= 16; warps *= 4)
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) bench_fence<<<1, 32 * warps>>>(d, mode);  // first launch warms the i-cache
      if (cudaDeviceSynchronize() != cudaSuccess) return 1;
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%2d warps  %-24s %8.1f clk/iter\n", warps, modes[mode], double(h[0]) / ITERS);
    }
  return 0;
}
