// Per-launch cost of persistent-kernel skeletons inside a CUDA graph (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I m3dssd_b200/csrc tools/microbench_launch.cu -o build/microbench_launch
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace m3d;

struct Big {
  char bytes[1216];
};

__global__ void k_trivial(int* out) {
  if (out != nullptr && threadIdx.x == 999) out[0] = 1;
}
__global__ void k_smem(int* out) {
  extern __shared__ uint8_t sm[];
  if (out != nullptr && threadIdx.x == 999) out[0] = sm[0];
}
__global__ void k_params(const __grid_constant__ Big b, int* out) {
  extern __shared__ uint8_t sm[];
  if (out != nullptr && threadIdx.x == 999) out[0] = sm[0] + b.bytes[5];
}
__global__ void __launch_bounds__(192, 1) k_tmem(const __grid_constant__ Big b, int* out) {
  extern __shared__ uint8_t sm[];
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 64);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = *slot;
  if (out != nullptr && threadIdx.x == 999) out[0] = b.bytes[5];
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(t);
  }
}

template <typename F>
float time_graph(F launch, cudaStream_t st, int n = 50) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < n; ++i) launch();
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0, st);
    cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best * 1e3f / n;
}

int main() {
  cudaStream_t st;
  cudaStreamCreate(&st);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_params, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Big b = {};
  printf("trivial 148x192            %6.2f us/launch\n", time_graph([&] { k_trivial<<<148, 192, 0, st>>>(nullptr); }, st));
  printf("trivial 1920x256           %6.2f us/launch\n", time_graph([&] { k_trivial<<<1920, 256, 0, st>>>(nullptr); }, st));
  printf("+200KB smem                %6.2f us/launch\n", time_graph([&] { k_smem<<<148, 192, smem, st>>>(nullptr); }, st));
  printf("+1.2KB params              %6.2f us/launch\n", time_graph([&] { k_params<<<148, 192, smem, st>>>(b, nullptr); }, st));
  printf("+tmem alloc/barriers       %6.2f us/launch\n", time_graph([&] { k_tmem<<<148, 192, smem, st>>>(b, nullptr); }, st));
  printf("alternating smem/no smem   %6.2f us/launch\n", time_graph([&] {
           k_tmem<<<148, 192, smem, st>>>(b, nullptr);
           k_trivial<<<1920, 256, 0, st>>>(nullptr);
         }, st, 25) / 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
