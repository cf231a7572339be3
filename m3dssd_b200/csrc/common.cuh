// Status codes and small host-side helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>

#include "../../include/m3dssd_b200.h"

namespace m3d {
void set_last_error(const char* fmt, ...);
// name of the kernel instantiation m3d_conv2d_nhwc dispatched to on this thread (bench.py labels its roofline with it)
void set_last_kernel(const char* fmt, ...);
// SMs the persistent kernels launched from now on may occupy (m3d_set_sm_limit; default: all of the device)
int persistent_sms();
// SMs left out by m3d_set_sm_limit (0 when no limit is set): their CTAs may each strand the sibling SM of a TPC,
// which matters to kernels launched as CTA pairs
int reserved_sms();
}

#define M3D_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::m3d::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return M3D_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define M3D_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) {                            \
      ::m3d::set_last_error(__VA_ARGS__);     \
      return M3D_ERR_INVALID;                 \
    }                                         \
  } while (0)

// Function attributes (cudaFuncSetAttribute) are per DEVICE: remember, per call site and thus per kernel
// instantiation, on which devices the enclosed statements have already run.  Usage:
//   M3D_ONCE_PER_DEVICE_BEGIN  M3D_CUDA_OK(cudaFuncSetAttribute(...));  M3D_ONCE_PER_DEVICE_END
#define M3D_ONCE_PER_DEVICE_BEGIN                                              \
  {                                                                            \
    static unsigned long long _m3d_done[4] = {0, 0, 0, 0};                     \
    int _m3d_dev = 0;                                                          \
    M3D_CUDA_OK(cudaGetDevice(&_m3d_dev));                                     \
    _m3d_dev &= 255;                                                           \
    if (!((_m3d_done[_m3d_dev >> 6] >> (_m3d_dev & 63)) & 1ull)) {
#define M3D_ONCE_PER_DEVICE_END                                                \
      _m3d_done[_m3d_dev >> 6] |= 1ull << (_m3d_dev & 63);                     \
    }                                                                          \
  }

// Kernel launch with programmatic dependent launch (PDL): the grid may be scheduled while its predecessor
// in the stream drains (SM by SM), runs its prologue (barrier init, TMEM allocation, descriptor prefetch)
// and blocks in griddepcontrol.wait until the predecessor has completed and flushed.  Only kernels that
// execute m3d::grid_dep_sync() before their first global access may be launched this way.
// M3D_PDL=0 falls back to ordinary stream serialisation.
#ifdef __CUDACC__
#include <cstdlib>
#include <utility>
namespace m3d {
// m3d_set_pdl(0 / 1): the engine switches programmatic dependent launch OFF for the launches it issues while a branch
// of its plan runs on a second stream (api_conv.cu).  Two persistent grids then share the device and one of them has
// unscheduled CTAs; a PDL successor whose CTAs sit blocked in griddepcontrol.wait on SMs those CTAs need can starve
// them (observed as a hang with 148-CTA grids beside the 8 CTAs of the detection tail).  Without PDL nothing ever
// occupies an SM while blocked, so progress is guaranteed whatever the hardware's CTA scheduling order.
bool pdl_switch();
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("M3D_PDL");
    on = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return on == 1 && pdl_switch();
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
// Device side: let the successor start launching, then wait for the predecessor's results.
__device__ __forceinline__ void grid_dep_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
}  // namespace m3d
#endif
