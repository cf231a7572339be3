// Status codes and small host-side helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>

#include "../../include/m3dssd_b200.h"

namespace m3d {
void set_last_error(const char* fmt, ...);
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
