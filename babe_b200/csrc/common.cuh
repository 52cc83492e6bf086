// Shared host/device helpers for the babe_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "babe_b200.h"

#ifndef BABE_HD
#define BABE_HD __host__ __device__ __forceinline__
#endif

namespace babe {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

#define BABE_REQUIRE(cond, code, ...)            \
  do {                                           \
    if (!(cond)) {                               \
      ::babe::set_error(__VA_ARGS__);            \
      return (code);                             \
    }                                            \
  } while (0)

// Barrier over the NT threads of one frame group (NT multiple of 32).
// One warp needs no hardware barrier.
template <int NT>
__device__ __forceinline__ void group_sync(int bar_id) {
  if (NT == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace babe
