// tba_segments.h -- run structure of a warp from the ballot of run heads (host/device, pure integer code; also compiled by
// tests/host_top2.cc for the CPU test of the "fast" segmented reduction).
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace tba {

// Last lane of the run (maximal group of contiguous lanes with one key) that `lane` belongs to, given the ballot `heads` of the
// lanes that start a run (bit 0 always set).
__host__ __device__ inline int run_last_lane(uint32_t heads, int lane) {
  const uint32_t above = lane >= 31 ? 0u : (heads & ~((2u << lane) - 1u));  // run heads strictly above `lane`
  if (above == 0u) return 31;
  int first = 0;
  uint32_t a = above;
  while (!(a & 1u)) { a >>= 1; ++first; }  // (device: compiles to a bit-scan; hosts have no __ffs)
  return first - 1;
}

}  // namespace tba
