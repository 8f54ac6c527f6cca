// tbm_decide.cuh -- MatchImagePair's decisions on the device (brute_force_feature_matcher.cc:63-116 + IntersectMatches,
// feature_matcher_utils.cc:48-71), one CTA per image pair, from the nearest / second-nearest results of both directions:
// ratio test (:78-81, double arithmetic on float distances), "not enough matches" early exit after the forward pass (:84-86),
// symmetric filtering, final count test (:116).  The match list of a pair is written in ascending feature1_ind order, exactly
// the list tbm_debug_postprocess (the host restatement the CPU tests pin against the oracle) produces; only the kept matches
// travel back to the host instead of three arrays per query.  Plain CUDA: also compiled by the SIMT emulation build.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/theia_matcher_b200.h"

namespace tbm {

struct PairSeg {
  long long f0, r0;  // first forward / reverse query of the pair in the chunk's result arrays (r0 unused when n_rev == 0)
  int n1, n2;        // descriptors of image 1 / image 2
  int n_rev;         // length of the reverse result arrays: n2 with symmetric matching, else 0
};

struct QuerySeg {  // one direction of one pair: queries [out0, out0 + nq) of the chunk are rows q_row0.. against rows [b_row0, b_row0 + b_rows)
  long long out0;
  int nq, q_row0, b_row0, b_rows;
};

struct DecideOptions {
  int symmetric, use_ratio, min_matches;
  float ratio_sq;  // lowes_ratio * lowes_ratio rounded to float (FeatureMatcherOptions::lowes_ratio is a float, :58-59)
};

constexpr int kDecideThreads = 256;

__device__ __forceinline__ bool passes_ratio(const DecideOptions& o, float best, float second, bool second_valid) {
  if (!o.use_ratio || !second_valid) return true;
  return (double)best < (double)o.ratio_sq * (double)second;
}

// per-query scatter of the segment description (the exact re-evaluation kernel reads one (row, range) triple per query)
__global__ void k_expand_segments(const QuerySeg* __restrict__ segs, int n_segs, int* __restrict__ q_row, int* __restrict__ b_row0,
                                  int* __restrict__ b_rows) {
  for (int s = blockIdx.x; s < n_segs; s += gridDim.x) {
    const QuerySeg g = segs[s];
    for (int i = threadIdx.x; i < g.nq; i += blockDim.x) {
      q_row[g.out0 + i] = g.q_row0 + i;
      b_row0[g.out0 + i] = g.b_row0;
      b_rows[g.out0 + i] = g.b_rows;
    }
  }
}

__global__ void __launch_bounds__(kDecideThreads) k_pair_decide(const PairSeg* __restrict__ segs, int n_pairs, const int* __restrict__ best_j,
                                                                const float* __restrict__ best_d, const float* __restrict__ second_d,
                                                                DecideOptions o, tbm_match* __restrict__ staged /*pair p: [f0, f0 + n1)*/,
                                                                int* __restrict__ count, uint8_t* __restrict__ ok) {
  __shared__ int s_cnt[kDecideThreads];
  __shared__ int s_warp[kDecideThreads / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
    const PairSeg g = segs[p];
    const bool f_second_valid = g.n2 >= 2, r_second_valid = g.n1 >= 2;
    auto forward = [&](int i) {
      return best_j[g.f0 + i] >= 0 && passes_ratio(o, best_d[g.f0 + i], second_d[g.f0 + i], f_second_valid);
    };
    // ---- forward matches (:63-82): how many?
    int c = 0;
    for (int i = tid; i < g.n1; i += kDecideThreads) c += forward(i) ? 1 : 0;
    s_cnt[tid] = c;
    __syncthreads();
    for (int s = kDecideThreads / 2; s > 0; s >>= 1) {
      if (tid < s) s_cnt[tid] += s_cnt[tid + s];
      __syncthreads();
    }
    const int n_fwd = s_cnt[0];
    __syncthreads();
    // :84-86: too few forward matches -> the pair fails and its list is the forward list; otherwise IntersectMatches
    const bool intersect = n_fwd >= o.min_matches && o.symmetric != 0;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < g.n1; base += kDecideThreads) {
      const int i = base + tid;
      bool keep = false;
      int j = -1;
      float d = 0.0f;
      if (i < g.n1 && forward(i)) {
        j = best_j[g.f0 + i];
        d = best_d[g.f0 + i];
        keep = true;
        if (intersect)
          keep = j < g.n_rev && best_j[g.r0 + j] == i && passes_ratio(o, best_d[g.r0 + j], second_d[g.r0 + j], r_second_valid);
      }
      const unsigned b = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_warp[warp] = __popc(b);
      __syncthreads();
      int off = s_base;
      for (int w = 0; w < warp; ++w) off += s_warp[w];
      if (keep) {
        tbm_match m;
        m.feature1_ind = i; m.feature2_ind = j; m.distance = d;
        staged[g.f0 + off + __popc(b & ((1u << lane) - 1u))] = m;
      }
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int w = 0; w < kDecideThreads / 32; ++w) t += s_warp[w];
        s_base += t;
      }
      __syncthreads();
    }
    if (tid == 0) {
      count[p] = s_base;
      ok[p] = (uint8_t)(n_fwd >= o.min_matches && s_base >= o.min_matches);  // :116
    }
    __syncthreads();
  }
}

// the kept matches of every pair, packed back to back in pair order: dst[dst_off[p] + k] = staged[f0(p) + k]
__global__ void k_gather_matches(const PairSeg* __restrict__ segs, int n_pairs, const int* __restrict__ count, const long long* __restrict__ dst_off,
                                 const tbm_match* __restrict__ staged, tbm_match* __restrict__ dst) {
  for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
    const int* src = reinterpret_cast<const int*>(staged + segs[p].f0);
    int* out = reinterpret_cast<int*>(dst + dst_off[p]);
    const int n = count[p] * 3;  // tbm_match = {int32, int32, float}
    for (int k = threadIdx.x; k < n; k += blockDim.x) out[k] = src[k];
  }
}

}  // namespace tbm
