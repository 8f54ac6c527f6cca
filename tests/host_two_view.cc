// host_two_view.cc -- runs the PRODUCT's per-pair two-view LM (theiasfm_b200/csrc/tba_two_view.cuh, the body of
// k_two_view_ba) on the host over the batch layout of tba_two_view_ba_batch, for the CPU test suite (tests/test_two_view.py).
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
#include <vector>
using std::atan; using std::atan2; using std::fabs; using std::fmax; using std::fmin; using std::isfinite; using std::sqrt; using std::tan;

#include "../include/theia_ba_b200.h"
#include "../theiasfm_b200/csrc/tba_two_view.cuh"

extern "C" void host_two_view_ba_batch(tba_two_view_batch* b, unsigned char* termination, double* initial_cost, double* final_cost, int* iterations) {
  tba::PointLmOptions o;
  o.loss_type = 0; o.loss_width = 1.0; o.max_num_iterations = 200;
  o.function_tolerance = 1e-6; o.gradient_tolerance = 1e-10; o.parameter_tolerance = 1e-8;
  o.initial_radius = 1e4; o.max_radius = 1e16; o.min_radius = 1e-32; o.min_relative_decrease = 1e-3; o.min_diag = 1e-6; o.max_diag = 1e32;
  o.jacobi_scaling = 1; o.max_consecutive_invalid = 5;
  const long long nc = b->pair_off[b->n_pairs];
  std::vector<double> sp((size_t)nc * 4), ptc((size_t)nc * 4);
  for (int p = 0; p < b->n_pairs; ++p) {
    const long long b0 = b->pair_off[p];
    tba::TwoViewPair P;
    P.ext1 = b->ext1 + (size_t)p * 6; P.ext2 = b->ext2 + (size_t)p * 6; P.k1 = b->intr1 + (size_t)p * 10; P.k2 = b->intr2 + (size_t)p * 10;
    P.model1 = b->model1[p]; P.model2 = b->model2[p]; P.free_f1 = b->constant_intrinsics1[p] ? 0 : 1; P.free_f2 = b->constant_intrinsics2[p] ? 0 : 1;
    P.n = (int)(b->pair_off[p + 1] - b0);
    P.pt = b->points + (size_t)b0 * 4; P.xy1 = b->xy1 + (size_t)b0 * 2; P.xy2 = b->xy2 + (size_t)b0 * 2; P.sp = sp.data() + (size_t)b0 * 4;
    P.pt_c = ptc.data() + (size_t)b0 * 4;
    const tba::PointLmResult r = tba::two_view_lm<true>(P, o);
    termination[p] = (unsigned char)r.termination; initial_cost[p] = r.initial_cost; final_cost[p] = r.final_cost; iterations[p] = r.iterations;
  }
}
