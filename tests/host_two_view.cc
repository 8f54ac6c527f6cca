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

#include <condition_variable>
#include <mutex>
#include <thread>

#include "../include/theia_ba_b200.h"
#include "../theiasfm_b200/csrc/tba_two_view.cuh"

// A 4-lane team emulated with host threads: the same strided loops and all-reduces WarpTeam performs with shuffles, so that
// the work decomposition of the warp kernel (ownership of points, placement of the reductions, uniform control flow) is
// exercised on the CPU.  Every reduction is a barrier + a fixed-order sum over the lanes' slots (all lanes get identical bits).
struct ThreadTeam {
  static constexpr int kLanes = 4;
  static thread_local int lane;
  static std::mutex mu; static std::condition_variable cv; static int waiting; static long generation; static double slot[kLanes];
  static void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const long g = generation;
    if (++waiting == kLanes) { waiting = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != g; });
  }
  static int rank() { return lane; }
  static int size() { return kLanes; }
  static double sum(double v) { slot[lane] = v; barrier(); double s = 0.0; for (int i = 0; i < kLanes; ++i) s += slot[i]; barrier(); return s; }
  static double max(double v) { slot[lane] = v; barrier(); double s = slot[0]; for (int i = 1; i < kLanes; ++i) s = std::fmax(s, slot[i]); barrier(); return s; }
  static bool all(bool v) { return sum(v ? 0.0 : 1.0) == 0.0; }
};
thread_local int ThreadTeam::lane = 0;
std::mutex ThreadTeam::mu; std::condition_variable ThreadTeam::cv; int ThreadTeam::waiting = 0; long ThreadTeam::generation = 0; double ThreadTeam::slot[ThreadTeam::kLanes];

extern "C" void host_two_view_ba_batch_team(tba_two_view_batch* b, unsigned char* termination, double* initial_cost, double* final_cost, int* iterations, int use_team);
extern "C" void host_two_view_ba_batch(tba_two_view_batch* b, unsigned char* termination, double* initial_cost, double* final_cost, int* iterations) {
  host_two_view_ba_batch_team(b, termination, initial_cost, final_cost, iterations, 0);
}
extern "C" void host_two_view_ba_batch_team(tba_two_view_batch* b, unsigned char* termination, double* initial_cost, double* final_cost, int* iterations, int use_team) {
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
    tba::PointLmResult r;
    if (use_team) {
      tba::PointLmResult rr[ThreadTeam::kLanes];
      std::vector<std::thread> th;
      for (int l = 0; l < ThreadTeam::kLanes; ++l) th.emplace_back([&, l] { ThreadTeam::lane = l; rr[l] = tba::two_view_lm<true, ThreadTeam>(P, o); });
      for (auto& t : th) t.join();
      r = rr[0];
      for (int l = 1; l < ThreadTeam::kLanes; ++l)   // every lane must have taken the same decisions
        if (rr[l].termination != r.termination || rr[l].iterations != r.iterations || rr[l].final_cost != r.final_cost) r.termination = 99;
    } else {
      r = tba::two_view_lm<true>(P, o);
    }
    if (b->inlier) tba::two_view_inliers<true>(P, b->final_max_reprojection_error_pixels * b->final_max_reprojection_error_pixels, b->inlier + (size_t)b0);
    termination[p] = (unsigned char)r.termination; initial_cost[p] = r.initial_cost; final_cost[p] = r.final_cost; iterations[p] = r.iterations;
  }
}
