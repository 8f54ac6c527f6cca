// tba_block_lm.h -- the per-block Levenberg-Marquardt of Ceres' inner iterations (CoordinateDescentMinimizer: one
// TrustRegionMinimizer with DEFAULT Solver::Options and DENSE_QR per parameter block, every other block constant) as a
// host-side state machine, so that ALL blocks of an independent set (all cameras, or all intrinsics groups) advance in
// lockstep while the observation passes run on the GPU:
//    NORMAL pass (k_block_normal): per block  H = J_b^T J_b, g = J_b^T r, cost   at the block's current value
//    COST   pass (k_block_cost):   per block  cost                               at the block's trial value
// Each block keeps its own trust-region radius, Jacobi scale, iteration count and termination, exactly as if it had been
// solved alone.  Pure C++ (no CUDA): also driven by the CPU test suite with host-evaluated passes
// (tests/host_inner.cc, tests/test_inner_iterations.py) against the oracle's recursive sub-solves.
// Semantics: DESIGN.md section 3 restricted to one dense block (same rules as tba_point_lm.cuh, dimension N <= 10, with a
// free-coordinate mask standing for the block's SubsetParameterization).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace tba {

constexpr int kBlkMaxN = 10;
constexpr int kBlkMaxH = kBlkMaxN * (kBlkMaxN + 1) / 2;

struct BlockLmOptions {  // ceres::Solver::Options defaults (what CoordinateDescentMinimizer::Solve runs with)
  int max_num_iterations = 50;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32, min_relative_decrease = 1e-3;
  double min_diag = 1e-6, max_diag = 1e32;
  int max_consecutive_invalid = 5;
};

enum BlockPhase { kBlkNeedNormal = 0, kBlkNeedCost = 1, kBlkDone = 2 };

struct BlockLm {
  int N = 0;
  bool free_[kBlkMaxN];
  double x[kBlkMaxN], cand[kBlkMaxN];
  double H[kBlkMaxH], g[kBlkMaxN], cost = 0.0;  // upper triangle row-major, masked, unscaled; at x
  double scale[kBlkMaxN];
  bool have_scale = false;
  double radius = 0.0, decrease = 2.0, xnorm = 0.0, mcc = 0.0;
  int iteration = 0, invalid = 0;
  bool last_successful = true;
  BlockPhase phase = kBlkDone;
  int termination = 1;  // 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE
  double initial_cost = -1.0, final_cost = -1.0;
};

inline int blk_tri(int N, int a, int b) { return a * N - a * (a - 1) / 2 + (b - a); }  // a <= b

inline void block_lm_init(BlockLm& B, int N, const bool* free_mask, const double* x0, const BlockLmOptions& o) {
  B = BlockLm();
  B.N = N;
  bool any = false;
  for (int j = 0; j < N; ++j) { B.free_[j] = free_mask[j]; B.x[j] = x0[j]; B.cand[j] = x0[j]; any |= free_mask[j]; }
  B.radius = o.initial_radius;
  B.phase = any ? kBlkNeedNormal : kBlkDone;
  B.termination = any ? 1 : 0;
}

// Solves the damped, Jacobi-scaled normal equations at the current (H, g, radius); fills cand, mcc.  false = invalid step.
inline bool block_lm_step(BlockLm& B, const BlockLmOptions& o) {
  const int N = B.N;
  double A[kBlkMaxN][kBlkMaxN], Hs[kBlkMaxN][kBlkMaxN], b[kBlkMaxN], y[kBlkMaxN], L[kBlkMaxN][kBlkMaxN];
  for (int a = 0; a < N; ++a)
    for (int c = 0; c < N; ++c) {
      const double h = a <= c ? B.H[blk_tri(N, a, c)] : B.H[blk_tri(N, c, a)];
      Hs[a][c] = (B.free_[a] && B.free_[c]) ? B.scale[a] * h * B.scale[c] : 0.0;
      A[a][c] = Hs[a][c];
    }
  for (int a = 0; a < N; ++a) {
    if (B.free_[a]) { b[a] = B.scale[a] * B.g[a]; A[a][a] += std::fmin(std::fmax(Hs[a][a], o.min_diag), o.max_diag) / B.radius; }
    else { b[a] = 0.0; A[a][a] = 1.0; }
  }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j <= i; ++j) {
      double acc = A[i][j];
      for (int k = 0; k < j; ++k) acc -= L[i][k] * L[j][k];
      if (i == j) { if (!(acc > 0.0)) return false; L[i][i] = std::sqrt(acc); }
      else L[i][j] = acc / L[j][j];
    }
  for (int i = 0; i < N; ++i) { double acc = b[i]; for (int k = 0; k < i; ++k) acc -= L[i][k] * y[k]; y[i] = acc / L[i][i]; }
  for (int i = N - 1; i >= 0; --i) { double acc = y[i]; for (int k = i + 1; k < N; ++k) acc -= L[k][i] * y[k]; y[i] = acc / L[i][i]; }
  double sg = 0.0, shs = 0.0;
  for (int a = 0; a < N; ++a) {
    const double sa = -y[a];
    sg += sa * b[a];
    for (int c = 0; c < N; ++c) shs += sa * Hs[a][c] * (-y[c]);
    const double d = B.free_[a] ? sa * B.scale[a] : 0.0;
    if (!std::isfinite(d)) return false;
    B.cand[a] = B.x[a] + d;
  }
  B.mcc = -(sg + 0.5 * shs);
  return B.mcc > 0.0;
}

// FinalizeIterationAndCheckIfMinimizerCanContinue + ComputeTrustRegionStep (retrying invalid steps, which need no new
// evaluation): leaves the block in kBlkNeedCost with a trial value, or kBlkDone.
inline void block_lm_advance(BlockLm& B, const BlockLmOptions& o) {
  for (;;) {
    if (B.iteration >= o.max_num_iterations) { B.termination = 1; B.phase = kBlkDone; return; }
    if (B.last_successful) {
      double gmax = 0.0;
      for (int j = 0; j < B.N; ++j) if (B.free_[j]) gmax = std::fmax(gmax, std::fabs(B.g[j]));
      if (gmax <= o.gradient_tolerance) { B.termination = 0; B.phase = kBlkDone; return; }
    }
    if (B.radius <= o.min_radius) { B.termination = 0; B.phase = kBlkDone; return; }
    ++B.iteration;
    if (block_lm_step(B, o)) { B.invalid = 0; B.phase = kBlkNeedCost; return; }
    if (++B.invalid >= o.max_consecutive_invalid) { B.termination = 2; B.phase = kBlkDone; return; }  // HandleInvalidStep
    B.radius /= B.decrease; B.decrease *= 2.0;
    B.last_successful = false;
  }
}

// Result of a NORMAL pass at B.x (ok = every residual of the block evaluated).
inline void block_lm_on_normal(BlockLm& B, bool ok, const double* H, const double* g, double cost, const BlockLmOptions& o) {
  const int N = B.N, NH = N * (N + 1) / 2;
  if (!ok) { B.termination = 2; B.phase = kBlkDone; return; }  // "Residual and Jacobian evaluation failed."
  for (int a = 0; a < N; ++a) for (int c = a; c < N; ++c) B.H[blk_tri(N, a, c)] = (B.free_[a] && B.free_[c]) ? H[blk_tri(N, a, c)] : 0.0;
  (void)NH;
  for (int j = 0; j < N; ++j) B.g[j] = B.free_[j] ? g[j] : 0.0;
  B.cost = cost;
  B.final_cost = cost;
  if (!B.have_scale) {
    B.initial_cost = cost;
    for (int j = 0; j < N; ++j) B.scale[j] = 1.0 / (1.0 + std::sqrt(B.H[blk_tri(N, j, j)]));
    B.have_scale = true;
  }
  double s = 0.0;
  for (int j = 0; j < N; ++j) s += B.x[j] * B.x[j];
  B.xnorm = std::sqrt(s);
  block_lm_advance(B, o);
}

// Result of a COST pass at B.cand.
inline void block_lm_on_cost(BlockLm& B, bool ok, double cand_cost, const BlockLmOptions& o) {
  if (!ok) cand_cost = 1.7976931348623157e308;
  double dn = 0.0;
  for (int j = 0; j < B.N; ++j) { const double d = B.cand[j] - B.x[j]; dn += d * d; }
  if (std::sqrt(dn) <= o.parameter_tolerance * (B.xnorm + o.parameter_tolerance)) { B.termination = 0; B.phase = kBlkDone; return; }
  const double cost_change = B.cost - cand_cost;
  if (std::fabs(cost_change) <= o.function_tolerance * B.cost) { B.termination = 0; B.phase = kBlkDone; return; }
  const double rho = cost_change / B.mcc;
  if (rho > o.min_relative_decrease) {  // HandleSuccessfulStep: a NORMAL pass at the new x follows
    for (int j = 0; j < B.N; ++j) B.x[j] = B.cand[j];
    const double t = 2.0 * rho - 1.0;
    B.radius = std::fmin(o.max_radius, B.radius / std::fmax(1.0 / 3.0, 1.0 - t * t * t));
    B.decrease = 2.0;
    B.last_successful = true;
    B.phase = kBlkNeedNormal;
    return;
  }
  B.radius /= B.decrease; B.decrease *= 2.0;  // HandleUnsuccessfulStep
  B.last_successful = false;
  block_lm_advance(B, o);
}

// Lockstep driver: all blocks advance together; exec(pass, active, vals, sum) runs one observation pass (pass 0: NORMAL at
// vals = current values, pass 1: COST at vals = trial values of the active blocks) and returns, per block b,
// sum[b * NA ..] = { H (upper triangle of the ND x ND block), g (ND), cost, failed evaluations }.  dims[b] <= ND is the block's
// true dimension.  Returns exec's first non-zero code.  On return B[b].x holds every block's result.
template <class Exec>
int block_lm_run_lockstep(std::vector<BlockLm>& B, const std::vector<int>& dims, int ND, int NA, const BlockLmOptions& lo, Exec&& exec) {
  const int nb = (int)B.size(), NHD = ND * (ND + 1) / 2;
  std::vector<double> vals((size_t)nb * ND, 0.0), sum((size_t)nb * NA, 0.0);
  std::vector<uint8_t> active((size_t)nb, 0);
  for (int round = 0; round < 4 * lo.max_num_iterations + 8; ++round) {
    bool any = false;
    for (int pass = 0; pass < 2; ++pass) {
      const BlockPhase want = pass == 0 ? kBlkNeedNormal : kBlkNeedCost;
      int n_act = 0;
      for (int b = 0; b < nb; ++b) {
        active[b] = B[b].phase == want ? 1 : 0;
        n_act += active[b];
        const double* src = (pass == 1 && active[b]) ? B[b].cand : B[b].x;
        for (int j = 0; j < dims[b]; ++j) vals[(size_t)b * ND + j] = src[j];
      }
      if (n_act == 0) continue;
      any = true;
      const int rc = exec(pass, active, vals, sum);
      if (rc) return rc;
      for (int b = 0; b < nb; ++b) {
        if (!active[b]) continue;
        const double* a = &sum[(size_t)b * NA];
        const bool ok = a[NHD + ND + 1] == 0.0;
        if (pass == 1) { block_lm_on_cost(B[b], ok, a[NHD + ND], lo); continue; }
        const int N = dims[b];
        double H[kBlkMaxH], g[kBlkMaxN];
        for (int p = 0; p < N; ++p) {
          g[p] = a[NHD + p];
          for (int q = p; q < N; ++q) H[blk_tri(N, p, q)] = a[blk_tri(ND, p, q)];
        }
        block_lm_on_normal(B[b], ok, H, g, a[NHD + ND], lo);
      }
    }
    if (!any) break;
  }
  return 0;
}

}  // namespace tba
