// tba_point_lm.cuh -- Levenberg-Marquardt on ONE homogeneous point with every camera held constant: the problem
// theia::BundleAdjustTrack (src/theia/sfm/bundle_adjustment/bundle_adjustment.cc:96-107; called per track from
// estimate_track.cc:241) hands to Ceres with DENSE_QR, and the "points" stage of Ceres' inner iterations.
// Host/device: k_adjust_tracks runs one instance per thread (thousands of independent tiny LM
// problems); the CPU test suite runs the very same body over the packed layout against the oracle
// (tests/host_point_lm.cc, tests/test_point_lm.py).
//
// Solver semantics = DESIGN.md section 3 restricted to one 4-dimensional block (Ceres trust-region LM: Jacobi scaling
// fixed at iteration 0, diagonal clamp [1e-6, 1e32], radius rules, parameter / function / gradient tolerances,
// failed evaluation = rejected step, at most 5 consecutive invalid steps); the 4x4 damped normal equations are solved by
// Cholesky (DENSE_QR solves the same least-squares problem).
#pragma once
#include <cstdint>

#include "tba_camera_models.cuh"
#include "tba_filter.cuh"  // FilterView: the read-only view of the packed problem

namespace tba {

struct PointLmOptions {
  int loss_type; double loss_width;
  int max_num_iterations;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_radius, max_radius, min_radius, min_relative_decrease, min_diag, max_diag;
  int jacobi_scaling, max_consecutive_invalid;
};

struct PointLmResult {
  double initial_cost, final_cost;
  int iterations;       // LM iterations performed (Ceres' iteration index of the last one)
  int termination;      // TBA_CONVERGENCE 0 / TBA_NO_CONVERGENCE 1 / TBA_FAILURE 2
};

// 4x4 SPD solve A y = b through Cholesky (A upper triangle: (0,0)=0 (0,1)=1 (0,2)=2 (0,3)=3 (1,1)=4 (1,2)=5 (1,3)=6 (2,2)=7 (2,3)=8 (3,3)=9)
__host__ __device__ inline bool spd4_solve(const double* A, const double* b, double* y) {
  const double a00 = A[0], a01 = A[1], a02 = A[2], a03 = A[3], a11 = A[4], a12 = A[5], a13 = A[6], a22 = A[7], a23 = A[8], a33 = A[9];
  if (!(a00 > 0.0)) return false;
  const double l00 = sqrt(a00);
  const double l10 = a01 / l00, l20 = a02 / l00, l30 = a03 / l00;
  const double d1 = a11 - l10 * l10;
  if (!(d1 > 0.0)) return false;
  const double l11 = sqrt(d1);
  const double l21 = (a12 - l20 * l10) / l11, l31 = (a13 - l30 * l10) / l11;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  if (!(d2 > 0.0)) return false;
  const double l22 = sqrt(d2);
  const double l32 = (a23 - l30 * l20 - l31 * l21) / l22;
  const double d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
  if (!(d3 > 0.0)) return false;
  const double l33 = sqrt(d3);
  const double z0 = b[0] / l00;
  const double z1 = (b[1] - l10 * z0) / l11;
  const double z2 = (b[2] - l20 * z0 - l21 * z1) / l22;
  const double z3 = (b[3] - l30 * z0 - l31 * z1 - l32 * z2) / l33;
  y[3] = z3 / l33;
  y[2] = (z2 - l32 * y[3]) / l22;
  y[1] = (z1 - l21 * y[2] - l31 * y[3]) / l11;
  y[0] = (z0 - l10 * y[1] - l20 * y[2] - l30 * y[3]) / l00;
  return true;
}

// cost, and optionally gradient g = J_p^T r (4) and H = J_p^T J_p (10, upper), of the point X over its observation slots.
// Returns false when the residual functor fails (||X - hC||^2 < 1e-8, reprojection_error.h:75-77).
template <bool EXT>
__host__ __device__ inline bool point_linearize(const FilterView& V, long long s0, int len, const double* X, int loss_type, double loss_width,
                                                bool need_derivatives, double* cost, double* g, double* H) {
  double c = 0.0;
  if (need_derivatives) {
    for (int j = 0; j < 4; ++j) g[j] = 0.0;
    for (int j = 0; j < 10; ++j) H[j] = 0.0;
  }
  for (int o = 0; o < len; ++o) {
    const long long s = s0 + o;
    const int cam = V.slot_cam[s];
    const int grp = V.cam_group[cam];
    const long long wq = s >> 5;
    const int l = (int)(s & 31);
    const double x = V.xy[(size_t)(wq * 2 + 0) * 32 + l], y = V.xy[(size_t)(wq * 2 + 1) * 32 + l];
    if (need_derivatives) {
      double r[2], rho0, Ja[6], Jw[6], Jh[2];
      if (!linearize_obs_any<0u, EXT>(V.group_model[grp], V.ext + (size_t)cam * 6, V.cam_rec + (size_t)cam * kCamRec, V.intr + (size_t)grp * 10, X[0],
                             X[1], X[2], X[3], x, y, loss_type, loss_width, r, rho0, Ja, Jw, Jh, nullptr))
        return false;
      c += 0.5 * rho0;
      const double j0[4] = {Ja[0], Ja[1], Ja[2], Jh[0]}, j1[4] = {Ja[3], Ja[4], Ja[5], Jh[1]};
      int n = 0;
      for (int a = 0; a < 4; ++a) {
        g[a] += j0[a] * r[0] + j1[a] * r[1];
        for (int b = a; b < 4; ++b) H[n++] += j0[a] * j0[b] + j1[a] * j1[b];
      }
    } else {
      double r0, r1;
      if (!reproject_any<EXT>(V.group_model[grp], V.ext + (size_t)cam * 6, V.cam_rec + (size_t)cam * kCamRec, V.intr + (size_t)grp * 10, X[0], X[1],
                     X[2], X[3], x, y, r0, r1))
        return false;
      double rho[3];
      loss_evaluate(loss_type, loss_width, r0 * r0 + r1 * r1, rho);
      c += 0.5 * rho[0];
    }
  }
  *cost = c;
  return true;
}

// Minimise over the point X (in/out).  Mirrors TrustRegionMinimizer::Minimize for a single 4-vector block.
template <bool EXT>
__host__ __device__ inline PointLmResult point_lm(const FilterView& V, long long s0, int len, double* X, const PointLmOptions& o) {
  PointLmResult res;
  res.initial_cost = res.final_cost = -1.0; res.iterations = 0; res.termination = 2;
  double cost, g[4], H[10];
  if (!point_linearize<EXT>(V, s0, len, X, o.loss_type, o.loss_width, true, &cost, g, H)) return res;  // "Residual and Jacobian evaluation failed."
  res.initial_cost = res.final_cost = cost;
  const int dg[4] = {0, 4, 7, 9};
  double sc[4];
  for (int j = 0; j < 4; ++j) sc[j] = o.jacobi_scaling ? 1.0 / (1.0 + sqrt(H[dg[j]])) : 1.0;
  double radius = o.initial_radius, decrease = 2.0;
  double xnorm = sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2] + X[3] * X[3]);
  int invalid = 0;
  bool last_successful = true;
  res.termination = 1;
  for (int it = 0;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it >= o.max_num_iterations) { res.termination = 1; break; }
    if (last_successful) {
      double gmax = 0.0;
      for (int j = 0; j < 4; ++j) gmax = fmax(gmax, fabs(g[j]));
      if (gmax <= o.gradient_tolerance) { res.termination = 0; break; }
    }
    if (radius <= o.min_radius) { res.termination = 0; break; }
    ++it;
    res.iterations = it;
    // scaled damped normal equations
    double A[10], b[4], y[4];
    int n = 0;
    for (int a = 0; a < 4; ++a) {
      b[a] = sc[a] * g[a];
      for (int c2 = a; c2 < 4; ++c2) { A[n] = sc[a] * H[n] * sc[c2]; ++n; }
    }
    double Hs[10];
    for (int j = 0; j < 10; ++j) Hs[j] = A[j];
    for (int a = 0; a < 4; ++a) A[dg[a]] += fmin(fmax(A[dg[a]], o.min_diag), o.max_diag) / radius;
    bool valid = spd4_solve(A, b, y);
    double step[4] = {0, 0, 0, 0}, mcc = 0.0;
    if (valid) {
      for (int a = 0; a < 4; ++a) step[a] = -y[a];
      // model_cost_change = -(J step).(r + J step / 2) = -(step.g~ + step^T H~ step / 2)
      const double sHs = Hs[0] * step[0] * step[0] + Hs[4] * step[1] * step[1] + Hs[7] * step[2] * step[2] + Hs[9] * step[3] * step[3] +
                         2.0 * (Hs[1] * step[0] * step[1] + Hs[2] * step[0] * step[2] + Hs[3] * step[0] * step[3] + Hs[5] * step[1] * step[2] +
                                Hs[6] * step[1] * step[3] + Hs[8] * step[2] * step[3]);
      mcc = -(step[0] * b[0] + step[1] * b[1] + step[2] * b[2] + step[3] * b[3] + 0.5 * sHs);
      valid = mcc > 0.0;
    }
    if (!valid) {  // HandleInvalidStep
      if (++invalid >= o.max_consecutive_invalid) { res.termination = 2; break; }
      radius /= decrease; decrease *= 2.0;
      last_successful = false;
      continue;
    }
    invalid = 0;
    double Xc[4], dn = 0.0;
    for (int a = 0; a < 4; ++a) { const double d = step[a] * sc[a]; Xc[a] = X[a] + d; dn += d * d; }
    double cand;
    if (!point_linearize<EXT>(V, s0, len, Xc, o.loss_type, o.loss_width, false, &cand, nullptr, nullptr)) cand = 1.7976931348623157e308;
    if (sqrt(dn) <= o.parameter_tolerance * (xnorm + o.parameter_tolerance)) { res.termination = 0; break; }
    const double cost_change = cost - cand;
    if (fabs(cost_change) <= o.function_tolerance * cost) { res.termination = 0; break; }
    const double rho = cost_change / mcc;
    if (rho > o.min_relative_decrease) {  // HandleSuccessfulStep
      for (int a = 0; a < 4; ++a) X[a] = Xc[a];
      xnorm = sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2] + X[3] * X[3]);
      if (!point_linearize<EXT>(V, s0, len, X, o.loss_type, o.loss_width, true, &cost, g, H)) { res.termination = 2; break; }
      res.final_cost = cost;
      const double t = 2.0 * rho - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
      radius = fmin(o.max_radius, radius);
      decrease = 2.0;
      last_successful = true;
    } else {  // HandleUnsuccessfulStep
      radius /= decrease; decrease *= 2.0;
      last_successful = false;
    }
  }
  return res;
}

}  // namespace tba
