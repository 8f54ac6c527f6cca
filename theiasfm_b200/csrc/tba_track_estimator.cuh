// tba_track_estimator.cuh -- per-track body of TrackEstimator::EstimateTrack (src/theia/sfm/estimate_track.cc:199-264):
//   rays from the observing cameras (Camera::PixelToUnitDepthRay, camera.cc:215-223; UndistortPoint of the two supported
//   models: pinhole_camera_model.h:259-296, pinhole_radial_tangential_camera_model.h:293-355)
//   -> SufficientTriangulationAngle (triangulation.cc:236-250) -> TriangulateMidpoint (triangulation.cc:130-157)
//   -> BundleAdjustTrack (bundle_adjustment.cc:96-107; tba_point_lm.cuh) -> AcceptableReprojectionError
//   (estimate_track.cc:88-113).
// Host/device: the GPU runs observation_ray in k_track_rays (one thread per observation slot) and estimate_track in
// k_estimate_tracks (one thread per point); the CPU test suite runs the same bodies on the packed layout
// (tests/host_point_lm.cc, tests/test_track_estimator.py) against oracle/ba_oracle.c:oracle_estimate_tracks.
#pragma once
#include <cstdint>

#include "tba_point_lm.cuh"

namespace tba {

// status codes of tba_estimate_tracks (include/theia_ba_b200.h)
enum : uint8_t { kTrackEstimated = 0, kTrackBadAngle = 1, kTrackTriangulationFailed = 2, kTrackBaFailed = 3, kTrackBadReprojection = 4, kTrackSkipped = 255 };

struct TrackEstimatorOptions {
  double max_sq_reprojection_error;  // max_acceptable_reprojection_error_pixels^2
  double cos_min_angle;              // cos(min_triangulation_angle_degrees)
  int bundle_adjustment;
  PointLmOptions lm;
};

// Fixed-point inversion of the lens distortion exactly as the reference iterates it (100 iterations, 1e-10 on both coordinates).
__host__ __device__ inline void undistort_point(int model, const double* __restrict__ k, double xd, double yd, double& xu, double& yu) {
  xu = xd; yu = yd;
  for (int i = 0; i < 100; ++i) {
    const double px = xu, py = yu;
    const double r2 = xu * xu + yu * yu;
    if (model == kModelPinhole) {
      const double d = 1.0 + r2 * (k[5] + k[6] * r2);
      xu = xd / d; yu = yd / d;
    } else {
      const double rd = 1.0 + k[5] * r2 + k[6] * r2 * r2 + k[7] * r2 * r2 * r2;
      const double tx = k[9] * (r2 + 2.0 * xu * xu) + 2.0 * k[8] * xu * yu;
      const double ty = k[8] * (r2 + 2.0 * yu * yu) + 2.0 * k[9] * xu * yu;
      xu = (xd - tx) / rd; yu = (yd - ty) / rd;
    }
    if (fabs(xu - px) < 1e-10 && fabs(yu - py) < 1e-10) break;
  }
}

// Unit world-frame ray of pixel (x, y): normalise(R^T [undistort(K^-1 pixel), 1]).
__host__ __device__ inline void observation_ray(int model, const double* __restrict__ R, const double* __restrict__ k, double x, double y, double d[3]) {
  double xu, yu;
  if (model >= kModelFisheye) pixel_to_camera_ext(model, k, x, y, xu, yu);
  else {
    const double yd = (y - k[4]) / (k[0] * k[1]);
    const double xd = (x - k[3] - yd * k[2]) / k[0];
    undistort_point(model, k, xd, yd, xu, yu);
  }
  const double d0 = R[0] * xu + R[3] * yu + R[6];
  const double d1 = R[1] * xu + R[4] * yu + R[7];
  const double d2 = R[2] * xu + R[5] * yu + R[8];
  const double n = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  d[0] = d0 / n; d[1] = d1 / n; d[2] = d2 / n;
}

__host__ __device__ inline void load_ray(const double* __restrict__ ray, long long s, double d[3]) {  // ray: [slot / 32][3][32]
  const long long wq = s >> 5;
  const int l = (int)(s & 31);
  d[0] = ray[(size_t)(wq * 3 + 0) * 32 + l]; d[1] = ray[(size_t)(wq * 3 + 1) * 32 + l]; d[2] = ray[(size_t)(wq * 3 + 2) * 32 + l];
}

// Estimates packed point X (in/out, written at the same stages the reference overwrites Track::MutablePoint) from the
// observations in slots [s0, s0 + len) whose unit rays are in `ray`.
template <bool EXT>
__host__ __device__ inline uint8_t estimate_track(const FilterView& V, const double* __restrict__ ray, long long s0, int len, double* X,
                                                  const TrackEstimatorOptions& o, PointLmResult* lm) {
  lm->initial_cost = lm->final_cost = -1.0; lm->iterations = 0; lm->termination = 2;
  if (len < 2) return kTrackBadAngle;
  bool wide = false;
  for (int i = 0; i < len && !wide; ++i) {
    double a[3];
    load_ray(ray, s0 + i, a);
    for (int j = i + 1; j < len; ++j) {
      double b[3];
      load_ray(ray, s0 + j, b);
      if (a[0] * b[0] + a[1] * b[1] + a[2] * b[2] < o.cos_min_angle) { wide = true; break; }
    }
  }
  if (!wide) return kTrackBadAngle;
  // midpoint: A = sum (I4 - [d;0][d;0]^T), b = sum A_i [origin;1]
  double A[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  for (int i = 0; i < len; ++i) {
    double d[3];
    load_ray(ray, s0 + i, d);
    const double* C = V.ext + (size_t)V.slot_cam[s0 + i] * 6;
    const double t00 = 1.0 - d[0] * d[0], t01 = 0.0 - d[0] * d[1], t02 = 0.0 - d[0] * d[2];
    const double t11 = 1.0 - d[1] * d[1], t12 = 0.0 - d[1] * d[2], t22 = 1.0 - d[2] * d[2];
    A[0] += t00; A[1] += t01; A[2] += t02; A[4] += t11; A[5] += t12; A[7] += t22; A[9] += 1.0;
    b[0] += t00 * C[0] + t01 * C[1] + t02 * C[2];
    b[1] += t01 * C[0] + t11 * C[1] + t12 * C[2];
    b[2] += t02 * C[0] + t12 * C[1] + t22 * C[2];
    b[3] += 1.0;
  }
  double Y[4];
  if (!spd4_solve(A, b, Y)) return kTrackTriangulationFailed;
  X[0] = Y[0]; X[1] = Y[1]; X[2] = Y[2]; X[3] = Y[3];
  if (o.bundle_adjustment) {
    *lm = point_lm<EXT>(V, s0, len, X, o.lm);
    if (lm->termination == 2) return kTrackBaFailed;
  }
  // AcceptableReprojectionError: any view behind -> false; mean squared error must be < max^2
  double sum = 0.0;
  for (int i = 0; i < len; ++i) {
    const long long s = s0 + i;
    const int cam = V.slot_cam[s];
    const int grp = V.cam_group[cam];
    const long long wq = s >> 5;
    const int l = (int)(s & 31);
    const double x = V.xy[(size_t)(wq * 2 + 0) * 32 + l], y = V.xy[(size_t)(wq * 2 + 1) * 32 + l];
    double px, py, qz, a_sq;
    project_pixel_any<EXT>(V.group_model[grp], V.ext + (size_t)cam * 6, V.cam_rec + (size_t)cam * kCamRec, V.intr + (size_t)grp * 10, X[0], X[1], X[2], X[3],
                  px, py, qz, a_sq);
    if (qz / X[3] < 0.0) return kTrackBadReprojection;
    sum += (x - px) * (x - px) + (y - py) * (y - py);
  }
  return (sum / (double)len < o.max_sq_reprojection_error) ? kTrackEstimated : kTrackBadReprojection;
}

}  // namespace tba
