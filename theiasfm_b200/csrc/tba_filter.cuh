// tba_filter.cuh -- per-track body of the post-BA outlier filter (N1), host/device so that the exact code the GPU
// runs is also exercised by the CPU test suite on the packed layout (tests/host_filter.cc, tests/test_track_filter.py).
// Reference: SetOutlierTracksToUnestimated (src/theia/sfm/set_outlier_tracks_to_unestimated.cc:62-136),
// SufficientTriangulationAngle (triangulation.cc:236-250), ComputeStatisticsForTrack
// (select_good_tracks_for_bundle_adjustment.cc:79-108).
#pragma once
#include <cstdint>

#include "tba_camera_models.cuh"

namespace tba {

struct FilterView {
  const double* ext;       // [n_cam][6]
  const double* cam_rec;   // [n_cam][kCamRec] for ext
  const double* intr;      // [n_group][10]
  const double* pt;        // [packed point][4]
  const double* xy;        // [tile][warp][2][32]
  const int* slot_cam;     // [slot]
  const int* cam_group;    // [n_cam]
  const int* group_model;  // [n_group]
};

// EXT: the problem may contain FISHEYE / FOV / DIVISION_UNDISTORTION groups (tba_camera_models_ext.cuh).
// Status of packed point k whose observations occupy slots [s0, s0 + len): 0 keep, 1 bad reprojection (a view sees the
// point at negative depth, or the mean squared reprojection error exceeds max_sq_err), 2 insufficient viewing angle
// (no pair of unit rays X/h - C with dot < cos_min_angle).  The reference "breaks" at the first negative depth, which
// only affects counters that are discarded for such a track: the result does not depend on its hash-map view order.
template <bool EXT>
__host__ __device__ inline uint8_t filter_track(const FilterView& V, int k, long long s0, int len, double max_sq_err, double cos_min_angle,
                                                double* mean_sq_err) {
  const double X0 = V.pt[(size_t)k * 4], X1 = V.pt[(size_t)k * 4 + 1], X2 = V.pt[(size_t)k * 4 + 2], h = V.pt[(size_t)k * 4 + 3];
  bool behind = false;
  double sum = 0.0;
  for (int o = 0; o < len; ++o) {
    const long long s = s0 + o;
    const int cam = V.slot_cam[s];
    const int grp = V.cam_group[cam];
    const long long wq = s >> 5;
    const int l = (int)(s & 31);
    const double x = V.xy[(size_t)(wq * 2 + 0) * 32 + l], y = V.xy[(size_t)(wq * 2 + 1) * 32 + l];
    double px, py, qz, a_sq;
    project_pixel_any<EXT>(V.group_model[grp], V.ext + (size_t)cam * 6, V.cam_rec + (size_t)cam * kCamRec, V.intr + (size_t)grp * 10, X0, X1, X2, h,
                  px, py, qz, a_sq);
    if (qz / h < 0.0) behind = true;
    sum += (px - x) * (px - x) + (py - y) * (py - y);
  }
  const double mean = sum / (double)len;  // len == 0 cannot happen for a packed point
  *mean_sq_err = mean;
  if (behind || mean > max_sq_err) return 1;
  const double ih = 1.0 / h;
  const double Xn0 = X0 * ih, Xn1 = X1 * ih, Xn2 = X2 * ih;
  for (int i = 0; i < len; ++i) {
    const double* Ci = V.ext + (size_t)V.slot_cam[s0 + i] * 6;
    double a0 = Xn0 - Ci[0], a1 = Xn1 - Ci[1], a2 = Xn2 - Ci[2];
    const double na = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    a0 /= na; a1 /= na; a2 /= na;
    for (int j = i + 1; j < len; ++j) {
      const double* Cj = V.ext + (size_t)V.slot_cam[s0 + j] * 6;
      double b0 = Xn0 - Cj[0], b1 = Xn1 - Cj[1], b2 = Xn2 - Cj[2];
      const double nb = sqrt(b0 * b0 + b1 * b1 + b2 * b2);
      b0 /= nb; b1 /= nb; b2 /= nb;
      if (a0 * b0 + a1 * b1 + a2 * b2 < cos_min_angle) return 0;  // wide enough: keep
    }
  }
  return 2;
}

}  // namespace tba
