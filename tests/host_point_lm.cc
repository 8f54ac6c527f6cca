// host_point_lm.cc -- runs the PRODUCT's per-point LM and track estimator (theiasfm_b200/csrc/tba_point_lm.cuh,
// tba_track_estimator.cuh: the bodies of k_adjust_tracks / k_track_rays / k_estimate_tracks) on
// the host over the packed layout of tba_debug_pack, for the CPU test suite (tests/test_point_lm.py).
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
#include <vector>
using std::fabs;
using std::fmax;
using std::fmin;
using std::sqrt;
using std::cos;

#include "../theiasfm_b200/csrc/tba_track_estimator.cuh"

extern "C" void host_adjust_tracks(int n_cam, const double* ext, const double* intr, const int* cam_group, const int* group_model, int n_pk,
                                   double* pt_packed, const double* xy, const int* slot_cam, const long long* pt_slot, const int* pt_len,
                                   int loss_type, double loss_width, int max_iters, double* initial_cost, double* final_cost, int* termination,
                                   int* iterations) {
  std::vector<double> rec((size_t)n_cam * tba::kCamRec);
  for (int c = 0; c < n_cam; ++c) tba::cam_prep(ext + (size_t)c * 6 + 3, rec.data() + (size_t)c * tba::kCamRec);
  tba::FilterView V;
  V.ext = ext; V.cam_rec = rec.data(); V.intr = intr; V.pt = pt_packed; V.xy = xy; V.slot_cam = slot_cam; V.cam_group = cam_group;
  V.group_model = group_model;
  tba::PointLmOptions o;
  o.loss_type = loss_type; o.loss_width = loss_width; o.max_num_iterations = max_iters;
  o.function_tolerance = 1e-6; o.gradient_tolerance = 1e-10; o.parameter_tolerance = 1e-8;
  o.initial_radius = 1e4; o.max_radius = 1e12; o.min_radius = 1e-32; o.min_relative_decrease = 1e-3; o.min_diag = 1e-6; o.max_diag = 1e32;
  o.jacobi_scaling = 1; o.max_consecutive_invalid = 5;
  for (int k = 0; k < n_pk; ++k) {
    const tba::PointLmResult r = tba::point_lm<true>(V, pt_slot[k], pt_len[k], pt_packed + (size_t)k * 4, o);
    initial_cost[k] = r.initial_cost; final_cost[k] = r.final_cost; termination[k] = r.termination; iterations[k] = r.iterations;
  }
}

static tba::PointLmOptions default_lm(int loss_type, double loss_width, int max_iters) {
  tba::PointLmOptions o;
  o.loss_type = loss_type; o.loss_width = loss_width; o.max_num_iterations = max_iters;
  o.function_tolerance = 1e-6; o.gradient_tolerance = 1e-10; o.parameter_tolerance = 1e-8;
  o.initial_radius = 1e4; o.max_radius = 1e12; o.min_radius = 1e-32; o.min_relative_decrease = 1e-3; o.min_diag = 1e-6; o.max_diag = 1e32;
  o.jacobi_scaling = 1; o.max_consecutive_invalid = 5;
  return o;
}

// k_track_rays + k_estimate_tracks on the host: ray scratch [n_slots/32][3][32], one estimate per packed point.
extern "C" void host_estimate_tracks(int n_cam, const double* ext, const double* intr, const int* cam_group, const int* group_model, int n_pk,
                                     long long n_slots, double* pt_packed, const double* xy, const int* slot_cam, const long long* pt_slot,
                                     const int* pt_len, int loss_type, double loss_width, int max_iters, double max_px, double min_angle_deg,
                                     int bundle_adjustment, unsigned char* status) {
  std::vector<double> rec((size_t)n_cam * tba::kCamRec);
  for (int c = 0; c < n_cam; ++c) tba::cam_prep(ext + (size_t)c * 6 + 3, rec.data() + (size_t)c * tba::kCamRec);
  std::vector<double> ray((size_t)n_slots * 3, 0.0);
  for (long long s = 0; s < n_slots; ++s) {
    const int cam = slot_cam[s];
    if (cam < 0) continue;
    const int grp = cam_group[cam];
    const long long wq = s >> 5; const int l = (int)(s & 31);
    double d[3];
    tba::observation_ray(group_model[grp], rec.data() + (size_t)cam * tba::kCamRec, intr + (size_t)grp * 10, xy[(size_t)(wq * 2) * 32 + l],
                         xy[(size_t)(wq * 2 + 1) * 32 + l], d);
    for (int j = 0; j < 3; ++j) ray[(size_t)(wq * 3 + j) * 32 + l] = d[j];
  }
  tba::FilterView V;
  V.ext = ext; V.cam_rec = rec.data(); V.intr = intr; V.pt = pt_packed; V.xy = xy; V.slot_cam = slot_cam; V.cam_group = cam_group;
  V.group_model = group_model;
  tba::TrackEstimatorOptions o;
  o.max_sq_reprojection_error = max_px * max_px; o.cos_min_angle = cos(min_angle_deg * 3.14159265358979323846 / 180.0);
  o.bundle_adjustment = bundle_adjustment; o.lm = default_lm(loss_type, loss_width, max_iters);
  for (int k = 0; k < n_pk; ++k) {
    tba::PointLmResult lm;
    status[k] = tba::estimate_track<true>(V, ray.data(), pt_slot[k], pt_len[k], pt_packed + (size_t)k * 4, o, &lm);
  }
}
