// host_filter.cc -- runs the PRODUCT's per-track filter body (theiasfm_b200/csrc/tba_filter.cuh, the code k_filter_tracks
// executes per thread) on the host over the PACKED layout produced by tba_debug_pack, so that slot indexing, the
// [tile][warp][2][32] measurement layout and the filter logic are all checked without a GPU (tests/test_track_filter.py).
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
#include <vector>
using std::fmax;
using std::sqrt;

#include "../theiasfm_b200/csrc/tba_filter.cuh"

extern "C" void host_filter(int n_cam, const double* ext, const double* intr, const int* cam_group, const int* group_model, int n_pk,
                            const double* pt_packed, const double* xy, const int* slot_cam, const long long* pt_slot, const int* pt_len,
                            double max_err, double min_angle_deg, unsigned char* status, double* mean) {
  std::vector<double> rec((size_t)n_cam * tba::kCamRec);
  for (int c = 0; c < n_cam; ++c) tba::cam_prep(ext + (size_t)c * 6 + 3, rec.data() + (size_t)c * tba::kCamRec);
  tba::FilterView V;
  V.ext = ext; V.cam_rec = rec.data(); V.intr = intr; V.pt = pt_packed; V.xy = xy; V.slot_cam = slot_cam; V.cam_group = cam_group;
  V.group_model = group_model;
  const double cos_min = std::cos(min_angle_deg * 3.14159265358979323846 / 180.0);
  for (int k = 0; k < n_pk; ++k) status[k] = tba::filter_track<true>(V, k, pt_slot[k], pt_len[k], max_err * max_err, cos_min, &mean[k]);
}
