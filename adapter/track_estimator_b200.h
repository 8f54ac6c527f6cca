// track_estimator_b200.h -- drop-in for src/theia/sfm/estimate_track.h: the same Options / Summary / methods, with the
// per-track ThreadPool loop of TrackEstimator::EstimateTracks (estimate_track.cc:133-197: triangulate + BundleAdjustTrack
// + reprojection test, one Ceres problem per track) replaced by ONE batched call, tba_estimate_tracks
// (include/theia_ba_b200.h) -- SURVEY 8f row N3: "needs a batch API upstream in TrackEstimator".
#ifndef THEIA_SFM_TRACK_ESTIMATOR_B200_H_
#define THEIA_SFM_TRACK_ESTIMATOR_B200_H_

#include <unordered_set>
#include <vector>

#include "bundle_adjuster_b200.h"
#ifdef THEIA_B200_INSIDE_THEIA
#include "theia/sfm/estimate_track.h"
#endif

namespace theia {

class TrackEstimatorB200 {
 public:
#ifdef THEIA_B200_INSIDE_THEIA
  typedef TrackEstimator::Options Options;
  typedef TrackEstimator::Summary Summary;
#else
  // estimate_track.h:56-80 (field for field, default for default)
  struct Options {
    int num_threads = 1;  // host threads of the reference; unused here
    double max_acceptable_reprojection_error_pixels = 5.0;
    double min_triangulation_angle_degrees = 3.0;
    bool bundle_adjustment = true;
    BundleAdjustmentOptions ba_options;
    int multithreaded_step_size = 100;  // unused here
  };
  // estimate_track.h:82-92
  struct Summary {
    int input_num_estimated_tracks = 0;
    int num_triangulation_attempts = 0;
    std::unordered_set<TrackId> estimated_tracks;
  };
#endif

  TrackEstimatorB200(const Options& options, Reconstruction* reconstruction) : options_(options), reconstruction_(reconstruction) {}

  // Attempts to estimate all unestimated tracks seen by estimated views (estimate_track.cc:116-131).
  Summary EstimateAllTracks();
  // Estimate only the tracks supplied by the user (estimate_track.cc:133-197).
  Summary EstimateTracks(const std::unordered_set<TrackId>& track_ids);

  // Counters the reference logs (estimate_track.cc:188-195) + the ones it drops, from the last EstimateTracks call.
  int num_bad_angles() const { return counts_[TBA_TRACK_BAD_ANGLE]; }
  int num_failed_triangulations() const { return counts_[TBA_TRACK_TRIANGULATION_FAILED]; }
  int num_failed_bundle_adjustments() const { return counts_[TBA_TRACK_BA_FAILED]; }
  int num_bad_reprojections() const { return counts_[TBA_TRACK_BAD_REPROJECTION]; }
  bool engine_ok() const { return engine_ok_; }  // false: no GPU / engine error -> nothing was estimated (never a CPU path)

 private:
  const Options options_;
  Reconstruction* reconstruction_;
  int32_t counts_[5] = {0, 0, 0, 0, 0};
  bool engine_ok_ = true;
};

}  // namespace theia
#endif
