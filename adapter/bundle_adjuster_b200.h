// bundle_adjuster_b200.h -- drop-in for src/theia/sfm/bundle_adjustment/{bundle_adjustment.h, bundle_adjuster.h}:
// the same option / summary structs, free functions and class, with the ceres::Solve call
// (bundle_adjuster.cc:205) replaced by the C-ABI of include/theia_ba_b200.h.
#ifndef THEIA_SFM_BUNDLE_ADJUSTMENT_BUNDLE_ADJUSTER_B200_H_
#define THEIA_SFM_BUNDLE_ADJUSTMENT_BUNDLE_ADJUSTER_B200_H_

#include <ceres/types.h>

#include <memory>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#ifdef THEIA_B200_INSIDE_THEIA
#include "theia/sfm/bundle_adjustment/bundle_adjustment.h"  // options / summary come from Theia itself
#include "theia/sfm/reconstruction.h"
#else
#include "theia/sfm/scene.h"  // theia_compat stand-in
#endif

#include "theia_ba_b200.h"

namespace theia {

#ifndef THEIA_B200_INSIDE_THEIA
// create_loss_function.h:51-58
enum class LossFunctionType { TRIVIAL = 0, HUBER = 1, SOFTLONE = 2, CAUCHY = 3, ARCTAN = 4, TUKEY = 5 };

// bundle_adjustment.h:78-122 (field for field, default for default)
struct BundleAdjustmentOptions {
  LossFunctionType loss_function_type = LossFunctionType::TRIVIAL;
  double robust_loss_width = 2.0;
  ceres::LinearSolverType linear_solver_type = ceres::SPARSE_SCHUR;
  ceres::PreconditionerType preconditioner_type = ceres::SCHUR_JACOBI;
  ceres::VisibilityClusteringType visibility_clustering_type = ceres::CANONICAL_VIEWS;
  bool verbose = false;
  bool constant_camera_orientation = false;
  bool constant_camera_position = false;
  OptimizeIntrinsicsType intrinsics_to_optimize = OptimizeIntrinsicsType::FOCAL_LENGTH | OptimizeIntrinsicsType::RADIAL_DISTORTION;
  int num_threads = 1;
  int max_num_iterations = 100;
  double max_solver_time_in_seconds = 3600.0;
  bool use_inner_iterations = true;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  double max_trust_region_radius = 1e12;
};

// bundle_adjustment.h:125-133
struct BundleAdjustmentSummary {
  bool success = false;
  double initial_cost = 0.0;
  double final_cost = 0.0;
  double setup_time_in_seconds = 0.0;
  double solve_time_in_seconds = 0.0;
};
#endif

// Same public interface as theia::BundleAdjuster (bundle_adjuster.h:60-77).  AddView must be called before AddTrack
// for views that are optimised (same requirement as the reference, bundle_adjuster.h:55-56).
class BundleAdjusterB200 {
 public:
  BundleAdjusterB200(const BundleAdjustmentOptions& options, Reconstruction* reconstruction);
  ~BundleAdjusterB200();
  void AddView(const ViewId view_id);
  void AddTrack(const TrackId track_id);
  BundleAdjustmentSummary Optimize();

  // For problems built with AddTrack only (every camera constant, bundle_adjuster.cc:156-175): each track is its own
  // 4-parameter problem; they are solved independently, one GPU thread per track (tba_adjust_tracks), which for one track
  // is exactly BundleAdjustTrack.  success = no track failed; costs are summed over the tracks.
  BundleAdjustmentSummary OptimizeTracks();

  // The flattened problem Optimize() hands to tba_solve (exposed for tests / tools).
  struct Flat {
    std::vector<double> ext, intr, pt, obs_xy;
    std::vector<uint8_t> ext_const, pt_const;
    std::vector<int32_t> cam_group, group_model, obs_cam, obs_pt;
    std::vector<uint32_t> group_const_mask;
    std::vector<ViewId> view_of_cam;
    std::vector<TrackId> track_of_pt;
    std::vector<CameraIntrinsicsGroupId> id_of_group;
    tba_problem AsProblem();
  };
  void Flatten(Flat* flat, tba_options* options) const;
  // N1: SetOutlierTracksToUnestimated (set_outlier_tracks_to_unestimated.cc:62-136) for the tracks of THIS problem,
  // evaluated on the device-resident copy the last Optimize() left on the GPU (no re-flattening, no re-upload).
  // Marks outlier tracks un-estimated in the Reconstruction and returns how many were removed, or -1 when there is
  // no device-resident problem (Optimize() not run, failed, ran through the multi-GPU entry point, or another
  // BundleAdjusterB200 / TrackEstimatorB200 has used the engine context since).
  int SetOutlierTracksToUnestimated(const double max_inlier_reprojection_error, const double min_triangulation_angle_degrees);

  // Detail of the last Optimize() (termination, iteration count, message).
  const tba_summary& last_summary() const { return last_summary_; }

 private:
  void AddResidual(ViewId view_id, TrackId track_id);
  const BundleAdjustmentOptions options_;
  Reconstruction* reconstruction_;
  double start_time_;
  std::unordered_set<ViewId> optimized_views_;
  std::unordered_set<TrackId> optimized_tracks_;
  std::unordered_set<CameraIntrinsicsGroupId> optimized_camera_intrinsics_groups_;
  std::unordered_set<CameraIntrinsicsGroupId> potentially_constant_camera_intrinsics_groups_;
  std::unordered_set<ViewId> constant_extrinsics_views_;  // reached through AddTrack before any AddView (bundle_adjuster.cc:166-168): the
                                                          // reference's SetParameterBlockConstant is never undone, a later AddView included
  std::unordered_set<TrackId> variable_tracks_;           // SetTrackVariable (:178) / SetTrackConstant (:137) in CALL ORDER: the last call wins
  std::vector<std::pair<ViewId, TrackId>> residuals_;     // in insertion order
  tba_summary last_summary_;
  std::vector<TrackId> resident_tracks_;  // point index -> TrackId of the problem left on the device by Optimize()
  bool resident_ = false;
  uint64_t generation_ = 0;               // b200::Generation() when that problem was uploaded
};

// Drop-ins for bundle_adjustment.h:145-155 (bundle_adjustment.cc:82-107).  The reference forces DENSE_QR and no inner
// iterations; the engine computes the same exact LM step (see tba_options::linear_solver_type).  BundleAdjustTrackB200
// is provided for API completeness: one call = one tiny GPU launch sequence; use TrackEstimatorB200 for batches.
BundleAdjustmentSummary BundleAdjustViewB200(const BundleAdjustmentOptions& options, const ViewId view_id, Reconstruction* reconstruction);
BundleAdjustmentSummary BundleAdjustTrackB200(const BundleAdjustmentOptions& options, const TrackId track_id, Reconstruction* reconstruction);

// Shared by the adapters of this directory: the process-wide engine context (guarded by Mutex()), a counter that
// identifies which flattened problem currently lives on the device, and the options copy.
namespace b200 {
std::mutex& Mutex();
tba_context* AcquireContext();   // creates the context on first use (THEIA_B200_DEVICE); nullptr without a usable GPU
tba_context* CurrentContext();
uint64_t& Generation();
void ToEngineOptions(const BundleAdjustmentOptions& in, tba_options* out);
}  // namespace b200

// Drop-ins for bundle_adjustment.h:136-143 (bundle_adjustment.cc:47-80).
BundleAdjustmentSummary BundleAdjustReconstructionB200(const BundleAdjustmentOptions& options, Reconstruction* reconstruction);
BundleAdjustmentSummary BundleAdjustPartialReconstructionB200(const BundleAdjustmentOptions& options,
                                                             const std::unordered_set<ViewId>& views_to_optimize,
                                                             const std::unordered_set<TrackId>& tracks_to_optimize,
                                                             Reconstruction* reconstruction);

}  // namespace theia
#endif
