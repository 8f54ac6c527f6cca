// bundle_adjust_two_views_b200.h -- drop-in for BundleAdjustTwoViews (src/theia/sfm/bundle_adjustment/
// bundle_adjust_two_views.h:51-69, .cc:54-191): camera 1 pose fixed, camera 2 pose free, each camera's intrinsics constant
// or focal-length-only (SubsetParameterization over indices 1..K-1, .cc:96-108), every triangulated point free, DENSE_SCHUR,
// 200 iterations, no robust loss, Ceres' default tolerances.  One engine call per image pair: launch-latency bound for the
// few hundred correspondences of a pair -- provided so that a build without Ceres has the call; a batched form (all
// pairs of the view graph in one launch sequence, each with its own trust region) is future work.
#ifndef THEIA_SFM_BUNDLE_ADJUSTMENT_BUNDLE_ADJUST_TWO_VIEWS_B200_H_
#define THEIA_SFM_BUNDLE_ADJUSTMENT_BUNDLE_ADJUST_TWO_VIEWS_B200_H_

#include <vector>

#include "bundle_adjuster_b200.h"
#ifdef THEIA_B200_INSIDE_THEIA
#include "theia/matching/feature_correspondence.h"
#include "theia/sfm/bundle_adjustment/bundle_adjust_two_views.h"
#endif

namespace theia {

#ifndef THEIA_B200_INSIDE_THEIA
// bundle_adjust_two_views.h:51-55
struct TwoViewBundleAdjustmentOptions {
  BundleAdjustmentOptions ba_options;
  bool constant_camera1_intrinsics = true;
  bool constant_camera2_intrinsics = true;
};
typedef Vector4d TwoViewPoint;
#else
typedef Eigen::Vector4d TwoViewPoint;
#endif

// The flattened two-view problem and the engine options (exposed for tests).
void FlattenTwoViewProblem(const TwoViewBundleAdjustmentOptions& options, const std::vector<FeatureCorrespondence>& correspondences,
                           Camera* camera1, Camera* camera2, std::vector<TwoViewPoint>* points3d, BundleAdjusterB200::Flat* flat,
                           tba_options* engine_options);

BundleAdjustmentSummary BundleAdjustTwoViewsB200(const TwoViewBundleAdjustmentOptions& options,
                                                 const std::vector<FeatureCorrespondence>& correspondences, Camera* camera1, Camera* camera2,
                                                 std::vector<TwoViewPoint>* points3d);

}  // namespace theia
#endif
