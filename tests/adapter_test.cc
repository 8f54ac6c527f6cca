// adapter_test.cc -- test driver for adapter/bundle_adjuster_b200.{h,cc} (built by adapter/Makefile, run by
// tests/test_adapter.py).  `adapter_test flatten` checks, without a GPU, that AddView / AddTrack / Optimize's
// parameterisation decisions reproduce bundle_adjuster.cc:102-180,223-287.  `adapter_test solve` (GPU) runs
// BundleAdjustReconstructionB200 / BundleAdjustPartialReconstructionB200 end to end and compares with the CPU
// oracle (dlopen'ed from oracle/libba_oracle.so; test infrastructure only).
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "bundle_adjuster_b200.h"
#include "track_estimator_b200.h"
#include "bundle_adjust_two_views_b200.h"

using namespace theia;

#define EXPECT(cond)                                                                      \
  do {                                                                                    \
    if (!(cond)) { std::fprintf(stderr, "FAILED: %s (%s:%d)\n", #cond, __FILE__, __LINE__); std::exit(1); } \
  } while (0)

// plain-double Camera::ProjectPoint for a PINHOLE camera with identity-ish intrinsics (test data synthesis only)
static void Project(const double* ext, const double* k, const double* X, double* pix) {
  const double a[3] = {X[0] - X[3] * ext[0], X[1] - X[3] * ext[1], X[2] - X[3] * ext[2]};
  const double* w = ext + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double q[3];
  if (th2 > 1e-16) {
    const double th = std::sqrt(th2), c = std::cos(th), s = std::sin(th);
    const double kx = w[0] / th, ky = w[1] / th, kz = w[2] / th;
    const double cr[3] = {ky * a[2] - kz * a[1], kz * a[0] - kx * a[2], kx * a[1] - ky * a[0]};
    const double d = (kx * a[0] + ky * a[1] + kz * a[2]) * (1 - c);
    q[0] = a[0] * c + cr[0] * s + kx * d; q[1] = a[1] * c + cr[1] * s + ky * d; q[2] = a[2] * c + cr[2] * s + kz * d;
  } else { q[0] = a[0]; q[1] = a[1]; q[2] = a[2]; }
  const double u = q[0] / q[2], v = q[1] / q[2], r2 = u * u + v * v, d = 1 + r2 * (k[5] + k[6] * r2);
  pix[0] = k[0] * u * d + k[2] * v * d + k[3];
  pix[1] = k[0] * k[1] * v * d + k[4];
}

struct Scene {
  Reconstruction rec;
  std::vector<ViewId> views;
  std::vector<TrackId> tracks;
};

// n_views cameras in front of a point cloud; views 0..n_shared-1 share one intrinsics group, the others own theirs.
static void BuildScene(Scene* sc, int n_views, int n_tracks, int obs_per_track, int n_shared, unsigned seed) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  std::normal_distribution<double> N(0.0, 1.0);
  std::vector<std::vector<double>> gt_ext(n_views, std::vector<double>(6));
  for (int i = 0; i < n_views; ++i) {
    const ViewId id = i < n_shared ? sc->rec.AddView("v" + std::to_string(i), 7) : sc->rec.AddView("v" + std::to_string(i));
    sc->views.push_back(id);
    View* v = sc->rec.MutableView(id);
    v->SetEstimated(true);
    double* e = v->MutableCamera()->mutable_extrinsics();
    for (int j = 0; j < 3; ++j) gt_ext[i][j] = 2.0 * U(rng);
    for (int j = 3; j < 6; ++j) gt_ext[i][j] = 0.1 * U(rng);
    for (int j = 0; j < 6; ++j) e[j] = gt_ext[i][j] + (j < 3 ? 0.02 : 0.004) * N(rng);
    double* k = v->MutableCamera()->mutable_intrinsics();
    k[0] = 800.0 * (i < n_shared ? 1.01 : 1.0 + 0.01 * U(rng)); k[1] = 1.0; k[2] = 0.0; k[3] = 500.0; k[4] = 500.0; k[5] = 0.0; k[6] = 0.0;
  }
  const double kgt[7] = {800.0, 1.0, 0.0, 500.0, 500.0, -0.05, 0.01};
  for (int t = 0; t < n_tracks; ++t) {
    double X[4] = {1.5 * U(rng), 1.5 * U(rng), 12.0 + 2.0 * U(rng), 1.0};
    std::vector<std::pair<ViewId, Feature>> obs;
    for (int o = 0; o < obs_per_track; ++o) {
      const int vi = (t * 3 + o * 5) % n_views;
      bool dup = false;
      for (auto& ob : obs) dup |= ob.first == sc->views[vi];
      if (dup) continue;
      double pix[2];
      Project(gt_ext[vi].data(), kgt, X, pix);
      obs.emplace_back(sc->views[vi], Feature(pix[0] + 0.3 * N(rng), pix[1] + 0.3 * N(rng)));
    }
    const TrackId id = sc->rec.AddTrack(obs);
    sc->tracks.push_back(id);
    Track* tr = sc->rec.MutableTrack(id);
    tr->SetEstimated(true);
    for (int j = 0; j < 3; ++j) (*tr->MutablePoint())[j] = X[j] + 0.03 * N(rng);
  }
}

static BundleAdjustmentOptions IterativeOptions() {
  BundleAdjustmentOptions o;
  o.linear_solver_type = ceres::ITERATIVE_SCHUR;
  o.use_inner_iterations = false;
  return o;
}

static int TestFlatten() {
  Scene sc;
  BuildScene(&sc, 10, 60, 4, 4, 11);
  // one un-estimated view and one un-estimated track must be skipped everywhere
  sc.rec.MutableView(sc.views[9])->SetEstimated(false);
  sc.rec.MutableTrack(sc.tracks[5])->SetEstimated(false);
  {
    // --- full BA (bundle_adjustment.cc:66-80): every estimated view and track
    BundleAdjusterB200 ba(IterativeOptions(), &sc.rec);
    for (ViewId v : sc.rec.ViewIds()) ba.AddView(v);
    for (TrackId t : sc.rec.TrackIds()) ba.AddTrack(t);
    BundleAdjusterB200::Flat f; tba_options o;
    ba.Flatten(&f, &o);
    EXPECT(o.linear_solver_type == TBA_ITERATIVE_SCHUR && o.use_inner_iterations == 0 && o.max_num_iterations == 100);
    EXPECT(o.intrinsics_to_optimize == (TBA_INTR_FOCAL_LENGTH | TBA_INTR_RADIAL_DISTORTION));
    EXPECT(f.view_of_cam.size() == 9);
    for (size_t i = 0; i < f.view_of_cam.size(); ++i) { EXPECT(f.view_of_cam[i] != sc.views[9]); EXPECT(f.ext_const[i] == 0); }
    for (size_t q = 0; q < f.track_of_pt.size(); ++q) { EXPECT(f.track_of_pt[q] != sc.tracks[5]); EXPECT(f.pt_const[q] == 0); }
    for (size_t g = 0; g < f.id_of_group.size(); ++g) { EXPECT(f.group_model[g] == TBA_MODEL_PINHOLE); EXPECT(f.group_const_mask[g] == 0x1Eu); }
    EXPECT(f.id_of_group.size() == 1 + 5);  // the shared group + 5 estimated single-view groups
    // every residual appears exactly once
    std::unordered_set<uint64_t> seen;
    for (size_t k = 0; k < f.obs_cam.size(); ++k) EXPECT(seen.insert((uint64_t)f.obs_cam[k] << 32 | (uint32_t)f.obs_pt[k]).second);
    size_t expected = 0;
    for (TrackId t : sc.rec.TrackIds()) { Track* tr = sc.rec.MutableTrack(t); if (!tr->IsEstimated()) continue; for (ViewId v : tr->ViewIds()) expected += sc.rec.MutableView(v)->IsEstimated(); }
    EXPECT(f.obs_cam.size() == expected);
  }
  {
    // --- partial BA (bundle_adjustment.cc:47-63): views {0,1}, tracks = a few tracks
    BundleAdjustmentOptions opt = IterativeOptions();
    opt.constant_camera_position = true;
    opt.intrinsics_to_optimize = OptimizeIntrinsicsType::FOCAL_LENGTH | OptimizeIntrinsicsType::PRINCIPAL_POINTS;
    BundleAdjusterB200 ba(opt, &sc.rec);
    ba.AddView(sc.views[0]);
    ba.AddView(sc.views[5]);
    ba.AddView(sc.views[0]);  // idempotent (:106-108)
    std::unordered_set<TrackId> opt_tracks = {sc.tracks[0], sc.tracks[1], sc.tracks[2]};
    for (TrackId t : opt_tracks) ba.AddTrack(t);
    BundleAdjusterB200::Flat f; tba_options o;
    ba.Flatten(&f, &o);
    EXPECT(o.constant_camera_position == 1);
    for (size_t i = 0; i < f.view_of_cam.size(); ++i) {
      const bool optimised = f.view_of_cam[i] == sc.views[0] || f.view_of_cam[i] == sc.views[5];
      EXPECT(f.ext_const[i] == (optimised ? TBA_EXT_POSITION_CONST : TBA_EXT_ALL_CONST));
      const CameraIntrinsicsGroupId gid = sc.rec.CameraIntrinsicsGroupIdFromViewId(f.view_of_cam[i]);
      const bool group_optimised = gid == sc.rec.CameraIntrinsicsGroupIdFromViewId(sc.views[0]) || gid == sc.rec.CameraIntrinsicsGroupIdFromViewId(sc.views[5]);
      // FOCAL_LENGTH | PRINCIPAL_POINTS free -> constant {1,2,5,6} = 0x66; groups seen only through constant cameras: all 7 constant
      EXPECT(f.group_const_mask[f.cam_group[i]] == (group_optimised ? 0x66u : 0x7Fu));
    }
    for (size_t q = 0; q < f.track_of_pt.size(); ++q) EXPECT(f.pt_const[q] == (opt_tracks.count(f.track_of_pt[q]) ? 0 : 1));
    // a view of the shared group that is NOT optimised still shares the optimised group's block (view 1 shares with view 0)
    bool found_shared_const_cam = false;
    for (size_t i = 0; i < f.view_of_cam.size(); ++i)
      if (f.view_of_cam[i] == sc.views[1] || f.view_of_cam[i] == sc.views[2] || f.view_of_cam[i] == sc.views[3]) {
        found_shared_const_cam = true;
        EXPECT(f.ext_const[i] == TBA_EXT_ALL_CONST && f.group_const_mask[f.cam_group[i]] == 0x66u);
      }
    EXPECT(found_shared_const_cam);
  }
  {
    // --- call order (not how Theia's own callers use the class, but what the reference class does): AddTrack BEFORE AddView
    BundleAdjusterB200 ba(IterativeOptions(), &sc.rec);
    const TrackId t0 = sc.tracks[0];
    ba.AddTrack(t0);                       // every view of t0 enters with constant extrinsics (:166-168), t0 variable (:178)
    const ViewId v_late = *sc.rec.MutableTrack(t0)->ViewIds().begin();
    ba.AddView(v_late);                    // the constant block stays constant; AddView re-freezes every track of the view (:137), t0 included
    ba.AddTrack(sc.tracks[1]);             // tracks[1] variable again even if v_late observes it
    BundleAdjusterB200::Flat f; tba_options o;
    ba.Flatten(&f, &o);
    bool saw_late = false;
    for (size_t i = 0; i < f.view_of_cam.size(); ++i) if (f.view_of_cam[i] == v_late) { saw_late = true; EXPECT(f.ext_const[i] == TBA_EXT_ALL_CONST); }
    EXPECT(saw_late);
    size_t n_t0 = 0;
    for (size_t q = 0; q < f.track_of_pt.size(); ++q) {
      if (f.track_of_pt[q] == t0) { EXPECT(f.pt_const[q] == 1); ++n_t0; }
      if (f.track_of_pt[q] == sc.tracks[1]) EXPECT(f.pt_const[q] == 0);
    }
    EXPECT(n_t0 == 1);
    // the residual (v_late, t0) was added twice -- once by AddTrack, once by AddView -- exactly as the reference's problem holds it twice
    size_t dup = 0;
    for (size_t k = 0; k < f.obs_cam.size(); ++k) dup += f.view_of_cam[f.obs_cam[k]] == v_late && f.track_of_pt[f.obs_pt[k]] == t0;
    EXPECT(dup == 2);
  }
  std::printf("flatten ok\n");
  return 0;
}

typedef int (*oracle_solve_fn)(const tba_options*, tba_problem*, tba_summary*);

static int TestSolve(const char* oracle_path) {
  Scene sc;
  BuildScene(&sc, 12, 400, 5, 5, 21);
  // an option set the engine does not implement (CGNR) is refused loudly, parameters untouched
  {
    Scene copy = sc;
    const double before = copy.rec.MutableTrack(copy.tracks[3])->Point().v[0];
    BundleAdjustmentOptions unsupported;
    unsupported.linear_solver_type = ceres::CGNR;
    BundleAdjustmentSummary s = BundleAdjustReconstructionB200(unsupported, &copy.rec);
    EXPECT(!s.success);
    EXPECT(copy.rec.MutableTrack(copy.tracks[3])->Point().v[0] == before);
  }
  // oracle on the SAME flattening (built from `sc` itself: a copied Reconstruction iterates its hash maps in another order, and the
  // inexact PCG + function-tolerance stop make the final cost order-dependent at the 1e-5 level); Flatten copies the parameters
  double oracle_final = -1, oracle_initial = -1;
  {
    BundleAdjusterB200 ba(IterativeOptions(), &sc.rec);
    for (ViewId v : sc.rec.ViewIds()) ba.AddView(v);
    for (TrackId t : sc.rec.TrackIds()) ba.AddTrack(t);
    BundleAdjusterB200::Flat f; tba_options o;
    ba.Flatten(&f, &o);
    void* h = dlopen(oracle_path, RTLD_NOW);
    EXPECT(h != nullptr);
    oracle_solve_fn solve = (oracle_solve_fn)dlsym(h, "oracle_solve");
    EXPECT(solve != nullptr);
    tba_problem p = f.AsProblem();
    tba_summary s; std::memset(&s, 0, sizeof s);
    EXPECT(solve(&o, &p, &s) == 0 && s.success);
    oracle_final = s.final_cost; oracle_initial = s.initial_cost;
  }
  BundleAdjustmentSummary s = BundleAdjustReconstructionB200(IterativeOptions(), &sc.rec);
  EXPECT(s.success);
  EXPECT(std::fabs(s.initial_cost - oracle_initial) <= 1e-10 * oracle_initial);
  // 42 LM iterations / 1000 CG iterations on this scene: rounding-level differences between two runs (another reduction order,
  // FMA contraction on the GPU) can move the stopping point -- same minimum, cost to 1e-3
  EXPECT(std::fabs(s.final_cost - oracle_final) <= 1e-3 * oracle_final);
  EXPECT(s.final_cost < 0.05 * s.initial_cost);
  // Theia's DEFAULT options (SPARSE_SCHUR + inner iterations, bundle_adjustment.h:78-122) run as they are
  {
    Scene copy0;  // a fresh scene (same seed): copies of a Reconstruction SHARE their CameraIntrinsicsModel objects (camera.cc:78-87)
    BuildScene(&copy0, 12, 400, 5, 5, 21);
    BundleAdjustmentOptions defaults;
    defaults.max_num_iterations = 15;
    BundleAdjustmentSummary sd = BundleAdjustReconstructionB200(defaults, &copy0.rec);
    EXPECT(sd.success && sd.final_cost < 0.05 * sd.initial_cost);
    EXPECT(std::fabs(sd.final_cost - oracle_final) <= 2e-2 * oracle_final);  // same minimum, different path
  }
  // the in-place update happened: reprojection with the refined parameters reproduces final_cost
  double cost = 0;
  for (TrackId t : sc.rec.TrackIds()) {
    Track* tr = sc.rec.MutableTrack(t);
    for (ViewId v : tr->ViewIds()) {
      View* view = sc.rec.MutableView(v);
      double pix[2];
      Project(view->MutableCamera()->extrinsics(), view->MutableCamera()->intrinsics(), tr->Point().data(), pix);
      const Feature* f = view->GetFeature(t);
      cost += 0.5 * ((pix[0] - f->x()) * (pix[0] - f->x()) + (pix[1] - f->y()) * (pix[1] - f->y()));
    }
  }
  EXPECT(std::fabs(cost - s.final_cost) <= 1e-9 * cost);
  // partial BA: two views + their tracks; everything else bit-identical
  {
    Scene copy = sc;
    std::unordered_set<ViewId> vs = {copy.views[2], copy.views[7]};
    std::unordered_set<TrackId> ts;
    for (ViewId v : vs) for (TrackId t : copy.rec.MutableView(v)->TrackIds()) ts.insert(t);
    for (ViewId v : vs) { double* e = copy.rec.MutableView(v)->MutableCamera()->mutable_extrinsics(); e[0] += 0.05; e[4] -= 0.01; }
    BundleAdjustmentSummary ps = BundleAdjustPartialReconstructionB200(IterativeOptions(), vs, ts, &copy.rec);
    EXPECT(ps.success && ps.final_cost < ps.initial_cost);
    for (ViewId v : copy.views) {
      if (vs.count(v)) continue;
      EXPECT(std::memcmp(copy.rec.MutableView(v)->MutableCamera()->extrinsics(), sc.rec.MutableView(v)->MutableCamera()->extrinsics(), 48) == 0);
    }
    for (TrackId t : copy.tracks) {
      if (ts.count(t)) continue;
      EXPECT(std::memcmp(copy.rec.MutableTrack(t)->Point().data(), sc.rec.MutableTrack(t)->Point().data(), 32) == 0);
    }
  }
  // N1: post-BA outlier filter on the device-resident problem (set_outlier_tracks_to_unestimated.cc:62-136)
  {
    Scene copy = sc;
    // one gross outlier track: shift all its measurements
    const TrackId bad_track = copy.tracks[10];
    for (ViewId v : copy.rec.MutableTrack(bad_track)->ViewIds()) {
      const Feature* f = copy.rec.MutableView(v)->GetFeature(bad_track);
      copy.rec.MutableView(v)->AddFeature(bad_track, Feature(f->x() + 80.0, f->y() - 60.0));
    }
    BundleAdjustmentOptions o = IterativeOptions();
    o.max_num_iterations = 0;  // evaluate only: parameters stay at the refined solution
    BundleAdjusterB200 ba(o, &copy.rec);
    for (ViewId v : copy.rec.ViewIds()) ba.AddView(v);
    for (TrackId t : copy.rec.TrackIds()) ba.AddTrack(t);
    EXPECT(ba.SetOutlierTracksToUnestimated(5.0, 0.0) == -1);  // nothing resident before Optimize()
    EXPECT(ba.Optimize().success);
    const int removed = ba.SetOutlierTracksToUnestimated(5.0, 0.0);
    EXPECT(removed == 1);
    EXPECT(!copy.rec.MutableTrack(bad_track)->IsEstimated());
    int still = 0;
    for (TrackId t : copy.tracks) still += copy.rec.MutableTrack(t)->IsEstimated();
    EXPECT(still == (int)copy.tracks.size() - 1);
    EXPECT(ba.SetOutlierTracksToUnestimated(5.0, 179.0) == (int)copy.tracks.size() - 1);  // impossible angle: every estimated track goes
  }
  std::printf("solve ok: cost %.6e -> %.6e (oracle %.6e), setup %.3f s, solve %.3f s\n", s.initial_cost, s.final_cost, oracle_final, s.setup_time_in_seconds, s.solve_time_in_seconds);
  return 0;
}

// N3: TrackEstimatorB200 (drop-in for estimate_track.h).  Without a GPU: loud refusal, reconstruction untouched.
static int TestTracksNoGpu() {
  Scene sc;
  BuildScene(&sc, 8, 60, 4, 3, 5);
  for (TrackId t : sc.tracks) sc.rec.MutableTrack(t)->SetEstimated(false);
  sc.rec.MutableTrack(sc.tracks[0])->SetEstimated(true);
  const double before = sc.rec.MutableTrack(sc.tracks[7])->Point().v[1];
  TrackEstimatorB200::Options o;
  TrackEstimatorB200 est(o, &sc.rec);
  TrackEstimatorB200::Summary s = est.EstimateAllTracks();
  EXPECT(s.input_num_estimated_tracks == 1 && s.num_triangulation_attempts == 59);
  if (tba_device_count() <= 0) {
    EXPECT(!est.engine_ok() && s.estimated_tracks.empty());
    EXPECT(sc.rec.MutableTrack(sc.tracks[7])->Point().v[1] == before && !sc.rec.MutableTrack(sc.tracks[7])->IsEstimated());
  }
  // nothing to do: no engine call at all
  for (TrackId t : sc.tracks) sc.rec.MutableTrack(t)->SetEstimated(true);
  TrackEstimatorB200 est2(o, &sc.rec);
  s = est2.EstimateAllTracks();
  EXPECT(s.num_triangulation_attempts == 0 && s.input_num_estimated_tracks == 60 && est2.engine_ok());
  std::printf("tracks-nogpu ok\n");
  return 0;
}

typedef int (*oracle_estimate_fn)(const tba_options*, tba_problem*, double, double, int, uint8_t*, int32_t*);

static int TestTracks(const char* oracle_path) {
  Scene sc;
  BuildScene(&sc, 12, 400, 5, 5, 21);
  EXPECT(BundleAdjustReconstructionB200(IterativeOptions(), &sc.rec).success);  // consistent cameras + points
  Scene refined = sc;
  // a track with a gross outlier measurement, an un-estimated view (its features are skipped), an already estimated track
  const TrackId outlier = sc.tracks[10], done = sc.tracks[30];
  {
    const ViewId v = *sc.rec.MutableTrack(outlier)->ViewIds().begin();
    const Feature* f = sc.rec.MutableView(v)->GetFeature(outlier);
    sc.rec.MutableView(v)->AddFeature(outlier, Feature(f->x() + 300.0, f->y()));
  }
  const ViewId unestimated_view = sc.views[11];
  sc.rec.MutableView(unestimated_view)->SetEstimated(false);
  for (TrackId t : sc.tracks) {
    Track* tr = sc.rec.MutableTrack(t);
    tr->SetEstimated(t == done);
    if (t != done) for (int j = 0; j < 4; ++j) (*tr->MutablePoint())[j] = 7.0 + j;  // must be ignored
  }
  // oracle on an independent flattening of the same job
  std::vector<double> ext, intr, pt, xy; std::vector<uint8_t> ec, pc; std::vector<int32_t> cg, gm, oc, op; std::vector<uint32_t> mk;
  std::vector<TrackId> order;
  {
    std::vector<ViewId> vs;
    for (ViewId v : sc.views) if (sc.rec.MutableView(v)->IsEstimated()) vs.push_back(v);
    for (size_t i = 0; i < vs.size(); ++i) {
      Camera* cam = sc.rec.MutableView(vs[i])->MutableCamera();
      for (int j = 0; j < 6; ++j) ext.push_back(cam->extrinsics()[j]);
      ec.push_back(TBA_EXT_ALL_CONST); cg.push_back((int32_t)i); gm.push_back(TBA_MODEL_PINHOLE); mk.push_back(0x7F);
      for (int j = 0; j < TBA_INTR_STRIDE; ++j) intr.push_back(j < 7 ? cam->intrinsics()[j] : 0.0);
    }
    for (TrackId t : sc.tracks) {
      if (t == done) continue;
      Track* tr = sc.rec.MutableTrack(t);
      for (int j = 0; j < 4; ++j) pt.push_back(tr->Point().data()[j]);
      pc.push_back(0);
      for (size_t i = 0; i < vs.size(); ++i) {
        if (!tr->ViewIds().count(vs[i])) continue;
        const Feature* f = sc.rec.MutableView(vs[i])->GetFeature(t);
        oc.push_back((int32_t)i); op.push_back((int32_t)order.size()); xy.push_back(f->x()); xy.push_back(f->y());
      }
      order.push_back(t);
    }
  }
  tba_problem p; std::memset(&p, 0, sizeof p);
  p.n_cam = (int32_t)ec.size(); p.ext = ext.data(); p.ext_const = ec.data(); p.cam_group = cg.data();
  p.n_group = p.n_cam; p.group_model = gm.data(); p.intr = intr.data(); p.group_const_mask = mk.data();
  p.n_pt = (int32_t)order.size(); p.pt = pt.data(); p.pt_const = pc.data(); p.n_obs = (int64_t)oc.size();
  p.obs_cam = oc.data(); p.obs_pt = op.data(); p.obs_xy = xy.data();
  void* h = dlopen(oracle_path, RTLD_NOW);
  EXPECT(h != nullptr);
  oracle_estimate_fn oracle_estimate = (oracle_estimate_fn)dlsym(h, "oracle_estimate_tracks");
  EXPECT(oracle_estimate != nullptr);
  tba_options oo; tba_options_init(&oo); oo.use_inner_iterations = 0; oo.linear_solver_type = TBA_ITERATIVE_SCHUR;
  std::vector<uint8_t> ost(order.size()); int32_t ocounts[5];
  EXPECT(oracle_estimate(&oo, &p, 5.0, 3.0, 1, ost.data(), ocounts) == 0);

  TrackEstimatorB200::Options o;
  o.ba_options = IterativeOptions();
  TrackEstimatorB200 est(o, &sc.rec);
  TrackEstimatorB200::Summary s = est.EstimateAllTracks();
  EXPECT(est.engine_ok());
  EXPECT(s.input_num_estimated_tracks == 1 && s.num_triangulation_attempts == (int)order.size());
  EXPECT((int)s.estimated_tracks.size() == ocounts[0] && est.num_bad_reprojections() == ocounts[4] && est.num_bad_angles() == ocounts[1]);
  EXPECT(ocounts[0] >= (int)order.size() - 5 && ocounts[4] >= 1);
  EXPECT(!s.estimated_tracks.count(outlier) && !sc.rec.MutableTrack(outlier)->IsEstimated());
  for (size_t q = 0; q < order.size(); ++q) {
    Track* tr = sc.rec.MutableTrack(order[q]);
    EXPECT(tr->IsEstimated() == (ost[q] == 0) && (s.estimated_tracks.count(order[q]) == 1) == (ost[q] == 0));
    if (ost[q] != 0) continue;
    const double* a = tr->Point().data();
    const double* b = &pt[q * 4];
    const double* r = refined.rec.MutableTrack(order[q])->Point().data();
    for (int j = 0; j < 3; ++j) {
      EXPECT(std::fabs(a[j] / a[3] - b[j] / b[3]) <= 1e-6 * (1.0 + std::fabs(b[j] / b[3])));   // = oracle
      EXPECT(std::fabs(a[j] / a[3] - r[j] / r[3]) <= 0.2);                                      // ~ the jointly refined point
    }
  }
  // cameras untouched
  for (ViewId v : sc.views) EXPECT(std::memcmp(sc.rec.MutableView(v)->MutableCamera()->extrinsics(), refined.rec.MutableView(v)->MutableCamera()->extrinsics(), 48) == 0);
  std::printf("tracks ok: %d of %d estimated, %d bad reprojections, %d bad angles\n", (int)s.estimated_tracks.size(), s.num_triangulation_attempts,
              est.num_bad_reprojections(), est.num_bad_angles());
  return 0;
}

// N3: BundleAdjustViewB200 / BundleAdjustTrackB200 (bundle_adjustment.cc:82-107) against the oracle on the same flattening.
static int TestMicro(const char* oracle_path, bool gpu) {
  Scene sc;
  BuildScene(&sc, 12, 400, 5, 5, 23);
  if (gpu) EXPECT(BundleAdjustReconstructionB200(IterativeOptions(), &sc.rec).success);
  void* h = dlopen(oracle_path, RTLD_NOW);
  EXPECT(h != nullptr);
  oracle_solve_fn solve = (oracle_solve_fn)dlsym(h, "oracle_solve");
  EXPECT(solve != nullptr);
  BundleAdjustmentOptions o;  // Theia defaults: the free functions override solver type and inner iterations themselves
  o.intrinsics_to_optimize = OptimizeIntrinsicsType::NONE;
  // ---- one view: pose disturbed, all its tracks constant
  {
    const ViewId v = sc.views[4];
    double* e = sc.rec.MutableView(v)->MutableCamera()->mutable_extrinsics();
    e[0] += 0.08; e[1] -= 0.05; e[3] += 0.01; e[5] -= 0.008;
    Scene copy = sc;
    BundleAdjustmentOptions oo = o; oo.linear_solver_type = ceres::DENSE_QR; oo.use_inner_iterations = false;
    BundleAdjusterB200 ba(oo, &copy.rec);
    ba.AddView(v);
    BundleAdjusterB200::Flat f; tba_options to;
    ba.Flatten(&f, &to);
    EXPECT(f.view_of_cam.size() == 1 && f.ext_const[0] == 0);
    for (uint8_t c : f.pt_const) EXPECT(c == 1);
    tba_problem p = f.AsProblem();
    tba_summary os; std::memset(&os, 0, sizeof os);
    EXPECT(solve(&to, &p, &os) == 0 && os.success);
    EXPECT(os.final_cost < 0.5 * os.initial_cost);
    const double before = sc.rec.MutableTrack(sc.tracks[0])->Point().v[0];
    if (gpu) {
    BundleAdjustmentSummary s = BundleAdjustViewB200(o, v, &sc.rec);
    EXPECT(s.success);
    EXPECT(std::fabs(s.initial_cost - os.initial_cost) <= 1e-10 * os.initial_cost);
    EXPECT(std::fabs(s.final_cost - os.final_cost) <= 1e-6 * os.final_cost);
    EXPECT(s.final_cost < 0.5 * s.initial_cost);
    const double* got = sc.rec.MutableView(v)->MutableCamera()->extrinsics();
    for (int j = 0; j < 6; ++j) EXPECT(std::fabs(got[j] - f.ext[j]) <= 1e-6 * (1.0 + std::fabs(f.ext[j])));
    EXPECT(sc.rec.MutableTrack(sc.tracks[0])->Point().v[0] == before);  // tracks constant
    }
  }
  // ---- one track: point disturbed, all its views constant
  {
    const TrackId t = sc.tracks[17];
    Track* tr = sc.rec.MutableTrack(t);
    (*tr->MutablePoint())[0] += 0.2; (*tr->MutablePoint())[2] -= 0.3;
    Scene copy = sc;
    BundleAdjustmentOptions oo = o; oo.linear_solver_type = ceres::DENSE_QR; oo.use_inner_iterations = false;
    BundleAdjusterB200 ba(oo, &copy.rec);
    ba.AddTrack(t);
    BundleAdjusterB200::Flat f; tba_options to;
    ba.Flatten(&f, &to);
    EXPECT(f.track_of_pt.size() == 1 && f.pt_const[0] == 0);
    for (uint8_t c : f.ext_const) EXPECT(c == TBA_EXT_ALL_CONST);
    tba_problem p = f.AsProblem();
    tba_summary os; std::memset(&os, 0, sizeof os);
    EXPECT(solve(&to, &p, &os) == 0 && os.success);
    const ViewId v0 = *tr->ViewIds().begin();
    double ext_before[6];
    std::memcpy(ext_before, sc.rec.MutableView(v0)->MutableCamera()->extrinsics(), 48);
    EXPECT(os.final_cost < os.initial_cost);
    if (gpu) {
    BundleAdjustmentSummary s = BundleAdjustTrackB200(o, t, &sc.rec);
    EXPECT(s.success);
    EXPECT(std::fabs(s.initial_cost - os.initial_cost) <= 1e-10 * os.initial_cost);
    EXPECT(std::fabs(s.final_cost - os.final_cost) <= 1e-6 * (1.0 + os.final_cost));
    const double* a = tr->Point().data();
    for (int j = 0; j < 3; ++j) EXPECT(std::fabs(a[j] / a[3] - f.pt[j] / f.pt[3]) <= 1e-6 * (1.0 + std::fabs(f.pt[j] / f.pt[3])));
    EXPECT(std::memcmp(ext_before, sc.rec.MutableView(v0)->MutableCamera()->extrinsics(), 48) == 0);
    }
  }
  std::printf(gpu ? "micro ok\n" : "micro-oracle ok\n");
  return 0;
}

// BundleAdjustTwoViewsB200 (bundle_adjust_two_views.cc:112-191): flattening rules on the CPU, solve vs oracle on the GPU.
static int TestTwoViews(const char* oracle_path, bool gpu) {
  std::mt19937 rng(77);
  std::normal_distribution<double> N(0.0, 1.0);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  Camera cam1, cam2;
  const double kgt[7] = {800.0, 1.0, 0.0, 500.0, 500.0, 0.0, 0.0};
  for (int j = 0; j < 7; ++j) { cam1.mutable_intrinsics()[j] = kgt[j]; cam2.mutable_intrinsics()[j] = kgt[j]; }
  const double e2gt[6] = {1.0, 0.1, -0.05, 0.02, -0.15, 0.01};
  const double e1[6] = {0, 0, 0, 0, 0, 0};
  std::vector<FeatureCorrespondence> corr;
  std::vector<TwoViewPoint> pts;
  for (int i = 0; i < 180; ++i) {
    const double X[4] = {2.0 * U(rng), 2.0 * U(rng), 7.0 + 2.0 * U(rng), 1.0};
    double p1[2], p2[2];
    Project(e1, kgt, X, p1); Project(e2gt, kgt, X, p2);
    corr.emplace_back(Feature(p1[0] + 0.3 * N(rng), p1[1] + 0.3 * N(rng)), Feature(p2[0] + 0.3 * N(rng), p2[1] + 0.3 * N(rng)));
    TwoViewPoint P; for (int j = 0; j < 3; ++j) P[j] = X[j] + 0.05 * N(rng); P[3] = 1.0;
    pts.push_back(P);
  }
  for (int j = 0; j < 6; ++j) cam2.mutable_extrinsics()[j] = e2gt[j] + (j < 3 ? 0.03 : 0.005) * N(rng);
  cam2.mutable_intrinsics()[0] = 780.0;  // wrong focal length: recovered only when camera 2's intrinsics are not constant
  void* h = dlopen(oracle_path, RTLD_NOW);
  EXPECT(h != nullptr);
  oracle_solve_fn solve = (oracle_solve_fn)dlsym(h, "oracle_solve");
  EXPECT(solve != nullptr);
  for (int variant = 0; variant < 2; ++variant) {
    TwoViewBundleAdjustmentOptions o;
    o.constant_camera2_intrinsics = variant == 0;
    Camera c1 = cam1, c2 = cam2;
    c1.MutableCameraIntrinsics().reset(new CameraIntrinsicsModel(*cam1.CameraIntrinsics()));  // private copies per variant
    c2.MutableCameraIntrinsics().reset(new CameraIntrinsicsModel(*cam2.CameraIntrinsics()));
    std::vector<TwoViewPoint> p = pts;
    BundleAdjusterB200::Flat f; tba_options to;
    FlattenTwoViewProblem(o, corr, &c1, &c2, &p, &f, &to);
    EXPECT(f.ext_const[0] == TBA_EXT_ALL_CONST && f.ext_const[1] == 0 && f.cam_group[0] == 0 && f.cam_group[1] == 1);
    EXPECT(f.group_const_mask[0] == 0x7Fu && f.group_const_mask[1] == (variant == 0 ? 0x7Fu : 0x7Eu));
    EXPECT(to.linear_solver_type == TBA_DENSE_SCHUR && to.max_num_iterations == 200 && to.use_inner_iterations == 0 && to.max_trust_region_radius == 1e16);
    EXPECT(f.obs_cam.size() == 360 && f.obs_cam[0] == 0 && f.obs_cam[1] == 1 && f.obs_pt[1] == 0 && f.obs_pt[2] == 1);
    tba_problem prob = f.AsProblem();
    tba_summary os; std::memset(&os, 0, sizeof os);
    EXPECT(solve(&to, &prob, &os) == 0 && os.success && os.final_cost < 0.2 * os.initial_cost);
    if (variant == 1) EXPECT(std::fabs(f.intr[TBA_INTR_STRIDE] - 800.0) < 8.0);  // focal length of camera 2 recovered
    else EXPECT(f.intr[TBA_INTR_STRIDE] == 780.0);
    if (!gpu) continue;
    BundleAdjustmentSummary s = BundleAdjustTwoViewsB200(o, corr, &c1, &c2, &p);
    EXPECT(s.success);
    EXPECT(std::fabs(s.initial_cost - os.initial_cost) <= 1e-10 * os.initial_cost);
    EXPECT(std::fabs(s.final_cost - os.final_cost) <= 1e-5 * os.final_cost);
    for (int j = 0; j < 6; ++j) EXPECT(std::fabs(c2.extrinsics()[j] - f.ext[6 + j]) <= 1e-4 * (1.0 + std::fabs(f.ext[6 + j])));
    EXPECT(std::fabs(c2.intrinsics()[0] - f.intr[TBA_INTR_STRIDE]) <= 1e-4 * 800.0);
    for (int j = 0; j < 6; ++j) EXPECT(c1.extrinsics()[j] == 0.0);
    EXPECT(c1.intrinsics()[0] == 800.0);
  }
  std::printf(gpu ? "twoview ok\n" : "twoview-oracle ok\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "flatten") return TestFlatten();
  if (argc >= 3 && std::string(argv[1]) == "solve") return TestSolve(argv[2]);
  if (argc >= 2 && std::string(argv[1]) == "tracks-nogpu") return TestTracksNoGpu();
  if (argc >= 3 && std::string(argv[1]) == "tracks") return TestTracks(argv[2]);
  if (argc >= 3 && std::string(argv[1]) == "micro") return TestMicro(argv[2], true);
  if (argc >= 3 && std::string(argv[1]) == "twoview") return TestTwoViews(argv[2], true);
  if (argc >= 3 && std::string(argv[1]) == "twoview-oracle") return TestTwoViews(argv[2], false);
  if (argc >= 3 && std::string(argv[1]) == "micro-oracle") return TestMicro(argv[2], false);  // CPU: flattening + oracle half only
  std::fprintf(stderr, "usage: adapter_test flatten | tracks-nogpu | solve <libba_oracle.so> | tracks <libba_oracle.so> | micro <libba_oracle.so>\n");
  return 2;
}
