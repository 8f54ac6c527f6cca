"""N1: SetOutlierTracksToUnestimated / ComputeStatisticsForTrack restated in the oracle
(set_outlier_tracks_to_unestimated.cc:62-136, triangulation.cc:236-250, select_good_tracks_for_bundle_adjustment.cc:79-108),
checked against a numpy restatement on synthetic scenes with injected outliers and against the reference's own saved
reconstruction (fountain-11), which must pass its own filter at Theia's default thresholds
(reconstruction_estimator_options.h:103,215: 5 px, 3 degrees)."""
import numpy as np

from helpers import fountain_problem
from theiasfm_b200 import _abi, synthetic


def _numpy_filter(p, max_err, min_angle_deg):
    pix, depth = synthetic.project(int(p.group_model[0]), p.ext[p.obs_cam], p.intr[p.cam_group[p.obs_cam]], p.pt[p.obs_pt])
    sq = ((pix - p.obs_xy) ** 2).sum(1)
    n = np.bincount(p.obs_pt, minlength=p.n_pt)
    mean = np.bincount(p.obs_pt, weights=sq, minlength=p.n_pt) / np.where(n > 0, n, np.nan)
    behind = np.bincount(p.obs_pt, weights=(depth < 0), minlength=p.n_pt) > 0
    rays = p.pt[p.obs_pt, :3] / p.pt[p.obs_pt, 3:4] - p.ext[p.obs_cam, :3]
    rays /= np.linalg.norm(rays, axis=1, keepdims=True)
    cos_min = np.cos(np.deg2rad(min_angle_deg))
    status = np.zeros(p.n_pt, np.uint8)
    order = np.argsort(p.obs_pt, kind="stable")
    off = np.concatenate([[0], np.cumsum(n)])
    for q in range(p.n_pt):
        if behind[q] or mean[q] > max_err ** 2:
            status[q] = 1
            continue
        r = rays[order[off[q]:off[q + 1]]]
        d = r @ r.T
        iu = np.triu_indices(len(r), 1)
        if not (len(r) >= 2 and (d[iu] < cos_min).any()):
            status[q] = 2
    return status, mean


def test_oracle_filter_matches_numpy_with_injected_outliers(oracle):
    p = synthetic.make_scene(n_cam=40, n_pt=600, obs_per_pt=6, seed=13, perturb=0.0)
    rng = np.random.default_rng(2)
    p.obs_xy[rng.choice(p.n_obs, 60, replace=False)] += 25.0           # gross reprojection outliers
    far = rng.choice(p.n_pt, 30, replace=False)
    for q in far:                                                       # far points on the far side of the ring, in front of
        cm = p.ext[p.obs_cam[p.obs_pt == q], :3].mean(0)                # all their cameras: tiny viewing angles ...
        p.pt[q, :3] = -2000.0 * cm / np.linalg.norm(cm)
    sel = np.isin(p.obs_pt, far)                                        # ... with consistent measurements
    pix, _ = synthetic.project(int(p.group_model[0]), p.ext[p.obs_cam[sel]], p.intr[p.cam_group[p.obs_cam[sel]]], p.pt[p.obs_pt[sel]])
    p.obs_xy[sel] = pix
    p.pt[7, 3] = -1.0                                                   # behind every camera (negative depth)
    keep = p.obs_pt != 11                                               # a point without observations
    p = _abi.Problem(p.ext, p.ext_const, p.cam_group, p.group_model, p.intr, p.group_const_mask, p.pt, p.pt_const,
                     p.obs_cam[keep], p.obs_pt[keep], p.obs_xy[keep])
    for max_err, angle in ((5.0, 3.0), (1.0, 0.5), (50.0, 10.0)):
        st, mean, removed = oracle.filter_tracks(p, max_err, angle)
        st_np, mean_np = _numpy_filter(p, max_err, angle)
        assert np.array_equal(st, st_np) and removed == int((st != 0).sum())
        ok = np.isfinite(mean_np)
        assert np.allclose(mean[ok], mean_np[ok], rtol=1e-10) and np.isnan(mean[11]) and st[11] == 2 and st[7] == 1
    st, _, _ = oracle.filter_tracks(p, 5.0, 3.0)
    assert (st == 1).sum() >= 20 and (st == 2).sum() >= 10 and (st == 0).sum() > 400


def test_reference_reconstruction_passes_its_own_filter(oracle):
    p, g = fountain_problem()
    st, mean, removed = oracle.filter_tracks(p, 5.0, 3.0)  # Theia's defaults
    assert removed <= 10 and (st == 2).sum() == 0            # 5 of 16 616 tracks in the saved file
    st_np, mean_np = _numpy_filter(p, 5.0, 3.0)
    assert np.array_equal(st, st_np) and np.allclose(mean, mean_np, rtol=1e-9)
    assert np.median(np.sqrt(mean)) < 0.5
