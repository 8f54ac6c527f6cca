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


def test_device_filter_body_on_packed_layout_matches_oracle(oracle):
    """theiasfm_b200/csrc/tba_filter.cuh (what k_filter_tracks runs per thread) compiled for the host and executed over the
    packed tile layout of tba_debug_pack: statuses and statistics equal the oracle's, on ragged synthetic tracks with
    outliers and on the reference's fountain reconstruction."""
    import ctypes as C
    import os
    import subprocess
    from theiasfm_b200 import engine
    here = os.path.dirname(os.path.abspath(__file__))
    so, src = os.path.join(here, "_host_filter.so"), os.path.join(here, "host_filter.cc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

    def run(p, max_err, angle):
        k = engine.debug_pack(p)
        assert k["rc"] == 0
        npk = k["n_packed_points"]
        valid = k["slot_cam"] >= 0
        slots = np.nonzero(valid)[0]
        pts = k["slot_pt"][valid]
        first = np.full(npk, -1, np.int64); cnt = np.bincount(pts, minlength=npk).astype(np.int32)
        first[pts[::-1]] = slots[::-1]
        pt_packed = np.ascontiguousarray(p.pt[k["pk2caller"]])
        status = np.zeros(max(npk, 1), np.uint8); mean = np.zeros(max(npk, 1))
        slot_cam = np.ascontiguousarray(k["slot_cam"]); xy = np.ascontiguousarray(k["xy"]); first = np.ascontiguousarray(first)
        L.host_filter(p.n_cam, p.ext.ctypes.data_as(dp), p.intr.ctypes.data_as(dp), p.cam_group.ctypes.data_as(ip), p.group_model.ctypes.data_as(ip),
                      npk, pt_packed.ctypes.data_as(dp), xy.ctypes.data_as(dp), slot_cam.ctypes.data_as(ip),
                      first.ctypes.data_as(C.POINTER(C.c_longlong)), cnt.ctypes.data_as(ip), C.c_double(max_err), C.c_double(angle),
                      status.ctypes.data_as(C.POINTER(C.c_ubyte)), mean.ctypes.data_as(dp))
        full = np.full(p.n_pt, 2, np.uint8); full[k["pk2caller"]] = status[:npk]
        fmean = np.full(p.n_pt, np.nan); fmean[k["pk2caller"]] = mean[:npk]
        return full, fmean

    p = synthetic.make_scene(n_cam=60, n_pt=900, obs_per_pt=40, seed=5, perturb=0.3)
    rng = np.random.default_rng(4)
    target = rng.choice([2, 3, 9, 31, 32, 33, 40], size=p.n_pt)
    seen = np.zeros(p.n_pt, int); keep = np.zeros(p.n_obs, bool)
    for i in range(p.n_obs):
        seen[p.obs_pt[i]] += 1
        keep[i] = seen[p.obs_pt[i]] <= target[p.obs_pt[i]]
    keep[p.obs_pt == 5] = False
    perm = rng.permutation(int(keep.sum()))
    p = _abi.Problem(p.ext, p.ext_const, p.cam_group, p.group_model, p.intr, p.group_const_mask, p.pt, p.pt_const,
                     p.obs_cam[keep][perm], p.obs_pt[keep][perm], p.obs_xy[keep][perm])
    p.obs_xy[rng.choice(p.n_obs, 200, replace=False)] += 40.0
    p.pt[9, 3] = -1.0
    for prob, cases in ((p, [(5.0, 3.0), (15.0, 25.0)]), (fountain_problem()[0], [(5.0, 3.0), (1.0, 8.0)])):
        for max_err, angle in cases:
            st, mean = run(prob, max_err, angle)
            st_o, mean_o, _ = oracle.filter_tracks(prob, max_err, angle)
            ok = np.isfinite(mean_o)
            assert np.allclose(mean[ok], mean_o[ok], rtol=1e-10)
            borderline = ok & (np.abs(mean_o - max_err ** 2) <= 1e-9 * max_err ** 2)
            assert np.array_equal(st[~borderline], st_o[~borderline])
    assert (st_o == 0).sum() > 1000
