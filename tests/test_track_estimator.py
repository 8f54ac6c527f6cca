"""N3: batched TrackEstimator::EstimateTrack (estimate_track.cc:199-264).  The product's device bodies
(tba_track_estimator.cuh: observation_ray = k_track_rays, estimate_track = k_estimate_tracks) run on the host over the
packed layout; the checker is oracle_estimate_tracks, which restates the reference's per-track pipeline on the caller's
layout with its own linear algebra and the oracle's LM solver."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import fountain_problem
from theiasfm_b200 import _abi, engine, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def H():
    so, src = os.path.join(HERE, "_host_point_lm.so"), os.path.join(HERE, "host_point_lm.cc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def run_host(H, p, loss=0, width=2.0, max_iters=100, max_px=5.0, min_angle=3.0, ba=True):
    k = engine.debug_pack(p)
    assert k["rc"] == 0
    npk = k["n_packed_points"]
    valid = k["slot_cam"] >= 0
    slots = np.nonzero(valid)[0]; pts = k["slot_pt"][valid]
    first = np.full(npk, -1, np.int64); cnt = np.bincount(pts, minlength=npk).astype(np.int32)
    first[pts[::-1]] = slots[::-1]
    pt_packed = np.ascontiguousarray(p.pt[k["pk2caller"]])
    status = np.zeros(npk, np.uint8)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    slot_cam = np.ascontiguousarray(k["slot_cam"]); xy = np.ascontiguousarray(k["xy"])
    H.host_estimate_tracks(p.n_cam, p.ext.ctypes.data_as(dp), p.intr.ctypes.data_as(dp), p.cam_group.ctypes.data_as(ip), p.group_model.ctypes.data_as(ip),
                           npk, C.c_longlong(len(slot_cam)), pt_packed.ctypes.data_as(dp), xy.ctypes.data_as(dp), slot_cam.ctypes.data_as(ip),
                           first.ctypes.data_as(C.POINTER(C.c_longlong)), cnt.ctypes.data_as(ip), loss, C.c_double(width), max_iters,
                           C.c_double(max_px), C.c_double(min_angle), int(ba), status.ctypes.data_as(C.POINTER(C.c_uint8)))
    out = p.pt.copy(); out[k["pk2caller"]] = pt_packed
    st = np.full(p.n_pt, 1, np.uint8); st[k["pk2caller"]] = status       # unobserved points: "fewer than 2 views"
    return out, st


def euclid(x):
    return x[:, :3] / x[:, 3:4]


def scene(seed, model, n_cam=24, n_pt=400):
    p = synthetic.make_scene(n_cam=n_cam, n_pt=n_pt, obs_per_pt=5, seed=seed, model=model, noise_px=0.5, perturb=0.0)   # cameras at their true poses
    p.ext_const[:] = _abi.EXT_ALL_CONST
    truth = p.pt.copy()
    p.pt[:] = np.random.default_rng(seed).normal(size=p.pt.shape)        # the incoming point value must be ignored
    return p, truth


@pytest.mark.parametrize("model", [_abi.MODEL_PINHOLE, _abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, _abi.MODEL_FISHEYE, _abi.MODEL_FOV,
                                   _abi.MODEL_DIVISION_UNDISTORTION])
@pytest.mark.parametrize("ba", [True, False])
def test_estimate_tracks_matches_oracle(H, oracle, model, ba):
    p, truth = scene(31, model)
    # engineered failures: an outlier observation (bad reprojection), a narrow-baseline track (bad angle), a single view
    o5 = np.nonzero(p.obs_pt == 5)[0]; p.obs_xy[o5[0]] += 400.0
    o9 = np.nonzero(p.obs_pt == 9)[0]; p.obs_cam[o9] = p.obs_cam[o9[0]]; p.obs_xy[o9] = p.obs_xy[o9[0]]   # identical rays: zero angle
    keep = np.ones(p.n_obs, bool); keep[np.nonzero(p.obs_pt == 11)[0][1:]] = False
    p = _abi.Problem(p.ext, p.ext_const, p.cam_group, p.group_model, p.intr, p.group_const_mask, p.pt, p.pt_const, p.obs_cam[keep], p.obs_pt[keep], p.obs_xy[keep])
    out, st = run_host(H, p, ba=ba)
    q = p.copy()
    st_o, counts = oracle.estimate_tracks(q, oracle.default_options(use_inner_iterations=0), bundle_adjustment=ba)
    assert np.array_equal(st, st_o), np.nonzero(st != st_o)
    assert st[5] == 4 and st[9] == 1 and st[11] == 1 and counts[0] > 0.95 * p.n_pt
    ok = st == 0
    tol = 1e-6 if ba else 1e-10                                           # LM stops on a 1e-6 function tolerance; triangulation is closed form
    assert np.abs(euclid(out[ok]) - euclid(q.pt[ok])).max() <= tol * np.abs(euclid(q.pt[ok])).max()
    assert np.array_equal(out[st == 1], p.pt[st == 1])                    # untouched, like Track::MutablePoint in the reference
    # and the estimate is the scene: triangulated + adjusted points land on the generating points (0.5 px noise)
    err = np.linalg.norm(euclid(out[ok]) - euclid(truth[ok]), axis=1)
    assert np.median(err) < 0.02


def test_estimate_tracks_on_the_reference_fountain(H, oracle):
    """Theia's own fountain-11 reconstruction: re-estimating every track from its cameras + features must give back
    (nearly) the points Theia stored, and accept the tracks its own TrackEstimator accepted (defaults 5 px / 3 deg)."""
    p, g = fountain_problem()
    q0 = p.copy(); q0.pt[:] = 0.0
    out, st = run_host(H, q0)
    assert (st == 0).mean() > 0.995
    ok = st == 0
    err = np.linalg.norm(euclid(out[ok]) - euclid(p.pt[ok]), axis=1)
    scale = np.linalg.norm(euclid(p.pt) - euclid(p.pt).mean(0), axis=1).mean()
    assert np.median(err) < 2e-3 * scale and np.percentile(err, 99) < 5e-2 * scale
    sub = np.arange(0, p.n_pt, 211)
    keep = np.isin(q0.obs_pt, sub)
    remap = -np.ones(p.n_pt, np.int64); remap[sub] = np.arange(len(sub))
    qs = _abi.Problem(q0.ext, q0.ext_const, q0.cam_group, q0.group_model, q0.intr, q0.group_const_mask, q0.pt[sub], q0.pt_const[sub],
                      q0.obs_cam[keep], remap[q0.obs_pt[keep]].astype(np.int32), q0.obs_xy[keep])
    st_o, _ = oracle.estimate_tracks(qs, oracle.default_options(use_inner_iterations=0))
    assert np.array_equal(st[sub], st_o)
    good = st_o == 0
    assert np.abs(euclid(out[sub][good]) - euclid(qs.pt[good])).max() <= 1e-6 * np.abs(euclid(qs.pt[good])).max()
