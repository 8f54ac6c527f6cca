"""N3: the per-track bundle adjustment BundleAdjustTrack performs (bundle_adjustment.cc:96-107, called per
track from estimate_track.cc:241): thousands of independent 4-parameter LM problems with every camera constant.  The
product's device body (theiasfm_b200/csrc/tba_point_lm.cuh, one instance per GPU thread in k_adjust_tracks) is compiled
for the host and run over the packed layout; the checker is the oracle solving each track as its own problem
(cameras constant, exact Schur type: the 4x4 back-substitution DENSE_QR amounts to)."""
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


def run_host(H, p, loss=0, width=2.0, max_iters=100):
    k = engine.debug_pack(p)
    assert k["rc"] == 0
    npk = k["n_packed_points"]
    valid = k["slot_cam"] >= 0
    slots = np.nonzero(valid)[0]; pts = k["slot_pt"][valid]
    first = np.full(npk, -1, np.int64); cnt = np.bincount(pts, minlength=npk).astype(np.int32)
    first[pts[::-1]] = slots[::-1]
    pt_packed = np.ascontiguousarray(p.pt[k["pk2caller"]])
    ic, fc = np.zeros(npk), np.zeros(npk); term, its = np.zeros(npk, np.int32), np.zeros(npk, np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    slot_cam = np.ascontiguousarray(k["slot_cam"]); xy = np.ascontiguousarray(k["xy"]); first = np.ascontiguousarray(first)
    H.host_adjust_tracks(p.n_cam, p.ext.ctypes.data_as(dp), p.intr.ctypes.data_as(dp), p.cam_group.ctypes.data_as(ip),
                         p.group_model.ctypes.data_as(ip), npk, pt_packed.ctypes.data_as(dp), xy.ctypes.data_as(dp), slot_cam.ctypes.data_as(ip),
                         first.ctypes.data_as(C.POINTER(C.c_longlong)), cnt.ctypes.data_as(ip), loss, C.c_double(width), max_iters,
                         ic.ctypes.data_as(dp), fc.ctypes.data_as(dp), term.ctypes.data_as(ip), its.ctypes.data_as(ip))
    out = p.pt.copy(); out[k["pk2caller"]] = pt_packed
    full = lambda a, fill: (lambda f: (f.__setitem__(k["pk2caller"], a), f)[1])(np.full(p.n_pt, fill, a.dtype))
    return out, full(ic, -1.0), full(fc, -1.0), full(term, 2), full(its, 0)


def oracle_per_track(oracle, p, tracks, loss=0, width=2.0, max_iters=100):
    res = {}
    for q in tracks:
        sel = p.obs_pt == q
        sub = _abi.Problem(p.ext, np.full(p.n_cam, _abi.EXT_ALL_CONST, np.uint8), p.cam_group, p.group_model, p.intr,
                           np.full(p.n_group, 0x3FF, np.uint32), p.pt[q:q + 1].copy(), [0], p.obs_cam[sel], np.zeros(int(sel.sum()), np.int32), p.obs_xy[sel])
        s = oracle.solve(sub, oracle.default_options(use_inner_iterations=0, linear_solver_type=_abi.DENSE_SCHUR, loss_function_type=loss,
                                                     robust_loss_width=width, max_num_iterations=max_iters))
        res[q] = (sub.pt[0].copy(), s)
    return res


@pytest.mark.parametrize("loss,model", [(_abi.LOSS_TRIVIAL, 0), (_abi.LOSS_HUBER, 0), (_abi.LOSS_HUBER, _abi.MODEL_FISHEYE),
                                        (_abi.LOSS_TRIVIAL, _abi.MODEL_FOV), (_abi.LOSS_CAUCHY, _abi.MODEL_DIVISION_UNDISTORTION)])
def test_point_lm_matches_per_track_oracle(H, oracle, loss, model):
    p = synthetic.make_scene(n_cam=30, n_pt=300, obs_per_pt=6, seed=19, model=model)   # points perturbed, cameras perturbed (held constant)
    if loss:
        p.obs_xy[::23] += 35.0
    p.pt[5, :3] = p.ext[p.obs_cam[p.obs_pt == 5][0], :3]; p.pt[5, 3] = 1.0  # on a camera centre: evaluation fails
    out, ic, fc, term, its = run_host(H, p, loss=loss)
    ref = oracle_per_track(oracle, p, range(0, 300, 7), loss=loss)
    for q, (pt_o, s) in ref.items():
        assert abs(ic[q] - s.initial_cost) <= 1e-11 * max(1.0, s.initial_cost)
        assert abs(fc[q] - s.final_cost) <= 1e-7 * max(1.0, s.final_cost), (q, fc[q], s.final_cost, s.message)
        # the homogeneous scale is a gauge direction (cost-invariant): compare the Euclidean point tightly, the 4-vector loosely
        assert np.abs(out[q, :3] / out[q, 3] - pt_o[:3] / pt_o[3]).max() <= 1e-6 * np.abs(pt_o[:3] / pt_o[3]).max()
        assert np.abs(out[q] - pt_o).max() <= 1e-3 * np.abs(pt_o).max()
        assert term[q] == s.termination_type
    so5 = oracle_per_track(oracle, p, [5], loss=loss)[5][1]
    assert term[5] == _abi.FAILURE and not so5.success and np.array_equal(out[5], p.pt[5])
    assert (fc[np.arange(300) != 5] <= ic[np.arange(300) != 5] + 1e-12).all()


def test_point_lm_on_the_reference_fountain_tracks(H, oracle):
    p, g = fountain_problem()
    rng = np.random.default_rng(6)
    q0 = p.copy()
    q0.pt[:, :3] += 0.01 * rng.normal(size=(p.n_pt, 3))          # disturb the triangulated points, cameras stay
    out, ic, fc, term, its = run_host(H, q0)
    assert (term != _abi.FAILURE).all() and (fc < ic).mean() > 0.99
    # every track returns to (nearly) the reference's point: its cameras are the reference's BA result
    err = np.linalg.norm(out[:, :3] / out[:, 3:4] - p.pt[:, :3] / p.pt[:, 3:4], axis=1)
    assert np.median(err) < 2e-3 and np.percentile(err, 99) < 3e-2
    ref = oracle_per_track(oracle, q0, range(0, p.n_pt, 997))
    for q, (pt_o, s) in ref.items():
        assert abs(fc[q] - s.final_cost) <= 1e-7 * max(1.0, s.final_cost)
        assert np.abs(out[q, :3] / out[q, 3] - pt_o[:3] / pt_o[3]).max() <= 1e-6 * np.abs(pt_o[:3] / pt_o[3]).max()
