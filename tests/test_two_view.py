"""N3: batched BundleAdjustTwoViews (bundle_adjust_two_views.cc:112-191).  The product's per-pair body
(theiasfm_b200/csrc/tba_two_view.cuh; on the GPU one warp per pair, k_two_view_ba) is compiled for the host (one-lane team) and run over
the batch layout; the checker is the oracle solving every pair as its own two-camera problem with the exact solver."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from theiasfm_b200 import _abi, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def H():
    so, src = os.path.join(HERE, "_host_two_view.so"), os.path.join(HERE, "host_two_view.cc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def run_host(H, batch):
    n = batch.n_pairs
    term = np.zeros(n, np.uint8); ic = np.zeros(n); fc = np.zeros(n); it = np.zeros(n, np.int32)
    st = batch.as_struct()
    dp = C.POINTER(C.c_double)
    H.host_two_view_ba_batch(C.byref(st), term.ctypes.data_as(C.POINTER(C.c_uint8)), ic.ctypes.data_as(dp), fc.ctypes.data_as(dp),
                             it.ctypes.data_as(C.POINTER(C.c_int32)))
    return term, ic, fc, it


def run_host_team(H, batch):
    """The same body as a 4-lane team (host threads + barriers standing for the warp's shuffles)."""
    n = batch.n_pairs
    term = np.zeros(n, np.uint8); ic = np.zeros(n); fc = np.zeros(n); it = np.zeros(n, np.int32)
    st = batch.as_struct()
    dp = C.POINTER(C.c_double)
    H.host_two_view_ba_batch_team(C.byref(st), term.ctypes.data_as(C.POINTER(C.c_uint8)), ic.ctypes.data_as(dp), fc.ctypes.data_as(dp),
                                  it.ctypes.data_as(C.POINTER(C.c_int32)), 1)
    return term, ic, fc, it


def compare(b_host, res_host, b_or, res_or):
    th, ich, fch, ith = res_host
    to, ico, fco, ito = res_or
    assert np.array_equal(th, to)
    assert np.allclose(ich, ico, rtol=1e-11) and np.allclose(fch, fco, rtol=1e-7)
    assert np.abs(ith - ito).max() <= 1                      # the oracle counts pushed iteration rows, the body its loop index
    assert np.abs(b_host.ext2 - b_or.ext2).max() <= 1e-6 * np.abs(b_or.ext2).max()
    assert np.abs(b_host.intr2[:, 0] - b_or.intr2[:, 0]).max() <= 1e-6 * 800.0
    assert np.abs(b_host.intr1[:, 0] - b_or.intr1[:, 0]).max() <= 1e-6 * 800.0
    eh = b_host.points[:, :3] / b_host.points[:, 3:4]; eo = b_or.points[:, :3] / b_or.points[:, 3:4]
    assert np.abs(eh - eo).max() <= 1e-5 * np.abs(eo).max()


def test_two_view_batch_matches_per_pair_oracle(H, oracle):
    b = synthetic.make_two_view_batch(24, min_corr=40, max_corr=160, seed=3)
    b.xy2[::17] += 6.0                                     # a few wrong matches: they must fail the post-BA inlier test
    b.final_max_reprojection_error_pixels = 2.0            # BundleAdjustRelativePose's final filter (:294-312)
    start = b.copy()
    bh, bo = b.copy(), b.copy()
    rh = run_host(H, bh)
    ro = oracle.two_view_ba_batch(bo)
    compare(bh, rh, bo, ro)
    th, ich, fch, ith = rh
    assert (th == _abi.CONVERGENCE).all() and (fch < 0.5 * ich).all()
    assert np.array_equal(bh.inlier, bo.inlier) and 0.85 < bh.inlier.mean() < 0.97 and bh.inlier[::17].mean() < 0.4
    # camera 1 never moves, constant intrinsics are bit-identical, free focal lengths move towards the truth
    assert np.array_equal(bh.ext1, start.ext1)
    c2 = start.const2 == 1
    assert np.array_equal(bh.intr2[c2], start.intr2[c2]) and np.array_equal(bh.intr1[start.const1 == 1], start.intr1[start.const1 == 1])
    assert np.array_equal(bh.intr2[:, 1:], start.intr2[:, 1:])       # only the focal length is ever free (.cc:96-108)
    assert (~c2).sum() >= 5 and (bh.intr2[~c2, 0] != start.intr2[~c2, 0]).all()


@pytest.mark.parametrize("model", [_abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, _abi.MODEL_FISHEYE, _abi.MODEL_FOV, _abi.MODEL_DIVISION_UNDISTORTION])
def test_two_view_batch_other_camera_models(H, oracle, model):
    b = synthetic.make_two_view_batch(6, min_corr=50, max_corr=120, seed=10 + model, models=(model,))
    bh, bo = b.copy(), b.copy()
    compare(bh, run_host(H, bh), bo, oracle.two_view_ba_batch(bo))


def test_degenerate_pairs(H, oracle):
    """A pair whose point sits on camera 2's centre fails at the initial evaluation (FAILURE, nothing touched); an empty pair
    converges trivially; neither disturbs its neighbours."""
    b = synthetic.make_two_view_batch(4, min_corr=30, max_corr=50, seed=5)
    o1 = int(b.pair_off[1])
    b.points[o1, :3] = b.ext2[1, :3]; b.points[o1, 3] = 1.0
    keep = np.ones(len(b.points), bool); keep[b.pair_off[2]:b.pair_off[3]] = False
    n = np.diff(b.pair_off); n[2] = 0
    b2 = _abi.TwoViewBatch(np.concatenate([[0], np.cumsum(n)]), b.ext1, b.ext2, b.intr1, b.intr2, b.model1, b.model2, b.const1, b.const2,
                           b.xy1[keep], b.xy2[keep], b.points[keep])
    start = b2.copy()
    bh, bo = b2.copy(), b2.copy()
    rh, ro = run_host(H, bh), oracle.two_view_ba_batch(bo)
    assert rh[0][1] == _abi.FAILURE and ro[0][1] == _abi.FAILURE
    s1, e1 = int(b2.pair_off[1]), int(b2.pair_off[2])
    assert np.array_equal(bh.points[s1:e1], start.points[s1:e1]) and np.array_equal(bh.ext2[1], start.ext2[1])
    assert rh[0][0] == _abi.CONVERGENCE and rh[0][3] == _abi.CONVERGENCE
    for p in (0, 3):
        assert abs(rh[2][p] - ro[2][p]) <= 1e-7 * ro[2][p]


def test_team_decomposition_matches_the_serial_body(H):
    """k_two_view_ba spreads one pair over the 32 lanes of a warp; the decomposition (strided point loops, all-reduced sums and
    flags, per-lane copies of the camera values, lane 0 writes back) is run here as a 4-lane team of host threads: every lane
    takes the same decisions, and the result equals the one-lane run up to the order of the floating-point sums."""
    b = synthetic.make_two_view_batch(5, min_corr=5, max_corr=60, seed=21, models=(_abi.MODEL_PINHOLE, _abi.MODEL_FISHEYE))
    o1 = int(b.pair_off[1])
    b.points[o1, :3] = b.ext2[1, :3]; b.points[o1, 3] = 1.0          # pair 1 fails at the initial evaluation (in one lane only)
    bs, bt = b.copy(), b.copy()
    ts, ics, fcs, its = run_host(H, bs)
    tt, ict, fct, itt = run_host_team(H, bt)
    assert np.array_equal(ts, tt) and ts[1] == _abi.FAILURE and (tt != 99).all()
    ok = ts != _abi.FAILURE
    assert np.allclose(ics[ok], ict[ok], rtol=1e-12) and np.allclose(fcs[ok], fct[ok], rtol=1e-9) and np.array_equal(its, itt)
    assert np.abs(bs.ext2 - bt.ext2).max() <= 1e-9 and np.abs(bs.points - bt.points).max() <= 1e-8 * np.abs(bs.points).max()
    assert np.abs(bs.intr2 - bt.intr2).max() <= 1e-9 * 800
