"""The product's device-side math (theiasfm_b200/csrc/tba_camera_models.cuh: cam_prep, linearize_obs, reproject,
loss_evaluate) compiled for the HOST (tests/host_models.cc) and checked, without a GPU, against the committed
torch.func.jacfwd golden vectors and against the oracle.  The same source runs on the GPU, where test_gpu_parity.py
checks it through the C-ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import golden_problem

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def H():
    so = os.path.join(HERE, "_host_models.so")
    src = os.path.join(HERE, "host_models.cc")
    hdr = os.path.join(HERE, "..", "theiasfm_b200", "csrc", "tba_camera_models.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.host_linearize.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, C.c_double, dp, dp, dp]
    L.host_reproject.argtypes = [C.c_int, dp, dp, dp, dp, dp]
    L.host_loss.argtypes = [C.c_int, C.c_double, C.c_double, dp]
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_analytic_jacobian_matches_golden(H):
    prob, g = golden_problem()
    for i in range(prob.n_obs):
        r = np.zeros(2); rho0 = C.c_double(); J = np.zeros((2, 20))
        ext, intr, pt, xy = (np.ascontiguousarray(g[k][i], np.float64) for k in ("ext", "intr", "pt", "xy"))
        ok = H.host_linearize(int(g["model"][i]), _dp(ext), _dp(intr), _dp(pt), _dp(xy), 0, 2.0, _dp(r), C.byref(rho0), _dp(J))
        assert ok == 1
        tag = str(g["tag"][i])
        scale_r = max(1.0, np.abs(g["r"][i]).max())
        assert np.abs(r - g["r"][i]).max() <= 1e-12 * scale_r, (i, tag)
        # theta ~ 2e-7: the jet (autodiff) evaluation of Rodrigues' formula cancels; the analytic form does not
        tol = 1e-7 if tag == "w_small_rodrigues" else 1e-12
        assert np.abs(J - g["J"][i]).max() <= tol * np.abs(g["J"][i]).max(), (i, tag)
        assert abs(rho0.value - (g["r"][i] ** 2).sum()) <= 1e-12 * max(1.0, (g["r"][i] ** 2).sum())


def test_small_angle_branch_is_the_exact_derivative_of_that_branch(H):
    # w = 0 and |w| ~ 1e-9 take ceres::AngleAxisRotatePoint's first-order branch: d q / d w = -[a]x exactly
    prob, g = golden_problem()
    idx = [i for i in range(prob.n_obs) if str(g["tag"][i]) in ("w_zero", "w_tiny")]
    assert len(idx) == 4
    for i in idx:
        r = np.zeros(2); rho0 = C.c_double(); J = np.zeros((2, 20))
        ext, intr, pt, xy = (np.ascontiguousarray(g[k][i], np.float64) for k in ("ext", "intr", "pt", "xy"))
        assert H.host_linearize(int(g["model"][i]), _dp(ext), _dp(intr), _dp(pt), _dp(xy), 0, 2.0, _dp(r), C.byref(rho0), _dp(J)) == 1
        assert np.abs(J[:, 3:6] - g["J"][i][:, 3:6]).max() <= 1e-13 * np.abs(g["J"][i]).max()


def test_reproject_and_failure_guard(H, oracle):
    prob, g = golden_problem()
    for i in range(0, prob.n_obs, 7):
        r = np.zeros(2)
        ext, intr, pt, xy = (np.ascontiguousarray(g[k][i], np.float64) for k in ("ext", "intr", "pt", "xy"))
        assert H.host_reproject(int(g["model"][i]), _dp(ext), _dp(intr), _dp(pt), _dp(xy), _dp(r)) == 1
        assert np.abs(r - g["r"][i]).max() <= 1e-12 * max(1.0, np.abs(g["r"][i]).max())
    ext = np.array([1.0, 2.0, 3.0, 0.1, 0.2, 0.3]); intr = np.array([800.0, 1, 0, 500, 500, 0, 0, 0, 0, 0]); xy = np.zeros(2)
    for d, expect in ((5e-5, 0), (2e-4, 1)):  # reprojection_error.h:75-77
        pt = np.array([1.0 + d, 2.0, 3.0, 1.0]); r = np.zeros(2)
        assert H.host_reproject(0, _dp(ext), _dp(intr), _dp(pt), _dp(xy), _dp(r)) == expect


def test_losses_match_oracle(H, oracle):
    for kind in range(6):
        for s in (0.0, 1e-3, 0.7, 3.99, 4.0, 4.01, 50.0, 1e4):
            rho = np.zeros(3)
            H.host_loss(kind, 2.0, s, _dp(rho))
            assert np.array_equal(rho, oracle.loss(kind, 2.0, s)), (kind, s)


def test_robust_corrector_scales_residual_and_jacobian(H):
    prob, g = golden_problem()
    i = 3
    ext, intr, pt = (np.ascontiguousarray(g[k][i], np.float64) for k in ("ext", "intr", "pt"))
    xy = np.ascontiguousarray(g["xy"][i] + g["r"][i] - np.array([30.0, -40.0]))  # residual (30, -40): |r| = 50 >> width
    r0 = np.zeros(2); J0 = np.zeros((2, 20)); rho = C.c_double()
    r1 = np.zeros(2); J1 = np.zeros((2, 20))
    H.host_linearize(int(g["model"][i]), _dp(ext), _dp(intr), _dp(pt), _dp(xy), 0, 2.0, _dp(r0), C.byref(rho), _dp(J0))
    H.host_linearize(int(g["model"][i]), _dp(ext), _dp(intr), _dp(pt), _dp(xy), 1, 2.0, _dp(r1), C.byref(rho), _dp(J1))
    s = float(r0 @ r0)
    w = np.sqrt(2.0 / np.sqrt(s))  # Huber: rho' = a / |r|, rho'' < 0 -> plain sqrt(rho') scaling (Ceres Corrector)
    assert np.allclose(r1, w * r0, rtol=1e-13) and np.allclose(J1, w * J0, rtol=1e-13)
    assert abs(rho.value - (2 * 2.0 * np.sqrt(s) - 4.0)) <= 1e-12 * rho.value
