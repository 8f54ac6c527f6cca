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
    hdrs = [os.path.join(HERE, "..", "theiasfm_b200", "csrc", h) for h in ("tba_camera_models.cuh", "tba_camera_models_ext.cuh")]
    if not os.path.exists(so) or max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]) > os.path.getmtime(so):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.host_linearize.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, C.c_double, dp, dp, dp]
    L.host_reproject.argtypes = [C.c_int, dp, dp, dp, dp, dp]
    L.host_loss.argtypes = [C.c_int, C.c_double, C.c_double, dp]
    L.host_pixel_to_camera_ext.argtypes = [C.c_int, dp, dp, dp]
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


def test_dual_number_jacobian_of_the_other_models_matches_golden(H, oracle):
    """FISHEYE / FOV / DIVISION_UNDISTORTION (tba_camera_models_ext.cuh: templated projection + forward-mode duals) against
    the torch jacfwd vectors, every DistortPoint branch, and with a robust loss against the oracle's corrector."""
    from theiasfm_b200 import _abi
    from theiasfm_b200.synthetic import rotation_from_angle_axis
    prob, g = golden_problem(ext=True)
    for i in range(prob.n_obs):
        r = np.zeros(2); rho0 = C.c_double(); J = np.zeros((2, 20))
        ext, intr, pt, xy = (np.ascontiguousarray(g[k][i], np.float64) for k in ("ext", "intr", "pt", "xy"))
        assert H.host_linearize(int(g["model"][i]), _dp(ext), _dp(intr), _dp(pt), _dp(xy), 0, 2.0, _dp(r), C.byref(rho0), _dp(J)) == 1
        tag = str(g["tag"][i])
        assert np.abs(r - g["r"][i]).max() <= 1e-12 * max(1.0, np.abs(g["r"][i]).max()), (i, tag)
        tol = np.full(20, 1e-12)
        if g["model"][i] == _abi.MODEL_DIVISION_UNDISTORTION and intr[4] != 0.0:   # cancellation, see tests/test_oracle_golden.py
            q = rotation_from_angle_axis(ext[3:6])[0] @ (pt[:3] - pt[3] * ext[:3])
            x = abs(4.0 * intr[4] * ((intr[0] * q[0] / q[2]) ** 2 + (intr[0] * intr[1] * q[1] / q[2]) ** 2))
            tol = np.maximum(tol, 2e-14 / x); tol[10] = max(1e-12, 2e-14 / x ** 2)
        assert (np.abs(J - g["J"][i]).max(axis=0) <= tol * np.abs(g["J"][i]).max()).all(), (i, tag)
        rr = np.zeros(2)
        assert H.host_reproject(int(g["model"][i]), _dp(ext), _dp(intr), _dp(pt), _dp(xy), _dp(rr)) == 1
        assert np.abs(rr - g["r"][i]).max() <= 1e-12 * max(1.0, np.abs(g["r"][i]).max())
    # viewing rays: PixelToCameraCoordinates against the oracle's restatement on a pixel grid
    for model, intr in ((2, [1200.0, 1.02, 0.3, 600, 400, 0.01, 0.001, 0.001, 0.001, 0]), (3, [1200.0, 0.98, 600, 400, 0.1, 0, 0, 0, 0, 0]),
                        (3, [1200.0, 1.0, 600, 400, 1e-4, 0, 0, 0, 0, 0]), (4, [1200.0, 1.01, 600, 400, -1e-6, 0, 0, 0, 0, 0])):
        k = np.array(intr, np.float64)
        for x in np.arange(0.0, 1200.0, 97.0):
            for y in np.arange(0.0, 800.0, 89.0):
                out = np.zeros(3); pix = np.array([x, y])
                H.host_pixel_to_camera_ext(model, _dp(k), _dp(pix), _dp(out))
                assert np.abs(out - oracle.pixel_to_camera(model, k, pix)).max() <= 1e-14


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
