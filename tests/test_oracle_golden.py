"""The oracle's jets vs the committed torch.func.jacfwd golden vectors (tests/golden/make_golden.py)."""
import numpy as np

from helpers import golden_problem
from theiasfm_b200 import _abi


def test_oracle_residual_and_jacobian_match_golden(oracle):
    prob, g = golden_problem()
    r, J, ok = oracle.residual_jacobian(prob)
    assert ok.all()
    # tolerance: fp64 reassociation only (two independent forward-mode ADs of the same expression)
    for i in range(prob.n_obs):
        scale_r = max(1.0, np.abs(g["r"][i]).max())
        assert np.abs(r[i] - g["r"][i]).max() <= 1e-12 * scale_r, (i, str(g["tag"][i]))
        scale_j = np.abs(g["J"][i]).max()
        assert np.abs(J[i] - g["J"][i]).max() <= 1e-12 * scale_j, (i, str(g["tag"][i]))


def test_oracle_matches_golden_for_fisheye_fov_division(oracle):
    """The three other camera models, every branch of their DistortPoint, against torch jacfwd."""
    prob, g = golden_problem(ext=True)
    r, J, ok = oracle.residual_jacobian(prob)
    assert ok.all() and set(g["model"].tolist()) == {2, 3, 4}
    from theiasfm_b200.synthetic import rotation_from_angle_axis
    for i in range(prob.n_obs):
        scale_r = max(1.0, np.abs(g["r"][i]).max())
        assert np.abs(r[i] - g["r"][i]).max() <= 1e-12 * scale_r, (i, str(g["tag"][i]))
        tol = np.full(20, 1e-12)
        if g["model"][i] == _abi.MODEL_DIVISION_UNDISTORTION and g["intr"][i][4] != 0.0:
            # (1 - sqrt(1 - x)) / (x / 2), x = 4 k r_u^2, cancels: two correct evaluations differ by eps / x, and by eps / x^2
            # in the derivative with respect to k (measured: err * x ~ 1e-15, err_k * x^2 ~ 2e-15)
            e, X, k = g["ext"][i], g["pt"][i], g["intr"][i]
            q = rotation_from_angle_axis(e[3:6])[0] @ (X[:3] - X[3] * e[:3])
            x = abs(4.0 * k[4] * ((k[0] * q[0] / q[2]) ** 2 + (k[0] * k[1] * q[1] / q[2]) ** 2))
            tol = np.maximum(tol, 2e-14 / x)
            tol[6 + 4] = max(1e-12, 2e-14 / x ** 2)
        scale_j = np.abs(g["J"][i]).max()
        assert (np.abs(J[i] - g["J"][i]).max(axis=0) <= tol * scale_j).all(), (i, str(g["tag"][i]))
    tags = set(str(t) for t in g["tag"])
    assert {"fisheye_r_sq_below_1e-8", "fov_small_omega", "fov_small_radius", "division_k_zero", "division_negative_sqrt_argument"} <= tags


def test_golden_covers_both_rotation_branches():
    _, g = golden_problem()
    th2 = (g["ext"][:, 3:6] ** 2).sum(1)
    assert (th2 <= np.finfo(float).eps).sum() >= 4 and (th2 > np.finfo(float).eps).sum() >= 80
    assert set(g["model"].tolist()) == {0, 1}


def test_too_close_to_camera_centre_fails(oracle):
    # reprojection_error.h:75-77 -- ||X - h C||^2 < 1e-8 makes the functor return false
    from theiasfm_b200 import _abi
    ext = np.array([[1.0, 2.0, 3.0, 0.1, 0.2, 0.3]])
    intr = np.zeros((1, 10)); intr[0, :5] = [800, 1, 0, 500, 500]
    for d, expect in ((5e-5, False), (2e-4, True)):
        pt = np.array([[1.0 + d, 2.0, 3.0, 1.0]])
        p = _abi.Problem(ext, [0], [0], [0], intr, [0], pt, [0], [0], [0], [[0.0, 0.0]])
        _, _, ok = oracle.residual_jacobian(p)
        assert bool(ok[0]) is expect
