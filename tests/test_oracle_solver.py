"""Checks of the oracle's solver algebra against dense numpy / scipy (independent implementations)."""
import numpy as np
import pytest

from helpers import dense_jacobian, free_masks, rel_err
from theiasfm_b200 import _abi, synthetic


def _tiny(model=_abi.MODEL_PINHOLE, shared=True, seed=3, **kw):
    return synthetic.make_scene(n_cam=8, n_pt=60, obs_per_pt=4, model=model, shared_intrinsics=shared, seed=seed, **kw)


def _dense_system(oracle, p, radius, opt):
    """Dense restatement: scaled Jacobian, LM diagonal, Schur complement, rhs."""
    r, Jobs, ok = oracle.residual_jacobian(p)
    assert ok.all()
    J = dense_jacobian(p, Jobs)
    nc, ng = p.n_cam, p.n_group
    ncs = 6 * nc + 10 * ng
    fc, fi, fp = free_masks(p)
    free = np.concatenate([fc.ravel(), fi.ravel(), fp.ravel()])
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    Js = J * scale
    diag = np.clip((Js * Js).sum(0), opt.min_lm_diagonal, opt.max_lm_diagonal)
    D2 = np.where(free, diag / radius, 0.0)
    H = Js.T @ Js + np.diag(D2)
    b = Js.T @ r.ravel()
    B, E, Cm = H[:ncs, :ncs], H[:ncs, ncs:], H[ncs:, ncs:]
    Cm = Cm + np.diag(np.where(free[ncs:], 0.0, 1.0))  # constant points: identity, E is zero there
    S = B - E @ np.linalg.solve(Cm, E.T)
    rhs = b[:ncs] - E @ np.linalg.solve(Cm, b[ncs:])
    return dict(J=J, Js=Js, scale=scale, D2=D2, H=H, b=b, S=S, rhs=rhs, free=free, ncs=ncs, r=r.ravel())


@pytest.mark.parametrize("model,shared,mask", [(_abi.MODEL_PINHOLE, True, _abi.INTR_FOCAL_LENGTH | _abi.INTR_RADIAL_DISTORTION),
                                               (_abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, False, _abi.INTR_ALL),
                                               (_abi.MODEL_PINHOLE, False, _abi.INTR_NONE)])
def test_linear_system_against_dense(oracle, model, shared, mask):
    p = _tiny(model, shared, intrinsics_to_optimize=mask)
    p.ext_const[1] = _abi.EXT_ALL_CONST
    p.ext_const[2] = _abi.EXT_POSITION_CONST
    p.ext_const[3] = _abi.EXT_ORIENTATION_CONST
    p.pt_const[5] = 1
    opt = oracle.default_options(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR)
    o = oracle.Oracle(p, opt)
    ok, cost = o.linearize()
    assert ok
    radius = 1e4
    d = _dense_system(oracle, p, radius, opt)
    assert abs(cost - 0.5 * d["r"] @ d["r"]) <= 1e-12 * cost
    # gradient of the unscaled Jacobian
    g = d["J"].T @ d["r"]
    nc, ng = p.n_cam, p.n_group
    assert rel_err(o.read(_abi.VEC_GRADIENT_CAM), g[:6 * nc]) < 1e-12
    assert rel_err(o.read(_abi.VEC_GRADIENT_INTR), g[6 * nc:d["ncs"]]) < 1e-12
    assert rel_err(o.read(_abi.VEC_GRADIENT_PT), g[d["ncs"]:]) < 1e-12
    assert o.prepare_linear_system(radius)
    rhs = np.concatenate([o.read(_abi.VEC_SCHUR_RHS_CAM), o.read(_abi.VEC_SCHUR_RHS_INTR)])
    assert rel_err(rhs, d["rhs"]) < 1e-10
    # S x for random x (constant coordinates of x zero, as in every CG vector)
    rng = np.random.default_rng(0)
    x = rng.normal(size=d["ncs"]) * d["free"][:d["ncs"]]
    yc, yi = o.schur_matvec(x[:6 * nc], x[6 * nc:])
    assert rel_err(np.concatenate([yc, yi]), d["S"] @ x) < 1e-10
    # SCHUR_JACOBI blocks = inverse of the diagonal blocks of S (identity on constant coordinates)
    Mc = o.read(_abi.VEC_PRECOND_CAM).reshape(nc, 6, 6)
    for c in range(nc):
        blk = d["S"][6 * c:6 * c + 6, 6 * c:6 * c + 6].copy()
        fr = d["free"][6 * c:6 * c + 6]
        blk[~fr, :] = 0; blk[:, ~fr] = 0; blk[~fr, ~fr] = 1.0
        assert rel_err(Mc[c], np.linalg.inv(blk)) < 1e-8
    Mi = o.read(_abi.VEC_PRECOND_INTR).reshape(ng, 10, 10)
    for gi in range(ng):
        s0 = 6 * nc + 10 * gi
        blk = d["S"][s0:s0 + 10, s0:s0 + 10].copy()
        fr = d["free"][s0:s0 + 10]
        blk[~fr, :] = 0; blk[:, ~fr] = 0; blk[~fr, ~fr] = 1.0
        assert rel_err(Mi[gi], np.linalg.inv(blk)) < 1e-7
    # CG to (near) convergence reproduces the dense LM step
    o.options.eta = 1e-14
    o2 = oracle.Oracle(p, oracle.default_options(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR, eta=1e-13, max_linear_solver_iterations=2000))
    o2.linearize(); o2.prepare_linear_system(radius)
    ok, iters, mcc = o2.solve_linear_system()
    assert ok and iters > 1
    Hf = d["H"] + np.diag(np.where(d["free"], 0.0, 1.0))
    y = np.linalg.solve(Hf, d["b"])
    delta = -(y * d["scale"])
    got = np.concatenate([o2.read(_abi.VEC_STEP_CAM), o2.read(_abi.VEC_STEP_INTR), o2.read(_abi.VEC_STEP_PT)])
    assert rel_err(got, delta) < 1e-4  # CG stops on the Q-test, not on the residual
    m = d["Js"] @ (-y)
    assert abs(mcc - (-(m @ (d["r"] + m / 2)))) <= 1e-6 * abs(mcc)
    ok, cand = o2.evaluate_step()
    assert ok and cand < cost


def test_final_cost_matches_scipy(oracle):
    from scipy.optimize import least_squares
    p = _tiny(seed=11)
    p0 = p.copy()
    opt = oracle.default_options(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR, function_tolerance=1e-14,
                                 parameter_tolerance=1e-14, gradient_tolerance=1e-14, max_num_iterations=200)
    s = oracle.solve(p, opt)
    assert s.success and s.final_cost < s.initial_cost
    fc, fi, fp = free_masks(p0)
    free = np.concatenate([fc.ravel(), fi.ravel(), fp.ravel()])
    x0 = np.concatenate([p0.ext.ravel(), p0.intr.ravel(), p0.pt.ravel()])

    def fun(z):
        x = x0.copy(); x[free] = z
        q = p0.copy()
        q.ext[:] = x[:p0.n_cam * 6].reshape(-1, 6); q.intr[:] = x[p0.n_cam * 6:p0.n_cam * 6 + p0.n_group * 10].reshape(-1, 10)
        q.pt[:] = x[p0.n_cam * 6 + p0.n_group * 10:].reshape(-1, 4)
        r, _, _ = oracle.residual_jacobian(q)
        return r.ravel()

    res = least_squares(fun, x0[free], method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-12, max_nfev=400)
    assert abs(res.cost - s.final_cost) <= 1e-6 * res.cost, (res.cost, s.final_cost)
    # monotone decrease over successful steps
    costs = [it["cost"] for it in s.iterations if it["step_is_successful"]]
    assert all(b <= a for a, b in zip(costs, costs[1:]))


def test_loss_functions_closed_form(oracle):
    a = 2.0
    for s in (0.0, 0.5, 3.9, 4.1, 100.0):
        rho = oracle.loss(_abi.LOSS_HUBER, a, s)
        exp = (s, 1.0) if s <= a * a else (2 * a * np.sqrt(s) - a * a, a / np.sqrt(s))
        assert np.allclose(rho[:2], exp, rtol=1e-14)
        rho = oracle.loss(_abi.LOSS_CAUCHY, a, s)
        assert np.allclose(rho[:2], (a * a * np.log1p(s / (a * a)), 1.0 / (1.0 + s / (a * a))), rtol=1e-13)
        rho = oracle.loss(_abi.LOSS_SOFTLONE, a, s)
        assert np.allclose(rho[:2], (2 * a * a * (np.sqrt(1 + s / (a * a)) - 1), 1 / np.sqrt(1 + s / (a * a))), rtol=1e-13)
        rho = oracle.loss(_abi.LOSS_ARCTAN, a, s)
        assert np.allclose(rho[:2], (a * np.arctan2(s, a), 1 / (1 + s * s / (a * a))), rtol=1e-13)
        rho = oracle.loss(_abi.LOSS_TRIVIAL, a, s)
        assert tuple(rho) == (s, 1.0, 0.0)
        rho = oracle.loss(_abi.LOSS_TUKEY, a, s)
        if s <= a * a:
            assert np.allclose(rho[:2], (a * a / 6 * (1 - (1 - s / (a * a)) ** 3), 0.5 * (1 - s / (a * a)) ** 2), rtol=1e-13)
        else:
            assert tuple(rho) == (a * a / 6, 0.0, 0.0)


def test_robust_loss_and_constant_blocks_solve(oracle):
    p = _tiny(seed=5)
    # a gross outlier observation
    p.obs_xy[7] += 300.0
    p.ext_const[0] = _abi.EXT_ALL_CONST
    opt = oracle.default_options(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR,
                                 loss_function_type=_abi.LOSS_HUBER, robust_loss_width=2.0, max_num_iterations=50)
    before = p.ext[0].copy()
    s = oracle.solve(p, opt)
    assert s.success and s.final_cost < s.initial_cost
    assert np.array_equal(p.ext[0], before)  # constant block untouched bit-for-bit


def test_unsupported_options_fail_loudly(oracle):
    p = _tiny()
    s = oracle.solve(p, oracle.default_options(linear_solver_type=6))  # CGNR: the one solver type that is not restated
    assert s.rc == _abi.ERR_UNSUPPORTED and not s.success
    s = oracle.solve(_tiny(), oracle.default_options(max_num_iterations=5))  # Theia defaults (SPARSE_SCHUR + inner iterations) run
    assert s.rc == 0 and s.success


@pytest.mark.parametrize("loss", [_abi.LOSS_HUBER, _abi.LOSS_SOFTLONE, _abi.LOSS_CAUCHY, _abi.LOSS_ARCTAN, _abi.LOSS_TUKEY])
def test_every_robust_loss_solves_and_resists_outliers(oracle, loss):
    """create_loss_function.cc:42-71: each robust loss runs end to end; with 5 % gross outliers the robust solve ends
    closer to the ground-truth cameras than the TRIVIAL (L2) solve."""
    p, truth = synthetic.make_scene(n_cam=10, n_pt=250, obs_per_pt=5, seed=17, return_truth=True, perturb=0.3)
    rng = np.random.default_rng(1)
    bad = rng.choice(p.n_obs, p.n_obs // 20, replace=False)
    p.obs_xy[bad] += rng.uniform(-60, 60, (len(bad), 2))
    from helpers import umeyama_align

    def solve(kind):
        q = p.copy()
        s = oracle.solve(q, oracle.default_options(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR,
                                                   loss_function_type=kind, robust_loss_width=2.0, max_num_iterations=60))
        assert s.success and s.final_cost < s.initial_cost
        aligned, _ = umeyama_align(q.ext[:, :3], truth["ext"][:, :3])
        return np.linalg.norm(aligned - truth["ext"][:, :3], axis=1).max()

    assert solve(loss) < 0.5 * solve(_abi.LOSS_TRIVIAL)


@pytest.mark.parametrize("solver", [_abi.SPARSE_SCHUR, _abi.DENSE_SCHUR])
def test_exact_schur_solver_types(oracle, solver):
    """DENSE_SCHUR / SPARSE_SCHUR (Theia's default linear solver, bundle_adjustment.h:86): the oracle factorises the
    explicit reduced camera matrix; its LM step equals the dense normal-equations step, and the solve converges to the
    same minimum as the ITERATIVE_SCHUR restatement and scipy."""
    p = _tiny(seed=11)
    opt = oracle.default_options(use_inner_iterations=0, linear_solver_type=solver)
    o = oracle.Oracle(p.copy(), opt)
    ok, cost = o.linearize()
    radius = 1e4
    assert ok and o.prepare_linear_system(radius)
    ok, iters, mcc = o.solve_linear_system()
    assert ok and iters == 1
    d = _dense_system(oracle, p, radius, opt)
    Hf = d["H"] + np.diag(np.where(d["free"], 0.0, 1.0))
    y = np.linalg.solve(Hf, d["b"])
    got = np.concatenate([o.read(_abi.VEC_STEP_CAM), o.read(_abi.VEC_STEP_INTR), o.read(_abi.VEC_STEP_PT)])
    assert rel_err(got, -(y * d["scale"])) < 1e-9
    pe, pi = p.copy(), p.copy()
    kw = dict(use_inner_iterations=0, function_tolerance=1e-12, max_num_iterations=100)
    se = oracle.solve(pe, oracle.default_options(linear_solver_type=solver, **kw))
    si = oracle.solve(pi, oracle.default_options(linear_solver_type=_abi.ITERATIVE_SCHUR, **kw))
    assert se.success and abs(se.final_cost - si.final_cost) <= 1e-7 * si.final_cost
    assert all(i["linear_solver_iterations"] in (0, 1) for i in se.iterations)
