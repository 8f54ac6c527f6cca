"""FISHEYE / FOV / DIVISION_UNDISTORTION on the GPU: the EXT instantiations k_linearize<0x3FF, true> /
k_cost<true> (templated projection + forward-mode dual numbers, tba_camera_models_ext.cuh) through the C-ABI, against
the committed torch jacfwd vectors, and stage / trajectory parity with the oracle.  The device bodies are also checked on
the host (tests/test_device_math_on_host.py).  Never executed on hardware in round 1 (GPU budget spent)."""
import numpy as np
import pytest

from helpers import golden_problem, rel_err
from theiasfm_b200 import _abi, engine, synthetic

pytestmark = pytest.mark.gpu
ITER = dict(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR)


def _opts(mod, **kw):
    d = dict(ITER); d.update(kw)
    return mod.default_options(**d)


def test_residuals_and_gradient_match_golden():
    prob, g = golden_problem(ext=True)
    prob.group_const_mask[:] = 0
    eng = engine.Engine()
    eng.upload(prob, _opts(engine, intrinsics_to_optimize=_abi.INTR_ALL))
    ok, cost = eng.linearize()
    assert ok
    res = eng.read(_abi.VEC_RESIDUALS).reshape(-1, 2)
    grad = np.concatenate([eng.read(_abi.VEC_GRADIENT_CAM).reshape(-1, 6), eng.read(_abi.VEC_GRADIENT_INTR).reshape(-1, 10),
                           eng.read(_abi.VEC_GRADIENT_PT).reshape(-1, 4)], axis=1)
    eng.close()
    scale = np.maximum(1.0, np.abs(g["r"]).max(axis=1))
    assert (np.abs(res - g["r"]).max(axis=1) / scale).max() < 1e-12
    assert abs(cost - 0.5 * (g["r"] ** 2).sum()) < 1e-11 * cost
    # gradient = J^T r per case (one camera / group / point per case)
    want = np.einsum("nij,ni->nj", g["J"], g["r"])
    for i in range(prob.n_obs):
        tol = 1e-11
        if g["model"][i] == _abi.MODEL_DIVISION_UNDISTORTION and g["intr"][i][4] != 0.0:
            tol = 1e-6   # (1 - sqrt(1 - x)) / x cancellation of that model's DistortPoint (tests/test_oracle_golden.py)
        assert np.abs(grad[i] - want[i]).max() <= tol * np.abs(want[i]).max(), (i, str(g["tag"][i]))


SCENES = {
    "fisheye_shared": dict(n_cam=12, n_pt=300, obs_per_pt=5, model=_abi.MODEL_FISHEYE, shared_intrinsics=True, seed=41),
    "fov_per_camera": dict(n_cam=10, n_pt=400, obs_per_pt=6, model=_abi.MODEL_FOV, shared_intrinsics=False, seed=42),
    "division_all": dict(n_cam=10, n_pt=400, obs_per_pt=6, model=_abi.MODEL_DIVISION_UNDISTORTION, shared_intrinsics=True, seed=43,
                         intrinsics_to_optimize=_abi.INTR_ALL),
}


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("loss", [_abi.LOSS_TRIVIAL, _abi.LOSS_HUBER])
def test_stage_and_trajectory_parity(oracle, name, loss):
    p = synthetic.make_scene(**SCENES[name])
    p.ext_const[1] = _abi.EXT_ALL_CONST; p.pt_const[[5, 17]] = 1
    if loss != _abi.LOSS_TRIVIAL:
        p.obs_xy[::37] += 40.0
    kw = dict(loss_function_type=loss, robust_loss_width=2.0, max_num_iterations=15)
    o = oracle.Oracle(p.copy(), _opts(oracle, **kw))
    eng = engine.Engine()
    eng.upload(p.copy(), _opts(engine, **kw))
    ok_o, cost_o = o.linearize()
    ok_g, cost_g = eng.linearize()
    stol = 1e-7 if "division" in name else 1e-10
    assert ok_o and ok_g and abs(cost_g - cost_o) <= 1e-11 * cost_o
    for which in (_abi.VEC_GRADIENT_CAM, _abi.VEC_GRADIENT_INTR, _abi.VEC_GRADIENT_PT, _abi.VEC_COLNORM2_CAM, _abi.VEC_COLNORM2_INTR):
        a, b = eng.read(which), o.read(which)
        assert np.abs(a - b).max() <= stol * max(np.abs(b).max(), 1e-300), which
    po, pg = p.copy(), p.copy()
    so = oracle.solve(po, _opts(oracle, **kw))
    sg = eng.solve(pg, _opts(engine, **kw))
    eng.close()
    assert sg.rc == 0 and sg.success and so.success
    assert abs(sg.num_iterations - so.num_iterations) <= 1
    n = min(len(sg.costs), len(so.costs))
    # Sensitivity measured on the CPU (the oracle on the same scene with its observations permuted = rounding-level noise): <= 1e-10
    # on costs and parameters for FISHEYE / FOV, but 7e-7 on costs and 5e-6 on parameters for DIVISION_UNDISTORTION, whose
    # (1 - sqrt(1 - x)) / x cancels -- hence the looser bound for that model
    ctol, ptol = (1e-4, 1e-3) if "division" in name else (1e-6, 1e-5)
    assert np.all(np.abs(sg.costs[:n] - so.costs[:n]) <= ctol * so.costs[:n])
    assert abs(sg.final_cost - so.final_cost) <= ctol * so.final_cost
    assert rel_err(pg.ext, po.ext) < ptol and rel_err(pg.intr, po.intr) < ptol
    assert sg.final_cost < (0.05 if loss == _abi.LOSS_TRIVIAL else 0.5) * sg.initial_cost


def test_mixed_models_in_one_problem(oracle):
    """Theia allows a different camera model per intrinsics group (bundle_adjuster.cc:242-287): PINHOLE and FISHEYE groups in
    one problem run the EXT instantiation for every observation and must agree with the oracle."""
    a = synthetic.make_scene(n_cam=8, n_pt=300, obs_per_pt=5, model=_abi.MODEL_PINHOLE, shared_intrinsics=True, seed=44)
    b = synthetic.make_scene(n_cam=8, n_pt=300, obs_per_pt=5, model=_abi.MODEL_FISHEYE, shared_intrinsics=True, seed=44)
    # same geometry (same seed): cameras 0..3 keep the pinhole group and their pinhole measurements, 4..7 become fisheye
    fish = a.obs_cam >= 4
    assert np.array_equal(a.obs_cam, b.obs_cam) and np.array_equal(a.obs_pt, b.obs_pt)
    xy = np.where(fish[:, None], b.obs_xy, a.obs_xy)
    cam_group = (np.arange(8) >= 4).astype(np.int32)
    p = _abi.Problem(a.ext, a.ext_const, cam_group, [_abi.MODEL_PINHOLE, _abi.MODEL_FISHEYE], np.concatenate([a.intr, b.intr]),
                     [a.group_const_mask[0], b.group_const_mask[0]], a.pt, a.pt_const, a.obs_cam, a.obs_pt, xy)
    kw = dict(max_num_iterations=15)
    po, pg = p.copy(), p.copy()
    so = oracle.solve(po, _opts(oracle, **kw))
    eng = engine.Engine()
    sg = eng.solve(pg, _opts(engine, **kw))
    eng.close()
    assert sg.rc == 0 and sg.success and so.success and abs(sg.num_iterations - so.num_iterations) <= 1
    assert abs(sg.final_cost - so.final_cost) <= 1e-6 * so.final_cost and sg.final_cost < 0.05 * sg.initial_cost
    assert rel_err(pg.intr, po.intr) < 1e-5
