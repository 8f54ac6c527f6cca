"""N4 on the GPU: use_inner_iterations = true (Theia's default, bundle_adjustment.h:114) through the C-ABI -- the lockstep
per-block LM driver with k_block_pass for cameras and intrinsics groups, k_adjust_tracks on the candidate for the points --
against the oracle, whose inner iterations build and solve every block's mini-program explicitly.  The driver and the
per-observation bodies are checked on the host by tests/test_inner_iterations.py.  Never executed on hardware in round 1."""
import numpy as np
import pytest

from helpers import rel_err
from theiasfm_b200 import _abi, engine, synthetic

pytestmark = pytest.mark.gpu

SCENES = {
    "pinhole_shared": dict(n_cam=12, n_pt=300, obs_per_pt=5, seed=61),
    "radtan_per_camera": dict(n_cam=10, n_pt=400, obs_per_pt=6, seed=62, model=_abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, shared_intrinsics=False,
                              intrinsics_to_optimize=_abi.INTR_ALL),
}


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("solver", [_abi.SPARSE_SCHUR, _abi.ITERATIVE_SCHUR])
def test_theia_default_options_match_the_oracle(oracle, name, solver):
    p = synthetic.make_scene(**SCENES[name])
    p.ext_const[1] = _abi.EXT_ALL_CONST; p.ext_const[2] = _abi.EXT_POSITION_CONST; p.pt_const[[4, 9]] = 1
    kw = dict(use_inner_iterations=1, linear_solver_type=solver, max_num_iterations=12)
    po, pg = p.copy(), p.copy()
    so = oracle.solve(po, oracle.default_options(**kw))
    eng = engine.Engine()
    sg = eng.solve(pg, engine.default_options(**kw))
    eng.close()
    assert sg.rc == 0 and sg.success and so.success
    assert abs(sg.initial_cost - so.initial_cost) <= 1e-11 * so.initial_cost
    # Once the fast initial descent is over (3-4 iterations) these scenes crawl along a flat valley (all ten RADTAN intrinsics free):
    # there even two runs of the ORACLE differ by 1e-3 in the intermediate costs (OpenMP reduction order amplified by the inexact CG
    # and the block re-solves -- seen with `pytest -m gpu --mock-engine`), so only the early iterations are compared tightly
    assert abs(sg.num_iterations - so.num_iterations) <= 3
    n = min(len(sg.costs), len(so.costs), 4)
    # (iteration 3 of the all-free RADTAN scene with the inexact PCG already sits at 1.07e-5 under the emulator's rounding: 3e-5 there)
    tol = np.array([1e-5, 1e-5, 1e-5, 3e-5])[:n]
    assert np.all(np.abs(sg.costs[:n] - so.costs[:n]) <= tol * so.costs[:n])
    # Sensitivity measured by running the ENGINE itself with a different rounding (tests/emu FMA-contracted build vs the plain build, both
    # against the oracle): every case stays within 5e-5 (final cost) / 3e-5 (parameters) after 12 iterations, EXCEPT all-free RADTAN
    # intrinsics with the inexact PCG (eta = 0.1), where a 1e-6 difference at iteration 3 grows to 7e-4 at iteration 4 and to 5.6 % in the
    # final cost (1.6e-2 in the parameters) at iteration 12 -- with the exact step (SPARSE_SCHUR) the same scene stays within 3e-6.
    chaotic = name == "radtan_per_camera" and solver == _abi.ITERATIVE_SCHUR
    assert abs(sg.final_cost - so.final_cost) <= (0.15 if chaotic else 2e-2) * so.final_cost
    assert sg.final_cost < 0.05 * sg.initial_cost and sg.costs[1] < 0.2 * sg.costs[0]
    assert rel_err(pg.ext, po.ext) < (0.1 if chaotic else 1e-2)
    if chaotic:  # the part of the trajectory that IS reproducible, compared tightly: 4 iterations (measured: 7e-4 cost, 8e-5 parameters)
        kw["max_num_iterations"] = 4
        po4, pg4 = p.copy(), p.copy()
        so4 = oracle.solve(po4, oracle.default_options(**kw))
        eng = engine.Engine()
        sg4 = eng.solve(pg4, engine.default_options(**kw))
        eng.close()
        assert sg4.rc == 0 and sg4.num_iterations == so4.num_iterations
        assert abs(sg4.final_cost - so4.final_cost) <= 5e-3 * so4.final_cost
        assert rel_err(pg4.ext, po4.ext) < 1e-3 and rel_err(pg4.pt, po4.pt) < 1e-3 and rel_err(pg4.intr, po4.intr) < 1e-3
    assert np.array_equal(pg.ext[1], p.ext[1]) and np.array_equal(pg.ext[2, :3], p.ext[2, :3]) and np.array_equal(pg.pt[[4, 9]], p.pt[[4, 9]])


def test_inner_iterations_lower_the_cost_of_every_iteration():
    p = synthetic.make_scene(n_cam=30, n_pt=3000, obs_per_pt=7, seed=63)
    a, b = p.copy(), p.copy()
    eng = engine.Engine()
    sa = eng.solve(a, engine.default_options(max_num_iterations=10))                                  # Theia defaults
    sb = eng.solve(b, engine.default_options(max_num_iterations=10, use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR))
    eng.close()
    assert sa.rc == 0 and sb.rc == 0 and sa.success and sb.success
    assert sa.costs[1] < sb.costs[1]
    assert abs(sa.final_cost - sb.final_cost) <= 5e-3 * sb.final_cost
