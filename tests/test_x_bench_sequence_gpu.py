"""The engine call sequence bench.py's timed region makes (bench.py:211-229): upload -> minimize (warm-up) -> reset_parameters ->
upload -> reset_parameters -> set_profiling -> minimize -> profile.  The warm-up must leave no trace: the timed minimise starts
from the initial estimate and reproduces a fresh solve (to the rounding of the unordered fp64 REDs); the profile counters describe what ran."""
import numpy as np
import pytest

from theiasfm_b200 import _abi, engine, synthetic

pytestmark = pytest.mark.gpu


def _kw(k):
    return dict(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR, max_num_iterations=k, function_tolerance=0.0,
                gradient_tolerance=0.0, parameter_tolerance=0.0)


def test_warmup_reset_timed_sequence_matches_a_fresh_solve():
    p = synthetic.make_scene(n_cam=16, n_pt=1200, obs_per_pt=6, seed=77)
    fresh = p.copy()
    e0 = engine.Engine()
    s0 = e0.solve(fresh, engine.default_options(**_kw(4)))
    e0.close()
    assert s0.rc == 0

    shard, init = p.copy(), p.copy()
    eng = engine.Engine()
    eng.upload(shard, engine.default_options(**_kw(3)))
    eng.minimize()
    eng.reset_parameters(init)
    eng.upload(shard, engine.default_options(**_kw(4)))
    eng.reset_parameters(init)
    eng.set_profiling(True)
    s = eng.minimize()
    prof = eng.profile()
    eng.set_profiling(False)
    eng.download()
    eng.close()
    # Two solves on the GPU agree only up to the order of the fp64 RED accumulations (not reproducible run to run): measured with the
    # emulator's random scheduling order (TBA_EMU_ORDER=random; the ascending order hid it and an exact comparison passed there) the costs
    # differ in the 12th digit here and by <= 1e-10 on a 100k-observation scene; a warm-up that leaked state would be off by far more.
    assert s.num_iterations == s0.num_iterations and np.all(np.abs(s.costs - s0.costs) <= 1e-9 * s0.costs)
    for a, b in ((shard.pt, fresh.pt), (shard.ext, fresh.ext), (shard.intr, fresh.intr)):
        assert np.abs(a - b).max() <= 1e-7 * np.abs(b).max()
    assert prof["observations"] == p.n_obs and prof["points"] == p.n_pt and prof["slots"] >= p.n_obs
    assert prof["linearize_launches"] == s.num_iterations          # one linearisation per evaluated point incl. the initial one
    assert prof["matvec_launches"] >= sum(it["linear_solver_iterations"] for it in s.iterations)   # + the rhs / residual products
    assert prof["matvec_ms"] >= 0.0 and prof["linearize_ms"] >= 0.0
