"""The kernel-selection switches against the oracle: TBA_MATVEC=tile (the tile-per-CTA Schur kernels instead of the persistent
streaming ones; they remain the shipped path for tracks of more than 32 observations), TBA_TRED=0 (lane-per-row REDs), TBA_LIN_OCC=2,
TBA_PCG=split (three vector kernels per CG iteration instead of the fused one with grid barriers).
They only change the ORDER of fp64 sums, so the default tolerances of tests/test_gpu_parity.py apply unchanged."""
import numpy as np
import pytest

from helpers import rel_err
from test_gpu_parity import _opts, _scene
from theiasfm_b200 import _abi, engine

pytestmark = pytest.mark.gpu

VARIANTS = {
    "tile_kernels": {"TBA_MATVEC": "tile"},                                        # tile-per-CTA k_schur everywhere (round-1 structure, TRED)
    "tile_no_tred": {"TBA_MATVEC": "tile", "TBA_TRED": "0"},                       # lane-per-row REDs
    "r1_kernels": {"TBA_MATVEC": "tile", "TBA_TRED": "0", "TBA_LIN_OCC": "2"},     # the round-1 defaults
    "lin_occ2": {"TBA_LIN_OCC": "2"},
    "split_pcg": {"TBA_PCG": "split"},                                             # k_pcg_c / k_pcg_a / k_pcg_b instead of k_pcg_fused
}
ALL = ("TBA_TRED", "TBA_MATVEC", "TBA_LIN_OCC", "TBA_ABLATE", "TBA_PCG")


@pytest.fixture
def variant_engine(request, monkeypatch):
    for k in ALL:
        monkeypatch.delenv(k, raising=False)
    for k, v in VARIANTS[request.param].items():
        monkeypatch.setenv(k, v)
    e = engine.Engine()  # the switches are read when the context is created
    yield e
    e.close()


@pytest.mark.parametrize("variant_engine", list(VARIANTS), indirect=True)
@pytest.mark.parametrize("name,loss", [("pinhole_shared", _abi.LOSS_TRIVIAL), ("radtan_per_camera", _abi.LOSS_HUBER)])
def test_switch_full_solve_parity(variant_engine, oracle, name, loss):
    p0 = _scene(name, constants=(loss == _abi.LOSS_HUBER))
    kw = dict(loss_function_type=loss, robust_loss_width=3.0, max_num_iterations=25)
    po, pg = p0.copy(), p0.copy()
    so = oracle.solve(po, _opts(oracle, **kw))
    sg = variant_engine.solve(pg, _opts(engine, **kw))
    assert sg.rc == 0 and sg.success and so.success, sg.message
    assert abs(sg.initial_cost - so.initial_cost) <= 1e-12 * so.initial_cost
    assert sg.num_iterations == so.num_iterations and sg.termination_type == so.termination_type, (sg.message, so.message)
    co, cg = so.costs, sg.costs
    n = min(len(co), 10)
    assert np.all(np.abs(cg[:n] - co[:n]) <= 1e-9 * co[:n])
    assert np.all(np.abs(cg - co) <= 1e-6 * co)
    assert rel_err(pg.ext, po.ext) < 1e-6 and rel_err(pg.pt, po.pt) < 1e-6 and rel_err(pg.intr, po.intr) < 1e-6
    assert np.array_equal(pg.pt[p0.pt_const != 0], p0.pt[p0.pt_const != 0])


@pytest.mark.parametrize("variant_engine", list(VARIANTS), indirect=True)
def test_switch_stage_parity(variant_engine, oracle):
    """Every stage at kernel level (gradient, column norms, reduced rhs, S*x, SCHUR_JACOBI blocks, PCG iteration count):
    the body of test_gpu_parity.test_stage_parity with the variant engine, per-camera RADTAN intrinsics + constants + HUBER."""
    from test_gpu_parity import test_stage_parity as stage_body
    stage_body(variant_engine, oracle, "radtan_per_camera", True, _abi.LOSS_HUBER)
    stage_body(variant_engine, oracle, "pinhole_shared", False, _abi.LOSS_TRIVIAL)


@pytest.mark.parametrize("variant_engine", ["tile_kernels", "tile_no_tred"], indirect=True)
def test_switch_long_tracks(variant_engine, oracle):
    """Long tiles (tracks > 32 observations) take the plain RED path in k_linearize and the staged one elsewhere."""
    from test_gpu_parity import test_long_tracks_use_the_cta_level_path as body
    body(variant_engine, oracle)
