"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/theia_ba_b200.h declares, its struct layouts match the ctypes mirror, and it fails
loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from theiasfm_b200 import _abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = engine.lib()
    hdr = open(os.path.join(ROOT, "include", "theia_ba_b200.h")).read()
    declared = set(re.findall(r"\b(tba_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(engine.EXPORTED_SYMBOLS), declared ^ set(engine.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_header():
    sizes = (C.c_int32 * 4)()
    engine.lib().tba_abi_sizes(sizes)
    assert list(sizes) == [C.sizeof(_abi.tba_options), C.sizeof(_abi.tba_problem), C.sizeof(_abi.tba_summary),
                           C.sizeof(_abi.tba_iteration)]
    assert engine.lib().tba_abi_size_two_view_batch() == C.sizeof(_abi.tba_two_view_batch)


def test_default_options_mirror_theia_defaults():
    # bundle_adjustment.h:78-122
    o = engine.default_options()
    assert o.loss_function_type == _abi.LOSS_TRIVIAL and o.robust_loss_width == 2.0
    assert o.linear_solver_type == _abi.SPARSE_SCHUR and o.preconditioner_type == _abi.PRECOND_SCHUR_JACOBI
    assert o.intrinsics_to_optimize == (_abi.INTR_FOCAL_LENGTH | _abi.INTR_RADIAL_DISTORTION)
    assert o.max_num_iterations == 100 and o.max_solver_time_in_seconds == 3600.0 and o.use_inner_iterations == 1
    assert (o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance, o.max_trust_region_radius) == (1e-6, 1e-10, 1e-8, 1e12)
    assert not o.constant_camera_orientation and not o.constant_camera_position and o.num_threads == 1


def test_oracle_and_engine_agree_on_defaults(oracle):
    a, b = engine.default_options(), oracle.default_options()
    for f, _ in _abi.tba_options._fields_:
        assert getattr(a, f) == getattr(b, f), f


def test_shard_points_matches_python_mirror():
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 30, 1000).astype(np.int32)
    for world in (1, 2, 3, 8):
        prev_end = 0
        for rank in range(world):
            b, e = engine.shard_points(counts, world, rank)
            assert (b, e) == _abi.shard_points(counts, world, rank)
            assert b == prev_end
            prev_end = e
        assert prev_end == len(counts)


@pytest.mark.skipif(engine.device_count() > 0, reason="only meaningful without a GPU")
def test_no_gpu_fails_loudly():
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine()
    assert ei.value.code == _abi.ERR_NO_DEVICE


def test_batched_entry_points_fail_loudly_without_a_gpu():
    """No CPU fallback anywhere: without a CUDA device the batched N3 entry points return TBA_ERR_NO_DEVICE."""
    from theiasfm_b200 import synthetic
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    b = synthetic.make_two_view_batch(3, min_corr=10, max_corr=12, seed=1)
    with pytest.raises(engine.EngineError):
        engine.two_view_ba_batch_multi(b, n_devices=0)
    before = b.points.copy()
    st = b.as_struct()
    term = (C.c_uint8 * 3)()
    assert engine.lib().tba_two_view_ba_batch_multi(C.byref(st), 0, term, None, None, None) == _abi.ERR_NO_DEVICE
    assert np.array_equal(b.points, before)
