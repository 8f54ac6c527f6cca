"""N > 1: one process per GPU, points+observations sharded, cameras replicated, NCCL all-reduce of the camera-space
sums.  The sharded solve must reproduce the single-GPU solve (same iteration counts, costs to 1e-9)."""
import os
import sys

import numpy as np
import pytest

from theiasfm_b200 import _abi, engine, synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR, max_num_iterations=12)


def _worker(rank, world, nccl_id, scene_kw, out):
    sys.path.insert(0, ROOT)
    from theiasfm_b200 import engine as eng_mod
    p = synthetic.make_scene(**scene_kw)
    shard, b, e = p.shard(rank, world)
    eng = eng_mod.Engine(device=rank, rank=rank, world_size=world, nccl_id=nccl_id)
    s = eng.solve(shard, eng_mod.default_options(**KW))
    out.put((rank, s.rc, s.message, list(s.costs), [i["linear_solver_iterations"] for i in s.iterations], b, e,
             shard.pt.copy(), shard.ext.copy(), shard.intr.copy()))
    eng.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("scene", ["shared", "per_camera"])
def test_sharded_solve_matches_single_gpu(world, scene):
    if engine.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    scene_kw = dict(n_cam=40, n_pt=6000, obs_per_pt=8, seed=31) if scene == "shared" else \
        dict(n_cam=30, n_pt=5000, obs_per_pt=6, seed=32, model=_abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, shared_intrinsics=False)
    p = synthetic.make_scene(**scene_kw)
    e1 = engine.Engine()
    ref = p.copy()
    s1 = e1.solve(ref, engine.default_options(**KW))
    e1.close()
    assert s1.rc == 0
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    nccl_id = engine.nccl_unique_id()
    procs = [ctx.Process(target=_worker, args=(r, world, nccl_id, scene_kw, out), daemon=True) for r in range(world)]
    for q in procs:
        q.start()
    try:
        res = sorted([out.get(timeout=150) for _ in range(world)])
        for q in procs:
            q.join(timeout=60)
            assert q.exitcode == 0
    finally:
        for q in procs:  # a hung rank must not hang the suite
            if q.is_alive():
                q.kill()
    for rank, rc, msg, costs, cg, b, e, pt, ext, intr in res:
        assert rc == 0, msg
        assert len(costs) == len(s1.costs) and cg == [i["linear_solver_iterations"] for i in s1.iterations]
        assert np.all(np.abs(np.array(costs) - s1.costs) <= 1e-9 * s1.costs)
        # every rank holds the same cameras; each rank owns its points
        assert np.abs(ext - ref.ext).max() <= 1e-7 * np.abs(ref.ext).max()
        assert np.abs(intr - ref.intr).max() <= 1e-7 * np.abs(ref.intr).max()
        assert np.abs(pt - ref.pt[b:e]).max() <= 1e-7 * np.abs(ref.pt).max()
    # replicated state is bit-identical across ranks (deterministic reductions + NCCL all-reduce)
    for r in res[1:]:
        assert np.array_equal(r[8], res[0][8]) and np.array_equal(r[9], res[0][9]) and r[3] == res[0][3]


_MULTI_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
from theiasfm_b200 import _abi, engine, synthetic
KW = dict(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR, max_num_iterations=12)
world = int(sys.argv[1])
p = synthetic.make_scene(n_cam=40, n_pt=6000, obs_per_pt=8, seed=31)
e1 = engine.Engine()
ref = p.copy()
s1 = e1.solve(ref, engine.default_options(**KW))
e1.close()
got = p.copy()
sm = engine.solve_multi(got, engine.default_options(**KW), n_devices=world)
assert sm.rc == 0 and sm.success, sm.message
assert len(sm.costs) == len(s1.costs) and np.all(np.abs(sm.costs - s1.costs) <= 1e-9 * s1.costs)
assert np.abs(got.pt - ref.pt).max() <= 1e-7 * np.abs(ref.pt).max()
assert np.abs(got.ext - ref.ext).max() <= 1e-7 * np.abs(ref.ext).max()
print("solve_multi ok")
"""


@pytest.mark.parametrize("world", [2, 8])
def test_single_process_multi_gpu_entry_point(world):
    """tba_solve_multi: the form the C++ adapter uses (one host thread per device inside the library).
    Runs in a subprocess with a hard timeout: NOT yet exercised on hardware (the round-1 GPU budget ran out on the
    call that first hit an NCCL-init deadlock in this path, fixed since by code inspection only)."""
    if engine.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import subprocess
    out = subprocess.run([sys.executable, "-c", _MULTI_SCRIPT % ROOT, str(world)], capture_output=True, text=True, timeout=180)
    assert out.returncode == 0 and "solve_multi ok" in out.stdout, out.stdout + out.stderr


def test_two_view_batch_sharded_over_gpus_is_bit_identical():
    """Pairs are independent units: tba_two_view_ba_batch_multi splits the batch over the GPUs with no collective; every pair runs
    the same single-thread LM wherever it lands, so the result equals the one-GPU result bit for bit."""
    if engine.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    b = synthetic.make_two_view_batch(400, min_corr=40, max_corr=200, seed=12)
    b1, b2 = b.copy(), b.copy()
    eng = engine.Engine()
    r1 = eng.two_view_ba_batch(b1)
    eng.close()
    r2 = engine.two_view_ba_batch_multi(b2, n_devices=2)
    for a, c in zip(r1, r2):
        assert np.array_equal(a, c)
    assert np.array_equal(b1.ext2, b2.ext2) and np.array_equal(b1.points, b2.points) and np.array_equal(b1.intr2, b2.intr2)
