"""Host-side logic of the multi-GPU path on CPU: the point sharding (tba_shard_points / Problem.shard) partitions
the problem exactly, and the quantities the ranks all-reduce (cost, camera-space gradient) sum to the unsharded
ones -- checked with a real world_size-2 gloo process group and the CPU oracle as the evaluator."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from theiasfm_b200 import _abi, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_points_and_observations():
    p = synthetic.make_scene(n_cam=30, n_pt=997, obs_per_pt=7, seed=5)
    for world in (2, 3, 8):
        seen_pts, n_obs, prev_end = 0, 0, 0
        for rank in range(world):
            s, b, e = p.shard(rank, world)
            assert b == prev_end
            prev_end = e
            assert s.n_pt == e - b and np.array_equal(s.pt, p.pt[b:e])
            assert s.n_cam == p.n_cam and np.array_equal(s.ext, p.ext)
            sel = (p.obs_pt >= b) & (p.obs_pt < e)
            assert np.array_equal(s.obs_pt + b, p.obs_pt[sel]) and np.array_equal(s.obs_cam, p.obs_cam[sel])
            seen_pts += s.n_pt
            n_obs += s.n_obs
        assert seen_pts == p.n_pt and n_obs == p.n_obs and prev_end == p.n_pt
        counts = [p.shard(r, world)[0].n_obs for r in range(world)]
        assert max(counts) - min(counts) <= 2 * 7  # balanced by observation count


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py
    oracle_py.set_num_threads(1)  # forked child: libgomp's thread pool of the parent does not exist here
    p = synthetic.make_scene(n_cam=12, n_pt=300, obs_per_pt=5, seed=21)
    shard, b, e = p.shard(rank, world)
    opt = oracle_py.default_options(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR)
    o = oracle_py.Oracle(shard, opt)
    ok, cost = o.linearize()
    t = torch.tensor([cost] + list(o.read(_abi.VEC_GRADIENT_CAM)) + list(o.read(_abi.VEC_GRADIENT_INTR)), dtype=torch.float64)
    dist.all_reduce(t)  # what the engine does with NCCL after k_linearize
    if rank == 0:
        full = oracle_py.Oracle(p, opt)
        ok2, cost2 = full.linearize()
        ref = np.array([cost2] + list(full.read(_abi.VEC_GRADIENT_CAM)) + list(full.read(_abi.VEC_GRADIENT_INTR)))
        out.put(float(np.max(np.abs(t.numpy() - ref) / np.maximum(1.0, np.abs(ref)))))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduced_shard_sums_equal_unsharded_world2(oracle):
    ctx = mp.get_context("fork")  # children inherit the imported modules (a fresh torch import per child is slow)
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for q in procs:
        q.start()
    err = out.get(timeout=180)
    for q in procs:
        q.join(timeout=60)
        assert q.exitcode == 0
    assert err < 1e-11
