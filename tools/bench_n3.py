"""Throughput of the N3 kernels (batched TrackEstimator::EstimateTrack, BundleAdjustTrack,
BundleAdjustTwoViews, and the linearisation of the camera models other than pinhole) on cuda:0, next to the oracle's
restatement of the same per-track / per-pair work on the host cores (OpenMP over tracks / pairs, the reference's thread
pool: estimate_track.cc:161-180).  GPU figures are API-level: wall clock of the C-ABI call on a device-resident problem
including the device->host copy of the per-track status; best of `--repeat` calls.  The CPU arm runs a bounded sample
(same per-track statistics, fewer tracks).  One JSON object on stdout (and in gpurun_out/n3_bench.json)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from theiasfm_b200 import _abi, engine, synthetic  # noqa: E402
from oracle import oracle_py  # noqa: E402

KW = dict(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR)


def best_of(f, n):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return min(ts)


def tracks_case(name, n_cam, n_pt, cpu_pt, repeat, model=_abi.MODEL_PINHOLE):
    out = {"case": name, "n_tracks": n_pt, "obs_per_track": 10}
    p = synthetic.make_scene(n_cam=n_cam, n_pt=n_pt, obs_per_pt=10, seed=5, model=model, noise_px=0.5, perturb=0.0)
    start = p.pt.copy()
    eng = engine.Engine()
    eng.upload(p, engine.default_options(**KW))
    opts = engine.default_options(**KW)
    for ba in (False, True):
        st, counts = eng.estimate_tracks(opts, 5.0, 3.0, ba)  # warm-up (module load, first-launch overheads)
        t = best_of(lambda: eng.estimate_tracks(opts, 5.0, 3.0, ba), repeat)
        out["estimate_tracks_%s" % ("with_ba" if ba else "triangulate_only")] = {
            "gpu_s": t, "gpu_tracks_per_s": n_pt / t, "estimated": int(counts[0])}
    # BundleAdjustTrack on perturbed points (cameras fixed)
    q = synthetic.make_scene(n_cam=n_cam, n_pt=n_pt, obs_per_pt=10, seed=5, model=model, noise_px=0.5, perturb=1.0)
    aopts = engine.default_options(**dict(KW, loss_function_type=_abi.LOSS_HUBER, robust_loss_width=3.0))
    eng.upload(q, aopts)
    eng.adjust_tracks(aopts)

    def adj():
        eng.reset_parameters(q)
        eng.adjust_tracks(aopts)
    t_reset = best_of(lambda: eng.reset_parameters(q), repeat)
    t = best_of(adj, repeat) - t_reset
    out["adjust_tracks"] = {"gpu_s": t, "gpu_tracks_per_s": n_pt / t}
    eng.close()
    # host arm: the oracle on a sample with the same track statistics
    threads = oracle_py.num_threads()
    ps = synthetic.make_scene(n_cam=min(n_cam, 1000), n_pt=cpu_pt, obs_per_pt=10, seed=6, model=model, noise_px=0.5, perturb=0.0)
    for ba in (False, True):
        r = ps.copy()
        t0 = time.perf_counter(); oracle_py.estimate_tracks(r, oracle_py.default_options(**KW), bundle_adjustment=ba); t = time.perf_counter() - t0
        k = "estimate_tracks_%s" % ("with_ba" if ba else "triangulate_only")
        out[k].update(cpu_s=t, cpu_tracks=cpu_pt, cpu_tracks_per_s=cpu_pt / t, cpu_threads=threads,
                      gpu_over_cpu=out[k]["gpu_tracks_per_s"] / (cpu_pt / t))
    qs = synthetic.make_scene(n_cam=min(n_cam, 1000), n_pt=cpu_pt, obs_per_pt=10, seed=6, model=model, noise_px=0.5, perturb=1.0)
    t0 = time.perf_counter()
    oracle_py.adjust_tracks(qs, oracle_py.default_options(**dict(KW, loss_function_type=_abi.LOSS_HUBER, robust_loss_width=3.0)))
    t = time.perf_counter() - t0
    out["adjust_tracks"].update(cpu_s=t, cpu_tracks=cpu_pt, cpu_tracks_per_s=cpu_pt / t, cpu_threads=threads,
                                gpu_over_cpu=out["adjust_tracks"]["gpu_tracks_per_s"] / (cpu_pt / t))
    del start
    return out


def two_view_case(n_pairs, cpu_pairs, repeat):
    b = synthetic.make_two_view_batch(n_pairs, seed=11)
    eng = engine.Engine()
    eng.two_view_ba_batch(b.copy())
    copies = [b.copy() for _ in range(repeat)]
    ts = []
    for c in copies:
        t = time.perf_counter(); term, ic, fc, it = eng.two_view_ba_batch(c); ts.append(time.perf_counter() - t)
    eng.close()
    t = min(ts)
    out = {"case": "two_view_ba", "n_pairs": n_pairs, "correspondences": int(b.pair_off[-1]), "gpu_s": t, "gpu_pairs_per_s": n_pairs / t,
           "mean_iterations": float(it.mean()), "converged": int((term == _abi.CONVERGENCE).sum())}
    s = synthetic.make_two_view_batch(cpu_pairs, seed=12)
    t0 = time.perf_counter(); oracle_py.two_view_ba_batch(s); tc = time.perf_counter() - t0
    out.update(cpu_s=tc, cpu_pairs=cpu_pairs, cpu_pairs_per_s=cpu_pairs / tc, cpu_threads=oracle_py.num_threads(),
               gpu_over_cpu=(n_pairs / t) / (cpu_pairs / tc))
    return out


def linearize_case(model, name, n_cam, n_pt, repeat):
    """Device time of one linearisation (residuals + compact Jacobian + gradient + column norms) per camera model."""
    p = synthetic.make_scene(n_cam=n_cam, n_pt=n_pt, obs_per_pt=10, seed=8, model=model, shared_intrinsics=False)
    eng = engine.Engine()
    eng.upload(p, engine.default_options(**KW))
    eng.linearize()
    eng.set_profiling(True)
    for _ in range(repeat):
        eng.linearize()
    st = eng.profile_stages()["linearize"]
    prof = eng.profile()
    eng.close()
    ms = st["ms"] / max(st["launches"], 1)
    return {"case": "linearize_" + name, "n_obs": p.n_obs, "ms_per_launch": ms, "obs_per_s": p.n_obs / (ms * 1e-3),
            "doubles_per_obs": prof["doubles_per_obs"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--big", type=int, default=2_000_000, help="tracks of the large case (config 3's point count)")
    ap.add_argument("--cpu-tracks", type=int, default=100_000)
    ap.add_argument("--pairs", type=int, default=20_000)
    ap.add_argument("--cpu-pairs", type=int, default=400)
    ap.add_argument("--only-big", action="store_true", help="under ncu: only the large track case and the two-view batch")
    a = ap.parse_args()
    res = {"threads": oracle_py.num_threads(), "cases": []}
    if not a.only_big:
        res["cases"].append(tracks_case("fountain_sized (8k tracks)", 50, 8_000, 8_000, a.repeat))
    res["cases"].append(tracks_case("config3_sized", 10_000, a.big, a.cpu_tracks, a.repeat))
    res["cases"].append(two_view_case(a.pairs, a.cpu_pairs, a.repeat))
    for m, nm in () if a.only_big else ((_abi.MODEL_PINHOLE, "pinhole"), (_abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, "radtan"), (_abi.MODEL_FISHEYE, "fisheye"),
                  (_abi.MODEL_FOV, "fov"), (_abi.MODEL_DIVISION_UNDISTORTION, "division_undistortion")):
        res["cases"].append(linearize_case(m, nm, 1000, 500_000, 5))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "n3_bench.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
