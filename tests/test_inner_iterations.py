"""N4: Ceres inner iterations (CoordinateDescentMinimizer in Theia's reversed ordering: extrinsics, intrinsics groups, points;
bundle_adjuster.cc:196-200).  The product's machinery -- the lockstep per-block LM driver the engine runs (tba_block_lm.h) with
its observation passes evaluated by the device bodies of tba_inner.cuh compiled for the host, then tba_point_lm.cuh -- against
the oracle, which builds every block's mini-program explicitly and solves it with oracle_solve."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from theiasfm_b200 import _abi, engine, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def H():
    so, src = os.path.join(HERE, "_host_inner.so"), os.path.join(HERE, "host_inner.cc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def run_host(H, p, loss=0, width=2.0):
    k = engine.debug_pack(p)
    assert k["rc"] == 0
    npk = k["n_packed_points"]
    valid = k["slot_cam"] >= 0
    slots = np.nonzero(valid)[0]; pts = k["slot_pt"][valid]
    first = np.full(npk, -1, np.int64); cnt = np.bincount(pts, minlength=npk).astype(np.int32)
    first[pts[::-1]] = slots[::-1]
    ext, intr = p.ext.copy(), p.intr.copy()
    pt_packed = np.ascontiguousarray(p.pt[k["pk2caller"]])
    ptc = np.ascontiguousarray(p.pt_const[k["pk2caller"]])
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    slot_cam = np.ascontiguousarray(k["slot_cam"]); slot_pt = np.ascontiguousarray(k["slot_pt"]); xy = np.ascontiguousarray(k["xy"])
    mask = np.ascontiguousarray(k["mask"])
    H.host_inner_iterations(p.n_cam, p.n_group, npk, C.c_longlong(len(slot_cam)), ext.ctypes.data_as(dp), intr.ctypes.data_as(dp),
                            pt_packed.ctypes.data_as(dp), p.ext_const.ctypes.data_as(C.POINTER(C.c_uint8)), p.cam_group.ctypes.data_as(ip),
                            p.group_model.ctypes.data_as(ip), p.group_const_mask.ctypes.data_as(C.POINTER(C.c_uint32)), mask.ctypes.data_as(dp),
                            xy.ctypes.data_as(dp), slot_cam.ctypes.data_as(ip), slot_pt.ctypes.data_as(ip),
                            first.ctypes.data_as(C.POINTER(C.c_longlong)), cnt.ctypes.data_as(ip), ptc.ctypes.data_as(C.POINTER(C.c_uint8)),
                            loss, C.c_double(width))
    pt = p.pt.copy(); pt[k["pk2caller"]] = pt_packed
    return ext, intr, pt


def euclid(x):
    return x[:, :3] / x[:, 3:4]


SCENES = {
    "pinhole_shared": dict(n_cam=10, n_pt=250, obs_per_pt=5, seed=51),
    "radtan_per_camera": dict(n_cam=8, n_pt=300, obs_per_pt=6, seed=52, model=_abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, shared_intrinsics=False,
                              intrinsics_to_optimize=_abi.INTR_ALL),
    "fisheye_shared": dict(n_cam=8, n_pt=200, obs_per_pt=5, seed=53, model=_abi.MODEL_FISHEYE),
}


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("loss", [_abi.LOSS_TRIVIAL, _abi.LOSS_HUBER])
def test_coordinate_descent_matches_the_oracle(H, oracle, name, loss):
    p = synthetic.make_scene(**SCENES[name])
    p.ext_const[1] = _abi.EXT_ALL_CONST; p.ext_const[2] = _abi.EXT_POSITION_CONST; p.ext_const[3] = _abi.EXT_ORIENTATION_CONST
    p.pt_const[[4, 9]] = 1
    if loss:
        p.obs_xy[::31] += 25.0
    ext, intr, pt = run_host(H, p, loss=loss, width=2.0)
    q = p.copy()
    L = oracle.lib()
    c0, c1 = C.c_double(), C.c_double()
    opts = oracle.default_options(use_inner_iterations=1, loss_function_type=loss, robust_loss_width=2.0)
    st = q.as_struct()
    L.oracle_debug_inner_iterations.argtypes = [C.POINTER(_abi.tba_options), C.POINTER(_abi.tba_problem), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    assert L.oracle_debug_inner_iterations(C.byref(opts), C.byref(st), C.byref(c0), C.byref(c1)) == 1
    assert c1.value < 0.7 * c0.value                         # one sweep of coordinate descent already helps a lot
    # constant blocks / coordinates untouched, bit for bit
    assert np.array_equal(ext[1], p.ext[1]) and np.array_equal(ext[2, :3], p.ext[2, :3]) and np.array_equal(ext[3, 3:], p.ext[3, 3:])
    assert np.array_equal(pt[[4, 9]], p.pt[[4, 9]])
    # every stage consumes the previous one's output, so agreement at the end means agreement throughout
    assert np.abs(ext - q.ext).max() <= 1e-6 * np.abs(q.ext).max()
    assert np.abs(intr - q.intr).max() <= 1e-6 * np.abs(q.intr).max()
    # (the point stage starts from cameras that agree to 1e-6 only: weakly triangulated points amplify that)
    d = np.abs(euclid(pt) - euclid(q.pt)).max(axis=1) / np.abs(euclid(q.pt)).max()
    assert np.median(d) <= 1e-6 and d.max() <= 1e-3
    # the refined state evaluates to the oracle's cost
    r = p.copy(); r.ext[:], r.intr[:], r.pt[:] = ext, intr, pt
    o = oracle.Oracle(r, oracle.default_options(use_inner_iterations=0, linear_solver_type=_abi.ITERATIVE_SCHUR, loss_function_type=loss,
                                                robust_loss_width=2.0))
    ok, cost = o.linearize()
    o.close()
    assert ok and abs(cost - c1.value) <= 1e-6 * c1.value


def test_oracle_solve_with_inner_iterations_converges_faster_per_iteration(oracle):
    """Theia's default option set (SPARSE_SCHUR + inner iterations) now runs in the oracle; with inner iterations every LM
    iteration ends at a lower cost than without (that is what they are for), and both reach the same minimum."""
    p = synthetic.make_scene(n_cam=10, n_pt=250, obs_per_pt=5, seed=51)
    a, b = p.copy(), p.copy()
    sa = oracle.solve(a, oracle.default_options(max_num_iterations=25))                                 # Theia defaults
    sb = oracle.solve(b, oracle.default_options(max_num_iterations=25, use_inner_iterations=0))
    assert sa.rc == 0 and sa.success and sb.success
    assert sa.costs[1] < sb.costs[1]
    assert abs(sa.final_cost - sb.final_cost) <= 2e-3 * sb.final_cost
