"""ctypes binding of oracle/libba_oracle.so (TEST INFRASTRUCTURE -- see ba_oracle.c header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under theiasfm_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from theiasfm_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libba_oracle.so")
    src = os.path.join(_HERE, "ba_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "theia_ba_b200.h")
    stale = (not os.path.exists(so)) or (os.path.exists(src) and max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "libba_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libba_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp = C.POINTER(C.c_double)
        L.oracle_options_init.argtypes = [C.POINTER(_abi.tba_options)]
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_abi.tba_options), C.POINTER(_abi.tba_problem)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_linearize.argtypes = [C.c_void_p]
        L.oracle_prepare_linear_system.argtypes = [C.c_void_p, C.c_double]
        L.oracle_solve_linear_system.argtypes = [C.c_void_p]
        L.oracle_evaluate_step.argtypes = [C.c_void_p, dp]
        L.oracle_schur_matvec_split.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.oracle_cost.restype = C.c_double
        L.oracle_cost.argtypes = [C.c_void_p]
        L.oracle_model_cost_change.restype = C.c_double
        L.oracle_model_cost_change.argtypes = [C.c_void_p]
        L.oracle_cg_iterations.argtypes = [C.c_void_p]
        L.oracle_read.argtypes = [C.c_void_p, C.c_int, dp, C.c_int64]
        L.oracle_solve.argtypes = [C.POINTER(_abi.tba_options), C.POINTER(_abi.tba_problem), C.POINTER(_abi.tba_summary)]
        L.oracle_residual_jacobian.argtypes = [C.POINTER(_abi.tba_problem), dp, dp, C.POINTER(C.c_uint8)]
        L.oracle_camera_to_pixel.argtypes = [C.c_int, dp, dp, dp]
        L.oracle_pixel_to_camera.argtypes = [C.c_int, dp, dp, dp]
        L.oracle_project_point.restype = C.c_double
        L.oracle_project_point.argtypes = [C.c_int, dp, dp, dp, dp]
        L.oracle_constant_intrinsics_mask.restype = C.c_uint32
        L.oracle_constant_intrinsics_mask.argtypes = [C.c_int, C.c_int]
        L.oracle_loss.argtypes = [C.c_int, C.c_double, C.c_double, dp]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_filter_tracks.argtypes = [C.POINTER(_abi.tba_problem), C.c_double, C.c_double, C.POINTER(C.c_uint8), dp]
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_options(**kw):
    o = _abi.tba_options()
    lib().oracle_options_init(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(n)


def residual_jacobian(problem):
    """Per-observation residual [n,2], full Jacobian [n,2,20] (ext|intr|pt columns), ok [n]."""
    n = problem.n_obs
    r = np.zeros((n, 2)); J = np.zeros((n, 2, 20)); ok = np.zeros(n, np.uint8)
    st = problem.as_struct()
    lib().oracle_residual_jacobian(C.byref(st), _dp(r), _dp(J), ok.ctypes.data_as(C.POINTER(C.c_uint8)))
    return r, J, ok.astype(bool)


def camera_to_pixel(model, intr, q):
    intr = np.ascontiguousarray(intr, np.float64); q = np.ascontiguousarray(q, np.float64); out = np.zeros(2)
    lib().oracle_camera_to_pixel(model, _dp(intr), _dp(q), _dp(out))
    return out


def pixel_to_camera(model, intr, pix):
    intr = np.ascontiguousarray(intr, np.float64); pix = np.ascontiguousarray(pix, np.float64); out = np.zeros(3)
    lib().oracle_pixel_to_camera(model, _dp(intr), _dp(pix), _dp(out))
    return out


def project_point(model, ext, intr, pt):
    ext = np.ascontiguousarray(ext, np.float64); intr = np.ascontiguousarray(intr, np.float64); pt = np.ascontiguousarray(pt, np.float64)
    out = np.zeros(2)
    depth = lib().oracle_project_point(model, _dp(ext), _dp(intr), _dp(pt), _dp(out))
    return out, depth


def filter_tracks(problem, max_inlier_reprojection_error, min_triangulation_angle_degrees):
    """oracle_filter_tracks: (status [n_pt] uint8, mean squared reprojection error [n_pt], number removed)."""
    status = np.zeros(max(problem.n_pt, 1), np.uint8); mean = np.zeros(max(problem.n_pt, 1))
    st = problem.as_struct()
    removed = lib().oracle_filter_tracks(C.byref(st), max_inlier_reprojection_error, min_triangulation_angle_degrees,
                                         status.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(mean))
    return status[:problem.n_pt], mean[:problem.n_pt], removed


def adjust_tracks(problem, options=None):
    """oracle_adjust_tracks (batched BundleAdjustTrack): updates problem.pt; (status [n_pt] uint8, initial_cost, final_cost, n_failed)."""
    options = options or default_options(use_inner_iterations=0)
    n = max(problem.n_pt, 1)
    status = np.zeros(n, np.uint8); ic = np.zeros(n); fc = np.zeros(n)
    st = problem.as_struct()
    L = lib()
    L.oracle_adjust_tracks.argtypes = [C.POINTER(_abi.tba_options), C.POINTER(_abi.tba_problem), C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    failed = L.oracle_adjust_tracks(C.byref(options), C.byref(st), status.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(ic), _dp(fc))
    return status[:problem.n_pt], ic[:problem.n_pt], fc[:problem.n_pt], failed


def estimate_tracks(problem, options=None, max_reprojection_error_pixels=5.0, min_triangulation_angle_degrees=3.0, bundle_adjustment=True):
    """oracle_estimate_tracks (batched TrackEstimator::EstimateTrack): updates problem.pt; (status [n_pt] uint8, counts [5])."""
    options = options or default_options(use_inner_iterations=0)
    status = np.zeros(max(problem.n_pt, 1), np.uint8); counts = np.zeros(5, np.int32)
    st = problem.as_struct()
    L = lib()
    L.oracle_estimate_tracks.argtypes = [C.POINTER(_abi.tba_options), C.POINTER(_abi.tba_problem), C.c_double, C.c_double, C.c_int,
                                         C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
    L.oracle_estimate_tracks(C.byref(options), C.byref(st), max_reprojection_error_pixels, min_triangulation_angle_degrees, int(bundle_adjustment),
                             status.ctypes.data_as(C.POINTER(C.c_uint8)), counts.ctypes.data_as(C.POINTER(C.c_int32)))
    return status[:problem.n_pt], counts


def two_view_ba_batch(batch):
    """oracle_two_view_ba_batch: every pair solved by oracle_solve; updates the batch in place.
    Returns (termination [n_pairs] uint8, initial_cost, final_cost, iterations)."""
    n = max(batch.n_pairs, 1)
    term = np.zeros(n, np.uint8); ic = np.zeros(n); fc = np.zeros(n); it = np.zeros(n, np.int32)
    st = batch.as_struct()
    L = lib()
    L.oracle_two_view_ba_batch.argtypes = [C.POINTER(_abi.tba_two_view_batch), C.POINTER(C.c_uint8), C.POINTER(C.c_double),
                                           C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    L.oracle_two_view_ba_batch(C.byref(st), term.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(ic), _dp(fc), it.ctypes.data_as(C.POINTER(C.c_int32)))
    n = batch.n_pairs
    return term[:n], ic[:n], fc[:n], it[:n]


def loss(kind, width, s):
    rho = np.zeros(3)
    lib().oracle_loss(kind, width, s, _dp(rho))
    return rho


class Summary:
    def __init__(self, s, iters):
        for f, _ in _abi.tba_summary._fields_:
            if f not in ("iterations", "message"):
                setattr(self, f, getattr(s, f))
        self.message = s.message.decode()
        n = min(s.num_iterations, len(iters))
        self.iterations = [{f: getattr(iters[i], f) for f, _ in _abi.tba_iteration._fields_} for i in range(n)]

    @property
    def costs(self):
        return np.array([it["cost"] for it in self.iterations])


def solve(problem, options=None, max_iterations_logged=1024):
    """oracle_solve: Ceres-semantics LM on the CPU; updates problem.ext/intr/pt in place."""
    options = options or default_options()
    iters = (_abi.tba_iteration * max_iterations_logged)()
    s = _abi.tba_summary()
    s.iterations = C.cast(iters, C.POINTER(_abi.tba_iteration))
    s.iterations_capacity = max_iterations_logged
    st = problem.as_struct()
    rc = lib().oracle_solve(C.byref(options), C.byref(st), C.byref(s))
    out = Summary(s, iters)
    out.rc = rc
    return out


class Oracle:
    """Stage-by-stage access (linearise / prepare / matvec / CG solve / step evaluation)."""

    def __init__(self, problem, options=None):
        self.problem = problem
        self.options = options or default_options()
        self._st = problem.as_struct()
        self._h = lib().oracle_create(C.byref(self.options), C.byref(self._st))

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def linearize(self):
        ok = lib().oracle_linearize(self._h)
        return bool(ok), lib().oracle_cost(self._h)

    def prepare_linear_system(self, radius):
        return bool(lib().oracle_prepare_linear_system(self._h, radius))

    def schur_matvec(self, x_cam, x_intr):
        x_cam = np.ascontiguousarray(x_cam, np.float64); x_intr = np.ascontiguousarray(x_intr, np.float64)
        y_cam = np.zeros_like(x_cam); y_intr = np.zeros_like(x_intr)
        lib().oracle_schur_matvec_split(self._h, _dp(x_cam), _dp(x_intr), _dp(y_cam), _dp(y_intr))
        return y_cam, y_intr

    def solve_linear_system(self):
        ok = lib().oracle_solve_linear_system(self._h)
        return bool(ok), lib().oracle_cg_iterations(self._h), lib().oracle_model_cost_change(self._h)

    def evaluate_step(self):
        c = C.c_double()
        ok = lib().oracle_evaluate_step(self._h, C.byref(c))
        return bool(ok), c.value

    def read(self, which):
        p = self.problem
        n = {_abi.VEC_GRADIENT_CAM: p.n_cam * 6, _abi.VEC_GRADIENT_INTR: p.n_group * 10, _abi.VEC_GRADIENT_PT: p.n_pt * 4,
             _abi.VEC_COLNORM2_CAM: p.n_cam * 6, _abi.VEC_COLNORM2_INTR: p.n_group * 10, _abi.VEC_COLNORM2_PT: p.n_pt * 4,
             _abi.VEC_RESIDUALS: p.n_obs * 2, _abi.VEC_SCHUR_RHS_CAM: p.n_cam * 6, _abi.VEC_SCHUR_RHS_INTR: p.n_group * 10,
             _abi.VEC_PRECOND_CAM: p.n_cam * 36, _abi.VEC_PRECOND_INTR: p.n_group * 100, _abi.VEC_STEP_CAM: p.n_cam * 6,
             _abi.VEC_STEP_INTR: p.n_group * 10, _abi.VEC_STEP_PT: p.n_pt * 4}[which]
        out = np.zeros(n)
        rc = lib().oracle_read(self._h, which, _dp(out), n)
        if rc != 0:
            raise RuntimeError("oracle_read(%d) failed" % which)
        return out
