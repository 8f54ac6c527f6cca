"""ctypes mirror of include/theia_ba_b200.h (tba_options / tba_problem / tba_summary).

Field order and types must match the header exactly; tests/test_abi.py checks the
struct sizes against the values compiled into the library (tba_abi_sizes).
"""
import ctypes as C

import numpy as np

EXT_SIZE = 6
INTR_STRIDE = 10
PT_SIZE = 4

MODEL_PINHOLE = 0
MODEL_PINHOLE_RADIAL_TANGENTIAL = 1
MODEL_FISHEYE = 2
MODEL_FOV = 3
MODEL_DIVISION_UNDISTORTION = 4
MODEL_NUM_PARAMS = {MODEL_PINHOLE: 7, MODEL_PINHOLE_RADIAL_TANGENTIAL: 10, MODEL_FISHEYE: 9, MODEL_FOV: 5, MODEL_DIVISION_UNDISTORTION: 5}

LOSS_TRIVIAL, LOSS_HUBER, LOSS_SOFTLONE, LOSS_CAUCHY, LOSS_ARCTAN, LOSS_TUKEY = range(6)

INTR_NONE = 0x00
INTR_FOCAL_LENGTH = 0x01
INTR_ASPECT_RATIO = 0x02
INTR_SKEW = 0x04
INTR_PRINCIPAL_POINTS = 0x08
INTR_RADIAL_DISTORTION = 0x10
INTR_TANGENTIAL_DISTORTION = 0x20
INTR_ALL = 0x3F

DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR = range(7)
PRECOND_IDENTITY, PRECOND_JACOBI, PRECOND_SCHUR_JACOBI, PRECOND_CLUSTER_JACOBI, PRECOND_CLUSTER_TRIDIAGONAL = range(5)

EXT_POSITION_CONST = 1
EXT_ORIENTATION_CONST = 2
EXT_ALL_CONST = 3

CONVERGENCE, NO_CONVERGENCE, FAILURE = range(3)

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_UNSUPPORTED = -2
ERR_CUDA = -3
ERR_NCCL = -4
ERR_NO_DEVICE = -5

(VEC_GRADIENT_CAM, VEC_GRADIENT_INTR, VEC_GRADIENT_PT, VEC_COLNORM2_CAM, VEC_COLNORM2_INTR, VEC_COLNORM2_PT,
 VEC_RESIDUALS, VEC_SCHUR_RHS_CAM, VEC_SCHUR_RHS_INTR, VEC_PRECOND_CAM, VEC_PRECOND_INTR, VEC_STEP_CAM,
 VEC_STEP_INTR, VEC_STEP_PT) = range(14)


class tba_options(C.Structure):
    _fields_ = [
        ("loss_function_type", C.c_int32),
        ("robust_loss_width", C.c_double),
        ("linear_solver_type", C.c_int32),
        ("preconditioner_type", C.c_int32),
        ("visibility_clustering_type", C.c_int32),
        ("verbose", C.c_int32),
        ("constant_camera_orientation", C.c_int32),
        ("constant_camera_position", C.c_int32),
        ("intrinsics_to_optimize", C.c_int32),
        ("num_threads", C.c_int32),
        ("max_num_iterations", C.c_int32),
        ("max_solver_time_in_seconds", C.c_double),
        ("use_inner_iterations", C.c_int32),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("initial_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("eta", C.c_double),
        ("min_linear_solver_iterations", C.c_int32),
        ("max_linear_solver_iterations", C.c_int32),
        ("jacobi_scaling", C.c_int32),
        ("max_num_consecutive_invalid_steps", C.c_int32),
        ("cg_residual_reset_period", C.c_int32),
    ]


class tba_problem(C.Structure):
    _fields_ = [
        ("n_cam", C.c_int32),
        ("ext", C.POINTER(C.c_double)),
        ("ext_const", C.POINTER(C.c_uint8)),
        ("cam_group", C.POINTER(C.c_int32)),
        ("n_group", C.c_int32),
        ("group_model", C.POINTER(C.c_int32)),
        ("intr", C.POINTER(C.c_double)),
        ("group_const_mask", C.POINTER(C.c_uint32)),
        ("n_pt", C.c_int32),
        ("pt", C.POINTER(C.c_double)),
        ("pt_const", C.POINTER(C.c_uint8)),
        ("n_obs", C.c_int64),
        ("obs_cam", C.POINTER(C.c_int32)),
        ("obs_pt", C.POINTER(C.c_int32)),
        ("obs_xy", C.POINTER(C.c_double)),
    ]


class tba_iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32),
        ("step_is_valid", C.c_int32),
        ("step_is_successful", C.c_int32),
        ("linear_solver_iterations", C.c_int32),
        ("cost", C.c_double),
        ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double),
        ("step_norm", C.c_double),
        ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double),
        ("iteration_time_in_seconds", C.c_double),
    ]


class tba_two_view_batch(C.Structure):
    """include/theia_ba_b200.h: tba_two_view_batch (batched BundleAdjustTwoViews)."""
    _fields_ = [
        ("n_pairs", C.c_int32),
        ("pair_off", C.POINTER(C.c_int64)),
        ("ext1", C.POINTER(C.c_double)),
        ("ext2", C.POINTER(C.c_double)),
        ("intr1", C.POINTER(C.c_double)),
        ("intr2", C.POINTER(C.c_double)),
        ("model1", C.POINTER(C.c_int32)),
        ("model2", C.POINTER(C.c_int32)),
        ("constant_intrinsics1", C.POINTER(C.c_uint8)),
        ("constant_intrinsics2", C.POINTER(C.c_uint8)),
        ("xy1", C.POINTER(C.c_double)),
        ("xy2", C.POINTER(C.c_double)),
        ("points", C.POINTER(C.c_double)),
        ("final_max_reprojection_error_pixels", C.c_double),
        ("inlier", C.POINTER(C.c_uint8)),
    ]


class TwoViewBatch:
    """Host arrays of a batch of two-view BA problems (pair p owns correspondences pair_off[p]:pair_off[p+1])."""

    def __init__(self, pair_off, ext1, ext2, intr1, intr2, model1, model2, const1, const2, xy1, xy2, points):
        self.pair_off = np.ascontiguousarray(pair_off, np.int64)
        self.ext1 = np.ascontiguousarray(ext1, np.float64).reshape(-1, 6)
        self.ext2 = np.ascontiguousarray(ext2, np.float64).reshape(-1, 6).copy()
        self.intr1 = np.ascontiguousarray(intr1, np.float64).reshape(-1, INTR_STRIDE).copy()
        self.intr2 = np.ascontiguousarray(intr2, np.float64).reshape(-1, INTR_STRIDE).copy()
        self.model1 = np.ascontiguousarray(model1, np.int32)
        self.model2 = np.ascontiguousarray(model2, np.int32)
        self.const1 = np.ascontiguousarray(const1, np.uint8)
        self.const2 = np.ascontiguousarray(const2, np.uint8)
        self.xy1 = np.ascontiguousarray(xy1, np.float64).reshape(-1, 2)
        self.xy2 = np.ascontiguousarray(xy2, np.float64).reshape(-1, 2)
        self.points = np.ascontiguousarray(points, np.float64).reshape(-1, 4).copy()
        self.n_pairs = len(self.pair_off) - 1
        self.final_max_reprojection_error_pixels = 0.0      # > 0: also compute self.inlier (post-BA reprojection test)
        self.inlier = None

    def copy(self):
        b = TwoViewBatch(self.pair_off, self.ext1, self.ext2, self.intr1, self.intr2, self.model1, self.model2, self.const1,
                         self.const2, self.xy1, self.xy2, self.points)
        b.final_max_reprojection_error_pixels = self.final_max_reprojection_error_pixels
        return b

    def as_struct(self):
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        s = tba_two_view_batch()
        s.n_pairs = self.n_pairs
        s.pair_off = self.pair_off.ctypes.data_as(C.POINTER(C.c_int64))
        s.ext1, s.ext2, s.intr1, s.intr2 = dp(self.ext1), dp(self.ext2), dp(self.intr1), dp(self.intr2)
        s.model1 = self.model1.ctypes.data_as(C.POINTER(C.c_int32)); s.model2 = self.model2.ctypes.data_as(C.POINTER(C.c_int32))
        s.constant_intrinsics1 = self.const1.ctypes.data_as(C.POINTER(C.c_uint8))
        s.constant_intrinsics2 = self.const2.ctypes.data_as(C.POINTER(C.c_uint8))
        s.xy1, s.xy2, s.points = dp(self.xy1), dp(self.xy2), dp(self.points)
        s.final_max_reprojection_error_pixels = self.final_max_reprojection_error_pixels
        if self.final_max_reprojection_error_pixels > 0.0:
            self.inlier = np.zeros(max(len(self.points), 1), np.uint8)
            s.inlier = self.inlier.ctypes.data_as(C.POINTER(C.c_uint8))
        return s


class tba_summary(C.Structure):
    _fields_ = [
        ("success", C.c_int32),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("setup_time_in_seconds", C.c_double),
        ("solve_time_in_seconds", C.c_double),
        ("termination_type", C.c_int32),
        ("num_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32),
        ("num_linear_solver_iterations", C.c_int32),
        ("num_kernel_launches", C.c_int64),
        ("h2d_bytes", C.c_double),
        ("d2h_bytes", C.c_double),
        ("iterations", C.POINTER(tba_iteration)),
        ("iterations_capacity", C.c_int32),
        ("message", C.c_char * 256),
    ]


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Problem:
    """The flattened BA problem (the IR shared by the engine, the oracle and the adapter).

    Arrays are owned numpy buffers with the exact dtypes/layout of ``tba_problem``;
    ``ext``/``intr``/``pt`` are updated in place by a solve.
    """

    def __init__(self, ext, ext_const, cam_group, group_model, intr, group_const_mask, pt, pt_const,
                 obs_cam, obs_pt, obs_xy):
        self.ext = np.ascontiguousarray(ext, dtype=np.float64).reshape(-1, EXT_SIZE)
        self.n_cam = self.ext.shape[0]
        self.ext_const = np.ascontiguousarray(ext_const, dtype=np.uint8).reshape(self.n_cam)
        self.cam_group = np.ascontiguousarray(cam_group, dtype=np.int32).reshape(self.n_cam)
        self.group_model = np.ascontiguousarray(group_model, dtype=np.int32).reshape(-1)
        self.n_group = self.group_model.shape[0]
        self.intr = np.ascontiguousarray(intr, dtype=np.float64).reshape(self.n_group, INTR_STRIDE)
        self.group_const_mask = np.ascontiguousarray(group_const_mask, dtype=np.uint32).reshape(self.n_group)
        self.pt = np.ascontiguousarray(pt, dtype=np.float64).reshape(-1, PT_SIZE)
        self.n_pt = self.pt.shape[0]
        self.pt_const = np.ascontiguousarray(pt_const, dtype=np.uint8).reshape(self.n_pt)
        self.obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32).reshape(-1)
        self.n_obs = self.obs_cam.shape[0]
        self.obs_pt = np.ascontiguousarray(obs_pt, dtype=np.int32).reshape(self.n_obs)
        self.obs_xy = np.ascontiguousarray(obs_xy, dtype=np.float64).reshape(self.n_obs, 2)

    def copy(self):
        return Problem(self.ext.copy(), self.ext_const.copy(), self.cam_group.copy(), self.group_model.copy(),
                       self.intr.copy(), self.group_const_mask.copy(), self.pt.copy(), self.pt_const.copy(),
                       self.obs_cam.copy(), self.obs_pt.copy(), self.obs_xy.copy())

    def as_struct(self):
        p = tba_problem()
        p.n_cam = self.n_cam
        p.ext = _ptr(self.ext, C.c_double)
        p.ext_const = _ptr(self.ext_const, C.c_uint8)
        p.cam_group = _ptr(self.cam_group, C.c_int32)
        p.n_group = self.n_group
        p.group_model = _ptr(self.group_model, C.c_int32)
        p.intr = _ptr(self.intr, C.c_double)
        p.group_const_mask = _ptr(self.group_const_mask, C.c_uint32)
        p.n_pt = self.n_pt
        p.pt = _ptr(self.pt, C.c_double)
        p.pt_const = _ptr(self.pt_const, C.c_uint8)
        p.n_obs = self.n_obs
        p.obs_cam = _ptr(self.obs_cam, C.c_int32)
        p.obs_pt = _ptr(self.obs_pt, C.c_int32)
        p.obs_xy = _ptr(self.obs_xy, C.c_double)
        return p

    def shard(self, rank, world_size):
        """Points [begin,end) of this rank (balanced by observation count) with their observations.

        Cameras / intrinsics groups are replicated; point indices are re-based to the shard.
        Mirrors tba_shard_points() in the C-ABI.
        """
        counts = np.bincount(self.obs_pt, minlength=self.n_pt).astype(np.int64)
        begin, end = shard_points(counts, world_size, rank)
        sel = (self.obs_pt >= begin) & (self.obs_pt < end)
        return Problem(self.ext.copy(), self.ext_const, self.cam_group, self.group_model, self.intr.copy(),
                       self.group_const_mask, self.pt[begin:end].copy(), self.pt_const[begin:end],
                       self.obs_cam[sel], self.obs_pt[sel] - begin, self.obs_xy[sel]), begin, end


def shard_points(pt_num_obs, world_size, rank):
    """Same split as tba_shard_points(): cut the prefix sum of per-point observation counts evenly."""
    cum = np.concatenate([[0], np.cumsum(np.asarray(pt_num_obs, dtype=np.int64))])
    total = int(cum[-1])
    n_pt = len(pt_num_obs)

    def cut(r):
        if r <= 0:
            return 0
        if r >= world_size:
            return n_pt
        target = (total * r) // world_size
        return int(np.searchsorted(cum, target, side="left"))

    return cut(rank), cut(rank + 1)


def constant_intrinsics_mask(model, intrinsics_to_optimize):
    """CameraIntrinsicsModel::GetSubsetFromOptimizeIntrinsicsType as a bitmask of CONSTANT indices
    (pinhole_camera_model.cc:132-162, pinhole_radial_tangential_camera_model.cc:150-188)."""
    m = 0
    if intrinsics_to_optimize == INTR_ALL:
        return 0
    if model in (MODEL_FOV, MODEL_DIVISION_UNDISTORTION):   # f, aspect, cx, cy, one distortion term (fov_camera_model.cc, division_...cc)
        if not intrinsics_to_optimize & INTR_FOCAL_LENGTH:
            m |= 1 << 0
        if not intrinsics_to_optimize & INTR_ASPECT_RATIO:
            m |= 1 << 1
        if not intrinsics_to_optimize & INTR_PRINCIPAL_POINTS:
            m |= (1 << 2) | (1 << 3)
        if not intrinsics_to_optimize & INTR_RADIAL_DISTORTION:
            m |= 1 << 4
        return m
    if not intrinsics_to_optimize & INTR_FOCAL_LENGTH:
        m |= 1 << 0
    if not intrinsics_to_optimize & INTR_ASPECT_RATIO:
        m |= 1 << 1
    if not intrinsics_to_optimize & INTR_SKEW:
        m |= 1 << 2
    if not intrinsics_to_optimize & INTR_PRINCIPAL_POINTS:
        m |= (1 << 3) | (1 << 4)
    if model == MODEL_PINHOLE:
        if not intrinsics_to_optimize & INTR_RADIAL_DISTORTION:
            m |= (1 << 5) | (1 << 6)
    elif model == MODEL_FISHEYE:                              # four radial terms, no tangential (fisheye_camera_model.cc)
        if not intrinsics_to_optimize & INTR_RADIAL_DISTORTION:
            m |= (1 << 5) | (1 << 6) | (1 << 7) | (1 << 8)
    else:
        if not intrinsics_to_optimize & INTR_RADIAL_DISTORTION:
            m |= (1 << 5) | (1 << 6) | (1 << 7)
        if not intrinsics_to_optimize & INTR_TANGENTIAL_DISTORTION:
            m |= (1 << 8) | (1 << 9)
    return m
