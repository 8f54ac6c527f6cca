"""Seeded synthetic BA scenes (SURVEY.md section 8d recipe; idiom of the reference's
global_pose_estimation/nonlinear_position_estimator_test.cc:66-74,181-200).

Cameras on a wavy ring of radius R looking at the origin, points in a ball of radius 0.4 R,
each point observed by L cameras drawn (stratified) from a window of angularly-near cameras,
observations = exact projection + pixel noise, initial estimate = ground truth + perturbation.
Everything is float64 numpy; the projection here is an independent vectorised restatement of
Camera::ProjectPoint (camera.cc:204-213) used only to synthesise measurements.
"""
import ctypes

import numpy as np

from . import _abi

try:  # keep freed large blocks in the heap: re-faulting fresh pages for every numpy temporary dominates generation
    _libc = ctypes.CDLL("libc.so.6")
    _libc.mallopt(-3, (1 << 31) - 1)  # M_MMAP_THRESHOLD
    _libc.mallopt(-1, (1 << 31) - 1)  # M_TRIM_THRESHOLD
except OSError:  # pragma: no cover
    pass

CONFIGS = {
    # BASELINE.json configs[0..3]
    "c1_50cam": dict(n_cam=50, n_pt=5_000, obs_per_pt=10, model=_abi.MODEL_PINHOLE, shared_intrinsics=True, seed=20240608),
    "c2_1kcam": dict(n_cam=1_000, n_pt=200_000, obs_per_pt=10, model=_abi.MODEL_PINHOLE, shared_intrinsics=True, seed=20240609),
    "c3_10kcam": dict(n_cam=10_000, n_pt=2_000_000, obs_per_pt=10, model=_abi.MODEL_PINHOLE, shared_intrinsics=True, seed=20240610),
    "c4_radtan": dict(n_cam=1_000, n_pt=500_000, obs_per_pt=10, model=_abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, shared_intrinsics=False, seed=20240611),
}


def rotation_from_angle_axis(w):
    """Rodrigues, vectorised: w [n,3] -> R [n,3,3] (world -> camera)."""
    w = np.asarray(w, dtype=np.float64).reshape(-1, 3)
    theta = np.linalg.norm(w, axis=1)
    small = theta < 1e-12
    k = w / np.where(small, 1.0, theta)[:, None]
    K = np.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(theta)[:, None, None], np.cos(theta)[:, None, None]
    KK = np.einsum("nij,njk->nik", K, K)  # einsum, not @: batched 3x3 matmul through BLAS is pathologically slow
    R = np.eye(3)[None] + s * K + (1.0 - c) * KK
    R[small] = np.eye(3)
    return R


def angle_axis_from_rotation(R):
    """Log map, vectorised; valid away from theta = pi (the generator never goes there)."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    cos_t = np.clip((np.trace(R, axis1=1, axis2=2) - 1.0) / 2.0, -1.0, 1.0)
    theta = np.arccos(cos_t)
    v = np.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], axis=1)
    s = np.sin(theta)
    f = np.where(s > 1e-12, theta / np.where(s > 1e-12, 2.0 * s, 1.0), 0.5)
    return v * f[:, None]


def project(model, ext, intr, pt, R=None):
    """Vectorised Camera::ProjectPoint: ext [n,6], intr [n,10], pt [n,4] -> pix [n,2], depth [n].
    R: optional precomputed world->camera rotations [n,3,3]."""
    a = pt[:, :3] - pt[:, 3:4] * ext[:, :3]
    if R is None:
        R = rotation_from_angle_axis(ext[:, 3:6])
    q = np.stack([(R[:, i, :] * a).sum(axis=1) for i in range(3)], axis=1)
    if model >= _abi.MODEL_FISHEYE:
        return _project_other_models(model, intr, q), q[:, 2] / pt[:, 3]
    u, v = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
    r2 = u * u + v * v
    if model == _abi.MODEL_PINHOLE:
        d = 1.0 + r2 * (intr[:, 5] + intr[:, 6] * r2)
        ud, vd = u * d, v * d
    else:
        rd = 1.0 + intr[:, 5] * r2 + intr[:, 6] * r2 * r2 + intr[:, 7] * r2 * r2 * r2
        tx = intr[:, 9] * (r2 + 2.0 * u * u) + 2.0 * intr[:, 8] * u * v
        ty = intr[:, 8] * (r2 + 2.0 * v * v) + 2.0 * intr[:, 9] * u * v
        ud, vd = u * rd + tx, v * rd + ty
    px = intr[:, 0] * ud + intr[:, 2] * vd + intr[:, 3]
    py = intr[:, 0] * intr[:, 1] * vd + intr[:, 4]
    return np.stack([px, py], axis=1), q[:, 2] / pt[:, 3]


def _project_other_models(model, k, q):
    """CameraToPixelCoordinates of FISHEYE / FOV / DIVISION_UNDISTORTION (fisheye_camera_model.h:160-270,
    fov_camera_model.h:157-258, division_undistortion_camera_model.h:171-286), vectorised; scene synthesis only."""
    with np.errstate(all="ignore"):
        if model == _abi.MODEL_FISHEYE:
            r = np.sqrt(q[:, 0] ** 2 + q[:, 1] ** 2)
            theta = np.arctan2(r, np.abs(q[:, 2]))
            t2 = theta * theta
            theta_d = theta * (1.0 + k[:, 5] * t2 + k[:, 6] * t2 * t2 + k[:, 7] * t2 ** 3 + k[:, 8] * t2 ** 4)
            s = np.where(r * r < 1e-8, 1.0, theta_d / np.where(r > 0, r, 1.0)) * np.where((q[:, 2] < 0) & (r * r >= 1e-8), -1.0, 1.0)
            ud, vd = s * q[:, 0], s * q[:, 1]
            return np.stack([k[:, 0] * ud + k[:, 2] * vd + k[:, 3], k[:, 0] * k[:, 1] * vd + k[:, 4]], axis=1)
        u, v = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
        if model == _abi.MODEL_FOV:
            w, r2 = k[:, 4], u * u + v * v
            ru = np.sqrt(r2)
            th = np.tan(w / 2.0)
            r_d = np.where(w < 1e-3, (w * w * r2) / 3.0 - w * w / 12.0 + 1.0,
                           np.where(r2 < 1e-3, (-2.0 * th * (4.0 * r2 * th * th - 3.0)) / (3.0 * np.where(w != 0, w, 1.0)),
                                    np.arctan(2.0 * ru * th) / np.where(ru * w != 0, ru * w, 1.0)))
            return np.stack([k[:, 0] * (r_d * u) + k[:, 2], k[:, 0] * k[:, 1] * (r_d * v) + k[:, 3]], axis=1)
        up0, up1 = k[:, 0] * u, k[:, 0] * k[:, 1] * v
        r2 = up0 * up0 + up1 * up1
        denom, inner = 2.0 * k[:, 4] * r2, 1.0 - 4.0 * k[:, 4] * r2
        ident = (np.abs(denom) < np.finfo(float).eps) | (inner < 0.0)
        scale = np.where(ident, 1.0, (1.0 - np.sqrt(np.where(inner > 0, inner, 0.0))) / np.where(ident, 1.0, denom))
        return np.stack([up0 * scale + k[:, 2], up1 * scale + k[:, 3]], axis=1)


def make_scene(n_cam, n_pt, obs_per_pt=10, model=_abi.MODEL_PINHOLE, shared_intrinsics=True, seed=0,
               noise_px=0.5, intrinsics_to_optimize=_abi.INTR_FOCAL_LENGTH | _abi.INTR_RADIAL_DISTORTION,
               perturb=1.0, return_truth=False):
    """Build a seeded scene; returns a ``Problem`` holding the perturbed initial estimate
    (all cameras / points variable, intrinsics masks from ``intrinsics_to_optimize``)."""
    rng = np.random.default_rng(seed)
    L = int(obs_per_pt)
    R0 = 10.0 * (n_cam / 50.0) ** (1.0 / 3.0)
    # --- ground-truth cameras: wavy ring, looking at the origin
    phi = 2.0 * np.pi * (np.arange(n_cam) + 0.25 * rng.uniform(-1, 1, n_cam)) / n_cam
    C = np.stack([R0 * np.cos(phi), R0 * np.sin(phi), 0.15 * R0 * np.sin(3.0 * phi)], axis=1)
    z = -C / np.linalg.norm(C, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(np.broadcast_to(up, z.shape), z)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    Rgt = np.stack([x, y, z], axis=1)  # rows = camera axes => world -> camera
    w = angle_axis_from_rotation(Rgt) + rng.uniform(-0.05, 0.05, (n_cam, 3))
    ext_gt = np.concatenate([C, w], axis=1)
    # --- ground-truth intrinsics
    n_group = 1 if shared_intrinsics else n_cam
    cam_group = np.zeros(n_cam, dtype=np.int32) if shared_intrinsics else np.arange(n_cam, dtype=np.int32)
    intr_gt = np.zeros((n_group, _abi.INTR_STRIDE))
    intr_gt[:, 0], intr_gt[:, 1], intr_gt[:, 3], intr_gt[:, 4] = 800.0, 1.0, 500.0, 500.0
    intr_gt[:, 5], intr_gt[:, 6] = -0.05, 0.01
    if model == _abi.MODEL_PINHOLE_RADIAL_TANGENTIAL:
        intr_gt[:, 7], intr_gt[:, 8], intr_gt[:, 9] = 0.001, 1e-3, -5e-4
    elif model == _abi.MODEL_FISHEYE:          # f, aspect, skew, cx, cy, k1..k4
        intr_gt[:, 5:9] = [-0.02, 0.004, -0.001, 0.0002]
    elif model == _abi.MODEL_FOV:              # f, aspect, cx, cy, omega
        intr_gt[:, :] = 0.0
        intr_gt[:, 0], intr_gt[:, 1], intr_gt[:, 2], intr_gt[:, 3], intr_gt[:, 4] = 800.0, 1.0, 500.0, 500.0, 0.35
    elif model == _abi.MODEL_DIVISION_UNDISTORTION:   # f, aspect, cx, cy, k (pixel units)
        intr_gt[:, :] = 0.0
        intr_gt[:, 0], intr_gt[:, 1], intr_gt[:, 2], intr_gt[:, 3], intr_gt[:, 4] = 800.0, 1.0, 500.0, 500.0, -2e-7
    # --- ground-truth points: uniform in a ball of radius 0.4 R0
    d = rng.normal(size=(n_pt, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    X = d * (0.4 * R0 * rng.uniform(0, 1, (n_pt, 1)) ** (1.0 / 3.0))
    pt_gt = np.concatenate([X, np.ones((n_pt, 1))], axis=1)
    # --- visibility: L cameras, one per stratum of a window of angularly-near cameras
    order = np.argsort(phi)
    psi = np.arctan2(X[:, 1], X[:, 0]) % (2.0 * np.pi)
    i0 = np.searchsorted(phi[order] % (2.0 * np.pi), psi) % n_cam
    half = min(max(L, int(round(0.06 * n_cam))), (n_cam - 1) // 2)
    width = 2 * half + 1
    if width < L:
        raise ValueError("need at least obs_per_pt cameras")
    # integer strata => the L cameras of a point are distinct (a view observes a track once: view.h:79-85)
    base = (np.arange(L + 1) * width) // L
    size = (base[1:] - base[:-1])[None, :]
    offs = base[None, :-1] + np.floor(rng.uniform(0, 1, (n_pt, L)) * size).astype(np.int64) - half
    obs_cam = order[(i0[:, None] + offs) % n_cam].astype(np.int32).reshape(-1)
    obs_pt = np.repeat(np.arange(n_pt, dtype=np.int32), L)
    R_cam = rotation_from_angle_axis(ext_gt[:, 3:6])
    pix, depth = project(model, ext_gt[obs_cam], intr_gt[cam_group[obs_cam]], pt_gt[obs_pt], R=R_cam[obs_cam])
    del R_cam
    keep = (depth > 0) & (pix[:, 0] >= 0) & (pix[:, 0] <= 1000) & (pix[:, 1] >= 0) & (pix[:, 1] <= 1000)
    obs_cam, obs_pt, pix = obs_cam[keep], obs_pt[keep], pix[keep]
    obs_xy = pix + noise_px * rng.normal(size=pix.shape)
    # --- initial estimate = perturbed ground truth
    ext0 = ext_gt.copy()
    ext0[:, :3] += perturb * 0.01 * R0 * rng.normal(size=(n_cam, 3))
    ext0[:, 3:] += perturb * 0.005 * rng.normal(size=(n_cam, 3))
    pt0 = pt_gt.copy()
    pt0[:, :3] += perturb * 0.005 * R0 * rng.normal(size=(n_pt, 3))
    intr0 = intr_gt.copy()
    intr0[:, 0] *= 1.0 + perturb * 0.01 * rng.normal(size=n_group)
    if perturb and model <= _abi.MODEL_FISHEYE:
        intr0[:, 5] = 0.0
        intr0[:, 6] = 0.0
    elif perturb:   # FOV / DIVISION_UNDISTORTION: the single distortion term starts 20 % off
        intr0[:, 4] *= 0.8
    group_model = np.full(n_group, model, dtype=np.int32)
    mask = np.full(n_group, _abi.constant_intrinsics_mask(model, intrinsics_to_optimize), dtype=np.uint32)
    prob = _abi.Problem(ext0, np.zeros(n_cam, np.uint8), cam_group, group_model, intr0, mask, pt0,
                        np.zeros(n_pt, np.uint8), obs_cam, obs_pt, obs_xy)
    if return_truth:
        return prob, dict(ext=ext_gt, intr=intr_gt, pt=pt_gt, radius=R0)
    return prob


def make_config(name, **overrides):
    kw = dict(CONFIGS[name])
    kw.update(overrides)
    return make_scene(**kw)


def make_two_view_batch(n_pairs, min_corr=60, max_corr=400, seed=0, models=(_abi.MODEL_PINHOLE,), noise_px=0.5, free_focal_fraction=0.5):
    """Seeded batch of two-view BA problems (BundleAdjustTwoViews inputs): camera 1 at a fixed pose, camera 2 displaced by a
    baseline and a small rotation, points in front of both, noisy correspondences, noisy triangulated points, camera 2 pose
    and (when its intrinsics are free) its focal length perturbed."""
    rng = np.random.default_rng(seed)
    n = rng.integers(min_corr, max_corr + 1, n_pairs)
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    model = np.asarray(models)[rng.integers(0, len(models), n_pairs)].astype(np.int32)
    ext1 = np.concatenate([0.1 * rng.normal(size=(n_pairs, 3)), 0.05 * rng.normal(size=(n_pairs, 3))], axis=1)
    ext2_gt = ext1.copy()
    ext2_gt[:, :3] += np.stack([rng.uniform(0.6, 1.4, n_pairs) * rng.choice([-1, 1], n_pairs), 0.15 * rng.normal(size=n_pairs), 0.1 * rng.normal(size=n_pairs)], 1)
    ext2_gt[:, 3:] += 0.08 * rng.normal(size=(n_pairs, 3))
    intr_gt = np.zeros((n_pairs, _abi.INTR_STRIDE))
    for p in range(n_pairs):
        m = model[p]
        if m <= _abi.MODEL_FISHEYE:
            intr_gt[p, :5] = [rng.uniform(600, 1000), 1.0, 0.0, 500.0, 500.0]
            intr_gt[p, 5:7] = [-0.03, 0.005]
        else:
            intr_gt[p, :4] = [rng.uniform(600, 1000), 1.0, 500.0, 500.0]
            intr_gt[p, 4] = 0.3 if m == _abi.MODEL_FOV else -1.5e-7
    const1 = np.ones(n_pairs, np.uint8)
    const2 = (rng.uniform(size=n_pairs) >= free_focal_fraction).astype(np.uint8)
    const1[rng.uniform(size=n_pairs) < 0.2 * free_focal_fraction] = 0
    total = int(off[-1])
    X = np.concatenate([rng.uniform(-2, 2, (total, 2)), rng.uniform(5, 9, (total, 1)), np.ones((total, 1))], axis=1)
    pair_of = np.repeat(np.arange(n_pairs), n)
    xy1 = np.zeros((total, 2)); xy2 = np.zeros((total, 2))
    for m in np.unique(model):
        sel = model[pair_of] == m
        xy1[sel] = project(int(m), ext1[pair_of[sel]], intr_gt[pair_of[sel]], X[sel])[0]
        xy2[sel] = project(int(m), ext2_gt[pair_of[sel]], intr_gt[pair_of[sel]], X[sel])[0]
    xy1 += noise_px * rng.normal(size=xy1.shape); xy2 += noise_px * rng.normal(size=xy2.shape)
    ext2 = ext2_gt + np.concatenate([0.03 * rng.normal(size=(n_pairs, 3)), 0.005 * rng.normal(size=(n_pairs, 3))], axis=1)
    intr1, intr2 = intr_gt.copy(), intr_gt.copy()
    intr2[const2 == 0, 0] *= 1.0 + 0.03 * rng.normal(size=int((const2 == 0).sum()))
    pts = X.copy(); pts[:, :3] += 0.05 * rng.normal(size=(total, 3))
    return _abi.TwoViewBatch(off, ext1, ext2, intr1, intr2, model, model, const1, const2, xy1, xy2, pts)
