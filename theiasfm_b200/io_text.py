"""N2: text dataset loaders into the flattened BA problem (``_abi.Problem``) -- harness / tooling, not the hot path.

* ``read_bundler(bundle_file)``: Noah Snavely's Bundler ``bundle.out`` (v0.3), converted exactly as Theia's
  ``ReadBundlerFiles`` does (src/theia/io/read_bundler_files.cc:62-189): PINHOLE cameras with their own intrinsics
  (focal length, k1, k2, principal point (0, 0) -- features are already centred, :116-118), axes flipped by
  diag(1, -1, -1) on rotation and translation (:94-96, :120-127), position = -R^T t, feature = (x, -y) (:155-157), cameras
  with focal length <= 0 dropped together with their observations (:106-110, :160-162), tracks with fewer than two remaining
  views or a zero position dropped (:167-169).
* ``read_bal(path)``: "Bundle Adjustment in the Large" problems (Agarwal et al.), the public large-scale BA benchmark: the same
  camera convention as Bundler (camera looks down -z, p = -P / P_z), angle-axis rotations, one (f, k1, k2) per camera.
* ``write_bal(problem, path)``: the inverse of ``read_bal`` for PINHOLE problems with per-camera intrinsics (tests, data exchange).
"""
import numpy as np

from . import _abi
from .synthetic import rotation_from_angle_axis

_FLIP = np.diag([1.0, -1.0, -1.0])  # bundler_to_theia, read_bundler_files.cc:94-96


def angle_axis_from_rotation(R):
    """Log map through the unit quaternion (Shepperd's branch selection), valid for every rotation including theta = pi --
    a Bundler camera with R = I becomes a 180-degree rotation after the axis flip."""
    R = np.asarray(R, np.float64).reshape(-1, 3, 3)
    out = np.zeros((len(R), 3))
    for i, M in enumerate(R):
        tr = M[0, 0] + M[1, 1] + M[2, 2]
        if tr > 0.0:
            S = 2.0 * np.sqrt(tr + 1.0)
            q = np.array([0.25 * S, (M[2, 1] - M[1, 2]) / S, (M[0, 2] - M[2, 0]) / S, (M[1, 0] - M[0, 1]) / S])
        else:
            a = int(np.argmax([M[0, 0], M[1, 1], M[2, 2]])); b, c = (a + 1) % 3, (a + 2) % 3
            S = 2.0 * np.sqrt(max(1.0 + M[a, a] - M[b, b] - M[c, c], 0.0))
            q = np.zeros(4)
            q[0] = (M[c, b] - M[b, c]) / S
            q[1 + a] = 0.25 * S; q[1 + b] = (M[b, a] + M[a, b]) / S; q[1 + c] = (M[c, a] + M[a, c]) / S
        if q[0] < 0.0:
            q = -q
        n = np.linalg.norm(q[1:])
        out[i] = 0.0 if n == 0.0 else q[1:] * (2.0 * np.arctan2(n, q[0]) / n)
    return out


def _problem(R, t, f, k1, k2, pts, obs_cam, obs_pt, obs_xy, keep_cam):
    """Shared tail: Bundler-convention cameras (x_cam = R X + t, looking down -z) -> Theia convention -> Problem."""
    Rt = np.einsum("ij,njk->nik", _FLIP, R)
    tt = t @ _FLIP.T
    C = -np.einsum("nji,nj->ni", Rt, tt)                       # position = -R^T t
    ext = np.concatenate([C, angle_axis_from_rotation(Rt)], axis=1)
    n_cam = len(R)
    intr = np.zeros((n_cam, _abi.INTR_STRIDE))
    intr[:, 0], intr[:, 1], intr[:, 5], intr[:, 6] = f, 1.0, k1, k2
    # drop cameras with an invalid focal length and everything they observe; re-index
    new_cam = -np.ones(n_cam, np.int64); new_cam[keep_cam] = np.arange(int(keep_cam.sum()))
    sel = keep_cam[obs_cam]
    obs_cam, obs_pt, obs_xy = new_cam[obs_cam[sel]], obs_pt[sel], obs_xy[sel]
    n_pt = len(pts)
    cnt = np.bincount(obs_pt, minlength=n_pt)
    keep_pt = (cnt >= 2) & ((pts ** 2).sum(axis=1) != 0.0)      # underconstrained / unset tracks are not added
    new_pt = -np.ones(n_pt, np.int64); new_pt[keep_pt] = np.arange(int(keep_pt.sum()))
    sel = keep_pt[obs_pt]
    obs_cam, obs_pt, obs_xy = obs_cam[sel], new_pt[obs_pt[sel]], obs_xy[sel]
    ext, intr = ext[keep_cam], intr[keep_cam]
    nc = len(ext)
    xy = obs_xy * np.array([1.0, -1.0])                          # Feature(x, -y)
    pt4 = np.concatenate([pts[keep_pt], np.ones((int(keep_pt.sum()), 1))], axis=1)
    mask = _abi.constant_intrinsics_mask(_abi.MODEL_PINHOLE, _abi.INTR_FOCAL_LENGTH | _abi.INTR_RADIAL_DISTORTION)
    return _abi.Problem(ext, np.zeros(nc, np.uint8), np.arange(nc, dtype=np.int32), np.zeros(nc, np.int32), intr,
                        np.full(nc, mask, np.uint32), pt4, np.zeros(len(pt4), np.uint8), obs_cam.astype(np.int32),
                        obs_pt.astype(np.int32), xy)


def read_bundler(bundle_file):
    with open(bundle_file) as fh:
        tok = [t for line in fh if not line.startswith("#") for t in line.split()]
    it = iter(tok)
    n_cam, n_pt = int(next(it)), int(next(it))
    cam = np.array([float(next(it)) for _ in range(n_cam * 15)]).reshape(n_cam, 15)   # f k1 k2 | R (9) | t (3)
    f, k1, k2 = cam[:, 0], cam[:, 1], cam[:, 2]
    R, t = cam[:, 3:12].reshape(n_cam, 3, 3), cam[:, 12:15]
    pts = np.zeros((n_pt, 3)); oc, op, oxy = [], [], []
    for q in range(n_pt):
        pts[q] = [float(next(it)) for _ in range(3)]
        for _ in range(3):
            next(it)                                                 # colour
        n_view = int(next(it))
        for _ in range(n_view):
            c = int(next(it)); next(it)                              # camera index, SIFT key index
            oc.append(c); op.append(q); oxy.append((float(next(it)), float(next(it))))
    return _problem(R, t, f, k1, k2, pts, np.array(oc, np.int64), np.array(op, np.int64), np.array(oxy, float).reshape(-1, 2), f > 0.0)


def read_bal(path):
    with open(path) as fh:
        tok = fh.read().split()
    n_cam, n_pt, n_obs = int(tok[0]), int(tok[1]), int(tok[2])
    o = np.array(tok[3:3 + 4 * n_obs], float).reshape(n_obs, 4)
    base = 3 + 4 * n_obs
    cam = np.array(tok[base:base + 9 * n_cam], float).reshape(n_cam, 9)       # angle-axis (3), t (3), f, k1, k2
    pts = np.array(tok[base + 9 * n_cam:base + 9 * n_cam + 3 * n_pt], float).reshape(n_pt, 3)
    R = rotation_from_angle_axis(cam[:, :3])
    return _problem(R, cam[:, 3:6], cam[:, 6], cam[:, 7], cam[:, 8], pts, o[:, 0].astype(np.int64), o[:, 1].astype(np.int64), o[:, 2:4],
                    cam[:, 6] > 0.0)


def write_bal(problem, path):
    p = problem
    if not (np.all(p.group_model == _abi.MODEL_PINHOLE) and p.n_group == p.n_cam and np.array_equal(p.cam_group, np.arange(p.n_cam))):
        raise ValueError("BAL stores one PINHOLE (f, k1, k2) camera per view")
    if np.any(p.intr[:, 1] != 1.0) or np.any(p.intr[:, 2:5] != 0.0):
        raise ValueError("BAL cameras have unit aspect ratio, no skew and a centred principal point")
    Rt = rotation_from_angle_axis(p.ext[:, 3:6])
    Rb = np.einsum("ij,njk->nik", _FLIP, Rt)                        # the flip is its own inverse
    tb = -np.einsum("nij,nj->ni", Rb, p.ext[:, :3])
    aa = angle_axis_from_rotation(Rb)
    X = p.pt[:, :3] / p.pt[:, 3:4]
    with open(path, "w") as fh:
        fh.write("%d %d %d\n" % (p.n_cam, p.n_pt, p.n_obs))
        for c, q, (x, y) in zip(p.obs_cam, p.obs_pt, p.obs_xy):
            fh.write("%d %d %.17g %.17g\n" % (c, q, x, -y))
        for i in range(p.n_cam):
            for v in (*aa[i], *tb[i], p.intr[i, 0], p.intr[i, 5], p.intr[i, 6]):
                fh.write("%.17g\n" % v)
        for q in range(p.n_pt):
            for v in X[q]:
                fh.write("%.17g\n" % v)
