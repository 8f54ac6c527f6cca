"""N2: Bundler / BAL text loaders (theiasfm_b200/io_text.py) against the conversion rules of the reference's ReadBundlerFiles
(src/theia/io/read_bundler_files.cc:62-189) and a round trip through the BAL writer."""
import numpy as np

from theiasfm_b200 import _abi, io_text, synthetic


def _bal_scene():
    p = synthetic.make_scene(n_cam=9, n_pt=200, obs_per_pt=5, seed=71, shared_intrinsics=False, perturb=0.0, noise_px=0.3)
    p.intr[:, 3:5] = 0.0                                 # BAL / Bundler measurements are centred on the principal point
    p.obs_xy -= 500.0
    return p


def test_bal_round_trip_preserves_the_residuals(oracle, tmp_path):
    p = _bal_scene()
    path = str(tmp_path / "problem.txt")
    io_text.write_bal(p, path)
    q = io_text.read_bal(path)
    assert q.n_cam == p.n_cam and q.n_pt == p.n_pt and q.n_obs == p.n_obs
    assert np.array_equal(q.obs_cam, p.obs_cam) and np.array_equal(q.obs_pt, p.obs_pt) and np.array_equal(q.obs_xy, p.obs_xy)
    assert np.allclose(q.ext, p.ext, rtol=0, atol=1e-12) and np.allclose(q.intr, p.intr, rtol=1e-15)
    rp, _, okp = oracle.residual_jacobian(p)
    rq, _, okq = oracle.residual_jacobian(q)
    assert okp.all() and okq.all() and np.abs(rp - rq).max() < 1e-9 and np.abs(rp).max() < 3.0   # 0.3 px noise
    assert (q.group_const_mask == _abi.constant_intrinsics_mask(_abi.MODEL_PINHOLE, _abi.INTR_FOCAL_LENGTH | _abi.INTR_RADIAL_DISTORTION)).all()


def test_bundler_conversion_rules(oracle, tmp_path):
    """A hand-written bundle.out: camera 1 looks down -z from the origin, camera 2 is translated; camera 3 has focal length 0
    (dropped with its observations); point 3 then has one view left (dropped); point 4 has a zero position (dropped)."""
    f = 700.0
    X = np.array([[0.2, -0.1, -5.0], [-0.4, 0.3, -6.0], [0.1, 0.1, -4.0], [0.0, 0.0, 0.0]])
    t2 = np.array([-1.0, 0.0, 0.0])                        # Bundler: x_cam = R X + t
    def proj(Xc):                                            # p = -P / P_z, pixel = f p (no distortion)
        return f * (-Xc[:2] / Xc[2])
    lines = ["# Bundle file v0.3", "3 4"]
    for ff, t in ((f, np.zeros(3)), (f, t2), (0.0, np.zeros(3))):
        lines += ["%g 0 0" % ff, "1 0 0", "0 1 0", "0 0 1", "%g %g %g" % tuple(t)]
    views = [[(0, X[0]), (1, X[0] + t2)], [(0, X[1]), (1, X[1] + t2), (2, X[1])], [(0, X[2]), (2, X[2])], [(0, X[0]), (1, X[0] + t2)]]
    for q in range(4):
        lines += ["%g %g %g" % tuple(X[q]), "255 255 255"]
        lines.append(" ".join(["%d" % len(views[q])] + ["%d 0 %.17g %.17g" % ((c,) + tuple(proj(Xc))) for c, Xc in views[q]]))
    path = tmp_path / "bundle.out"
    path.write_text("\n".join(lines) + "\n")
    p = io_text.read_bundler(str(path))
    assert p.n_cam == 2 and p.n_pt == 2 and p.n_obs == 4     # camera 3, point 3 (one view left) and point 4 (zero) are gone
    assert np.allclose(p.ext[0, :3], 0.0) and np.allclose(np.abs(p.ext[0, 3:]), [np.pi, 0, 0], atol=1e-12)   # R = I flips to a half turn about x
    assert np.allclose(p.ext[1, :3], [1.0, 0.0, 0.0], atol=1e-12)    # position = -R^T t, flipped axes
    assert np.allclose(p.pt[:, :3], X[:2]) and np.all(p.pt[:, 3] == 1.0)
    assert np.array_equal(p.intr[:, 0], [f, f]) and np.all(p.intr[:, 3:5] == 0.0)
    r, _, ok = oracle.residual_jacobian(p)                    # exact projections -> zero residuals in Theia's convention
    assert ok.all() and np.abs(r).max() < 1e-9
    # the log map is right for general rotations too
    rng = np.random.default_rng(0)
    w = rng.normal(size=(50, 3)); w *= (rng.uniform(0, np.pi, 50) / np.linalg.norm(w, axis=1))[:, None]
    assert np.abs(io_text.angle_axis_from_rotation(synthetic.rotation_from_angle_axis(w)) - w).max() < 1e-9
    for i in range(p.n_obs):                                  # points are in front of the converted cameras
        _, depth = oracle.project_point(0, p.ext[p.obs_cam[i]], p.intr[p.obs_cam[i]], p.pt[p.obs_pt[i]])
        assert depth > 0
