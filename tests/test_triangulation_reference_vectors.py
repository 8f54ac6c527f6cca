"""Pins the track-estimation restatement (oracle_estimate_tracks) AND the product's device bodies (host-compiled,
tests/host_point_lm.cc) to the reference's own known-answer tests for the two geometric pieces of
TrackEstimator::EstimateTrack: src/theia/sfm/triangulation/triangulation_test.cc
  TriangulationMidpoint.BasicTest (:328-351)        two views, exact projections, squared reprojection error <= 1e-12
  TriangulationNView-style many-view scene (:138-218) reused for the midpoint method with the same poses / points
  SufficientTriangulationAngle.* (:432-497)         rays on a unit circle at known angles
The cameras are given identity calibration so that pixels are the normalised image points the reference test uses."""
import numpy as np
import pytest

from test_track_estimator import H, run_host  # noqa: F401  (H is a fixture)
from theiasfm_b200 import _abi, synthetic

IDENT = np.array([[1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]])


def rot_y(a):
    return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])


def axis_angle_matrix(deg, axis):
    k = np.asarray(axis, float); k /= np.linalg.norm(k)
    a = np.radians(deg)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def problem(Rs, ts, pixels_per_point):
    """cameras x_cam = R X + t with identity calibration; pixels_per_point: list of [n_view][2] arrays."""
    Rs = np.asarray(Rs); ts = np.asarray(ts)
    C = -np.einsum("nji,nj->ni", Rs, ts)
    ext = np.concatenate([C, synthetic.angle_axis_from_rotation(Rs)], axis=1)
    n_cam, n_pt = len(Rs), len(pixels_per_point)
    obs_cam = np.tile(np.arange(n_cam, dtype=np.int32), n_pt)
    obs_pt = np.repeat(np.arange(n_pt, dtype=np.int32), n_cam)
    xy = np.concatenate(pixels_per_point, axis=0)
    return _abi.Problem(ext, np.full(n_cam, _abi.EXT_ALL_CONST, np.uint8), np.zeros(n_cam, np.int32), [_abi.MODEL_PINHOLE], IDENT.copy(),
                        [0x7F], np.zeros((n_pt, 4)), np.zeros(n_pt, np.uint8), obs_cam, obs_pt, xy)


def sq_reprojection_errors(Rs, ts, X4, pixels):
    out = []
    for R, t, px in zip(Rs, ts, pixels):
        q = R @ X4[:3] + t * X4[3]
        out.append(((q[:2] / q[2] - px) ** 2).sum())
    return np.array(out)


def both(H, oracle, p, **kw):
    q = p.copy()
    st_o, _ = oracle.estimate_tracks(q, oracle.default_options(use_inner_iterations=0), kw.get("max_px", 5.0), kw.get("min_angle", 3.0), False)
    out, st = run_host(H, p, ba=False, max_px=kw.get("max_px", 5.0), min_angle=kw.get("min_angle", 3.0))
    return (st_o, q.pt), (st, out)


def test_midpoint_basic_test_of_the_reference(H, oracle):
    t = np.array([-3.0, 1.5, 11.0]); t /= np.linalg.norm(t)                # pose1 = [R | t.normalized()], pose2 = identity
    Rs, ts = [rot_y(0.15), np.eye(3)], [t, np.zeros(3)]
    pts = [np.array([5.0, 20.0, 23.0]), np.array([-6.0, 16.0, 33.0])]
    pix = [np.array([(R @ X + tt)[:2] / (R @ X + tt)[2] for R, tt in zip(Rs, ts)]) for X in pts]
    p = problem(Rs, ts, pix)
    for st, out in both(H, oracle, p, max_px=1e-5, min_angle=0.1):
        assert (st == 0).all()
        for X4, px in zip(out, pix):
            assert (sq_reprojection_errors(Rs, ts, X4, px) <= 1e-12).all()  # kReprojectionTolerance


def test_midpoint_on_the_reference_many_view_scene(H, oracle):
    rot = [(7, (0, 0, 1)), (12, (0, 1, 0)), (15, (1, 0, 0)), (20, (1, 0, 1)), (11, (0, 1, 1)), (0, (1, 1, 1)), (5, (0, 1, 1)), (0, (1, 1, 1))]
    Rs = [axis_angle_matrix(d, a) for d, a in rot]
    ts = [np.array(v, float) for v in [(1, 1, 1), (3, 2, 13), (4, 5, 11), (1, 2, 15), (3, 1.5, 91), (1, 7, 11), (0, 0, 0), (0, 0, 0)]]
    pts = np.array([[-1.62, -2.99, 6.12], [4.42, -1.53, 9.83], [1.45, -0.59, 5.29], [1.89, -1.10, 8.22], [-0.21, 2.38, 5.63], [0.61, -0.97, 7.49],
                    [0.48, 0.70, 8.94], [1.65, -2.56, 8.63], [2.44, -0.20, 7.78], [2.84, -2.58, 7.35], [-1.35, -2.84, 7.33], [-0.42, 1.54, 8.86],
                    [2.56, 1.72, 7.86], [1.75, -1.39, 5.73], [2.08, -3.91, 8.37], [-0.91, 1.36, 9.16], [2.84, 1.54, 8.74], [-1.01, 3.02, 8.18],
                    [-3.73, -0.62, 7.81], [-2.98, -1.88, 6.23], [2.39, -0.19, 6.47], [-0.63, -1.05, 7.11], [-1.76, -0.55, 5.18], [-3.19, 3.27, 8.18],
                    [0.31, -2.77, 7.54], [0.54, -3.77, 9.77]])
    pix = [np.array([(R @ X + t)[:2] / (R @ X + t)[2] for R, t in zip(Rs, ts)]) for X in pts]
    p = problem(Rs, ts, pix)
    for st, out in both(H, oracle, p, max_px=1e-5, min_angle=0.1):
        assert (st == 0).all()
        for X4, px, X in zip(out, pix, pts):
            assert (sq_reprojection_errors(Rs, ts, X4, px) <= 1e-12).all()
            assert np.abs(X4[:3] / X4[3] - X).max() < 1e-9


def ray_cameras(angles_deg):
    """Cameras at the origin whose principal ray (pixel (0,0), identity calibration) is (cos a, sin a, 0)."""
    a = np.radians(np.asarray(angles_deg, float))
    Rs = np.stack([np.stack([-np.sin(a), np.cos(a), 0 * a], 1), np.tile([0.0, 0.0, 1.0], (len(a), 1)), np.stack([np.cos(a), np.sin(a), 0 * a], 1)], 1)
    return problem(Rs, np.zeros((len(a), 3)), [np.zeros((len(a), 2))])


def test_sufficient_triangulation_angle_cases_of_the_reference(H, oracle):
    def sufficient(angles):
        (st_o, _), (st, _) = both(H, oracle, ray_cameras(angles), min_angle=4.0)
        assert st_o[0] == st[0]
        return st[0] != 1                                                   # 1 = rejected by the angle test
    for n in range(2, 50):
        assert sufficient(np.arange(n) * 5.0)                               # AllSufficient
        assert not sufficient(np.arange(n) * (4.0 / (n + 1e-4)))            # AllInsufficient
    assert sufficient([0.0, 5.0, 1.0])                                      # SomeInsufficient
    assert not sufficient([0.0, 1.0])                                       # TwoInsufficient
