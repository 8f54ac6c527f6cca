"""Ports of the reference's own tests that touch the hot path's projection functions:
  pinhole_camera_model_test.cc:218-298          (ReprojectionTest, 3 distortion settings)
  pinhole_radial_tangential_camera_model_test.cc:241-343
  pinhole_camera_model_test.cc:158-213          (GetSubsetFromOptimizeIntrinsicsType)
  camera_test.cc:183-227                        (ProjectPoint o PixelToUnitDepthRay, tol 1e-5)
  fisheye_camera_model_test.cc:227-327, fov_camera_model_test.cc:192-285, division_undistortion_camera_model_test.cc:320-415
run against the oracle's restatement (same grids, same tolerances)."""
import numpy as np
import pytest

from theiasfm_b200 import _abi


def _reprojection_test(oracle, model, intr):
    tol = 1e-5
    ntol = tol / intr[0]
    # image -> camera -> image on the reference's 1200x980 grid (step 10), depths 2..24
    for x in np.arange(0.0, 1200.0, 50.0):       # grid thinned 5x per axis to keep the CPU suite fast
        for y in np.arange(0.0, 980.0, 50.0):
            ray = oracle.pixel_to_camera(model, intr, [x, y])
            for depth in (2.0, 7.0, 24.0):
                pix = oracle.camera_to_pixel(model, intr, ray * depth)
                assert np.hypot(pix[0] - x, pix[1] - y) < tol
    # camera -> image -> camera
    for x in np.arange(-0.8, 0.8, 0.2):
        for y in np.arange(-0.8, 0.8, 0.2):
            for depth in (2.0, 11.0, 24.0):
                pt = np.array([x, y, depth])
                pix = oracle.camera_to_pixel(model, intr, pt)
                ray = oracle.pixel_to_camera(model, intr, pix)
                assert np.linalg.norm(pt - ray * depth) < ntol * depth * 2.0 + 1e-9


@pytest.mark.parametrize("k1,k2", [(0.0, 0.0), (0.01, 0.0), (0.01, 0.001)])
def test_pinhole_reprojection(oracle, k1, k2):
    intr = np.array([1200.0, 1.0, 0.0, 600.0, 400.0, k1, k2, 0, 0, 0])
    _reprojection_test(oracle, _abi.MODEL_PINHOLE, intr)


@pytest.mark.parametrize("rad,tan", [((0, 0, 0), (0, 0)), ((0.01, 0, 0), (0, 0)), ((0.01, 0.001, 0.0001), (0, 0)),
                                     ((0, 0, 0), (0.01, 0.001)), ((0.01, 0.001, 0.0001), (0.01, 0.001))])
def test_radtan_reprojection(oracle, rad, tan):
    intr = np.array([1200.0, 1.0, 0.0, 600.0, 400.0, *rad, *tan])
    _reprojection_test(oracle, _abi.MODEL_PINHOLE_RADIAL_TANGENTIAL, intr)


def _reprojection_test_ref_tol(oracle, model, intr, tol, height):
    """ReprojectionTest of the three other models, with the reference's own tolerances (kTolerance on pixels,
    kTolerance / focal on camera points); grids thinned to keep the CPU suite fast."""
    ntol = tol / intr[0]
    for x in np.arange(0.0, 1200.0, 50.0):
        for y in np.arange(0.0, height, 50.0):
            ray = oracle.pixel_to_camera(model, intr, [x, y])
            for depth in (2.0, 7.0, 24.0):
                pix = oracle.camera_to_pixel(model, intr, ray * depth)
                assert np.hypot(pix[0] - x, pix[1] - y) < tol, (x, y, depth, pix)
    for x in np.arange(-0.8, 0.8, 0.1):
        for y in np.arange(-0.8, 0.8, 0.1):
            for depth in (2.0, 11.0, 24.0):
                pt = np.array([x, y, depth])
                ray = oracle.pixel_to_camera(model, intr, oracle.camera_to_pixel(model, intr, pt))
                assert np.linalg.norm(pt - ray * depth) < ntol, (x, y, depth)


@pytest.mark.parametrize("rad", [(0, 0, 0, 0), (0.01, 0, 0, 0), (0.01, 0.001, 0, 0), (0.01, 0.001, 0.001, 0), (0.01, 0.001, 0.001, 0.001)])
def test_fisheye_reprojection(oracle, rad):
    intr = np.array([1200.0, 1.0, 0.0, 600.0, 400.0, *rad, 0.0])
    _reprojection_test_ref_tol(oracle, _abi.MODEL_FISHEYE, intr, 1e-5, 980.0)


@pytest.mark.parametrize("omega", [0.0, 0.0001, 0.001, 0.1])
def test_fov_reprojection(oracle, omega):
    intr = np.array([1200.0, 1.0, 600.0, 400.0, omega, 0, 0, 0, 0, 0])
    _reprojection_test_ref_tol(oracle, _abi.MODEL_FOV, intr, 1e-5, 980.0)


@pytest.mark.parametrize("k", [0.0, -1e-8, -1e-7, -1e-6])
def test_division_undistortion_reprojection(oracle, k):
    intr = np.array([1200.0, 1.0, 600.0, 400.0, k, 0, 0, 0, 0, 0])
    _reprojection_test_ref_tol(oracle, _abi.MODEL_DIVISION_UNDISTORTION, intr, 1e-6, 800.0)


def test_constant_subset_masks_of_the_other_models(oracle):
    """GetSubsetFromOptimizeIntrinsicsType tests: fisheye_camera_model_test.cc:168-225, fov_camera_model_test.cc:141-189,
    division_undistortion_camera_model_test.cc:163-216."""
    L = oracle.lib()
    bits = lambda m: [j for j in range(10) if (m >> j) & 1]
    F, V, D = _abi.MODEL_FISHEYE, _abi.MODEL_FOV, _abi.MODEL_DIVISION_UNDISTORTION
    assert bits(L.oracle_constant_intrinsics_mask(F, _abi.INTR_NONE)) == list(range(9))
    assert bits(L.oracle_constant_intrinsics_mask(V, _abi.INTR_NONE)) == list(range(5))
    assert bits(L.oracle_constant_intrinsics_mask(D, _abi.INTR_NONE)) == list(range(5))
    for flag, ff, fv in ((_abi.INTR_FOCAL_LENGTH, [0], [0]), (_abi.INTR_ASPECT_RATIO, [1], [1]), (_abi.INTR_SKEW, [2], []),
                         (_abi.INTR_PRINCIPAL_POINTS, [3, 4], [2, 3]), (_abi.INTR_RADIAL_DISTORTION, [5, 6, 7, 8], [4]),
                         (_abi.INTR_TANGENTIAL_DISTORTION, [], [])):
        assert bits(L.oracle_constant_intrinsics_mask(F, flag)) == [j for j in range(9) if j not in ff]
        assert bits(L.oracle_constant_intrinsics_mask(V, flag)) == [j for j in range(5) if j not in fv]
        assert bits(L.oracle_constant_intrinsics_mask(D, flag)) == [j for j in range(5) if j not in fv]
    for model in (F, V, D):
        assert L.oracle_constant_intrinsics_mask(model, _abi.INTR_ALL) == 0
        for m in range(64):
            assert _abi.constant_intrinsics_mask(model, m) == L.oracle_constant_intrinsics_mask(model, m)


def test_constant_subset_masks(oracle):
    L = oracle.lib()
    P, R = _abi.MODEL_PINHOLE, _abi.MODEL_PINHOLE_RADIAL_TANGENTIAL
    bits = lambda m: [j for j in range(10) if (m >> j) & 1]
    # NONE -> every parameter constant
    assert bits(L.oracle_constant_intrinsics_mask(P, _abi.INTR_NONE)) == list(range(7))
    assert bits(L.oracle_constant_intrinsics_mask(R, _abi.INTR_NONE)) == list(range(10))
    # single flags: size = K - (#freed), freed indices absent
    for flag, freed_p, freed_r in ((_abi.INTR_FOCAL_LENGTH, [0], [0]), (_abi.INTR_ASPECT_RATIO, [1], [1]),
                                   (_abi.INTR_SKEW, [2], [2]), (_abi.INTR_PRINCIPAL_POINTS, [3, 4], [3, 4]),
                                   (_abi.INTR_RADIAL_DISTORTION, [5, 6], [5, 6, 7]),
                                   (_abi.INTR_TANGENTIAL_DISTORTION, [], [8, 9])):
        mp = bits(L.oracle_constant_intrinsics_mask(P, flag)); mr = bits(L.oracle_constant_intrinsics_mask(R, flag))
        assert mp == [j for j in range(7) if j not in freed_p]
        assert mr == [j for j in range(10) if j not in freed_r]
    assert L.oracle_constant_intrinsics_mask(P, _abi.INTR_ALL) == 0
    # host mirror used by the adapter/python wrapper agrees
    for model in (P, R):
        for m in range(64):
            assert _abi.constant_intrinsics_mask(model, m) == L.oracle_constant_intrinsics_mask(model, m)


def test_project_point_round_trip(oracle):
    rng = np.random.default_rng(157)
    intr = np.array([800.0, 1.0, 0.0, 500.0, 500.0, 0.01, 0.001, 0, 0, 0])
    from theiasfm_b200.synthetic import rotation_from_angle_axis
    for _ in range(100):
        ext = np.concatenate([rng.uniform(-1, 1, 3), 0.2 * rng.uniform(-1, 1, 3)])
        pix = rng.uniform(0, 1000, 2)
        ray_cam = oracle.pixel_to_camera(_abi.MODEL_PINHOLE, intr, pix)
        R = rotation_from_angle_axis(ext[3:6])[0]
        depth = rng.uniform(2, 25)
        X = ext[:3] + depth * (R.T @ ray_cam)
        out, d = oracle.project_point(_abi.MODEL_PINHOLE, ext, intr, np.append(X, 1.0))
        assert np.linalg.norm(out - pix) < 1e-5 and abs(d - depth) < 1e-9
        # homogeneous scale invariance of the pixel, depth = q_z / h
        out2, d2 = oracle.project_point(_abi.MODEL_PINHOLE, ext, intr, np.append(X, 1.0) * 2.5)
        assert np.linalg.norm(out2 - out) < 1e-9 and abs(d2 - d) < 1e-9
