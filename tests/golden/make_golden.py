"""Generates tests/golden/reprojection_golden.npz.

An INDEPENDENT float64 restatement of ReprojectionError<CameraModel>::operator()
(reference: src/theia/sfm/camera/reprojection_error.h:51-95, pinhole_camera_model.h:181-257,
pinhole_radial_tangential_camera_model.h:190-291, ceres rotation.h AngleAxisRotatePoint) in
torch, differentiated with torch.func.jacfwd (forward-mode AD = the mathematical object
Ceres' Jets compute).  The reference itself cannot be imported (C++, needs Ceres/Eigen --
absent), so these vectors pin the oracle's jets and the CUDA analytic Jacobian against a
second, independently written autodiff.  Run:  python tests/golden/make_golden.py
"""
import os

import numpy as np
import torch

torch.set_default_dtype(torch.float64)
EPS = np.finfo(np.float64).eps


def rotate(w, a, small):
    if small:  # first-order branch: a + w x a
        return a + torch.linalg.cross(w, a)
    theta = torch.sqrt((w * w).sum())
    k = w / theta
    return a * torch.cos(theta) + torch.linalg.cross(k, a) * torch.sin(theta) + k * (k * a).sum() * (1.0 - torch.cos(theta))


def residual(ext, intr, pt, xy, model, small):
    a = pt[:3] - pt[3] * ext[:3]
    q = rotate(ext[3:6], a, small)
    u, v = q[0] / q[2], q[1] / q[2]
    r2 = u * u + v * v
    if model == 0:
        d = 1.0 + r2 * (intr[5] + intr[6] * r2)
        ud, vd = u * d, v * d
    else:
        rd = 1.0 + intr[5] * r2 + intr[6] * r2 * r2 + intr[7] * r2 * r2 * r2
        tx = intr[9] * (r2 + 2.0 * u * u) + 2.0 * intr[8] * u * v
        ty = intr[8] * (r2 + 2.0 * v * v) + 2.0 * intr[9] * u * v
        ud, vd = u * rd + tx, v * rd + ty
    px = intr[0] * ud + intr[2] * vd + intr[3]
    py = intr[0] * intr[1] * vd + intr[4]
    return torch.stack([px - xy[0], py - xy[1]])


def residual_ext(ext, intr, pt, xy, model, small):
    """FISHEYE (2) / FOV (3) / DIVISION_UNDISTORTION (4): fisheye_camera_model.h:160-187,224-270;
    fov_camera_model.h:157-181,212-258; division_undistortion_camera_model.h:171-202,256-286.  Branches are chosen on the
    VALUES (as the reference's templated code does on Jet values)."""
    a = pt[:3] - pt[3] * ext[:3]
    q = rotate(ext[3:6], a, small)
    if model == 2:
        r_sq = q[0] * q[0] + q[1] * q[1]
        if float(r_sq) < 1e-8:
            ud, vd = q[0], q[1]
        else:
            r = torch.sqrt(r_sq)
            theta = torch.atan2(r, torch.abs(q[2]))
            t2 = theta * theta
            theta_d = theta * (1.0 + intr[5] * t2 + intr[6] * t2 * t2 + intr[7] * t2 * t2 * t2 + intr[8] * t2 * t2 * t2 * t2)
            ud, vd = theta_d * q[0] / r, theta_d * q[1] / r
            if float(q[2]) < 0.0:
                ud, vd = -ud, -vd
        px = intr[0] * ud + intr[2] * vd + intr[3]
        py = intr[0] * intr[1] * vd + intr[4]
    elif model == 3:
        u, v = q[0] / q[2], q[1] / q[2]
        omega = intr[4]
        r_u_sq = u * u + v * v
        if float(omega) < 1e-3:
            r_d = (omega * omega * r_u_sq) / 3.0 - omega * omega / 12.0 + 1.0
        elif float(r_u_sq) < 1e-3:
            th = torch.tan(omega / 2.0)
            r_d = (-2.0 * th * (4.0 * r_u_sq * th * th - 3.0)) / (3.0 * omega)
        else:
            r_u = torch.sqrt(r_u_sq)
            r_d = torch.atan(2.0 * r_u * torch.tan(omega / 2.0)) / (r_u * omega)
        px = intr[0] * (r_d * u) + intr[2]
        py = intr[0] * intr[1] * (r_d * v) + intr[3]
    else:
        u, v = q[0] / q[2], q[1] / q[2]
        up0, up1 = intr[0] * u, intr[0] * intr[1] * v
        r_u_sq = up0 * up0 + up1 * up1
        denom = 2.0 * intr[4] * r_u_sq
        inner = 1.0 - 4.0 * intr[4] * r_u_sq
        if abs(float(denom)) < EPS or float(inner) < 0.0:
            d0, d1 = up0, up1
        else:
            scale = (1.0 - torch.sqrt(inner)) / denom
            d0, d1 = up0 * scale, up1 * scale
        px = d0 + intr[2]
        py = d1 + intr[3]
    return torch.stack([px - xy[0], py - xy[1]])


def main_ext():
    """tests/golden/reprojection_golden_ext.npz: the three other camera models, every branch of their DistortPoint."""
    rng = np.random.default_rng(777)
    cases = []

    def add(model, ext, intr, pt, xy, tag):
        cases.append((model, np.array(ext, float), np.array(intr, float), np.array(pt, float), np.array(xy, float), tag))

    def intr_of(model, rng):
        k = np.zeros(10)
        if model == 2:
            k[:5] = [rng.uniform(300, 1500), rng.uniform(0.9, 1.1), rng.uniform(-2, 2), rng.uniform(300, 700), rng.uniform(300, 700)]
            k[5:9] = [rng.uniform(-0.05, 0.05), rng.uniform(-0.01, 0.01), rng.uniform(-0.003, 0.003), rng.uniform(-0.001, 0.001)]
        elif model == 3:
            k[:4] = [rng.uniform(300, 1500), rng.uniform(0.9, 1.1), rng.uniform(300, 700), rng.uniform(300, 700)]
            k[4] = rng.choice([rng.uniform(0.01, 1.2), rng.uniform(0.0, 9e-4)])
        else:
            k[:4] = [rng.uniform(300, 1500), rng.uniform(0.9, 1.1), rng.uniform(300, 700), rng.uniform(300, 700)]
            k[4] = -10.0 ** rng.uniform(-9, -6.5)
        return k

    for model in (2, 3, 4):
        for i in range(40):
            C = rng.uniform(-2, 2, 3)
            w = rng.uniform(-1.5, 1.5, 3) * rng.choice([1.0, 0.1, 1e-3])
            X = rng.uniform(-1, 1, 3) + np.array([0, 0, rng.uniform(4, 20)])
            h = rng.choice([1.0, 1.0, 0.5, 2.5, -1.0])
            add(model, np.concatenate([C, w]), intr_of(model, rng), np.concatenate([X * h, [h]]), rng.uniform(0, 1000, 2), "random")
        k = intr_of(model, rng)
        add(model, [0.1, -0.2, 0.3, 0.0, 0.0, 0.0], k, [0.5, 0.2, 9.0, 1.0], [510.0, 480.0], "w_zero")
        add(model, [0.0, 0.0, 0.0, 0.02, 0.01, -0.03], k, [0.4, -0.3, -6.0, 1.0], [100.0, 900.0], "behind_camera")
        add(model, [0.0, 0.0, 0.0, 0.3, -0.2, 0.1], k, [6.0, 5.0, 7.0, 1.0], [2000.0, 1900.0], "wide_angle")
        add(model, [0.0, 0.0, 0.0, 0.0, 0.0, 0.0], k, [2e-5, -3e-5, 5.0, 1.0], [500.0, 500.0], "on_axis")
    # branch cases
    add(2, [0, 0, 0, 0, 0, 0], [800, 1, 0.5, 500, 500, 0.02, 0.003, 0.001, 0.0005, 0], [3e-5, 4e-5, 2.0, 1.0], [500, 500], "fisheye_r_sq_below_1e-8")
    add(2, [0, 0, 0, 0.1, 0.0, 0.0], [800, 1, 0.5, 500, 500, 0.02, 0.003, 0.001, 0.0005, 0], [4.0, 1.0, 0.05, 1.0], [900, 600], "fisheye_near_90_degrees")
    add(3, [0, 0, 0, 0, 0, 0], [800, 1, 500, 500, 5e-4, 0, 0, 0, 0, 0], [1.0, 0.5, 6.0, 1.0], [600, 550], "fov_small_omega")
    add(3, [0, 0, 0, 0, 0, 0], [800, 1, 500, 500, 0.9, 0, 0, 0, 0, 0], [0.05, 0.02, 6.0, 1.0], [505, 503], "fov_small_radius")
    add(3, [0, 0, 0, 0, 0, 0], [800, 1, 500, 500, 0.9, 0, 0, 0, 0, 0], [2.0, 1.5, 6.0, 1.0], [700, 650], "fov_regular")
    add(4, [0, 0, 0, 0, 0, 0], [800, 1, 500, 500, 0.0, 0, 0, 0, 0, 0], [1.0, 0.5, 6.0, 1.0], [600, 550], "division_k_zero")
    add(4, [0, 0, 0, 0, 0, 0], [800, 1, 500, 500, 2e-6, 0, 0, 0, 0, 0], [4.0, 3.0, 6.0, 1.0], [900, 800], "division_negative_sqrt_argument")
    add(4, [0, 0, 0, 0, 0, 0], [800, 1, 500, 500, -1e-6, 0, 0, 0, 0, 0], [2.0, 1.5, 6.0, 1.0], [700, 650], "division_regular")
    out = dict(model=[], ext=[], intr=[], pt=[], xy=[], r=[], J=[], tag=[])
    for model, ext, intr, pt, xy, tag in cases:
        small = float(ext[3:6] @ ext[3:6]) <= EPS
        args = tuple(torch.tensor(a) for a in (ext, intr, pt))
        xy_t = torch.tensor(xy)
        f = lambda e, k, x: residual_ext(e, k, x, xy_t, model, small)
        r = f(*args)
        Je, Jk, Jx = torch.func.jacfwd(f, argnums=(0, 1, 2))(*args)
        J = torch.cat([Je, Jk, Jx], dim=1)
        npar = {2: 9, 3: 5, 4: 5}[model]
        assert float(J[:, 6 + npar:16].abs().max()) == 0.0
        for k, v in zip(("model", "ext", "intr", "pt", "xy", "r", "J", "tag"), (model, ext, intr, pt, xy, r.numpy(), J.numpy(), tag)):
            out[k].append(v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reprojection_golden_ext.npz")
    np.savez_compressed(path, model=np.array(out["model"], np.int32), ext=np.array(out["ext"]), intr=np.array(out["intr"]),
                        pt=np.array(out["pt"]), xy=np.array(out["xy"]), r=np.array(out["r"]), J=np.array(out["J"]),
                        tag=np.array(out["tag"]))
    print("wrote", path, len(cases), "cases")


def main():
    rng = np.random.default_rng(4242)
    cases = []

    def add(model, ext, intr, pt, xy, tag):
        cases.append((model, np.array(ext, float), np.array(intr, float), np.array(pt, float), np.array(xy, float), tag))

    for model in (0, 1):
        for i in range(40):
            C = rng.uniform(-2, 2, 3)
            w = rng.uniform(-1.5, 1.5, 3) * rng.choice([1.0, 0.1, 1e-3])
            X = rng.uniform(-1, 1, 3) + np.array([0, 0, rng.uniform(4, 20)])
            h = rng.choice([1.0, 1.0, 0.5, 2.5, -1.0])
            intr = np.zeros(10)
            intr[:5] = [rng.uniform(300, 1500), rng.uniform(0.9, 1.1), rng.uniform(-2, 2), rng.uniform(300, 700), rng.uniform(300, 700)]
            intr[5:7] = [rng.uniform(-0.2, 0.2), rng.uniform(-0.05, 0.05)]
            if model == 1:
                intr[7:10] = [rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01)]
            add(model, np.concatenate([C, w]), intr, np.concatenate([X * h, [h]]), rng.uniform(0, 1000, 2), "random")
        base_intr = np.array([800.0, 1.0, 0.0, 500.0, 500.0, -0.05, 0.01, 0.001 * model, 1e-3 * model, -5e-4 * model])
        add(model, [0.1, -0.2, 0.3, 0.0, 0.0, 0.0], base_intr, [0.5, 0.2, 9.0, 1.0], [510.0, 480.0], "w_zero")
        add(model, [0.1, -0.2, 0.3, 1e-9, -2e-9, 5e-10], base_intr, [0.5, 0.2, 9.0, 1.0], [510.0, 480.0], "w_tiny")
        add(model, [0.1, -0.2, 0.3, 1e-7, -2e-7, 5e-8], base_intr, [0.5, 0.2, 9.0, 1.0], [510.0, 480.0], "w_small_rodrigues")
        add(model, [0.0, 0.0, 0.0, 0.02, 0.01, -0.03], base_intr, [0.4, -0.3, -6.0, 1.0], [100.0, 900.0], "behind_camera")
        add(model, [0.0, 0.0, 0.0, 0.3, -0.2, 0.1], base_intr, [6.0, 5.0, 7.0, 1.0], [2000.0, 1900.0], "wide_angle")
        add(model, [1.0, 2.0, 3.0, 3.0, 0.5, -0.4], base_intr, [1.0, 2.0, 15.0, 3.0], [400.0, 600.0], "large_rotation")
    out = dict(model=[], ext=[], intr=[], pt=[], xy=[], r=[], J=[], tag=[])
    for model, ext, intr, pt, xy, tag in cases:
        small = float(ext[3:6] @ ext[3:6]) <= EPS
        args = tuple(torch.tensor(a) for a in (ext, intr, pt))
        xy_t = torch.tensor(xy)
        f = lambda e, k, x: residual(e, k, x, xy_t, model, small)
        r = f(*args)
        Je, Jk, Jx = torch.func.jacfwd(f, argnums=(0, 1, 2))(*args)
        J = torch.cat([Je, Jk, Jx], dim=1)  # [2, 6+10+4]
        if model == 0:
            assert float(J[:, 13:16].abs().max()) == 0.0
        for k, v in zip(("model", "ext", "intr", "pt", "xy", "r", "J", "tag"), (model, ext, intr, pt, xy, r.numpy(), J.numpy(), tag)):
            out[k].append(v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reprojection_golden.npz")
    np.savez_compressed(path, model=np.array(out["model"], np.int32), ext=np.array(out["ext"]), intr=np.array(out["intr"]),
                        pt=np.array(out["pt"]), xy=np.array(out["xy"]), r=np.array(out["r"]), J=np.array(out["J"]),
                        tag=np.array(out["tag"]))
    print("wrote", path, len(cases), "cases")


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "ext":
        main_ext()       # python tests/golden/make_golden.py ext
    else:
        main()
