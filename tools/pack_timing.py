"""Dump a synthetic config for tools/pack_timing.cc:  python tools/pack_timing.py c3_10kcam /tmp/c3.bin [shuffle]"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from theiasfm_b200 import synthetic
if sys.argv[1].startswith("custom:"):   # custom:n_cam,n_pt,obs_per_pt,shared_intrinsics(0/1),seed
    nc, npt, L, shared, seed = (int(v) for v in sys.argv[1][7:].split(","))
    p = synthetic.make_scene(n_cam=nc, n_pt=npt, obs_per_pt=L, shared_intrinsics=bool(shared), seed=seed)
    p.pt_const[::7] = 1; p.ext_const[::5] = 3; p.ext_const[1::5] = 1
else:
    p = synthetic.make_config(sys.argv[1])
if len(sys.argv) > 3:  # observation order of a per-view flattening (what the adapter produces): sorted by camera
    o = np.argsort(p.obs_cam, kind="stable") if sys.argv[3] == "byview" else np.random.default_rng(1).permutation(p.n_obs)
    p.obs_cam, p.obs_pt, p.obs_xy = p.obs_cam[o].copy(), p.obs_pt[o].copy(), p.obs_xy[o].copy()
with open(sys.argv[2], "wb") as f:
    np.array([p.n_cam, p.n_group, p.n_pt, p.n_obs], np.int64).tofile(f)
    for a in (p.ext, p.ext_const, p.cam_group, p.group_model, p.intr, p.group_const_mask, p.pt, p.pt_const, p.obs_cam, p.obs_pt, p.obs_xy):
        np.ascontiguousarray(a).tofile(f)
print(p.n_cam, p.n_pt, p.n_obs)
