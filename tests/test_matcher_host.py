"""Secondary path, host side (no GPU): the matcher library loads and exports its ABI, fails loudly without a device,
and its host post-processing (ratio test, early exits, IntersectMatches) reproduces the CPU oracle when fed exact
nearest-neighbour results."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from theiasfm_b200 import engine, matcher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_matcher_library_exports_declared_symbols():
    L = matcher.lib()
    hdr = open(os.path.join(ROOT, "include", "theia_matcher_b200.h")).read()
    declared = set(re.findall(r"\b(tbm_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(matcher.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name)
    o = matcher.default_options()
    assert (o.keep_only_symmetric_matches, o.use_lowes_ratio, o.min_num_feature_matches) == (1, 1, 30) and abs(o.lowes_ratio - 0.8) < 1e-7


@pytest.mark.skipif(engine.device_count() > 0, reason="only meaningful without a GPU")
def test_matcher_has_no_cpu_fallback():
    d = [np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)]
    rc, _, _ = matcher.match_all(d, [(0, 1)])
    assert rc == -5


def _nn2_float32(A, B):
    """left-to-right float32 accumulation without FMA: the arithmetic of L2::operator() term by term"""
    s = np.zeros((len(A), len(B)), np.float32)
    for k in range(A.shape[1]):
        d = (A[:, None, k] - B[None, :, k]).astype(np.float32)
        s = (s + (d * d).astype(np.float32)).astype(np.float32)
    order = np.argsort(s, axis=1, kind="stable")
    bj = order[:, 0].astype(np.int32)
    bd = s[np.arange(len(A)), bj]
    sd = s[np.arange(len(A)), order[:, 1]] if len(B) > 1 else np.zeros(len(A), np.float32)
    return bj, np.ascontiguousarray(bd), np.ascontiguousarray(sd)


def _oracle_match(d1, d2, **kw):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libmatcher_oracle.so"], stdout=subprocess.DEVNULL)
    M = C.CDLL(os.path.join(ROOT, "oracle", "libmatcher_oracle.so"))
    o = matcher.default_options(**kw)
    out = (matcher.tbm_match * max(len(d1), 1))(); n = C.c_int()
    fp = C.POINTER(C.c_float)
    ok = M.matcher_match_image_pair(d1.ctypes.data_as(fp), len(d1), d2.ctypes.data_as(fp), len(d2), d1.shape[1], C.byref(o), out, C.byref(n))
    return bool(ok), [(out[i].feature1_ind, out[i].feature2_ind, out[i].distance) for i in range(n.value)]


@pytest.mark.parametrize("kw", [dict(), dict(keep_only_symmetric_matches=0), dict(use_lowes_ratio=0, min_num_feature_matches=0),
                                dict(min_num_feature_matches=500)])
def test_postprocess_matches_oracle(kw):
    rng = np.random.default_rng(8)
    d1 = rng.normal(size=(160, 32)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 = np.concatenate([d1[:110] + 0.04 * rng.normal(size=(110, 32)).astype(np.float32), rng.normal(size=(70, 32)).astype(np.float32)])
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    d1, d2 = np.ascontiguousarray(d1), np.ascontiguousarray(d2.astype(np.float32))
    fj, fd, fs = _nn2_float32(d1, d2)
    rj, rd, rs = _nn2_float32(d2, d1)
    o = matcher.default_options(**kw)
    out = (matcher.tbm_match * len(d1))(); n = C.c_int32()
    ip, fp = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    ok = matcher.lib().tbm_debug_postprocess(fj.ctypes.data_as(ip), fd.ctypes.data_as(fp), fs.ctypes.data_as(fp), len(d1), 1,
                                             rj.ctypes.data_as(ip), rd.ctypes.data_as(fp), rs.ctypes.data_as(fp), len(d2), 1,
                                             C.byref(o), out, C.byref(n))
    got = [(out[i].feature1_ind, out[i].feature2_ind, out[i].distance) for i in range(n.value)]
    ok_o, exp = _oracle_match(d1, d2, **kw)
    assert bool(ok) == ok_o and got == exp
    if not kw:
        assert len(got) > 80


def test_top2_scan_and_merge_equal_sequential_scan_with_ties():
    """theiasfm_b200/csrc/tbm_top2.h (used by the CUDA kernel) compiled for the host: the kernel's 8-scanner order +
    merge gives exactly the sequential-scan answer, including heavy ties and tiny candidate counts."""
    so = os.path.join(ROOT, "tests", "_host_top2.so")
    src = os.path.join(ROOT, "tests", "host_top2.cc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++14", "-fPIC", "-shared", src, "-o", so])
    L = C.CDLL(so)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    rng = np.random.default_rng(0)
    for trial in range(3000):
        n = int(rng.integers(0, 140))
        levels = int(rng.choice([1, 2, 3, 5, 1000]))
        d = rng.integers(0, levels, n).astype(np.float32) * np.float32(0.25) if levels < 1000 else rng.random(n).astype(np.float32)
        d = np.ascontiguousarray(d if n else np.zeros(1, np.float32))
        outs = []
        for fn in (L.emulate_kernel, L.sequential):
            bj, h2 = C.c_int(), C.c_int(); bd, sd = C.c_float(), C.c_float()
            fn(d.ctypes.data_as(fp), n, C.byref(bj), C.byref(bd), C.byref(sd), C.byref(h2))
            outs.append((bj.value, bd.value, sd.value, h2.value))
        assert outs[0] == outs[1], (trial, n, d[:n].tolist(), outs)
        if n >= 2:  # and the sequential scan is the plain definition
            order = np.argsort(d[:n], kind="stable")
            assert outs[1] == (int(order[0]), float(d[order[0]]), float(d[order[1]]), 1)
