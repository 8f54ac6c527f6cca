"""Secondary path: the CPU restatement of BruteForceFeatureMatcher pinned against the reference's
own tests -- brute_force_feature_matcher_test.cc:54-181 (NoOptions, RatioTest, SymmetricMatches),
feature_matcher_utils_test.cc:44-54 (IntersectMatches known answer), distance_test.cc:53-77 (L2) -- and against a
numpy restatement on random unit descriptors.  The CUDA matcher itself is round-2 work."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(HERE), "oracle")


class Match(C.Structure):
    _fields_ = [("feature1_ind", C.c_int32), ("feature2_ind", C.c_int32), ("distance", C.c_float)]


class Options(C.Structure):
    _fields_ = [("keep_only_symmetric_matches", C.c_int32), ("use_lowes_ratio", C.c_int32), ("lowes_ratio", C.c_float),
                ("min_num_feature_matches", C.c_int32)]


@pytest.fixture(scope="module")
def M():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "libmatcher_oracle.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ORACLE_DIR, "libmatcher_oracle.so"))
    fp = C.POINTER(C.c_float)
    L.matcher_l2.restype = C.c_float
    L.matcher_l2.argtypes = [fp, fp, C.c_int]
    L.matcher_options_init.argtypes = [C.POINTER(Options)]
    L.matcher_match_image_pair.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(Match), C.POINTER(C.c_int)]
    L.matcher_intersect.argtypes = [C.POINTER(Match), C.c_int, C.POINTER(Match), C.c_int, C.c_int]
    return L


def _match(M, d1, d2, **kw):
    d1 = np.ascontiguousarray(d1, np.float32); d2 = np.ascontiguousarray(d2, np.float32)
    o = Options(); M.matcher_options_init(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    out = (Match * max(len(d1), 1))(); n = C.c_int()
    fp = C.POINTER(C.c_float)
    ok = M.matcher_match_image_pair(d1.ctypes.data_as(fp), len(d1), d2.ctypes.data_as(fp), len(d2), d1.shape[1], C.byref(o), out, C.byref(n))
    return bool(ok), [(out[i].feature1_ind, out[i].feature2_ind, out[i].distance) for i in range(n.value)]


def _unit(v):
    v = np.asarray(v, np.float32)
    return v / np.linalg.norm(v)


def test_defaults_mirror_feature_matcher_options(M):
    o = Options(); M.matcher_options_init(C.byref(o))
    assert (o.keep_only_symmetric_matches, o.use_lowes_ratio, o.min_num_feature_matches) == (1, 1, 30) and abs(o.lowes_ratio - 0.8) < 1e-7


def test_reference_no_options_case(M):
    d = np.stack([_unit(np.ones(10))] * 10)
    ok, m = _match(M, d, d, min_num_feature_matches=0, keep_only_symmetric_matches=0, use_lowes_ratio=0)
    assert ok and len(m) > 0  # EXPECT_GT(database.NumMatches(), 0)


def test_reference_ratio_test_case(M):
    d1 = _unit(np.ones(10))[None]
    a = np.ones(10, np.float32); a[0] = 0.9
    b = np.ones(10, np.float32); b[0] = 0.89
    ok, m = _match(M, d1, np.stack([_unit(a), _unit(b)]), min_num_feature_matches=0, keep_only_symmetric_matches=0, use_lowes_ratio=1)
    # the reference only asserts NumMatches() > 0 at the image-pair level (the pair is stored even with 0 feature matches);
    # the two candidates are nearly equidistant, so the squared-ratio test rejects the feature match
    assert ok and m == []


def test_reference_symmetric_case(M):
    d1 = np.stack([_unit(np.ones(10)), np.eye(10, dtype=np.float32)[0]])
    a = np.ones(10, np.float32); a[0] = 0
    b = np.ones(10, np.float32); b[1] = 0; b[2] = 0
    ok, m = _match(M, d1, np.stack([_unit(a), _unit(b)]), min_num_feature_matches=0, keep_only_symmetric_matches=1, use_lowes_ratio=0)
    assert ok and len(m) == 1  # "the symmetric matching produces only 1 match"


def test_reference_intersect_known_answer(M):
    fwd = (Match * 2)(Match(0, 1, 0.8), Match(1, 2, 1.0))
    back = (Match * 2)(Match(1, 0, 0.8), Match(2, 3, 1.0))
    n = M.matcher_intersect(back, 2, fwd, 2, 4)
    assert n == 1 and (fwd[0].feature1_ind, fwd[0].feature2_ind) == (0, 1)


def test_reference_l2_cases(M):
    rng = np.random.default_rng(62)
    fp = C.POINTER(C.c_float)
    for _ in range(100):
        a = rng.normal(size=128).astype(np.float32); b = rng.normal(size=128).astype(np.float32)
        a /= np.linalg.norm(a); b /= np.linalg.norm(b)
        assert M.matcher_l2(a.ctypes.data_as(fp), a.ctypes.data_as(fp), 128) == 0.0
        d = M.matcher_l2(a.ctypes.data_as(fp), b.ctypes.data_as(fp), 128)
        assert abs(d - float(((a - b) ** 2).sum())) < 1e-5


def test_against_numpy_restatement(M):
    rng = np.random.default_rng(5)
    d1 = rng.normal(size=(300, 128)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 = np.concatenate([d1[:200] + 0.05 * rng.normal(size=(200, 128)).astype(np.float32), rng.normal(size=(150, 128)).astype(np.float32)])
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    D = ((d1[:, None, :].astype(np.float64) - d2[None, :, :]) ** 2).sum(-1)
    def one_way(D):
        idx = np.argsort(D, axis=1)[:, :2]
        best, second = D[np.arange(len(D)), idx[:, 0]], D[np.arange(len(D)), idx[:, 1]]
        keep = best < (np.float64(np.float32(0.8)) ** 2) * second
        return {int(i): int(idx[i, 0]) for i in np.nonzero(keep)[0]}
    f, b = one_way(D), one_way(D.T)
    expect = sorted((i, j) for i, j in f.items() if b.get(j) == i)
    ok, m = _match(M, d1, d2)
    assert ok and sorted((i, j) for i, j, _ in m) == expect and len(expect) >= 150
    # early exit: too few forward matches
    ok, m = _match(M, d1[:10], d2, min_num_feature_matches=30)
    assert not ok
