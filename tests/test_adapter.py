"""The C++ adapter (adapter/bundle_adjuster_b200.{h,cc}: drop-in for Theia's BundleAdjuster / BundleAdjust*Reconstruction)
compiled against the theia_compat stand-in headers."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "adapter")


@pytest.fixture(scope="module")
def adapter_test_bin(oracle):
    subprocess.check_call(["make", "-C", ADAPTER], stdout=subprocess.DEVNULL)
    return os.path.join(ADAPTER, "adapter_test")


def test_problem_construction_matches_reference_rules(adapter_test_bin):
    out = subprocess.run([adapter_test_bin, "flatten"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "flatten ok" in out.stdout


def test_track_estimator_host_logic_and_loud_refusal_without_gpu(adapter_test_bin):
    """TrackEstimatorB200 (drop-in for estimate_track.h): bookkeeping of EstimateTracks (:142-158) and, in a container
    without a GPU, a loud refusal that leaves the reconstruction untouched (no CPU path)."""
    out = subprocess.run([adapter_test_bin, "tracks-nogpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "tracks-nogpu ok" in out.stdout


def test_view_only_and_track_only_problems_flatten_and_solve_in_the_oracle(adapter_test_bin):
    """The CPU half of the BundleAdjustViewB200 / BundleAdjustTrackB200 test (bundle_adjustment.cc:82-107): AddView-only and
    AddTrack-only flattenings have the reference's constness pattern and the oracle solves them with the exact solver type."""
    out = subprocess.run([adapter_test_bin, "micro-oracle", os.path.join(ROOT, "oracle", "libba_oracle.so")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "micro-oracle ok" in out.stdout


def test_two_view_flattening_and_oracle_solve(adapter_test_bin):
    """BundleAdjustTwoViewsB200 (bundle_adjust_two_views.cc:112-191): camera 1 fixed, focal-only intrinsics subsets, DENSE_SCHUR,
    200 iterations, Ceres-default tolerances -- the CPU half (flattening + oracle); the GPU half is tests/test_z_adapter_gpu.py."""
    out = subprocess.run([adapter_test_bin, "twoview-oracle", os.path.join(ROOT, "oracle", "libba_oracle.so")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "twoview-oracle ok" in out.stdout


@pytest.mark.parametrize("mode,marker", [("solve", "solve ok"), ("tracks", "tracks ok"), ("micro", "micro ok"), ("twoview", "twoview ok")])
def test_adapters_end_to_end_against_a_mock_engine(oracle, mode, marker):
    """The adapters' OWN logic end to end without a GPU: tests/adapter_test.cc's GPU modes linked against tests/mock_engine.c, a
    stand-in for the C-ABI backed by the oracle (test infrastructure only).  Checks flattening, scatter back into the
    Reconstruction, Theia's shallow intrinsics sharing, residency of the device problem, status mapping, default options."""
    subprocess.check_call(["make", "-C", ADAPTER, "mock"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "tests", "adapter_test_mock"), mode, os.path.join(ROOT, "oracle", "libba_oracle.so")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert marker in out.stdout
