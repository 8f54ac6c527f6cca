"""bench.py and __graft_entry__.smoke() need a GPU; their Python logic (argument handling, the JSON line and every derived field,
the oracle comparison of smoke) is exercised here on the CPU by running them in a subprocess against tests/mock_engine_py.py, a
stand-in for the engine answered by the oracle.  Says nothing about the CUDA engine or about performance."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = """
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import mock_engine_py; mock_engine_py.install()
import torch
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None
""" % (os.path.join(ROOT, "tests"), ROOT)


def _run(code):
    return subprocess.run([sys.executable, "-c", PRELUDE + code], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_bench_line_has_every_contract_field():
    out = _run("""
import bench
bench.run_microbench = lambda d: {"fp64_fma_tflops": 37.0, "fp64_red_gops": 200.0, "gather48_grows": 50.0, "how": "mock"}
sys.argv = ["bench.py", "--workload", "c1_50cam", "--steps", "2", "--warmup", "1", "--no-cpu-baseline"]
bench.main()
""")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["unit"] == "obs/s" and line["dtype"] == "f64" and line["data"] == "synthetic" and "workload" in line["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["steps_run"] == 2 and line["value"] > 0 and line["obs_passes_per_s"] >= line["value"]


def test_smoke_logic():
    out = _run("import __graft_entry__ as g\ng.smoke()\n")
    assert out.returncode == 0, out.stderr[-2000:]
    assert "smoke ok" in out.stdout
