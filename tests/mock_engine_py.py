"""TEST INFRASTRUCTURE: `pytest -m gpu --mock-engine` runs the GPU test files on a machine WITHOUT a GPU against this stand-in for
theiasfm_b200.engine.Engine, which answers every call from the CPU oracle.  It proves nothing about the CUDA engine (the answers are
the oracle's own); it checks the TESTS: scene construction, thresholds derived from scene statistics, attribute names, array
shapes -- everything that otherwise only surfaces on the GPU box at round end.  Never imported by the product."""
import numpy as np

from oracle import oracle_py
from theiasfm_b200 import _abi, engine


class MockEngine:
    def __init__(self, device=0, rank=0, world_size=1, nccl_id=None, max_iterations_logged=2048):
        self.rank, self.world_size = rank, world_size
        self._p = self._w = self._opts = self._o = None

    def close(self):
        if self._o is not None:
            self._o.close()
            self._o = None

    def _unsupported(self, problem, options):
        if options.linear_solver_type == _abi.CGNR:
            return "CGNR"
        if engine.debug_pack(problem)["rc"] == _abi.ERR_UNSUPPORTED:   # the engine's 256-observation track limit (host pack)
            return "track too long"
        return None

    def solve(self, problem, options=None):
        options = options or engine.default_options()
        why = self._unsupported(problem, options)
        if why:
            class S:
                pass
            s = S(); s.rc = _abi.ERR_UNSUPPORTED; s.success = 0; s.message = why; s.termination_type = _abi.FAILURE
            return s
        self.upload(problem, options)
        s = self.minimize()
        self.download(problem)
        return s

    def upload(self, problem, options=None):
        self.close()
        self._p, self._opts = problem, options or engine.default_options()
        self._w = problem.copy()

    def _staged(self):
        if self._o is None:
            self._o = oracle_py.Oracle(self._w, self._opts)
        return self._o

    def minimize(self):
        self.close()
        s = oracle_py.solve(self._w, self._opts)
        s.num_kernel_launches = 1
        self._last = s
        return s

    def download(self, problem=None):
        problem = problem or self._p
        problem.ext[:], problem.intr[:], problem.pt[:] = self._w.ext, self._w.intr, self._w.pt

    def reset_parameters(self, problem):
        self.close()
        self._w.ext[:], self._w.intr[:], self._w.pt[:] = problem.ext, problem.intr, problem.pt

    def filter_tracks(self, max_err, min_angle):
        st, mean, removed = oracle_py.filter_tracks(self._w, max_err, min_angle)
        return st, mean, int((st == 1).sum()), int((st == 2).sum())

    def adjust_tracks(self, options):
        return oracle_py.adjust_tracks(self._w, options)

    def estimate_tracks(self, options, max_reprojection_error_pixels=5.0, min_triangulation_angle_degrees=3.0, bundle_adjustment=True):
        return oracle_py.estimate_tracks(self._w, options, max_reprojection_error_pixels, min_triangulation_angle_degrees, bundle_adjustment)

    def two_view_ba_batch(self, batch):
        return oracle_py.two_view_ba_batch(batch)

    def set_max_iterations(self, n):
        self._opts.max_num_iterations = int(n)

    def set_profiling(self, enable=True):
        pass

    def profile(self):
        s = getattr(self, "_last", None)
        n_it = s.num_iterations if s is not None else 0
        n_mv = sum(it["linear_solver_iterations"] + 2 for it in s.iterations) if s is not None else 0
        return dict(matvec_ms=0.0, matvec_launches=n_mv, linearize_ms=0.0, linearize_launches=n_it, slots=self._w.n_obs, observations=self._w.n_obs,
                    points=self._w.n_pt, doubles_per_obs=20)

    def profile_stages(self):
        p = self.profile()
        z = {"ms": 0.0, "launches": p["linearize_launches"]}
        out = {k: dict(z) for k in ("linearize", "precond_ext", "precond_intr", "rhs", "backsub", "candidate_cost", "prepare_fused")}
        out["matvec"] = {"ms": 0.0, "launches": p["matvec_launches"]}
        return out

    # staged hooks
    def linearize(self):
        return self._staged().linearize()

    def prepare_linear_system(self, radius):
        return self._staged().prepare_linear_system(radius)

    def schur_matvec(self, x_cam, x_intr):
        return self._staged().schur_matvec(x_cam, x_intr)

    def solve_linear_system(self):
        return self._staged().solve_linear_system()

    def evaluate_step(self):
        return self._staged().evaluate_step()

    def read(self, which):
        return self._staged().read(which)


def install(monkeypatch_target=engine):
    monkeypatch_target.Engine = MockEngine
    monkeypatch_target.device_count = lambda: 1
    monkeypatch_target.two_view_ba_batch_multi = lambda batch, n_devices=0: oracle_py.two_view_ba_batch(batch)
