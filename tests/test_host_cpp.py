"""The C++ host layer: `MPCPlanner::Solver` over the C ABI (mpc_planner_b200/host), checked with
the assertions of the reference's own mpc_planner_solver/test/test_solver.cpp:52-134."""
import os
import struct
import subprocess

import numpy as np
import pytest

from mpc_planner_b200 import problem as pb
from mpc_planner_b200 import scenes
from mpc_planner_b200.host import build_host


@pytest.fixture(scope="module")
def exe():
    return build_host.build_test("cfg2")


def test_reference_solver_semantics_cpu(exe):
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


def test_generated_header_mirrors_reference_glue(exe):
    text = open(os.path.join(os.path.dirname(exe), "mpc_planner_generated.h")).read()
    for needle in ("#define MPC_PLANNER_CUDA_SOLVER", "#define SOLVER_N 20", "#define SOLVER_NP 135", "setSolverParameterEllipsoidObstX(",
                   "setSolverParameterLinConstraintA1(", "setSolverParameterSplineStart(", "setSolverParameterEgoDiscRadius(",
                   '#define WEIGHT_PARAMS {"acceleration", "angular_velocity", "velocity", "reference_velocity"}'):
        assert needle in text, needle


@pytest.mark.gpu
def test_cpp_solver_matches_python_binding(exe, tmp_path):
    from mpc_planner_b200 import binding

    spec = pb.config_tmpc()
    P = 8
    b = scenes.make_scene_batch(spec, 1, P, seed=9)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("i", P))
        for i in range(P):
            f.write(b["xinit"][i].tobytes())
            f.write(b["x0"][i].tobytes())
            f.write(b["params"][i].tobytes())
    r = subprocess.run([exe, "gpu", str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    solver = binding.BatchSolver(spec, P)
    xt, ut, info, best = solver.solve(b["xinit"], b["x0"], b["params"])
    raw = open(fout, "rb").read()
    off = 0
    (cbest,) = struct.unpack_from("i", raw, off)
    off += 4
    assert cbest == best
    nxt, nut = (spec.N + 1) * spec.nx, spec.N * spec.nu
    for i in range(P):
        (code,) = struct.unpack_from("i", raw, off)
        off += 4
        (pobj,) = struct.unpack_from("d", raw, off)
        off += 8
        cx = np.frombuffer(raw, dtype=np.float64, count=nxt, offset=off)
        off += 8 * nxt
        cu = np.frombuffer(raw, dtype=np.float64, count=nut, offset=off)
        off += 8 * nut
        assert code == info[i]["exit_code"]
        assert pobj == info[i]["pobj"]  # same kernel, same inputs: bit-identical
        np.testing.assert_array_equal(cx, xt[i].ravel())
        np.testing.assert_array_equal(cu, ut[i].ravel())
