"""The solver generator for the CUDA backend (mpc_planner_b200/generate.py): a module stack that is NOT among the
shipped instantiations is emitted as a Config<> registry header, compiled into its own library (nvcc cross-compiles
here) and -- through the host emulation of the same generated instantiation -- solves like the oracle.
Reference: solver_generator/generate_solver.py:13-61, generate_acados_solver.py:190-216."""
import ctypes
import os

import numpy as np
import pytest

from mpc_planner_b200 import binding, generate, problem as pb, scenes
from oracle import oracle
from tests.emu import emu

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_generated")


def custom_spec():
    spec = pb.config_tmpc(N=16, n_obstacles=5)
    spec.name = "custom_n16_o5"
    return spec


def test_registry_header_names_the_stack():
    spec = custom_spec()
    path = generate.generate_cuda_registry(spec, OUT)
    text = open(path).read()
    assert "Config<MPCGPU_MODEL_UNICYCLE, MPCGPU_OBJ_CONTOURING, 16, 5, MPCGPU_CON_LINEAR, 5, MPCGPU_CON_ELLIPSOID, 5>" in text
    assert "#define MPCGPU_CONFIG_LIST(X) X(cfg_generated_custom_n16_o5)" in text


@pytest.mark.parametrize("bad", [dict(N=40), dict(N=1)])
def test_generator_rejects_shapes_outside_the_kernel_templates(bad):
    spec = pb.config_tmpc(**bad)
    with pytest.raises(ValueError):
        generate.generate_cuda_registry(spec, OUT)


def test_generated_instantiation_solves_like_the_oracle_in_emulation():
    spec = custom_spec()
    reg = generate.generate_cuda_registry(spec, OUT)
    so = emu.build_generated(reg, os.path.join(OUT, "libemu_custom.so"))
    b = scenes.make_scene_batch(spec, 2, 3, seed=5)
    xt, ut, info = emu.solve_batch(spec, b["xinit"], b["x0"], b["params"], library=so)
    oxt, out, oinfo = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    assert [i["exit_code"] for i in info] == [i["exit_code"] for i in oinfo]
    assert [i["qp_iter"] for i in info] == [i["qp_iter"] for i in oinfo]
    ok = np.array([i["exit_code"] == 1 for i in info])
    assert ok.any()
    np.testing.assert_allclose(xt[ok], oxt[ok], rtol=1e-6, atol=1e-6)


def test_generated_cuda_library_builds_and_serves_only_its_stack():
    spec = custom_spec()
    so = generate.build_solver_library(spec, OUT)
    lib = binding.load_library(so)
    for sym in binding.EXPORTED_SYMBOLS:
        assert hasattr(lib, sym)
    assert lib.mpcgpu_problem_supported(ctypes.byref(spec.desc())) == 1
    assert lib.mpcgpu_problem_supported(ctypes.byref(pb.CONFIGS["cfg2"]().desc())) == 0
    # the shipped library does not carry this stack
    assert binding.load_library().mpcgpu_problem_supported(ctypes.byref(spec.desc())) == 0
    assert os.path.exists(os.path.join(OUT, "mpc_planner_generated.h"))


@pytest.mark.gpu
def test_generated_cuda_library_matches_oracle_on_gpu():
    spec = custom_spec()
    so = os.path.join(OUT, "libmpcgpu_custom_n16_o5.so")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    deps = [os.path.join(root, "include", "mpcgpu.h")] + [
        os.path.join(root, "mpc_planner_b200", "csrc", f) for f in os.listdir(os.path.join(root, "mpc_planner_b200", "csrc"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        so = generate.build_solver_library(spec, OUT)  # nvcc is on the GPU box too
    b = scenes.make_scene_batch(spec, 4, 4, seed=8)
    solver = binding.BatchSolver(spec, 16, library=so)
    solver.set_group(4)
    xt, ut, info, best = solver.solve(b["xinit"], b["x0"], b["params"])
    oxt, out, oinfo = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    assert [i["exit_code"] for i in info] == [i["exit_code"] for i in oinfo]
    ok = np.array([i["exit_code"] == 1 for i in info])
    assert ok.any()
    np.testing.assert_allclose(xt[ok], oxt[ok], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ut[ok], out[ok], rtol=1e-6, atol=1e-6)
    solver.close()


def _build_if_stale(spec, out_dir):
    so = os.path.join(out_dir, "libmpcgpu_%s.so" % "".join(c if c.isalnum() else "_" for c in spec.name))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    deps = [os.path.join(root, "include", "mpcgpu.h"), spec.header()] + [
        os.path.join(root, "mpc_planner_b200", "csrc", f) for f in os.listdir(os.path.join(root, "mpc_planner_b200", "csrc"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        so = generate.build_solver_library(spec, out_dir)
    return so


def test_emitted_stage_functions_compile_into_per_problem_cuda_libraries():
    """nvcc cross-compiles the SQP kernel against the stage functions EMITTED by codegen.py (sm_100a) for both generated
    stacks; the libraries serve exactly their stack."""
    for name in ("tmpc", "lmpcc_goal"):
        spec = pb.GeneratedSpec.load(name)
        so = _build_if_stale(spec, os.path.join(OUT, spec.name))
        lib = binding.load_library(so)
        assert lib.mpcgpu_problem_supported(ctypes.byref(spec.desc())) == 1
        assert lib.mpcgpu_problem_supported(ctypes.byref(pb.CONFIGS["cfg2"]().desc())) == 0
        text = open(os.path.join(OUT, spec.name, "mpcgpu_generated_registry.cuh")).read()
        assert "MPCGPU_OBJ_GENERATED" in text and "MPCGPU_CON_GENERATED, %d" % spec.nh in text


@pytest.mark.gpu
def test_kernel_on_emitted_tmpc_code_matches_oracle_on_gpu():
    gspec = pb.GeneratedSpec.load("tmpc")
    cfg2 = pb.config_tmpc()
    so = _build_if_stale(gspec, os.path.join(OUT, gspec.name))
    b = scenes.make_scene_batch(cfg2, 4, 4, seed=8)
    solver = binding.BatchSolver(gspec, 16, library=so)
    xt, ut, info, best = solver.solve(b["xinit"], b["x0"], b["params"])
    oxt, out, oinfo = oracle.solve_batch(cfg2, b["xinit"], b["x0"], b["params"])
    assert [i["exit_code"] for i in info] == [i["exit_code"] for i in oinfo]
    ok = np.array([i["exit_code"] == 1 for i in info])
    assert ok.sum() >= 4
    np.testing.assert_allclose(xt[ok], oxt[ok], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ut[ok], out[ok], rtol=1e-6, atol=1e-6)
    solver.close()


@pytest.mark.gpu
def test_kernel_on_the_emitted_goal_stack_matches_oracle_on_gpu():
    from tests.test_codegen import goal_batch

    gspec = pb.GeneratedSpec.load("lmpcc_goal", oracle_objective=pb.OBJ_GOAL)
    so = _build_if_stale(gspec, os.path.join(OUT, gspec.name))
    xinit, x0, params = goal_batch(gspec, 8, seed=3)
    solver = binding.BatchSolver(gspec, 8, library=so)
    xt, ut, info, best = solver.solve(xinit, x0, params)

    class OracleView:
        N, nx, nu, nz, npar, n_sqp = gspec.N, gspec.nx, gspec.nu, gspec.nz, gspec.npar, gspec.n_sqp

        @staticmethod
        def desc():
            return gspec.desc(for_oracle=True)

    oxt, out, oinfo = oracle.solve_batch(OracleView, xinit, x0, params)
    assert [i["exit_code"] for i in info] == [i["exit_code"] for i in oinfo]
    ok = np.array([i["exit_code"] == 1 for i in info])
    assert ok.sum() >= 4
    np.testing.assert_allclose(xt[ok], oxt[ok], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ut[ok], out[ok], rtol=1e-6, atol=1e-6)
    solver.close()
