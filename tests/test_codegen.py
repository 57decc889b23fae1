"""Generator back end: CUDA stage functions emitted by mpc_planner_b200/codegen.py from the
reference's UNMODIFIED module scripts (sympy trace -> derivatives -> CSE -> straight-line code).  The committed headers under
mpc_planner_b200/generated_stacks/ are compiled for the host here:
  (a) the emitted functions of the T-MPC stack (BASELINE cfg 2) equal the hand-derived ones of csrc/stage_functions.cuh --
      value, gradient, Hessian, constraint values, Jacobian and weighted Hessian -- to 1e-12 relative;
  (b) a stack the hand-written templates cannot express (configuration_lmpcc: Goal objective,
      generate_jackalsimulator_solver.py:105-121) reproduces golden values computed by the reference's own Python, and its
      derivatives agree with central differences of those values;
  (c) with the reference tree present, re-running the generator reproduces the committed headers byte for byte."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from mpc_planner_b200 import problem as pb
from tests.emu import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "mpc_planner_b200", "generated_stacks")
_dp = ctypes.POINTER(ctypes.c_double)


def gen_lib(name):
    out = os.path.join(ROOT, "tests", "_generated", f"libgen_{name}.so")
    hdr = os.path.join(GEN, f"stage_{name}.cuh")
    src = os.path.join(ROOT, "tests", "emu", "emu_gen.cpp")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(hdr), os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "mpc_planner_b200", "csrc", "stage_functions.cuh"))):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
                               f'-DMPCGPU_GENERATED_STAGE="{hdr}"', "-o", out, src])
    return ctypes.CDLL(out)


def gen_eval(lib, z, p, w=None):
    nz, nh, npar = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.gen_dims(ctypes.byref(nz), ctypes.byref(nh), ctypes.byref(npar))
    nz, nh, npar = nz.value, nh.value, npar.value
    z = np.ascontiguousarray(z, float).reshape(-1, nz)
    M = z.shape[0]
    p = np.ascontiguousarray(p, float).reshape(M, npar)
    out = dict(cost=np.zeros(M), cost_grad=np.zeros((M, nz)), cost_hess=np.zeros((M, nz, nz)), h=np.zeros((M, nh)), h_jac3=np.zeros((M, nh, 3)),
               h_hess6=np.zeros((M, 6)))
    wp = None if w is None else np.ascontiguousarray(w, float).reshape(M, nh)
    rc = lib.gen_eval(M, z.ctypes.data_as(_dp), p.ctypes.data_as(_dp), None if wp is None else wp.ctypes.data_as(_dp),
                      *[out[k].ctypes.data_as(_dp) for k in ("cost", "cost_grad", "cost_hess", "h", "h_jac3", "h_hess6")])
    assert rc == 0, rc
    return out


def test_emitted_tmpc_functions_equal_the_hand_derived_ones(golden):
    lib = gen_lib("tmpc")
    spec = pb.config_tmpc()
    d = spec.desc()
    g = golden["cfg2"]
    z = np.array([s["z"] for s in g["samples"]])
    p = np.array([s["p"] for s in g["samples"]])
    rng = np.random.default_rng(5)
    w = rng.uniform(-1.0, 1.0, size=(len(z), spec.nh))
    gen = gen_eval(lib, z, p, w)
    hand = emu.eval_stage(d, z, p, lam=w)
    iz = {n: i for i, n in enumerate(spec.inputs + spec.states)}
    cols = [iz["x"], iz["y"], iz["psi"]]

    def close(a, b, what):
        scale = 1.0 + np.abs(b).max()
        assert np.abs(a - b).max() <= 1e-12 * scale, (what, np.abs(a - b).max(), scale)

    close(gen["cost"], hand["cost"], "cost")
    close(gen["cost_grad"], hand["cost_grad"], "cost gradient")
    close(gen["cost_hess"], hand["cost_hess"], "cost Hessian")
    close(gen["h"], hand["h"], "h")
    close(gen["h_jac3"], hand["h_jac"][:, :, cols], "Jacobian (x, y, psi)")
    hh = np.array([[hand["h_hess"][m][cols[a], cols[b]] for a in range(3) for b in range(a + 1)] for m in range(len(z))])
    close(gen["h_hess6"], hh, "weighted constraint Hessian")
    # and both equal the reference's own Python values
    np.testing.assert_allclose(gen["cost"], [s["cost"] for s in g["samples"]], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(gen["h"], [s["h"] for s in g["samples"]], rtol=1e-11, atol=1e-10)


def test_generated_goal_stack_reproduces_the_reference_python(golden):
    lib = gen_lib("lmpcc_goal")
    g = golden["lmpcc_goal"]
    spec = json.load(open(os.path.join(GEN, "stage_lmpcc_goal.json")))
    assert spec["parameter_map"] == g["parameter_map"] and spec["npar"] == g["npar"] == 55 and spec["nh"] == g["nh"] == 4
    assert spec["modules"] == ["MPCBaseModule", "GoalModule", "PathReferenceVelocity", "EllipsoidConstraints"]
    z = np.array([s["z"] for s in g["samples"]])
    p = np.array([s["p"] for s in g["samples"]])
    rng = np.random.default_rng(6)
    w = rng.uniform(-1.0, 1.0, size=(len(z), 4))
    out = gen_eval(lib, z, p, w)
    np.testing.assert_allclose(out["cost"], [s["cost"] for s in g["samples"]], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["h"], [s["h"] for s in g["samples"]], rtol=1e-11, atol=1e-10)
    # derivatives: central differences of the generated VALUES (which equal the reference's)
    nz, eps = z.shape[1], 1e-6
    names = spec["inputs"] + spec["states"]
    cols = [names.index("x"), names.index("y"), names.index("psi")]
    for i in range(nz):
        zp, zm = z.copy(), z.copy()
        zp[:, i] += eps
        zm[:, i] -= eps
        op, om = gen_eval(lib, zp, p, w), gen_eval(lib, zm, p, w)
        np.testing.assert_allclose((op["cost"] - om["cost"]) / (2 * eps), out["cost_grad"][:, i], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose((op["cost_grad"] - om["cost_grad"]) / (2 * eps), out["cost_hess"][:, i], rtol=1e-5, atol=1e-5)
        if i in cols:
            c = cols.index(i)
            np.testing.assert_allclose((op["h"] - om["h"]) / (2 * eps), out["h_jac3"][:, :, c], rtol=1e-5, atol=1e-5)
            # weighted Hessian row c: d/dz_i of sum_j w_j dh_j/d(x,y,psi)
            dj = np.einsum("mj,mjc->mc", w, (op["h_jac3"] - om["h_jac3"]) / (2 * eps))
            for b in range(3):
                a_, b_ = max(c, b), min(c, b)
                np.testing.assert_allclose(dj[:, b], out["h_hess6"][:, a_ * (a_ + 1) // 2 + b_], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name,npar,nh", [("tmpc_2discs", 136, 24), ("tmpc_dynvel", 155, 16)])
def test_generated_settings_variants_reproduce_the_reference_python(golden, name, npar, nh):
    """n_discs = 2 (one ellipsoid row per obstacle AND disc, ellipsoid_constraints.py:88-96) and
    contouring/dynamic_velocity_reference (velocity spline in the contouring cost, contouring.py:61-67, 80-81): settings the
    hand-written templates do not cover exist as generated stacks."""
    lib = gen_lib(name)
    g = golden[name]
    spec = json.load(open(os.path.join(GEN, f"stage_{name}.json")))
    assert spec["parameter_map"] == g["parameter_map"] and spec["npar"] == g["npar"] == npar and spec["nh"] == g["nh"] == nh
    z = np.array([s["z"] for s in g["samples"]])
    p = np.array([s["p"] for s in g["samples"]])
    w = np.random.default_rng(7).uniform(-1.0, 1.0, size=(len(z), nh))
    out = gen_eval(lib, z, p, w)
    np.testing.assert_allclose(out["cost"], [s["cost"] for s in g["samples"]], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(out["h"], [s["h"] for s in g["samples"]], rtol=1e-11, atol=1e-10)
    eps = 1e-6
    for i in range(z.shape[1]):
        zp, zm = z.copy(), z.copy()
        zp[:, i] += eps
        zm[:, i] -= eps
        op, om = gen_eval(lib, zp, p, w), gen_eval(lib, zm, p, w)
        scale = 1.0 + np.abs(out["cost_grad"][:, i])
        assert (np.abs((op["cost"] - om["cost"]) / (2 * eps) - out["cost_grad"][:, i]) / scale).max() < 1e-5
        scale = 1.0 + np.abs(out["cost_hess"][:, i]).max(axis=1, keepdims=True)
        assert (np.abs((op["cost_grad"] - om["cost_grad"]) / (2 * eps) - out["cost_hess"][:, i]) / scale).max() < 1e-4
    # the kernel instantiated on the emitted code solves a scene: every solve that succeeds is dynamically consistent and
    # satisfies the emitted constraints
    gspec = pb.GeneratedSpec.load(name)
    so = _generated_emulation(gspec)
    from mpc_planner_b200 import scenes

    cfg2 = pb.config_tmpc()
    b = scenes.make_scene_batch(cfg2, 2, 3, seed=9)
    pm2, pmg = cfg2.parameter_map(), gspec.parameter_map()
    par2 = b["params"].reshape(6, cfg2.N, cfg2.npar)
    par = np.zeros((6, gspec.N, gspec.npar))
    for nm, i in pmg.items():
        if nm in pm2:
            par[:, :, i] = par2[:, :, pm2[nm]]
    if name == "tmpc_2discs":
        par[:, :, pmg["ego_disc_0_offset"]] = 0.1
        par[:, :, pmg["ego_disc_1_offset"]] = -0.1
    else:
        for k in range(gspec.n_segments):
            par[:, :, pmg[f"spline_v{k}_d"]] = 2.0  # constant reference velocity 2 m/s along the path
    xt, ut, info = emu.solve_batch(gspec, b["xinit"], b["x0"], par.reshape(6, -1), library=so)
    ok = [i for i, inf in enumerate(info) if inf["exit_code"] == 1]
    assert len(ok) >= 3
    lh = np.array(spec["lh"])
    uh = np.array(spec["uh"])
    for i in ok:
        assert info[i]["res_eq"] < 1e-2
        zz = np.concatenate([ut[i], xt[i, :-1]], axis=1)[1:]
        hv = gen_eval(lib, zz, par[i, 1:])["h"]
        assert (hv >= lh - 1e-3).all() and (hv <= uh + 1e-3).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree")
def test_regenerating_from_the_reference_reproduces_the_committed_headers(tmp_path):
    code = ("import sys; sys.path.insert(0, %r); from mpc_planner_b200 import codegen; "
            "[print(codegen.generate(n, %r)) for n in ('tmpc', 'lmpcc_goal')]") % (ROOT, str(tmp_path))  # (the settings variants take a minute more)
    subprocess.check_call([os.environ.get("PYTHON", "python"), "-c", code], timeout=1800)
    for n in ("tmpc", "lmpcc_goal"):
        for ext in ("cuh", "json"):
            assert open(tmp_path / f"stage_{n}.{ext}").read() == open(os.path.join(GEN, f"stage_{n}.{ext}")).read(), (n, ext)


def _generated_emulation(spec):
    from mpc_planner_b200 import generate

    out_dir = os.path.join(ROOT, "tests", "_generated", spec.name)
    reg = generate.generate_cuda_registry(spec, out_dir)
    lib = os.path.join(out_dir, "libemu.so")
    srcs = [reg, spec.header(), os.path.join(ROOT, "tests", "emu", "emu.cpp")] + [os.path.join(ROOT, "mpc_planner_b200", "csrc", f) for f in ("sqp_core.cuh", "stage_functions.cuh", "registry.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(s) for s in srcs):
        emu.build_generated(reg, lib, stage_header=spec.header())
    return lib


def test_sqp_kernel_on_emitted_tmpc_code_solves_like_the_oracle():
    """The SQP kernel source instantiated on the EMITTED stage functions (Config<..., MPCGPU_OBJ_GENERATED, ...,
    MPCGPU_CON_GENERATED, nh>; host emulation here, the same per-problem library on the GPU in tests/test_generator.py) solves
    the T-MPC stack like the oracle does with its hand-restated functions: same parameter vector, same NLP."""
    from mpc_planner_b200 import scenes
    from oracle import oracle
    from tests.util import same_status

    gspec = pb.GeneratedSpec.load("tmpc")
    cfg2 = pb.config_tmpc()
    assert gspec.parameter_map() == cfg2.parameter_map() and gspec.nh == cfg2.nh and gspec.npar == cfg2.npar
    lib = _generated_emulation(gspec)
    b = scenes.make_scene_batch(cfg2, 2, 3, seed=9)
    xt, ut, info = emu.solve_batch(gspec, b["xinit"], b["x0"], b["params"], library=lib)
    oxt, out, oinfo = oracle.solve_batch(cfg2, b["xinit"], b["x0"], b["params"])
    n_ok = 0
    for i in range(len(info)):
        assert same_status(info[i], oinfo[i]), (i, info[i], oinfo[i])
        if oinfo[i]["exit_code"] == 1:
            n_ok += 1
            np.testing.assert_allclose(xt[i], oxt[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ut[i], out[i], rtol=1e-6, atol=1e-6)
            assert abs(info[i]["pobj"] - oinfo[i]["pobj"]) <= 1e-6 * (1 + abs(oinfo[i]["pobj"]))
    assert n_ok >= 3


def goal_batch(gspec, n, seed):
    """Scenes for configuration_lmpcc: the T-MPC scene generator's robot state, warm start and obstacles (cfg 1 shapes: 4
    obstacles), goal 12-18 m ahead on the reference path, Goal / PathReferenceVelocity parameters by name."""
    from mpc_planner_b200 import scenes

    cfg1 = pb.config_basic()
    b = scenes.make_scene_batch(cfg1, n, 1, seed=seed)
    pm1, pmg = cfg1.parameter_map(), gspec.parameter_map()
    N = gspec.N
    par1 = b["params"].reshape(n, N, cfg1.npar)
    par = np.zeros((n, N, gspec.npar))
    rng = np.random.default_rng(seed)
    for name, i in pmg.items():
        if name in pm1:
            par[:, :, i] = par1[:, :, pm1[name]]
    x0 = b["x0"].reshape(n, N + 1, gspec.nz)
    for s in range(n):
        ahead = x0[s, N, gspec.nu:gspec.nu + 2] + rng.uniform(2.0, 6.0) * (x0[s, N, gspec.nu:gspec.nu + 2] - x0[s, N - 1, gspec.nu:gspec.nu + 2]) / 0.3
        par[s, :, pmg["goal_weight"]] = 1.5
        par[s, :, pmg["goal_x"]], par[s, :, pmg["goal_y"]] = ahead
    return b["xinit"], b["x0"], par.reshape(n, -1)


def test_sqp_kernel_on_the_emitted_goal_stack_solves_like_the_oracle():
    """configuration_lmpcc (Goal objective): no hand-written template exists for it; the kernel runs on emitted code only.
    Checker: the oracle with the Goal objective restated by hand (oracle/mpc_oracle.cpp, goal_module.py:26-41)."""
    from oracle import oracle
    from tests.util import same_status

    gspec = pb.GeneratedSpec.load("lmpcc_goal", oracle_objective=pb.OBJ_GOAL)
    lib = _generated_emulation(gspec)
    xinit, x0, params = goal_batch(gspec, 4, seed=3)
    xt, ut, info = emu.solve_batch(gspec, xinit, x0, params, library=lib)

    class OracleView:  # the same problem seen through the oracle's descriptor
        N, nx, nu, nz, npar, n_sqp = gspec.N, gspec.nx, gspec.nu, gspec.nz, gspec.npar, gspec.n_sqp

        @staticmethod
        def desc():
            return gspec.desc(for_oracle=True)

    oxt, out, oinfo = oracle.solve_batch(OracleView, xinit, x0, params)
    n_ok = 0
    for i in range(len(info)):
        assert same_status(info[i], oinfo[i]), (i, info[i], oinfo[i])
        if oinfo[i]["exit_code"] == 1:
            n_ok += 1
            np.testing.assert_allclose(xt[i], oxt[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ut[i], out[i], rtol=1e-6, atol=1e-6)
    assert n_ok >= 2
    # the goal pulls: the final position is closer to the goal than the initial one
    pm = gspec.parameter_map()
    par = params.reshape(4, gspec.N, gspec.npar)
    for i in range(4):
        if oinfo[i]["exit_code"] == 1:
            goal = par[i, 0, [pm["goal_x"], pm["goal_y"]]]
            assert np.linalg.norm(xt[i, -1, :2] - goal) < np.linalg.norm(xt[i, 0, :2] - goal)


# ---- SecondOrderUnicycleModel (solver_model.py:170-190) as the Goal stack with an idle spline state --------------------------
def test_second_order_unicycle_is_the_contouring_model_with_an_idle_spline_state(golden):
    """The reference's OWN Python (fixture `second_order_unicycle`, tests/golden/make_golden.py): Goal-stack cost, constraints and
    the first four rows of f are the same numbers on SecondOrderUnicycleModel and on ContouringSecondOrderUnicycleModel at the
    point extended by ANY spline value; the fifth row is v.  The emitted Goal stack reproduces the 4-state model's numbers."""
    g = golden["second_order_unicycle"]
    assert g["states"] == ["x", "y", "psi", "v"] and g["inputs"] == ["a", "w"]
    lib = gen_lib("lmpcc_goal")
    for s in g["samples"]:
        assert s["cost4"] == s["cost5"] and s["h4"] == s["h5"] and s["f4"] == s["f5"][:4]
        assert s["f5"][4] == s["z5"][5]  # spline' = v
        out = gen_eval(lib, np.array(s["z5"]), np.array(s["p"]))
        np.testing.assert_allclose(out["cost"][0], s["cost4"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(out["h"][0], s["h4"], rtol=1e-11, atol=1e-10)
    spec = pb.second_order_unicycle_goal()
    lo, up = spec.model_bounds()
    assert lo[:6] == g["lower_bound"] and up[:6] == g["upper_bound"]  # the 4-state model's own bounds (:177-178)
    assert pb.GeneratedSpec.load("lmpcc_goal").model_bounds()[0][1] == -0.8  # the contouring model's table is untouched


def test_second_order_unicycle_goal_solve_ignores_the_spline_state():
    """The 5-state solve IS the 4-state problem's solution: started from spline = 0 and from spline = 7.3 (initial state and warm
    start shifted alike) the kernel (host emulation) returns the same (x, y, psi, v) and (a, w) trajectories to the solver's
    tolerance, and the oracle agrees with it."""
    from oracle import oracle
    from tests.util import same_status

    gspec = pb.second_order_unicycle_goal(oracle_objective=pb.OBJ_GOAL)
    lib = _generated_emulation(gspec)
    xinit, x0, params = goal_batch(gspec, 3, seed=11)
    N, nz, nu = gspec.N, gspec.nz, gspec.nu
    xinit2, x02 = xinit.copy(), x0.copy().reshape(3, N + 1, nz)
    xinit2[:, 4] += 7.3
    x02[:, :, nu + 4] += 7.3
    x02 = x02.reshape(3, -1)
    xt, ut, info = emu.solve_batch(gspec, xinit, x0, params, library=lib)
    xt2, ut2, info2 = emu.solve_batch(gspec, xinit2, x02, params, library=lib)

    class OracleView:
        N, nx, nu, nz, npar, n_sqp = gspec.N, gspec.nx, gspec.nu, gspec.nz, gspec.npar, gspec.n_sqp

        @staticmethod
        def desc():
            return gspec.desc(for_oracle=True)

    oxt, out, oinfo = oracle.solve_batch(OracleView, xinit, x0, params)
    n_ok = 0
    for i in range(3):
        assert same_status(info[i], oinfo[i]) and same_status(info[i], info2[i])
        if oinfo[i]["exit_code"] == 1:
            n_ok += 1
            # to the interior-point tolerance (qp_tol = 1e-5; observed <= 1.2e-5): the barrier of the spline state's never-active
            # bound [-1, 10000] is the only place its value enters
            np.testing.assert_allclose(xt[i][:, :4], xt2[i][:, :4], rtol=0, atol=1e-4)
            np.testing.assert_allclose(ut[i], ut2[i], rtol=0, atol=1e-4)
            np.testing.assert_allclose(xt2[i][:, 4] - xt[i][:, 4], 7.3, rtol=0, atol=1e-4)
            np.testing.assert_allclose(xt[i], oxt[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ut[i], out[i], rtol=1e-6, atol=1e-6)
    assert n_ok >= 2
