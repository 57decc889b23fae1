"""BicycleModel2ndOrder on the device: BASELINE configs[0] as written -- bicycle model,
MPCC, 4 ellipsoidal obstacles, N = 20.  Device code: ERK4 x 3 of the (x, y, psi) subsystem with first- and second-order
sensitivities propagated as Taylor jets (csrc/stage_functions.cuh bicycle_step), dense Riccati sweeps for nu = 3
(csrc/sqp_core.cuh riccati_*_dense).  Checked here through the host emulation of the kernel source; on the GPU through the
C ABI (tests below marked gpu)."""
import ctypes

import numpy as np
import pytest

from mpc_planner_b200 import problem as pb, scenes
from oracle import oracle
from tests.emu import emu
from tests.util import same_status


def samples(spec, M, seed):
    rng = np.random.default_rng(seed)
    b = scenes.make_bicycle_batch(spec, M, seed=seed)
    z = b["x0"].reshape(M, spec.N + 1, spec.nz)[:, 3].copy()
    z[:, 0], z[:, 1], z[:, 2] = rng.uniform(-2, 2, M), rng.uniform(-1, 1, M), rng.uniform(0, 1, M)
    z[:, spec.nu + 4] = rng.uniform(-0.5, 0.5, M)  # steering angle
    p = b["params"].reshape(M, spec.N, spec.npar)[:, 3]
    return z, p, rng.uniform(0, 1, (M, spec.nh)), rng.uniform(-1, 1, (M, spec.nx))


def test_continuous_model_is_the_reference_one(golden):
    """scenes._bicycle_rk4's right-hand side (used to build warm starts) against the reference's own Python values."""
    for s in golden["bicycle"]["samples"]:
        z = np.array(s["z"])
        lr = 2.79 / 2.0
        a, w = z[0], z[1]
        x = z[3:]
        beta = np.arctan(0.5 * np.tan(x[4]))
        f = [x[3] * np.cos(x[2] + beta), x[3] * np.sin(x[2] + beta), x[3] / lr * np.sin(beta), a, w, x[3]]
        np.testing.assert_allclose(f, s["f"], rtol=1e-13, atol=1e-13)
    spec = pb.config_bicycle()
    assert list(spec.desc().lower_bound)[:9] == golden["bicycle"]["lower_bound"] and list(spec.desc().upper_bound)[:9] == golden["bicycle"]["upper_bound"]


def test_bicycle_stage_functions_match_oracle_ad():
    spec = pb.config_bicycle()
    d = spec.desc()
    z, p, lam, pi = samples(spec, 8, seed=2)
    e = emu.eval_stage(d, z, p, lam=lam, pi=pi)
    for m in range(len(z)):
        o = oracle.eval_stage(d, z[m], p[m], lam[m], pi[m])
        for k in ("cost", "cost_grad", "cost_hess", "h", "h_jac", "h_hess", "xnext", "AB", "dyn_hess"):
            a, b = np.asarray(e[k][m]), np.asarray(o[k])
            assert np.abs(a - b).max() <= 1e-9 * (1 + np.abs(b).max()), (m, k)
    # the discrete map is ERK4 x 3 of the reference's continuous model: compare with an independent numpy integration
    for m in range(len(z)):
        xn = scenes._bicycle_rk4(z[m, spec.nu:].copy(), (z[m, 0], z[m, 1]), spec.dt)
        np.testing.assert_allclose(e["xnext"][m], xn, rtol=1e-12, atol=1e-12)


def test_emulated_kernel_solves_the_bicycle_problem_like_the_oracle():
    spec = pb.config_bicycle()
    b = scenes.make_bicycle_batch(spec, 6, seed=4)
    xt, ut, info = emu.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    oxt, out, oinfo = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    n_ok = 0
    for i in range(6):
        assert same_status(info[i], oinfo[i]), (i, info[i], oinfo[i])
        if oinfo[i]["exit_code"] == 1:
            n_ok += 1
            np.testing.assert_allclose(xt[i], oxt[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ut[i], out[i], rtol=1e-6, atol=1e-6)
            assert abs(info[i]["pobj"] - oinfo[i]["pobj"]) <= 1e-6 * (1 + abs(oinfo[i]["pobj"]))
    assert n_ok >= 4


def test_bicycle_is_a_shipped_kernel():
    from mpc_planner_b200 import binding

    lib = binding.load_library()
    assert lib.mpcgpu_problem_supported(ctypes.byref(pb.config_bicycle().desc())) == 1


@pytest.mark.gpu
def test_gpu_bicycle_stage_functions_and_solves_match_oracle():
    from mpc_planner_b200 import binding

    spec = pb.config_bicycle()
    d = spec.desc()
    solver = binding.BatchSolver(spec, 16)
    z, p, lam, pi = samples(spec, 8, seed=2)
    g = solver.eval_stage(z, p, lam, pi)
    for m in range(len(z)):
        o = oracle.eval_stage(d, z[m], p[m], lam[m], pi[m])
        for k in ("cost", "cost_grad", "cost_hess", "h", "h_jac", "h_hess", "xnext", "AB", "dyn_hess"):
            a, b = np.asarray(g[k][m]), np.asarray(o[k])
            assert np.abs(a - b).max() <= 1e-9 * (1 + np.abs(b).max()), (m, k)
    b = scenes.make_bicycle_batch(spec, 16, seed=4)
    xt, ut, info, best = solver.solve(b["xinit"], b["x0"], b["params"])
    oxt, out, oinfo = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    n_ok = 0
    for i in range(16):
        assert same_status(info[i], oinfo[i]), (i, info[i], oinfo[i])
        if oinfo[i]["exit_code"] == 1:
            n_ok += 1
            np.testing.assert_allclose(xt[i], oxt[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ut[i], out[i], rtol=1e-6, atol=1e-6)
    assert n_ok >= 10
    solver.close()
