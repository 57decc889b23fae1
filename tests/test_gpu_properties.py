"""Size-independent properties of the CUDA path at BASELINE sizes, and edge cases of the C ABI."""
import ctypes

import numpy as np
import pytest

from mpc_planner_b200 import problem as pb
from mpc_planner_b200 import scenes
from tests.util import same_status

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding():
    from mpc_planner_b200 import binding as b

    assert b.load_library().mpcgpu_device_count() >= 1
    return b


@pytest.fixture(scope="module")
def big_batch():
    spec = pb.config_tmpc()
    base = scenes.make_scene_batch(spec, 64, 8, seed=77)
    reps = 8  # 4096 solves = the bench step
    return spec, {k: np.tile(base[k], (reps, 1)) for k in ("xinit", "x0", "params")}


def _rk4(x, u, dt, steps=3):
    def f(x):
        return np.stack([x[:, 3] * np.cos(x[:, 2]), x[:, 3] * np.sin(x[:, 2]), u[:, 1], u[:, 0], x[:, 3]], 1)

    h = dt / steps
    for _ in range(steps):
        k1 = f(x)
        k2 = f(x + 0.5 * h * k1)
        k3 = f(x + 0.5 * h * k2)
        k4 = f(x + h * k3)
        x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def test_full_size_batch_properties(binding, big_batch):
    spec, b = big_batch
    P = len(b["xinit"])
    solver = binding.BatchSolver(spec, P)
    solver.set_group(8)
    xt, ut, info, best = solver.solve(b["xinit"], b["x0"], b["params"])
    codes = np.array([i["exit_code"] for i in info])
    ok = codes == 1
    assert ok.mean() > 0.7 and set(np.unique(codes)) <= {0, 1, 2, 3, 4}
    # x0 = xinit exactly (node 0 is eliminated from the QP and re-imposed by the step)
    np.testing.assert_allclose(xt[ok, 0], b["xinit"][ok], atol=1e-12)
    # shooting defects of every successful solve are below the reference's res_eq threshold (1e-2): 10 RTI
    # steps are not full convergence, what remains is the second-order term of the last full step
    d = spec.desc()
    lo, up = np.array(list(d.lower_bound)[: spec.nz]), np.array(list(d.upper_bound)[: spec.nz])
    worst = 0.0
    for k in range(spec.N):
        worst = max(worst, np.abs(_rk4(xt[ok, k], ut[ok, k], spec.dt) - xt[ok, k + 1]).max())
    assert worst < 1e-2
    assert max(i["res_eq"] for i, o in zip(info, ok) if o) < 1e-2
    # box constraints hold up to the QP tolerance at every stage
    assert (ut[ok] >= lo[: spec.nu] - 1e-3).all() and (ut[ok] <= up[: spec.nu] + 1e-3).all()
    assert (xt[ok, 1:spec.N] >= lo[spec.nu:] - 1e-3).all() and (xt[ok, 1:spec.N] <= up[spec.nu:] + 1e-3).all()
    # collision constraints of stage k >= 1: ellipsoid value >= 1 (up to tolerance)
    pm = spec.parameter_map()
    par = b["params"].reshape(P, spec.N, spec.npar)
    r = scenes.ROBOT_RADIUS + scenes.OBSTACLE_RADIUS
    for o in range(8):
        ox, oy = par[ok, 1:, pm[f"ellipsoid_obst_{o}_x"]], par[ok, 1:, pm[f"ellipsoid_obst_{o}_y"]]
        h = ((xt[ok, 1:spec.N, 0] - ox) ** 2 + (xt[ok, 1:spec.N, 1] - oy) ** 2) / r ** 2
        assert h.min() > 1.0 - 5e-2  # linearised constraint + full step: violated at most to second order
    # per-scene winner = argmin objective over successful guesses (FindBestPlanner)
    obj = np.array([i["pobj"] for i in info]).reshape(-1, 8)
    okg = ok.reshape(-1, 8)
    for g in range(P // 8):
        if okg[g].any():
            want = int(np.argmin(np.where(okg[g], obj[g], np.inf)))
            assert best[g] == want or obj[g, best[g]] == obj[g, want]
        else:
            assert best[g] == -1
    # determinism: the same inputs give bit-identical outputs; the batch is tiled 8x, so replicas must agree too
    xt2, ut2, info2, best2 = solver.solve(b["xinit"], b["x0"], b["params"])
    np.testing.assert_array_equal(xt, xt2)
    np.testing.assert_array_equal(xt[:512], xt[512:1024])
    solver.close()


def test_batch_order_does_not_matter(binding, big_batch):
    spec, b = big_batch
    idx = np.random.default_rng(0).permutation(256)
    solver = binding.BatchSolver(spec, 256)
    xt, ut, info, _ = solver.solve(b["xinit"][:256], b["x0"][:256], b["params"][:256])
    xtp, utp, infop, _ = solver.solve(b["xinit"][idx], b["x0"][idx], b["params"][idx])
    np.testing.assert_array_equal(xt[idx], xtp)
    assert [info[i]["exit_code"] for i in idx] == [i["exit_code"] for i in infop]
    solver.close()


def test_ragged_and_tiny_batches(binding, big_batch):
    spec, b = big_batch
    solver = binding.BatchSolver(spec, 64)
    ref_xt, _, ref_info, _ = solver.solve(b["xinit"][:64], b["x0"][:64], b["params"][:64])
    for P in (1, 3, 13):  # P = 1 is the plain Solver::solve(); 13 is not a multiple of the group size
        solver.set_group(8 if P > 8 else 0)
        xt, ut, info, best = solver.solve(b["xinit"][:P], b["x0"][:P], b["params"][:P])
        np.testing.assert_array_equal(xt, ref_xt[:P])
        if P == 13:
            assert len(best) == 2
    solver.close()


def test_errors_are_reported_not_swallowed(binding, big_batch):
    spec, b = big_batch
    solver = binding.BatchSolver(spec, 4)
    with pytest.raises(binding.MpcGpuError, match="max_batch"):
        solver.solve(b["xinit"][:8], b["x0"][:8], b["params"][:8])
    lib = binding.load_library()
    assert lib.mpcgpu_solve_host(solver.handle, 0, None, None, None, None, None, None, None) != 0  # empty batch is an argument error
    solver.close()


def test_nan_inputs_give_failure_exit_codes(binding, big_batch):
    spec, b = big_batch
    xi, x0, par = b["xinit"][:8].copy(), b["x0"][:8].copy(), b["params"][:8].copy()
    par[3, 5] = np.nan  # poison one solve's parameters
    x0[5, 10] = np.inf
    solver = binding.BatchSolver(spec, 8)
    xt, ut, info, best = solver.solve(xi, x0, par)
    assert info[3]["exit_code"] != 1 and info[5]["exit_code"] != 1
    clean, _, cinfo, _ = solver.solve(b["xinit"][:8], b["x0"][:8], b["params"][:8])
    for i in (0, 1, 2, 4, 6, 7):  # the neighbours are unaffected
        np.testing.assert_array_equal(xt[i], clean[i])
    assert best != 3 and best != 5
    solver.close()


def test_scene_without_obstacles(binding):
    """All constraint slots hold the reference's dummies (far-away obstacles, halfspace x <= x0 + 100)."""
    from oracle import oracle

    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 2, 2, seed=5)
    pm = spec.parameter_map()
    par = b["params"].reshape(4, spec.N, spec.npar).copy()
    for o in range(8):
        par[:, :, pm[f"ellipsoid_obst_{o}_x"]] = b["xinit"][:, None, 0] + 50
        par[:, :, pm[f"ellipsoid_obst_{o}_y"]] = b["xinit"][:, None, 1] + 50
        par[:, :, pm[f"lin_constraint_{o}_a1"]], par[:, :, pm[f"lin_constraint_{o}_a2"]] = 1.0, 0.0
        par[:, :, pm[f"lin_constraint_{o}_b"]] = b["xinit"][:, None, 0] + 100
    solver = binding.BatchSolver(spec, 4)
    xt, ut, info, _ = solver.solve(b["xinit"], b["x0"], par.reshape(4, -1))
    oxt, _, oinfo = oracle.solve_batch(spec, b["xinit"], b["x0"], par.reshape(4, -1))
    assert all(i["exit_code"] == 1 for i in info)
    np.testing.assert_allclose(xt, oxt, rtol=1e-6, atol=1e-6)
    solver.close()


def test_solver_timeout_bounds_the_rti_iterations(binding):
    """`_params.solver_timeout` (acados_solver_interface.cpp:91-117): a budget no iteration fits in stops every solve after
    its first RTI iteration -- the result is then exactly the oracle's with n_sqp = 1; a generous budget changes nothing;
    0 switches the check off again."""
    import dataclasses

    from oracle import oracle

    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 4, 8, seed=9)
    solver = binding.BatchSolver(spec, 32)
    xt10, ut10, info10, _ = solver.solve(b["xinit"], b["x0"], b["params"])
    assert all(i["sqp_iter"] == spec.n_sqp for i in info10 if i["qp_status"] == 0)
    solver.set_timeout(1e-9)
    xt1, ut1, info1, _ = solver.solve(b["xinit"], b["x0"], b["params"])
    assert all(i["sqp_iter"] == 1 for i in info1)
    spec1 = dataclasses.replace(spec, n_sqp=1)
    oxt, out, oinfo = oracle.solve_batch(spec1, b["xinit"], b["x0"], b["params"])
    for i in range(32):
        assert same_status(info1[i], oinfo[i])
    np.testing.assert_allclose(xt1, oxt, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ut1, out, rtol=1e-6, atol=1e-6)
    solver.set_timeout(10.0)
    xt, ut, info, _ = solver.solve(b["xinit"], b["x0"], b["params"])
    assert np.array_equal(xt, xt10) and np.array_equal(ut, ut10)
    solver.set_timeout(0.0)
    xt, ut, info, _ = solver.solve(b["xinit"], b["x0"], b["params"])
    assert np.array_equal(xt, xt10) and [i["sqp_iter"] for i in info] == [i["sqp_iter"] for i in info10]
    solver.close()


def _rk4_family(spec, x, u):
    """ERK4 x 3 of the unicycle family (states x, y, psi, v, spline[, slack]; slack' = 0), vectorised over solves."""
    out = _rk4(x[:, :5], u, spec.dt)
    return np.concatenate([out, x[:, 5:]], axis=1) if x.shape[1] > 5 else out


def check_batch_properties(spec, b, xt, ut, info, best, group):
    """Size-independent properties of one solved batch (shared by the GPU test below and its CPU dry run on the emulation)."""
    P = len(info)
    codes = np.array([i["exit_code"] for i in info])
    ok = codes == 1
    assert ok.mean() > 0.5 and set(np.unique(codes)) <= {0, 1, 2, 3, 4}
    np.testing.assert_allclose(xt[ok, 0], b["xinit"][ok], atol=1e-12)
    worst = max(np.abs(_rk4_family(spec, xt[ok, k], ut[ok, k]) - xt[ok, k + 1]).max() for k in range(spec.N))
    assert worst < 1e-2 and max(i["res_eq"] for i, o in zip(info, ok) if o) <= 1e-2
    d = spec.desc()
    lo, up = np.array(list(d.lower_bound)[: spec.nz]), np.array(list(d.upper_bound)[: spec.nz])
    assert (ut[ok] >= lo[: spec.nu] - 1e-3).all() and (ut[ok] <= up[: spec.nu] + 1e-3).all()
    assert (xt[ok, 1:spec.N] >= lo[spec.nu:] - 1e-3).all() and (xt[ok, 1:spec.N] <= up[spec.nu:] + 1e-3).all()
    obj = np.array([i["pobj"] for i in info]).reshape(-1, group)
    okg = ok.reshape(-1, group)
    for g in range(P // group):
        if okg[g].any():
            want = int(np.argmin(np.where(okg[g], obj[g], np.inf)))
            assert best[g] == want or obj[g, best[g]] == obj[g, want]
        else:
            assert best[g] == -1


@pytest.mark.parametrize("name,S,G", [("cfg3", 1, 64), ("cfg4", 256, 4)])
def test_baseline_size_batches_of_the_other_configs(binding, name, S, G):
    """BASELINE configs[2] (T-MPC++ P = 64, N = 30, decomp halfspaces, CA cost) and configs[3] (SH-MPC: 256 scenario sets x 4
    guesses) at their full batch sizes: size-independent properties, a sample against the oracle, determinism."""
    from oracle import oracle

    spec = pb.CONFIGS[name]()
    base = scenes.make_scene_batch(spec, min(S, 16), G, seed=21)  # the batches scripts/config_sweep.py measures
    reps = S // min(S, 16)
    b = {k: np.tile(base[k], (reps, 1)) for k in ("xinit", "x0", "params")}
    P = S * G
    solver = binding.BatchSolver(spec, P)
    solver.set_group(G)
    xt, ut, info, best = solver.solve(b["xinit"], b["x0"], b["params"])
    check_batch_properties(spec, b, xt, ut, info, best, G)
    n = 6
    oxt, out, oinfo = oracle.solve_batch(spec, b["xinit"][:n], b["x0"][:n], b["params"][:n])
    for i in range(n):
        assert same_status(info[i], oinfo[i])
        if oinfo[i]["exit_code"] == 1:
            np.testing.assert_allclose(xt[i], oxt[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ut[i], out[i], rtol=1e-6, atol=1e-6)
    xt2, ut2, info2, best2 = solver.solve(b["xinit"], b["x0"], b["params"])
    np.testing.assert_array_equal(xt, xt2)
    np.testing.assert_array_equal(best, best2)
    if reps > 1:  # the batch is tiled: replicas must agree bit for bit
        np.testing.assert_array_equal(xt[: P // reps], xt[P // reps: 2 * (P // reps)])
    solver.close()
