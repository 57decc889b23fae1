"""Persistent NLP multipliers per solver slot: `Solver::operator=` copies parameters only, so a
reference LocalPlanner's acados memory -- the multipliers that enter the first exact Hessian -- survives from one
control step to the next (acados_solver_interface.cpp:67-77).  Oracle and kernel (emulated here, real on the GPU)
implement the same state [pi | multipliers of h] and must agree over consecutive control steps."""
import numpy as np
import pytest

from mpc_planner_b200 import problem as pb, scenes
from oracle import oracle
from tests.emu import emu


def two_steps(spec, solve, state):
    """control step 1 on the scene, control step 2 warm-started at step 1's solution shifted by the applied input"""
    b = scenes.make_scene_batch(spec, 2, 2, seed=21)
    N, nu = spec.N, spec.nu
    r1 = solve(b["xinit"], b["x0"], b["params"], state)
    xt, ut = r1[0], r1[1]
    z = np.zeros((xt.shape[0], N + 1, spec.nz))
    z[:, :N, :nu] = ut
    z[:, :, nu:] = xt
    r2 = solve(b["xinit"], z.reshape(xt.shape[0], -1), b["params"], state)
    return r1, r2


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_emulated_kernel_and_oracle_agree_over_two_stateful_steps(name):
    spec = pb.CONFIGS[name]()
    P = 4
    so = np.zeros((P, oracle.state_size(spec)))
    se = np.zeros_like(so)
    o1, o2 = two_steps(spec, lambda a, b, c, st: oracle.solve_batch(spec, a, b, c, state=st), so)
    e1, e2 = two_steps(spec, lambda a, b, c, st: emu.solve_batch(spec, a, b, c, state=st), se)
    ok = np.array([i["exit_code"] == 1 for i in o2[2]]) & np.array([i["exit_code"] == 1 for i in o1[2]])
    assert ok.any()
    assert [i["exit_code"] for i in e2[2]] == [i["exit_code"] for i in o2[2]]
    np.testing.assert_allclose(e2[0][ok], o2[0][ok], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(se[ok], so[ok], rtol=1e-5, atol=1e-6)
    assert np.abs(so[ok]).max() > 1e-3  # active constraints / dynamics leave non-zero multipliers
    # the state matters: a stateless second step takes a different first exact Hessian
    z = np.zeros((P, spec.N + 1, spec.nz))
    z[:, : spec.N, : spec.nu] = o1[1]
    z[:, :, spec.nu:] = o1[0]
    b = scenes.make_scene_batch(spec, 2, 2, seed=21)
    fresh = oracle.solve_batch(spec, b["xinit"], z.reshape(P, -1), b["params"])
    assert np.abs(fresh[0][ok] - o2[0][ok]).max() > 1e-13  # (both runs converge towards the same KKT point)


def test_failed_solve_clears_the_slot():
    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 1, 2, seed=3)
    bad = b["params"].copy()
    bad[0, :] = np.nan  # non-finite parameters: QP failure
    st = np.ones((2, oracle.state_size(spec)))
    _, _, info = oracle.solve_batch(spec, b["xinit"], b["x0"], bad, state=st)
    assert info[0]["exit_code"] == 4 and not st[0].any()
    se = np.ones_like(st)
    _, _, info_e = emu.solve_batch(spec, b["xinit"], b["x0"], bad, state=se)
    assert info_e[0]["exit_code"] == 4 and not se[0].any()
    assert np.isfinite(se).all() and np.isfinite(st).all()


def demoted_batch(spec):
    """Solves whose single RTI step leaves a dynamics defect above 1e-2: the reference then turns ACADOS_SUCCESS into
    ACADOS_QP_FAILURE (acados_solver_interface.cpp:176-181) although the QP itself was solved (qp_status 0 or 2)."""
    import copy

    sp = copy.deepcopy(spec)
    sp.n_sqp = 1
    b = scenes.make_scene_batch(sp, 4, 2, seed=8)
    rng = np.random.default_rng(8)
    x0 = b["x0"].reshape(8, sp.N + 1, sp.nz).copy()
    x0[:, 1:, sp.nu + 2] += rng.uniform(-1.5, 1.5, size=x0[:, 1:, 0].shape)  # heading of the warm start far off
    x0[:, 1:, sp.nu + 3] += rng.uniform(0.0, 1.5, size=x0[:, 1:, 0].shape)   # and its speed
    return sp, b["xinit"], x0.reshape(8, -1), b["params"]


def test_res_eq_demotion_clears_the_slot_too():
    """the reference resets the solver memory on EVERY non-success exit (Solver_acados_reset(capsule, 1),
    acados_solver_interface.cpp:183-191), not only on QP failures."""
    sp, xinit, x0, params = demoted_batch(pb.config_tmpc())
    st = np.ones((8, oracle.state_size(sp)))
    _, _, info = oracle.solve_batch(sp, xinit, x0, params, state=st)
    demoted = [i for i, inf in enumerate(info) if inf["exit_code"] == 4 and inf["qp_status"] in (0, 2)]
    kept = [i for i, inf in enumerate(info) if inf["exit_code"] == 1]
    assert demoted, [(i["exit_code"], i["qp_status"], i["res_eq"]) for i in info]
    for i in demoted:
        assert info[i]["res_eq"] > 1e-2 and not st[i].any()
    for i in kept:
        assert st[i].any()
    se = np.ones_like(st)
    _, _, info_e = emu.solve_batch(sp, xinit, x0, params, state=se)
    assert [i["exit_code"] for i in info_e] == [i["exit_code"] for i in info]
    for i in demoted:
        assert not se[i].any()
    np.testing.assert_allclose(se, st, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_gpu_stateful_slots_match_oracle(name):
    from mpc_planner_b200 import binding

    spec = pb.CONFIGS[name]()
    P = 4
    solver = binding.BatchSolver(spec, P)
    solver.set_stateful(True)
    assert solver.state_size() == oracle.state_size(spec)
    so = np.zeros((P, oracle.state_size(spec)))
    o1, o2 = two_steps(spec, lambda a, b, c, st: oracle.solve_batch(spec, a, b, c, state=st), so)
    g1, g2 = two_steps(spec, lambda a, b, c, st: solver.solve(a, b, c)[:3], None)
    ok = np.array([i["exit_code"] == 1 for i in o2[2]]) & np.array([i["exit_code"] == 1 for i in o1[2]])
    assert ok.any()
    assert [i["exit_code"] for i in g2[2]] == [i["exit_code"] for i in o2[2]]
    np.testing.assert_allclose(g2[0][ok], o2[0][ok], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(solver.get_state()[ok], so[ok], rtol=1e-5, atol=1e-6)
    # reset == fresh solver; set_state restores a checkpoint
    saved = solver.get_state()
    solver.reset_state()
    assert not solver.get_state().any()
    b = scenes.make_scene_batch(spec, 2, 2, seed=21)
    fresh_gpu = solver.solve(b["xinit"], b["x0"], b["params"])
    np.testing.assert_allclose(fresh_gpu[0][ok], o1[0][ok], rtol=1e-6, atol=1e-6)
    solver.set_state(saved)
    np.testing.assert_array_equal(solver.get_state(), saved)
    solver.close()


@pytest.mark.gpu
def test_gpu_res_eq_demotion_clears_the_slot():
    from mpc_planner_b200 import binding

    sp, xinit, x0, params = demoted_batch(pb.config_tmpc())
    st = np.ones((8, oracle.state_size(sp)))
    _, _, info = oracle.solve_batch(sp, xinit, x0, params, state=st)
    solver = binding.BatchSolver(sp, 8)
    solver.set_stateful(True)
    solver.set_state(np.ones_like(st))
    _, _, ginfo, _ = solver.solve(xinit, x0, params)
    assert [i["exit_code"] for i in ginfo] == [i["exit_code"] for i in info]
    demoted = [i for i, inf in enumerate(info) if inf["exit_code"] == 4 and inf["qp_status"] in (0, 2)]
    assert demoted
    gs = solver.get_state()
    for i in demoted:
        assert not gs[i].any()
    np.testing.assert_allclose(gs, st, rtol=1e-5, atol=1e-6)
    solver.close()
