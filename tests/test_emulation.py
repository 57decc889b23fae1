"""The kernel SOURCE (mpc_planner_b200/csrc/*.cuh), compiled for the host as a sequential warp
emulation (tests/emu), against the CPU oracle.  This exercises the hand-derived derivatives and
the warp-phased SQP / IPM / Riccati logic on the CPU; the GPU parity tests (-m gpu) run the same
source through nvcc and the C ABI."""
import numpy as np
import pytest

from mpc_planner_b200 import problem as pb
from mpc_planner_b200 import scenes
from oracle import oracle
from tests.emu import emu
from tests.util import same_status

SPECS = {
    "cfg1": pb.config_basic(),
    "cfg2": pb.config_tmpc(),
    "cfg2_gaussian": pb.config_tmpc(gaussian=True),
    "cfg3": pb.config_tmpc_decomp(),
    "cfg4": pb.config_safe_horizon(),
}


@pytest.mark.parametrize("name", list(SPECS))
def test_hand_derived_derivatives_match_oracle_ad(golden, name):
    spec, g = SPECS[name], golden[name]
    d = spec.desc()
    rng = np.random.default_rng(0)
    for s in g["samples"]:
        z, p = np.array(s["z"]), np.array(s["p"])
        lam, pi = rng.uniform(0, 1, spec.nh), rng.uniform(-1, 1, spec.nx)
        o = oracle.eval_stage(d, z, p, lam, pi)
        e = emu.eval_stage(d, z, p, lam, pi)
        assert e["cost"][0] == pytest.approx(s["cost"], rel=1e-12, abs=1e-12)  # the reference's own value
        np.testing.assert_allclose(e["h"][0], s["h"], rtol=1e-12, atol=1e-12)
        for key in ("cost_grad", "cost_hess", "h_jac", "h_hess", "xnext", "AB", "dyn_hess"):
            ref = np.asarray(o[key])
            np.testing.assert_allclose(e[key][0], ref, rtol=1e-10, atol=1e-11 * (1 + np.abs(ref).max()), err_msg=key)


@pytest.mark.parametrize("name,S,P", [("cfg2", 3, 8), ("cfg1", 6, 1), ("cfg2_gaussian", 2, 4), ("cfg3", 2, 4), ("cfg4", 2, 4)])
def test_emulated_kernel_solves_match_oracle(name, S, P):
    spec = SPECS[name]
    b = scenes.make_scene_batch(spec, S, P, seed=31)
    xt, ut, info = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    ext, eut, einfo = emu.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    n_ok = 0
    for i in range(S * P):
        assert same_status(einfo[i], info[i]), (i, einfo[i], info[i])
        if info[i]["exit_code"] == 1:
            n_ok += 1
            np.testing.assert_allclose(ext[i], xt[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(eut[i], ut[i], rtol=1e-6, atol=1e-6)
            assert einfo[i]["pobj"] == pytest.approx(info[i]["pobj"], rel=1e-6, abs=1e-6)
    assert n_ok >= max(1, S * P // 3)


def test_lane_cooperative_factorisation_variant_matches_oracle(tmp_path):
    """-DMPC_RICCATI_LANES (csrc/sqp_core.cuh riccati_factor_lanes: one warp, lane = column of the stage KKT block, kept as
    a measured alternative to the single-thread sweep) must solve the same QPs: emulated lane by lane against the oracle."""
    lib = emu.build_variant(["MPC_RICCATI_LANES"], str(tmp_path / "libmpcgpu_emu_lanes.so"))
    for name, S, P in (("cfg2", 1, 8), ("cfg4", 1, 4)):
        spec = SPECS[name]
        b = scenes.make_scene_batch(spec, S, P, seed=31)
        xt, ut, info = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
        ext, eut, einfo = emu.solve_batch(spec, b["xinit"], b["x0"], b["params"], library=lib)
        n_ok = 0
        for i in range(S * P):
            assert same_status(einfo[i], info[i])
            if info[i]["exit_code"] == 1:
                n_ok += 1
                np.testing.assert_allclose(ext[i], xt[i], rtol=1e-6, atol=1e-6)
                np.testing.assert_allclose(eut[i], ut[i], rtol=1e-6, atol=1e-6)
        assert n_ok >= 1


@pytest.mark.parametrize("m", [2, 4, 6, 8, 10])
def test_tournament_schedule_covers_every_pair_once_with_disjoint_rounds(m):
    """The parallel cyclic Jacobi relies on it: m - 1 rounds of m/2 pairwise DISJOINT index pairs, every pair exactly once."""
    import ctypes

    L = emu.lib()
    seen = set()
    for r in range(m - 1):
        used = set()
        for j in range(m // 2):
            p, q = ctypes.c_int(), ctypes.c_int()
            L.emu_tournament_pair(m, r, j, ctypes.byref(p), ctypes.byref(q))
            assert 0 <= p.value < q.value < m
            assert p.value not in used and q.value not in used  # disjoint within the round
            used |= {p.value, q.value}
            assert (p.value, q.value) not in seen
            seen.add((p.value, q.value))
    assert len(seen) == m * (m - 1) // 2


def test_general_mirror_path_runs_the_parallel_jacobi(tmp_path):
    """cfg 3 (CA objective: stage Hessians are not block diagonal) must reach mirror_general_cta -- the round-robin parallel
    cyclic Jacobi -- and still match the oracle's eigen-decomposition based MIRROR."""
    import ctypes

    lib = emu.build_variant(["MPC_EMU_TRACE"], str(tmp_path / "libmpcgpu_emu_trace.so"))
    spec = SPECS["cfg3"]
    b = scenes.make_scene_batch(spec, 1, 4, seed=31)
    xt, ut, info = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    ext, eut, einfo = emu.solve_batch(spec, b["xinit"], b["x0"], b["params"], library=lib)
    L = ctypes.CDLL(lib)
    L.emu_mirror_general_count.restype = ctypes.c_long
    assert L.emu_mirror_general_count() > 0
    n_ok = 0
    for i in range(4):
        assert einfo[i]["exit_code"] == info[i]["exit_code"]
        if info[i]["exit_code"] == 1:
            n_ok += 1
            np.testing.assert_allclose(ext[i], xt[i], rtol=1e-6, atol=1e-6)
    assert n_ok >= 1


def test_solver_timeout_iteration_count_logic():
    """Solver::solve()'s timeout (acados_solver_interface.cpp:91-117) as restated in csrc/sqp_core.cuh: stop after iteration
    i once elapsed + average iteration time >= budget.  The emulation's clock advances 1 ms per read, so: start = 1 ms,
    after iteration i: elapsed = i ms, average = 1 ms."""
    import dataclasses

    spec = SPECS["cfg2"]
    b = scenes.make_scene_batch(spec, 1, 4, seed=31)
    try:
        for budget_s, expect in ((1.0, 10), (0.0101, 10), (0.0035, 3), (0.0005, 1)):
            emu.set_timeout(budget_s, 1_000_000)
            xt, ut, info = emu.solve_batch(spec, b["xinit"], b["x0"], b["params"])
            assert [i["sqp_iter"] for i in info] == [expect] * 4, (budget_s, info)
            if expect == 3:  # a timed-out solve equals the plain solve with n_sqp = 3 (oracle)
                oxt, out, oinfo = oracle.solve_batch(dataclasses.replace(spec, n_sqp=3), b["xinit"], b["x0"], b["params"])
                np.testing.assert_allclose(xt, oxt, rtol=1e-6, atol=1e-6)
                assert [i["exit_code"] for i in info] == [i["exit_code"] for i in oinfo]
    finally:
        emu.set_timeout(0.0, 0)


def test_scene_generator_invariants():
    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 2, 8, seed=4)
    pm = spec.parameter_map()
    N, npar, nz, nu = spec.N, spec.npar, spec.nz, spec.nu
    par = b["params"].reshape(16, N, npar)
    x0 = b["x0"].reshape(16, N + 1, nz)
    # stage-0 dummies (ellipsoid_constraints.cpp:42-55, linearized_constraints.cpp:155-166)
    np.testing.assert_allclose(par[:, 0, pm["ellipsoid_obst_0_x"]], b["xinit"][:, 0] + 50)
    np.testing.assert_allclose(par[:, 0, pm["lin_constraint_0_a1"]], 1.0)
    np.testing.assert_allclose(par[:, 0, pm["lin_constraint_0_b"]], b["xinit"][:, 0] + 100)
    # guess 0 is the non-guided planner: dummy halfspaces at every stage
    np.testing.assert_allclose(par[0, :, pm["lin_constraint_3_a2"]], 0.0)
    # warm starts are dynamically consistent and start at xinit
    np.testing.assert_allclose(x0[:, 0, nu:], b["xinit"])
    # topology halfspaces: unit normal towards the obstacle, plane 1e-3 + robot_radius from its centre
    for i in range(1, 8):
        for k in range(1, N):
            for o in range(8):
                a1, a2, bb = (par[i, k, pm[f"lin_constraint_{o}_{s}"]] for s in ("a1", "a2", "b"))
                ox, oy = par[i, k, pm[f"ellipsoid_obst_{o}_x"]], par[i, k, pm[f"ellipsoid_obst_{o}_y"]]
                assert a1 * a1 + a2 * a2 == pytest.approx(1.0)
                assert bb == pytest.approx(a1 * ox + a2 * oy - (1e-3 + scenes.ROBOT_RADIUS))
                dist = np.hypot(ox - x0[i, k, nu], oy - x0[i, k, nu + 1])
                assert a1 * x0[i, k, nu] + a2 * x0[i, k, nu + 1] - bb == pytest.approx(-dist + 1e-3 + scenes.ROBOT_RADIUS)


@pytest.mark.parametrize("name", list(SPECS))
def test_structural_sparsity_assumed_by_the_riccati_sweeps(golden, name):
    """csrc/sqp_core.cuh (RicIdx) hard-codes the non-zero pattern of [B A] of the unicycle family;
    check it against the actual Jacobians."""
    spec, g = SPECS[name], golden[name]
    d = spec.desc()
    nx, nu = spec.nx, spec.nu
    enz = lambda l, c: (l in (0, 1)) if c == 0 else (l in (0, 1, 4))  # noqa: E731
    bnz = lambda l, u: (l in (0, 1, 3, 4)) if u == 0 else (l in (0, 1, 2))  # noqa: E731
    for s in g["samples"][:8]:
        AB = emu.eval_stage(d, np.array(s["z"]), np.array(s["p"]))["AB"][0]
        B, A = AB[:, :nu], AB[:, nu:]
        E = A - np.eye(nx)
        for l in range(nx):
            for u in range(nu):
                if not bnz(l, u):
                    assert B[l, u] == 0.0
            for j in range(nx):
                c = {2: 0, 3: 1}.get(j, -1)
                if c < 0 or not enz(l, c):
                    assert E[l, j] == 0.0


def test_compare_select_forms_equal_fmax():
    """The row bodies use `(v > floor) ? v : floor` and `if (v > acc) acc = v` where they used fmax (half of their instructions):
    same value for every input the callers can produce -- any v, NaN included, against a non-NaN floor / running maximum."""
    import ctypes
    import math

    L = emu.lib()
    for f in (L.emu_floor_at, L.emu_raise):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double, ctypes.c_double]
    vals = [0.0, -0.0, 1e-300, -1e-300, 1.0, -1.0, 1e-10, 1e-16, 3.5, -7.25, 1e300, -1e300, math.inf, -math.inf, math.nan]
    libm = ctypes.CDLL("libm.so.6")  # C fmax: the function the kernel source called
    libm.fmax.restype = ctypes.c_double
    libm.fmax.argtypes = [ctypes.c_double, ctypes.c_double]
    same = lambda a, b: (a != a and b != b) or a == b
    for v in vals:
        for c in vals:
            if c == c:
                assert same(L.emu_floor_at(v, c), libm.fmax(v, c)), (v, c)
                assert same(L.emu_raise(c, v), libm.fmax(c, v)), (c, v)


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_constraint_rows_are_dealt_to_the_four_linearisation_threads_exactly_once(cfg):
    """linearize_part gives thread q of a stage share q of 4 of EVERY constraint block (stage_constraints(..., q, 4)): over
    q = 0..3 each row of h is evaluated exactly once, whatever the block sizes (4; 8 + 8; 12 + 12 + 12; 24)."""
    import ctypes

    L = emu.lib()
    hit = (ctypes.c_int * 64)()
    nh = -1
    for q in range(4):
        nh = L.emu_quarter_rows(cfg, q, hit)
    assert nh > 0 and list(hit[:nh]) == [1] * nh and not any(hit[nh:])
