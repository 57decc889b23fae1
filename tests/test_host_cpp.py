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

CONFIG_N_PATHS = 4  # host Config::n_paths: index of the non-guided planner in GuidanceConstraints::planners_


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
    # second control step of solver 0 with its own multipliers kept (Solver::operator= is parameters-only):
    # equals the oracle run with persistent state
    from oracle import oracle

    (code2,) = struct.unpack_from("i", raw, off)
    off += 4
    (pobj2,) = struct.unpack_from("d", raw, off)
    off += 8
    cx2 = np.frombuffer(raw, dtype=np.float64, count=nxt, offset=off)
    st = np.zeros((1, oracle.state_size(spec)))
    o1 = oracle.solve_batch(spec, b["xinit"][:1], b["x0"][:1], b["params"][:1], state=st)
    z = np.zeros((1, spec.N + 1, spec.nz))
    z[:, : spec.N, : spec.nu] = o1[1]
    z[:, :, spec.nu:] = o1[0]
    o2 = oracle.solve_batch(spec, b["xinit"][:1], z.reshape(1, -1), b["params"][:1], state=st)
    assert code2 == o2[2][0]["exit_code"]
    if code2 == 1:
        np.testing.assert_allclose(cx2, o2[0][0].ravel(), rtol=1e-6, atol=1e-6)
        assert pobj2 == pytest.approx(o2[2][0]["pobj"], rel=1e-6)


# ---------------------------------------------------------------------------------------------
# module layer: parameter packing (MPCBase / Contouring / Ellipsoid / Linearized) + batched T-MPC++ loop
# ---------------------------------------------------------------------------------------------
def _make_scene(tmp_path, consistency=1.0, selected=-1):
    from mpc_planner_b200.scenes import _fit_path, _path_eval

    spec = pb.config_tmpc()
    N, dt = spec.N, spec.dt
    coef, starts = _fit_path(1.2, 6, phase=0.4)  # 6 segments: the module picks 5 starting at the closest one
    s0 = 1.0
    p0, d0 = _path_eval(coef, starts, s0)
    state = dict(x=p0[0], y=p0[1] + 0.1, psi=float(np.arctan2(d0[1], d0[0])), v=1.5, spline=0.0)
    t = dt * np.arange(N + 1)
    base_s = s0 + 2.0 * t
    base, der = _path_eval(coef, starts, base_s)
    head = np.arctan2(der[:, 1], der[:, 0])
    nrm = np.stack([-np.sin(head), np.cos(head)], 1)
    # 8 obstacle slots (ensureObstacleSize pads to max_obstacles): 3 real ones + far-away dummies
    obs = np.zeros((8, N, 2))  # prediction step j = position at time (j+1) dt
    for o in range(8):
        if o < 3:
            start = base[6 + 3 * o] + (0.3 - 0.3 * o) * nrm[6 + 3 * o]
            vel = np.array([-0.3, 0.2 * (o - 1)])
            obs[o] = start[None] + (t[1:, None] - t[6 + 3 * o]) * vel[None]
        else:
            obs[o] = np.array([state["x"] + 100.0 + o, state["y"] + 100.0])[None]
    guid = []
    for side in (+1.0, -1.0):
        lat = side * 1.3 * np.clip(t / 1.0, 0, 1)
        pos = base + lat[:, None] * nrm
        vel = np.gradient(pos, dt, axis=0)
        guid.append((pos, vel))
    f = tmp_path / "scene.bin"
    vals = [state[k] for k in ("x", "y", "psi", "v", "spline")] + [6]
    for i in range(6):
        vals += list(coef[i, 0]) + list(coef[i, 1]) + [starts[i]]
    vals += [30.0, 8, N]
    for o in range(8):
        vals += list(obs[o].ravel())
    vals += [len(guid), N + 1]
    for g, (pos, vel) in enumerate(guid):
        vals += [1.0 if g == selected else 0.0] + list(pos.ravel()) + list(vel.ravel())
    vals += [consistency]
    np.array(vals, dtype=np.float64).tofile(f)
    return spec, f, state, coef, starts, obs, guid


def _read_modules_output(path, spec):
    raw = np.fromfile(path, dtype=np.float64)
    n, exit_code, best = int(raw[0]), int(raw[1]), int(raw[2])
    off, planners = 3, []
    sz = spec.sizes()
    for _ in range(n):
        d = dict(id=int(raw[off]), original=bool(raw[off + 1]))
        off += 2
        for key in ("xinit", "x0", "params"):
            d[key] = raw[off:off + sz[key]].copy()
            off += sz[key]
        d["exit_code"], d["objective"] = int(raw[off]), raw[off + 1]
        off += 2
        d["xtraj"] = raw[off:off + sz["xtraj"]].copy()
        off += sz["xtraj"]
        d["utraj"] = raw[off:off + sz["utraj"]].copy()
        off += sz["utraj"]
        planners.append(d)
    main_xtraj = raw[off:off + sz["xtraj"]].copy()
    main_pobj = raw[off + sz["xtraj"]]
    return exit_code, best, planners, main_xtraj, main_pobj


@pytest.fixture(scope="module")
def modules_exe():
    return build_host.build_modules_test("cfg2")


def test_module_parameter_packing_follows_the_reference(modules_exe, tmp_path):
    spec, scene, state, coef, starts, obs, guid = _make_scene(tmp_path)
    out = tmp_path / "out.bin"
    r = subprocess.run([modules_exe, "pack", str(scene), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "homotopy mapping and timeout checks passed" in r.stdout  # guidance_constraints.cpp:192-250, 331-333
    _, _, planners, _, _ = _read_modules_output(out, spec)
    assert [p["id"] for p in planners] == [0, 1, 2] and [p["original"] for p in planners] == [False, False, True]
    pm, N, npar, nz, nu = spec.parameter_map(), spec.N, spec.npar, spec.nz, spec.nu
    from mpc_planner_b200.scenes import WEIGHTS, ROBOT_RADIUS, OBSTACLE_RADIUS

    for p in planners:
        par = p["params"].reshape(N, npar)
        x0 = p["x0"].reshape(N + 1, nz)
        # Contouring::update put the closest point into the state / xinit (contouring.cpp:37,42)
        assert abs(p["xinit"][4] - 1.0) < 0.05  # closest-point search on a 1 cm grid; the state is 0.1 m off the path
        for k in range(N):
            for w in ("acceleration", "angular_velocity", "velocity", "reference_velocity", "contour", "lag"):  # mpc_base.cpp:23-35
                assert par[k, pm[w]] == WEIGHTS[w]
            for i in range(5):  # contouring.cpp:94-124: the same 5 segments at every stage, from the closest one (0)
                np.testing.assert_array_equal(par[k, pm[f"spline_x{i}_a"]:pm[f"spline_x{i}_a"] + 4], coef[i, 0])
                np.testing.assert_array_equal(par[k, pm[f"spline_y{i}_a"]:pm[f"spline_y{i}_a"] + 4], coef[i, 1])
                assert par[k, pm[f"spline{i}_start"]] == starts[i]
            assert par[k, pm["ego_disc_radius"]] == ROBOT_RADIUS and par[k, pm["ego_disc_0_offset"]] == 0.0
            for o in range(8):  # ellipsoid_constraints.cpp:34-90
                e = {s: par[k, pm[f"ellipsoid_obst_{o}_{s}"]] for s in ("x", "y", "psi", "major", "minor", "chi", "r")}
                if k == 0:
                    assert (e["x"], e["y"], e["r"], e["chi"]) == (state["x"] + 50, state["y"] + 50, 0.1, 1.0)
                else:  # prediction step k-1 feeds stage k
                    assert (e["x"], e["y"]) == tuple(obs[o, k - 1]) and e["r"] == OBSTACLE_RADIUS and e["chi"] == 1.0 and e["major"] == 0.0
                a1, a2, b = (par[k, pm[f"lin_constraint_{o}_{s}"]] for s in ("a1", "a2", "b"))
                if k == 0 or p["original"]:  # linearized_constraints.cpp:155-166, guidance_constraints.cpp:301-305
                    assert (a1, a2, b) == (1.0, 0.0, state["x"] + 100.0)
                else:  # :85-105, topology mode: r = 1e-3
                    gp = guid[p["id"]][0][k]
                    dvec = obs[o, k - 1] - gp
                    a = dvec / np.linalg.norm(dvec)
                    np.testing.assert_allclose([a1, a2, b], [a[0], a[1], a @ obs[o, k - 1] - (1e-3 + ROBOT_RADIUS)], rtol=1e-12)
        if not p["original"]:  # guidance warm start k = 1..N-1 (guidance_constraints.cpp:401-413)
            pos, vel = guid[p["id"]]
            for k in range(1, N):
                np.testing.assert_allclose(x0[k, nu:nu + 2], pos[k], rtol=1e-14)
                assert x0[k, nu + 2] == pytest.approx(np.arctan2(vel[k, 1], vel[k, 0]))
                assert x0[k, nu + 3] == pytest.approx(np.linalg.norm(vel[k]))
        else:  # braking plan of the main solver (acados_solver_interface.cpp:303-342)
            assert x0[1, nu + 3] == pytest.approx(1.5 - 3.0 * spec.dt) and x0[3, 0] == -3.0


@pytest.mark.gpu
@pytest.mark.parametrize("selected,consistency", [(-1, 1.0), (1, 1e-3)])
def test_batched_tmpc_loop_matches_oracle(modules_exe, tmp_path, selected, consistency):
    from oracle import oracle

    spec, scene, *_ = _make_scene(tmp_path, consistency=consistency, selected=selected)
    out = tmp_path / "out.bin"
    r = subprocess.run([modules_exe, "gpu", str(scene), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exit_code, best, planners, main_xtraj, main_pobj = _read_modules_output(out, spec)
    xi = np.array([p["xinit"] for p in planners])
    x0 = np.array([p["x0"] for p in planners])
    par = np.array([p["params"] for p in planners])
    oxt, out_, oinfo = oracle.solve_batch(spec, xi, x0, par)
    objs = []
    for i, p in enumerate(planners):
        assert p["exit_code"] == oinfo[i]["exit_code"]
        w = consistency if (p["id"] == selected and not p["original"]) else 1.0
        if oinfo[i]["exit_code"] == 1:
            np.testing.assert_allclose(p["xtraj"], oxt[i].ravel(), rtol=1e-6, atol=1e-6)
            assert p["objective"] == pytest.approx(oinfo[i]["pobj"] * w, rel=1e-6)
            objs.append((oinfo[i]["pobj"] * w, i))
    assert objs, "at least one planner must be feasible in this scene"
    want = min(objs)[1]
    assert best == want and exit_code == 1
    np.testing.assert_array_equal(main_xtraj, planners[best]["xtraj"])  # guidance_constraints.cpp:382-384
    if selected >= 0:
        assert best == selected  # the consistency weight makes the previously selected guess win (:358-359)


# ---------------------------------------------------------------------------------------------
# Planner::solveMPC in closed loop (mpc_planner/src/planner.cpp:37-158)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def planner_exe():
    return build_host.build_planner_test("cfg2")


def test_planner_without_data_reports_failure_cpu(planner_exe):
    r = subprocess.run([planner_exe, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_planner_closed_loop_avoids_obstacles_and_follows_the_path(planner_exe, tmp_path):
    """50 control steps of Planner::solveMPC on the CUDA backend (T-MPC++: 2 guided + 1 non-guided guesses per step in
    one launch, warm start shifted forward, multipliers kept per local solver): the robot must stay clear of the three
    crossing obstacles, make progress along the reference path and keep the MPC feasible."""
    from mpc_planner_b200.scenes import OBSTACLE_RADIUS, ROBOT_RADIUS, _fit_path, _path_eval

    n_steps, v_ref, dt = 50, 2.0, 0.2
    coef, starts = _fit_path(1.0, 9, phase=0.3)
    p0, d0 = _path_eval(coef, starts, np.array(0.5))
    state = [p0[0], p0[1] + 0.05, float(np.arctan2(d0[1], d0[0])), 1.0, 0.5]
    obstacles = []
    for s_cross, t_cross, side, speed in ((6.0, 3.0, +1.0, 0.8), (12.0, 6.0, -1.0, 1.0), (18.0, 9.5, +1.0, 0.6)):
        q, dq = _path_eval(coef, starts, np.array(s_cross))
        nrm = np.array([-dq[1], dq[0]]) / np.hypot(*dq)
        vel = -side * speed * nrm  # walks across the path, reaching it at t_cross
        obstacles.append((q - vel * t_cross, vel))
    scene, out = tmp_path / "scene.bin", tmp_path / "out.bin"
    with open(scene, "wb") as f:
        def wd(*vals):
            f.write(struct.pack(f"{len(vals)}d", *[float(v) for v in vals]))
        wd(*state)
        wd(len(starts))
        for i in range(len(starts)):
            wd(*coef[i, 0], *coef[i, 1], starts[i])
        wd(starts[-1] + 5.0)
        wd(len(obstacles))
        for pos, vel in obstacles:
            wd(*pos, *vel)
        wd(n_steps, v_ref)
    r = subprocess.run([planner_exe, "gpu", str(scene), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.frombuffer(open(out, "rb").read(), dtype=np.float64)
    assert int(raw[0]) == n_steps
    rec = raw[1:].reshape(n_steps, 11)
    success, exit_flag, best, x, y, psi, v, spline, a, w, clearance = rec.T
    assert success.mean() >= 0.9, (success, exit_flag)
    assert success[0] == 1.0
    assert clearance.min() >= OBSTACLE_RADIUS + ROBOT_RADIUS - 0.05, clearance.min()
    assert spline[-1] - spline[0] >= 0.6 * v_ref * dt * n_steps  # progress along the path
    # contouring: the robot stays near the reference path except while passing an obstacle
    pos_ref, _ = _path_eval(coef, starts, spline)
    lateral = np.hypot(x - pos_ref[:, 0], y - pos_ref[:, 1])
    assert lateral.max() < 3.0 and np.median(lateral) < 0.8
    assert np.all(v >= -1e-6) and np.all(v <= 3.0 + 1e-6) and np.all(np.abs(w[success == 1]) <= 0.8 + 1e-6)
    assert set(best[success == 1].astype(int)) <= {0, 1, CONFIG_N_PATHS}


# ---- Safe Horizon MPC control flow (scenario_constraints.cpp:38-108): `parallel_solvers` local solvers as one batch ----
@pytest.fixture(scope="module")
def scenario_exe():
    return build_host.build_scenario_test("cfg4")


def test_scenario_module_packing_cpu(scenario_exe):
    r = subprocess.run([scenario_exe, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_batched_scenario_solvers_match_python_binding(scenario_exe, tmp_path):
    """Each of the 4 local solvers copies the main solver (parameters + initial guess), gets its own scenario halfspaces and
    is solved in the same launch; the lowest-cost feasible one is loaded into the main solver (:88-105)."""
    from mpc_planner_b200 import binding

    spec = pb.config_safe_horizon()
    P = 4
    b = scenes.make_scene_batch(spec, 1, P, seed=12)
    pm, N, npar = spec.parameter_map(), spec.N, spec.npar
    n_rows = 24
    cols = np.array([[pm[f"disc_0_scenario_constraint_{i}_{s}"] for s in ("a1", "a2", "b")] for i in range(n_rows)])
    par = b["params"].reshape(P, N, npar)
    # what the reference does: every local solver starts from the MAIN solver's parameters and initial guess
    xinit = np.tile(b["xinit"][0], (P, 1))
    x0 = np.tile(b["x0"][0], (P, 1))
    blocks = np.tile(par[0], (P, 1, 1))
    blocks[:, 1:, cols.ravel()] = par[:, 1:][:, :, cols.ravel()]
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("i", P))
        f.write(xinit[0].tobytes())
        f.write(x0[0].tobytes())
        f.write(par[0].tobytes())
        f.write(struct.pack("i", n_rows))
        for p in range(P):
            f.write(np.ascontiguousarray(par[p][:, cols]).tobytes())  # [k][row][a1, a2, b]
    r = subprocess.run([scenario_exe, "gpu", str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    solver = binding.BatchSolver(spec, P)
    xt, ut, info, best = solver.solve(xinit, x0, blocks.reshape(P, -1))
    raw = open(fout, "rb").read()
    cbest, cexit = struct.unpack_from("ii", raw, 0)
    off = 8
    nxt, nut = (N + 1) * spec.nx, N * spec.nu
    for p in range(P):
        (ec,) = struct.unpack_from("i", raw, off)
        off += 4
        (pobj,) = struct.unpack_from("d", raw, off)
        off += 8
        cx = np.frombuffer(raw, np.float64, nxt, off)
        off += 8 * nxt
        cu = np.frombuffer(raw, np.float64, nut, off)
        off += 8 * nut
        assert ec == info[p]["exit_code"]
        assert np.array_equal(cx, xt[p].ravel()) and np.array_equal(cu, ut[p].ravel()) and pobj == info[p]["pobj"]
    ok = [p for p in range(P) if info[p]["exit_code"] == 1]
    expect = min(ok, key=lambda p: (info[p]["pobj"], p)) if ok else -1
    assert cbest == expect == best
    if ok:
        main_x = np.frombuffer(raw, np.float64, nxt, off)
        assert cexit == 1 and np.array_equal(main_x, xt[expect].ravel())
    assert len(ok) >= 1
    solver.close()


@pytest.mark.parametrize("name", ["cfg2_gaussian", "cfg3"])
def test_gaussian_and_decomp_packing_follow_the_reference_cpu(name):
    """gaussian_constraints.cpp:31-79 (prediction step k-1, static obstacles without uncertainty, k = 0 dummies) and
    decomp_constraints.cpp:86-114,150-189 (polyhedron k -> stage k+1, rows end at a zero / NaN normal, dummies after)."""
    exe = build_host.build_packing_test(name)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
