"""Synthetic obstacle / reference-path scenes at the *parameter level* (SURVEY.md 8d).

The reference's guesses, halfspaces and splines come from un-vendored packages (guidance_planner,
DecompUtil, scenario_module, ros_tools), so benchmark inputs are synthesised directly in the
layout the solver consumes (`AcadosParameters`, acados_solver_interface.h:51-56), following the
packing semantics of the reference's C++ modules:

* weights: settings.yaml:75-88 via MPCBaseModule::setParameters (mpc_base.cpp:23-35) and
  Contouring::setParameters (contouring.cpp:50-92); the same spline coefficients at every stage
  (contouring.cpp:94-124);
* stage 0 dummies: obstacles at (x0+50, y0+50), r=0.1 (ellipsoid_constraints.cpp:30-31,42-55);
  halfspace (1, 0, x0+100) (linearized_constraints.cpp:53,155-166);
* stage k uses the obstacle prediction step k-1, i.e. the position at time k*dt
  (ellipsoid_constraints.cpp:67-70);
* guided guesses: warm start x,y,psi,v for k=1..N-1 from a guidance trajectory
  (guidance_constraints.cpp:390-414) and topology halfspaces a=(o-p)/|o-p|,
  b=a.o-(1e-3+robot_radius) (linearized_constraints.cpp:85-105); guess 0 is the non-guided
  planner with all-dummy halfspaces (guidance_constraints.cpp:301-305,323-324).

Everything is seeded numpy (FP64) so the CPU oracle and the GPU consume identical bytes.
This is input preparation for tests and benchmarks; it contains no solver arithmetic.
"""
from __future__ import annotations

import numpy as np

from . import problem as pb

WEIGHTS = dict(  # mpc_planner_jackalsimulator/config/settings.yaml:75-88
    acceleration=0.34, angular_velocity=0.85, velocity=0.55, reference_velocity=2.0, contour=0.05, lag=0.75,
    slack=10000.0, terminal_angle=100.0, terminal_contouring=10.0,
)
ROBOT_RADIUS = 0.325  # settings.yaml:35
OBSTACLE_RADIUS = 0.4  # settings.yaml:40
SEGMENT_LENGTH = 5.0


def _fit_path(amplitude, n_segments, phase=0.0):
    """Arc-length cubic spline through y = A sin(2 pi x / 40 + phase); returns per-segment
    coefficients [n_seg, 2, 4] (a,b,c,d per axis, local coordinate s - start) and starts."""
    from scipy.interpolate import CubicSpline

    xs = np.linspace(-2.0, SEGMENT_LENGTH * n_segments + 12.0, 4000)
    ys = amplitude * (np.sin(2 * np.pi * xs / 40.0 + phase) - np.sin(phase))
    ds = np.hypot(np.diff(xs), np.diff(ys))
    s = np.concatenate([[0.0], np.cumsum(ds)])
    i0 = np.searchsorted(xs, 0.0)
    s = s - s[i0]
    knots = SEGMENT_LENGTH * np.arange(n_segments + 1)
    kx, ky = np.interp(knots, s, xs), np.interp(knots, s, ys)
    csx, csy = CubicSpline(knots, kx, bc_type="natural"), CubicSpline(knots, ky, bc_type="natural")
    coef = np.zeros((n_segments, 2, 4))
    coef[:, 0, :] = csx.c.T  # PPoly: c[0] s^3 + c[1] s^2 + c[2] s + c[3] in local coordinates
    coef[:, 1, :] = csy.c.T
    return coef, knots[:-1].copy()


def _path_eval(coef, starts, s):
    """Piecewise evaluation (hard segment switch; only used to build guesses)."""
    s = np.asarray(s, float)
    i = np.clip(np.searchsorted(starts, s, side="right") - 1, 0, len(starts) - 1)
    t = s - starts[i]
    c = coef[i]
    pos = c[..., 0] * t[..., None] ** 3 + c[..., 1] * t[..., None] ** 2 + c[..., 2] * t[..., None] + c[..., 3]
    der = 3 * c[..., 0] * t[..., None] ** 2 + 2 * c[..., 1] * t[..., None] + c[..., 2]
    return pos, der


def _unicycle_rk4(x, u, dt, steps=3):
    def f(x):
        return np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), u[1], u[0], x[3]])

    h = dt / steps
    for _ in range(steps):
        k1 = f(x)
        k2 = f(x + 0.5 * h * k1)
        k3 = f(x + 0.5 * h * k2)
        k4 = f(x + h * k3)
        x = x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def _rollout(spec, iz, state, pos, dt, N):
    """Track the polyline `pos` (one point per stage) with bounded a, w; returns [N+1, nz]."""
    nz, nu = spec.nz, spec.nu
    traj = np.zeros((N + 1, nz))
    x = np.array([state["x"], state["y"], state["psi"], state["v"], state["spline"]])
    for k in range(N + 1):
        traj[k, iz["x"]], traj[k, iz["y"]], traj[k, iz["psi"]], traj[k, iz["v"]], traj[k, iz["spline"]] = x
        if "slack" in iz and iz["slack"] >= nu:
            traj[k, iz["slack"]] = 0.0
        if k == N:
            break
        tgt = pos[min(k + 2, N)]
        nxt = pos[min(k + 1, N)]
        dvec = tgt - x[:2]
        heading = np.arctan2(dvec[1], dvec[0]) if np.linalg.norm(dvec) > 1e-6 else x[2]
        err = (heading - x[2] + np.pi) % (2 * np.pi) - np.pi
        w = float(np.clip(1.5 * err / (2 * dt), -0.75, 0.75))
        v_des = float(np.clip(np.linalg.norm(nxt - x[:2]) / dt, 0.0, 2.8))
        a = float(np.clip((v_des - x[3]) / dt, -1.9, 1.9))
        if x[3] + a * dt < 0.0:
            a = -x[3] / dt
        traj[k, iz["a"]], traj[k, iz["w"]] = a, w
        x = _unicycle_rk4(x, (a, w), dt)
    return traj


def make_scene_batch(spec: pb.ProblemSpec, n_scenes: int, n_guesses: int, seed: int = 0, gaussian_obstacles=False):
    """Returns dict(xinit[S*P,nx], x0[S*P,nz*(N+1)], params[S*P,npar*N], scene[S*P], guess[S*P])."""
    rng = np.random.default_rng(seed)
    N, nx, nu, nz, npar, dt = spec.N, spec.nx, spec.nu, spec.nz, spec.npar, spec.dt
    pm = spec.parameter_map()
    mm = spec.model_map()
    iz = {k: v[1] for k, v in mm.items()}
    kinds = dict(spec.constraints)
    n_lin = kinds.get(pb.CON_LINEAR, 0)
    n_ell = kinds.get(pb.CON_ELLIPSOID, 0)
    n_gau = kinds.get(pb.CON_GAUSSIAN, 0)
    n_dec = kinds.get(pb.CON_DECOMP, 0)
    n_sce = kinds.get(pb.CON_SCENARIO, 0)
    n_obs = max(n_ell, n_gau, n_lin)
    if n_sce and n_obs == 0:
        n_obs = 4

    S, P = n_scenes, n_guesses
    xinit = np.zeros((S, P, nx))
    x0 = np.zeros((S, P, N + 1, nz))
    params = np.zeros((S, P, N, npar))

    for si in range(S):
        # ---- reference path and initial state ----
        A = rng.uniform(0.0, 2.0)
        coef, starts = _fit_path(A, spec.n_segments, phase=rng.uniform(0, 2 * np.pi))
        s_init = rng.uniform(0.5, 1.5)
        p0, d0 = _path_eval(coef, starts, s_init)
        heading = np.arctan2(d0[1], d0[0])
        normal = np.array([-np.sin(heading), np.cos(heading)])
        lateral = rng.uniform(-0.3, 0.3)
        pos0 = p0 + lateral * normal
        psi0 = heading + rng.uniform(-0.1, 0.1)
        v0 = rng.uniform(0.5, 2.5)
        state = dict(x=pos0[0], y=pos0[1], psi=psi0, v=v0, spline=s_init, slack=0.0)

        # ---- obstacles: constant velocity, positions at t = k*dt for stage k ----
        t = dt * np.arange(N + 1)
        v_ref = WEIGHTS["reference_velocity"]
        base_s = s_init + np.cumsum(np.concatenate([[0.0], np.full(N, 0.0)]))  # filled below
        speed = np.minimum(v_ref, v0 + 1.0 * t)  # ramp towards the reference velocity
        speed = np.where(v0 > v_ref, np.maximum(v_ref, v0 - 1.0 * t), speed)
        base_s = s_init + np.concatenate([[0.0], np.cumsum(0.5 * (speed[1:] + speed[:-1]) * dt)])
        base_pos, base_der = _path_eval(coef, starts, base_s)
        base_head = np.arctan2(base_der[:, 1], base_der[:, 0])
        base_nrm = np.stack([-np.sin(base_head), np.cos(base_head)], axis=1)

        obs_pos = np.zeros((n_obs, N + 1, 2))
        for o in range(n_obs):
            for _ in range(100):
                along, across = rng.uniform(2.0, 14.0), rng.uniform(-3.0, 3.0)
                q, dq = _path_eval(coef, starts, s_init + along)
                hq = np.arctan2(dq[1], dq[0])
                start = q + across * np.array([-np.sin(hq), np.cos(hq)])
                sp = rng.uniform(0.5, 1.5)
                # head towards / across the path
                target_dir = -np.sign(across if across != 0 else 1.0) * np.array([-np.sin(hq), np.cos(hq)])
                mix = rng.uniform(-1.0, 0.3)
                direction = target_dir + mix * np.array([np.cos(hq), np.sin(hq)])
                direction /= np.linalg.norm(direction)
                traj = start[None, :] + (sp * t)[:, None] * direction[None, :]
                # not on top of the robot now, and not across its nominal motion during the first 1.6 s (an obstacle
                # that close cannot be avoided with the bounded inputs: every guess would be infeasible)
                early = np.linalg.norm(traj[:9] - base_pos[:9], axis=1).min()
                if np.linalg.norm(traj[0] - pos0) > 1.8 and np.linalg.norm(traj[1] - pos0) > 1.5 and early > 1.5:
                    obs_pos[o] = traj
                    break
            else:
                obs_pos[o] = start[None, :] + 100.0
        sigma = np.sqrt(np.cumsum(np.full(N + 1, (0.3 * dt) ** 2)))  # data_preparation.cpp:170-186
        chi = 5.991464547107979 if gaussian_obstacles else 1.0  # ExponentialQuantile(0.5, 0.95) = -2 ln(0.05)

        # ---- shared (scene-level) parameters ----
        shared = np.zeros((N, npar))
        for k in range(N):
            for name in ("acceleration", "angular_velocity", "velocity", "reference_velocity", "contour", "lag", "slack",
                         "terminal_angle", "terminal_contouring"):
                if name in pm:
                    shared[k, pm[name]] = WEIGHTS[name]
            for i in range(spec.n_segments):
                for ax, axn in enumerate("xy"):
                    for ci, cn in enumerate("abcd"):
                        shared[k, pm[f"spline_{axn}{i}_{cn}"]] = coef[i, ax, ci]
                shared[k, pm[f"spline{i}_start"]] = starts[i]
            if "ego_disc_radius" in pm:
                shared[k, pm["ego_disc_radius"]] = ROBOT_RADIUS
            if "ego_disc_0_offset" in pm:
                shared[k, pm["ego_disc_0_offset"]] = 0.0
            for o in range(n_ell):
                pre = f"ellipsoid_obst_{o}_"
                if k == 0:
                    vals = dict(x=pos0[0] + 50, y=pos0[1] + 50, psi=0.0, r=0.1, major=0.0, minor=0.0, chi=1.0)
                else:
                    mj = sigma[k - 1] if gaussian_obstacles else 0.0
                    vals = dict(x=obs_pos[o, k, 0], y=obs_pos[o, k, 1], psi=0.0, r=OBSTACLE_RADIUS, major=mj, minor=mj, chi=chi)
                for kk, vv in vals.items():
                    shared[k, pm[pre + kk]] = vv
            for o in range(n_gau):  # gaussian_constraints.cpp:31-79
                pre = f"gaussian_obst_{o}_"
                if k == 0:
                    vals = dict(x=pos0[0] + 100, y=pos0[1] + 100, major=0.1, minor=0.1, risk=0.05, r=0.1)  # (:23-29, 39-52)
                else:
                    vals = dict(x=obs_pos[o, k, 0], y=obs_pos[o, k, 1], major=max(sigma[k - 1], 1e-3), minor=max(sigma[k - 1], 1e-3),
                                risk=0.05, r=OBSTACLE_RADIUS)
                for kk, vv in vals.items():
                    shared[k, pm[pre + kk]] = vv
            dummy = (1.0, 0.0, pos0[0] + 100.0)
            for i in range(n_lin):
                for sfx, vv in zip(("a1", "a2", "b"), dummy):
                    shared[k, pm[f"lin_constraint_{i}_{sfx}"]] = vv
            for i in range(n_dec):
                for sfx, vv in zip(("a1", "a2", "b"), dummy):
                    shared[k, pm[f"disc_0_decomp_{i}_{sfx}"]] = vv
            for i in range(n_sce):
                for sfx, vv in zip(("a1", "a2", "b"), dummy):
                    shared[k, pm[f"disc_0_scenario_constraint_{i}_{sfx}"]] = vv
        # decomp corridor (cfg 3): convex polygon around the predicted path point, stage k >= 1
        if n_dec:
            n_sides = min(8, n_dec)
            rot = rng.uniform(0, 2 * np.pi)
            widths = rng.uniform(3.5, 5.0, n_sides)  # corridor wide enough to contain every guess (max lateral offset 3 m)
            for k in range(1, N):
                for i in range(n_sides):
                    th = rot + 2 * np.pi * i / n_sides
                    a = np.array([np.cos(th), np.sin(th)])
                    # long along the path, `widths` across: scale by alignment with the heading
                    align = abs(a @ np.array([np.cos(base_head[k]), np.sin(base_head[k])]))
                    b = a @ base_pos[k] + widths[i] * (1 + 4 * align)
                    shared[k, pm[f"disc_0_decomp_{i}_a1"]] = a[0]
                    shared[k, pm[f"disc_0_decomp_{i}_a2"]] = a[1]
                    shared[k, pm[f"disc_0_decomp_{i}_b"]] = b

        # ---- relevance of obstacles to the base trajectory ----
        d_obs = np.linalg.norm(obs_pos - base_pos[None], axis=2)  # [n_obs, N+1]
        k_star = d_obs.argmin(axis=1)
        d_star = d_obs.min(axis=1)
        relevant = [o for o in np.argsort(d_star) if d_star[o] < 2.5]

        for gi in range(P):
            par = shared.copy()
            traj = np.zeros((N + 1, nz))
            guided = gi > 0 and n_obs > 0
            best_clear, best_traj = -np.inf, None
            for attempt in range(12):
                bits = max(gi - 1, 0) if attempt == 0 else int(rng.integers(0, 2 ** max(len(relevant), 1)))
                amp_jitter = 1.0 if (attempt == 0 and gi - 1 < 2 ** max(len(relevant), 1)) else rng.uniform(0.8, 1.5)
                bump = 1.2 if attempt == 0 else rng.uniform(1.0, 2.0)
                lat = lateral * np.exp(-t / 1.0)
                for j, o in enumerate(relevant):
                    width = (OBSTACLE_RADIUS + ROBOT_RADIUS + 0.6) * amp_jitter
                    rel = (obs_pos[o, k_star[o]] - base_pos[k_star[o]]) @ base_nrm[k_star[o]]
                    if guided:  # pass on the side given by the guess's sign pattern
                        side = 1.0 if (bits >> j) & 1 else -1.0
                    else:  # non-guided planner: the side needing the smaller detour
                        side = -1.0 if rel >= 0 else 1.0
                    offset = rel + side * width
                    if (not guided) and abs(rel) > width:
                        offset = 0.0  # the path itself clears this obstacle
                    offset = float(np.clip(offset, -3.0, 3.0))
                    lat = lat + offset * np.exp(-0.5 * ((t - t[k_star[o]]) / bump) ** 2) * (t / max(t[k_star[o]], dt)).clip(0, 1)
                pos = base_pos + lat[:, None] * base_nrm
                # keep the guess collision-free (guidance trajectories are; stand-in for the
                # Douglas-Rachford projection of linearized_constraints.cpp:130-148, ros_tools un-vendored)
                R = 1e-3 + ROBOT_RADIUS
                R_safe = OBSTACLE_RADIUS + ROBOT_RADIUS + 0.25
                for _ in range(5):
                    for o in range(n_obs):
                        for k in range(1, N + 1):
                            dvec = pos[k] - obs_pos[o, k]
                            dist = np.linalg.norm(dvec)
                            if dist < R_safe:
                                pos[k] = obs_pos[o, k] + R_safe * (dvec / dist if dist > 1e-9 else base_nrm[k])
                # dynamically consistent warm start: roll the unicycle out along the guess with a bounded
                # tracking controller (guidance trajectories of the reference are smooth and feasible)
                traj = _rollout(spec, iz, state, pos, dt, N)
                clear = min(
                    (np.linalg.norm(traj[1:, [iz["x"], iz["y"]]] - obs_pos[o, 1:], axis=1).min() for o in range(n_obs)),
                    default=np.inf,
                )
                if clear > best_clear:
                    best_clear, best_traj = clear, traj
                if clear >= OBSTACLE_RADIUS + ROBOT_RADIUS - 0.12:  # warm start (nearly) collision-free
                    break
            traj = best_traj  # the clearest attempt if none was collision-free
            if guided and n_lin:
                for k in range(1, N):
                    for o in range(min(n_obs, n_lin)):
                        dvec = obs_pos[o, k] - traj[k, [iz["x"], iz["y"]]]
                        if np.linalg.norm(dvec) < 1e-9:  # projectToSafety would move the point; any direction will do
                            dvec = np.array([1e-9, 0.0])
                        a = dvec / np.linalg.norm(dvec)
                        b = a @ obs_pos[o, k] - R
                        par[k, pm[f"lin_constraint_{o}_a1"]] = a[0]
                        par[k, pm[f"lin_constraint_{o}_a2"]] = a[1]
                        par[k, pm[f"lin_constraint_{o}_b"]] = b
            if n_sce:
                # SH-MPC stand-in: samples of a 2-mode mixture per obstacle; keep the most constraining rows
                srng = np.random.default_rng(seed * 7919 + si * 131 + gi)
                for k in range(1, N):
                    pk = traj[k, [iz["x"], iz["y"]]]
                    rows = []
                    for o in range(n_obs):
                        mode = srng.random(64) < 0.6
                        spread = 0.05 + 0.05 * t[k]
                        turn = np.where(mode, 0.0, np.deg2rad(30.0) * np.sign(srng.standard_normal(64)))
                        rel = obs_pos[o, k] - obs_pos[o, 0]
                        c, s_ = np.cos(turn), np.sin(turn)
                        samp = obs_pos[o, 0] + np.stack([c * rel[0] - s_ * rel[1], s_ * rel[0] + c * rel[1]], 1)
                        samp = samp + spread * srng.standard_normal((64, 2))
                        dv = samp - pk
                        dist = np.linalg.norm(dv, axis=1)
                        a = dv / dist[:, None]
                        b = np.einsum("ij,ij->i", a, samp) - (OBSTACLE_RADIUS + ROBOT_RADIUS)
                        margin = b - a @ pk  # > 0: the guess satisfies the halfspace
                        keep_mask = margin > 0.05  # scenarios the guess already violates are discarded (removed as support)
                        rows.append(np.column_stack([a, b, margin])[keep_mask])
                    rows = np.concatenate(rows) if rows else np.zeros((0, 4))
                    keep = rows[np.argsort(rows[:, 3])[:n_sce]]
                    for i, r in enumerate(keep):
                        par[k, pm[f"disc_0_scenario_constraint_{i}_a1"]] = r[0]
                        par[k, pm[f"disc_0_scenario_constraint_{i}_a2"]] = r[1]
                        par[k, pm[f"disc_0_scenario_constraint_{i}_b"]] = r[2]
            xinit[si, gi] = [state[nme] for nme in spec.states]
            x0[si, gi] = traj
            params[si, gi] = par

    SP = S * P
    return dict(
        xinit=xinit.reshape(SP, nx), x0=x0.reshape(SP, (N + 1) * nz), params=params.reshape(SP, N * npar),
        scene=np.repeat(np.arange(S), P), guess=np.tile(np.arange(P), S), n_scenes=S, n_guesses=P,
    )


def _chunk_job(args):
    spec, n, n_guesses, seed, gaussian = args
    return make_scene_batch(spec, n, n_guesses, seed=seed, gaussian_obstacles=gaussian)


def make_scene_batch_chunked(spec: pb.ProblemSpec, n_scenes: int, n_guesses: int, seed: int = 0, gaussian_obstacles=False, chunk=16,
                             workers=None):
    """`n_scenes` DISTINCT scenes generated in chunks of `chunk` scenes, chunk c seeded with `seed * 100003 + c` -- the same
    bytes whatever the number of worker processes (generation is host-side numpy, ~70 ms per scene; the workers only make
    the benchmark start faster).  Any prefix of whole chunks of a larger batch equals the smaller batch."""
    import multiprocessing as mp
    import os

    jobs = []
    for c in range((n_scenes + chunk - 1) // chunk):
        n = min(chunk, n_scenes - c * chunk)
        jobs.append((spec, n, n_guesses, seed * 100003 + c, gaussian_obstacles))
    if workers is None:
        try:
            workers = len(os.sched_getaffinity(0))
        except Exception:
            workers = os.cpu_count() or 1
        workers = max(1, min(workers, 32, len(jobs)))
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_chunk_job, jobs)
    else:
        parts = [_chunk_job(j) for j in jobs]
    out = {}
    for k in ("xinit", "x0", "params"):
        out[k] = np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=0))
    return out


def compact_scene(spec: pb.ProblemSpec, batch, n_scenes: int, n_guesses: int):
    """The compact scene description `mpcgpu_solve_tmpc_scene_host` consumes, recovered from a full batch: what the
    modules' setParameters READ (obstacle predictions, radii, polyhedra rows, the stage-invariant weights / spline / disc
    parameters) instead of the N x npar parameter matrix they write.  Returns a dict with base_params [S, npar],
    obstacles [S, n_obs, N, 5], obstacle_radius [S, n_obs], obstacle_flags [S, n_obs], risk, halfspaces [S, N, n_hs, 3] or
    None, guided [G], xinit [S, nx]."""
    S, G, N, npar, nx = n_scenes, n_guesses, spec.N, spec.npar, spec.nx
    pm = spec.parameter_map()
    kinds = dict(spec.constraints)
    par = batch["params"].reshape(S, G, N, npar)[:, 0]  # guess 0 = the main solver's parameters of the control step
    base = par[:, 1].copy()
    n_ell, n_gau, n_dec = kinds.get(pb.CON_ELLIPSOID, 0), kinds.get(pb.CON_GAUSSIAN, 0), kinds.get(pb.CON_DECOMP, 0)
    n_obs = max(n_ell, n_gau)
    obstacles = np.zeros((S, n_obs, N, 5))
    radius = np.zeros((S, n_obs))
    flags = np.zeros((S, n_obs), dtype=np.int32)
    risk = 0.05
    for o in range(n_obs):
        if n_ell:
            for c, nm in enumerate(("x", "y", "psi", "major", "minor")):
                obstacles[:, o, : N - 1, c] = par[:, 1:, pm[f"ellipsoid_obst_{o}_{nm}"]]
            radius[:, o] = par[:, 1, pm[f"ellipsoid_obst_{o}_r"]]
            gauss = par[:, 1, pm[f"ellipsoid_obst_{o}_chi"]] != 1.0
            flags[:, o] = gauss.astype(np.int32)
            if gauss.any():
                risk = float(np.exp(-0.5 * par[gauss, 1, pm[f"ellipsoid_obst_{o}_chi"]][0]))
        else:
            obstacles[:, o, : N - 1, 0] = par[:, 1:, pm[f"gaussian_obst_{o}_x"]]
            obstacles[:, o, : N - 1, 1] = par[:, 1:, pm[f"gaussian_obst_{o}_y"]]
            obstacles[:, o, : N - 1, 3] = par[:, 1:, pm[f"gaussian_obst_{o}_major"]]
            obstacles[:, o, : N - 1, 4] = par[:, 1:, pm[f"gaussian_obst_{o}_minor"]]
            radius[:, o] = par[:, 1, pm[f"gaussian_obst_{o}_r"]]
            flags[:, o] = 1
            risk = float(par[0, 1, pm[f"gaussian_obst_{o}_risk"]])
        obstacles[:, o, N - 1] = obstacles[:, o, N - 2]  # the last prediction step feeds no stage
    halfspaces = None
    if n_dec:
        halfspaces = np.zeros((S, N, n_dec, 3))
        for i in range(n_dec):
            for c, nm in enumerate(("a1", "a2", "b")):
                halfspaces[:, :, i, c] = par[:, :, pm[f"disc_0_decomp_{i}_{nm}"]]
    guided = np.array([0] + [1] * (G - 1), dtype=np.int32)
    return dict(base_params=base, obstacles=obstacles, obstacle_radius=radius, obstacle_flags=flags, risk=risk, halfspaces=halfspaces,
                guided=guided, xinit=batch["xinit"].reshape(S, G, nx)[:, 0].copy())


def _bicycle_rk4(x, u, dt, steps=3):
    """BicycleModel2ndOrder.continuous_model (solver_model.py:319-351), states x, y, psi, v, delta, spline; ERK4 x 3."""
    lr = 2.79 / 2.0

    def f(x):
        beta = np.arctan(0.5 * np.tan(x[4]))
        return np.array([x[3] * np.cos(x[2] + beta), x[3] * np.sin(x[2] + beta), x[3] / lr * np.sin(beta), u[0], u[1], x[3]])

    h = dt / steps
    for _ in range(steps):
        k1 = f(x)
        k2 = f(x + 0.5 * h * k1)
        k3 = f(x + 0.5 * h * k2)
        k4 = f(x + h * k3)
        x = x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def make_bicycle_batch(spec: pb.ProblemSpec, n: int, seed: int = 0):
    """BASELINE configs[0]: bicycle model, MPCC, ellipsoidal obstacles, one solve per scene (P = 1).  Robot on the reference
    path with a small lateral / heading / steering offset, warm start = a steering-controlled bicycle roll-out towards the path,
    obstacles moving alongside the path at >= 2.5 m lateral distance from it (the model is a car: radius 0.325 m disc as in
    the jackal settings, to keep the parameter semantics of the T-MPC scenes)."""
    assert spec.model == pb.MODEL_BICYCLE
    rng = np.random.default_rng(seed)
    N, nx, nu, nz, npar, dt = spec.N, spec.nx, spec.nu, spec.nz, spec.npar, spec.dt
    pm = spec.parameter_map()
    iz = {k: v[1] for k, v in spec.model_map().items()}
    n_ell = dict(spec.constraints).get(pb.CON_ELLIPSOID, 0)
    xinit, x0, params = np.zeros((n, nx)), np.zeros((n, N + 1, nz)), np.zeros((n, N, npar))
    for si in range(n):
        coef, starts = _fit_path(rng.uniform(0.0, 1.5), spec.n_segments, phase=rng.uniform(0, 2 * np.pi))
        s0 = rng.uniform(0.5, 1.5)
        p0, d0 = _path_eval(coef, starts, s0)
        heading = np.arctan2(d0[1], d0[0])
        normal = np.array([-np.sin(heading), np.cos(heading)])
        pos0 = p0 + rng.uniform(-0.3, 0.3) * normal
        x = np.array([pos0[0], pos0[1], heading + rng.uniform(-0.1, 0.1), rng.uniform(1.0, 2.5), rng.uniform(-0.05, 0.05), s0])
        xinit[si] = x
        traj = np.zeros((N + 1, nz))
        for k in range(N + 1):
            traj[k, nu:] = x
            if k == N:
                break
            tgt, _ = _path_eval(coef, starts, x[5] + 2.0)
            err = (np.arctan2(tgt[1] - x[1], tgt[0] - x[0]) - x[2] + np.pi) % (2 * np.pi) - np.pi
            delta_des = float(np.clip(1.2 * err, -0.4, 0.4))
            w = float(np.clip((delta_des - x[4]) / dt, -1.2, 1.2))
            a = float(np.clip((2.0 - x[3]) / (4 * dt), -1.0, 1.0))
            traj[k, iz["a"]], traj[k, iz["w"]], traj[k, iz["slack"]] = a, w, 0.0
            x = _bicycle_rk4(x, (a, w), dt)
        x0[si] = traj
        par = np.zeros((N, npar))
        for name in ("acceleration", "angular_velocity", "velocity", "reference_velocity", "contour", "lag", "slack", "terminal_angle", "terminal_contouring"):
            if name in pm:
                par[:, pm[name]] = WEIGHTS[name]
        for i in range(spec.n_segments):
            for ax, axn in enumerate("xy"):
                for ci, cn in enumerate("abcd"):
                    par[:, pm[f"spline_{axn}{i}_{cn}"]] = coef[i, ax, ci]
            par[:, pm[f"spline{i}_start"]] = starts[i]
        par[:, pm["ego_disc_radius"]] = ROBOT_RADIUS
        par[:, pm["ego_disc_0_offset"]] = 0.0
        t = dt * np.arange(N + 1)
        for o in range(n_ell):
            so = rng.uniform(3.0, 18.0)
            po, do = _path_eval(coef, starts, so)
            ho = np.arctan2(do[1], do[0])
            side = 1.0 if rng.random() < 0.5 else -1.0
            start = po + side * rng.uniform(2.5, 4.0) * np.array([-np.sin(ho), np.cos(ho)])
            vel = rng.uniform(0.3, 1.0) * np.array([np.cos(ho), np.sin(ho)])
            pos = start[None, :] + t[:, None] * vel[None, :]
            pre = f"ellipsoid_obst_{o}_"
            for k in range(N):
                vals = (dict(x=pos0[0] + 50, y=pos0[1] + 50, psi=0.0, r=0.1, major=0.0, minor=0.0, chi=1.0) if k == 0 else
                        dict(x=pos[k, 0], y=pos[k, 1], psi=0.0, r=OBSTACLE_RADIUS, major=0.0, minor=0.0, chi=1.0))
                for kk, vv in vals.items():
                    par[k, pm[pre + kk]] = vv
        params[si] = par
    return dict(xinit=xinit, x0=x0.reshape(n, -1), params=params.reshape(n, -1))
