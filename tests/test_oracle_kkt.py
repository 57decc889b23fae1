"""Oracle hardening: checks that share NO solver code with the oracle's interior-point method, on
every shipped module stack and several seeds.

1. Every QP of a solve.  For each of the 10 RTI steps the QP is assembled independently in numpy (exact Lagrangian Hessian
   from the multipliers the oracle holds at that step -- read back through the persistent-multiplier state --, MIRROR by
   `eigh`, linearised ERK4 dynamics, boxes, linearised h) and the oracle's step is checked against the QP's optimality
   conditions: primal feasibility, and multipliers (equality: free, active inequalities: >= 0) found by NON-NEGATIVE LEAST
   SQUARES that make the stationarity residual vanish.  The QP is strictly convex (MIRROR), so these conditions are
   sufficient for the global minimum: "kernel == oracle" no longer rests on two copies of one IPM.
2. NLP level.  The converged point (same algorithm, 40 steps) satisfies the KKT conditions of the NLP with multipliers
   found the same way (scipy NNLS), for cfg1, cfg2, cfg2_gaussian, cfg3, cfg4."""
import dataclasses

import numpy as np
import pytest
from scipy.optimize import nnls

from mpc_planner_b200 import problem as pb
from mpc_planner_b200 import scenes
from oracle import oracle

CONFIGS = ["cfg1", "cfg2", "cfg2_gaussian", "cfg3", "cfg4"]
SEEDS = [31, 32, 33, 34, 35]


def bounds(spec):
    d = spec.desc()
    nz = spec.nz
    lh, uh = map(np.array, spec.constraint_bounds())
    return np.array(list(d.lower_bound)[:nz]), np.array(list(d.upper_bound)[:nz]), lh, uh


def inequality_rows(spec, k, zk, ev, lo, up, lh, uh):
    """rows (value >= 0, gradient wrt z_k) of node k: boxes (inputs only at node 0) and h (nodes >= 1)"""
    nz, nu = spec.nz, spec.nu
    vals, grads = [], []
    for i in (range(nz) if k >= 1 else range(nu)):
        e = np.zeros(nz)
        e[i] = 1.0
        if lo[i] > -1e10:
            vals.append(zk[i] - lo[i]); grads.append(e)
        if up[i] < 1e10:
            vals.append(up[i] - zk[i]); grads.append(-e)
    if k >= 1:
        for j in range(spec.nh):
            if np.isfinite(lh[j]) and lh[j] > -1e14:
                vals.append(ev["h"][j] - lh[j]); grads.append(ev["h_jac"][j])
            if np.isfinite(uh[j]) and uh[j] < 1e14:
                vals.append(uh[j] - ev["h"][j]); grads.append(-ev["h_jac"][j])
    return np.array(vals), np.array(grads).reshape(len(vals), nz)


def kkt_residuals(grad, J_eq, J_in, slack):
    """Multipliers (nu free, mu >= 0 on ALL inequality rows) that minimise  |grad + J_eq' nu - J_in' mu|^2 + sum_j (mu_j
    slack_j)^2  (non-negative least squares: stationarity and complementarity in one problem -- an interior-point solution
    leaves every row a multiplier ~ tol / slack, so a hard active-set threshold would throw real multipliers away).
    Returns (stationarity residual, complementarity residual), both inf-norms."""
    n, m = grad.size, slack.size
    A = np.zeros((n + m, 2 * J_eq.shape[0] + m))
    A[:n, :J_eq.shape[0]] = J_eq.T
    A[:n, J_eq.shape[0]:2 * J_eq.shape[0]] = -J_eq.T
    A[:n, 2 * J_eq.shape[0]:] = -J_in.T
    A[n:, 2 * J_eq.shape[0]:] = np.diag(np.maximum(slack, 0.0))
    rhs = np.concatenate([-grad, np.zeros(m)])
    x, _ = nnls(A, rhs, maxiter=20 * A.shape[1])
    r = A @ x - rhs
    return np.abs(r[:n]).max(), (np.abs(r[n:]).max() if m else 0.0)


def qp_of_step(spec, xinit, params, z, xN, pi, lam_h):
    """independent assembly of the QP the SQP-RTI step solves at iterate (z, xN) with multipliers (pi, lam_h)"""
    d = spec.desc()
    N, nx, nu, nz, npar, nh = spec.N, spec.nx, spec.nu, spec.nz, spec.npar, spec.nh
    lo, up, lh, uh = bounds(spec)
    par = params.reshape(N, npar)
    w_h = np.where(lh > -1e14, -lam_h, lam_h)  # L += lam_hi (h - uh) + lam_lo (lh - h)
    ev = [oracle.eval_stage(d, z[k], par[k], w_h[k] if k >= 1 else np.zeros(max(nh, 1)), pi[k + 1]) for k in range(N)]
    nvar = N * nz + nx
    H = np.zeros((nvar, nvar))
    g = np.zeros(nvar)
    indefinite = 0
    for k in range(N):
        Hk = spec.dt * ev[k]["cost_hess"] + ev[k]["dyn_hess"] + (ev[k]["h_hess"] if k >= 1 else 0.0)
        Hk = 0.5 * (Hk + Hk.T)
        wv, V = np.linalg.eigh(Hk)
        indefinite += int(wv.min() < -spec.reg_epsilon)
        wm = np.where(np.abs(wv) <= spec.reg_epsilon, spec.reg_epsilon, np.abs(wv))
        H[k * nz:(k + 1) * nz, k * nz:(k + 1) * nz] = (V * wm) @ V.T
        g[k * nz:(k + 1) * nz] = spec.dt * ev[k]["cost_grad"]
    H[N * nz:, N * nz:] = spec.reg_epsilon * np.eye(nx)
    # equalities: x_0 = xinit;  Phi(z_k) + [B A] dz_k - x_{k+1} - dx_{k+1} = 0
    Aeq = np.zeros(((N + 1) * nx, nvar))
    beq = np.zeros((N + 1) * nx)
    Aeq[:nx, nu:nz] = np.eye(nx)
    beq[:nx] = z[0, nu:] - xinit
    for k in range(N):
        r = slice((k + 1) * nx, (k + 2) * nx)
        Aeq[r, k * nz:(k + 1) * nz] = ev[k]["AB"]
        if k + 1 < N:
            Aeq[r, (k + 1) * nz + nu:(k + 2) * nz] -= np.eye(nx)
            beq[r] = ev[k]["xnext"] - z[k + 1, nu:]
        else:
            Aeq[r, N * nz:] -= np.eye(nx)
            beq[r] = ev[k]["xnext"] - xN
    vals, rows = [], []
    for k in range(N):
        v, G = inequality_rows(spec, k, z[k], ev[k], lo, up, lh, uh)
        R = np.zeros((len(v), nvar))
        R[:, k * nz:(k + 1) * nz] = G
        vals.append(v); rows.append(R)
    return H, g, Aeq, beq, np.concatenate(vals), np.concatenate(rows), indefinite


@pytest.mark.parametrize("name", CONFIGS)
def test_every_qp_of_a_solve_is_solved_to_its_optimality_conditions(name):
    spec = pb.CONFIGS[name]()
    N, nx, nu, nz, nh = spec.N, spec.nx, spec.nu, spec.nz, spec.nh
    checked = indefinite_blocks = 0
    worst = 0.0
    for seed in SEEDS[:2] if name in ("cfg3",) else SEEDS[:3]:
        b = scenes.make_scene_batch(spec, 1, 2, seed=seed, gaussian_obstacles=(name == "cfg2_gaussian"))
        for i in range(2):
            xi, x0, par = b["xinit"][i:i + 1], b["x0"][i:i + 1], b["params"][i:i + 1]
            full = oracle.solve_batch(spec, xi, x0, par, iterates=True)
            if full[2][0]["exit_code"] != 1:
                continue
            its = full[3][0]  # [n_sqp, N+1, nz] iterates after each step
            for k_done in range(spec.n_sqp):
                if k_done == 0:
                    z = x0[0].reshape(N + 1, nz).copy()
                    zN = z[N, nu:].copy()
                    pi, lam = np.zeros((N + 1, nx)), np.zeros((N, max(nh, 1)))
                else:
                    st = np.zeros((1, oracle.state_size(spec)))
                    xt_k, ut_k, inf_k = oracle.solve_batch(dataclasses.replace(spec, n_sqp=k_done, res_eq_fail=1e9), xi, x0, par, state=st)
                    assert inf_k[0]["exit_code"] == 1 or inf_k[0]["qp_status"] == 0
                    if not st.any():
                        continue  # demoted intermediate solve: the state was cleared
                    z = np.zeros((N + 1, nz))
                    z[:N, :nu], z[:N, nu:] = ut_k[0], xt_k[0][:N]
                    zN = xt_k[0][N]
                    pi = st[0, :(N + 1) * nx].reshape(N + 1, nx)
                    lam = st[0, (N + 1) * nx:].reshape(N, nh) if nh else np.zeros((N, 1))
                H, g, Aeq, beq, cval, C, ind = qp_of_step(spec, xi[0], par[0], z, zN, pi, lam)
                indefinite_blocks += ind
                step = np.concatenate([(its[k_done, :N] - z[:N]).ravel(), its[k_done, N, :nx] - zN])
                # primal feasibility of the oracle's step
                assert np.abs(Aeq @ step + beq).max() < 1e-7
                slack = cval + C @ step
                assert slack.min() > -1e-6
                # optimality: multipliers by NNLS (stationarity + complementarity); the IPM stops at 1e-5 per row in both
                near = slack < 1.0  # rows a unit away from their bound cannot carry a multiplier above qp_tol
                res, comp = kkt_residuals(H @ step + g, Aeq, C[near], slack[near])
                worst = max(worst, res, comp)
                assert res < 1e-4 and comp < 1e-4, (name, seed, i, k_done, res, comp)
                checked += 1
    assert checked >= 20
    print(f"{name}: {checked} QPs checked, worst stationarity residual with NNLS multipliers {worst:.2e}, indefinite stage blocks mirrored {indefinite_blocks}")


@pytest.mark.parametrize("name", CONFIGS)
def test_converged_points_satisfy_the_kkt_conditions_of_the_nlp(name):
    spec = dataclasses.replace(pb.CONFIGS[name](), n_sqp=40)
    d = spec.desc()
    N, nx, nu, nz, npar, nh = spec.N, spec.nx, spec.nu, spec.nz, spec.npar, spec.nh
    lo, up, lh, uh = bounds(spec)
    checked = 0
    worst = 0.0
    for seed in SEEDS:
        b = scenes.make_scene_batch(spec, 1, 2, seed=seed, gaussian_obstacles=(name == "cfg2_gaussian"))
        xt, ut, info = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
        xt1, _, info1 = oracle.solve_batch(dataclasses.replace(spec, n_sqp=41), b["xinit"], b["x0"], b["params"])
        for i in range(2):
            if info[i]["exit_code"] != 1 or info1[i]["exit_code"] != 1 or info[i]["qp_status"] != 0 or info[i]["sqp_iter"] != spec.n_sqp:
                continue  # (a QP that hits its iteration limit ends the RTI loop early with exit code 1, as in acados)
            if np.abs(xt1[i] - xt[i]).max() > 1e-6:
                continue  # the 41st step still moves the iterate: not a converged point (MIRROR'd SQP converges linearly)
            par = b["params"][i].reshape(N, npar)
            z = np.concatenate([ut[i], xt[i, :N]], axis=1)
            ev = [oracle.eval_stage(d, z[k], par[k]) for k in range(N)]
            nvar = N * nz + nx
            grad = np.zeros(nvar)
            Jeq = np.zeros(((N + 1) * nx, nvar))
            Jeq[:nx, nu:nz] = np.eye(nx)
            defect = [z[0, nu:] - b["xinit"][i]]
            vals, rows = [], []
            for k in range(N):
                grad[k * nz:(k + 1) * nz] = spec.dt * ev[k]["cost_grad"]
                r = slice((k + 1) * nx, (k + 2) * nx)
                Jeq[r, k * nz:(k + 1) * nz] = ev[k]["AB"]
                if k + 1 < N:
                    Jeq[r, (k + 1) * nz + nu:(k + 2) * nz] -= np.eye(nx)
                    defect.append(ev[k]["xnext"] - z[k + 1, nu:])
                else:
                    Jeq[r, N * nz:] -= np.eye(nx)
                    defect.append(ev[k]["xnext"] - xt[i, N])
                v, G = inequality_rows(spec, k, z[k], ev[k], lo, up, lh, uh)
                R = np.zeros((len(v), nvar))
                R[:, k * nz:(k + 1) * nz] = G
                vals.append(v); rows.append(R)
            vals, rows = np.concatenate(vals), np.concatenate(rows)
            assert np.abs(np.concatenate(defect)).max() < 1e-4 and vals.min() > -1e-5  # feasible (linear convergence: 40 steps)
            near = vals < 1.0
            res, comp = kkt_residuals(grad, Jeq, rows[near], vals[near])
            worst = max(worst, res, comp)
            # 40 steps of the MIRROR-regularised SQP converge linearly: stationarity to ~1e-4 of the gradient scale
            assert res < 5e-4 * max(1.0, np.abs(grad).max()) and comp < 1e-4, (name, seed, i, res, comp)
            checked += 1
    assert checked >= 4
    print(f"{name}: {checked} converged points, worst KKT stationarity residual {worst:.2e}")
