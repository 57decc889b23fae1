"""Solver-level checks of the CPU oracle (T2 of SURVEY.md 8c): the SQP-RTI restatement must land
on KKT points of the NLP, verified independently with scipy.optimize on the same functions."""
import numpy as np
import pytest

from mpc_planner_b200 import problem as pb
from mpc_planner_b200 import scenes
from oracle import oracle


def _nlp_pieces(spec, xinit, params):
    d = spec.desc()
    N, nx, nu, nz, npar, nh = spec.N, spec.nx, spec.nu, spec.nz, spec.npar, spec.nh
    par = params.reshape(N, npar)
    lh, uh = map(np.array, spec.constraint_bounds())
    lo, up = np.array(list(d.lower_bound)[:nz]), np.array(list(d.upper_bound)[:nz])
    nvar = N * nz + nx

    def unpack(w):
        return w[: N * nz].reshape(N, nz), w[N * nz:]

    def cost(w):
        z, _ = unpack(w)
        return sum(spec.dt * oracle.eval_stage(d, z[k], par[k])["cost"] for k in range(N))

    def cost_grad(w):
        z, _ = unpack(w)
        g = np.zeros(nvar)
        for k in range(N):
            g[k * nz:(k + 1) * nz] = spec.dt * oracle.eval_stage(d, z[k], par[k])["cost_grad"]
        return g

    def eq(w):
        z, xN = unpack(w)
        out = [z[0, nu:] - xinit]
        for k in range(N):
            xn = z[k + 1, nu:] if k + 1 < N else xN
            out.append(oracle.eval_stage(d, z[k], par[k])["xnext"] - xn)
        return np.concatenate(out)

    def ineq(w):  # >= 0
        z, _ = unpack(w)
        out = []
        for k in range(N):
            sl = slice(0, nz) if k >= 1 else slice(0, nu)
            out += [z[k, sl] - lo[sl], up[sl] - z[k, sl]]
            if k >= 1:
                h = oracle.eval_stage(d, z[k], par[k])["h"]
                out += [(h - lh)[np.isfinite(lh)], (uh - h)[np.isfinite(uh)]]
        return np.concatenate(out)

    return nvar, cost, cost_grad, eq, ineq


@pytest.mark.parametrize("seed", [11])
def test_converged_points_are_kkt_points(seed):
    from scipy.optimize import minimize

    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 2, 3, seed=seed)
    # MIRROR-regularised SQP converges linearly: 10 RTI steps (the reference's setting) stop
    # ~1e-4 short of the optimum in objective, so run the same algorithm to convergence here.
    spec.n_sqp = 40
    xt, ut, info = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    N, nx, nu, nz = spec.N, spec.nx, spec.nu, spec.nz
    checked = 0
    for i, inf in enumerate(info):
        if inf["exit_code"] != 1 or checked >= 2:
            continue
        nvar, cost, cost_grad, eq, ineq = _nlp_pieces(spec, b["xinit"][i], b["params"][i])
        w0 = np.concatenate([np.concatenate([ut[i], xt[i, :N]], axis=1).ravel(), xt[i, N]])
        assert cost(w0) == pytest.approx(inf["pobj"], rel=1e-12)
        assert np.abs(eq(w0)).max() < 1e-6  # dynamics satisfied
        assert ineq(w0).min() > -1e-6  # feasible
        res = minimize(cost, w0, jac=cost_grad, method="SLSQP", constraints=[dict(type="eq", fun=eq), dict(type="ineq", fun=ineq)],
                       options=dict(maxiter=60, ftol=1e-12))
        # an independent NLP solver started at the oracle's point cannot improve on it
        # the IPM stops at complementarity <= qp_tol = 1e-5 per row: objective gap <= n_active * 1e-5
        assert res.fun == pytest.approx(inf["pobj"], abs=5e-5)
        assert res.fun <= inf["pobj"] + 1e-9
        assert np.abs(res.x - w0).max() < 5e-3
        checked += 1
    assert checked >= 1


def test_exit_codes_and_shapes():
    spec = pb.config_basic()
    b = scenes.make_scene_batch(spec, 2, 2, seed=3)
    xt, ut, info = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"], nthreads=2)
    assert xt.shape == (4, spec.N + 1, spec.nx) and ut.shape == (4, spec.N, spec.nu)
    for i, inf in enumerate(info):
        assert inf["exit_code"] in (0, 1, 2, 3, 4)
        if inf["exit_code"] == 1:
            np.testing.assert_allclose(xt[i, 0], b["xinit"][i], atol=1e-12)  # x0 constraint
            assert inf["res_eq"] < 1e-2 and inf["sqp_iter"] == spec.n_sqp
    assert oracle.find_best(info) in range(-1, 4)


def test_mirror_regularisation():
    import ctypes

    rng = np.random.default_rng(0)
    for n in (5, 7, 8):
        A = rng.standard_normal((n, n))
        A = A + A.T
        w, V = np.linalg.eigh(A)
        eps = 1e-4
        wm = np.where(np.abs(w) <= eps, eps, np.abs(w))
        want = (V * wm) @ V.T
        got = np.ascontiguousarray(A.copy())
        oracle.lib().oracle_mirror(n, got.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_double(eps))
        np.testing.assert_allclose(got, want, atol=1e-12)


def test_first_rti_step_solves_the_qp_an_independent_solver_finds():
    """QP level (assembly + Mehrotra IPM + Riccati): with zero multipliers the first SQP-RTI step is
    z0 + argmin of the QP built from the stage linearisations at z0 (MIRROR-regularised dt-scaled cost Hessian,
    linearised ERK4 dynamics, boxes, linearised h).  The QP is convex, so scipy's SLSQP -- which shares no code with
    the oracle's interior-point method -- must return the same step up to the IPM's tolerance (qp_tol = 1e-5)."""
    import ctypes

    from scipy.optimize import minimize

    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 1, 3, seed=4)
    d = spec.desc()
    N, nx, nu, nz, npar = spec.N, spec.nx, spec.nu, spec.nz, spec.npar
    xt, ut, info, its = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"], iterates=True)
    lh, uh = map(np.array, spec.constraint_bounds())
    lo, up = np.array(list(d.lower_bound)[:nz]), np.array(list(d.upper_bound)[:nz])
    checked = 0
    for i in range(len(info)):
        if info[i]["exit_code"] != 1:
            continue
        par = b["params"][i].reshape(N, npar)
        z0 = b["x0"][i].reshape(N + 1, nz)
        xN0 = z0[N, nu:]  # warm start of node N: [u (unused); x]
        ev = [oracle.eval_stage(d, z0[k], par[k]) for k in range(N)]
        H = []
        for k in range(N):
            Hk = np.ascontiguousarray(spec.dt * ev[k]["cost_hess"])
            oracle.lib().oracle_mirror(nz, Hk.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_double(spec.reg_epsilon))
            H.append(Hk)
        g = [spec.dt * ev[k]["cost_grad"] for k in range(N)]
        nvar = N * nz + nx

        def split(w):
            return w[: N * nz].reshape(N, nz), w[N * nz:]

        def obj(w):
            dz, dxN = split(w)
            return sum(0.5 * dz[k] @ H[k] @ dz[k] + g[k] @ dz[k] for k in range(N)) + 0.5 * spec.reg_epsilon * dxN @ dxN

        def obj_grad(w):
            dz, dxN = split(w)
            return np.concatenate([np.concatenate([H[k] @ dz[k] + g[k] for k in range(N)]), spec.reg_epsilon * dxN])

        def eq(w):
            dz, dxN = split(w)
            out = [z0[0, nu:] + dz[0, nu:] - b["xinit"][i]]
            for k in range(N):
                nxt = (z0[k + 1, nu:] + dz[k + 1, nu:]) if k + 1 < N else (xN0 + dxN)
                out.append(ev[k]["xnext"] + ev[k]["AB"] @ dz[k] - nxt)
            return np.concatenate(out)

        def ineq(w):
            dz, _ = split(w)
            out = []
            for k in range(N):
                sl = slice(0, nz) if k >= 1 else slice(0, nu)
                out += [z0[k, sl] + dz[k, sl] - lo[sl], up[sl] - z0[k, sl] - dz[k, sl]]
                if k >= 1:
                    h = ev[k]["h"] + ev[k]["h_jac"] @ dz[k]
                    out += [(h - lh)[np.isfinite(lh)], (uh - h)[np.isfinite(uh)]]
            return np.concatenate(out)

        # start SLSQP from zero (NOT from the oracle's step)
        res = minimize(obj, np.zeros(nvar), jac=obj_grad, method="SLSQP", constraints=[dict(type="eq", fun=eq), dict(type="ineq", fun=ineq)],
                       options=dict(maxiter=400, ftol=1e-14))
        assert res.success, res.message
        dz_ref, dxN_ref = split(res.x)
        step = its[i, 0, :N] - z0[:N]
        stepN = its[i, 0, N, :nx] - xN0
        w_or = np.concatenate([step.ravel(), stepN])
        assert np.abs(eq(w_or)).max() < 1e-8 and ineq(w_or).min() > -1e-8  # the oracle's step is QP-feasible
        assert obj(w_or) == pytest.approx(res.fun, abs=2e-5)  # complementarity <= qp_tol per row
        assert obj(w_or) >= res.fun - 1e-7
        np.testing.assert_allclose(step, dz_ref, atol=5e-4)
        np.testing.assert_allclose(stepN, dxN_ref, atol=5e-4)
        print('qp check: obj gap', obj(w_or) - res.fun, 'max step diff', np.abs(step - dz_ref).max())
        checked += 1
        if checked >= 2:
            break
    assert checked >= 1


@pytest.mark.parametrize("k_done", [1, 3])
def test_later_rti_steps_solve_the_exact_hessian_qp_an_independent_solver_finds(k_done):
    """RTI step k+1 with NON-zero multipliers: the oracle's step from iterate k equals the argmin of the QP an independent
    assembly builds there -- Hessian of the Lagrangian = dt * cost Hessian + sum_j w_j Hess h_j + sum_i pi_i Hess Phi_i with the
    multipliers the oracle holds after k steps (w = lam_hi - lam_lo; read back through the persistent-multiplier state),
    MIRROR by an eigen-decomposition, linearised ERK4 dynamics, boxes, linearised h -- solved by scipy's SLSQP from zero.
    This pins exact-Hessian assembly, MIRROR on indefinite blocks and the warm-started interior-point solve together."""
    import dataclasses

    from scipy.optimize import minimize

    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 1, 3, seed=4)
    d = spec.desc()
    N, nx, nu, nz, npar, nh = spec.N, spec.nx, spec.nu, spec.nz, spec.npar, spec.nh
    P = len(b["xinit"])
    state = np.zeros((P, oracle.state_size(spec)))
    # (res_eq_fail lifted for the intermediate run: a demoted solve would clear the multipliers we want to read back)
    xt_k, ut_k, info_k = oracle.solve_batch(dataclasses.replace(spec, n_sqp=k_done, res_eq_fail=1e9), b["xinit"], b["x0"], b["params"], state=state)
    xt_1, ut_1, info_1 = oracle.solve_batch(dataclasses.replace(spec, n_sqp=k_done + 1), b["xinit"], b["x0"], b["params"])
    lh, uh = map(np.array, spec.constraint_bounds())
    lo, up = np.array(list(d.lower_bound)[:nz]), np.array(list(d.upper_bound)[:nz])
    checked, indefinite = 0, 0
    for i in range(P):
        if info_k[i]["qp_status"] != 0 or info_1[i]["qp_status"] != 0:
            continue
        par = b["params"][i].reshape(N, npar)
        z0 = np.zeros((N + 1, nz))
        z0[:N, :nu], z0[:N, nu:] = ut_k[i], xt_k[i][:N]
        xN0 = xt_k[i][N]
        pi = state[i, : (N + 1) * nx].reshape(N + 1, nx)
        lam = state[i, (N + 1) * nx:].reshape(N, nh)
        w_h = np.where(np.isfinite(lh), -lam, lam)  # L += lam_hi (h - uh) + lam_lo (lh - h)
        assert np.abs(pi).max() > 0 and lam.max() >= 0
        ev = [oracle.eval_stage(d, z0[k], par[k], w_h[k] if k >= 1 else np.zeros(nh), pi[k + 1]) for k in range(N)]
        H = []
        for k in range(N):
            Hk = spec.dt * ev[k]["cost_hess"] + ev[k]["dyn_hess"] + (ev[k]["h_hess"] if k >= 1 else 0.0)
            Hk = 0.5 * (Hk + Hk.T)
            wv, V = np.linalg.eigh(Hk)
            indefinite += int(wv.min() < -spec.reg_epsilon)
            wm = np.where(np.abs(wv) <= spec.reg_epsilon, spec.reg_epsilon, np.abs(wv))
            H.append((V * wm) @ V.T)
        g = [spec.dt * ev[k]["cost_grad"] for k in range(N)]
        nvar = N * nz + nx

        def split(w):
            return w[: N * nz].reshape(N, nz), w[N * nz:]

        def obj(w):
            dz, dxN = split(w)
            return sum(0.5 * dz[k] @ H[k] @ dz[k] + g[k] @ dz[k] for k in range(N)) + 0.5 * spec.reg_epsilon * dxN @ dxN

        def obj_grad(w):
            dz, dxN = split(w)
            return np.concatenate([np.concatenate([H[k] @ dz[k] + g[k] for k in range(N)]), spec.reg_epsilon * dxN])

        def eq(w):
            dz, dxN = split(w)
            out = [z0[0, nu:] + dz[0, nu:] - b["xinit"][i]]
            for k in range(N):
                nxt = (z0[k + 1, nu:] + dz[k + 1, nu:]) if k + 1 < N else (xN0 + dxN)
                out.append(ev[k]["xnext"] + ev[k]["AB"] @ dz[k] - nxt)
            return np.concatenate(out)

        def ineq(w):
            dz, _ = split(w)
            out = []
            for k in range(N):
                sl = slice(0, nz) if k >= 1 else slice(0, nu)
                out += [z0[k, sl] + dz[k, sl] - lo[sl], up[sl] - z0[k, sl] - dz[k, sl]]
                if k >= 1:
                    h = ev[k]["h"] + ev[k]["h_jac"] @ dz[k]
                    out += [(h - lh)[np.isfinite(lh)], (uh - h)[np.isfinite(uh)]]
            return np.concatenate(out)

        res = minimize(obj, np.zeros(nvar), jac=obj_grad, method="SLSQP", constraints=[dict(type="eq", fun=eq), dict(type="ineq", fun=ineq)],
                       options=dict(maxiter=1500, ftol=1e-13))
        assert res.success, res.message
        dz_ref, dxN_ref = split(res.x)
        step = np.concatenate([ut_1[i], xt_1[i][:N]], axis=1) - z0[:N]
        stepN = xt_1[i][N] - xN0
        w_or = np.concatenate([step.ravel(), stepN])
        assert np.abs(eq(w_or)).max() < 1e-8 and ineq(w_or).min() > -1e-6
        assert obj(w_or) == pytest.approx(res.fun, abs=2e-5)
        np.testing.assert_allclose(step, dz_ref, atol=5e-4)
        np.testing.assert_allclose(stepN, dxN_ref, atol=5e-4)
        print("qp check (step %d): obj gap %.2e, max step diff %.2e, max |step| %.2e" % (k_done + 1, obj(w_or) - res.fun, np.abs(step - dz_ref).max(), np.abs(step).max()))
        checked += 1
    assert checked >= 1
    print("stage blocks with an eigenvalue below -eps (MIRROR active):", indefinite)


def test_an_independent_nlp_solver_reaches_the_same_optimum_from_the_same_warm_start():
    """NLP level, no shared starting information: scipy's SLSQP (quasi-Newton SQP on the dense NLP, no MIRROR, no IPM, no
    Riccati) is started from the WARM START -- not from the oracle's solution -- and must end at the oracle's converged point
    whenever both land in the same basin (the NLP is non-convex: obstacles can be passed on either side; where the two end
    in different basins nothing can be compared and the oracle's basin must then simply be feasible)."""
    from scipy.optimize import minimize

    spec = pb.config_tmpc()
    spec.n_sqp = 40  # the oracle's algorithm run to convergence (10 RTI steps stop ~1e-4 short, see above)
    b = scenes.make_scene_batch(spec, 2, 3, seed=11)
    xt, ut, info = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    N, nu, nz = spec.N, spec.nu, spec.nz
    same_basin = 0
    for i in (0, 3, 4):
        assert info[i]["exit_code"] == 1
        nvar, cost, cost_grad, eq, ineq = _nlp_pieces(spec, b["xinit"][i], b["params"][i])
        z0 = b["x0"][i].reshape(N + 1, nz)
        w_start = np.concatenate([z0[:N].ravel(), z0[N, nu:]])
        w_or = np.concatenate([np.concatenate([ut[i], xt[i, :N]], axis=1).ravel(), xt[i, N]])
        res = minimize(cost, w_start, jac=cost_grad, method="SLSQP", constraints=[dict(type="eq", fun=eq), dict(type="ineq", fun=ineq)],
                       options=dict(maxiter=300, ftol=1e-12))
        assert res.success, res.message
        assert np.abs(eq(w_or)).max() < 1e-6 and ineq(w_or).min() > -1e-6
        if np.abs(res.x - w_or).max() < 5e-2:  # same basin
            same_basin += 1
            assert np.abs(res.x - w_or).max() < 2e-3
            assert res.fun == pytest.approx(info[i]["pobj"], abs=5e-5)  # complementarity <= qp_tol per active row
    assert same_basin >= 2
