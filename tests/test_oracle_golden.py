"""Pin the oracle's stage functions and the host parameter layout against golden vectors produced
by the reference's own Python (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from mpc_planner_b200 import problem as pb
from oracle import oracle

SPECS = {
    "cfg1": pb.config_basic(),
    "cfg2": pb.config_tmpc(),
    "cfg2_gaussian": pb.config_tmpc(gaussian=True),
    "cfg3": pb.config_tmpc_decomp(),
    "cfg4": pb.config_safe_horizon(),
}


@pytest.mark.parametrize("name", list(SPECS))
def test_parameter_map_matches_reference(golden, name):
    spec, g = SPECS[name], golden[name]
    assert spec.parameter_map() == g["parameter_map"]
    assert spec.parameter_bundles() == g["parameter_bundles"]
    assert spec.npar == g["npar"] and spec.nh == g["nh"]
    assert spec.nx == g["nx"] and spec.nu == g["nu"]
    assert spec.states == g["states"] and spec.inputs == g["inputs"]
    d = spec.desc()
    assert list(d.lower_bound)[: spec.nz] == g["lower_bound"]
    assert list(d.upper_bound)[: spec.nz] == g["upper_bound"]
    lh, uh = spec.constraint_bounds()
    assert [(-1e15 if v == -np.inf else v) for v in lh] == g["lh"]
    assert [(1e15 if v == np.inf else v) for v in uh] == g["uh"]


def test_reference_parameter_counts():
    # solver_generator/test/test_control_modules.py:86-87: one ellipsoid obstacle -> npar = 7 + 2 extra
    spec1 = pb.ProblemSpec("one", constraints=[(pb.CON_ELLIPSOID, 1)])
    spec0 = pb.ProblemSpec("none", constraints=[])
    assert spec1.npar - spec0.npar == 7 + 2
    # solver_generator/test/test_acados.py:68-74: N=20, 12 obstacles -> len(lh)=len(uh)=12, nx=5, nu=2
    s = pb.config_basic(N=20, n_obstacles=12, n_segments=8)
    lh, uh = s.constraint_bounds()
    assert len(lh) == len(uh) == 12 and s.nx == 5 and s.nu == 2


@pytest.mark.parametrize("name", list(SPECS))
def test_stage_function_values(golden, name):
    spec, g = SPECS[name], golden[name]
    d = spec.desc()
    for s in g["samples"]:
        z, p = np.array(s["z"]), np.array(s["p"])
        o = oracle.eval_stage(d, z, p)
        assert o["cost"] == pytest.approx(s["cost"], rel=1e-12, abs=1e-12)
        np.testing.assert_allclose(o["h"], s["h"], rtol=1e-12, atol=1e-12)
        # continuous model: check through a tiny-step consistency of the ERK map is done below;
        # here compare f(x,u) via finite RK identity: xnext(dt->0) is not exposed, so compare
        # the first-order term using AB? -> covered by test_dynamics_against_reference_f.


def test_appendix_c(golden):
    g = golden["appendix_c"]
    o = oracle.eval_stage(SPECS["cfg2"].desc(), np.array(g["z"]), np.array(g["p"]))
    assert o["cost"] == pytest.approx(0.145025, abs=1e-9)
    assert o["cost"] == pytest.approx(g["cost"], rel=1e-13)
    np.testing.assert_allclose(o["h"], g["h"], rtol=1e-13)
    assert o["h"][8] == pytest.approx(20.337693, abs=1e-6)
    assert all(v == pytest.approx(-99.0) for v in o["h"][:8])


def _rk4_numpy(f, x, u, dt, steps):
    h = dt / steps
    for _ in range(steps):
        k1 = f(x, u)
        k2 = f(x + 0.5 * h * k1, u)
        k3 = f(x + 0.5 * h * k2, u)
        k4 = f(x + h * k3, u)
        x = x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_dynamics_against_reference_f(golden, name):
    """f(x,u) golden values pin the continuous model; the discretisation (ERK4 x 3,
    generate_acados_solver.py:148-150) is restated here in numpy independently of the oracle."""
    spec, g = SPECS[name], golden[name]
    d = spec.desc()
    nu = spec.nu

    def f(x, u):  # numpy restatement checked against golden f right below
        out = [x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), u[1], u[0], x[3]]
        if spec.nx == 6:
            out.append(0.0)
        return np.array(out)

    for s in g["samples"]:
        z, p = np.array(s["z"]), np.array(s["p"])
        np.testing.assert_allclose(f(z[nu:], z[:nu]), s["f"], rtol=1e-14, atol=1e-14)
        o = oracle.eval_stage(d, z, p)
        np.testing.assert_allclose(o["xnext"], _rk4_numpy(f, z[nu:], z[:nu], spec.dt, 3), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("name", list(SPECS))
def test_derivatives_by_finite_differences(golden, name):
    spec, g = SPECS[name], golden[name]
    d = spec.desc()
    nz, nh, nx = spec.nz, spec.nh, spec.nx
    rng = np.random.default_rng(5)
    for s in g["samples"][:6]:
        z, p = np.array(s["z"]), np.array(s["p"])
        lam, pi = rng.uniform(0, 1, nh), rng.uniform(-1, 1, nx)
        o = oracle.eval_stage(d, z, p, lam, pi)
        eps = 1e-6
        gnum, Jnum, ABnum = np.zeros(nz), np.zeros((nh, nz)), np.zeros((nx, nz))
        Hc, Hh, Hd = np.zeros((nz, nz)), np.zeros((nz, nz)), np.zeros((nz, nz))
        for i in range(nz):
            zp, zm = z.copy(), z.copy()
            zp[i] += eps
            zm[i] -= eps
            op, om = oracle.eval_stage(d, zp, p, lam, pi), oracle.eval_stage(d, zm, p, lam, pi)
            gnum[i] = (op["cost"] - om["cost"]) / (2 * eps)
            Jnum[:, i] = (op["h"] - om["h"]) / (2 * eps)
            ABnum[:, i] = (op["xnext"] - om["xnext"]) / (2 * eps)
            Hc[:, i] = (op["cost_grad"] - om["cost_grad"]) / (2 * eps)
            Hh[:, i] = (lam @ op["h_jac"] - lam @ om["h_jac"]) / (2 * eps)
            Hd[:, i] = (pi @ op["AB"] - pi @ om["AB"]) / (2 * eps)
        scale = 1 + np.abs(o["cost_hess"]).max()
        np.testing.assert_allclose(o["cost_grad"], gnum, rtol=1e-5, atol=1e-6 * scale)
        np.testing.assert_allclose(o["h_jac"], Jnum, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(o["AB"], ABnum, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(o["cost_hess"], Hc, rtol=1e-4, atol=1e-6 * scale)
        np.testing.assert_allclose(o["h_hess"], Hh, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["dyn_hess"], Hd, rtol=1e-4, atol=1e-7)


def test_erfinv_restatement():
    from scipy.special import erfinv

    for risk in (0.01, 0.05, 0.2, 0.4):
        assert oracle.lib().oracle_erfinv(risk) == pytest.approx(erfinv(1 - 2 * risk), rel=1e-6)
