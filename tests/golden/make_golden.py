#!/usr/bin/env python
"""Generate golden stage-function vectors FROM THE REFERENCE'S OWN PYTHON (run in the build
container only: needs /root/reference, which does not exist on the GPU box).

The reference's module scripts (mpc_planner_modules/scripts/*.py) and generator
(solver_generator/solver_definition.py, solver_model.py, spline.py) are imported UNMODIFIED under
a ~20-line numeric stand-in for `casadi` (numpy functions), so that
`objective(...)`, `constraints(...)` and `model.continuous_model(...)` return plain floats.

Output: tests/golden/stage_functions.json -- per config: the reference's parameter map, nh,
bounds, and M seeded random samples (z, p) with the reference's l(z,p), h(z,p), f(x,u).
"""
import json
import math
import os
import sys
import types

import numpy as np

REF = "/root/reference"
np.Inf = np.inf  # gaussian_constraints.py:65 uses the NumPy<2 alias

c = types.ModuleType("casadi")
for n in ("sqrt", "cos", "sin", "exp", "log", "tan", "arctan", "fmod", "fabs"):
    setattr(c, n, getattr(np, n))
c.atan2, c.fmax, c.pi, c.erf = np.arctan2, np.maximum, np.pi, np.vectorize(math.erf)


class SX(np.ndarray):
    def __new__(cls, *a):
        if len(a) == 2 and all(isinstance(i, int) for i in a):
            return np.zeros(a).view(cls)
        return np.asarray(a[0], float).view(cls)


c.SX, c.vertcat = SX, (lambda *a: np.array(a, float))
sys.modules["casadi"] = c
sys.path[:0] = [REF + "/solver_generator", REF + "/mpc_planner_modules/scripts"]
sys.argv[0] = REF + "/mpc_planner_jackalsimulator/scripts/x.py"  # util/files.py resolves paths from argv[0]

from control_modules import ModuleManager  # noqa: E402
from contouring import ContouringModule  # noqa: E402
from curvature_aware_contouring import CurvatureAwareContouringModule  # noqa: E402
from decomp_constraints import DecompConstraintModule  # noqa: E402
from ellipsoid_constraints import EllipsoidConstraintModule  # noqa: E402
from gaussian_constraints import GaussianConstraintModule  # noqa: E402
from goal_module import GoalModule  # noqa: E402
from guidance_constraints import GuidanceConstraintModule  # noqa: E402
from mpc_base import MPCBaseModule  # noqa: E402
from path_reference_velocity import PathReferenceVelocityModule  # noqa: E402
from scenario_constraints import ScenarioConstraintModule  # noqa: E402
from solver_definition import (  # noqa: E402
    constraint_lower_bounds,
    constraint_number,
    constraint_upper_bounds,
    constraints,
    define_parameters,
    objective,
)
from solver_model import (  # noqa: E402
    BicycleModel2ndOrder,
    ContouringSecondOrderUnicycleModel,
    ContouringSecondOrderUnicycleModelWithSlack,
    SecondOrderUnicycleModel,
)
from util.parameters import Parameters  # noqa: E402


def base(settings, slack=False):
    modules = ModuleManager()
    b = modules.add_module(MPCBaseModule(settings))
    b.weigh_variable(var_name="a", weight_names="acceleration")
    b.weigh_variable(var_name="w", weight_names="angular_velocity")
    if slack:
        b.weigh_variable(var_name="slack", weight_names="slack")
    b.weigh_variable(var_name="v", weight_names=["velocity", "reference_velocity"], cost_function=lambda x, w: w[0] * (x - w[1]) ** 2)
    return modules


def settings_for(N, n_obs, n_seg=5, n_decomp=12):
    return {
        "N": N,
        "n_discs": 1,
        "max_obstacles": n_obs,
        "contouring": {"num_segments": n_seg, "dynamic_velocity_reference": False},
        "linearized_constraints": {"add_halfspaces": 0},
        "decomp": {"max_constraints": n_decomp},
    }


def build(name):
    if name == "cfg1":  # configuration_basic, generate_jackalsimulator_solver.py:61-66
        s = settings_for(20, 4)
        m = base(s)
        m.add_module(ContouringModule(s))
        m.add_module(EllipsoidConstraintModule(s))
        return ContouringSecondOrderUnicycleModel(), m, s
    if name in ("cfg2", "cfg2_gaussian"):  # configuration_tmpc, :94-103
        s = settings_for(20, 8)
        m = base(s)
        m.add_module(ContouringModule(s))
        sub = GaussianConstraintModule if name.endswith("gaussian") else EllipsoidConstraintModule
        m.add_module(GuidanceConstraintModule(s, constraint_submodule=sub))
        return ContouringSecondOrderUnicycleModel(), m, s
    if name == "cfg3":  # generate_rosnavigation_solver.py:91-112 with the CA cost
        s = settings_for(30, 12)
        m = base(s, slack=True)
        m.add_module(CurvatureAwareContouringModule(s))
        m.add_module(GuidanceConstraintModule(s, constraint_submodule=EllipsoidConstraintModule))
        m.add_module(DecompConstraintModule(s))
        return ContouringSecondOrderUnicycleModelWithSlack(), m, s
    if name == "cfg4":  # configuration_safe_horizon, generate_jackalsimulator_solver.py:69-91
        s = settings_for(20, 0)
        m = base(s, slack=True)
        m.add_module(ContouringModule(s))
        m.add_module(ScenarioConstraintModule(s))
        return ContouringSecondOrderUnicycleModelWithSlack(), m, s
    if name in ("tmpc_2discs", "tmpc_dynvel"):  # configuration_tmpc with n_discs = 2 / dynamic_velocity_reference (generated code only)
        s = settings_for(20, 8)
        dyn = name == "tmpc_dynvel"
        if name == "tmpc_2discs":
            s["n_discs"] = 2
        s["contouring"]["dynamic_velocity_reference"] = dyn
        m = ModuleManager()
        b = m.add_module(MPCBaseModule(s))
        b.weigh_variable(var_name="a", weight_names="acceleration")
        b.weigh_variable(var_name="w", weight_names="angular_velocity")
        if not dyn:
            b.weigh_variable(var_name="v", weight_names=["velocity", "reference_velocity"], cost_function=lambda x, w: w[0] * (x - w[1]) ** 2)
        m.add_module(ContouringModule(s))
        if dyn:
            m.add_module(PathReferenceVelocityModule(s))
        m.add_module(GuidanceConstraintModule(s, constraint_submodule=EllipsoidConstraintModule))
        return ContouringSecondOrderUnicycleModel(), m, s
    if name == "lmpcc_goal":  # configuration_lmpcc, generate_jackalsimulator_solver.py:105-121 (Goal objective: generated code only)
        s = settings_for(20, 4)
        m = ModuleManager()
        b = m.add_module(MPCBaseModule(s))
        b.weigh_variable(var_name="a", weight_names="acceleration")
        b.weigh_variable(var_name="w", weight_names="angular_velocity")
        m.add_module(GoalModule(s))
        m.add_module(PathReferenceVelocityModule(s))
        m.add_module(EllipsoidConstraintModule(s))
        return ContouringSecondOrderUnicycleModel(), m, s
    raise KeyError(name)


def random_params(pmap, rng):
    """Physically sensible random parameter vector, chosen by parameter name."""
    p = np.zeros(len(pmap))
    seg_len = 5.0
    for name, i in pmap.items():
        if name in ("acceleration", "angular_velocity", "velocity", "contour", "lag"):
            p[i] = rng.uniform(0.05, 1.0)
        elif name == "slack":
            p[i] = rng.uniform(10.0, 1000.0)
        elif name == "goal_weight":
            p[i] = rng.uniform(0.5, 5.0)
        elif name in ("goal_x", "goal_y"):
            p[i] = rng.uniform(-10.0, 20.0)
        elif name.startswith("spline_v"):
            p[i] = rng.uniform(-1, 1)
        elif name == "reference_velocity":
            p[i] = rng.uniform(0.5, 2.5)
        elif name in ("terminal_angle", "terminal_contouring"):
            p[i] = rng.uniform(1.0, 100.0)
        elif name.startswith("spline") and name.endswith("_start"):
            k = int(name[len("spline"):-len("_start")])
            p[i] = seg_len * k
        elif name.startswith("spline_"):
            coef = name[-1]
            scale = {"a": 0.01, "b": 0.05, "c": 1.0, "d": 5.0}[coef]
            p[i] = rng.uniform(-1, 1) * scale + (1.0 if (coef == "c" and name.startswith("spline_x")) else 0.0)
        elif name == "ego_disc_radius":
            p[i] = 0.325
        elif name.startswith("ego_disc_") and name.endswith("_offset"):
            p[i] = rng.uniform(-0.3, 0.3)
        elif name.endswith(("_a1", "_a2")):
            p[i] = rng.uniform(-1, 1)
        elif name.endswith("_b"):
            p[i] = rng.uniform(-5, 20)
        elif name.endswith(("_x", "_y")):
            p[i] = rng.uniform(-10, 10)
        elif name.endswith("_psi"):
            p[i] = rng.uniform(-3, 3)
        elif name.endswith(("_major", "_minor")):
            p[i] = rng.uniform(0.0, 0.6)
        elif name.endswith("_chi"):
            p[i] = rng.uniform(0.5, 6.0)
        elif name.endswith("_risk"):
            p[i] = rng.uniform(0.01, 0.3)
        elif name.endswith("_r"):
            p[i] = rng.uniform(0.2, 0.6)
        else:
            raise KeyError(name)
    return p


def random_z(model, rng):
    z = np.zeros(model.nu + model.nx)
    for nme in model.inputs + model.states:
        idx = (model.inputs.index(nme) if nme in model.inputs else model.nu + model.states.index(nme))
        z[idx] = {
            "a": rng.uniform(-2, 2), "w": rng.uniform(-0.8, 0.8), "x": rng.uniform(-3, 12), "y": rng.uniform(-4, 4),
            "psi": rng.uniform(-3, 3), "v": rng.uniform(0, 3), "spline": rng.uniform(0.0, 24.0),
            "slack": rng.uniform(0, 1.0), "delta": rng.uniform(-0.5, 0.5),
        }[nme]
    return z


def main(M=24):
    out = {}
    for ci, name in enumerate(["cfg1", "cfg2", "cfg2_gaussian", "cfg3", "cfg4", "lmpcc_goal", "tmpc_2discs", "tmpc_dynvel"]):
        model, modules, settings = build(name)
        params = Parameters()
        define_parameters(modules, params, settings)
        settings["params"] = params
        pmap = {k: int(v) for k, v in params._params.items()}
        nh = constraint_number(modules)
        lh = [(-1e15 if v == -np.inf else float(v)) for v in constraint_lower_bounds(modules)]
        uh = [(1e15 if v == np.inf else float(v)) for v in constraint_upper_bounds(modules)]
        rng = np.random.default_rng(1000 + ci)
        samples = []
        for _ in range(M):
            z = random_z(model, rng)
            p = random_params(pmap, rng)
            l1 = float(objective(modules, z, p, model, settings, 1))
            h = [float(v) for v in constraints(modules, z, p, model, settings, 1)]
            f = [float(v) for v in model.continuous_model(z[model.nu:], z[: model.nu])]
            samples.append(dict(z=z.tolist(), p=p.tolist(), cost=l1, h=h, f=f))
        out[name] = dict(
            parameter_map=pmap, parameter_bundles={k: [int(i) for i in v] for k, v in params.parameter_bundles.items()}, npar=params.length(), nh=nh, lh=lh, uh=uh, nx=model.nx, nu=model.nu,
            states=model.states, inputs=model.inputs, lower_bound=list(map(float, model.lower_bound)),
            upper_bound=list(map(float, model.upper_bound)), samples=samples,
        )
        print(name, "npar", params.length(), "nh", nh)

    # Appendix C hand-checked scene: cfg2, straight path, obstacle o at (4+o, (-1)^o 1.5)
    model, modules, settings = build("cfg2")
    params = Parameters(); define_parameters(modules, params, settings); settings["params"] = params
    pm = params._params
    p = np.zeros(params.length())
    for k, v in dict(acceleration=.34, angular_velocity=.85, velocity=.55, reference_velocity=2., contour=.05, lag=.75,
                     terminal_angle=100., terminal_contouring=10., ego_disc_radius=.325, ego_disc_0_offset=0.).items():
        p[pm[k]] = v
    for i in range(5):
        p[pm[f"spline_x{i}_c"]] = 1.0; p[pm[f"spline_x{i}_d"]] = 5.0 * i; p[pm[f"spline{i}_start"]] = 5.0 * i
    for o in range(8):
        p[pm[f"ellipsoid_obst_{o}_x"]] = 4 + o; p[pm[f"ellipsoid_obst_{o}_y"]] = (-1) ** o * 1.5
        p[pm[f"ellipsoid_obst_{o}_r"]] = .4; p[pm[f"ellipsoid_obst_{o}_chi"]] = 1.
        p[pm[f"lin_constraint_{o}_a1"]] = 1.; p[pm[f"lin_constraint_{o}_b"]] = 100.
    z = np.array([.1, -.05, 1, .2, .1, 1.5, 1])
    out["appendix_c"] = dict(z=z.tolist(), p=p.tolist(), cost=float(objective(modules, z, p, model, settings, 1)),
                             cost_terminal_expr=float(objective(modules, z, p, model, settings, 19)),
                             h=[float(v) for v in constraints(modules, z, p, model, settings, 1)])
    print("appendix C:", out["appendix_c"]["cost"], out["appendix_c"]["cost_terminal_expr"], out["appendix_c"]["h"][8])

    # bicycle model continuous dynamics (solver_model.py:302-351)
    bm = BicycleModel2ndOrder()
    rng = np.random.default_rng(77)
    bs = []
    for _ in range(8):
        z = random_z(bm, rng)
        bs.append(dict(z=z.tolist(), f=[float(v) for v in bm.continuous_model(z[bm.nu:], z[: bm.nu])]))
    out["bicycle"] = dict(samples=bs, lower_bound=list(map(float, bm.lower_bound)), upper_bound=list(map(float, bm.upper_bound)))

    # SecondOrderUnicycleModel (solver_model.py:170-190; no generator script of the reference instantiates it): the Goal stack of
    # configuration_lmpcc evaluated on it AND on ContouringSecondOrderUnicycleModel at the same point extended by an arbitrary
    # spline state -- the reference's own numbers show the 4-state problem is the 5-state one with the spline state idle
    m4, m5 = SecondOrderUnicycleModel(), ContouringSecondOrderUnicycleModel()
    rng = np.random.default_rng(4242)
    so = []
    for _ in range(12):
        _, modules, settings = build("lmpcc_goal")
        params = Parameters(); define_parameters(modules, params, settings); settings["params"] = params
        pmap = {k: int(v) for k, v in params._params.items()}
        z4 = random_z(m4, rng)
        z5 = np.append(z4, rng.uniform(0.0, 24.0))
        p = random_params(pmap, rng)
        rec = dict(z4=z4.tolist(), z5=z5.tolist(), p=p.tolist())
        for tag, mdl, z in (("4", m4, z4), ("5", m5, z5)):
            rec["cost" + tag] = float(objective(modules, z, p, mdl, settings, 1))
            rec["h" + tag] = [float(v) for v in constraints(modules, z, p, mdl, settings, 1)]
            rec["f" + tag] = [float(v) for v in mdl.continuous_model(z[mdl.nu:], z[: mdl.nu])]
        so.append(rec)
    out["second_order_unicycle"] = dict(samples=so, states=m4.states, inputs=m4.inputs, lower_bound=list(map(float, m4.lower_bound)),
                                        upper_bound=list(map(float, m4.upper_bound)))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage_functions.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
