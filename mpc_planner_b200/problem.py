"""Problem definition (the NLP as data) for the B200 SQP backend.

Host-side mirror of what the reference's offline generator fixes at generation time:

* the module stack and the order in which its parameters are registered
  (solver_generator/solver_definition.py:5-16 `define_parameters`; duplicates registered once,
  solver_generator/util/parameters.py:43-44), i.e. the content of `parameter_map.yaml`
  (solver_generator/generate_solver.py:34-46);
* the model's variable order / bounds, i.e. `model_map.yaml` (solver_generator/solver_model.py:117-128);
* the acados options (solver_generator/generate_acados_solver.py:143-173).

`ProblemSpec.desc()` packs all of it into the C struct `mpcgpu_problem_desc` (include/mpcgpu.h).
No arithmetic lives here.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

MAX_NZ = 9
MAX_CON_MODULES = 4

MODEL_UNICYCLE, MODEL_UNICYCLE_SLACK, MODEL_BICYCLE = 0, 1, 2
OBJ_CONTOURING, OBJ_CURVATURE_AWARE, OBJ_GENERATED, OBJ_GOAL = 0, 1, 2, 3
CON_LINEAR, CON_ELLIPSOID, CON_GAUSSIAN, CON_DECOMP, CON_SCENARIO, CON_GENERATED = 0, 1, 2, 3, 4, 5


class ProblemDesc(ctypes.Structure):
    """ctypes image of `mpcgpu_problem_desc` (include/mpcgpu.h)."""

    _fields_ = [
        ("model", ctypes.c_int),
        ("objective", ctypes.c_int),
        ("N", ctypes.c_int),
        ("nx", ctypes.c_int),
        ("nu", ctypes.c_int),
        ("npar", ctypes.c_int),
        ("nh", ctypes.c_int),
        ("dt", ctypes.c_double),
        ("erk_steps", ctypes.c_int),
        ("n_segments", ctypes.c_int),
        ("p_acceleration", ctypes.c_int),
        ("p_angular_velocity", ctypes.c_int),
        ("p_slack", ctypes.c_int),
        ("p_velocity", ctypes.c_int),
        ("p_reference_velocity", ctypes.c_int),
        ("weigh_velocity", ctypes.c_int),
        ("p_contour", ctypes.c_int),
        ("p_lag", ctypes.c_int),
        ("p_spline", ctypes.c_int),
        ("p_ego_disc_radius", ctypes.c_int),
        ("p_ego_disc_offset", ctypes.c_int),
        ("n_con_modules", ctypes.c_int),
        ("con_kind", ctypes.c_int * MAX_CON_MODULES),
        ("con_count", ctypes.c_int * MAX_CON_MODULES),
        ("con_param", ctypes.c_int * MAX_CON_MODULES),
        ("lower_bound", ctypes.c_double * MAX_NZ),
        ("upper_bound", ctypes.c_double * MAX_NZ),
        ("n_sqp", ctypes.c_int),
        ("qp_iter_max", ctypes.c_int),
        ("qp_tol", ctypes.c_double),
        ("reg_epsilon", ctypes.c_double),
        ("qp_mu0", ctypes.c_double),
        ("qp_thr0", ctypes.c_double),
        ("qp_warm_start", ctypes.c_int),
        ("qp_ws_thr", ctypes.c_double),
        ("qp_ws_mu", ctypes.c_double),
        ("qp_relax", ctypes.c_double),
        ("res_eq_fail", ctypes.c_double),
        ("p_goal", ctypes.c_int),
    ]


class SolveInfo(ctypes.Structure):
    """ctypes image of `mpcgpu_info` (include/mpcgpu.h)."""

    _fields_ = [
        ("pobj", ctypes.c_double),
        ("res_eq", ctypes.c_double),
        ("res_stat_qp", ctypes.c_double),
        ("res_comp_qp", ctypes.c_double),
        ("exit_code", ctypes.c_int),
        ("qp_status", ctypes.c_int),
        ("sqp_iter", ctypes.c_int),
        ("qp_iter", ctypes.c_int),
    ]


# Variable order and bounds per model: solver_generator/solver_model.py
_MODELS = {
    MODEL_UNICYCLE: dict(  # ContouringSecondOrderUnicycleModel :193-205
        inputs=["a", "w"],
        states=["x", "y", "psi", "v", "spline"],
        lower=[-2.0, -0.8, -2000.0, -2000.0, -math.pi * 4, -0.01, -1.0],
        upper=[2.0, 0.8, 2000.0, 2000.0, math.pi * 4, 3.0, 10000.0],
    ),
    MODEL_UNICYCLE_SLACK: dict(  # ContouringSecondOrderUnicycleModelWithSlack :274-285
        inputs=["a", "w"],
        states=["x", "y", "psi", "v", "spline", "slack"],
        lower=[-2.0, -0.8, -2000.0, -2000.0, -math.pi * 4, -0.01, -1.0, 0.0],
        upper=[2.0, 0.8, 2000.0, 2000.0, math.pi * 4, 3.0, 10000.0, 5000.0],
    ),
    MODEL_BICYCLE: dict(  # BicycleModel2ndOrder :302-317
        inputs=["a", "w", "slack"],
        states=["x", "y", "psi", "v", "delta", "spline"],
        lower=[-3.0, -1.5, 0.0, -1.0e6, -1.0e6, -math.pi * 4, -0.01, -0.55, -1.0],
        upper=[3.0, 1.5, 1.0e2, 1.0e6, 1.0e6, math.pi * 4, 5.0, 0.55, 5000.0],
    ),
}


@dataclass
class ProblemSpec:
    """A module stack + model + solver options; the analogue of one `generate_*_solver.py`."""

    name: str
    model: int = MODEL_UNICYCLE
    objective: int = OBJ_CONTOURING
    N: int = 20
    dt: float = 0.2  # integrator_step, settings.yaml:3
    n_segments: int = 5  # contouring/num_segments, settings.yaml:60
    weigh_slack: bool = False  # base_module.weigh_variable("slack", "slack")
    weigh_velocity: bool = True  # lambda cost on v (dynamic_velocity_reference == False)
    constraints: List[Tuple[int, int]] = field(default_factory=list)  # (kind, rows)
    # acados options
    n_sqp: int = 10  # settings.yaml:17
    qp_iter_max: int = 50  # generate_acados_solver.py:172
    qp_tol: float = 1e-5  # :162
    reg_epsilon: float = 1e-4  # acados default reg_epsilon for MIRROR (:157)
    qp_mu0: float = 1e1
    qp_thr0: float = 1e-1
    qp_warm_start: int = 2  # generate_acados_solver.py:173
    qp_ws_thr: float = 1e-4
    qp_ws_mu: float = 1e-5
    qp_relax: float = 1e-7
    res_eq_fail: float = 1e-2  # acados_solver_interface.cpp:176
    # model bounds that differ from the model class's table, by variable name: {"w": (-2.0, 2.0), ...}  (the reference's
    # DynamicsModel carries lower_bound / upper_bound as plain data, solver_model.py:129-167)
    bounds_override: Dict[str, Tuple[float, float]] = field(default_factory=dict)

    # ---- model_map ----
    @property
    def inputs(self) -> List[str]:
        return _MODELS[self.model]["inputs"]

    @property
    def states(self) -> List[str]:
        return _MODELS[self.model]["states"]

    @property
    def nu(self) -> int:
        return len(self.inputs)

    @property
    def nx(self) -> int:
        return len(self.states)

    @property
    def nz(self) -> int:
        return self.nu + self.nx

    @property
    def nh(self) -> int:
        return sum(c for _, c in self.constraints)

    def model_bounds(self):
        """(lower, upper) over z = [u; x]: the model class's table with `bounds_override` applied by variable name."""
        lo, up = list(_MODELS[self.model]["lower"]), list(_MODELS[self.model]["upper"])
        names = self.inputs + self.states
        for name, (l, u) in self.bounds_override.items():
            lo[names.index(name)], up[names.index(name)] = float(l), float(u)
        return lo, up

    def model_map(self) -> Dict[str, list]:
        """name -> ["x"|"u", index in z, lower, upper]  (solver_model.py:117-128 save_map)."""
        lo, up = _MODELS[self.model]["lower"], _MODELS[self.model]["upper"]
        m = {}
        for i, s in enumerate(self.states):
            m[s] = ["x", i + self.nu, lo[self.nu + i], up[self.nu + i]]
        for i, u in enumerate(self.inputs):
            m[u] = ["u", i, lo[i], up[i]]
        return m

    # ---- parameter_map ----
    def parameter_map(self) -> Dict[str, int]:
        """name -> index, registered in the reference's order (objectives, then constraints)."""
        pm: Dict[str, int] = {}

        def add(name):
            if name not in pm:  # util/parameters.py:43-44
                pm[name] = len(pm)

        # MPCBaseModule (mpc_base.py:24-28), weights in weigh_variable order
        add("acceleration")
        add("angular_velocity")
        if self.weigh_slack:
            add("slack")
        if self.weigh_velocity:
            add("velocity")
            add("reference_velocity")
        # Contouring / CurvatureAwareContouring (contouring.py:22-46)
        add("contour")
        add("lag")
        add("velocity")
        add("reference_velocity")
        add("terminal_angle")
        add("terminal_contouring")
        for i in range(self.n_segments):
            for ax in ("x", "y"):
                for c in "abcd":
                    add(f"spline_{ax}{i}_{c}")
            add(f"spline{i}_start")
        for kind, n in self.constraints:
            if kind == CON_LINEAR:  # guidance_constraints.py:73-78
                for i in range(n):
                    for s in ("a1", "a2", "b"):
                        add(f"lin_constraint_{i}_{s}")
            elif kind == CON_ELLIPSOID:  # ellipsoid_constraints.py:36-49
                add("ego_disc_radius")
                add("ego_disc_0_offset")
                for i in range(n):
                    for s in ("x", "y", "psi", "major", "minor", "chi", "r"):
                        add(f"ellipsoid_obst_{i}_{s}")
            elif kind == CON_GAUSSIAN:  # gaussian_constraints.py:37-48
                add("ego_disc_radius")
                add("ego_disc_0_offset")
                for i in range(n):
                    for s in ("x", "y", "major", "minor", "risk", "r"):
                        add(f"gaussian_obst_{i}_{s}")
            elif kind == CON_DECOMP:  # decomp_constraints.py:43-51
                add("ego_disc_0_offset")
                for i in range(n):
                    for s in ("a1", "a2", "b"):
                        add(f"disc_0_decomp_{i}_{s}")
            elif kind == CON_SCENARIO:  # scenario_constraints.py:39-47
                add("ego_disc_0_offset")
                for i in range(n):
                    for s in ("a1", "a2", "b"):
                        add(f"disc_0_scenario_constraint_{i}_{s}")
            else:
                raise ValueError(kind)
        return pm

    @property
    def npar(self) -> int:
        return len(self.parameter_map())

    def parameter_bundles(self) -> Dict[str, List[int]]:
        """bundle name -> parameter indices, as `Parameters.parameter_bundles` of the reference
        (util/parameters.py:46-53): the generated C++ `setSolverParameter<Bundle>(k, params, value, index)`
        setters are keyed by these (generate_cpp_files.py:235-254)."""
        import re

        rules = [
            (r"^spline_([xy])\d+_([abcd])$", lambda m: f"spline_{m.group(1)}_{m.group(2)}"),
            (r"^spline\d+_start$", lambda m: "spline_start"),
            (r"^lin_constraint_\d+_(a1|a2|b)$", lambda m: f"lin_constraint_{m.group(1)}"),
            (r"^ego_disc_\d+_offset$", lambda m: "ego_disc_offset"),
            (r"^ellipsoid_obst_\d+_(\w+)$", lambda m: f"ellipsoid_obst_{m.group(1)}"),
            (r"^gaussian_obst_\d+_(\w+)$", lambda m: f"gaussian_obst_{m.group(1)}"),
            (r"^disc_\d+_decomp_\d+_(a1|a2|b)$", lambda m: f"decomp_{m.group(1)}"),
        ]
        bundles: Dict[str, List[int]] = {}
        for name, idx in self.parameter_map().items():
            bundle = name  # scenario constraints and scalar weights are their own bundle
            for pat, fn in rules:
                m = re.match(pat, name)
                if m:
                    bundle = fn(m)
                    break
            bundles.setdefault(bundle, []).append(idx)
        return bundles

    def constraint_first_param(self) -> List[int]:
        pm = self.parameter_map()
        first = {
            CON_LINEAR: "lin_constraint_0_a1",
            CON_ELLIPSOID: "ellipsoid_obst_0_x",
            CON_GAUSSIAN: "gaussian_obst_0_x",
            CON_DECOMP: "disc_0_decomp_0_a1",
            CON_SCENARIO: "disc_0_scenario_constraint_0_a1",
        }
        return [pm[first[k]] for k, _ in self.constraints]

    def constraint_bounds(self):
        """(lh, uh) per row: get_lower_bound/get_upper_bound of every constraint module."""
        lh, uh = [], []
        for kind, n in self.constraints:
            if kind == CON_ELLIPSOID:
                lh += [1.0] * n
                uh += [math.inf] * n
            elif kind == CON_GAUSSIAN:
                lh += [0.0] * n
                uh += [math.inf] * n
            else:
                lh += [-math.inf] * n
                uh += [0.0] * n
        return lh, uh

    def desc(self) -> ProblemDesc:
        pm = self.parameter_map()
        d = ProblemDesc()
        d.model, d.objective, d.N = self.model, self.objective, self.N
        d.nx, d.nu, d.npar, d.nh = self.nx, self.nu, len(pm), self.nh
        d.dt, d.erk_steps, d.n_segments = self.dt, 3, self.n_segments
        d.p_acceleration = pm["acceleration"]
        d.p_angular_velocity = pm["angular_velocity"]
        d.p_slack = pm.get("slack", -1)
        d.p_velocity = pm["velocity"]
        d.p_reference_velocity = pm["reference_velocity"]
        d.weigh_velocity = 1 if self.weigh_velocity else 0
        d.p_contour, d.p_lag = pm["contour"], pm["lag"]
        d.p_spline = pm["spline_x0_a"]
        d.p_ego_disc_radius = pm.get("ego_disc_radius", -1)
        d.p_ego_disc_offset = pm.get("ego_disc_0_offset", -1)
        if len(self.constraints) > MAX_CON_MODULES:
            raise ValueError("too many constraint modules")
        d.n_con_modules = len(self.constraints)
        for i, ((k, n), p0) in enumerate(zip(self.constraints, self.constraint_first_param())):
            d.con_kind[i], d.con_count[i], d.con_param[i] = k, n, p0
        lo, up = self.model_bounds()
        for i in range(self.nz):
            d.lower_bound[i], d.upper_bound[i] = lo[i], up[i]
        d.n_sqp, d.qp_iter_max, d.qp_tol = self.n_sqp, self.qp_iter_max, self.qp_tol
        d.reg_epsilon, d.qp_mu0, d.qp_thr0 = self.reg_epsilon, self.qp_mu0, self.qp_thr0
        d.res_eq_fail = self.res_eq_fail
        d.qp_warm_start, d.qp_ws_thr, d.qp_ws_mu = self.qp_warm_start, self.qp_ws_thr, self.qp_ws_mu
        d.qp_relax = self.qp_relax
        d.p_goal = -1
        return d

    # sizes of the reference's flat arrays (AcadosParameters, acados_solver_interface.h:51-56)
    def sizes(self):
        return dict(
            xinit=self.nx,
            x0=self.nz * (self.N + 1),
            params=self.npar * self.N,
            xtraj=self.nx * (self.N + 1),
            utraj=self.nu * self.N,
        )


@dataclass
class GeneratedSpec(ProblemSpec):
    """A module stack whose stage functions were EMITTED by mpc_planner_b200/codegen.py (`stage_<name>.cuh` + `.json`): the
    parameter map, bounds and row sides are the ones the reference's own `define_parameters` / `constraint_*_bounds`
    produced when the generator traced the stack.  `oracle_objective` (optional) names the hand-restated objective the CPU
    oracle checks this stack with."""

    gen: dict = field(default_factory=dict)
    oracle_objective: int = OBJ_GENERATED

    @staticmethod
    def load(name, directory=None, **kw):
        import json
        import os

        directory = directory or os.path.join(os.path.dirname(os.path.abspath(__file__)), "generated_stacks")
        with open(os.path.join(directory, f"stage_{name}.json")) as f:
            g = json.load(f)
        model = {("a", "w", "x", "y", "psi", "v", "spline"): MODEL_UNICYCLE, ("a", "w", "x", "y", "psi", "v", "spline", "slack"): MODEL_UNICYCLE_SLACK}[
            tuple(g["inputs"] + g["states"])]
        return GeneratedSpec("generated_" + name, model=model, objective=OBJ_GENERATED, N=g["N"], n_segments=g["n_segments"],
                             constraints=[(CON_GENERATED, g["nh"])], gen=g, **kw)

    def header(self, directory=None):
        import os

        directory = directory or os.path.join(os.path.dirname(os.path.abspath(__file__)), "generated_stacks")
        return os.path.join(directory, f"stage_{self.gen['name']}.cuh")

    def parameter_map(self):
        return dict(self.gen["parameter_map"])

    def parameter_bundles(self):
        return {k: [v] for k, v in self.gen["parameter_map"].items()}

    def constraint_first_param(self):
        return [0]

    def constraint_bounds(self):
        lh = [(-math.inf if v <= -1e14 else v) for v in self.gen["lh"]]
        uh = [(math.inf if v >= 1e14 else v) for v in self.gen["uh"]]
        return lh, uh

    def desc(self, for_oracle=False) -> ProblemDesc:
        pm = self.parameter_map()
        d = ProblemDesc()
        d.model, d.objective, d.N = self.model, self.objective, self.N
        d.nx, d.nu, d.npar, d.nh = self.nx, self.nu, len(pm), self.nh
        d.dt, d.erk_steps, d.n_segments = self.dt, 3, self.n_segments
        for f in ("p_acceleration", "p_angular_velocity", "p_slack", "p_velocity", "p_reference_velocity", "p_contour", "p_lag", "p_spline",
                  "p_ego_disc_radius", "p_ego_disc_offset", "p_goal"):
            setattr(d, f, -1)
        d.weigh_velocity = 0
        d.n_con_modules = 1
        d.con_kind[0], d.con_count[0], d.con_param[0] = CON_GENERATED, self.nh, 0
        lo, up = self.model_bounds()
        for i in range(self.nz):
            d.lower_bound[i], d.upper_bound[i] = lo[i], up[i]
        d.n_sqp, d.qp_iter_max, d.qp_tol = self.n_sqp, self.qp_iter_max, self.qp_tol
        d.reg_epsilon, d.qp_mu0, d.qp_thr0 = self.reg_epsilon, self.qp_mu0, self.qp_thr0
        d.res_eq_fail = self.res_eq_fail
        d.qp_warm_start, d.qp_ws_thr, d.qp_ws_mu = self.qp_warm_start, self.qp_ws_thr, self.qp_ws_mu
        d.qp_relax = self.qp_relax
        if for_oracle:
            # the hand-restated NLP the oracle checks this stack with (same parameter vector)
            d.objective = self.oracle_objective
            d.p_acceleration, d.p_angular_velocity = pm["acceleration"], pm["angular_velocity"]
            if self.oracle_objective == OBJ_GOAL:
                d.p_goal = pm["goal_weight"]
                assert pm["goal_x"] == d.p_goal + 1 and pm["goal_y"] == d.p_goal + 2
                n = sum(1 for k in pm if k.startswith("ellipsoid_obst_") and k.endswith("_x"))
                d.con_kind[0], d.con_count[0], d.con_param[0] = CON_ELLIPSOID, n, pm["ellipsoid_obst_0_x"]
                d.p_ego_disc_radius, d.p_ego_disc_offset = pm["ego_disc_radius"], pm["ego_disc_0_offset"]
            else:
                raise ValueError("no oracle restatement for this generated stack")
        return d


def second_order_unicycle_goal(**kw) -> "GeneratedSpec":
    """`SecondOrderUnicycleModel` (solver_model.py:170-190: states x, y, psi, v; no generator script of the reference instantiates
    it) under the Goal stack of configuration_lmpcc.  It runs as its `ContouringSecondOrderUnicycleModel` twin with the spline
    state idle: the reference's own Python gives the same l, h and the same first four rows of f for the two models at every
    point (tests/golden `second_order_unicycle`), the fifth row (spline' = v) feeds nothing back, and the spline bound
    [-1, 10000] cannot bind -- so (x, y, psi, v, a, w) of the 5-state solve are the 4-state problem's solution (to the interior-point
    tolerance: the barrier of that never-active bound is the only place the spline value enters).  The model's own bounds
    (:177-178) replace the contouring model's."""
    return GeneratedSpec.load("lmpcc_goal", bounds_override={"a": (-2.0, 2.0), "w": (-2.0, 2.0), "x": (-200.0, 200.0), "y": (-200.0, 200.0),
                                                             "psi": (-math.pi * 4, math.pi * 4), "v": (-2.0, 3.0)}, **kw)


# ---- the BASELINE.json configs ----
def config_basic(N=20, n_obstacles=4, n_segments=5) -> ProblemSpec:
    """cfg 1: MPCBase + Contouring + Ellipsoid (generate_jackalsimulator_solver.py:61-66)."""
    return ProblemSpec("basic", N=N, n_segments=n_segments, constraints=[(CON_ELLIPSOID, n_obstacles)])


def config_tmpc(N=20, n_obstacles=8, n_segments=5, gaussian=False) -> ProblemSpec:
    """cfg 2/5: ... + Guidance(linear + Ellipsoid|Gaussian) (generate_jackalsimulator_solver.py:94-103)."""
    sub = CON_GAUSSIAN if gaussian else CON_ELLIPSOID
    return ProblemSpec(
        "tmpc_gaussian" if gaussian else "tmpc",
        N=N,
        n_segments=n_segments,
        constraints=[(CON_LINEAR, n_obstacles), (sub, n_obstacles)],
    )


def config_tmpc_decomp(N=30, n_obstacles=12, n_decomp=12, n_segments=5) -> ProblemSpec:
    """cfg 3: slack model + CA contouring + Guidance + Decomp (generate_rosnavigation_solver.py:91-112)."""
    return ProblemSpec(
        "tmpc_decomp",
        model=MODEL_UNICYCLE_SLACK,
        objective=OBJ_CURVATURE_AWARE,
        N=N,
        n_segments=n_segments,
        weigh_slack=True,
        constraints=[(CON_LINEAR, n_obstacles), (CON_ELLIPSOID, n_obstacles), (CON_DECOMP, n_decomp)],
    )


def config_safe_horizon(N=20, n_scenario=24, n_segments=5) -> ProblemSpec:
    """cfg 4: slack model + Contouring + Scenario halfspaces (generate_jackalsimulator_solver.py:69-91)."""
    return ProblemSpec(
        "safe_horizon",
        model=MODEL_UNICYCLE_SLACK,
        N=N,
        n_segments=n_segments,
        weigh_slack=True,
        constraints=[(CON_SCENARIO, n_scenario)],
    )


def config_bicycle(N=20, n_obstacles=4, n_segments=5) -> ProblemSpec:
    """BASELINE configs[0] as written: BicycleModel2ndOrder (solver_model.py:302-351; inputs a, w, slack; states x, y, psi, v,
    delta, spline) + MPCBase(a, w, slack, v) + Contouring + Ellipsoid(4).  No shipped generator script of the reference uses this
    model; the stack is the jackalsimulator `configuration_basic` with the model swapped."""
    return ProblemSpec("bicycle", model=MODEL_BICYCLE, N=N, n_segments=n_segments, weigh_slack=True, constraints=[(CON_ELLIPSOID, n_obstacles)])


CONFIGS = {
    "cfg0_bicycle": config_bicycle,
    "cfg1": config_basic,
    "cfg2": config_tmpc,
    "cfg3": config_tmpc_decomp,
    "cfg4": config_safe_horizon,
    "cfg2_gaussian": lambda: config_tmpc(gaussian=True),  # T-MPC with the CC-MPC submodule (generate_jackal_solver.py:53-73)
}
