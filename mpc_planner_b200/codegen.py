"""Generator back end: emit CUDA stage-function source from the reference's module definitions.

Replaces, for the B200 backend, what `solver_generator/generate_acados_solver.py:27-65` (`create_acados_model`: symbolic
cost = `objective(modules, z, p, model, settings, 1)`, constraints = `constraints(...)`) hands to acados / CasADi's code
generator.  The reference's module scripts (`mpc_planner_modules/scripts/*.py`) and `solver_definition.py`,
`solver_model.py`, `spline.py`, `util/parameters.py` are imported UNMODIFIED -- the `ModuleManager / ObjectiveModule /
ConstraintModule / DynamicsModel` API of `control_modules.py:4-118` is the front end, as in the reference -- under a
sympy-backed stand-in for the small part of `casadi` they use.  The traced expressions
    l(z, p)      stage cost                      -> value, gradient, upper Hessian
    h(z, p)      stacked constraints             -> values, Jacobian, multiplier-weighted Hessian
are differentiated symbolically, common sub-expressions are eliminated (sympy.cse) and straight-line `__device__`
functions are emitted (`emit_stage_header`) with exactly the signatures the SQP kernel calls (csrc/stage_functions.cuh
`stage_cost`, `stage_constraints`), next to a JSON description of the problem (parameter map, bounds, row signs) that
`problem.ProblemSpec.from_generated` turns into the C-ABI descriptor.

Needs the reference tree (a maintainer runs it where the module scripts live): `python -m mpc_planner_b200.codegen`.
The emitted files of the shipped stacks are committed under mpc_planner_b200/generated_stacks/ so that the GPU box (no
reference tree there) can compile and test them."""
from __future__ import annotations

import json
import os
import sys
import types

REF = os.environ.get("MPC_PLANNER_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "generated_stacks")


# ---------------------------------------------------------------------------------------------
# sympy stand-in for casadi (only what the reference's scripts touch) + import of the reference's Python
# ---------------------------------------------------------------------------------------------
def install_sympy_casadi():
    import numpy as np
    import sympy as sp

    np.Inf = np.inf  # gaussian_constraints.py:65 uses the NumPy < 2 alias
    # spline.py:37 calls np.exp on a symbolic expression: numpy dispatches to a method of that name
    for fn in ("exp", "sqrt", "sin", "cos", "log", "tan"):
        if not hasattr(sp.Expr, fn):
            setattr(sp.Expr, fn, (lambda f: lambda self: getattr(sp, f)(self))(fn))
    c = types.ModuleType("casadi")
    for n in ("sqrt", "cos", "sin", "exp", "log", "tan", "erf"):
        setattr(c, n, getattr(sp, n))
    c.arctan = sp.atan
    c.atan2 = sp.atan2
    c.fabs = sp.Abs
    c.fmax = sp.Max
    c.fmod = lambda a, b: sp.Mod(a, b)
    c.pi = sp.pi

    class SX(sp.Matrix):  # cd.SX(r, c) zero matrix / cd.SX(array)
        def __new__(cls, *a):
            if len(a) == 2 and all(isinstance(i, int) for i in a):
                return sp.zeros(*a)
            return sp.Matrix(a[0])

    c.SX = SX
    c.vertcat = lambda *a: sp.Matrix([x for x in a])
    sys.modules["casadi"] = c
    for pth in (REF + "/solver_generator", REF + "/mpc_planner_modules/scripts"):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    sys.argv[0] = REF + "/mpc_planner_jackalsimulator/scripts/x.py"  # util/files.py resolves paths from argv[0]
    return sp


def trace_stack(name, model, modules, settings):
    """Run the reference's own `define_parameters`, `objective`, `constraints`, bounds on symbols; return a dict."""
    import sympy as sp
    from solver_definition import constraint_lower_bounds, constraint_number, constraint_upper_bounds, constraints, define_parameters, objective
    from util.parameters import Parameters

    params = Parameters()
    define_parameters(modules, params, settings)
    settings = dict(settings)
    settings["params"] = params
    npar = params.length()
    nz = model.get_nvar()
    z = sp.symbols(f"z0:{nz}", real=True)
    p = sp.symbols(f"p0:{npar}", real=True)
    cost = objective(modules, list(z), list(p), model, settings, 1)  # every node uses stage_idx = 1 (generate_acados_solver.py:41,48)
    h = constraints(modules, list(z), list(p), model, settings, 1)
    def scalar(e):  # quadratic forms come back as 1 x 1 matrices (ellipsoid_constraints.py:116)
        e = sp.sympify(e)
        if isinstance(e, sp.MatrixBase):
            assert e.shape == (1, 1), e.shape
            e = e[0, 0]
        return e

    h = [scalar(e) for e in h]
    nh = constraint_number(modules)
    assert nh == len(h)
    lh = [float(v) for v in constraint_lower_bounds(modules)]
    uh = [float(v) for v in constraint_upper_bounds(modules)]
    return dict(name=name, nz=nz, nx=model.nx, nu=model.nu, npar=npar, nh=nh, z=z, p=p, cost=scalar(cost), h=h, lh=lh, uh=uh,
                parameter_map=dict(params._params), states=list(model.states), inputs=list(model.inputs),
                lower_bound=[float(v) for v in model.lower_bound], upper_bound=[float(v) for v in model.upper_bound],
                modules=[m.module_name for m in modules.modules])


# ---------------------------------------------------------------------------------------------
# emission
# ---------------------------------------------------------------------------------------------
def _tri(i, j):
    return i * (i + 1) // 2 + j


class _Printer:
    """C printing of a sympy expression over z[i], pa(i) and CSE temporaries."""

    def __init__(self, sp, z, p):
        from sympy.printing.c import C99CodePrinter

        class P(C99CodePrinter):
            def _print_Pow(self, expr):
                b, e = expr.as_base_exp()
                if e.is_Integer and 2 <= int(e) <= 4:
                    s = self._print(b)
                    return "(" + "*".join(["(" + s + ")"] * int(e)) + ")"
                if e == -1:
                    return "(1.0/(" + self._print(b) + "))"
                if e.is_Integer and -4 <= int(e) <= -2:
                    s = self._print(b)
                    return "(1.0/(" + "*".join(["(" + s + ")"] * (-int(e))) + "))"
                return super()._print_Pow(expr)

        self.p = P(dict(contract=False))
        self.names = {s: f"z[{i}]" for i, s in enumerate(z)}
        self.names.update({s: f"P{i}" for i, s in enumerate(p)})
        self.sp = sp

    def __call__(self, e):
        return self.p.doprint(e.xreplace({k: self.sp.Symbol(v) for k, v in self.names.items()}))


def _emit_block(sp, pr, outputs, used_p, indent="    "):
    """CSE over `outputs` (list of (lhs string, expr)); returns C lines and records the parameters that occur."""
    exprs = [e for _, e in outputs]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("t"), optimizations="basic")
    lines = []
    for e in exprs:
        used_p.update(s for s in e.free_symbols if str(s).startswith("p"))
    for s, e in repl:
        lines.append(f"{indent}const double {s} = {pr(e)};")
    for (lhs, _), e in zip(outputs, red):
        lines.append(f"{indent}{lhs} ({pr(e)});")
    ops = int(sum(sp.count_ops(e) for _, e in repl) + sum(sp.count_ops(e) for e in red))
    return lines, (len(repl), ops)


def emit_stage_header(tr, path, structure):
    """Write the CUDA header with the stage functions of the traced stack `tr`.
    `structure`: dict(ix, iy, ipsi, islack (or -1)) -- the columns the kernel keeps of the constraint Jacobian; a
    constraint that depends on any other variable cannot be expressed by the kernel's row image and is refused."""
    import sympy as sp

    z, p, nz, nh = tr["z"], tr["p"], tr["nz"], tr["nh"]
    pr = _Printer(sp, z, p)
    cost = tr["cost"]
    grad = [sp.diff(cost, zi) for zi in z]
    hess = [(i, j, sp.diff(grad[i], z[j])) for i in range(nz) for j in range(i + 1)]
    cols = [structure["ix"], structure["iy"], structure["ipsi"]]
    islack = structure.get("islack", -1)
    # row signs / bounds: one-sided rows only (every reference module), folded to  s (h - bound) >= 0
    sign, bound, uses_slack = [], [], []
    for j in range(nh):
        lo_fin, up_fin = tr["lh"][j] > -1e14, tr["uh"][j] < 1e14
        if lo_fin == up_fin:
            raise ValueError(f"row {j}: two-sided or unbounded rows are not supported (lh = {tr['lh'][j]}, uh = {tr['uh'][j]})")
        sign.append(1.0 if lo_fin else -1.0)
        bound.append(tr["lh"][j] if lo_fin else tr["uh"][j])
    jac = [[sp.diff(tr["h"][j], zi) for zi in z] for j in range(nh)]
    for j in range(nh):
        for i in range(nz):
            if i in cols:
                continue
            if i == islack:
                if jac[j][i] != 0 and sp.simplify(jac[j][i] + 1) != 0:
                    raise ValueError(f"row {j}: d h / d slack must be 0 or -1 (is {jac[j][i]})")
                continue
            if sp.simplify(jac[j][i]) != 0:
                raise ValueError(f"row {j} depends on z[{i}] ({(tr['inputs'] + tr['states'])[i]}): outside the kernel's row image (x, y, psi, slack)")
        uses_slack.append(bool(islack >= 0 and jac[j][islack] != 0))
    w = sp.symbols(f"w0:{max(nh, 1)}", real=True)
    pr.names.update({s: f"wq{i}" for i, s in enumerate(w)})

    L = []
    L.append(f"// GENERATED by mpc_planner_b200/codegen.py from the reference's module scripts -- do not edit.")
    L.append(f"// stack '{tr['name']}': modules {', '.join(tr['modules'])};  z = [{', '.join(tr['inputs'] + tr['states'])}];  npar = {tr['npar']}, nh = {nh}")
    L.append("#pragma once")
    L.append(f"#define MPCGPU_GEN_NAME \"{tr['name']}\"")
    L.append(f"#define MPCGPU_GEN_NZ {nz}")
    L.append(f"#define MPCGPU_GEN_NH {nh}")
    L.append(f"#define MPCGPU_GEN_NPAR {tr['npar']}")
    L.append("namespace mpcgpu {")
    L.append("namespace gen {")
    L.append("MPC_HD double row_sign(int j) {")
    L.append("    switch (j) {")
    for j in range(nh):
        if sign[j] < 0:
            L.append(f"    case {j}: return -1.0;")
    L.append("    default: return 1.0;\n    }\n}")
    L.append("MPC_HD double row_bound(int j) {")
    L.append("    switch (j) {")
    for j in range(nh):
        if bound[j] != 0.0:
            L.append(f"    case {j}: return {bound[j]!r};")
    L.append("    default: return 0.0;\n    }\n}")
    L.append("MPC_HD bool row_uses_slack(int j) {")
    L.append("    switch (j) {")
    for j in range(nh):
        if uses_slack[j]:
            L.append(f"    case {j}: return true;")
    L.append("    default: return false;\n    }\n}")

    stats = {}
    # ---- cost ----
    used = set()
    val_lines, _ = _emit_block(sp, pr, [("const double cost_value =", cost)], used)
    outs = [(f"g[{i}] += scale *", grad[i]) for i in range(nz) if grad[i] != 0]
    outs += [(f"Hp[{_tri(i, j)}] += scale *", e) for (i, j, e) in hess if e != 0]
    outs = [("const double cost_value =", cost)] + outs
    der_lines, (n_tmp, n_ops) = _emit_block(sp, pr, outs, used)
    stats["cost_cse_temporaries"] = n_tmp
    stats["cost_ops_after_cse"] = n_ops  # value + gradient + lower Hessian (count F_cost from the generator)
    L.append("// value (g == nullptr) or value + gradient + packed lower Hessian, both accumulated scaled by `scale`")
    L.append("template <class PA>\nMPC_HD double stage_cost(const PA &pa, const double *z, double scale, double *g, double *Hp) {")
    for s in sorted(used, key=lambda s: int(str(s)[1:])):
        L.append(f"    const double P{str(s)[1:]} = pa({str(s)[1:]});")
    L.append("    if (!g) {")
    L += ["    " + l for l in val_lines]
    L.append("        return cost_value;\n    }")
    L += der_lines
    L.append("    return cost_value;\n}")

    # ---- constraints: four row quarters (the kernel runs one thread per quarter) ----
    L.append("// rows [NH q / 4, NH (q + 1) / 4): raw h, d h / d (x, y, psi), and  Hxyp += sum_j wq_j * Hess_{x,y,psi} h_j  with")
    L.append("// wq_j = -sign_j * wgt[j] (LAGRANGE: multipliers of the folded rows) or wgt[j] (plain weights)")
    ops = 0
    for q in range(4):
        lo, hi = nh * q // 4, nh * (q + 1) // 4
        used = set()
        outs = []
        for j in range(lo, hi):
            outs.append((f"hv[{j}] =", tr["h"][j]))
            for c, col in enumerate(cols):
                outs.append((f"J[{3 * j + c}] =", jac[j][col]))
        hx = []
        for a in range(3):
            for b in range(a + 1):
                e = sum(w[j] * sp.diff(jac[j][cols[a]], z[cols[b]]) for j in range(lo, hi))
                hx.append((a, b, e))
        outs_h = [(f"Hxyp[{_tri(a, b)}] +=", e) for (a, b, e) in hx if e != 0]
        lines, (_, q_ops) = _emit_block(sp, pr, outs + outs_h, used)
        ops += q_ops
        L.append(f"template <bool LAGRANGE, class PA>\nMPC_HD void stage_constraints_q{q}(const PA &pa, const double *z, double *hv, double *J, const double *wgt, double *Hxyp) {{")
        for s in sorted(used, key=lambda s: int(str(s)[1:])):
            L.append(f"    const double P{str(s)[1:]} = pa({str(s)[1:]});")
        for j in range(lo, hi):
            sg = "-" if sign[j] > 0 else ""
            L.append(f"    const double wq{j} = (Hxyp && wgt) ? (LAGRANGE ? {sg}wgt[{j}] : wgt[{j}]) : 0.0;")
        L += lines
        L.append("    (void)hv; (void)J;")
        L.append("}")
    stats["constraint_ops_after_cse"] = ops  # values + Jacobian + weighted Hessian, four row quarters
    L.append("template <bool LAGRANGE, class PA>\nMPC_HD void stage_constraints(const PA &pa, const double *z, double *hv, double *J, const double *wgt, double *Hxyp, int lo, int hi) {")
    L.append("    double dummy[6] = {0, 0, 0, 0, 0, 0};")
    L.append("    double *H = Hxyp ? Hxyp : dummy;")
    for q in range(4):
        lo, hi = nh * q // 4, nh * (q + 1) // 4
        if hi > lo:
            L.append(f"    if (lo <= {lo} && hi >= {hi}) stage_constraints_q{q}<LAGRANGE>(pa, z, hv, J, Hxyp ? wgt : nullptr, H);")
    L.append("}")
    L.append("}  // namespace gen\n}  // namespace mpcgpu")
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")
    tr["row_sign"], tr["row_bound"], tr["row_uses_slack"] = sign, bound, uses_slack
    return stats


def write_spec_json(tr, path, extra=None):
    d = {k: tr[k] for k in ("name", "nz", "nx", "nu", "npar", "nh", "lh", "uh", "parameter_map", "states", "inputs", "lower_bound", "upper_bound",
                            "modules", "row_sign", "row_bound", "row_uses_slack")}
    d["lh"] = [v if abs(v) < 1e14 else (-1e15 if v < 0 else 1e15) for v in d["lh"]]
    d["uh"] = [v if abs(v) < 1e14 else (-1e15 if v < 0 else 1e15) for v in d["uh"]]
    d.update(extra or {})
    with open(path, "w") as f:
        json.dump(d, f, indent=1)


# ---------------------------------------------------------------------------------------------
# the stacks generated here
# ---------------------------------------------------------------------------------------------
def settings_for(N, n_obs, n_seg=5):
    return {"N": N, "n_discs": 1, "max_obstacles": n_obs, "contouring": {"num_segments": n_seg, "dynamic_velocity_reference": False},
            "linearized_constraints": {"add_halfspaces": 0}, "decomp": {"max_constraints": 12}}


def build_stack(name):
    """(model, modules, settings) of a stack, built with the reference's own classes exactly as its generator scripts do."""
    from control_modules import ModuleManager
    from contouring import ContouringModule
    from ellipsoid_constraints import EllipsoidConstraintModule
    from goal_module import GoalModule
    from guidance_constraints import GuidanceConstraintModule
    from mpc_base import MPCBaseModule
    from path_reference_velocity import PathReferenceVelocityModule
    from solver_model import ContouringSecondOrderUnicycleModel

    if name == "tmpc":  # configuration_tmpc, mpc_planner_jackalsimulator/scripts/generate_jackalsimulator_solver.py:94-103 (BASELINE cfg 2)
        s = settings_for(20, 8)
        m = ModuleManager()
        b = m.add_module(MPCBaseModule(s))
        b.weigh_variable(var_name="a", weight_names="acceleration")
        b.weigh_variable(var_name="w", weight_names="angular_velocity")
        b.weigh_variable(var_name="v", weight_names=["velocity", "reference_velocity"], cost_function=lambda x, w: w[0] * (x - w[1]) ** 2)
        m.add_module(ContouringModule(s))
        m.add_module(GuidanceConstraintModule(s, constraint_submodule=EllipsoidConstraintModule))
        return ContouringSecondOrderUnicycleModel(), m, s
    if name in ("tmpc_2discs", "tmpc_dynvel"):
        # configuration_tmpc (generate_jackalsimulator_solver.py:35-58, 94-103) with settings the hand-written templates do not
        # cover: two ego discs (settings n_discs: 2 -> one ellipsoid row per obstacle and disc, ellipsoid_constraints.py:88-96),
        # or contouring/dynamic_velocity_reference: true (velocity spline tracked in the contouring cost, contouring.py:61-67,
        # 80-81, parameters from PathReferenceVelocity)
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
    if name == "lmpcc_goal":  # configuration_lmpcc, generate_jackalsimulator_solver.py:105-121: Goal objective (no template for it)
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


def generate(name, out_dir=OUT_DIR):
    install_sympy_casadi()
    model, modules, settings = build_stack(name)
    tr = trace_stack(name, model, modules, settings)
    names = tr["inputs"] + tr["states"]
    structure = dict(ix=names.index("x"), iy=names.index("y"), ipsi=names.index("psi"), islack=names.index("slack") if "slack" in names else -1)
    os.makedirs(out_dir, exist_ok=True)
    header = os.path.join(out_dir, f"stage_{name}.cuh")
    stats = emit_stage_header(tr, header, structure)
    write_spec_json(tr, os.path.join(out_dir, f"stage_{name}.json"), dict(N=settings["N"], n_segments=settings["contouring"]["num_segments"], stats=stats))
    return header, stats


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["tmpc", "lmpcc_goal", "tmpc_2discs", "tmpc_dynvel"]):
        hdr, st = generate(nm)
        print(nm, "->", hdr, st)
