"""FLOP count of one solve, MEASURED from the kernel source (no guessed constants).

Builds tests/emu/emu_count.cpp -- the host emulation of csrc/*.cuh with `double` replaced by a counting wrapper -- and runs
it on a sample of the bench batch (bench.py set 0).  Writes profiles/r02_flops.json:
  per_ipm_iteration / per_sqp_iteration / per_solve_fixed   FLOPs (add = mul = div = sqrt = 1, FMA = 2; transcendental calls
  are listed separately and not included), so that bench.py computes
      flops_per_solve = sqp_iter * per_sqp_iteration + qp_iter * per_ipm_iteration + per_solve_fixed
  from the iteration counts it measures.  TEST INFRASTRUCTURE (runs in the build container, no GPU)."""
import ctypes, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mpc_planner_b200 import problem as pb, scenes

EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "libmpcgpu_emu_count.so")
PHASES = ["lin", "assemble+mirror", "ipm_init", "PA", "PB rows", "acc0", "reduce", "factor", "solve", "PD rows", "acc1", "PF rows", "sqp step", "final"]
SQP_PHASES, FIXED_PHASES = {"lin", "assemble+mirror", "ipm_init", "sqp step"}, {"final"}


def build():
    srcs = [os.path.join(EMU, f) for f in ("emu_count.cpp", "counting_double.h")] + [os.path.join(ROOT, "mpc_planner_b200", "csrc", f) for f in ("sqp_core.cuh", "stage_functions.cuh", "registry.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", LIB, srcs[0]])
    return LIB


def count(spec, batch, n):
    lib = ctypes.CDLL(build())
    dp = ctypes.POINTER(ctypes.c_double)
    info = (pb.SolveInfo * n)()
    phase = np.zeros((16, 5), dtype=np.int64)
    xi, x0, par = (np.ascontiguousarray(batch[k][:n]) for k in ("xinit", "x0", "params"))
    desc = spec.desc()
    rc = lib.emu_count_solve(ctypes.byref(desc), n, xi.ctypes.data_as(dp), x0.ctypes.data_as(dp), par.ctypes.data_as(dp), info,
                             phase.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    assert rc == 0, rc
    return phase, info


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n_scenes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    spec = pb.CONFIGS[cfg]()
    batch = scenes.make_scene_batch_chunked(spec, n_scenes, 8, seed=1)  # bench.py set 0, first chunk(s)
    n = n_scenes * 8
    phase, info = count(spec, batch, n)
    sqp_it = sum(i.sqp_iter for i in info)
    qp_it = sum(i.qp_iter for i in info)
    flops = phase[:, :4].sum(axis=1)  # add + mul + div + sqrt
    rows = {}
    per_sqp = per_ipm = fixed = 0.0
    for i, name in enumerate(PHASES):
        rows[name] = {"flops_per_solve": float(flops[i]) / n, "add": int(phase[i, 0]), "mul": int(phase[i, 1]), "div": int(phase[i, 2]),
                      "sqrt": int(phase[i, 3]), "transcendental_calls": int(phase[i, 4])}
        if name in SQP_PHASES:
            per_sqp += flops[i]
        elif name in FIXED_PHASES:
            fixed += flops[i]
        else:
            per_ipm += flops[i]
    out = {
        "config": cfg, "solves": n, "sample": f"first {n_scenes} scenes x 8 guesses of bench.py set 0 (scenes.make_scene_batch_chunked, seed 1)",
        "how": "tests/emu/emu_count.cpp: kernel source (csrc/*.cuh) compiled for the host with `double` -> counting wrapper; add = mul = div = sqrt = 1 FLOP (FMA = 2); all 128 emulated threads counted (includes the per-thread redundant combination of CTA reductions); single-thread Riccati sweep (structural zeros skipped)",
        "mean_sqp_iterations": sqp_it / n, "mean_ipm_iterations": qp_it / n,
        "per_sqp_iteration": per_sqp / sqp_it, "per_ipm_iteration": per_ipm / qp_it, "per_solve_fixed": fixed / n,
        "flops_per_solve_sample": float(flops.sum()) / n,
        "transcendental_calls_per_solve": float(phase[:, 4].sum()) / n,
        "phases": rows,
    }
    tag = "" if cfg == "cfg2" else "_" + cfg
    path = os.path.join(ROOT, "profiles", f"r02_flops{tag}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"{cfg}: {out['flops_per_solve_sample'] / 1e6:.3f} MFLOP/solve on the sample ({out['mean_ipm_iterations']:.1f} IPM its, {out['mean_sqp_iterations']:.1f} SQP its)")
    print(f"  per IPM iteration {out['per_ipm_iteration']:.0f}   per SQP iteration {out['per_sqp_iteration']:.0f}   fixed {out['per_solve_fixed']:.0f}")
    for name, r in rows.items():
        print(f"  {name:18s} {r['flops_per_solve']:12.0f} FLOP/solve   transcendental calls {r['transcendental_calls'] / n:8.1f}")
    print("wrote", path)


if __name__ == "__main__":
    main()
