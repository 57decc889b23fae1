"""Development aid: build a -DMPC_PROFILE variant of the library and print the per-phase cycle
breakdown of the SQP kernel (thread 0's clock64 deltas).  Not part of the product."""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import build, binding, problem as pb, scenes

# PROF_TAG / PROF_FLAGS: extra development flags for the profiled build (e.g. -DMPC_LIN_SKIP=1)
tag = os.environ.get("PROF_TAG", "")
lib_prof = os.path.join(build.LIB_DIR, "libmpcgpu_prof%s.so" % tag)
if "--build" in sys.argv or not os.path.exists(lib_prof):
    cmd = [build.nvcc_path()] + build.NVCC_FLAGS + ["-DMPC_PROFILE"] + os.environ.get("PROF_FLAGS", "").split() + ["-I" + os.path.join(build.ROOT, "include"), "-o", lib_prof, os.path.join(build.CSRC, "mpcgpu.cu")]
    subprocess.check_call(cmd, cwd=build.CSRC)
    if "--build" in sys.argv:
        sys.exit(0)
binding.LIB_PATH = lib_prof
P = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
cfg = next((a for a in sys.argv[1:] if a in pb.CONFIGS), "cfg2")
spec = pb.CONFIGS[cfg]()
base = scenes.make_scene_batch(spec, max(1, min(P // 8, 32)), 8, seed=1)
reps = (P + len(base["xinit"]) - 1) // len(base["xinit"])
xi, x0, par = (np.tile(base[k], (reps, 1))[:P] for k in ("xinit", "x0", "params"))
s = binding.BatchSolver(spec, P)
s.solve(xi, x0, par)
xt, ut, info, best = s.solve(xi, x0, par)
names = ["lin", "assemble(mirror)", "ipm_init", "PA", "PB rows", "acc0", "reduce", "factor", "solve", "PD rows", "acc1", "PF rows", "sqp step", "final"]
prof = ut.reshape(P, -1)[:, :16]
tot = prof[:, :14].sum(axis=1)
print(f"{cfg} P={P}{tag}: mean total cycles {tot.mean():.0f}; mean qp_iter {np.mean([i['qp_iter'] for i in info]):.1f}; mean sqp_iter {np.mean([i['sqp_iter'] for i in info]):.2f}")
for i, n in enumerate(names):
    print(f"  {n:18s} {prof[:, i].mean():12.0f} cycles  {100 * prof[:, i].mean() / tot.mean():5.1f}%")
