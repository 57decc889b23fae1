"""Development aid: P copies of the SAME 8 solves (isolates co-residency effects from solve-to-solve variation)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import binding, problem as pb, scenes

if os.environ.get("MPC_LIB"):
    binding.LIB_PATH = os.path.abspath(os.environ["MPC_LIB"])
spec = pb.config_tmpc()
base = scenes.make_scene_batch(spec, 1, 8, seed=1)
for P in [int(a) for a in sys.argv[1:]] or [8]:
    reps = (P + 7) // 8
    xinit, x0, par = (np.tile(base[k], (reps, 1))[:P] for k in ("xinit", "x0", "params"))
    s = binding.BatchSolver(spec, P)
    s.solve(xinit, x0, par)
    s.set_timing(True)
    for _ in range(3):
        xt, ut, info, best = s.solve(xinit, x0, par)
    ms, nl = s.kernel_ms()
    print(f"P={P:5d} kernel {ms:8.3f} ms  mean qp_iter {np.mean([i['qp_iter'] for i in info]):.1f}  per-wave ms {ms / max(1.0, P / 592):.3f}")
    s.close()
