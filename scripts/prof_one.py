"""Run a few batches of P solves (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import binding, problem as pb, scenes

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
spec = pb.CONFIGS[cfg]()
distinct = len(sys.argv) > 3 and sys.argv[3] == "distinct"  # the bench batch: P / 8 distinct scenes
base = scenes.make_scene_batch_chunked(spec, P // 8, 8, seed=1) if distinct else scenes.make_scene_batch(spec, min(max(P // 8, 1), 32), 8, seed=1)
reps = (P + len(base["xinit"]) - 1) // len(base["xinit"])
xinit = np.tile(base["xinit"], (reps, 1))[:P]
x0 = np.tile(base["x0"], (reps, 1))[:P]
par = np.tile(base["params"], (reps, 1))[:P]
s = binding.BatchSolver(spec, P)
for _ in range(3):
    xt, ut, info, best = s.solve(xinit, x0, par)
print("ok", sum(i["exit_code"] == 1 for i in info), "of", P)
