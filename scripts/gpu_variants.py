"""Time library variants (development aid)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import binding, problem as pb, scenes
spec = pb.config_tmpc()
base = scenes.make_scene_batch(spec, 32, 8, seed=1)
for var in sys.argv[1:]:
    binding._lib = None
    binding.LIB_PATH = os.path.join(os.path.dirname(binding.LIB_PATH), var)
    for P in (8, 1184, 4096):
        reps = (P + 255) // 256
        xi, x0, par = (np.tile(base[k], (reps, 1))[:P] for k in ("xinit", "x0", "params"))
        s = binding.BatchSolver(spec, P)
        s.solve(xi, x0, par)
        s.set_timing(True)
        for _ in range(5):
            xt, ut, info, best = s.solve(xi, x0, par)
        ms, nl = s.kernel_ms()
        print(f"{var:22s} P={P:5d} kernel {ms:8.3f} ms -> {P/(ms*1e-3):10.0f} solves/s  ok={sum(i['exit_code']==1 for i in info)}")
        s.close()
