"""Quick GPU timing sweep (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import binding, problem as pb, scenes

spec = pb.config_tmpc()
lib = binding.load_library()
print("fp64 peak GFLOP/s:", lib.mpcgpu_measure_fp64_gflops(0))
base = scenes.make_scene_batch(spec, 32, 8, seed=1)
for P in (8, 64, 256, 592, 1184, 4096):
    reps = (P + 255) // 256
    xinit = np.tile(base["xinit"], (reps, 1))[:P]
    x0 = np.tile(base["x0"], (reps, 1))[:P]
    par = np.tile(base["params"], (reps, 1))[:P]
    s = binding.BatchSolver(spec, P)  # group stays 0 -> single chunk, kernel timing available
    s.solve(xinit, x0, par)
    s.set_timing(True)
    t = time.time()
    n = 5
    for _ in range(n):
        xt, ut, info, best = s.solve(xinit, x0, par)
    wall = (time.time() - t) / n
    ms, nl = s.kernel_ms()
    ok = sum(i["exit_code"] == 1 for i in info)
    qp = np.mean([i["qp_iter"] for i in info])
    print(f"P={P:5d} kernel {ms:8.3f} ms  wall {wall*1e3:8.3f} ms  -> {P/(ms*1e-3):10.0f} solves/s (kernel) {P/wall:10.0f} (e2e)  ok={ok} mean qp_iter={qp:.1f}")
    s.close()
