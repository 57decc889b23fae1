"""BASELINE configs[4]: P sweep 1 -> 4096 parallel solves (N=20, 8 obstacles) on one GPU: per-batch latency and
solves/s through the host-buffer C ABI (pinned-less numpy buffers) and kernel-only."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import binding, problem as pb, scenes

spec = pb.config_tmpc()
base = scenes.make_scene_batch(spec, 64, 8, seed=3)
rows = []
for P in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    reps = (P + 511) // 512
    xi, x0, par = (np.tile(base[k], (reps, 1))[:P] for k in ("xinit", "x0", "params"))
    s = binding.BatchSolver(spec, P)
    for _ in range(3):
        s.solve(xi, x0, par)
    s.set_timing(True)
    lat = []
    for _ in range(10):
        t = time.perf_counter()
        xt, ut, info, best = s.solve(xi, x0, par)
        lat.append(time.perf_counter() - t)
    ms, _ = s.kernel_ms()
    wall = float(np.median(lat))
    rows.append(dict(P=P, kernel_ms=ms, host_call_ms=wall * 1e3, solves_per_s_kernel=P / (ms * 1e-3), solves_per_s_host_call=P / wall,
                     success=float(np.mean([i["exit_code"] == 1 for i in info]))))
    print(rows[-1], flush=True)
    s.close()
json.dump(rows, open(os.path.join("gpurun_out", "p_sweep_r01.json"), "w"), indent=1)
