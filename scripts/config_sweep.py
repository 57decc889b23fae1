"""BASELINE configs[0], [2], [3] (parity-test cases; the bench line is configs[1]): kernel time, host-call latency and
solves/s of one batch of the size BASELINE.json names, through the host-buffer C ABI.  Writes gpurun_out/configs_r01.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import binding, problem as pb, scenes

CASES = [
    ("cfg1", "configs[0]: MPCC + 4 ellipsoids, N=20, P=1", 1, 1),
    ("cfg2", "configs[1]: T-MPC++ P=8, N=20, 8 obstacles (one control step)", 1, 8),
    ("cfg3", "configs[2]: T-MPC++ P=64, N=30, 12 obstacles + decomp halfspaces + CA cost", 1, 64),
    ("cfg4", "configs[3]: SH-MPC 256 scenario sets x 4 guesses, N=20", 256, 4),
]
rows = []
for name, what, S, G in CASES:
    spec = pb.CONFIGS[name]()
    b = scenes.make_scene_batch(spec, min(S, 16), G, seed=21)
    reps = (S + min(S, 16) - 1) // min(S, 16)
    xi, x0, par = (np.tile(b[k], (reps, 1))[: S * G] for k in ("xinit", "x0", "params"))
    P = S * G
    s = binding.BatchSolver(spec, P)
    s.set_group(G)
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
    rows.append(dict(config=name, what=what, solves=P, kernel_ms=ms, host_call_ms=wall * 1e3, solves_per_s_host_call=P / wall,
                     success=float(np.mean([i["exit_code"] == 1 for i in info])), mean_qp_iter=float(np.mean([i["qp_iter"] for i in info]))))
    print(rows[-1], flush=True)
    s.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(os.path.join("gpurun_out", "configs_r01.json"), "w"), indent=1)
