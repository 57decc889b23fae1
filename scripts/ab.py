"""A/B of library builds on the GPU box (development aid): parity against the CPU oracle on a small batch, then kernel time
of the bench batch (4096 distinct solves) and of one control step (P = 8).

    python scripts/ab.py [cfg=cfg2] lib1.so lib2.so ...        (no library: the in-tree build)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from mpc_planner_b200 import binding, problem as pb, scenes
from oracle import oracle

args = sys.argv[1:]
cfg = "cfg2"
if args and args[0].startswith("cfg="):
    cfg = args.pop(0)[4:]
libs = args or [binding.LIB_PATH]
spec = pb.CONFIGS[cfg]()
small = scenes.make_scene_batch(spec, 8, 8, seed=7)
oxt, out, oinfo = oracle.solve_batch(spec, small["xinit"], small["x0"], small["params"])
big = scenes.make_scene_batch_chunked(spec, 512, 8, seed=1)
res = []
for lib in libs:
    s = binding.BatchSolver(spec, 4096, library=lib)
    xt, ut, info, best = s.solve(small["xinit"], small["x0"], small["params"])
    worst, codes = 0.0, 0
    for i in range(len(oinfo)):
        codes += info[i]["exit_code"] != oinfo[i]["exit_code"]
        if oinfo[i]["exit_code"] == 1 and info[i]["exit_code"] == 1:
            worst = max(worst, float(np.max(np.abs(xt[i] - oxt[i]) / (1 + np.abs(oxt[i])))), float(np.max(np.abs(ut[i] - out[i]) / (1 + np.abs(out[i])))))
    rec = dict(lib=os.path.relpath(lib, ROOT), cfg=cfg, parity_max_rel=worst, exit_code_mismatches=codes, n_ok=sum(i["exit_code"] == 1 for i in oinfo))
    for name, b in (("P8", {k: small[k][:8] for k in ("xinit", "x0", "params")}), ("P4096", big)):
        s.solve(b["xinit"], b["x0"], b["params"])
        s.set_timing(True)
        ms = []
        for _ in range(6):
            xt, ut, info, best = s.solve(b["xinit"], b["x0"], b["params"])
            ms.append(s.kernel_ms()[0])
        s.set_timing(False)
        rec[name + "_ms"] = min(ms)
        rec[name + "_ok"] = sum(i["exit_code"] == 1 for i in info)
        rec[name + "_qp_iter"] = float(np.mean([i["qp_iter"] for i in info]))
    rec["solves_per_s"] = 4096 / (rec["P4096_ms"] * 1e-3)
    print(json.dumps(rec), flush=True)
    res.append(rec)
    s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "ab.json"), "w") as f:
    json.dump(res, f, indent=1)
