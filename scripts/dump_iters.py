"""Development aid (profiles/r02_tail.md): IPM / SQP iteration counts of every solve of the bench batch, for n_sqp = 1, 2, 3, 4, 6, 10
(cumulative counts: how well the first QPs predict the length of a solve).  Writes gpurun_out/iters_by_nsqp.json."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpc_planner_b200 import binding, problem as pb, scenes
out = {}
for n in (1, 2, 3, 4, 6, 10):
    spec = pb.config_tmpc()
    spec.n_sqp = n
    big = scenes.make_scene_batch_chunked(spec, 512, 8, seed=1)
    s = binding.BatchSolver(spec, 4096)
    xt, ut, info, best = s.solve(big["xinit"], big["x0"], big["params"])
    out[n] = [[i["qp_iter"], i["sqp_iter"], i["exit_code"], i["qp_status"]] for i in info]
    s.close()
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'iters_by_nsqp.json'), 'w'))
