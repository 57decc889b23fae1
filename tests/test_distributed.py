"""World-size-2 gloo tests (CPU) of the multi-process path: the P guesses of one control step are
sharded over ranks, solved locally (here by the CPU oracle, standing in for the per-GPU CUDA
solver) and reduced to the best trajectory exactly as GuidanceConstraints::optimize /
FindBestPlanner do (guidance_constraints.cpp:344-387, 416-434)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_planner_b200 import problem as pb
from mpc_planner_b200 import scenes
from mpc_planner_b200 import tmpc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, P, weights, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle

    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 1, P, seed=12)
    t = tmpc.DistributedTMPC(spec, lambda xi, x0, par: oracle.solve_batch(spec, xi, x0, par, nthreads=2))
    res = t.optimize(b["xinit"], b["x0"], b["params"], consistency_weights=weights)
    ret[rank] = (res.exit_code, res.best, res.objective, None if res.xtraj is None else res.xtraj.copy())
    dist.destroy_process_group()


@pytest.mark.parametrize("weights", [None, "prefer_last"])
def test_sharded_optimize_matches_serial_find_best(weights):
    from oracle import oracle

    P, world = 5, 2
    spec = pb.config_tmpc()
    b = scenes.make_scene_batch(spec, 1, P, seed=12)
    xt, ut, infos = oracle.solve_batch(spec, b["xinit"], b["x0"], b["params"])
    w = None
    if weights == "prefer_last":
        w = np.ones(P)
        ok = [i for i, inf in enumerate(infos) if inf["exit_code"] == 1]
        w[ok[-1]] = 1e-3  # consistency weight makes the last feasible guess win (guidance_constraints.cpp:358-359)
    want, want_obj = tmpc.find_best_local(infos, w)
    assert want >= 0
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, P, w, ret), nprocs=world, join=True)
    for rank in range(world):
        code, best, obj, x = ret[rank]
        assert best == want and code == infos[want]["exit_code"]
        assert obj == pytest.approx(want_obj, rel=1e-12)
        np.testing.assert_allclose(x, xt[want], rtol=0, atol=0)
    if w is not None:
        assert want == ok[-1]


def test_shard_bounds_cover_everything():
    for P in (1, 5, 8, 13):
        for world in (1, 2, 4, 8):
            blocks = [tmpc.shard_bounds(P, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == P
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_no_feasible_guess_returns_planner0_exit_code():
    spec = pb.config_tmpc()
    fake = lambda xi, x0, par: (np.zeros((len(xi), spec.N + 1, spec.nx)), np.zeros((len(xi), spec.N, spec.nu)),
                                [dict(pobj=1.0, exit_code=4) for _ in range(len(xi))])
    t = tmpc.DistributedTMPC(spec, fake)
    res = t.optimize(np.zeros((3, spec.nx)), np.zeros((3, spec.nz * (spec.N + 1))), np.zeros((3, spec.npar * spec.N)))
    assert res.best == -1 and res.exit_code == 4 and res.xtraj is None
