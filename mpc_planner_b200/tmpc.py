"""Batch replacement of the T-MPC++ parallel loop, `GuidanceConstraints::optimize`
(mpc_planner_modules/src/guidance_constraints.cpp:264-388), with the P guesses of one control step
sharded over the ranks of a `torch.distributed` job (one process per GPU).

Reference semantics kept:
  * every planner solves its own NLP (one row of the batch); result.success = exit_code == 1,
    result.objective = pobj, multiplied by `selection_weight_consistency` for the previously
    selected guidance trajectory (:344-360);
  * FindBestPlanner (:416-434): argmin objective over successful planners, first index wins ties,
    -1 if none; the winner's output / info / params become the main solver's (:382-384); if nothing
    is feasible the exit code of planner 0 is returned (:369-373).

Communication: the solves are independent; the only exchange is the final
best-trajectory reduction -- an all_gather of one (objective, global index) pair per rank followed by
a broadcast of the winner's trajectory (~1.2 KB) from its owner.  With backend "nccl" both run over
NVLink; there is no data-path collective.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .problem import ProblemSpec

SENTINEL = 1e10  # the reference's "no feasible planner" objective (guidance_constraints.cpp:419)


@dataclass
class TMPCResult:
    exit_code: int  # best planner's exit code, or planner 0's if none is feasible
    best: int  # global index of the winning guess, -1 if none
    objective: float
    xtraj: Optional[np.ndarray]  # [N+1, nx] of the winner (None if best == -1)
    utraj: Optional[np.ndarray]  # [N, nu]
    local_info: list  # this rank's per-guess info dicts


def shard_bounds(P: int, world: int, rank: int):
    """Contiguous block partition of P guesses over `world` ranks."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def find_best_local(infos, weights=None):
    """FindBestPlanner over a list of info dicts; returns (local index or -1, weighted objective)."""
    best, best_obj = -1, SENTINEL
    for i, inf in enumerate(infos):
        obj = inf["pobj"] * (1.0 if weights is None else float(weights[i]))
        if inf["exit_code"] == 1 and obj < best_obj:
            best, best_obj = i, obj
    return best, best_obj


class DistributedTMPC:
    """optimize() for one control step whose P guesses are split over the process group.

    local_solve(xinit, x0, params) -> (xtraj, utraj, infos[, best]) solves this rank's shard; the
    product path passes `BatchSolver.solve` (CUDA); CPU tests pass the oracle."""

    def __init__(self, spec: ProblemSpec, local_solve: Callable, group=None, device=None):
        self.spec = spec
        self.local_solve = local_solve
        self.group = group
        self.device = device

    def optimize(self, xinit, x0, params, consistency_weights=None) -> TMPCResult:
        import torch
        import torch.distributed as dist

        s = self.spec
        distributed = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size(self.group) if distributed else 1
        rank = dist.get_rank(self.group) if distributed else 0
        xinit = np.asarray(xinit, float).reshape(-1, s.nx)
        P = xinit.shape[0]
        x0 = np.asarray(x0, float).reshape(P, -1)
        params = np.asarray(params, float).reshape(P, -1)
        lo, hi = shard_bounds(P, world, rank)
        infos, xt, ut = [], None, None
        if hi > lo:
            out = self.local_solve(xinit[lo:hi], x0[lo:hi], params[lo:hi])
            xt, ut, infos = out[0], out[1], out[2]
        w = None if consistency_weights is None else np.asarray(consistency_weights, float)[lo:hi]
        lbest, lobj = find_best_local(infos, w)
        # (objective, global index, exit code of global planner 0 if owned here)
        exit0 = float(infos[0]["exit_code"]) if (lo == 0 and hi > 0) else -1.0
        key = torch.tensor([lobj, float(lo + lbest) if lbest >= 0 else -1.0, exit0], dtype=torch.float64, device=self.device)
        if distributed and world > 1:
            keys = [torch.empty_like(key) for _ in range(world)]
            dist.all_gather(keys, key, group=self.group)
            keys = torch.stack(keys).cpu().numpy()
        else:
            keys = key.cpu().numpy()[None, :]
        # identical argmin on every rank; ranks hold increasing index blocks, so the first minimum
        # is the lowest global index, as in the reference's serial loop
        best, best_obj, owner = -1, SENTINEL, -1
        for r in range(world):
            if keys[r, 1] >= 0 and keys[r, 0] < best_obj:
                best_obj, best, owner = float(keys[r, 0]), int(keys[r, 1]), r
        exit0 = int(max(keys[:, 2]))
        if best < 0:
            return TMPCResult(exit0, -1, SENTINEL, None, None, infos)
        # broadcast the winner's trajectory + exit code from its owner
        n_x, n_u = (s.N + 1) * s.nx, s.N * s.nu
        buf = torch.empty(n_x + n_u + 1, dtype=torch.float64, device=self.device)
        if rank == owner:
            li = best - lo
            payload = np.concatenate([np.asarray(xt[li]).ravel(), np.asarray(ut[li]).ravel(), [float(infos[li]["exit_code"])]])
            buf.copy_(torch.from_numpy(payload))
        if distributed and world > 1:
            dist.broadcast(buf, src=dist.get_global_rank(self.group, owner) if self.group is not None else owner, group=self.group)
        h = buf.cpu().numpy()
        return TMPCResult(int(h[-1]), best, best_obj, h[:n_x].reshape(s.N + 1, s.nx), h[n_x:n_x + n_u].reshape(s.N, s.nu), infos)
