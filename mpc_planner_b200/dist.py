"""ctypes binding of the multi-GPU best-trajectory reduction (include/mpcgpu_dist.h, lib/libmpcgpu_dist.so).

One process per GPU.  The guesses of a T-MPC++ control step are split over the ranks in contiguous blocks; each rank
solves its block with its `BatchSolver` and `ShardedTMPC.solve_device` leaves the winner (global index, trajectory, info)
on every rank -- NCCL all-gather of (objective, index) pairs + all-reduce of the owner's record, no host round trip.
Replaces GuidanceConstraints::FindBestPlanner + the winner copy across GPUs
(mpc_planner_modules/src/guidance_constraints.cpp:416-434, 382-384)."""
from __future__ import annotations

import ctypes
import os

from .binding import BatchSolver, MpcGpuError, load_library

_HERE = os.path.dirname(os.path.abspath(__file__))
DIST_LIB_PATH = os.path.join(_HERE, "lib", "libmpcgpu_dist.so")
ID_BYTES = 128
EXPORTED_SYMBOLS = ["mpcgpu_dist_strerror", "mpcgpu_dist_unique_id", "mpcgpu_dist_create", "mpcgpu_dist_destroy",
                    "mpcgpu_dist_solve_tmpc_device", "mpcgpu_dist_reduce_best"]
_lib = None


def load_dist_library():
    global _lib
    if _lib is None:
        if not os.path.exists(DIST_LIB_PATH):
            raise MpcGpuError(f"{DIST_LIB_PATH} is missing: build it with `python -m mpc_planner_b200.build`")
        load_library()  # the single-GPU library first (same soname the dist library needs)
        lib = ctypes.CDLL(DIST_LIB_PATH)
        lib.mpcgpu_dist_strerror.restype = ctypes.c_char_p
        lib.mpcgpu_dist_strerror.argtypes = [ctypes.c_int]
        lib.mpcgpu_dist_unique_id.argtypes = [ctypes.c_void_p]
        lib.mpcgpu_dist_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.mpcgpu_dist_destroy.argtypes = [ctypes.c_void_p]
        lib.mpcgpu_dist_solve_tmpc_device.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 11
        lib.mpcgpu_dist_reduce_best.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 9
        _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        raise MpcGpuError(f"{what} failed ({rc}): {load_dist_library().mpcgpu_dist_strerror(rc).decode()}")


def unique_id() -> bytes:
    """ncclGetUniqueId (rank 0); ship the 128 bytes to the other ranks with whatever the host program has."""
    buf = ctypes.create_string_buffer(ID_BYTES)
    _check(load_dist_library().mpcgpu_dist_unique_id(buf), "mpcgpu_dist_unique_id")
    return buf.raw


class ShardedTMPC:
    """The reduction object of one rank (ncclCommInitRank is collective: construct on every rank)."""

    def __init__(self, solver: BatchSolver, rank: int, nranks: int, uid: bytes):
        self.lib = load_dist_library()
        self.solver = solver
        self.rank, self.nranks = int(rank), int(nranks)
        h = ctypes.c_void_p()
        _check(self.lib.mpcgpu_dist_create(solver.handle, self.rank, self.nranks, ctypes.c_char_p(uid), ctypes.byref(h)), "mpcgpu_dist_create")
        self.handle = h

    def solve_device(self, S, G_local, g0, d_xinit, d_x0, d_params, d_xtraj, d_utraj, d_info, d_best_global, d_win_xtraj, d_win_utraj,
                     d_win_info, stream=None):
        """Raw device pointers (e.g. torch tensors' data_ptr()); asynchronous on `stream`."""
        _check(self.lib.mpcgpu_dist_solve_tmpc_device(self.handle, int(S), int(G_local), int(g0), d_xinit, d_x0, d_params, d_xtraj, d_utraj,
                                                      d_info, d_best_global, d_win_xtraj, d_win_utraj, d_win_info, stream),
               "mpcgpu_dist_solve_tmpc_device")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mpcgpu_dist_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
