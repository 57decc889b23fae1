"""ctypes binding of the C ABI (include/mpcgpu.h) of mpc_planner_b200/lib/libmpcgpu.so.

This is the Python face of the product path: there is no CPU fallback.  If the library is missing
or no CUDA device is present, construction fails loudly."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .problem import ProblemDesc, ProblemSpec, SolveInfo

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmpcgpu.so")
_libs = {}
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)

# every symbol include/mpcgpu.h declares
EXPORTED_SYMBOLS = [
    "mpcgpu_version", "mpcgpu_strerror", "mpcgpu_device_count", "mpcgpu_problem_supported", "mpcgpu_create", "mpcgpu_destroy", "mpcgpu_describe",
    "mpcgpu_solve_host", "mpcgpu_solve_host_stateful", "mpcgpu_solve_device", "mpcgpu_solve_tmpc_host", "mpcgpu_solve_tmpc_scene_host", "mpcgpu_get_packed_params", "mpcgpu_set_group", "mpcgpu_set_timeout",
    "mpcgpu_set_stateful", "mpcgpu_state_size", "mpcgpu_reset_state", "mpcgpu_get_state", "mpcgpu_set_state", "mpcgpu_host_alloc", "mpcgpu_host_free", "mpcgpu_launch_count",
    "mpcgpu_set_timing", "mpcgpu_kernel_ms", "mpcgpu_eval_stage", "mpcgpu_measure_fp64_gflops",
]


class MpcGpuError(RuntimeError):
    pass


def load_library(path: str | None = None):
    """dlopen the CUDA library (never builds, never falls back).  `path` selects a per-problem library emitted by
    `generate.build_solver_library`; the default is the in-tree library with the shipped module stacks."""
    path = os.path.abspath(path or LIB_PATH)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise MpcGpuError(f"{path} is missing: build it with `python -m mpc_planner_b200.build` (there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    lib.mpcgpu_version.restype = ctypes.c_char_p
    lib.mpcgpu_strerror.restype = ctypes.c_char_p
    lib.mpcgpu_strerror.argtypes = [ctypes.c_int]
    lib.mpcgpu_device_count.restype = ctypes.c_int
    lib.mpcgpu_problem_supported.argtypes = [ctypes.POINTER(ProblemDesc)]
    lib.mpcgpu_create.argtypes = [ctypes.POINTER(ProblemDesc), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.mpcgpu_destroy.argtypes = [ctypes.c_void_p]
    lib.mpcgpu_describe.argtypes = [ctypes.c_void_p, ctypes.POINTER(ProblemDesc), _ip, _ip]
    lib.mpcgpu_solve_host.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, ctypes.POINTER(SolveInfo), _ip]
    lib.mpcgpu_solve_host_stateful.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp, ctypes.POINTER(SolveInfo), _ip]
    lib.mpcgpu_solve_device.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 8
    lib.mpcgpu_solve_tmpc_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, _ip, ctypes.c_double, _dp, _dp,
                                           ctypes.POINTER(SolveInfo), _ip]
    lib.mpcgpu_solve_tmpc_scene_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int, _dp, _dp, _ip, ctypes.c_double,
                                                 ctypes.c_int, _dp, _ip, ctypes.c_double, _dp, _dp, ctypes.POINTER(SolveInfo), _ip]
    lib.mpcgpu_get_packed_params.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp]
    lib.mpcgpu_set_group.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.mpcgpu_set_stateful.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.mpcgpu_set_timeout.argtypes = [ctypes.c_void_p, ctypes.c_double]
    lib.mpcgpu_state_size.argtypes = [ctypes.c_void_p]
    lib.mpcgpu_reset_state.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.mpcgpu_get_state.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp]
    lib.mpcgpu_set_state.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp]
    lib.mpcgpu_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_ulonglong]
    lib.mpcgpu_host_free.argtypes = [ctypes.c_void_p]
    lib.mpcgpu_launch_count.restype = ctypes.c_longlong
    lib.mpcgpu_launch_count.argtypes = [ctypes.c_void_p]
    lib.mpcgpu_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.mpcgpu_kernel_ms.restype = ctypes.c_double
    lib.mpcgpu_kernel_ms.argtypes = [ctypes.c_void_p, _ip]
    lib.mpcgpu_eval_stage.argtypes = [ctypes.c_void_p, ctypes.c_int] + [_dp] * 13
    lib.mpcgpu_measure_fp64_gflops.restype = ctypes.c_double
    lib.mpcgpu_measure_fp64_gflops.argtypes = [ctypes.c_int]
    _libs[path] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _check(rc, what):
    if rc != 0:
        raise MpcGpuError(f"{what} failed ({rc}): {load_library().mpcgpu_strerror(rc).decode()}")


def infos_to_dicts(info):
    return [dict(pobj=i.pobj, res_eq=i.res_eq, res_stat_qp=i.res_stat_qp, res_comp_qp=i.res_comp_qp, exit_code=i.exit_code,
                 qp_status=i.qp_status, sqp_iter=i.sqp_iter, qp_iter=i.qp_iter) for i in info]


INFO_DTYPE = np.dtype([("pobj", "f8"), ("res_eq", "f8"), ("res_stat_qp", "f8"), ("res_comp_qp", "f8"), ("exit_code", "i4"),
                       ("qp_status", "i4"), ("sqp_iter", "i4"), ("qp_iter", "i4")])
assert INFO_DTYPE.itemsize == ctypes.sizeof(SolveInfo)


class BatchSolver:
    """P simultaneous `Solver::solve()` calls on one GPU (mpcgpu_create / mpcgpu_solve_*)."""

    def __init__(self, spec: ProblemSpec, max_batch: int, device: int = 0, library: str | None = None):
        self.lib = load_library(library)
        self.spec = spec
        self.desc = spec.desc()
        self.max_batch = int(max_batch)
        self.device = int(device)
        if not self.lib.mpcgpu_problem_supported(ctypes.byref(self.desc)):
            raise MpcGpuError(f"problem '{spec.name}' has no compiled kernel specialisation (csrc/registry.cuh)")
        h = ctypes.c_void_p()
        _check(self.lib.mpcgpu_create(ctypes.byref(self.desc), self.device, self.max_batch, ctypes.byref(h)), "mpcgpu_create")
        self.handle = h
        self.group = 0

    def set_group(self, group: int):
        """Guesses per control step: `best` becomes one winner index per group of consecutive solves."""
        _check(self.lib.mpcgpu_set_group(self.handle, int(group)), "mpcgpu_set_group")
        self.group = int(group)

    def set_timeout(self, seconds: float):
        """`_params.solver_timeout` (acados_solver_interface.cpp:91-117): per-solve time budget checked on the device
        after every SQP-RTI iteration; <= 0 disables it."""
        _check(self.lib.mpcgpu_set_timeout(self.handle, float(seconds)), "mpcgpu_set_timeout")

    # ---- persistent NLP multipliers per solve slot (Solver::operator= keeps the destination's acados memory) ----
    def set_stateful(self, enable=True):
        _check(self.lib.mpcgpu_set_stateful(self.handle, int(bool(enable))), "mpcgpu_set_stateful")

    def state_size(self):
        return int(self.lib.mpcgpu_state_size(self.handle))

    def reset_state(self, first=0, count=None):
        _check(self.lib.mpcgpu_reset_state(self.handle, int(first), int(self.max_batch - first if count is None else count)), "mpcgpu_reset_state")

    def get_state(self, first=0, count=None):
        count = self.max_batch - first if count is None else count
        out = np.empty((count, self.state_size()))
        _check(self.lib.mpcgpu_get_state(self.handle, int(first), int(count), _p(out)), "mpcgpu_get_state")
        return out

    def set_state(self, state, first=0):
        state = np.ascontiguousarray(state, float).reshape(-1, self.state_size())
        _check(self.lib.mpcgpu_set_state(self.handle, int(first), int(state.shape[0]), _p(state)), "mpcgpu_set_state")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mpcgpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer path: what MPCPlanner::Solver::solve() / GuidanceConstraints::optimize use ----
    def solve(self, xinit, x0, params):
        """Arrays [P, ...] in the reference's flat layouts.  Returns (xtraj, utraj, infos, best)."""
        s = self.spec
        xinit = np.ascontiguousarray(xinit, float).reshape(-1, s.nx)
        P = xinit.shape[0]
        x0 = np.ascontiguousarray(x0, float).reshape(P, s.nz * (s.N + 1))
        params = np.ascontiguousarray(params, float).reshape(P, s.npar * s.N)
        xtraj, utraj = np.empty((P, s.N + 1, s.nx)), np.empty((P, s.N, s.nu))
        info = (SolveInfo * P)()
        group = self.group if self.group > 0 else P
        best = np.full((P + group - 1) // group, -1, dtype=np.int32)
        _check(self.lib.mpcgpu_solve_host(self.handle, P, _p(xinit), _p(x0), _p(params), _p(xtraj), _p(utraj), info,
                                          best.ctypes.data_as(_ip)), "mpcgpu_solve_host")
        return xtraj, utraj, infos_to_dicts(info), (int(best[0]) if self.group == 0 else best)

    # ---- T-MPC++ control steps with device-side halfspace construction and parameter packing ----
    def solve_tmpc(self, xinit, x0, scene_params, guided, robot_radius=0.325):
        """S control steps x G guesses.  xinit [S,nx], x0 [S*G, ...] (guess-minor), scene_params [S, N*npar] (the main
        solver's parameters per step), guided [G] flags.  Returns (xtraj, utraj, infos, best[S])."""
        s = self.spec
        xinit = np.ascontiguousarray(xinit, float).reshape(-1, s.nx)
        S = xinit.shape[0]
        guided = np.ascontiguousarray(guided, np.int32)
        G = len(guided)
        P = S * G
        x0 = np.ascontiguousarray(x0, float).reshape(P, s.nz * (s.N + 1))
        scene_params = np.ascontiguousarray(scene_params, float).reshape(S, s.npar * s.N)
        xtraj, utraj = np.empty((P, s.N + 1, s.nx)), np.empty((P, s.N, s.nu))
        info = (SolveInfo * P)()
        best = np.full(S, -1, dtype=np.int32)
        _check(self.lib.mpcgpu_solve_tmpc_host(self.handle, S, G, _p(xinit), _p(x0), _p(scene_params), guided.ctypes.data_as(_ip),
                                               float(robot_radius), _p(xtraj), _p(utraj), info, best.ctypes.data_as(_ip)),
               "mpcgpu_solve_tmpc_host")
        return xtraj, utraj, infos_to_dicts(info), best

    def solve_tmpc_scene(self, xinit, x0, base_params, obstacles, obstacle_radius, obstacle_flags, guided, risk=0.05, halfspaces=None,
                         robot_radius=0.325):
        """S control steps x G guesses from the compact scene description (mpcgpu_solve_tmpc_scene_host).
        obstacles [S, n_obs, N, 5] (x, y, psi, major, minor per prediction step), halfspaces [S, N, n_hs, 3] or None."""
        s = self.spec
        xinit = np.ascontiguousarray(xinit, float).reshape(-1, s.nx)
        S = xinit.shape[0]
        guided = np.ascontiguousarray(guided, np.int32)
        G = len(guided)
        P = S * G
        x0 = np.ascontiguousarray(x0, float).reshape(P, s.nz * (s.N + 1))
        base_params = np.ascontiguousarray(base_params, float).reshape(S, s.npar)
        obstacles = np.ascontiguousarray(obstacles, float)
        n_obs = obstacles.shape[1] if obstacles.size else 0
        assert obstacles.size == 0 or obstacles.shape == (S, n_obs, s.N, 5)
        obstacle_radius = np.ascontiguousarray(obstacle_radius, float).reshape(S, n_obs)
        obstacle_flags = np.ascontiguousarray(obstacle_flags, np.int32).reshape(S, n_obs)
        n_hs = 0
        if halfspaces is not None:
            halfspaces = np.ascontiguousarray(halfspaces, float)
            n_hs = halfspaces.shape[2]
            assert halfspaces.shape == (S, s.N, n_hs, 3)
        xtraj, utraj = np.empty((P, s.N + 1, s.nx)), np.empty((P, s.N, s.nu))
        info = (SolveInfo * P)()
        best = np.full(S, -1, dtype=np.int32)
        _check(self.lib.mpcgpu_solve_tmpc_scene_host(self.handle, S, G, _p(xinit), _p(x0), _p(base_params), n_obs, _p(obstacles), _p(obstacle_radius),
                                                     obstacle_flags.ctypes.data_as(_ip), float(risk), n_hs, _p(halfspaces),
                                                     guided.ctypes.data_as(_ip), float(robot_radius), _p(xtraj), _p(utraj), info,
                                                     best.ctypes.data_as(_ip)), "mpcgpu_solve_tmpc_scene_host")
        return xtraj, utraj, infos_to_dicts(info), best

    def packed_params(self, P):
        """Parameter blocks of the last device-packed call, reference layout [P, N * npar] (debug / parity)."""
        out = np.empty((P, self.spec.N * self.spec.npar))
        _check(self.lib.mpcgpu_get_packed_params(self.handle, int(P), _p(out)), "mpcgpu_get_packed_params")
        return out

    # ---- device-buffer path (raw pointers, e.g. torch tensors' data_ptr()) ----
    def solve_device(self, P, d_xinit, d_x0, d_params, d_xtraj, d_utraj, d_info, d_best=None, stream=None):
        _check(self.lib.mpcgpu_solve_device(self.handle, int(P), d_xinit, d_x0, d_params, d_xtraj, d_utraj, d_info, d_best, stream),
               "mpcgpu_solve_device")

    def eval_stage(self, z, p, lam=None, pi=None):
        d = self.desc
        nz, nx, nh = d.nx + d.nu, d.nx, d.nh
        z = np.ascontiguousarray(z, float).reshape(-1, nz)
        M = z.shape[0]
        p = np.ascontiguousarray(p, float).reshape(M, d.npar)
        lam = np.zeros((M, max(nh, 1))) if lam is None else np.ascontiguousarray(lam, float).reshape(M, -1)
        pi = np.zeros((M, nx)) if pi is None else np.ascontiguousarray(pi, float).reshape(M, nx)
        out = dict(cost=np.zeros(M), cost_grad=np.zeros((M, nz)), cost_hess=np.zeros((M, nz, nz)), h=np.zeros((M, max(nh, 1))),
                   h_jac=np.zeros((M, max(nh, 1), nz)), h_hess=np.zeros((M, nz, nz)), xnext=np.zeros((M, nx)),
                   AB=np.zeros((M, nx, nz)), dyn_hess=np.zeros((M, nz, nz)))
        _check(self.lib.mpcgpu_eval_stage(self.handle, M, _p(z), _p(p), _p(lam), _p(pi), _p(out["cost"]), _p(out["cost_grad"]),
                                          _p(out["cost_hess"]), _p(out["h"]), _p(out["h_jac"]), _p(out["h_hess"]), _p(out["xnext"]),
                                          _p(out["AB"]), _p(out["dyn_hess"])), "mpcgpu_eval_stage")
        out["h"], out["h_jac"] = out["h"][:, :nh], out["h_jac"][:, :nh]
        return out

    def set_timing(self, enabled=True):
        _check(self.lib.mpcgpu_set_timing(self.handle, int(enabled)), "mpcgpu_set_timing")

    def kernel_ms(self):
        n = ctypes.c_int(0)
        ms = self.lib.mpcgpu_kernel_ms(self.handle, ctypes.byref(n))
        return float(ms), int(n.value)

    @property
    def launches(self):
        return int(self.lib.mpcgpu_launch_count(self.handle))
