"""ctypes binding of the host emulation of the kernel source (tests/emu/emu.cpp).  TEST ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libmpcgpu_emu.so")
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def build_generated(registry_header, out, stage_header=None):
    """Host emulation of a GENERATED kernel instantiation (generate.generate_cuda_registry); `stage_header`: stage
    functions emitted by codegen.py."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    extra = ['-DMPCGPU_GENERATED_STAGE="%s"' % os.path.abspath(stage_header)] if stage_header else []
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(_ROOT, "include"),
                           '-DMPCGPU_GENERATED_REGISTRY="%s"' % os.path.abspath(registry_header)] + extra + ["-o", out, os.path.join(_HERE, "emu.cpp")])
    return out


def build_variant(defines, out):
    """Host emulation compiled with extra -D flags (e.g. MPC_RICCATI_LANES: the lane-cooperative factorisation)."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(_ROOT, "include")] + ["-D" + d for d in defines]
                          + ["-o", out, os.path.join(_HERE, "emu.cpp")])
    return out


def build(force=False):
    srcs = [os.path.join(_HERE, "emu.cpp")] + [
        os.path.join(_ROOT, "mpc_planner_b200", "csrc", f) for f in ("sqp_core.cuh", "stage_functions.cuh", "registry.cuh")
    ]
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(s) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(_ROOT, "include"), "-o", _LIB, srcs[0]])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def set_timeout(timeout_s, step_ns, library=None):
    """Fake clock of the emulation (csrc/sqp_core.cuh mpc_clock_ns): every read advances by step_ns."""
    L = ctypes.CDLL(library) if library else lib()
    L.emu_set_timeout.argtypes = [ctypes.c_double, ctypes.c_longlong]
    L.emu_set_timeout(float(timeout_s), int(step_ns))


def eval_stage(desc, z, p, lam=None, pi=None):
    nz, nx, nh = desc.nx + desc.nu, desc.nx, desc.nh
    z = np.ascontiguousarray(z, float).reshape(-1, nz)
    M = z.shape[0]
    p = np.ascontiguousarray(p, float).reshape(M, desc.npar)
    lam = np.zeros((M, max(nh, 1))) if lam is None else np.ascontiguousarray(lam, float).reshape(M, -1)
    pi = np.zeros((M, nx)) if pi is None else np.ascontiguousarray(pi, float).reshape(M, nx)
    out = dict(cost=np.zeros(M), cost_grad=np.zeros((M, nz)), cost_hess=np.zeros((M, nz, nz)), h=np.zeros((M, max(nh, 1))),
               h_jac=np.zeros((M, max(nh, 1), nz)), h_hess=np.zeros((M, nz, nz)), xnext=np.zeros((M, nx)), AB=np.zeros((M, nx, nz)),
               dyn_hess=np.zeros((M, nz, nz)))
    rc = lib().emu_eval_stage(ctypes.byref(desc), M, _p(z), _p(p), _p(lam), _p(pi), _p(out["cost"]), _p(out["cost_grad"]),
                              _p(out["cost_hess"]), _p(out["h"]), _p(out["h_jac"]), _p(out["h_hess"]), _p(out["xnext"]),
                              _p(out["AB"]), _p(out["dyn_hess"]))
    if rc != 0:
        raise RuntimeError(f"emu_eval_stage: {rc}")
    out["h"], out["h_jac"] = out["h"][:, :nh], out["h_jac"][:, :nh]
    return out


def solve_batch(spec, xinit, x0, params, iterates=False, library=None, state=None):
    """`state` [P, state_size] (in/out, C-contiguous float64): persistent multipliers of the P solve slots."""
    from mpc_planner_b200.problem import SolveInfo

    desc = spec.desc()
    N, nx, nu, nz = spec.N, spec.nx, spec.nu, spec.nz
    xinit = np.ascontiguousarray(xinit, float).reshape(-1, nx)
    P = xinit.shape[0]
    x0 = np.ascontiguousarray(x0, float).reshape(P, -1)
    params = np.ascontiguousarray(params, float).reshape(P, -1)
    xtraj, utraj = np.zeros((P, N + 1, nx)), np.zeros((P, N, nu))
    info = (SolveInfo * P)()
    its = np.zeros((P, spec.n_sqp, N + 1, nz)) if iterates else None
    L = ctypes.CDLL(library) if library else lib()
    if state is not None:
        assert not iterates and state.flags.c_contiguous and state.dtype == np.float64 and state.shape[0] == P
        rc = L.emu_solve_stateful(ctypes.byref(desc), P, _p(xinit), _p(x0), _p(params), _p(xtraj), _p(utraj), info, _p(state))
    else:
        rc = L.emu_solve(ctypes.byref(desc), P, _p(xinit), _p(x0), _p(params), _p(xtraj), _p(utraj), info, _p(its))
    if rc != 0:
        raise RuntimeError(f"emu_solve: {rc}")
    infos = [dict(pobj=i.pobj, res_eq=i.res_eq, res_stat_qp=i.res_stat_qp, res_comp_qp=i.res_comp_qp, exit_code=i.exit_code,
                  qp_status=i.qp_status, sqp_iter=i.sqp_iter, qp_iter=i.qp_iter) for i in info]
    return (xtraj, utraj, infos, its) if iterates else (xtraj, utraj, infos)
