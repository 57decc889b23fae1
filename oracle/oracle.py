"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (mpc_planner_b200/) never imports this module."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "mpc_oracle.cpp")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_erfinv.restype = ctypes.c_double
        _lib.oracle_erfinv.argtypes = [ctypes.c_double]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def eval_stage(desc, z, p, lam=None, pi=None):
    """Stage functions and derivatives at one (z, p).  Returns a dict of numpy arrays."""
    nz, nx, nh = desc.nx + desc.nu, desc.nx, desc.nh
    z = np.ascontiguousarray(z, float)
    p = np.ascontiguousarray(p, float)
    lam = np.zeros(max(nh, 1)) if lam is None else np.ascontiguousarray(lam, float)
    pi = np.zeros(nx) if pi is None else np.ascontiguousarray(pi, float)
    out = dict(
        cost=np.zeros(1), cost_grad=np.zeros(nz), cost_hess=np.zeros((nz, nz)), h=np.zeros(max(nh, 1)),
        h_jac=np.zeros((max(nh, 1), nz)), h_hess=np.zeros((nz, nz)), xnext=np.zeros(nx), AB=np.zeros((nx, nz)),
        dyn_hess=np.zeros((nz, nz)),
    )
    rc = lib().oracle_eval_stage(
        ctypes.byref(desc), _p(z), _p(p), _p(lam), _p(pi), _p(out["cost"]), _p(out["cost_grad"]), _p(out["cost_hess"]),
        _p(out["h"]), _p(out["h_jac"]), _p(out["h_hess"]), _p(out["xnext"]), _p(out["AB"]), _p(out["dyn_hess"]),
    )
    if rc != 0:
        raise RuntimeError(f"oracle_eval_stage failed: {rc}")
    out["cost"] = float(out["cost"][0])
    out["h"] = out["h"][:nh]
    out["h_jac"] = out["h_jac"][:nh]
    return out


def state_size(spec):
    return int(lib().oracle_state_size(ctypes.byref(spec.desc())))


def solve_batch(spec, xinit, x0, params, nthreads=0, iterates=False, state=None):
    """Solve P problems.  Arrays are [P, ...] in the reference's flat layouts.
    Returns (xtraj[P,N+1,nx], utraj[P,N,nu], info(list of dict))  (+ iterates[P,n_sqp,N+1,nz]).
    `state` [P, state_size] (in/out): persistent NLP multipliers of the P solvers (Solver::operator= semantics)."""
    from mpc_planner_b200.problem import SolveInfo

    desc = spec.desc()
    N, nx, nu = spec.N, spec.nx, spec.nu
    nz = nx + nu
    xinit = np.ascontiguousarray(xinit, float).reshape(-1, nx)
    P = xinit.shape[0]
    x0 = np.ascontiguousarray(x0, float).reshape(P, -1)
    params = np.ascontiguousarray(params, float).reshape(P, -1)
    assert x0.shape[1] == nz * (N + 1) and params.shape[1] == spec.npar * N
    xtraj = np.zeros((P, N + 1, nx))
    utraj = np.zeros((P, N, nu))
    info = (SolveInfo * P)()
    if state is not None:
        assert not iterates and state.flags.c_contiguous and state.dtype == np.float64 and state.shape == (P, state_size(spec))
        for i in range(P):
            rc = lib().oracle_solve_stateful(ctypes.byref(desc), _p(xinit[i]), _p(x0[i]), _p(params[i]), _p(xtraj[i]), _p(utraj[i]),
                                             ctypes.byref(info[i]), _p(state[i]))
            assert rc == 0
    elif iterates:
        its = np.zeros((P, spec.n_sqp, N + 1, nz))
        for i in range(P):
            rc = lib().oracle_solve(ctypes.byref(desc), _p(xinit[i]), _p(x0[i]), _p(params[i]), _p(xtraj[i]), _p(utraj[i]),
                                    ctypes.byref(info[i]), _p(its[i]))
            assert rc == 0
    else:
        rc = lib().oracle_solve_batch(ctypes.byref(desc), P, _p(xinit), _p(x0), _p(params), _p(xtraj), _p(utraj), info,
                                      int(nthreads))
        assert rc == 0
    infos = [
        dict(pobj=i.pobj, res_eq=i.res_eq, res_stat_qp=i.res_stat_qp, res_comp_qp=i.res_comp_qp, exit_code=i.exit_code,
             qp_status=i.qp_status, sqp_iter=i.sqp_iter, qp_iter=i.qp_iter)
        for i in info
    ]
    if iterates:
        return xtraj, utraj, infos, its
    return xtraj, utraj, infos


def find_best(infos):
    best, best_obj = -1, 1e50
    for i, inf in enumerate(infos):
        if inf["exit_code"] == 1 and inf["pobj"] < best_obj:
            best, best_obj = i, inf["pobj"]
    return best


def max_threads():
    return int(lib().oracle_max_threads())
