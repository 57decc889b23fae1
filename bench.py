#!/usr/bin/env python
"""bench.py -- MPC solves/sec of the batched SQP hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic scenes: S scenes x P=8 T-MPC++
guesses of BASELINE.json configs[1] (N=20, 8 dynamic obstacles, nx=5, nu=2, npar=135, nh=16),
i.e. S*8 full `Solver::solve()` calls (10 SQP-RTI iterations each) plus the per-scene best-guess
selection.  Per GPU work is fixed as N grows (weak scaling; scenes are independent, so there is no
data-path collective -- SURVEY.md 8e).

Prints ONE JSON line (rank 0):
  value   solves/s with inputs resident in HBM (device-pointer C-ABI call, CUDA-event timed)
  e2e     solves/s through the host-buffer C-ABI call (pinned host buffers, H2D + D2H inside)
  roofline / fp64   achieved vs measured peaks for the SQP kernel
  cpu_baseline      the CPU oracle (acados-equivalent restatement) on this box's host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mpc_planner_b200 import problem as pb  # noqa: E402
from mpc_planner_b200 import scenes  # noqa: E402

METRIC = "MPC solves/sec (N=20, 8 obs, P guesses)"
UNIT = "solves/s"
GUESSES = 8
ALG_BYTES_PER_SOLVE = 22816 + 1224  # xinit + x0 + all_parameters in, xtraj + utraj + info out


def flops_per_solve(spec, mean_qp_iter, n_sqp):
    """Algorithmic FLOP model of one solve (DESIGN.md 'FLOP model'): counted from the kernel's
    arithmetic, FMA = 2 flops, sparse structure not discounted."""
    nx, nu, nz, nh, N = spec.nx, spec.nu, spec.nz, spec.nh, spec.N
    nrow = 2 * nz + nh
    f_factor = (nz * (nz + 1) // 2) * (nx * nx + nx) * 2 + (nx * (nx + 1) // 2) * 8 + nx * nx * 2
    f_solve = 2 * ((nz * nx + nx * nu * 2) * 2 + (nx * nu * 2 + nx * nz + nx * nx) * 2)
    f_rows = nrow * (14 + 2 * 12 + 10 + 12)
    f_resid = (nz * nz + 2 * nx * nz) * 2 + nrow * 8
    f_ipm = f_factor + f_solve + f_rows + f_resid
    f_lin = 3000 + 60 * nh + 2 * nz ** 3 // 3
    return N * (n_sqp * f_lin + mean_qp_iter * f_ipm)


def make_batch(spec, n_scenes, seed):
    """n_scenes x 8 guesses; a pool of distinct scenes is generated and tiled (generation is
    host-side numpy and not part of the measured path)."""
    pool = min(n_scenes, 64)
    base = scenes.make_scene_batch(spec, pool, GUESSES, seed=seed)
    reps = (n_scenes + pool - 1) // pool
    out = {}
    for k in ("xinit", "x0", "params"):
        out[k] = np.ascontiguousarray(np.tile(base[k], (reps, 1))[: n_scenes * GUESSES])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device_index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for (t, line) in self.lines:
            if not (t0 - 0.05 <= t <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: use everything collected
            for (_, line) in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_solve_pass(spec, batch, n_solves, threads):
    """One OpenMP pass of the oracle over the first n_solves problems; returns (seconds, infos)."""
    from oracle import oracle

    xi, x0, par = batch["xinit"][:n_solves], batch["x0"][:n_solves], batch["params"][:n_solves]
    t = time.perf_counter()
    _, _, info = oracle.solve_batch(spec, xi, x0, par, nthreads=threads)
    return time.perf_counter() - t, info


def calibrate_threads(spec, batch):
    """Pick the OpenMP thread count that solves a small sample fastest (hyper-threads / cgroup quotas
    make 'all logical CPUs' a bad default on some boxes)."""
    cores = usable_cores()
    cands = sorted({cores, max(1, cores // 2)}, reverse=True)
    best_t, best_rate = cands[0], 0.0
    for t in cands:
        n = len(batch["xinit"])  # the whole pool: small samples are dominated by the slowest solve
        cpu_solve_pass(spec, batch, min(n, t), t)  # spin the team up
        dt, _ = cpu_solve_pass(spec, batch, n, t)
        if n / dt > best_rate:
            best_t, best_rate = t, n / dt
    return best_t


def cpu_arm(spec, target_seconds=12.0, pool_scenes=64, seed=2024):
    """The CPU arm: acados-equivalent oracle, OpenMP over the batch, bounded sample of the workload."""
    batch = make_batch(spec, pool_scenes, seed=seed)
    n = len(batch["xinit"])
    threads = calibrate_threads(spec, batch)
    total, t0 = 0, time.perf_counter()
    while True:
        cpu_solve_pass(spec, batch, n, threads)
        total += n
        if time.perf_counter() - t0 >= target_seconds:
            break
    dt = time.perf_counter() - t0
    return total / dt, threads, total, dt


def run_reference(args, spec):
    """--impl reference: the CPU implementation of the path (acados itself cannot be built here:
    un-pinned third-party HEAD, absent from /root/reference; the arm runs the acados-equivalent
    oracle port with the host threads that serve it best).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool = int(os.environ.get("MPC_BENCH_REF_SCENES", "64"))
    batch = make_batch(spec, pool, seed=2024)
    sample = len(batch["xinit"])
    threads = calibrate_threads(spec, batch)
    for _ in range(args.warmup):
        cpu_solve_pass(spec, batch, min(sample, 2 * threads), threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_solve_pass(spec, batch, sample, threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(sample // GUESSES), "sample": f"{sample} solves per step (bounded CPU sample of the same workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} solves/step x {args.steps} steps in {dt:.2f} s, OpenMP over the batch ({threads} threads of {usable_cores()} usable CPUs), acados-equivalent CPU restatement (oracle/), not acados"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(n_scenes):
    return (f"BASELINE configs[1]: T-MPC++ P=8 guesses, N=20, 8 dynamic obstacles (unicycle nx=5 nu=2, npar=135, nh=16), "
            f"{n_scenes} scenes x 8 guesses = {n_scenes * GUESSES} solves per step per GPU, 10 SQP-RTI iterations per solve")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenes", type=int, default=int(os.environ.get("MPC_BENCH_SCENES", "512")), help="scenes per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    spec = pb.config_tmpc()
    if args.impl == "reference":
        run_reference(args, spec)
        return

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from mpc_planner_b200 import binding

    S, P = args.scenes, args.scenes * GUESSES
    N, nx, nu, nz, npar = spec.N, spec.nx, spec.nu, spec.nz, spec.npar
    solver = binding.BatchSolver(spec, max_batch=P, device=local_rank)
    solver.set_group(GUESSES)

    # ---- inputs: NSETS distinct batches resident in HBM; rotating through them makes every step's
    # inputs cold in L2 (NSETS x 93 MB > 126 MB L2) ----
    NSETS = 4
    host_sets = [make_batch(spec, S, seed=1000 * rank + 17 * j + 1) for j in range(NSETS)]
    dev = torch.device("cuda", local_rank)
    d_sets = []
    for hs in host_sets:
        d_sets.append({k: torch.from_numpy(hs[k]).to(dev) for k in ("xinit", "x0", "params")})
    d_xtraj = torch.empty((P, N + 1, nx), dtype=torch.float64, device=dev)
    d_utraj = torch.empty((P, N, nu), dtype=torch.float64, device=dev)
    d_info = torch.empty((P, binding.INFO_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_best = torch.empty((S,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step_device(j):
        ds = d_sets[j % NSETS]
        solver.solve_device(P, ds["xinit"].data_ptr(), ds["x0"].data_ptr(), ds["params"].data_ptr(), d_xtraj.data_ptr(),
                            d_utraj.data_ptr(), d_info.data_ptr(), d_best.data_ptr(), stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for j in range(args.warmup):
        step_device(j)
    barrier()

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = solver.launches
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    for j in range(args.steps):
        step_device(j)
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    gpu_launches = solver.launches - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * P / (ms_per_step * 1e-3)

    # per-solve statistics of the last step (for the FLOP model)
    info = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=binding.INFO_DTYPE)
    mean_qp_iter = float(info["qp_iter"].mean())
    frac_success = float((info["exit_code"] == 1).mean())

    # ---- kernel-only duration (CUDA events on the launching stream around the SQP kernel) ----
    solver.set_timing(True)
    for j in range(max(3, min(args.steps, 5))):
        step_device(j)
    torch.cuda.synchronize(dev)
    kernel_ms, n_k = solver.kernel_ms()
    solver.set_timing(False)

    # ---- e2e: host buffers through the C ABI (pinned), H2D + D2H inside the timed region ----
    lib = binding.load_library()

    def pinned(shape, dtype=np.float64):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = ctypes.c_void_p()
        rc = lib.mpcgpu_host_alloc(ctypes.byref(ptr), nbytes)
        if rc != 0:
            raise RuntimeError("mpcgpu_host_alloc failed")
        buf = (ctypes.c_char * nbytes).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), ptr

    pins = []
    h_in = []
    for hs in host_sets[:2]:
        d = {}
        for k in ("xinit", "x0", "params"):
            arr, ptr = pinned(hs[k].shape)
            arr[...] = hs[k]
            d[k] = arr
            pins.append(ptr)
        h_in.append(d)
    h_xtraj, p1 = pinned((P, N + 1, nx))
    h_utraj, p2 = pinned((P, N, nu))
    h_info, p3 = pinned((P * binding.INFO_DTYPE.itemsize,), np.uint8)
    h_best, p4 = pinned((S,), np.int32)
    pins += [p1, p2, p3, p4]
    _dp = ctypes.POINTER(ctypes.c_double)

    def step_host(j):
        hs = h_in[j % 2]
        rc = lib.mpcgpu_solve_host(solver.handle, P, hs["xinit"].ctypes.data_as(_dp), hs["x0"].ctypes.data_as(_dp),
                                   hs["params"].ctypes.data_as(_dp), h_xtraj.ctypes.data_as(_dp), h_utraj.ctypes.data_as(_dp),
                                   ctypes.cast(h_info.ctypes.data, ctypes.POINTER(pb.SolveInfo)),
                                   h_best.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        if rc != 0:
            raise RuntimeError(f"mpcgpu_solve_host: {rc}")

    # The T-MPC++ call: one parameter block per control step (scene), the per-guess blocks and topology halfspaces are
    # built on the device (mpcgpu_solve_tmpc_host) -- the call GuidanceConstraints::optimize makes.
    t_in = []
    for hs in host_sets[:2]:
        d = {}
        src = {"xinit": hs["xinit"].reshape(S, GUESSES, -1)[:, 0], "x0": hs["x0"],
               "params": hs["params"].reshape(S, GUESSES, -1)[:, 0]}
        for k, v in src.items():
            arr, ptr = pinned(v.shape)
            arr[...] = v
            d[k] = arr
            pins.append(ptr)
        t_in.append(d)
    guided = np.array([0] + [1] * (GUESSES - 1), dtype=np.int32)

    def step_tmpc(j):
        hs = t_in[j % 2]
        rc = lib.mpcgpu_solve_tmpc_host(solver.handle, S, GUESSES, hs["xinit"].ctypes.data_as(_dp), hs["x0"].ctypes.data_as(_dp),
                                        hs["params"].ctypes.data_as(_dp), guided.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                        float(scenes.ROBOT_RADIUS), h_xtraj.ctypes.data_as(_dp), h_utraj.ctypes.data_as(_dp),
                                        ctypes.cast(h_info.ctypes.data, ctypes.POINTER(pb.SolveInfo)),
                                        h_best.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        if rc != 0:
            raise RuntimeError(f"mpcgpu_solve_tmpc_host: {rc}")

    def time_host(fn):
        for j in range(2):
            fn(j)
        barrier()
        t0 = time.perf_counter()
        for j in range(args.steps):
            fn(j)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * P * args.steps / float(te.item())

    e2e_generic = time_host(step_host)
    e2e_value = time_host(step_tmpc)
    input_bytes_resident = sum(int(h_in[0][k].nbytes) for k in ("xinit", "x0", "params"))
    h2d = sum(int(t_in[0][k].nbytes) for k in ("xinit", "x0", "params")) + int(guided.nbytes)
    d2h = int(h_xtraj.nbytes + h_utraj.nbytes + h_info.nbytes + h_best.nbytes)
    for ptr in pins:
        lib.mpcgpu_host_free(ptr)

    # ---- latency of ONE control step (P = 8 guesses, the reference's real-time use) ----
    lat_solver = binding.BatchSolver(spec, max_batch=GUESSES, device=local_rank)
    one = {k: host_sets[0][k][:GUESSES] for k in ("xinit", "x0", "params")}
    for _ in range(3):
        lat_solver.solve(one["xinit"], one["x0"], one["params"])
    lat = []
    for _ in range(10):
        t0 = time.perf_counter()
        lat_solver.solve(one["xinit"], one["x0"], one["params"])
        lat.append(time.perf_counter() - t0)
    lat_solver.close()
    latency_ms = 1e3 * float(np.median(lat))

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        fp64_peak = float(lib.mpcgpu_measure_fp64_gflops(local_rank))
        achieved_gbs = ALG_BYTES_PER_SOLVE * P / (kernel_ms * 1e-3) / 1e9
        fl = flops_per_solve(spec, mean_qp_iter, spec.n_sqp)
        achieved_gflops = fl * P / (kernel_ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            v, threads, n_cpu, dt = cpu_arm(spec, target_seconds=float(os.environ.get("MPC_BENCH_CPU_SECONDS", "12")))
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{n_cpu} solves of the same workload in {dt:.1f} s, OpenMP over the batch ({threads} threads of {usable_cores()} usable CPUs); acados-equivalent CPU restatement (oracle/), not acados"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one sqp_kernel launch (ncu --set full)
            tj = json.load(open(tpath))
            if tj.get("grid") == P:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(S), "l2": f"inputs larger than L2: rotating {NSETS} resident input sets of {input_bytes_resident / 1e6:.0f} MB",
                       "solves_per_step_per_gpu": P, "success_fraction": frac_success, "mean_ipm_iterations_per_solve": mean_qp_iter,
                       "control_step_latency_ms_P8_host_call": latency_ms,
                       "e2e_call": "mpcgpu_solve_tmpc_host (one parameter block per control step; per-guess blocks + topology halfspaces built on the device)",
                       "e2e_generic_call_solves_per_s": e2e_generic,
                       "e2e_generic_call": "mpcgpu_solve_host (per-guess parameter blocks packed on the host)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": ALG_BYTES_PER_SOLVE * P, "peak_source": peak_src, "kernel": "sqp_kernel", "kernel_ms": kernel_ms,
                         "note": "latency/FP64-issue bound by design (SURVEY.md 8d); see fp64"},
            "fp64": {"achieved": achieved_gflops, "peak": fp64_peak, "unit": "GFLOP/s", "frac": achieved_gflops / fp64_peak if fp64_peak > 0 else None,
                     "flops_per_solve": fl, "peak_source": "measured in this run (dependent-free DFMA chains, all SMs)"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    solver.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
