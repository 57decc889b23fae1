"""Throughput of the SQP kernel against the number of co-resident solves per SM (development aid, profiles/r02_residency.md).

The shared-memory image of one solve fixes the residency (cfg 2, N = 20: 56.6 KB -> 4 solves per SM).  To learn what a smaller
image would buy BEFORE restructuring the kernel, the same kernel is built for a short horizon (N = 9: 27 KB -> up to 8 solves
per SM) and the residency is lowered step by step with unused dynamic shared memory (MPCGPU_SMEM_PAD, -DMPC_DEV_KNOBS builds).

    python scripts/residency_sweep.py build      # here (nvcc cross-compiles): build/variants/*.so
    python scripts/residency_sweep.py run        # on the GPU box
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build", "variants")

# (tag, N, extra nvcc flags)
VARIANTS = [
    ("n9_t64_r8", 9, ["-DMPC_DEV_KNOBS", "-DMPC_NT_OVERRIDE=64", "-DMPC_MIN_CTAS=8"]),
    ("n13_t64_r6", 13, ["-DMPC_DEV_KNOBS", "-DMPC_NT_OVERRIDE=64", "-DMPC_MIN_CTAS=6"]),
    ("n9_t128_r4", 9, ["-DMPC_DEV_KNOBS", "-DMPC_NT_OVERRIDE=128", "-DMPC_MIN_CTAS=4"]),
    ("n20_t64_r4", 20, ["-DMPC_DEV_KNOBS", "-DMPC_NT_OVERRIDE=64", "-DMPC_MIN_CTAS=4"]),
    ("n20_t128_r4", 20, ["-DMPC_DEV_KNOBS", "-DMPC_NT_OVERRIDE=128", "-DMPC_MIN_CTAS=4"]),
]
SMEM_SM = 233472


def spec_of(N):
    from mpc_planner_b200 import problem as pb

    s = pb.config_tmpc(N=N)
    s.name = "tmpc_n%d" % N
    return s


def build():
    from mpc_planner_b200 import generate

    for tag, N, flags in VARIANTS:
        d = os.path.join(OUT, tag)
        path = generate.build_solver_library(spec_of(N), d, verbose=True, extra_flags=flags, suffix="_" + tag)
        print(tag, "->", path)


def run():
    import numpy as np

    from mpc_planner_b200 import binding, scenes

    P = 4096
    results = []
    for tag, N, flags in VARIANTS:
        spec = spec_of(N)
        lib = os.path.join(OUT, tag, "libmpcgpu_%s_%s.so" % (spec.name, tag))
        batch = scenes.make_scene_batch_chunked(spec, P // 8, 8, seed=1)
        nz = spec.nz
        image = (N + 1) * 337 * 8  # SL<cfg 2>::TOTAL doubles: odd stride 337 per stage (sqp_core.cuh)
        for r in (8, 7, 6, 5, 4, 3, 2):
            os.environ.pop("MPCGPU_SMEM_PAD", None)
            s = binding.BatchSolver(spec, P, library=lib)
            r_nat = SMEM_SM // (image + 1024)
            if r > r_nat:
                s.close()
                continue
            # largest per-CTA footprint that still admits r CTAs (and not r + 1)
            pad = 0 if r == r_nat else max(0, (SMEM_SM // r - 1024 - 16) // 128 * 128 - image - 16)  # 128-byte allocation granularity, 16 B static
            os.environ["MPCGPU_SMEM_PAD"] = str(pad)
            s.solve(batch["xinit"], batch["x0"], batch["params"])
            s.set_timing(True)
            ms = []
            for _ in range(4):
                xt, ut, info, best = s.solve(batch["xinit"], batch["x0"], batch["params"])
                ms.append(s.kernel_ms()[0])
            ok = sum(i["exit_code"] == 1 for i in info)
            qp = float(np.mean([i["qp_iter"] for i in info]))
            rec = dict(variant=tag, N=N, image_bytes=image, ctas_per_sm=r, pad=pad, kernel_ms=min(ms), solves_per_s=P / (min(ms) * 1e-3), ok=ok, qp_iter=qp)
            results.append(rec)
            print(json.dumps(rec), flush=True)
            s.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "residency_sweep.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    (build if sys.argv[1:] == ["build"] else run)()
