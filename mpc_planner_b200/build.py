"""Build the CUDA shared library IN-TREE: mpc_planner_b200/lib/libmpcgpu.so (sm_100a only).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with
the repo snapshot."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmpcgpu.so")
DIST_LIB = os.path.join(LIB_DIR, "libmpcgpu_dist.so")  # multi-GPU best-trajectory reduction (links libnccl.so.2)
SOURCES = ["mpcgpu.cu"]
HEADERS = ["sqp_core.cuh", "stage_functions.cuh", "registry.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA backend cannot be built")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "mpcgpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, output=None, registry_header=None, stage_header=None, extra_flags=()):
    """Compile the library.  With `registry_header` (a file emitted by generate.generate_cuda_registry) the kernels
    are instantiated for that module stack only and the result is written to `output`."""
    out = output or LIB
    if output is None and not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(out), exist_ok=True)
    extra = []
    if registry_header is not None:
        extra = ["-DMPCGPU_GENERATED_REGISTRY=\"%s\"" % os.path.abspath(registry_header)]
    if stage_header is not None:  # stage functions emitted by codegen.py (MPCGPU_OBJ_GENERATED / MPCGPU_CON_GENERATED)
        extra += ["-DMPCGPU_GENERATED_STAGE=\"%s\"" % os.path.abspath(stage_header)]
    extra += list(extra_flags)  # development variants (scripts/): -DMPC_NT_OVERRIDE=..., -DMPC_MIN_CTAS=...
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-I" + os.path.join(ROOT, "include"), "-o", out] + [
        os.path.join(CSRC, s) for s in SOURCES
    ]
    subprocess.check_call(cmd, cwd=CSRC)
    return out


def build_dist(force=False):
    """libmpcgpu_dist.so: the NCCL best-trajectory reduction behind its own C ABI (include/mpcgpu_dist.h).  Links the
    single-GPU library (found next to it through $ORIGIN) and libnccl.so.2."""
    build()
    src = os.path.join(CSRC, "mpcgpu_dist.cu")
    deps = [src, os.path.join(ROOT, "include", "mpcgpu_dist.h"), os.path.join(ROOT, "include", "mpcgpu.h"), LIB]
    if not force and os.path.exists(DIST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(DIST_LIB) for d in deps):
        return DIST_LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-o", DIST_LIB, src, "-L" + LIB_DIR, "-lmpcgpu", "-lnccl",
                                        "-Xlinker", "-rpath=$ORIGIN"]
    subprocess.check_call(cmd, cwd=CSRC)
    return DIST_LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_dist(force=True))
