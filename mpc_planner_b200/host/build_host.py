"""Build the C++ host layer (MPCPlanner::Solver over the C ABI) and its test for one problem."""
import os
import subprocess

from .. import build as libbuild
from .. import problem as pb
from ..generate import generate_solver

HERE = os.path.dirname(os.path.abspath(__file__))


def build_test(name="cfg2", out_dir=None):
    spec = pb.CONFIGS[name]()
    out_dir = out_dir or os.path.join(HERE, "generated", name)
    generate_solver(spec, out_dir)
    exe = os.path.join(out_dir, "test_solver")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I" + out_dir, "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(libbuild.ROOT, "include"),
           os.path.join(HERE, "src", "cuda_solver_interface.cpp"), os.path.join(HERE, "test", "test_solver.cpp"),
           "-L" + libbuild.LIB_DIR, "-lmpcgpu", "-Wl,-rpath," + libbuild.LIB_DIR, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def build_modules_test(name="cfg2", out_dir=None):
    """The module layer (parameter packing + batched T-MPC++ loop) and its test driver."""
    spec = pb.CONFIGS[name]()
    out_dir = out_dir or os.path.join(HERE, "generated", name)
    generate_solver(spec, out_dir)
    exe = os.path.join(out_dir, "test_modules")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I" + out_dir, "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(libbuild.ROOT, "include"),
           os.path.join(HERE, "src", "cuda_solver_interface.cpp"), os.path.join(HERE, "src", "modules.cpp"),
           os.path.join(HERE, "test", "test_modules.cpp"), "-L" + libbuild.LIB_DIR, "-lmpcgpu", "-Wl,-rpath," + libbuild.LIB_DIR, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def build_planner_test(name="cfg2", out_dir=None, nvtx=False):
    """Planner::solveMPC (host/src/planner.cpp) + the closed-loop driver."""
    spec = pb.CONFIGS[name]()
    out_dir = out_dir or os.path.join(HERE, "generated", name)
    generate_solver(spec, out_dir)
    exe = os.path.join(out_dir, "test_planner_nvtx" if nvtx else "test_planner")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    extra = ["-DMPC_PLANNER_NVTX", "-I/usr/local/cuda/include"] if nvtx else []  # PROFILE_SCOPE -> NVTX ranges (header-only nvtx3 + libdl)
    cmd = [cxx, "-O2", "-std=c++17"] + extra + ["-I" + out_dir, "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(libbuild.ROOT, "include"),
           os.path.join(HERE, "src", "cuda_solver_interface.cpp"), os.path.join(HERE, "src", "modules.cpp"), os.path.join(HERE, "src", "planner.cpp"),
           os.path.join(HERE, "test", "test_planner.cpp"), "-L" + libbuild.LIB_DIR, "-lmpcgpu", "-Wl,-rpath," + libbuild.LIB_DIR, "-ldl", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def build_scenario_test(name="cfg4", out_dir=None):
    """ScenarioConstraints (Safe Horizon MPC control flow, host/src/modules.cpp) + its test driver."""
    spec = pb.CONFIGS[name]()
    out_dir = out_dir or os.path.join(HERE, "generated", name)
    generate_solver(spec, out_dir)
    exe = os.path.join(out_dir, "test_scenario")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I" + out_dir, "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(libbuild.ROOT, "include"),
           os.path.join(HERE, "src", "cuda_solver_interface.cpp"), os.path.join(HERE, "src", "modules.cpp"),
           os.path.join(HERE, "test", "test_scenario.cpp"), "-L" + libbuild.LIB_DIR, "-lmpcgpu", "-Wl,-rpath," + libbuild.LIB_DIR, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def build_packing_test(name):
    """Parameter packing of the Gaussian / decomp modules (host/test/test_packing.cpp) for the stack `name`."""
    spec = pb.CONFIGS[name]()
    out_dir = os.path.join(HERE, "generated", name)
    generate_solver(spec, out_dir)
    exe = os.path.join(out_dir, "test_packing")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I" + out_dir, "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(libbuild.ROOT, "include"),
           os.path.join(HERE, "src", "cuda_solver_interface.cpp"), os.path.join(HERE, "src", "modules.cpp"),
           os.path.join(HERE, "test", "test_packing.cpp"), "-L" + libbuild.LIB_DIR, "-lmpcgpu", "-Wl,-rpath," + libbuild.LIB_DIR, "-o", exe]
    subprocess.check_call(cmd)
    return exe
