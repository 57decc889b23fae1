"""CPU-only checks of the drop-in boundary: the CUDA library loads without a GPU, exports every
symbol include/mpcgpu.h declares, knows the compiled problem shapes and refuses to run without a
device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from mpc_planner_b200 import binding
from mpc_planner_b200 import problem as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(binding.LIB_PATH):
        from mpc_planner_b200 import build

        build.build()
    return binding.load_library()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "mpcgpu.h")).read()
    declared = set(re.findall(r"\b(mpcgpu_[a-z0-9_]+)\s*\(", header))
    declared -= {"mpcgpu_solver"}
    assert declared == set(binding.EXPORTED_SYMBOLS), declared ^ set(binding.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_struct_layouts_match_header():
    # sizeof computed by hand from include/mpcgpu.h (ints 4, doubles 8, natural alignment)
    assert ctypes.sizeof(pb.SolveInfo) == 4 * 8 + 4 * 4
    d = pb.config_tmpc().desc()
    assert ctypes.sizeof(d) % 8 == 0
    assert pb.ProblemDesc.dt.offset == 32 and pb.ProblemDesc.lower_bound.offset % 8 == 0


@pytest.mark.parametrize("name", list(pb.CONFIGS))
def test_baseline_configs_have_kernels(lib, name):
    d = pb.CONFIGS[name]().desc()
    assert lib.mpcgpu_problem_supported(ctypes.byref(d)) == 1


def test_unsupported_shape_is_reported(lib):
    d = pb.config_tmpc(N=17).desc()
    assert lib.mpcgpu_problem_supported(ctypes.byref(d)) == 0
    h = ctypes.c_void_p()
    rc = lib.mpcgpu_create(ctypes.byref(d), 0, 4, ctypes.byref(h))
    assert rc != 0 and b"specialisation" in lib.mpcgpu_strerror(rc)


def test_no_cpu_fallback(lib):
    if lib.mpcgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(binding.MpcGpuError, match="no CUDA device"):
        binding.BatchSolver(pb.config_tmpc(), 8)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mpc_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
                assert "tests.emu" not in text and "libmpcgpu_emu" not in text, f


@pytest.mark.parametrize("name", list(pb.CONFIGS))
def test_descriptor_validation_accepts_the_shipped_stacks(lib, name):
    """mpcgpu_create range-checks every parameter index / block extent against npar before it looks for a device
    (a wrong descriptor must be refused at the ABI, not read out of bounds by the kernel)."""
    if lib.mpcgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    d = pb.CONFIGS[name]().desc()
    if not lib.mpcgpu_problem_supported(ctypes.byref(d)):
        pytest.skip("oracle-only stack")
    h = ctypes.c_void_p()
    rc = lib.mpcgpu_create(ctypes.byref(d), 0, 4, ctypes.byref(h))
    assert b"no CUDA device" in lib.mpcgpu_strerror(rc), lib.mpcgpu_strerror(rc)


@pytest.mark.parametrize("field,value", [("p_spline", 130), ("p_ego_disc_radius", -1), ("p_contour", 135), ("dt", 0.0), ("n_sqp", 0),
                                         ("npar", 100), ("p_acceleration", -1), ("qp_tol", 0.0)])
def test_descriptor_validation_rejects_out_of_range_indices(lib, field, value):
    d = pb.config_tmpc().desc()
    setattr(d, field, value)
    h = ctypes.c_void_p()
    rc = lib.mpcgpu_create(ctypes.byref(d), 0, 4, ctypes.byref(h))
    assert lib.mpcgpu_strerror(rc) == b"invalid argument", (field, lib.mpcgpu_strerror(rc))
    # a constraint block that runs past the end of the parameter vector
    d = pb.config_tmpc().desc()
    d.con_param[1] = d.npar - 10
    assert lib.mpcgpu_strerror(lib.mpcgpu_create(ctypes.byref(d), 0, 4, ctypes.byref(h))) == b"invalid argument"
    d = pb.config_tmpc().desc()
    d.lower_bound[0] = float("nan")
    assert lib.mpcgpu_strerror(lib.mpcgpu_create(ctypes.byref(d), 0, 4, ctypes.byref(h))) == b"invalid argument"
