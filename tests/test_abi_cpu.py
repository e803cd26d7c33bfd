"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import parmgmc_b200 as pmg


def header_functions():
    src = open(pmg.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmg_[a-z0-9_]+)\s*\(", src)) - {"pmg_sample_cb", "pmg_ctx_deleter"})


def test_library_exports_every_declared_symbol():
    assert os.path.exists(pmg.LIB_PATH), "build the extension first (__graft_entry__.build())"
    L = ctypes.CDLL(pmg.LIB_PATH)
    names = header_functions()
    assert len(names) >= 45
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/parmgmc_b200.h but not exported"
    # the python mirror binds exactly the declared surface
    assert sorted(pmg.SIGNATURES) == names


def test_no_torch_types_in_the_boundary():
    src = open(pmg.HEADER_PATH).read()
    assert "#include <torch" not in src and "at::" not in src and "#include <petsc" not in src
    assert 'extern "C"' in src


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert pmg.device_count() == 0
    with pytest.raises(pmg.PMGError) as e:
        pmg.Context(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(pmg.HEADER_PATH)
    pkg = os.path.join(os.path.dirname(root), "parmgmc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "import oracle" not in txt and "oracle/oracle.h" not in txt.replace("FP contract: oracle/oracle.h", ""), f


def test_petsc_shim_builds_and_fails_loudly_without_device():
    """shim/petsc: the PETSc-facing registration layer + an ex1-shaped C host program link against the C ABI
    (PETSc objects from oracle/petsc_stub, since PETSc is not in this image); without a GPU the program must stop
    with the no-CPU-fallback error instead of computing anything."""
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(pmg.HEADER_PATH))
    shim = os.path.join(root, "shim", "petsc")
    subprocess.check_call(["make", "-s", "-C", shim])
    exe = os.path.join(shim, "build", "host_ex1")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is tests/test_shim_host.py")
    r = subprocess.run([exe, "mcgibbs", "10", "0.02"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
