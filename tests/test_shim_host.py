"""The drop-in boundary exercised from C: shim/petsc/host_ex1.c is an examples/ex1.c-shaped host program that only uses
ParMGMCInitialize, -pc_type style PC creation, option keys, PCSetSampleCallback and PCApplyRichardson; underneath it the
PETSc shim forwards to libparmgmc_b200.so.  Acceptance is the reference's own: relative error of the sample mean <= 0.02
(examples/ex1.c:20-44, :135)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim", "petsc")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-s", "-C", SHIM])
    return os.path.join(SHIM, "build", "host_ex1")


@pytest.mark.parametrize("args", [
    ["mcgibbs", "1000000", "0.02"],                                        # examples/ex1.c:20, the reference's own sample count and tolerance
    ["mcgibbs", "200000", "0.05", "-pc_mcgibbs_backward", "", "-pc_mcgibbs_omega", "1.2"],   # :21 (fewer samples, 1/sqrt(N)-scaled tolerance)
    ["mcgibbs", "200000", "0.05", "-pc_mcgibbs_symmetric", ""],            # :22
    ["sorgibbs", "200000", "0.05"],                                        # :23
    ["cholsampler", "200000", "0.05"],                                     # :26
    ["gamgmc", "100000", "0.07", "-pc_gamgmc_mg_type", "mg", "-gamgmc_pc_mg_levels", "2", "-pc_b200_grid", "9,9"],  # :41
    ["gamgmc", "100000", "0.07", "-pc_gamgmc_mg_type", "mg", "-gamgmc_pc_mg_levels", "3", "-pc_b200_grid", "9,9", "-gamgmc_mg_coarse_pc_type", "mcgibbs", "-gamgmc_mg_coarse_ksp_max_it", "2"],  # :44
])
def test_ex1_shaped_c_host_program(exe, args):
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "relative mean error" in r.stdout


@pytest.mark.parametrize("args", [[], ["17", "1.3"]])
def test_ex5_shaped_c_host_program_typed_mcsor_api(exe, args):
    """examples/ex5.c through the PETSc-typed MCSOR API of the shim (forward + backward == symmetric to 1e-15), the same
    operator as a one-rank MATMPIAIJ, MCSOR on a MATLRC operator against MCSORBuildLRCCorrection (examples/ex3.c -with_lr),
    MCSORGetISColoring / GetNumColors."""
    r = subprocess.run([os.path.join(SHIM, "build", "host_ex5")] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_ex5 ok" in r.stdout


def test_c_host_program_rest_of_the_exported_api(exe):
    """ParMGMCGetPetscRandom / VecSetRandomStandardNormal, PC woodbury (option keys and PCWoodburySet*), mcgibbs on MATLRC,
    PCGAMGMCGet/SetInternalPC + "PCMGGetLevels_C", cholsampler presolve / postsolve callback rules, IACT / Autocorrelation."""
    r = subprocess.run([os.path.join(SHIM, "build", "host_api"), "100000", "0.05"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_api ok" in r.stdout
