"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
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


def test_petsc_shim_exports_the_reference_api_of_the_sampling_path():
    """nm -D of the shim object must contain every function the reference's headers export for the sampling path
    (include/parmgmc/parmgmc.h:33-44, mc_sor.h:21-30, iact.h:14-15, pc/pc_mcgibbs.h:16-18, pc/pc_sorgibbs.h:15,
    pc/pc_gamgmc.h:15-18, pc/pc_chols.h:15-16, pc/woodbury.h:15-17).  Out of scope: ms.h, obs.h, problems.h,
    stats.h (DM / FE set-up and host-side post-processing) and pc/pc_parsor.h."""
    import subprocess
    root = os.path.dirname(os.path.dirname(pmg.HEADER_PATH))
    shim = os.path.join(root, "shim", "petsc")
    subprocess.check_call(["make", "-s", "-C", shim])
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(shim, "build", "libparmgmc_b200_petsc.so")], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    wanted = """PARMGMC_CLASSID MULTICOL_SOR VEC_SET_RANDOM_NORMAL ParMGMCInitialize ParMGMCFinalize PCRegisterSetSampleCallback
        PCSetSampleCallback ParMGMCGetPetscRandom VecSetRandomStandardNormal
        MCSORCreate MCSORSetUp MCSORDestroy MCSORApply MCSORSetOmega MCSORSetSweepType MCSORGetSweepType MCSORGetISColoring
        MCSORGetNumColors MCSORBuildLRCCorrection Autocorrelation IACT
        PCCreate_MulticolorGibbs PCMulticolorGibbsSetOmega PCMulticolorGibbsSetSweepType PCCreate_SORGibbs
        PCGAMGMCSetLevels PCCreate_GAMGMC PCGAMGMCGetInternalPC PCGAMGMCSetInternalPC
        PCCreate_CholSampler PCCholSamplerSetIsCoarseGAMG PCCreate_Woodbury PCWoodburySetSolver PCWoodburySetSampler""".split()
    missing = [w for w in wanted if w not in exported]
    assert not missing, missing
    for exe in ("host_ex1", "host_ex5", "host_api"):
        assert os.path.exists(os.path.join(shim, "build", exe))


# ---- host logic of the fused 3D sweep: the work list (pmg_plan_sweep3d; csrc/stencil_op.cu sweep3d_plan) -----------------------
@pytest.mark.parametrize("dims,slab,bz,nw", [
    ((512, 512, 512), None, 64, 16),      # config 3: 4 full strips + a 32-column narrow strip, thin z-edge bands
    ((513, 513, 513), None, 64, 16),      # config 4's per-GPU grid
    ((150, 40, 48), (24, 48), 64, 16),    # upper slab of a 2-rank split: thin band only at the grid's last planes
    ((150, 40, 48), (0, 24), 64, 16),
    ((57, 31, 12), None, 64, 16),         # last strip wider than 56 columns: no narrow tiles
    ((36, 30, 5), None, 64, 16),          # short in z: plain bands
    ((50, 64, 16), None, 64, 8),
])
def test_sweep3d_work_list_tiles_the_slab_exactly_once(dims, slab, bz, nw):
    import parmgmc_b200 as pmg
    nx, ny, nz = dims
    slo, shi = slab if slab else (0, nz)
    items = pmg.plan_sweep3d(nx, ny, nz, slab, bz, nw)
    assert items.shape[1] == 5 and len(items) > 0
    cover = np.zeros((shi - slo, ny, (nx + 119) // 120), np.int32)
    nstrips = (nx + 119) // 120
    for s, ya, ka, kb, narrow in items:
        rows = (2 * nw - 2) if narrow else (nw - 2)
        assert 0 <= s < nstrips and 0 <= ya < ny and slo <= ka < kb <= shi
        if narrow:  # only the last, short strip, and only where every plane touched has both z neighbours
            assert s == nstrips - 1 and nx - 120 * s <= 56 and ka - 1 >= 1 and kb <= nz - 2
        cover[ka - slo:kb - slo, ya:min(ya + rows, ny), s] += 1
    assert cover.min() == 1 and cover.max() == 1
    # planes without a z neighbour sit in bands of at most 4 planes when the slab is long enough
    if shi - slo > 10:
        for s, ya, ka, kb, narrow in items:
            if ka < 2 or kb > nz - 2:
                assert kb - ka <= 4
    if (nx % 120) and nx - 120 * (nstrips - 1) <= 56 and ny >= 2 * nw - 2 and shi - slo > 10:
        assert items[:, 4].sum() > 0


def test_sweep3d_work_list_rejects_bad_geometry():
    import parmgmc_b200 as pmg
    with pytest.raises(pmg.PMGError):
        pmg.plan_sweep3d(64, 64, 64, (10, 5))
    with pytest.raises(pmg.PMGError):
        pmg.plan_sweep3d(64, 64, 64, None, 64, 2)


def test_sass_backs_the_design_claims():
    """DESIGN.md's hardware claims, checked in the built library's SASS (cuobjdump; no GPU needed): sm_100a only, TMA tensor loads and
    mbarriers in the fused sweeps, setmaxnreg in the warp-specialised ones, cp.async in the staged 27-point sweep, system-scope
    accesses in the peer-memory halo kernel, griddepcontrol.wait for the programmatic dependent launches."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    elf = subprocess.run(["cuobjdump", "-lelf", pmg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), elf
    sass = subprocess.run(["cuobjdump", "-sass", pmg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    per = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = []
        elif cur is not None and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            per[cur].append(line)

    def count(kernel_pat, instr_pat):
        return sum(len([ln for ln in lines if re.search(instr_pat, ln)]) for k, lines in per.items() if re.search(kernel_pat, k))

    assert count(r"sweep2d_kernel", r"\bUTMALDG") > 0 and count(r"sweep2d_kernel", r"\bSYNCS") > 0
    assert count(r"sweep3d_ws_kernel", r"\bUTMALDG") > 0 and count(r"sweep3d_ws_kernel", r"\bUSETMAXREG") == 2
    assert count(r"box2d_kernel", r"\bUTMALDG") > 0 and count(r"box2d_kernel", r"\bACQBULK") > 0
    assert count(r"box3_sweep_smem_kernel", r"\bLDGSTS") > 0 and count(r"box3_sweep_smem_kernel", r"\bLDS\.128") > 0
    assert count(r"halo_kernel", r"\.SYS\b") > 0


# ---- host logic of the fused 2D sweep on slabs: the work list with the thin boundary bands of the overlapped exchange -------------
@pytest.mark.parametrize("restrict_mode", [False, True])
@pytest.mark.parametrize("nx,ny,slab,by", [
    (4097, 32769, (0, 4096), 82),          # first rank of the 8-GPU headline: a neighbour above only
    (4097, 32769, (12288, 16384), 82),     # a middle rank: thin bands at both ends
    (4097, 32769, (28672, 32769), 82),     # last rank (owns the closing row)
    (129, 257, (96, 128), 4),              # bench.py's parity case on 8 ranks: 32-row slabs
    (129, 257, (224, 257), 4),
    (300, 140, (40, 70), 6),               # 30 rows: too short for thin bands
    (4097, 4097, None, 82),                # one GPU
])
def test_sweep2d_work_list_tiles_the_slab_exactly_once(nx, ny, slab, by, restrict_mode):
    import parmgmc_b200 as pmg
    slo, shi = slab if slab else (0, ny)
    for overlap in (True, False):
        items, nohalo = pmg.plan_sweep2d(nx, ny, slab, by, restrict_mode, overlap)
        nstrips = (nx + 119) // 120
        cover = np.zeros((shi - slo, nstrips), np.int32)
        for s, ja, jb in items:
            assert 0 <= s < nstrips and slo <= ja < jb <= shi
            cover[ja - slo:jb - slo, s] += 1
        assert cover.min() == 1 and cover.max() == 1
        lo, hi = (5, 3) if restrict_mode else (3, 1)
        reads_ghost = [(ja - lo < slo and slo > 0) or (jb + hi >= shi and shi < ny) for _, ja, jb in items]
        # the tiles that read no ghost row come first, and `nohalo` counts exactly them
        assert not any(reads_ghost[:nohalo]) and all(reads_ghost[nohalo:])
        if slab is None:
            assert nohalo == len(items)
        elif overlap and shi - slo >= 32:
            # every tile that reads ghost rows is an 8-row band at a slab end that has a neighbour: one per strip and such end
            ends = (1 if slo > 0 else 0) + (1 if shi < ny else 0)
            assert len(items) - nohalo == ends * nstrips
            assert all(jb - ja == 8 for (_, ja, jb), g in zip(items, reads_ghost) if g)
