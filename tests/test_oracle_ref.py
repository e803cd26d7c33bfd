"""Pins the oracle restatement against the reference's OWN code.

oracle/_ref/libparmgmc_ref.so is /root/reference/src/{mc_sor.c, pc_mcgibbs.c, parmgmc.c} compiled unmodified against
the container-only PETSc API stub (oracle/petsc_stub).  It exists only where /root/reference does (this container);
elsewhere the committed golden vectors (tests/golden/, written by tests/golden/make_golden.py from the same library)
stand in for it -- see test_golden.py.

Tolerance: the reference writes `sum -= a*y` / `(1-omega)*y + idiag*sum` in plain C, so the compiler decides about
fused multiply-adds; the oracle (and the CUDA path) fix the contraction explicitly (oracle.h "Floating-point
contract").  The two therefore agree to rounding, not bitwise: 1e-13 relative here, against the 1e-12 of north_star.
"""
import numpy as np
import pytest

import oracle as orc
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
RTOL = 1e-13
SEED = 20260625


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("shape,kappa", [((129, 129), 10.0), ((129, 129), 1.0), ((33, 17), 1.0), ((9, 9), 10.0)])
@pytest.mark.parametrize("omega", [1.0, 1.2, 1.6])
@pytest.mark.parametrize("sweep", [orc.SOR_FORWARD, orc.SOR_BACKWARD, orc.SOR_SYMMETRIC])
def test_seq_sweep_one_colour(shape, kappa, omega, sweep):
    """MCSORApply_SEQAIJ with the reference's own 1-rank colouring (config 1: lexicographic Gauss-Seidel)."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, *shape, kappa=kappa)
    assert ref.mcsor_num_colors(A) == 1
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.mcsor_apply(A, b, y0.copy(), None, omega, sweep, nsweeps=3)
    mc = orc.MCSOR(A, None, omega, sweep)
    y = y0.copy()
    for _ in range(3):
        mc.apply(b, y)
    assert rel(y, y_ref) < RTOL


@pytest.mark.parametrize("dim,shape", [(2, (129, 129)), (2, (40, 23)), (3, (12, 9, 7))])
@pytest.mark.parametrize("omega,sweep", [(1.0, orc.SOR_FORWARD), (1.3, orc.SOR_SYMMETRIC), (0.7, orc.SOR_BACKWARD)])
def test_seq_sweep_injected_colouring(dim, shape, omega, sweep):
    """The same colouring injected into both sides: red-black, greedy and level-set colourings."""
    rng = np.random.default_rng(SEED + 1)
    A = orc.laplace(dim, *shape, kappa=1.0)
    for col in (orc.Coloring.parity(shape), orc.Coloring.greedy(A), orc.Coloring.levelset(A)):
        assert col.violations(A) == 0
        b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
        y_ref = ref.mcsor_apply(A, b, y0.copy(), col, omega, sweep, nsweeps=2)
        mc = orc.MCSOR(A, col, omega, sweep)
        y = y0.copy()
        mc.apply(b, y)
        mc.apply(b, y)
        assert rel(y, y_ref) < RTOL


def test_seq_sweep_galerkin_operator():
    """A 9-point Galerkin coarse operator (non-constant near the boundary) with 4 colours."""
    rng = np.random.default_rng(SEED + 2)
    mg = orc.MG.geometric(2, 33, 33, 1, 1.0, 3)
    A = mg.level_csr(1)
    col = orc.Coloring.parity(mg.level_dims(1)[:2], 4)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.mcsor_apply(A, b, y0.copy(), col, 1.0, orc.SOR_FORWARD)
    y = orc.MCSOR(A, col, 1.0, orc.SOR_FORWARD).apply(b, y0.copy())
    assert rel(y, y_ref) < RTOL


@pytest.mark.parametrize("nranks", [2, 3, 4])
@pytest.mark.parametrize("omega,sweep", [(1.0, orc.SOR_FORWARD), (1.4, orc.SOR_BACKWARD), (1.2, orc.SOR_SYMMETRIC)])
def test_partitioned_sweep(nranks, omega, sweep):
    """MCSORApply_MPIAIJ + MatCreateScatters (ghost packing order, backward ghost walk) on emulated ranks."""
    rng = np.random.default_rng(SEED + 3)
    shape = (21, 17)
    A = orc.laplace(2, *shape, kappa=1.0)
    col = orc.Coloring.parity(shape)
    cuts = np.sort(rng.choice(np.arange(1, A.n), nranks - 1, replace=False))  # ragged row blocks, not aligned to grid rows
    rowstart = np.concatenate([[0], cuts, [A.n]])
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.mcsor_apply_mpi(A, rowstart, col, b, y0.copy(), omega, sweep, nsweeps=2)
    part = orc.Partitioned(A, rowstart, col, omega)
    y = y0.copy()
    part.sweep(b, y, sweep)
    part.sweep(b, y, sweep)
    assert rel(y, y_ref) < RTOL
    # a distributed multicolour sweep equals the sequential one with the same colouring
    y_seq = ref.mcsor_apply(A, b, y0.copy(), col, omega, sweep, nsweeps=2)
    assert rel(y_ref, y_seq) < RTOL


def test_box_muller_stream():
    """VecSetRandomStandardNormal, Box-Muller branch (src/parmgmc.c:100-110): pairing, odd tail, call order."""
    for n in (1, 2, 7, 1000):
        z_ref = ref.normal_fill(0x12345678, n, ncalls=3)
        ns = orc.Noise.rander48(0x12345678)
        z = np.concatenate([orc.noise_fill(ns, n) for _ in range(3)])
        assert np.abs(z - z_ref).max() < 1e-14 * max(1.0, np.abs(z_ref).max())


@pytest.mark.parametrize("omega,sweep_opt,sweep", [(None, "", orc.SOR_FORWARD), (1.2, "-pc_mcgibbs_backward", orc.SOR_BACKWARD), (1.6, "-pc_mcgibbs_symmetric", orc.SOR_SYMMETRIC)])
@pytest.mark.parametrize("colored", [False, True])
def test_mcgibbs_richardson(omega, sweep_opt, sweep, colored):
    """PCMCGIBBS end to end: options, sqrtdiag, PrepareRHS_Default, the noise tape order (symmetric = two fills per
    sample), the callback after every sample (src/pc_mcgibbs.c:119-188)."""
    rng = np.random.default_rng(SEED + 4)
    shape = (33, 21)
    A = orc.laplace(2, *shape, kappa=10.0)
    col = orc.Coloring.parity(shape) if colored else None
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    seen_ref, seen = [], []
    y_ref = ref.mcgibbs_richardson(A, b, y0.copy(), 4, 777, col, omega, sweep_opt, callback=lambda it, y: seen_ref.append((it, y.copy())))
    y = orc.gibbs_richardson(A, b, y0.copy(), 4, orc.Noise.rander48(777), col, 1.0 if omega is None else omega, sweep, callback=lambda it, y: seen.append((it, y.copy())))
    assert rel(y, y_ref) < RTOL
    assert [i for i, _ in seen] == [i for i, _ in seen_ref] == [0, 1, 2, 3]
    for (_, a), (_, r) in zip(seen, seen_ref):
        assert rel(a, r) < RTOL
    # b = NULL (prior sampling, examples/ex8.c:47-49)
    y_ref = ref.mcgibbs_richardson(A, None, y0.copy(), 2, 5, col, omega, sweep_opt)
    y = orc.gibbs_richardson(A, None, y0.copy(), 2, orc.Noise.rander48(5), col, 1.0 if omega is None else omega, sweep)
    assert rel(y, y_ref) < RTOL


def test_ex5_identity_on_reference():
    """examples/ex5.c:60-70 on the reference itself: (forward; backward) == symmetric."""
    rng = np.random.default_rng(SEED + 5)
    A = orc.laplace(2, 9, 9, kappa=1.0)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y1 = ref.mcsor_apply(A, b, y0.copy(), None, 1.0, orc.SOR_FORWARD)
    ref.mcsor_apply(A, b, y1, None, 1.0, orc.SOR_BACKWARD)
    y2 = ref.mcsor_apply(A, b, y0.copy(), None, 1.0, orc.SOR_SYMMETRIC)
    assert np.linalg.norm(y1 - y2) < 1e-15
