"""Golden vectors written by the reference's OWN compiled code (tests/golden/make_golden.py, oracle/_ref).

not-gpu part: the oracle restatement reproduces them (this is what pins the oracle on the GPU box, where
/root/reference and oracle/_ref's sources do not exist).
gpu part: the CUDA path, through the C ABI, reproduces them -- injected colouring and, for the sampler cases, the
reference's own normal stream as the injected noise tape (SURVEY 8(c) tape contract).  Tolerance 1e-12 relative
(north_star); the reference leaves fused-multiply-add contraction to the compiler, so bitwise equality is not defined.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL_ORACLE = 1e-13
RTOL = 1e-12


def cases():
    import ast
    src = open(os.path.join(HERE, "golden", "make_golden.py")).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and node.targets[0].id in ("SWEEP_CASES", "GIBBS_CASES", "SEED"):
            out[node.targets[0].id] = ast.literal_eval(node.value)
    return out


C = cases()
SWEEPS = np.load(os.path.join(HERE, "golden", "mcsor_sweeps.npz"))
GIBBS = np.load(os.path.join(HERE, "golden", "mcgibbs_samples.npz"))
PART = np.load(os.path.join(HERE, "golden", "mcsor_partitioned.npz"))


def inputs(name, n):
    rng = np.random.default_rng([C["SEED"], sum(map(ord, name))])
    return rng.standard_normal(n), rng.standard_normal(n)


def coloring(orc, kind, A, shape):
    return {"single": lambda: None, "parity": lambda: orc.Coloring.parity(shape), "greedy": lambda: orc.Coloring.greedy(A)}[kind]()


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ---- the oracle against the reference's outputs (CPU) ---------------------------------------------------------
@pytest.mark.parametrize("name", sorted(C["SWEEP_CASES"]))
def test_oracle_sweep_matches_reference_output(orc, name):
    dim, shape, kappa, ckind, omega, sweep = C["SWEEP_CASES"][name]
    A = orc.laplace(dim, *shape, kappa=kappa)
    b, y = inputs(name, A.n)
    mc = orc.MCSOR(A, coloring(orc, ckind, A, shape), omega, sweep)
    mc.apply(b, y)
    mc.apply(b, y)
    assert rel(y, SWEEPS[name]) < RTOL_ORACLE


@pytest.mark.parametrize("name", sorted(C["GIBBS_CASES"]))
def test_oracle_sampler_matches_reference_output(orc, name):
    shape, kappa, ckind, omega, _opt, sweep, its, seed = C["GIBBS_CASES"][name]
    A = orc.laplace(2, *shape, kappa=kappa)
    b, y = inputs(name, A.n)
    col = coloring(orc, ckind, A, shape)
    w = 1.0 if omega is None else omega
    # (1) the oracle's own rander48 + Box-Muller stream equals the reference's
    ns = orc.Noise.rander48(seed)
    z = np.concatenate([orc.noise_fill(ns, A.n) for _ in range(GIBBS[name + "__z"].size // A.n)])
    assert np.abs(z - GIBBS[name + "__z"]).max() < 1e-14 * np.abs(z).max()
    # (2) samples with that stream, and with the reference's stream injected as a tape
    y1 = orc.gibbs_richardson(A, b, y.copy(), its, orc.Noise.rander48(seed), col, w, sweep)
    y2 = orc.gibbs_richardson(A, b, y.copy(), its, orc.Noise.tape(GIBBS[name + "__z"]), col, w, sweep)
    assert rel(y1, GIBBS[name + "__y"]) < RTOL_ORACLE and rel(y2, GIBBS[name + "__y"]) < RTOL_ORACLE


@pytest.mark.parametrize("nr", [2, 4])
def test_oracle_partitioned_matches_reference_output(orc, nr):
    A = orc.laplace(2, 21, 17, kappa=1.0)
    b, y = inputs("part", A.n)
    part = orc.Partitioned(A, PART[f"part_21x17_r{nr}__rowstart"], orc.Coloring.parity((21, 17)), 1.2)
    part.sweep(b, y, orc.SOR_SYMMETRIC)
    part.sweep(b, y, orc.SOR_SYMMETRIC)
    assert rel(y, PART[f"part_21x17_r{nr}"]) < RTOL_ORACLE


# ---- the CUDA path against the reference's outputs (GPU) ------------------------------------------------------------
@pytest.fixture(scope="module")
def pmg():
    import parmgmc_b200 as m
    if m.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return m


@pytest.fixture(scope="module")
def ctx(pmg):
    c = pmg.Context(0, seed=0xCAFE)
    yield c
    c.close()


def device_mat(pmg, ctx, orc, A, ckind, shape):
    m = pmg.Mat.from_csr(ctx, A.rowptr, A.col, A.val)
    if ckind == "single":
        m.set_coloring_auto(pmg.COLORING_LEXICOGRAPHIC)  # level sets == the one-colour natural-order sweep
    else:
        col = coloring(orc, ckind, A, shape)
        m.set_coloring(col.color, col.ncolors)
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(C["SWEEP_CASES"]))
def test_cuda_sweep_matches_reference_output(pmg, ctx, orc, name):
    dim, shape, kappa, ckind, omega, sweep = C["SWEEP_CASES"][name]
    A = orc.laplace(dim, *shape, kappa=kappa)
    b, y = inputs(name, A.n)
    mc = pmg.MCSOR(device_mat(pmg, ctx, orc, A, ckind, shape))
    mc.set_omega(omega)
    mc.set_sweep_type(sweep)
    y = mc.apply(b, y)
    y = mc.apply(b, y)
    assert rel(y, SWEEPS[name]) < RTOL
    if ckind == "parity":  # the matrix-free operator too
        lap = pmg.Mat.laplace(ctx, dim, *shape, kappa=kappa)
        mc = pmg.MCSOR(lap)
        mc.set_omega(omega)
        mc.set_sweep_type(sweep)
        _, y2 = inputs(name, A.n)
        y2 = mc.apply(b, y2)
        y2 = mc.apply(b, y2)
        assert rel(y2, SWEEPS[name]) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(C["GIBBS_CASES"]))
def test_cuda_sampler_matches_reference_output(pmg, ctx, orc, name):
    shape, kappa, ckind, omega, opt, sweep, its, seed = C["GIBBS_CASES"][name]
    A = orc.laplace(2, *shape, kappa=kappa)
    b, y = inputs(name, A.n)
    mats = [device_mat(pmg, ctx, orc, A, ckind, shape)]
    if ckind == "parity":
        mats.append(pmg.Mat.laplace(ctx, 2, *shape, kappa=kappa))  # matrix-free: per-colour and fused streaming kernels
    for mat in mats:
        pc = pmg.PC(ctx, "mcgibbs")
        pc.set_operator(mat)
        if omega is not None:
            pc.set_option("-pc_mcgibbs_omega", omega)
        if opt:
            pc.set_option(opt, "")
        pc.setup()
        pc.set_noise_tape(GIBBS[name + "__z"])  # the reference's own normal stream, in its call order
        yy = y.copy()
        pc.apply_richardson(b, yy, its=its)
        assert rel(yy, GIBBS[name + "__y"]) < RTOL
