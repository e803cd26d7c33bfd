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


# ---- round 2: pc_sorgibbs.c, pc_chols.c (dense branch), iact.c, stats.c and the MATLRC branches (tests/golden/round2_pins.npz,
#      written by make_golden.main_round2 from the same reference library) ---------------------------------------------------
R2 = np.load(os.path.join(HERE, "golden", "round2_pins.npz"))


def lrc_problem(orc):
    rng = np.random.default_rng([C["SEED"], 4242])
    A = orc.laplace(2, 13, 11, kappa=2.0)
    B = rng.standard_normal((A.n, 4)) * (rng.random((A.n, 4)) < 0.2)
    S = 1.0 + 10.0 * rng.random(4)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    return A, B, S, b, y0


def ar1(n, seed):
    rng = np.random.default_rng([C["SEED"], seed])
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):
        x[i] = 0.8 * x[i - 1] + rng.standard_normal()
    return x


def test_oracle_sorgibbs_and_cholsampler_match_reference_output(orc):
    A = orc.laplace(2, 33, 21, kappa=3.0)
    b, y0 = inputs("sorgibbs", A.n)
    y = orc.gibbs_richardson(A, b, y0.copy(), 3, orc.Noise.tape(R2["sorgibbs_33x21__z"]), None, 1.0, orc.SOR_FORWARD)
    assert rel(y, R2["sorgibbs_33x21__y"]) < RTOL_ORACLE
    y = orc.gibbs_richardson(A, b, np.zeros(A.n), 1, orc.Noise.rander48(4711), None, 1.0, orc.SOR_FORWARD)  # PCApply zeroes y
    assert rel(y, R2["sorgibbs_33x21__pcapply"]) < RTOL_ORACLE
    Ac = orc.laplace(2, 7, 9, kappa=2.0)
    bc, _ = inputs("chol", Ac.n)
    lflat = orc.potrf_lower(Ac.to_scipy().toarray())
    ns = orc.Noise.tape(R2["chol_7x9__z"])
    y1 = orc.chol_sample(lflat, Ac.n, ns, bc)
    assert rel(y1, R2["chol_7x9__y1"]) < 1e-12
    y3 = orc.chol_sample(lflat, Ac.n, ns, bc)
    y3 = orc.chol_sample(lflat, Ac.n, ns, bc)
    assert rel(y3, R2["chol_7x9__y3"]) < 1e-12


@pytest.mark.parametrize("n", [500, 5000])
def test_oracle_iact_matches_reference_output(orc, n):
    x = ar1(n, n)
    tau, valid = orc.iact(x)
    assert abs(tau - R2[f"iact_{n}__tau_valid"][0]) < 1e-10 * abs(tau) and float(valid) == R2[f"iact_{n}__tau_valid"][1]
    assert np.abs(orc.autocorrelation(x) - R2[f"iact_{n}__acf"]).max() < 1e-12


def test_oracle_cov_errors_match_reference_output(orc):
    As = orc.laplace(2, 4, 5, kappa=1.5)
    samples = np.random.default_rng([C["SEED"], 12]).standard_normal((6, 9, As.n))
    errs = orc.cov_errors(As.to_scipy().toarray(), samples)
    assert np.abs(errs - R2["cov_4x5__errs"]).max() < 1e-12 * np.abs(errs).max()


def test_oracle_matlrc_paths_match_reference_output(orc):
    A, B, S, b, y0 = lrc_problem(orc)
    Bb = {d: orc.lrc_build_correction(A, B, S, None, 1.0, d) for d in (orc.SOR_FORWARD, orc.SOR_BACKWARD)}
    for name, sweep in (("fwd", orc.SOR_FORWARD), ("bwd", orc.SOR_BACKWARD), ("sym", orc.SOR_SYMMETRIC)):
        y = y0.copy()
        orc.lrc_mcsor_apply(A, B, Bb, b, y, None, 1.0, sweep)
        orc.lrc_mcsor_apply(A, B, Bb, b, y, None, 1.0, sweep)
        assert rel(y, R2[f"lrc_mcsor_{name}"]) < 1e-11
    z = R2["lrc_gibbs__z"]
    y = orc.lrc_gibbs_richardson(A, B, S, b, y0.copy(), 3, orc.Noise.tape(z), None, 1.3, orc.SOR_SYMMETRIC)
    assert rel(y, R2["lrc_mcgibbs_sym_w13__y"]) < 1e-11
    y = orc.lrc_gibbs_richardson(A, B, S, b, y0.copy(), 3, orc.Noise.tape(z), None, 1.0, orc.SOR_FORWARD)
    assert rel(y, R2["lrc_sorgibbs__y"]) < 1e-11
    P = A.to_scipy().toarray() + B @ np.diag(S) @ B.T
    y = orc.chol_sample(orc.potrf_lower(P), A.n, orc.Noise.tape(z), b)
    assert rel(y, R2["lrc_chol__y"]) < 1e-11


@pytest.mark.gpu
def test_cuda_sorgibbs_and_cholsampler_match_reference_output(pmg, ctx, orc):
    A = orc.laplace(2, 33, 21, kappa=3.0)
    b, y0 = inputs("sorgibbs", A.n)
    mat = device_mat(pmg, ctx, orc, A, "single", (33, 21))
    pc = pmg.PC(ctx, "sorgibbs")
    pc.set_operator(mat)
    pc.setup()
    pc.set_noise_tape(R2["sorgibbs_33x21__z"])
    y = y0.copy()
    pc.apply_richardson(b, y, its=3)
    assert rel(y, R2["sorgibbs_33x21__y"]) < RTOL
    pc.set_noise_tape(R2["sorgibbs_33x21__z"])
    assert rel(pc.apply(b), R2["sorgibbs_33x21__pcapply"]) < RTOL  # PCApply_SORGibbs zeroes y (src/pc_sorgibbs.c:110)
    Ac = orc.laplace(2, 7, 9, kappa=2.0)
    bc, _ = inputs("chol", Ac.n)
    for solve in ("trsv", "gemv"):
        ch = pmg.PC(ctx, "cholsampler")
        ch.set_operator(pmg.Mat.from_csr(ctx, Ac.rowptr, Ac.col, Ac.val))
        ch.set_option("-pc_cholsampler_b200_solve", solve)
        ch.setup()
        ch.set_noise_tape(R2["chol_7x9__z"])
        y1 = np.zeros(Ac.n)
        ch.apply_richardson(bc, y1, its=1)
        assert rel(y1, R2["chol_7x9__y1"]) < RTOL
        ch.set_noise_tape(R2["chol_7x9__z"])
        y3 = np.zeros(Ac.n)
        ch.apply_richardson(bc, y3, its=3)
        assert rel(y3, R2["chol_7x9__y3"]) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("n", [500, 5000])
def test_cuda_iact_matches_reference_output(pmg, ctx, n):
    x = ar1(n, n)
    tau, valid = pmg.iact(ctx, x)
    assert abs(tau - R2[f"iact_{n}__tau_valid"][0]) < 1e-9 * abs(tau) and float(valid) == R2[f"iact_{n}__tau_valid"][1]
    assert np.abs(pmg.autocorrelation(ctx, x) - R2[f"iact_{n}__acf"]).max() < 1e-11


@pytest.mark.gpu
def test_cuda_matlrc_paths_match_reference_output(pmg, ctx, orc):
    A, B, S, b, y0 = lrc_problem(orc)
    base = device_mat(pmg, ctx, orc, A, "single", (13, 11))
    mat = pmg.Mat.lrc(base, B, S)
    for name, sweep in (("fwd", 1), ("bwd", 2), ("sym", 3)):
        mc = pmg.MCSOR(mat)
        mc.set_sweep_type(sweep)
        y = mc.apply(b, y0.copy())
        y = mc.apply(b, y)
        assert rel(y, R2[f"lrc_mcsor_{name}"]) < 1e-10
    z = R2["lrc_gibbs__z"]
    for pctype, opts, key in (("mcgibbs", {"-pc_mcgibbs_omega": 1.3, "-pc_mcgibbs_symmetric": ""}, "lrc_mcgibbs_sym_w13__y"), ("sorgibbs", {}, "lrc_sorgibbs__y")):
        pc = pmg.PC(ctx, pctype)
        pc.set_operator(mat)
        pc.set_options(opts)
        pc.setup()
        pc.set_noise_tape(z)
        y = y0.copy()
        pc.apply_richardson(b, y, its=3)
        assert rel(y, R2[key]) < 1e-10
    ch = pmg.PC(ctx, "cholsampler")
    ch.set_operator(mat)
    ch.setup()
    ch.set_noise_tape(z)
    y = np.zeros(A.n)
    ch.apply_richardson(b, y, its=1)
    assert rel(y, R2["lrc_chol__y"]) < 1e-10


# ---- late round 2: src/woodbury.c (tests/golden/woodbury_pins.npz, written by make_golden.main_woodbury from the reference's own
#      woodbury.c compiled against the stub) -------------------------------------------------------------------------------------
WB = np.load(os.path.join(HERE, "golden", "woodbury_pins.npz"))


def _woodbury_restatement(orc, A, B, S, b, y0, sampler, its, z):
    """PCApplyRichardson_Woodbury (src/woodbury.c:259-286) with G = C (S^-1 + B^T C)^-1, C = A^-1 B (:21-86, exact solver)."""
    Ad = A.to_scipy().toarray()
    Cm = np.linalg.solve(Ad, B)
    G = Cm @ np.linalg.inv(np.diag(1.0 / S) + B.T @ Cm)
    noise = orc.Noise.tape(z)
    lflat = orc.potrf_lower(Ad) if sampler == "cholsampler" else None
    y = y0.copy()
    for _ in range(its):
        w = b + B @ (np.sqrt(np.abs(S)) * orc.noise_fill(noise, B.shape[1]))
        y = orc.chol_sample(lflat, A.n, noise, w) if sampler == "cholsampler" else orc.gibbs_richardson(A, w, y, 1, noise, None, 1.0, orc.SOR_FORWARD)
        y = y - G @ (B.T @ y)
    return y


@pytest.mark.parametrize("sampler", ["mcgibbs", "sorgibbs", "cholsampler"])
def test_oracle_woodbury_matches_reference_output(orc, sampler):
    A, B, S, b, y0 = lrc_problem(orc)
    assert rel(_woodbury_restatement(orc, A, B, S, b, y0, sampler, 3, WB["woodbury__z"]), WB[f"woodbury_{sampler}__y"]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", ["mcgibbs", "sorgibbs", "cholsampler"])
def test_cuda_woodbury_matches_reference_output(pmg, ctx, orc, sampler):
    """The CUDA PCWOODBURY against what the reference's own woodbury.c produced (one colour = the reference's 1-rank ordering)."""
    A, B, S, b, y0 = lrc_problem(orc)
    mat = pmg.Mat.lrc(device_mat(pmg, ctx, orc, A, "single", (13, 11)), B, S)
    pc = pmg.PC(ctx, "woodbury")
    pc.set_operator(mat)
    pc.set_options({"-pc_woodbury_sampler": sampler, "-pc_woodbury_solver": "cholesky"})
    pc.setup()
    pc.set_noise_tape(WB["woodbury__z"])
    y = y0.copy()
    pc.apply_richardson(b, y, its=3)
    assert rel(y, WB[f"woodbury_{sampler}__y"]) < 1e-9
