"""Writes tests/golden/*.npz from the reference's OWN compiled code (oracle/_ref, i.e. /root/reference/src/mc_sor.c,
pc_mcgibbs.c, parmgmc.c, pc_sorgibbs.c, pc_chols.c, iact.c and stats.c built unmodified against oracle/petsc_stub).  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py
The GPU box has no /root/reference; there the committed files are the reference's voice (tests/test_golden.py).
Inputs are regenerated from the seeds below by the tests; only reference OUTPUTS (and the reference's own normal
stream, which the CUDA path takes as an injected noise tape) are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as orc  # noqa: E402  (problem generator and colourings only)
from oracle import ref  # noqa: E402

SEED = 20260625  # examples/ex13.py:38

# name -> (dim, shape, kappa, colouring, omega, sweep)
SWEEP_CASES = {
    "c1_129_k10_lex_fwd": (2, (129, 129), 10.0, "single", 1.0, 1),
    "c1_129_k1_lex_sym_w12": (2, (129, 129), 1.0, "single", 1.2, 3),
    "c1_129_k10_rb_bwd_w16": (2, (129, 129), 10.0, "parity", 1.6, 2),
    "g2d_40x23_greedy_sym": (2, (40, 23), 1.0, "greedy", 1.3, 3),
    "g3d_12x9x7_rb_fwd": (3, (12, 9, 7), 1.0, "parity", 1.0, 1),
    "g3d_12x9x7_lex_sym": (3, (12, 9, 7), 2.0, "single", 0.8, 3),
}
# name -> (shape, kappa, colouring, omega, option, sweep, its, seed)
GIBBS_CASES = {
    "gibbs_33x21_lex_fwd": ((33, 21), 10.0, "single", None, "", 1, 3, 777),
    "gibbs_33x21_rb_sym_w16": ((33, 21), 10.0, "parity", 1.6, "-pc_mcgibbs_symmetric", 3, 3, 778),
    "gibbs_129_rb_bwd_w12": ((129, 129), 10.0, "parity", 1.2, "-pc_mcgibbs_backward", 2, 2, 779),
}


def coloring(kind, A, shape):
    if kind == "single":
        return None
    if kind == "parity":
        return orc.Coloring.parity(shape)
    if kind == "greedy":
        return orc.Coloring.greedy(A)
    raise ValueError(kind)


def sweep_inputs(name, n):
    rng = np.random.default_rng([SEED, sum(map(ord, name))])
    return rng.standard_normal(n), rng.standard_normal(n)


def main():
    out = {}
    for name, (dim, shape, kappa, ckind, omega, sweep) in SWEEP_CASES.items():
        A = orc.laplace(dim, *shape, kappa=kappa)
        b, y0 = sweep_inputs(name, A.n)
        out[name] = ref.mcsor_apply(A, b, y0.copy(), coloring(ckind, A, shape), omega, sweep, nsweeps=2)
    np.savez(os.path.join(HERE, "mcsor_sweeps.npz"), **out)
    out = {}
    for name, (shape, kappa, ckind, omega, opt, sweep, its, seed) in GIBBS_CASES.items():
        A = orc.laplace(2, *shape, kappa=kappa)
        b, y0 = sweep_inputs(name, A.n)
        fills = its * (2 if sweep == 3 else 1)
        out[name + "__z"] = ref.normal_fill(seed, A.n, fills)  # the stream PCMCGIBBS consumes, in call order
        out[name + "__y"] = ref.mcgibbs_richardson(A, b, y0.copy(), its, seed, coloring(ckind, A, shape), omega, opt)
    np.savez(os.path.join(HERE, "mcgibbs_samples.npz"), **out)
    # the partitioned sweep on ragged row blocks (MCSORApply_MPIAIJ)
    out = {}
    A = orc.laplace(2, 21, 17, kappa=1.0)
    b, y0 = sweep_inputs("part", A.n)
    for nr, cuts in ((2, [100]), (4, [50, 151, 300])):
        rs = np.array([0] + cuts + [A.n])
        out[f"part_21x17_r{nr}"] = ref.mcsor_apply_mpi(A, rs, orc.Coloring.parity((21, 17)), b, y0.copy(), 1.2, 3, nsweeps=2)
        out[f"part_21x17_r{nr}__rowstart"] = rs
    np.savez(os.path.join(HERE, "mcsor_partitioned.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


def lrc_problem(seed=SEED):
    """The MATLRC operator of the round-2 golden cases: 13x11 grid, k = 4 sparse observation columns."""
    rng = np.random.default_rng([seed, 4242])
    A = orc.laplace(2, 13, 11, kappa=2.0)
    B = rng.standard_normal((A.n, 4)) * (rng.random((A.n, 4)) < 0.2)
    S = 1.0 + 10.0 * rng.random(4)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    return A, B, S, b, y0


def ar1(n, seed):
    rng = np.random.default_rng([SEED, seed])
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):
        x[i] = 0.8 * x[i - 1] + rng.standard_normal()
    return x


def main_round2():
    """pc_sorgibbs.c, pc_chols.c (dense branch), iact.c, stats.c and the MATLRC branches, from the same library."""
    out = {}
    A = orc.laplace(2, 33, 21, kappa=3.0)
    b, y0 = sweep_inputs("sorgibbs", A.n)
    out["sorgibbs_33x21__z"] = ref.normal_fill(4711, A.n, 3)
    out["sorgibbs_33x21__y"] = ref.sampler_run("sorgibbs", A, b, y0.copy(), 3, 4711)
    out["sorgibbs_33x21__pcapply"] = ref.sampler_run("sorgibbs", A, b, y0.copy(), 0, 4711)
    Ac = orc.laplace(2, 7, 9, kappa=2.0)
    bc, _ = sweep_inputs("chol", Ac.n)
    out["chol_7x9__z"] = ref.normal_fill(31337, Ac.n, 3)
    out["chol_7x9__y1"] = ref.sampler_run("cholsampler", Ac, bc, np.zeros(Ac.n), 1, 31337)
    out["chol_7x9__y3"] = ref.sampler_run("cholsampler", Ac, bc, np.zeros(Ac.n), 3, 31337)
    for n in (500, 5000):
        tau, valid, acf = ref.iact(ar1(n, n))
        out[f"iact_{n}__tau_valid"] = np.array([tau, float(valid)])
        out[f"iact_{n}__acf"] = acf
    As = orc.laplace(2, 4, 5, kappa=1.5)
    samples = np.random.default_rng([SEED, 12]).standard_normal((6, 9, As.n))
    out["cov_4x5__errs"] = ref.cov_errors(As, samples)
    A, B, S, b, y0 = lrc_problem()
    for name, sweep in (("fwd", 1), ("bwd", 2), ("sym", 3)):
        out[f"lrc_mcsor_{name}"] = ref.mcsor_apply_lrc(A, B, S, b, y0.copy(), None, 1.0, sweep, nsweeps=2)
    # stream consumed per directional sweep: a fill of n normals, then one of k (PrepareRHS_LRC).  The reference library only
    # exposes equal-sized fills from a fresh seed (ref.normal_fill), so this tape comes from the oracle's rander48 + Box-Muller,
    # which tests/test_oracle_ref.py::test_box_muller_stream pins to the reference's stream
    ns = orc.Noise.rander48(2024)
    out["lrc_gibbs__z"] = np.concatenate([orc.noise_fill(ns, m) for _ in range(6) for m in (A.n, 4)])
    out["lrc_mcgibbs_sym_w13__y"] = ref.sampler_run("mcgibbs", A, b, y0.copy(), 3, 2024, opts=(("-pc_mcgibbs_omega", 1.3), ("-pc_mcgibbs_symmetric", "")), lrc=(B, S))
    out["lrc_sorgibbs__y"] = ref.sampler_run("sorgibbs", A, b, y0.copy(), 3, 2024, lrc=(B, S))
    out["lrc_chol__y"] = ref.sampler_run("cholsampler", A, b, np.zeros(A.n), 1, 2024, lrc=(B, S))
    np.savez(os.path.join(HERE, "round2_pins.npz"), **out)


def main_woodbury():
    """src/woodbury.c compiled into the same library (late round 2): PCWOODBURY on the round-2 MATLRC operator with an exact solver
    and the samplers mcgibbs (forward, omega = 1) / sorgibbs / cholsampler.  Stream per sample: a fill of k normals (the observation
    noise comes first, src/woodbury.c:273), then the sampler's fill of n."""
    out = {}
    A, B, S, b, y0 = lrc_problem()
    ns = orc.Noise.rander48(5150)
    out["woodbury__z"] = np.concatenate([orc.noise_fill(ns, m) for _ in range(3) for m in (4, A.n)])
    for sampler in ("mcgibbs", "sorgibbs", "cholsampler"):
        out[f"woodbury_{sampler}__y"] = ref.sampler_run("woodbury", A, b, y0.copy(), 3, 5150, opts=(("-pc_woodbury_solver", "cholesky"), ("-pc_woodbury_sampler", sampler)), lrc=(B, S))
    np.savez(os.path.join(HERE, "woodbury_pins.npz"), **out)


if __name__ == "__main__":
    main()
    main_round2()
    main_woodbury()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
