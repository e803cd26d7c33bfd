"""Writes tests/golden/*.npz from the reference's OWN compiled code (oracle/_ref, i.e. /root/reference/src/mc_sor.c,
pc_mcgibbs.c and parmgmc.c built unmodified against oracle/petsc_stub).  Run in the build container, where
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


if __name__ == "__main__":
    main()
