#!/usr/bin/env python
"""Benchmark driver in the shape of the reference's examples/benchmark/main.cc: same option names, same
output lines, so that a run is comparable line for line with a PETSc run of the reference.

  python tools/benchmark.py -pc_type gamgmc -n 1025 -kappa 1e-2 -n_burnin 1000 -n_samples 10000 -measure_sampling_time -measure_iact [-with_lr] [-est_mean_and_var]

Problem: the shifted Laplacian of src/problems.c on an n^dim grid (the FE / lshape.msh problem needs PETSc's DMPlex and is out of
scope); -with_lr adds 17 ball observations with noise variance 1e-5 as a MATLRC term (examples/benchmark/lshape.opts:4-8);
the QOI is the mean of the field over a ball around the centre (lshape.opts:11-13).  All other options are passed to the PC
(-pc_mcgibbs_omega, -gamgmc_pc_mg_levels, -pc_woodbury_sampler ...).  Statistics run on the device (pmg_pc_set_qoi, pmg_iact)."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import parmgmc_b200 as pmg  # noqa: E402


def parse(argv):
    own = {"n": 1025, "dim": 2, "kappa": 1.0, "n_burnin": 1000, "n_samples": 10000, "seed": 1, "pc_type": "gamgmc"}
    flags = {"measure_sampling_time": False, "measure_iact": False, "with_lr": False, "est_mean_and_var": False, "print_acf": False}
    pcopts, i = {}, 0
    while i < len(argv):
        k = argv[i].lstrip("-")
        nxt = argv[i + 1] if i + 1 < len(argv) and not (argv[i + 1].startswith("-") and not argv[i + 1][1:2].isdigit() and argv[i + 1][1:2] != ".") else None
        if k in flags:
            flags[k] = True if nxt is None else nxt.lower() in ("1", "true", "yes")
        elif k in own:
            own[k] = type(own[k])(nxt)
        else:
            pcopts["-" + k] = "" if nxt is None else nxt
        i += 1 if nxt is None else 2
    return own, flags, pcopts


def ball_indicator(dims, centre, radius):
    """Normalised indicator of the nodes within `radius` of `centre` (unit square / cube coordinates)."""
    ax = [np.linspace(0.0, 1.0, d) for d in dims if d > 1]
    grid = np.meshgrid(*ax, indexing="ij")
    r2 = sum((g - c) ** 2 for g, c in zip(grid, centre))
    ind = (r2 <= radius * radius).astype(np.float64)
    return np.ravel(ind, order="F")  # natural order: x fastest


def main():
    own, flags, pcopts = parse(sys.argv[1:])
    if not (flags["measure_iact"] or flags["measure_sampling_time"]):
        raise SystemExit("Pass at least one of -measure_sampling_time or -measure_iact")  # main.cc:215
    bar = "#" * 80
    print(bar + "\n#############                Benchmark Test Program                #############\n" + bar)
    n, dim = own["n"], own["dim"]
    dims = (n, n, n if dim == 3 else 1)
    ctx = pmg.Context(0, seed=own["seed"])
    t0 = time.time()
    print("Starting assembly of operator... ", end="")
    A = pmg.Mat.laplace(ctx, dim, *dims, kappa=own["kappa"])
    op, N = A, A.n
    b = np.zeros(N)
    if flags["with_lr"]:  # 17 observations, sigma^2 = 1e-5 (lshape.opts:4-8): B column = indicator of a small ball, S = 1 / sigma^2
        rng = np.random.default_rng(17)
        k = 17
        B = np.column_stack([ball_indicator(dims, rng.uniform(0.15, 0.85, dim), 0.03) for _ in range(k)])
        B /= np.maximum(B.sum(0), 1.0)
        S = np.full(k, 1e5)
        obs = rng.standard_normal(k)
        b = B @ (S * obs)  # posterior right-hand side B S y_obs
        op = pmg.Mat.lrc(A, B, S)
    meas = ball_indicator(dims, (0.5,) * dim, 0.2)
    meas /= max(meas.sum(), 1.0)
    print(f"done. Took {time.time() - t0:.4f}s.")
    print("Starting Setup sampler... ", end="")
    t0 = time.time()
    pc = pmg.PC(ctx, own["pc_type"])
    pc.set_operator(op)
    pc.set_options(pcopts)
    pc.set_option("-pc_b200_noise", "philox")
    pc.setup()
    print(f" done. Took {time.time() - t0:.4f}s.")
    x = np.zeros(N)

    def timed(name, fn):
        print(f"Starting {name}... ", end="", flush=True)
        t = time.time()
        fn()
        ctx.synchronize()
        dt = time.time() - t
        print(f" done. Took {dt:.4f}s.")
        return dt

    import torch
    xd = torch.zeros(N, dtype=torch.float64, device="cuda")
    bd = torch.from_numpy(b).cuda()
    if flags["measure_sampling_time"]:
        print(bar + "\n                              Measure sampling time\n" + bar)
        timed("Burn-in", lambda: pc.apply_richardson_dev(bd, xd, its=own["n_burnin"]))
        dt = timed("Sampling", lambda: pc.apply_richardson_dev(bd, xd, its=own["n_samples"]))
        print(f"Time per sample [ms]: {dt / own['n_samples'] * 1000:.6f}\n")
    if flags["measure_iact"]:
        print(bar + "\n                                  Measure IACT\n" + bar)
        timed("Burn-in", lambda: pc.apply_richardson_dev(bd, xd, its=own["n_burnin"]))
        pc.set_qoi(meas, own["n_samples"] + 1, flags["est_mean_and_var"])
        timed("Sampling", lambda: pc.apply_richardson_dev(bd, xd, its=own["n_samples"]))
        qois = pc.get_qoi()
        tau, valid = pmg.iact(ctx, qois)
        if not valid:
            print(f"WARNING: Chain is too short to give reliable IACT estimate (need at least {math.ceil(500 * tau)})")
        if flags["print_acf"]:
            print("ACF: " + " ".join(f"{v:.6f}" for v in pmg.autocorrelation(ctx, qois)[:50]))
        print(f"IACT: {tau:.5f}")
        print(f"QOI mean: {qois.mean():.6e}  QOI variance: {qois.var(ddof=1):.6e}")
        if flags["est_mean_and_var"]:
            mean, var, seen = pc.get_mean_var()
            print(f"Field mean (norm): {np.linalg.norm(mean):.6e}  field variance (mean): {var.mean():.6e}  samples: {seen}")


if __name__ == "__main__":
    main()
