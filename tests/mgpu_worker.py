"""One rank of a multi-GPU parity run (launched by tests/test_multigpu.py with torch.distributed.run, one process per
GPU).  Every case runs the slab-partitioned sampler over NCCL, gathers the slabs on rank 0 and compares them BITWISE
with the same sampler on one GPU: Philox is keyed on the global row and the colouring is partition independent, so the
distributed result must not depend on the number of ranks (SURVEY 8(e) determinism)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parmgmc_b200 as pmg  # noqa: E402

CASES = {
    # name: (pc type, dim, dims, options, its, ragged cut?)
    "sorgibbs2d": ("sorgibbs", 2, (65, 67, 1), {}, 3, True),
    "mcgibbs3d_sym": ("mcgibbs", 3, (17, 12, 20), {"-pc_mcgibbs_symmetric": "", "-pc_mcgibbs_omega": 1.3}, 2, True),
    "mcgibbs3d_narrow_strip": ("mcgibbs", 3, (150, 40, 48), {"-pc_mcgibbs_symmetric": "", "-pc_mcgibbs_omega": 1.2}, 2, False),  # 16-lane last strip, thin z-edge bands on slabs
    "gamgmc2d_deep": ("gamgmc", 2, (129, 257, 1), {"-gamgmc_pc_mg_levels": 5, "-pc_b200_replicate_below": 500}, 3, True),
    "gamgmc2d_default": ("gamgmc", 2, (129, 257, 1), {"-gamgmc_pc_mg_levels": 4}, 2, False),
    "gamgmc3d": ("gamgmc", 3, (33, 33, 65), {"-gamgmc_pc_mg_levels": 3, "-pc_b200_replicate_below": 3000}, 2, False),
    "gamgmc2d_its2_literal": ("gamgmc", 2, (65, 129, 1), {"-gamgmc_pc_mg_levels": 3, "-gamgmc_mg_levels_ksp_max_it": 2, "-pc_b200_cycle": "literal", "-pc_b200_replicate_below": 100}, 2, True),
}


def run(ctx, pctype, dim, dims, opts, its, slab, b_full, y0_full):
    nx, ny, nz = dims
    mat = pmg.Mat.laplace(ctx, dim, nx, ny, nz, kappa=1.0, slab=slab)
    nloc, _, row0 = mat.size
    pc = pmg.PC(ctx, pctype)
    pc.set_operator(mat)
    pc.set_options(dict(opts, **{"-pc_b200_noise": "philox"}))
    pc.setup()
    ctx.set_seed(4242)
    y = y0_full[row0:row0 + nloc].copy()
    pc.apply_richardson(b_full[row0:row0 + nloc].copy(), y, its=its)
    return row0, y, ctx.draw_counter


def csr_dist_case(ctx, rank, world, user_coloring):
    """Row-partitioned general CSR operator (MCSORApply_MPIAIJ, src/mc_sor.c:298-381): a randomly permuted 2D Laplacian, so
    that every rank has ghost columns on every other rank.  Checked against (a) the same sampler (device Philox noise, keyed on
    the global row) on one GPU with the gathered colouring and (b) the oracle's thread-emulated MCSORApply_MPIAIJ (pinned
    against the reference binary); both to rounding, 1e-13."""
    import oracle as orc
    import scipy.sparse as sp
    A0 = orc.laplace(2, 33, 29, kappa=1.0)
    n = A0.n
    rng = np.random.default_rng(7)
    perm = rng.permutation(n)
    M = sp.csr_matrix(A0.to_scipy())[perm][:, perm].tocsr()
    M.sort_indices()
    A = orc.CSR(n, M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.astype(np.float64))
    starts = np.round(np.linspace(0, n, world + 1) + (np.arange(world + 1) % 2) * 3).astype(np.int64)
    starts[0], starts[-1] = 0, n
    r0, r1 = int(starts[rank]), int(starts[rank + 1])
    rp = (M.indptr[r0:r1 + 1] - M.indptr[r0]).astype(np.int64)
    cols = M.indices[M.indptr[r0]:M.indptr[r1]].astype(np.int64)
    vals = M.data[M.indptr[r0]:M.indptr[r1]].astype(np.float64)
    mat = pmg.Mat.from_csr_dist(ctx, n, r0, rp, cols, vals)
    assert mat.size == (r1 - r0, n, r0)
    if user_coloring == "lex":  # global level sets: the sweep over all ranks IS the natural-order Gauss-Seidel sweep (PCPARSOR's result)
        mat.set_coloring_auto(pmg.COLORING_LEXICOGRAPHIC)
    elif user_coloring:
        gcol = orc.Coloring.greedy(A)
        mat.set_coloring(gcol.color[r0:r1], gcol.ncolors)
        bad = gcol.color.copy()
        bad[:] = 0
        try:
            mat.set_coloring(bad[r0:r1], 1)  # every rank must reject a colouring that is invalid somewhere
            raise AssertionError("invalid colouring accepted")
        except pmg.PMGError:
            pass
    k, mycol = mat.get_coloring()
    b_full, y0_full, x_full = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    # sampler with device noise
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(mat)
    pc.set_options({"-pc_mcgibbs_symmetric": "", "-pc_mcgibbs_omega": 1.2, "-pc_b200_noise": "philox"})
    pc.setup()
    ctx.set_seed(4242)
    y = y0_full[r0:r1].copy()
    pc.apply_richardson(b_full[r0:r1].copy(), y, its=3)
    # deterministic MCSORApply and the operator product
    mc = pmg.MCSOR(mat)
    mc.set_omega(1.3)
    mc.set_sweep_type(3)
    ys = y0_full[r0:r1].copy()
    mc.apply(b_full[r0:r1].copy(), ys)
    ax = mat.mult(x_full[r0:r1].copy())
    parts = [None] * world
    dist.all_gather_object(parts, (r0, y, ys, ax, mycol, k))
    ok = True
    if rank == 0:
        got, gots, gax, col = np.empty(n), np.empty(n), np.empty(n), np.empty(n, np.int32)
        for q0, yy, yys, aax, cc, _ in parts:
            got[q0:q0 + yy.size], gots[q0:q0 + yy.size], gax[q0:q0 + yy.size], col[q0:q0 + yy.size] = yy, yys, aax, cc
        kk = max(p[5] for p in parts)
        coloring = orc.Coloring(col, kk)
        assert coloring.violations(A) == 0
        single = pmg.Context(0, seed=4242)
        m1 = pmg.Mat.from_csr(single, A.rowptr, A.col, A.val)
        m1.set_coloring(col, kk)
        p1 = pmg.PC(single, "mcgibbs")
        p1.set_operator(m1)
        p1.set_options({"-pc_mcgibbs_symmetric": "", "-pc_mcgibbs_omega": 1.2, "-pc_b200_noise": "philox"})
        p1.setup()
        single.set_seed(4242)
        ref = y0_full.copy()
        p1.apply_richardson(b_full.copy(), ref, its=3)
        part = orc.Partitioned(A, starts, coloring, 1.3)
        refs = y0_full.copy()
        part.sweep(b_full, refs, orc.SOR_SYMMETRIC, nthreads=world)
        # a row accumulates its diagonal-block entries before its off-diagonal-block entries (src/mc_sor.c:323-334), one GPU
        # accumulates in global column order: same noise, same colouring, results equal to rounding
        e1, e2, e3 = np.abs(got - ref).max() / np.abs(ref).max(), np.abs(gots - refs).max(), np.abs(gax - M @ x_full).max()
        ok = e1 < 1e-13 and e2 < 1e-13 and e3 < 1e-12
        extra = ""
        if user_coloring == "lex":
            # the colouring must be the level-set colouring of the WHOLE matrix, and the sweep the one-colour natural-order sweep
            # of the reference's 1-rank path (src/mc_sor.c:397-410) -- what PCPARSOR reproduces in parallel (src/pc_parsor.c:703-878)
            lex = orc.Coloring.levelset(A)
            nat = y0_full.copy()
            orc.MCSOR(A, orc.Coloring.single(n), 1.3, orc.SOR_SYMMETRIC).apply(b_full, nat)
            e4 = np.abs(gots - nat).max()
            same = bool(np.array_equal(lex.color, col)) and lex.ncolors == kk
            ok = ok and same and e4 < 1e-13
            extra = f" global_levelset={same} vs_natural_order_sweep={e4:.2e}"
        tag = "lexicographic" if user_coloring == "lex" else ("user" if user_coloring else "auto")
        print(f"[mgpu] csr_dist_{tag}: world={world} starts={starts.tolist()} colours={kk} sampler_vs_1gpu={e1:.2e} "
              f"mcsor_vs_oracle_MPIAIJ={e2:.2e} mult={e3:.2e}{extra} -> {'OK' if ok else 'FAIL'}", flush=True)
        single.close()
    dist.barrier()
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pmg.Context(local, seed=4242)
    uid = [pmg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    if rank == 0:
        print(f"[mgpu] slab halo exchange: {'peer memory (CUDA IPC mailboxes)' if ctx.comm_p2p() else 'NCCL send/recv'}", flush=True)
    names = sys.argv[1:] or sorted(CASES)
    failed = []
    for name in names:
        pctype, dim, dims, opts, its, ragged = CASES[name]
        nslow = dims[2] if dim == 3 else dims[1]
        slabs = pmg.partition_slabs(nslow, world, align=1 if ragged else 2)
        if ragged and world > 1:  # move the first cut to an odd unit
            lo, hi = slabs[0]
            slabs[0] = (lo, hi + 1 - (hi % 2))
            slabs[1] = (slabs[0][1], slabs[1][1])
        n = dims[0] * dims[1] * dims[2]
        rng = np.random.default_rng(99)
        b_full, y0_full = rng.standard_normal(n), rng.standard_normal(n)
        row0, y, draws = run(ctx, pctype, dim, dims, opts, its, slabs[rank], b_full, y0_full)
        parts = [None] * world
        dist.all_gather_object(parts, (row0, y, draws))
        if rank == 0:
            got = np.empty(n)
            for r0, yy, _ in parts:
                got[r0:r0 + yy.size] = yy
            single = pmg.Context(0, seed=4242)  # no communicator: the one-GPU path (fused kernels where available)
            _, ref, ref_draws = run(single, pctype, dim, dims, opts, its, None, b_full, y0_full)
            same = np.array_equal(got, ref)
            counters = {d for _, _, d in parts}
            ok = same and counters == {ref_draws}
            print(f"[mgpu] {name}: world={world} slabs={slabs} bitwise_equal={same} max|d|={np.abs(got - ref).max():.3e} draw_counters={sorted(counters)} single={ref_draws} -> {'OK' if ok else 'FAIL'}", flush=True)
            if not ok:
                failed.append(name)
            single.close()
        dist.barrier()
    if not sys.argv[1:]:
        for user in (False, True, "lex"):
            if not csr_dist_case(ctx, rank, world, user) and rank == 0:
                failed.append("csr_dist")
    flag = torch.tensor([len(failed)], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
