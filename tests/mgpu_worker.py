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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pmg.Context(local, seed=4242)
    uid = [pmg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
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
    flag = torch.tensor([len(failed)], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
