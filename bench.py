#!/usr/bin/env python
"""bench.py -- headline benchmark of the ParMGMC sampling hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port)

Workload (BASELINE.json configs[1]): 2D 5-point 4097x4097 shifted-Laplacian GMRF (src/problems.c:14-75
semantics), PCGAMGMC V(1,1) cycle with SOR-Gibbs smoothers, Galerkin geometric hierarchy, dense Cholesky
sampler on the coarsest level; b = 0 (prior sampling, examples/ex8.c:47-49), device Philox noise.
One "step" = `--samples-per-step` MGMC samples (outer Richardson iterations, src/pc_gamgmc.c:242-259).
For N > 1 the grid is weak-scaled in y (one 4097 x 4096 slab per rank, halo exchange over NCCL); a unit of
work is one sample of one 4097^2-DOF slab, so `value` = N * samples/s of the N-slab grid.

Prints ONE JSON line (see the keys below).  The oracle is used only for the cpu_baseline leg and for
--impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, nranks):
    n = args.n
    ny_local = n if nranks == 1 else n - 1
    return {"workload": f"2D 5-point {n}x{n} shifted-Laplacian GMRF per GPU, PCGAMGMC V(1,1), SOR-Gibbs smoother (red-black), Galerkin Q1 hierarchy, dense Cholesky coarsest",
            "grid_per_gpu": [n, ny_local], "grid_global": [n, ny_local * nranks + (1 if nranks > 1 else 0)], "kappa": args.kappa, "levels": args.levels,
            "samples_per_step": args.samples_per_step, "noise": "device Philox4x32-10 + Box-Muller", "rhs": "b = 0 (prior sampling)",
            "l2": "working set (>= 134 MB per fine vector, > 1 GB per sample) exceeds the 126 MB L2; no explicit flush",
            "parallelism": f"row-slab x{nranks}" if nranks > 1 else "single GPU"}


def cpu_baseline_leg(args, samples):
    """The reference's 1-rank CPU arithmetic (oracle port): lexicographic SOR-Gibbs sweeps, rander48 Box-Muller,
    PCMG V-cycle.  Timed on a bounded number of samples of the SAME workload."""
    import oracle as orc
    n = args.n
    t0 = time.time()
    mg = orc.MG.geometric(2, n, n, 1, args.kappa, args.levels)
    mg.setup()
    setup_s = time.time() - t0
    b, y = np.zeros(n * n), np.zeros(n * n)
    ns = orc.Noise.rander48()
    mg.richardson(ns, b, y, 1)  # warm caches / page in
    t0 = time.time()
    mg.richardson(ns, b, y, samples)
    dt = time.time() - t0
    return {"value": samples / dt, "unit": "samples/s", "cores": 1, "kind": "port",
            "sample": f"{samples} MGMC samples of the full {n}x{n} workload after 1 warm-up sample (setup {setup_s:.1f} s not timed); "
                      "single thread = the reference's 1-rank path (one-colour lexicographic sweeps, rander48 Box-Muller)",
            "ms_per_sample": 1e3 * dt / samples}


def _ref_worker(n, kappa, levels, per_step, warmup, steps, start_evt, ready_q, done_q):
    """One independent chain of the reference's 1-rank CPU arithmetic (what a rank of `mpirun -np T` replicas would run)."""
    import oracle as orc
    mg = orc.MG.geometric(2, n, n, 1, kappa, levels)
    mg.setup()
    b, y = np.zeros(n * n), np.zeros(n * n)
    ns = orc.Noise.rander48()
    for _ in range(warmup):
        mg.richardson(ns, b, y, per_step)
    ready_q.put(1)
    start_evt.wait()
    t0 = time.time()
    for _ in range(steps):
        mg.richardson(ns, b, y, per_step)
    done_q.put(time.time() - t0)


def run_reference(args):
    """The reference's CPU path on the box's host cores.  PETSc + MPI cannot be built here, so the
    arithmetic is the oracle port of the 1-rank reference; to use every host thread it runs T independent chains (the
    replica-parallel pattern of examples/ex7.c:136-205), T = min(cores, --ref-procs, memory / 3 GB)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    n = args.n
    per_step = max(1, args.ref_samples_per_step)
    cores = os.cpu_count() or 1
    try:
        mem_gb = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
    except (ValueError, OSError):
        mem_gb = 16.0
    per_proc_gb = 3.0 * (n / 4097.0) ** 2
    T = max(1, min(cores, args.ref_procs if args.ref_procs > 0 else 16, int(mem_gb * 0.6 / per_proc_gb)))
    mpc = mp.get_context("spawn")
    start_evt, ready_q, done_q = mpc.Event(), mpc.Queue(), mpc.Queue()
    procs = [mpc.Process(target=_ref_worker, args=(n, args.kappa, args.levels, per_step, args.warmup, args.steps, start_evt, ready_q, done_q)) for _ in range(T)]
    for p in procs:
        p.start()
    for _ in range(T):
        ready_q.get()
    t0 = time.time()
    start_evt.set()
    times = [done_q.get() for _ in range(T)]
    dt = time.time() - t0
    for p in procs:
        p.join()
    value = T * args.steps * per_step / dt
    cfg = workload_config(args, 1)
    cfg["samples_per_step"] = per_step
    cfg["parallelism"] = f"{T} independent CPU chains (one per host thread used)"
    cfg["noise"] = "rander48 Box-Muller (PETSc's default PetscRandom, src/parmgmc.c:100-110)"
    cfg["workload"] = cfg["workload"].replace("SOR-Gibbs smoother (red-black)", "SOR-Gibbs smoother (lexicographic, the reference's 1-rank order)")
    out = {"impl": "reference", "metric": "mgmc_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": value, "unit": "samples/s", "cores": T, "kind": "port", "host_cores": cores,
                            "sample": f"{T} chains x {args.steps} steps x {per_step} sample(s) of the full {n}x{n} workload (slowest chain {max(times):.1f} s); the reference (PETSc+MPI) cannot be "
                                      "built here, so each chain is the oracle port of its 1-rank path (one-colour lexicographic sweeps, rander48 Box-Muller)"},
           "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist

    import parmgmc_b200 as pmg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    ctx = pmg.Context(local, stream=stream.cuda_stream, seed=0xCAFE)
    if world > 1:
        uid = [pmg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])

    n = args.n
    if world == 1:
        ny, slab = n, None
    else:  # weak scaling in y: one (n-1)-row slab per rank, the last rank also owns the closing row
        ny = (n - 1) * world + 1
        slab = pmg.partition_slabs(ny, world)[rank]
    mat = pmg.Mat.laplace(ctx, 2, n, ny, 1, args.kappa, slab=slab)
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(mat)
    pc.set_options({"-gamgmc_pc_mg_levels": args.levels, "-pc_b200_noise": "philox"})
    t0 = time.time()
    pc.setup()
    setup_s = time.time() - t0
    nloc = mat.n
    S = args.samples_per_step

    y = torch.zeros(nloc, dtype=torch.float64, device="cuda")
    b = torch.zeros(nloc, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        pc.apply_richardson_dev(b, y, its=S)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = updates = 0
    e0.record(stream)
    for _ in range(args.steps):
        pc.apply_richardson_dev(b, y, its=S)
        st = pc.last_stats()
        launches += st["launches"]; updates += st["dof_updates"]
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clk = clocks.stop() if rank == 0 else None
    value = world * args.steps * S / (ms * 1e-3)

    # ---- end to end through the host-pointer C-ABI call (pinned host buffers, H2D + D2H inside) ----
    hb = torch.zeros(nloc, dtype=torch.float64).pin_memory().numpy()
    hy = torch.zeros(nloc, dtype=torch.float64).pin_memory().numpy()
    pc.apply_richardson(hb, hy, its=S)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e0.record(stream)
    for _ in range(e2e_steps):
        pc.apply_richardson(hb, hy, its=S)  # H2D of b and of the chain state y, S samples, D2H of y
    e1.record(stream)
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * S / (float(ms2.item()) * 1e-3)

    # ---- roofline of the dominant kernel: the fine-level colour sweep, timed alone on the same stream ----
    gibbs = pmg.PC(ctx, "sorgibbs")
    gibbs.set_operator(mat)
    gibbs.set_option("-pc_b200_noise", "philox")
    gibbs.setup()
    for _ in range(3):
        gibbs.apply_richardson_dev(b, y, its=4)
    barrier()
    nsweeps = 40
    e0.record(stream)
    gibbs.apply_richardson_dev(b, y, its=nsweeps)
    e1.record(stream)
    barrier()
    gst = gibbs.last_stats()
    sweep_ms = e0.elapsed_time(e1) / nsweeps
    launches_per_sweep = gst["launches"] / nsweeps
    bytes_per_update = args.bytes_per_update
    alg_bytes_launch = bytes_per_update * nloc / launches_per_sweep
    achieved = alg_bytes_launch / (sweep_ms * 1e-3 / launches_per_sweep) / 1e9
    peak, peak_src = measured_peaks()
    # measured DRAM traffic of one launch of that kernel (ncu --set full, profiles/): only for the configuration it was taken on
    traffic = None
    tj = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if world == 1 and os.path.exists(tj):
        t = json.load(open(tj)).get("sweep2d_kernel", {})
        if t.get("n") == n:
            traffic = t.get("dram_bytes_per_launch")

    # ---- BASELINE's target kernel: the fused 3D 7-point sweep on 512^3 per GPU (z-slabs, NCCL halo for N > 1) ----
    gibbs3d = None
    view = pc.view().strip().splitlines()[:2]
    if not args.no_gibbs3d:
        del gibbs, pc
        n3 = args.n3
        nz = n3 * world
        slab3 = pmg.partition_slabs(nz, world)[rank] if world > 1 else None
        mat3 = pmg.Mat.laplace(ctx, 3, n3, n3, nz, args.kappa, slab=slab3)
        g3 = pmg.PC(ctx, "sorgibbs")
        g3.set_operator(mat3)
        g3.set_option("-pc_b200_noise", "philox")
        g3.setup()
        y3 = torch.zeros(mat3.n, dtype=torch.float64, device="cuda")
        b3 = torch.zeros(mat3.n, dtype=torch.float64, device="cuda")
        g3.apply_richardson_dev(b3, y3, its=3)
        barrier()
        e0.record(stream)
        g3.apply_richardson_dev(b3, y3, its=20)
        e1.record(stream)
        barrier()
        ms3 = torch.tensor([e0.elapsed_time(e1) / 20], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        ms3 = float(ms3.item())
        ach3 = bytes_per_update * mat3.n / (ms3 * 1e-3) / 1e9
        gibbs3d = {"workload": f"3D 7-point {n3}^3 per GPU, fused red-black sweep (sweep3d_kernel)", "sweep_ms": ms3, "dof_updates_per_s": world * mat3.n / (ms3 * 1e-3),
                   "roofline": {"bound": "hbm", "achieved": ach3, "peak": peak, "unit": "GB/s", "frac": ach3 / peak, "frac_of_nominal_8TBs": ach3 / 8000.0, "algorithmic_bytes_per_dof_update": bytes_per_update}}

        del g3, mat3, y3, b3

    # ---- config 4's building block at N = 1: one MGMC V-cycle sample on a 3D grid (fused fine level, Galerkin 27-point levels) ----
    mgmc3d = None
    if world == 1 and not args.no_mgmc3d:
        nm = args.n3 + 1 if args.n3 % 2 == 0 else args.n3  # 2^k + 1 nodes per direction for the Q1 hierarchy
        lv3 = 1
        while ((nm - 1) >> (lv3 - 1)) % 2 == 0 and (((nm - 1) >> (lv3 - 1)) + 1) ** 3 > 4096:
            lv3 += 1
        matm = pmg.Mat.laplace(ctx, 3, nm, nm, nm, args.kappa)
        m3 = pmg.PC(ctx, "gamgmc")
        m3.set_operator(matm)
        m3.set_options({"-gamgmc_pc_mg_levels": lv3, "-pc_b200_noise": "philox"})
        m3.setup()
        ym = torch.zeros(matm.n, dtype=torch.float64, device="cuda")
        bm = torch.zeros(matm.n, dtype=torch.float64, device="cuda")
        m3.apply_richardson_dev(bm, ym, its=2)
        barrier()
        e0.record(stream)
        m3.apply_richardson_dev(bm, ym, its=8)
        e1.record(stream)
        barrier()
        msm = e0.elapsed_time(e1) / 8
        mgmc3d = {"workload": f"3D 7-point {nm}^3, PCGAMGMC V(1,1), {lv3} levels, SOR-Gibbs smoother, dense Cholesky coarsest", "ms_per_sample": msm, "samples_per_s": 1e3 / msm,
                  "launches_per_sample": m3.last_stats()["launches"] / 8}
        del m3, matm, ym, bm

    if rank == 0:
        cpu = cpu_baseline_leg(args, args.cpu_samples) if (world == 1 and not args.no_cpu_baseline) else None
        out = {"metric": "mgmc_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": workload_config(args, world),
               "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(2 * 8 * nloc), "d2h_bytes_per_step": int(8 * nloc)},
               "gpu_launches": int(launches), "clocks": clk,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "note": "achieved = ALGORITHMIC bytes (32 B / DOF-update, SURVEY 8(d)) / time; the one-pass kernel actually moves ~24 B / update (traffic), so frac can exceed 1",
                            "kernel": "fine-level fused red-black sweep (sweep2d_kernel: both colours + Philox normals in one TMA-fed pass)", "algorithmic_bytes_per_dof_update": bytes_per_update,
                            "launch_ms": sweep_ms / launches_per_sweep, "peak_source": peak_src,
                            "frac_of_nominal_8TBs": achieved / 8000.0},
               "gibbs_dof_updates_per_s": world * nloc / (sweep_ms * 1e-3), "setup_s": setup_s,
               "mgmc_ms_per_sample": ms / (args.steps * S), "view": view}
        if gibbs3d is not None:
            out["gibbs3d"] = gibbs3d
        if mgmc3d is not None:
            out["mgmc3d"] = mgmc3d
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=4097)
    ap.add_argument("--n3", type=int, default=512, help="edge of the 3D grid per GPU for the gibbs3d measurement")
    ap.add_argument("--levels", type=int, default=0, help="0: 8 + log2(gpus): the coarsest grid is 33 nodes wide (SURVEY 8(d): cut at <= 33x33 + dense Cholesky) as the grid grows")
    ap.add_argument("--kappa", type=float, default=1.0)
    ap.add_argument("--samples-per-step", type=int, default=5)
    ap.add_argument("--ref-samples-per-step", type=int, default=1)
    ap.add_argument("--ref-procs", type=int, default=0, help="CPU chains of the reference arm (0: min(cores, 16, memory / 3 GB))")
    ap.add_argument("--no-gibbs3d", action="store_true", help="skip the 3D 7-point sweep measurement")
    ap.add_argument("--no-mgmc3d", action="store_true", help="skip the 3D V-cycle measurement (N = 1 only)")
    ap.add_argument("--cpu-samples", type=int, default=4)
    ap.add_argument("--bytes-per-update", type=float, default=32.0, help="algorithmic bytes per DOF update of the fine sweep (DESIGN.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.levels <= 0:
        args.levels = 8 + max(0, (max(1, args.gpus) - 1).bit_length())
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
