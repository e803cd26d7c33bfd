"""Weak-scaled fused Gibbs sweeps on N GPUs (BASELINE config 3: 3D 7-point, z-slabs, NCCL halo).
usage: torchrun --nproc-per-node N tools/bench_gibbs_mgpu.py [dim] [n per GPU edge] [sweeps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import parmgmc_b200 as pmg

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.current_stream()
ctx = pmg.Context(local, stream=stream.cuda_stream, seed=0xCAFE)
if world > 1:
    uid = [pmg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
nslow = n * world
slab = pmg.partition_slabs(nslow, world)[rank] if world > 1 else None
mat = pmg.Mat.laplace(ctx, dim, n, n if dim == 3 else nslow, nslow if dim == 3 else 1, kappa=1.0, slab=slab)
pc = pmg.PC(ctx, "sorgibbs"); pc.set_operator(mat); pc.set_option("-pc_b200_noise", "philox"); pc.setup()
y = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
b = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
pc.apply_richardson_dev(b, y, its=3)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); pc.apply_richardson_dev(b, y, its=reps); e1.record(stream); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
us = 1e3 * float(ms.item()) / reps
if rank == 0:
    tot = mat.n * world
    print(json.dumps({"n_gpus": world, "dim": dim, "grid_per_gpu": [n] * dim, "sweep_us": round(us, 1), "GDOF_per_s": round(tot / us / 1e3, 2), "alg_GB_per_s_per_gpu": round(32 * mat.n / us / 1e3, 1),
                      "launches_per_sweep": pc.last_stats()["launches"] / reps}), flush=True)
if world > 1:
    dist.destroy_process_group()
