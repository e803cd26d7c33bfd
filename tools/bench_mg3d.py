"""Times one MGMC V-cycle sample on a 3D grid (device resident).  usage: bench_mg3d.py [n] [reps] [levels]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import parmgmc_b200 as pmg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 6
stream = torch.cuda.current_stream()
ctx = pmg.Context(0, stream=stream.cuda_stream, seed=0xCAFE)
mat = pmg.Mat.laplace(ctx, 3, n, n, n, kappa=1.0)
y = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
b = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
mg = pmg.PC(ctx, "gamgmc"); mg.set_operator(mat); mg.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_noise": "philox"}); mg.setup()
mg.apply_richardson_dev(b, y, its=2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); mg.apply_richardson_dev(b, y, its=reps); e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print({"n": n, "levels": levels, "fused_top": os.environ.get("PMG_NO_FUSED_MG3") is None, "ms_per_sample": round(ms, 3), "samples_per_s": round(1e3 / ms, 2),
       "launches_per_sample": mg.last_stats()["launches"] / reps}, flush=True)
if os.environ.get("PMG_PROFILE"):
    mg.set_option("-pc_b200_profile", "1")
    mg.apply_richardson_dev(b, y, its=reps)
    torch.cuda.synchronize()
    import json
    for row in mg.profile():
        print(row, flush=True)
