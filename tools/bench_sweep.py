"""Times the fine-level fused sweep and one MGMC sample (device resident).  usage: bench_sweep.py [n] [reps] [levels]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import parmgmc_b200 as pmg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 9
what = sys.argv[4] if len(sys.argv) > 4 else "both"
dim = int(os.environ.get("DIM", "2"))
stream = torch.cuda.current_stream()
ctx = pmg.Context(0, stream=stream.cuda_stream, seed=0xCAFE)
mat = pmg.Mat.laplace(ctx, dim, n, n, n if dim == 3 else 1, kappa=1.0)
y = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
b = None if os.environ.get("NOB") else torch.zeros(mat.n, dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {}
if what in ("both", "gibbs"):
    pc = pmg.PC(ctx, "sorgibbs"); pc.set_operator(mat); pc.set_option("-pc_b200_noise", os.environ.get("NOISE", "philox")); pc.setup()
    pc.apply_richardson_dev(b, y, its=5)
    torch.cuda.synchronize()
    e0.record(stream); pc.apply_richardson_dev(b, y, its=reps); e1.record(stream); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    out["sweep_us"] = round(us, 2); out["GDOF/s"] = round(mat.n / us / 1e3, 2); out["alg_GB/s"] = round(32 * mat.n / us / 1e3, 1)
if what in ("both", "mg"):
    mg = pmg.PC(ctx, "gamgmc"); mg.set_operator(mat); mg.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_noise": "philox"}); mg.setup()
    mg.apply_richardson_dev(b, y, its=3)
    torch.cuda.synchronize()
    e0.record(stream); mg.apply_richardson_dev(b, y, its=reps); e1.record(stream); torch.cuda.synchronize()
    out["mg_ms_per_sample"] = round(e0.elapsed_time(e1) / reps, 4); out["launches_per_sample"] = mg.last_stats()["launches"] / reps
print(os.environ.get("PMG_STREAM_BY", "-"), out, flush=True)
