"""Small driver for ncu: a few fused Gibbs sweeps (or MGMC samples) on the config-2 grid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import parmgmc_b200 as pmg

what = sys.argv[1] if len(sys.argv) > 1 else "gibbs"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4097
its = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ctx = pmg.Context(0, stream=torch.cuda.current_stream().cuda_stream, seed=0xCAFE)
mat = pmg.Mat.laplace(ctx, 2, n, n, kappa=1.0)
pc = pmg.PC(ctx, "sorgibbs" if what == "gibbs" else "gamgmc")
pc.set_operator(mat)
if what != "gibbs":
    pc.set_option("-gamgmc_pc_mg_levels", 10)
pc.set_option("-pc_b200_noise", os.environ.get("NOISE", "philox"))
pc.setup()
y = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
b = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
pc.apply_richardson_dev(b, y, its=its)
torch.cuda.synchronize()
print(pc.last_stats())
