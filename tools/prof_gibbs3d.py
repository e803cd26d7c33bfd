"""Small driver for ncu: a few fused 3D Gibbs sweeps.  usage: prof_gibbs3d.py [n] [its]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import parmgmc_b200 as pmg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
its = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = pmg.Context(0, stream=torch.cuda.current_stream().cuda_stream, seed=0xCAFE)
mat = pmg.Mat.laplace(ctx, 3, n, n, n, kappa=1.0)
pc = pmg.PC(ctx, "sorgibbs")
pc.set_operator(mat)
pc.set_option("-pc_b200_noise", os.environ.get("NOISE", "philox"))
pc.setup()
y = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
b = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
pc.apply_richardson_dev(b, y, its=its)
torch.cuda.synchronize()
print(pc.last_stats())
