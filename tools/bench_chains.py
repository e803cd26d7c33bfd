"""C independent MGMC chains on one GPU, one stream and one host thread each (the small levels of one chain can fill the SMs
the other chain's small levels leave idle).  usage: bench_chains.py [n] [chains] [samples per call] [calls]"""
import json, os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import parmgmc_b200 as pmg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
C = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 120
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 5
levels = int(os.environ.get("LEVELS", "10"))
chains = []
for c in range(C):
    st = torch.cuda.Stream()
    ctx = pmg.Context(0, stream=st.cuda_stream, seed=0xCAFE + c)
    mat = pmg.Mat.laplace(ctx, 2, n, n, 1, kappa=1.0)
    pc = pmg.PC(ctx, "gamgmc"); pc.set_operator(mat); pc.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_noise": "philox"}); pc.setup()
    y = torch.zeros(mat.n, dtype=torch.float64, device="cuda")
    pc.apply_richardson_dev(None, y, its=3)
    chains.append((st, ctx, mat, pc, y))
torch.cuda.synchronize()

def work(ch):
    st, ctx, mat, pc, y = ch
    for _ in range(calls):
        pc.apply_richardson_dev(None, y, its=S)

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
th = [threading.Thread(target=work, args=(ch,)) for ch in chains]
for t in th: t.start()
for t in th: t.join()
torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"chains": C, "samples": C * S * calls, "ms": round(ms, 2), "samples_per_s": round(1e3 * C * S * calls / ms, 1), "ms_per_sample": round(ms / (C * S * calls), 4)}), flush=True)
