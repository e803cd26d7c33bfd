"""Times the CSR/SELL colour sweep (K2 of SURVEY 8(d)).  usage: bench_csr.py [dim] [n] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch
import parmgmc_b200 as pmg

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 160
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
stream = torch.cuda.current_stream()
ctx = pmg.Context(0, stream=stream.cuda_stream, seed=0xCAFE)
dims = (n, n, n) if dim == 3 else (n, n, 1)
# the shifted Laplacian of src/problems.c:14-75 assembled here (no oracle outside the tests): kappa^2 I + h (T x I x I + ...),
# T = tridiag(-1, number of existing neighbours, -1), h = 1 / (n - 1)^2
def _t(m):
    deg = np.full(m, 2.0); deg[0] = deg[-1] = 1.0
    return sp.diags([-np.ones(m - 1), deg, -np.ones(m - 1)], [-1, 0, 1], format="csr")
h, I, T = 1.0 / (n - 1) ** 2, sp.identity(n, format="csr"), _t(n)
L = sp.kron(sp.kron(I, I), T) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(T, I), I) if dim == 3 else sp.kron(I, T) + sp.kron(T, I)
As = (1.0 * sp.identity(L.shape[0], format="csr") + h * L).tocsr(); As.sort_indices()
class A:
    n, nnz, rowptr, col, val = As.shape[0], As.nnz, As.indptr.astype(np.int64), As.indices.astype(np.int32), As.data
mat = pmg.Mat.from_csr(ctx, A.rowptr, A.col, A.val)
idx = np.arange(A.n)
i, j, k = idx % n, (idx // n) % n, idx // (n * n)
mat.set_coloring(((i + j + k) & 1).astype(np.int32), 2)
pc = pmg.PC(ctx, "sorgibbs"); pc.set_operator(mat); pc.set_option("-pc_b200_noise", os.environ.get("NOISE", "philox")); pc.setup()
y = torch.zeros(A.n, dtype=torch.float64, device="cuda"); b = torch.zeros(A.n, dtype=torch.float64, device="cuda")
pc.apply_richardson_dev(b, y, its=3); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); pc.apply_richardson_dev(b, y, its=reps); e1.record(stream); torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / reps
nnz_row = A.nnz / A.n
bytes_row = 12 * nnz_row + 48
print({"dim": dim, "n": n, "rows": A.n, "sweep_us": round(us, 1), "GDOF/s": round(A.n / us / 1e3, 2), "K2_alg_GB/s": round(bytes_row * A.n / us / 1e3, 1), "bytes_per_row": round(bytes_row, 1)})
