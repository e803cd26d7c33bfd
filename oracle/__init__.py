"""ctypes front-end of the CPU oracle (oracle/build/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py -- never by parmgmc_b200.
Each wrapper names the reference file:line its C body restates (see oracle.h and the .c files).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")

SOR_FORWARD, SOR_BACKWARD, SOR_SYMMETRIC = 1, 2, 3
KIND_SORGIBBS, KIND_MCGIBBS, KIND_CHOL = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no dependencies)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None

i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class Noise(C.Structure):
    """orc_noise: the single process-global stream of the reference (src/parmgmc.c:38-42)."""

    _fields_ = [("mode", C.c_int), ("tape_ptr", C.c_void_p), ("tape_len", C.c_int64), ("tape_pos", C.c_int64),
                ("seed_", C.c_uint64), ("call", C.c_uint64), ("x48", C.c_uint64),
                ("grid_n", C.c_int64), ("grid_nx", C.c_int64), ("grid_pad", C.c_int64)]

    @staticmethod
    def tape(z) -> "Noise":
        z = np.ascontiguousarray(z, dtype=np.float64).ravel()
        ns = Noise()
        lib().orc_noise_init_tape(C.byref(ns), z.ctypes.data, z.size)
        ns._keep = z
        return ns

    @staticmethod
    def philox(seed: int, grid=None) -> "Noise":
        """grid = (nx, ny, nz) of a matrix-free grid operator: its blocks are keyed on the padded index
        (k ny + j) pitch + i, pitch = nx rounded up to 4 (parmgmc_b200/csrc/philox.cuh)."""
        ns = Noise()
        lib().orc_noise_init_philox(C.byref(ns), seed)
        if grid is not None:
            nx, ny, nz = grid
            ns.grid_n, ns.grid_nx, ns.grid_pad = nx * ny * nz, nx, (-nx) % 4
        return ns

    @staticmethod
    def rander48(seed: int = 0x12345678) -> "Noise":
        ns = Noise()
        lib().orc_noise_init_rander48(C.byref(ns), seed)
        return ns


CB = C.CFUNCTYPE(C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_void_p)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.orc_laplace_nnz.restype = C.c_int64
    L.orc_laplace_nnz.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64]
    L.orc_laplace_csr.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, i64p, i32p, f64p]
    L.orc_diag_ptrs.argtypes = [C.c_int64, i64p, i32p, i64p]
    L.orc_idiag.argtypes = [C.c_int64, f64p, i64p, C.c_double, f64p]
    L.orc_sweep_seq.argtypes = [C.c_int64, i64p, i32p, f64p, i64p, f64p, C.c_double, C.c_int, i64p, i32p, C.c_int, f64p, f64p]
    L.orc_mcsor_apply.argtypes = L.orc_sweep_seq.argtypes
    L.orc_coloring_lists.argtypes = [C.c_int64, i32p, C.c_int, i64p, i32p]
    L.orc_coloring_greedy.argtypes = [C.c_int64, i64p, i32p, i32p]
    L.orc_coloring_levelset.argtypes = [C.c_int64, i64p, i32p, i32p]
    L.orc_coloring_valid.argtypes = [C.c_int64, i64p, i32p, f64p, i32p]
    L.orc_part_create.restype = C.c_void_p
    L.orc_part_create.argtypes = [C.c_int64, i64p, i32p, f64p, C.c_int, i64p, C.c_int, i32p, C.c_double]
    L.orc_part_destroy.argtypes = [C.c_void_p]
    L.orc_part_sweep.argtypes = [C.c_void_p, C.c_int, f64p, f64p, C.c_int]
    L.orc_part_ghost_count.restype = C.c_int64
    L.orc_part_ghost_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_part_ghost_index.argtypes = [C.c_void_p, C.c_int, C.c_int, i64p]
    L.orc_philox4x32_10.argtypes = [np.ctypeslib.ndpointer(np.uint32), np.ctypeslib.ndpointer(np.uint32), np.ctypeslib.ndpointer(np.uint32)]
    L.orc_normal_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, f64p]
    L.orc_noise_init_tape.argtypes = [C.POINTER(Noise), C.c_void_p, C.c_int64]
    L.orc_noise_init_philox.argtypes = [C.POINTER(Noise), C.c_uint64]
    L.orc_noise_init_rander48.argtypes = [C.POINTER(Noise), C.c_uint64]
    L.orc_noise_fill.argtypes = [C.POINTER(Noise), C.c_int64, C.c_int64, f64p]
    L.orc_sqrtdiag.argtypes = [C.c_int64, f64p, i64p, C.c_double, f64p]
    L.orc_prepare_rhs.argtypes = [C.c_int64, C.c_void_p, f64p, f64p, f64p]
    L.orc_gibbs_richardson.argtypes = [C.c_int64, i64p, i32p, f64p, C.c_double, C.c_int, i64p, i32p, C.c_int, C.POINTER(Noise), C.c_void_p, f64p, C.c_int64, CB, C.c_void_p]
    L.orc_potrf_lower.argtypes = [C.c_int64, f64p]
    L.orc_trsv_lower.argtypes = [C.c_int64, f64p, C.c_int, f64p]
    L.orc_chol_sample.argtypes = [C.c_int64, f64p, C.POINTER(Noise), f64p, f64p]
    L.orc_spmv.argtypes = [C.c_int64, i64p, i32p, f64p, f64p, f64p]
    L.orc_q1_nnz.restype = C.c_int64
    L.orc_q1_nnz.argtypes = [C.c_int, i64p, i64p]
    L.orc_q1_coarse_dims.argtypes = [C.c_int, i64p, i64p]
    L.orc_q1_interp.argtypes = [C.c_int, i64p, i64p, i64p, i32p, f64p]
    L.orc_mg_create.restype = C.c_void_p
    L.orc_mg_create.argtypes = [C.c_int]
    L.orc_mg_destroy.argtypes = [C.c_void_p]
    L.orc_mg_set_fine.argtypes = [C.c_void_p, C.c_int64, i64p, i32p, f64p]
    L.orc_mg_set_interp.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, i64p, i32p, f64p]
    L.orc_mg_galerkin.argtypes = [C.c_void_p]
    L.orc_mg_build_geometric.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double]
    L.orc_mg_level_dims.argtypes = [C.c_void_p, C.c_int, i64p]
    L.orc_mg_level_n.restype = C.c_int64
    L.orc_mg_level_n.argtypes = [C.c_void_p, C.c_int]
    L.orc_mg_level_nnz.restype = C.c_int64
    L.orc_mg_level_nnz.argtypes = [C.c_void_p, C.c_int]
    L.orc_mg_level_csr.argtypes = [C.c_void_p, C.c_int, i64p, i32p, f64p]
    L.orc_mg_set_smoother.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orc_mg_set_coarse.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orc_mg_setup.argtypes = [C.c_void_p]
    L.orc_mg_apply.argtypes = [C.c_void_p, C.POINTER(Noise), f64p, f64p]
    L.orc_gamgmc_richardson.argtypes = [C.c_void_p, C.POINTER(Noise), f64p, f64p, C.c_int64, C.c_int, CB, C.c_void_p]
    L.orc_autocorrelation.argtypes = [C.c_int64, f64p, f64p]
    L.orc_iact.argtypes = [C.c_int64, f64p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int)]
    L.orc_cov_errors.argtypes = [C.c_int64, f64p, C.c_int64, C.c_int64, f64p, f64p]
    L.orc_gelman_rubin.restype = C.c_double
    L.orc_gelman_rubin.argtypes = [C.c_int64, C.c_int64, C.c_int64, f64p]
    _lib = L
    return L


_NULL_CB = C.cast(None, CB)


class CSR:
    """CSR with int64 row offsets, int32 columns (ascending per row), float64 values."""

    def __init__(self, n, rowptr, col, val):
        self.n = int(n)
        self.rowptr = np.ascontiguousarray(rowptr, np.int64)
        self.col = np.ascontiguousarray(col, np.int32)
        self.val = np.ascontiguousarray(val, np.float64)

    @property
    def nnz(self):
        return int(self.rowptr[-1])

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val, self.col, self.rowptr), shape=(self.n, self.n))

    def diag_ptrs(self):
        d = np.empty(self.n, np.int64)
        missing = lib().orc_diag_ptrs(self.n, self.rowptr, self.col, d)
        if missing:
            raise ValueError(f"{missing} rows have no diagonal entry")
        return d


def laplace(dim: int, nx: int, ny: int, nz: int = 1, kappa: float = 1.0) -> CSR:
    """src/problems.c:14-75 (+ the 3D extension of SURVEY F8)."""
    L = lib()
    if dim == 2:
        nz = 1
    n = nx * ny * nz
    nnz = L.orc_laplace_nnz(dim, nx, ny, nz)
    rowptr, col, val = np.empty(n + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    L.orc_laplace_csr(dim, nx, ny, nz, kappa, rowptr, col, val)
    return CSR(n, rowptr, col, val)


class Coloring:
    """ISColoring analogue: colour of each row plus per-colour ascending row lists."""

    def __init__(self, color, ncolors=None):
        self.color = np.ascontiguousarray(color, np.int32)
        self.ncolors = int(self.color.max()) + 1 if ncolors is None else int(ncolors)
        n = self.color.size
        self.ptr = np.empty(self.ncolors + 1, np.int64)
        self.rows = np.empty(n, np.int32)
        if lib().orc_coloring_lists(n, self.color, self.ncolors, self.ptr, self.rows):
            raise ValueError("colour out of range")

    @staticmethod
    def single(n):
        """src/mc_sor.c:397-410: every row colour 0 (the 1-rank reference)."""
        return Coloring(np.zeros(n, np.int32), 1)

    @staticmethod
    def greedy(A: CSR):
        c = np.empty(A.n, np.int32)
        k = lib().orc_coloring_greedy(A.n, A.rowptr, A.col, c)
        return Coloring(c, k)

    @staticmethod
    def levelset(A: CSR):
        c = np.empty(A.n, np.int32)
        k = lib().orc_coloring_levelset(A.n, A.rowptr, A.col, c)
        return Coloring(c, k)

    @staticmethod
    def parity(dims, ncolors_per_dim=2):
        """red-black (i+j+k)%2 for star stencils / 2^d colours (i%2,j%2,k%2) for box stencils."""
        nx, ny, nz = (list(dims) + [1, 1])[:3]
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        if ncolors_per_dim == 2:
            return Coloring(((i + j + k) % 2).astype(np.int32).ravel(), 2)
        c = (i % 2) + 2 * (j % 2) + (4 * (k % 2) if nz > 1 else 0)
        return Coloring(c.astype(np.int32).ravel(), 8 if nz > 1 else 4)

    def violations(self, A: CSR) -> int:
        return lib().orc_coloring_valid(A.n, A.rowptr, A.col, A.val, self.color)


class MCSOR:
    """src/mc_sor.c MCSOR object on one rank (MCSORCreate/SetUp/Apply/SetOmega/SetSweepType)."""

    def __init__(self, A: CSR, coloring: Coloring | None = None, omega: float = 1.0, sweep: int = SOR_FORWARD):
        self.A = A
        self.coloring = coloring or Coloring.single(A.n)
        self.diagptr = A.diag_ptrs()
        self.sweep = sweep
        self.set_omega(omega)

    def set_omega(self, omega):
        self.omega = float(omega)
        self.idiag = np.empty(self.A.n, np.float64)
        lib().orc_idiag(self.A.n, self.A.val, self.diagptr, self.omega, self.idiag)

    def apply(self, b, y, sweep=None):
        """MCSORApply (src/mc_sor.c:216-239), in place on y."""
        A, c = self.A, self.coloring
        lib().orc_mcsor_apply(A.n, A.rowptr, A.col, A.val, self.diagptr, self.idiag, self.omega, c.ncolors, c.ptr, c.rows,
                              self.sweep if sweep is None else sweep, np.ascontiguousarray(b, np.float64), y)
        return y


def sqrtdiag(A: CSR, omega: float):
    out = np.empty(A.n, np.float64)
    lib().orc_sqrtdiag(A.n, A.val, A.diag_ptrs(), omega, out)
    return out


def _wrap_cb(cb, n):
    if cb is None:
        return _NULL_CB
    def tramp(it, yptr, _ctx):
        y = np.ctypeslib.as_array(yptr, shape=(n,))
        r = cb(int(it), y)
        return int(r or 0)
    return CB(tramp)


def gibbs_richardson(A: CSR, b, y, its, noise: Noise, coloring: Coloring | None = None, omega=1.0, sweep=SOR_FORWARD, callback=None):
    """PCApplyRichardson_MulticolorGibbs (src/pc_mcgibbs.c:155-188); with omega=1, forward, this is
    also PCApplyRichardson_SORGibbs (src/pc_sorgibbs.c:115-134)."""
    c = coloring or Coloring.single(A.n)
    bptr = None if b is None else np.ascontiguousarray(b, np.float64).ctypes.data
    cbf = _wrap_cb(callback, A.n)
    err = lib().orc_gibbs_richardson(A.n, A.rowptr, A.col, A.val, omega, c.ncolors, c.ptr, c.rows, sweep, C.byref(noise), bptr, y, its, cbf, None)
    if err:
        raise RuntimeError(f"oracle gibbs_richardson failed ({err})")
    return y


# ---- MATLRC operators A + B diag(S) B^T (parity of this part is UNPINNED: the PETSc stub of oracle/_ref has no MATLRC;
#      tests/test_oracle.py validates it through exact invariance of N(0, (A + B S B^T)^-1)) ---------------------------------
def lrc_build_correction(A: CSR, B, S, coloring: Coloring | None, omega: float, sweep: int):
    """MCSORBuildLRCCorrection (src/mc_sor.c:480-544): Bb = M^-1 B (S^-1 + B^T M^-1 B)^-1, where M^-1 is one deterministic
    sweep from zero (MCSORApplyAsDetSOR, :546-551) in direction `sweep`."""
    B = np.asarray(B, np.float64)
    n, k = B.shape
    mc = MCSOR(A, coloring, omega, sweep)
    Cm = np.empty((n, k))
    for j in range(k):  # :499-510, column by column
        Cm[:, j] = mc.apply(np.ascontiguousarray(B[:, j]), np.zeros(n))
    T = B.T @ Cm + np.diag(1.0 / np.asarray(S, np.float64))  # :513-527
    return Cm @ np.linalg.inv(T)  # :528-535


def lrc_mcsor_apply(A: CSR, B, Bb, b, y, coloring, omega, sweep):
    """MCSORApply with postsor = MCSORPostSOR_LRC (src/mc_sor.c:216-239, :101-112); Bb = {direction: matrix}."""
    mc = MCSOR(A, coloring, omega, sweep)
    for d in ((SOR_FORWARD, SOR_BACKWARD) if sweep == SOR_SYMMETRIC else (sweep,)):
        mc.apply(b, y, d)
        y -= Bb[d] @ (B.T @ y)
    return y


def lrc_gibbs_richardson(A: CSR, B, S, b, y, its, noise: Noise, coloring: Coloring | None = None, omega=1.0, sweep=SOR_FORWARD, omega_build=1.0):
    """PCApplyRichardson_MulticolorGibbs (src/pc_mcgibbs.c:155-188) on a MATLRC operator: per directional sweep
    w = b + sqrtdiag z + B (sqrt|S| eta) (PrepareRHS_LRC :130-140: n draws, then k), MCSORApply on the base matrix, then
    y -= Bb_dir (B^T y).  Bb is built with a temporary MCSOR at its own default omega (src/mc_sor.c:583-593) = omega_build."""
    B = np.asarray(B, np.float64)
    S = np.asarray(S, np.float64)
    n, k = B.shape
    Bb = {d: lrc_build_correction(A, B, S, coloring, omega_build, d) for d in (SOR_FORWARD, SOR_BACKWARD)}
    sd = sqrtdiag(A, omega)
    sqrtS = np.sqrt(np.abs(S))
    mc = MCSOR(A, coloring, omega, SOR_FORWARD)
    b = np.zeros(n) if b is None else np.asarray(b, np.float64)
    for _ in range(its):
        for d in ((SOR_FORWARD, SOR_BACKWARD) if sweep == SOR_SYMMETRIC else (sweep,)):
            w = noise_fill(noise, n) * sd + b
            w = B @ (noise_fill(noise, k) * sqrtS) + w
            mc.apply(w, y, d)
            y -= Bb[d] @ (B.T @ y)
    return y


def normal_philox(seed, call, row0, n):
    out = np.empty(n, np.float64)
    lib().orc_normal_philox(seed, call, row0, n, out)
    return out


def noise_fill(noise: Noise, n, row0=0):
    out = np.empty(n, np.float64)
    if lib().orc_noise_fill(C.byref(noise), row0, n, out):
        raise RuntimeError("noise tape exhausted")
    return out


class Partitioned:
    """MCSORApply_MPIAIJ with ranks emulated in one process (src/mc_sor.c:152-214, :298-381)."""

    def __init__(self, A: CSR, rowstart, coloring: Coloring, omega=1.0):
        self.A, self.rowstart = A, np.ascontiguousarray(rowstart, np.int64)
        self.nranks = self.rowstart.size - 1
        self.coloring = coloring
        self._h = lib().orc_part_create(A.n, A.rowptr, A.col, A.val, self.nranks, self.rowstart, coloring.ncolors, coloring.color, omega)

    def sweep(self, b, y, sweep=SOR_FORWARD, nthreads=1):
        lib().orc_part_sweep(self._h, sweep, np.ascontiguousarray(b, np.float64), y, nthreads)
        return y

    def ghost_index(self, rank, color):
        n = lib().orc_part_ghost_count(self._h, rank, color)
        out = np.empty(n, np.int64)
        if n:
            lib().orc_part_ghost_index(self._h, rank, color, out)
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_part_destroy(self._h)
            self._h = None


class MG:
    """PCGAMGMC with the geometric (`-pc_gamgmc_mg_type mg`) hierarchy: src/pc_gamgmc.c:227-356 +
    PCMG semantics of SURVEY Appendix A.  Level 0 = coarsest."""

    def __init__(self, nlevels):
        self.nlevels = nlevels
        self._h = lib().orc_mg_create(nlevels)
        self._keep = []

    @staticmethod
    def geometric(dim, nx, ny, nz, kappa, nlevels):
        mg = MG(nlevels)
        err = lib().orc_mg_build_geometric(mg._h, dim, nx, ny, nz, kappa)
        if err:
            raise RuntimeError(f"orc_mg_build_geometric failed ({err})")
        return mg

    def level_dims(self, l):
        d = np.empty(3, np.int64)
        lib().orc_mg_level_dims(self._h, l, d)
        return tuple(int(x) for x in d)

    def level_csr(self, l) -> CSR:
        L = lib()
        n, nnz = L.orc_mg_level_n(self._h, l), L.orc_mg_level_nnz(self._h, l)
        rp, col, val = np.empty(n + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
        L.orc_mg_level_csr(self._h, l, rp, col, val)
        return CSR(n, rp, col, val)

    def set_smoother(self, level, kind=KIND_SORGIBBS, omega=1.0, sweep=SOR_FORWARD, its=1, coloring: Coloring | None = None):
        cp = None
        if coloring is not None:
            self._keep.append(coloring)
            cp = coloring.color.ctypes.data
        fn = lib().orc_mg_set_coarse if level == 0 else None
        if level == 0:
            err = fn(self._h, kind, omega, sweep, its, coloring.ncolors if coloring else 0, cp)
        else:
            err = lib().orc_mg_set_smoother(self._h, level, kind, omega, sweep, its, coloring.ncolors if coloring else 0, cp)
        if err:
            raise RuntimeError(f"set_smoother failed ({err})")

    def setup(self):
        err = lib().orc_mg_setup(self._h)
        if err:
            raise RuntimeError(f"orc_mg_setup failed ({err})")

    def apply(self, noise: Noise, b, x):
        err = lib().orc_mg_apply(self._h, C.byref(noise), np.ascontiguousarray(b, np.float64), x)
        if err:
            raise RuntimeError(f"orc_mg_apply failed ({err})")
        return x

    def richardson(self, noise: Noise, b, y, its, guesszero=False, callback=None):
        n = lib().orc_mg_level_n(self._h, self.nlevels - 1)
        err = lib().orc_gamgmc_richardson(self._h, C.byref(noise), np.ascontiguousarray(b, np.float64), y, its, int(guesszero), _wrap_cb(callback, n), None)
        if err:
            raise RuntimeError(f"orc_gamgmc_richardson failed ({err})")
        return y

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_mg_destroy(self._h)
            self._h = None


def potrf_lower(a_colmajor):
    a = np.array(a_colmajor, np.float64, order="F", copy=True)
    n = a.shape[0]
    flat = np.ascontiguousarray(a.T).ravel()  # column-major bytes
    info = lib().orc_potrf_lower(n, flat)
    if info:
        raise np.linalg.LinAlgError(f"leading minor {info} not positive definite")
    return flat


def chol_sample(lflat, n, noise: Noise, b):
    y = np.empty(n, np.float64)
    if lib().orc_chol_sample(n, lflat, C.byref(noise), np.ascontiguousarray(b, np.float64), y):
        raise RuntimeError("noise tape exhausted")
    return y


def iact(x):
    x = np.ascontiguousarray(x, np.float64)
    tau, valid = C.c_double(), C.c_int()
    if lib().orc_iact(x.size, x, C.byref(tau), None, C.byref(valid)):
        raise ValueError("too few data points")
    return tau.value, bool(valid.value)


def autocorrelation(x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    lib().orc_autocorrelation(x.size, x, out)
    return out


def cov_errors(A_dense, samples):
    """samples[s, chain, :] -> relative Frobenius error per sample index (src/stats.c:94-117)."""
    samples = np.ascontiguousarray(samples, np.float64)
    S, K, n = samples.shape
    errs = np.empty(S, np.float64)
    if lib().orc_cov_errors(n, np.ascontiguousarray(A_dense, np.float64), K, S, samples, errs):
        raise np.linalg.LinAlgError("singular")
    return errs


def gelman_rubin(vals):
    vals = np.ascontiguousarray(vals, np.float64)
    chains, n = vals.shape
    return lib().orc_gelman_rubin(0, chains, n, vals)
