"""ctypes front-end of oracle/_ref/libparmgmc_ref.so: the reference's OWN mc_sor.c / pc_mcgibbs.c / parmgmc.c / pc_sorgibbs.c /
pc_chols.c / iact.c / stats.c, compiled unmodified from /root/reference against oracle/petsc_stub (see oracle/Makefile target `ref`).

TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_ref.py to pin the oracle restatement, and by
tests/golden/make_golden.py to write the committed golden vectors.  Never imported by parmgmc_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libparmgmc_ref.so")
_lib = None

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
CB = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p)


def available() -> bool:
    """True when the library exists or can be built (the reference tree is present in this container only)."""
    if os.path.exists(_PATH):
        return True
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libparmgmc_ref.so is not built (needs /root/reference)")
        L = C.CDLL(_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_mcsor_seq.argtypes = [C.c_int, i32p, i32p, f64p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int, f64p, f64p]
        L.ref_mcsor_num_colors.argtypes = [C.c_int, i32p, i32p, f64p, C.POINTER(C.c_int)]
        L.ref_mcsor_mpi.argtypes = [C.c_int, i32p, i32p, f64p, C.c_int, i32p, C.c_int, u16p, C.c_double, C.c_int, C.c_int, f64p, f64p]
        L.ref_mcgibbs_richardson.argtypes = [C.c_int, i32p, i32p, f64p, C.c_int, C.c_void_p, C.c_char_p, C.c_char_p, C.c_longlong, C.c_int, C.c_void_p, f64p, CB, C.c_void_p]
        L.ref_normal_fill.argtypes = [C.c_longlong, C.c_int, C.c_int, f64p]
        L.ref_sampler_run.argtypes = [C.c_char_p, C.c_int, i32p, i32p, f64p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                      C.c_longlong, C.c_int, C.c_void_p, f64p, CB, C.c_void_p]
        L.ref_mcsor_lrc.argtypes = [C.c_int, i32p, i32p, f64p, C.c_int, f64p, f64p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int, f64p, f64p]
        L.ref_iact.argtypes = [C.c_int, f64p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int)]
        L.ref_autocorrelation.argtypes = [C.c_int, f64p, f64p]
        L.ref_cov_errors.argtypes = [C.c_int, i32p, i32p, f64p, C.c_int, C.c_int, f64p, f64p]
        _lib = L
    return _lib


def _check(rc):
    if rc:
        raise RuntimeError(f"reference call failed ({rc}): {lib().ref_last_error().decode()}")


def _csr32(A):
    return A.n, np.ascontiguousarray(A.rowptr, np.int32), np.ascontiguousarray(A.col, np.int32), A.val


def mcsor_apply(A, b, y, coloring=None, omega=1.0, sweep=1, nsweeps=1):
    """MCSORCreate/SetUp/SetOmega/SetSweepType + nsweeps x MCSORApply (src/mc_sor.c:216-296) on one rank, in place on y.
    coloring=None: the reference's own 1-rank colouring (one colour)."""
    n, rp, cj, va = _csr32(A)
    col16 = None if coloring is None else np.ascontiguousarray(coloring.color, np.uint16)
    _check(lib().ref_mcsor_seq(n, rp, cj, va, 0 if coloring is None else coloring.ncolors, None if col16 is None else col16.ctypes.data, omega, sweep, nsweeps,
                               np.ascontiguousarray(b, np.float64), y))
    return y


def mcsor_num_colors(A) -> int:
    n, rp, cj, va = _csr32(A)
    out = C.c_int(0)
    _check(lib().ref_mcsor_num_colors(n, rp, cj, va, C.byref(out)))
    return out.value


def mcsor_apply_mpi(A, rowstart, coloring, b, y, omega=1.0, sweep=1, nsweeps=1):
    """MCSORApply_MPIAIJ (src/mc_sor.c:298-381) on len(rowstart)-1 emulated ranks, in place on the global y."""
    n, rp, cj, va = _csr32(A)
    rs = np.ascontiguousarray(rowstart, np.int32)
    _check(lib().ref_mcsor_mpi(n, rp, cj, va, rs.size - 1, rs, coloring.ncolors, np.ascontiguousarray(coloring.color, np.uint16), omega, sweep, nsweeps,
                               np.ascontiguousarray(b, np.float64), y))
    return y


def mcgibbs_richardson(A, b, y, its, seed, coloring=None, omega=None, sweep_opt="", callback=None):
    """PCSetFromOptions + PCSetUp + PCApplyRichardson of PCMCGIBBS (src/pc_mcgibbs.c:155-251), noise from the library's
    global rander48 stream through VecSetRandomStandardNormal (src/parmgmc.c:100-110)."""
    n, rp, cj, va = _csr32(A)
    col16 = None if coloring is None else np.ascontiguousarray(coloring.color, np.uint16)
    bptr = None if b is None else np.ascontiguousarray(b, np.float64).ctypes.data
    if callback is None:
        cbf = C.cast(None, CB)
    else:
        cbf = CB(lambda it, yp, m, _c: int(callback(int(it), np.ctypeslib.as_array(yp, shape=(m,))) or 0))
    _check(lib().ref_mcgibbs_richardson(n, rp, cj, va, 0 if coloring is None else coloring.ncolors, None if col16 is None else col16.ctypes.data,
                                        b"" if omega is None else repr(float(omega)).encode(), sweep_opt.encode(), seed, its, bptr, y, cbf, None))
    return y


def normal_fill(seed, n, ncalls=1):
    out = np.empty(n * ncalls, np.float64)
    _check(lib().ref_normal_fill(seed, n, ncalls, out))
    return out


def sampler_run(pctype, A, b, y, its, seed, coloring=None, opts=(), lrc=None, callback=None):
    """PCSetFromOptions + PCSetUp + PCApplyRichardson (its > 0) or PCApply (its == 0) of `pctype` in {"mcgibbs", "sorgibbs",
    "cholsampler"} on one rank (src/pc_sorgibbs.c:76-134, :181-262; src/pc_chols.c:100-342), optionally on the MATLRC operator
    A + B diag(S) B^T (lrc = (B, S)); at most two (option, value) pairs; noise from the library's global rander48 stream."""
    n, rp, cj, va = _csr32(A)
    col16 = None if coloring is None else np.ascontiguousarray(coloring.color, np.uint16)
    bb = None if b is None else np.ascontiguousarray(b, np.float64)
    k, Bf, Sf = 0, None, None
    if lrc is not None:
        Bf = np.asfortranarray(np.asarray(lrc[0], np.float64))
        Sf = np.ascontiguousarray(lrc[1], np.float64)
        k = Bf.shape[1]
    o = [(str(a).encode(), str(v).encode()) for a, v in opts] + [(b"", b""), (b"", b"")]
    cbf = C.cast(None, CB) if callback is None else CB(lambda it, yp, m, _c: int(callback(int(it), np.ctypeslib.as_array(yp, shape=(m,))) or 0))
    _check(lib().ref_sampler_run(pctype.encode(), n, rp, cj, va, k, None if Bf is None else Bf.ctypes.data, None if Sf is None else Sf.ctypes.data,
                                 0 if coloring is None else coloring.ncolors, None if col16 is None else col16.ctypes.data, o[0][0], o[0][1], o[1][0], o[1][1],
                                 seed, its, None if bb is None else bb.ctypes.data, y, cbf, None))
    return y


def assemble_laplace2d(mx, my, kappa):
    """MatAssembleShiftedLaplaceFD (src/problems.c:14-75) on an mx x my DMDA, one rank: the dense (mx my) x (mx my) operator."""
    n = mx * my
    d = np.zeros(n * n, np.float64)
    f = lib().ref_assemble_laplace2d
    f.argtypes = [C.c_int, C.c_int, C.c_double, np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
    f.restype = C.c_int
    _check(f(mx, my, float(kappa), d))
    return d.reshape(n, n).T.copy()  # column-major -> [row, col]


def mcsor_apply_lrc(A, B, S, b, y, coloring=None, omega=1.0, sweep=1, nsweeps=1):
    """MCSORCreate / SetUp / Apply on the MATLRC operator A + B diag(S) B^T: the sweep on A followed by MCSORPostSOR_LRC with the
    correction MCSORBuildLRCCorrection built at set-up (src/mc_sor.c:101-112, :480-544, :565-595)."""
    n, rp, cj, va = _csr32(A)
    Bf = np.asfortranarray(np.asarray(B, np.float64))
    col16 = None if coloring is None else np.ascontiguousarray(coloring.color, np.uint16)
    _check(lib().ref_mcsor_lrc(n, rp, cj, va, Bf.shape[1], Bf.ravel(order="K"), np.ascontiguousarray(S, np.float64), 0 if coloring is None else coloring.ncolors,
                               None if col16 is None else col16.ctypes.data, omega, sweep, nsweeps, np.ascontiguousarray(b, np.float64), y))
    return y


def iact(x):
    x = np.ascontiguousarray(x, np.float64)
    tau, valid = C.c_double(), C.c_int()
    acf = np.empty_like(x)
    _check(lib().ref_iact(x.size, x, C.byref(tau), acf.ctypes.data, C.byref(valid)))
    return tau.value, bool(valid.value), acf


def autocorrelation(x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    _check(lib().ref_autocorrelation(x.size, x, out))
    return out


def cov_errors(A, samples):
    """EstimateCovarianceMatErrors (src/stats.c:94-117); samples[s, chain, :]."""
    samples = np.ascontiguousarray(samples, np.float64)
    S, K, n = samples.shape
    _, rp, cj, va = _csr32(A)
    errs = np.empty(S, np.float64)
    _check(lib().ref_cov_errors(n, rp, cj, va, K, S, samples.ravel(), errs))
    return errs
