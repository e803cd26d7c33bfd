"""pymgmc -- the reference's Python module (python/main.cc:20-52) for the device samplers.

The reference builds a pybind11 module with two functions for petsc4py users: PCSetSampleCallback(pc, fn), which installs a
Python callable fn(it, y) as the sample callback of a ParMGMC PC (:26-40), and seed(s), which seeds the library's global random
number stream (:42-51).  petsc4py / PETSc do not exist in this image; the same two functions are offered for this package's own
PC objects (parmgmc_b200.PC, the ctypes mirror of the C ABI), so a Python driver written against the reference module keeps its shape:

    import parmgmc_b200 as pmg
    from parmgmc_b200 import pymgmc
    ctx = pmg.Context(0); pymgmc.use(ctx)
    pymgmc.seed(1234)
    pymgmc.PCSetSampleCallback(pc, lambda it, y: ...)
"""
_ctx = None


def use(ctx):
    """The context whose noise stream seed() addresses (the reference has one global PetscRandom, src/parmgmc.c:38-68)."""
    global _ctx
    _ctx = ctx


def PCSetSampleCallback(pc, cb):  # python/main.cc:26-40
    pc.set_sample_callback(cb)


def seed(s):  # python/main.cc:42-51: PetscRandomSetSeed + PetscRandomSeed on the global stream
    ctx = _ctx
    if ctx is None:
        raise RuntimeError("pymgmc.use(ctx) first: the device library keeps its noise stream per context")
    ctx.set_seed(int(s))
