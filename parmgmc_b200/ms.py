"""MS -- the Matern sampler object of the reference (include/parmgmc/ms.h:22-41, src/ms.c), on top of the device samplers.

The reference's MS owns a DMPlex mesh, assembles the P1 finite-element precision matrix kappa^2 M + K on it (src/ms.c:86-164),
wraps it in a Richardson KSP with a PCGAMGMC preconditioner under the option prefix "ms_" (src/ms.c:327-359) and exposes
sample / save-samples / mean-and-variance / QOI calls.  DMPlex and the FE assembly are PETSc set-up code and out of scope here
(DESIGN.md section 8); everything from the assembled operator on is mirrored:

    ms = MS(ctx)                                   # MSCreate
    ms.set_from_options({"-matern_kappa": 2.0, "-ms_gamgmc_pc_mg_levels": 4})   # MSSetFromOptions (+ the "ms_" KSP / PC options)
    ms.set_grid(2, 129, 129)                       # stands in for MSSetDM: a structured grid -> the shifted Laplacian of src/problems.c
    # or ms.set_precision_matrix(mat)              # an operator assembled elsewhere (tests/golden/make_lshape.py: kappa^2 M + K on data/lshape.msh)
    ms.setup()                                     # MSSetUp
    ms.set_num_samples(1000)                       # MSSetNumSamples
    ms.set_qoi(lambda it, y: y.mean())             # MSSetQOI
    ms.begin_save_samples(); ms.sample(x); ms.end_save_samples()
    mean, var = ms.get_mean_and_var()              # MSGetMeanAndVar (mean 1/n, variance 1/(n-1), src/ms.c:220-249)
    q = ms.get_qoi_values()                        # MSGetQOIValues
"""
import numpy as np

from . import PC, Mat, NOISE_PHILOX


class MS:
    def __init__(self, ctx):  # MSCreate (src/ms.c:410-426)
        self.ctx = ctx
        self.kappa = 1.0
        self.assemble_only = False
        self.A = None
        self.pc = None
        self._grid = None
        self._pc_opts = {}
        self._pc_type = "gamgmc"  # src/ms.c:343
        self.nsamples = 1         # KSPSetTolerances(..., 1) in MSSetUp (src/ms.c:350)
        self._save = False
        self._keep = True
        self._samples = None
        self._qoi = None
        self._qois = None
        self.mean = self.var = None

    # ---- configuration --------------------------------------------------------------------------------------------------
    def set_from_options(self, opts: dict):  # MSSetFromOptions (src/ms.c:397-408); "-ms_*" are the KSP / PC options (KSPSetOptionsPrefix "ms_", :344)
        for k, v in opts.items():
            key = k.lstrip("-")
            if key == "matern_kappa":
                self.set_kappa(float(v))
            elif key == "matern_assemble_only":
                self.assemble_only = str(v).lower() not in ("0", "false", "no")
            elif key == "ms_pc_type":
                self._pc_type = str(v)
            elif key.startswith("ms_"):
                self._pc_opts["-" + key[3:]] = v
            else:
                raise ValueError(f"MS: unknown option {k}")

    def set_kappa(self, kappa):  # MSSetKappa (src/ms.c:267-276)
        if kappa < 0:
            raise ValueError("Range parameter kappa must be nonnegative")
        self.kappa = float(kappa)

    def set_assembly_only(self, flag):  # MSSetAssemblyOnly (src/ms.c:388-395)
        self.assemble_only = bool(flag)

    def set_grid(self, dim, nx, ny, nz=1, slab=None):  # in place of MSSetDM: a structured grid
        self._grid = (dim, nx, ny, nz, slab)

    def set_precision_matrix(self, mat: Mat):
        self.A = mat

    def get_precision_matrix(self):  # MSGetPrecisionMatrix (src/ms.c:379-386)
        return self.A

    def setup(self):  # MSSetUp (src/ms.c:327-359)
        if self.A is None:
            if self._grid is None:
                self._grid = (2, 17, 17, 1, None)  # CreateMeshDefault: a 4 x 4 box mesh, refined by options (src/ms.c:296-325)
            dim, nx, ny, nz, slab = self._grid
            self.A = Mat.laplace(self.ctx, dim, nx, ny, nz, kappa=self.kappa, slab=slab)
        if self.assemble_only:
            return
        self.pc = PC(self.ctx, self._pc_type)
        self.pc.set_operator(self.A)
        self.pc.set_options(self._pc_opts)
        self.pc.setup()
        self.pc.set_noise_mode(NOISE_PHILOX)
        self.set_num_samples(1)

    # ---- sampling ---------------------------------------------------------------------------------------------------------
    def set_num_samples(self, nsamples):  # MSSetNumSamples (src/ms.c:185-194)
        self.nsamples = int(nsamples)
        self._qois = np.zeros(self.nsamples)

    def set_qoi(self, qoi):  # MSSetQOI (src/ms.c:361-370): qoi(it, y) -> float
        self._qoi = qoi

    def get_qoi_values(self):  # MSGetQOIValues (src/ms.c:372-377)
        return self._qois

    def _callback(self, it, y):  # MS_SampleCallback (src/ms.c:166-174)
        if self._save and self._keep:
            self._samples[it][:] = y
        if self._qoi is not None:
            self._qois[it] = float(self._qoi(it, y))
        return 0

    def sample(self, x):  # MSSample (src/ms.c:176-183): nsamples Richardson iterations = samples, continuing the chain in x; b = 0
        if self.pc is None:
            raise RuntimeError("MS.setup() has not been called (or the sampler is assembly-only)")
        self.pc.set_sample_callback(self._callback)  # MSSetUp installs MS_SampleCallback once (src/ms.c:352); re-installed here in case a caller replaced it
        self.pc.apply_richardson(None, x, its=self.nsamples)
        return x

    def begin_save_samples(self, keep=True):  # MSBeginSaveSamples (src/ms.c:205-218)
        """keep = True stores every sample on the host like the reference; keep = False keeps only the running mean / variance,
        accumulated on the device (Welford), which is what MSEndSaveSamples needs."""
        self._save, self._keep = True, bool(keep)
        if keep:
            self._samples = [np.empty(self.A.n) for _ in range(self.nsamples)]
        else:
            self.pc.set_qoi(np.zeros(self.A.n), self.nsamples, est_mean_and_var=True)

    def get_samples(self):  # MSGetSamples (src/ms.c:196-203)
        if not (self._save and self._keep):
            raise RuntimeError("Samples can only be obtained between a call to MSBeginSaveSamples and a call to MSEndSaveSamples")
        return self._samples

    def end_save_samples(self):  # MSEndSaveSamples + MS_ComputeMeanAndVar (src/ms.c:220-261)
        n = self.nsamples
        if n <= 1:
            raise ValueError("Need at least 2 samples for variance computation")
        if self._keep:
            S = np.stack(self._samples)
            self.mean = S.sum(axis=0) / n
            self.var = ((S - self.mean) ** 2).sum(axis=0) / (n - 1)
            self._samples = None
        else:
            self.mean, self.var, seen = self.pc.get_mean_var()
            assert seen == n, (seen, n)
            self.pc.set_qoi(None, 0)
        self._save = False

    def get_mean_and_var(self):  # MSGetMeanAndVar (src/ms.c:258-265)
        return self.mean, self.var
