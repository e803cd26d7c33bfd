/* host_common.h -- helpers shared by the example-shaped host programs (test infrastructure). */
#ifndef HOST_COMMON_H
#define HOST_COMMON_H
#include <math.h>
#include <petsc.h>

/* MatAssembleShiftedLaplaceFD semantics (src/problems.c:14-75) on an nx x nx grid, SEQAIJ */
static PetscErrorCode host_assemble(PetscInt nx, double kappa, Mat *A)
{
  const PetscInt n = nx * nx;
  const double   h = 1.0 / (double)((nx - 1) * (nx - 1));
  PetscInt      *ia, *ja, nnz = 0;
  double        *va;
  PetscCall(PetscMalloc1(n + 1, &ia));
  PetscCall(PetscMalloc1(5 * n, &ja));
  PetscCall(PetscMalloc1(5 * n, &va));
  ia[0] = 0;
  for (PetscInt j = 0; j < nx; ++j)
    for (PetscInt i = 0; i < nx; ++i) {
      double d = kappa * kappa;
      if (j > 0) { ja[nnz] = i + nx * (j - 1); va[nnz++] = -h; d += h; }
      if (i > 0) { ja[nnz] = i - 1 + nx * j; va[nnz++] = -h; d += h; }
      const PetscInt dpos = nnz++;
      ja[dpos] = i + nx * j;
      if (i < nx - 1) { ja[nnz] = i + 1 + nx * j; va[nnz++] = -h; d += h; }
      if (j < nx - 1) { ja[nnz] = i + nx * (j + 1); va[nnz++] = -h; d += h; }
      va[dpos] = d;
      ia[i + nx * j + 1] = nnz;
    }
  PetscCall(MatStubCreateSeqAIJ(MPI_COMM_WORLD, n, n, ia, ja, va, A));
  PetscCall(PetscFree(ia));
  PetscCall(PetscFree(ja));
  PetscCall(PetscFree(va));
  return PETSC_SUCCESS;
}

/* a reproducible stand-in for VecSetRandom(v, NULL): uniform in (0, 1) */
static PetscErrorCode host_fill_uniform(Vec v, unsigned long long seed)
{
  PetscScalar *a;
  PetscInt     n;
  PetscCall(VecGetLocalSize(v, &n));
  PetscCall(VecGetArray(v, &a));
  for (PetscInt i = 0; i < n; ++i) {
    seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    a[i] = (double)((seed >> 11) + 1) / 9007199254740994.0;
  }
  PetscCall(VecRestoreArray(v, &a));
  return PETSC_SUCCESS;
}

static PetscErrorCode host_diffnorm(Vec x, Vec y, double *rel)
{
  const PetscScalar *a, *b;
  PetscInt           n;
  double             num = 0, den = 0;
  PetscCall(VecGetLocalSize(x, &n));
  PetscCall(VecGetArrayRead(x, &a));
  PetscCall(VecGetArrayRead(y, &b));
  for (PetscInt i = 0; i < n; ++i) {
    num += (a[i] - b[i]) * (a[i] - b[i]);
    den += b[i] * b[i];
  }
  PetscCall(VecRestoreArrayRead(x, &a));
  PetscCall(VecRestoreArrayRead(y, &b));
  *rel = sqrt(num / (den > 0 ? den : 1));
  return PETSC_SUCCESS;
}
#endif
