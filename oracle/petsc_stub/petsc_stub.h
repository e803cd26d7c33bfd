/* petsc_stub.h -- a container-only stand-in for the slice of the PETSc + MPI API that the reference's
 * sweep engine touches (TEST INFRASTRUCTURE ONLY).
 *
 * Purpose: /root/reference cannot be linked here (PETSc, MPI and MKL are absent).  This stub
 * declares exactly the types, macros and functions that /root/reference/src/{mc_sor.c, pc_mcgibbs.c, parmgmc.c}
 * use, so that those files compile UNMODIFIED, from where they lie, into oracle/_ref/libparmgmc_ref.so
 * (oracle/Makefile target `ref`).  tests/test_oracle_ref.py then runs the reference's own loops
 * (MCSORApply_SEQAIJ, MCSORApply_MPIAIJ, PrepareRHS_Default, VecSetRandomStandardNormal, the symmetric-sweep
 * handling of PCApplyRichardson_MulticolorGibbs) next to the oracle restatement on the same inputs.
 *
 * What is real and what is emulated:
 *   real      every line of the three reference files above
 *   emulated  Vec/Mat/IS containers (plain arrays), the MPIAIJ layout (diagonal block with local columns,
 *             off-diagonal block with compressed columns + colmap), VecScatter (a gather from one shared
 *             array), ranks (threads of one process with a barrier inside VecScatterBegin/End),
 *             MatColoringApply (returns the colouring the driver injected instead of PETSc's randomised JP),
 *             PetscRandom (rander48), the options database (a small key/value table)
 * Nothing of PETSc's source is copied: these are the public API signatures re-declared with toy bodies.
 */
#ifndef PETSC_STUB_H
#define PETSC_STUB_H
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PETSC_EXTERN extern
#define PETSC_INTERN extern
#define PETSC_VERSION_LT(a, b, c) 0
#define PETSC_VERSION_GE(a, b, c) 1

typedef int       PetscInt;
typedef long long PetscInt64;
typedef double    PetscReal;
typedef double    PetscScalar;
typedef int       PetscMPIInt;
typedef int       PetscErrorCode;
typedef int       PetscClassId;
typedef int       PetscLogEvent;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef enum { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER } PetscCopyMode;
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { SCATTER_FORWARD = 0, SCATTER_REVERSE = 1 } ScatterMode;
typedef enum { MAT_DO_NOT_COPY_VALUES, MAT_COPY_VALUES, MAT_SHARE_NONZERO_PATTERN } MatDuplicateOption;
typedef enum { MAT_INITIAL_MATRIX, MAT_REUSE_MATRIX } MatReuse;
typedef enum { PETSC_MEMTYPE_HOST = 0 } PetscMemType;
typedef enum {
  SOR_FORWARD_SWEEP         = 1,
  SOR_BACKWARD_SWEEP        = 2,
  SOR_SYMMETRIC_SWEEP       = 3,
  SOR_LOCAL_FORWARD_SWEEP   = 4,
  SOR_LOCAL_BACKWARD_SWEEP  = 8,
  SOR_LOCAL_SYMMETRIC_SWEEP = 12,
  SOR_ZERO_INITIAL_GUESS    = 16
} MatSORType;
typedef enum { IS_COLORING_GLOBAL, IS_COLORING_LOCAL } ISColoringType;
typedef enum { PCRICHARDSON_NOT_SET = 0, PCRICHARDSON_CONVERGED_RTOL = 2, PCRICHARDSON_CONVERGED_ATOL = 3, PCRICHARDSON_CONVERGED_ITS = 4 } PCRichardsonConvergedReason;
typedef unsigned short ISColoringValue;
typedef const char    *MatType;
typedef const char    *MatColoringType;

#define PETSC_SUCCESS 0
#define PETSC_ERR_SUP 56
#define PETSC_ERR_PLIB 77
#define PETSC_ERR_LIB 76
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ORDER 58
#define PetscInt_FMT "d"
#define PETSC_PI 3.1415926535897932384626433832795029
#define PetscSqrtReal(a) sqrt(a)
#define PetscLogReal(a) log(a)
#define PetscCosReal(a) cos(a)
#define PetscSinReal(a) sin(a)
#define PetscAbsReal(a) fabs(a)

/* ---- MPI ---- */
typedef int MPI_Comm;
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_COMM_SELF MPI_COMM_SELF
#define MPI_SUCCESS 0
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
#define MPI_BYTE 1
int MPI_Bcast(void *buf, int count, int datatype, int root, MPI_Comm comm); /* one rank per process in the stub: a no-op */

/* ---- error handling / memory ---- */
#define PetscFunctionBegin
#define PetscFunctionBeginUser
#define PetscFunctionReturn(x) return (x)
#define PetscCall(...)                    \
  do {                                    \
    PetscErrorCode ierr_ = (__VA_ARGS__); \
    if (ierr_) return ierr_;              \
  } while (0)
#define PetscCallMPI(...) PetscCall(__VA_ARGS__)
PetscErrorCode PetscStubError(int code, const char *file, int line, const char *fmt, ...);
#define SETERRQ(comm, code, ...) return PetscStubError((code), __FILE__, __LINE__, __VA_ARGS__)
#define PetscCheck(cond, comm, code, ...)                                      \
  do {                                                                         \
    if (!(cond)) return PetscStubError((code), __FILE__, __LINE__, __VA_ARGS__); \
  } while (0)
#define PetscAssert(cond, comm, code, ...) PetscCheck(cond, comm, code, __VA_ARGS__)
#define PetscMalloc1(n, p) ((*(p) = malloc(sizeof(**(p)) * (size_t)((n) > 0 ? (n) : 1))) ? PETSC_SUCCESS : 55)
#define PetscCalloc1(n, p) ((*(p) = calloc((size_t)((n) > 0 ? (n) : 1), sizeof(**(p)))) ? PETSC_SUCCESS : 55)
#define PetscNew(p) PetscCalloc1(1, p)
#define PetscFree(p) (free(p), (p) = NULL, PETSC_SUCCESS)

/* ---- objects ---- */
typedef struct _p_PetscObject *PetscObject;
typedef struct _p_Vec         *Vec;
typedef struct _p_Mat         *Mat;
typedef struct _p_IS          *IS;
typedef struct _n_ISColoring  *ISColoring;
typedef struct _p_VecScatter  *VecScatter;
typedef struct _p_MatColoring *MatColoring;
typedef struct _p_KSP         *KSP;
typedef struct _p_PC          *PC;
typedef struct _p_PetscRandom *PetscRandom;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_PetscOptions *PetscOptions;
typedef struct _p_PetscOptionItems *PetscOptionItems;

struct _p_PetscObject {
  MPI_Comm comm;
  int      refct;
  const char *prefix;
  struct { const char *name; void (*f)(void); } composed[4]; /* PetscObjectComposeFunction table ("PCSetSampleCallback_C", "PCMGGetLevels_C") */
};
void (*PetscStubQueryFunction(PetscObject o, const char *name))(void);
MPI_Comm       PetscObjectComm(PetscObject o);
PetscErrorCode PetscObjectGetComm(PetscObject o, MPI_Comm *comm);
PetscErrorCode PetscObjectReference(PetscObject o);
PetscErrorCode PetscObjectComposeFunction_Stub(PetscObject o, const char *name, void (*f)(void));
#define PetscObjectComposeFunction(o, name, f) PetscObjectComposeFunction_Stub((o), (name), (void (*)(void))(f))
#define PetscUseMethod(obj, name, proto, args)                                                 \
  do {                                                                                         \
    PetscErrorCode(*f_) proto = (PetscErrorCode(*) proto)PetscStubQueryFunction((obj), (name)); \
    PetscCheck(f_, PETSC_COMM_SELF, PETSC_ERR_SUP, "no method %s", name);                      \
    PetscCall((*f_)args);                                                                      \
  } while (0)

/* logging: no-ops */
#define PetscLogEventBegin(e, a, b, c, d) PETSC_SUCCESS
#define PetscLogEventEnd(e, a, b, c, d) PETSC_SUCCESS
PetscErrorCode PetscClassIdRegister(const char *name, PetscClassId *id);
PetscErrorCode PetscLogEventRegister(const char *name, PetscClassId id, PetscLogEvent *e);

/* options database: key/value table filled by the driver (PetscStubOptionsSet) */
PetscErrorCode PetscStubOptionsSet(const char *key, const char *value);
PetscErrorCode PetscStubOptionsClear(void);
PetscErrorCode PetscOptionsGetString(PetscOptions o, const char *pre, const char *name, char *buf, size_t len, PetscBool *set);
PetscErrorCode PetscOptionsGetReal(PetscOptions o, const char *pre, const char *name, PetscReal *v, PetscBool *set);
#define PetscOptionsHeadBegin(obj, title) (void)(obj)
#define PetscOptionsHeadEnd() (void)0
PetscErrorCode PetscOptionsRangeReal(const char *opt, const char *text, const char *man, PetscReal cur, PetscReal *v, PetscBool *set, PetscReal lo, PetscReal hi);
PetscErrorCode PetscOptionsBool(const char *opt, const char *text, const char *man, PetscBool cur, PetscBool *v, PetscBool *set);
PetscErrorCode PetscOptionsString(const char *opt, const char *text, const char *man, const char *cur, char *v, size_t len, PetscBool *set);

/* ---- Vec ---- */
struct _p_Vec {
  struct _p_PetscObject hdr;
  PetscInt              n, N;   /* local, global size */
  PetscInt              rstart; /* global index of local entry 0 */
  double               *a;      /* local array; for a distributed vector a - rstart addresses the whole vector */
  int                   own;
};
PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec *v);
PetscErrorCode VecCreateMPIWithArray(MPI_Comm comm, PetscInt bs, PetscInt n, PetscInt N, const PetscScalar *array, Vec *v);
PetscErrorCode VecStubCreate(MPI_Comm comm, PetscInt n, PetscInt N, PetscInt rstart, double *array, Vec *v); /* driver helper */
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecDuplicate(Vec v, Vec *w);
PetscErrorCode VecGetArray(Vec v, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a);
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a);
PetscErrorCode VecGetSize(Vec v, PetscInt *N);
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n);
PetscErrorCode VecReciprocal(Vec v);
PetscErrorCode VecScale(Vec v, PetscScalar s);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecZeroEntries(Vec v);
PetscErrorCode VecSqrtAbs(Vec v);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);

/* ---- IS / ISColoring ---- */
struct _p_IS {
  struct _p_PetscObject hdr;
  PetscInt              n;
  PetscInt             *idx;
};
PetscErrorCode ISCreateGeneral(MPI_Comm comm, PetscInt n, const PetscInt idx[], PetscCopyMode mode, IS *is);
PetscErrorCode ISCreateStride(MPI_Comm comm, PetscInt n, PetscInt first, PetscInt step, IS *is);
PetscErrorCode ISDestroy(IS *is);
PetscErrorCode ISGetLocalSize(IS is, PetscInt *n);
PetscErrorCode ISGetIndices(IS is, const PetscInt **idx);
PetscErrorCode ISRestoreIndices(IS is, const PetscInt **idx);
struct _n_ISColoring {
  MPI_Comm         comm;
  PetscInt         ncolors, n;
  ISColoringValue *colors;
  IS              *is;
};
PetscErrorCode ISColoringCreate(MPI_Comm comm, PetscInt ncolors, PetscInt n, const ISColoringValue colors[], PetscCopyMode mode, ISColoring *isc);
PetscErrorCode ISColoringSetType(ISColoring isc, ISColoringType t);
PetscErrorCode ISColoringGetIS(ISColoring isc, PetscCopyMode mode, PetscInt *n, IS *iss[]);
PetscErrorCode ISColoringRestoreIS(ISColoring isc, PetscCopyMode mode, IS *iss[]);
PetscErrorCode ISColoringDestroy(ISColoring *isc);

/* ---- Mat ---- */
#define MATSEQAIJ "seqaij"
#define MATMPIAIJ "mpiaij"
#define MATLRC "lrc"
#define MATCOLORINGJP "jp"
struct _p_Mat {
  struct _p_PetscObject hdr;
  const char           *type;
  PetscInt              m, n, M, N, rstart; /* local rows/cols, global rows/cols, first owned row */
  /* seqaij */
  PetscInt *i, *j;
  double   *a;
  /* mpiaij */
  Mat       Ad, Ao;
  PetscInt *colmap; /* compressed off-diagonal column -> global column (PETSc's garray) */
  /* seqdense: column-major values, leading dimension m */
  double *d;
  /* lrc: A + U diag(c) U^T (MatCreateLRC); borrowed references */
  Mat lrc_A, lrc_U;
  Vec lrc_c;
  /* factor objects (MatGetFactor): dense LU with partial pivoting / dense Cholesky of the matrix given to the numeric phase */
  double *fac;
  int    *piv;
  int     fac_kind; /* 0 none, 1 LU, 2 Cholesky (lower) */
  /* colouring the driver wants MatColoringApply to return for this matrix (local rows), and its global colour count */
  const ISColoringValue *inject_colors;
  PetscInt               inject_ncolors;
};
PetscErrorCode MatStubCreateSeqAIJ(MPI_Comm comm, PetscInt m, PetscInt n, const PetscInt *i, const PetscInt *j, const double *a, Mat *A); /* copies */
PetscErrorCode MatStubCreateMPIAIJ(PetscInt m, PetscInt M, PetscInt rstart, Mat Ad, Mat Ao, const PetscInt *colmap, PetscInt ncolmap, Mat *A);
PetscErrorCode MatStubInjectColoring(Mat A, PetscInt ncolors, const ISColoringValue *colors);
PetscErrorCode MatDestroy(Mat *A);
PetscErrorCode MatGetType(Mat A, MatType *t);
PetscErrorCode MatGetSize(Mat A, PetscInt *M, PetscInt *N);
PetscErrorCode MatGetLocalSize(Mat A, PetscInt *m, PetscInt *n);
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt *lo, PetscInt *hi);
PetscErrorCode MatMPIAIJGetSeqAIJ(Mat A, Mat *Ad, Mat *Ao, const PetscInt **colmap);
PetscErrorCode MatSeqAIJGetCSRAndMemType(Mat A, const PetscInt **i, const PetscInt **j, PetscScalar **a, PetscMemType *mt);
PetscErrorCode MatGetDiagonal(Mat A, Vec d);
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z);
/* MATLRC / dense / KSP pieces of the low-rank path and of the estimators (petsc_stub_dense.c): small dense column-major
 * matrices, an exact dense solve behind KSP, PETSc's MatSOR_SeqAIJ semantics restated */
PetscErrorCode MatLRCGetMats(Mat A, Mat *base, Mat *U, Vec *c, Mat *V);
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption o, Mat *B);
PetscErrorCode MatDenseGetColumnVecRead(Mat A, PetscInt c, Vec *v);
PetscErrorCode MatDenseRestoreColumnVecRead(Mat A, PetscInt c, Vec *v);
PetscErrorCode MatDenseGetColumnVecWrite(Mat A, PetscInt c, Vec *v);
PetscErrorCode MatDenseRestoreColumnVecWrite(Mat A, PetscInt c, Vec *v);
PetscErrorCode MatTransposeMatMult(Mat A, Mat B, MatReuse r, PetscReal fill, Mat *C);
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse r, PetscReal fill, Mat *C);
PetscErrorCode MatDiagonalSet(Mat A, Vec d, InsertMode m);
PetscErrorCode MatShift(Mat A, PetscScalar s);
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P);
PetscErrorCode KSPMatSolve(KSP ksp, Mat B, Mat X);
PetscErrorCode KSPDestroy(KSP *ksp);

#define MATAIJ "aij"
#define MATSBAIJ "sbaij"
#define MATSEQDENSE "seqdense"
#define MATDENSE "dense"
#define PETSC_DECIDE (-1)
#define PETSC_DEFAULT (-2)
#define PETSC_ERR_SUP_SYS 57
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_MAT_CH_ZRPVT 81
typedef enum { DIFFERENT_NONZERO_PATTERN, SUBSET_NONZERO_PATTERN, SAME_NONZERO_PATTERN, UNKNOWN_NONZERO_PATTERN } MatStructure;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { MAT_FACTOR_NONE, MAT_FACTOR_LU, MAT_FACTOR_CHOLESKY } MatFactorType;
typedef enum { MAT_SYMMETRIC = 1, MAT_SPD = 2 } MatOption;
typedef const char *MatSolverType;
typedef const char *MatOrderingType;
#define MATSOLVERPETSC "petsc"
#define MATORDERINGNATURAL "natural"
#define MATORDERINGMETISND "metisnd"
#define MATORDERINGEXTERNAL "external"
typedef struct { PetscReal fill, dtcol; } MatFactorInfo;
typedef struct { PetscReal nz_used, nz_allocated, memory; } MatInfo;
typedef enum { MAT_LOCAL = 1, MAT_GLOBAL_MAX = 2, MAT_GLOBAL_SUM = 3 } MatInfoType;
PetscErrorCode MatCreateDense(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscScalar *data, Mat *A);
PetscErrorCode MatCreateSeqDense(MPI_Comm comm, PetscInt m, PetscInt n, PetscScalar *data, Mat *A);
PetscErrorCode MatCreateLRC(Mat A, Mat U, Vec c, Mat V, Mat *N);
PetscErrorCode MatDenseGetArray(Mat A, PetscScalar **a);
PetscErrorCode MatDenseRestoreArray(Mat A, PetscScalar **a);
PetscErrorCode MatDenseGetArrayRead(Mat A, const PetscScalar **a);
PetscErrorCode MatDenseRestoreArrayRead(Mat A, const PetscScalar **a);
PetscErrorCode MatZeroEntries(Mat A);
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure s);
PetscErrorCode MatNorm(Mat A, NormType t, PetscReal *nrm);
PetscErrorCode MatConvert(Mat A, MatType t, MatReuse r, Mat *B);
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r);
PetscErrorCode MatMatTransposeMult(Mat A, Mat B, MatReuse r, PetscReal fill, Mat *C);
PetscErrorCode MatSetOption(Mat A, MatOption o, PetscBool v);
PetscErrorCode MatFactorInfoInitialize(MatFactorInfo *info);
PetscErrorCode MatGetFactor(Mat A, MatSolverType st, MatFactorType ft, Mat *F);
PetscErrorCode MatGetOrdering(Mat A, MatOrderingType t, IS *r, IS *c);
PetscErrorCode MatLUFactorSymbolic(Mat F, Mat A, IS r, IS c, const MatFactorInfo *info);
PetscErrorCode MatLUFactorNumeric(Mat F, Mat A, const MatFactorInfo *info);
PetscErrorCode MatCholeskyFactorSymbolic(Mat F, Mat A, IS perm, const MatFactorInfo *info);
PetscErrorCode MatCholeskyFactorNumeric(Mat F, Mat A, const MatFactorInfo *info);
PetscErrorCode MatMatSolve(Mat F, Mat B, Mat X);
PetscErrorCode MatForwardSolve(Mat F, Vec b, Vec x);
PetscErrorCode MatBackwardSolve(Mat F, Vec b, Vec x);
PetscErrorCode MatGetInfo(Mat A, MatInfoType t, MatInfo *info);
/* PETSc's MatSOR on a SeqAIJ matrix: its x (lits) directional sweeps, no SOR_ZERO_INITIAL_GUESS */
PetscErrorCode MatSOR(Mat A, Vec b, PetscReal omega, MatSORType flag, PetscReal shift, PetscInt its, PetscInt lits, Vec x);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *lo, PetscInt *hi);
PetscErrorCode VecGetLocalVector(Vec v, Vec w);
PetscErrorCode VecRestoreLocalVector(Vec v, Vec w);
PetscErrorCode VecGetLocalVectorRead(Vec v, Vec w);
PetscErrorCode VecRestoreLocalVectorRead(Vec v, Vec w);
PetscErrorCode PetscStrcmp(const char *a, const char *b, PetscBool *flg);
#define PetscArraycpy(dst, src, n) (memcpy((dst), (src), sizeof(*(dst)) * (size_t)(n)), PETSC_SUCCESS)
#define PetscRealPart(a) creal(a)
/* BLAS / LAPACK (reference implementations, column-major, petsc_stub_dense.c) */
typedef int PetscBLASInt;
#define PetscBLASInt_FMT "d"
#define PetscBLASIntCast(a, b) (*(b) = (PetscBLASInt)(a), PETSC_SUCCESS)
#define PetscCallBLAS(name, call) do { call; } while (0)
typedef enum { PETSC_FP_TRAP_OFF = 0, PETSC_FP_TRAP_ON = 1 } PetscFPTrap;
#define PetscFPTrapPush(t) PETSC_SUCCESS
#define PetscFPTrapPop() PETSC_SUCCESS
void LAPACKpotrf_(const char *uplo, const PetscBLASInt *n, PetscScalar *a, const PetscBLASInt *lda, PetscBLASInt *info);
void BLAStrsv_(const char *uplo, const char *trans, const char *diag, const PetscBLASInt *n, const PetscScalar *a, const PetscBLASInt *lda, PetscScalar *x, const PetscBLASInt *incx);
PetscErrorCode PetscOptionsInt(const char *opt, const char *text, const char *man, PetscInt cur, PetscInt *v, PetscBool *set);

/* ---- MatColoring ---- */
PetscErrorCode MatColoringCreate(Mat A, MatColoring *mc);
PetscErrorCode MatColoringSetDistance(MatColoring mc, PetscInt d);
PetscErrorCode MatColoringSetType(MatColoring mc, MatColoringType t);
PetscErrorCode MatColoringApply(MatColoring mc, ISColoring *isc);
PetscErrorCode MatColoringDestroy(MatColoring *mc);

/* ---- VecScatter: global vector -> sequential ghost vector ---- */
PetscErrorCode VecScatterCreate(Vec x, IS ix, Vec y, IS iy, VecScatter *sct);
PetscErrorCode VecScatterBegin(VecScatter sct, Vec x, Vec y, InsertMode im, ScatterMode sm);
PetscErrorCode VecScatterEnd(VecScatter sct, Vec x, Vec y, InsertMode im, ScatterMode sm);
PetscErrorCode VecScatterDestroy(VecScatter *sct);

/* ---- PetscRandom (rander48) ---- */
PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r);
PetscErrorCode PetscRandomSetFromOptions(PetscRandom r);
PetscErrorCode PetscRandomSetSeed(PetscRandom r, PetscInt64 seed);
PetscErrorCode PetscRandomGetSeed(PetscRandom r, PetscInt64 *seed);
PetscErrorCode PetscRandomSeed(PetscRandom r);
PetscErrorCode PetscRandomGetValueReal(PetscRandom r, PetscReal *v);
PetscErrorCode PetscRandomDestroy(PetscRandom *r);

/* ---- PC (petsc/private/pcimpl.h) ---- */
typedef struct _PCOps *PCOps;
struct _PCOps {
  PetscErrorCode (*setup)(PC);
  PetscErrorCode (*apply)(PC, Vec, Vec);
  PetscErrorCode (*applyrichardson)(PC, Vec, Vec, Vec, PetscReal, PetscReal, PetscReal, PetscInt, PetscBool, PetscInt *, PCRichardsonConvergedReason *);
  PetscErrorCode (*setfromoptions)(PC, PetscOptionItems);
  PetscErrorCode (*reset)(PC);
  PetscErrorCode (*destroy)(PC);
  PetscErrorCode (*view)(PC, PetscViewer);
  PetscErrorCode (*presolve)(PC, KSP, Vec, Vec);
  PetscErrorCode (*postsolve)(PC, KSP, Vec, Vec);
};
struct _p_PC {
  struct _p_PetscObject hdr;
  struct _PCOps         ops[1];
  void                 *data;
  Mat                   mat, pmat;
  PetscBool             setupcalled;
};
PetscErrorCode PCRegister(const char *name, PetscErrorCode (*create)(PC));
PetscErrorCode PCStubCreate(const char *type, Mat pmat, PC *pc); /* PCCreate + PCSetType + PCSetOperators */
typedef const char *PCType;
PetscErrorCode PCCreate(MPI_Comm comm, PC *pc);
PetscErrorCode PCSetType(PC pc, PCType type);
PetscErrorCode PCSetOperators(PC pc, Mat A, Mat P);
PetscErrorCode PCSetUp(PC pc);
PetscErrorCode PCDestroy(PC *pc);
PetscErrorCode PCStubDestroy(PC *pc);
/* what src/woodbury.c needs of a PC used as a member object: the type-independent entry points, the options prefix (kept, not
 * used for look-ups: the stub's options table is flat) and a built-in exact solver type "cholesky" / "lu" (dense LU of pmat) */
PetscErrorCode PCApply(PC pc, Vec x, Vec y);
PetscErrorCode PCApplyRichardson(PC pc, Vec b, Vec y, Vec w, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt its, PetscBool guesszero, PetscInt *outits, PCRichardsonConvergedReason *reason);
PetscErrorCode PCSetFromOptions(PC pc);
PetscErrorCode PCReset(PC pc);
PetscErrorCode PCSetOptionsPrefix(PC pc, const char *prefix);
PetscErrorCode PCAppendOptionsPrefix(PC pc, const char *prefix);
PetscErrorCode PCGetOptionsPrefix(PC pc, const char **prefix);
#define PetscObjectIncrementTabLevel(obj, parent, n) 0
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char *fmt, ...);

/* ---- DMDA: just enough for src/problems.c (a one-rank 2D grid; the Mat is a dense mx my x mx my matrix, row = j mx + i) ---- */
typedef struct _p_DM *DM;
struct _p_DM {
  PetscInt mx, my;
};
typedef struct {
  PetscInt k, j, i, c;
} MatStencil;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
PetscErrorCode DMStubCreate2d(PetscInt mx, PetscInt my, DM *dm);
PetscErrorCode DMDestroy(DM *dm);
PetscErrorCode DMDAGetInfo(DM dm, PetscInt *dim, PetscInt *M, PetscInt *N, PetscInt *P, PetscInt *m, PetscInt *n, PetscInt *p, PetscInt *dof, PetscInt *s, void *bx, void *by, void *bz, void *st);
PetscErrorCode DMDAGetCorners(DM dm, PetscInt *xs, PetscInt *ys, PetscInt *zs, PetscInt *xm, PetscInt *ym, PetscInt *zm);
PetscErrorCode MatSetValuesStencil(Mat A, PetscInt m, const MatStencil *rows, PetscInt n, const MatStencil *cols, const PetscScalar *v, InsertMode mode);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);

/* ---- rank emulation: threads of one process ---- */
void PetscStubWorldBegin(int nranks);          /* called once before the rank threads start */
void PetscStubWorldEnd(void);
void PetscStubSetRank(int rank);               /* called by each rank thread */
void PetscStubBarrier(void);
#endif
