// stencil_op.cu -- structured-grid operators.
#include "common.hpp"

void laplace_assemble(int dim, int64_t nx, int64_t ny, int64_t nz, double kappa, HostCsr &a);

int make_laplace_op(pmg_ctx ctx, int dim, int64_t nx, int64_t ny, int64_t nz, double kappa, int64_t slab_lo, int64_t slab_hi, std::unique_ptr<LevelOp> &op)
{
  const int64_t nslow = dim == 3 ? nz : ny;
  if (slab_lo != 0 || (slab_hi != nslow && slab_hi != 0)) PMG_FAIL(PMG_ERR_SUP, "slab-partitioned Laplace operator needs the matrix-free path");
  HostCsr a;
  laplace_assemble(dim, nx, ny, nz, kappa, a);
  const int64_t dims[3] = {nx, ny, dim == 3 ? nz : 1};
  PMG_TRY(make_csr_grid_op(ctx, std::move(a), dim, dims, op));
  return op->set_coloring_auto(PMG_COLORING_PARITY);
}

int build_structured_hierarchy(pmg_ctx, LevelOp *, int, std::vector<std::unique_ptr<LevelOp>> &, std::vector<std::unique_ptr<Transfer>> &)
{
  PMG_FAIL(PMG_ERR_SUP, "matrix-free hierarchy not available");
}
