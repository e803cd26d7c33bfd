// laplace.cu -- the synthetic shifted-Laplacian GMRF operator of the reference's generator.
//
// Reference: MatAssembleShiftedLaplaceFD, src/problems.c:14-75 (2D 5-point; hinv2 = 1/(mx-1)^2 is really h^2;
// off-diagonals -hinv2 to existing neighbours; diag = kappa^2 + one hinv2 per existing neighbour, added one by
// one).  dim = 3 is the 7-point extension.
#include "common.hpp"

void laplace_assemble(int dim, int64_t nx, int64_t ny, int64_t nz, double kappa, HostCsr &a)
{
  const double h = 1.0 / (double)((nx - 1) * (nx - 1));
  if (dim == 2) nz = 1;
  a.n = a.m = nx * ny * nz;
  a.rowptr.assign((size_t)a.n + 1, 0);
  a.col.clear();
  a.val.clear();
  a.col.reserve((size_t)a.n * (2 * dim + 1));
  a.val.reserve((size_t)a.n * (2 * dim + 1));
  for (int64_t k = 0; k < nz; ++k)
    for (int64_t j = 0; j < ny; ++j)
      for (int64_t i = 0; i < nx; ++i) {
        const int64_t r    = i + nx * (j + ny * k);
        const bool    nb[6] = {dim == 3 && k > 0, j > 0, i > 0, i < nx - 1, j < ny - 1, dim == 3 && k < nz - 1};
        const int64_t off[6] = {-nx * ny, -nx, -1, 1, nx, nx * ny};
        double        diag = kappa * kappa;
        for (int q = 0; q < 6; ++q)
          if (nb[q]) diag += h;
        for (int q = 0; q < 3; ++q)
          if (nb[q]) {
            a.col.push_back((int32_t)(r + off[q]));
            a.val.push_back(-h);
          }
        a.col.push_back((int32_t)r);
        a.val.push_back(diag);
        for (int q = 3; q < 6; ++q)
          if (nb[q]) {
            a.col.push_back((int32_t)(r + off[q]));
            a.val.push_back(-h);
          }
        a.rowptr[(size_t)r + 1] = (int64_t)a.col.size();
      }
}
