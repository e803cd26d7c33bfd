// host_sparse.cpp -- one-off set-up work on the host: colourings, grid transfers, Galerkin
// products and the dense factorisation of the coarsest operator.  Nothing here runs per sample.
//
// Reference behaviour being provided (the reference delegates all of it to PETSc):
//   colouring          src/mc_sor.c:383-410  (MatColoring JP / one colour)
//   Q1 interpolation   PETSc DMCreateInterpolation on a DMDA
//   A_c = P^T A P      -pc_mg_galerkin both, src/pc_gamgmc.c:345-349
//   potrf "L"          src/pc_chols.c:173-195
#include <algorithm>
#include <cmath>

#include "common.hpp"

void host_transpose(const HostCsr &a, HostCsr &t)
{
  const int64_t nnz = a.nnz();
  t.n               = a.m;
  t.m               = a.n;
  t.rowptr.assign((size_t)a.m + 1, 0);
  t.col.resize((size_t)nnz);
  t.val.resize((size_t)nnz);
  for (int64_t k = 0; k < nnz; ++k) t.rowptr[(size_t)a.col[k] + 1]++;
  for (int64_t c = 0; c < a.m; ++c) t.rowptr[c + 1] += t.rowptr[c];
  std::vector<int64_t> pos(t.rowptr.begin(), t.rowptr.end() - 1);
  for (int64_t r = 0; r < a.n; ++r)
    for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k) {
      const int64_t q = pos[a.col[k]]++;
      t.col[q]        = (int32_t)r;
      t.val[q]        = a.val[k];
    }
}

// Row-by-row sparse product with a dense accumulator; output rows sorted by column.
void host_matmul(const HostCsr &a, const HostCsr &b, HostCsr &c)
{
  c.n = a.n;
  c.m = b.m;
  c.rowptr.assign((size_t)a.n + 1, 0);
  c.col.clear();
  c.val.clear();
  c.col.reserve((size_t)a.nnz() * 2);
  c.val.reserve((size_t)a.nnz() * 2);
  std::vector<int64_t> stamp((size_t)b.m, -1);
  std::vector<double>  acc((size_t)b.m, 0.0);
  std::vector<int32_t> touched;
  for (int64_t i = 0; i < a.n; ++i) {
    touched.clear();
    for (int64_t k = a.rowptr[i]; k < a.rowptr[i + 1]; ++k) {
      const int32_t mid = a.col[k];
      const double  av  = a.val[k];
      for (int64_t l = b.rowptr[mid]; l < b.rowptr[mid + 1]; ++l) {
        const int32_t j = b.col[l];
        if (stamp[j] != i) {
          stamp[j] = i;
          acc[j]   = 0.0;
          touched.push_back(j);
        }
        acc[j] = std::fma(av, b.val[l], acc[j]);
      }
    }
    std::sort(touched.begin(), touched.end());
    for (int32_t j : touched) {
      c.col.push_back(j);
      c.val.push_back(acc[j]);
    }
    c.rowptr[i + 1] = (int64_t)c.col.size();
  }
}

void host_q1_dims(int dim, const int64_t nf[3], int64_t nc[3])
{
  for (int d = 0; d < 3; ++d) nc[d] = (d < dim && nf[d] > 1) ? (nf[d] + 1) / 2 : 1;
}

namespace {
struct Stencil1 {
  int     cnt;
  int64_t idx[2];
  double  w[2];
};
// vertex-centred linear interpolation, ratio 2: even fine node copies coarse i/2, odd fine node
// averages its existing coarse neighbours
Stencil1 q1_line(int64_t i, int64_t nc, bool active)
{
  Stencil1 s{1, {0, 0}, {1.0, 0.0}};
  if (!active || nc == 1) {
    if (active && i > 0) s.w[0] = 0.5; // nf == 2: second node sees only coarse 0
    return s;
  }
  if ((i & 1) == 0) {
    s.idx[0] = i / 2;
    return s;
  }
  s.idx[0] = (i - 1) / 2;
  s.w[0]   = 0.5;
  if ((i + 1) / 2 < nc) {
    s.cnt    = 2;
    s.idx[1] = (i + 1) / 2;
    s.w[1]   = 0.5;
  }
  return s;
}
} // namespace

void host_q1_interp(int dim, const int64_t nf[3], const int64_t nc[3], HostCsr &p)
{
  p.n = nf[0] * nf[1] * nf[2];
  p.m = nc[0] * nc[1] * nc[2];
  p.rowptr.assign((size_t)p.n + 1, 0);
  p.col.clear();
  p.val.clear();
  for (int64_t z = 0; z < nf[2]; ++z) {
    const Stencil1 sz = q1_line(z, nc[2], dim >= 3 && nf[2] > 1);
    for (int64_t j = 0; j < nf[1]; ++j) {
      const Stencil1 sy = q1_line(j, nc[1], dim >= 2 && nf[1] > 1);
      for (int64_t i = 0; i < nf[0]; ++i) {
        const Stencil1 sx = q1_line(i, nc[0], nf[0] > 1);
        for (int c = 0; c < sz.cnt; ++c)
          for (int b = 0; b < sy.cnt; ++b)
            for (int a = 0; a < sx.cnt; ++a) {
              p.col.push_back((int32_t)(sx.idx[a] + nc[0] * (sy.idx[b] + nc[1] * sz.idx[c])));
              p.val.push_back(sx.w[a] * sy.w[b] * sz.w[c]);
            }
        p.rowptr[(size_t)(i + nf[0] * (j + nf[1] * z)) + 1] = (int64_t)p.col.size();
      }
    }
  }
}

int host_coloring_greedy(const HostCsr &a, std::vector<int32_t> &color)
{
  color.assign((size_t)a.n, -1);
  std::vector<int64_t> forbidden;
  int                  ncolors = 0;
  for (int64_t r = 0; r < a.n; ++r) {
    for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k) {
      const int32_t c = a.col[k];
      if (c != r && color[c] >= 0) {
        if ((size_t)color[c] >= forbidden.size()) forbidden.resize((size_t)color[c] + 1, -1);
        forbidden[color[c]] = r;
      }
    }
    int c = 0;
    while (c < (int)forbidden.size() && forbidden[c] == r) ++c;
    color[r] = c;
    ncolors  = std::max(ncolors, c + 1);
  }
  return ncolors;
}

int host_coloring_levelset(const HostCsr &a, std::vector<int32_t> &color)
{
  color.assign((size_t)a.n, 0);
  int ncolors = a.n > 0 ? 1 : 0;
  for (int64_t r = 0; r < a.n; ++r) {
    int32_t lvl = 0;
    for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k)
      if (a.col[k] < r) lvl = std::max(lvl, color[a.col[k]] + 1);
    color[r] = lvl;
    ncolors  = std::max(ncolors, lvl + 1);
  }
  return ncolors;
}

int64_t host_coloring_violations(const HostCsr &a, const std::vector<int32_t> &color)
{
  int64_t bad = 0;
  for (int64_t r = 0; r < a.n; ++r)
    for (int64_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k)
      if (a.col[k] != r && color[a.col[k]] == color[r]) ++bad;
  return bad;
}

// column-major lower Cholesky; returns the order of the first non-positive leading minor, 0 on success
int host_potrf_lower(int64_t n, std::vector<double> &a)
{
  for (int64_t j = 0; j < n; ++j) {
    double d = a[j + j * n];
    for (int64_t k = 0; k < j; ++k) d = std::fma(-a[j + k * n], a[j + k * n], d);
    if (!(d > 0)) return (int)(j + 1);
    d            = std::sqrt(d);
    a[j + j * n] = d;
    for (int64_t i = j + 1; i < n; ++i) {
      double s = a[i + j * n];
      for (int64_t k = 0; k < j; ++k) s = std::fma(-a[i + k * n], a[j + k * n], s);
      a[i + j * n] = s / d;
    }
  }
  return 0;
}
