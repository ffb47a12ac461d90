// Thread bodies for K1 (factor assignment), K4 (factored energy) and K5
// (regularisation).  One thread = one batch element of one cluster / sepset.
#pragma once
#include "pgbp_internal.h"
#include "pgbp_shapes.h"

namespace pgbp {

#ifndef PGBP_LOG2PI
#define PGBP_LOG2PI 1.8378770664093454835606594728112
#define PGBP_EPS 2.220446049250313e-16
#endif

// Inverse and log-determinant of a small SPD matrix (column-major p x p, upper
// triangle read) through its upper Cholesky factor, as PDMats/`inv` do for the
// rate matrices R_c and the hybrid variances.  `w` is overwritten (U, then
// V = U^-1); `out` receives the full symmetric inverse.  Returns 0, or the
// 1-based failing pivot.
PGBP_HD int spd_inverse_logdet(double* w, double* out, int p, double* logdet) {
#define W_(r, c) w[(c)*p + (r)]
  double ld = 0.0;
  for (int k = 0; k < p; k++) {
    const double d = W_(k, k);
    if (!(d > 0.0)) return k + 1;
    ld += log(d);
    const double r = sqrt(d);
    W_(k, k) = r;
    for (int c = k + 1; c < p; c++) W_(k, c) /= r;
    for (int c = k + 1; c < p; c++)
      for (int rr = k + 1; rr <= c; rr++) W_(rr, c) -= W_(k, rr) * W_(k, c);
  }
  // V = U^-1, in place: columns descending, rows descending
  for (int c = p - 1; c >= 0; c--) {
    W_(c, c) = 1.0 / W_(c, c);
    for (int r = c - 1; r >= 0; r--) {
      double s = W_(r, c) * W_(c, c);  // U(r,c) V(c,c)
      for (int k = r + 1; k < c; k++) s += W_(r, k) * W_(k, c);  // U(r,k) V(k,c), r<k<c
      W_(r, c) = -s / W_(r, r);
    }
  }
  // out = V V'
  for (int j = 0; j < p; j++)
    for (int i = 0; i <= j; i++) {
      double s = 0.0;
      for (int k = j; k < p; k++) s += W_(i, k) * W_(j, k);
      out[j * p + i] = s;
      out[i * p + j] = s;
    }
#undef W_
  *logdet = ld;
  return 0;
}

struct FamDev {
  const int32_t* node_cluster;
  const int32_t* mem_off;
  const int32_t* mem_pos;
  const double* mem_length;
  const double* mem_gamma;
  const int32_t* mem_color;
  const int32_t* node_datarow;
  const int32_t* clu_off;
  const int32_t* clu_node;
  const int64_t* cl_jslot;  // per cluster
  const int64_t* cl_hslot;
  const int64_t* cl_gslot;
  const int32_t* cl_dim;
  int32_t p, ncolors, root_fixed;
};

// rows of the per-parameter-set table `theta` (SoA over parameter sets)
struct ThetaRows {
  int p, nc;
  PGBP_HD int R(int c) const { return c * p * p; }
  PGBP_HD int P(int c) const { return nc * p * p + c * p * p; }
  PGBP_HD int g0(int c) const { return 2 * nc * p * p + c; }
  PGBP_HD int mu() const { return 2 * nc * p * p + nc; }
  PGBP_HD int rootP() const { return mu() + p; }
  PGBP_HD int rooth() const { return rootP() + p * p; }
  PGBP_HD int rootg() const { return rooth() + p; }
  PGBP_HD int kind() const { return rootg() + 1; }  // 0 fixed, 1 proper, 2 improper, <0 invalid (-pivot)
  PGBP_HD int nrows() const { return kind() + 1; }
};

}  // namespace pgbp
