/*
 * CPU oracle, C twin of oracle/*.py -- TEST INFRASTRUCTURE ONLY (see
 * oracle/__init__.py): used by tests/ as a fast checker at full batch sizes and
 * by bench.py as the timed CPU baseline ("port": the reference is Julia and
 * cannot run in this image).  Never linked or loaded by the product.
 *
 * It restates the reference's arithmetic in the reference's own formulation --
 * dense column-major m x m precision matrices, one ClusterGraphBelief per
 * replicate, messages strictly in the reference's sequential order -- which is
 * deliberately NOT how the CUDA path computes (packed storage, permuted
 * right-looking partial Cholesky, level-parallel steps):
 *   marginalize        src/beliefupdates.jl:55-83   (PDMat(Symmetric(J_I)) -> U upper,
 *                      X_invA_Xt = (J_KI/U)(J_KI/U)', J_I \ h_I, logdet)
 *   divide! / mult!    src/beliefupdates.jl:579-587, 483-488
 *   propagate_belief!  src/beliefupdates.jl:634-665
 *   residual flags     src/beliefs.jl:994-1003
 *   traversals         src/calibration.jl:111-161 ; calibrate! :72-84
 *   integratebelief    src/beliefupdates.jl:187-200
 *   free_energy        src/score.jl:162-182
 *   BM factors + evidence  src/beliefs.jl:786-861, src/evomodels/homogeneousbrownianmotion.jl:222-351,
 *                      src/evomodels/heterogeneousmodels.jl:128-150, src/evomodels/evomodels.jl:377-396,
 *                      src/beliefupdates.jl:210-231
 * Replicates are independent: OpenMP `parallel for` over them (the
 * "Threads.@threads over replicates" baseline of BASELINE.json's north_star).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Arithmetic type.  Default: IEEE binary64, the reference's Float64.  -DPGBPO_QUAD builds the SAME code in
 * IEEE binary128 (113-bit significand, libquadmath): the ">= 100-bit" adjudicator used by tests/ and
 * tools/adjudicate_c3.py to decide which double-precision restatement is closer to the exact answer of an
 * ill-conditioned configuration.  Inputs and outputs stay binary64 at the boundary; constants that the
 * reference defines in Float64 (eps(Float64), the 1e-5 calibration tolerance) keep their binary64 values. */
#ifdef PGBPO_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_SQRT sqrtq
#define R_LOG logq
#define R_FABS fabsq
#define R_ISINF isinfq
#define LOG2PI 1.8378770664093454835606594728112352797227949472755668Q
#else
typedef double real;
#define R_SQRT sqrt
#define R_LOG log
#define R_FABS fabs
#define R_ISINF isinf
#define LOG2PI 1.8378770664093454835606594728112
#endif
#define EPS 2.220446049250313e-16
#define MAXM 64

typedef struct {
  int32_t nclusters, nsepsets, ntraits;
  const int32_t* dim;      /* [nb] */
  const int64_t* off;      /* [nb+1]: belief b = J (m*m col-major) | h (m) | g (1) at state + off[b] */
  const int32_t* sep_a;    /* [ns] cluster indices */
  const int32_t* sep_b;
  const int32_t* up_off;   /* [2*ns+1] */
  const int32_t* up;       /* scopeindex(sepset, cluster_a) then (sepset, cluster_b) */
  const int64_t* roff;     /* [2*ns+1]: residual d = dJ (s*s) | dh (s) at resid + roff[d] */
} og_graph;

typedef struct {
  int32_t nnodes, ntips, root_fixed;
  const int32_t* node_cluster;
  const int32_t* mem_off;
  const int32_t* mem_pos;
  const double* mem_length;
  const double* mem_gamma;
  const int32_t* mem_color;
  const int32_t* node_datarow;
} og_families;

int pgbpo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* upper Cholesky, unblocked LAPACK dpotf2('U') order; returns 0 or 1-based info */
static int chol_upper(real* A, int n, int lda) {
  for (int j = 0; j < n; j++) {
    real ajj = A[j + j * lda];
    for (int k = 0; k < j; k++) ajj -= A[k + j * lda] * A[k + j * lda];
    if (!(ajj > 0.0)) return j + 1;
    ajj = R_SQRT(ajj);
    A[j + j * lda] = ajj;
    for (int c = j + 1; c < n; c++) {
      real s = A[j + c * lda];
      for (int k = 0; k < j; k++) s -= A[k + j * lda] * A[k + c * lda];
      A[j + c * lda] = s / ajj;
    }
  }
  return 0;
}
static void solve_ut(const real* U, int n, int lda, real* x) { /* U' x = b */
  for (int r = 0; r < n; r++) {
    real s = x[r];
    for (int k = 0; k < r; k++) s -= U[k + r * lda] * x[k];
    x[r] = s / U[r + r * lda];
  }
}
static void solve_u(const real* U, int n, int lda, real* x) { /* U x = b */
  for (int r = n - 1; r >= 0; r--) {
    real s = x[r];
    for (int k = r + 1; k < n; k++) s -= U[r + k * lda] * x[k];
    x[r] = s / U[r + r * lda];
  }
}

/* marginalize: message (mh, mJ[s*s], mg) of belief (h,J,g) of dimension m keeping `keep` (ascending) */
static int marginalize(const real* J, const real* h, real g, int m, const int32_t* keep, int s,
                       real* mJ, real* mh, real* mg) {
  int integ[MAXM], ni = 0, kk = 0;
  for (int v = 0; v < m; v++) {
    if (kk < s && keep[kk] == v) kk++;
    else integ[ni++] = v;
  }
  for (int c = 0; c < s; c++) {
    mh[c] = h[keep[c]];
    for (int r = 0; r < s; r++) mJ[r + c * s] = J[keep[r] + keep[c] * m];
  }
  *mg = g;
  if (ni == 0) return 0;
  real Ji[MAXM * MAXM], hi[MAXM], mui[MAXM], Z[MAXM];
  int allzero = 1;
  for (int c = 0; c < ni; c++) {
    hi[c] = h[integ[c]];
    if (!(R_FABS(hi[c]) <= EPS)) allzero = 0;
    for (int r = 0; r < ni; r++) {
      Ji[r + c * ni] = J[integ[r] + integ[c] * m];
      if (!(R_FABS(Ji[r + c * ni]) <= EPS)) allzero = 0;
    }
    for (int r = 0; r < s; r++)
      if (!(R_FABS(J[keep[r] + integ[c] * m]) <= EPS)) allzero = 0;
  }
  if (allzero) return 0;
  const int info = chol_upper(Ji, ni, ni);
  if (info) return info;
  /* messageJ = Jk - (Jki/U)(Jki/U)' : row r of Z solves U' z = Jki[r,:]' */
  real Zall[MAXM * MAXM];
  for (int r = 0; r < s; r++) {
    for (int c = 0; c < ni; c++) Z[c] = J[keep[r] + integ[c] * m];
    solve_ut(Ji, ni, ni, Z);
    memcpy(Zall + (size_t)r * ni, Z, sizeof(real) * ni);
  }
  for (int c = 0; c < s; c++)
    for (int r = 0; r < s; r++) {
      real d = 0.0;
      for (int k = 0; k < ni; k++) d += Zall[(size_t)r * ni + k] * Zall[(size_t)c * ni + k];
      mJ[r + c * s] -= d;
    }
  memcpy(mui, hi, sizeof(real) * ni);
  solve_ut(Ji, ni, ni, mui);
  solve_u(Ji, ni, ni, mui);
  real logdet = 0.0, quad = 0.0;
  for (int k = 0; k < ni; k++) {
    logdet += R_LOG(Ji[k + k * ni]);
    quad += hi[k] * mui[k];
  }
  logdet *= 2.0;
  for (int r = 0; r < s; r++) {
    real d = 0.0;
    for (int c = 0; c < ni; c++) d += J[keep[r] + integ[c] * m] * mui[c];
    mh[r] -= d;
  }
  *mg = g + (ni * LOG2PI - logdet + quad) / 2;
  return 0;
}

/* propagate_belief!(to, sepset j, from, residual): returns 0 or Cholesky info */
int pgbpo_propagate(const og_graph* G, real* state, real* resid, uint8_t* flags, int from, int j, int to,
                    int update_residnorm) {
  const int nc = G->nclusters;
  int side_from, side_to;
  if (from == G->sep_a[j] && to == G->sep_b[j]) { side_from = 0; side_to = 1; }
  else if (from == G->sep_b[j] && to == G->sep_a[j]) { side_from = 1; side_to = 0; }
  else return -1;
  const int mF = G->dim[from], mT = G->dim[to], s = G->dim[nc + j];
  const int32_t* upF = G->up + G->up_off[2 * j + side_from];
  const int32_t* upT = G->up + G->up_off[2 * j + side_to];
  real* F = state + G->off[from];
  real* S = state + G->off[nc + j];
  real* T = state + G->off[to];
  real mJ[MAXM * MAXM], mh[MAXM], mg;
  const int info = marginalize(F, F + mF * mF, F[mF * mF + mF], mF, upF, s, mJ, mh, &mg);
  if (info) return info;
  real* SJ = S; real* Sh = S + s * s; real* Sg = S + s * s + s;
  real* TJ = T; real* Th = T + mT * mT; real* Tg = T + mT * mT + mT;
  const int d = 2 * j + side_to; /* residual of the message INTO `to`: key (to, from) */
  real* RJ = resid ? resid + G->roff[d] : NULL;
  real* Rh = RJ ? RJ + s * s : NULL;
  real maxJ = 0.0, maxh = 0.0;
  for (int c = 0; c < s; c++) {
    for (int r = 0; r < s; r++) {
      const real dJ = mJ[r + c * s] - SJ[r + c * s];
      SJ[r + c * s] = mJ[r + c * s];
      TJ[upT[r] + upT[c] * mT] += dJ;
      if (RJ) RJ[r + c * s] = dJ;
      const real a = R_FABS(dJ / R_SQRT((real)(s * s)));
      if (a > maxJ || a != a) maxJ = a;
    }
    const real dh = mh[c] - Sh[c];
    Sh[c] = mh[c];
    Th[upT[c]] += dh;
    if (Rh) Rh[c] = dh;
    const real a = R_FABS(dh / R_SQRT((real)s));
    if (a > maxh || a != a) maxh = a;
  }
  const real dg = mg - *Sg;
  *Sg = mg;
  *Tg += dg;
  if (update_residnorm && flags) flags[d] = (s == 0) ? 1 : ((maxh <= 1e-5) && (maxJ <= 1e-5));
  return 0;
}

/* one traversal; returns 0, or ((ref+1)<<8 | info) at the first failing message (then stops) */
static int traverse(const og_graph* G, real* state, real* resid, uint8_t* flags, const int32_t* tsep,
                    const int32_t* tpar, const int32_t* tchi, int n, int preorder, int upd, int ref_base) {
  for (int r = 0; r < n; r++) {
    const int i = preorder ? r : n - 1 - r;
    const int from = preorder ? tpar[i] : tchi[i], to = preorder ? tchi[i] : tpar[i];
    const int info = pgbpo_propagate(G, state, resid, flags, from, tsep[i], to, upd);
    if (info) return ((ref_base + r + 1) << 8) | (info & 0xff);
  }
  return 0;
}

/* calibrate!(beliefs, schedule, niter; auto): trees concatenated; returns status word (0 ok) */
int pgbpo_calibrate(const og_graph* G, real* state, real* resid, uint8_t* flags, int ntrees,
                    const int32_t* tree_off, const int32_t* tsep, const int32_t* tpar, const int32_t* tchi, int niter,
                    int do_post, int do_pre, int upd, int autostop, int32_t* iscal_out, int32_t* iter_tree) {
  int ref = 0, iscal = 0;
  if (iter_tree) iter_tree[0] = iter_tree[1] = 0;
  for (int it = 1; it <= niter; it++)
    for (int t = 0; t < ntrees; t++) {
      const int o = tree_off[t], n = tree_off[t + 1] - o;
      /* the reference runs the preorder pass even if the postorder pass failed
         (src/calibration.jl:79-80); a failed element is reported, its beliefs are not compared */
      if (do_post) { const int st = traverse(G, state, resid, flags, tsep + o, tpar + o, tchi + o, n, 0, upd, ref); if (st) return st; ref += n; }
      if (do_pre) { const int st = traverse(G, state, resid, flags, tsep + o, tpar + o, tchi + o, n, 1, upd, ref); if (st) return st; ref += n; }
      iscal = 1;
      if (flags) for (int d = 0; d < 2 * G->nsepsets; d++) if (!flags[d]) { iscal = 0; break; }
      if (!flags) iscal = 0;
      if (iscal) {
        if (iter_tree && iter_tree[0] == 0) { iter_tree[0] = it; iter_tree[1] = t + 1; }
        if (autostop) { if (iscal_out) *iscal_out = 1; return 0; }
      }
    }
  if (iscal_out) *iscal_out = iscal;
  return 0;
}

/* integratebelief: returns info; mu may be NULL */
int pgbpo_integrate(const og_graph* G, const real* state, int b, real* mu, real* norm) {
  const int m = G->dim[b];
  const real* J = state + G->off[b];
  const real* h = J + m * m;
  const real g = h[m];
  int zero = 1;
  for (int k = 0; k < m * m; k++) if (J[k] != 0.0) zero = 0;
  for (int k = 0; k < m; k++) if (h[k] != 0.0) zero = 0;
  if (zero) {
    if (mu) for (int k = 0; k < m; k++) mu[k] = INFINITY;
    *norm = g;
    return 0;
  }
  real U[MAXM * MAXM], x[MAXM];
  memcpy(U, J, sizeof(real) * m * m);
  const int info = chol_upper(U, m, m);
  if (info) { *norm = (real)NAN; return info; }
  memcpy(x, h, sizeof(real) * m);
  solve_ut(U, m, m, x);
  solve_u(U, m, m, x);
  real logdet = 0.0, quad = 0.0;
  for (int k = 0; k < m; k++) { logdet += R_LOG(U[k + k * m]); quad += h[k] * x[k]; }
  if (mu) memcpy(mu, x, sizeof(real) * m);
  *norm = g + (m * LOG2PI - 2.0 * logdet + quad) / 2;
  return 0;
}

/* free_energy: out = (energy, entropy, factored energy = -(energy - entropy)) */
int pgbpo_factored_energy(const og_graph* G, const real* state, const real* factor, real* out) {
  real en = 0.0, ent = 0.0;
  real U[MAXM * MAXM], mu[MAXM], col[MAXM];
  for (int c = 0; c < G->nclusters; c++) {
    const int m = G->dim[c];
    const real* fJ = factor + G->off[c]; const real* fh = fJ + m * m; const real fg = fh[m];
    if (m == 0) { en -= fg; continue; }
    const real* bJ = state + G->off[c]; const real* bh = bJ + m * m;
    memcpy(U, bJ, sizeof(real) * m * m);
    if (chol_upper(U, m, m)) { out[0] = out[1] = out[2] = NAN; return 1; }
    memcpy(mu, bh, sizeof(real) * m);
    solve_ut(U, m, m, mu); solve_u(U, m, m, mu);
    real tr = 0.0, quad = 0.0, hm = 0.0, logdet = 0.0;
    for (int k = 0; k < m; k++) {
      memcpy(col, fJ + k * m, sizeof(real) * m);
      real fm = 0.0;
      for (int r = 0; r < m; r++) fm += col[r] * mu[r];
      quad += mu[k] * fm;
      solve_ut(U, m, m, col); solve_u(U, m, m, col);
      tr += col[k];
      hm += fh[k] * mu[k];
      logdet += R_LOG(U[k + k * m]);
    }
    en += (tr + quad) / 2 - hm - fg;
    ent += (m * (LOG2PI + 1) - 2.0 * logdet) / 2;
  }
  for (int j = 0; j < G->nsepsets; j++) {
    const int m = G->dim[G->nclusters + j];
    if (m == 0) continue;
    memcpy(U, state + G->off[G->nclusters + j], sizeof(real) * m * m);
    real logdet = 0.0;
    if (chol_upper(U, m, m)) logdet = NAN;
    else { for (int k = 0; k < m; k++) logdet += R_LOG(U[k + k * m]); logdet *= 2.0; }
    ent -= (m * (LOG2PI + 1) - logdet) / 2;
  }
  out[0] = en; out[1] = ent; out[2] = -(en - ent);
  return 0;
}

/* small SPD inverse + logdet (Cholesky), n <= 16 */
/* regularizebeliefs_bycluster! (src/clustergraphbeliefs.jl:235-249) with
 * regularizebeliefs_1clustersepset! (:264-275): clusters in index order; eps_c = max(eps, max|J_c|)
 * taken once before the neighbour loop; neighbours in increasing cluster index (neighbor_labels).
 * nbr_off / nbr_sep: CSR list of the sepsets incident to each cluster in that order (pgbpo_neighbours). */
void pgbpo_neighbours(const og_graph* G, int32_t* nbr_off, int32_t* nbr_sep) {
  const int nc = G->nclusters, ns = G->nsepsets;
  int32_t* cnt = (int32_t*)calloc((size_t)nc + 1, sizeof(int32_t));
  for (int j = 0; j < ns; j++) { cnt[G->sep_a[j]]++; cnt[G->sep_b[j]]++; }
  nbr_off[0] = 0;
  for (int c = 0; c < nc; c++) nbr_off[c + 1] = nbr_off[c] + cnt[c];
  memset(cnt, 0, sizeof(int32_t) * (size_t)nc);
  for (int j = 0; j < ns; j++) {
    nbr_sep[nbr_off[G->sep_a[j]] + cnt[G->sep_a[j]]++] = j;
    nbr_sep[nbr_off[G->sep_b[j]] + cnt[G->sep_b[j]]++] = j;
  }
  for (int c = 0; c < nc; c++) {  /* insertion sort by the other end's cluster index */
    for (int x = nbr_off[c] + 1; x < nbr_off[c + 1]; x++) {
      const int j = nbr_sep[x];
      const int key = (G->sep_a[j] == c) ? G->sep_b[j] : G->sep_a[j];
      int y = x - 1;
      while (y >= nbr_off[c]) {
        const int jj = nbr_sep[y];
        const int k2 = (G->sep_a[jj] == c) ? G->sep_b[jj] : G->sep_a[jj];
        if (k2 <= key) break;
        nbr_sep[y + 1] = jj;
        y--;
      }
      nbr_sep[y + 1] = j;
    }
  }
  free(cnt);
}
void pgbpo_regularize_bycluster(const og_graph* G, real* state, const int32_t* nbr_off, const int32_t* nbr_sep) {
  const int nc = G->nclusters;
  for (int c = 0; c < nc; c++) {
    const int m = G->dim[c];
    real* J = state + G->off[c];
    real eps = EPS;
    for (int q = 0; q < m * m; q++) { const real a = R_FABS(J[q]); if (a > eps || a != a) eps = a; }
    for (int x = nbr_off[c]; x < nbr_off[c + 1]; x++) {
      const int j = nbr_sep[x];
      const int s = G->dim[nc + j];
      if (s == 0) continue;
      const int side = (G->sep_a[j] == c) ? 0 : 1;
      const int32_t* up = G->up + G->up_off[2 * j + side];
      real* Js = state + G->off[nc + j];
      for (int k = 0; k < s; k++) {
        J[(size_t)up[k] * m + up[k]] += eps;
        Js[(size_t)k * s + k] += eps;
      }
    }
  }
}

static int spd_inv(const real* A, int n, real* inv, real* logdet) {
  real U[16 * 16], e[16];
  memcpy(U, A, sizeof(real) * n * n);
  const int info = chol_upper(U, n, n);
  if (info) return info;
  real ld = 0.0;
  for (int k = 0; k < n; k++) ld += R_LOG(U[k + k * n]);
  *logdet = 2.0 * ld;
  for (int c = 0; c < n; c++) {
    memset(e, 0, sizeof e); e[c] = 1.0;
    solve_ut(U, n, n, e); solve_u(U, n, n, e);
    memcpy(inv + c * n, e, sizeof(real) * n);
  }
  return 0;
}

/* assignfactors! for Brownian motion without missing data, in the reference's order of
 * operations: build phi_v (child block first, parents next), absorb the leaf's data, then the
 * fixed root's mean (absorbevidence!), then mult! into the cluster.  params = one parameter set
 * laid out as in pgbp_assign_factors; tip = one data set [ntips][p]. */
int pgbpo_assign_bm(const og_graph* G, const og_families* F, int ncolors, const double* params_d, const double* tip,
                    real* state) {
  const int p = G->ntraits, pp = p * p;
  const int nb = G->nclusters + G->nsepsets;
  memset(state, 0, sizeof(real) * (size_t)G->off[nb]);
  real params[8 * 16 * 16 + 16 + 16 * 16];
  if (ncolors > 8 || p > 16) return -1;
  for (int k = 0; k < ncolors * pp + p + pp; k++) params[k] = params_d[k];
  const real* mu = params + ncolors * pp;
  const real* v = mu + p;
  real Pinv[8][16 * 16], g0[8];
  if (ncolors > 8 || p > 16) return -1;
  for (int c = 0; c < ncolors; c++) {
    real ld;
    const int info = spd_inv(params + c * pp, p, Pinv[c], &ld);
    if (info) return info;
    g0[c] = -(p * LOG2PI + ld) / 2;
  }
  int rootkind = 1, allzero = 1;
  for (int k = 0; k < pp; k++) if (v[k] != 0.0) allzero = 0;
  if (allzero) rootkind = 0;
  else for (int k = 0; k < p; k++) if (R_ISINF(v[k + k * p])) rootkind = 2;
  for (int node = 0; node < F->nnodes; node++) {
    const int c = F->node_cluster[node], m = G->dim[c];
    real* J = state + G->off[c]; real* h = J + m * m; real* g = h + m;
    const int k0 = F->mem_off[node], nm = F->mem_off[node + 1] - k0;
    if (nm == 1) {
      const int pos = F->mem_pos[k0];
      if (pos < 0 || rootkind != 1) continue;
      real jr[16 * 16], ld;
      const int info = spd_inv(v, p, jr, &ld);
      if (info) return info;
      real quad = 0.0;
      for (int r = 0; r < p; r++) {
        real s = 0.0;
        for (int cc = 0; cc < p; cc++) { s += jr[r + cc * p] * mu[cc]; J[(pos + r) + (pos + cc) * m] += jr[r + cc * p]; }
        h[pos + r] += s; quad += mu[r] * s;
      }
      *g += (-p * LOG2PI - ld - quad) / 2;
      continue;
    }
    /* phi = (hh, JJ, gg) on nm*p variables */
    const int n = nm * p;
    real JJ[(8 * 16) * (8 * 16)], hh[8 * 16], gg, j[16 * 16], cf[8];
    if (nm > 8) return -2;
    int same = 1;
    for (int k = k0 + 2; k < k0 + nm; k++) if (F->mem_color[k] != F->mem_color[k0 + 1]) same = 0;
    if (same) {
      const int col = F->mem_color[k0 + 1];
      real t0 = 0.0;
      if (nm == 2) t0 = F->mem_length[k0 + 1];
      else for (int k = k0 + 1; k < k0 + nm; k++) t0 += F->mem_gamma[k] * F->mem_gamma[k] * F->mem_length[k];
      for (int q = 0; q < pp; q++) j[q] = Pinv[col][q] / t0;
      gg = g0[col] - p * R_LOG(t0) / 2;
    } else {
      real V[16 * 16], ld;
      memset(V, 0, sizeof V);
      for (int k = k0 + 1; k < k0 + nm; k++) {
        const real f = F->mem_gamma[k] * F->mem_gamma[k] * F->mem_length[k];
        for (int q = 0; q < pp; q++) V[q] += f * params[F->mem_color[k] * pp + q];
      }
      const int info = spd_inv(V, p, j, &ld);
      if (info) return info;
      gg = -(p * LOG2PI + ld) / 2;
    }
    cf[0] = 1.0;
    for (int a = 1; a < nm; a++) cf[a] = (nm == 2) ? -1.0 : -F->mem_gamma[k0 + a];
    for (int a = 0; a < nm; a++)
      for (int b = 0; b < nm; b++)
        for (int tb = 0; tb < p; tb++)
          for (int ta = 0; ta < p; ta++) JJ[(a * p + ta) + (b * p + tb) * n] = cf[a] * cf[b] * j[ta + tb * p];
    memset(hh, 0, sizeof(real) * n);
    /* absorbevidence! (src/beliefupdates.jl:210-231), one fixed member at a time, child first */
    char gone[8 * 16];
    memset(gone, 0, sizeof gone);
    for (int a = 0; a < nm; a++) {
      if (F->mem_pos[k0 + a] >= 0) continue;
      real y[16];
      for (int t = 0; t < p; t++) y[t] = (a == 0) ? tip[F->node_datarow[node] * p + t] : mu[t];
      real hay = 0.0, yJy = 0.0;
      for (int ta = 0; ta < p; ta++) {
        hay += hh[a * p + ta] * y[ta];
        for (int tb = 0; tb < p; tb++) yJy += y[ta] * JJ[(a * p + ta) + (a * p + tb) * n] * y[tb];
      }
      gg += hay - yJy / 2;
      for (int r = 0; r < n; r++) {
        if (gone[r] || (r / p) == a) continue;
        real s = 0.0;
        for (int t = 0; t < p; t++) s += JJ[r + (a * p + t) * n] * y[t];
        hh[r] -= s;
      }
      for (int t = 0; t < p; t++) gone[a * p + t] = 1;
    }
    for (int a = 0; a < nm; a++) {
      const int pa = F->mem_pos[k0 + a];
      if (pa < 0) continue;
      for (int ta = 0; ta < p; ta++) {
        h[pa + ta] += hh[a * p + ta];
        for (int b = 0; b < nm; b++) {
          const int pb = F->mem_pos[k0 + b];
          if (pb < 0) continue;
          for (int tb = 0; tb < p; tb++) J[(pa + ta) + (pb + tb) * m] += JJ[(a * p + ta) + (b * p + tb) * n];
        }
      }
    }
    *g += gg;
  }
  return 0;
}

/* Batched driver: for each replicate e (OpenMP): assign factors, calibrate, integrate
 * `root_belief`, optionally factored energy, optionally copy out the final state.
 * pairing as in pgbp_assign_factors.  Outputs: loglik[B], status[B], fe[3B] (may be NULL),
 * state_out [B][state_size] (may be NULL). */
int pgbpo_run_batch(const og_graph* G, const og_families* F, int ncolors, const double* params, int64_t np_,
                    const double* tip, int64_t nd, int pairing, int ntrees, const int32_t* tree_off,
                    const int32_t* tsep, const int32_t* tpar, const int32_t* tchi, int niter, int do_post, int do_pre,
                    int upd, int autostop, int root_belief, int64_t B, double* loglik, int32_t* status, double* fe,
                    double* state_out, int32_t* iscal_out, int nthreads, int reg_bycluster) {
  const int nb = G->nclusters + G->nsepsets;
  const int64_t ssize = G->off[nb], rsize = G->roff[2 * G->nsepsets];
  const int p = G->ntraits;
  const int64_t plen = (int64_t)ncolors * p * p + p + (int64_t)p * p, tlen = (int64_t)F->ntips * p;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int32_t* nbr_off = NULL; int32_t* nbr_sep = NULL;
  if (reg_bycluster) {
    nbr_off = (int32_t*)malloc(sizeof(int32_t) * ((size_t)G->nclusters + 1));
    nbr_sep = (int32_t*)malloc(sizeof(int32_t) * (2 * (size_t)G->nsepsets + 1));
    pgbpo_neighbours(G, nbr_off, nbr_sep);
  }
#pragma omp parallel
  {
    real* state = (real*)malloc(sizeof(real) * (size_t)(ssize > 0 ? ssize : 1));
    real* factor = fe ? (real*)malloc(sizeof(real) * (size_t)(ssize > 0 ? ssize : 1)) : NULL;
    real* resid = (real*)malloc(sizeof(real) * (size_t)(rsize > 0 ? rsize : 1));
    uint8_t* flags = (uint8_t*)malloc((size_t)(2 * G->nsepsets + 1));
#pragma omp for schedule(static)
    for (int64_t e = 0; e < B; e++) {
      int64_t ip, id;
      if (pairing == 1) { ip = e / nd; id = e % nd; }
      else { ip = np_ == 1 ? 0 : e; id = nd == 1 ? 0 : e; }
      int st = pgbpo_assign_bm(G, F, ncolors, params + ip * plen, tip + id * tlen, state);
      if (st) { status[e] = (0x7ffffd << 8) | (st & 0xff); loglik[e] = NAN; continue; }
      if (factor) memcpy(factor, state, sizeof(real) * (size_t)ssize);
      if (reg_bycluster) pgbpo_regularize_bycluster(G, state, nbr_off, nbr_sep);
      for (int j = 0; j < G->nsepsets; j++) flags[2 * j] = flags[2 * j + 1] = (G->dim[G->nclusters + j] == 0);
      int32_t isc = 0;
      st = pgbpo_calibrate(G, state, resid, flags, ntrees, tree_off, tsep, tpar, tchi, niter, do_post, do_pre, upd,
                           autostop, &isc, NULL);
      status[e] = st;
      if (iscal_out) iscal_out[e] = st ? 0 : isc;
      if (st) { loglik[e] = NAN; continue; }
      real nrm, fe3[3];
      if (pgbpo_integrate(G, state, root_belief, NULL, &nrm)) status[e] = (0x7ffffe << 8) | 1;
      loglik[e] = (double)nrm;
      if (fe) {
        pgbpo_factored_energy(G, state, factor, fe3);
        for (int k = 0; k < 3; k++) fe[3 * e + k] = (double)fe3[k];
      }
      if (state_out) for (int64_t k = 0; k < ssize; k++) state_out[e * ssize + k] = (double)state[k];
    }
    free(state); free(factor); free(resid); free(flags);
  }
  free(nbr_off); free(nbr_sep);
  return 0;
}
