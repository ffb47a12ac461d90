// Per-thread bodies of the hot kernels.  One thread = one batch element of one
// message / belief.  All global accesses are `base[slot * ld + e]`: consecutive
// threads touch consecutive doubles (batch-innermost SoA), so every warp load
// or store is one fully coalesced 256-byte transaction.
#pragma once
#include "pgbp_internal.h"

namespace pgbp {

#define PGBP_LOG2PI 1.8378770664093454835606594728112
#define PGBP_EPS 2.220446049250313e-16

struct MsgArgs {
  const MsgDesc* msgs;
  const int32_t* tab;
  double* state;
  double* resid;     // may be null
  uint8_t* calflag;  // may be null
  int32_t* status;
  const uint8_t* done;  // may be null
  int64_t B, ld;
  uint32_t opts;     // PGBP_CAL_RESIDNORM
  int32_t ref_base;
};

// NaN-propagating running maximum of |x| (Julia's maximum(abs, x))
PGBP_HD void absmax(double& m, double x) {
  const double a = fabs(x);
  if (a > m || a != a) m = a;
}

struct TrailJ {
  const double* A;
  int I;
  PGBP_HD double operator()(int r, int c, int) const { return A[pk(I + r, I + c)]; }
};
struct TrailH {
  const double* hv;
  int I;
  PGBP_HD double operator()(int k) const { return hv[I + k]; }
};
struct GatherJ {
  const double* st;
  const int32_t* gat;
  int64_t base, ld;
  PGBP_HD double operator()(int, int, int q) const { return st[(base + gat[q]) * ld]; }
};
struct GatherH {
  const double* st;
  const int32_t* gat;
  int64_t base, ld;
  PGBP_HD double operator()(int k) const { return st[(base + gat[k]) * ld]; }
};

// Divide by the sepset, multiply into the receiver, store the residual and the
// calibration flag: src/beliefupdates.jl:579-587 (divide!), :483-488 (mult!),
// :646-647 (residual), src/beliefs.jl:994-1003 (iscalibrated_residnorm!).
// `newJ(q)`, `newh(k)` give the outgoing message in sepset order.
template <class FJ, class FH>
PGBP_HD void divide_mult_store(const MsgArgs& a, const MsgDesc& md, int64_t e, int S, FJ newJ, FH newh,
                               double newg) {
  const int64_t ld = a.ld;
  double* st = a.state + e;
  double* rs = a.resid ? a.resid + e : nullptr;
  const int32_t* sca = a.tab + md.sca;
  const int SS = tri(S);
  double maxJ = 0.0, maxh = 0.0;
#pragma unroll
  for (int c = 0; c < S; c++) {
#pragma unroll
    for (int r = 0; r <= c; r++) {
      const int q = pk(r, c);
      const double nv = newJ(r, c, q);
      double* sp = st + (md.sJ + q) * ld;
      const double d = nv - *sp;
      *sp = nv;
      st[(md.tJ + sca[q]) * ld] += d;
      if (rs) rs[(md.rJ + q) * ld] = d;
      absmax(maxJ, d);
    }
  }
#pragma unroll
  for (int k = 0; k < S; k++) {
    const double nv = newh(k);
    double* sp = st + (md.sh + k) * ld;
    const double d = nv - *sp;
    *sp = nv;
    st[(md.th + sca[SS + k]) * ld] += d;
    if (rs) rs[(md.rh + k) * ld] = d;
    absmax(maxh, d);
  }
  {
    double* sp = st + md.sg * ld;
    const double d = newg - *sp;
    *sp = newg;
    st[md.tg * ld] += d;
  }
  if ((a.opts & PGBP_CAL_RESIDNORM) && a.calflag) {
    bool ok = true;
    if (S > 0) {
      // max_i |x_i| / sqrt(n) == max_i (|x_i| / sqrt(n)) exactly (monotone rounding)
      ok = (maxh / sqrt((double)S) <= 1e-5) && (maxJ / (double)S <= 1e-5);
    }
    a.calflag[(int64_t)md.dmsg * ld + e] = ok ? 1 : 0;
  }
}

// Message with i >= 1 variables integrated out (marginalize,
// src/beliefupdates.jl:55-83).  The sender is gathered in [I;K] order so that
// the Schur complement is the first i pivots of a right-looking Cholesky; the
// trailing s x s block then holds J_K - J_KI J_I^-1 J_IK, the eliminated h gives
// h_K - J_KI J_I^-1 h_I, and w = U^-T h_I gives h_I' J_I^-1 h_I = |w|^2.
// CI/CS >= 0: compile-time shape, arrays live in registers.  CI < 0: runtime
// shape, arrays live in thread-local memory sized for MAXM.
template <int CI, int CS, int MAXM>
PGBP_HD void message_thread(const MsgArgs& a, int msg_index, int64_t e) {
  constexpr bool RT = (CI < 0);
  constexpr int CM = RT ? MAXM : (CI + CS);
  constexpr int NA = CM * (CM + 1) / 2;
  const MsgDesc& md = a.msgs[msg_index];
  if (a.status[e] != 0) return;
  if (a.done && a.done[e]) return;
  const int I = RT ? (md.mF - md.s) : CI;
  const int S = RT ? md.s : CS;
  const int M = I + S;
  const int64_t ld = a.ld;
  const double* st = a.state + e;
  const int32_t* gat = a.tab + md.gat;
  double A[NA > 0 ? NA : 1];
  double hv[CM > 0 ? CM : 1];
  const int SM = tri(M);
#pragma unroll
  for (int q = 0; q < SM; q++) A[q] = st[(md.fJ + gat[q]) * ld];
#pragma unroll
  for (int k = 0; k < M; k++) hv[k] = st[(md.fh + gat[SM + k]) * ld];
  double g = st[md.fg * ld];

  // "Ji = Jki = hi = 0 if missing data" shortcut, src/beliefupdates.jl:62-66
  bool allzero = true;
#pragma unroll
  for (int c = 0; c < M; c++) {
    const int rmax = c < I ? c + 1 : I;
#pragma unroll
    for (int r = 0; r < rmax; r++)
      if (!(fabs(A[pk(r, c)]) <= PGBP_EPS)) allzero = false;
  }
#pragma unroll
  for (int k = 0; k < I; k++)
    if (!(fabs(hv[k]) <= PGBP_EPS)) allzero = false;

  if (!allzero) {
    double logdet = 0.0, ww = 0.0;
#pragma unroll
    for (int k = 0; k < I; k++) {
      const double d = A[pk(k, k)];
      if (!(d > 0.0)) {  // LAPACK potrf: info = k+1 (also catches NaN)
        status_fail(a.status, e, PGBP_STATUS(a.ref_base + md.ref, k + 1));
        return;
      }
      logdet += log(d);
      const double rinv = 1.0 / sqrt(d);
#pragma unroll
      for (int c = k + 1; c < M; c++) A[pk(k, c)] *= rinv;
      const double wk = hv[k] * rinv;
      ww += wk * wk;
#pragma unroll
      for (int c = k + 1; c < M; c++) {
        const double akc = A[pk(k, c)];
#pragma unroll
        for (int r = k + 1; r <= c; r++) A[pk(r, c)] -= A[pk(k, r)] * akc;
        hv[c] -= akc * wk;
      }
    }
    g += 0.5 * ((double)I * PGBP_LOG2PI - logdet + ww);
  }
  // trailing block -> sepset order
  divide_mult_store(a, md, e, S, TrailJ{A, I}, TrailH{hv, I}, g);
}

// Message with nothing to integrate out (src/beliefupdates.jl:56): the outgoing
// message is the sender's belief re-ordered; streamed, no local storage.
PGBP_HD void message_copy_thread(const MsgArgs& a, int msg_index, int64_t e) {
  const MsgDesc& md = a.msgs[msg_index];
  if (a.status[e] != 0) return;
  if (a.done && a.done[e]) return;
  const int S = md.s;
  const int64_t ld = a.ld;
  const double* st = a.state + e;
  const int32_t* gat = a.tab + md.gat;
  const int SS = tri(S);
  const double g = st[md.fg * ld];
  divide_mult_store(a, md, e, S, GatherJ{st, gat, md.fJ, ld}, GatherH{st, gat + SS, md.fh, ld}, g);
}

// integratebelief (src/beliefupdates.jl:187-200): mu = J^-1 h,
// norm = g + (m log2pi - logdet J + h'mu)/2; all-zero (h,J) -> (Inf.., g).
template <int MAXM>
PGBP_HD void integrate_thread(const double* state, int32_t* status, int64_t ld, int64_t e, int64_t jslot,
                              int64_t hslot, int64_t gslot, int M, double* mu_soa, double* norm, int64_t ld_out) {
  constexpr int NA = MAXM * (MAXM + 1) / 2;
  double A[NA > 0 ? NA : 1];
  double hv[MAXM > 0 ? MAXM : 1];
  const double* st = state + e;
  const int SM = tri(M);
  bool zero = true;
  for (int q = 0; q < SM; q++) {
    A[q] = st[(jslot + q) * ld];
    if (A[q] != 0.0) zero = false;
  }
  for (int k = 0; k < M; k++) {
    hv[k] = st[(hslot + k) * ld];
    if (hv[k] != 0.0) zero = false;
  }
  const double g = st[gslot * ld];
  if (zero) {
    if (mu_soa)
      for (int k = 0; k < M; k++) mu_soa[k * ld_out + e] = INFINITY;
    norm[e] = g;
    return;
  }
  double logdet = 0.0, ww = 0.0;
  for (int k = 0; k < M; k++) {
    const double d = A[pk(k, k)];
    if (!(d > 0.0)) {
      status_fail(status, e, PGBP_STATUS(0x7ffffe, k + 1));
      if (mu_soa)
        for (int q = 0; q < M; q++) mu_soa[q * ld_out + e] = NAN;
      norm[e] = NAN;
      return;
    }
    logdet += log(d);
    const double rinv = 1.0 / sqrt(d);
    A[pk(k, k)] = rinv;
    for (int c = k + 1; c < M; c++) A[pk(k, c)] *= rinv;
    const double wk = hv[k] * rinv;
    hv[k] = wk;
    ww += wk * wk;
    for (int c = k + 1; c < M; c++) {
      const double akc = A[pk(k, c)];
      for (int r = k + 1; r <= c; r++) A[pk(r, c)] -= A[pk(k, r)] * akc;
      hv[c] -= akc * wk;
    }
  }
  norm[e] = g + 0.5 * ((double)M * PGBP_LOG2PI - logdet + ww);
  if (mu_soa) {
    for (int k = M - 1; k >= 0; k--) {  // U mu = w
      double s = hv[k];
      for (int c = k + 1; c < M; c++) s -= A[pk(k, c)] * hv[c];
      hv[k] = s * A[pk(k, k)];
      mu_soa[k * ld_out + e] = hv[k];
    }
  }
}

}  // namespace pgbp
