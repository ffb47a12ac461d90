// Shape classes of the message kernel (K2).
//  T0  register-resident, fully unrolled: i >= 1 variables integrated out, i + s <= PGBP_T0_MAX
//  COPY  i == 0 (nothing to integrate): streaming, any s
//  MEDIUM  i + s > PGBP_T0_MAX, class id ci = -2, launch groups uniform in (i, s) = (maxm, cs):
//      T0S  one thread per element, chol(J_II) / Z / w in shared memory (k_message_smem), used
//           whenever 32 * 8 * (tri(i) + i*s + i) bytes fit the SM's shared memory and i <= 32;
//      COOP G lanes of a warp per element, sender belief column-cyclic in registers
//           (k_message_coop), i + s <= 32;
//      T1   generic runtime dimensions, matrix in thread-local memory (what is left).
// The X-macro below is generated (see the bottom of this file for the recipe).
#pragma once
#define PGBP_MAX_DIM 64
#define PGBP_MAX_FAMILY 8
#define PGBP_MAX_TRAITS 16
#define PGBP_T0_MAX 12
#define PGBP_COOP_MAX 48
#define PGBP_WALK_MAXP 4
#define PGBP_SCOPED_MAXN 48  // variables of one node family in the scoped (missing-data) K1 path

namespace pgbp {
inline void shape_class(int i, int s, int* ci, int* cs, int* maxm) {
  if (i == 0) { *ci = 0; *cs = -1; *maxm = 0; return; }
  if (i + s <= PGBP_T0_MAX) { *ci = i; *cs = s; *maxm = 0; return; }
  *ci = -2; *cs = s; *maxm = i;
}
// "Walk" family for ntraits = P: a message whose integrated / kept dimensions are
// a*P and b*P (a <= 3, b <= 3, whole nodes in scope) gets id 4a+b; anything else -1
// (then the traversal uses the level-parallel launches).
inline int walk_shape_id(int P, int i, int s) {
  if (P < 1 || P > PGBP_WALK_MAXP || i % P || s % P) return -1;
  const int a = i / P, b = s / P;
  if (a > 3 || b > 3) return -1;
  if (a > 0 && (a + b) * P > PGBP_T0_MAX) return -1;
  return 4 * a + b;
}
}  // namespace pgbp

#define PGBP_T0_SHAPES(X) \
  X(1,0) X(1,1) X(1,2) X(1,3) X(1,4) X(1,5) X(1,6) X(1,7) X(1,8) X(1,9) X(1,10) X(1,11) \
  X(2,0) X(2,1) X(2,2) X(2,3) X(2,4) X(2,5) X(2,6) X(2,7) X(2,8) X(2,9) X(2,10) \
  X(3,0) X(3,1) X(3,2) X(3,3) X(3,4) X(3,5) X(3,6) X(3,7) X(3,8) X(3,9) \
  X(4,0) X(4,1) X(4,2) X(4,3) X(4,4) X(4,5) X(4,6) X(4,7) X(4,8) \
  X(5,0) X(5,1) X(5,2) X(5,3) X(5,4) X(5,5) X(5,6) X(5,7) \
  X(6,0) X(6,1) X(6,2) X(6,3) X(6,4) X(6,5) X(6,6) \
  X(7,0) X(7,1) X(7,2) X(7,3) X(7,4) X(7,5) \
  X(8,0) X(8,1) X(8,2) X(8,3) X(8,4) \
  X(9,0) X(9,1) X(9,2) X(9,3) \
  X(10,0) X(10,1) X(10,2) \
  X(11,0) X(11,1) \
  X(12,0)

// recipe: for i in 1..PGBP_T0_MAX: for s in 0..PGBP_T0_MAX-i: X(i,s)
