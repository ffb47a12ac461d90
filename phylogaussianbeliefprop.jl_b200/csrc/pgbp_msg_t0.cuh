// K2 register-resident ("T0") and thread-local generic ("T1") message kernels + their launcher.
// Included by several translation units (the 78 T0 shapes are split over three of them so that the
// library builds in about a minute instead of three).
#pragma once
#include "pgbp_kernels.cuh"
#include "pgbp_launch.h"
#include "pgbp_shapes.h"

namespace pgbp {

#define PGBP_MSG_THREADS 128
#ifndef PGBP_HOST_EMUL
#ifndef PGBP_MSG_MINBLOCKS
#define PGBP_MSG_MINBLOCKS 1  // measured on C2 (B200): 1 -> 0.641 of HBM roofline, 2 -> 0.641, 3 -> 0.594, 4 -> 0.548
#endif
template <int CI, int CS, int MAXM>
__global__ void __launch_bounds__(PGBP_MSG_THREADS, (CI >= 0 ? PGBP_MSG_MINBLOCKS : 1)) k_message(MsgArgs a) {
  const int64_t e = a.e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.B) return;
  if constexpr (CI >= 0) message_thread_t0<CI, CS>(a, blockIdx.y, e);
  else message_thread_rt<MAXM>(a, blockIdx.y, e);
}
#endif

template <int CI, int CS, int MAXM>
static inline int launch_message(pgbp_batch* b, const MsgArgs& a, int nmsg) {
#ifdef PGBP_HOST_EMUL
  for (int m = 0; m < nmsg; m++)
    for (int64_t e = a.e0; e < a.B; e++) {
      if constexpr (CI >= 0) message_thread_t0<CI, CS>(a, m, e);
      else message_thread_rt<MAXM>(a, m, e);
    }
#else
  dim3 grid((unsigned)((a.B - a.e0 + PGBP_MSG_THREADS - 1) / PGBP_MSG_THREADS), (unsigned)nmsg);
  k_message<CI, CS, MAXM><<<grid, PGBP_MSG_THREADS, 0, b->stream>>>(a);
#endif
  b->launches++;
  return check_launch("k_message");
}

// parts of the T0 shape table (pgbp_message_t0.cu compiled with -DPGBP_T0_PART=0,1,2): return
// PGBP_NOT_MINE when (ci, cs) belongs to another part
#define PGBP_NOT_MINE 12345
int launch_t0_part0(pgbp_batch* b, const MsgArgs& a, int nmsg, int ci, int cs);
int launch_t0_part1(pgbp_batch* b, const MsgArgs& a, int nmsg, int ci, int cs);
int launch_t0_part2(pgbp_batch* b, const MsgArgs& a, int nmsg, int ci, int cs);
// medium / large shapes (pgbp_message_medium.cu): shared-memory, cooperative or generic kernel
int launch_medium(pgbp_batch* b, const MsgArgs& a, int nmsg, int I, int S);

}  // namespace pgbp
