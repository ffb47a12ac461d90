// Multi-GPU gather of per-element results over NVLink peer memory.
//
// The path shards by batch element and has exactly one exchange: the gather of the per-replicate
// log-likelihoods (<= 0.5 MB per rank and step).  As an NCCL all-gather that is a separate collective launch per
// step whose fixed cost (~0.2 ms of launch latency and SM contention against 0.7 ms of message kernels, measured
// in round 1) sat on the step's critical path.  Here the exchange is FUSED into the producing kernel instead:
// every rank owns a window [nbuffers][nranks][ld] of doubles (+ one sequence flag per (buffer, rank)), exported
// with CUDA IPC and mapped by every peer; integratebelief! (k_integrate_gather) writes each log-likelihood into
// slot `rank` of EVERY rank's window with plain stores through the peer mappings (NVLink writes are posted:
// the kernel does not wait for them), then a one-warp kernel publishes the sequence number.  A consumer calls
// pgbp_comm_wait (a spin on its own local flags) before reading its window.  No collective, no extra stream.
#include <vector>

#include "pgbp_kernels.cuh"
#include "pgbp_launch.h"
#include "pgbp_shapes.h"

#define PGBP_COMM_MAXRANKS 16

struct pgbp_comm {
  int32_t device = 0, rank = 0, nranks = 1, nbuffers = 2;
  int64_t ld = 0;
  size_t bytes = 0;
  char* local = nullptr;                       // this rank's window
  char* peer[PGBP_COMM_MAXRANKS] = {nullptr};  // every rank's window as mapped here (peer[rank] == local)
  bool opened[PGBP_COMM_MAXRANKS] = {false};
  uint64_t seq[8] = {0};                       // per buffer: sequence number of this rank's last put
  int32_t* d_err = nullptr;                    // set by a wait that timed out
};

namespace pgbp {

struct PeerWin {
  double* data[PGBP_COMM_MAXRANKS];                 // slot `rank` of the chosen buffer in every rank's window
  unsigned long long* flag[PGBP_COMM_MAXRANKS];     // flag (buffer, rank) in every rank's window
};

static size_t comm_data_bytes(const pgbp_comm* c) { return sizeof(double) * (size_t)c->nbuffers * c->nranks * (size_t)c->ld; }
static double* comm_slot(const pgbp_comm* c, char* win, int buffer, int r) {
  return (double*)win + ((size_t)buffer * c->nranks + r) * (size_t)c->ld;
}
static unsigned long long* comm_flag(const pgbp_comm* c, char* win, int buffer, int r) {
  return (unsigned long long*)(win + comm_data_bytes(c)) + (size_t)buffer * c->nranks + r;
}
static PeerWin peer_win(const pgbp_comm* c, int buffer) {
  PeerWin w;
  for (int r = 0; r < PGBP_COMM_MAXRANKS; r++) {
    w.data[r] = r < c->nranks ? comm_slot(c, c->peer[r], buffer, c->rank) : nullptr;
    w.flag[r] = r < c->nranks ? comm_flag(c, c->peer[r], buffer, c->rank) : nullptr;
  }
  return w;
}

#ifndef PGBP_HOST_EMUL
// integratebelief! of one belief for every element, log-likelihood stored into slot `rank` of every rank's window
template <int MAXM>
__global__ void __launch_bounds__(128) k_integrate_gather(const double* state, int32_t* status, int64_t B, int64_t ld,
                                                          int64_t jslot, int64_t hslot, int64_t gslot, int M, JSide js,
                                                          PeerWin w, int rank, int nranks) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  double* mine = w.data[rank];
  integrate_thread<MAXM>(state, status, ld, e, jslot, hslot, gslot, M, nullptr, mine, ld, nullptr, jcolumn(js, e), js.ld);
  const double v = mine[e];
  for (int r = 0; r < nranks; r++)
    if (r != rank) w.data[r][e] = v;  // posted NVLink store
}
__global__ void k_comm_put(const double* __restrict__ src, int64_t n, PeerWin w, int nranks) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double v = src[e];
  for (int r = 0; r < nranks; r++) w.data[r][e] = v;
}
// runs after the producing kernel in stream order: its stores are complete and visible system-wide
__global__ void k_comm_signal(PeerWin w, int nranks, unsigned long long seq) {
  const int r = threadIdx.x;
  if (r < nranks) {
    __threadfence_system();
    *(volatile unsigned long long*)w.flag[r] = seq;
    __threadfence_system();
  }
}
__global__ void k_comm_wait(const unsigned long long* flags, int nranks, unsigned long long seq, long long timeout_cycles,
                            int32_t* err) {
  const int r = threadIdx.x;
  if (r >= nranks) return;
  const long long t0 = clock64();
  while (*(volatile const unsigned long long*)(flags + r) < seq) {
    if (clock64() - t0 > timeout_cycles) { atomicExch(err, r + 1); break; }
    __nanosleep(200);
  }
  __threadfence_system();
}
#endif

}  // namespace pgbp

using namespace pgbp;

extern "C" {

int32_t pgbp_comm_create(int32_t device, int32_t rank, int32_t nranks, int64_t ld, int32_t nbuffers, pgbp_comm** out) {
  if (!out || nranks < 1 || nranks > PGBP_COMM_MAXRANKS || rank < 0 || rank >= nranks || ld <= 0 || nbuffers < 1 || nbuffers > 8)
    PGBP_FAIL(PGBP_EINVAL, "bad arguments");
  PGBP_TRY(set_device(device));
  std::unique_ptr<pgbp_comm> c(new pgbp_comm);
  c->device = device; c->rank = rank; c->nranks = nranks; c->ld = ld; c->nbuffers = nbuffers;
  c->bytes = comm_data_bytes(c.get()) + sizeof(unsigned long long) * (size_t)nbuffers * nranks;
  void* v = nullptr;
  PGBP_TRY(dev_malloc(&v, c->bytes));
  c->local = (char*)v;
  c->peer[rank] = c->local;
  int rc = dev_memset(c->local, 0, c->bytes, 0);
  if (!rc) rc = dev_malloc(&v, sizeof(int32_t));
  if (!rc) { c->d_err = (int32_t*)v; rc = dev_memset(c->d_err, 0, sizeof(int32_t), 0); }
  if (!rc) rc = stream_sync(0);
  if (rc) { dev_free(c->local); dev_free(c->d_err); return rc; }
  *out = c.release();
  return 0;
}

// 64-byte CUDA IPC handle of this rank's window; the host side exchanges the handles of all ranks
// (torch.distributed / MPI all-gather of 64 bytes) and passes them to pgbp_comm_connect in rank order
int32_t pgbp_comm_handle(pgbp_comm* c, uint8_t* handle64) {
  if (!c || !handle64) PGBP_FAIL(PGBP_EINVAL, "null argument");
#ifdef PGBP_HOST_EMUL
  memset(handle64, 0, 64);
  memcpy(handle64, &c->local, sizeof(char*));  // same address space: the "handle" is the pointer
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  PGBP_TRY(set_device(c->device));
  cudaIpcMemHandle_t h;
  PGBP_CUDA(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle64, &h, 64);
#endif
  return 0;
}

int32_t pgbp_comm_connect(pgbp_comm* c, const uint8_t* handles) {
  if (!c || !handles) PGBP_FAIL(PGBP_EINVAL, "null argument");
  PGBP_TRY(set_device(c->device));
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank || c->peer[r]) continue;
#ifdef PGBP_HOST_EMUL
    memcpy(&c->peer[r], handles + 64 * (size_t)r, sizeof(char*));
#else
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    PGBP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = (char*)p;
    c->opened[r] = true;
#endif
  }
  return 0;
}

// Unmap the peers' windows (after this, puts / gathers fail until pgbp_comm_connect is called again).  Multi-process
// teardown order: every rank disconnects, the ranks synchronise (barrier), then every rank destroys -- an exported
// allocation must not be freed while another process still maps it.
int32_t pgbp_comm_disconnect(pgbp_comm* c) {
  if (!c) PGBP_FAIL(PGBP_EINVAL, "null argument");
  PGBP_TRY(set_device(c->device));
#ifndef PGBP_HOST_EMUL
  PGBP_CUDA(cudaDeviceSynchronize());
#endif
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank) continue;
#ifndef PGBP_HOST_EMUL
    if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
#endif
    c->opened[r] = false;
    c->peer[r] = nullptr;
  }
  return 0;
}

int32_t pgbp_comm_destroy(pgbp_comm* c) {
  if (!c) return 0;
  pgbp_comm_disconnect(c);
  dev_free(c->local);
  dev_free(c->d_err);
  delete c;
  return 0;
}

int32_t pgbp_comm_window(pgbp_comm* c, int32_t buffer, double** d_ptr, int64_t* ld) {
  if (!c || buffer < 0 || buffer >= c->nbuffers) PGBP_FAIL(PGBP_EINVAL, "bad arguments");
  if (d_ptr) *d_ptr = comm_slot(c, c->local, buffer, 0);
  if (ld) *ld = c->ld;
  return 0;
}

static int comm_ready(const pgbp_comm* c, const pgbp_batch* b, int32_t buffer) {
  if (!c || !b || buffer < 0 || buffer >= c->nbuffers) PGBP_FAIL(PGBP_EINVAL, "bad arguments");
  if (c->device != b->device) PGBP_FAIL(PGBP_EINVAL, "window and batch live on different devices");
  if (c->ld < b->B) PGBP_FAIL(PGBP_EINVAL, "window rows (%lld) shorter than the batch (%lld)", (long long)c->ld, (long long)b->B);
  for (int r = 0; r < c->nranks; r++)
    if (!c->peer[r]) PGBP_FAIL(PGBP_ESTATE, "window of rank %d not connected (pgbp_comm_connect)", r);
  return 0;
}

static int comm_signal(pgbp_comm* c, pgbp_batch* b, int32_t buffer, const PeerWin& w) {
  const unsigned long long seq = ++c->seq[buffer];
#ifdef PGBP_HOST_EMUL
  for (int r = 0; r < c->nranks; r++) *w.flag[r] = seq;
#else
  k_comm_signal<<<1, 32, 0, b->stream>>>(w, c->nranks, seq);
#endif
  b->launches++;
  return check_launch("k_comm_signal");
}

int32_t pgbp_integrate_gather(pgbp_batch* b, int32_t belief, pgbp_comm* c, int32_t buffer) {
  PGBP_TRY(comm_ready(c, b, buffer));
  const pgbp_plan* p = b->plan;
  if (belief < 0 || belief >= p->nbeliefs) PGBP_FAIL(PGBP_EINVAL, "belief index out of range");
  PGBP_TRY(set_device(b->device));
  if (belief >= p->nclusters) PGBP_TRY(batch_materialize_sepsets(b));
  const int M = p->dim[belief];
  const int64_t js = p->jslot[belief], hs = batch_hrow(b, belief), gs = batch_grow(b, belief);
  const JSide jside = batch_jside(b);
  const PeerWin w = peer_win(c, buffer);
#ifdef PGBP_HOST_EMUL
  for (int64_t e = 0; e < b->B; e++) {
    integrate_thread<PGBP_MAX_DIM>(b->state, b->status, b->ld, e, js, hs, gs, M, nullptr, w.data[c->rank], b->ld, nullptr, jcolumn(jside, e), jside.ld);
    for (int r = 0; r < c->nranks; r++) w.data[r][e] = w.data[c->rank][e];
  }
#else
  const unsigned grid = (unsigned)((b->B + 127) / 128);
#define PGBP_IG(MAXM) k_integrate_gather<MAXM><<<grid, 128, 0, b->stream>>>(b->state, b->status, b->B, b->ld, js, hs, gs, M, jside, w, c->rank, c->nranks)
  if (M <= 4) PGBP_IG(4);
  else if (M <= 12) PGBP_IG(12);
  else if (M <= 32) PGBP_IG(32);
  else PGBP_IG(PGBP_MAX_DIM);
#undef PGBP_IG
#endif
  b->launches++;
  PGBP_TRY(check_launch("k_integrate_gather"));
  return comm_signal(c, b, buffer, w);
}

int32_t pgbp_comm_put(pgbp_comm* c, pgbp_batch* b, int32_t buffer, const double* d_src) {
  PGBP_TRY(comm_ready(c, b, buffer));
  if (!d_src) PGBP_FAIL(PGBP_EINVAL, "null source");
  PGBP_TRY(set_device(b->device));
  const PeerWin w = peer_win(c, buffer);
#ifdef PGBP_HOST_EMUL
  for (int r = 0; r < c->nranks; r++)
    for (int64_t e = 0; e < b->B; e++) w.data[r][e] = d_src[e];
#else
  k_comm_put<<<(unsigned)((b->B + 255) / 256), 256, 0, b->stream>>>(d_src, b->B, w, c->nranks);
#endif
  b->launches++;
  PGBP_TRY(check_launch("k_comm_put"));
  return comm_signal(c, b, buffer, w);
}

// Enqueue (on the batch's stream) a wait until every rank's put number seq[buffer] of THIS rank's count has
// landed in the local window; SPMD callers put the same number of times per buffer.  timeout_ms bounds the spin
// (a lost peer): the next pgbp_comm_check then reports which rank was missing.
int32_t pgbp_comm_wait(pgbp_comm* c, pgbp_batch* b, int32_t buffer, int32_t timeout_ms) {
  PGBP_TRY(comm_ready(c, b, buffer));
  PGBP_TRY(set_device(b->device));
#ifdef PGBP_HOST_EMUL
  (void)timeout_ms;
  for (int r = 0; r < c->nranks; r++)
    if (*comm_flag(c, c->local, buffer, r) < c->seq[buffer]) PGBP_FAIL(PGBP_ESTATE, "rank %d has not put buffer %d", r, buffer);
#else
  const long long cycles = (long long)(timeout_ms > 0 ? timeout_ms : 2000) * 2000000LL;  // ~2 GHz
  k_comm_wait<<<1, 32, 0, b->stream>>>(comm_flag(c, c->local, buffer, 0), c->nranks, c->seq[buffer], cycles, c->d_err);
#endif
  b->launches++;
  return check_launch("k_comm_wait");
}

// host copy of this rank's window of `buffer` ([nranks][ld] doubles), synchronous on the batch's stream
int32_t pgbp_comm_read(pgbp_comm* c, pgbp_batch* b, int32_t buffer, double* host) {
  if (!c || !b || !host || buffer < 0 || buffer >= c->nbuffers) PGBP_FAIL(PGBP_EINVAL, "bad arguments");
  PGBP_TRY(set_device(b->device));
  PGBP_TRY(d2h(host, comm_slot(c, c->local, buffer, 0), sizeof(double) * (size_t)c->nranks * (size_t)c->ld, b->stream));
  return stream_sync(b->stream);
}

// synchronises the batch's stream; error if a wait timed out (names the rank)
int32_t pgbp_comm_check(pgbp_comm* c, pgbp_batch* b) {
  if (!c || !b) PGBP_FAIL(PGBP_EINVAL, "null argument");
  PGBP_TRY(set_device(b->device));
  int32_t err = 0;
  PGBP_TRY(d2h(&err, c->d_err, sizeof err, b->stream));
  PGBP_TRY(stream_sync(b->stream));
  if (err) PGBP_FAIL(PGBP_ESTATE, "pgbp_comm_wait timed out waiting for rank %d", err - 1);
  return 0;
}

}  // extern "C"
