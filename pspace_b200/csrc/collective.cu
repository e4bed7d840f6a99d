// (e) Multi-GPU exchange step behind the C-ABI: q-sharded projection + ONE NCCL collective over the coefficient blocks.
//
// Quadrature points shard across ranks: rank r projects its contiguous q-range into a full partial
// window; the partial windows are summed by ncclAllReduce (every rank ends with the total) or ncclReduceScatter over
// basis-row blocks (rank r ends with its block of rows).  A C++ caller shaped like the reference's MPI drivers
// (examples/stacs/smd/projection.cpp:36-39: MPI_Init / MPI_Comm_rank) creates the communicator itself
// (ncclCommInitRank under MPI, or ncclCommInitAll in one process) and hands it over as an opaque pointer.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy already loaded into the process — torch's bundled one,
// or the system library — so the communicator and the calls come from the same library); libpspace_cuda.so itself has no
// link-time NCCL dependency and single-GPU users never touch it.  The handful of NCCL declarations below are the stable
// public ABI of nccl.h (ncclFloat64 = 8, ncclSum = 0).
#include <dlfcn.h>
#include <algorithm>
#include <mutex>
#include "ctx.h"

namespace {
typedef int nccl_result_t;
typedef void *nccl_comm_t;
typedef nccl_result_t (*AllReduceFn)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef nccl_result_t (*ReduceScatterFn)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char *(*ErrStrFn)(nccl_result_t);
typedef nccl_result_t (*CommCountFn)(nccl_comm_t, int *);
typedef nccl_result_t (*CommRankFn)(nccl_comm_t, int *);
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct Nccl {
  void *handle = nullptr;
  AllReduceFn all_reduce = nullptr;
  ReduceScatterFn reduce_scatter = nullptr;
  ErrStrFn err = nullptr;
  CommCountFn count = nullptr;
  CommRankFn rank = nullptr;
  bool ok = false;
};

const Nccl &nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (!n.handle) return;
    n.all_reduce = (AllReduceFn)dlsym(n.handle, "ncclAllReduce");
    n.reduce_scatter = (ReduceScatterFn)dlsym(n.handle, "ncclReduceScatter");
    n.err = (ErrStrFn)dlsym(n.handle, "ncclGetErrorString");
    n.count = (CommCountFn)dlsym(n.handle, "ncclCommCount");
    n.rank = (CommRankFn)dlsym(n.handle, "ncclCommUserRank");
    n.ok = n.all_reduce && n.reduce_scatter && n.err && n.count && n.rank;
  });
  return n;
}

int need_nccl() {
  if (!nccl().ok) return ps_fail(PSPACE_ECUDA, "NCCL is not available (dlopen libnccl.so.2 failed: %s)", dlerror() ? dlerror() : "symbols missing");
  return PSPACE_OK;
}
#define PS_NCCL(call)                                                                              \
  do {                                                                                             \
    nccl_result_t r__ = (call);                                                                    \
    if (r__ != 0) return ps_fail(PSPACE_ECUDA, "%s failed: %s", #call, nccl().err(r__));           \
  } while (0)

int comm_geometry(void *comm, int *world, int *rank) {
  int rc = need_nccl();
  if (rc) return rc;
  if (!comm) return ps_fail(PSPACE_EINVAL, "null NCCL communicator");
  PS_NCCL(nccl().count(comm, world));
  PS_NCCL(nccl().rank(comm, rank));
  return PSPACE_OK;
}
}  // namespace

extern "C" {

// Contiguous, near-equal q-ranges with boundaries on multiples of 32 points (the last range takes the remainder);
// the same rule as pspace_b200/sharded.py:shard_bounds, so C++ and Python callers shard identically.
int pspace_cuda_shard_bounds(long long Q, int world, int rank, long long *q0, long long *q1) {
  if (!q0 || !q1 || world < 1 || rank < 0 || rank >= world || Q < 0) return ps_fail(PSPACE_EINVAL, "shard_bounds(%lld, %d, %d)", Q, world, rank);
  const long long align = 32, units = (Q + align - 1) / align;
  long long start = 0;
  for (int r = 0; r <= rank; r++) {
    const long long n_units = units / world + (r < units % world ? 1 : 0);
    const long long end = std::min(Q, start + n_units * align);
    if (r == rank) { *q0 = start; *q1 = end; }
    start = end;
  }
  return PSPACE_OK;
}

// equal row blocks for the reduce-scatter: rank r owns rows [r*per, min(rows, (r+1)*per)), per = ceil(rows / world)
int pspace_cuda_row_block(int rows, int world, int rank, int *r0, int *r1, int *rows_per_rank) {
  if (world < 1 || rank < 0 || rank >= world || rows < 0) return ps_fail(PSPACE_EINVAL, "row_block(%d, %d, %d)", rows, world, rank);
  const int per = (rows + world - 1) / world;
  if (r0) *r0 = std::min(rows, rank * per);
  if (r1) *r1 = std::min(rows, (rank + 1) * per);
  if (rows_per_rank) *rows_per_rank = per;
  return PSPACE_OK;
}

int pspace_cuda_nccl_allreduce(double *d_buf, size_t count, void *nccl_comm, void *stream) {
  int rc = need_nccl();
  if (rc) return rc;
  if (!nccl_comm || (!d_buf && count)) return ps_fail(PSPACE_EINVAL, "null argument");
  if (count == 0) return PSPACE_OK;
  PS_NCCL(nccl().all_reduce(d_buf, d_buf, count, NCCL_FLOAT64, NCCL_SUM, nccl_comm, (cudaStream_t)stream));
  return PSPACE_OK;
}

int pspace_cuda_project_matrix_allreduce(pspace_ctx *c, const pspace_project_args *a, const double *d_J, double *d_C,
                                         void *d_ws, size_t ws_bytes, void *nccl_comm, void *stream) {
  if (!c || !a) return ps_fail(PSPACE_EINVAL, "null argument");
  int rc = need_nccl();
  if (rc) return rc;
  if ((rc = pspace_cuda_project_matrix(c, a, d_J, d_C, d_ws, ws_bytes, stream))) return rc;
  const size_t n = (size_t)std::max(0, a->k1 - a->k0) * (size_t)std::max(0, a->j1 - a->j0) * a->ncols * c->width;
  return pspace_cuda_nccl_allreduce(d_C, n, nccl_comm, stream);
}

int pspace_cuda_project_vector_allreduce(pspace_ctx *c, const pspace_project_args *a, const double *d_f, double *d_F,
                                         void *d_ws, size_t ws_bytes, void *nccl_comm, void *stream) {
  if (!c || !a) return ps_fail(PSPACE_EINVAL, "null argument");
  int rc = need_nccl();
  if (rc) return rc;
  if ((rc = pspace_cuda_project_vector(c, a, d_f, d_F, d_ws, ws_bytes, stream))) return rc;
  const size_t n = (size_t)std::max(0, a->k1 - a->k0) * a->ncols * c->width;
  return pspace_cuda_nccl_allreduce(d_F, n, nccl_comm, stream);
}

// d_C_partial: this rank's partial window, [world * rows_per_rank][(j1-j0)][ncols] scalars (the rows past i1-i0 are
// padding, zeroed here); d_C_rows: [rows_per_rank][(j1-j0)][ncols] scalars, the rank's row block of the TOTAL.
int pspace_cuda_project_matrix_reduce_scatter(pspace_ctx *c, const pspace_project_args *a, const double *d_J,
                                              double *d_C_partial, double *d_C_rows, void *d_ws, size_t ws_bytes,
                                              void *nccl_comm, void *stream) {
  if (!c || !a || !d_C_partial || !d_C_rows) return ps_fail(PSPACE_EINVAL, "null argument");
  int world = 1, rank = 0;
  int rc = comm_geometry(nccl_comm, &world, &rank);
  if (rc) return rc;
  if ((rc = pspace_cuda_project_matrix(c, a, d_J, d_C_partial, d_ws, ws_bytes, stream))) return rc;
  const int rows = std::max(0, a->k1 - a->k0), per = (rows + world - 1) / world;
  const size_t rowlen = (size_t)std::max(0, a->j1 - a->j0) * a->ncols * c->width;
  if (per == 0 || rowlen == 0) return PSPACE_OK;
  const size_t pad_rows = (size_t)per * world - rows;
  if (pad_rows) PS_CUDA(cudaMemsetAsync(d_C_partial + (size_t)rows * rowlen, 0, pad_rows * rowlen * sizeof(double), (cudaStream_t)stream));
  PS_NCCL(nccl().reduce_scatter(d_C_partial, d_C_rows, (size_t)per * rowlen, NCCL_FLOAT64, NCCL_SUM, nccl_comm, (cudaStream_t)stream));
  return PSPACE_OK;
}

}  // extern "C"
