// Multi-GPU C++ caller: quadrature points sharded over R GPUs, partial coefficient blocks combined by NCCL through the
// C-ABI (include/pspace_cuda.h section (e)) — no Python, no torch.
//
// The reference's parallel drivers are MPI programs (examples/stacs/smd/projection.cpp:36-39: MPI_Init, MPI_Comm_rank) that
// never distribute the quadrature loop; here every rank owns a q-range.  MPI is not in this image, so the R "ranks" are R
// host threads of one process with one communicator each from ncclCommInitAll — under MPI the only difference is
// ncclCommInitRank with an id broadcast by rank 0.  Each rank:
//     ParameterContainer pc; pc.setDevice(r); pc.initialize();                     same container on every rank
//     pspace_cuda_shard_bounds(Q, R, r, &q0, &q1)                                   its quadrature-point range
//     evaluate the deterministic blocks J_q, f_q for q in [q0, q1) only             (the caller's element)
//     pspace_cuda_project_matrix_allreduce / _vector_allreduce / _matrix_reduce_scatter
// `--check` compares every rank's result with the single-GPU projection of all points (<= 1e-12, normwise).
//
//   make -C examples sharded_projection && ./examples/sharded_projection --nranks 2 --check
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ParameterContainer.h"
#include "ParameterFactory.h"
#include "pspace_cuda.h"

namespace {
const int M = 4, M2 = M * M;

// synthetic deterministic element: a 4x4 block and a 4-vector, smooth in the parameter values y
void element(const scalar *y, int nv, scalar *A, scalar *f) {
  for (int r = 0; r < M; r++) {
    f[r] = 0.5 * r - y[r % nv] * y[(r + 1) % nv];
    for (int c = 0; c < M; c++) A[r * M + c] = (r == c ? 2.0 + y[0] : 0.0) + 0.1 * (r + 1) * y[c % nv] - 0.05 * c * y[(r + c) % nv] * y[1 % nv];
  }
}

ParameterContainer *make_container(ParameterFactory *factory, int device) {
  ParameterContainer *pc = new ParameterContainer(1 /* complete basis */);
  pc->addParameter(factory->createNormalParameter(1.0, 0.3, 3));
  pc->addParameter(factory->createUniformParameter(-1.0, 2.0, 3));
  pc->addParameter(factory->createExponentialParameter(0.5, 0.7, 2));
  pc->addParameter(factory->createNormalParameter(-0.5, 0.2, 2));
  pc->setDevice(device);
  pc->initialize();
  return pc;
}

#define MUST(call)                                                                                \
  do {                                                                                            \
    int rc__ = (call);                                                                            \
    if (rc__ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, pspace_cuda_last_error()); exit(2); } \
  } while (0)

double maxabs(const std::vector<scalar> &v) { double m = 0; for (size_t i = 0; i < v.size(); i++) m = std::max(m, (double)std::abs(v[i])); return m; }
double maxdiff(const scalar *a, const scalar *b, size_t n) { double m = 0; for (size_t i = 0; i < n; i++) m = std::max(m, (double)std::abs(a[i] - b[i])); return m; }

struct RankResult { double err_c = 0, err_f = 0, err_rows = 0; long long q0 = 0, q1 = 0; int r0 = 0, r1 = 0; };

void rank_main(int rank, int world, ncclComm_t comm, const std::vector<scalar> *Cref, const std::vector<scalar> *Fref, RankResult *out) {
  ParameterFactory factory;
  ParameterContainer *pc = make_container(&factory, rank);
  pspace_ctx *ctx = pc->deviceContext();
  const int N = pc->getNumBasisTerms(), nv = pc->getNumParameters();
  const long long Q = pc->getNumQuadraturePoints();
  long long q0, q1;
  MUST(pspace_cuda_shard_bounds(Q, world, rank, &q0, &q1));
  const long long nq = q1 - q0;
  // the caller's deterministic element, only at this rank's points
  std::vector<scalar> J((size_t)std::max(1LL, nq) * M2), f((size_t)std::max(1LL, nq) * M), zq(nv), yq(nv);
  for (long long q = q0; q < q1; q++) {
    pc->quadrature((int)q, zq.data(), yq.data());
    element(yq.data(), nv, &J[(size_t)(q - q0) * M2], &f[(size_t)(q - q0) * M]);
  }
  const size_t w = sizeof(scalar);
  int per = 0, r0 = 0, r1 = 0;
  MUST(pspace_cuda_row_block(N, world, rank, &r0, &r1, &per));
  void *dJ = 0, *df = 0, *dC = 0, *dF = 0, *dCp = 0, *dCr = 0, *dws = 0;
  MUST(pspace_cuda_activate(ctx));
  MUST(pspace_cuda_malloc(&dJ, J.size() * w));
  MUST(pspace_cuda_malloc(&df, f.size() * w));
  MUST(pspace_cuda_malloc(&dC, (size_t)N * N * M2 * w));
  MUST(pspace_cuda_malloc(&dF, (size_t)N * M * w));
  MUST(pspace_cuda_malloc(&dCp, (size_t)per * world * N * M2 * w));
  MUST(pspace_cuda_malloc(&dCr, (size_t)per * N * M2 * w));
  MUST(pspace_cuda_memcpy(dJ, J.data(), J.size() * w, 1));
  MUST(pspace_cuda_memcpy(df, f.data(), f.size() * w, 1));
  pspace_project_args a;
  memset(&a, 0, sizeof a);
  a.k0 = 0; a.k1 = N; a.j0 = 0; a.j1 = N; a.q0 = q0; a.q1 = q1; a.ncols = M2;
  pspace_project_args av = a;
  av.ncols = M;
  const size_t ws = std::max(pspace_cuda_project_workspace(ctx, 1, &a), pspace_cuda_project_workspace(ctx, 0, &av));
  MUST(pspace_cuda_malloc(&dws, ws));
  MUST(pspace_cuda_project_matrix_allreduce(ctx, &a, (const double *)dJ, (double *)dC, dws, ws, comm, 0));
  MUST(pspace_cuda_project_vector_allreduce(ctx, &av, (const double *)df, (double *)dF, dws, ws, comm, 0));
  MUST(pspace_cuda_project_matrix_reduce_scatter(ctx, &a, (const double *)dJ, (double *)dCp, (double *)dCr, dws, ws, comm, 0));
  std::vector<scalar> C((size_t)N * N * M2), F((size_t)N * M), Crows((size_t)per * N * M2);
  MUST(pspace_cuda_memcpy(C.data(), dC, C.size() * w, 0));          // blocking copies: ordered after the stream-0 work
  MUST(pspace_cuda_memcpy(F.data(), dF, F.size() * w, 0));
  MUST(pspace_cuda_memcpy(Crows.data(), dCr, Crows.size() * w, 0));
  out->q0 = q0; out->q1 = q1; out->r0 = r0; out->r1 = r1;
  if (Cref) {
    out->err_c = maxdiff(C.data(), Cref->data(), C.size()) / maxabs(*Cref);
    out->err_f = maxdiff(F.data(), Fref->data(), F.size()) / maxabs(*Fref);
    out->err_rows = maxdiff(Crows.data(), Cref->data() + (size_t)r0 * N * M2, (size_t)(r1 - r0) * N * M2) / maxabs(*Cref);
  }
  pspace_cuda_free(dJ); pspace_cuda_free(df); pspace_cuda_free(dC); pspace_cuda_free(dF);
  pspace_cuda_free(dCp); pspace_cuda_free(dCr); pspace_cuda_free(dws);
  delete pc;
}
}  // namespace

int main(int argc, char **argv) {
  int world = 0;
  bool check = false;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--check")) check = true;
    else if (!strcmp(argv[i], "--nranks") && i + 1 < argc) world = atoi(argv[++i]);
  }
  const int ndev = pspace_cuda_device_count();
  if (ndev < 1) { fprintf(stderr, "sharded_projection: no CUDA device (%s): the pspace GPU path has no CPU fallback\n", pspace_cuda_last_error()); return 2; }
  if (world <= 0) world = ndev;
  if (world > ndev) { fprintf(stderr, "sharded_projection: %d ranks asked, %d devices visible\n", world, ndev); return 2; }

  // single-GPU projection of all points: what every rank must end up with
  std::vector<scalar> Cref, Fref;
  int N = 0;
  long long Q = 0;
  {
    ParameterFactory factory;
    ParameterContainer *pc = make_container(&factory, 0);
    N = pc->getNumBasisTerms(); Q = pc->getNumQuadraturePoints();
    const int nv = pc->getNumParameters();
    std::vector<scalar> J((size_t)Q * M2), f((size_t)Q * M), zq(nv), yq(nv);
    for (long long q = 0; q < Q; q++) {
      pc->quadrature((int)q, zq.data(), yq.data());
      element(yq.data(), nv, &J[(size_t)q * M2], &f[(size_t)q * M]);
    }
    Cref.resize((size_t)N * N * M2); Fref.resize((size_t)N * M);
    pc->projectMatrix(M2, J.data(), Cref.data());
    pc->projectVector(M, f.data(), Fref.data());
    delete pc;
  }
  printf("sharded projection: N = %d basis terms, Q = %lld quadrature points, %dx%d blocks, %d rank(s)\n", N, Q, M, M, world);

  std::vector<int> devs(world);
  for (int r = 0; r < world; r++) devs[r] = r;
  std::vector<ncclComm_t> comms(world);
  ncclResult_t nr = ncclCommInitAll(comms.data(), world, devs.data());
  if (nr != ncclSuccess) { fprintf(stderr, "ncclCommInitAll: %s\n", ncclGetErrorString(nr)); return 2; }

  std::vector<RankResult> res(world);
  std::vector<std::thread> th;
  for (int r = 0; r < world; r++) th.emplace_back(rank_main, r, world, comms[r], &Cref, &Fref, &res[r]);
  for (int r = 0; r < world; r++) th[r].join();
  int bad = 0;
  for (int r = 0; r < world; r++) {
    printf("rank %d: q in [%lld, %lld)  allreduce C err %.2e  F err %.2e | reduce-scatter rows [%d, %d) err %.2e\n", r, res[r].q0, res[r].q1,
           res[r].err_c, res[r].err_f, res[r].r0, res[r].r1, res[r].err_rows);
    if (!(res[r].err_c < 1e-12 && res[r].err_f < 1e-12 && res[r].err_rows < 1e-12)) bad = 1;
  }
  for (int r = 0; r < world; r++) ncclCommDestroy(comms[r]);
  if (check) printf(bad ? "FAILED\n" : "PASSED (<= 1e-12 on every rank)\n");
  return check ? bad : 0;
}
