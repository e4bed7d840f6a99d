// C-ABI layer (include/pspace_cuda.h): context management, argument checking, error reporting and the
// host-buffer entry points.  All numerics live in basis.cu / tables.cu / project.cu.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "ctx.h"

static thread_local char g_err[512] = "";

int ps_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

extern "C" {

const char *pspace_cuda_last_error(void) { return g_err; }
int pspace_cuda_version(void) { return 100; }

int pspace_cuda_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return ps_fail(PSPACE_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

int pspace_cuda_create(pspace_ctx **out, int device, int is_complex, int nvars, const int *kinds, const double *h_p1,
                       const double *h_p2, const int *dmax, int basis_type) {
  if (!out || !kinds || !h_p1 || !h_p2 || !dmax) return ps_fail(PSPACE_EINVAL, "null argument");
  if (nvars < 1 || nvars > PSPACE_MAX_VARS) return ps_fail(PSPACE_EINVAL, "nvars %d outside 1..%d", nvars, PSPACE_MAX_VARS);
  int ndev = pspace_cuda_device_count();
  if (ndev < 0) return ndev;
  if (ndev == 0) return ps_fail(PSPACE_ECUDA, "no CUDA device: the pspace GPU path has no CPU fallback");
  if (device < 0 || device >= ndev) return ps_fail(PSPACE_EINVAL, "device %d outside 0..%d", device, ndev - 1);
  PS_CUDA(cudaSetDevice(device));
  {
    // scratch of the small kernels comes from the stream-ordered allocator: keep freed blocks in the pool instead of
    // returning them to the driver at every synchronisation (a cudaMallocAsync would otherwise cost milliseconds)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  pspace_ctx *c = new pspace_ctx();
  c->device = device;
  c->is_complex = is_complex ? 1 : 0;
  c->width = is_complex ? 2 : 1;
  c->nvars = nvars;
  c->basis_type = basis_type;
  for (int d = 0; d < nvars; d++) {
    if (kinds[d] < 0 || kinds[d] > 2) { delete c; return ps_fail(PSPACE_EINVAL, "kinds[%d]=%d", d, kinds[d]); }
    c->kinds[d] = kinds[d];
    c->dmax[d] = dmax[d];
    for (int w = 0; w < c->width; w++) {
      c->p1[c->width * d + w] = h_p1[c->width * d + w];
      c->p2[c->width * d + w] = h_p2[c->width * d + w];
    }
  }
  *out = c;
  return PSPACE_OK;
}

int pspace_cuda_destroy(pspace_ctx *c) {
  if (!c) return PSPACE_OK;
  cudaSetDevice(c->device);
  cudaFree(c->d_deg); cudaFree(c->d_zp); cudaFree(c->d_yp); cudaFree(c->d_wp); cudaFree(c->d_T); cudaFree(c->d_Tw);
  cudaFree(c->d_qdig); cudaFree(c->d_W);
  cudaDeviceSynchronize();
  cudaFree(c->panel_u.d); cudaFree(c->panel_w.d);
  if (c->panel_u.ev) cudaEventDestroy(c->panel_u.ev);
  if (c->panel_w.ev) cudaEventDestroy(c->panel_w.ev);
  cudaFree(c->d_scratch_in); cudaFree(c->d_scratch_out); cudaFree(c->d_scratch_ws);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) {
    cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(c->ev_copy[i]); cudaEventDestroy(c->ev_comp[i]); }
  }
  delete c;
  return PSPACE_OK;
}

int pspace_cuda_basis_count(int nvars, const int *pmax, int basis_type, long long *nbasis) {
  if (!pmax || !nbasis) return ps_fail(PSPACE_EINVAL, "null argument");
  return ps_basis_count_host(nvars, pmax, basis_type, nbasis);
}

int pspace_cuda_basis_degrees(int nvars, const int *pmax, int basis_type, int *nbasis, int *d_out, long long capacity_rows,
                              void *stream) {
  if (!pmax || !nbasis || !d_out) return ps_fail(PSPACE_EINVAL, "null argument");
  return ps_basis_degrees_device(nvars, pmax, basis_type, nbasis, d_out, capacity_rows, (cudaStream_t)stream, nullptr);
}

int pspace_cuda_initialize_basis(pspace_ctx *c, const int *pmax, void *stream) {
  if (!c || !pmax) return ps_fail(PSPACE_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  PS_CUDA(cudaSetDevice(c->device));
  long long n = 0;
  int rc = ps_basis_count_host(c->nvars, pmax, c->basis_type, &n);
  if (rc) return rc;
  if (n > 0x7fffffffLL / c->nvars) return ps_fail(PSPACE_EINVAL, "basis of %lld terms is too large", n);
  if (c->d_deg) { PS_CUDA(cudaFree(c->d_deg)); c->d_deg = nullptr; }
  PS_CUDA(cudaMalloc(&c->d_deg, sizeof(int) * (size_t)n * c->nvars));
  int nb = 0;
  rc = ps_basis_degrees_device(c->nvars, pmax, c->basis_type, &nb, c->d_deg, n, st, &c->launches);
  if (rc) return rc;
  if (nb != n) return ps_fail(PSPACE_ECUDA, "internal: device produced %d basis terms, expected %lld", nb, n);
  c->N = nb;
  for (int d = 0; d < c->nvars; d++) c->pmax[d] = pmax[d];
  c->h_deg.resize((size_t)nb * c->nvars);
  PS_CUDA(cudaMemcpyAsync(c->h_deg.data(), c->d_deg, sizeof(int) * c->h_deg.size(), cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  c->basis_ready = true;
  c->invalidate_panels();
  rc = ps_build_tables(c, st);
  if (rc) return rc;
  // complete on return, like initialize_quadrature: the host getters copy on the legacy stream, which does not order
  // against non-blocking streams
  PS_CUDA(cudaStreamSynchronize(st));
  return PSPACE_OK;
}

int pspace_cuda_initialize_quadrature(pspace_ctx *c, const int *nqpts, void *stream) {
  if (!c || !nqpts) return ps_fail(PSPACE_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  PS_CUDA(cudaSetDevice(c->device));
  long long Q = 1;
  for (int d = 0; d < c->nvars; d++) {
    if (nqpts[d] < 1 || nqpts[d] > PSPACE_MAX_NQ) return ps_fail(PSPACE_EINVAL, "nqpts[%d]=%d outside 1..%d", d, nqpts[d], PSPACE_MAX_NQ);
    if (Q > (1LL << 40) / nqpts[d]) return ps_fail(PSPACE_EINVAL, "quadrature grid exceeds 2^40 points");
    Q *= nqpts[d];
  }
  for (int d = 0; d < c->nvars; d++) c->nq[d] = nqpts[d];
  c->Q = Q;
  c->quad_ready = true;
  c->invalidate_panels();
  int rc = ps_build_tables(c, st);
  if (rc) return rc;
  rc = ps_build_grid(c, st);
  if (rc) return rc;
  PS_CUDA(cudaStreamSynchronize(st));
  return PSPACE_OK;
}

int pspace_cuda_initialize(pspace_ctx *c, void *stream) {
  if (!c) return ps_fail(PSPACE_EINVAL, "null argument");
  int pm[PSPACE_MAX_VARS], nq[PSPACE_MAX_VARS];
  for (int d = 0; d < c->nvars; d++) { pm[d] = c->dmax[d]; nq[d] = c->dmax[d] + 1; }
  int rc = pspace_cuda_initialize_basis(c, pm, stream);
  if (rc) return rc;
  return pspace_cuda_initialize_quadrature(c, nq, stream);
}

int pspace_cuda_num_basis(const pspace_ctx *c) { return c ? c->N : 0; }
int pspace_cuda_num_params(const pspace_ctx *c) { return c ? c->nvars : 0; }
long long pspace_cuda_num_quad(const pspace_ctx *c) { return c ? c->Q : 0; }
int pspace_cuda_is_complex(const pspace_ctx *c) { return c ? c->is_complex : 0; }

int pspace_cuda_get_degrees(const pspace_ctx *c, int *h_out) {
  if (!c || !h_out) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->basis_ready) return ps_fail(PSPACE_ESTATE, "basis not initialised");
  memcpy(h_out, c->h_deg.data(), sizeof(int) * c->h_deg.size());
  return PSPACE_OK;
}

int pspace_cuda_get_rules(const pspace_ctx *c, int pid, double *h_z, double *h_y, double *h_w) {
  if (!c) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->quad_ready) return ps_fail(PSPACE_ESTATE, "quadrature not initialised");
  if (pid < 0 || pid >= c->nvars) return ps_fail(PSPACE_EINVAL, "parameter id %d", pid);
  PS_CUDA(cudaSetDevice(c->device));
  size_t off = (size_t)c->roff[pid] * c->width, bytes = sizeof(double) * c->width * c->nq[pid];
  if (h_z) PS_CUDA(cudaMemcpy(h_z, c->d_zp + off, bytes, cudaMemcpyDeviceToHost));
  if (h_y) PS_CUDA(cudaMemcpy(h_y, c->d_yp + off, bytes, cudaMemcpyDeviceToHost));
  if (h_w) PS_CUDA(cudaMemcpy(h_w, c->d_wp + off, bytes, cudaMemcpyDeviceToHost));
  return PSPACE_OK;
}

int pspace_cuda_get_table(const pspace_ctx *c, int pid, double *h_T) {
  if (!c || !h_T) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->tables_ready) return ps_fail(PSPACE_ESTATE, "tables need both basis and quadrature");
  if (pid < 0 || pid >= c->nvars) return ps_fail(PSPACE_EINVAL, "parameter id %d", pid);
  PS_CUDA(cudaSetDevice(c->device));
  PS_CUDA(cudaMemcpy(h_T, c->d_T + (size_t)c->toff[pid] * c->width,
                     sizeof(double) * c->width * (c->pmax[pid] + 1) * c->nq[pid], cudaMemcpyDeviceToHost));
  return PSPACE_OK;
}

int pspace_cuda_get_weights(const pspace_ctx *c, long long q0, long long q1, double *h_W) {
  if (!c || !h_W) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->quad_ready) return ps_fail(PSPACE_ESTATE, "quadrature not initialised");
  if (q0 < 0 || q1 > c->Q || q1 < q0) return ps_fail(PSPACE_EINVAL, "q range");
  PS_CUDA(cudaSetDevice(c->device));
  PS_CUDA(cudaMemcpy(h_W, c->d_W + (size_t)q0 * c->width, sizeof(double) * c->width * (size_t)(q1 - q0), cudaMemcpyDeviceToHost));
  return PSPACE_OK;
}

int pspace_cuda_tensor_grid(const pspace_ctx *c, long long q0, long long q1, double *d_Z, double *d_Y, double *d_W,
                            void *stream) {
  if (!c) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->quad_ready) return ps_fail(PSPACE_ESTATE, "quadrature not initialised");
  if (q0 < 0 || q1 > c->Q || q1 < q0) return ps_fail(PSPACE_EINVAL, "q range");
  PS_CUDA(cudaSetDevice(c->device));
  return ps_tensor_grid(c, q0, q1, d_Z, d_Y, d_W, (cudaStream_t)stream);
}

int pspace_cuda_eval_basis_grid(const pspace_ctx *c, int k0, int k1, long long q0, long long q1, double *d_psi, void *stream) {
  if (!c || !d_psi) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->tables_ready) return ps_fail(PSPACE_ESTATE, "eval_basis needs basis and quadrature");
  if (k0 < 0 || k1 > c->N || k1 < k0 || q0 < 0 || q1 > c->Q || q1 < q0) return ps_fail(PSPACE_EINVAL, "range");
  PS_CUDA(cudaSetDevice(c->device));
  return ps_eval_basis_grid(c, k0, k1, q0, q1, d_psi, (cudaStream_t)stream);
}

int pspace_cuda_eval_basis_points(const pspace_ctx *c, int k0, int k1, long long npts, const double *d_z, double *d_psi,
                                  void *stream) {
  if (!c || !d_psi || !d_z) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->basis_ready) return ps_fail(PSPACE_ESTATE, "basis not initialised");
  if (k0 < 0 || k1 > c->N || k1 < k0 || npts < 0) return ps_fail(PSPACE_EINVAL, "range");
  PS_CUDA(cudaSetDevice(c->device));
  return ps_eval_basis_points(c, k0, k1, npts, d_z, d_psi, (cudaStream_t)stream);
}

size_t pspace_cuda_project_workspace(const pspace_ctx *c, int is_matrix, const pspace_project_args *a) {
  if (!c || !a) return 0;
  return ps_project_workspace(c, is_matrix, a);
}

int pspace_cuda_project_vector(pspace_ctx *c, const pspace_project_args *a, const double *d_f, double *d_F,
                               void *d_ws, size_t ws_bytes, void *stream) {
  if (!c || !a) return ps_fail(PSPACE_EINVAL, "null argument");
  if ((!d_f && a->q1 > a->q0) || (!d_F && a->k1 > a->k0)) return ps_fail(PSPACE_EINVAL, "null buffer for a non-empty range");
  PS_CUDA(cudaSetDevice(c->device));
  return ps_project(c, 0, a, d_f, d_F, d_ws, ws_bytes, (cudaStream_t)stream);
}

int pspace_cuda_project_matrix(pspace_ctx *c, const pspace_project_args *a, const double *d_J, double *d_C,
                               void *d_ws, size_t ws_bytes, void *stream) {
  if (!c || !a) return ps_fail(PSPACE_EINVAL, "null argument");
  if ((!d_J && a->q1 > a->q0) || (!d_C && a->k1 > a->k0 && a->j1 > a->j0))
    return ps_fail(PSPACE_EINVAL, "null buffer for a non-empty range");
  PS_CUDA(cudaSetDevice(c->device));
  return ps_project(c, 1, a, d_J, d_C, d_ws, ws_bytes, (cudaStream_t)stream);
}

int pspace_cuda_evaluate_expansion(pspace_ctx *c, int k0, int k1, long long q0, long long q1, int ncols,
                                   int accumulate, const double *d_V, double *d_U, void *stream) {
  if (!c) return ps_fail(PSPACE_EINVAL, "null argument");
  if ((!d_V && k1 > k0) || (!d_U && q1 > q0)) return ps_fail(PSPACE_EINVAL, "null buffer for a non-empty range");
  PS_CUDA(cudaSetDevice(c->device));
  return ps_expand(c, k0, k1, q0, q1, ncols, accumulate, d_V, d_U, (cudaStream_t)stream);
}

int pspace_cuda_peer_layout(const pspace_ctx *c, const pspace_project_args *a, int world, long long *recv_doubles) {
  if (!c || !a || !recv_doubles) return ps_fail(PSPACE_EINVAL, "null argument");
  return ps_peer_layout(c, a, world, recv_doubles);
}
int pspace_cuda_project_matrix_peer(pspace_ctx *c, const pspace_project_args *a, const double *d_J,
                                    double *const *peer_recv, int world, int rank, void *d_ws, size_t ws_bytes, void *stream) {
  if (!c || !a || !d_J || !peer_recv) return ps_fail(PSPACE_EINVAL, "null argument");
  if (world < 1 || world > PSPACE_MAX_PEERS || rank < 0 || rank >= world) return ps_fail(PSPACE_EINVAL, "world/rank");
  PS_CUDA(cudaSetDevice(c->device));
  PeerInfo pi;
  pi.world = world; pi.rank = rank;
  for (int w = 0; w < PSPACE_MAX_PEERS; w++) pi.recv[w] = w < world ? peer_recv[w] : nullptr;
  return ps_project_ex(c, 1, a, d_J, nullptr, d_ws, ws_bytes, (cudaStream_t)stream, &pi);
}
int pspace_cuda_peer_reduce_gather(const pspace_ctx *c, const pspace_project_args *a, const double *d_recv_local,
                                   double *const *peer_out, int world, int rank, void *stream) {
  if (!c || !a || !d_recv_local || !peer_out) return ps_fail(PSPACE_EINVAL, "null argument");
  if (world < 1 || world > PSPACE_MAX_PEERS || rank < 0 || rank >= world) return ps_fail(PSPACE_EINVAL, "world/rank");
  PS_CUDA(cudaSetDevice(c->device));
  return ps_peer_reduce_gather(c, a, d_recv_local, peer_out, world, rank, (cudaStream_t)stream);
}

int pspace_cuda_project_matrix_batched(pspace_ctx *c, const pspace_project_args *a, int nelem, const double *d_J,
                                       long long j_elem_stride, double *d_C, long long c_elem_stride, void *d_ws,
                                       size_t ws_bytes, void *stream) {
  if (!c || !a || nelem < 0) return ps_fail(PSPACE_EINVAL, "bad argument");
  if (nelem > 0 && (!d_J || !d_C)) return ps_fail(PSPACE_EINVAL, "null buffer");
  PS_CUDA(cudaSetDevice(c->device));
  pspace_ctx *mc = c;
  const int saved = mc->panel_cache;
  int rc = PSPACE_OK;
  for (int e = 0; e < nelem && rc == PSPACE_OK; e++) {
    if (e == 1) mc->panel_cache = 1;          // elements share the container: the basis panels are built once
    rc = ps_project(c, 1, a, d_J + (size_t)e * j_elem_stride * c->width, d_C + (size_t)e * c_elem_stride * c->width, d_ws,
                    ws_bytes, (cudaStream_t)stream);
  }
  mc->panel_cache = saved;
  if (!saved) mc->invalidate_panels();
  return rc;
}

static int ensure(void **buf, size_t *have, size_t need) {
  if (need <= *have) return PSPACE_OK;
  if (*buf) { PS_CUDA(cudaFree(*buf)); *buf = nullptr; *have = 0; }
  if (need) PS_CUDA(cudaMalloc(buf, need));
  *have = need;
  return PSPACE_OK;
}

// Host-input projection: the rows of f / J stream from host memory in slabs of quadrature points through a
// double buffer — slab s+1 is uploaded on the copy stream while slab s is projected (accumulating into d_out) on the
// compute stream — so only the first slab's upload is exposed and the device never holds more than two slabs of J.
// Pinned host memory makes the uploads truly asynchronous; pageable memory works too (staged by the driver).
static int ensure_streams(pspace_ctx *c) {
  PS_CUDA(cudaSetDevice(c->device));
  if (!c->own_stream) PS_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  if (!c->copy_stream) {
    PS_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      PS_CUDA(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
      PS_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
    }
  }
  return PSPACE_OK;
}

static size_t slab_target(const pspace_ctx *c) {
  return c->slab_bytes != 0 ? (size_t)(c->slab_bytes < 0 ? -c->slab_bytes : c->slab_bytes) : ((size_t)384 << 20);
}

static int project_hostin(pspace_ctx *c, int is_matrix, const pspace_project_args *a, const double *h_in, double *d_out) {
  int rc0 = ensure_streams(c);
  if (rc0) return rc0;
  cudaStream_t st = c->own_stream, cs = c->copy_stream;
  const size_t w = c->width;
  const long long nq = a->q1 - a->q0;
  const size_t row_bytes = sizeof(double) * w * a->ncols;
  const size_t in_bytes = row_bytes * (size_t)nq;
  const size_t target = slab_target(c);
  // slab sizes: the first upload is the only one not hidden behind a projection, so the slabs START small (1/16 of the
  // target) and double up to the target — the exposed copy shrinks 16x, the later slabs keep the kernel's stream-K waves long
  long long max_rows = (long long)((target + row_bytes - 1) / (row_bytes ? row_bytes : 1));
  max_rows = (max_rows + 63) / 64 * 64;                          // keep every 64-point stage of a slab full
  if (max_rows < 64) max_rows = 64;
  long long first_rows = c->slab_bytes < 0 ? max_rows : std::max<long long>(64, max_rows / 16 / 64 * 64);   // slab_mb < 0: equal slabs
  std::vector<long long> starts;
  {
    long long r = 0, sz = first_rows;
    while (r < nq && starts.size() < 4096) { starts.push_back(r); r += sz; sz = std::min(max_rows, sz * 2); }
  }
  const long long nslab = (long long)starts.size();
  int rc;
  if ((rc = ensure(&c->d_scratch_in, &c->scratch_in_bytes, (nslab > 1 ? 2 : 1) * row_bytes * (size_t)std::min(max_rows, std::max(nq, 1LL))))) return rc;
  pspace_project_args as = *a;
  size_t ws = 0;                                                  // workspace: the largest need over the slabs
  for (long long s = 0; s < nslab; s++) {
    as.q0 = a->q0 + starts[s];
    as.q1 = a->q0 + (s + 1 < nslab ? starts[s + 1] : nq);
    ws = std::max(ws, ps_project_workspace(c, is_matrix, &as));
  }
  if (nslab == 0) { as = *a; ws = ps_project_workspace(c, is_matrix, &as); }
  if ((rc = ensure(&c->d_scratch_ws, &c->scratch_ws_bytes, ws))) return rc;
  if (nslab == 0) return ps_project(c, is_matrix, a, nullptr, d_out, c->d_scratch_ws, ws, st);
  const size_t buf_stride = row_bytes * (size_t)std::min(max_rows, std::max(nq, 1LL));
  for (long long s = 0; s < nslab; s++) {
    const int b = (int)(s & 1);
    const long long r0 = starts[s], r1 = s + 1 < nslab ? starts[s + 1] : nq;
    double *buf = (double *)((unsigned char *)c->d_scratch_in + (size_t)b * buf_stride);
    if (s >= 2) PS_CUDA(cudaStreamWaitEvent(cs, c->ev_comp[b], 0));        // buffer b is free again
    PS_CUDA(cudaMemcpyAsync(buf, (const unsigned char *)h_in + (size_t)r0 * row_bytes, (size_t)(r1 - r0) * row_bytes,
                            cudaMemcpyHostToDevice, cs));
    PS_CUDA(cudaEventRecord(c->ev_copy[b], cs));
    PS_CUDA(cudaStreamWaitEvent(st, c->ev_copy[b], 0));
    as = *a;
    as.q0 = a->q0 + r0; as.q1 = a->q0 + r1;
    as.accumulate = (s > 0) ? 1 : a->accumulate;
    rc = ps_project(c, is_matrix, &as, buf, d_out, c->d_scratch_ws, ws, st);
    if (rc) return rc;
    PS_CUDA(cudaEventRecord(c->ev_comp[b], st));
  }
  return PSPACE_OK;
}

// Output-bound matrix projections (cfg3: 6 MB of J in, 7.7 GB of C out per element): J is uploaded once and the window is
// projected in blocks of basis rows i through two device buffers; block s is read back on the copy stream while block s+1 is
// projected, so the call costs max(projection, read-back) + one block instead of their sum, and the device holds two blocks
// of C instead of the whole window.  A row block is an ordinary window call on whole 16-row pair tiles.
static int project_host_rowblocks(pspace_ctx *c, const pspace_project_args *a, const double *h_in, double *h_out) {
  int rc = ensure_streams(c);
  if (rc) return rc;
  cudaStream_t st = c->own_stream, cs = c->copy_stream;
  const size_t w = c->width;
  const size_t in_bytes = sizeof(double) * w * a->ncols * (size_t)(a->q1 - a->q0);
  const size_t row_out = sizeof(double) * w * (size_t)(a->j1 - a->j0) * a->ncols;      // bytes of C per basis row i
  const int ni = a->k1 - a->k0;
  int blk = (int)std::max<size_t>(16, c->out_block_bytes / row_out / 16 * 16);            // whole 16-row pair tiles
  blk = std::min(blk, ni);
  if ((rc = ensure(&c->d_scratch_in, &c->scratch_in_bytes, in_bytes))) return rc;
  if ((rc = ensure(&c->d_scratch_out, &c->scratch_out_bytes, 2 * (size_t)blk * row_out))) return rc;
  pspace_project_args as = *a;
  as.k1 = a->k0 + blk;
  size_t ws = ps_project_workspace(c, 1, &as);
  if (ni % blk) { as.k0 = a->k0 + ni / blk * blk; as.k1 = a->k1; ws = std::max(ws, ps_project_workspace(c, 1, &as)); }
  if ((rc = ensure(&c->d_scratch_ws, &c->scratch_ws_bytes, ws))) return rc;
  PS_CUDA(cudaMemcpyAsync(c->d_scratch_in, h_in, in_bytes, cudaMemcpyHostToDevice, st));
  // Block s is projected BEFORE the read-back of block s-1 is issued: with pinned host memory both are asynchronous anyway;
  // with pageable memory cudaMemcpyAsync holds the host thread until the copy is staged, and this order keeps the GPU busy
  // with block s meanwhile.
  auto read_back = [&](int sb, int r0, int r1) -> int {
    const int b = sb & 1;
    double *buf = (double *)((unsigned char *)c->d_scratch_out + (size_t)b * blk * row_out);
    PS_CUDA(cudaStreamWaitEvent(cs, c->ev_comp[b], 0));
    PS_CUDA(cudaMemcpyAsync((unsigned char *)h_out + (size_t)r0 * row_out, buf, (size_t)(r1 - r0) * row_out, cudaMemcpyDeviceToHost, cs));
    PS_CUDA(cudaEventRecord(c->ev_copy[b], cs));
    return PSPACE_OK;
  };
  int s = 0, prev_r0 = 0, prev_r1 = 0;
  for (int r0 = 0; r0 < ni; r0 += blk, s++) {
    const int b = s & 1, r1 = std::min(ni, r0 + blk);
    double *buf = (double *)((unsigned char *)c->d_scratch_out + (size_t)b * blk * row_out);
    if (s >= 2) PS_CUDA(cudaStreamWaitEvent(st, c->ev_copy[b], 0));                  // read-back of block s-2 has drained buffer b
    if (a->accumulate)
      PS_CUDA(cudaMemcpyAsync(buf, (unsigned char *)h_out + (size_t)r0 * row_out, (size_t)(r1 - r0) * row_out, cudaMemcpyHostToDevice, st));
    as = *a;
    as.k0 = a->k0 + r0; as.k1 = a->k0 + r1;
    if ((rc = ps_project(c, 1, &as, (const double *)c->d_scratch_in, buf, c->d_scratch_ws, ws, st))) return rc;
    PS_CUDA(cudaEventRecord(c->ev_comp[b], st));
    if (s > 0 && (rc = read_back(s - 1, prev_r0, prev_r1))) return rc;
    prev_r0 = r0; prev_r1 = r1;
  }
  if (s > 0 && (rc = read_back(s - 1, prev_r0, prev_r1))) return rc;
  PS_CUDA(cudaStreamSynchronize(st));
  PS_CUDA(cudaStreamSynchronize(cs));
  return PSPACE_OK;
}

static int project_host(pspace_ctx *c, int is_matrix, const pspace_project_args *a, const double *h_in, double *h_out) {
  if (!c || !a || !h_in || !h_out) return ps_fail(PSPACE_EINVAL, "null argument");
  PS_CUDA(cudaSetDevice(c->device));
  if (a->q1 < a->q0 || a->k1 < a->k0 || (is_matrix && a->j1 < a->j0)) return ps_fail(PSPACE_EINVAL, "range");
  const size_t w = c->width;
  const size_t rows = is_matrix ? (size_t)(a->k1 - a->k0) * (a->j1 - a->j0) : (size_t)(a->k1 - a->k0);
  const size_t out_bytes = sizeof(double) * w * rows * a->ncols;
  int rc;
  // one upload slab of J and at least four output blocks: pipeline the read-back over blocks of basis rows
  if (is_matrix && !a->symmetric && out_bytes >= 4 * c->out_block_bytes && a->k1 - a->k0 >= 64 &&
      sizeof(double) * w * a->ncols * (size_t)(a->q1 - a->q0) <= slab_target(c))
    return project_host_rowblocks(c, a, h_in, h_out);
  if ((rc = ensure(&c->d_scratch_out, &c->scratch_out_bytes, out_bytes))) return rc;
  if (!c->own_stream) PS_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  cudaStream_t st = c->own_stream;
  if (a->accumulate) PS_CUDA(cudaMemcpyAsync(c->d_scratch_out, h_out, out_bytes, cudaMemcpyHostToDevice, st));
  if ((rc = project_hostin(c, is_matrix, a, h_in, (double *)c->d_scratch_out))) return rc;
  PS_CUDA(cudaMemcpyAsync(h_out, c->d_scratch_out, out_bytes, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  return PSPACE_OK;
}

int pspace_cuda_project_vector_host(pspace_ctx *c, const pspace_project_args *a, const double *h_f, double *h_F) {
  return project_host(c, 0, a, h_f, h_F);
}
int pspace_cuda_project_matrix_host(pspace_ctx *c, const pspace_project_args *a, const double *h_J, double *h_C) {
  return project_host(c, 1, a, h_J, h_C);
}
int pspace_cuda_project_matrix_hostin(pspace_ctx *c, const pspace_project_args *a, const double *h_J, double *d_C) {
  if (!c || !a || !h_J || !d_C) return ps_fail(PSPACE_EINVAL, "null argument");
  if (a->q1 < a->q0 || a->k1 < a->k0 || a->j1 < a->j0) return ps_fail(PSPACE_EINVAL, "range");
  int rc = project_hostin(c, 1, a, h_J, d_C);
  if (rc) return rc;
  PS_CUDA(cudaStreamSynchronize(c->own_stream));
  return PSPACE_OK;
}

int pspace_cuda_activate(const pspace_ctx *c) {
  if (!c) return ps_fail(PSPACE_EINVAL, "null argument");
  PS_CUDA(cudaSetDevice(c->device));
  return PSPACE_OK;
}

/* CUDA-runtime shims for host code that is compiled without CUDA headers (libpspace.so) */
int pspace_cuda_malloc(void **p, size_t bytes) {
  if (!p) return ps_fail(PSPACE_EINVAL, "null argument");
  *p = nullptr;
  if (bytes == 0) return PSPACE_OK;
  PS_CUDA(cudaMalloc(p, bytes));
  return PSPACE_OK;
}
int pspace_cuda_free(void *p) {
  if (p) PS_CUDA(cudaFree(p));
  return PSPACE_OK;
}
int pspace_cuda_memcpy(void *dst, const void *src, size_t bytes, int to_device) {
  if (bytes == 0) return PSPACE_OK;
  PS_CUDA(cudaMemcpy(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
  return PSPACE_OK;
}

long long pspace_cuda_launch_count(const pspace_ctx *c) { return c ? c->launches : 0; }

int pspace_cuda_set_option(pspace_ctx *c, const char *name, int value) {
  if (!c || !name) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!strcmp(name, "force_generic")) { c->force_generic = value; return PSPACE_OK; }
  if (!strcmp(name, "out_block_mb")) { c->out_block_bytes = (size_t)std::max(1, value) << 20; return PSPACE_OK; }
  if (!strcmp(name, "no_ragged")) { c->no_ragged = value; return PSPACE_OK; }
  if (!strcmp(name, "force_panel")) { c->force_panel = value; return PSPACE_OK; }
  if (!strcmp(name, "slab_mb")) { c->slab_bytes = value < 0 ? -((long long)(-value) << 20) : ((long long)value << 20); return PSPACE_OK; }
  if (!strcmp(name, "panel_cache")) { c->panel_cache = value; c->invalidate_panels(); return PSPACE_OK; }
  return ps_fail(PSPACE_EINVAL, "unknown option %s", name);
}

int pspace_cuda_last_plan(const pspace_ctx *c, char *buf, int buflen) {
  if (!c || !buf || buflen < 1) return ps_fail(PSPACE_EINVAL, "null argument");
  snprintf(buf, buflen, "%s flop=%.6e", c->plan.c_str(), c->plan_flop);
  return PSPACE_OK;
}

}  // extern "C"
