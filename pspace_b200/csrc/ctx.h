// Internal: device-side state of one ParameterContainer (behind the opaque pspace_ctx of include/pspace_cuda.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/pspace_cuda.h"

// multi-GPU fused projection: the peer-mapped receive buffers of all ranks (project.cu, EPI_PEER)
struct PeerInfo {
  int world, rank;
  double *recv[PSPACE_MAX_PEERS];
};

struct pspace_ctx {
  int device = 0, is_complex = 0, nvars = 0, basis_type = 0, width = 1;  // width = doubles per scalar
  int kinds[PSPACE_MAX_VARS] = {0}, dmax[PSPACE_MAX_VARS] = {0};
  double p1[2 * PSPACE_MAX_VARS] = {0}, p2[2 * PSPACE_MAX_VARS] = {0};
  bool basis_ready = false, quad_ready = false, tables_ready = false;
  int pmax[PSPACE_MAX_VARS] = {0}, nq[PSPACE_MAX_VARS] = {0};
  int N = 0;
  long long Q = 0;
  int *d_deg = nullptr;                  // [N][nvars]
  std::vector<int> h_deg;
  int roff[PSPACE_MAX_VARS + 1] = {0};   // rule offsets (scalars): rule of parameter d at roff[d], nq[d] entries
  double *d_zp = nullptr, *d_yp = nullptr, *d_wp = nullptr;
  int toff[PSPACE_MAX_VARS + 1] = {0};   // table offsets (scalars): T_d at toff[d], (pmax[d]+1) x nq[d], row-major
  int tsz = 0;
  double *d_T = nullptr;
  double *d_Tw = nullptr;                // weighted tables Tw_d[alpha][i] = T_d[alpha][i] * w_d[i]  (same offsets as d_T)
  int DP = 16;                           // bytes per row of the digit table (16 or 32)
  uint8_t *d_qdig = nullptr;             // [Q][DP] mixed-radix digits of q (last variable fastest)
  double *d_W = nullptr;                 // [Q] scalars
  size_t cap_rules = 0, cap_T = 0, cap_qdig = 0, cap_W = 0;   // bytes allocated (buffers are grow-only: re-initialising
                                                              // with the same sizes does not touch the allocator)
  long long launches = 0;
  std::string plan;
  const char *plan_kernel = "";
  int plan_chunk_pts = 32;               // quadrature points per pipeline stage of the kernel that ran
  // Basis panels (project.cu ensure_panel): one slot for the unweighted rows psi_r(z_q) — the i side of the matrix
  // projection and the left operand of the expansion — and one for the weighted rows w_q psi_r(z_q) — the j side of
  // both projections.  Grow-only buffers; a slot is reused when its (row range, q-range) repeats, so a time-stepping
  // loop of expand -> caller -> project builds each panel once.
  struct PanelSlot {
    double *d = nullptr;
    size_t bytes = 0;
    bool valid = false;
    int r0 = 0, nr = 0;
    long long q0 = 0, q1 = 0;
    // stream ordering: `ev` is recorded after the last kernel that wrote or read the slot; a call on ANY stream waits
    // on it before it reuses, rebuilds or frees the slot (calls may come from different streams; the host side of a
    // context is for one thread at a time — see include/pspace_cuda.h)
    cudaEvent_t ev = nullptr;
    bool ev_pending = false;
  };
  PanelSlot panel_u, panel_w;
  void invalidate_panels() { panel_u.valid = panel_w.valid = false; }
  int panel_cache = 1;                   // reuse a panel when the (row range, q-range) of a call repeats
  int force_panel = 0;                   // tests / bench: one-pass operators through the round-1 panel kernels
  size_t out_block_bytes = (size_t)256 << 20;   // host-output matrix form: read-back block (option out_block_mb)
  int no_ragged = 0;                     // tests / A-B timing: keep the full last chunk in k_project_ws (see RAG there)
  int force_generic = 0;                 // tests: force the cp.async kernel even where TMA is usable
  double plan_flop = 0.0;
  // scratch owned by the context for the host-buffer entry points
  void *d_scratch_in = nullptr, *d_scratch_out = nullptr, *d_scratch_ws = nullptr;
  size_t scratch_in_bytes = 0, scratch_out_bytes = 0, scratch_ws_bytes = 0;
  cudaStream_t own_stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr};
  long long slab_bytes = 0;              // host-input path: bytes of J per upload slab (0 = 384 MiB)
};

// error plumbing (capi.cu)
int ps_fail(int code, const char *fmt, ...);
#define PS_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return ps_fail(PSPACE_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define PS_LAUNCH_CHECK(ctxp)                                                                \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess)                                                                  \
      return ps_fail(PSPACE_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
    if (ctxp) ((pspace_ctx *)(ctxp))->launches++;                                            \
  } while (0)

// The last variable is the fastest index of the tensor grid, so the product over the leading variables [0, dsplit) is
// shared by tail_stride consecutive points.  Picks the tail length (<= max_tail trailing variables) that minimises the
// table reads per basis value, tail + (nvars - tail) / tail_stride.  Used by k_psi_panel and k_eval_basis_grid.
static inline void ps_choose_tail(const pspace_ctx *c, int max_tail, int *dsplit, int *tail_stride) {
  double best_cost = (double)c->nvars;
  long long stride = 1;
  *dsplit = c->nvars;
  *tail_stride = 1;
  for (int t = 1; t <= max_tail && t <= c->nvars; t++) {
    stride *= c->nq[c->nvars - t];
    if (stride > (1 << 20)) break;
    const double cost = t + (double)(c->nvars - t) / (double)stride;
    if (cost < best_cost) { best_cost = cost; *dsplit = c->nvars - t; *tail_stride = (int)stride; }
  }
}

// implemented in basis.cu / tables.cu / project.cu, called from capi.cu
int ps_basis_degrees_device(int nvars, const int *pmax, int basis_type, int *nbasis, int *d_out,
                            long long capacity_rows, cudaStream_t st, long long *launches);
int ps_basis_count_host(int nvars, const int *pmax, int basis_type, long long *nbasis);
int ps_build_tables(pspace_ctx *ctx, cudaStream_t st);
int ps_build_grid(pspace_ctx *ctx, cudaStream_t st);
int ps_tensor_grid(const pspace_ctx *ctx, long long q0, long long q1, double *d_Z, double *d_Y, double *d_W, cudaStream_t st);
int ps_eval_basis_grid(const pspace_ctx *ctx, int k0, int k1, long long q0, long long q1, double *d_psi, cudaStream_t st);
int ps_eval_basis_points(const pspace_ctx *ctx, int k0, int k1, long long npts, const double *d_z, double *d_psi, cudaStream_t st);
size_t ps_project_workspace(const pspace_ctx *ctx, int is_matrix, const pspace_project_args *a);
int ps_expand(const pspace_ctx *c, int k0, int k1, long long q0, long long q1, int ncols, int accumulate,
              const double *d_V, double *d_U, cudaStream_t st);
int ps_project(const pspace_ctx *ctx, int is_matrix, const pspace_project_args *a, const double *d_in, double *d_out,
               void *d_ws, size_t ws_bytes, cudaStream_t st);
int ps_project_ex(const pspace_ctx *ctx, int is_matrix, const pspace_project_args *a, const double *d_in, double *d_out,
                  void *d_ws, size_t ws_bytes, cudaStream_t st, const PeerInfo *peer);
// factored-operand vector projection (vproject.cu)
size_t ps_vproject_workspace(const pspace_ctx *ctx, const pspace_project_args *a);
bool ps_vproject_usable(const pspace_ctx *ctx, const pspace_project_args *a, const double *d_in);
int ps_vproject(const pspace_ctx *ctx, const pspace_project_args *a, const double *d_in, double *d_out, void *d_ws,
                size_t ws_bytes, cudaStream_t st);
// records that the kernels enqueued on `st` so far use the panel slots (project.cu)
void ps_panels_mark_used(pspace_ctx *c, cudaStream_t st);
int ps_peer_layout(const pspace_ctx *ctx, const pspace_project_args *a, int world, long long *recv_elems);
int ps_peer_reduce_gather(const pspace_ctx *ctx, const pspace_project_args *a, const double *d_recv, double *const *peer_out,
                          int world, int rank, cudaStream_t st);
