/* pspace_cuda.h — C-ABI of the B200-native stochastic-Galerkin projection path (libpspace_cuda.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, `extern "C"`, int status returns, no C++ or
 * torch types.  The host C++ library (libpspace.so: ParameterFactory / ParameterContainer with the
 * reference's own headers and signatures) calls ONLY these entry points for everything batched; a
 * maintainer of the reference binds them the same way.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Scalars: `double` in the real build; in the USE_COMPLEX build (reference src/include/scalar.h:18-24)
 * a scalar is an interleaved (re,im) pair of doubles.  A context is created real or complex
 * (`is_complex`), and every `scalar` array argument below is `const double*` / `double*` holding
 * 1 or 2 doubles per scalar accordingly.
 *
 * Memory: arguments named d_* are DEVICE pointers owned by the caller; h_* are HOST pointers.
 * Streams: `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls on one context may use
 * different streams: context-owned scratch (the basis panels of the matrix projection) is ordered across streams with
 * events.  The HOST side of a context is not re-entrant: one thread at a time per context (entry points that change
 * context state take a non-const `pspace_ctx *`).
 * Errors: every function returns PSPACE_OK (0) or a negative PSPACE_E* code; pspace_cuda_last_error()
 * returns a thread-local message.  There is NO CPU fallback: without a usable CUDA device the calls
 * fail with PSPACE_ECUDA.
 */
#ifndef PSPACE_CUDA_H
#define PSPACE_CUDA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PSPACE_OK 0
#define PSPACE_EINVAL (-1)   /* bad argument */
#define PSPACE_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define PSPACE_ESTATE (-3)   /* call out of order (e.g. project before initialize) */
#define PSPACE_ENOMEM (-4)
#define PSPACE_MAX_VARS 32
#define PSPACE_MAX_PEERS 16  /* ranks of the fused multi-GPU exchange */
#define PSPACE_MAX_NQ 20     /* 1-D rules exist for npoints 1..20 (GaussianQuadrature.cpp:31-1483) */

/* parameter kinds = the three ParameterFactory products (src/include/ParameterFactory.h:21-28) */
#define PSPACE_NORMAL 0       /* (mu, sigma)  Gauss-Hermite / unit Hermite  */
#define PSPACE_UNIFORM 1      /* (a, b)       Gauss-Legendre / unit shifted Legendre */
#define PSPACE_EXPONENTIAL 2  /* (mu, beta)   Gauss-Laguerre / Laguerre */

typedef struct pspace_ctx pspace_ctx;

const char *pspace_cuda_last_error(void);
int pspace_cuda_version(void);
/* number of visible CUDA devices, or a negative error code */
int pspace_cuda_device_count(void);

/* ---- context = device-side state of one ParameterContainer (src/include/ParameterContainer.h:17-58) --
 * kinds[nvars], p1/p2[nvars] scalars (host), dmax[nvars]; basis_type 0 tensor / !=0 complete
 * (BasisHelper.cpp:40-61).  Replaces: ParameterContainer ctor + addParameter (ParameterContainer.cpp:11-15,50-56). */
int pspace_cuda_create(pspace_ctx **out, int device, int is_complex, int nvars, const int *kinds,
                       const double *h_p1, const double *h_p2, const int *dmax, int basis_type);
int pspace_cuda_destroy(pspace_ctx *ctx);

/* ---- (1) multi-index basis construction: integer kernel, bit-exact ------------------------------
 * Replaces BasisHelper::basisDegrees (BasisHelper.cpp:33-66 and the unrolled variants :98-584) as
 * called by ParameterContainer::initializeBasis (ParameterContainer.cpp:84-109).
 * Stateless form: writes d_out[nbasis][nvars] (capacity_rows rows available) and *nbasis.  */
int pspace_cuda_basis_degrees(int nvars, const int *pmax, int basis_type, int *nbasis, int *d_out,
                              long long capacity_rows, void *stream);
/* number of basis terms without producing them (for sizing d_out) */
int pspace_cuda_basis_count(int nvars, const int *pmax, int basis_type, long long *nbasis);
/* context form: builds and keeps the degree table on the device */
int pspace_cuda_initialize_basis(pspace_ctx *ctx, const int *pmax, void *stream);

/* ---- (2) 1-D rules, orthonormal 1-D basis tables, tensor grid ------------------------------------
 * Replaces ParameterContainer::initializeQuadrature (ParameterContainer.cpp:116-160):
 * {Normal,Uniform,Exponential}Parameter::quadrature -> GaussianQuadrature::*Quadrature
 * (GaussianQuadrature.cpp:510-514,1026-1032,1479-1482), QuadratureHelper::tensorProduct
 * (QuadratureHelper.cpp:34-132), and tabulates unit_{hermite,legendre,laguerre}
 * (OrthogonalPolynomials.cpp:42-44,66-68,90-92) at the 1-D nodes. */
int pspace_cuda_initialize_quadrature(pspace_ctx *ctx, const int *nqpts, void *stream);
/* ParameterContainer::initialize (ParameterContainer.cpp:224-237): pmax = dmax, nq = dmax+1 */
int pspace_cuda_initialize(pspace_ctx *ctx, void *stream);

/* accessors (ParameterContainer.cpp:61-77) */
int pspace_cuda_num_basis(const pspace_ctx *ctx);
int pspace_cuda_num_params(const pspace_ctx *ctx);
long long pspace_cuda_num_quad(const pspace_ctx *ctx);
int pspace_cuda_is_complex(const pspace_ctx *ctx);
/* copy results back to the host (the host library mirrors them for the per-call accessors
 * getBasisParamDeg / quadrature(q,zq,yq), ParameterContainer.cpp:169-176,200-205) */
int pspace_cuda_get_degrees(const pspace_ctx *ctx, int *h_out /*[N][nvars]*/);
int pspace_cuda_get_rules(const pspace_ctx *ctx, int pid, double *h_z, double *h_y, double *h_w /*[nq_pid] scalars*/);
int pspace_cuda_get_table(const pspace_ctx *ctx, int pid, double *h_T /*[(pmax+1)][nq] scalars*/);
int pspace_cuda_get_weights(const pspace_ctx *ctx, long long q0, long long q1, double *h_W /*[q1-q0] scalars*/);
/* materialise grid columns q0..q1-1: d_Z,d_Y [q1-q0][nvars] scalars (point-major), d_W [q1-q0]; any may be NULL
 * (QuadratureHelper.cpp:105-129; layout as handed out by quadrature(q,zq,yq)) */
int pspace_cuda_tensor_grid(const pspace_ctx *ctx, long long q0, long long q1, double *d_Z, double *d_Y,
                            double *d_W, void *stream);

/* ---- (2b) batched basis evaluation ---------------------------------------------------------------
 * Replaces N*npts calls of ParameterContainer::basis(k, z) (ParameterContainer.cpp:184-192).
 * Grid form: d_psi[(k1-k0)][(q1-q0)] = psi_k(z_q) at quadrature points (table products).
 * Point form: arbitrary points d_z[npts][nvars]; polynomials evaluated with the reference formulas. */
int pspace_cuda_eval_basis_grid(const pspace_ctx *ctx, int k0, int k1, long long q0, long long q1,
                                double *d_psi, void *stream);
int pspace_cuda_eval_basis_points(const pspace_ctx *ctx, int k0, int k1, long long npts, const double *d_z,
                                  double *d_psi, void *stream);

/* ---- (3) Galerkin projections ---------------------------------------------------------------------
 * Vector: F[k][c] = sum_q w_q psi_k(z_q) f[q][c]        (TACSStochasticElement.cpp:283-311 addResidual,
 *         :196-238 getInitConditions, :487-530 evalPointQuantity; pspace/core.py:750-818)
 * Matrix: C[i][j][c] = sum_q w_q psi_i(z_q) psi_j(z_q) J[q][c]   (TACSStochasticElement.cpp:380-414
 *         addJacobian; pspace/core.py:840-911), all ordered pairs of the window [i0,i1) x [j0,j1).
 * q runs over [q0,q1) (a quadrature-point shard; the full projection is q0=0,q1=Q); d_f / d_J hold ONLY
 * rows q0..q1-1: d_f[(q1-q0)][ncols], d_J[(q1-q0)][ncols] scalars, row-major, ncols = m resp. m*m.
 * Output: d_F[(k1-k0)][ncols]; d_C[(i1-i0)][(j1-j0)][ncols] scalars, row-major.
 * accumulate = 0 overwrites the output, 1 adds to it (the reference callers add into res/mat).
 * Workspace: the call may need scratch for its split over q; query with *_workspace and pass a device
 * buffer of at least that many bytes (may be NULL when the query returns 0). */
typedef struct {
  int k0, k1;               /* vector: basis range; matrix: i range */
  int j0, j1;               /* matrix only: j range */
  long long q0, q1;         /* quadrature-point range of this call */
  int ncols;                /* m (vector) or m*m (matrix), in scalars */
  int accumulate;
  int ksplit;               /* 0 = let the library choose; >0 forces the number of q-slices */
  int variant;              /* 0 = auto; otherwise a tile-variant id (bench / tuning) */
  /* matrix only, optional: DEVICE [N][N] bytes from pspace_cuda_pair_mask (the reference's nonzero() filter,
   * TACSStochasticElement.cpp:7-21).  Pairs with mask 0 are structurally zero: their blocks are written as 0
   * (left untouched when accumulate = 1) and tiles without any kept pair are skipped.  NULL = all pairs. */
  const unsigned char *d_pair_mask;
  /* matrix only, optional: 1 = exploit C[i][j] == C[j][i] (psi_i psi_j = psi_j psi_i): only tiles holding a pair
   * j >= i are computed and each block is stored to (i,j) and (j,i).  Needs a square window (k0==j0, k1==j1).  The
   * result is the same full window; about half the arithmetic.  The reference computes all N^2 ordered pairs
   * (TACSStochasticElement.cpp:380-388), so this is off by default and reported separately by bench.py. */
  int symmetric;
} pspace_project_args;

size_t pspace_cuda_project_workspace(const pspace_ctx *ctx, int is_matrix, const pspace_project_args *a);
int pspace_cuda_project_vector(pspace_ctx *ctx, const pspace_project_args *a, const double *d_f,
                               double *d_F, void *d_workspace, size_t workspace_bytes, void *stream);
int pspace_cuda_project_matrix(pspace_ctx *ctx, const pspace_project_args *a, const double *d_J,
                               double *d_C, void *d_workspace, size_t workspace_bytes, void *stream);
/* Many elements sharing one container (BASELINE.json configs[2]: "24x24 element Jacobians over 10k elements"): element e
 * reads d_J + e*j_elem_stride and writes d_C + e*c_elem_stride (strides in scalars); same window / q-range / workspace
 * for all.  The basis panels are built for the first element and reused. */
int pspace_cuda_project_matrix_batched(pspace_ctx *ctx, const pspace_project_args *a, int nelem, const double *d_J,
                                       long long j_elem_stride, double *d_C, long long c_elem_stride, void *d_workspace,
                                       size_t workspace_bytes, void *stream);
/* Host-buffer forms (the call a host-side user of the reference makes: pageable or pinned host
 * arrays in, host array out; H2D + kernels + D2H inside, synchronous). */
int pspace_cuda_project_vector_host(pspace_ctx *ctx, const pspace_project_args *a, const double *h_f, double *h_F);
int pspace_cuda_project_matrix_host(pspace_ctx *ctx, const pspace_project_args *a, const double *h_J, double *h_C);
/* Host input, DEVICE output (complete on return): for callers that combine the result on the device first
 * (multi-GPU allreduce) before reading it back.  The host forms upload J in slabs of quadrature points through a
 * double buffer, overlapping the upload of slab s+1 with the projection of slab s (option "slab_mb"). */
int pspace_cuda_project_matrix_hostin(pspace_ctx *ctx, const pspace_project_args *a, const double *h_J, double *d_C);

/* ---- (e) multi-GPU: q-sharded projection + ONE NCCL collective over the coefficient blocks ---------------------------
 * Quadrature points shard across ranks (one process or thread per GPU); rank r projects q in [q0_r, q1_r) — d_J / d_f hold
 * only those rows — into a full partial window, and the partial windows are summed by ncclAllReduce (every rank ends with
 * the total in d_C) or by ncclReduceScatter over blocks of basis rows (rank r ends with rows [r*per, (r+1)*per) of the total;
 * for large N^2 m^2 or row-sharded consumers).  `nccl_comm` is the caller's ncclComm_t (ncclCommInitRank under MPI, as the
 * reference's drivers initialise MPI: examples/stacs/smd/projection.cpp:36-39, or ncclCommInitAll); NCCL is resolved at run
 * time from the libnccl.so.2 already in the process, libpspace_cuda.so has no link-time dependency on it.  All calls are
 * asynchronous on `stream`.  pspace_cuda_shard_bounds gives the q-range of a rank (boundaries on multiples of 32 points,
 * the rule of pspace_b200/sharded.py); pspace_cuda_row_block the row block of the reduce-scatter (per = ceil(rows/world)).
 * reduce_scatter buffers: d_C_partial [world*per][(j1-j0)][ncols] scalars (scratch: the rank's partial window, padded),
 * d_C_rows [per][(j1-j0)][ncols] scalars (result). */
int pspace_cuda_shard_bounds(long long Q, int world, int rank, long long *q0, long long *q1);
int pspace_cuda_row_block(int rows, int world, int rank, int *r0, int *r1, int *rows_per_rank);
int pspace_cuda_nccl_allreduce(double *d_buf, size_t count_doubles, void *nccl_comm, void *stream);
int pspace_cuda_project_matrix_allreduce(pspace_ctx *ctx, const pspace_project_args *a, const double *d_J, double *d_C,
                                         void *d_workspace, size_t workspace_bytes, void *nccl_comm, void *stream);
int pspace_cuda_project_vector_allreduce(pspace_ctx *ctx, const pspace_project_args *a, const double *d_f, double *d_F,
                                         void *d_workspace, size_t workspace_bytes, void *nccl_comm, void *stream);
int pspace_cuda_project_matrix_reduce_scatter(pspace_ctx *ctx, const pspace_project_args *a, const double *d_J,
                                              double *d_C_partial, double *d_C_rows, void *d_workspace,
                                              size_t workspace_bytes, void *nccl_comm, void *stream);

/* ---- fused multi-GPU projection: compute + reduce-scatter/all-gather over peer-mapped memory (NVLink) -----------------
 * Each rank projects its quadrature-point shard; instead of keeping a full partial C and calling a collective, every
 * finished partial TILE is pushed by the projection kernel's epilogue straight into the receive slot of the rank that
 * owns the tile (peer stores, overlapped with the math of the following tiles).  After a cross-rank barrier,
 * pspace_cuda_peer_reduce_gather adds the `world` partial blocks of each owned element in rank order (deterministic)
 * and stores the total into the C window of every rank.  The caller supplies peer-mapped buffers (e.g.
 * torch.distributed._symmetric_memory) and the two barriers:
 *     project_matrix_peer  ->  barrier  ->  peer_reduce_gather  ->  barrier
 * peer_recv[w] / peer_out[w] are HOST arrays of `world` device pointers (rank w's receive buffer / C window as mapped
 * into this process).  pspace_cuda_peer_layout gives the receive-buffer size in doubles.  Restrictions: matrix mode,
 * no mask / symmetric / accumulate, even number of real columns (TMA path). */
int pspace_cuda_peer_layout(const pspace_ctx *ctx, const pspace_project_args *a, int world, long long *recv_doubles);
int pspace_cuda_project_matrix_peer(pspace_ctx *ctx, const pspace_project_args *a, const double *d_J,
                                    double *const *peer_recv, int world, int rank, void *d_workspace,
                                    size_t workspace_bytes, void *stream);
int pspace_cuda_peer_reduce_gather(const pspace_ctx *ctx, const pspace_project_args *a, const double *d_recv_local,
                                   double *const *peer_out, int world, int rank, void *stream);

/* ---- (f1) polynomial-chaos expansion evaluated at quadrature points -------------------------------------
 * U[q][c] = sum_{k in [k0,k1)} psi_k(z_q) V[k-k0][c]  for q in [q0,q1);  d_V[(k1-k0)][ncols], d_U[(q1-q0)][ncols]
 * scalars, row-major.  Replaces getDeterministicStates / getDeterministicAdjoint of the reference caller
 * (examples/stacs/cpp/TACSStochasticElement.cpp:53-120; pspace/core.py:788-792), which recompute it with N calls
 * of basis(k,zq) per node inside every (i,j,q) iteration. */
int pspace_cuda_evaluate_expansion(pspace_ctx *ctx, int k0, int k1, long long q0, long long q1, int ncols,
                                   int accumulate, const double *d_V, double *d_U, void *stream);

/* ---- (f2) TACS-interleaved layouts and the basis-pair sparsity filter (pure index work, bit-exact) ---------------
 * m = nnodes*ndvpn local dofs; stochastic vectors have nsvpn = N*ndvpn variables per node.
 * scatter_vector: res[n*nsvpn + i*ndvpn + d] (+)= F[i][n*ndvpn + d]        (TACSStochasticElement.cpp:316-322)
 * gather_vector : V[k][n*ndvpn + d] = v[n*nsvpn + k*ndvpn + d]             (input of getDeterministicStates :84-120)
 * scatter_matrix: mat[(ni*nsvpn+i*ndvpn+di)*nsdof + nj*nsvpn+j*ndvpn+dj] (+)= C[i][j][(ni*ndvpn+di)*m + nj*ndvpn+dj]
 *                 with nsdof = N*m                                          (TACSStochasticElement.cpp:417-432)
 * point_quantity_layout: out[d*N + i] = F[i][d]                             (TACSStochasticElement.cpp:525)
 * pair_mask: d_mask[i*N+j] = prod_d (|deg_i[d]-deg_j[d]| <= dmapf[d])       (nonzero(), TACSStochasticElement.cpp:7-21;
 *            BasisHelper::sparse, BasisHelper.cpp:76-88).  dmapf is a host array of nvars ints. */
int pspace_cuda_scatter_vector_tacs(int N, int nnodes, int ndvpn, int is_complex, const double *d_F, double *d_res,
                                    int accumulate, void *stream);
int pspace_cuda_gather_vector_tacs(int N, int nnodes, int ndvpn, int is_complex, const double *d_v, double *d_V, void *stream);
int pspace_cuda_scatter_matrix_tacs(int N, int nnodes, int ndvpn, int is_complex, const double *d_C, double *d_mat,
                                    int accumulate, void *stream);
int pspace_cuda_point_quantity_layout(int N, int m, int is_complex, const double *d_F, double *d_out, void *stream);
int pspace_cuda_pair_mask(const pspace_ctx *ctx, const int *dmapf, unsigned char *d_mask, void *stream);

/* ---- (f3) moments at the quadrature points -------------------------------------------------------------------
 * mean[c] = sum_q (w_q psi_0(z_q)) f[q][c],  second[c] = sum_q (w_q psi_0(z_q)) f[q][c]^2,  var = second - mean^2
 * over q in [q0,q1); d_f[(q1-q0)][ncols] scalars; outputs [ncols] scalars, any may be NULL.  Replaces
 * TACSStochasticFunction::getFunctionValue (examples/stacs/cpp/TACSStochasticFunction.cpp:209-238).  Sharded
 * use: combine mean/second across ranks, then var = second - mean^2. */
int pspace_cuda_moments(const pspace_ctx *ctx, long long q0, long long q1, int ncols, const double *d_f,
                        double *d_mean, double *d_second, double *d_var, void *stream);

/* ---- (f4) adjoint-residual product ------------------------------------------------------------------------------
 * TACSStochasticElement::addAdjResProduct (examples/stacs/cpp/TACSStochasticElement.cpp:548-637), three steps around the
 * caller's deterministic element (m = nnodes*ndvpn local dofs; v, adj are TACS-interleaved stochastic vectors
 * [nnodes*N*ndvpn] scalars on the device):
 *   adjoint_states : d_UP[(q1-q0)][2m], columns [0,m) = deterministic states u_q, [m,2m) = deterministic adjoint psi_q
 *                    (getDeterministicStates :84-120, getDeterministicAdjoint :53-82) — one expansion pass for both;
 *   (caller)       : G[q][n] = g_n(psi_q, u_q, y_q), n < dvLen — the deterministic element's addAdjResProduct integrand;
 *   adjoint_product: d_dfdx[n] (+)= scale * sum_{q in [q0,q1)} (w_q psi_0(z_q)) G[q][n]   (wt = basis(0,zq)*wq :590; the
 *                    reference projects on basis index 0 only, :583).  h_scale points at ONE scalar on the host.
 * A q-range = a rank's shard; the partial dfdx of the ranks add (pspace_cuda_nccl_allreduce). */
int pspace_cuda_adjoint_states(pspace_ctx *ctx, int nnodes, int ndvpn, long long q0, long long q1, const double *d_v,
                               const double *d_adj, double *d_UP, void *stream);
int pspace_cuda_adjoint_product(pspace_ctx *ctx, int dvlen, const double *h_scale, long long q0, long long q1,
                                const double *d_G, double *d_dfdx, int accumulate, void *stream);

/* make the context's device current for the calling thread (the shims below allocate on the CURRENT device) */
int pspace_cuda_activate(const pspace_ctx *ctx);
/* CUDA-runtime shims so that host code built without CUDA headers can stage buffers (to_device: 1 = H2D, 0 = D2H) */
int pspace_cuda_malloc(void **p, size_t bytes);
int pspace_cuda_free(void *p);
int pspace_cuda_memcpy(void *dst, const void *src, size_t bytes, int to_device);

/* launches of this library's kernels since the context was created (bench "gpu_launches") */
long long pspace_cuda_launch_count(const pspace_ctx *ctx);
/* tuning / test switches.  "force_generic": 0 (default) = the planner chooses between the TMA warp-specialised stream-K
 * kernel and the single-launch cp.async kernel by size (matrix windows under ~1 GFLOP, and operands TMA cannot address
 * - odd column counts, unaligned bases - take the single-launch kernel); 1 = always the single-launch kernel; -1 = the
 * TMA kernel wherever it can run.  "force_panel" = 1: one-pass operators through the panel kernels.
 * "panel_cache" = 0: rebuild the basis panels on every call.  "slab_mb" = n: upload slab size of the host-input forms.
 * "out_block_mb" = n (default 256): pspace_cuda_project_matrix_host reads C back in blocks of basis rows of about n MB,
 * overlapped with the projection of the next block, when J fits one upload slab and C is at least four blocks.
 * "no_ragged" = 1: keep the full last TMA stage in the matrix kernel (A/B timing; same bits either way). */
int pspace_cuda_set_option(pspace_ctx *ctx, const char *name, int value);
/* description + FLOP count of the kernel the last project call used */
int pspace_cuda_last_plan(const pspace_ctx *ctx, char *buf, int buflen);

/* ---- measurement helpers (bench.py) --------------------------------------------------------------
 * FP64 peak of this device measured with a register-resident loop: mode 0 = DFMA pipe,
 * 1 = DMMA (mma.sync.m8n8k4.f64).  Runs ~ms_target milliseconds; returns TFLOP/s. */
int pspace_cuda_fp64_peak(int device, int mode, double ms_target, double *tflops);
/* splitmix64 synthetic blocks written on the device: out[i] = u(seed ^ (offset+i)) in (-0.5,0.5) */
int pspace_cuda_fill_synthetic(double *d_out, long long n, unsigned long long seed, unsigned long long offset,
                               void *stream);

#ifdef __cplusplus
}
#endif
#endif
