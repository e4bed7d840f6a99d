"""DeviceContainer — Python handle on the device-side state of one ParameterContainer.

This is the layer the tests, smoke() and bench.py drive: it owns a `pspace_ctx*` of the C-ABI and moves
data with torch (device memory, streams).  Names follow the reference container
(src/include/ParameterContainer.h:17-41): initialize / initializeBasis / initializeQuadrature /
getNumBasisTerms / getNumQuadraturePoints / getBasisParamDeg / quadrature / basis, plus the batched
additions projectVector / projectMatrix / evalBasis.
"""
import ctypes

import numpy as np
import torch

from . import cabi
from .cabi import ProjectArgs, check

KIND = {"normal": 0, "uniform": 1, "exponential": 2}


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceContainer:
    def __init__(self, params, basis_type=0, complex_=False, device=0):
        """params: [(kind, p1, p2, dmax)] in ParameterFactory order (ids 0..nvars-1)."""
        self.L = cabi.lib()
        self.is_complex = bool(complex_)
        self.np_dtype = np.complex128 if complex_ else np.float64
        self.t_dtype = torch.complex128 if complex_ else torch.float64
        self.device = torch.device("cuda", device)
        self.nvars = len(params)
        kinds = np.array([KIND[p[0]] if isinstance(p[0], str) else int(p[0]) for p in params], dtype=np.intc)
        p1 = np.array([p[1] for p in params], dtype=self.np_dtype)
        p2 = np.array([p[2] for p in params], dtype=self.np_dtype)
        dmax = np.array([p[3] for p in params], dtype=np.intc)
        h = ctypes.c_void_p()
        check(self.L.pspace_cuda_create(ctypes.byref(h), device, int(complex_), self.nvars, cabi.np_ptr(kinds),
                                        cabi.np_ptr(p1), cabi.np_ptr(p2), cabi.np_ptr(dmax), basis_type))
        self.h = h
        self._ws = None

    def close(self):
        if getattr(self, "h", None):
            self.L.pspace_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- initialisation (ParameterContainer.cpp:84-160, 224-237) ---------------------------------
    def initialize(self):
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_initialize(self.h, _stream_ptr()))
        return self

    def initializeBasis(self, pmax):
        a = np.ascontiguousarray(pmax, dtype=np.intc)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_initialize_basis(self.h, cabi.np_ptr(a), _stream_ptr()))
        return self

    def initializeQuadrature(self, nqpts):
        a = np.ascontiguousarray(nqpts, dtype=np.intc)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_initialize_quadrature(self.h, cabi.np_ptr(a), _stream_ptr()))
        return self

    # ---- accessors -------------------------------------------------------------------------------
    def getNumBasisTerms(self):
        return self.L.pspace_cuda_num_basis(self.h)

    def getNumQuadraturePoints(self):
        return self.L.pspace_cuda_num_quad(self.h)

    def getNumParameters(self):
        return self.L.pspace_cuda_num_params(self.h)

    N = property(getNumBasisTerms)
    Q = property(getNumQuadraturePoints)

    def degrees(self):
        out = np.zeros((self.N, self.nvars), dtype=np.intc)
        check(self.L.pspace_cuda_get_degrees(self.h, cabi.np_ptr(out)))
        return out

    def getBasisParamDeg(self, k):
        return self.degrees()[k].copy()

    def rules(self, pid, n):
        z = np.zeros(n, dtype=self.np_dtype); y = np.zeros(n, dtype=self.np_dtype); w = np.zeros(n, dtype=self.np_dtype)
        check(self.L.pspace_cuda_get_rules(self.h, pid, cabi.np_ptr(z), cabi.np_ptr(y), cabi.np_ptr(w)))
        return z, y, w

    def table(self, pid, rows, n):
        T = np.zeros((rows, n), dtype=self.np_dtype)
        check(self.L.pspace_cuda_get_table(self.h, pid, cabi.np_ptr(T)))
        return T

    def weights(self, q0=0, q1=None):
        q1 = self.Q if q1 is None else q1
        W = np.zeros(q1 - q0, dtype=self.np_dtype)
        check(self.L.pspace_cuda_get_weights(self.h, q0, q1, cabi.np_ptr(W)))
        return W

    def grid(self, q0=0, q1=None):
        """Materialised (Z, Y, W) for q in [q0,q1) as device tensors, point-major."""
        q1 = self.Q if q1 is None else q1
        n = q1 - q0
        Z = torch.empty((n, self.nvars), dtype=self.t_dtype, device=self.device)
        Y = torch.empty((n, self.nvars), dtype=self.t_dtype, device=self.device)
        W = torch.empty((n,), dtype=self.t_dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_tensor_grid(self.h, q0, q1, Z.data_ptr(), Y.data_ptr(), W.data_ptr(), _stream_ptr()))
        return Z, Y, W

    # ---- batched basis evaluation ------------------------------------------------------------------
    def evalBasisGrid(self, k0=0, k1=None, q0=0, q1=None):
        k1 = self.N if k1 is None else k1
        q1 = self.Q if q1 is None else q1
        psi = torch.empty((k1 - k0, q1 - q0), dtype=self.t_dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_eval_basis_grid(self.h, k0, k1, q0, q1, psi.data_ptr(), _stream_ptr()))
        return psi

    def evalBasis(self, z, k0=0, k1=None):
        """psi[k][p] = basis(k, z_p) for arbitrary points z [npts][nvars] (device or host array)."""
        k1 = self.N if k1 is None else k1
        zt = torch.as_tensor(z, dtype=self.t_dtype).reshape(-1, self.nvars).contiguous().to(self.device)
        psi = torch.empty((k1 - k0, zt.shape[0]), dtype=self.t_dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_eval_basis_points(self.h, k0, k1, zt.shape[0], zt.data_ptr(), psi.data_ptr(), _stream_ptr()))
        return psi

    def evaluateExpansion(self, V, k0=0, k1=None, q0=0, q1=None, out=None, accumulate=False):
        """U[q][c] = sum_k psi_k(z_q) V[k][c]: the PC expansion with coefficient rows V (device [k1-k0][m]) evaluated
        at quadrature points q0..q1-1 (getDeterministicStates, TACSStochasticElement.cpp:84-120)."""
        k1 = self.N if k1 is None else k1
        q1 = self.Q if q1 is None else q1
        assert V.is_cuda and V.dtype == self.t_dtype and V.is_contiguous() and V.shape[0] == k1 - k0
        m = V.shape[1]
        if out is None:
            out = torch.empty((q1 - q0, m), dtype=self.t_dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_evaluate_expansion(self.h, k0, k1, q0, q1, m, int(accumulate), V.data_ptr(),
                                                        out.data_ptr(), _stream_ptr()))
        return out

    # ---- projections ---------------------------------------------------------------------------------
    def _args(self, k0, k1, j0, j1, q0, q1, ncols, accumulate, ksplit, variant, mask=None, symmetric=False):
        return ProjectArgs(k0, k1, j0, j1, q0, q1, ncols, int(accumulate), ksplit, variant,
                           mask.data_ptr() if mask is not None else None, int(symmetric))

    def _workspace(self, is_matrix, a):
        need = self.L.pspace_cuda_project_workspace(self.h, is_matrix, ctypes.byref(a))
        self._last_ws_need = need                    # tests: bytes of the workspace the call may touch
        if need == 0:
            return None, 0
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    def projectVector(self, f, k0=0, k1=None, q0=0, q1=None, out=None, accumulate=False, ksplit=0, variant=0):
        """F[k][c] = sum_q w_q psi_k(z_q) f[q][c]; f holds rows q0..q1-1 (device tensor [q1-q0][m])."""
        k1 = self.N if k1 is None else k1
        q1 = self.Q if q1 is None else q1
        assert f.is_cuda and f.dtype == self.t_dtype and f.is_contiguous() and f.shape[0] == q1 - q0
        m = f.shape[1]
        if out is None:
            out = torch.empty((k1 - k0, m), dtype=self.t_dtype, device=self.device)
        a = self._args(k0, k1, 0, 0, q0, q1, m, accumulate, ksplit, variant)
        ws, need = self._workspace(0, a)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_project_vector(self.h, ctypes.byref(a), f.data_ptr(), out.data_ptr(),
                                                    ws.data_ptr() if ws is not None else None, need, _stream_ptr()))
        return out

    def projectMatrix(self, J, i0=0, i1=None, j0=0, j1=None, q0=0, q1=None, out=None, accumulate=False, ksplit=0,
                      variant=0, mask=None, symmetric=False):
        """C[i][j][c] = sum_q w_q psi_i psi_j J[q][c]; J holds rows q0..q1-1 (device tensor [q1-q0][m2])."""
        i1 = self.N if i1 is None else i1
        j1 = self.N if j1 is None else j1
        q1 = self.Q if q1 is None else q1
        assert J.is_cuda and J.dtype == self.t_dtype and J.is_contiguous() and J.shape[0] == q1 - q0
        m2 = J.shape[1]
        if out is None:
            out = torch.empty((i1 - i0, j1 - j0, m2), dtype=self.t_dtype, device=self.device)
        if mask is not None:
            assert mask.is_cuda and mask.dtype == torch.uint8 and mask.shape == (self.N, self.N) and mask.is_contiguous()
        a = self._args(i0, i1, j0, j1, q0, q1, m2, accumulate, ksplit, variant, mask, symmetric)
        ws, need = self._workspace(1, a)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_project_matrix(self.h, ctypes.byref(a), J.data_ptr(), out.data_ptr(),
                                                    ws.data_ptr() if ws is not None else None, need, _stream_ptr()))
        return out

    def projectMatrixBatched(self, J, i0=0, i1=None, j0=0, j1=None, q0=0, q1=None, out=None):
        """Elements sharing this container: J device [E][q1-q0][m2] -> C [E][i1-i0][j1-j0][m2] (panels built once)."""
        i1 = self.N if i1 is None else i1
        j1 = self.N if j1 is None else j1
        q1 = self.Q if q1 is None else q1
        assert J.is_cuda and J.dtype == self.t_dtype and J.is_contiguous() and J.dim() == 3 and J.shape[1] == q1 - q0
        E, _, m2 = J.shape
        if out is None:
            out = torch.empty((E, i1 - i0, j1 - j0, m2), dtype=self.t_dtype, device=self.device)
        a = self._args(i0, i1, j0, j1, q0, q1, m2, False, 0, 0)
        ws, need = self._workspace(1, a)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_project_matrix_batched(self.h, ctypes.byref(a), E, J.data_ptr(), (q1 - q0) * m2, out.data_ptr(),
                                                            (i1 - i0) * (j1 - j0) * m2, ws.data_ptr() if ws is not None else None,
                                                            need, _stream_ptr()))
        return out

    def projectMatrixHost(self, J, i0=0, i1=None, j0=0, j1=None, q0=0, q1=None, out=None, accumulate=False):
        """Host-buffer form (numpy or pinned torch CPU tensors in/out); copies are inside the call (J rows q0..q1-1)."""
        i1 = self.N if i1 is None else i1
        j1 = self.N if j1 is None else j1
        q1 = self.Q if q1 is None else q1
        Jn = J.numpy() if isinstance(J, torch.Tensor) else np.ascontiguousarray(J, dtype=self.np_dtype)
        m2 = Jn.shape[1]
        if out is None:
            out = np.zeros((i1 - i0, j1 - j0, m2), dtype=self.np_dtype)
        on = out.numpy() if isinstance(out, torch.Tensor) else out
        a = self._args(i0, i1, j0, j1, q0, q1, m2, accumulate, 0, 0)
        check(self.L.pspace_cuda_project_matrix_host(self.h, ctypes.byref(a), Jn.ctypes.data, on.ctypes.data))
        return out

    def projectMatrixHostIn(self, J, out, i0=0, i1=None, j0=0, j1=None, q0=0, q1=None, accumulate=False):
        """Host J rows (numpy / pinned torch CPU tensor) -> DEVICE tensor `out`, slab-pipelined upload; complete on return."""
        i1 = self.N if i1 is None else i1
        j1 = self.N if j1 is None else j1
        q1 = self.Q if q1 is None else q1
        Jn = J.numpy() if isinstance(J, torch.Tensor) else np.ascontiguousarray(J, dtype=self.np_dtype)
        assert out.is_cuda and out.is_contiguous() and Jn.shape[0] == q1 - q0
        a = self._args(i0, i1, j0, j1, q0, q1, Jn.shape[1], accumulate, 0, 0)
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
            check(self.L.pspace_cuda_project_matrix_hostin(self.h, ctypes.byref(a), Jn.ctypes.data, out.data_ptr()))
        return out

    def projectVectorHost(self, f, k0=0, k1=None, q0=0, q1=None, out=None, accumulate=False):
        k1 = self.N if k1 is None else k1
        q1 = self.Q if q1 is None else q1
        fn = f.numpy() if isinstance(f, torch.Tensor) else np.ascontiguousarray(f, dtype=self.np_dtype)
        m = fn.shape[1]
        if out is None:
            out = np.zeros((k1 - k0, m), dtype=self.np_dtype)
        on = out.numpy() if isinstance(out, torch.Tensor) else out
        a = self._args(k0, k1, 0, 0, q0, q1, m, accumulate, 0, 0)
        check(self.L.pspace_cuda_project_vector_host(self.h, ctypes.byref(a), fn.ctypes.data, on.ctypes.data))
        return out

    def adjointStates(self, v, adj, nnodes, ndvpn, q0=0, q1=None, out=None):
        """(f4 step 1) deterministic states and adjoint at the quadrature points from TACS-interleaved stochastic vectors
        v, adj (device, [nnodes*N*ndvpn]): returns UP [q1-q0][2m], columns [0,m) = u_q, [m,2m) = psi_q."""
        q1 = self.Q if q1 is None else q1
        m = nnodes * ndvpn
        assert v.is_cuda and adj.is_cuda and v.dtype == self.t_dtype and adj.dtype == self.t_dtype
        assert v.is_contiguous() and adj.is_contiguous() and v.numel() == self.N * m and adj.numel() == self.N * m
        if out is None:
            out = torch.empty((q1 - q0, 2 * m), dtype=self.t_dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_adjoint_states(self.h, nnodes, ndvpn, q0, q1, v.data_ptr(), adj.data_ptr(), out.data_ptr(), _stream_ptr()))
        return out

    def adjointProduct(self, G, scale=1.0, q0=0, q1=None, out=None, accumulate=True):
        """(f4 step 3) dfdx[n] (+)= scale * sum_q w_q psi_0(z_q) G[q][n]; G device [q1-q0][dvLen] from the caller's element."""
        q1 = self.Q if q1 is None else q1
        assert G.is_cuda and G.dtype == self.t_dtype and G.is_contiguous() and G.shape[0] == q1 - q0
        dvlen = G.shape[1]
        if out is None:
            out = torch.zeros(dvlen, dtype=self.t_dtype, device=self.device)
        sc = np.array([scale], dtype=self.np_dtype)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_adjoint_product(self.h, dvlen, cabi.np_ptr(sc), q0, q1, G.data_ptr(), out.data_ptr(), int(accumulate),
                                                     _stream_ptr()))
        return out

    def setOption(self, name, value):
        check(self.L.pspace_cuda_set_option(self.h, name.encode(), int(value)))
        return self

    def moments(self, f, q0=0, q1=None):
        """(mean, second, variance) over quadrature points of f (device [q1-q0][m]); TACSStochasticFunction.cpp:209-238."""
        q1 = self.Q if q1 is None else q1
        assert f.is_cuda and f.dtype == self.t_dtype and f.is_contiguous() and f.shape[0] == q1 - q0
        m = f.shape[1]
        out = [torch.empty(m, dtype=self.t_dtype, device=self.device) for _ in range(3)]
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_moments(self.h, q0, q1, m, f.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                             out[2].data_ptr(), _stream_ptr()))
        return tuple(out)

    def pairMask(self, dmapf):
        """mask[i][j] = prod_d(|deg_i[d]-deg_j[d]| <= dmapf[d]) as a device uint8 tensor (nonzero(), TACSStochasticElement.cpp:7-21)."""
        dm = np.ascontiguousarray(dmapf, dtype=np.intc)
        assert dm.shape[0] == self.nvars
        mask = torch.empty((self.N, self.N), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.pspace_cuda_pair_mask(self.h, cabi.np_ptr(dm), mask.data_ptr(), _stream_ptr()))
        return mask

    def launchCount(self):
        return self.L.pspace_cuda_launch_count(self.h)

    def lastPlan(self):
        buf = ctypes.create_string_buffer(512)
        check(self.L.pspace_cuda_last_plan(self.h, buf, 512))
        return buf.value.decode()


# ---- (f2) TACS-interleaved layouts / pair mask -------------------------------------------------------------------
def _cplx(t):
    return int(t.is_complex())


def scatter_vector_tacs(F, nnodes, ndvpn, res=None, accumulate=True):
    """res[n*nsvpn + i*ndvpn + d] (+)= F[i][n*ndvpn + d]; F device [N][nnodes*ndvpn]."""
    N = F.shape[0]
    assert F.is_cuda and F.is_contiguous() and F.shape[1] == nnodes * ndvpn
    if res is None:
        res = torch.zeros(N * nnodes * ndvpn, dtype=F.dtype, device=F.device)
    check(cabi.lib().pspace_cuda_scatter_vector_tacs(N, nnodes, ndvpn, _cplx(F), F.data_ptr(), res.data_ptr(), int(accumulate), _stream_ptr()))
    return res


def gather_vector_tacs(v, N, nnodes, ndvpn):
    """V[k][n*ndvpn + d] = v[n*nsvpn + k*ndvpn + d]."""
    assert v.is_cuda and v.is_contiguous() and v.numel() == N * nnodes * ndvpn
    V = torch.empty((N, nnodes * ndvpn), dtype=v.dtype, device=v.device)
    check(cabi.lib().pspace_cuda_gather_vector_tacs(N, nnodes, ndvpn, _cplx(v), v.data_ptr(), V.data_ptr(), _stream_ptr()))
    return V


def scatter_matrix_tacs(C, nnodes, ndvpn, mat=None, accumulate=True):
    """mat[(ni*nsvpn+i*ndvpn+di)*nsdof + nj*nsvpn+j*ndvpn+dj] (+)= C[i][j][(ni*ndvpn+di)*m + nj*ndvpn+dj]."""
    N = C.shape[0]
    m = nnodes * ndvpn
    assert C.is_cuda and C.is_contiguous() and C.shape[1] == N and C.shape[2] == m * m
    if mat is None:
        mat = torch.zeros((N * m, N * m), dtype=C.dtype, device=C.device)
    check(cabi.lib().pspace_cuda_scatter_matrix_tacs(N, nnodes, ndvpn, _cplx(C), C.data_ptr(), mat.data_ptr(), int(accumulate), _stream_ptr()))
    return mat


def point_quantity_layout(F):
    """out[d*N + i] = F[i][d]  (evalPointQuantity, TACSStochasticElement.cpp:525)."""
    N, m = F.shape
    out = torch.empty(N * m, dtype=F.dtype, device=F.device)
    check(cabi.lib().pspace_cuda_point_quantity_layout(N, m, _cplx(F), F.data_ptr(), out.data_ptr(), _stream_ptr()))
    return out


def fill_synthetic(n, seed, offset=0, device=0):
    """Device tensor of n doubles: splitmix64 uniform(-0.5,0.5) (same stream of numbers as the oracle's)."""
    out = torch.empty(n, dtype=torch.float64, device=torch.device("cuda", device))
    with torch.cuda.device(device):
        check(cabi.lib().pspace_cuda_fill_synthetic(out.data_ptr(), n, seed, offset, _stream_ptr()))
    return out


def fp64_peak(mode=1, ms_target=300.0, device=0):
    v = ctypes.c_double()
    check(cabi.lib().pspace_cuda_fp64_peak(device, mode, ms_target, ctypes.byref(v)))
    return v.value
