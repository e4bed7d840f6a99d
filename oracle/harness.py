"""TEST INFRASTRUCTURE — ctypes driver shared by the restated oracle (libpspace_oracle.so) and the
compiled reference shim (oracle/_ref/libpspace_ref{,_z}.so).  Both export the same C API (see
pspace_oracle.h / ref_shim.cpp), so `CpuContainer(lib=...)` drives either.

Importers allowed: tests/, tools/gen_golden.py, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference legs).  Never imported by pspace_b200/.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KIND = {"normal": 0, "uniform": 1, "exponential": 2}

_vp, _i, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong


def build_oracle(quiet=True):
    """(Re)build libpspace_oracle.so and, if /root/reference is present, oracle/_ref."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=out)
    subprocess.call(["make", "-C", HERE, "ref"], stdout=out, stderr=out)


class _Api:
    """Binds one (library, prefix, suffix) triple."""

    def __init__(self, path, prefix, suffix, is_complex, kind):
        self.lib = ctypes.CDLL(path)
        self.is_complex = is_complex
        self.kind = kind  # "port" (restated oracle) or "reference" (compiled reference)
        self.path = path
        sig = {
            "create": (_vp, [_i, _vp, _vp, _vp, _vp, _i]),
            "initialize": (None, [_vp]),
            "initialize_basis": (None, [_vp, _vp]),
            "initialize_quadrature": (None, [_vp, _vp]),
            "num_basis": (_i, [_vp]),
            "num_quad": (_i, [_vp]),
            "num_params": (_i, [_vp]),
            "degrees": (None, [_vp, _vp]),
            "max_degrees": (None, [_vp, _vp]),
            "grid": (None, [_vp, _vp, _vp, _vp]),
            "basis_at": (None, [_vp, _i, _vp, _vp]),
            "param_quadrature": (None, [_vp, _i, _i, _vp, _vp, _vp]),
            "param_basis": (None, [_vp, _i, _i, _vp, _i, _vp]),
            "project_vector": (None, [_vp, _i, _vp, _i, _i, _vp]),
            "project_matrix": (None, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i]),
            "evaluate_expansion": (None, [_vp, _i, _vp, _i, _i, _ll, _ll, _vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(self.lib, prefix + name + suffix)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)
        fn = getattr(self.lib, prefix + "moments" + suffix)
        fn.restype, fn.argtypes = None, [_vp, _i, _vp, _ll, _ll, _vp, _vp, _vp]
        self.moments = fn
        fn = getattr(self.lib, prefix + "adj_res_product" + suffix)
        fn.restype, fn.argtypes = None, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]
        self.adj_res_product = fn
        if prefix == "oracle_":
            fn = getattr(self.lib, "oracle_project_matrix_qrange" + suffix)
            fn.restype, fn.argtypes = None, [_vp, _i, _vp, _i, _i, _i, _i, _ll, _ll, _vp, _i]
            self.project_matrix_qrange = fn
            self.synthetic_adjres_element = ctypes.cast(getattr(self.lib, "oracle_synthetic_adjres_element" + suffix), _vp)
            fs = self.lib.oracle_fill_synthetic
            fs.restype, fs.argtypes = None, [_vp, _ll, ctypes.c_ulonglong, ctypes.c_ulonglong]
            self.fill_synthetic = fs


def oracle_api(complex_=False):
    path = os.path.join(HERE, "libpspace_oracle.so")
    if not os.path.exists(path):
        build_oracle()
    return _Api(path, "oracle_", "_z" if complex_ else "", complex_, "port")


def ref_available(complex_=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libpspace_ref_z.so" if complex_ else "libpspace_ref.so"))


def ref_api(complex_=False):
    path = os.path.join(HERE, "_ref", "libpspace_ref_z.so" if complex_ else "libpspace_ref.so")
    return _Api(path, "refshim_", "", complex_, "reference")


def _ptr(a):
    return a.ctypes.data


class CpuContainer:
    """A ParameterContainer living in `api` (oracle or compiled reference).

    params: list of (kind, p1, p2, dmax) with kind in {"normal","uniform","exponential"}.
    """

    def __init__(self, api, params, basis_type=0):
        self.api = api
        self.dtype = np.complex128 if api.is_complex else np.float64
        self.nvars = len(params)
        kinds = np.array([KIND[p[0]] if isinstance(p[0], str) else p[0] for p in params], dtype=np.intc)
        p1 = np.array([p[1] for p in params], dtype=self.dtype)
        p2 = np.array([p[2] for p in params], dtype=self.dtype)
        dmax = np.array([p[3] for p in params], dtype=np.intc)
        self.h = api.create(self.nvars, _ptr(kinds), _ptr(p1), _ptr(p2), _ptr(dmax), basis_type)

    # -- initialisation -------------------------------------------------------------------------
    def initialize(self):
        self.api.initialize(self.h)
        return self

    def initialize_basis(self, pmax):
        a = np.ascontiguousarray(pmax, dtype=np.intc)
        self.api.initialize_basis(self.h, _ptr(a))
        return self

    def initialize_quadrature(self, nq):
        a = np.ascontiguousarray(nq, dtype=np.intc)
        self.api.initialize_quadrature(self.h, _ptr(a))
        return self

    # -- accessors ------------------------------------------------------------------------------
    @property
    def N(self):
        return self.api.num_basis(self.h)

    @property
    def Q(self):
        return self.api.num_quad(self.h)

    def degrees(self):
        out = np.zeros((self.N, self.nvars), dtype=np.intc)
        self.api.degrees(self.h, _ptr(out))
        return out

    def grid(self):
        Q = self.Q
        Z = np.zeros((Q, self.nvars), dtype=self.dtype)
        Y = np.zeros((Q, self.nvars), dtype=self.dtype)
        W = np.zeros(Q, dtype=self.dtype)
        self.api.grid(self.h, _ptr(Z), _ptr(Y), _ptr(W))
        return Z, Y, W

    def basis_at(self, Z):
        Z = np.ascontiguousarray(Z, dtype=self.dtype).reshape(-1, self.nvars)
        out = np.zeros((self.N, Z.shape[0]), dtype=self.dtype)
        self.api.basis_at(self.h, Z.shape[0], _ptr(Z), _ptr(out))
        return out

    def param_quadrature(self, pid, n):
        z = np.zeros(n + 8, dtype=self.dtype)  # slack: the reference's Laguerre 12/17 branches overrun
        y = np.zeros(n + 8, dtype=self.dtype)
        w = np.zeros(n + 8, dtype=self.dtype)
        self.api.param_quadrature(self.h, pid, n, _ptr(z), _ptr(y), _ptr(w))
        return z[:n].copy(), y[:n].copy(), w[:n].copy()

    def param_basis(self, pid, z, d):
        z = np.ascontiguousarray(z, dtype=self.dtype).ravel()
        out = np.zeros(z.shape[0], dtype=self.dtype)
        self.api.param_basis(self.h, pid, z.shape[0], _ptr(z), d, _ptr(out))
        return out

    # -- projections (reference loop order) -----------------------------------------------------
    def project_vector(self, f, k0=0, k1=None):
        f = np.ascontiguousarray(f, dtype=self.dtype).reshape(self.Q, -1)
        k1 = self.N if k1 is None else k1
        F = np.zeros((k1 - k0, f.shape[1]), dtype=self.dtype)
        self.api.project_vector(self.h, f.shape[1], _ptr(f), k0, k1, _ptr(F))
        return F

    def project_matrix(self, J, i0=0, i1=None, j0=0, j1=None, nthreads=1):
        J = np.ascontiguousarray(J, dtype=self.dtype).reshape(self.Q, -1)
        i1 = self.N if i1 is None else i1
        j1 = self.N if j1 is None else j1
        C = np.zeros((i1 - i0, j1 - j0, J.shape[1]), dtype=self.dtype)
        self.api.project_matrix(self.h, J.shape[1], _ptr(J), i0, i1, j0, j1, _ptr(C), nthreads)
        return C

    def evaluate_expansion(self, V, k0=0, k1=None, q0=0, q1=None):
        """U[q][c] = sum_k psi_k(z_q) V[k][c]  (getDeterministicStates)."""
        k1 = self.N if k1 is None else k1
        q1 = self.Q if q1 is None else q1
        V = np.ascontiguousarray(V, dtype=self.dtype).reshape(k1 - k0, -1)
        U = np.zeros((q1 - q0, V.shape[1]), dtype=self.dtype)
        self.api.evaluate_expansion(self.h, V.shape[1], _ptr(V), k0, k1, q0, q1, _ptr(U))
        return U

    def adj_res_product(self, v, adj, nnodes, ndvpn, dvlen, scale=1.0, element=None, dfdx=None):
        """(f4) TACSStochasticElement::addAdjResProduct (:548-637) over the TACS-interleaved stochastic state and adjoint
        vectors v, adj [nnodes * N * ndvpn]; `element` = C callback of the deterministic element (default: the oracle's
        synthetic one).  Adds into and returns dfdx[dvlen]."""
        v = np.ascontiguousarray(v, dtype=self.dtype)
        adj = np.ascontiguousarray(adj, dtype=self.dtype)
        assert v.size == nnodes * self.N * ndvpn and adj.size == v.size
        if dfdx is None:
            dfdx = np.zeros(dvlen, dtype=self.dtype)
        sc = np.array([scale], dtype=self.dtype)
        self.api.adj_res_product(self.h, nnodes, ndvpn, dvlen, _ptr(sc), _ptr(v), _ptr(adj),
                                 element if element is not None else self.api.synthetic_adjres_element, None, _ptr(dfdx))
        return dfdx

    def moments(self, f, q0=0, q1=None):
        q1 = self.Q if q1 is None else q1
        f = np.ascontiguousarray(f, dtype=self.dtype).reshape(q1 - q0, -1)
        out = [np.zeros(f.shape[1], dtype=self.dtype) for _ in range(3)]
        self.api.moments(self.h, f.shape[1], _ptr(f), q0, q1, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]))
        return tuple(out)

    def project_matrix_qrange(self, Jslice, q0, q1, i0=0, i1=None, j0=0, j1=None, nthreads=1):
        Jslice = np.ascontiguousarray(Jslice, dtype=self.dtype).reshape(q1 - q0, -1)
        i1 = self.N if i1 is None else i1
        j1 = self.N if j1 is None else j1
        C = np.zeros((i1 - i0, j1 - j0, Jslice.shape[1]), dtype=self.dtype)
        self.api.project_matrix_qrange(self.h, Jslice.shape[1], _ptr(Jslice), i0, i1, j0, j1, q0, q1, _ptr(C), nthreads)
        return C


def synthetic(n, seed, offset=0):
    """uniform(-0.5,0.5) doubles from splitmix64(seed ^ (offset+i)) — SURVEY.md §8d."""
    out = np.zeros(n, dtype=np.float64)
    oracle_api(False).fill_synthetic(_ptr(out), n, seed, offset)
    return out


def tacs_scatter_vector(F, nnodes, ndvpn, res=None):
    """res[n*nsvpn + i*ndvpn + d] += F[i][n*ndvpn + d]  (reference caller loops, TACSStochasticElement.cpp:316-322)."""
    F = np.ascontiguousarray(F)
    N = F.shape[0]
    width = 2 if np.iscomplexobj(F) else 1
    res = np.zeros(N * nnodes * ndvpn, dtype=F.dtype) if res is None else res
    lib = oracle_api(False).lib
    lib.oracle_scatter_vector_tacs.argtypes = [_i, _i, _i, _i, _vp, _vp]
    lib.oracle_scatter_vector_tacs(N, nnodes, ndvpn, width, _ptr(F), _ptr(res))
    return res


def tacs_scatter_matrix(C, nnodes, ndvpn, mat=None):
    """mat[...] += C[i][j] blocks in the stochastic-element layout (TACSStochasticElement.cpp:417-432)."""
    C = np.ascontiguousarray(C)
    N = C.shape[0]
    width = 2 if np.iscomplexobj(C) else 1
    nsdof = N * nnodes * ndvpn
    mat = np.zeros((nsdof, nsdof), dtype=C.dtype) if mat is None else mat
    lib = oracle_api(False).lib
    lib.oracle_scatter_matrix_tacs.argtypes = [_i, _i, _i, _i, _vp, _vp]
    lib.oracle_scatter_matrix_tacs(N, nnodes, ndvpn, width, _ptr(C), _ptr(mat))
    return mat


def pair_mask(deg, dmapf, reference=False):
    """nonzero() pair filter on a degree table; reference=True: through the compiled reference's BasisHelper::sparse."""
    deg = np.ascontiguousarray(deg, dtype=np.intc)
    dm = np.ascontiguousarray(dmapf, dtype=np.intc)
    N, nv = deg.shape
    mask = np.zeros((N, N), dtype=np.uint8)
    fn = ref_api(False).lib.refshim_pair_mask if reference else oracle_api(False).lib.oracle_pair_mask
    fn.restype, fn.argtypes = None, [_i, _i, _vp, _vp, _vp]
    fn(N, nv, _ptr(deg), _ptr(dm), _ptr(mask))
    return mask
