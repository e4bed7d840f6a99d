"""Python face of libpspace.so: PyParameterFactory / PyAbstractParameter / PyParameterContainer with the
method names, arguments and return conventions of the reference's Cython wrapper (pspace/PSPACE.pyx:35-131),
implemented with ctypes over the extern-"C" handles of pspace_b200/host/src/host_capi.cpp.

The reference chooses real or complex scalars when the extension is compiled (USE_COMPLEX, setup.py:33-46);
here both host libraries are built and the choice is made at import: environment PSPACE_COMPLEX=1 selects
lib/complex/libpspace.so, otherwise lib/libpspace.so.  `bind(complex_=...)` returns the three classes for an
explicit choice.  Additive methods (not in the reference wrapper): evalBasis, projectVector, projectMatrix.

Everything below initialise*/project*/evalBasis runs on the GPU; there is no CPU implementation to fall back on.
"""
import ctypes
import os
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp, _i = ctypes.c_void_p, ctypes.c_int


def _load(complex_):
    path = os.path.join(_HERE, "lib", "complex" if complex_ else "", "libpspace.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run __graft_entry__.build() (make -C pspace_b200/host/src)")
    L = ctypes.CDLL(path)
    sig = {
        "pspace_host_is_complex": (_i, []),
        "pspace_host_factory_new": (_vp, []),
        "pspace_host_create_normal": (_vp, [_vp, _vp, _vp, _i]),
        "pspace_host_create_uniform": (_vp, [_vp, _vp, _vp, _i]),
        "pspace_host_create_exponential": (_vp, [_vp, _vp, _vp, _i]),
        "pspace_host_param_basis": (None, [_vp, _vp, _i, _vp]),
        "pspace_host_param_quadrature": (None, [_vp, _i, _vp, _vp, _vp]),
        "pspace_host_param_id": (_i, [_vp]),
        "pspace_host_param_max_degree": (_i, [_vp]),
        "pspace_host_container_new": (_vp, [_i, _i]),
        "pspace_host_container_delete": (None, [_vp]),
        "pspace_host_add_parameter": (None, [_vp, _vp]),
        "pspace_host_initialize": (None, [_vp]),
        "pspace_host_initialize_basis": (None, [_vp, _vp]),
        "pspace_host_initialize_quadrature": (None, [_vp, _vp]),
        "pspace_host_num_basis": (_i, [_vp]),
        "pspace_host_num_params": (_i, [_vp]),
        "pspace_host_num_quad": (_i, [_vp]),
        "pspace_host_quadrature": (None, [_vp, _i, _vp, _vp, _vp]),
        "pspace_host_basis": (None, [_vp, _i, _vp, _vp]),
        "pspace_host_basis_param_deg": (None, [_vp, _i, _vp]),
        "pspace_host_basis_param_max_deg": (None, [_vp, _vp]),
        "pspace_host_eval_basis": (None, [_vp, _i, _vp, _vp]),
        "pspace_host_evaluate_expansion": (None, [_vp, _i, _vp, _vp]),
        "pspace_host_project_vector": (None, [_vp, _i, _vp, _vp]),
        "pspace_host_project_matrix": (None, [_vp, _i, _vp, _vp]),
        "pspace_host_project_matrix_window": (None, [_vp, _i, _vp, _i, _i, _i, _i, _vp]),
        "pspace_host_device_context": (_vp, [_vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    assert L.pspace_host_is_complex() == int(complex_)
    return L


def bind(complex_=False):
    """Return a namespace with PyParameterFactory, PyAbstractParameter, PyParameterContainer for one scalar type."""
    L = _load(complex_)
    dtype = np.complex128 if complex_ else np.float64

    def _s(x):  # python scalar -> 1-element array of the scalar type
        return np.array([x], dtype=dtype)

    def _py(a):  # 1-element array -> python scalar (complex in the complex build, float otherwise)
        return complex(a[0]) if complex_ else float(a[0])

    class PyAbstractParameter:
        def __init__(self):
            self.ptr = None

        def basis(self, z, d):
            out = np.zeros(1, dtype=dtype)
            zz = _s(z)
            L.pspace_host_param_basis(self.ptr, zz.ctypes.data, int(d), out.ctypes.data)
            return _py(out)

        def quadrature(self, npoints):   # additive: the wrapper comments this binding out (PSPACE.pxd:14)
            z = np.zeros(npoints, dtype=dtype); y = np.zeros(npoints, dtype=dtype); w = np.zeros(npoints, dtype=dtype)
            L.pspace_host_param_quadrature(self.ptr, int(npoints), z.ctypes.data, y.ctypes.data, w.ctypes.data)
            return z, y, w

    class PyParameterFactory:
        def __init__(self):
            self.ptr = L.pspace_host_factory_new()

        def _wrap(self, p):
            q = PyAbstractParameter()
            q.ptr = p
            return q

        def createNormalParameter(self, mu, sigma, dmax):
            a, b = _s(mu), _s(sigma)
            return self._wrap(L.pspace_host_create_normal(self.ptr, a.ctypes.data, b.ctypes.data, int(dmax)))

        def createUniformParameter(self, a, b, dmax):
            x, y = _s(a), _s(b)
            return self._wrap(L.pspace_host_create_uniform(self.ptr, x.ctypes.data, y.ctypes.data, int(dmax)))

        def createExponentialParameter(self, mu, beta, dmax):
            a, b = _s(mu), _s(beta)
            return self._wrap(L.pspace_host_create_exponential(self.ptr, a.ctypes.data, b.ctypes.data, int(dmax)))

    class PyParameterContainer:
        def __init__(self, basis_type=0, quadrature_type=0):
            self.ptr = L.pspace_host_container_new(int(basis_type), int(quadrature_type))
            self._params = []   # keep the Python parameter objects alive

        def addParameter(self, param):
            self._params.append(param)
            L.pspace_host_add_parameter(self.ptr, param.ptr)

        def basis(self, k, z):
            zz = np.ascontiguousarray(z, dtype=dtype)
            out = np.zeros(1, dtype=dtype)
            L.pspace_host_basis(self.ptr, int(k), zz.ctypes.data, out.ctypes.data)
            return _py(out)

        def quadrature(self, q, with_weight=False):
            n = self.getNumParameters()
            zq = np.zeros(n, dtype=dtype); yq = np.zeros(n, dtype=dtype); wq = np.zeros(1, dtype=dtype)
            L.pspace_host_quadrature(self.ptr, int(q), zq.ctypes.data, yq.ctypes.data, wq.ctypes.data)
            if with_weight:
                return _py(wq), zq, yq
            return zq, yq

        def getNumBasisTerms(self):
            return L.pspace_host_num_basis(self.ptr)

        def getNumParameters(self):
            return L.pspace_host_num_params(self.ptr)

        def getNumQuadraturePoints(self):
            return L.pspace_host_num_quad(self.ptr)

        def getBasisParamDeg(self, k):
            degs = np.zeros(self.getNumParameters(), dtype=np.intc)
            L.pspace_host_basis_param_deg(self.ptr, int(k), degs.ctypes.data)
            return degs

        def getBasisParamMaxDeg(self):
            pmax = np.zeros(self.getNumParameters(), dtype=np.intc)
            L.pspace_host_basis_param_max_deg(self.ptr, pmax.ctypes.data)
            return pmax

        def initialize(self):
            L.pspace_host_initialize(self.ptr)

        def initializeBasis(self, pmax):
            a = np.ascontiguousarray(pmax, dtype=np.intc)
            L.pspace_host_initialize_basis(self.ptr, a.ctypes.data)

        def initializeQuadrature(self, nqpts):
            a = np.ascontiguousarray(nqpts, dtype=np.intc)
            L.pspace_host_initialize_quadrature(self.ptr, a.ctypes.data)

        # ---- additive batched methods (host arrays in/out, GPU inside) ----------------------------------
        def evalBasis(self, z):
            zz = np.ascontiguousarray(z, dtype=dtype).reshape(-1, self.getNumParameters())
            psi = np.zeros((self.getNumBasisTerms(), zz.shape[0]), dtype=dtype)
            L.pspace_host_eval_basis(self.ptr, zz.shape[0], zz.ctypes.data, psi.ctypes.data)
            return psi

        def evaluateExpansion(self, V):
            VV = np.ascontiguousarray(V, dtype=dtype).reshape(self.getNumBasisTerms(), -1)
            U = np.zeros((self.getNumQuadraturePoints(), VV.shape[1]), dtype=dtype)
            L.pspace_host_evaluate_expansion(self.ptr, VV.shape[1], VV.ctypes.data, U.ctypes.data)
            return U

        def adjointStates(self, v, adj, nnodes, ndvpn):
            """(u_q, psi_q) at every quadrature point from TACS-interleaved stochastic vectors: (Q, 2m), m = nnodes*ndvpn."""
            m = nnodes * ndvpn
            vv, aa = np.ascontiguousarray(v, dtype=dtype).ravel(), np.ascontiguousarray(adj, dtype=dtype).ravel()
            assert vv.size == aa.size == self.getNumBasisTerms() * m
            UP = np.zeros((self.getNumQuadraturePoints(), 2 * m), dtype=dtype)
            L.pspace_host_adjoint_states(self.ptr, nnodes, ndvpn, vv.ctypes.data, aa.ctypes.data, UP.ctypes.data)
            return UP

        def adjointProduct(self, G, scale=1.0, dfdx=None):
            """dfdx[n] += scale * sum_q w_q basis(0, z_q) G[q, n]  (addAdjResProduct, TACSStochasticElement.cpp:548-637)."""
            GG = np.ascontiguousarray(G, dtype=dtype).reshape(self.getNumQuadraturePoints(), -1)
            out = np.zeros(GG.shape[1], dtype=dtype) if dfdx is None else dfdx
            sc = np.array([scale], dtype=dtype)
            L.pspace_host_adjoint_product(self.ptr, GG.shape[1], sc.ctypes.data, GG.ctypes.data, out.ctypes.data)
            return out

        def projectVector(self, f):
            ff = np.ascontiguousarray(f, dtype=dtype).reshape(self.getNumQuadraturePoints(), -1)
            F = np.zeros((self.getNumBasisTerms(), ff.shape[1]), dtype=dtype)
            L.pspace_host_project_vector(self.ptr, ff.shape[1], ff.ctypes.data, F.ctypes.data)
            return F

        def projectMatrix(self, J, window=None):
            JJ = np.ascontiguousarray(J, dtype=dtype).reshape(self.getNumQuadraturePoints(), -1)
            N = self.getNumBasisTerms()
            i0, i1, j0, j1 = (0, N, 0, N) if window is None else window
            C = np.zeros((i1 - i0, j1 - j0, JJ.shape[1]), dtype=dtype)
            L.pspace_host_project_matrix_window(self.ptr, JJ.shape[1], JJ.ctypes.data, i0, i1, j0, j1, C.ctypes.data)
            return C

    return types.SimpleNamespace(PyAbstractParameter=PyAbstractParameter, PyParameterFactory=PyParameterFactory,
                                 PyParameterContainer=PyParameterContainer, USE_COMPLEX=bool(complex_), dtype=dtype)


def __getattr__(name):
    """Module-level PyParameterFactory / PyAbstractParameter / PyParameterContainer (lazy: loads libpspace.so)."""
    if name in ("PyAbstractParameter", "PyParameterFactory", "PyParameterContainer"):
        ns = bind(os.environ.get("PSPACE_COMPLEX") == "1")
        globals().update(PyAbstractParameter=ns.PyAbstractParameter, PyParameterFactory=ns.PyParameterFactory,
                         PyParameterContainer=ns.PyParameterContainer)
        return globals()[name]
    raise AttributeError(name)
