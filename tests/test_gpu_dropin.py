"""GPU tests of the drop-in boundary: the C++ host library (ParameterFactory / ParameterContainer, reference
API) driven (a) through pspace_b200.PSPACE, the Python mirror of the reference's Cython wrapper, and (b) through
the reference's OWN unchanged PSPACE.pyx compiled against this repo's headers and libpspace.so
(tools/build_dropin.py -> oracle/_ref/dropin, present when built in the container that has /root/reference)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.util import ROOT, golden_params, load_golden, normwise, same_bits  # noqa: E402


def _container(ns, params, basis_type):
    f = ns.PyParameterFactory()
    make = {"normal": f.createNormalParameter, "uniform": f.createUniformParameter, "exponential": f.createExponentialParameter}
    pc = ns.PyParameterContainer(basis_type)
    for kind, a, b, d in params:
        pc.addParameter(make[kind](a, b, d))
    pc.initialize()
    return pc


# ---- the reference's own tests/test_pspace.py:6-43, restated against the Python mirror ---------------------
def _test_pspace_container(ns):
    f = ns.PyParameterFactory()
    y1 = f.createNormalParameter(mu=1.0, sigma=0.1, dmax=2)
    y2 = f.createUniformParameter(a=1.0, b=0.1, dmax=2)
    y3 = f.createExponentialParameter(mu=1.0, beta=0.1, dmax=2)
    pc = ns.PyParameterContainer(1)
    for y in (y1, y2, y3):
        pc.addParameter(y)
    pc.initialize()
    return pc


@pytest.mark.parametrize("cplx", [False, True])
def test_reference_wrapper_contract(cplx):
    from pspace_b200 import PSPACE
    pc = _test_pspace_container(PSPACE.bind(cplx))
    zq, yq = pc.quadrature(0)
    assert isinstance(zq, np.ndarray) and isinstance(yq, np.ndarray)
    assert zq.shape == yq.shape == (pc.getNumParameters(),)
    wq, zq_w, yq_w = pc.quadrature(0, with_weight=True)
    assert np.isscalar(wq)
    assert np.allclose(zq_w, zq) and np.allclose(yq_w, yq)
    assert np.isscalar(pc.basis(0, zq))
    assert pc.getNumBasisTerms() > 0
    # and the numbers, against the golden vectors of the same container generated from the compiled reference
    g = load_golden("test_pspace_complex" if cplx else "test_pspace_real")
    assert pc.getNumBasisTerms() == int(g["N"]) and pc.getNumQuadraturePoints() == int(g["Q"])
    for k in range(pc.getNumBasisTerms()):
        assert np.array_equal(pc.getBasisParamDeg(k), g["degrees"][k])
    assert np.array_equal(pc.getBasisParamMaxDeg(), [2, 2, 2])
    for q in range(pc.getNumQuadraturePoints()):
        w, z, y = pc.quadrature(q, with_weight=True)
        if cplx:
            assert abs(w - g["W"][q]) < 1e-15 and normwise(z, g["Z"][q]) < 1e-15 and normwise(y, g["Y"][q]) < 1e-15
        else:
            assert w == g["W"][q] and same_bits(z, g["Z"][q]) and same_bits(y, g["Y"][q])


@pytest.mark.parametrize("name", ["cfg1_real", "cfg1_complex", "cfg2_real"])
def test_host_library_against_golden(name):
    from oracle import harness as H
    from pspace_b200 import PSPACE
    g = load_golden(name)
    cplx = bool(int(g["is_complex"]))
    ns = PSPACE.bind(cplx)
    pc = _container(ns, golden_params(g), int(g["basis_type"]))
    N, Q = pc.getNumBasisTerms(), pc.getNumQuadraturePoints()
    assert (N, Q) == (int(g["N"]), int(g["Q"]))
    # per-call basis(k, z) at grid points and arbitrary points
    for col, q in enumerate(g["qsel"][:6]):
        z, _ = pc.quadrature(int(q))
        for k in (0, 1, N // 2, N - 1):
            v = pc.basis(k, z)
            ref = g["psi_sel"][k, col]
            assert (abs(v - ref) <= 1e-14 * max(1.0, abs(ref))) if cplx else (v == ref)
    psi = pc.evalBasis(g["zarb"])
    assert normwise(psi, g["psi_arb"]) < 1e-13
    for p_, z in enumerate(g["zarb"][:2]):
        assert abs(pc.basis(3, z) - g["psi_arb"][3, p_]) <= 1e-13 * max(1.0, abs(g["psi_arb"][3, p_]))
    # batched projections through ParameterContainer::projectVector / projectMatrixWindow (host buffers)
    m, seed = int(g["m"]), int(g["seed"])
    syn = (lambda n, s: H.synthetic(2 * n, s).view(np.complex128)) if cplx else H.synthetic
    f = syn(Q * m, seed).reshape(Q, m)
    J = syn(Q * m * m, seed + 1000).reshape(Q, m * m)
    assert normwise(pc.projectVector(f), g["F"]) < 1e-12
    win = tuple(int(v) for v in g["win"])
    assert normwise(pc.projectMatrix(J, window=win), g["C"]) < 1e-12
    # (f4) ParameterContainer::adjointStates / adjointProduct (host buffers) against the oracle's restatement of
    # TACSStochasticElement::addAdjResProduct with the oracle's synthetic deterministic element
    o = H.CpuContainer(H.oracle_api(cplx), golden_params(g), int(g["basis_type"])).initialize()
    nnodes, ndvpn, dvlen, nv = 2, 2, 3, len(golden_params(g))
    mm = nnodes * ndvpn
    v, adj = syn(N * mm, seed + 5), syn(N * mm, seed + 6)
    UP = pc.adjointStates(v, adj, nnodes, ndvpn)
    U, P = UP[:, :mm], UP[:, mm:]
    _, Yo, _ = o.grid()
    G = np.stack([(P * U).sum(1) * (1.0 + 0.1 * n) + Yo[:, n % nv] * U.sum(1) for n in range(dvlen)], axis=1)
    assert normwise(pc.adjointProduct(G, 0.5), o.adj_res_product(v, adj, nnodes, ndvpn, dvlen, 0.5)) < 1e-12


DROPIN_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import pspace.PSPACE as uq                      # the reference's unchanged wrapper, linked to OUR libpspace.so
from tests.util import load_golden
cplx = sys.argv[3] == "complex"
f = uq.PyParameterFactory()
pc = uq.PyParameterContainer(1)
for y in (f.createNormalParameter(mu=1.0, sigma=0.1, dmax=2), f.createUniformParameter(a=1.0, b=0.1, dmax=2),
          f.createExponentialParameter(mu=1.0, beta=0.1, dmax=2)):
    pc.addParameter(y)
pc.initialize()
zq, yq = pc.quadrature(0)
assert isinstance(zq, np.ndarray) and zq.shape == yq.shape == (pc.getNumParameters(),)
wq, zq_w, yq_w = pc.quadrature(0, with_weight=True)
assert np.isscalar(wq) and np.allclose(zq_w, zq) and np.allclose(yq_w, yq)
assert np.isscalar(pc.basis(0, zq)) and pc.getNumBasisTerms() > 0
g = load_golden("test_pspace_complex" if cplx else "test_pspace_real")
assert pc.getNumBasisTerms() == int(g["N"]) and pc.getNumQuadraturePoints() == int(g["Q"])
for k in range(pc.getNumBasisTerms()):
    assert np.array_equal(pc.getBasisParamDeg(k), g["degrees"][k])
for q in range(pc.getNumQuadraturePoints()):
    w, z, y = pc.quadrature(q, with_weight=True)
    assert abs(w - g["W"][q]) < 1e-15 and np.max(np.abs(z - g["Z"][q])) < 1e-14 and np.max(np.abs(y - g["Y"][q])) < 1e-14
for col, q in enumerate(g["qsel"]):
    z, _ = pc.quadrature(int(q))
    for k in range(pc.getNumBasisTerms()):
        assert abs(pc.basis(k, z) - g["psi_sel"][k, col]) < 1e-13 * max(1.0, abs(g["psi_sel"][k, col]))
print("DROPIN-OK")
'''


@pytest.mark.parametrize("tag", ["real", "complex"])
def test_unchanged_reference_cython_wrapper_runs_on_this_library(tag):
    pkg = os.path.join(ROOT, "oracle", "_ref", "dropin", tag)
    if not any(f.startswith("PSPACE") and f.endswith(".so") for f in os.listdir(os.path.join(pkg, "pspace"))
               ) if os.path.isdir(os.path.join(pkg, "pspace")) else True:
        pytest.skip("oracle/_ref/dropin not built (tools/build_dropin.py needs /root/reference)")
    r = subprocess.run([sys.executable, "-c", DROPIN_SCRIPT, pkg, ROOT, tag], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout + r.stderr


ADDITIVE_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import pspace.PSPACE as uq             # the reference wrapper + the additive binding of integration/PSPACE_additive.*.inc
from oracle import harness as H
cplx = sys.argv[3] == "complex"
ps = [("normal", 1.0 + (1e-30j if cplx else 0.0), 0.1, 3), ("uniform", 1.0, 2.0, 2), ("exponential", 1.0, 0.1, 2)]
f = uq.PyParameterFactory()
pc = uq.PyParameterContainer(1)
pc.addParameter(f.createNormalParameter(mu=ps[0][1], sigma=0.1, dmax=3))
pc.addParameter(f.createUniformParameter(a=1.0, b=2.0, dmax=2))
pc.addParameter(f.createExponentialParameter(mu=1.0, beta=0.1, dmax=2))
pc.initialize()
o = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()
N, Q = pc.getNumBasisTerms(), pc.getNumQuadraturePoints()
assert (N, Q) == (o.N, o.Q)
rng = np.random.default_rng(3)
def rnd(*shape):
    a = rng.normal(size=shape).astype(o.dtype)
    return a + 1j * rng.normal(size=shape) if cplx else a
nw = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
J, fv, V = rnd(Q, 16), rnd(Q, 4), rnd(N, 4)
C = pc.projectMatrix(J)
assert C.shape == (N, N, 16) and nw(C, o.project_matrix(J, nthreads=4)) < 1e-12
Cw = pc.projectMatrix(J, window=(2, 7, 0, N))
assert nw(Cw, o.project_matrix(J, 2, 7, 0, N)) < 1e-12
assert nw(pc.projectVector(fv), o.project_vector(fv)) < 1e-12
U = pc.evaluateExpansion(V)
assert U.shape == (Q, 4) and nw(U, o.evaluate_expansion(V)) < 1e-12
Z, _, _ = o.grid()
assert nw(pc.evalBasis(Z[:9]), o.basis_at(Z[:9])) < 1e-13
print("ADDITIVE-OK")
'''


@pytest.mark.parametrize("tag", ["real", "complex"])
def test_additive_cython_binding_reaches_the_gpu_projections(tag):
    """The reference's PSPACE.pyx / PSPACE.pxd with the additive binding applied (integration/PSPACE_additive.{pxd,pyx}.inc,
    the patch INTEGRATION.md carries; built by tools/build_dropin.py from a scratch copy): a user of the reference's own
    wrapper calls projectMatrix / projectVector / evaluateExpansion / evalBasis and gets the GPU path, checked against
    the oracle."""
    pkg = os.path.join(ROOT, "oracle", "_ref", "dropin_additive", tag)
    if not os.path.isdir(os.path.join(pkg, "pspace")) or not any(f.startswith("PSPACE") and f.endswith(".so") for f in os.listdir(os.path.join(pkg, "pspace"))):
        pytest.skip("oracle/_ref/dropin_additive not built (tools/build_dropin.py needs /root/reference)")
    r = subprocess.run([sys.executable, "-c", ADDITIVE_SCRIPT, pkg, ROOT, tag], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ADDITIVE-OK" in r.stdout, r.stdout + r.stderr
