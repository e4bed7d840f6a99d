"""CPU, build container only: the restated oracle against the compiled reference itself (oracle/_ref)
on random cases (skipped where oracle/_ref was not built, e.g. no /root/reference)."""
import numpy as np
import pytest

from oracle import harness as H
from tests.util import same_bits

pytestmark = pytest.mark.skipif(not (H.ref_available(False) and H.ref_available(True)),
                                reason="compiled reference (oracle/_ref) not present")
KINDS = ["normal", "uniform", "exponential"]


def _rand_params(rng, nv, cplx, dhi):
    ps = []
    for _ in range(nv):
        k = KINDS[rng.integers(3)]
        a = rng.normal()
        b = abs(rng.normal()) + 0.1
        if k == "uniform":
            b = a + rng.normal()       # a > b is legal (tests/test_pspace.py:9)
        if cplx and rng.random() < 0.5:
            a = a + 1e-30j
        if cplx and rng.random() < 0.3:
            b = b + 1e-30j
        ps.append((k, a, b, int(rng.integers(0, dhi + 1))))
    return ps


@pytest.mark.parametrize("cplx", [False, True])
def test_random_containers_bit_exact(cplx):
    rng = np.random.default_rng(1234 + cplx)
    o, r = H.oracle_api(cplx), H.ref_api(cplx)
    for _ in range(60):
        nv = int(rng.integers(1, 6))
        bt = int(rng.integers(0, 2))
        ps = _rand_params(rng, nv, cplx, 6 if nv <= 3 else 3)
        # Laguerre 12/13/17/18-point rules are broken in the reference; dmax <= 6 stays clear of them
        co = H.CpuContainer(o, ps, bt).initialize()
        cr = H.CpuContainer(r, ps, bt).initialize()
        assert (co.N, co.Q) == (cr.N, cr.Q)
        assert np.array_equal(co.degrees(), cr.degrees())
        for a, b in zip(co.grid(), cr.grid()):
            assert same_bits(a, b)
        Z = co.grid()[0]
        sel = rng.choice(co.Q, min(co.Q, 40), replace=False)
        assert same_bits(co.basis_at(Z[sel]), cr.basis_at(Z[sel]))
        zs = rng.normal(size=(6, nv)).astype(co.dtype) + (1e-30j if cplx else 0.0)
        assert same_bits(co.basis_at(zs), cr.basis_at(zs))


@pytest.mark.parametrize("cplx", [False, True])
def test_projection_loops_bit_exact(cplx):
    rng = np.random.default_rng(99)
    o, r = H.oracle_api(cplx), H.ref_api(cplx)
    ps = [("normal", 0.3, 1.2, 2), ("exponential", 0.0, 0.5, 3), ("uniform", -1.0, 2.0, 2)]
    co = H.CpuContainer(o, ps, 1).initialize()
    cr = H.CpuContainer(r, ps, 1).initialize()
    m = 3
    f = rng.normal(size=(co.Q, m)).astype(co.dtype)
    J = rng.normal(size=(co.Q, m * m)).astype(co.dtype)
    if cplx:
        f = f + 1j * rng.normal(size=f.shape)
        J = J + 1j * rng.normal(size=J.shape)
    assert same_bits(co.project_vector(f), cr.project_vector(f))
    assert same_bits(co.project_matrix(J, nthreads=3), cr.project_matrix(J, nthreads=2))
    assert same_bits(co.project_matrix(J, 2, 5, 1, 7), cr.project_matrix(J, 2, 5, 1, 7))
    V = rng.normal(size=(co.N, m)).astype(co.dtype) * (1 + (0.5j if cplx else 0))
    assert same_bits(co.evaluate_expansion(V), cr.evaluate_expansion(V))
    assert same_bits(co.evaluate_expansion(V[1:4], 1, 4, 5, 17), cr.evaluate_expansion(V[1:4], 1, 4, 5, 17))
    # q-range restriction used by the sharded tests: ranges sum to the whole (same order inside a range)
    half = co.Q // 2
    Ca = co.project_matrix_qrange(J[:half], 0, half)
    Cb = co.project_matrix_qrange(J[half:], half, co.Q)
    assert np.max(np.abs(Ca + Cb - cr.project_matrix(J))) < 1e-14


@pytest.mark.parametrize("cplx", [False, True])
def test_moments_and_adjoint_product_loops_bit_exact(cplx):
    """(f3) getFunctionValue moments and (f4) addAdjResProduct: the oracle's restated loops against the same loops driven by the
    compiled reference's quadrature() / basis() (oracle/ref_shim.cpp), with the same synthetic deterministic element."""
    rng = np.random.default_rng(123)
    o, r = H.oracle_api(cplx), H.ref_api(cplx)
    ps = [("normal", 0.3 + (1e-30j if cplx else 0), 1.2, 2), ("exponential", 0.0, 0.5, 3), ("uniform", -1.0, 2.0, 2)]
    co = H.CpuContainer(o, ps, 1).initialize()
    cr = H.CpuContainer(r, ps, 1).initialize()
    f = rng.normal(size=(co.Q, 5)).astype(co.dtype)
    if cplx:
        f = f + 1j * rng.normal(size=f.shape)
    for (q0, q1) in ((0, co.Q), (3, 20)):
        for a, b in zip(co.moments(f[q0:q1], q0, q1), cr.moments(f[q0:q1], q0, q1)):
            assert same_bits(a, b)
    nnodes, ndvpn, dvlen = 3, 2, 7
    v = rng.normal(size=nnodes * co.N * ndvpn).astype(co.dtype)
    adj = rng.normal(size=v.size).astype(co.dtype)
    if cplx:
        v = v + 1j * rng.normal(size=v.size)
        adj = adj - 0.5j * rng.normal(size=v.size)
    elem = o.synthetic_adjres_element                    # one element implementation for both arms (a C callback)
    a = co.adj_res_product(v, adj, nnodes, ndvpn, dvlen, 0.7, element=elem)
    b = cr.adj_res_product(v, adj, nnodes, ndvpn, dvlen, 0.7, element=elem)
    assert same_bits(a, b) and np.all(np.isfinite(a.view(np.float64))) and np.any(a != 0)
    base = rng.normal(size=dvlen).astype(co.dtype)
    assert same_bits(co.adj_res_product(v, adj, nnodes, ndvpn, dvlen, 1.0, element=elem, dfdx=base.copy()),
                     cr.adj_res_product(v, adj, nnodes, ndvpn, dvlen, 1.0, element=elem, dfdx=base.copy()))


def test_pair_filter_equals_the_reference_library_function():
    """(f2) nonzero() filter: the oracle's restatement against BasisHelper::sparse of the compiled reference, on the degree
    tables of tensor and complete bases with uniform and per-variable integrand degrees."""
    o = H.oracle_api(False)
    for bt, dmax in ((0, [3, 2, 4]), (1, [4, 4, 4, 4]), (1, [2, 5])):
        ps = [("normal", 0.0, 1.0, d) for d in dmax]
        deg = H.CpuContainer(o, ps, bt).initialize().degrees()
        for dmapf in ([0] * len(dmax), [1] * len(dmax), [3] * len(dmax), list(range(len(dmax)))):
            a, b = H.pair_mask(deg, dmapf), H.pair_mask(deg, dmapf, reference=True)
            assert np.array_equal(a, b) and a.any() and (dmapf[0] >= max(dmax) or not a.all())


def test_high_degree_single_variable_bit_exact():
    """d up to 15: factorial loop (n > 10), deep naive recursions, Legendre pow sum."""
    for cplx in (False, True):
        o, r = H.oracle_api(cplx), H.ref_api(cplx)
        for kind in KINDS:
            ps = [(kind, 0.2, 1.5, 15)]
            co = H.CpuContainer(o, ps, 0).initialize_basis([15])
            cr = H.CpuContainer(r, ps, 0).initialize_basis([15])
            zs = np.linspace(0.05, 0.95, 7).astype(co.dtype).reshape(-1, 1)
            assert same_bits(co.basis_at(zs), cr.basis_at(zs)), (kind, cplx)
