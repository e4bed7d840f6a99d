"""CPU: the restated oracle (oracle/pspace_oracle.c) against the golden vectors generated from the
compiled, unmodified reference (tools/gen_golden.py).  Bit-exact everywhere: integer tables, 1-D
rules, grids, basis values AND the projected F / C (the oracle keeps the reference's loop order)."""
import numpy as np
import pytest

from oracle import harness as H
from tests.util import golden_params, load_golden, same_bits

CONTAINER_CASES = ["cfg1_real", "cfg1_complex", "test_pspace_real", "test_pspace_complex", "cfg2_real",
                   "cfg2_complex", "cfg3_real", "highdeg_real", "highdeg_complex"]


def _synth(n, seed, cplx):
    return H.synthetic(2 * n, seed).view(np.complex128) if cplx else H.synthetic(n, seed)


@pytest.mark.parametrize("name", CONTAINER_CASES)
def test_container_case_bit_exact(name):
    g = load_golden(name)
    cplx = bool(int(g["is_complex"]))
    c = H.CpuContainer(H.oracle_api(cplx), golden_params(g), int(g["basis_type"])).initialize()
    assert c.N == int(g["N"]) and c.Q == int(g["Q"])
    assert np.array_equal(c.degrees(), g["degrees"])
    Z, Y, W = c.grid()
    assert same_bits(W, g["W"])
    if "Z" in g.files:
        assert same_bits(Z, g["Z"]) and same_bits(Y, g["Y"])
    else:
        assert same_bits(Z[g["qsel"]], g["Z_sel"]) and same_bits(Y[g["qsel"]], g["Y_sel"])
    assert same_bits(c.basis_at(Z[g["qsel"]]), g["psi_sel"])
    assert same_bits(c.basis_at(g["zarb"]), g["psi_arb"])
    m, seed, Q = int(g["m"]), int(g["seed"]), c.Q
    f = _synth(Q * m, seed, cplx).reshape(Q, m)
    J = _synth(Q * m * m, seed + 1000, cplx).reshape(Q, m * m)
    assert same_bits(c.project_vector(f), g["F"])
    i0, i1, j0, j1 = (int(v) for v in g["win"])
    assert same_bits(c.project_matrix(J, i0, i1, j0, j1, nthreads=4), g["C"])


def test_basis_degree_tables_bit_exact():
    g = load_golden("basis_degrees")
    api = H.oracle_api(False)
    for idx in range(int(g["ncases"])):
        pmax = g[f"pmax_{idx}"]
        ps = [("normal", 0.0, 1.0, int(p)) for p in pmax]
        c = H.CpuContainer(api, ps, int(g[f"bt_{idx}"])).initialize_basis(pmax)
        assert np.array_equal(c.degrees(), g[f"deg_{idx}"]), (pmax, int(g[f"bt_{idx}"]))


@pytest.mark.parametrize("cplx", [False, True])
def test_one_dimensional_rules_and_polynomials_bit_exact(cplx):
    g = load_golden("onedim_z" if cplx else "onedim")
    ps = golden_params(g)
    if not cplx:
        ps = [(k, float(np.real(a)), float(np.real(b)), d) for k, a, b, d in ps]
    c = H.CpuContainer(H.oracle_api(cplx), ps, 0)
    nrules = 0
    for pid in range(len(ps)):
        for n in range(1, 21):
            key = f"rule_{pid}_{n}"
            if key not in g.files:
                continue
            z, y, w = c.param_quadrature(pid, n)
            assert same_bits(np.stack([z, y, w]), g[key]), key
            nrules += 1
        zs = g[f"z_{pid}"]
        for d in range(13):
            assert same_bits(c.param_basis(pid, zs, d), g[f"poly_{pid}"][d]), (pid, d)
    assert nrules == 20 + 20 + 16 + 20


def test_survey_known_answers():
    """KATs recorded from the real reference during the survey."""
    ps = [("normal", 1.0, 0.1, 3), ("uniform", 1.0, 2.0, 3), ("exponential", 1.0, 0.1, 3)]
    c = H.CpuContainer(H.oracle_api(False), ps, 0).initialize()
    assert (c.N, c.Q) == (64, 64)
    Z, Y, W = c.grid()
    assert W[5] == 0.005346575540382012
    assert W[37] == 0.05292564159227238
    assert tuple(Z[37]) == (0.741963784302726, 0.33000947820757176, 1.7457611011583465)
    assert tuple(Y[37]) == (1.0741963784302726, 1.3300094782075718, 1.1745761101158347)
    assert tuple(c.degrees()[17]) == (2, 0, 1)
    assert c.basis_at(Z[37])[17, 0] == 0.2370306539663418
    f = (Y[:, 0] * Y[:, 1] * Y[:, 2]).reshape(-1, 1)
    F = c.project_vector(f)[:, 0]
    assert F[0] == 1.65 and F[1] == -0.14999999999999986 and F[2] == 0.31754264805429416
    assert F[3] == 0.16500000000000026
    assert abs(np.sum(np.abs(F)) - 2.3610511776651575) < 1e-14
    psi = c.basis_at(Z)
    G = (psi * W) @ psi.T
    assert np.max(np.abs(G - np.eye(64))) < 1e-14      # orthonormality, tests/test.py:131


def test_main_cpp_complete_43_has_14_terms():
    c = H.CpuContainer(H.oracle_api(False), [("normal", -4.0, 0.5, 4), ("uniform", -5.0, 4.0, 3)], 1).initialize()
    assert c.N == 14 and c.Q == 20


def test_extended_oracle_counts_beyond_five_variables():
    """nvars > 5 has no reference; the enumeration rule must still give C(n+p,p) / prod(p+1) terms,
    graded order, and weights that sum to one (pspace/core.py:181,229,276)."""
    from math import comb
    api = H.oracle_api(False)
    from pspace_b200.workloads import mixed_params
    c = H.CpuContainer(api, mixed_params(7, 2), 1).initialize()
    assert c.N == comb(9, 2) and c.Q == 3 ** 7
    deg = c.degrees()
    tot = deg.sum(axis=1)
    assert np.all(np.diff(tot) >= 0) and tot.max() == 2
    assert len({tuple(r) for r in deg}) == c.N
    _, _, W = c.grid()
    assert abs(W.sum() - 1.0) < 1e-13
    c2 = H.CpuContainer(api, mixed_params(6, 1), 0).initialize()
    assert c2.N == 64 and np.array_equal(c2.degrees()[0], np.zeros(6, dtype=np.intc))
