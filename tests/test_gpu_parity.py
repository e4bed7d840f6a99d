"""GPU parity tests: the CUDA path, driven through the C-ABI (pspace_b200.cabi / DeviceContainer), against
(a) the golden vectors generated from the compiled reference and (b) the restated oracle on the same
seeded inputs.

Bars (BASELINE.json north_star): integer/index work bit-exact; FP64 <= 1e-12 relative.  What is asserted:
  * degree tables .................. bit-exact
  * real 1-D rules, grids (Z,Y,W), 1-D tables, basis values at grid / arbitrary points ... bit-exact
    (Legendre degree >= 5 goes through a double-double integer power instead of libm pow: <= 4 ulp asserted)
  * complex rules/tables/basis ...... <= 1e-14 relative (complex division is Smith's, not libgcc's)
  * projected F / C ................. normwise max|C - C_ref| / max|C_ref| <= 1e-12  (summation
    order differs from the reference's sequential q loop, element-wise error is meaningless for entries
    that cancel to ~0)
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import harness as H  # noqa: E402
from tests.util import golden_params, load_golden, normwise, same_bits  # noqa: E402

TOL = 1e-12


def _dc(params, basis_type, cplx):
    from pspace_b200.device import DeviceContainer
    return DeviceContainer(params, basis_type, cplx)


def _synth(n, seed, cplx):
    return H.synthetic(2 * n, seed).view(np.complex128) if cplx else H.synthetic(n, seed)


def _ulps(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.spacing(np.abs(b)), 1e-300))


CASES = ["cfg1_real", "cfg1_complex", "test_pspace_real", "test_pspace_complex", "cfg2_real", "cfg2_complex",
         "cfg3_real", "highdeg_real", "highdeg_complex"]


@pytest.mark.parametrize("name", CASES)
def test_container_against_golden(name):
    g = load_golden(name)
    cplx = bool(int(g["is_complex"]))
    ps = golden_params(g)
    c = _dc(ps, int(g["basis_type"]), cplx).initialize()
    assert c.N == int(g["N"]) and c.Q == int(g["Q"])
    # (1) integer kernel: bit-exact
    assert np.array_equal(c.degrees(), g["degrees"])
    # (2) grid
    Z, Y, W = (t.cpu().numpy() for t in c.grid())
    has_deg5_legendre = any(k == "uniform" and d >= 5 for k, _, _, d in ps)
    if not cplx:
        assert same_bits(W, g["W"])
        if "Z" in g.files:
            assert same_bits(Z, g["Z"]) and same_bits(Y, g["Y"])
        else:
            assert same_bits(Z[g["qsel"]], g["Z_sel"]) and same_bits(Y[g["qsel"]], g["Y_sel"])
    else:
        assert normwise(W, g["W"]) < 1e-14
        Zg = g["Z"] if "Z" in g.files else None
        if Zg is not None:
            assert normwise(Z, Zg) < 1e-14 and normwise(Y, g["Y"]) < 1e-14
    assert np.array_equal(c.weights(), W)
    # (2b) batched basis evaluation
    psi = c.evalBasisGrid().cpu().numpy()[:, g["qsel"]]
    psi_arb = c.evalBasis(g["zarb"]).cpu().numpy()
    if not cplx and not has_deg5_legendre:
        assert same_bits(psi, g["psi_sel"])
        assert same_bits(psi_arb, g["psi_arb"])
    else:
        assert normwise(psi, g["psi_sel"]) < 1e-13
        assert normwise(psi_arb, g["psi_arb"]) < 1e-13
    # (3) projections
    m, seed, Q = int(g["m"]), int(g["seed"]), c.Q
    f = _synth(Q * m, seed, cplx).reshape(Q, m)
    J = _synth(Q * m * m, seed + 1000, cplx).reshape(Q, m * m)
    fd, Jd = torch.from_numpy(f).cuda(), torch.from_numpy(J).cuda()
    F = c.projectVector(fd).cpu().numpy()
    assert normwise(F, g["F"]) < TOL
    i0, i1, j0, j1 = (int(v) for v in g["win"])
    C = c.projectMatrix(Jd, i0, i1, j0, j1).cpu().numpy()
    assert C.shape == g["C"].shape
    assert normwise(C, g["C"]) < TOL
    # host-buffer entry points give the same answer
    Ch = c.projectMatrixHost(J, i0, i1, j0, j1)
    assert np.array_equal(Ch, C)
    Fh = c.projectVectorHost(f)
    assert np.array_equal(Fh, F)
    c.close()


def test_one_dimensional_rules_bit_exact():
    """(z, y, w) of every 1-D rule, npoints 1..20, each family (one single-parameter container per rule)."""
    for cplx in (False, True):
        g = load_golden("onedim_z" if cplx else "onedim")
        ps = golden_params(g)
        if not cplx:
            ps = [(k, float(np.real(a)), float(np.real(b)), d) for k, a, b, d in ps]
        checked = 0
        for pid, p in enumerate(ps):
            for n in range(1, 21):
                key = f"rule_{pid}_{n}"
                if key not in g.files:
                    continue
                c1 = _dc([(p[0], p[1], p[2], 1)], 0, cplx).initializeQuadrature([n])
                got = np.stack(c1.rules(0, n))
                if not cplx:
                    assert same_bits(got, g[key]), key
                else:
                    assert normwise(got, g[key]) < 1e-15, key
                c1.close()
                checked += 1
        assert checked == 76


def test_polynomials_d0_to_12_against_golden():
    """unit polynomials of every family at arbitrary z, degrees 0..12 (golden from the compiled reference)."""
    for cplx in (False, True):
        g = load_golden("onedim_z" if cplx else "onedim")
        ps = golden_params(g)
        if not cplx:
            ps = [(k, float(np.real(a)), float(np.real(b)), d) for k, a, b, d in ps]
        for pid, p in enumerate(ps):
            c = _dc([(p[0], p[1], p[2], 12)], 0, cplx).initializeBasis([12])
            zs = g[f"z_{pid}"].reshape(-1, 1)
            psi = c.evalBasis(zs).cpu().numpy()          # [13][9]
            ref = g[f"poly_{pid}"]
            if cplx:
                assert normwise(psi, ref) < 1e-13, (pid, p[0])
            elif p[0] == "uniform":
                assert same_bits(psi[:5], ref[:5])
                assert _ulps(psi[5:], ref[5:]) <= 4, (pid, _ulps(psi[5:], ref[5:]))
            else:
                assert same_bits(psi, ref), (pid, p[0])
            c.close()


def test_basis_degree_tables_bit_exact():
    from pspace_b200 import cabi
    import ctypes
    g = load_golden("basis_degrees")
    L = cabi.lib()
    for idx in range(int(g["ncases"])):
        pmax = np.ascontiguousarray(g[f"pmax_{idx}"], dtype=np.intc)
        bt = int(g[f"bt_{idx}"])
        ref = g[f"deg_{idx}"]
        n = ctypes.c_longlong()
        cabi.check(L.pspace_cuda_basis_count(len(pmax), pmax.ctypes.data, bt, ctypes.byref(n)))
        assert n.value == ref.shape[0]
        out = torch.zeros((n.value, len(pmax)), dtype=torch.int32, device="cuda")
        nb = ctypes.c_int()
        cabi.check(L.pspace_cuda_basis_degrees(len(pmax), pmax.ctypes.data, bt, ctypes.byref(nb), out.data_ptr(),
                                               n.value, None))
        assert nb.value == ref.shape[0]
        assert np.array_equal(out.cpu().numpy(), ref), (pmax, bt)


@pytest.mark.parametrize("nv,p,bt", [(6, 2, 1), (7, 2, 0), (8, 3, 0), (10, 3, 1), (9, 1, 0)])
def test_basis_degrees_beyond_five_variables_vs_extended_oracle(nv, p, bt):
    from pspace_b200.workloads import mixed_params
    ps = mixed_params(nv, p)
    o = H.CpuContainer(H.oracle_api(False), ps, bt).initialize_basis([p] * nv)
    c = _dc(ps, bt, False).initializeBasis([p] * nv)
    assert c.N == o.N
    assert np.array_equal(c.degrees(), o.degrees())
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_random_windows_and_splits_vs_oracle(cplx):
    """ragged windows, odd column counts, forced q-splits, accumulate, q-shards."""
    rng = np.random.default_rng(5 + cplx)
    ps = [("normal", 0.3 + (1e-30j if cplx else 0), 1.2, 3), ("exponential", 0.0, 0.5, 2), ("uniform", -1.0, 2.0, 4),
          ("normal", -1.0, 0.4, 2)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()
    c = _dc(ps, 1, cplx).initialize()
    assert (c.N, c.Q) == (o.N, o.Q)
    for m2 in (1, 3, 8, 9, 25, 64, 100):
        J = rng.normal(size=(o.Q, m2)).astype(o.dtype)
        if cplx:
            J = J + 1j * rng.normal(size=J.shape)
        Jd = torch.from_numpy(J).cuda()
        i0, i1 = sorted(rng.choice(o.N + 1, 2, replace=False))
        j0, j1 = sorted(rng.choice(o.N + 1, 2, replace=False))
        ref = o.project_matrix(J, i0, i1, j0, j1, nthreads=4)
        for ksplit in (0, 1, 3):
            C = c.projectMatrix(Jd, i0, i1, j0, j1, ksplit=ksplit).cpu().numpy()
            assert normwise(C, ref) < TOL, (m2, ksplit)
        # accumulate adds to what is there
        base = torch.from_numpy(np.ascontiguousarray(ref)).cuda().clone()
        C2 = c.projectMatrix(Jd, i0, i1, j0, j1, out=base, accumulate=True).cpu().numpy()
        assert normwise(C2, 2 * ref) < TOL
        # two q-shards add up to the whole (multi-GPU decomposition)
        qs = int(rng.integers(1, o.Q))
        Ca = c.projectMatrix(Jd[:qs].contiguous(), i0, i1, j0, j1, q0=0, q1=qs)
        Cb = c.projectMatrix(Jd[qs:].contiguous(), i0, i1, j0, j1, q0=qs, q1=o.Q)
        assert normwise((Ca + Cb).cpu().numpy(), ref) < TOL
        # vector projection with the same column counts
        f = J
        F = c.projectVector(Jd).cpu().numpy()
        assert normwise(F, o.project_vector(f)) < TOL, m2
        k0, k1 = sorted(rng.choice(o.N + 1, 2, replace=False))
        Fw = c.projectVector(Jd, k0, k1, ksplit=2).cpu().numpy()
        assert normwise(Fw, o.project_vector(f, k0, k1)) < TOL
    # empty ranges
    assert c.projectMatrix(Jd, 2, 2, 0, 3).numel() == 0
    z = c.projectMatrix(Jd[:0].contiguous(), 0, 2, 0, 2, q0=5, q1=5)
    assert float(z.abs().max()) == 0.0
    c.close()


def test_every_tile_variant_matches():
    rng = np.random.default_rng(11)
    ps = [("normal", 0.0, 1.0, 3), ("uniform", 0.0, 1.0, 3), ("exponential", 0.0, 1.0, 3)]
    o = H.CpuContainer(H.oracle_api(False), ps, 0).initialize()
    c = _dc(ps, 0, False).initialize()
    J = rng.normal(size=(o.Q, 40))
    Jd = torch.from_numpy(J).cuda()
    ref = o.project_matrix(J, 0, 20, 3, 30, nthreads=4)
    for variant in (1, 2, 3, 4, 5, 6):
        C = c.projectMatrix(Jd, 0, 20, 3, 30, variant=variant).cpu().numpy()
        assert normwise(C, ref) < TOL, variant
    refF = o.project_vector(J)
    for variant in range(1, 9):                      # factored-operand kernel: 128- and 96-row tiles x 8/24/32/48 columns
        F = c.projectVector(Jd, variant=variant).cpu().numpy()
        assert normwise(F, refF) < TOL, variant
        assert "factored" in c.lastPlan()
    c.setOption("force_panel", 1)                    # round-1 panel kernels
    for variant in (1, 2, 3):
        F = c.projectVector(Jd, variant=variant).cpu().numpy()
        assert normwise(F, refF) < TOL, variant
        assert "factored" not in c.lastPlan()
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_tma_and_generic_kernels_agree_with_oracle(cplx):
    """Both projection kernels (TMA warp-specialised; cp.async all-warps) on the same input, incl. ragged q."""
    rng = np.random.default_rng(21)
    ps = [("uniform", -1.0, 2.0, 4), ("normal", 0.5, 0.4, 4), ("exponential", 1.0, 0.3, 3)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 0).initialize()     # Q = 100 (not a multiple of 32), N = 100
    c = _dc(ps, 0, cplx).initialize()
    for m2 in (24, 64, 576):
        J = rng.normal(size=(o.Q, m2)).astype(o.dtype)
        if cplx:
            J = J + 1j * rng.normal(size=J.shape)
        Jd = torch.from_numpy(J).cuda()
        ref = o.project_matrix(J, 5, 31, 90, 100, nthreads=8)
        c.setOption("force_generic", -1)
        C1 = c.projectMatrix(Jd, 5, 31, 90, 100).cpu().numpy()
        assert "tma" in c.lastPlan()
        c.setOption("force_generic", 1)
        C2 = c.projectMatrix(Jd, 5, 31, 90, 100).cpu().numpy()
        assert "cp.async" in c.lastPlan()
        assert normwise(C1, ref) < TOL and normwise(C2, ref) < TOL
        refF = o.project_vector(J)
        c.setOption("force_generic", 0)
        F1 = c.projectVector(Jd).cpu().numpy()
        c.setOption("force_generic", 1)
        F2 = c.projectVector(Jd).cpu().numpy()
        assert normwise(F1, refF) < TOL and normwise(F2, refF) < TOL
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_trimmed_last_chunk_keeps_the_bits(cplx):
    """Point counts that are not a multiple of the TMA stage: the dense matrix kernel skips the k-steps of the last chunk
    that hold only zero-filled rows.  Adding those zeros changed nothing, so the output is bitwise the same as with the
    full last chunk (option no_ragged), for whole tiles and for the stream-K remainder (partial tiles through the fix-up)."""
    rng = np.random.default_rng(77)
    ps = [("uniform", -1.0, 2.0, 5), ("normal", 0.5, 0.4, 5), ("exponential", 1.0, 0.3, 4)]      # N = Q = 180 = 2 x 64 + 52 ... 5 x 32 + 20
    o = H.CpuContainer(H.oracle_api(cplx), ps, 0).initialize()
    c = _dc(ps, 0, cplx).initialize()
    c.setOption("force_generic", -1)
    for m2, q1 in ((576, o.Q), (64, o.Q), (96, 77), (576, 145)):
        J = rng.normal(size=(o.Q, m2)).astype(o.dtype)
        if cplx:
            J = J + 1j * rng.normal(size=J.shape)
        Jd = torch.from_numpy(J).cuda()
        c.setOption("no_ragged", 1)
        full = c.projectMatrix(Jd[:q1], q0=0, q1=q1).clone()
        c.setOption("no_ragged", 0)
        trimmed = c.projectMatrix(Jd[:q1], q0=0, q1=q1)
        assert "tma" in c.lastPlan()
        assert torch.equal(full, trimmed), (m2, q1)
        if q1 == o.Q:
            assert normwise(trimmed[3:9, 100:].cpu().numpy(), o.project_matrix(J, 3, 9, 100, o.N, nthreads=8)) < TOL
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_expansion_evaluation_vs_oracle(cplx):
    """(f1) U[q][c] = sum_k psi_k(z_q) V[k][c] (getDeterministicStates), incl. ranges, odd columns, accumulate, and the
    round trip project(evaluate(V)) == V (orthonormality: exact quadrature of degree <= 2p)."""
    rng = np.random.default_rng(41 + cplx)
    ps = [("normal", 0.3 + (1e-30j if cplx else 0), 1.2, 3), ("exponential", 0.0, 0.5, 2), ("uniform", -1.0, 2.0, 4),
          ("normal", -1.0, 0.4, 2)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 0).initialize()      # tensor basis: N = Q = 180
    c = _dc(ps, 0, cplx).initialize()
    for m in (1, 3, 8, 24, 28, 33, 60, 70):          # every warp-tile width of k_expand
        V = rng.normal(size=(o.N, m)).astype(o.dtype)
        if cplx:
            V = V + 1j * rng.normal(size=V.shape)
        Vd = torch.from_numpy(V).cuda()
        U = c.evaluateExpansion(Vd)
        assert normwise(U.cpu().numpy(), o.evaluate_expansion(V)) < TOL, m
        k0, k1, q0, q1 = 3, 41, 17, 150
        Uw = c.evaluateExpansion(Vd[k0:k1].contiguous(), k0, k1, q0, q1).cpu().numpy()
        assert normwise(Uw, o.evaluate_expansion(V[k0:k1], k0, k1, q0, q1)) < TOL
        U2 = c.evaluateExpansion(Vd, out=U.clone(), accumulate=True)
        assert normwise(U2.cpu().numpy(), 2 * U.cpu().numpy()) < TOL
        # projection of the evaluated expansion returns the coefficients
        back = c.projectVector(U).cpu().numpy()
        assert normwise(back, V) < 1e-11, m
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_factored_operand_kernels_vs_oracle_and_panel_kernels(cplx):
    """Vector projection and expansion evaluation through the factored-operand kernels (psi = prefix x tail, formed
    in shared memory; no basis panel): every tile variant, rule sizes that do not divide the 64-point stage (tail
    lengths 3, 5, 15, 20, 36 ...), q-windows that start inside a tail run, basis windows, accumulate, a one-variable
    container (empty prefix), and agreement with the round-1 panel kernels on the same input."""
    rng = np.random.default_rng(123 + cplx)
    pert = 1e-30j if cplx else 0.0
    containers = [
        ([("normal", 0.3 + pert, 1.2, 3), ("uniform", -1.0, 2.0, 2), ("exponential", 0.0, 0.5, 4), ("normal", 1.0, 0.4, 2)], 1, None),
        ([("uniform", 0.0, 1.0 + pert, 5), ("normal", 0.0, 1.0, 5), ("exponential", 1.0, 0.5, 5)], 0, None),      # nq = 6: L = 36
        ([("normal", 0.5 + pert, 1.5, 19)], 0, None),                                                       # one rule of 20 points
        ([("normal", 0.0, 1.0, 1), ("uniform", 0.0, 1.0, 1), ("exponential", 0.0, 1.0, 1)] * 3, 1, None),     # 9 variables, nq = 2
        ([("uniform", -1.0, 1.0, 2), ("normal", 0.2, 0.7 + pert, 2), ("exponential", 0.0, 1.0, 2)], 0, [7, 3, 5]),  # nq != dmax+1
        ([("normal", 0.1 * d + pert * (d == 0), 1.0 + 0.1 * d, 3) for d in range(4)] + [("uniform", -1.0, 1.0, 3), ("exponential", 0.0, 1.0, 3)],
         1, None),                                                                    # 6 variables, nq = 4: L = 64, N = 84, Q = 4096
        ([("uniform", 0.0, 1.0, 7), ("normal", 0.0, 1.0 + pert, 7), ("exponential", 0.0, 1.0, 3)], 0, None),   # tensor: N = Q = 256, rules of 8, 8, 4 points
    ]
    nvar_v = 6 if cplx else 8
    for ps, bt, nq in containers:
        o = H.CpuContainer(H.oracle_api(cplx), ps, bt).initialize()
        c = _dc(ps, bt, cplx).initialize()
        if nq is not None:
            o.initialize_quadrature(np.array(nq, dtype=np.intc))
            c.initializeQuadrature(nq)
        assert (c.N, c.Q) == (o.N, o.Q)
        tol = 5e-11 if len(ps) == 1 else TOL
        for m in (2, 12, 24, 26):
            f = rng.normal(size=(o.Q, m)).astype(o.dtype)
            V = rng.normal(size=(o.N, m)).astype(o.dtype)
            if cplx:
                f = f + 1j * rng.normal(size=f.shape)
                V = V + 1j * rng.normal(size=V.shape)
            fd, Vd = torch.from_numpy(f).cuda(), torch.from_numpy(V).cuda()
            refF, refU = o.project_vector(f), o.evaluate_expansion(V)
            for variant in range(0, nvar_v + 1):
                F = c.projectVector(fd, variant=variant)
                assert "factored" in c.lastPlan()
                assert normwise(F.cpu().numpy(), refF) < tol, (ps, m, variant)
            U = c.evaluateExpansion(Vd)
            assert "factored" in c.lastPlan()
            assert normwise(U.cpu().numpy(), refU) < tol, (ps, m)
            # windows: q range starting inside a tail run / ending ragged, basis sub-range, accumulate
            q0, q1 = (1, o.Q - 1) if o.Q > 4 else (0, o.Q)
            k0, k1 = (1, o.N - 1) if o.N > 3 else (0, o.N)
            Fw = c.projectVector(fd[q0:q1].contiguous(), k0, k1, q0=q0, q1=q1)
            fz = f.copy()
            fz[:q0] = 0
            fz[q1:] = 0                                           # the oracle sums over all q: zero rows outside the window
            assert normwise(Fw.cpu().numpy(), o.project_vector(fz, k0, k1)) < tol, (ps, m)
            Fa = c.projectVector(fd[q0:q1].contiguous(), k0, k1, q0=q0, q1=q1, out=Fw.clone(), accumulate=True)
            assert normwise(Fa.cpu().numpy(), 2 * Fw.cpu().numpy()) < 1e-15
            Uw = c.evaluateExpansion(Vd[k0:k1].contiguous(), k0, k1, q0, q1)
            assert normwise(Uw.cpu().numpy(), o.evaluate_expansion(V[k0:k1], k0, k1, q0, q1)) < tol, (ps, m)
            Ua = c.evaluateExpansion(Vd[k0:k1].contiguous(), k0, k1, q0, q1, out=Uw.clone(), accumulate=True)
            assert normwise(Ua.cpu().numpy(), 2 * Uw.cpu().numpy()) < 1e-15
            # the panel kernels of round 1 on the same input
            c.setOption("force_panel", 1)
            Fp, Up = c.projectVector(fd), c.evaluateExpansion(Vd)
            c.setOption("force_panel", 0)
            assert normwise(Fp.cpu().numpy(), F.cpu().numpy()) < 1e-13 and normwise(Up.cpu().numpy(), U.cpu().numpy()) < 1e-13
        c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_basis_panel_is_bit_identical_to_batched_basis(cplx):
    """The q-major panel kernel (k_psi_panel: prefix product over the slow variables kept per thread) must produce the
    very values of the k-major batched evaluation (k_eval_basis_grid): expanding the identity returns the panel itself
    (every sum has one non-zero term), for whole grids and for q-windows that start inside a prefix run."""
    for ps, bt in (([("normal", 0.0, 1.0, 2), ("exponential", 0.0, 1.0, 0), ("uniform", -1.0, 1.0, 4), ("normal", 1.0, 2.0, 1)], 0),
                   ([("uniform", 0.0, 1.0, 3), ("normal", 0.0, 1.0, 3), ("exponential", 0.5, 1.0, 0)], 1),
                   ([("normal", 0.2, 0.7, 4)], 0),
                   ([("uniform", 0.0, 1.0, 1)] * 6 + [("normal", 0.0, 1.0, 2)], 1),
                   ([("normal", 0.0, 1.0, 8), ("exponential", 0.0, 1.0, 8), ("uniform", 0.0, 1.0, 8)], 0)):   # N = 729 > 512 rows
        c = _dc(ps, bt, cplx).initialize()
        dt = torch.complex128 if cplx else torch.float64
        eye = torch.eye(c.N, dtype=dt, device="cuda")
        psi = c.evalBasisGrid()                                  # [N][Q]
        # the factored-operand kernel multiplies (prefix product) x (tail product): same factors, other association
        Uf = c.evaluateExpansion(eye)
        assert ("factored" in c.lastPlan()) == (c.N * (2 if cplx else 1) % 2 == 0)     # odd real column counts: panel kernel
        assert normwise(Uf.cpu().numpy(), psi.T.cpu().numpy()) < 1e-15, ps
        c.setOption("force_panel", 1)
        U = c.evaluateExpansion(eye)                             # [Q][N]
        assert "panel" in c.lastPlan()

        def same(a, b):      # real: pure products in the same order; complex: the panel kernel contracts to FMAs
            return torch.equal(a, b) if not cplx else normwise(a.cpu().numpy(), b.cpu().numpy()) < 1e-14

        assert same(U, psi.T.contiguous()), ps
        for q0, q1 in ((1, c.Q), (c.Q // 3, c.Q - 1), (c.Q - 1, c.Q)):
            for k0, k1 in ((0, c.N), (1, min(c.N, 3))):
                Uw = c.evaluateExpansion(torch.eye(k1 - k0, dtype=dt, device="cuda"), k0, k1, q0, q1)
                assert same(Uw, psi[k0:k1, q0:q1].T.contiguous()), (ps, q0, q1, k0, k1)
        c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_tacs_layouts_and_pair_mask_bit_exact(cplx):
    """(f2) interleaved scatter/gather layouts and the sparsity filter: index work, bit-exact vs the oracle loops."""
    from pspace_b200 import device as D
    rng = np.random.default_rng(51)
    ps = [("normal", 0.0, 1.0, 2), ("uniform", 0.0, 1.0, 3), ("exponential", 0.0, 1.0, 1)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()
    c = _dc(ps, 1, cplx).initialize()
    N = o.N
    for nnodes, ndvpn in ((1, 1), (2, 3), (3, 2), (4, 6)):
        m = nnodes * ndvpn
        F = rng.normal(size=(N, m)).astype(o.dtype) * (1 + (0.3j if cplx else 0))
        C = rng.normal(size=(N, N, m * m)).astype(o.dtype) * (1 - (0.7j if cplx else 0))
        res0 = rng.normal(size=N * m).astype(o.dtype)
        ref = H.tacs_scatter_vector(F, nnodes, ndvpn, res0.copy())
        got = D.scatter_vector_tacs(torch.from_numpy(F).cuda(), nnodes, ndvpn, torch.from_numpy(res0.copy()).cuda())
        assert same_bits(got.cpu().numpy(), ref)
        V = D.gather_vector_tacs(torch.from_numpy(ref).cuda(), N, nnodes, ndvpn).cpu().numpy()
        assert same_bits(H.tacs_scatter_vector(V, nnodes, ndvpn), ref)          # gather is the inverse permutation
        refm = H.tacs_scatter_matrix(C, nnodes, ndvpn)
        gotm = D.scatter_matrix_tacs(torch.from_numpy(C).cuda(), nnodes, ndvpn)
        assert same_bits(gotm.cpu().numpy(), refm)
        pq = D.point_quantity_layout(torch.from_numpy(F).cuda()).cpu().numpy()
        assert same_bits(pq, np.ascontiguousarray(F.T).ravel())
    for dmapf in ([3, 3, 3], [1, 0, 1], [0, 0, 0], [2, 1, 0]):
        assert np.array_equal(c.pairMask(dmapf).cpu().numpy(), H.pair_mask(o.degrees(), dmapf)), dmapf
    c.close()


@pytest.mark.parametrize("force_panel", [0, 1])
@pytest.mark.parametrize("cplx", [False, True])
def test_panel_cache_slots_reuse_and_invalidate(cplx, force_panel):
    """The two panel slots (unweighted rows: expansion + i side; weighted rows: j side) are reused when a call repeats
    their (row range, q-range) and rebuilt otherwise: an interleaved sequence of expansions, vector and matrix
    projections with the cache on must return bit-for-bit what the same calls return with the cache off, also across
    a re-initialisation of the quadrature."""
    rng = np.random.default_rng(91)
    ps = [("normal", 0.3 + (1e-30j if cplx else 0), 1.2, 3), ("exponential", 0.0, 0.5, 2), ("uniform", -1.0, 2.0, 3)]
    c = _dc(ps, 0, cplx).initialize()
    c.setOption("force_panel", force_panel)      # 1: the one-pass operators use the panel slots too (round-1 kernels)
    dt = torch.complex128 if cplx else torch.float64

    def operands():
        V = torch.randn((c.N, 6), dtype=torch.float64, device="cuda").to(dt)
        J = torch.randn((c.Q, 16), dtype=torch.float64, device="cuda").to(dt)
        return V, J

    def sequence(V, J):
        q0, q1 = 5, c.Q - 3
        out = []
        for rep in range(2):
            out.append(c.evaluateExpansion(V))                                            # U slot: all rows, all q
            out.append(c.projectMatrix(J, 0, c.N, 2, 20))                                 # U hit, W built
            out.append(c.projectVector(J, 2, 20))                                         # W hit
            out.append(c.projectVector(J))                                                # W replaced
            out.append(c.projectMatrix(J[q0:q1].contiguous(), 3, 11, 0, c.N, q0=q0, q1=q1))   # both replaced (q-window)
            out.append(c.evaluateExpansion(V[3:11].contiguous(), 3, 11, q0, q1))          # U hit
            out.append(c.projectMatrix(J, 0, c.N, 2, 20))                                 # both rebuilt
        return [o.clone() for o in out]

    for nq in (None, [5, 4, 6]):
        if nq is not None:
            c.initializeQuadrature(nq)           # new grid: every cached panel is stale
        V, J = operands()
        c.setOption("panel_cache", 0)
        ref = sequence(V, J)
        c.setOption("panel_cache", 1)
        got = sequence(V, J)
        for k, (a, b) in enumerate(zip(ref, got)):
            assert torch.equal(a, b), (nq, k)
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_adjoint_residual_product_vs_oracle(cplx):
    """(f4) TACSStochasticElement::addAdjResProduct (TACSStochasticElement.cpp:548-637) as a component: the device path
    pspace_cuda_adjoint_states -> the caller's deterministic element (here the oracle's synthetic one, restated in torch)
    -> pspace_cuda_adjoint_product, against the oracle's point-by-point restatement of the reference loop
    (oracle_adj_res_product).  Also over two q-shards, whose partial dfdx add up (multi-GPU decomposition)."""
    from pspace_b200 import device as D
    rng = np.random.default_rng(77 + cplx)
    ps = [("normal", 0.3 + (1e-30j if cplx else 0), 1.2, 3), ("uniform", -1.0, 2.0, 2), ("exponential", 0.0, 0.5, 2)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()
    c = _dc(ps, 1, cplx).initialize()
    N, nv = o.N, len(ps)
    nnodes, ndvpn, dvlen, scale = 3, 2, 5, 0.75
    m = nnodes * ndvpn

    def rnd(n):
        a = rng.normal(size=n).astype(o.dtype)
        return a + 1j * rng.normal(size=n) if cplx else a

    v, adj = rnd(N * m), rnd(N * m)                 # TACS layout: [node][basis term][dof]
    ref = o.adj_res_product(v, adj, nnodes, ndvpn, dvlen, scale)
    _, Y, _ = c.grid()

    def element(UP, Yq):                            # the synthetic element's integrand, dvLen values per point
        U, P = UP[:, :m], UP[:, m:]
        spu, su = (P * U).sum(dim=1), U.sum(dim=1)
        return torch.stack([spu * (1.0 + 0.1 * n) + Yq[:, n % nv] * su for n in range(dvlen)], dim=1).contiguous()

    vd, ad = torch.from_numpy(v).cuda(), torch.from_numpy(adj).cuda()
    UP = c.adjointStates(vd, ad, nnodes, ndvpn)
    # the two halves are the plain expansions of the gathered vectors
    assert normwise(UP[:, :m].cpu().numpy(), c.evaluateExpansion(D.gather_vector_tacs(vd, N, nnodes, ndvpn)).cpu().numpy()) < 1e-15
    dfdx = c.adjointProduct(element(UP, Y), scale)
    assert normwise(dfdx.cpu().numpy(), ref) < TOL
    # accumulate semantics (the reference adds into dfdx) and q-shards
    twice = c.adjointProduct(element(UP, Y), scale, out=dfdx.clone(), accumulate=True)
    assert normwise(twice.cpu().numpy(), 2 * ref) < TOL
    qs = c.Q // 3 + 1
    part = c.adjointProduct(element(c.adjointStates(vd, ad, nnodes, ndvpn, 0, qs), Y[:qs]), scale, 0, qs)
    part = c.adjointProduct(element(c.adjointStates(vd, ad, nnodes, ndvpn, qs, c.Q), Y[qs:]), scale, qs, c.Q, out=part, accumulate=True)
    assert normwise(part.cpu().numpy(), ref) < TOL
    c.close()


def test_complex_step_verifies_the_adjoint_product():
    """The reference's verification flow (examples/stacs/smd/projection.cpp:228-236: RealPart(deriv) - ImagPart(f)/1e-30):
    the USE_COMPLEX build carries a 1e-30j perturbation of a distribution parameter through rules, grid, expansion
    evaluation, the element and the weighted reduction; Im(E[F])/1e-30 must equal the REAL build's adjoint-product
    projection of dF/dmu.  F(y; u, psi) = (psi.u) y_0^2 + sum(u) y_0 y_1, y_0 = mu_0 + sigma_0 z  =>  dF/dmu_0 =
    2 (psi.u) y_0 + sum(u) y_1 (u_q, psi_q do not depend on mu_0: the expansion lives in the standard variables z)."""
    rng = np.random.default_rng(99)
    h = 1e-30
    nnodes, ndvpn = 2, 3
    m = nnodes * ndvpn

    def params(cplx):
        return [("normal", 0.8 + (1j * h if cplx else 0.0), 0.35, 3), ("uniform", -0.5, 1.5, 3), ("exponential", 0.2, 0.6, 2)]

    cr = _dc(params(False), 1, False).initialize()
    cz = _dc(params(True), 1, True).initialize()
    N = cr.N
    v, adj = rng.normal(size=N * m), rng.normal(size=N * m)
    # real build: adjoint-product projection of the derivative integrand
    UP = cr.adjointStates(torch.from_numpy(v).cuda(), torch.from_numpy(adj).cuda(), nnodes, ndvpn)
    _, Y, _ = cr.grid()
    U, P = UP[:, :m], UP[:, m:]
    G = (2.0 * (P * U).sum(1) * Y[:, 0] + U.sum(1) * Y[:, 1])[:, None].contiguous()
    deriv = float(cr.adjointProduct(G)[0])
    # complex build: the function itself, perturbed parameter
    UPz = cz.adjointStates(torch.from_numpy(v.astype(np.complex128)).cuda(), torch.from_numpy(adj.astype(np.complex128)).cuda(), nnodes, ndvpn)
    _, Yz, _ = cz.grid()
    Uz, Pz = UPz[:, :m], UPz[:, m:]
    Fz = ((Pz * Uz).sum(1) * Yz[:, 0] ** 2 + Uz.sum(1) * Yz[:, 0] * Yz[:, 1])[:, None].contiguous()
    mean = complex(cz.adjointProduct(Fz)[0])
    cs = mean.imag / h
    assert abs(cs - deriv) <= 1e-10 * abs(deriv), (cs, deriv)
    # and the real part is the real build's mean of F
    Fr = ((P * U).sum(1) * Y[:, 0] ** 2 + U.sum(1) * Y[:, 0] * Y[:, 1])[:, None].contiguous()
    assert abs(mean.real - float(cr.adjointProduct(Fr)[0])) <= 1e-12 * abs(mean.real)
    cr.close()
    cz.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_moments_vs_oracle(cplx):
    """(f3) mean / second moment / variance over the quadrature grid (TACSStochasticFunction.cpp:209-238)."""
    rng = np.random.default_rng(61)
    ps = [("normal", 0.3 + (1e-30j if cplx else 0), 1.2, 4), ("exponential", 0.0, 0.5, 3), ("uniform", -1.0, 2.0, 5)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()
    c = _dc(ps, 1, cplx).initialize()
    for m in (1, 5, 32, 77):
        f = rng.normal(size=(o.Q, m)).astype(o.dtype) + 2.0
        if cplx:
            f = f + 1j * rng.normal(size=f.shape)
        fd = torch.from_numpy(f).cuda()
        for (q0, q1) in ((0, o.Q), (7, 93)):
            ref = o.moments(f[q0:q1], q0, q1)
            got = c.moments(fd[q0:q1].contiguous(), q0, q1)
            for g_, r_ in zip(got, ref):
                assert normwise(g_.cpu().numpy(), r_) < 1e-12, (m, q0)
    # mean of a PC expansion is its zeroth coefficient, variance the sum of squares of the rest (orthonormal basis)
    V = rng.normal(size=(o.N, 3)).astype(o.dtype)
    U = c.evaluateExpansion(torch.from_numpy(V).cuda())
    mean, _, var = c.moments(U)
    assert normwise(mean.cpu().numpy(), V[0]) < 1e-11
    assert normwise(var.cpu().numpy(), (V[1:] * V[1:]).sum(axis=0)) < 1e-10
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_masked_projection_sparsity_filter(cplx):
    """(f2) projection restricted by the nonzero() pair filter: kept pairs equal the dense result, filtered pairs are
    exact zeros (or untouched under accumulate); both kernels; J polynomial of degree <= dmapf reproduces the dense
    projection entirely (that is what the filter is for)."""
    rng = np.random.default_rng(71)
    ps = [("normal", 0.2 + (1e-30j if cplx else 0), 0.8, 3), ("uniform", -1.0, 1.5, 3), ("exponential", 0.5, 0.7, 3)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 0).initialize()            # N = Q = 64
    c = _dc(ps, 0, cplx).initialize()
    N, Q = o.N, o.Q
    for dmapf in ([1, 1, 1], [0, 2, 1], [3, 3, 3], [0, 0, 0]):
        maskh = H.pair_mask(o.degrees(), dmapf)
        mask = c.pairMask(dmapf)
        for m2 in (4, 64, 9, 576):        # 576: more kept tiles than SMs -> full data-parallel wave + stream-K remainder
            J = rng.normal(size=(Q, m2)).astype(o.dtype)
            if cplx:
                J = J + 1j * rng.normal(size=J.shape)
            Jd = torch.from_numpy(J).cuda()
            dense = o.project_matrix(J, nthreads=8)
            ref = dense * maskh[:, :, None]
            for force in (-1, 0, 1):       # TMA kernel / planner's choice / single-launch kernel
                c.setOption("force_generic", force)
                got = c.projectMatrix(Jd, mask=mask).cpu().numpy()
                assert normwise(got, ref) < TOL, (dmapf, m2, force)
                assert np.all(got[maskh == 0] == 0)
                base = torch.full((N, N, m2), 3.0, dtype=Jd.dtype, device="cuda")
                acc = c.projectMatrix(Jd, out=base, accumulate=True, mask=mask).cpu().numpy()
                assert normwise(acc, ref + 3.0) < TOL
                assert np.all(acc[maskh == 0] == 3.0)
                w = c.projectMatrix(Jd, 5, 40, 9, 33, mask=mask).cpu().numpy()
                assert normwise(w, ref[5:40, 9:33]) < TOL
            c.setOption("force_generic", 0)
    # J(y) linear in every y_d: blocks with |deg_i - deg_j| > 1 in some variable vanish, so the filtered projection
    # with dmapf = 1 equals the dense one
    _, Y, _ = o.grid()
    Jlin = (1.0 + Y[:, 0] * 0.5 - Y[:, 1] * 0.25 + Y[:, 2]).reshape(Q, 1) * np.ones((1, 4))
    dense = o.project_matrix(Jlin.astype(o.dtype), nthreads=4)
    got = c.projectMatrix(torch.from_numpy(np.ascontiguousarray(Jlin.astype(o.dtype))).cuda(), mask=c.pairMask([1, 1, 1])).cpu().numpy()
    assert np.max(np.abs(got - dense)) < 1e-12 * np.max(np.abs(dense))
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_host_input_slab_pipeline(cplx):
    """Host-buffer entry points upload J in slabs overlapped with the projection: many small slabs (forced with
    slab_mb) must give the one-shot device result, incl. accumulate, ragged last slab and pageable memory."""
    rng = np.random.default_rng(81)
    ps = [("normal", 0.2 + (1e-30j if cplx else 0), 0.8, 4), ("uniform", -1.0, 1.5, 6), ("exponential", 0.5, 0.7, 5),
          ("normal", 0.0, 1.0, 3)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()          # Q = 5*7*6*4 = 840
    c = _dc(ps, 1, cplx).initialize()
    m2 = 576
    J = rng.normal(size=(o.Q, m2)).astype(o.dtype)
    if cplx:
        J = J + 1j * rng.normal(size=J.shape)
    ref = c.projectMatrix(torch.from_numpy(J).cuda(), 0, 20, 0, 30).cpu().numpy()
    assert normwise(ref, o.project_matrix(J, 0, 20, 0, 30, nthreads=8)) < TOL
    c.setOption("slab_mb", 1)                                          # 1 MiB slabs -> 4-8 slabs of 224/96 points
    got = c.projectMatrixHost(J, 0, 20, 0, 30)
    assert normwise(got, ref) < 1e-13
    Jp = torch.from_numpy(J).pin_memory()
    out = torch.full((20, 30, m2), 2.0, dtype=Jp.dtype).pin_memory()
    c.projectMatrixHost(Jp, 0, 20, 0, 30, out=out, accumulate=True)
    assert normwise(out.numpy(), ref + 2.0) < 1e-13
    dev = torch.empty((20, 30, m2), dtype=Jp.dtype, device="cuda")
    c.projectMatrixHostIn(Jp[100:700], dev, 0, 20, 0, 30, q0=100, q1=700)
    part = c.projectMatrix(torch.from_numpy(J[100:700]).cuda(), 0, 20, 0, 30, q0=100, q1=700)
    assert normwise(dev.cpu().numpy(), part.cpu().numpy()) < 1e-13
    f = J[:, :24].copy()
    assert normwise(c.projectVectorHost(f), c.projectVector(torch.from_numpy(f).cuda()).cpu().numpy()) < 1e-13
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_host_output_row_block_pipeline(cplx):
    """Output-bound host form (cfg3's shape: small J, large C): C is projected and read back in blocks of basis rows through two
    device buffers.  Forced here with out_block_mb = 1: the device-form result (to summation order: the stream-K split of a
    block differs from the whole window's), incl. a ragged last block, a row-offset
    window, a pair mask, accumulate into pinned and pageable host buffers; the oracle pins a sub-window."""
    rng = np.random.default_rng(83)
    ps = [("normal", 0.2 + (1e-30j if cplx else 0), 0.8, 4), ("uniform", -1.0, 1.5, 5), ("exponential", 0.5, 0.7, 4)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 0).initialize()          # N = Q = 150
    c = _dc(ps, 0, cplx).initialize()
    N, m2 = o.N, 96
    J = rng.normal(size=(o.Q, m2)).astype(o.dtype)
    if cplx:
        J = J + 1j * rng.normal(size=J.shape)
    Jd = torch.from_numpy(J).cuda()
    c.setOption("out_block_mb", 1)            # row of C = 150 x 96 x 8 (16) B = 115 (230) kB -> blocks of 16 rows, 10 blocks, last one ragged
    for (i0, i1, kw) in ((0, N, {}), (3, N - 2, {}), (0, N, {"mask": True})):
        mask = c.pairMask([2] * len(ps)) if kw else None
        ref = c.projectMatrix(Jd, i0, i1, 0, N, mask=mask)
        a = c._args(i0, i1, 0, N, 0, o.Q, m2, False, 0, 0, mask, False)
        host = torch.empty((i1 - i0, N, m2), dtype=Jd.dtype).pin_memory()
        Jp = torch.from_numpy(J).pin_memory()
        from pspace_b200.cabi import check
        check(c.L.pspace_cuda_project_matrix_host(c.h, __import__("ctypes").byref(a), Jp.data_ptr(), host.data_ptr()))
        assert normwise(host.numpy(), ref.cpu().numpy()) < 1e-13, (i0, i1, kw)
        if mask is not None:
            assert bool((host[mask[i0:i1].cpu() == 0] == 0).all())
    full = c.projectMatrix(Jd).cpu().numpy()
    assert normwise(full[40:52, 100:], o.project_matrix(J, 40, 52, 100, N, nthreads=8)) < TOL
    got = c.projectMatrixHost(J)                                        # pageable numpy in / out
    assert normwise(got, full) < 1e-13
    acc = np.full((N, N, m2), 2.0, dtype=o.dtype)
    c.projectMatrixHost(J, out=acc, accumulate=True)
    assert normwise(acc, full + 2.0) < 1e-13
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_symmetric_projection_matches_dense(cplx):
    """symmetric=1 computes pairs j >= i only and mirrors: same full window as the dense projection (the mirrored
    half is bitwise equal to its source block), for full and sub-square windows, with accumulate."""
    rng = np.random.default_rng(91)
    ps = [("normal", 0.2 + (1e-30j if cplx else 0), 0.8, 3), ("uniform", -1.0, 1.5, 4), ("exponential", 0.5, 0.7, 3)]
    o = H.CpuContainer(H.oracle_api(cplx), ps, 0).initialize()            # N = Q = 80
    c = _dc(ps, 0, cplx).initialize()
    N, Q = o.N, o.Q
    for m2 in (4, 64, 576):
        J = rng.normal(size=(Q, m2)).astype(o.dtype)
        if cplx:
            J = J + 1j * rng.normal(size=J.shape)
        Jd = torch.from_numpy(J).cuda()
        for (a0, a1) in ((0, N), (7, 53), (N - 3, N)):
            ref = o.project_matrix(J, a0, a1, a0, a1, nthreads=8)
            got = c.projectMatrix(Jd, a0, a1, a0, a1, symmetric=True).cpu().numpy()
            assert normwise(got, ref) < TOL, (m2, a0)
            assert np.array_equal(got, np.swapaxes(got, 0, 1))               # exact mirror
            base = torch.full((a1 - a0, a1 - a0, m2), -1.5, dtype=Jd.dtype, device="cuda")
            acc = c.projectMatrix(Jd, a0, a1, a0, a1, out=base, accumulate=True, symmetric=True).cpu().numpy()
            assert normwise(acc, ref - 1.5) < TOL
    from pspace_b200.cabi import PspaceCudaError
    with pytest.raises(PspaceCudaError):
        c.projectMatrix(Jd, 0, 10, 1, 11, symmetric=True)                     # not a square window
    c.close()


@pytest.mark.parametrize("case", ["one_var_20pts", "eighteen_vars", "big_tables", "degree_zero_vars"])
def test_edge_shapes_vs_oracle(case):
    """Extremes of the supported range: a single parameter with the 20-point rules and degree 19; more than 16
    parameters (32-byte digit rows); the largest 1-D tables the staged kernels accept; parameters of degree 0."""
    rng = np.random.default_rng(101)
    if case == "one_var_20pts":
        ps, bt, m2 = [("normal", 0.5, 1.5, 19)], 0, 6
    elif case == "eighteen_vars":
        ps = [(("normal", "uniform", "exponential")[d % 3], 0.1 * d, 1.0 + 0.1 * d, 1) for d in range(18)]
        ps[3] = (ps[3][0], ps[3][1], ps[3][2], 2)
        bt, m2 = 1, 4                                            # complete degree 2: C(20,2) - (18-1)... N from the oracle
        ps = [(k, a, b, d if i < 4 else (1 if i < 9 else 0)) for i, (k, a, b, d) in enumerate(ps)]   # Q = 2^8*3*... keep it small
    elif case == "big_tables":
        ps, bt, m2 = [("uniform", -1.0, 1.0, 9), ("exponential", 0.0, 1.0, 10), ("normal", 0.0, 1.0, 8)], 1, 8
    else:
        ps, bt, m2 = [("normal", 0.0, 1.0, 0), ("uniform", 0.0, 2.0, 3), ("exponential", 1.0, 1.0, 0), ("normal", 1.0, 2.0, 2)], 0, 5
    o = H.CpuContainer(H.oracle_api(False), ps, bt).initialize()
    c = _dc(ps, bt, False).initialize()
    assert (c.N, c.Q) == (o.N, o.Q)
    assert np.array_equal(c.degrees(), o.degrees())
    Zo, Yo, Wo = o.grid()
    Z, Y, W = (t.cpu().numpy() for t in c.grid())
    deg5_legendre = any(k == "uniform" and d >= 5 for k, _, _, d in ps)
    assert same_bits(W, Wo) and same_bits(Z, Zo) and same_bits(Y, Yo)
    psi = c.evalBasisGrid().cpu().numpy()
    ref_psi = o.basis_at(Zo)
    if deg5_legendre:
        assert normwise(psi, ref_psi) < 1e-12
    else:
        assert same_bits(psi, ref_psi)
    J = rng.normal(size=(o.Q, m2))
    Jd = torch.from_numpy(J).cuda()
    i1, j1 = min(o.N, 24), min(o.N, 40)
    ref = o.project_matrix(J, 0, i1, 0, j1, nthreads=8)
    tol = 5e-11 if case == "one_var_20pts" else TOL      # degree-19 Hermite at |z| ~ 7: values ~1e6, reference-order cancellation
    for force in (-1, 0, 1):       # TMA kernel / planner's choice / single-launch kernel
        c.setOption("force_generic", force)
        assert normwise(c.projectMatrix(Jd, 0, i1, 0, j1).cpu().numpy(), ref) < tol, (case, force)
        assert normwise(c.projectVector(Jd).cpu().numpy(), o.project_vector(J)) < tol, (case, force)
    V = rng.normal(size=(o.N, 3))
    assert normwise(c.evaluateExpansion(torch.from_numpy(V).cuda()).cpu().numpy(), o.evaluate_expansion(V)) < tol
    c.close()


def test_batched_elements_share_one_container():
    """configs[2] shape: several elements' 24x24 Jacobians over one 4-parameter tensor container (degree lowered so the
    oracle finishes in seconds); the panels are built once and each element equals its own projection."""
    from pspace_b200.workloads import mixed_params
    rng = np.random.default_rng(111)
    ps = mixed_params(4, 2)                                                # N = Q = 81
    o = H.CpuContainer(H.oracle_api(False), ps, 0).initialize()
    c = _dc(ps, 0, False).initialize()
    E, m2 = 5, 576
    J = rng.normal(size=(E, o.Q, m2))
    C = c.projectMatrixBatched(torch.from_numpy(J).cuda(), 3, 30, 0, o.N)
    launches0 = c.launchCount()
    c.projectMatrixBatched(torch.from_numpy(J).cuda(), 3, 30, 0, o.N)
    per_call = c.launchCount() - launches0
    assert per_call <= 2 + 2 * E                                           # 2 panel kernels once + (project, fix-up) per element
    for e in range(E):
        assert normwise(C[e].cpu().numpy(), o.project_matrix(J[e], 3, 30, 0, o.N, nthreads=8)) < TOL, e
    c.close()


def test_no_write_outside_the_output_window():
    """compute-sanitizer is closed on this pool: own bounds check.  The output sits inside a larger buffer of
    sentinels; every kernel path (TMA stream-K incl. fix-up, generic incl. split reduce, vector) must leave the
    guard bands untouched."""
    rng = np.random.default_rng(31)
    ps = [("uniform", -1.0, 2.0, 3), ("normal", 0.5, 0.4, 3), ("exponential", 1.0, 0.3, 3), ("normal", 0.0, 1.0, 2)]
    c = _dc(ps, 1, False).initialize()
    N, Q = c.N, c.Q
    guard = 4096
    for m2 in (6, 25, 64, 576):
        J = torch.from_numpy(rng.normal(size=(Q, m2))).cuda()
        for (i0, i1, j0, j1) in ((0, N, 0, N), (3, 20, 5, 6), (N - 1, N, 0, N)):
            n = (i1 - i0) * (j1 - j0) * m2
            for force in (-1, 0, 1):       # TMA kernel / planner's choice / single-launch kernel
                c.setOption("force_generic", force)
                for ks in (0, 2):
                    buf = torch.full((n + 2 * guard,), -7.25, dtype=torch.float64, device="cuda")
                    out = buf[guard:guard + n].view(i1 - i0, j1 - j0, m2)
                    c.projectMatrix(J, i0, i1, j0, j1, out=out, ksplit=ks)
                    torch.cuda.synchronize()
                    assert bool((buf[:guard] == -7.25).all()) and bool((buf[guard + n:] == -7.25).all()), (m2, i0, i1, force, ks)
                    assert not bool((out == -7.25).any())
        for (k0, k1) in ((0, N), (7, 9)):
            n = (k1 - k0) * m2
            for force in (-1, 0, 1):       # TMA kernel / planner's choice / single-launch kernel
                c.setOption("force_generic", force)
                buf = torch.full((n + 2 * guard,), -7.25, dtype=torch.float64, device="cuda")
                out = buf[guard:guard + n].view(k1 - k0, m2)
                c.projectVector(J, k0, k1, out=out)
                torch.cuda.synchronize()
                assert bool((buf[:guard] == -7.25).all()) and bool((buf[guard + n:] == -7.25).all())
    c.close()


@pytest.mark.parametrize("cplx", [False, True])
def test_no_write_outside_workspace_or_expansion_output(cplx):
    """Own bounds check, continued: (i) the caller-provided workspace is used only up to the size the query returned (stream-K
    partials of the matrix and the factored vector kernels, q-split slices of the single-launch kernel); (ii) the expansion
    kernels (factored and panel) write only rows q0..q1-1 of U; (iii) the host-output row-block path writes only C."""
    rng = np.random.default_rng(33)
    ps = [("uniform", -1.0, 2.0, 3), ("normal", 0.5 + (1e-30j if cplx else 0), 0.4, 4), ("exponential", 1.0, 0.3, 3), ("normal", 0.0, 1.0, 2)]
    c = _dc(ps, 0, cplx).initialize()              # N = Q = 240
    N, Q = c.N, c.Q
    guard = 1 << 16

    def rand(shape):
        a = rng.normal(size=shape)
        return torch.from_numpy(a + 1j * rng.normal(size=shape) if cplx else a).cuda()

    import ctypes

    def guarded(kind, a, fn, tag):
        need = c.L.pspace_cuda_project_workspace(c.h, kind, ctypes.byref(a))
        c._ws = torch.full((need + (1 << 20),), 0xA5, dtype=torch.uint8, device="cuda")
        fn()
        torch.cuda.synchronize()
        assert c._last_ws_need == need and bool((c._ws[need:] == 0xA5).all()), tag

    for m2 in (24, 64, 576):
        J = rand((Q, m2))
        for force in (-1, 0, 1):
            c.setOption("force_generic", force)
            for ks in (0, 3):
                for (q0, q1) in ((0, Q), (17, 203)):
                    Jq = J[q0:q1].contiguous()
                    guarded(1, c._args(0, 40, 3, N, q0, q1, m2, False, ks, 0),
                            lambda: c.projectMatrix(Jq, 0, 40, 3, N, q0=q0, q1=q1, ksplit=ks), ("matrix", m2, force, ks, q0))
                    guarded(0, c._args(0, N, 0, 0, q0, q1, m2, False, ks, 0),
                            lambda: c.projectVector(Jq, 0, N, q0=q0, q1=q1, ksplit=ks), ("vector", m2, force, ks, q0))
    c.setOption("force_generic", 0)
    sent = -7.25
    for m in (3, 24, 50):
        V = rand((N, m))
        for panel in (0, 1):
            c.setOption("force_panel", panel)
            for (q0, q1) in ((0, Q), (5, 6), (31, 222)):
                n = (q1 - q0) * m
                buf = torch.full((n + 2 * guard,), sent, dtype=V.dtype, device="cuda")
                out = buf[guard:guard + n].view(q1 - q0, m)
                c.evaluateExpansion(V, q0=q0, q1=q1, out=out)
                torch.cuda.synchronize()
                assert bool((buf[:guard] == sent).all()) and bool((buf[guard + n:] == sent).all()), (m, panel, q0, q1)
                assert not bool((out == sent).any())
    c.setOption("force_panel", 0)
    # host-output row blocks: C inside a larger pinned buffer of sentinels
    c.setOption("out_block_mb", 1)
    m2 = 64
    Jh = rand((Q, m2)).cpu().pin_memory()
    n = N * N * m2
    hbuf = torch.full((n + 2 * guard,), sent, dtype=Jh.dtype).pin_memory()
    hout = hbuf[guard:guard + n].view(N, N, m2)
    c.projectMatrixHost(Jh, out=hout)
    assert bool((hbuf[:guard] == sent).all()) and bool((hbuf[guard + n:] == sent).all()) and not bool((hout == sent).any())
    c.close()


def test_error_paths_are_loud():
    from pspace_b200.cabi import PspaceCudaError
    ps = [("normal", 0.0, 1.0, 2)]
    c = _dc(ps, 0, False)
    with pytest.raises(PspaceCudaError):
        c.projectVector(torch.zeros((3, 1), dtype=torch.float64, device="cuda"), 0, 0, 0, 3)   # not initialised
    with pytest.raises(PspaceCudaError):
        c.initializeQuadrature([21])
    with pytest.raises(PspaceCudaError):
        c.initializeBasis([25])
    c.initialize()
    with pytest.raises(PspaceCudaError):
        c.projectVector(torch.zeros((3, 1), dtype=torch.float64, device="cuda"), 0, 9, 0, 3)   # k1 > N
    c.close()


@pytest.mark.parametrize("mode", ["dense", "masked", "symmetric", "complex"])
def test_stage_release_ordering_under_memory_pressure(mode):
    """The TMA ring of k_project_ws refills a stage only after the consumers have released the chunk AFTER the one that
    occupied it (project.cu, producer): every fragment load has provably returned by then.  Stress: 50 repetitions of
    a multi-chunk, multi-segment projection while a second stream saturates HBM with large copies (the condition that
    exposed round 1's hazard: a busy load/store path); every repetition must be bit-identical to the first, and the
    first must match the oracle."""
    rng = np.random.default_rng(7)
    cplx = mode == "complex"
    ps = [("normal", 0.3 + (1e-30j if cplx else 0), 1.2, 3), ("uniform", -1.0, 2.0, 3), ("exponential", 0.0, 0.5, 3),
          ("normal", -1.0, 0.4, 3), ("uniform", 0.0, 1.0, 3), ("normal", 1.0, 0.3, 3)]          # Q = 4096: 64 (128) chunks per tile
    o = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()
    c = _dc(ps, 1, cplx).initialize()
    m2 = 96
    J = rng.normal(size=(o.Q, m2)).astype(o.dtype)
    if cplx:
        J = J + 1j * rng.normal(size=J.shape)
    Jd = torch.from_numpy(J).cuda()
    kw = {}
    if mode == "masked":
        kw["mask"] = c.pairMask([2] * len(ps))
    if mode == "symmetric":
        kw["symmetric"] = True
    first = c.projectMatrix(Jd, **kw).clone()
    assert "tma" in c.lastPlan()
    ref = o.project_matrix(J, 0, 12, 60, o.N, nthreads=8)
    got = first[0:12, 60:].cpu().numpy()
    if mode == "masked":
        got = c.projectMatrix(Jd)[0:12, 60:].cpu().numpy()
    assert normwise(got, ref) < TOL
    side = torch.cuda.Stream()
    big_a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")        # 2 GiB each
    big_b = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    big_a.fill_(1.0)
    torch.cuda.synchronize()
    for rep in range(50):
        with torch.cuda.stream(side):
            for _ in range(3):
                big_b.copy_(big_a, non_blocking=True)                       # ~1.3 ms of HBM-saturating traffic beside each projection
        C = c.projectMatrix(Jd, **kw)
        torch.cuda.synchronize()
        assert torch.equal(C, first), (mode, rep)
    c.close()
