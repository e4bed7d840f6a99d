#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref, built from
/root/reference by `make -C oracle ref`).  Runs only in the build container (the GPU box has no
/root/reference; it consumes the committed .npz files).

Every array below is an output of the reference's own code: degree tables, 1-D rules, grids, basis
values come straight from ParameterContainer / AbstractParameter calls; projected F / C come from the
reference-order loops of oracle/ref_shim.cpp (TACSStochasticElement.cpp:283-311, 380-414) over
those calls.  Inputs f / J are the splitmix64 synthetic blocks (seed = 20260 + config id).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402
from pspace_b200 import workloads as WL  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synth(n, seed, cplx):
    if cplx:
        return H.synthetic(2 * n, seed).view(np.complex128)
    return H.synthetic(n, seed)


def params_arrays(params):
    return dict(kinds=np.array([H.KIND[p[0]] for p in params], dtype=np.intc),
                p1=np.array([p[1] for p in params], dtype=np.complex128),
                p2=np.array([p[2] for p in params], dtype=np.complex128),
                dmax=np.array([p[3] for p in params], dtype=np.intc))


def container_case(name, params, basis_type, cplx, m, window, nsample, seed, full_grid=True):
    api = H.ref_api(cplx)
    c = H.CpuContainer(api, params, basis_type).initialize()
    N, Q = c.N, c.Q
    Z, Y, W = c.grid()
    rng = np.random.default_rng(seed)
    qsel = np.sort(rng.choice(Q, min(nsample, Q), replace=False))
    psi = c.basis_at(Z[qsel])                       # [N][nsel]
    zarb = rng.normal(size=(5, c.nvars)).astype(c.dtype)
    if cplx:
        zarb = zarb + 1e-30j
    psi_arb = c.basis_at(zarb)
    f = synth(Q * m, seed, cplx).reshape(Q, m)
    J = synth(Q * m * m, seed + 1000, cplx).reshape(Q, m * m)
    F = c.project_vector(f)
    wi, wj = min(window[0], N), min(window[1], N)
    i0 = (N - wi) // 2                               # window away from the origin
    j0 = N - wj
    C = c.project_matrix(J, i0, i0 + wi, j0, j0 + wj, nthreads=8)
    d = dict(basis_type=basis_type, is_complex=int(cplx), m=m, seed=seed, N=N, Q=Q,
             degrees=c.degrees(), W=W, qsel=qsel, psi_sel=psi, zarb=zarb, psi_arb=psi_arb,
             F=F, win=np.array([i0, i0 + wi, j0, j0 + wj]), C=C, **params_arrays(params))
    if full_grid:
        d.update(Z=Z, Y=Y)
    else:
        d.update(Z_sel=Z[qsel], Y_sel=Y[qsel])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: N={N} Q={Q} window={wi}x{wj} |C|max={np.abs(C).max():.3e}")


def degree_cases():
    out = {}
    api = H.ref_api(False)
    cases = [((4, 3), 1), ((1, 2), 0), ((2, 2, 2), 1), ((3,), 0), ((3,), 1), ((0, 2), 1), ((2, 0, 1), 0),
             ((3, 3, 3), 0), ((3, 3, 3), 1), ((4, 1, 2), 1), ((2, 3, 1, 2), 0), ((2, 3, 1, 2), 1),
             ((4, 4, 4, 4, 4), 1), ((1, 2, 1, 2, 1), 0), ((3, 2, 3, 2, 3), 1), ((5, 5, 5, 5), 0), ((0, 0, 0), 1)]
    for idx, (pmax, bt) in enumerate(cases):
        ps = [("normal", 0.0, 1.0, p) for p in pmax]
        c = H.CpuContainer(api, ps, bt).initialize_basis(pmax)
        out[f"pmax_{idx}"] = np.array(pmax, dtype=np.intc)
        out[f"bt_{idx}"] = bt
        out[f"deg_{idx}"] = c.degrees()
    out["ncases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "basis_degrees.npz"), **out)
    print("basis_degrees:", len(cases), "cases")


def onedim_cases():
    """1-D rules (npoints 1..20) and unit polynomials (d = 0..12) of each family, real and complex."""
    lag_ok = [n for n in range(1, 21) if n not in (12, 13, 17, 18)]
    for cplx in (False, True):
        api = H.ref_api(cplx)
        eps = 1e-30j if cplx else 0.0
        ps = [("normal", 1.3 + eps, 0.7, 1), ("uniform", -0.4, 2.1 + eps, 1), ("exponential", 0.25 + eps, 1.7, 1),
              ("uniform", 1.0, 0.1, 1)]  # a > b as in tests/test_pspace.py:9
        c = H.CpuContainer(api, ps, 0)
        out = params_arrays(ps)
        for pid, p in enumerate(ps):
            ns = lag_ok if p[0] == "exponential" else list(range(1, 21))
            for n in ns:
                z, y, w = c.param_quadrature(pid, n)
                out[f"rule_{pid}_{n}"] = np.stack([z, y, w])
            rng = np.random.default_rng(7 + pid)
            zs = rng.normal(size=9) * (0.5 if p[0] == "uniform" else 1.5) + (0.5 if p[0] == "uniform" else 0.0)
            if p[0] == "exponential":
                zs = np.abs(zs) * 2
            zs = zs.astype(c.dtype) + eps
            out[f"z_{pid}"] = zs
            out[f"poly_{pid}"] = np.stack([c.param_basis(pid, zs, d) for d in range(13)])
        np.savez_compressed(os.path.join(OUT, "onedim_z.npz" if cplx else "onedim.npz"), **out)
    print("onedim: done")


def main():
    os.makedirs(OUT, exist_ok=True)
    assert H.ref_available(False) and H.ref_available(True), "run `make -C oracle ref` first"
    degree_cases()
    onedim_cases()
    w1 = WL.cfg1()
    container_case("cfg1_real", w1.params, w1.basis_type, False, 1, (64, 64), 64, w1.seed)
    p1c = [(k, a + (1e-30j if i == 0 else 0), b, d) for i, (k, a, b, d) in enumerate(w1.params)]
    container_case("cfg1_complex", p1c, w1.basis_type, True, 1, (64, 64), 64, w1.seed)
    # the container of the reference's own tests/test_pspace.py:6-17 (complete basis, Uniform with a > b)
    tp = [("normal", 1.0, 0.1, 2), ("uniform", 1.0, 0.1, 2), ("exponential", 1.0, 0.1, 2)]
    container_case("test_pspace_real", tp, 1, False, 2, (10, 10), 27, 777)
    container_case("test_pspace_complex", tp, 1, True, 2, (10, 10), 27, 777)
    w2 = WL.cfg2()
    container_case("cfg2_real", w2.params, w2.basis_type, False, w2.m, (8, 8), 40, w2.seed, full_grid=False)
    container_case("cfg2_complex", w2.params, w2.basis_type, True, w2.m, (4, 4), 20, w2.seed, full_grid=False)
    w3 = WL.cfg3()   # degree-5 paths: Legendre binomial sum with pow, Hermite/Laguerre naive recursion
    container_case("cfg3_real", w3.params, w3.basis_type, False, 24, (3, 3), 24, w3.seed, full_grid=False)
    # mixed dmax, 5 variables, tensor, high degree in one variable (factorial loop path n > 10)
    hp = [("uniform", 0.0, 1.0, 6), ("normal", 0.0, 1.0, 7), ("exponential", 0.0, 1.0, 5)]
    container_case("highdeg_real", hp, 0, False, 3, (6, 6), 30, 4242, full_grid=False)
    container_case("highdeg_complex", [(k, a + 1e-30j, b, d) for (k, a, b, d) in hp], 1, True, 3, (6, 6), 30, 4242,
                   full_grid=False)


if __name__ == "__main__":
    main()
