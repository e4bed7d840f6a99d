"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares, the
host C++ library exports what the reference's wrapper links to, per-parameter host arithmetic against the
golden vectors, loud failure without a GPU, workload sizes."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests.util import ROOT, golden_params, load_golden, same_bits

LIB = os.path.join(ROOT, "pspace_b200", "lib")


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "pspace_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pspace_cuda_[a-z0-9_]+)\s*\(", text)))


def test_cabi_library_exports_every_declared_symbol():
    import ctypes
    from pspace_b200 import cabi
    L = ctypes.CDLL(cabi.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    # the ctypes table covers the header (no undeclared / unbound entry point)
    assert set(cabi.SIGNATURES) <= set(names)
    missing = set(names) - set(cabi.SIGNATURES) - {"pspace_cuda_malloc", "pspace_cuda_free", "pspace_cuda_memcpy"}
    assert not missing, missing


def test_cabi_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pspace_b200 import cabi
    L = cabi.lib()
    assert L.pspace_cuda_device_count() < 0
    import ctypes
    h = ctypes.c_void_p()
    kinds = np.zeros(1, dtype=np.intc); p = np.ones(1); d = np.ones(1, dtype=np.intc)
    rc = L.pspace_cuda_create(ctypes.byref(h), 0, 0, 1, kinds.ctypes.data, p.ctypes.data, p.ctypes.data, d.ctypes.data, 0)
    assert rc == -2 and b"CUDA" in L.pspace_cuda_last_error()
    with pytest.raises(cabi.PspaceCudaError):
        cabi.check(rc)


def test_host_container_aborts_without_gpu():
    """ParameterContainer::initialize() has no CPU path: the process reports and aborts."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from pspace_b200 import PSPACE\n"
            "ns = PSPACE.bind(False); f = ns.PyParameterFactory(); pc = ns.PyParameterContainer(0)\n"
            "pc.addParameter(f.createNormalParameter(0.0, 1.0, 2)); pc.initialize(); print('survived')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "survived" not in r.stdout
    assert "pspace" in r.stderr and "failed" in r.stderr


def test_libpspace_exports_reference_cxx_symbols():
    """The symbols the reference's compiled wrapper imports (nm -D probe)."""
    for sub in ("", "complex"):
        out = subprocess.check_output(["nm", "-DC", os.path.join(LIB, sub, "libpspace.so")], text=True)
        sc = "std::complex<double>" if sub else "double"
        for sym in ["ParameterFactory::ParameterFactory()",
                    f"ParameterFactory::createNormalParameter({sc}, {sc}, int)",
                    f"ParameterFactory::createUniformParameter({sc}, {sc}, int)",
                    f"ParameterFactory::createExponentialParameter({sc}, {sc}, int)",
                    f"ParameterFactory::createNormalParameter({sc}, {sc})",
                    "ParameterContainer::ParameterContainer(int, int)",
                    "ParameterContainer::addParameter(AbstractParameter*)",
                    "ParameterContainer::initialize()",
                    "ParameterContainer::initializeBasis(int const*)",
                    "ParameterContainer::initializeQuadrature(int const*)",
                    f"ParameterContainer::quadrature(int, {sc}*, {sc}*)",
                    f"ParameterContainer::basis(int, {sc}*)",
                    "ParameterContainer::getNumBasisTerms()",
                    "ParameterContainer::getNumParameters()",
                    "ParameterContainer::getNumQuadraturePoints()",
                    "ParameterContainer::getBasisParamDeg(int, int*)",
                    "ParameterContainer::getBasisParamMaxDeg(int*)",
                    "AbstractParameter::getParameterID()", "AbstractParameter::setMaxDegree(int)"]:
            assert f" T {sym}" in out, (sub, sym)


def _ulps(a, b):
    a, b = np.asarray(a).view(np.float64), np.asarray(b).view(np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.spacing(np.abs(b)), 1e-300))


@pytest.mark.parametrize("cplx", [False, True])
def test_host_parameters_against_golden(cplx):
    """AbstractParameter::quadrature / ::basis on the host (same templates as the device tables)."""
    from pspace_b200 import PSPACE
    ns = PSPACE.bind(cplx)
    g = load_golden("onedim_z" if cplx else "onedim")
    ps = golden_params(g)
    f = ns.PyParameterFactory()
    make = {"normal": f.createNormalParameter, "uniform": f.createUniformParameter, "exponential": f.createExponentialParameter}
    for pid, (kind, a, b, _) in enumerate(ps):
        if not cplx:
            a, b = float(np.real(a)), float(np.real(b))
        p = make[kind](a, b, 3)
        for n in range(1, 21):
            key = f"rule_{pid}_{n}"
            if key not in g.files:
                continue
            got = np.stack(p.quadrature(n))
            if cplx:
                assert np.max(np.abs(got - g[key])) <= 1e-15 * np.max(np.abs(g[key])), key
            else:
                assert same_bits(got, g[key]), key
        zs = g[f"z_{pid}"]
        for d in range(13):
            got = np.array([p.basis(z, d) for z in zs], dtype=ns.dtype)
            ref = g[f"poly_{pid}"][d]
            if cplx:
                assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref)), (pid, d)
            elif kind == "uniform" and d >= 5:
                assert _ulps(got, ref) <= 4, (pid, d)       # double-double integer power vs libm pow
            else:
                assert same_bits(got, ref), (pid, d)


def test_workload_sizes_match_survey_table():
    from pspace_b200 import workloads as WL
    sizes = {k: (f().N, f().Q, f().m) for k, f in WL.ALL.items()}
    assert sizes == {1: (64, 64, 1), 2: (126, 3125, 8), 3: (1296, 1296, 24), 4: (65536, 65536, 24), 5: (286, 1048576, 24)}
    assert WL.cfg4().complex_ and WL.cfg5().seed == 20265


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (CPU arm of the driver's comparison) on the small cfg2 workload, shortened."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2", "--steps", "1",
                        "--warmup", "0", "--ref-seconds", "0.3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "updates/s" and line["value"] > 1e5
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    # rank != 0 under torchrun exits silently
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                        text=True, timeout=60, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_header_is_plain_c_and_matches_the_ctypes_binding(tmp_path):
    """The boundary is a C ABI: include/pspace_cuda.h must compile as C99, and the struct the Python binding passes
    (ProjectArgs) must have the size and field offsets the C compiler gives pspace_project_args."""
    import ctypes
    from pspace_b200.cabi import ProjectArgs
    src = tmp_path / "abi.c"
    fields = [f[0] for f in ProjectArgs._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(pspace_project_args, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pspace_cuda.h"\n'
                   'int main(void){ printf("%zu\\n", sizeof(pspace_project_args)); ' + body + ' return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert out[0] == ctypes.sizeof(ProjectArgs)
    assert out[1:] == [getattr(ProjectArgs, f).offset for f in fields]


def test_dense_dmma_loop_keeps_the_fast_ptxas_schedule():
    """Static guard (no GPU): ptxas emits one of two schedules for the DMMA loop of the dense 128x96 real tile depending on
    unrelated code in the kernel; the slower one (about 40 more static stall cycles per 32 points and warp) measures 1.2 % slower on
    the B200.  tools/sass_loop_stats.py decodes the control words of the built library."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    tile = "Li2ELi12ELi8ELi1ELb0ELi0ELi4ELi1EEE"                        # Tile<2, 12, 8, 1, real, matrix, 4, 1>
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "sass_loop_stats.py"),
                                   tile + "Li0ELb0E", tile + "Li0ELb1E", tile + "Li2ELb0E"], text=True)

    def stats(suffix):
        m = re.search(re.escape(tile + suffix) + r".*DMMA\s+(\d+)\s+static stall cycles/chunk\s+(\d+)", out)
        assert m, out
        return int(m.group(1)), int(m.group(2))

    # dense, full last chunk: 16 k-steps (64-point stage) x 2 row tiles x 12 column tiles; 384 x 16 = 6144 is the floor,
    # 6256 today (the slower schedule: >= 6300)
    assert stats("Li0ELb0E")[0] == 384 and stats("Li0ELb0E")[1] <= 6290, out
    # trimmed last chunk (point counts off the stage size, cfg3): 6338 as first written, 6257 with the (inactive)
    # start-of-segment lag statement that only this and the next instantiation contain (cfg3 71.1 -> 70.6 ms per element)
    assert stats("Li0ELb1E")[0] == 384 and stats("Li0ELb1E")[1] <= 6290, out
    # fused multi-GPU epilogue (EPI_PEER): 6254 (6308 without the statement: 1.5 % at 8 GPUs)
    assert stats("Li2ELb0E")[0] == 384 and stats("Li2ELb0E")[1] <= 6290, out


def test_hot_kernels_do_not_spill():
    """Static guard (no GPU): the DMMA kernels hold their accumulators and fragments in registers; a local-memory frame in one of
    them means ptxas spilled (seen twice during development: a `__noinline__` helper, a register-resident tail table)."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "pspace_b200", "lib", "libpspace_cuda.so")
    out = subprocess.check_output(["cuobjdump", "-res-usage", lib], text=True)
    seen = 0
    for name, stack in re.findall(r"Function (\S+):\s*\n\s*REG:\d+ STACK:(\d+)", out):
        if "k_project_ws" in name or "k_expand_f" in name:
            seen += 1
            assert int(stack) == 0, (name, stack)
        elif "k_project_vf" in name:
            seen += 1
            assert int(stack) <= 16, (name, stack)      # two variants keep one 8/16-byte temporary
    assert seen >= 40, seen


def test_cpp_caller_example_builds_against_the_reference_headers():
    """examples/stochastic_element.cpp — a caller shaped like the reference's TACSStochasticElement — compiles and links
    against host/include + libpspace.so in both scalar builds, and (no CPU path) aborts loudly without a GPU."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for exe in ("stochastic_element", "stochastic_element_z"):
        path = os.path.join(ROOT, "examples", exe)
        assert os.path.exists(path)
        import torch
        if not torch.cuda.is_available():
            run = subprocess.run([path], capture_output=True, text=True)
            assert run.returncode != 0 and "CUDA" in (run.stdout + run.stderr)


def test_basis_count_is_host_arithmetic_and_matches_golden_and_oracle():
    """pspace_cuda_basis_count sizes the caller's degree table without touching the device (an exact count, not the
    reference's overflowing factorial, BasisHelper.cpp:361-367): it must equal the row counts of the golden degree
    tables (compiled reference), of the extended oracle beyond five variables, and SURVEY.md's table of sizes."""
    import ctypes
    from oracle import harness as H
    from pspace_b200 import cabi
    from pspace_b200.workloads import mixed_params
    L = ctypes.CDLL(cabi.LIB_PATH)
    L.pspace_cuda_basis_count.restype = ctypes.c_int
    L.pspace_cuda_last_error.restype = ctypes.c_char_p

    def count(pmax, bt):
        a = np.ascontiguousarray(pmax, dtype=np.intc)
        n = ctypes.c_longlong(-1)
        rc = L.pspace_cuda_basis_count(ctypes.c_int(len(a)), ctypes.c_void_p(a.ctypes.data), ctypes.c_int(bt), ctypes.byref(n))
        return rc, n.value

    g = load_golden("basis_degrees")
    for idx in range(int(g["ncases"])):
        rc, n = count(g[f"pmax_{idx}"], int(g[f"bt_{idx}"]))
        assert rc == 0 and n == g[f"deg_{idx}"].shape[0]
    for nv, p, bt in ((6, 2, 1), (7, 2, 0), (8, 3, 0), (10, 3, 1), (9, 1, 0), (12, 2, 1)):
        o = H.CpuContainer(H.oracle_api(False), mixed_params(nv, p), bt).initialize_basis([p] * nv)
        assert count([p] * nv, bt) == (0, o.N)
    # SURVEY.md section 8 table
    assert count([3, 3, 3], 0) == (0, 64)
    assert count([4] * 5, 1) == (0, 126)
    assert count([5] * 4, 0) == (0, 1296)
    assert count([3] * 8, 0) == (0, 65536)
    assert count([3] * 10, 1) == (0, 286)
    # mixed caps: complete basis keeps sum <= max(pmax) with the per-variable caps still applied
    assert count([4, 3], 1) == (0, 14)
    # loud on bad input
    rc, _ = count([3] * 33, 0)
    assert rc != 0 and L.pspace_cuda_last_error()
    rc, _ = count([3, -1], 0)
    assert rc != 0
