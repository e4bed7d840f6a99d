"""Multi-GPU parity where the driver runs it (`pytest -m gpu`): the torchrun scripts are self-launched on
min(device_count, 8) GPUs (skipped on a one-GPU box), every error asserted < 1e-12 against the CPU ORACLE
(tests/run_sharded_check.py: NCCL allreduce, reduce-scatter, host-input slab path, vector projection, fused peer-memory
exchange incl. a repeated call that must be bitwise equal) and against the rank-order sum at cfg5 size
(tests/run_fused_check.py).  The C++ caller of the C-ABI collective (examples/sharded_projection.cpp: NCCL through
include/pspace_cuda.h section (e), no Python in the data path) runs on every box: with one rank on a single GPU."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return min(torch.cuda.device_count(), 8)


def _torchrun(script, n, *args, timeout=900):
    port = 29600 + (os.getpid() % 1500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    return r, lines


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_projection_matches_the_oracle_on_all_gpus():
    r, lines = _torchrun("run_sharded_check.py", _ngpu())
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert len(lines) == 2, r.stdout
    for ln in lines:
        assert ln["pass"], ln
        for k in ("allreduce_err", "reduce_scatter_err", "host_input_err", "vector_err", "fused_peer_err", "cabi_allreduce_err"):
            assert ln[k] < 1e-12, (k, ln)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_peer_exchange_equals_rank_order_sum_at_cfg5_size():
    r, lines = _torchrun("run_fused_check.py", _ngpu(), "8")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ln = lines[-1]
    assert ln["nccl_vs_rank_order_sum"] < 1e-12 and ln["fused_vs_rank_order_sum"] < 1e-14 and ln["fused_repeatable"], ln


@pytest.mark.parametrize("nranks", [1, 0])
def test_cpp_caller_of_the_cabi_collective(nranks):
    """examples/sharded_projection.cpp --check: q-shards on `nranks` GPUs (0 = all visible), NCCL allreduce and
    reduce-scatter through pspace_cuda_project_matrix_allreduce / _reduce_scatter, every rank <= 1e-12 against the
    single-GPU projection."""
    if nranks == 0 and torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = os.path.join(ROOT, "examples", "sharded_projection")
    if not os.path.exists(exe):
        b = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "sharded_projection"], capture_output=True, text=True)
        assert b.returncode == 0, b.stdout + b.stderr
    n = nranks if nranks else _ngpu()
    r = subprocess.run([exe, "--nranks", str(n), "--check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr
