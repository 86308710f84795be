"""GPU tier (-m gpu): the row-partitioned mode (BASELINE.json configs[1] at N > 1).

* one GPU: tests/dist_selftest.py runs the partitioned code path with world = 1 (collectives degenerate to
  copies) against the compiled reference / numpy oracle and the ordinary single-GPU path;
* two GPUs (skipped on a one-GPU box): tests/dist_gpu_check.py under torchrun with NCCL.
Both run in subprocesses: the communicator is process-wide state.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partitioned_code_path_on_one_gpu(gpu):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_selftest.py")], capture_output=True, text=True,
                       timeout=1500, cwd=ROOT)
    assert r.returncode == 0 and "dist selftest ok" in r.stdout, r.stdout[-4000:] + r.stderr[-3000:]


def test_row_partitioned_two_gpus(gpu):
    """2-rank NCCL run of tests/dist_gpu_check.py: the row-partitioned solve (Anderson acceleration on) agrees
    with the compiled reference / oracle (status, objectives 1e-6) and with the single-GPU solve on cone QP /
    LASSO / SOCP / SDP, and a verbose run that stops at max_iters does not hang."""
    if gpu < 2:
        pytest.skip("needs 2 GPUs on the box (one-GPU coverage: test_row_partitioned_code_path_on_one_gpu)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29671",
                        os.path.join(ROOT, "tests", "dist_gpu_check.py")],
                       capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0 and "dist check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
