"""GPU, >= 2 devices (self-skips otherwise): the data-parallel paths on real peer memory / NCCL, run as `torchrun tools/dp_check.py`.

Checked: three critic steps (bf16 whole-step kernel, gradient all-reduce inside the kernel over NVLink peer memory) and three
frozen-critic Hourglass steps (cgs_p2p_stage + cgs_p2p_allreduce_adam) on 2 ranks leave BIT-IDENTICAL parameters on every rank, agree
with the NCCL all-reduce variant and with one process training on the global batch (reference main.py:185-200 / 344-463 semantics:
the gradient of the global-batch mean), and stay identical through CUDA-graph replays.  The host-side sharding logic has its own
CPU tests (tests/test_dp.py, gloo)."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_two_ranks_bit_identical_and_equal_to_global_batch():
    env = dict(os.environ, DP_CHECK_TIMEOUT="150")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(ROOT, "tools", "dp_check.py")], capture_output=True, text=True, timeout=300, env=env)
    out = r.stdout + r.stderr
    assert "done" in out, out[-3000:]
    assert out.count("params equal across ranks: True") >= 2, out[-3000:]                      # rank 0 and rank 1
    assert out.count("params equal across ranks after graph replays: True") >= 2, out[-3000:]
    assert out.count("hourglass params equal across ranks: True") >= 2, out[-3000:]
    for tag, tol in (("DP vs single-process global batch", 2e-6), ("hourglass DP vs single-process global batch", 2e-5), ("p2p vs NCCL", 2e-6)):
        m = re.search(re.escape(tag) + r": max \|dparam\| = ([0-9.e+-]+)", out)
        assert m is not None, (tag, out[-3000:])
        assert float(m.group(1)) <= tol, (tag, m.group(1))
