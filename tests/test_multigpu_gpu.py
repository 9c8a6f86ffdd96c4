"""Multi-GPU paths (run only when the box exposes >= 2 devices): node-partitioned GENConv with
NCCL halo exchange, and the batch-sharded dense bench line."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29533", script] + list(args)
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_partitioned_genconv_matches_single_gpu():
    r = _torchrun(2, os.path.join(ROOT, "tests", "multigpu_sparse_check.py"))
    assert r.returncode == 0 and "MULTIGPU_SPARSE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_bench_two_gpus_weak_scaling_line():
    r = _torchrun(2, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "20", "--warmup", "3", "--cpu-seconds",
                  "0.5")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["value"] > 0
