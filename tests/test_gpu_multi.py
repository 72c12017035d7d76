"""Multi-GPU checks (need >= 2 CUDA devices; skipped on a single-GPU box): the one real collective of the path — the fitting
loop's all-reduce of its objective partials (fitting.py, BASELINE config 5) — over NCCL against the un-sharded loop."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_sharded_fit_over_nccl_equals_whole_batch_fit(tmp_path):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    out = tmp_path / "fit.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_fit_worker.py"), str(out)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-3000:]
    res = json.load(open(out))
    assert res["fused"]
    # same arithmetic per hand; only the fp64 partial sums are added in another order (ranks instead of blocks)
    assert res["max_param_diff"] < 1e-5, res
    for a, b in zip(res["loss_sharded"], res["loss_whole"]):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), res
