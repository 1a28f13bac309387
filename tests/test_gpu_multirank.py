"""Multi-rank GPU test (-m gpu, needs two GPUs): image-sharded loss + the normaliser all-reduce over NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_loss_allreduce_nccl():
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py")]
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, BG_MULTIRANK_OUT=tmp))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        res = [json.load(open(os.path.join(tmp, "rank%d.json" % k))) for k in range(2)]
    assert all(d["current_device"] == 0 for d in res)           # nobody called set_device: ops followed the tensors
    for form in ("decoded", "raw"):
        big = res[0][form]["big"]
        for rk in res:
            assert rk[form]["grad_ok"]
            assert abs(rk[form]["combined"] - big) <= 1e-6 * abs(big), (form, rk)
            assert rk[form]["big"] == big                       # every rank computed the same big-batch loss
        assert res[0][form]["local"] != res[1][form]["local"]   # the shards really differ
    print("NCCL sharded loss:", res)
