"""The real multi-process slab path -- one rank per GPU, NCCL migration / column tables, CUDA-IPC peer-memory halo or NCCL halo
messages -- against a single-context run (scripts/check_multi_gpu.py under torchrun). Needs two GPUs on the box; the same
comparison also runs inside `bench.py --gpus N` (parity_check in the scaling lines)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p,workload,axis", [("1", "config2", 0), ("0", "config2", 0), ("1", "config4", 2), ("0", "config4", 2), ("0", "config4", 0)])
def test_slabs_over_nccl_match_single_context(p2p, workload, axis):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs on the box (bench.py --gpus N runs the same check as parity_check)")
    env = dict(os.environ, CCB_HALO_P2P=p2p)
    port = 29600 + int(p2p) + 2 * axis + (10 if workload == "config4" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "check_multi_gpu.py"), "--workload", workload, "--axis", str(axis)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith('{"check"')]
    assert r.returncode == 0 and line, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(line[-1])
    assert res["ok"] and res["btrack_equal"] and res["worst_rel_err"] < 1e-11, res
