"""The sharded operator behind the C ABI (include/cozo_gpu.h: cozo_gpu_shards_*, cozo_gpu_hnsw_*_sharded),
one process per GPU.  World size 1 runs everywhere; world size 2 needs two GPUs (`gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_sharded_abi_world1(gpu):
    out = _run(1, 29611)
    assert "world=1" in out


def test_sharded_abi_world2(gpu):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run(2, 29612)
    assert "world=2" in out
