"""Slab decomposition on the GPU: world-size-2 (and 4 when the box has the GPUs) runs must reproduce the single-GPU
step bit for bit -- per-floe forces, torques, overlap areas, stress, kill/transfer and every contact row.  With fewer
GPUs than ranks the ranks share cuda:0 and exchange halos over gloo (host-staged); with enough GPUs it is NCCL."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run(world, n, seed, kind):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world * 7 + seed), os.path.join(HERE, "multi_worker.py"), str(n), str(seed), kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    assert line and line[0].startswith("RESULT OK"), r.stdout[-3000:]
    return line[0]


def test_two_slabs_voronoi():
    print(run(2, 20000, 1, "voronoi"))


def test_three_slabs_voronoi():
    print(run(3, 6000, 2, "voronoi"))


def test_two_slabs_real_shapes_with_merges():
    out = run(2, 14, 1, "real")
    assert "kill_events=0" not in out


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs for the NCCL path at world 4")
def test_four_slabs_nccl():
    print(run(4, 40000, 3, "voronoi"))
