"""Slab decomposition on the GPU with the list built on the device (subzero_b200.slabs.DeviceSlab): world-size-2/3 (and 4
when the box has the GPUs) runs of a COUPLED time loop -- contact step + integrator, floes moving, rotating, thinning, crossing
the periodic boundary and the slab edges -- must reproduce the single-GPU loop bit for bit every step: per-floe forces, torques,
overlap areas, stress, kill/transfer, every contact row, and the integrated state.  With fewer GPUs than ranks the ranks share
cuda:0 and exchange over gloo (host-staged); with enough GPUs it is NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run(world, n, seed, kind, steps=6):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world * 7 + seed), os.path.join(HERE, "slab_worker.py"), str(n), str(seed), kind, str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    assert line and line[0].startswith("RESULT OK"), r.stdout[-3000:]
    return line[0]


def test_device_slab_world1_equals_plain_step():
    """the device-built list with a single rank (no halo) against the library's own K0 on the same field, over coupled steps"""
    import subzero_b200 as sz
    from subzero_b200 import slabs
    prm, field = sz.voronoi_field(4000, seed=4)
    prm.dt = 10.0
    field.u[:] *= 100.0
    field.v[:] *= 100.0
    mass = field.area * field.h * 920.0
    inertia = mass * field.rmax ** 2 / 4
    dev = torch.device("cuda", 0)
    with sz.ContactContext(0) as a, sz.ContactContext(0) as b:
        slab = slabs.DeviceSlab(prm, field, np.arange(1, field.n + 1, dtype=np.int32), field.n, slabs.Comm(None, 0, 1, dev), a)
        slab.trajectory_init(mass, inertia, nz=2, dXi_p=field.u, dYi_p=field.v)
        b.upload(prm, field)
        b.trajectory_init(mass, inertia, nz=2, dXi_p=field.u, dYi_p=field.v)
        for it in range(4):
            s = slab.run()
            s1 = b.step_resident()
            o, o1 = slab.outputs(), b.floe_outputs()
            for k in o1:
                assert np.array_equal(o[k], o1[k], equal_nan=True), (it, k)
            off, rows = slab.rows()
            off1, rows1 = b.rows()
            assert np.array_equal(off, off1[:field.n + 1]) and np.array_equal(rows, rows1[:off1[field.n]], equal_nan=True)
            _, nl = slab.positions()
            assert nl == s1.n and s.n_pairs == s1.n_pairs and s.collision_count == s1.collision_count
            assert slab.trajectory_step(prm.dt, 1e-4) == b.trajectory_step(prm.dt, 1e-4)
            st, st1 = a.trajectory_state(nverts=field.vx.shape[0]), b.trajectory_state(nverts=field.vx.shape[0])
            for k in st1:
                assert np.array_equal(st[k], st1[k], equal_nan=True), (it, k)
        assert s1.n > s1.n0


def test_two_slabs_voronoi_coupled():
    print(run(2, 20000, 1, "voronoi"))


def test_three_slabs_voronoi_coupled():
    print(run(3, 6000, 2, "voronoi"))


def test_three_slabs_with_floe_migration():
    """DeviceSlab.repartition(): floes that crossed a slab edge move to the new owner with their whole state (rotated outline,
    integrator history, stress history) and keep their global numbers; the coupled loop stays bit-identical to one GPU"""
    out = run(3, 9000, 7, "migrate", steps=6)
    assert "migrated=0" not in out


def test_two_slabs_walls_coupled():
    """non-periodic domain: wall contacts and floes leaving the domain, resolved by the owner of each floe"""
    print(run(2, 8000, 5, "walls"))


def test_two_slabs_real_shapes_with_merges():
    out = run(2, 14, 1, "real", steps=2)
    assert "kill_events=0" not in out


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs for the NCCL path at world 4")
def test_four_slabs_nccl():
    print(run(4, 40000, 3, "voronoi"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs for the NCCL path")
def test_two_slabs_nccl():
    print(run(2, 30000, 6, "voronoi"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs for the NCCL path")
def test_two_slabs_nccl_cuda_graph():
    """the whole step -- library kernels and NCCL collectives -- captured once and replayed: same bits as one GPU, every step of
    a coupled loop (the graph reads counts from device memory, so moving floes do not invalidate it)"""
    out = run(2, 30000, 8, "graph", steps=8)
    assert "graph_replays=0" not in out
