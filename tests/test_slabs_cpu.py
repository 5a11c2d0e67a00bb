"""The slab decomposition's host logic with world_size 2 and 3 over gloo (CPU tensors): global numbering of the
periodic images, halo selection/exchange and the merged local lists, checked against the oracle's global extended
list and candidate pairs.  The GPU step itself is covered by tests/test_gpu_multi.py (-m gpu)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import subzero_b200 as sz
from subzero_b200 import slabs

N_FLOES, SEED = 1200, 21


def _field(world, balance):
    """the test field renumbered slab by slab.  balance: a field of non-uniform density (three quarters of the floes east
    of x = 0 removed) cut at the quantiles of the centroids instead of at equal widths"""
    prm, field = sz.voronoi_field(N_FLOES, seed=SEED)
    if balance:
        keep = np.nonzero((field.x < 0) | (np.arange(field.n) % 4 == 0))[0]
        field = slabs.select(field, keep)
    field, starts = slabs.sort_by_slab(field, prm.Lx, world, balance=balance)
    return prm, field, starts


def _worker(rank, world, port, outdir, periodic, balance=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prm, field, starts = _field(world, balance)
        N_FLOES = field.n
        mine = slabs.take_range(field, int(starts[rank]), int(starts[rank + 1]))
        st = slabs.SlabState.from_soa(mine, int(starts[rank]), N_FLOES, torch.device("cpu"))
        comm = slabs.Comm(dist, rank, world, torch.device("cpu"))
        reach = 2.0 * float(comm.all_gather(st.rmax.max().reshape(1)).max())
        L = slabs.build_local_list(st, prm.Lx, prm.Ly, periodic, reach, comm)
        # the kill/transfer fix-up with synthetic merge events: entry gid g "kills" floe (g*7 % N)+1 on every 97th owned entry
        kill_i = torch.zeros_like(L.gid)
        ev = (L.owned != 0) & (L.gid % 97 == 0)
        kill_i[ev] = (L.gid[ev] * 7) % N_FLOES + 1
        k, t = slabs.fix_kill_transfer(L.gid, L.floe_num, L.owned, kill_i, torch.zeros_like(L.gid), st.id0, st.n, comm)
        np.savez(os.path.join(outdir, "r%d.npz" % rank), gid=L.gid.numpy(), floe_num=L.floe_num.numpy(), x=L.x.numpy(), y=L.y.numpy(), root_x=L.root_x.numpy(),
                 root_y=L.root_y.numpy(), body=L.body.numpy(), alive=L.alive.numpy(), owned=L.owned.numpy(), parent=L.parent.numpy(), voff=L.voff.numpy(),
                 vx=L.vx.numpy(), vy=L.vy.numpy(), n_ext=L.n_ext_global, id0=st.id0, n_own=st.n, kill=k.numpy(), transfer=t.numpy(), kill_i=kill_i.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,periodic,balance", [(2, True, False), (3, True, False), (2, False, False), (3, True, True)])
def test_local_lists_reproduce_the_global_extended_list(tmp_path, world, periodic, balance):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), periodic, balance), nprocs=world, join=True)
    R = [np.load(tmp_path / ("r%d.npz" % r)) for r in range(world)]
    prm, field, starts = _field(world, balance)
    N_FLOES = field.n
    prm.periodic = int(periodic)
    if balance:
        # slabs cut at the quantiles: equal floe counts although the eastern half holds a quarter of the density
        assert np.diff(starts).max() - np.diff(starts).min() <= 1 and N_FLOES < 0.7 * 1200
        width = np.array([field.x[starts[k]:starts[k + 1]].max() - field.x[starts[k]:starts[k + 1]].min() for k in range(world)])
        assert width.max() > 1.5 * width.min()
    ref = oracle.OracleStep(prm, field, broad_mode=0)
    n0, n = ref.summary.n0, ref.summary.n
    g = ref.ghosts()
    gx = np.concatenate([field.x, g["x"]])
    gy = np.concatenate([field.y, g["y"]])
    gnum = np.concatenate([np.arange(1, n0 + 1), g["floe_num"]])
    gpar = np.concatenate([np.zeros(n0, np.int64), g["parent"]])
    # 1. every entry of the global list is owned exactly once, with the reference's number, image centroid and parent
    seen = np.zeros(n, int)
    for r in R:
        assert int(r["n_ext"]) == n
        assert np.all(np.diff(r["gid"]) > 0)
        own = r["owned"] != 0
        seen[r["gid"][own]] += 1
        np.testing.assert_array_equal(r["floe_num"], gnum[r["gid"]])
        np.testing.assert_array_equal(r["x"], gx[r["gid"]])
        np.testing.assert_array_equal(r["y"], gy[r["gid"]])
        root = np.abs(r["floe_num"]) - 1
        np.testing.assert_array_equal(r["root_x"], field.x[root])
        np.testing.assert_array_equal(r["root_y"], field.y[root])
        np.testing.assert_array_equal(r["body"][:, 0], field.rmax[root])
        np.testing.assert_array_equal(r["body"][:, 2], field.area[root])
        has = r["parent"] > 0
        assert np.all(own[has]) and np.all(r["floe_num"][has] < 0)
        np.testing.assert_array_equal(r["gid"][r["parent"][has] - 1] + 1, gpar[r["gid"][has]])
        assert np.all(has[own & (r["floe_num"] < 0)])                # every owned image has its parent here
        # originals are owned by the rank whose id range holds them
        oo = own & (r["floe_num"] > 0)
        assert np.all((r["gid"][oo] >= int(r["id0"])) & (r["gid"][oo] < int(r["id0"]) + int(r["n_own"])))
        # outlines travelled intact
        for e in range(0, len(root), 7):
            a, b = r["voff"][e], r["voff"][e + 1]
            fx_, fy_ = field.outline(root[e])
            np.testing.assert_array_equal(r["vx"][a:b], fx_)
            np.testing.assert_array_equal(r["vy"][a:b], fy_)
    assert np.all(seen == 1)
    # 2. every candidate pair of the reference is resolvable on the ranks that own its floes
    pr = ref.pairs()
    owner = np.empty(n, int)
    for k, r in enumerate(R):
        owner[r["gid"][r["owned"] != 0]] = k
    have = [set(r["gid"].tolist()) for r in R]
    assert len(pr["i"]) > (1.5 if balance else 4) * N_FLOES
    straddling = 0
    for i, j in zip(pr["i"] - 1, pr["j"] - 1):
        for k in {owner[i], owner[j]}:
            assert i in have[k] and j in have[k], (i, j, k)
        straddling += owner[i] != owner[j]
    assert straddling > 0
    # the halo is a thin layer, not the whole field
    for r in R:
        assert (r["owned"] == 0).sum() < (0.6 if balance else 0.45) * len(r["gid"])     # (balanced: 250 floes per rank in narrow slabs)
    # 3. kill/transfer fix-up across ranks == the serial loop of floe_interactions_all.m:175-179
    kill_all = np.zeros(n, np.int64)
    for r in R:
        own = r["owned"] != 0
        kill_all[r["gid"][own]] = r["kill_i"][own]
    transfer = np.zeros(n0, np.int64)
    for i in range(n):
        if kill_all[i] > 0 and kill_all[i] != i + 1:
            transfer[kill_all[i] - 1] = i + 1
    got_k = np.concatenate([r["kill"] for r in R])
    got_t = np.concatenate([r["transfer"] for r in R])
    np.testing.assert_array_equal(got_k, kill_all[:n0])
    np.testing.assert_array_equal(got_t, transfer)
    assert (transfer > 0).sum() > (1 if balance else 3)


def test_sort_by_slab_is_a_stable_renumbering():
    prm, field = sz.voronoi_field(500, seed=2)
    out, starts = slabs.sort_by_slab(field, prm.Lx, 4)
    assert starts[0] == 0 and starts[-1] == 500 and np.all(np.diff(starts) > 0)
    s = slabs.slab_of(out.x, prm.Lx, 4)
    assert np.all(np.diff(s) >= 0)
    assert np.isclose(out.area.sum(), field.area.sum())
    i = 123
    x, y = out.outline(i)
    assert x[0] == x[-1] and len(x) == out.voff[i + 1] - out.voff[i]
    one = slabs.take_range(out, int(starts[1]), int(starts[2]))
    assert one.n == starts[2] - starts[1] and one.voff[0] == 0 and one.voff[-1] == one.vx.shape[0]
    # balanced cuts: the same floes, equal counts, a floe on an edge goes to the upper slab, NaN centroids to slab 0
    bal, bstarts = slabs.sort_by_slab(field, prm.Lx, 4, balance=True)
    assert list(np.diff(bstarts)) == [125, 125, 125, 125] and np.isclose(bal.area.sum(), field.area.sum())
    edges = slabs.slab_edges(field.x, prm.Lx, 4, balance=True)
    assert np.all(np.diff(slabs.slab_of(bal.x, prm.Lx, 4, edges)) >= 0)
    assert list(slabs.slab_of([edges[0], np.nan, -prm.Lx, prm.Lx], prm.Lx, 4, edges)) == [1, 0, 0, 3]
    assert np.allclose(slabs.slab_edges(field.x, prm.Lx, 4), [-prm.Lx / 2, 0, prm.Lx / 2])
