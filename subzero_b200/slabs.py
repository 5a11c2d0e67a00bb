"""Work partition of the contact step over the GPUs of one box: spatial slabs with a halo exchange.

One process per GPU.  The domain [-Lx, Lx) is cut into `world` slabs along x (equal widths, or the quantiles of the centroids
for a field of non-uniform density); a floe is owned by the rank whose slab held its centroid when the field was (re)partitioned
-- the floes are numbered slab by slab, so a rank usually owns a contiguous range of the global floe numbers, but the device
path works for any ascending subset (after `DeviceSlab.repartition` moved floes between ranks, numbers stay what they were).

The product path is `DeviceSlab`: the list surgery runs in the library's kernels (sz_slab_prepare / _pack / _build) every step,
from the current state, around two collectives (an all-gather of small per-rank meta records and an all-to-all of fixed-size
blocks with the halo entries and their current outlines), all on one CUDA stream.

`build_local_list` / `fix_kill_transfer` are the same list logic written with torch tensor operations: the reference
implementation that runs under gloo on CPU tensors (tests/test_slabs_cpu.py: against the oracle's global extended list and
candidate pairs) and that the GPU worker compares the device-built list with (tests/slab_worker.py).

What every rank's list holds, and why results are the single-GPU ones:
  1. the periodic images of ITS floes (floe_interactions_all.m:16-66), numbered in the GLOBAL extended list [originals |
     x-ghosts | y-ghosts] -- so every `j > i` comparison, partner id and row order is the single-GPU one;
  2. every entry of another rank that lies within reach (2 max(rmax)) of this rank's x-extent: the halo;
  3. every pair with at least one owned floe is resolved locally -- pairs straddling two slabs on both sides (about
     2 * reach / slab_width of all pairs), which keeps each floe's rows bit-identical to the single-GPU result and replaces the
     return exchange of partial forces; the forces on an owned periodic image are folded into its parent by the owner (:242-245);
  4. the serial kill/transfer fix-up of :175-179 runs across ranks on the (rare) merge events.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import abi
from .contact import ContactContext
from .field import voronoi_field

F64, I64 = torch.float64, torch.int64
N_STATE = 14          # gid, floe_num, x, y, root_x, root_y, rmax, h, area, u, v, ksi, alive, nverts


# ------------------------------------------------------------------------------------------------ host-side layout
def slab_edges(x, Lx, world, balance=False):
    """interior slab edges [world - 1] over [-Lx, Lx): equal widths, or (balance) the quantiles of the centroids, so that
    every rank owns the same number of floes when the density is not uniform (the halo logic only needs each
    rank's entries to be contiguous in x)"""
    x = np.asarray(x, np.float64)
    xs = np.sort(x[~np.isnan(x)])
    if not balance or xs.shape[0] == 0:
        return -Lx + (2.0 * Lx / world) * np.arange(1, world)
    n = xs.shape[0]
    return np.array([xs[min(n - 1, (k * n) // world)] for k in range(1, world)])


def slab_of(x, Lx, world, edges=None):
    """slab index of a centroid: equal-width slabs, or the slabs between `edges` (a floe on an edge belongs to the upper slab)"""
    if edges is None:
        w = 2.0 * Lx / world
        s = np.floor((np.asarray(x) + Lx) / w)
        s = np.where(np.isnan(s), 0, s)
        return np.clip(s, 0, world - 1).astype(np.int64)
    x = np.asarray(x, np.float64)
    s = np.searchsorted(np.asarray(edges, np.float64), x, side="right")
    return np.where(np.isnan(x), 0, s).astype(np.int64)


def select(soa, order):
    """the floes `order` (indices) of a FloesSoA, in that order, with their outlines"""
    return soa.take(order)


def sort_by_slab(soa, Lx, world, balance=False):
    """Renumber the floes slab by slab (stable).  Returns (permuted FloesSoA, id_start [world+1]).  balance: slab edges at
    the quantiles of the centroids instead of equal widths."""
    slab = slab_of(soa.x, Lx, world, slab_edges(soa.x, Lx, world, True) if balance else None)
    order = np.argsort(slab, kind="stable")
    out = select(soa, order)
    starts = np.searchsorted(slab[order], np.arange(world + 1), side="left").astype(np.int64)
    return out, starts


def take_range(soa, a, b):
    v0, v1 = int(soa.voff[a]), int(soa.voff[b])
    return abi.FloesSoA(*(getattr(soa, k)[a:b] for k in abi.FloesSoA.FIELDS), soa.alive[a:b], (soa.voff[a:b + 1] - v0).astype(np.int32), soa.vx[v0:v1], soa.vy[v0:v1])


# ------------------------------------------------------------------------------------------------ communication
class Comm:
    """the three exchanges the slab step needs, over torch.distributed (nccl or gloo) or nothing (world 1)"""

    def __init__(self, dist, rank, world, device):
        self.dist, self.rank, self.world, self.device = dist, rank, world, device
        # gloo moves CPU tensors only: CUDA tensors are staged through the host (tests that emulate two ranks on one GPU)
        self.stage = world > 1 and dist.get_backend() == "gloo" and torch.device(device).type == "cuda"

    def all_gather(self, t):
        """t [k] -> [world, k]"""
        if self.world == 1:
            return t.unsqueeze(0)
        src = t.cpu() if self.stage else t
        out = [torch.empty_like(src) for _ in range(self.world)]
        self.dist.all_gather(out, src.contiguous())
        return torch.stack(out).to(t.device)

    def all_max(self, t):
        """elementwise maximum over the ranks (one all-reduce)"""
        if self.world == 1:
            return t
        src = t.cpu() if self.stage else t.clone()
        self.dist.all_reduce(src, op=self.dist.ReduceOp.MAX)
        return src.to(t.device)

    # ---- the three collectives of the device-built slab step, in place on preallocated tensors (one NCCL call each; under
    # gloo with CUDA tensors they are staged through the host, which is what the two-ranks-on-one-GPU tests use)
    def all_gather_into(self, out, t):
        """out [world * k] <- every rank's t [k]"""
        if self.world == 1:
            out.copy_(t)
        elif self.stage:
            out.copy_(self.all_gather(t).reshape(-1))
        else:
            self.dist.all_gather_into_tensor(out, t)

    def all_to_all_into(self, out, t):
        """equal splits: block p of t goes to rank p, block p of out comes from rank p"""
        if self.world == 1:
            return
        if self.stage or self.dist.get_backend() == "gloo":
            k = t.numel() // self.world
            out.copy_(self.exchange(t.view(self.world, k), [1] * self.world, [1] * self.world).reshape(-1))
        else:
            self.dist.all_to_all_single(out, t)

    def all_max_(self, t):
        """elementwise maximum over the ranks, in place"""
        if self.world == 1:
            return
        if self.stage:
            t.copy_(self.all_max(t))
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)

    def exchange(self, send, send_counts, recv_counts):
        """variable all-to-all of rows: send [sum(send_counts), k] grouped by destination -> [sum(recv_counts), k]"""
        k = send.shape[1:]
        dev = send.device
        if self.stage:
            send = send.cpu()
        recv = torch.empty((int(sum(recv_counts)),) + tuple(k), dtype=send.dtype, device=send.device)
        if self.world == 1:
            return recv
        ops, so, ro = [], 0, 0
        for p in range(self.world):
            sc, rc = int(send_counts[p]), int(recv_counts[p])
            if p != self.rank:
                if sc:
                    ops.append(self.dist.P2POp(self.dist.isend, send[so:so + sc].contiguous(), p))
                if rc:
                    ops.append(self.dist.P2POp(self.dist.irecv, recv[ro:ro + rc], p))
            so += sc
            ro += rc
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        return recv.to(dev)


# ------------------------------------------------------------------------------------------------ per-rank state
@dataclass
class SlabState:
    """this rank's floes (a contiguous range of global ids) as torch tensors on the compute device"""
    id0: int
    n_global: int
    x: torch.Tensor
    y: torch.Tensor
    rmax: torch.Tensor
    h: torch.Tensor
    area: torch.Tensor
    u: torch.Tensor
    v: torch.Tensor
    ksi: torch.Tensor
    alive: torch.Tensor      # uint8
    voff: torch.Tensor       # int64 [n+1]
    vx: torch.Tensor
    vy: torch.Tensor
    ext: tuple = None        # (minvx, maxvx, minvy, maxvy) per floe: static while the outlines do not change

    @staticmethod
    def from_soa(soa, id0, n_global, device):
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(device)
        s = SlabState(id0, n_global, *(t(getattr(soa, k), F64) for k in abi.FloesSoA.FIELDS), t(soa.alive, torch.uint8), t(soa.voff, I64), t(soa.vx, F64), t(soa.vy, F64))
        s.update_outline_extents()
        return s

    @property
    def n(self):
        return self.x.shape[0]

    def update_outline_extents(self):
        n = self.n
        nv = self.voff[1:] - self.voff[:-1]
        seg = torch.repeat_interleave(torch.arange(n, device=self.x.device), nv)
        big = torch.full((n,), float("inf"), dtype=F64, device=self.x.device)
        mn = lambda v: big.clone().scatter_reduce(0, seg, v, "amin", include_self=True)
        mx = lambda v: (-big).scatter_reduce(0, seg, v, "amax", include_self=True)
        self.ext = (mn(self.vx), mx(self.vx), mn(self.vy), mx(self.vy))


@dataclass
class LocalList:
    """this rank's part of the global extended floe list, ascending gid (what sz_upload_extended takes)"""
    gid: torch.Tensor        # int64, 0-based global position
    floe_num: torch.Tensor   # int64, FloeNums
    x: torch.Tensor
    y: torch.Tensor
    root_x: torch.Tensor
    root_y: torch.Tensor
    body: torch.Tensor       # [n, 6] rmax h area u v ksi
    alive: torch.Tensor      # uint8
    owned: torch.Tensor      # uint8
    parent: torch.Tensor     # int64, 1-based local index of an owned image's parent, 0 otherwise
    voff: torch.Tensor       # int64 [n+1]
    vx: torch.Tensor
    vy: torch.Tensor
    n_ext_global: int
    halo_sent: int
    halo_bytes: int
    plan: dict = None        # bookkeeping of the build (image parents, halo selection), kept for inspection


def _sgn(t):
    return (t > 0).to(F64) - (t < 0).to(F64)


def build_local_list(st, Lx, Ly, periodic, reach, comm, skin=0.0):
    """The rank's part of the global extended floe list in torch (CPU or CUDA tensors): the REFERENCE IMPLEMENTATION of what
    the library's sz_slab_prepare / sz_slab_pack / sz_slab_build kernels do on the device.  It runs under gloo on CPU tensors
    (tests/test_slabs_cpu.py holds it to the oracle's global list and candidate pairs) and the GPU worker compares the
    device-built list with it entry by entry.  `reach` = 2 * max(rmax) over all ranks; `skin` optionally widens the halo."""
    reach = reach + skin
    dev = st.x.device
    n = st.n
    ar = torch.arange(n, device=dev)
    alive = st.alive != 0
    minvx, maxvx, minvy, maxvy = st.ext
    if periodic:
        # x pass over the originals (:28-39): max_v |c_alpha(1,v) + Xi| > Lx; fl(v + X) is monotone in v, so the maximum
        # over the vertices is attained at an extreme of the outline
        fx = alive & (torch.maximum((maxvx + st.x).abs(), (minvx + st.x).abs()) > Lx)
        fy = None
        xg_par = fx.nonzero().squeeze(1)
        xg_x, xg_y = st.x[xg_par] - 2 * Lx * _sgn(st.x[xg_par]), st.y[xg_par]
        # y pass over originals + x-ghosts (:49-60)
        src_all = torch.cat([ar, xg_par])
        x_all, y_all = torch.cat([st.x, xg_x]), torch.cat([st.y, xg_y])
        fy = alive[src_all] & (torch.maximum((maxvy[src_all] + y_all).abs(), (minvy[src_all] + y_all).abs()) > Ly)
        yg_par = fy.nonzero().squeeze(1)
        yg_x, yg_y = x_all[yg_par], y_all[yg_par] - 2 * Ly * _sgn(y_all[yg_par])
    else:
        xg_par = yg_par = torch.zeros(0, dtype=I64, device=dev)
        fx = fy = None
        xg_x = xg_y = yg_x = yg_y = torch.zeros(0, dtype=F64, device=dev)
        src_all = ar
    cx, cy = xg_par.shape[0], yg_par.shape[0]
    cyo = int((yg_par < n).sum()) if cy else 0
    # x-extents of this rank's entries: A = originals (and their y-images), B = x-images (and theirs)
    inf = float("inf")
    fin = st.x[~torch.isnan(st.x)]
    a_lo, a_hi = (float(fin.min()), float(fin.max())) if fin.numel() else (inf, -inf)
    b_lo, b_hi = (float(xg_x.min()), float(xg_x.max())) if cx else (inf, -inf)
    meta = comm.all_gather(torch.tensor([cx, cyo, cy - cyo, a_lo, a_hi, b_lo, b_hi], dtype=F64, device=dev)).cpu()
    cnt = meta[:, :3].to(I64)
    r = comm.rank
    n0 = st.n_global
    n1 = n0 + int(cnt[:, 0].sum())
    base_yo, base_yx = n1, n1 + int(cnt[:, 1].sum())
    n_ext = base_yx + int(cnt[:, 2].sum())
    gid = torch.cat([st.id0 + ar,
                     n0 + int(cnt[:r, 0].sum()) + torch.arange(cx, device=dev),
                     base_yo + int(cnt[:r, 1].sum()) + torch.arange(cyo, device=dev),
                     base_yx + int(cnt[:r, 2].sum()) + torch.arange(cy - cyo, device=dev)])
    src = torch.cat([ar, xg_par, src_all[yg_par]])                       # local original each own entry images
    ox, oy = torch.cat([st.x, xg_x, yg_x]), torch.cat([st.y, xg_y, yg_y])
    n_own = n + cx + cy
    fnum = torch.cat([st.id0 + ar + 1, -(st.id0 + src[n:] + 1)])
    par_own = torch.cat([torch.full((n,), -1, dtype=I64, device=dev), xg_par, yg_par])   # index into the own list

    # ---- halo selection and exchange
    send_idx, send_counts = [], [0] * comm.world
    for p in range(comm.world):
        if p == r:
            continue
        alo, ahi, blo, bhi = (float(v) for v in meta[p, 3:7])
        m = ((ox >= alo - reach) & (ox <= ahi + reach)) | ((ox >= blo - reach) & (ox <= bhi + reach))
        idx = m.nonzero().squeeze(1)
        send_idx.append(idx)
        send_counts[p] = idx.shape[0]
    sidx = torch.cat(send_idx) if send_idx else torch.zeros(0, dtype=I64, device=dev)
    ssrc = src[sidx]
    nv_own = st.voff[1:] - st.voff[:-1]
    s_nv = nv_own[ssrc]
    state = torch.stack([gid[sidx].to(F64), fnum[sidx].to(F64), ox[sidx], oy[sidx], st.x[ssrc], st.y[ssrc], st.rmax[ssrc], st.h[ssrc], st.area[ssrc],
                         st.u[ssrc], st.v[ssrc], st.ksi[ssrc], st.alive[ssrc].to(F64), s_nv.to(F64)], 1) if sidx.numel() else torch.zeros((0, N_STATE), dtype=F64, device=dev)
    s_voff = torch.zeros(sidx.shape[0] + 1, dtype=I64, device=dev)
    torch.cumsum(s_nv, 0, out=s_voff[1:])
    vidx = torch.repeat_interleave(st.voff[:-1][ssrc] - s_voff[:-1], s_nv) + torch.arange(int(s_voff[-1]), device=dev)
    sverts = torch.stack([st.vx[vidx], st.vy[vidx]], 1)
    vert_counts, o = [0] * comm.world, 0
    for p in range(comm.world):
        vert_counts[p] = int(s_voff[o + send_counts[p]] - s_voff[o])
        o += send_counts[p]
    theirs = comm.all_gather(torch.tensor(send_counts + vert_counts, dtype=I64, device=dev)).cpu()
    recv_counts = [int(theirs[p, r]) for p in range(comm.world)]
    recv_vcounts = [int(theirs[p, comm.world + r]) for p in range(comm.world)]
    rstate = comm.exchange(state, send_counts, recv_counts)
    rverts = comm.exchange(sverts, vert_counts, recv_vcounts)

    # ---- merge own + halo, ascending gid
    nh = rstate.shape[0]
    all_gid = torch.cat([gid, rstate[:, 0].to(I64)])
    order = torch.argsort(all_gid)
    pos = torch.empty_like(order)
    pos[order] = torch.arange(order.shape[0], device=dev)                # position of every entry in the sorted list
    parent = torch.zeros(n_own + nh, dtype=I64, device=dev)
    has_par = par_own >= 0
    parent[:n_own][has_par] = pos[par_own[has_par]] + 1
    own_body = torch.stack([st.rmax[src], st.h[src], st.area[src], st.u[src], st.v[src], st.ksi[src]], 1)
    body = torch.cat([own_body, rstate[:, 6:12]])[order]
    nv_all = torch.cat([nv_own[src], rstate[:, 13].to(I64)])
    start_all = torch.cat([st.voff[:-1][src], st.vx.shape[0] + (torch.cumsum(rstate[:, 13].to(I64), 0) - rstate[:, 13].to(I64))])
    nv_s, start_s = nv_all[order], start_all[order]
    voff = torch.zeros(order.shape[0] + 1, dtype=I64, device=dev)
    torch.cumsum(nv_s, 0, out=voff[1:])
    gidx = torch.repeat_interleave(start_s - voff[:-1], nv_s) + torch.arange(int(voff[-1]), device=dev)
    pool_x, pool_y = torch.cat([st.vx, rverts[:, 0]]), torch.cat([st.vy, rverts[:, 1]])
    return LocalList(
        gid=all_gid[order], floe_num=torch.cat([fnum, rstate[:, 1].to(I64)])[order],
        x=torch.cat([ox, rstate[:, 2]])[order], y=torch.cat([oy, rstate[:, 3]])[order],
        root_x=torch.cat([st.x[src], rstate[:, 4]])[order], root_y=torch.cat([st.y[src], rstate[:, 5]])[order],
        body=body, alive=torch.cat([st.alive[src], rstate[:, 12].to(torch.uint8)])[order],
        owned=torch.cat([torch.ones(n_own, dtype=torch.uint8, device=dev), torch.zeros(nh, dtype=torch.uint8, device=dev)])[order],
        parent=parent[order], voff=voff, vx=pool_x[gidx], vy=pool_y[gidx], n_ext_global=n_ext,
        halo_sent=int(sidx.shape[0]), halo_bytes=int(state.numel() * 8 + sverts.numel() * 8),
        plan=dict(xg_par=xg_par, yg_par=yg_par, src=src, src_all=src_all, fx=fx, fy=fy, sidx=sidx, ssrc=ssrc, send_counts=send_counts, recv_counts=recv_counts,
                  order=order, n_own=n_own, x0=st.x.clone(), y0=st.y.clone(), skin=skin))


def fix_kill_transfer(gid, floe_num, owned, kill_i, transfer_i, id0, n_own, comm):
    """floe_interactions_all.m:175-179 across ranks:  for i = 1:length(kill): if kill(i) ~= i && kill(i) > 0,
    transfer(kill(i)) = i  (serial, so the largest i wins).  Inputs are per local entry; returns (kill, transfer)
    for this rank's original floes."""
    dev = gid.device
    mine = (owned != 0) & (floe_num > 0)
    kill = kill_i[mine].clone()
    transfer = transfer_i[mine].clone()
    ev = (owned != 0) & (kill_i > 0) & (kill_i != gid + 1)
    rec = torch.stack([gid[ev] + 1, kill_i[ev]], 1).to(I64)
    cnts = comm.all_gather(torch.tensor([rec.shape[0]], dtype=I64, device=dev)).cpu().flatten()
    mx = int(cnts.max())
    if mx == 0:
        return kill, transfer
    pad = torch.zeros((mx, 2), dtype=I64, device=dev)
    pad[:rec.shape[0]] = rec
    allrec = comm.all_gather(pad.flatten()).reshape(comm.world, mx, 2)
    recs = torch.cat([allrec[p, :int(cnts[p])] for p in range(comm.world)])
    tgt = recs[:, 1] - 1 - id0
    ok = (tgt >= 0) & (tgt < n_own)
    if ok.any():
        win = torch.zeros(n_own, dtype=I64, device=dev).scatter_reduce(0, tgt[ok], recs[ok, 0], "amax", include_self=True)
        transfer = torch.where(win > 0, win.to(transfer.dtype), transfer)
    return kill, transfer


# ------------------------------------------------------------------------------------------------ device-built slab step
class DeviceSlab:
    """One rank's part of a multi-GPU run with the list surgery on the device (sz_slab_* of the C ABI): the rank OWNS the floes
    `gid` (ascending global floe numbers, 1-based -- the numbering of the single-GPU run; any subset), their state and the
    integrator's stay resident in the library, and every step the rank's part of the global extended floe list is rebuilt from
    the CURRENT state -- positions, rotated outlines, thickness, alive flags -- so the step is exact for a moving, rotating,
    thinning field (calc_trajectory.m:75-79,170-222).  Per step: two library kernels groups around two collectives --

        sz_slab_prepare -> all-gather of the meta records (image counts and numbers, x-extents, largest rmax)
        sz_slab_pack    -> all-to-all of fixed-size blocks: every own entry (state + CURRENT outline) another rank can reach
        sz_slab_build   -> received entries merged by global position; sz_step_resident; (sz_trajectory_step)

    -- all on torch's current stream, with no host synchronisation before the step's own final counter read.  Pairs that
    straddle two slabs are resolved on both sides (bit-identical rows for every floe, no return exchange of partial forces);
    forces on an owned periodic image are folded into its parent by the owner (floe_interactions_all.m:242-245).  Capacities
    of the fixed-size blocks are agreed at plan time from measured counts (x 1.5); a step that outgrows them is detected on
    the device (one all-reduced status word) and repeated after a new plan."""

    def __init__(self, prm, owned, gid, n_global, comm, ctx, bnd=None, slack=1.5):
        self.prm, self.comm, self.ctx, self.bnd, self.slack = prm, comm, ctx, bnd, slack
        self.n_global = int(n_global)
        self.plans = self.steps = 0
        self.summary = None
        import os
        self.timing = {"ms": [0.0] * 7, "n": 0} if os.environ.get("SZ_SLAB_TIMING") else None
        self.dev = torch.device(comm.device) if comm.device is not None else torch.device("cuda", torch.cuda.current_device())
        # One stream for everything of a step: the library launches there and torch's collectives are issued there, so kernels
        # and collectives are ordered by the stream alone.  Under gloo (host-staged tests) it is torch's current stream.
        self.stream = torch.cuda.Stream(self.dev) if (comm.world > 1 and not comm.stage) else torch.cuda.current_stream(self.dev)
        ctx.set_stream(self.stream.cuda_stream)
        self.graph, self.want_graph, self.graph_launches, self.graph_replays = None, False, 0, 0
        self.upload(owned, gid)

    # ---- state
    def upload(self, owned, gid, plan=True):
        self.owned = owned
        self.gid = np.ascontiguousarray(gid, np.int32)
        assert self.gid.shape[0] == owned.n and (owned.n < 2 or np.all(np.diff(self.gid) > 0)), "global floe numbers must ascend"
        fs = owned.struct()
        bs = self.bnd.struct() if self.bnd is not None else None
        if plan:
            self.graph = None        # (plan=False: same floes, same capacities, same buffers -- only their contents change, a captured step stays valid)
        self.stream.synchronize()
        abi.check(abi.lib().sz_slab_upload(self.ctx._h, C.byref(self.prm), C.byref(fs), C.byref(bs) if bs is not None else None, abi._ptr(self.gid, abi.c_ip),
                                           self.n_global, self.comm.rank, self.comm.world))
        self.ctx._n0 = owned.n
        if plan:
            self.plan()
        else:               # same capacities and buffers as before (a step that outgrows them is detected and re-planned)
            abi.check(abi.lib().sz_slab_configure(self.ctx._h, self.cap_img, self.cap_rec, self.cap_vert))

    def plan(self, grow=1.0):
        """measure what the current state needs, agree on the capacities across ranks, allocate the exchange buffers"""
        lib, W = abi.lib(), self.comm.world
        self.graph = None
        local8 = np.zeros(8)
        abi.check(lib.sz_slab_measure(self.ctx._h, abi._ptr(local8, abi.c_dp)))
        all8 = self.comm.all_gather(torch.from_numpy(local8).to(self.dev)).cpu().numpy().reshape(W, 8).copy()
        rec, vert = np.zeros(W, np.int64), np.zeros(W, np.int64)
        abi.check(lib.sz_slab_measure_halo(self.ctx._h, abi._ptr(all8, abi.c_dp), abi._ptr(rec, abi.c_lp), abi._ptr(vert, abi.c_lp)))
        need = torch.tensor([float(all8[:, :3].max()), float(rec.max()), float(vert.max())], dtype=F64, device=self.dev)
        need = self.comm.all_max(need).cpu().numpy()
        k = self.slack * grow
        self.cap_img, self.cap_rec, self.cap_vert = (int(k * need[0]) + 64, int(k * need[1]) + 64, int(k * need[2]) + 1024)
        abi.check(lib.sz_slab_configure(self.ctx._h, self.cap_img, self.cap_rec, self.cap_vert))
        self.meta_n = int(lib.sz_slab_meta_doubles(self.cap_img))
        self.block = int(lib.sz_slab_block_doubles(self.cap_rec, self.cap_vert))
        z = lambda n: torch.zeros(n, dtype=F64, device=self.dev)
        self.meta, self.all_meta = z(self.meta_n), z(W * self.meta_n)
        self.send, self.recv = z(W * self.block), z(W * self.block)
        self.status = torch.zeros(4, dtype=torch.int32, device=self.dev)      # [overflow, outside | list length, -]: the first two are all-reduced
        self.flag = self.status[:2]
        P = lambda x, typ: C.cast(C.c_void_p(x.data_ptr()), typ)
        self._p_meta, self._p_all_meta, self._p_send, self._p_recv, self._p_status = (P(self.meta, abi.c_dp), P(self.all_meta, abi.c_dp), P(self.send, abi.c_dp),
                                                                                      P(self.recv, abi.c_dp), P(self.status, abi.c_ip))
        self.halo_measured = (int(rec.sum()), int(vert.sum()))
        torch.cuda.synchronize(self.dev)          # the buffers were made on torch's current stream, the step uses self.stream
        self.plans += 1

    # ---- one step
    def exchange(self):
        lib, h, comm = abi.lib(), self.ctx._h, self.comm
        ev = self._stage_events() if self.timing is not None else None
        abi.check(lib.sz_slab_prepare(h, self._p_meta))
        if ev: ev[1].record()
        comm.all_gather_into(self.all_meta, self.meta)
        if ev: ev[2].record()
        abi.check(lib.sz_slab_pack(h, self._p_all_meta, self._p_send))
        if ev: ev[3].record()
        comm.all_to_all_into(self.recv, self.send)
        if ev: ev[4].record()
        abi.check(lib.sz_slab_build(h, self._p_recv, self._p_status))
        if ev: ev[5].record()
        comm.all_max_(self.flag)       # agreed by all ranks: [capacity overflow on some rank, owned floes that left their rank's extent]
        if ev: ev[6].record()

    STAGES = ("prepare", "all_gather", "pack", "all_to_all", "build", "all_reduce", "contact_step")

    def _stage_events(self):
        """SZ_SLAB_TIMING=1: CUDA events between the stages of a step (diagnostic; read back by stage_ms())"""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        ev[0].record()
        self._ev = ev
        return ev

    def stage_ms(self):
        """mean device ms per stage over the steps timed so far (SZ_SLAB_TIMING=1), or None"""
        if not self.timing or not self.timing["n"]:
            return None
        return {k: self.timing["ms"][i] / self.timing["n"] for i, k in enumerate(self.STAGES)}

    def enable_graph(self, on=True):
        """Capture a whole step -- the library's kernels AND the NCCL collectives between them -- into one CUDA graph and replay
        it every step: one launch from the host instead of ~60, so the GPU is never waiting for the host to enqueue the next
        kernel or collective (at 8 GPUs / 125k floes per rank that wait was a quarter of the step).  Needs NCCL (not the
        host-staged gloo path).  Everything the graph bakes in is a capacity, never a count: list, pair and row capacities,
        the cell grid, the exchange blocks; counts are read from device memory.  A step that outgrows a capacity is detected
        as usual, repeated on the ordinary path, and the graph is captured again with the new sizes."""
        self.want_graph = bool(on) and self.comm.world > 1 and not self.comm.stage
        if not self.want_graph:
            self.graph = None
        self.ctx.set_option("graph_safe", 1 if self.want_graph else 0)

    def _capture(self):
        lib, h = abi.lib(), self.ctx._h
        g = torch.cuda.CUDAGraph()
        l0 = lib.sz_launch_count()
        with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
            self.exchange()
            abi.check(lib.sz_step_enqueue(h))
        self.graph_launches = int(lib.sz_launch_count() - l0)
        self.graph = g

    def run(self, allow_pair_errors=False):
        """one contact step on the current state; returns the step's SzSummary (n = the padded list length)"""
        with torch.cuda.stream(self.stream):
            if self.want_graph and self.graph is None and self.steps >= 2:
                try:
                    self._capture()       # (the sizes of the step before are carried over: capture from the third step on)
                except abi.SzError:
                    self.graph = None     # no sizes to carry over yet (a re-plan changed the list capacity): ordinary step first
            if self.graph is not None:
                self.graph.replay()
                self.graph_replays += 1
                abi.lib().sz_add_launches(self.graph_launches)
                s = self.ctx.step_finish(allow_pair_errors=allow_pair_errors)
                flag = self.flag.cpu()
                if s is not None and int(flag[0]) == 0:
                    self.summary, self.steps, self.n_outside = s, self.steps + 1, int(flag[1])
                    return s
                self.graph = None         # a capacity was outgrown: the ordinary path below repeats the step and re-plans
            for attempt in range(4):
                self.exchange()
                s = self.ctx.step_resident(allow_pair_errors=allow_pair_errors)
                if self.timing is not None:
                    self._ev[7].record()
                    self._ev[7].synchronize()
                    self.timing["seen"] = self.timing.get("seen", 0) + 1
                    if self.timing["seen"] > 3:         # the first steps carry NCCL's connection set-up and the planning step
                        for i in range(7):
                            self.timing["ms"][i] += self._ev[i].elapsed_time(self._ev[i + 1])
                        self.timing["n"] += 1
                flag = self.flag.cpu()
                if int(flag[0]) == 0:
                    self.summary, self.steps, self.n_outside = s, self.steps + 1, int(flag[1])
                    return s
                self.plan(grow=2.0 ** (attempt + 1))           # the field outgrew the blocks: new capacities, repeat the step
        raise RuntimeError("slab step: capacities kept overflowing")

    def trajectory_init(self, mass, inertia, nz=1000, **fields):
        self.ctx.trajectory_init(mass, inertia, nz=nz, **fields)
        self._nz = int(nz)
        self._c0 = (np.array(fields["c0x"], np.float64), np.array(fields["c0y"], np.float64)) if "c0x" in fields else (self.owned.vx.copy(), self.owned.vy.copy())

    def set_extent(self, xlo, xhi, tol=0.0):
        """the x-range this rank's floes belong to (its slab, widened by `tol`): after a step `n_outside` says how many live
        owned floes have left it on some rank -- the cue for repartition()"""
        self.extent, self.extent_tol = (float(xlo), float(xhi)), float(tol)
        abi.check(abi.lib().sz_slab_set_extent(self.ctx._h, float(xlo) - tol, float(xhi) + tol))

    def repartition(self, edges):
        """Floe migration between slabs: every floe goes to the rank whose slab [edges[r-1], edges[r]) holds its CURRENT centroid
        (interior edges [world - 1] as from slab_edges()); floes keep their global numbers, so results stay those of the
        single-GPU run.  A moving floe takes everything along: state, rotated outline and c0, the integrator's previous-step
        values and forcing, its stress history.  Host-mediated (rare: a floe has to cross a slab first); needs
        trajectory_init.  Returns the number of floes this rank gave away."""
        comm, W, r = self.comm, self.comm.world, self.comm.rank
        n, nv = self.owned.n, self.owned.vx.shape[0]
        st = self.ctx.trajectory_state(nverts=nv)
        fo = self.ctx.trajectory_forcing()
        sh, sc = self.ctx.stress_history()
        dest = slab_of(st["x"], self.prm.Lx, W, edges)
        dest = np.where(np.isnan(st["x"]), r, dest)
        voff = self.owned.voff.astype(np.int64)
        per_floe = {k: st[k] for k in ("x", "y", "u", "v", "ksi", "h", "alive", "mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p")}
        per_floe.update(rmax=self.owned.rmax, area=self.owned.area, FxOA=fo["FxOA"], FyOA=fo["FyOA"], torqueOA=fo["torqueOA"], gid=self.gid, stress_count=sc, nv=np.diff(voff))
        per_floe["stress_h"] = sh.reshape(n, -1)

        def take(idx):
            d = {k: np.ascontiguousarray(v[idx]) for k, v in per_floe.items()}
            vi = np.repeat(voff[idx] - np.concatenate([[0], np.cumsum(per_floe["nv"][idx])[:-1]]), per_floe["nv"][idx]) + np.arange(int(per_floe["nv"][idx].sum())) if len(idx) else np.zeros(0, np.int64)
            d.update(cax=st["cax"][vi], cay=st["cay"][vi], c0x=self._c0[0][vi], c0y=self._c0[1][vi])
            return d
        out = {p: take(np.nonzero(dest == p)[0]) for p in range(W) if p != r and (dest == p).any()}
        gathered = [None] * W
        if W > 1:
            comm.dist.all_gather_object(gathered, out)
        parts = [take(np.nonzero(dest == r)[0])] + [g[r] for p, g in enumerate(gathered) if p != r and g and r in g]
        cat = lambda k: np.concatenate([p_[k] for p_ in parts])
        order = np.argsort(cat("gid"), kind="stable")
        nvs = cat("nv")
        starts = np.concatenate([[0], np.cumsum(nvs)[:-1]]).astype(np.int64)
        vi = np.repeat(starts[order] - np.concatenate([[0], np.cumsum(nvs[order])[:-1]]), nvs[order]) + np.arange(int(nvs.sum())) if len(order) else np.zeros(0, np.int64)
        g = lambda k: cat(k)[order]
        new_voff = np.concatenate([[0], np.cumsum(nvs[order])]).astype(np.int32)
        owned = abi.FloesSoA(g("x"), g("y"), g("rmax"), g("h"), g("area"), g("u"), g("v"), g("ksi"), g("alive").astype(np.uint8), new_voff, cat("cax")[vi], cat("cay")[vi])
        gave = int((dest != r).sum())
        self.upload(owned, g("gid").astype(np.int32), plan=True)
        nz = self._nz
        self.trajectory_init(g("mass"), g("inertia"), nz=nz, c0x=cat("c0x")[vi], c0y=cat("c0y")[vi], stress_h=g("stress_h").reshape(owned.n, nz, 4), stress_count=g("stress_count"),
                             **{k: g(k) for k in ("alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "FxOA", "FyOA", "torqueOA")})
        lo = -np.inf if r == 0 else float(edges[r - 1])
        hi = np.inf if r == W - 1 else float(edges[r])
        if hasattr(self, "extent_tol"):
            self.set_extent(lo, hi, self.extent_tol)
        return gave

    def trajectory_step(self, dt, HFo=0.0, *bounds):
        return self.ctx.trajectory_step(dt, HFo, *bounds)

    # ---- results of the owned floes
    def outputs(self, into=None):
        n = self.owned.n
        o = into if into is not None else {"fx": np.empty(n), "fy": np.empty(n), "torque": np.empty(n), "overlap_area": np.empty(n), "stress": np.empty((n, 2, 2)), "xi": np.empty(n), "yi": np.empty(n),
             "alive": np.empty(n, np.uint8), "kill": np.empty(n, np.int32), "transfer": np.empty(n, np.int32)}
        p = abi._ptr
        abi.check(abi.lib().sz_slab_get_outputs(self.ctx._h, p(o["fx"], abi.c_dp), p(o["fy"], abi.c_dp), p(o["torque"], abi.c_dp), p(o["overlap_area"], abi.c_dp), p(o["stress"], abi.c_dp),
                                                p(o["xi"], abi.c_dp), p(o["yi"], abi.c_dp), p(o["alive"], abi.c_bp), p(o["kill"], abi.c_ip), p(o["transfer"], abi.c_ip)))
        if self.summary is not None and self._any_kill():
            o["kill"], o["transfer"] = self._fix_kill_transfer(o["kill"], o["transfer"])
        return o

    def positions(self):
        pos, nl = np.empty(self.owned.n, np.int32), C.c_int32()
        abi.check(abi.lib().sz_slab_get_positions(self.ctx._h, abi._ptr(pos, abi.c_ip), C.byref(nl)))
        return pos, nl.value

    def local_list(self):
        _, nl = self.positions()
        o = {"gid": np.empty(nl, np.int32), "floe_num": np.empty(nl, np.int32), "owned": np.empty(nl, np.uint8), "x": np.empty(nl), "y": np.empty(nl)}
        p = abi._ptr
        abi.check(abi.lib().sz_slab_get_list(self.ctx._h, p(o["gid"], abi.c_ip), p(o["floe_num"], abi.c_ip), p(o["owned"], abi.c_bp), p(o["x"], abi.c_dp), p(o["y"], abi.c_dp)))
        return o

    def rows(self, pinned=None):
        """(row_off [n_owned + 1], rows [K, 7]) of the owned floes in their order (gathered on the device); partner ids are global
        list positions.  pinned: optional dict caching page-locked receive buffers between calls."""
        n = self.owned.n
        if pinned is not None:
            if pinned.get("n") != n:
                pinned.update(n=n, off=torch.empty(n + 1, dtype=I64).pin_memory().numpy(), rows=None)
            cap = 0 if pinned["rows"] is None else pinned["rows"].shape[0]
            need = int(self.summary.n_rows) if self.summary is not None else 0
            if cap < need:
                pinned["rows"] = torch.empty((int(need * 1.1) + 16, 7), dtype=F64).pin_memory().numpy()
            off, buf = pinned["off"], pinned["rows"]
        else:
            off, buf = np.empty(n + 1, np.int64), np.empty((int(self.summary.n_rows) + 1, 7))
        nr = C.c_int64()
        abi.check(abi.lib().sz_slab_get_rows(self.ctx._h, abi._ptr(off, abi.c_lp), abi._ptr(buf, abi.c_dp), buf.shape[0], C.byref(nr)))
        return off, buf[:nr.value]

    def _any_kill(self):
        t = torch.tensor([float(self.summary.n_kill_events)], dtype=F64, device=self.dev)
        return float(self.comm.all_max(t)) > 0 if self.comm.world > 1 else self.summary.n_kill_events > 0

    def _fix_kill_transfer(self, kill, transfer):
        """floe_interactions_all.m:175-179 across ranks (rare: only when some floe merged): for i = 1:length(kill), if kill(i) ~= i
        && kill(i) > 0, transfer(kill(i)) = i -- serial, so the largest i wins; i runs over the whole extended list (an image's
        entry counts)."""
        L = self.local_list()
        n_list = L["gid"].shape[0]
        ko, to = np.empty(self.summary.n, np.int32), np.empty(self.summary.n, np.int32)
        abi.check(abi.lib().sz_get_floe_outputs(self.ctx._h, None, None, None, None, None, None, None, None, abi._ptr(ko, abi.c_ip), abi._ptr(to, abi.c_ip)))
        ko = ko[:n_list]
        ev = (L["owned"] != 0) & (ko > 0) & (ko != L["gid"])
        rec = np.stack([L["gid"][ev], ko[ev]], 1).astype(np.int64) if ev.any() else np.zeros((0, 2), np.int64)
        cnts = self.comm.all_gather(torch.tensor([rec.shape[0]], dtype=I64, device=self.dev)).cpu().flatten()
        mx = int(cnts.max())
        if mx == 0:
            return kill, transfer
        pad = torch.zeros((mx, 2), dtype=I64, device=self.dev)
        pad[:rec.shape[0]] = torch.from_numpy(rec).to(self.dev)
        allrec = self.comm.all_gather(pad.flatten()).reshape(self.comm.world, mx, 2).cpu().numpy()
        recs = np.concatenate([allrec[p, :int(cnts[p])] for p in range(self.comm.world)])
        transfer = transfer.copy()
        k = np.searchsorted(self.gid, recs[:, 1])
        ok = (k < self.gid.shape[0]) & (self.gid[np.minimum(k, self.gid.shape[0] - 1)] == recs[:, 1])
        win = np.zeros(self.gid.shape[0], np.int64)
        np.maximum.at(win, k[ok], recs[ok, 0])
        transfer = np.where(win > 0, win.astype(np.int32), transfer)
        return kill, transfer


class SlabJob:
    """bench.py's workload: the synthetic periodic Voronoi field (BASELINE.json configs[4]) on `world` GPUs.  The floes are
    numbered slab by slab (a stable sort of the generator's numbering by x-slab; at world 1 the generator's own), rank r owns
    the r-th slab: floe ownership by centroid."""

    def __init__(self, n_floes, seed, rank, world, local_rank, dist, order="site", number_for_world=None, graph=True):
        self.rank, self.world, self.dist = rank, world, dist
        self.prm, field = voronoi_field(n_floes, seed=seed, order=order)
        self.ctx = ContactContext(local_rank)
        self.summary = None
        self.slab = None
        nw = number_for_world or world
        if nw > 1:
            field, starts = sort_by_slab(field, self.prm.Lx, nw)
        else:
            starts = np.array([0, field.n], np.int64)
        self.field, self.starts = field, starts
        if world == 1:
            self.floes = field
            self.gid = np.arange(1, field.n + 1, dtype=np.int32)
            self.ctx.upload(self.prm, self.floes)
        else:
            a, b = int(starts[rank]), int(starts[rank + 1])
            self.floes = take_range(field, a, b)
            self.gid = np.arange(a + 1, b + 1, dtype=np.int32)
            dev = torch.device("cuda", local_rank)
            self.comm = Comm(dist, rank, world, dev)
            self.slab = DeviceSlab(self.prm, self.floes, self.gid, field.n, self.comm, self.ctx)
            self.slab.enable_graph(graph)
        self._pin()

    def _pin(self):
        """pinned host copies of the step's inputs (e2e leg)"""
        f = self.floes
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        self.pinned = abi.FloesSoA(*(pin(getattr(f, k)) for k in abi.FloesSoA.FIELDS), pin(f.alive), pin(f.voff), pin(f.vx), pin(f.vy))
        self._out = None
        self._rows = None

    def step_resident(self):
        """one step with this rank's floes resident in HBM; returns (device ms, phase ms)"""
        if self.slab is None:
            s = self.ctx.step_resident()
            ms = s.ms_device
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.slab.stream)
            s = self.slab.run()
            e1.record(self.slab.stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)          # meta all-gather + halo all-to-all + list build + local step, on the device timeline
        self.summary = s
        self.pairs_owned, self.pairs_force_total = int(s.n_pairs_owned), int(s.n_pairs_force)
        ph = self.ctx.phase_ms()
        ph["classes"] = self.ctx.narrow_class_ms()     # per size class: (kernel ms, pairs)
        self.rows_owned = int(s.n_rows)
        if self.slab is None:
            self.ext_entries_owned = int(s.n)
        else:
            self.ext_entries_owned = int(self.slab.status[2].item())      # list length incl. halo (the halo share is in describe())
        return ms, ph

    def e2e_step(self):
        """host buffers in (pinned), per-floe outputs and all contact rows out; returns (h2d, d2h) bytes"""
        f = self.pinned
        h2d = sum(getattr(f, k).nbytes for k in abi.FloesSoA.FIELDS) + f.alive.nbytes + f.voff.nbytes + f.vx.nbytes + f.vy.nbytes
        if self.slab is None:
            s = self.ctx.step(self.prm, f)
            n = f.n
            if self._out is None:
                z = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
                self._out = {"fx": z(n, F64), "fy": z(n, F64), "torque": z(n, F64), "overlap_area": z(n, F64), "stress": z((n, 2, 2), F64),
                             "xi": z(n, F64), "yi": z(n, F64), "alive": z(n, torch.uint8), "kill": z(n, torch.int32), "transfer": z(n, torch.int32)}
            if self._rows is None or self._rows[1].shape[0] < s.n_rows or self._rows[0].shape[0] != s.n + 1:
                self._rows = (torch.empty(s.n + 1, dtype=I64).pin_memory().numpy(), torch.empty((int(s.n_rows * 1.1) + 16, 7), dtype=F64).pin_memory().numpy())
            self.ctx.floe_outputs(into=self._out)
            abi.check(abi.lib().sz_get_rows(self.ctx._h, abi._ptr(self._rows[0], abi.c_lp), abi._ptr(self._rows[1], abi.c_dp)))
            d2h = sum(v.nbytes for v in self._out.values()) + (s.n + 1) * 8 + int(s.n_rows) * 56
        else:
            # this rank's floes travel host -> device, the slab step runs, its floes' results travel back
            self.slab.upload(f, self.gid, plan=False)
            s = self.slab.run()
            n = f.n
            if self._out is None:
                z = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
                self._out = {"fx": z(n, F64), "fy": z(n, F64), "torque": z(n, F64), "overlap_area": z(n, F64), "stress": z((n, 2, 2), F64),
                             "xi": z(n, F64), "yi": z(n, F64), "alive": z(n, torch.uint8), "kill": z(n, torch.int32), "transfer": z(n, torch.int32)}
                self._rows = {}
            out = self.slab.outputs(into=self._out)
            row_off, rows = self.slab.rows(pinned=self._rows)
            d2h = sum(v.nbytes for v in out.values()) + row_off.nbytes + rows.nbytes
        self.summary = s
        return h2d, d2h

    def describe(self):
        if self.world == 1:
            return "1 GPU, whole field"
        return ("%d x-slabs, floe ownership by centroid; every step each rank rebuilds its part of the global extended floe list on the device from the current state: "
                "one all-gather of per-rank meta records (periodic-image counts and numbers, x-extents, largest rmax) and one all-to-all over NCCL of fixed-size blocks "
                "carrying every entry within reach of another slab -- state, FloeNums, root centroid and its current (rotated) outline; straddling pairs are resolved on "
                "both sides instead of returning partial forces to the owner (bit-identical rows, no second exchange)") % self.world

    def close(self):
        if self.slab is not None:
            self.slab.graph = None            # a captured graph holds NCCL kernels: release it before the process group goes away
            torch.cuda.synchronize()
        self.ctx.close()
