"""Work partition of one contact step over the GPUs of one box (one process per GPU).

world == 1: the whole field on one device.
"""
import numpy as np

from . import abi
from .contact import ContactContext
from .field import voronoi_field


class SlabJob:
    def __init__(self, n_floes, seed, rank, world, local_rank, dist):
        self.rank, self.world, self.dist = rank, world, dist
        if world != 1:
            raise NotImplementedError("multi-GPU slabs: see DESIGN.md (e)")
        self.prm, self.floes = voronoi_field(n_floes, seed=seed)
        self.ctx = ContactContext(local_rank)
        self.ctx.upload(self.prm, self.floes)
        self.summary = None
        self._pin()

    def _pin(self):
        """pinned host copies of the step's inputs and result buffers (e2e leg)"""
        import torch
        f = self.floes
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        self.pinned = abi.FloesSoA(*(pin(getattr(f, k)) for k in abi.FloesSoA.FIELDS), pin(f.alive), pin(f.voff), pin(f.vx), pin(f.vy))
        self._out = None
        self._rows = None

    def step_resident(self):
        s = self.ctx.step_resident()
        self.summary = s
        self.pairs_owned, self.rows_owned, self.pairs_force_total = int(s.n_pairs), int(s.n_rows), int(s.n_pairs_force)
        return s.ms_device, self.ctx.phase_ms()

    def e2e_step(self):
        """host buffers in, per-floe outputs and all contact rows out; returns (h2d, d2h) bytes"""
        import torch
        f = self.pinned
        s = self.ctx.step(self.prm, f)
        n = f.n
        if self._out is None:
            z = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
            self._out = {"fx": z(n, torch.float64), "fy": z(n, torch.float64), "torque": z(n, torch.float64), "overlap_area": z(n, torch.float64),
                         "stress": z((n, 2, 2), torch.float64), "xi": z(n, torch.float64), "yi": z(n, torch.float64), "alive": z(n, torch.uint8),
                         "kill": z(n, torch.int32), "transfer": z(n, torch.int32)}
        if self._rows is None or self._rows[1].shape[0] < s.n_rows or self._rows[0].shape[0] != s.n + 1:
            self._rows = (torch.empty(s.n + 1, dtype=torch.int64).pin_memory().numpy(), torch.empty((int(s.n_rows * 1.1) + 16, 7), dtype=torch.float64).pin_memory().numpy())
        self.ctx.floe_outputs(into=self._out)
        abi.check(abi.lib().sz_get_rows(self.ctx._h, abi._ptr(self._rows[0], abi.c_lp), abi._ptr(self._rows[1], abi.c_dp)))
        h2d = sum(getattr(f, k).nbytes for k in abi.FloesSoA.FIELDS) + f.alive.nbytes + f.voff.nbytes + f.vx.nbytes + f.vy.nbytes
        d2h = sum(v.nbytes for v in self._out.values()) + (s.n + 1) * 8 + int(s.n_rows) * 56
        return h2d, d2h

    def describe(self):
        return "1 GPU, whole field"

    def close(self):
        self.ctx.close()
