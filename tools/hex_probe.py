"""GPU probe: a perfect hexagonal lattice (every floe the same hexagon, same start vertex) vs the random Voronoi field:
upper bound of what aligned lanes would buy the narrow phase."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import subzero_b200 as sz
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ny = nx // 2 * 2
a = 1240.0                                  # hexagon circumradius: area 4e6 m^2
w, hgt = np.sqrt(3) * a, 1.5 * a
ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
x = (ix + 0.5 * (iy % 2)) * w; y = iy * hgt
Lx, Ly = nx * w / 2, ny * hgt / 2
x = (x - Lx + w / 4).ravel(); y = (y - Ly + hgt / 2).ravel()
rng = np.random.default_rng(0)
jit = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
x += rng.uniform(-jit, jit, x.shape); y += rng.uniform(-jit, jit, y.shape)
n = x.shape[0]
ang = np.pi / 2 - 2 * np.pi * np.arange(7) / 6          # clockwise, closed
hx, hy = 1.02 * a * np.cos(ang), 1.02 * a * np.sin(ang)
hx[6], hy[6] = hx[0], hy[0]
prm = sz.default_params(Lx=Lx, Ly=Ly, modulus=1.5e3 * 2 * 2000.0, dt=10.0, periodic=1, collision=1)
area = 1.5 * np.sqrt(3) * (1.02 * a) ** 2
soa = sz.FloesSoA(x, y, np.full(n, 1.02 * a), np.full(n, 0.25), np.full(n, area), rng.uniform(-.1, .1, n), rng.uniform(-.1, .1, n), rng.uniform(-1e-5, 1e-5, n),
                  np.ones(n, np.uint8), (7 * np.arange(n + 1)).astype(np.int32), np.tile(hx, n), np.tile(hy, n))
ctx = sz.ContactContext(0); ctx.upload(prm, soa)
for it in range(3):
    s = ctx.step_resident()
print("hex lattice", n, "floes pairs", s.n_pairs, "force", s.n_pairs_force, "narrow ms", round(ctx.phase_ms()["narrow"], 3), "ns/pair %.2f" % (ctx.phase_ms()["narrow"] * 1e6 / s.n_pairs))
