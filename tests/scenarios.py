"""Floe sets used by the parity tests (test infrastructure)."""
import os

import numpy as np

import subzero_b200 as sz

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def polyshape_area_centroid(v):
    """vertex-0-relative shoelace, numpy restatement used only to BUILD inputs"""
    x, y = v[:, 0] - v[0, 0], v[:, 1] - v[0, 1]
    xj, yj = np.roll(x, -1), np.roll(y, -1)
    c = x * yj - xj * y
    a2 = c.sum()
    return abs(a2) / 2, v[0, 0] + ((x + xj) * c).sum() / (3 * a2), v[0, 1] + ((y + yj) * c).sum() / (3 * a2)


def floe_from_polygon(verts, h=0.25, u=0.0, v=0.0, ksi=0.0):
    """Initialize_Model/initialize_floe_values.m:12-52: centroid/area of the polygon, c_alpha about the centroid, CLOSED"""
    verts = np.asarray(verts, float)
    area, cx, cy = polyshape_area_centroid(verts)
    ca = np.vstack([verts - [cx, cy], verts[:1] - [cx, cy]]).T          # 2 x (n+1)
    return {"c_alpha": ca, "Xi": cx, "Yi": cy, "area": area, "h": h, "rmax": float(np.sqrt(((verts - [cx, cy]) ** 2).sum(1).max())),
            "Ui": u, "Vi": v, "ksi_ice": ksi, "alive": 1}


def floe_shapes():
    d = np.load(os.path.join(GOLDEN, "floe_shapes.npz"))
    off, verts = d["off"], d["verts"]
    return [verts[off[i]:off[i + 1]] for i in range(len(off) - 1)], d["boundary_info"], float(d["modulus"])


def domain(Lx, Ly):
    """c2_boundary (2x5 closed, Subzero.m:68) and the boundary floe of Subzero.m:69-70"""
    c2 = np.array([[-Lx, -Lx, Lx, Lx, -Lx], [-Ly, Ly, Ly, -Ly, -Ly]], float)
    hole = np.array([[-Lx, Lx, Lx, -Lx], [-Ly, -Ly, Ly, Ly]], float)   # holes(floebound.poly).Vertices: the domain rectangle
    floebound = {"c": hole, "area": 16 * Lx * Ly - 4 * Lx * Ly, "h": 0.25, "Xi": 0.0, "Yi": 0.0, "Ui": 0.0, "Vi": 0.0, "ksi_ice": 0.0}
    return c2, floebound


def conservation_cases():
    """the five set-ups of the reference's test/conservation_test.m:22-54 (shapes :5-7,16-17,51; height.mean 0.25)"""
    polys, _, modulus = floe_shapes()
    x1, y1 = np.array([2, 2, 5, 5]) * 1e4, np.array([2, 5, 5, 2]) * 1e4
    x2, y2 = np.array([6, 6, 9, 9]) * 1e4, np.array([2, 5, 5, 2]) * 1e4
    x3, y3 = np.array([5.5, 5.25, 5.75]) * 1e4, np.array([2, 4, 4]) * 1e4
    p1, p2, p3 = np.stack([x1, y1], 1), np.stack([x2, y2], 1), np.stack([x3, y3], 1)
    c1, c2 = polys[4], polys[3] - np.array([1e4, 4e4])                 # poly(5), translate(poly(4), -[1e4 4e4])
    F = floe_from_polygon
    cases = {
        "head_on": [F(p1, u=0.15, v=0.02), F(p2, u=-0.1, v=0.02)],
        "offset": [F(p1 + [0, 1e4], u=0.11, v=0.02), F(p2, u=-0.1, v=0.02)],
        "triangle_between": [F(p1, u=0.11, v=0.001), F(p2, u=-0.1, v=0.001), F(p3, u=0.0, v=0.001)],
        "complex_pair": [F(c1, u=-0.11, v=0.02), F(c2, u=0.1, v=0.02)],
        "complex_wall": [F(c1 + [7.75e4, 0], u=0.11, v=0.02)],
    }
    return cases, modulus


def advance(Floe, t):
    """move every floe rigidly for t seconds (so the scenario floes actually touch)"""
    out = []
    for f in Floe:
        g = dict(f)
        g["Xi"], g["Yi"] = f["Xi"] + f["Ui"] * t, f["Yi"] + f["Vi"] * t
        out.append(g)
    return out


def real_shape_field(n_side, seed=0, spacing=0.8, periodic=True, max_vertices=None):
    """tile FloeShapes.mat polygons (7..591 vertices, concave) on a jittered grid so neighbours overlap"""
    polys, _, modulus = floe_shapes()
    rng = np.random.default_rng(seed)
    pitch = 9000.0 * spacing
    L = 0.5 * n_side * pitch
    Floe = []
    for a in range(n_side):
        for b in range(n_side):
            v = polys[int(rng.integers(0, len(polys)))]
            if max_vertices is not None and len(v) > max_vertices:       # crude FloeSimplify: keep every k-th vertex
                v = v[np.linspace(0, len(v), max_vertices, endpoint=False).astype(int)]
            area, cx, cy = polyshape_area_centroid(v)
            s = np.sqrt(5.5e7 / area)                                    # similar footprint, keeps the vertex count
            th = rng.uniform(0, 2 * np.pi)
            R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            w = (v - [cx, cy]) @ R.T * s
            pos = np.array([-L + (a + 0.5) * pitch, -L + (b + 0.5) * pitch]) + rng.uniform(-0.1, 0.1, 2) * pitch
            if len(Floe) % 9 == 8:                                       # a few floes nearly on top of their predecessor: merge (+-Inf) outcomes
                pos = np.array([Floe[-1]["Xi"], Floe[-1]["Yi"]]) + rng.uniform(-0.05, 0.05, 2) * pitch
            Floe.append(floe_from_polygon(w + pos, u=rng.uniform(-0.1, 0.1), v=rng.uniform(-0.1, 0.1), ksi=rng.uniform(-1e-5, 1e-5)))
    prm = sz.default_params()
    prm.Lx = prm.Ly = L
    prm.modulus, prm.dt, prm.periodic, prm.collision = modulus, 10.0, int(periodic), 1
    return prm, Floe


def soa_and_boundary(Floe, prm, periodic):
    soa = sz.floes_to_soa(Floe)
    bnd = None
    if not periodic:
        c2, fb = domain(prm.Lx, prm.Ly)
        bnd = sz.Boundary(fb["c"][0], fb["c"][1], c2[0], c2[1], fb["area"], fb["h"])
    return soa, bnd
