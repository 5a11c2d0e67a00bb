"""The C-ABI library loads and exports every symbol include/subzero_b200.h declares; the product path fails
loudly without a GPU / without the extension (no CPU fallback).  CPU only -- no compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import subzero_b200 as sz
from subzero_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "subzero_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sz_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(abi.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    l = C.CDLL(abi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(l, name), name
    assert abi.lib().sz_abi_version() == 1


def test_struct_layouts_match_the_header():
    """sizes implied by the header's field lists (8-byte alignment)"""
    assert C.sizeof(abi.SzParams) == 15 * 8 + 6 * 4          # five int32 fields + padding
    assert C.sizeof(abi.SzFloesSoA) == 8 + 8 + 12 * 8
    assert C.sizeof(abi.SzBoundary) == 8 + 8 + 8 + 8 + 8 + 8 + 7 * 8
    assert C.sizeof(abi.SzSummary) == 8 + 5 * 8 + 8 + 4 + 4 + 8 + 8 + 8          # ... ms_device (+pad), n_pairs_owned, n_kill_events (+pad)
    assert C.sizeof(abi.SzExtendedList) == 6 * 8


def test_default_params_are_the_reference_constants():
    p = sz.default_params()
    # floe_interactions.m:20-21 nu, mu; :55-58 0.55; :37 0.75; :79 100/1.75; :99 1; :127 1e-8; :141 0.1; :15 1e5; :54 0.95
    assert (p.nu, p.mu, p.merge_frac, p.wall_frac, p.amin_per_vertex) == (0.3, 0.2, 0.55, 0.75, 100 / 1.75)
    assert (p.vertex_match_tol, p.on_edge_tol, p.dl_min, p.close_gap, p.big_floe_r, p.domain_area_frac) == (1, 1e-8, 0.1, 1, 1e5, 0.95)
    q = abi.SzParams()
    abi.lib().sz_default_params(C.byref(q))                      # the host restatement and the library agree field by field
    assert bytes(p) == bytes(q)


def test_reference_arm_does_not_map_the_product_library():
    """bench.py --impl reference: the synthetic field comes from libsz_field.so and the parameter block from the host, so the CPU
    arm never loads libsubzero_b200.so"""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import subzero_b200 as sz, oracle\n"
            "prm, f = sz.voronoi_field(300, seed=1)\n"
            "r = oracle.OracleStep(prm, f, nthreads=2, broad_mode=1)\n"
            "maps = open('/proc/self/maps').read()\n"
            "print('PAIRS', r.summary.n_pairs, 'PRODUCT' if 'libsubzero_b200.so' in maps else 'CLEAN', 'FIELD' if 'libsz_field.so' in maps else 'NOFIELD')\n") % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "CLEAN FIELD" in r.stdout and "PAIRS" in r.stdout, r.stdout + r.stderr


def test_bad_arguments_are_rejected_without_a_device():
    assert abi.lib().sz_create(None, 0) == abi.SZ_ERR_ARG
    assert b"NULL" in abi.lib().sz_last_error()
    assert abi.lib().sz_step_resident(None, None) == abi.SZ_ERR_ARG
    # the consumers of the contact rows / resident state validate their context before touching the device too
    assert abi.lib().sz_corner_mask(None, 0, None, 0, None) == abi.SZ_ERR_ARG and b"sz_corner_mask" in abi.lib().sz_last_error()
    assert abi.lib().sz_get_corner_mask(None, None, None) == abi.SZ_ERR_ARG
    assert abi.lib().sz_eulerian_data(None, 4, 4, -1.0, 1.0, -1.0, 1.0, 1, None, None, None, None, None, None, None) == abi.SZ_ERR_ARG
    assert b"sz_eulerian_data" in abi.lib().sz_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sz.SzError) as e:
        sz.ContactContext(0)
    assert e.value.code == abi.SZ_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_missing_extension_fails_loudly(tmp_path):
    code = ("import sys; sys.path.insert(0, %r); from subzero_b200 import abi; abi.LIB_PATH = %r\n"
            "try:\n    abi.lib()\nexcept RuntimeError as e:\n    print('RAISED', e)\n") % (ROOT, str(tmp_path / "nope.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "RAISED" in r.stdout and "no CPU fallback" in r.stdout


def test_product_does_not_reference_the_oracle():
    """nothing under subzero_b200/ or include/ may import, link or name the checker"""
    bad = []
    for base in ("subzero_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "_lib" in dp or "__pycache__" in dp:
                continue
            for f in files:
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"sz_oracle|libsz_oracle|libclipper_ref|szo_|szref_|import oracle|from oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert bad == []
    out = subprocess.run(["ldd", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "clipper_ref" not in out


def test_field_generator_shapes():
    prm, f = sz.voronoi_field(4000, seed=0)
    nv = f.voff[1:] - f.voff[:-1]
    assert f.n == 4000 and nv.min() >= 4 and nv.max() <= 20          # closed outlines: 3..~13 vertices + the repeat
    assert abs(nv.mean() - 7.0) < 0.1                                 # mean 6 vertices per cell
    assert prm.Lx == prm.Ly == 0.5 * (4000 * 4e6) ** 0.5
    for i in (0, 17, 3999):
        x, y = f.outline(i)
        assert x[0] == x[-1] and y[0] == y[-1]                         # closed (initialize_floe_values.m:17)
        area2 = (x[:-1] * y[1:] - x[1:] * y[:-1]).sum()
        assert area2 < 0                                               # clockwise, like FloeShapes.mat
    import numpy as np
    assert abs(f.area.sum() / (4 * prm.Lx * prm.Ly) - 1.02 ** 2) < 1e-9   # cells tile the domain, inflated by 1.02


def test_mex_gateway_source_compiles_against_the_stub_header():
    """MATLAB is not installed: the gateway (subzero_b200/matlab/sz_contact_mex.cpp, modelled on private/mexclipper.cpp)
    is syntax-checked against a stub mex.h so that it cannot rot"""
    for src in ("sz_contact_mex.cpp", "sz_resident_mex.cpp"):          # one-shot contact step; device-resident timestep (integrator, ocean forcing, fracture deformation)
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "host", "mex_stub"), "-I" + os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "subzero_b200", "matlab", src)], capture_output=True, text=True)
        assert r.returncode == 0, src + "\n" + r.stderr


def test_standalone_driver_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = str(tmp_path / "sz_driver")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "sz_driver.cpp"),
                        "-L" + os.path.dirname(abi.LIB_PATH), "-lsubzero_b200", "-Wl,-rpath," + os.path.dirname(abi.LIB_PATH), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "1000", "1", "1"], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_morton_order_is_a_spatial_renumbering_of_the_same_field():
    """voronoi_field(order="morton"): the same floes (a permutation, outlines intact), consecutive numbers close in space"""
    import numpy as np
    from subzero_b200 import field
    prm, a = sz.voronoi_field(3000, seed=5)
    _, b = sz.voronoi_field(3000, seed=5, order="morton")
    order = field.morton_order(a.x, a.y, prm.Lx, prm.Ly)
    assert sorted(order.tolist()) == list(range(3000)) and np.array_equal(b.x, a.x[order]) and np.array_equal(b.area, a.area[order])
    for k in (0, 1234, 2999):
        xa, ya = a.outline(int(order[k])); xb, yb = b.outline(k)
        assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    hop = lambda f: np.hypot(np.diff(f.x), np.diff(f.y)).mean()
    assert hop(b) < 0.1 * hop(a)                     # neighbours in number are neighbours in space
    with pytest.raises(ValueError):
        sz.voronoi_field(10, order="hilbert")


def test_matlab_drop_in_keeps_the_reference_signature_and_its_pieces_exist():
    """subzero_b200/matlab/floe_interactions_all.m is the file a maintainer puts ahead of the reference's on the path: same
    function line as floe_interactions_all.m:1, the pre-step `x` that the kill rule of :282 reads is defined, ghosts are
    materialised for the ridging / rafting tail, and the tail itself is the generated function (never shipped)."""
    import re
    d = os.path.join(ROOT, "subzero_b200", "matlab")
    src = open(os.path.join(d, "floe_interactions_all.m")).read()
    norm = lambda s: re.sub(r"\s+", "", s)
    want = ("function[Floe,dissolvedNEW]=floe_interactions_all(Floe,floebound,ocean,winds,c2_boundary,dt,HFo,min_floe_size,Nx,Ny,Nb,"
            "dissolvedNEW,doInt,COLLISION,PERIODIC,RIDGING,RAFTING)")
    assert norm(src.splitlines()[0]) == want
    ref = "/root/reference/floe_interactions_all.m"
    if os.path.exists(ref):
        assert norm(open(ref).readline()) == want
    body = norm(src)
    assert "x=cat(1,Floe.Xi);" in body and body.index("x=cat(1,Floe.Xi);") < body.index("sz_contact_step(")          # :282 needs x
    assert "isnan(x(i))" in body and "calc_trajectory(dt,ocean,winds,Floe(i),HFo,doInt)" in body                     # :281-282
    assert "sz_floe_interactions_tail(" in body and "Floe=[FloeG];" in body
    for f in ("sz_contact_step.m", "sz_contact_mex.cpp", "sz_make_tail.m"):
        assert os.path.exists(os.path.join(d, f)), f
    assert "ghost_parent" in open(os.path.join(d, "sz_contact_mex.cpp")).read()
    # nothing of the reference's tail is stored in this repository
    for f in os.listdir(d):
        assert "Ridged" not in open(os.path.join(d, f), errors="ignore").read(), f


def test_tail_generator_cuts_the_reference_tail(tmp_path):
    """tools/make_matlab_tail.py (twin of sz_make_tail.m) on the reference as surveyed: the tail is floe_interactions_all.m
    :288-511, and every variable it reads before writing is an argument of the generated function"""
    import re
    import sys
    ref = "/root/reference/floe_interactions_all.m"
    if not os.path.exists(ref):
        pytest.skip("the reference tree is only present in the build container")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_matlab_tail as mt
    out = mt.make_tail(ref, str(tmp_path / "sz_floe_interactions_tail.m"))
    text = open(out).read().splitlines()
    assert text[0].startswith("function [Floe,dissolvedNEW] = sz_floe_interactions_tail(") and text[-1] == "end"
    first, last, body = mt.cut_tail(open(ref).read())
    assert (first, last) == (288, 511) and text[2:-1] == body
    used = set(re.findall(r"[A-Za-z_]\w*", "\n".join(l.split("%")[0] for l in body)))
    for v in ("Floe", "floebound", "c2_boundary", "c2_boundary_poly", "min_floe_size", "Nx", "Ny", "Nb", "N0", "kill", "transfer", "dissolvedNEW", "doInt",
              "PERIODIC", "RIDGING", "RAFTING", "id", "id3"):
        assert v in used and v in mt.ARGS, v
    # variables of the replaced head that the tail must NOT need
    for v in ("x", "y", "rmax", "FloeNums", "parent", "N", "Lx", "Ly", "Modulus", "ocean", "winds", "dt", "HFo", "COLLISION"):
        assert v not in used, v
