"""A MATLAB stand-in for the test-suite (test infrastructure): loads the UNMODIFIED mex gateways of subzero_b200/matlab/
(compiled by tests/host/Makefile against tests/host/mex_mock/mex_mock.cpp, a small implementation of the mx/mex calls they
use) and calls their mexFunction(nlhs, plhs, nrhs, prhs) the way MATLAB would -- private/mexclipper.cpp:83 is the
reference's entry point of the same kind.  Python values map to MATLAB values like this:

    dict            <-> 1 x 1 struct                    str   -> char row
    1-D float array <-> N x 1 double column             float -> 1 x 1 double
    2-D float array <-> m x n double, column-major      (what comes back is always a 2-D Fortran-ordered array or a dict)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "host")
_vp, _dp = C.c_void_p, C.POINTER(C.c_double)


class MexError(RuntimeError):
    """what MATLAB would raise from mexErrMsgIdAndTxt(id, ...)"""

    def __init__(self, ident, msg):
        super().__init__("%s: %s" % (ident, msg))
        self.identifier, self.message = ident, msg


def build():
    """the three shared objects of `make -C tests/host mex` (built by __graft_entry__.build(); rebuilt here when absent)"""
    need = [os.path.join(HOST, f) for f in ("libmexmock.so", "sz_contact_mex_host.so", "sz_resident_mex_host.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-s", "-C", HOST, "mex"], check=True)
    return need


class Session:
    def __init__(self):
        paths = build()
        self.mm = mm = C.CDLL(paths[0], mode=C.RTLD_GLOBAL)
        for name, res, args in (("mm_double", _vp, [C.c_size_t, C.c_size_t, _vp]), ("mm_string", _vp, [C.c_char_p]), ("mm_struct", _vp, []),
                                ("mm_set", None, [_vp, C.c_char_p, _vp]), ("mm_kind", C.c_int, [_vp]), ("mm_rows", C.c_size_t, [_vp]),
                                ("mm_cols", C.c_size_t, [_vp]), ("mm_nfields", C.c_int, [_vp]), ("mm_field_name", C.c_char_p, [_vp, C.c_int]),
                                ("mm_call", C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.c_int, C.POINTER(_vp), C.c_char_p, C.c_size_t]),
                                ("mm_clear", C.c_int, []), ("mm_free_all", None, []), ("mxGetPr", _dp, [_vp]), ("mxGetField", _vp, [_vp, C.c_size_t, C.c_char_p])):
            f = getattr(mm, name)
            f.restype, f.argtypes = res, args
        self._gateways = {}

    def gateway(self, name):
        """address of mexFunction in tests/host/<name>_host.so (name: sz_contact_mex / sz_resident_mex)"""
        if name not in self._gateways:
            lib = C.CDLL(os.path.join(HOST, name + "_host.so"))
            self._gateways[name] = (lib, C.cast(lib.mexFunction, _vp))
        return self._gateways[name][1]

    # ---- Python -> mxArray
    def to_mx(self, v):
        if v is None:
            return None
        if isinstance(v, dict):
            s = self.mm.mm_struct()
            for k, x in v.items():
                self.mm.mm_set(s, k.encode(), self.to_mx(x))
            return s
        if isinstance(v, str):
            return self.mm.mm_string(v.encode())
        a = np.asarray(v, dtype=np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        a = np.asfortranarray(a)
        return self.mm.mm_double(a.shape[0], a.shape[1], a.ctypes.data_as(_vp))

    # ---- mxArray -> Python
    def from_mx(self, p):
        if not p:
            return None
        kind = self.mm.mm_kind(p)
        if kind == 1:
            return {self.mm.mm_field_name(p, k).decode(): self.from_mx(self.mm.mxGetField(p, 0, self.mm.mm_field_name(p, k))) for k in range(self.mm.mm_nfields(p))}
        if kind == 0:
            m, n = self.mm.mm_rows(p), self.mm.mm_cols(p)
            if m * n == 0:
                return np.zeros((m, n), order="F")
            return np.ctypeslib.as_array(self.mm.mxGetPr(p), shape=(n, m)).T.copy(order="F")
        raise TypeError("char results are not used by the gateways")

    def call(self, name, *args, nlhs=1):
        """[out] = name(args...); returns the Python form of plhs[0] (None when the gateway left it unset)"""
        prhs = (_vp * max(1, len(args)))(*[self.to_mx(a) for a in args])
        plhs = (_vp * max(1, nlhs))()
        err = C.create_string_buffer(4096)
        rc = self.mm.mm_call(self.gateway(name), nlhs, plhs, len(args), prhs, err, len(err))
        if rc:
            ident, _, msg = err.value.decode("utf-8", "replace").partition("|")
            raise MexError(ident, msg)
        return self.from_mx(plhs[0])

    def clear(self):
        """`clear mex`: the mexAtExit handlers run (device contexts are released), the session's arrays are freed"""
        k = self.mm.mm_clear()
        self.mm.mm_free_all()
        return k


def col(a):
    """N x 1 result -> 1-D array"""
    return np.asarray(a).reshape(-1, order="F")


def soa_struct(soa):
    """the `soa` argument of the gateways, exactly what sz_contact_step.m builds from the Floe struct array"""
    d = {k: getattr(soa, k) for k in ("x", "y", "rmax", "h", "area", "u", "v", "ksi", "vx", "vy")}
    d["alive"] = soa.alive.astype(np.float64)
    d["voff"] = soa.voff.astype(np.float64)
    return d


def prm_struct(prm):
    return {"Lx": prm.Lx, "Ly": prm.Ly, "modulus": prm.modulus, "dt": prm.dt, "Nb": float(prm.Nb), "periodic": float(prm.periodic), "collision": float(prm.collision)}


def bnd_struct(b):
    return {"x": b.x, "y": b.y, "box_x": b.box_x, "box_y": b.box_y, "area": b.area, "h": b.h, "xi": b.xi, "yi": b.yi, "u": b.u, "v": b.v, "ksi": b.ksi}
