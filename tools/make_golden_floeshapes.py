"""Decode the reference fixture test/test_conservation/FloeShapes.mat (a 1x462 MATLAB polyshape
array, MCOS-opaque) into a plain .npz under tests/golden/.

Run in the BUILD container only (needs /root/reference); the .npz is committed so that the tests
never read /root/reference at run time.  Per polygon we keep the Vertices and MATLAB's own cached
BoundaryInfo = [_, _, 1, area, perimeter, cx, cy, xmin, ymin, xmax, ymax, -1] -- the 462 known
answers that pin the polyshape area/centroid restatement.
"""
import io, sys
import numpy as np
import scipy.io
from scipy.io.matlab._mio5 import MatFile5Reader

REF = "/root/reference/test/test_conservation"

def decode(path):
    raw = scipy.io.loadmat(path)
    fw = raw["__function_workspace__"].tobytes()
    # the function workspace is itself a v5 MAT stream without the 128-byte file header
    hdr = bytearray(128)
    hdr[0:4] = b"MATL"
    hdr[124:126] = (0x0100).to_bytes(2, "little")
    hdr[126:128] = b"IM"
    # bytes 0..7 of the stream are a sub-header (version/endian + padding); variables start after it
    stream = io.BytesIO(bytes(hdr) + fw[8:])
    rdr = MatFile5Reader(stream, struct_as_record=True, squeeze_me=False)
    rdr.initialize_read()
    stream.seek(128)
    hdrv, _ = rdr.read_var_header()
    var = rdr.read_var_array(hdrv, process=False)
    return var

def main(out):
    var = decode(f"{REF}/FloeShapes.mat")
    mcos = var["MCOS"][0, 0] if "MCOS" in var.dtype.names else var
    cells = mcos["_ObjectMetadata"] if hasattr(mcos, "dtype") and mcos.dtype.names and "_ObjectMetadata" in mcos.dtype.names else None
    # robust walk: collect every struct that has Vertices + BoundaryInfo fields
    found = []
    def walk(o, depth=0):
        if depth > 8: return
        if isinstance(o, np.ndarray):
            if o.dtype.names:
                names = o.dtype.names
                if "Vertices" in names and "BoundaryInfo" in names:
                    for el in o.flat:
                        found.append((np.asarray(el["Vertices"], float), np.asarray(el["BoundaryInfo"], float).ravel()))
                    return
                for el in o.flat:
                    for nm in names: walk(el[nm], depth + 1)
            elif o.dtype == object:
                for el in o.flat: walk(el, depth + 1)
    walk(var)
    print("polygons decoded:", len(found))
    nv = np.array([v.shape[0] for v, _ in found], np.int32)
    off = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
    verts = np.concatenate([v for v, _ in found], 0)
    binfo = np.stack([b for _, b in found], 0)
    modulus = float(scipy.io.loadmat(f"{REF}/Modulus.mat")["Modulus"].ravel()[0])
    np.savez_compressed(out, off=off, verts=verts, boundary_info=binfo, modulus=modulus)
    print("vertices min/median/max:", nv.min(), int(np.median(nv)), nv.max(), "modulus", modulus)

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/floe_shapes.npz")
