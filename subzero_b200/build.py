"""Builds the product library subzero_b200/_lib/libsubzero_b200.so for sm_100a, in-tree.

nvcc cross-compiles without a GPU.  The narrow-phase size classes live in separate translation units
(sz_narrow_{S,M,L}.cu) so that their template instantiations compile in parallel.  Device code is built
with -fmad=false: Clipper's int64 output and the force law's thresholds depend on individually rounded
FP64 operations, and the reference binary contains no FMA.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("SZ_BUILD_DIR", os.path.join(HERE, "_lib"))       # SZ_BUILD_DIR / SZ_EXTRA_NVCC: experiment builds
LIB = os.path.join(OUT, "libsubzero_b200.so")
FIELD_LIB = os.path.join(OUT, "libsz_field.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2"] + os.environ.get("SZ_EXTRA_NVCC", "").split()
CU = ["sz_contact.cu", "sz_narrow_C.cu", "sz_narrow_S.cu", "sz_narrow_T.cu", "sz_narrow_M.cu", "sz_narrow_L.cu"]
CPP = ["sz_field.cpp"]
HEADERS = ["sz_clip.cuh", "sz_convex.cuh", "sz_pairforce.cuh", "sz_narrow.cuh", "sz_corners.cuh", "sz_euler.cuh", "sz_apart.cuh", os.path.join("..", "..", "include", "subzero_b200.h")]


def kernel_stamp(files=("sz_narrow_C.cu", "sz_narrow.cuh", "sz_convex.cuh", "sz_pairforce.cuh", "sz_clip.cuh")):
    """sha256 over the sources of the dominant kernel (class C of the narrow phase).  profiles/narrow_traffic.json records it
    when the ncu capture is published; bench.py reports `roofline.traffic` only while the stamp still matches, i.e. while the
    measured DRAM traffic belongs to the kernel that is being timed."""
    import hashlib
    h = hashlib.sha256()
    for f in files:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:16]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(verbose=False, force=False):
    os.makedirs(OUT, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    objs = []
    for f in CU:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OUT, f + ".o")
        objs.append(obj)
        if force or _newer([src] + hdrs, obj):
            jobs.append([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj])
    for f in CPP:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OUT, f + ".o")
        objs.append(obj)
        if force or _newer([src] + hdrs, obj):
            jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread", "-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out.strip():
                    print(out)
    if jobs or not os.path.exists(LIB):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread"])
    # the synthetic-field generator alone (host code, no CUDA): what the CPU reference arm of bench.py loads, so that it never
    # maps the product library
    if _newer([os.path.join(CSRC, "sz_field.cpp")] + hdrs, FIELD_LIB):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread", "-shared", "-DSZ_FIELD_STANDALONE", os.path.join(CSRC, "sz_field.cpp"), "-o", FIELD_LIB])
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
