// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin C wrapper around the UNMODIFIED reference Clipper 6.4.2
// (/root/reference/private/clipper.cpp, compiled where it lies by oracle/Makefile
// into oracle/_ref/libclipper_ref.so).  It performs exactly the call sequence of the
// reference mex gateway's boolean branch (private/mexclipper.cpp:291-298):
//     Clipper c; c.AddPaths(subj, ptSubject, true); c.AddPaths(clip, ptClip, true);
//     c.Execute(CT, solution, pftEvenOdd, pftEvenOdd);
// with the method numbering of private/mexclipper.cpp:206-230 / polyclip.m:52-58
// (0 = difference, 1 = intersection, 2 = xor, 3 = union).
#include "clipper.hpp"
#include <cstdint>
#include <cstring>

using namespace ClipperLib;

extern "C" {

// Returns the number of output paths (>=0), -1 on "Clipper Error." (Execute false or
// exception, mexclipper.cpp:303-304), -2 if the caller's buffers are too small.
// Output: path k occupies out_x/out_y[out_off[k] .. out_off[k+1]).
int szref_clip(const int64_t* sx, const int64_t* sy, int ns,
               const int64_t* cx, const int64_t* cy, int nc,
               int method,
               int64_t* out_x, int64_t* out_y, int out_cap,
               int* out_off, int off_cap)
{
    Paths subj(1), clip(1), sol;
    subj[0].resize(ns);
    for (int i = 0; i < ns; ++i) { subj[0][i].X = sx[i]; subj[0][i].Y = sy[i]; }
    clip[0].resize(nc);
    for (int i = 0; i < nc; ++i) { clip[0][i].X = cx[i]; clip[0][i].Y = cy[i]; }
    ClipType ct;
    switch (method) {
        case 0: ct = ctDifference; break;
        case 1: ct = ctIntersection; break;
        case 2: ct = ctXor; break;
        case 3: ct = ctUnion; break;
        default: return -1;
    }
    bool ok = false;
    try {
        Clipper c;
        c.AddPaths(subj, ptSubject, true);
        c.AddPaths(clip, ptClip, true);
        ok = c.Execute(ct, sol, pftEvenOdd, pftEvenOdd);
    } catch (...) {
        return -1;
    }
    if (!ok) return -1;
    int np = (int)sol.size();
    if (np + 1 > off_cap) return -2;
    int pos = 0;
    out_off[0] = 0;
    for (int k = 0; k < np; ++k) {
        int m = (int)sol[k].size();
        if (pos + m > out_cap) return -2;
        for (int v = 0; v < m; ++v) { out_x[pos + v] = sol[k][v].X; out_y[pos + v] = sol[k][v].Y; }
        pos += m;
        out_off[k + 1] = pos;
    }
    return np;
}

const char* szref_version() { return CLIPPER_VERSION; }

}  // extern "C"
