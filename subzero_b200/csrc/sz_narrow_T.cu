// class T instantiation of the narrow phase (see sz_narrow.cuh): <= 47 vertices per outline, arena in local memory
#include "sz_narrow.cuh"
using namespace sznarrow;
extern "C" void sz_launch_narrow_T(const NarrowArgs* a, cudaStream_t stream)
{
    if (a->n_work <= 0) return;
    const int tpb = SZ_S_TPB;
    narrow_local_kernel<PairT><<<(a->n_work + tpb - 1) / tpb, tpb, 0, stream>>>(*a);
}
