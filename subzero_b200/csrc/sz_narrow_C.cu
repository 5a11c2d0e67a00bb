// class C instantiation of the narrow phase (see sz_narrow.cuh): strictly convex pairs, no arena
#include "sz_narrow.cuh"
using namespace sznarrow;
extern "C" void sz_launch_narrow_C(const NarrowArgs* a, cudaStream_t stream)
{
    if (a->n_work <= 0) return;
    const int tpb = SZ_C_TPB;
    // the kernel uses no shared memory: ask for the whole unified array as L1 (its working set is per-thread local memory)
    // (experiment switch SZ_C_SMEM_EDGES: the sweep's edge records in dynamic shared memory, see sz_convex.cuh)
    const size_t smem = szcvx::smem_edge_bytes(tpb);
    static bool once = false;
    if (!once) {
        if (smem == 0) cudaFuncSetAttribute(narrow_convex_kernel<PairS>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        else cudaFuncSetAttribute(narrow_convex_kernel<PairS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        once = true;
    }
    narrow_convex_kernel<PairS><<<(a->n_work + tpb - 1) / tpb, tpb, smem, stream>>>(*a);
}
