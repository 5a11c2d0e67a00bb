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
// experiment: class C as a sweep kernel and a force-law kernel (sz_narrow.cuh); the same grid twice, the second launch reads
// what the first one left in the handoff buffers (stream order)
extern "C" void sz_launch_narrow_C_split(const NarrowArgs* a, cudaStream_t stream)
{
    if (a->n_work <= 0) return;
    const int tpb = SZ_C_TPB;
    const size_t smem = szcvx::smem_edge_bytes(tpb);      // non-zero only in builds with SZ_C_SMEM_EDGES: the sweep kernel's edge records
    static bool once = false;
    if (!once) {
        if (smem == 0) cudaFuncSetAttribute(narrow_convex_sweep_kernel<PairS>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        else cudaFuncSetAttribute(narrow_convex_sweep_kernel<PairS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(narrow_convex_force_kernel<PairS>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        once = true;
    }
    narrow_convex_sweep_kernel<PairS><<<(a->n_work + tpb - 1) / tpb, tpb, smem, stream>>>(*a);
    narrow_convex_force_kernel<PairS><<<(a->n_work + tpb - 1) / tpb, tpb, 0, stream>>>(*a);
}
