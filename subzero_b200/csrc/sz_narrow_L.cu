// class L instantiation of the narrow phase (see sz_narrow.cuh)
#include "sz_narrow.cuh"
using namespace sznarrow;
extern "C" size_t sz_workspace_bytes_L(void) { return sizeof(szpf::Workspace<PairL>); }
extern "C" void sz_launch_narrow_L(const NarrowArgs* a, cudaStream_t stream)
{
    if (a->n_threads <= 0) return;
    const int tpb = 64;
    narrow_scratch_kernel<PairL><<<(a->n_threads + tpb - 1) / tpb, tpb, 0, stream>>>(*a);
}
extern "C" void sz_launch_clip_L(const ClipArgs* a, cudaStream_t stream)
{
    if (a->n_threads <= 0) return;
    const int tpb = 64;
    clip_scratch_kernel<PairL><<<(a->n_threads + tpb - 1) / tpb, tpb, 0, stream>>>(*a);
}
extern "C" void sz_launch_fracture_L(const FractureArgs* a, cudaStream_t stream)
{
    if (a->n_threads <= 0 || a->count <= 0) return;
    const int tpb = 64;
    fracture_deform_kernel<PairL><<<(a->n_threads + tpb - 1) / tpb, tpb, 0, stream>>>(*a);
}
extern "C" void sz_launch_euler_L(const szeul::EulerArgs* a, const int* list, const int* list_count, void* scratch, int n_threads, cudaStream_t stream)
{
    if (n_threads <= 0) return;
    const int tpb = 64;
    euler_item_scratch_kernel<PairL><<<(n_threads + tpb - 1) / tpb, tpb, 0, stream>>>(*a, list, list_count, scratch, n_threads);
}
