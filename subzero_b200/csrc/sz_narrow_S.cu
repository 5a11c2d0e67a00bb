// class S instantiation of the narrow phase (see sz_narrow.cuh)
#include "sz_narrow.cuh"
using namespace sznarrow;
extern "C" void sz_launch_narrow_S(const NarrowArgs* a, cudaStream_t stream)
{
    if (a->n_work <= 0) return;
    const int tpb = SZ_S_TPB;
    narrow_local_kernel<PairS><<<(a->n_work + tpb - 1) / tpb, tpb, 0, stream>>>(*a);
}
extern "C" void sz_launch_clip_S(const ClipArgs* a, cudaStream_t stream)
{
    if (a->count <= 0) return;
    const int tpb = 128;
    clip_local_kernel<ClipS><<<(a->count + tpb - 1) / tpb, tpb, 0, stream>>>(*a);
}
extern "C" void sz_launch_euler_S(const szeul::EulerArgs* a, int* next_list, int* next_count, cudaStream_t stream)
{
    if (a->n_items <= 0) return;
    const int tpb = 128;
    euler_item_local_kernel<ClipS><<<(a->n_items + tpb - 1) / tpb, tpb, 0, stream>>>(*a, next_list, next_count);
}
