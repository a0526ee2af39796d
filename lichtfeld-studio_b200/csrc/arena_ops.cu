// lfs_b200 -- densification-side state surgery on the planar arenas.
// The reference edits its six AoS parameter tensors and the Adam moments with index_select + cat per tensor
// (src/training/strategies/default_strategy.cpp:49-230 duplicate / split / prune, mcmc.cpp:112-347 relocate / add,
// strategy_utils.cpp:57-131 update_param_with_optimizer): 18 gathers + 18 concatenations per operation.  On the planar
// layout all of them are ONE row gather: output Gaussian i takes every plane of source Gaussian index[i] -- for the
// parameters and both Adam moments in the same launch, with the moments of new Gaussians zeroed in passing.
#include "common.cuh"

namespace lfs {

constexpr int kArThreads = 256;

// grid.y = plane; one thread per output Gaussian: coalesced writes, gathered reads (L2-friendly: sorted-ish indices)
__global__ void __launch_bounds__(kArThreads)
    k_arena_gather(const float* __restrict__ src_p, const float* __restrict__ src_m, const float* __restrict__ src_v,
                   float* __restrict__ dst_p, float* __restrict__ dst_m, float* __restrict__ dst_v,
                   const int32_t* __restrict__ index, const uint8_t* __restrict__ zero_state, const uint32_t n_out,
                   const uint64_t src_np, const uint64_t dst_np, const uint32_t n_src) {
    const uint32_t i = blockIdx.x * kArThreads + threadIdx.x;
    if (i >= n_out)
        return;
    const int32_t s = __ldg(index + i);
    if (s < 0) // keep the destination slot (in-place relocation leaves untouched Gaussians alone)
        return;
    if ((uint32_t)s >= n_src)
        return;
    const size_t so = (size_t)blockIdx.y * src_np + (uint32_t)s, d = (size_t)blockIdx.y * dst_np + i;
    dst_p[d] = __ldg(src_p + so);
    const bool z = zero_state && zero_state[i];
    if (dst_m)
        dst_m[d] = z ? 0.f : __ldg(src_m + so);
    if (dst_v)
        dst_v[d] = z ? 0.f : __ldg(src_v + so);
}

// writes AoS rows [n, n_planes] into slots index[n] of planes [first_plane, first_plane + n_planes)
__global__ void __launch_bounds__(kArThreads)
    k_arena_set_rows(float* __restrict__ arena, const uint64_t np, const uint32_t first_plane, const uint32_t n_planes,
                     const int32_t* __restrict__ index, const float* __restrict__ rows, const uint32_t n,
                     const uint32_t n_slots) {
    const uint32_t i = blockIdx.x * kArThreads + threadIdx.x;
    if (i >= n)
        return;
    const int32_t s = index ? __ldg(index + i) : (int32_t)i;
    if (s < 0 || (uint32_t)s >= n_slots)
        return;
    for (uint32_t p = 0; p < n_planes; ++p)
        arena[(size_t)(first_plane + p) * np + (uint32_t)s] = __ldg(rows + (size_t)i * n_planes + p);
}

// reads slots index[n] of planes [first_plane, first_plane + n_planes) into AoS rows [n, n_planes]
__global__ void __launch_bounds__(kArThreads)
    k_arena_get_rows(const float* __restrict__ arena, const uint64_t np, const uint32_t first_plane,
                     const uint32_t n_planes, const int32_t* __restrict__ index, float* __restrict__ rows,
                     const uint32_t n, const uint32_t n_slots) {
    const uint32_t i = blockIdx.x * kArThreads + threadIdx.x;
    if (i >= n)
        return;
    const int32_t s = index ? __ldg(index + i) : (int32_t)i;
    for (uint32_t p = 0; p < n_planes; ++p)
        rows[(size_t)i * n_planes + p] =
            (s < 0 || (uint32_t)s >= n_slots) ? 0.f : __ldg(arena + (size_t)(first_plane + p) * np + (uint32_t)s);
}

// zero every element this rank does NOT own under the lfs_adam_step_multi_p2p rule (chunk c -> rank c % world): the sum
// over the ranks of the result is the complete moment arena
constexpr int64_t kOwnChunk4_ = 1024; // must match adam.cu kOwnChunk4
__global__ void __launch_bounds__(kArThreads)
    k_zero_unowned(float4* __restrict__ arena4, const int64_t n4, const int world, const int rank) {
    const int64_t stride = (int64_t)gridDim.x * kArThreads;
    for (int64_t i = (int64_t)blockIdx.x * kArThreads + threadIdx.x; i < n4; i += stride)
        if ((int)((i / kOwnChunk4_) % world) != rank)
            arena4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

} // namespace lfs

extern "C" int lfs_arena_gather(const float* src_params, const float* src_m, const float* src_v, float* dst_params,
                                float* dst_m, float* dst_v, const int32_t* index, const uint8_t* zero_state,
                                uint32_t n_out, uint32_t n_src, uint32_t planes, uint64_t src_plane_elems,
                                uint64_t dst_plane_elems, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(src_params && dst_params && index, "arena_gather: null pointer");
    LFS_CHECK_ARG((src_m != nullptr) == (dst_m != nullptr) && (src_v != nullptr) == (dst_v != nullptr),
                  "arena_gather: moment arenas must be given as source AND destination, or not at all");
    LFS_CHECK_ARG(n_out <= dst_plane_elems && n_src <= src_plane_elems, "arena_gather: plane too short");
    if (n_out == 0 || planes == 0)
        return LFS_OK;
    k_arena_gather<<<dim3(div_up(n_out, kArThreads), planes), kArThreads, 0, (cudaStream_t)stream>>>(
        src_params, src_m, src_v, dst_params, dst_m, dst_v, index, zero_state, n_out, src_plane_elems, dst_plane_elems, n_src);
    LFS_LAUNCH_OK("k_arena_gather");
    return LFS_OK;
}

extern "C" int lfs_arena_set_rows(float* arena, uint64_t plane_elems, uint32_t first_plane, uint32_t n_planes,
                                  const int32_t* index, const float* rows, uint32_t n, uint32_t n_slots, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(arena && rows, "arena_set_rows: null pointer");
    LFS_CHECK_ARG(n_slots <= plane_elems, "arena_set_rows: n_slots exceeds the plane length");
    if (n == 0 || n_planes == 0)
        return LFS_OK;
    k_arena_set_rows<<<div_up(n, kArThreads), kArThreads, 0, (cudaStream_t)stream>>>(arena, plane_elems, first_plane,
                                                                                      n_planes, index, rows, n, n_slots);
    LFS_LAUNCH_OK("k_arena_set_rows");
    return LFS_OK;
}

extern "C" int lfs_arena_get_rows(const float* arena, uint64_t plane_elems, uint32_t first_plane, uint32_t n_planes,
                                  const int32_t* index, float* rows, uint32_t n, uint32_t n_slots, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(arena && rows, "arena_get_rows: null pointer");
    LFS_CHECK_ARG(n_slots <= plane_elems, "arena_get_rows: n_slots exceeds the plane length");
    if (n == 0 || n_planes == 0)
        return LFS_OK;
    k_arena_get_rows<<<div_up(n, kArThreads), kArThreads, 0, (cudaStream_t)stream>>>(arena, plane_elems, first_plane,
                                                                                      n_planes, index, rows, n, n_slots);
    LFS_LAUNCH_OK("k_arena_get_rows");
    return LFS_OK;
}

extern "C" int lfs_adam_p2p_zero_unowned(float* arena, int64_t n_floats, int world, int rank, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(arena && (n_floats & 3) == 0 && (reinterpret_cast<uintptr_t>(arena) & 15u) == 0,
                  "adam_p2p_zero_unowned: arena must be 16-byte aligned, n_floats a multiple of 4");
    LFS_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "adam_p2p_zero_unowned: bad rank %d of %d", rank, world);
    if (n_floats == 0 || world == 1)
        return LFS_OK;
    const int64_t n4 = n_floats / 4, want = (n4 + kArThreads - 1) / kArThreads, cap = (int64_t)num_sms() * 16;
    k_zero_unowned<<<(unsigned)(want < cap ? want : cap), kArThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(arena), n4, world, rank);
    LFS_LAUNCH_OK("k_zero_unowned");
    return LFS_OK;
}
