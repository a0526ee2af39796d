// lfs_b200 -- tile intersection: building blocks + the gsplat-surface ops lfs_intersect_tile /
// lfs_intersect_offset (drop-ins for gsplat::intersect_tile / intersect_offset, reference
// gsplat/Intersect.cpp:15-137, kernels gsplat/IntersectTile.cu:24-114, :206-252, CUB sort :290-328).
// Integer outputs are bit-identical to the reference for identical (means2d, radii, depths).
#include "intersect.cuh"
#include "projection.cuh"

#include <vector>

namespace lfs {

constexpr int kIsThreads = 256;

__global__ void __launch_bounds__(kIsThreads)
    k_tile_count(const float* __restrict__ means2d, const int32_t* __restrict__ radii,
                 const float* __restrict__ depths, const uint32_t n, const float tile_size, const uint32_t tile_w,
                 const uint32_t tile_h, int32_t* __restrict__ tiles_per_gauss, TileRect* __restrict__ rects,
                 uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ ident) {
    const uint32_t idx = blockIdx.x * kIsThreads + threadIdx.x;
    if (idx >= n)
        return;
    const int2 r = __ldg(reinterpret_cast<const int2*>(radii) + idx);
    int32_t cnt = 0;
    TileRect tr = {0, 0, 0, 0};
    if (r.x > 0 && r.y > 0) {
        const float2 m = __ldg(reinterpret_cast<const float2*>(means2d) + idx);
        uint32_t x0, y0, x1, y1;
        tile_rect(m.x, m.y, (float)r.x, (float)r.y, tile_size, tile_w, tile_h, x0, y0, x1, y1);
        cnt = (int32_t)((y1 - y0) * (x1 - x0));
        tr.x0 = (unsigned short)x0, tr.y0 = (unsigned short)y0, tr.x1 = (unsigned short)x1, tr.y1 = (unsigned short)y1;
    }
    tiles_per_gauss[idx] = cnt;
    rects[idx] = tr;
    if (depth_keys)
        depth_keys[idx] = cnt > 0 ? __float_as_uint(__ldg(depths + idx)) : 0xFFFFFFFFu;
    if (ident)
        ident[idx] = idx;
}

int launch_tile_count(const float* means2d, const int32_t* radii, const float* depths, uint32_t n, float tile_size,
                      uint32_t tile_w, uint32_t tile_h, int32_t* tiles_per_gauss, TileRect* rects,
                      uint32_t* depth_keys, uint32_t* ident, cudaStream_t stream) {
    if (n == 0)
        return LFS_OK;
    k_tile_count<<<div_up(n, kIsThreads), kIsThreads, 0, stream>>>(means2d, radii, depths, n, tile_size, tile_w,
                                                                   tile_h, tiles_per_gauss, rects, depth_keys, ident);
    LFS_LAUNCH_OK("k_tile_count");
    return LFS_OK;
}

// largest i in [0, n) with off[i] <= j   (off is non-decreasing, off[0] == 0)
__device__ __forceinline__ uint32_t upper_slot(const uint32_t* __restrict__ off, uint32_t n, uint32_t j) {
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= j)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// Emission: one thread per slot of the depth order writes that Gaussian's run of instances (contiguous at off[slot],
// ~9 entries = two 32-B sectors per output array: the scattered stores merge in L2).  History: the instance-parallel
// version (20-step binary search over `off` per instance) took 0.150 ms per 1080p view, a warp-cooperative version that
// serialised 32 Gaussians per warp 0.085 ms (ncu r01c / r01f).
__global__ void __launch_bounds__(kIsThreads)
    k_emit_instances(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ off, const uint32_t n_gauss,
                     const TileRect* __restrict__ rects, const uint32_t tile_w, const uint32_t id_offset,
                     const uint32_t n_cap, const uint32_t* __restrict__ n_dev, uint32_t* __restrict__ tile_keys,
                     uint32_t* __restrict__ vals) {
    uint32_t n = n_cap;
    if (n_dev) {
        const uint32_t nd = *n_dev;
        n = nd < n_cap ? nd : n_cap;
    }
    for (uint32_t slot = blockIdx.x * kIsThreads + threadIdx.x; slot < n_gauss; slot += gridDim.x * kIsThreads) {
        const uint32_t g = perm ? __ldg(perm + slot) : slot;
        const TileRect r = rects[g];
        if (r.x1 <= r.x0 || r.y1 <= r.y0)
            continue;
        uint32_t pos = __ldg(off + slot);
        for (uint32_t ty = r.y0; ty < r.y1; ++ty)
            for (uint32_t tx = r.x0; tx < r.x1; ++tx, ++pos)
                if (pos < n) {
                    tile_keys[pos] = ty * tile_w + tx;
                    vals[pos] = g + id_offset;
                }
    }
}

int launch_emit_instances(const uint32_t* perm, const uint32_t* off, uint32_t n_gauss, const TileRect* rects,
                          uint32_t tile_w, uint32_t id_offset, uint32_t n_cap, const uint32_t* n_dev,
                          uint32_t* tile_keys, uint32_t* vals, cudaStream_t stream) {
    if (n_cap == 0 || n_gauss == 0)
        return LFS_OK;
    const unsigned want = div_up(n_gauss, kIsThreads);
    const unsigned grid = want < (unsigned)(num_sms() * 16) ? want : (unsigned)(num_sms() * 16);
    k_emit_instances<<<grid, kIsThreads, 0, stream>>>(perm, off, n_gauss, rects, tile_w, id_offset, n_cap, n_dev,
                                                      tile_keys, vals);
    LFS_LAUNCH_OK("k_emit_instances");
    return LFS_OK;
}

// emission with the exact tile test (see CullRec): ONE THREAD per slot of the depth order walks the rows of its
// rectangle, computes each row span once (the very predicate the count used) and writes its run.  A Gaussian's run is
// ~7 instances, i.e. one 32-B sector per output array, so the scattered 4-B stores merge in L2; the warp-cooperative
// variants tried first (32 Gaussians serialised per warp, span per tile) cost 0.2 ms per 1080p view, this one ~0.05.
__global__ void __launch_bounds__(kIsThreads)
    k_emit_instances_cull(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ off, const uint32_t n_gauss,
                          const TileRect* __restrict__ rects, const int32_t* __restrict__ counts,
                          const CullRec* __restrict__ cull, const uint32_t tile_w, const uint32_t n_cap,
                          const uint32_t* __restrict__ n_dev, uint32_t* __restrict__ tile_keys,
                          uint32_t* __restrict__ vals) {
    uint32_t n = n_cap;
    if (n_dev) {
        const uint32_t nd = *n_dev;
        n = nd < n_cap ? nd : n_cap;
    }
    for (uint32_t slot = blockIdx.x * kIsThreads + threadIdx.x; slot < n_gauss; slot += gridDim.x * kIsThreads) {
        const uint32_t g = perm ? __ldg(perm + slot) : slot;
        if (counts[g] <= 0)
            continue;
        const TileRect r = rects[g];
        const float4* cp = reinterpret_cast<const float4*>(cull + g);
        const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
        const CullRec cr{c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        uint32_t pos = __ldg(off + slot);
        for (uint32_t ty = r.y0; ty < r.y1; ++ty) {
            int first, last;
            cull_row_span(cr, ty, r.x0, r.x1, first, last);
            for (int tx = first; tx <= last; ++tx, ++pos)
                if (pos < n) {
                    tile_keys[pos] = ty * tile_w + (uint32_t)tx;
                    vals[pos] = g;
                }
        }
    }
}

int launch_emit_instances_cull(const uint32_t* perm, const uint32_t* off, uint32_t n_gauss, const TileRect* rects,
                               const int32_t* counts, const CullRec* cull, uint32_t tile_w, uint32_t n_cap,
                               const uint32_t* n_dev, uint32_t* tile_keys, uint32_t* vals, cudaStream_t stream) {
    if (n_cap == 0 || n_gauss == 0)
        return LFS_OK;
    const unsigned want = div_up(n_gauss, kIsThreads);
    const unsigned grid = want < (unsigned)(num_sms() * 16) ? want : (unsigned)(num_sms() * 16);
    k_emit_instances_cull<<<grid, kIsThreads, 0, stream>>>(perm, off, n_gauss, rects, counts, cull, tile_w, n_cap, n_dev,
                                                           tile_keys, vals);
    LFS_LAUNCH_OK("k_emit_instances_cull");
    return LFS_OK;
}

__global__ void __launch_bounds__(kIsThreads)
    k_tile_offsets(const uint32_t* __restrict__ keys, const uint32_t n_cap, const uint32_t* __restrict__ n_dev,
                   const uint32_t n_tiles, int32_t* __restrict__ offsets) {
    uint32_t n = n_cap;
    if (n_dev) {
        const uint32_t nd = *n_dev;
        n = nd < n_cap ? nd : n_cap;
    }
    if (n == 0) {
        for (uint32_t t = blockIdx.x * kIsThreads + threadIdx.x; t <= n_tiles; t += gridDim.x * kIsThreads)
            offsets[t] = 0;
        return;
    }
    for (uint32_t j = blockIdx.x * kIsThreads + threadIdx.x; j < n; j += gridDim.x * kIsThreads) {
        const uint32_t cur = __ldg(keys + j);
        if (j == 0) {
            for (uint32_t t = 0; t <= cur && t <= n_tiles; ++t)
                offsets[t] = 0;
        } else {
            const uint32_t prev = __ldg(keys + j - 1);
            for (uint32_t t = prev + 1; t <= cur && t <= n_tiles; ++t)
                offsets[t] = (int32_t)j;
        }
        if (j == n - 1)
            for (uint32_t t = cur + 1; t <= n_tiles; ++t)
                offsets[t] = (int32_t)n;
    }
}

int launch_tile_offsets(const uint32_t* sorted_tile_keys, uint32_t n_cap, const uint32_t* n_dev, uint32_t n_tiles,
                        int32_t* offsets, cudaStream_t stream) {
    const unsigned want = div_up(n_cap > n_tiles + 1 ? n_cap : n_tiles + 1, kIsThreads);
    const unsigned grid = want < (unsigned)(num_sms() * 16) ? want : (unsigned)(num_sms() * 16);
    k_tile_offsets<<<grid, kIsThreads, 0, stream>>>(sorted_tile_keys, n_cap, n_dev, n_tiles, offsets);
    LFS_LAUNCH_OK("k_tile_offsets");
    return LFS_OK;
}

// ---- gsplat-surface only -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kIsThreads)
    k_make_isect_ids(const uint32_t* __restrict__ tile_keys, const uint32_t* __restrict__ vals,
                     const float* __restrict__ depths, const uint32_t n, const int64_t cid_enc,
                     int64_t* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids) {
    const uint32_t j = blockIdx.x * kIsThreads + threadIdx.x;
    if (j >= n)
        return;
    const uint32_t g = __ldg(vals + j);
    const int64_t depth_enc = (int64_t)__float_as_uint(__ldg(depths + g));
    isect_ids[j] = cid_enc | ((int64_t)__ldg(tile_keys + j) << 32) | depth_enc;
    flatten_ids[j] = (int32_t)g;
}

// unsorted emission in flattened-index order (the reference's second pass, IntersectTile.cu:95-113)
__global__ void __launch_bounds__(kIsThreads)
    k_emit_unsorted(const uint32_t* __restrict__ off, const uint32_t n_gauss, const uint32_t N,
                    const TileRect* __restrict__ rects, const float* __restrict__ depths, const uint32_t tile_w,
                    const uint32_t tile_n_bits, const uint32_t n, int64_t* __restrict__ isect_ids,
                    int32_t* __restrict__ flatten_ids) {
    for (uint32_t j = blockIdx.x * kIsThreads + threadIdx.x; j < n; j += gridDim.x * kIsThreads) {
        const uint32_t g = upper_slot(off, n_gauss, j);
        const uint32_t k = j - __ldg(off + g);
        const TileRect r = rects[g];
        const uint32_t w = (uint32_t)r.x1 - (uint32_t)r.x0;
        const int64_t tile_id = (int64_t)((r.y0 + k / w) * tile_w + (r.x0 + k % w));
        const int64_t cid_enc = (int64_t)(g / N) << (32 + tile_n_bits);
        isect_ids[j] = cid_enc | (tile_id << 32) | (int64_t)__float_as_uint(__ldg(depths + g));
        flatten_ids[j] = (int32_t)g;
    }
}

// ---- packed layout (gsplat/Intersect.cpp:32-39, IntersectTile.cu:83-90): element e belongs to camera camera_ids[e] ----
// the camera id goes above the tile bits of the second-level key, as in the reference's 64-bit key
__global__ void __launch_bounds__(kIsThreads)
    k_add_camera_bits(uint32_t* __restrict__ tile_keys, const uint32_t* __restrict__ vals,
                      const int64_t* __restrict__ camera_ids, const uint32_t n, const uint32_t tile_n_bits) {
    const uint32_t j = blockIdx.x * kIsThreads + threadIdx.x;
    if (j < n)
        tile_keys[j] |= (uint32_t)__ldg(camera_ids + __ldg(vals + j)) << tile_n_bits;
}
__global__ void __launch_bounds__(kIsThreads)
    k_emit_unsorted_packed(const uint32_t* __restrict__ off, const uint32_t n_gauss, const int64_t* __restrict__ camera_ids,
                           const TileRect* __restrict__ rects, const float* __restrict__ depths, const uint32_t tile_w,
                           const uint32_t tile_n_bits, const uint32_t n, int64_t* __restrict__ isect_ids,
                           int32_t* __restrict__ flatten_ids) {
    for (uint32_t j = blockIdx.x * kIsThreads + threadIdx.x; j < n; j += gridDim.x * kIsThreads) {
        const uint32_t g = upper_slot(off, n_gauss, j);
        const uint32_t k = j - __ldg(off + g);
        const TileRect r = rects[g];
        const uint32_t w = (uint32_t)r.x1 - (uint32_t)r.x0;
        const int64_t tile_id = (int64_t)((r.y0 + k / w) * tile_w + (r.x0 + k % w));
        const int64_t cid_enc = __ldg(camera_ids + g) << (32 + tile_n_bits);
        isect_ids[j] = cid_enc | (tile_id << 32) | (int64_t)__float_as_uint(__ldg(depths + g));
        flatten_ids[j] = (int32_t)g;
    }
}

__global__ void __launch_bounds__(kIsThreads)
    k_intersect_offset(const int64_t* __restrict__ isect_ids, const uint32_t n, const uint32_t n_tiles,
                       const uint32_t tile_n_bits, const uint32_t total, int32_t* __restrict__ offsets) {
    const uint32_t j = blockIdx.x * kIsThreads + threadIdx.x;
    if (j >= n)
        return;
    const int64_t hi = isect_ids[j] >> 32;
    const int64_t cur = (hi >> tile_n_bits) * n_tiles + (hi & (((int64_t)1 << tile_n_bits) - 1));
    int64_t first;
    if (j == 0) {
        first = 0;
    } else {
        const int64_t ph = isect_ids[j - 1] >> 32;
        first = (ph >> tile_n_bits) * n_tiles + (ph & (((int64_t)1 << tile_n_bits) - 1)) + 1;
    }
    for (int64_t t = first; t <= cur && t < total; ++t)
        offsets[t] = (int32_t)j;
    if (j == n - 1)
        for (int64_t t = cur + 1; t < total; ++t)
            offsets[t] = (int32_t)n;
}

} // namespace lfs

extern "C" int lfs_intersect_tile(const float* means2d, const int32_t* radii, const float* depths, uint32_t C,
                                  uint32_t N, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                  int sort, int32_t* tiles_per_gauss, lfs_alloc_fn alloc, void* alloc_ctx,
                                  int64_t** isect_ids, int32_t** flatten_ids, int64_t* n_isects_host, void* stream_) {
    using namespace lfs;
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(alloc && isect_ids && flatten_ids && n_isects_host, "intersect_tile: null callback / out pointer");
    LFS_CHECK_ARG(tile_size > 0 && tile_width > 0 && tile_height > 0, "intersect_tile: bad tile geometry");
    LFS_CHECK_ARG(tile_width < 65536 && tile_height < 65536, "intersect_tile: more than 65535 tiles per axis");
    const uint64_t n64 = (uint64_t)C * N;
    LFS_CHECK_ARG(n64 < (1ull << 31), "intersect_tile: C*N too large");
    const uint32_t n = (uint32_t)n64;
    const uint32_t n_tiles = tile_width * tile_height;
    const uint32_t tnb = ref_tile_n_bits(n_tiles);
    *isect_ids = nullptr;
    *flatten_ids = nullptr;
    *n_isects_host = 0;
    if (n == 0)
        return LFS_OK;
    LFS_CHECK_ARG(means2d && radii && depths && tiles_per_gauss, "intersect_tile: null input");

    // scratch A: rects, depth keys x2, perm x2, offsets, per-camera totals, scan + sort scratch
    Carver ca(nullptr);
    ca.take<TileRect>(n);
    ca.take<uint32_t>(n), ca.take<uint32_t>(n), ca.take<uint32_t>(n), ca.take<uint32_t>(n);
    ca.take<uint32_t>(n);
    ca.take<uint32_t>(C + 1);
    ca.take<char>(scan_scratch_bytes(n));
    ca.take<char>(radix_scratch_bytes(N));
    void* blob_a = alloc(alloc_ctx, LFS_TAG_SCRATCH, ca.total());
    if (!blob_a) {
        set_error("intersect_tile: scratch allocation of %zu bytes failed", ca.total());
        return LFS_ERR_ALLOC;
    }
    Carver a(blob_a);
    TileRect* rects = a.take<TileRect>(n);
    uint32_t* dk_a = a.take<uint32_t>(n);
    uint32_t* dk_b = a.take<uint32_t>(n);
    uint32_t* pm_a = a.take<uint32_t>(n);
    uint32_t* pm_b = a.take<uint32_t>(n);
    uint32_t* off = a.take<uint32_t>(n);
    uint32_t* totals = a.take<uint32_t>(C + 1);
    void* scan_scr = a.take<char>(scan_scratch_bytes(n));
    void* sort_scr = a.take<char>(radix_scratch_bytes(N));

    int rc = launch_tile_count(means2d, radii, depths, n, (float)tile_size, tile_width, tile_height, tiles_per_gauss,
                               rects, sort ? dk_a : nullptr, sort ? pm_a : nullptr, stream);
    if (rc)
        return rc;
    const uint32_t* counts = reinterpret_cast<const uint32_t*>(tiles_per_gauss);

    if (!sort) {
        rc = exclusive_scan_u32(counts, nullptr, off, totals, n, nullptr, scan_scr, stream);
        if (rc)
            return rc;
        uint32_t total = 0;
        LFS_CUDA_OK(cudaMemcpyAsync(&total, totals, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        LFS_CUDA_OK(cudaStreamSynchronize(stream));
        *n_isects_host = total;
        if (total == 0)
            return LFS_OK;
        int64_t* ids = static_cast<int64_t*>(alloc(alloc_ctx, LFS_TAG_ISECT_IDS, sizeof(int64_t) * (size_t)total));
        int32_t* flat = static_cast<int32_t*>(alloc(alloc_ctx, LFS_TAG_FLATTEN_IDS, sizeof(int32_t) * (size_t)total));
        if (!ids || !flat) {
            set_error("intersect_tile: output allocation failed (n_isects=%u)", total);
            return LFS_ERR_ALLOC;
        }
        const unsigned want = div_up(total, kIsThreads);
        const unsigned grid = want < (unsigned)(num_sms() * 16) ? want : (unsigned)(num_sms() * 16);
        k_emit_unsorted<<<grid, kIsThreads, 0, stream>>>(off, n, N, rects, depths, tile_width, tnb, total, ids, flat);
        LFS_LAUNCH_OK("k_emit_unsorted");
        *isect_ids = ids;
        *flatten_ids = flat;
        return LFS_OK;
    }

    // sorted path: per camera depth sort + depth-ordered offsets
    std::vector<const uint32_t*> perm_of(C);
    for (uint32_t c = 0; c < C; ++c) {
        int in_b = 0;
        rc = radix_sort_pairs(dk_a + (size_t)c * N, pm_a + (size_t)c * N, dk_b + (size_t)c * N, pm_b + (size_t)c * N, N,
                              nullptr, 0, 32, sort_scr, &in_b, stream);
        if (rc)
            return rc;
        perm_of[c] = (in_b ? pm_b : pm_a) + (size_t)c * N;
        rc = exclusive_scan_u32(counts, perm_of[c], off + (size_t)c * N, totals + c, N, nullptr, scan_scr, stream);
        if (rc)
            return rc;
    }
    std::vector<uint32_t> h_tot(C);
    LFS_CUDA_OK(cudaMemcpyAsync(h_tot.data(), totals, sizeof(uint32_t) * C, cudaMemcpyDeviceToHost, stream));
    LFS_CUDA_OK(cudaStreamSynchronize(stream));
    uint64_t total64 = 0;
    uint32_t max_cam = 0;
    for (uint32_t c = 0; c < C; ++c) {
        total64 += h_tot[c];
        max_cam = h_tot[c] > max_cam ? h_tot[c] : max_cam;
    }
    LFS_CHECK_ARG(total64 < (1ull << 31), "intersect_tile: n_isects %llu does not fit int32 ids",
                  (unsigned long long)total64);
    *n_isects_host = (int64_t)total64;
    if (total64 == 0)
        return LFS_OK;
    int64_t* ids = static_cast<int64_t*>(alloc(alloc_ctx, LFS_TAG_ISECT_IDS, sizeof(int64_t) * (size_t)total64));
    int32_t* flat = static_cast<int32_t*>(alloc(alloc_ctx, LFS_TAG_FLATTEN_IDS, sizeof(int32_t) * (size_t)total64));
    Carver cb(nullptr);
    cb.take<uint32_t>(max_cam), cb.take<uint32_t>(max_cam), cb.take<uint32_t>(max_cam), cb.take<uint32_t>(max_cam);
    cb.take<char>(radix_scratch_bytes(max_cam));
    void* blob_b = alloc(alloc_ctx, LFS_TAG_SCRATCH, cb.total());
    if (!ids || !flat || !blob_b) {
        set_error("intersect_tile: output allocation failed (n_isects=%llu)", (unsigned long long)total64);
        return LFS_ERR_ALLOC;
    }
    Carver b(blob_b);
    uint32_t* tk_a = b.take<uint32_t>(max_cam);
    uint32_t* tk_b = b.take<uint32_t>(max_cam);
    uint32_t* tv_a = b.take<uint32_t>(max_cam);
    uint32_t* tv_b = b.take<uint32_t>(max_cam);
    void* sort_scr_b = b.take<char>(radix_scratch_bytes(max_cam));
    const int tbits = tile_key_bits(n_tiles);
    uint64_t cam_off = 0;
    for (uint32_t c = 0; c < C; ++c) {
        const uint32_t nc = h_tot[c];
        if (nc == 0)
            continue;
        rc = launch_emit_instances(perm_of[c], off + (size_t)c * N, N, rects, tile_width, 0, nc, nullptr, tk_a, tv_a,
                                   stream);
        if (rc)
            return rc;
        int in_b = 0;
        rc = radix_sort_pairs(tk_a, tv_a, tk_b, tv_b, nc, nullptr, 0, tbits, sort_scr_b, &in_b, stream);
        if (rc)
            return rc;
        const int64_t cid_enc = (int64_t)c << (32 + tnb);
        k_make_isect_ids<<<div_up(nc, kIsThreads), kIsThreads, 0, stream>>>(in_b ? tk_b : tk_a, in_b ? tv_b : tv_a,
                                                                            depths, nc, cid_enc, ids + cam_off,
                                                                            flat + cam_off);
        LFS_LAUNCH_OK("k_make_isect_ids");
        cam_off += nc;
    }
    *isect_ids = ids;
    *flatten_ids = flat;
    return LFS_OK;
}

extern "C" int lfs_intersect_tile_packed(const float* means2d, const int32_t* radii, const float* depths,
                                         const int64_t* camera_ids, uint32_t nnz, uint32_t C, uint32_t tile_size,
                                         uint32_t tile_width, uint32_t tile_height, int sort, int32_t* tiles_per_gauss,
                                         lfs_alloc_fn alloc, void* alloc_ctx, int64_t** isect_ids, int32_t** flatten_ids,
                                         int64_t* n_isects_host, void* stream_) {
    using namespace lfs;
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(alloc && isect_ids && flatten_ids && n_isects_host, "intersect_tile_packed: null callback / out pointer");
    LFS_CHECK_ARG(tile_size > 0 && tile_width > 0 && tile_height > 0 && C > 0, "intersect_tile_packed: bad geometry");
    LFS_CHECK_ARG(tile_width < 65536 && tile_height < 65536, "intersect_tile_packed: more than 65535 tiles per axis");
    LFS_CHECK_ARG(nnz < (1u << 31), "intersect_tile_packed: nnz too large");
    const uint32_t n = nnz, n_tiles = tile_width * tile_height;
    const uint32_t tnb = ref_tile_n_bits(n_tiles), cnb = ref_tile_n_bits(C); // floor(log2) + 1, Intersect.cpp:46-47
    LFS_CHECK_ARG(tnb + cnb <= 32, "intersect_tile_packed: camera and tile ids need %u bits (32 available)", tnb + cnb);
    *isect_ids = nullptr;
    *flatten_ids = nullptr;
    *n_isects_host = 0;
    if (n == 0)
        return LFS_OK;
    LFS_CHECK_ARG(means2d && radii && depths && camera_ids && tiles_per_gauss, "intersect_tile_packed: null input");

    Carver ca(nullptr);
    ca.take<TileRect>(n);
    ca.take<uint32_t>(n), ca.take<uint32_t>(n), ca.take<uint32_t>(n), ca.take<uint32_t>(n), ca.take<uint32_t>(n);
    ca.take<uint32_t>(4);
    ca.take<char>(scan_scratch_bytes(n));
    ca.take<char>(radix_scratch_bytes(n));
    void* blob_a = alloc(alloc_ctx, LFS_TAG_SCRATCH, ca.total());
    if (!blob_a) {
        set_error("intersect_tile_packed: scratch allocation of %zu bytes failed", ca.total());
        return LFS_ERR_ALLOC;
    }
    Carver a(blob_a);
    TileRect* rects = a.take<TileRect>(n);
    uint32_t *dk_a = a.take<uint32_t>(n), *dk_b = a.take<uint32_t>(n), *pm_a = a.take<uint32_t>(n), *pm_b = a.take<uint32_t>(n);
    uint32_t* off = a.take<uint32_t>(n);
    uint32_t* totals = a.take<uint32_t>(4);
    void* scan_scr = a.take<char>(scan_scratch_bytes(n));
    void* sort_scr = a.take<char>(radix_scratch_bytes(n));

    int rc = launch_tile_count(means2d, radii, depths, n, (float)tile_size, tile_width, tile_height, tiles_per_gauss, rects,
                               sort ? dk_a : nullptr, sort ? pm_a : nullptr, stream);
    if (rc)
        return rc;
    const uint32_t* counts = reinterpret_cast<const uint32_t*>(tiles_per_gauss);
    const uint32_t* perm = nullptr;
    if (sort) { // level 1: all elements by depth bits (stable)
        int in_b = 0;
        rc = radix_sort_pairs(dk_a, pm_a, dk_b, pm_b, n, nullptr, 0, 32, sort_scr, &in_b, stream);
        if (rc)
            return rc;
        perm = in_b ? pm_b : pm_a;
    }
    rc = exclusive_scan_u32(counts, perm, off, totals, n, nullptr, scan_scr, stream);
    if (rc)
        return rc;
    uint32_t total = 0;
    LFS_CUDA_OK(cudaMemcpyAsync(&total, totals, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    LFS_CUDA_OK(cudaStreamSynchronize(stream));
    LFS_CHECK_ARG(total < (1u << 31), "intersect_tile_packed: n_isects %u does not fit int32 ids", total);
    *n_isects_host = total;
    if (total == 0)
        return LFS_OK;
    int64_t* ids = static_cast<int64_t*>(alloc(alloc_ctx, LFS_TAG_ISECT_IDS, sizeof(int64_t) * (size_t)total));
    int32_t* flat = static_cast<int32_t*>(alloc(alloc_ctx, LFS_TAG_FLATTEN_IDS, sizeof(int32_t) * (size_t)total));
    if (!ids || !flat) {
        set_error("intersect_tile_packed: output allocation failed (n_isects=%u)", total);
        return LFS_ERR_ALLOC;
    }
    const unsigned want = div_up(total, kIsThreads);
    const unsigned grid = want < (unsigned)(num_sms() * 16) ? want : (unsigned)(num_sms() * 16);
    if (!sort) {
        k_emit_unsorted_packed<<<grid, kIsThreads, 0, stream>>>(off, n, camera_ids, rects, depths, tile_width, tnb, total, ids,
                                                                flat);
        LFS_LAUNCH_OK("k_emit_unsorted_packed");
    } else { // level 2: (camera | tile) bits only, stable -> the order of the reference's 64-bit sort
        Carver cb(nullptr);
        cb.take<uint32_t>(total), cb.take<uint32_t>(total), cb.take<uint32_t>(total), cb.take<uint32_t>(total);
        cb.take<char>(radix_scratch_bytes(total));
        void* blob_b = alloc(alloc_ctx, LFS_TAG_SCRATCH, cb.total());
        if (!blob_b) {
            set_error("intersect_tile_packed: sort scratch allocation failed (n_isects=%u)", total);
            return LFS_ERR_ALLOC;
        }
        Carver b(blob_b);
        uint32_t *tk_a = b.take<uint32_t>(total), *tk_b = b.take<uint32_t>(total), *tv_a = b.take<uint32_t>(total),
                 *tv_b = b.take<uint32_t>(total);
        void* sort_scr_b = b.take<char>(radix_scratch_bytes(total));
        rc = launch_emit_instances(perm, off, n, rects, tile_width, 0, total, nullptr, tk_a, tv_a, stream);
        if (rc)
            return rc;
        k_add_camera_bits<<<div_up(total, kIsThreads), kIsThreads, 0, stream>>>(tk_a, tv_a, camera_ids, total, tnb);
        LFS_LAUNCH_OK("k_add_camera_bits");
        int in_b = 0;
        rc = radix_sort_pairs(tk_a, tv_a, tk_b, tv_b, total, nullptr, 0, (int)(tnb + cnb), sort_scr_b, &in_b, stream);
        if (rc)
            return rc;
        k_make_isect_ids<<<div_up(total, kIsThreads), kIsThreads, 0, stream>>>(in_b ? tk_b : tk_a, in_b ? tv_b : tv_a, depths,
                                                                               total, 0, ids, flat);
        LFS_LAUNCH_OK("k_make_isect_ids");
    }
    *isect_ids = ids;
    *flatten_ids = flat;
    return LFS_OK;
}

extern "C" int lfs_intersect_offset(const int64_t* isect_ids, int64_t n_isects, uint32_t C, uint32_t tile_width,
                                    uint32_t tile_height, int32_t* offsets, void* stream_) {
    using namespace lfs;
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(offsets != nullptr, "intersect_offset: offsets is null");
    const uint32_t n_tiles = tile_width * tile_height;
    const uint64_t total = (uint64_t)C * n_tiles;
    if (total == 0)
        return LFS_OK;
    if (n_isects <= 0) { // reference: offsets.fill_(0) (IntersectTile.cu:268-271)
        LFS_CUDA_OK(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * total, stream));
        return LFS_OK;
    }
    LFS_CHECK_ARG(isect_ids != nullptr, "intersect_offset: isect_ids is null");
    LFS_CHECK_ARG(n_isects < (1ll << 31) && total < (1ull << 31), "intersect_offset: sizes exceed int32");
    k_intersect_offset<<<div_up((uint64_t)n_isects, kIsThreads), kIsThreads, 0, stream>>>(
        isect_ids, (uint32_t)n_isects, n_tiles, ref_tile_n_bits(n_tiles), (uint32_t)total, offsets);
    LFS_LAUNCH_OK("k_intersect_offset");
    return LFS_OK;
}
