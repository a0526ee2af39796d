// lfs_b200 -- tile-intersection building blocks shared by the gsplat-surface op and the fused trainer.
//
// Pipeline (per camera), all device-side, no host sync:
//   rects[g], count[g]            tile rectangle (AABB rule of the reference) and its area per Gaussian
//   depth sort of the Gaussians   stable radix sort, key = fp32 depth bits (culled: 0xFFFFFFFF), value = g
//   off = exclusive_scan(count[perm])      write offset of every Gaussian in depth order
//   emit                          one thread per INSTANCE (binary search in off) -> (tile key, gaussian id)
//   tile sort                     stable radix sort over the tile bits only (13 bits @1080p -> 2 passes)
//   offsets                       per-tile [start, end) from adjacent key differences
// The result order (tile, depth, gaussian index) is identical to the reference's single 64-bit-key CUB
// sort (gsplat/IntersectTile.cu:95-108, :290-328) while moving ~6x fewer bytes through HBM:
// 24 B/Gaussian/pass x 3 passes + 16 B/instance/pass x 2 passes instead of 24 B/instance/pass x 6 passes.
#pragma once
#include "common.cuh"
#include "sort_scan.cuh"

namespace lfs {

struct TileRect {
    unsigned short x0, y0, x1, y1;
};
static_assert(sizeof(TileRect) == 8, "TileRect must be 8 bytes");

// bits needed to encode a tile id < n_tiles
static inline int tile_key_bits(uint32_t n_tiles) {
    int b = 1;
    while ((1u << b) < n_tiles && b < 31)
        ++b;
    return b;
}
// the reference's key layout constant (gsplat/IntersectTile.cu:150): floor(log2(n)) + 1
static inline uint32_t ref_tile_n_bits(uint32_t n_tiles) {
    uint32_t b = 0;
    while ((n_tiles >> b) > 1)
        ++b;
    return b + 1;
}

// rects + counts (+ depth keys / identity values for the depth sort) from projected means2d / radii.
int launch_tile_count(const float* means2d, const int32_t* radii, const float* depths, uint32_t n, float tile_size,
                      uint32_t tile_w, uint32_t tile_h, int32_t* tiles_per_gauss, TileRect* rects,
                      uint32_t* depth_keys /* nullable */, uint32_t* ident /* nullable */, cudaStream_t stream);

// instance emission; perm == nullptr means identity order. n_inst comes from n_dev (clamped to n_cap) or n_cap.
int launch_emit_instances(const uint32_t* perm, const uint32_t* off, uint32_t n_gauss, const TileRect* rects,
                          uint32_t tile_w, uint32_t id_offset, uint32_t n_cap, const uint32_t* n_dev,
                          uint32_t* tile_keys, uint32_t* vals, cudaStream_t stream);

// offsets[t] = first sorted instance with tile key >= t, t in [0, n_tiles]; offsets[n_tiles] = n_inst.
int launch_tile_offsets(const uint32_t* sorted_tile_keys, uint32_t n_cap, const uint32_t* n_dev, uint32_t n_tiles,
                        int32_t* offsets, cudaStream_t stream);

} // namespace lfs
