/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * CPU oracle for the 3DGS training hot path of MrNeRF/LichtFeld-Studio: a plain-C restatement of the
 * reference algorithms (reference file:line cited per function).  Built by oracle/Makefile into
 * oracle/liblfs_oracle.so.  Floating-point stages exist in a float (orc32_) and a double (orc64_)
 * variant (lfs_oracle_impl.h); integer stages (tile intersection, keys, offsets) exist once and follow
 * the CUDA semantics bit for bit (float arithmetic, saturating float->uint conversion).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define PFX orc32_
#include "lfs_oracle_impl.h"
#include "lfs_oracle_fastgs_impl.h"
#include "lfs_oracle_legacy2d_impl.h"
#undef REAL
#undef PFX

#define REAL double
#define PFX orc64_
#include "lfs_oracle_impl.h"
#include "lfs_oracle_fastgs_impl.h"
#include "lfs_oracle_legacy2d_impl.h"
#undef REAL
#undef PFX

/* ------------------------------------------------------------------------------------------------
 * (a5) tile intersection.  gsplat/IntersectTile.cu:54-113 (kernel), gsplat/Intersect.cpp:41-121 (host
 * two-pass + cumsum + sort).  cvt.rzi.u32.f32 saturates: negative / NaN -> 0.
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t f2u_sat(float x) {
    if (!(x > 0.0f))
        return 0u;
    if (x >= 4294967296.0f)
        return 0xFFFFFFFFu;
    return (uint32_t)x;
}
static inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

uint32_t orc_tile_n_bits(uint32_t n_tiles) { return (uint32_t)floor(log2((double)n_tiles)) + 1; }

static inline int tile_rect(const float* means2d, const int32_t* radii, size_t idx, uint32_t tile_size,
                            uint32_t tile_width, uint32_t tile_height, uint32_t* x0, uint32_t* y0, uint32_t* x1,
                            uint32_t* y1) {
    const float radius_x = (float)radii[idx * 2];
    const float radius_y = (float)radii[idx * 2 + 1];
    if (radius_x <= 0 || radius_y <= 0)
        return 0;
    const float tile_radius_x = radius_x / (float)tile_size;
    const float tile_radius_y = radius_y / (float)tile_size;
    const float tile_x = means2d[idx * 2] / (float)tile_size;
    const float tile_y = means2d[idx * 2 + 1] / (float)tile_size;
    *x0 = umin(f2u_sat(floorf(tile_x - tile_radius_x)), tile_width);
    *y0 = umin(f2u_sat(floorf(tile_y - tile_radius_y)), tile_height);
    *x1 = umin(f2u_sat(ceilf(tile_x + tile_radius_x)), tile_width);
    *y1 = umin(f2u_sat(ceilf(tile_y + tile_radius_y)), tile_height);
    return 1;
}

/* first pass: tiles_per_gauss[C*N]; returns n_isects */
int64_t orc_intersect_count(int C, int N, const float* means2d, const int32_t* radii, uint32_t tile_size,
                            uint32_t tile_width, uint32_t tile_height, int32_t* tiles_per_gauss) {
    int64_t total = 0;
    for (size_t idx = 0; idx < (size_t)C * N; ++idx) {
        uint32_t x0, y0, x1, y1;
        int32_t n = 0;
        if (tile_rect(means2d, radii, idx, tile_size, tile_width, tile_height, &x0, &y0, &x1, &y1))
            n = (int32_t)((y1 - y0) * (x1 - x0));
        tiles_per_gauss[idx] = n;
        total += n;
    }
    return total;
}

static void stable_sort_pairs_u64(uint64_t* keys, int32_t* vals, int64_t n, int end_bit) {
    uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    int32_t* v2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int shift = 0; shift < end_bit; shift += 8) {
        int64_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; ++i)
            cnt[((keys[i] >> shift) & 0xFF) + 1]++;
        for (int d = 0; d < 256; ++d)
            cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; ++i) {
            int64_t p = cnt[(keys[i] >> shift) & 0xFF]++;
            k2[p] = keys[i];
            v2[p] = vals[i];
        }
        memcpy(keys, k2, sizeof(uint64_t) * (size_t)n);
        memcpy(vals, v2, sizeof(int32_t) * (size_t)n);
    }
    free(k2);
    free(v2);
}

/* second pass (+ optional stable sort over 32 + tile_n_bits + cam_n_bits bits, as
 * cub::DeviceRadixSort::SortPairs does in gsplat/IntersectTile.cu:290-328).
 * isect_ids / flatten_ids sized n_isects (from orc_intersect_count). */
void orc_intersect_emit(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                        uint32_t tile_size, uint32_t tile_width, uint32_t tile_height, int sort, int64_t* isect_ids,
                        int32_t* flatten_ids) {
    const uint32_t n_tiles = tile_width * tile_height;
    const uint32_t tile_n_bits = orc_tile_n_bits(n_tiles);
    const uint32_t cam_n_bits = (uint32_t)floor(log2((double)C)) + 1;
    int64_t cur = 0;
    for (size_t idx = 0; idx < (size_t)C * N; ++idx) {
        uint32_t x0, y0, x1, y1;
        if (!tile_rect(means2d, radii, idx, tile_size, tile_width, tile_height, &x0, &y0, &x1, &y1))
            continue;
        const int64_t cid = (int64_t)(idx / (size_t)N);
        const int64_t cid_enc = cid << (32 + tile_n_bits);
        uint32_t dbits;
        memcpy(&dbits, &depths[idx], 4);
        const int64_t depth_id_enc = (int64_t)dbits;
        for (uint32_t i = y0; i < y1; ++i)
            for (uint32_t j = x0; j < x1; ++j) {
                const int64_t tile_id = (int64_t)i * tile_width + j;
                isect_ids[cur] = cid_enc | (tile_id << 32) | depth_id_enc;
                flatten_ids[cur] = (int32_t)idx;
                ++cur;
            }
    }
    if (sort)
        stable_sort_pairs_u64((uint64_t*)isect_ids, flatten_ids, cur, (int)(32 + tile_n_bits + cam_n_bits));
}

/* (a6) gsplat/IntersectTile.cu:206-252: offsets[c*n_tiles + t] = first sorted index with (cid,tile) >= (c,t);
 * n_isects == 0 -> all zeros (IntersectTile.cu:268-271). */
void orc_intersect_offset(int64_t n_isects, const int64_t* isect_ids, int C, uint32_t tile_width,
                          uint32_t tile_height, int32_t* offsets) {
    const uint32_t n_tiles = tile_width * tile_height;
    const uint32_t tile_n_bits = orc_tile_n_bits(n_tiles);
    const int64_t total = (int64_t)C * n_tiles;
    int64_t next = 0; /* next offsets slot to fill */
    for (int64_t i = 0; i < n_isects; ++i) {
        const int64_t hi = isect_ids[i] >> 32;
        const int64_t cid = hi >> tile_n_bits;
        const int64_t tid = hi & (((int64_t)1 << tile_n_bits) - 1);
        const int64_t id = cid * n_tiles + tid;
        while (next <= id && next < total)
            offsets[next++] = (int32_t)i;
    }
    while (next < total)
        offsets[next++] = (int32_t)n_isects;
}
