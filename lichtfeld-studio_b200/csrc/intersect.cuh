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

// Exact (conservative) tile culling for the from-world response of a PINHOLE camera.  A pixel can only receive
// alpha >= 1/255 from a Gaussian if power = -1/2 N/D >= -ln(255 opacity), i.e. Q = N - tau D <= 0 with
// tau = 2 ln(255 opacity): a conic region of the image plane (N, D are the quadratics of raster.cuh).  CullRec holds Q
// re-centred on its minimiser: Q(u,v) = a u^2 + 2 b u v + c v^2 - lim, (u,v) = pixel - (xc,yc).  When the region is not
// an ellipse (camera inside the tau-ellipsoid) lim = +inf and every tile of the AABB is kept.  tau carries a safety
// margin, so a culled tile provably holds no pixel centre with alpha >= 1/255: results are unchanged, instances drop.
struct alignas(16) CullRec {
    float a, b, c, xc, yc, lim, idet, ia; // idet = 1 / (a c - b^2), ia = 1 / a
};
static_assert(sizeof(CullRec) == 32, "CullRec must be 32 bytes");

// Tiles [first, last] of tile row `ty` (within [x0, x1)) whose box of pixel centres meets the region; last < first: none.
// The region cut by the row band is convex, so its u-extent [u_min, u_max] decides every tile of the row exactly.
// The count (k_preprocess_fwd) and the emission (k_emit_instances_cull) MUST agree tile for tile, so every operation is
// an explicit intrinsic with one fixed hardware meaning (round-to-nearest FMA/MUL/ADD, MUFU.RSQ, MUFU.RCP): nothing can
// be contracted or re-associated differently in the two kernels.  Accuracy is not critical (tau carries a margin).
__device__ __forceinline__ float cull_sqrt(const float x) { // deterministic ~1 ulp sqrt, 0 for x <= 0
    const float y = fmaxf(x, 1e-30f);
    return __fmul_rn(y, __frsqrt_rn(y));
}
__device__ __forceinline__ void cull_row_span(const CullRec& r, const uint32_t ty, const uint32_t x0, const uint32_t x1,
                                              int& first, int& last) {
    first = (int)x0, last = (int)x1 - 1;
    if (!(r.lim < 3.0e38f)) // not an ellipse: keep the whole AABB row
        return;
    const float va = __fsub_rn(__fadd_rn((float)(ty * kTile), 0.5f), r.yc), vb = __fadd_rn(va, (float)(kTile - 1));
    const float la = __fmul_rn(r.lim, r.a);
    const float vext = cull_sqrt(__fmul_rn(la, r.idet));
    if (va > vext || vb < -vext) {
        last = first - 1;
        return;
    }
    const float det = __fmaf_rn(r.a, r.c, -__fmul_rn(r.b, r.b));
    const float vlo = fmaxf(va, -vext), vhi = fminf(vb, vext);
    const float uext = cull_sqrt(__fmul_rn(__fmul_rn(r.lim, r.c), r.idet)); // half-width of the whole ellipse
    const float vL = __fdividef(__fmul_rn(r.b, uext), r.c);                   // v of the leftmost point (u = -uext)
    const float sl = cull_sqrt(__fmaf_rn(-det, __fmul_rn(vlo, vlo), la));
    const float sh = cull_sqrt(__fmaf_rn(-det, __fmul_rn(vhi, vhi), la));
    const float bl = __fmul_rn(-r.b, vlo), bh = __fmul_rn(-r.b, vhi);
    float umin = fminf(__fmul_rn(__fsub_rn(bl, sl), r.ia), __fmul_rn(__fsub_rn(bh, sh), r.ia));
    float umax = fmaxf(__fmul_rn(__fadd_rn(bl, sl), r.ia), __fmul_rn(__fadd_rn(bh, sh), r.ia));
    if (vlo <= vL && vL <= vhi)
        umin = -uext;
    if (vlo <= -vL && -vL <= vhi)
        umax = uext;
    // tile tx holds pixel centres u in [16 tx + 0.5 - xc, 16 tx + 15.5 - xc]; 0.02 px of slack for the approximations
    const float f = ceilf(__fmul_rn(__fsub_rn(__fadd_rn(umin, r.xc), 15.52f), 1.0f / (float)kTile));
    const float l = floorf(__fmul_rn(__fadd_rn(__fadd_rn(umax, r.xc), -0.48f), 1.0f / (float)kTile));
    first = max(first, (int)fmaxf(f, -1.0e6f));
    last = min(last, (int)fminf(l, 1.0e6f));
}

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
int launch_emit_instances_cull(const uint32_t* perm, const uint32_t* off, uint32_t n_gauss, const TileRect* rects,
                               const int32_t* counts, const CullRec* cull, uint32_t tile_w, uint32_t n_cap,
                               const uint32_t* n_dev, uint32_t* tile_keys, uint32_t* vals, cudaStream_t stream);
int launch_emit_instances(const uint32_t* perm, const uint32_t* off, uint32_t n_gauss, const TileRect* rects,
                          uint32_t tile_w, uint32_t id_offset, uint32_t n_cap, const uint32_t* n_dev,
                          uint32_t* tile_keys, uint32_t* vals, cudaStream_t stream);

// offsets[t] = first sorted instance with tile key >= t, t in [0, n_tiles]; offsets[n_tiles] = n_inst.
int launch_tile_offsets(const uint32_t* sorted_tile_keys, uint32_t n_cap, const uint32_t* n_dev, uint32_t n_tiles,
                        int32_t* offsets, cudaStream_t stream);

} // namespace lfs
