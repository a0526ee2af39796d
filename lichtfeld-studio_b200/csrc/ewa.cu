// lfs_b200 -- fastgs (EWA) rasterizer surface: lfs_fastgs_forward / lfs_fastgs_backward, drop-ins for
// fast_gs::rasterization::forward / backward (reference fastgs/rasterization/src/forward.cu:15-199,
// src/backward.cu:14-116; kernels include/kernels_forward.cuh, include/kernels_backward.cuh, include/kernel_utils.cuh).
//
// Same pipeline shape as the from-world path and the same blend kernels (raster.cu, EWA instantiation): the 2-D conic
// response is a plain quadratic in the tile-local pixel offset, so a record is 6 coefficients + opacity + colour.
//   preprocess (1 thread / primitive, exact tile count) -> depth radix sort of all primitives (invisible = 0xFFFFFFFF,
//   ties by index: deterministic, unlike the reference's atomic compaction) -> scan -> one-thread-per-primitive instance
//   emission with the exact ellipse/tile test -> tile-bits radix sort -> offsets / buckets -> blend.
#include "intersect.cuh"
#include "raster.cuh"
#include "sort_scan.cuh"

namespace lfs {

constexpr float kFgDilation = 0.3f;          // rasterization_config.h:16
constexpr float kFgMinAlphaRcp = 255.0f;     // :17

// kernel_utils.cuh:108-143 (mean already shifted by -0.5, :152): does the primitive reach alpha >= 1/255 in the tile?
// The reference compiles with --use_fast_math, so its two divisions are div.approx: __fdividef here.
__device__ __forceinline__ bool fg_tile_test(const float mx, const float my, const float ca, const float cb, const float cc,
                                             const uint32_t tile_x, const uint32_t tile_y, const float power_threshold) {
    const float rminx = (float)(tile_x * kTile), rminy = (float)(tile_y * kTile);
    const float rmaxx = (float)((tile_x + 1) * kTile - 1), rmaxy = (float)((tile_y + 1) * kTile - 1);
    const float x_min_diff = rminx - mx;
    const float x_left = x_min_diff > 0.f ? 1.f : 0.f;
    const float not_in_x = x_left + (mx > rmaxx ? 1.f : 0.f);
    const float y_min_diff = rminy - my;
    const float y_above = y_min_diff > 0.f ? 1.f : 0.f;
    const float not_in_y = y_above + (my > rmaxy ? 1.f : 0.f);
    if (not_in_y + not_in_x == 0.f)
        return true;
    const float ccx = rmaxx + x_left * (rminx - rmaxx), ccy = rmaxy + y_above * (rminy - rmaxy);
    const float diffx = mx - ccx, diffy = my - ccy;
    const float dx = copysignf((float)(kTile - 1), x_min_diff), dy = copysignf((float)(kTile - 1), y_min_diff);
    const float tx = not_in_y * __saturatef(__fdividef(dx * ca * diffx + dx * cb * diffy, dx * ca * dx));
    const float ty = not_in_x * __saturatef(__fdividef(dy * cb * diffx + dy * cc * diffy, dy * cc * dy));
    const float ex = mx - (ccx + tx * dx), ey = my - (ccy + ty * dy);
    const float max_power = 0.5f * (ca * ex * ex + cc * ey * ey) + cb * ex * ey;
    return max_power <= power_threshold;
}
// Rectangles of more than 64 tiles (rare) are tested twice -- counted in k_fg_preprocess, re-walked in k_fg_emit -- and both
// must take the same decision for the same inputs: ONE compiled body.  Smaller rectangles are tested once and remembered
// as a 64-bit mask.
__device__ __noinline__ bool fg_will_contribute(const float mx, const float my, const float ca, const float cb,
                                                   const float cc, const uint32_t tile_x, const uint32_t tile_y,
                                                   const float power_threshold) {
    return fg_tile_test(mx, my, ca, cb, cc, tile_x, tile_y, power_threshold);
}

struct FgGeom { // shared by forward (kernels_forward.cuh:60-147) and backward (kernels_backward.cuh:56-126)
    float depth, x, y, tx, ty, j11, j13, j22, j23;
    float rot[9], rs[9], cov3d[6], variance[3];
    float qn2, q[4], q2[9];
    f3 jw1, jw2, jwc1, jwc2;
    float a, b, c;
};

__device__ __forceinline__ void fg_geometry(const f3 mean, const f3 raw_scale, const float4 rq, const float4 r1,
                                            const float4 r2, const float4 r3, const float w, const float h,
                                            const float fx, const float fy, const float cx, const float cy, FgGeom& G) {
    G.depth = r3.x * mean.x + r3.y * mean.y + r3.z * mean.z + r3.w;
    G.variance[0] = expf(2.0f * raw_scale.x), G.variance[1] = expf(2.0f * raw_scale.y), G.variance[2] = expf(2.0f * raw_scale.z);
    const float qr = rq.x, qx = rq.y, qy = rq.z, qz = rq.w;
    const float n2 = qr * qr + qx * qx + qy * qy + qz * qz;
    G.qn2 = n2;
    G.q[0] = qr, G.q[1] = qx, G.q[2] = qy, G.q[3] = qz;
    const float s2 = 2.0f / n2;
    const float qxx = s2 * qx * qx, qyy = s2 * qy * qy, qzz = s2 * qz * qz;
    const float qxy = s2 * qx * qy, qxz = s2 * qx * qz, qyz = s2 * qy * qz;
    const float qrx = s2 * qr * qx, qry = s2 * qr * qy, qrz = s2 * qr * qz;
    G.q2[0] = qxx, G.q2[1] = qyy, G.q2[2] = qzz, G.q2[3] = qxy, G.q2[4] = qxz, G.q2[5] = qyz, G.q2[6] = qrx, G.q2[7] = qry,
    G.q2[8] = qrz;
    const float R[9] = {1.0f - (qyy + qzz), qxy - qrz, qry + qxz, qrz + qxy, 1.0f - (qxx + qzz), qyz - qrx,
                        qxz - qry, qrx + qyz, 1.0f - (qxx + qyy)};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            G.rot[3 * i + j] = R[3 * i + j];
            G.rs[3 * i + j] = R[3 * i + j] * G.variance[j];
        }
    int t = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j)
            G.cov3d[t++] = G.rs[3 * i] * R[3 * j] + G.rs[3 * i + 1] * R[3 * j + 1] + G.rs[3 * i + 2] * R[3 * j + 2];
    G.x = (r1.x * mean.x + r1.y * mean.y + r1.z * mean.z + r1.w) / G.depth;
    G.y = (r2.x * mean.x + r2.y * mean.y + r2.z * mean.z + r2.w) / G.depth;
    const float cl = (-0.15f * w - cx) / fx, cr = (1.15f * w - cx) / fx;
    const float ct = (-0.15f * h - cy) / fy, cb = (1.15f * h - cy) / fy;
    G.tx = fminf(fmaxf(G.x, cl), cr), G.ty = fminf(fmaxf(G.y, ct), cb);
    G.j11 = fx / G.depth, G.j13 = -G.j11 * G.tx, G.j22 = fy / G.depth, G.j23 = -G.j22 * G.ty;
    G.jw1 = mk3(G.j11 * r1.x + G.j13 * r3.x, G.j11 * r1.y + G.j13 * r3.y, G.j11 * r1.z + G.j13 * r3.z);
    G.jw2 = mk3(G.j22 * r2.x + G.j23 * r3.x, G.j22 * r2.y + G.j23 * r3.y, G.j22 * r2.z + G.j23 * r3.z);
    const float* S = G.cov3d; // m11 m12 m13 m22 m23 m33
    G.jwc1 = mk3(G.jw1.x * S[0] + G.jw1.y * S[1] + G.jw1.z * S[2], G.jw1.x * S[1] + G.jw1.y * S[3] + G.jw1.z * S[4],
                 G.jw1.x * S[2] + G.jw1.y * S[4] + G.jw1.z * S[5]);
    G.jwc2 = mk3(G.jw2.x * S[0] + G.jw2.y * S[1] + G.jw2.z * S[2], G.jw2.x * S[1] + G.jw2.y * S[3] + G.jw2.z * S[4],
                 G.jw2.x * S[2] + G.jw2.y * S[4] + G.jw2.z * S[5]);
    G.a = dot(G.jwc1, G.jw1) + kFgDilation;
    G.b = dot(G.jwc1, G.jw2);
    G.c = dot(G.jwc2, G.jw2) + kFgDilation;
}

// SH basis values of kernel_utils.cuh:15-39 (bases 1..15 of sh_coefficients_rest; band 0 handled by the caller)
__device__ __forceinline__ int fg_sh_basis(const int active, const float x, const float y, const float z, float b[15]) {
    if (active <= 1)
        return 0;
    b[0] = -0.48860251190291987f * y, b[1] = 0.48860251190291987f * z, b[2] = -0.48860251190291987f * x;
    if (active <= 4)
        return 3;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    b[3] = 1.0925484305920792f * xy, b[4] = -1.0925484305920792f * yz;
    b[5] = 0.94617469575755997f * zz - 0.31539156525251999f;
    b[6] = -1.0925484305920792f * xz, b[7] = 0.54627421529603959f * xx - 0.54627421529603959f * yy;
    if (active <= 9)
        return 8;
    b[8] = 0.59004358992664352f * y * (-3.0f * xx + yy), b[9] = 2.8906114426405538f * xy * z;
    b[10] = 0.45704579946446572f * y * (1.0f - 5.0f * zz), b[11] = 0.3731763325901154f * z * (5.0f * zz - 3.0f);
    b[12] = 0.45704579946446572f * x * (1.0f - 5.0f * zz), b[13] = 1.4453057213202769f * z * (xx - yy);
    b[14] = 0.59004358992664352f * x * (-xx + 3.0f * yy);
    return 15;
}

// Block-cooperative copy of `rows` consecutive rows (row_floats floats each, first row `row0`) between an AoS tensor and
// shared memory, 128-bit where the range allows (every block starts at a multiple of 128 rows).
__device__ __forceinline__ void fg_stage_rows(const float* __restrict__ src, float* __restrict__ smem, const uint32_t row0,
                                              const uint32_t n_rows_total, const uint32_t row_floats) {
    if (row0 >= n_rows_total)
        return;
    const uint32_t rows = min(128u, n_rows_total - row0);
    const uint32_t nfl = rows * row_floats;
    const float* g = src + (size_t)row0 * row_floats;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const uint32_t n4 = nfl >> 2;
        for (uint32_t k = threadIdx.x; k < n4; k += blockDim.x)
            reinterpret_cast<float4*>(smem)[k] = __ldg(reinterpret_cast<const float4*>(g) + k);
        for (uint32_t k = (n4 << 2) + threadIdx.x; k < nfl; k += blockDim.x)
            smem[k] = __ldg(g + k);
    } else {
        for (uint32_t k = threadIdx.x; k < nfl; k += blockDim.x)
            smem[k] = __ldg(g + k);
    }
}
__device__ __forceinline__ void fg_unstage_rows(float* __restrict__ dst, const float* __restrict__ smem, const uint32_t row0,
                                                const uint32_t n_rows_total, const uint32_t row_floats) {
    if (row0 >= n_rows_total)
        return;
    const uint32_t rows = min(128u, n_rows_total - row0);
    const uint32_t nfl = rows * row_floats;
    float* g = dst + (size_t)row0 * row_floats;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const uint32_t n4 = nfl >> 2;
        for (uint32_t k = threadIdx.x; k < n4; k += blockDim.x)
            reinterpret_cast<float4*>(g)[k] = reinterpret_cast<const float4*>(smem)[k];
        for (uint32_t k = (n4 << 2) + threadIdx.x; k < nfl; k += blockDim.x)
            g[k] = smem[k];
    } else {
        for (uint32_t k = threadIdx.x; k < nfl; k += blockDim.x)
            g[k] = smem[k];
    }
}

// ------------------------------------------------------------------------------------------------------
// preprocess (kernels_forward.cuh:18-205)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    k_fg_preprocess(const float* __restrict__ means, const float* __restrict__ raw_scales,
                    const float4* __restrict__ raw_rotations, const float* __restrict__ raw_opacities,
                    const float* __restrict__ sh0, const float* __restrict__ sh_rest, const float4* __restrict__ w2c,
                    const float* __restrict__ cam_position, const uint32_t N, const uint32_t grid_w,
                    const uint32_t grid_h, const int active, const int total_rest, const float w, const float h,
                    const float fx, const float fy, const float cx, const float cy, const float near_, const float far_,
                    GaussRec* __restrict__ gauss, TileRect* __restrict__ rects, int32_t* __restrict__ counts,
                    unsigned long long* __restrict__ masks, uint32_t* __restrict__ depth_keys,
                    uint32_t* __restrict__ ident, uint32_t* __restrict__ n_visible) {
    // the higher SH bands of the block's 128 primitives are one contiguous range of the AoS tensor: copied with coalesced
    // 128-bit loads (a thread reading its own 180-B row touches 32 sectors per load instruction of the warp)
    __shared__ __align__(16) float s_sh[128 * 45];
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (active > 1)
        fg_stage_rows(sh_rest, s_sh, blockIdx.x * 128u, N, 3u * (uint32_t)total_rest);
    __syncthreads();
    if (i >= N)
        return;
    ident[i] = i;
    counts[i] = 0;
    rects[i] = TileRect{0, 0, 0, 0};
    depth_keys[i] = 0xFFFFFFFFu;
    const f3 mean = mk3(means[3 * (size_t)i], means[3 * (size_t)i + 1], means[3 * (size_t)i + 2]);
    const float4 r1 = w2c[0], r2 = w2c[1], r3 = w2c[2];
    FgGeom G;
    fg_geometry(mean, mk3(raw_scales[3 * (size_t)i], raw_scales[3 * (size_t)i + 1], raw_scales[3 * (size_t)i + 2]),
                raw_rotations[i], r1, r2, r3, w, h, fx, fy, cx, cy, G);
    if (G.depth < near_ || G.depth > far_) // :61-62
        return;
    const float opacity = 1.0f / (1.0f + expf(-raw_opacities[i]));
    if (opacity < 1.0f / kFgMinAlphaRcp || G.qn2 < 1e-8f) // :75, :84
        return;
    const float det = G.a * G.c - G.b * G.b;
    if (!(det >= 1e-8f)) // :146 (also rejects NaN)
        return;
    const float ca = G.c / det, cb = -G.b / det, cc = G.a / det;
    const float mx = G.x * fx + cx, my = G.y * fy + cy;
    const float pt = logf(opacity * kFgMinAlphaRcp);
    const float ptf = sqrtf(2.0f * pt);
    const float ex = fmaxf(ptf * sqrtf(G.a) - 0.5f, 0.0f), ey = fmaxf(ptf * sqrtf(G.c) - 0.5f, 0.0f);
    const uint32_t x0 = min(grid_w, (uint32_t)max(0, __float2int_rd((mx - ex) / (float)kTile)));
    const uint32_t x1 = min(grid_w, (uint32_t)max(0, __float2int_ru((mx + ex) / (float)kTile)));
    const uint32_t y0 = min(grid_h, (uint32_t)max(0, __float2int_rd((my - ey) / (float)kTile)));
    const uint32_t y1 = min(grid_h, (uint32_t)max(0, __float2int_ru((my + ey) / (float)kTile)));
    if (x1 <= x0 || y1 <= y0)
        return;
    uint32_t nt = 0;
    unsigned long long mask = 0ull;
    if ((x1 - x0) * (y1 - y0) <= 64u) { // row-major bit per tile of the rectangle; k_fg_emit walks the bits
        unsigned long long bit = 1ull;
        for (uint32_t ty = y0; ty < y1; ++ty)
            for (uint32_t tx = x0; tx < x1; ++tx, bit <<= 1)
                if (fg_tile_test(mx - 0.5f, my - 0.5f, ca, cb, cc, tx, ty, pt))
                    mask |= bit;
        nt = (uint32_t)__popcll(mask);
    } else {
        for (uint32_t ty = y0; ty < y1; ++ty)
            for (uint32_t tx = x0; tx < x1; ++tx)
                nt += fg_will_contribute(mx - 0.5f, my - 0.5f, ca, cb, cc, tx, ty, pt) ? 1u : 0u;
    }
    if (nt == 0)
        return;
    masks[i] = mask;
    // colour, kernel_utils.cuh:15-39
    f3 col = mk3(0.5f + 0.28209479177387814f * sh0[3 * (size_t)i], 0.5f + 0.28209479177387814f * sh0[3 * (size_t)i + 1],
                 0.5f + 0.28209479177387814f * sh0[3 * (size_t)i + 2]);
    if (active > 1) {
        const f3 d = mk3(mean.x - cam_position[0], mean.y - cam_position[1], mean.z - cam_position[2]);
        const float inv = rsqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        float b[15];
        const int nb = fg_sh_basis(active, d.x * inv, d.y * inv, d.z * inv, b);
        const float* c = s_sh + 3 * (size_t)threadIdx.x * total_rest; // row stride 3 * total_rest floats: odd (9, 45), hence
                                                                     // conflict-free, for SH degree 1 and 3
#pragma unroll
        for (int j = 0; j < 15; ++j)
            if (j < nb) {
                col.x = fmaf(b[j], c[3 * j], col.x), col.y = fmaf(b[j], c[3 * j + 1], col.y);
                col.z = fmaf(b[j], c[3 * j + 2], col.z);
            }
    }
    counts[i] = (int32_t)nt;
    rects[i] = TileRect{(unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1};
    depth_keys[i] = __float_as_uint(G.depth);
    float4* og = reinterpret_cast<float4*>(gauss + i);
    og[0] = make_float4(mx, my, ca, cb);
    og[1] = make_float4(cc, opacity, col.x, col.y);
    og[2] = make_float4(col.z, pt, 0.f, 0.f);
    atomicAdd(n_visible, 1u);
}

__global__ void __launch_bounds__(128)
    k_fg_preprocess_coop(const float* __restrict__ means, const float* __restrict__ raw_scales,
                    const float4* __restrict__ raw_rotations, const float* __restrict__ raw_opacities,
                    const float* __restrict__ sh0, const float* __restrict__ sh_rest, const float4* __restrict__ w2c,
                    const float* __restrict__ cam_position, const uint32_t N, const uint32_t grid_w,
                    const uint32_t grid_h, const int active, const int total_rest, const float w, const float h,
                    const float fx, const float fy, const float cx, const float cy, const float near_, const float far_,
                    GaussRec* __restrict__ gauss, TileRect* __restrict__ rects, int32_t* __restrict__ counts,
                    unsigned long long* __restrict__ masks, uint32_t* __restrict__ depth_keys,
                    uint32_t* __restrict__ ident, uint32_t* __restrict__ n_visible) {
    // A/B variant (fg_variant = 1, measured slower, see k_fg_emit_coop).  Same results as k_fg_preprocess; the exact tile
    // tests of a warp's 32 primitives are spread evenly over its lanes (k_fg_preprocess: every lane walks its own
    // rectangle, 11.7 of 32 lanes active per instruction, 68 % issue-bound).
    __shared__ __align__(16) float s_sh[128 * 45];
    __shared__ float s_par[4][6][32];          // mx - 0.5, my - 0.5, conic a b c, power threshold
    __shared__ uint32_t s_geo[4][32];          // x0 | y0 << 12 | w << 24
    __shared__ uint32_t s_rcp[4][32];          // ceil(65536 / w)
    __shared__ uint32_t s_pre[4][33];
    __shared__ unsigned long long s_mask[4][32];
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (active > 1)
        fg_stage_rows(sh_rest, s_sh, blockIdx.x * 128u, N, 3u * (uint32_t)total_rest);
    bool alive = i < N;
    f3 mean = mk3(0.f, 0.f, 0.f);
    float depth = 0.f, opacity = 0.f, ca = 0.f, cb = 0.f, cc = 0.f, mx = 0.f, my = 0.f, pt = 0.f;
    uint32_t x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    if (alive) {
        ident[i] = i;
        counts[i] = 0;
        rects[i] = TileRect{0, 0, 0, 0};
        depth_keys[i] = 0xFFFFFFFFu;
        mean = mk3(means[3 * (size_t)i], means[3 * (size_t)i + 1], means[3 * (size_t)i + 2]);
        const float4 r1 = w2c[0], r2 = w2c[1], r3 = w2c[2];
        FgGeom G;
        fg_geometry(mean, mk3(raw_scales[3 * (size_t)i], raw_scales[3 * (size_t)i + 1], raw_scales[3 * (size_t)i + 2]),
                    raw_rotations[i], r1, r2, r3, w, h, fx, fy, cx, cy, G);
        depth = G.depth;
        opacity = 1.0f / (1.0f + expf(-raw_opacities[i]));
        const float det = G.a * G.c - G.b * G.b;
        alive = !(G.depth < near_ || G.depth > far_) && !(opacity < 1.0f / kFgMinAlphaRcp || G.qn2 < 1e-8f) &&
                (det >= 1e-8f); // :61-62, :75, :84, :146 (also rejects NaN)
        if (alive) {
            ca = G.c / det, cb = -G.b / det, cc = G.a / det;
            mx = G.x * fx + cx, my = G.y * fy + cy;
            pt = logf(opacity * kFgMinAlphaRcp);
            const float ptf = sqrtf(2.0f * pt);
            const float ex = fmaxf(ptf * sqrtf(G.a) - 0.5f, 0.0f), ey = fmaxf(ptf * sqrtf(G.c) - 0.5f, 0.0f);
            x0 = min(grid_w, (uint32_t)max(0, __float2int_rd((mx - ex) / (float)kTile)));
            x1 = min(grid_w, (uint32_t)max(0, __float2int_ru((mx + ex) / (float)kTile)));
            y0 = min(grid_h, (uint32_t)max(0, __float2int_rd((my - ey) / (float)kTile)));
            y1 = min(grid_h, (uint32_t)max(0, __float2int_ru((my + ey) / (float)kTile)));
            alive = x1 > x0 && y1 > y0;
        }
    }
    const uint32_t rw = alive ? x1 - x0 : 0u, area = alive ? rw * (y1 - y0) : 0u;
    const uint32_t a_coop = area <= 64u ? area : 0u;
    uint32_t incl = a_coop;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o)
            incl += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    s_pre[warp][lane] = incl - a_coop;
    if (lane == 31)
        s_pre[warp][32] = total;
    s_par[warp][0][lane] = mx - 0.5f, s_par[warp][1][lane] = my - 0.5f, s_par[warp][2][lane] = ca;
    s_par[warp][3][lane] = cb, s_par[warp][4][lane] = cc, s_par[warp][5][lane] = pt;
    s_geo[warp][lane] = x0 | (y0 << 12) | (rw << 24);
    s_rcp[warp][lane] = rw ? (65536u + rw - 1u) / rw : 0u;
    s_mask[warp][lane] = 0ull;
    __syncwarp();
    for (uint32_t t = lane; t < total; t += 32u) {
        uint32_t lo = 0, hi = 32;
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_pre[warp][mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        const uint32_t k = t - s_pre[warp][lo];
        const uint32_t geo = s_geo[warp][lo];
        const uint32_t row = (k * s_rcp[warp][lo]) >> 16, col = k - row * (geo >> 24);
        if (fg_tile_test(s_par[warp][0][lo], s_par[warp][1][lo], s_par[warp][2][lo], s_par[warp][3][lo], s_par[warp][4][lo],
                         (geo & 0xfffu) + col, ((geo >> 12) & 0xfffu) + row, s_par[warp][5][lo]))
            atomicOr(&s_mask[warp][lo], 1ull << k);
    }
    __syncwarp();
    __syncthreads(); // the staged SH rows
    if (!alive)
        return;
    uint32_t nt = 0;
    unsigned long long mask = 0ull;
    if (area <= 64u) {
        mask = s_mask[warp][lane];
        nt = (uint32_t)__popcll(mask);
    } else {
        for (uint32_t ty = y0; ty < y1; ++ty)
            for (uint32_t tx = x0; tx < x1; ++tx)
                nt += fg_will_contribute(mx - 0.5f, my - 0.5f, ca, cb, cc, tx, ty, pt) ? 1u : 0u;
    }
    if (nt == 0)
        return;
    masks[i] = mask;
    // colour, kernel_utils.cuh:15-39
    f3 col = mk3(0.5f + 0.28209479177387814f * sh0[3 * (size_t)i], 0.5f + 0.28209479177387814f * sh0[3 * (size_t)i + 1],
                 0.5f + 0.28209479177387814f * sh0[3 * (size_t)i + 2]);
    if (active > 1) {
        const f3 d = mk3(mean.x - cam_position[0], mean.y - cam_position[1], mean.z - cam_position[2]);
        const float inv = rsqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        float b[15];
        const int nb = fg_sh_basis(active, d.x * inv, d.y * inv, d.z * inv, b);
        const float* c = s_sh + 3 * (size_t)threadIdx.x * total_rest;
#pragma unroll
        for (int j = 0; j < 15; ++j)
            if (j < nb) {
                col.x = fmaf(b[j], c[3 * j], col.x), col.y = fmaf(b[j], c[3 * j + 1], col.y);
                col.z = fmaf(b[j], c[3 * j + 2], col.z);
            }
    }
    counts[i] = (int32_t)nt;
    rects[i] = TileRect{(unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1};
    depth_keys[i] = __float_as_uint(depth);
    float4* og = reinterpret_cast<float4*>(gauss + i);
    og[0] = make_float4(mx, my, ca, cb);
    og[1] = make_float4(cc, opacity, col.x, col.y);
    og[2] = make_float4(col.z, pt, 0.f, 0.f);
    atomicAdd(n_visible, 1u);
}

// Instance emission with the exact tile test (kernels_forward.cuh:221-320): one thread per slot of the depth order
// re-evaluates the test over its rectangle (same function, same inputs as the count in k_fg_preprocess) and writes its
// run; see k_emit_instances for why this beats the warp-cooperative walk.
__global__ void __launch_bounds__(256)
    k_fg_emit(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ off, const uint32_t n_gauss,
              const TileRect* __restrict__ rects, const int32_t* __restrict__ counts,
              const unsigned long long* __restrict__ masks, const GaussRec* __restrict__ gauss, const uint32_t tile_w,
              const uint32_t n_cap, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ vals) {
    for (uint32_t slot = blockIdx.x * 256 + threadIdx.x; slot < n_gauss; slot += gridDim.x * 256) {
        const uint32_t g = __ldg(perm + slot);
        const int32_t cnt = counts[g];
        if (cnt <= 0)
            continue;
        const TileRect r = rects[g];
        uint32_t pos = __ldg(off + slot);
        const uint32_t end = pos + (uint32_t)cnt; // never write past this Gaussian's share
        if ((uint32_t)(r.x1 - r.x0) * (uint32_t)(r.y1 - r.y0) <= 64u) {
            unsigned long long mask = masks[g];
            const uint32_t w = (uint32_t)(r.x1 - r.x0);
            const unsigned long long row_mask = w == 64u ? ~0ull : ((1ull << w) - 1ull);
            for (uint32_t ty = r.y0; ty < r.y1 && mask; ++ty, mask = w == 64u ? 0ull : (mask >> w)) {
                unsigned long long row = mask & row_mask;
                const uint32_t key0 = ty * tile_w + r.x0;
                while (row) { // one iteration per contributing tile
                    const uint32_t k = (uint32_t)__ffsll((long long)row) - 1u;
                    row &= row - 1ull;
                    if (pos < end && pos < n_cap) {
                        tile_keys[pos] = key0 + k;
                        vals[pos] = g;
                        ++pos;
                    }
                }
            }
            continue;
        }
        const float4* gp = reinterpret_cast<const float4*>(gauss + g);
        const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
        const float mx = g0.x - 0.5f, my = g0.y - 0.5f, ca = g0.z, cb = g0.w, cc = g1.x, pt = g2.y;
        for (uint32_t ty = r.y0; ty < r.y1; ++ty)
            for (uint32_t tx = r.x0; tx < r.x1; ++tx)
                if (fg_will_contribute(mx, my, ca, cb, cc, tx, ty, pt) && pos < end && pos < n_cap) {
                    tile_keys[pos] = ty * tile_w + tx;
                    vals[pos] = g;
                    ++pos;
                }
    }
}

// Warp-cooperative emission (A/B variant, fg_variant = 1; measured SLOWER on C3: the op surface takes 1.96 instead of 1.90 ms
// per view with this and k_fg_preprocess_coop -- the owner search and the shared-memory traffic cost more than the idle
// lanes they remove).  k_fg_emit gives every primitive to one thread, so a warp runs as long as its largest
// rectangle and its stores scatter (measured: 6.6 of 32 lanes active per instruction, 0.16 ms at C3).  Here a warp takes 32
// consecutive slots of the depth order, the instances of all of them form ONE contiguous output range (off is their
// exclusive scan), and instance t of that range is written by lane t mod 32: its owner is found by a binary search in the
// warp's prefix sums, its tile is the k-th set bit of the owner's mask.  Rectangles of more than 64 tiles (no mask) are
// walked by their owner lane afterwards.
constexpr int kEmitWarps = 8;
__global__ void __launch_bounds__(kEmitWarps * 32)
    k_fg_emit_coop(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ off, const uint32_t n_gauss,
                   const TileRect* __restrict__ rects, const int32_t* __restrict__ counts,
                   const unsigned long long* __restrict__ masks, const GaussRec* __restrict__ gauss, const uint32_t tile_w,
                   const uint32_t n_cap, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ vals) {
    __shared__ uint32_t s_pre[kEmitWarps][33];
    __shared__ unsigned long long s_mask[kEmitWarps][32];
    __shared__ uint32_t s_geo[kEmitWarps][32];  // x0 | y0 << 12 | w << 24
    __shared__ uint32_t s_rcp[kEmitWarps][32];  // ceil(65536 / w): bit / w == (bit * rcp) >> 16 for bit < 64
    __shared__ uint32_t s_base[kEmitWarps][32]; // first output position of the owner
    __shared__ uint32_t s_gid[kEmitWarps][32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n_groups = (n_gauss + 31u) / 32u;
    for (uint32_t grp = blockIdx.x * kEmitWarps + warp; grp < n_groups; grp += gridDim.x * kEmitWarps) {
        const uint32_t slot = grp * 32u + lane;
        uint32_t g = 0, cnt = 0, base = 0;
        TileRect r{0, 0, 0, 0};
        if (slot < n_gauss) {
            g = __ldg(perm + slot);
            const int32_t c = counts[g];
            if (c > 0) {
                cnt = (uint32_t)c;
                r = rects[g];
                base = __ldg(off + slot);
            }
        }
        const uint32_t w = (uint32_t)(r.x1 - r.x0), area = w * (uint32_t)(r.y1 - r.y0);
        const bool big = cnt > 0 && area > 64u;
        const uint32_t c_coop = big ? 0u : cnt;
        uint32_t incl = c_coop;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o)
                incl += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        s_pre[warp][lane] = incl - c_coop;
        if (lane == 31)
            s_pre[warp][32] = total;
        s_mask[warp][lane] = c_coop ? masks[g] : 0ull;
        s_geo[warp][lane] = (uint32_t)r.x0 | ((uint32_t)r.y0 << 12) | (w << 24);
        s_rcp[warp][lane] = w ? (65536u + w - 1u) / w : 0u;
        s_base[warp][lane] = base;
        s_gid[warp][lane] = g;
        __syncwarp();
        for (uint32_t t = lane; t < total; t += 32u) {
            // owner = last lane whose exclusive prefix is <= t (lanes without instances share their successor's prefix and
            // are skipped by taking the LAST such lane ... that has instances: pre[o] <= t < pre[o + 1])
            uint32_t lo = 0, hi = 32;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_pre[warp][mid] <= t)
                    lo = mid;
                else
                    hi = mid;
            }
            const uint32_t o = lo;
            const uint32_t k = t - s_pre[warp][o];
            const unsigned long long m = s_mask[warp][o];
            const uint32_t mlo = (uint32_t)m, mhi = (uint32_t)(m >> 32);
            const uint32_t clo = (uint32_t)__popc(mlo);
            const uint32_t bit = k < clo ? __fns(mlo, 0, (int)k + 1) : 32u + __fns(mhi, 0, (int)(k - clo) + 1);
            const uint32_t geo = s_geo[warp][o];
            const uint32_t row = (bit * s_rcp[warp][o]) >> 16;
            const uint32_t col = bit - row * (geo >> 24);
            const uint32_t pos = s_base[warp][o] + k;
            if (pos < n_cap) {
                tile_keys[pos] = (((geo >> 12) & 0xfffu) + row) * tile_w + (geo & 0xfffu) + col;
                vals[pos] = s_gid[warp][o];
            }
        }
        if (big) { // the old walk, by the owner
            const float4* gp = reinterpret_cast<const float4*>(gauss + g);
            const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
            const float mx = g0.x - 0.5f, my = g0.y - 0.5f, ca = g0.z, cb = g0.w, cc = g1.x, pt = g2.y;
            uint32_t pos = base;
            const uint32_t end = base + cnt;
            for (uint32_t ty = r.y0; ty < r.y1; ++ty)
                for (uint32_t tx = r.x0; tx < r.x1; ++tx)
                    if (fg_will_contribute(mx, my, ca, cb, cc, tx, ty, pt) && pos < end && pos < n_cap) {
                        tile_keys[pos] = ty * tile_w + tx;
                        vals[pos] = g;
                        ++pos;
                    }
        }
        __syncwarp();
    }
}

// image [3,H,W] (no background: composited by the caller, fast_rasterizer.cpp:63), alpha [1,H,W]
__global__ void __launch_bounds__(256)
    k_fg_export(const float4* __restrict__ pix_state, const uint32_t npix, float* __restrict__ image,
                float* __restrict__ alpha) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npix)
        return;
    const float4 s = pix_state[i];
    image[i] = s.x, image[npix + i] = s.y, image[2 * (size_t)npix + i] = s.z;
    alpha[i] = 1.0f - s.w;
}

// v_pix = (dL/drgb, dL/dalpha * T_final)   (grad_alpha_common, kernels_backward.cuh:343-345)
__global__ void __launch_bounds__(256)
    k_fg_pack_vpix(const float* __restrict__ grad_image, const float* __restrict__ grad_alpha,
                   const float4* __restrict__ pix_state, const uint32_t npix, float4* __restrict__ v_pix) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npix)
        return;
    v_pix[i] = make_float4(grad_image[i], grad_image[npix + i], grad_image[2 * (size_t)npix + i],
                           grad_alpha[i] * pix_state[i].w);
}

// ------------------------------------------------------------------------------------------------------
// preprocess backward (kernels_backward.cuh:18-237 + kernel_utils.cuh:41-105)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    k_fg_preprocess_bwd(const float* __restrict__ means, const float* __restrict__ raw_scales,
                        const float4* __restrict__ raw_rotations, const float* __restrict__ sh_rest,
                        const float4* __restrict__ w2c, const float* __restrict__ cam_position,
                        const int32_t* __restrict__ counts, const float* __restrict__ v_m2d,
                        const float* __restrict__ v_conic, const float* __restrict__ v_col,
                        const float* __restrict__ v_op, float* __restrict__ g_means, float* __restrict__ g_scales,
                        float4* __restrict__ g_rot, float* __restrict__ g_opac, float* __restrict__ g_sh0,
                        float* __restrict__ g_shN, float* __restrict__ g_w2c, float* __restrict__ densification_info,
                        const uint32_t N, const int active, const int total_rest, const float w, const float h,
                        const float fx, const float fy, const float cx, const float cy) {
    // sh_rest rows of the block come in, and the g_shN rows go out, through one shared buffer with coalesced 128-bit
    // accesses: each thread overwrites its own row (coefficients -> gradients) once it has read it
    __shared__ __align__(16) float s_sh[128 * 45];
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    const bool on = i < N && counts[i] > 0;
    const uint32_t row_floats = 3u * (uint32_t)total_rest;
    if (active > 1)
        fg_stage_rows(sh_rest, s_sh, blockIdx.x * 128u, N, row_floats);
    __syncthreads();
    f3 dcam = mk3(0.f, 0.f, 0.f), mean = mk3(0.f, 0.f, 0.f);
    if (i < N && !on) { // the reference leaves torch::zeros there (rasterization_api.cu:120-125)
        g_means[3 * (size_t)i] = g_means[3 * (size_t)i + 1] = g_means[3 * (size_t)i + 2] = 0.f;
        g_scales[3 * (size_t)i] = g_scales[3 * (size_t)i + 1] = g_scales[3 * (size_t)i + 2] = 0.f;
        g_rot[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        g_opac[i] = 0.f;
        g_sh0[3 * (size_t)i] = g_sh0[3 * (size_t)i + 1] = g_sh0[3 * (size_t)i + 2] = 0.f;
        for (uint32_t j = 0; j < row_floats; ++j)
            s_sh[threadIdx.x * row_floats + j] = 0.f;
    }
    if (on) {
        mean = mk3(means[3 * (size_t)i], means[3 * (size_t)i + 1], means[3 * (size_t)i + 2]);
        // ---- SH backward
        const f3 gc = mk3(v_col[3 * (size_t)i], v_col[3 * (size_t)i + 1], v_col[3 * (size_t)i + 2]);
        g_sh0[3 * (size_t)i] = 0.28209479177387814f * gc.x, g_sh0[3 * (size_t)i + 1] = 0.28209479177387814f * gc.y;
        g_sh0[3 * (size_t)i + 2] = 0.28209479177387814f * gc.z;
        f3 dpos = mk3(0.f, 0.f, 0.f);
        float* gN = s_sh + threadIdx.x * row_floats;
        int nb = 0;
        if (active > 1) {
            const float xr = mean.x - cam_position[0], yr = mean.y - cam_position[1], zr = mean.z - cam_position[2];
            const float inv = rsqrtf(xr * xr + yr * yr + zr * zr);
            const float x = xr * inv, y = yr * inv, z = zr * inv;
            float b[15];
            nb = fg_sh_basis(active, x, y, z, b);
            const float* c = gN; // this thread's row: read here, overwritten by the gradients below
            float cg[15]; // <coefficient_j, grad_color>
#pragma unroll
            for (int j = 0; j < 15; ++j) {
                cg[j] = 0.f;
                if (j < nb) {
                    cg[j] = c[3 * j] * gc.x + c[3 * j + 1] * gc.y + c[3 * j + 2] * gc.z;
                    gN[3 * j] = b[j] * gc.x, gN[3 * j + 1] = b[j] * gc.y, gN[3 * j + 2] = b[j] * gc.z;
                }
            }
            float gdx = -0.48860251190291987f * cg[2], gdy = -0.48860251190291987f * cg[0], gdz = 0.48860251190291987f * cg[1];
            if (active > 4) {
                gdx += 1.0925484305920792f * (y * cg[3] - z * cg[6] + x * cg[7]);
                gdy += 1.0925484305920792f * (x * cg[3] - z * cg[4] - y * cg[7]);
                gdz += -1.0925484305920792f * y * cg[4] + 1.8923493915151202f * z * cg[5] - 1.0925484305920792f * x * cg[6];
                if (active > 9) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
                    gdx += -3.5402615395598609f * xy * cg[8] + 2.8906114426405538f * yz * cg[9] +
                           (0.45704579946446572f - 2.2852289973223288f * zz) * cg[12] + 2.8906114426405538f * xz * cg[13] +
                           (-1.7701307697799304f * xx + 1.7701307697799304f * yy) * cg[14];
                    gdy += (-1.7701307697799304f * xx + 1.7701307697799304f * yy) * cg[8] + 2.8906114426405538f * xz * cg[9] +
                           (0.45704579946446572f - 2.2852289973223288f * zz) * cg[10] - 2.8906114426405538f * yz * cg[13] +
                           3.5402615395598609f * xy * cg[14];
                    gdz += 2.8906114426405538f * xy * cg[9] - 4.5704579946446566f * yz * cg[10] +
                           (5.597644988851731f * zz - 1.1195289977703462f) * cg[11] - 4.5704579946446566f * xz * cg[12] +
                           (1.4453057213202769f * xx - 1.4453057213202769f * yy) * cg[13];
                }
            }
            const float xx = xr * xr, yy = yr * yr, zz = zr * zr, xy = xr * yr, xz = xr * zr, yz = yr * zr;
            const float n2 = xx + yy + zz;
            const float s3 = rsqrtf(n2 * n2 * n2);
            dpos = mk3(((yy + zz) * gdx - xy * gdy - xz * gdz) * s3, (-xy * gdx + (xx + zz) * gdy - yz * gdz) * s3,
                       (-xz * gdx - yz * gdy + (xx + yy) * gdz) * s3);
        }
        for (int j = 3 * nb; j < 3 * total_rest; ++j) // inactive bands: torch::zeros in the reference
            gN[j] = 0.f;

        // ---- EWA chain
        const float4 r1 = w2c[0], r2 = w2c[1], r3 = w2c[2];
        FgGeom G;
        fg_geometry(mean, mk3(raw_scales[3 * (size_t)i], raw_scales[3 * (size_t)i + 1], raw_scales[3 * (size_t)i + 2]),
                    raw_rotations[i], r1, r2, r3, w, h, fx, fy, cx, cy, G);
        const float a = G.a, b = G.b, c = G.c;
        const float det = a * c - b * b, dr = 1.0f / det, dr2 = dr * dr;
        // reference convention: dL_dconic.y is HALF the true derivative w.r.t. the off-diagonal conic entry
        const float dcx = v_conic[3 * (size_t)i], dcy = 0.5f * v_conic[3 * (size_t)i + 1], dcz = v_conic[3 * (size_t)i + 2];
        const float dcov0 = dr2 * (2.0f * b * c * dcy - c * c * dcx - b * b * dcz);
        const float dcov1 = dr2 * (b * c * dcx - (a * c + b * b) * dcy + a * b * dcz);
        const float dcov2 = dr2 * (2.0f * a * b * dcy - b * b * dcx - a * a * dcz);
        const float u[3] = {G.jw1.x, G.jw1.y, G.jw1.z}, v[3] = {G.jw2.x, G.jw2.y, G.jw2.z};
        float dS[6];
        {
            int t = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = r; s < 3; ++s)
                    dS[t++] = (r == s) ? (u[r] * u[r] * dcov0 + 2.0f * u[r] * v[r] * dcov1 + v[r] * v[r] * dcov2)
                                       : (u[r] * u[s] * dcov0 + (u[r] * v[s] + u[s] * v[r]) * dcov1 + v[r] * v[s] * dcov2);
        }
        const f3 djw1 = (G.jwc1 * dcov0 + G.jwc2 * dcov1) * 2.0f, djw2 = (G.jwc1 * dcov1 + G.jwc2 * dcov2) * 2.0f;
        const f3 w1 = mk3(r1.x, r1.y, r1.z), w2_ = mk3(r2.x, r2.y, r2.z), w3 = mk3(r3.x, r3.y, r3.z);
        const float dj11 = dot(w1, djw1), dj22 = dot(w2_, djw2), dj13 = dot(w3, djw1), dj23 = dot(w3, djw2);
        const float h1 = dj11 - 2.0f * G.tx * dj13, h2 = dj22 - 2.0f * G.ty * dj23;
        const float dmx = v_m2d[2 * (size_t)i], dmy = v_m2d[2 * (size_t)i + 1];
        dcam = mk3(G.j11 * (dmx - dj13 / G.depth), G.j22 * (dmy - dj23 / G.depth),
                   -G.j11 * (G.x * dmx + h1 / G.depth) - G.j22 * (G.y * dmy + h2 / G.depth));
        g_means[3 * (size_t)i] = r1.x * dcam.x + r2.x * dcam.y + r3.x * dcam.z + dpos.x;
        g_means[3 * (size_t)i + 1] = r1.y * dcam.x + r2.y * dcam.y + r3.y * dcam.z + dpos.y;
        g_means[3 * (size_t)i + 2] = r1.z * dcam.x + r2.z * dcam.y + r3.z * dcam.z + dpos.z;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float R0 = G.rot[k], R1 = G.rot[3 + k], R2 = G.rot[6 + k];
            const float dvar = R0 * R0 * dS[0] + R1 * R1 * dS[3] + R2 * R2 * dS[5] +
                               2.0f * (R0 * R1 * dS[1] + R0 * R2 * dS[2] + R1 * R2 * dS[4]);
            g_scales[3 * (size_t)i + k] = 2.0f * G.variance[k] * dvar;
        }
        const float Sm[9] = {dS[0], dS[1], dS[2], dS[1], dS[3], dS[4], dS[2], dS[4], dS[5]};
        float dR[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                dR[3 * r + k] = 2.0f * (G.rs[k] * Sm[3 * r] + G.rs[3 + k] * Sm[3 * r + 1] + G.rs[6 + k] * Sm[3 * r + 2]);
        const float dqxx = -dR[4] - dR[8], dqyy = -dR[0] - dR[8], dqzz = -dR[0] - dR[4];
        const float dqxy = dR[1] + dR[3], dqxz = dR[2] + dR[6], dqyz = dR[5] + dR[7];
        const float dqrx = dR[7] - dR[5], dqry = dR[2] - dR[6], dqrz = dR[3] - dR[1];
        const float* q2 = G.q2;
        const float nh = q2[0] * dqxx + q2[1] * dqyy + q2[2] * dqzz + q2[3] * dqxy + q2[4] * dqxz + q2[5] * dqyz +
                         q2[6] * dqrx + q2[7] * dqry + q2[8] * dqrz;
        const float qr = G.q[0], qx = G.q[1], qy = G.q[2], qz = G.q[3];
        const float iq = 2.0f / G.qn2;
        g_rot[i] = make_float4(iq * (qx * dqrx + qy * dqry + qz * dqrz - qr * nh),
                               iq * (2.0f * qx * dqxx + qy * dqxy + qz * dqxz + qr * dqrx - qx * nh),
                               iq * (2.0f * qy * dqyy + qx * dqxy + qz * dqyz + qr * dqry - qy * nh),
                               iq * (2.0f * qz * dqzz + qx * dqxz + qy * dqyz + qr * dqrz - qz * nh));
        g_opac[i] = v_op[i];
        if (densification_info) { // kernels_backward.cuh:233-236
            densification_info[i] += 1.0f;
            const float sx = dmx * 0.5f * w, sy = dmy * 0.5f * h;
            densification_info[(size_t)N + i] += sqrtf(sx * sx + sy * sy);
        }
    }
    __syncthreads();
    if (total_rest > 0)
        fg_unstage_rows(g_shN, s_sh, blockIdx.x * 128u, N, row_floats);
    if (g_w2c) { // kernels_backward.cuh:162-175, one atomic per warp and entry instead of one per primitive
        float vals[12] = {dcam.x * mean.x, dcam.x * mean.y, dcam.x * mean.z, dcam.x, dcam.y * mean.x, dcam.y * mean.y,
                          dcam.y * mean.z, dcam.y, dcam.z * mean.x, dcam.z * mean.y, dcam.z * mean.z, dcam.z};
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            float s = vals[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0 && s != 0.f)
                atomicAdd(g_w2c + k, s);
        }
    }
}

// ---- blob layouts (the four buffers of the reference's forward, buffer_utils.h:38-151, with our own contents) -------
struct FgPrim {
    GaussRec* gauss;
    TileRect* rects;
    int32_t* counts;
    unsigned long long* masks; // contributing tiles of a rectangle of <= 64 tiles, row-major
    uint32_t *dk_a, *dk_b, *pm_a, *pm_b, *off, *counters;
    void *sort_scr, *scan_scr;
    size_t bytes;
};
static FgPrim carve_prim(void* blob, uint32_t N) {
    Carver c(blob);
    FgPrim p;
    p.gauss = c.take<GaussRec>(N);
    p.rects = c.take<TileRect>(N);
    p.counts = c.take<int32_t>(N);
    p.masks = c.take<unsigned long long>(N);
    p.dk_a = c.take<uint32_t>(N), p.dk_b = c.take<uint32_t>(N);
    p.pm_a = c.take<uint32_t>(N), p.pm_b = c.take<uint32_t>(N);
    p.off = c.take<uint32_t>(N);
    p.counters = c.take<uint32_t>(8); // [0] n_instances, [1] n_visible
    p.sort_scr = c.take<char>(radix_scratch_bytes(N));
    p.scan_scr = c.take<char>(scan_scratch_bytes(N));
    p.bytes = c.total();
    return p;
}
struct FgTile {
    int32_t* tile_off;
    uint32_t *bucket_off, *counts_tmp, *tile_max, *n_buckets;
    float4* pix_state;
    int32_t* n_contrib;
    void* scan_scr;
    size_t bytes;
};
static FgTile carve_tile(void* blob, uint32_t n_tiles, uint32_t npix) {
    Carver c(blob);
    FgTile t;
    t.tile_off = c.take<int32_t>(n_tiles + 1);
    t.bucket_off = c.take<uint32_t>(n_tiles + 1);
    t.counts_tmp = c.take<uint32_t>(n_tiles + 1);
    t.tile_max = c.take<uint32_t>(n_tiles);
    t.n_buckets = c.take<uint32_t>(4);
    t.pix_state = c.take<float4>(npix);
    t.n_contrib = c.take<int32_t>(npix);
    t.scan_scr = c.take<char>(scan_scratch_bytes(n_tiles + 1));
    t.bytes = c.total();
    return t;
}
struct FgInst {
    uint32_t *tk_a, *tk_b, *tv_a, *tv_b;
    void* sort_scr;
    size_t bytes;
};
static FgInst carve_inst(void* blob, uint32_t n_inst) {
    Carver c(blob);
    FgInst s;
    const uint32_t n = n_inst ? n_inst : 1;
    s.tk_a = c.take<uint32_t>(n), s.tk_b = c.take<uint32_t>(n);
    s.tv_a = c.take<uint32_t>(n), s.tv_b = c.take<uint32_t>(n);
    s.sort_scr = c.take<char>(radix_scratch_bytes(n));
    s.bytes = c.total();
    return s;
}
struct FgBucket {
    uint32_t* bucket_tile;
    uint32_t* live;
    float4* ckpt;
    size_t bytes;
};
static FgBucket carve_bucket(void* blob, uint32_t n_buckets) {
    Carver c(blob);
    FgBucket b;
    const uint32_t n = n_buckets ? n_buckets : 1;
    b.bucket_tile = c.take<uint32_t>(n);
    b.live = c.take<uint32_t>(live_list_words(n));
    b.ckpt = c.take<float4>((size_t)n * kTilePix);
    b.bytes = c.total();
    return b;
}

} // namespace lfs

using namespace lfs;

extern "C" int lfs_fastgs_forward(const float* means, const float* scales_raw, const float* rotations_raw,
                                  const float* opacities_raw, const float* sh_coefficients_0,
                                  const float* sh_coefficients_rest, const float* w2c, const float* cam_position,
                                  uint32_t n_primitives, int active_sh_bases, int total_bases_sh_rest, int width,
                                  int height, float focal_x, float focal_y, float center_x, float center_y,
                                  float near_plane, float far_plane, float* image, float* alpha, lfs_alloc_fn alloc,
                                  void* alloc_ctx, int* n_visible_primitives, int* n_instances, int* n_buckets,
                                  int* instance_selector, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(means && scales_raw && rotations_raw && opacities_raw && sh_coefficients_0 && w2c && cam_position,
                  "fastgs_forward: null input");
    LFS_CHECK_ARG(image && alpha && alloc && n_visible_primitives && n_instances && n_buckets && instance_selector,
                  "fastgs_forward: null output");
    LFS_CHECK_ARG(width > 0 && height > 0 && n_primitives > 0, "fastgs_forward: empty problem");
    LFS_CHECK_ARG(active_sh_bases == 1 || active_sh_bases == 4 || active_sh_bases == 9 || active_sh_bases == 16,
                  "fastgs_forward: active_sh_bases must be 1, 4, 9 or 16 (got %d)", active_sh_bases);
    LFS_UNSUPPORTED(total_bases_sh_rest > 15, "fastgs_forward: %d higher SH bases (at most 15 = degree 3, as the reference)",
                    total_bases_sh_rest);
    LFS_CHECK_ARG(active_sh_bases - 1 <= total_bases_sh_rest && (total_bases_sh_rest == 0 || sh_coefficients_rest),
                  "fastgs_forward: sh_coefficients_rest holds %d bases, %d are active", total_bases_sh_rest,
                  active_sh_bases - 1);
    const uint32_t N = n_primitives;
    const uint32_t tile_w = (width + kTile - 1) / kTile, tile_h = (height + kTile - 1) / kTile, n_tiles = tile_w * tile_h;
    const uint32_t npix = (uint32_t)width * (uint32_t)height;
    LFS_UNSUPPORTED(n_tiles > 65535u, "fastgs_forward: %u tiles exceed the reference's 16-bit tile keys", n_tiles);
    LFS_UNSUPPORTED(tile_w > 4095u || tile_h > 4095u, "fastgs_forward: more than 4095 tiles along one axis");

    void* prim_blob = alloc(alloc_ctx, LFS_TAG_FG_PER_PRIMITIVE, carve_prim(nullptr, N).bytes);
    void* tile_blob = alloc(alloc_ctx, LFS_TAG_FG_PER_TILE, carve_tile(nullptr, n_tiles, npix).bytes);
    if (!prim_blob || !tile_blob) {
        set_error("fastgs_forward: buffer allocation failed");
        return LFS_ERR_ALLOC;
    }
    const FgPrim P = carve_prim(prim_blob, N);
    const FgTile T = carve_tile(tile_blob, n_tiles, npix);
    LFS_CUDA_OK(cudaMemsetAsync(P.counters, 0, sizeof(uint32_t) * 8, stream));
    if (raster_options().fg_variant != 1)
        k_fg_preprocess<<<div_up(N, 128), 128, 0, stream>>>(
            means, scales_raw, reinterpret_cast<const float4*>(rotations_raw), opacities_raw, sh_coefficients_0,
            sh_coefficients_rest, reinterpret_cast<const float4*>(w2c), cam_position, N, tile_w, tile_h, active_sh_bases,
            total_bases_sh_rest, (float)width, (float)height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
            P.gauss, P.rects, P.counts, P.masks, P.dk_a, P.pm_a, P.counters + 1);
    else
        k_fg_preprocess_coop<<<div_up(N, 128), 128, 0, stream>>>(
            means, scales_raw, reinterpret_cast<const float4*>(rotations_raw), opacities_raw, sh_coefficients_0,
            sh_coefficients_rest, reinterpret_cast<const float4*>(w2c), cam_position, N, tile_w, tile_h, active_sh_bases,
            total_bases_sh_rest, (float)width, (float)height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
            P.gauss, P.rects, P.counts, P.masks, P.dk_a, P.pm_a, P.counters + 1);
    LFS_LAUNCH_OK("k_fg_preprocess");
    int in_b = 0;
    int rc = radix_sort_pairs(P.dk_a, P.pm_a, P.dk_b, P.pm_b, N, nullptr, 0, 32, P.sort_scr, &in_b, stream);
    if (rc)
        return rc;
    const uint32_t* perm = in_b ? P.pm_b : P.pm_a;
    rc = exclusive_scan_u32(reinterpret_cast<const uint32_t*>(P.counts), perm, P.off, P.counters, N, nullptr, P.scan_scr,
                            stream);
    if (rc)
        return rc;
    uint32_t h_cnt[2] = {0, 0}; // the reference blocks here too (src/forward.cu:97-100)
    LFS_CUDA_OK(cudaMemcpyAsync(h_cnt, P.counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, stream));
    LFS_CUDA_OK(cudaStreamSynchronize(stream));
    const uint32_t n_inst = h_cnt[0];
    LFS_CHECK_ARG(n_inst < (1u << 31), "fastgs_forward: %u instances do not fit int32", n_inst);
    *n_instances = (int)n_inst;
    *n_visible_primitives = (int)h_cnt[1];

    void* inst_blob = alloc(alloc_ctx, LFS_TAG_FG_PER_INSTANCE, carve_inst(nullptr, n_inst).bytes);
    if (!inst_blob) {
        set_error("fastgs_forward: per-instance allocation failed (%u instances)", n_inst);
        return LFS_ERR_ALLOC;
    }
    const FgInst I = carve_inst(inst_blob, n_inst);
    in_b = 0;
    if (n_inst > 0) {
        const unsigned want = div_up(N, 256);
        const unsigned grid = want < (unsigned)(num_sms() * 16) ? want : (unsigned)(num_sms() * 16);
        if (raster_options().fg_variant != 1) // one thread per primitive: measured faster than the cooperative walk
            k_fg_emit<<<grid, 256, 0, stream>>>(perm, P.off, N, P.rects, P.counts, P.masks, P.gauss, tile_w, n_inst, I.tk_a,
                                                I.tv_a);
        else
        {
            const unsigned want_c = div_up(div_up(N, 32), kEmitWarps), cap_c = (unsigned)(num_sms() * 8);
            k_fg_emit_coop<<<want_c < cap_c ? want_c : cap_c, kEmitWarps * 32, 0, stream>>>(
                perm, P.off, N, P.rects, P.counts, P.masks, P.gauss, tile_w, n_inst, I.tk_a, I.tv_a);
        }
        LFS_LAUNCH_OK("k_fg_emit");
        rc = radix_sort_pairs(I.tk_a, I.tv_a, I.tk_b, I.tv_b, n_inst, nullptr, 0, tile_key_bits(n_tiles), I.sort_scr, &in_b,
                              stream);
        if (rc)
            return rc;
    }
    *instance_selector = in_b;
    const uint32_t* sorted_keys = in_b ? I.tk_b : I.tk_a;
    const uint32_t* sorted_vals = in_b ? I.tv_b : I.tv_a;
    rc = launch_tile_offsets(sorted_keys, n_inst, nullptr, n_tiles, T.tile_off, stream);
    if (rc)
        return rc;
    RasterBuffers rb{};
    rb.gauss = P.gauss;
    rb.tile_off = T.tile_off;
    rb.inst_gid = reinterpret_cast<const int32_t*>(sorted_vals);
    rb.bucket_off = T.bucket_off;
    rb.tile_max_contrib = T.tile_max;
    rb.pix_state = T.pix_state;
    rb.n_contrib = T.n_contrib;
    rc = launch_bucket_offsets(rb, n_tiles, T.n_buckets, T.scan_scr, T.counts_tmp, stream);
    if (rc)
        return rc;
    uint32_t h_nb = 0; // src/forward.cu:175
    LFS_CUDA_OK(cudaMemcpyAsync(&h_nb, T.n_buckets, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    LFS_CUDA_OK(cudaStreamSynchronize(stream));
    *n_buckets = (int)h_nb;
    void* bucket_blob = alloc(alloc_ctx, LFS_TAG_FG_PER_BUCKET, carve_bucket(nullptr, h_nb).bytes);
    if (!bucket_blob) {
        set_error("fastgs_forward: per-bucket allocation failed (%u buckets)", h_nb);
        return LFS_ERR_ALLOC;
    }
    const FgBucket B = carve_bucket(bucket_blob, h_nb);
    rb.bucket_tile = B.bucket_tile;
    rb.live = B.live;
    rb.ckpt = B.ckpt;
    rc = launch_blend_fwd_ewa(rb, (uint32_t)width, (uint32_t)height, tile_w, tile_h, true, stream);
    if (rc)
        return rc;
    k_fg_export<<<div_up(npix, 256), 256, 0, stream>>>(T.pix_state, npix, image, alpha);
    LFS_LAUNCH_OK("k_fg_export");
    return LFS_OK;
}

extern "C" int lfs_fastgs_backward(const float* grad_image, const float* grad_alpha, const float* means,
                                   const float* scales_raw, const float* rotations_raw, const float* sh_coefficients_rest,
                                   const float* w2c, const float* cam_position, const void* per_primitive_buffers,
                                   const void* per_tile_buffers, const void* per_instance_buffers,
                                   const void* per_bucket_buffers, float* grad_means, float* grad_scales_raw,
                                   float* grad_rotations_raw, float* grad_opacities_raw, float* grad_sh_coefficients_0,
                                   float* grad_sh_coefficients_rest, float* grad_w2c, float* densification_info,
                                   uint32_t n_primitives, int n_instances, int n_buckets, int instance_selector,
                                   int active_sh_bases, int total_bases_sh_rest, int width, int height, float focal_x,
                                   float focal_y, float center_x, float center_y, lfs_alloc_fn alloc, void* alloc_ctx,
                                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(grad_image && grad_alpha && means && scales_raw && rotations_raw && w2c && cam_position,
                  "fastgs_backward: null input");
    LFS_CHECK_ARG(per_primitive_buffers && per_tile_buffers && per_instance_buffers && per_bucket_buffers,
                  "fastgs_backward: a forward buffer is missing");
    LFS_CHECK_ARG(grad_means && grad_scales_raw && grad_rotations_raw && grad_opacities_raw && grad_sh_coefficients_0 &&
                      (total_bases_sh_rest == 0 || grad_sh_coefficients_rest) && alloc,
                  "fastgs_backward: null output");
    LFS_CHECK_ARG(n_instances >= 0 && n_buckets >= 0 && width > 0 && height > 0 && n_primitives > 0,
                  "fastgs_backward: bad sizes");
    LFS_UNSUPPORTED(total_bases_sh_rest > 15, "fastgs_backward: %d higher SH bases (at most 15 = degree 3, as the reference)",
                    total_bases_sh_rest);
    const uint32_t N = n_primitives;
    const uint32_t tile_w = (width + kTile - 1) / kTile, tile_h = (height + kTile - 1) / kTile, n_tiles = tile_w * tile_h;
    const uint32_t npix = (uint32_t)width * (uint32_t)height;
    const FgPrim P = carve_prim(const_cast<void*>(per_primitive_buffers), N);
    const FgTile T = carve_tile(const_cast<void*>(per_tile_buffers), n_tiles, npix);
    const FgInst I = carve_inst(const_cast<void*>(per_instance_buffers), (uint32_t)n_instances);
    const FgBucket B = carve_bucket(const_cast<void*>(per_bucket_buffers), (uint32_t)n_buckets);

    Carver cs(nullptr);
    cs.take<float4>(npix), cs.take<float>(2 * (size_t)N), cs.take<float>(3 * (size_t)N), cs.take<float>(3 * (size_t)N),
        cs.take<float>(N);
    void* scratch = alloc(alloc_ctx, LFS_TAG_SCRATCH, cs.total());
    if (!scratch) {
        set_error("fastgs_backward: scratch allocation failed");
        return LFS_ERR_ALLOC;
    }
    Carver c(scratch);
    float4* v_pix = c.take<float4>(npix);
    float* v_m2d = c.take<float>(2 * (size_t)N);
    float* v_conic = c.take<float>(3 * (size_t)N);
    float* v_col = c.take<float>(3 * (size_t)N);
    float* v_op = c.take<float>(N);
    LFS_CUDA_OK(cudaMemsetAsync(v_m2d, 0, reinterpret_cast<char*>(v_op + N) - reinterpret_cast<char*>(v_m2d), stream));
    k_fg_pack_vpix<<<div_up(npix, 256), 256, 0, stream>>>(grad_image, grad_alpha, T.pix_state, npix, v_pix);
    LFS_LAUNCH_OK("k_fg_pack_vpix");
    RasterBuffers rb{};
    rb.gauss = P.gauss;
    rb.tile_off = T.tile_off;
    rb.inst_gid = reinterpret_cast<const int32_t*>(instance_selector ? I.tv_b : I.tv_a);
    rb.bucket_off = T.bucket_off;
    rb.bucket_tile = B.bucket_tile;
    rb.live = B.live;
    rb.ckpt = B.ckpt;
    rb.tile_max_contrib = T.tile_max;
    rb.pix_state = T.pix_state;
    rb.n_contrib = T.n_contrib;
    int rc = launch_blend_bwd_ewa(rb, v_pix, N, (uint32_t)width, (uint32_t)height, tile_w, tile_h, (uint32_t)n_buckets,
                                  T.n_buckets, v_m2d, v_conic, v_col, v_op, stream);
    if (rc)
        return rc;
    k_fg_preprocess_bwd<<<div_up(N, 128), 128, 0, stream>>>(
        means, scales_raw, reinterpret_cast<const float4*>(rotations_raw), sh_coefficients_rest,
        reinterpret_cast<const float4*>(w2c), cam_position, P.counts, v_m2d, v_conic, v_col, v_op, grad_means,
        grad_scales_raw, reinterpret_cast<float4*>(grad_rotations_raw), grad_opacities_raw, grad_sh_coefficients_0,
        grad_sh_coefficients_rest, grad_w2c, densification_info, N, active_sh_bases, total_bases_sh_rest, (float)width,
        (float)height, focal_x, focal_y, center_x, center_y);
    LFS_LAUNCH_OK("k_fg_preprocess_bwd");
    return LFS_OK;
}
