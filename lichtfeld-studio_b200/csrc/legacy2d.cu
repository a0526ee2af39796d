// lfs_b200 -- the legacy 2-D operator surface of the reference's gtest files:
//   gsplat::quat_scale_to_covar_preci_fwd / _bwd   (tests/test_basic.cpp:54-81, tests/test_gsplat_ops.cpp:76-96)
//   gsplat::projection_ewa_3dgs_fused_fwd          (tests/test_basic.cpp:114-128, tests/test_gsplat_ops.cpp:174-189)
//   gsplat::rasterize_to_pixels_3dgs_fwd / _bwd    (tests/test_basic.cpp:347-358; launchers declared in
//                                                   gsplat/Rasterization.h:16-63)
// The reference tree declares / calls these ops but no longer carries their kernels; the semantics implemented here are
// the reference's CPU statement tests/torch_impl.cpp:38-218 (quat/scale -> covariance, pinhole EWA projection), the culling
// tail of its surviving projection kernel (ProjectionUT3DGSFused.cu:142-199) and the blend loop of its from-world kernels
// (RasterizeToPixelsFromWorld3DGSFwd.cu:193-279, ...Bwd.cu:196-370) with the 2-D conic response.  These ops are a test /
// tooling surface, not the training hot path (the trainer and the fastgs surface run the per-tile polynomial kernels of
// raster.cu): the kernels below are the plain one-thread-per-pixel formulation, written for clarity and exact semantics.
#include "common.cuh"

namespace lfs {
namespace {

constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTStop = 1e-4f;

// R (row-major) of the normalised quaternion (w,x,y,z), gsplat/Utils.cuh:80-102
__device__ __forceinline__ void quat_to_R(const float4 q, float R[9], float& inv_norm) {
    inv_norm = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float w = q.x * inv_norm, x = q.y * inv_norm, y = q.z * inv_norm, z = q.w * inv_norm;
    const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.f - 2.f * (y2 + z2), R[1] = 2.f * (xy - wz), R[2] = 2.f * (xz + wy);
    R[3] = 2.f * (xy + wz), R[4] = 1.f - 2.f * (x2 + z2), R[5] = 2.f * (yz - wx);
    R[6] = 2.f * (xz - wy), R[7] = 2.f * (yz + wx), R[8] = 1.f - 2.f * (x2 + y2);
}

// S = (R diag(f)) (R diag(f))^T, symmetric: returns (xx, xy, xz, yy, yz, zz)
__device__ __forceinline__ void rs_outer(const float R[9], const float f0, const float f1, const float f2, float s[6]) {
    const float m[9] = {R[0] * f0, R[1] * f1, R[2] * f2, R[3] * f0, R[4] * f1, R[5] * f2, R[6] * f0, R[7] * f1, R[8] * f2};
    s[0] = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    s[1] = m[0] * m[3] + m[1] * m[4] + m[2] * m[5];
    s[2] = m[0] * m[6] + m[1] * m[7] + m[2] * m[8];
    s[3] = m[3] * m[3] + m[4] * m[4] + m[5] * m[5];
    s[4] = m[3] * m[6] + m[4] * m[7] + m[5] * m[8];
    s[5] = m[6] * m[6] + m[7] * m[7] + m[8] * m[8];
}

__device__ __forceinline__ void store_sym(float* out, const uint32_t i, const float s[6], const bool triu) {
    if (triu) {
        float* o = out + 6 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            o[k] = s[k];
    } else {
        float* o = out + 9 * (size_t)i;
        o[0] = s[0], o[1] = s[1], o[2] = s[2], o[3] = s[1], o[4] = s[3], o[5] = s[4], o[6] = s[2], o[7] = s[4], o[8] = s[5];
    }
}

__global__ void __launch_bounds__(256)
    k_qs_covar_fwd(const float4* __restrict__ quats, const float* __restrict__ scales, const uint32_t N, const bool triu,
                   float* __restrict__ covars, float* __restrict__ precis) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N)
        return;
    float R[9], inv_norm, s[6];
    quat_to_R(__ldg(quats + i), R, inv_norm);
    const float s0 = scales[3 * (size_t)i], s1 = scales[3 * (size_t)i + 1], s2 = scales[3 * (size_t)i + 2];
    if (covars) {
        rs_outer(R, s0, s1, s2, s);
        store_sym(covars, i, s, triu);
    }
    if (precis) {
        rs_outer(R, 1.0f / s0, 1.0f / s1, 1.0f / s2, s);
        store_sym(precis, i, s, triu);
    }
}

// symmetrised upstream gradient V + V^T as (xx, xy, xz, yy, yz, zz) of the full symmetric matrix
__device__ __forceinline__ void load_vsym(const float* v, const uint32_t i, const bool triu, float g[6]) {
    if (triu) { // out_k = (S_rc + S_cr) / 2: dS_rc = dS_cr = v/2 off the diagonal -> (V + V^T)_rc = v
        const float* p = v + 6 * (size_t)i;
        g[0] = 2.f * p[0], g[1] = p[1], g[2] = p[2], g[3] = 2.f * p[3], g[4] = p[4], g[5] = 2.f * p[5];
    } else {
        const float* p = v + 9 * (size_t)i;
        g[0] = 2.f * p[0], g[1] = p[1] + p[3], g[2] = p[2] + p[6], g[3] = 2.f * p[4], g[4] = p[5] + p[7], g[5] = 2.f * p[8];
    }
}

__global__ void __launch_bounds__(256)
    k_qs_covar_bwd(const float4* __restrict__ quats, const float* __restrict__ scales, const uint32_t N, const bool triu,
                   const float* __restrict__ v_covars, const float* __restrict__ v_precis, float4* __restrict__ v_quats,
                   float* __restrict__ v_scales) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N)
        return;
    float R[9], inv_norm;
    const float4 q = __ldg(quats + i);
    quat_to_R(q, R, inv_norm);
    const float sc[3] = {scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2]};
    float G[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; // dL/dR, row-major
    float vs[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const float* vin = which ? v_precis : v_covars;
        if (!vin)
            continue;
        float g[6];
        load_vsym(vin, i, triu, g);
        const float W[9] = {g[0], g[1], g[2], g[1], g[3], g[4], g[2], g[4], g[5]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float f = which ? 1.0f / sc[a] : sc[a]; // M[r][a] = R[r][a] f
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float vM = (W[3 * r] * R[a] + W[3 * r + 1] * R[3 + a] + W[3 * r + 2] * R[6 + a]) * f;
                G[3 * r + a] = fmaf(vM, f, G[3 * r + a]);
                acc = fmaf(R[3 * r + a], vM, acc);
            }
            vs[a] += which ? -acc * f * f : acc;
        }
    }
    // quat_to_rotmat VJP incl. the normalisation (gsplat/Utils.cuh:104-126; G[r][c] = dL/dR[r][c])
    const float w = q.x * inv_norm, x = q.y * inv_norm, y = q.z * inv_norm, z = q.w * inv_norm;
    float vq[4];
    vq[0] = 2.f * (x * (G[7] - G[5]) + y * (G[2] - G[6]) + z * (G[3] - G[1]));
    vq[1] = 2.f * (-2.f * x * (G[4] + G[8]) + y * (G[3] + G[1]) + z * (G[6] + G[2]) + w * (G[7] - G[5]));
    vq[2] = 2.f * (x * (G[3] + G[1]) - 2.f * y * (G[0] + G[8]) + z * (G[7] + G[5]) + w * (G[2] - G[6]));
    vq[3] = 2.f * (x * (G[6] + G[2]) + y * (G[7] + G[5]) - 2.f * z * (G[0] + G[4]) + w * (G[3] - G[1]));
    const float d = vq[0] * w + vq[1] * x + vq[2] * y + vq[3] * z;
    v_quats[i] = make_float4((vq[0] - d * w) * inv_norm, (vq[1] - d * x) * inv_norm, (vq[2] - d * y) * inv_norm,
                             (vq[3] - d * z) * inv_norm);
    v_scales[3 * (size_t)i] = vs[0], v_scales[3 * (size_t)i + 1] = vs[1], v_scales[3 * (size_t)i + 2] = vs[2];
}

// fused pinhole EWA projection, one thread per (camera, Gaussian); torch_impl.cpp:80-218 + ProjectionUT3DGSFused.cu:142-199
__global__ void __launch_bounds__(256)
    k_proj_ewa(const float* __restrict__ means, const float* __restrict__ covars, const float4* __restrict__ quats,
               const float* __restrict__ scales, const float* __restrict__ opacities, const float* __restrict__ viewmats,
               const float* __restrict__ Ks, const uint32_t N, const uint32_t C, const float width, const float height,
               const float eps2d, const float near_plane, const float far_plane, const float radius_clip,
               int32_t* __restrict__ radii, float* __restrict__ means2d, float* __restrict__ depths,
               float* __restrict__ conics, float* __restrict__ compensations) {
    const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= N * C)
        return;
    const uint32_t cid = idx / N, gid = idx - cid * N;
    radii[2 * (size_t)idx] = 0, radii[2 * (size_t)idx + 1] = 0;
    const float* vm = viewmats + 16 * cid;
    const float* K = Ks + 9 * cid;
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const float mx = means[3 * (size_t)gid], my = means[3 * (size_t)gid + 1], mz = means[3 * (size_t)gid + 2];
    const float pc[3] = {vm[0] * mx + vm[1] * my + vm[2] * mz + vm[3], vm[4] * mx + vm[5] * my + vm[6] * mz + vm[7],
                         vm[8] * mx + vm[9] * my + vm[10] * mz + vm[11]};
    if (pc[2] < near_plane || pc[2] > far_plane)
        return;
    float s[6];
    if (covars) {
        const float* p = covars + 9 * (size_t)gid;
        s[0] = p[0], s[1] = p[1], s[2] = p[2], s[3] = p[4], s[4] = p[5], s[5] = p[8];
    } else {
        float R[9], inv_norm;
        quat_to_R(__ldg(quats + gid), R, inv_norm);
        rs_outer(R, scales[3 * (size_t)gid], scales[3 * (size_t)gid + 1], scales[3 * (size_t)gid + 2], s);
    }
    // camera-space covariance W S W^T
    const float S[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
    float WS[9], Sc[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            WS[3 * r + c] = vm[4 * r] * S[c] + vm[4 * r + 1] * S[3 + c] + vm[4 * r + 2] * S[6 + c];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            Sc[3 * r + c] = WS[3 * r] * vm[4 * c] + WS[3 * r + 1] * vm[4 * c + 1] + WS[3 * r + 2] * vm[4 * c + 2];
    const float tz = pc[2], rz = 1.0f / tz, rz2 = rz * rz;
    const float tan_fovx = 0.5f * width / fx, tan_fovy = 0.5f * height / fy;
    const float lim_x_pos = (width - cx) / fx + 0.3f * tan_fovx, lim_x_neg = cx / fx + 0.3f * tan_fovx;
    const float lim_y_pos = (height - cy) / fy + 0.3f * tan_fovy, lim_y_neg = cy / fy + 0.3f * tan_fovy;
    const float tx = tz * fminf(lim_x_pos, fmaxf(-lim_x_neg, pc[0] * rz));
    const float ty = tz * fminf(lim_y_pos, fmaxf(-lim_y_neg, pc[1] * rz));
    const float J[6] = {fx * rz, 0.f, -fx * tx * rz2, 0.f, fy * rz, -fy * ty * rz2};
    float JS[6];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            JS[3 * r + c] = J[3 * r] * Sc[c] + J[3 * r + 1] * Sc[3 + c] + J[3 * r + 2] * Sc[6 + c];
    float c00 = JS[0] * J[0] + JS[1] * J[1] + JS[2] * J[2];
    const float c01 = JS[0] * J[3] + JS[1] * J[4] + JS[2] * J[5];
    const float c10 = JS[3] * J[0] + JS[4] * J[1] + JS[5] * J[2];
    float c11 = JS[3] * J[3] + JS[4] * J[4] + JS[5] * J[5];
    const float u = fx * pc[0] * rz + cx, v = fy * pc[1] * rz + cy;
    const float det_orig = c00 * c11 - c01 * c10;
    c00 += eps2d, c11 += eps2d;
    const float det = c00 * c11 - c01 * c10;
    if (det <= 0.f)
        return;
    const float compensation = sqrtf(fmaxf(0.f, det_orig / det));
    float extend = 3.33f;
    if (opacities) {
        float op = opacities[gid];
        if (compensations)
            op *= compensation;
        if (op < kAlphaMin)
            return;
        extend = fminf(extend, sqrtf(2.0f * logf(op / kAlphaMin)));
    }
    const float radius_x = ceilf(extend * sqrtf(c00)), radius_y = ceilf(extend * sqrtf(c11));
    if (radius_x <= radius_clip && radius_y <= radius_clip)
        return;
    if (u + radius_x <= 0.f || u - radius_x >= width || v + radius_y <= 0.f || v - radius_y >= height)
        return;
    const float ood = 1.0f / det;
    radii[2 * (size_t)idx] = (int32_t)radius_x, radii[2 * (size_t)idx + 1] = (int32_t)radius_y;
    means2d[2 * (size_t)idx] = u, means2d[2 * (size_t)idx + 1] = v;
    depths[idx] = pc[2];
    conics[3 * (size_t)idx] = c11 * ood, conics[3 * (size_t)idx + 1] = -0.5f * (c01 + c10) * ood;
    conics[3 * (size_t)idx + 2] = c00 * ood;
    if (compensations)
        compensations[idx] = compensation;
}

// ---- 2-D blend ------------------------------------------------------------------------------------------------------
// One CTA per tile, one thread per pixel; the tile's Gaussians are staged 256 at a time (xy + opacity, conic, id).  CP is
// the channel count padded to the template grid; colours are read through the read-only path (the same address for the
// whole warp: one transaction).
template <int CP>
__global__ void __launch_bounds__(kTilePix)
    k_raster2d_fwd(const float2* __restrict__ means2d, const float* __restrict__ conics, const float* __restrict__ colors,
                   const float* __restrict__ opacities, const float* __restrict__ backgrounds,
                   const uint8_t* __restrict__ masks, const int channels, const uint32_t width, const uint32_t height,
                   const uint32_t tile_w, const uint32_t tile_h, const int32_t* __restrict__ tile_offsets,
                   const int32_t* __restrict__ flatten_ids, const int64_t n_isects, const uint32_t C,
                   float* __restrict__ renders, float* __restrict__ alphas, int32_t* __restrict__ last_ids) {
    __shared__ float4 s_xyo[kTilePix]; // (x, y, opacity, -)
    __shared__ float4 s_con[kTilePix]; // (a, b, c, -)
    __shared__ int32_t s_id[kTilePix];
    const uint32_t cid = blockIdx.z, tile_id = blockIdx.y * tile_w + blockIdx.x;
    const uint32_t tflat = cid * tile_w * tile_h + tile_id;
    const uint32_t j = blockIdx.x * kTile + (threadIdx.x & 15), i = blockIdx.y * kTile + (threadIdx.x >> 4);
    const bool inside = i < height && j < width;
    const size_t pix = ((size_t)cid * height + i) * width + j;
    const float* bg = backgrounds ? backgrounds + (size_t)cid * channels : nullptr;
    if (masks && !masks[tflat]) { // RasterizeToPixelsFromWorld3DGSFwd.cu:116-128: background, alpha 0
        if (inside) {
            for (int k = 0; k < channels; ++k)
                renders[pix * channels + k] = bg ? bg[k] : 0.f;
            alphas[pix] = 0.f;
            last_ids[pix] = 0;
        }
        return;
    }
    const int64_t start = tile_offsets[tflat];
    const int64_t end = (cid == C - 1 && tile_id == tile_w * tile_h - 1) ? n_isects : (int64_t)tile_offsets[tflat + 1];
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    float T = 1.0f;
    int32_t cur_idx = 0;
    float out[CP];
#pragma unroll
    for (int k = 0; k < CP; ++k)
        out[k] = 0.f;
    bool done = !inside;
    for (int64_t b0 = start; b0 < end; b0 += kTilePix) {
        if (__syncthreads_count(done ? 1 : 0) == kTilePix)
            break;
        const int64_t idx = b0 + threadIdx.x;
        if (idx < end) {
            const int32_t g = __ldg(flatten_ids + idx);
            const float2 xy = __ldg(means2d + g);
            s_id[threadIdx.x] = g;
            s_xyo[threadIdx.x] = make_float4(xy.x, xy.y, __ldg(opacities + g), 0.f);
            s_con[threadIdx.x] = make_float4(__ldg(conics + 3 * (size_t)g), __ldg(conics + 3 * (size_t)g + 1),
                                             __ldg(conics + 3 * (size_t)g + 2), 0.f);
        }
        __syncthreads();
        const int n = (int)min((int64_t)kTilePix, end - b0);
        for (int t = 0; t < n && !done; ++t) {
            const float4 xyo = s_xyo[t], cn = s_con[t];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = 0.5f * (cn.x * dx * dx + cn.z * dy * dy) + cn.y * dx * dy;
            const float alpha = fminf(kAlphaMax, xyo.z * __expf(-sigma));
            if (sigma < 0.f || alpha < kAlphaMin)
                continue;
            const float next_T = T * (1.0f - alpha);
            if (next_T <= kTStop) {
                done = true;
                break;
            }
            const float vis = alpha * T;
            const float* c = colors + (size_t)s_id[t] * channels;
#pragma unroll
            for (int k = 0; k < CP; ++k)
                if (k < channels)
                    out[k] = fmaf(__ldg(c + k), vis, out[k]);
            cur_idx = (int32_t)(b0 + t);
            T = next_T;
        }
    }
    if (inside) {
        alphas[pix] = 1.0f - T;
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < channels)
                renders[pix * channels + k] = bg ? fmaf(T, bg[k], out[k]) : out[k];
        last_ids[pix] = cur_idx;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Back to front from the tile's deepest last contributor; every Gaussian's gradients are summed over the warp and added
// with one atomic per warp and component (the scheme of the reference's backward, ...Bwd.cu:300-370).
template <int CP>
__global__ void __launch_bounds__(kTilePix)
    k_raster2d_bwd(const float2* __restrict__ means2d, const float* __restrict__ conics, const float* __restrict__ colors,
                   const float* __restrict__ opacities, const float* __restrict__ backgrounds,
                   const uint8_t* __restrict__ masks, const int channels, const uint32_t width, const uint32_t height,
                   const uint32_t tile_w, const uint32_t tile_h, const int32_t* __restrict__ tile_offsets,
                   const int32_t* __restrict__ flatten_ids, const float* __restrict__ render_alphas,
                   const int32_t* __restrict__ last_ids, const float* __restrict__ v_render_colors,
                   const float* __restrict__ v_render_alphas, float* __restrict__ v_means2d_abs,
                   float* __restrict__ v_means2d, float* __restrict__ v_conics, float* __restrict__ v_colors,
                   float* __restrict__ v_opacities) {
    __shared__ float4 s_xyo[kTilePix];
    __shared__ float4 s_con[kTilePix];
    __shared__ int32_t s_id[kTilePix];
    __shared__ int32_t s_last[kTilePix / 32];
    const uint32_t cid = blockIdx.z, tile_id = blockIdx.y * tile_w + blockIdx.x;
    const uint32_t tflat = cid * tile_w * tile_h + tile_id;
    if (masks && !masks[tflat])
        return;
    const uint32_t j = blockIdx.x * kTile + (threadIdx.x & 15), i = blockIdx.y * kTile + (threadIdx.x >> 4);
    const bool inside = i < height && j < width;
    const size_t pix = ((size_t)cid * height + min(i, height - 1)) * width + min(j, width - 1);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int32_t start = tile_offsets[tflat];
    const int32_t bin_final = inside ? last_ids[pix] : start - 1;
    const float T_final = 1.0f - render_alphas[pix];
    float T = T_final;
    float buffer[CP], vrc[CP];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        buffer[k] = 0.f;
        vrc[k] = (inside && k < channels) ? v_render_colors[pix * channels + k] : 0.f;
        if (backgrounds && k < channels)
            bg_dot = fmaf(backgrounds[(size_t)cid * channels + k], vrc[k], bg_dot);
    }
    const float vra = inside ? v_render_alphas[pix] : 0.f;
    // deepest last contributor of the tile
    int32_t m = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0)
        s_last[threadIdx.x >> 5] = m;
    __syncthreads();
    int32_t tile_last = s_last[0];
#pragma unroll
    for (int w = 1; w < kTilePix / 32; ++w)
        tile_last = max(tile_last, s_last[w]);
    const bool lane0 = (threadIdx.x & 31) == 0;
    for (int32_t b_hi = tile_last; b_hi >= start; b_hi -= kTilePix) {
        __syncthreads();
        const int32_t idx = b_hi - (int32_t)threadIdx.x; // slot t holds element b_hi - t
        if (idx >= start) {
            const int32_t g = __ldg(flatten_ids + idx);
            const float2 xy = __ldg(means2d + g);
            s_id[threadIdx.x] = g;
            s_xyo[threadIdx.x] = make_float4(xy.x, xy.y, __ldg(opacities + g), 0.f);
            s_con[threadIdx.x] = make_float4(__ldg(conics + 3 * (size_t)g), __ldg(conics + 3 * (size_t)g + 1),
                                             __ldg(conics + 3 * (size_t)g + 2), 0.f);
        }
        __syncthreads();
        const int n = min((int32_t)kTilePix, b_hi - start + 1);
        for (int t = 0; t < n; ++t) {
            const int32_t cur = b_hi - t;
            bool valid = inside && cur <= bin_final;
            const float4 xyo = s_xyo[t], cn = s_con[t];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = 0.5f * (cn.x * dx * dx + cn.z * dy * dy) + cn.y * dx * dy;
            const float vis = __expf(-sigma);
            const float opac = xyo.z;
            const float alpha = fminf(kAlphaMax, opac * vis);
            if (sigma < 0.f || alpha < kAlphaMin)
                valid = false;
            if (!__any_sync(0xffffffffu, valid))
                continue;
            const int32_t g = s_id[t];
            float v_col[CP];
            float v_con0 = 0.f, v_con1 = 0.f, v_con2 = 0.f, v_x = 0.f, v_y = 0.f, v_ax = 0.f, v_ay = 0.f, v_op = 0.f;
#pragma unroll
            for (int k = 0; k < CP; ++k)
                v_col[k] = 0.f;
            if (valid) {
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;
                const float fac = alpha * T;
                const float* c = colors + (size_t)g * channels;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < CP; ++k)
                    if (k < channels) {
                        const float ck = __ldg(c + k);
                        v_col[k] = fac * vrc[k];
                        v_alpha = fmaf(ck * T - buffer[k] * ra, vrc[k], v_alpha);
                        buffer[k] = fmaf(ck, fac, buffer[k]);
                    }
                v_alpha += T_final * ra * vra;
                if (backgrounds)
                    v_alpha += -T_final * ra * bg_dot;
                if (opac * vis <= kAlphaMax) {
                    const float v_sigma = -opac * vis * v_alpha;
                    v_con0 = 0.5f * v_sigma * dx * dx, v_con1 = v_sigma * dx * dy, v_con2 = 0.5f * v_sigma * dy * dy;
                    v_x = v_sigma * (cn.x * dx + cn.y * dy), v_y = v_sigma * (cn.y * dx + cn.z * dy);
                    v_ax = fabsf(v_x), v_ay = fabsf(v_y);
                    v_op = vis * v_alpha;
                }
            }
#pragma unroll
            for (int k = 0; k < CP; ++k)
                if (k < channels) {
                    const float s = warp_sum(v_col[k]);
                    if (lane0)
                        atomicAdd(v_colors + (size_t)g * channels + k, s);
                }
            v_con0 = warp_sum(v_con0), v_con1 = warp_sum(v_con1), v_con2 = warp_sum(v_con2);
            v_x = warp_sum(v_x), v_y = warp_sum(v_y), v_op = warp_sum(v_op);
            if (v_means2d_abs)
                v_ax = warp_sum(v_ax), v_ay = warp_sum(v_ay);
            if (lane0) {
                atomicAdd(v_conics + 3 * (size_t)g, v_con0);
                atomicAdd(v_conics + 3 * (size_t)g + 1, v_con1);
                atomicAdd(v_conics + 3 * (size_t)g + 2, v_con2);
                atomicAdd(v_means2d + 2 * (size_t)g, v_x);
                atomicAdd(v_means2d + 2 * (size_t)g + 1, v_y);
                if (v_means2d_abs) {
                    atomicAdd(v_means2d_abs + 2 * (size_t)g, v_ax);
                    atomicAdd(v_means2d_abs + 2 * (size_t)g + 1, v_ay);
                }
                atomicAdd(v_opacities + g, v_op);
            }
        }
    }
}

int pad_channels(const uint32_t channels) { return channels <= 4 ? 4 : channels <= 8 ? 8 : channels <= 16 ? 16 : 40; }

} // namespace
} // namespace lfs

using namespace lfs;

extern "C" int lfs_quat_scale_to_covar_preci_fwd(const float* quats, const float* scales, uint32_t N, int triu,
                                                 float* covars, float* precis, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(quats && scales, "quat_scale_to_covar_preci_fwd: null input");
    LFS_CHECK_ARG(covars || precis, "quat_scale_to_covar_preci_fwd: neither covars nor precis requested");
    if (N == 0)
        return LFS_OK;
    k_qs_covar_fwd<<<div_up(N, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(quats), scales, N, triu != 0, covars,
                                                        precis);
    LFS_LAUNCH_OK("k_qs_covar_fwd");
    return LFS_OK;
}

extern "C" int lfs_quat_scale_to_covar_preci_bwd(const float* quats, const float* scales, uint32_t N, int triu,
                                                 const float* v_covars, const float* v_precis, float* v_quats,
                                                 float* v_scales, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(quats && scales && v_quats && v_scales, "quat_scale_to_covar_preci_bwd: null pointer");
    if (N == 0)
        return LFS_OK;
    k_qs_covar_bwd<<<div_up(N, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(quats), scales, N, triu != 0,
                                                        v_covars, v_precis, reinterpret_cast<float4*>(v_quats), v_scales);
    LFS_LAUNCH_OK("k_qs_covar_bwd");
    return LFS_OK;
}

extern "C" int lfs_projection_ewa_3dgs_fused_fwd(const float* means, const float* covars, const float* quats,
                                                 const float* scales, const float* opacities, const float* viewmats,
                                                 const float* Ks, uint32_t N, uint32_t C, uint32_t image_width,
                                                 uint32_t image_height, float eps2d, float near_plane, float far_plane,
                                                 float radius_clip, int camera_model, int32_t* radii, float* means2d,
                                                 float* depths, float* conics, float* compensations, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(means && viewmats && Ks && radii && means2d && depths && conics, "projection_ewa: null pointer");
    LFS_CHECK_ARG(covars || (quats && scales), "projection_ewa: either covars or quats + scales are required");
    LFS_UNSUPPORTED(camera_model != LFS_PINHOLE, "projection_ewa: only the pinhole camera model (the reference's CPU "
                                                 "statement tests/torch_impl.cpp:165-169 has no other either)");
    if ((uint64_t)N * C == 0)
        return LFS_OK;
    k_proj_ewa<<<div_up((uint64_t)N * C, 256), 256, 0, stream>>>(
        means, covars, reinterpret_cast<const float4*>(quats), scales, opacities, viewmats, Ks, N, C, (float)image_width,
        (float)image_height, eps2d, near_plane, far_plane, radius_clip, radii, means2d, depths, conics, compensations);
    LFS_LAUNCH_OK("k_proj_ewa");
    return LFS_OK;
}

extern "C" int lfs_rasterize_to_pixels_3dgs_fwd(const float* means2d, const float* conics, const float* colors,
                                                const float* opacities, const float* backgrounds, const uint8_t* masks,
                                                uint32_t C, uint32_t N, uint32_t channels, uint32_t image_width,
                                                uint32_t image_height, uint32_t tile_size, const int32_t* tile_offsets,
                                                const int32_t* flatten_ids, int64_t n_isects, float* renders,
                                                float* alphas, int32_t* last_ids, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(means2d && conics && colors && opacities && tile_offsets && renders && alphas && last_ids,
                  "rasterize_to_pixels_3dgs_fwd: null pointer");
    LFS_CHECK_ARG(flatten_ids || n_isects == 0, "rasterize_to_pixels_3dgs_fwd: null flatten_ids");
    LFS_UNSUPPORTED(tile_size != kTile, "rasterize_to_pixels_3dgs_fwd: tile_size %u (only 16)", tile_size);
    LFS_UNSUPPORTED(channels < 1 || channels > 40, "rasterize_to_pixels_3dgs_fwd: %u channels (1..40)", channels);
    (void)N;
    if (C == 0 || image_width == 0 || image_height == 0)
        return LFS_OK;
    const uint32_t tw = div_up(image_width, kTile), th = div_up(image_height, kTile);
    const dim3 grid(tw, th, C);
#define LFS_R2D_FWD(CP)                                                                                                  \
    k_raster2d_fwd<CP><<<grid, kTilePix, 0, stream>>>(reinterpret_cast<const float2*>(means2d), conics, colors, opacities, \
                                                      backgrounds, masks, (int)channels, image_width, image_height, tw, th, \
                                                      tile_offsets, flatten_ids, n_isects, C, renders, alphas, last_ids)
    switch (pad_channels(channels)) {
    case 4: LFS_R2D_FWD(4); break;
    case 8: LFS_R2D_FWD(8); break;
    case 16: LFS_R2D_FWD(16); break;
    default: LFS_R2D_FWD(40); break;
    }
#undef LFS_R2D_FWD
    LFS_LAUNCH_OK("k_raster2d_fwd");
    return LFS_OK;
}

extern "C" int lfs_rasterize_to_pixels_3dgs_bwd(const float* means2d, const float* conics, const float* colors,
                                                const float* opacities, const float* backgrounds, const uint8_t* masks,
                                                uint32_t C, uint32_t N, uint32_t channels, uint32_t image_width,
                                                uint32_t image_height, uint32_t tile_size, const int32_t* tile_offsets,
                                                const int32_t* flatten_ids, int64_t n_isects, const float* render_alphas,
                                                const int32_t* last_ids, const float* v_render_colors,
                                                const float* v_render_alphas, float* v_means2d_abs, float* v_means2d,
                                                float* v_conics, float* v_colors, float* v_opacities, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(means2d && conics && colors && opacities && tile_offsets && render_alphas && last_ids && v_render_colors &&
                      v_render_alphas && v_means2d && v_conics && v_colors && v_opacities,
                  "rasterize_to_pixels_3dgs_bwd: null pointer");
    LFS_UNSUPPORTED(tile_size != kTile, "rasterize_to_pixels_3dgs_bwd: tile_size %u (only 16)", tile_size);
    LFS_UNSUPPORTED(channels < 1 || channels > 40, "rasterize_to_pixels_3dgs_bwd: %u channels (1..40)", channels);
    // the outputs are fully written: zero-filled, then accumulated (the reference ops allocate with at::zeros)
    const size_t G = (size_t)C * N;
    LFS_CUDA_OK(cudaMemsetAsync(v_means2d, 0, sizeof(float) * 2 * G, stream));
    LFS_CUDA_OK(cudaMemsetAsync(v_conics, 0, sizeof(float) * 3 * G, stream));
    LFS_CUDA_OK(cudaMemsetAsync(v_colors, 0, sizeof(float) * channels * G, stream));
    LFS_CUDA_OK(cudaMemsetAsync(v_opacities, 0, sizeof(float) * G, stream));
    if (v_means2d_abs)
        LFS_CUDA_OK(cudaMemsetAsync(v_means2d_abs, 0, sizeof(float) * 2 * G, stream));
    if (C == 0 || n_isects == 0 || image_width == 0 || image_height == 0)
        return LFS_OK;
    LFS_CHECK_ARG(flatten_ids, "rasterize_to_pixels_3dgs_bwd: null flatten_ids");
    const uint32_t tw = div_up(image_width, kTile), th = div_up(image_height, kTile);
    const dim3 grid(tw, th, C);
#define LFS_R2D_BWD(CP)                                                                                                  \
    k_raster2d_bwd<CP><<<grid, kTilePix, 0, stream>>>(                                                                   \
        reinterpret_cast<const float2*>(means2d), conics, colors, opacities, backgrounds, masks, (int)channels,          \
        image_width, image_height, tw, th, tile_offsets, flatten_ids, render_alphas, last_ids, v_render_colors,          \
        v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities)
    switch (pad_channels(channels)) {
    case 4: LFS_R2D_BWD(4); break;
    case 8: LFS_R2D_BWD(8); break;
    case 16: LFS_R2D_BWD(16); break;
    default: LFS_R2D_BWD(40); break;
    }
#undef LFS_R2D_BWD
    LFS_LAUNCH_OK("k_raster2d_bwd");
    return LFS_OK;
}
