// lfs_b200 -- unscented-transform projection of one Gaussian (PINHOLE / GLOBAL shutter), device side.
// Own formulation of the algorithm in the reference's gsplat/ProjectionUT3DGSFused.cu:47-203 and
// gsplat/Cameras.cuh:1034-1150; the arithmetic order of the depth / mean / covariance sums follows the
// reference so results agree to fp32 rounding (radii +-1 is the reference's own tolerance).
#pragma once
#include "cameras.cuh"
#include "common.cuh"
#include <float.h>
#include <math.h>

namespace lfs {

__host__ __device__ inline ViewCam make_viewcam(const float* vm, const float* K, int width, int height) {
    ViewCam c;
    c.R[0] = vm[0], c.R[1] = vm[1], c.R[2] = vm[2];
    c.R[3] = vm[4], c.R[4] = vm[5], c.R[5] = vm[6];
    c.R[6] = vm[8], c.R[7] = vm[9], c.R[8] = vm[10];
    c.t[0] = vm[3], c.t[1] = vm[7], c.t[2] = vm[11];
    const quat4 q = quat_from_rowmajor_rot(vm);
    c.q[0] = q.w, c.q[1] = q.x, c.q[2] = q.y, c.q[3] = q.z;
    // camera centre -R^T t
    c.org[0] = -(c.R[0] * c.t[0] + c.R[3] * c.t[1] + c.R[6] * c.t[2]);
    c.org[1] = -(c.R[1] * c.t[0] + c.R[4] * c.t[1] + c.R[7] * c.t[2]);
    c.org[2] = -(c.R[2] * c.t[0] + c.R[5] * c.t[1] + c.R[8] * c.t[2]);
    c.fx = K[0], c.fy = K[4], c.cx = K[2], c.cy = K[5];
    c.width = width, c.height = height;
    c.tile_w = (width + kTile - 1) / kTile;
    c.tile_h = (height + kTile - 1) / kTile;
    return c;
}

struct UTOut {
    bool ok;
    float mx, my;       // 2-D mean (pixels)
    float rx, ry;       // integer-valued radii
    float depth;        // camera-space z of the centre
    float c00, c01, c11; // conic (inverse blurred covariance)
    float comp;         // blur compensation
};

__device__ __forceinline__ UTOut ut_project_pinhole(const ViewCam& cam, f3 mean, float4 quat_wxyz, f3 scale,
                                                    bool has_opacity, float opacity, float eps2d,
                                                    float near_plane, float far_plane, float radius_clip,
                                                    const lfs_ut_params ut) {
    UTOut o;
    o.ok = false;
    const quat4 cq = quat4{cam.q[0], cam.q[1], cam.q[2], cam.q[3]};
    const f3 ct = mk3(cam.t[0], cam.t[1], cam.t[2]);

    // centre in camera space through the quaternion route (Cameras.cuh:76-77)
    const f3 mean_c = quat_rotate(cq, mean) + ct;
    o.depth = mean_c.z;
    if (mean_c.z < near_plane || mean_c.z > far_plane)
        return o;

    // glm::normalize(quat)
    float qw = quat_wxyz.x, qx = quat_wxyz.y, qy = quat_wxyz.z, qz = quat_wxyz.w;
    {
        const float len = sqrtf((qw * qw + qx * qx) + (qy * qy + qz * qz));
        if (len <= 0.f) {
            qw = 1.f;
            qx = qy = qz = 0.f;
        } else {
            const float il = 1.0f / len;
            qw *= il, qx *= il, qy *= il, qz *= il;
        }
    }
    // columns of R = mat3_cast(q)
    const float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz,
                qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    const f3 col0 = mk3(1.f - 2.f * (qyy + qzz), 2.f * (qxy + qwz), 2.f * (qxz - qwy));
    const f3 col1 = mk3(2.f * (qxy - qwz), 1.f - 2.f * (qxx + qzz), 2.f * (qyz + qwx));
    const f3 col2 = mk3(2.f * (qxz + qwy), 2.f * (qyz - qwx), 1.f - 2.f * (qxx + qyy));

    const float D = 3.0f;
    const float lambda = ut.alpha * ut.alpha * (D + ut.kappa) - D;
    const float sq = sqrtf(D + lambda);
    const float w_m0 = lambda / (D + lambda);
    const float w_c0 = lambda / (D + lambda) + (1.f - ut.alpha * ut.alpha + ut.beta);
    const float w_i = 1.f / (2.f * (D + lambda));

    const f3 delta[3] = {col0 * (sq * scale.x), col1 * (sq * scale.y), col2 * (sq * scale.z)};
    const float MX = (float)cam.width * ut.in_image_margin_factor;
    const float MY = (float)cam.height * ut.in_image_margin_factor;

    float px[7], py[7];
    float mx = 0.f, my = 0.f;
    bool any_valid = false;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        f3 p = mean;
        if (i >= 1 && i <= 3)
            p = mean + delta[i - 1];
        else if (i >= 4)
            p = mean - delta[i - 4];
        const f3 pc = quat_rotate(cq, p) + ct;
        float u = 0.f, v = 0.f;
        bool pv = false;
        if (pc.z > 0.f) {
            u = (pc.x / pc.z) * cam.fx + cam.cx;
            v = (pc.y / pc.z) * cam.fy + cam.cy;
            pv = (-MX <= u) && (u < (float)cam.width + MX) && (-MY <= v) && (v < (float)cam.height + MY);
        }
        if (ut.require_all_sigma_points_valid) {
            if (!pv)
                return o;
        } else {
            any_valid |= pv;
        }
        px[i] = u;
        py[i] = v;
        const float w = (i == 0) ? w_m0 : w_i;
        mx += w * u;
        my += w * v;
    }
    if (!ut.require_all_sigma_points_valid && !any_valid)
        return o;

    float cxx = 0.f, cxy = 0.f, cyy = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float w = (i == 0) ? w_c0 : w_i;
        const float dx = px[i] - mx, dy = py[i] - my;
        cxx += w * (dx * dx);
        cxy += w * (dx * dy);
        cyy += w * (dy * dy);
    }
    // add_blur (Utils.cuh:171-179)
    const float det_orig = cxx * cyy - cxy * cxy;
    cxx += eps2d;
    cyy += eps2d;
    const float det = cxx * cyy - cxy * cxy;
    const float compensation = sqrtf(fmaxf(0.f, det_orig / det));
    if (!(det > 0.f))
        return o;
    const float ood = 1.0f / det;

    float extend = 3.33f;
    if (has_opacity) {
        const float op = opacity * compensation;
        if (op < (1.f / 255.f))
            return o;
        extend = fminf(extend, sqrtf(2.0f * __logf(op * 255.f)));
    }
    const float b = 0.5f * (cxx + cyy);
    const float tmp = sqrtf(fmaxf(0.01f, b * b - det));
    const float v1 = b + tmp;
    const float r1 = extend * sqrtf(v1);
    const float radius_x = ceilf(fminf(extend * sqrtf(cxx), r1));
    const float radius_y = ceilf(fminf(extend * sqrtf(cyy), r1));
    if (radius_x <= radius_clip && radius_y <= radius_clip)
        return o;
    if (mx + radius_x <= 0.f || mx - radius_x >= (float)cam.width || my + radius_y <= 0.f ||
        my - radius_y >= (float)cam.height)
        return o;

    o.ok = true;
    o.mx = mx, o.my = my;
    o.rx = radius_x, o.ry = radius_y;
    o.c00 = cyy * ood;
    o.c01 = -cxy * ood;
    o.c11 = cxx * ood;
    o.comp = compensation;
    return o;
}

// The same projection for every camera model / shutter of the reference (OpenCV pinhole distortion, fisheye, rolling
// shutter): the seven sigma points go through world_to_image (camera model + rolling-shutter fixed point), the depth is
// taken at the centre-of-frame pose (ProjectionUT3DGSFused.cu:74-78).  Same summation order as ut_project_pinhole.
__device__ inline UTOut ut_project_general(const CamModel& cam, f3 mean, float4 quat_wxyz, f3 scale, bool has_opacity,
                                           float opacity, float eps2d, float near_plane, float far_plane,
                                           float radius_clip, const lfs_ut_params ut) {
    UTOut o;
    o.ok = false;
    quat4 qc;
    f3 tc;
    shutter_pose<true>(cam, 0.5f, qc, tc); // one pose, the same input for every Gaussian: reproduces the reference's depth bits
    const f3 mean_c = quat_rotate(qc, mean) + tc;
    o.depth = mean_c.z;
    if (mean_c.z < near_plane || mean_c.z > far_plane)
        return o;
    float qw = quat_wxyz.x, qx = quat_wxyz.y, qy = quat_wxyz.z, qz = quat_wxyz.w;
    {
        const float len = sqrtf((qw * qw + qx * qx) + (qy * qy + qz * qz));
        if (len <= 0.f) {
            qw = 1.f;
            qx = qy = qz = 0.f;
        } else {
            const float il = 1.0f / len;
            qw *= il, qx *= il, qy *= il, qz *= il;
        }
    }
    const float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz,
                qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    const f3 col0 = mk3(1.f - 2.f * (qyy + qzz), 2.f * (qxy + qwz), 2.f * (qxz - qwy));
    const f3 col1 = mk3(2.f * (qxy - qwz), 1.f - 2.f * (qxx + qzz), 2.f * (qyz + qwx));
    const f3 col2 = mk3(2.f * (qxz + qwy), 2.f * (qyz - qwx), 1.f - 2.f * (qxx + qyy));
    const float D = 3.0f;
    const float lambda = ut.alpha * ut.alpha * (D + ut.kappa) - D;
    const float sq = sqrtf(D + lambda);
    const float w_m0 = lambda / (D + lambda);
    const float w_c0 = lambda / (D + lambda) + (1.f - ut.alpha * ut.alpha + ut.beta);
    const float w_i = 1.f / (2.f * (D + lambda));
    const f3 delta[3] = {col0 * (sq * scale.x), col1 * (sq * scale.y), col2 * (sq * scale.z)};
    float px[7], py[7];
    float mx = 0.f, my = 0.f;
    bool any_valid = false;
    for (int i = 0; i < 7; ++i) {
        f3 p = mean;
        if (i >= 1 && i <= 3)
            p = mean + delta[i - 1];
        else if (i >= 4)
            p = mean - delta[i - 4];
        float u, v;
        const bool pv = world_to_image(cam, p, ut.in_image_margin_factor, u, v);
        if (ut.require_all_sigma_points_valid) {
            if (!pv)
                return o;
        } else {
            any_valid |= pv;
        }
        px[i] = u;
        py[i] = v;
        const float w = (i == 0) ? w_m0 : w_i;
        mx += w * u;
        my += w * v;
    }
    if (!ut.require_all_sigma_points_valid && !any_valid)
        return o;
    float cxx = 0.f, cxy = 0.f, cyy = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float w = (i == 0) ? w_c0 : w_i;
        const float dx = px[i] - mx, dy = py[i] - my;
        cxx += w * (dx * dx);
        cxy += w * (dx * dy);
        cyy += w * (dy * dy);
    }
    const float det_orig = cxx * cyy - cxy * cxy;
    cxx += eps2d;
    cyy += eps2d;
    const float det = cxx * cyy - cxy * cxy;
    const float compensation = sqrtf(fmaxf(0.f, det_orig / det));
    if (!(det > 0.f))
        return o;
    const float ood = 1.0f / det;
    float extend = 3.33f;
    if (has_opacity) {
        const float op = opacity * compensation;
        if (op < (1.f / 255.f))
            return o;
        extend = fminf(extend, sqrtf(2.0f * __logf(op * 255.f)));
    }
    const float b = 0.5f * (cxx + cyy);
    const float tmp = sqrtf(fmaxf(0.01f, b * b - det));
    const float r1 = extend * sqrtf(b + tmp);
    const float radius_x = ceilf(fminf(extend * sqrtf(cxx), r1));
    const float radius_y = ceilf(fminf(extend * sqrtf(cyy), r1));
    if (radius_x <= radius_clip && radius_y <= radius_clip)
        return o;
    if (mx + radius_x <= 0.f || mx - radius_x >= (float)cam.width || my + radius_y <= 0.f ||
        my - radius_y >= (float)cam.height)
        return o;
    o.ok = true;
    o.mx = mx, o.my = my;
    o.rx = radius_x, o.ry = radius_y;
    o.c00 = cyy * ood;
    o.c01 = -cxy * ood;
    o.c11 = cxx * ood;
    o.comp = compensation;
    return o;
}

// Tile rectangle of a projected Gaussian: the reference's AABB rule (gsplat/IntersectTile.cu:65-76).
// float->uint conversions saturate (cvt.rzi.u32.f32), negative -> 0.
__device__ __forceinline__ void tile_rect(float mx, float my, float rx, float ry, float tile_size, uint32_t tile_w,
                                          uint32_t tile_h, uint32_t& x0, uint32_t& y0, uint32_t& x1,
                                          uint32_t& y1) {
    const float trx = rx / tile_size, try_ = ry / tile_size;
    const float tx = mx / tile_size, ty = my / tile_size;
    x0 = min(__float2uint_rz(floorf(tx - trx)), tile_w);
    y0 = min(__float2uint_rz(floorf(ty - try_)), tile_h);
    x1 = min(__float2uint_rz(ceilf(tx + trx)), tile_w);
    y1 = min(__float2uint_rz(ceilf(ty + try_)), tile_h);
}

} // namespace lfs
