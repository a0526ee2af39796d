// lfs_b200 -- camera models of the 3DGUT path: perfect pinhole, OpenCV pinhole (radial 6 / tangential 2 / thin-prism 4),
// OpenCV fisheye (4 radial), global and rolling shutter.  Behaviour follows the reference's gsplat/Cameras.cuh
// (pinhole :416-470, OpenCV pinhole :473-757, fisheye :760-1024, shutter pose / rolling-shutter iteration :253-413); the
// code is organised differently: one plain struct filled once per camera (host or device) instead of a CRTP class
// hierarchy, so that the projection kernel and the per-pixel ray kernel take it by value.
#pragma once
#include "common.cuh"
#include <float.h>
#include <math.h>

namespace lfs {

// shutter types: values of the reference's global ShutterType enum (gsplat/Cameras.h:16-22)
enum : int { kRsTopBottom = 0, kRsLeftRight = 1, kRsBottomTop = 2, kRsRightLeft = 3, kRsGlobal = 4 };

struct CamModel {
    int model;          // LFS_PINHOLE / LFS_FISHEYE
    int shutter;        // kRs*
    int distorted;      // pinhole: any of radial / tangential / thin-prism given
    uint32_t width, height;
    float fx, fy, cx, cy;
    float k[6], p[2], s[4]; // pinhole distortion (k: radial, p: tangential, s: thin prism); fisheye uses k[0..3]
    float fish_max_angle, fish_back1; // fisheye: FOV limit, slope of the crude backward polynomial
    quat4 q0, q1;       // world->camera rotation at the start / end of the frame (glm::quat_cast of the row-major pose)
    f3 t0, t1;
};

// ---- quaternion helpers with glm's semantics (the reference interpolates poses with glm::slerp) ---------------------
__host__ __device__ __forceinline__ quat4 quat_conj(const quat4 q) { return quat4{q.w, -q.x, -q.y, -q.z}; }
__host__ __device__ __forceinline__ quat4 quat_inverse(const quat4 q) { // glm::inverse = conjugate / dot(q, q)
    const float d = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    return quat4{q.w / d, -q.x / d, -q.y / d, -q.z / d};
}
// glm::slerp.  The reference builds its kernels with --use_fast_math (gsplat/CMakeLists.txt:76), which turns glm's sin()
// and the division into sin.approx / div.approx: for the small angle between the two shutter poses sin.approx's ABSOLUTE
// error (2^-21.4) is a RELATIVE error of ~2e-5 of sin(angle), i.e. the interpolated quaternion comes out with a norm that
// is off by that much, and mat3_cast / rotate of it move a ray by ~1e-4 of a Gaussian's extent.
//   FAST = true  reproduces that arithmetic (device only).  Used for the per-pixel rays of the blend, where it is the
//                difference between 83 % and 99.7 % of the rolling-shutter render inside the 1e-4 band of the reference
//                build (measured, tests/test_gpu_boundary.py).
//   FAST = false exact sinf.  Used by the UT projection: there the sigma points land on different rows, hence different
//                poses, hence independent sin.approx errors, which the +-99 sigma-point weights amplify to pixels -- noise
//                that two implementations only share if they are bit-identical; the exact variant differs from the reference
//                by the reference's noise alone (measured: 98 % of the means inside 1e-4, against 61 % with FAST).
template <bool FAST>
__host__ __device__ inline quat4 quat_slerp(const quat4 x, const quat4 y, const float a) {
    float c = x.w * y.w + x.x * y.x + x.y * y.y + x.z * y.z;
    quat4 z = y;
    if (c < 0.f) { // take the short way round
        z = quat4{-y.w, -y.x, -y.y, -y.z};
        c = -c;
    }
    if (c > 1.f - FLT_EPSILON) { // nearly parallel: glm::mix per component (glm does not renormalise)
        const float b = 1.f - a;
        return quat4{x.w * b + z.w * a, x.x * b + z.x * a, x.y * b + z.y * a, x.z * b + z.z * a};
    }
    const float ang = acosf(c);
#ifdef __CUDA_ARCH__
    if (FAST) {
        const float s0 = __sinf((1.f - a) * ang), s1 = __sinf(a * ang), sd = __sinf(ang);
        return quat4{__fdividef(s0 * x.w + s1 * z.w, sd), __fdividef(s0 * x.x + s1 * z.x, sd),
                     __fdividef(s0 * x.y + s1 * z.y, sd), __fdividef(s0 * x.z + s1 * z.z, sd)};
    }
#endif
    const float s0 = sinf((1.f - a) * ang), s1 = sinf(a * ang), sd = sinf(ang);
    return quat4{(s0 * x.w + s1 * z.w) / sd, (s0 * x.x + s1 * z.x) / sd, (s0 * x.y + s1 * z.y) / sd,
                 (s0 * x.z + s1 * z.z) / sd};
}
// rotation matrix of a quaternion (glm::mat3_cast), rows R[0..2]
__host__ __device__ __forceinline__ void quat_to_mat3(const quat4 q, float R[9]) {
    const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xz = q.x * q.z, xy = q.x * q.y, yz = q.y * q.z,
                wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    R[0] = 1.f - 2.f * (yy + zz), R[1] = 2.f * (xy - wz), R[2] = 2.f * (xz + wy);
    R[3] = 2.f * (xy + wz), R[4] = 1.f - 2.f * (xx + zz), R[5] = 2.f * (yz - wx);
    R[6] = 2.f * (xz - wy), R[7] = 2.f * (yz + wx), R[8] = 1.f - 2.f * (xx + yy);
}

// ---- fisheye set-up (Cameras.cuh:760-892) ---------------------------------------------------------------------------
__host__ __device__ inline float fisheye_max_angle_cubic(const float a, const float b, const float c) {
    // smallest positive root of 1 + a x + b x^2 + c x^3 = 0 (x = theta^2), FLT_MAX if none
    const float INF = FLT_MAX;
    if (c == 0.0f) {
        if (b == 0.0f)
            return a >= 0.0f ? INF : -1.0f / a;
        float delta = a * a - 4.0f * b;
        if (delta >= 0.0f) {
            delta = sqrtf(delta) - a;
            if (delta > 0.0f)
                return 2.0f / delta;
        }
        return INF;
    }
    const float boc = b / c, boc2 = boc * boc;
    const float t1 = (9.0f * a * boc - 2.0f * b * boc2 - 27.0f) / c;
    const float t2 = 3.0f * a / c - boc2;
    const float delta = t1 * t1 + 4.0f * t2 * t2 * t2;
    if (delta >= 0.0f) {
        const float d2 = sqrtf(delta);
        const float cr = cbrtf((d2 + t1) / 2.0f);
        if (cr != 0.0f) {
            const float soln = (cr - (t2 / cr) - boc) / 3.0f;
            if (soln > 0.0f)
                return soln;
        }
        return INF;
    }
    const float theta = atan2f(sqrtf(-delta), t1) / 3.0f;
    const float two_third_pi = 2.0f * 3.14159265358979323846f / 3.0f;
    const float t3 = 2.0f * sqrtf(-t2);
    float soln = INF;
    for (int i = -1; i <= 1; ++i) {
        const float s = (t3 * cosf(theta + (float)i * two_third_pi) - boc) / 3.0f;
        if (s > 0.0f)
            soln = fminf(soln, s);
    }
    return soln;
}
__host__ __device__ __forceinline__ float fish_fwd(const CamModel& c, const float th) { // theta (1 + k1 th^2 + ... )
    const float t2 = th * th;
    return th * (1.f + t2 * (c.k[0] + t2 * (c.k[1] + t2 * (c.k[2] + t2 * c.k[3]))));
}
__host__ __device__ __forceinline__ float fish_dfwd(const CamModel& c, const float th) {
    const float t2 = th * th;
    return 1.f + t2 * (3.f * c.k[0] + t2 * (5.f * c.k[1] + t2 * (7.f * c.k[2] + t2 * 9.f * c.k[3])));
}
__host__ __device__ inline void fisheye_setup(CamModel& c) {
    const float mdx = fmaxf((float)c.width - c.cx, c.cx), mdy = fmaxf((float)c.height - c.cy, c.cy);
    const float max_radius_pixels = sqrtf(mdx * mdx + mdy * mdy);
    float max_angle;
    if (c.k[3] == 0.f) {
        max_angle = sqrtf(fisheye_max_angle_cubic(3.f * c.k[0], 5.f * c.k[1], 7.f * c.k[2]));
    } else { // Newton on d(forward)/d(theta) = 0 from 1.57 (20 iterations, |dx| < 1e-6 stops)
        float x = 1.57f;
        bool converged = false;
        for (int j = 0; j < 20; ++j) {
            const float x2 = x * x;
            const float ddf = x * (6.f * c.k[0] + x2 * (20.f * c.k[1] + x2 * (42.f * c.k[2] + x2 * 72.f * c.k[3])));
            const float dx = fish_dfwd(c, x) / ddf;
            x -= dx;
            if (fabsf(dx) < 1e-6f) {
                converged = true;
                break;
            }
        }
        max_angle = (!converged || x <= 0.f) ? FLT_MAX : x;
    }
    max_angle = fminf(max_angle, fmaxf(max_radius_pixels / c.fx, max_radius_pixels / c.fy));
    const float max_normalized_dist = fmaxf((float)c.width / 2.f / c.fx, (float)c.height / 2.f / c.fy);
    c.fish_max_angle = max_angle;
    c.fish_back1 = max_angle / max_normalized_dist;
}

// ---- model construction ---------------------------------------------------------------------------------------------
// viewmat0 / viewmat1: row-major [4,4] world->camera at the start / end of the frame (viewmat1 may be null).
__host__ __device__ inline CamModel make_cam_model(const float* vm0, const float* vm1, const float* K, const uint32_t width,
                                                   const uint32_t height, const int model, const int shutter,
                                                   const float* radial, const float* tangential, const float* prism) {
    CamModel c;
    c.model = model, c.shutter = shutter;
    c.width = width, c.height = height;
    c.fx = K[0], c.fy = K[4], c.cx = K[2], c.cy = K[5];
    for (int i = 0; i < 6; ++i) c.k[i] = 0.f;
    c.p[0] = c.p[1] = 0.f;
    for (int i = 0; i < 4; ++i) c.s[i] = 0.f;
    c.distorted = 0;
    c.fish_max_angle = FLT_MAX, c.fish_back1 = 0.f;
    if (model == LFS_FISHEYE) {
        if (radial)
            for (int i = 0; i < 4; ++i) c.k[i] = radial[i];
        fisheye_setup(c);
    } else {
        c.distorted = (radial || tangential || prism) ? 1 : 0;
        if (radial)
            for (int i = 0; i < 6; ++i) c.k[i] = radial[i];
        if (tangential)
            c.p[0] = tangential[0], c.p[1] = tangential[1];
        if (prism)
            for (int i = 0; i < 4; ++i) c.s[i] = prism[i];
    }
    c.q0 = quat_from_rowmajor_rot(vm0);
    c.t0 = f3{vm0[3], vm0[7], vm0[11]};
    if (vm1) {
        c.q1 = quat_from_rowmajor_rot(vm1);
        c.t1 = f3{vm1[3], vm1[7], vm1[11]};
    } else {
        c.q1 = c.q0, c.t1 = c.t0;
    }
    return c;
}

__host__ __device__ __forceinline__ bool in_image_margin(const CamModel& c, const float u, const float v,
                                                         const float margin_factor) {
    const float MX = (float)c.width * margin_factor, MY = (float)c.height * margin_factor;
    return (-MX <= u) && (u < (float)c.width + MX) && (-MY <= v) && (v < (float)c.height + MY);
}

// OpenCV distortion of a normalised point (Cameras.cuh:504-533): returns icD, writes delta
__host__ __device__ __forceinline__ float opencv_distortion(const CamModel& c, const float x, const float y, float& dxo,
                                                            float& dyo) {
    const float xx = x * x, yy = y * y, r2 = xx + yy;
    const float a1 = 2.f * x * y, a2 = r2 + 2.f * xx, a3 = r2 + 2.f * yy;
    const float num = 1.f + r2 * (c.k[0] + r2 * (c.k[1] + r2 * c.k[2]));
    const float den = 1.f + r2 * (c.k[3] + r2 * (c.k[4] + r2 * c.k[5]));
    dxo = c.p[0] * a1 + c.p[1] * a2 + r2 * (c.s[0] + r2 * c.s[1]);
    dyo = c.p[0] * a3 + c.p[1] * a1 + r2 * (c.s[2] + r2 * c.s[3]);
    return num / den;
}

// camera-space point -> image point; false if behind the camera, flipped by the distortion, outside FOV / image + margin
__host__ __device__ inline bool cam_project(const CamModel& c, const f3 pc, const float margin_factor, float& u, float& v) {
    u = v = 0.f;
    if (pc.z <= 0.f)
        return false;
    if (c.model == LFS_FISHEYE) {
        const float ax = fabsf(pc.x), ay = fabsf(pc.y);
        const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
        float nrm = 0.f; // numerically stable |(x, y)|
        if (mx > 0.f) {
            const float rr = mn / mx;
            nrm = mx * sqrtf(1.f + rr * rr);
        }
        if (nrm <= 0.f)
            nrm = FLT_EPSILON;
        const float theta_full = atan2f(nrm, pc.z);
        const float theta = theta_full < c.fish_max_angle ? theta_full : c.fish_max_angle;
        const float delta = fish_fwd(c, theta) / nrm;
        if (delta <= 0.f)
            return false;
        u = c.fx * delta * pc.x + c.cx;
        v = c.fy * delta * pc.y + c.cy;
        return in_image_margin(c, u, v, margin_factor) && theta <= c.fish_max_angle; // as written in the reference
    }
    const float x = pc.x / pc.z, y = pc.y / pc.z;
    if (!c.distorted) {
        u = x * c.fx + c.cx;
        v = y * c.fy + c.cy;
        return in_image_margin(c, u, v, margin_factor);
    }
    float dx, dy;
    const float icD = opencv_distortion(c, x, y, dx, dy);
    u = (icD * x + dx) * c.fx + c.cx;
    v = (icD * y + dy) * c.fy + c.cy;
    return (icD > 0.8f) && in_image_margin(c, u, v, margin_factor);
}

// image point -> unit camera ray; false when the inverse does not converge / leaves the valid cone
__host__ __device__ inline bool cam_unproject(const CamModel& c, const float u, const float v, f3& ray) {
    const float xd = (u - c.cx) / c.fx, yd = (v - c.cy) / c.fy;
    if (c.model == LFS_FISHEYE) {
        const float delta = sqrtf(xd * xd + yd * yd);
        float th = c.fish_back1 * delta; // crude start (approx_backward_poly = {0, max_angle / max_normalized_dist})
        bool converged = false;
        for (int j = 0; j < 20; ++j) {
            const float dth = (fish_fwd(c, th) - delta) / fish_dfwd(c, th);
            th -= dth;
            if (fabsf(dth) < 1e-6f) {
                converged = true;
                break;
            }
        }
        if (th < 0.f || th >= c.fish_max_angle || !converged) {
            ray = f3{0.f, 0.f, 1.f};
            return false;
        }
        if (delta >= 1e-6f) {
            const float sf = sinf(th) / delta;
            ray = f3{sf * xd, sf * yd, cosf(th)};
        } else {
            ray = f3{0.f, 0.f, 1.f};
        }
        return true;
    }
    float x = xd, y = yd;
    bool ok = true;
    if (c.distorted) { // Newton undistortion, at most 5 iterations (Cameras.cuh:698-740)
        ok = false;
        for (int it = 0; it < 5; ++it) {
            const float r = x * x + y * y, r2 = r * r;
            const float alpha = 1.0f + r * (c.k[0] + r * (c.k[1] + r * c.k[2]));
            const float beta = 1.0f + r * (c.k[3] + r * (c.k[4] + r * c.k[5]));
            const float d = alpha / beta;
            if (d <= 0.f)
                break;
            const float fx_ = d * x + 2.f * c.p[0] * x * y + c.p[1] * (r + 2.f * x * x) + c.s[0] * r + c.s[1] * r2 - xd;
            const float fy_ = d * y + 2.f * c.p[1] * x * y + c.p[0] * (r + 2.f * y * y) + c.s[2] * r + c.s[3] * r2 - yd;
            const float alpha_r = c.k[0] + r * (2.0f * c.k[1] + r * (3.0f * c.k[2]));
            const float beta_r = c.k[3] + r * (2.0f * c.k[4] + r * (3.0f * c.k[5]));
            const float d_r = (alpha_r * beta - alpha * beta_r) / (beta * beta);
            const float d_x = 2.0f * x * d_r, d_y = 2.0f * y * d_r;
            const float fx_x = d + d_x * x + 2.0f * c.p[0] * y + 6.0f * c.p[1] * x + 2.0f * x * (c.s[0] + 2.0f * c.s[1] * r);
            const float fx_y = d_y * x + 2.0f * c.p[0] * x + 2.0f * c.p[1] * y + 2.0f * y * (c.s[0] + 2.0f * c.s[1] * r);
            const float fy_x = d_x * y + 2.0f * c.p[1] * y + 2.0f * c.p[0] * x + 2.0f * x * (c.s[2] + 2.0f * c.s[3] * r);
            const float fy_y = d + d_y * y + 2.0f * c.p[1] * x + 6.0f * c.p[0] * y + 2.0f * y * (c.s[2] + 2.0f * c.s[3] * r);
            const float det = fx_y * fy_x - fx_x * fy_y;
            if (fabsf(det) < 1e-6f)
                break;
            const float dx = (fx_ * fy_y - fy_ * fx_y) / det, dy = (fy_ * fx_x - fx_ * fy_x) / det;
            x += dx;
            y += dy;
            if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) {
                ok = true;
                break;
            }
        }
    }
    const float il = 1.0f / sqrtf(x * x + y * y + 1.f);
    ray = f3{x * il, y * il, il};
    return ok;
}

// relative frame time of an image point (Cameras.cuh:293-318)
__host__ __device__ __forceinline__ float shutter_time(const CamModel& c, const float u, const float v) {
    switch (c.shutter) {
    case kRsTopBottom: return floorf(v) / (float)(c.height - 1);
    case kRsLeftRight: return floorf(u) / (float)(c.width - 1);
    case kRsBottomTop: return ((float)c.height - ceilf(v)) / (float)(c.height - 1);
    case kRsRightLeft: return ((float)c.width - ceilf(u)) / (float)(c.width - 1);
    default: return 0.f;
    }
}
template <bool FAST>
__host__ __device__ __forceinline__ void shutter_pose(const CamModel& c, const float t, quat4& q, f3& tr) {
    tr = f3{(1.f - t) * c.t0.x + t * c.t1.x, (1.f - t) * c.t0.y + t * c.t1.y, (1.f - t) * c.t0.z + t * c.t1.z};
    q = quat_slerp<FAST>(c.q0, c.q1, t);
}

// world point -> image point through the (rolling) shutter pose: the reference's 10 fixed-point iterations
// (Cameras.cuh:346-413)
__host__ __device__ inline bool world_to_image(const CamModel& c, const f3 pw, const float margin_factor, float& u, float& v) {
    float us, vs;
    const bool valid_start = cam_project(c, quat_rotate(c.q0, pw) + c.t0, margin_factor, us, vs);
    if (c.shutter == kRsGlobal) {
        u = us, v = vs;
        return valid_start;
    }
    float ue, ve;
    const bool valid_end = cam_project(c, quat_rotate(c.q1, pw) + c.t1, margin_factor, ue, ve);
    float pu, pv;
    if (valid_start) {
        pu = us, pv = vs;
    } else if (valid_end) {
        pu = ue, pv = ve;
    } else {
        u = ue, v = ve;
        return false;
    }
    for (int j = 0; j < 10; ++j) {
        quat4 q;
        f3 tr;
        shutter_pose<false>(c, shutter_time(c, pu, pv), q, tr);
        float nu, nv;
        cam_project(c, quat_rotate(q, pw) + tr, margin_factor, nu, nv);
        pu = nu, pv = nv;
    }
    u = pu, v = pv;
    return true;
}

// pixel -> world ray through the shutter pose of that pixel (Cameras.cuh:322-339, :253-266)
__host__ __device__ inline bool pixel_to_world_ray(const CamModel& c, const float u, const float v, f3& org, f3& dir) {
    f3 cr;
    if (!cam_unproject(c, u, v, cr)) {
        org = dir = f3{0.f, 0.f, 0.f};
        return false;
    }
    quat4 q;
    f3 tr;
    shutter_pose<true>(c, shutter_time(c, u, v), q, tr);
    float R[9];
    quat_to_mat3(quat_inverse(q), R);
    org = f3{-(R[0] * tr.x + R[1] * tr.y + R[2] * tr.z), -(R[3] * tr.x + R[4] * tr.y + R[5] * tr.z),
             -(R[6] * tr.x + R[7] * tr.y + R[8] * tr.z)};
    dir = f3{R[0] * cr.x + R[1] * cr.y + R[2] * cr.z, R[3] * cr.x + R[4] * cr.y + R[5] * cr.z,
             R[6] * cr.x + R[7] * cr.y + R[8] * cr.z};
    return true;
}

} // namespace lfs
