/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 *
 * Type-generic body of the CPU oracle; included by lfs_oracle.c once with REAL=float (PFX=orc32_) and
 * once with REAL=double (PFX=orc64_).  Each function restates the reference algorithm and cites the
 * reference file:line it follows (paths relative to /root/reference).  Matrices are stored row-major
 * m[r][c]; glm in the reference is column-major m[c][r] -- translations are noted where it matters.
 *
 * Parity status: SH, tile intersection and quat->covar are pinned against the reference's own CPU
 * restatement tests/torch_impl.cpp (golden vectors under tests/golden/, generated in-container from the
 * unmodified reference file).  UT projection and the from-world blend fwd/bwd have no CPU reference in
 * the tree ("we don't have a reference rasterizer", tests/test_rasterization.cpp:96-98): for those the
 * oracle is pinned (a) on the GPU box against oracle/_ref/libgsplat_ref.so (the unmodified reference CUDA
 * kernels) and (b) here by finite differences of its own forward (tests/test_oracle_grad.py).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define ORC(name) CAT(PFX, name)
#define real REAL

#define R_SQRT(x) ((real)sqrt((double)(x)))
#define R_EXP(x) ((real)exp((double)(x)))
#define R_LOG(x) ((real)log((double)(x)))

/* ------------------------------------------------------------------------------------------------
 * small linear-algebra helpers
 * ---------------------------------------------------------------------------------------------- */
static inline void ORC(cross3)(const real a[3], const real b[3], real o[3]) {
    real x = a[1] * b[2] - b[1] * a[2];
    real y = a[2] * b[0] - b[2] * a[0];
    real z = a[0] * b[1] - b[0] * a[1];
    o[0] = x;
    o[1] = y;
    o[2] = z;
}
static inline real ORC(dot3)(const real a[3], const real b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* gsplat/Utils.cuh:80-102 quat_to_rotmat (wxyz, normalised inside). Row-major R. */
static inline void ORC(quat_to_rotmat)(const real q[4], real R[3][3]) {
    real w = q[0], x = q[1], y = q[2], z = q[3];
    real inv_norm = (real)1 / R_SQRT(x * x + y * y + z * z + w * w);
    x *= inv_norm;
    y *= inv_norm;
    z *= inv_norm;
    w *= inv_norm;
    real x2 = x * x, y2 = y * y, z2 = z * z;
    real xy = x * y, xz = x * z, yz = y * z;
    real wx = w * x, wy = w * y, wz = w * z;
    R[0][0] = 1 - 2 * (y2 + z2);
    R[1][0] = 2 * (xy + wz);
    R[2][0] = 2 * (xz - wy);
    R[0][1] = 2 * (xy - wz);
    R[1][1] = 1 - 2 * (x2 + z2);
    R[2][1] = 2 * (yz + wx);
    R[0][2] = 2 * (xz + wy);
    R[1][2] = 2 * (yz - wx);
    R[2][2] = 1 - 2 * (x2 + y2);
}

/* gsplat/Utils.cuh:104-126 quat_to_rotmat_vjp.  G[r][c] = dL/dR[r][c] (the reference's v_R[c][r]). */
static inline void ORC(quat_to_rotmat_vjp)(const real q[4], const real G[3][3], real v_quat[4]) {
    real w = q[0], x = q[1], y = q[2], z = q[3];
    real inv_norm = (real)1 / R_SQRT(x * x + y * y + z * z + w * w);
    x *= inv_norm;
    y *= inv_norm;
    z *= inv_norm;
    w *= inv_norm;
    real vq[4];
    vq[0] = 2 * (x * (G[2][1] - G[1][2]) + y * (G[0][2] - G[2][0]) + z * (G[1][0] - G[0][1]));
    vq[1] = 2 * (-2 * x * (G[1][1] + G[2][2]) + y * (G[1][0] + G[0][1]) + z * (G[2][0] + G[0][2]) +
                 w * (G[2][1] - G[1][2]));
    vq[2] = 2 * (x * (G[1][0] + G[0][1]) - 2 * y * (G[0][0] + G[2][2]) + z * (G[2][1] + G[1][2]) +
                 w * (G[0][2] - G[2][0]));
    vq[3] = 2 * (x * (G[2][0] + G[0][2]) + y * (G[2][1] + G[1][2]) - 2 * z * (G[0][0] + G[1][1]) +
                 w * (G[1][0] - G[0][1]));
    real qn[4] = {w, x, y, z};
    real d = vq[0] * qn[0] + vq[1] * qn[1] + vq[2] * qn[2] + vq[3] * qn[3];
    for (int k = 0; k < 4; ++k)
        v_quat[k] += (vq[k] - d * qn[k]) * inv_norm;
}

/* glm quaternion helpers used by gsplat/Cameras.cuh:33-70,253-280 (q stored w,x,y,z). */
static inline void ORC(quat_cast)(const real m[3][3] /* row-major rotation */, real q[4]) {
    /* glm::quat_cast on the column-major matrix built from the row-major se3: m_glm[c][r] = m[r][c] */
#define GM(c, r) m[r][c]
    real fx = GM(0, 0) - GM(1, 1) - GM(2, 2);
    real fy = GM(1, 1) - GM(0, 0) - GM(2, 2);
    real fz = GM(2, 2) - GM(0, 0) - GM(1, 1);
    real fw = GM(0, 0) + GM(1, 1) + GM(2, 2);
    int bi = 0;
    real fb = fw;
    if (fx > fb) {
        fb = fx;
        bi = 1;
    }
    if (fy > fb) {
        fb = fy;
        bi = 2;
    }
    if (fz > fb) {
        fb = fz;
        bi = 3;
    }
    real bv = R_SQRT(fb + 1) * (real)0.5;
    real mult = (real)0.25 / bv;
    switch (bi) {
    case 0:
        q[0] = bv;
        q[1] = (GM(1, 2) - GM(2, 1)) * mult;
        q[2] = (GM(2, 0) - GM(0, 2)) * mult;
        q[3] = (GM(0, 1) - GM(1, 0)) * mult;
        break;
    case 1:
        q[0] = (GM(1, 2) - GM(2, 1)) * mult;
        q[1] = bv;
        q[2] = (GM(0, 1) + GM(1, 0)) * mult;
        q[3] = (GM(2, 0) + GM(0, 2)) * mult;
        break;
    case 2:
        q[0] = (GM(2, 0) - GM(0, 2)) * mult;
        q[1] = (GM(0, 1) + GM(1, 0)) * mult;
        q[2] = bv;
        q[3] = (GM(1, 2) + GM(2, 1)) * mult;
        break;
    default:
        q[0] = (GM(0, 1) - GM(1, 0)) * mult;
        q[1] = (GM(2, 0) + GM(0, 2)) * mult;
        q[2] = (GM(1, 2) + GM(2, 1)) * mult;
        q[3] = bv;
        break;
    }
#undef GM
}
/* glm::rotate(q, v) = v + 2 (w (u x v) + u x (u x v)) */
static inline void ORC(quat_rotate)(const real q[4], const real v[3], real o[3]) {
    real u[3] = {q[1], q[2], q[3]};
    real uv[3], uuv[3];
    ORC(cross3)(u, v, uv);
    ORC(cross3)(u, uv, uuv);
    for (int k = 0; k < 3; ++k)
        o[k] = v[k] + ((uv[k] * q[0]) + uuv[k]) * 2;
}
/* glm::mat3_cast -> row-major R (R[r][c] = Result[c][r]) */
static inline void ORC(mat3_cast)(const real q[4], real R[3][3]) {
    real w = q[0], x = q[1], y = q[2], z = q[3];
    real qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y,
         qwz = w * z;
    R[0][0] = 1 - 2 * (qyy + qzz);
    R[1][0] = 2 * (qxy + qwz);
    R[2][0] = 2 * (qxz - qwy);
    R[0][1] = 2 * (qxy - qwz);
    R[1][1] = 1 - 2 * (qxx + qzz);
    R[2][1] = 2 * (qyz + qwx);
    R[0][2] = 2 * (qxz + qwy);
    R[1][2] = 2 * (qyz - qwx);
    R[2][2] = 1 - 2 * (qxx + qyy);
}

/* Camera pose of one view as the reference derives it for a GLOBAL shutter
 * (gsplat/Cameras.cuh:39-56 RollingShutterParameters, :268-280 interpolate_shutter_pose with
 * q_start == q_end: slerp takes the lerp branch, mix(a,a,.5) == a). */
typedef struct {
    real q[4];
    real t[3];
} ORC(pose_t);
static inline void ORC(pose_from_viewmat)(const real* vm /* [4,4] row-major */, ORC(pose_t) * p) {
    real Rm[3][3] = {{vm[0], vm[1], vm[2]}, {vm[4], vm[5], vm[6]}, {vm[8], vm[9], vm[10]}};
    ORC(quat_cast)(Rm, p->q);
    p->t[0] = vm[3];
    p->t[1] = vm[7];
    p->t[2] = vm[11];
}

/* ------------------------------------------------------------------------------------------------
 * (a3) UT projection, PINHOLE / GLOBAL shutter / no distortion.
 * Follows gsplat/ProjectionUT3DGSFused.cu:47-203, gsplat/Cameras.cuh:431-455 (pinhole projection +
 * margin test), :1034-1083 (sigma points), :1091-1150 (weighted mean / covariance),
 * gsplat/Utils.cuh:171-179 (add_blur).  Culled entries: radii = 0 and all other outputs 0
 * (the reference leaves them uninitialised).
 * ---------------------------------------------------------------------------------------------- */
void ORC(projection_ut)(int C, int N, const real* means, const real* quats, const real* scales,
                        const real* opacities /* nullable */, const real* viewmats, const real* Ks, int width,
                        int height, real eps2d, real near_plane, real far_plane, real radius_clip, real ut_alpha,
                        real ut_beta, real ut_kappa, real margin_factor, int require_all_valid, int32_t* radii,
                        real* means2d, real* depths, real* conics, real* compensations /* nullable */) {
    const real ALPHA_THRESHOLD = (real)1 / (real)255;
    for (int cid = 0; cid < C; ++cid) {
        ORC(pose_t) pose;
        ORC(pose_from_viewmat)(viewmats + 16 * cid, &pose);
        const real fx = Ks[cid * 9 + 0], fy = Ks[cid * 9 + 4], cx = Ks[cid * 9 + 2], cy = Ks[cid * 9 + 5];
#pragma omp parallel for schedule(static)
        for (int gid = 0; gid < N; ++gid) {
            const size_t idx = (size_t)cid * N + gid;
            radii[idx * 2] = radii[idx * 2 + 1] = 0;
            means2d[idx * 2] = means2d[idx * 2 + 1] = 0;
            depths[idx] = 0;
            conics[idx * 3] = conics[idx * 3 + 1] = conics[idx * 3 + 2] = 0;
            if (compensations)
                compensations[idx] = 0;

            const real* mean = means + 3 * gid;
            const real* scale = scales + 3 * gid;
            real quat[4] = {quats[4 * gid], quats[4 * gid + 1], quats[4 * gid + 2], quats[4 * gid + 3]};
            { /* glm::normalize(quat) */
                real len = R_SQRT((quat[1] * quat[1] + quat[2] * quat[2]) + (quat[3] * quat[3] + quat[0] * quat[0]));
                if (len <= 0) {
                    quat[0] = 1;
                    quat[1] = quat[2] = quat[3] = 0;
                } else {
                    real il = (real)1 / len;
                    for (int k = 0; k < 4; ++k)
                        quat[k] *= il;
                }
            }
            real mean_c[3];
            ORC(quat_rotate)(pose.q, mean, mean_c);
            for (int k = 0; k < 3; ++k)
                mean_c[k] += pose.t[k];
            if (mean_c[2] < near_plane || mean_c[2] > far_plane)
                continue;

            /* sigma points (Cameras.cuh:1034-1083) */
            const real D = 3;
            const real lambda = ut_alpha * ut_alpha * (D + ut_kappa) - D;
            real Rg[3][3];
            ORC(mat3_cast)(quat, Rg);
            real pts[7][3], w_mean[7], w_cov[7];
            for (int k = 0; k < 3; ++k)
                pts[0][k] = mean[k];
            const real sq = R_SQRT(D + lambda);
            for (int i = 0; i < 3; ++i) {
                for (int k = 0; k < 3; ++k) {
                    real delta = sq * scale[i] * Rg[k][i]; /* column i of R */
                    pts[i + 1][k] = mean[k] + delta;
                    pts[i + 4][k] = mean[k] - delta;
                }
            }
            w_mean[0] = lambda / (D + lambda);
            w_cov[0] = lambda / (D + lambda) + (1 - ut_alpha * ut_alpha + ut_beta);
            for (int i = 0; i < 6; ++i) {
                w_mean[i + 1] = 1 / (2 * (D + lambda));
                w_cov[i + 1] = 1 / (2 * (D + lambda));
            }
            /* project (Cameras.cuh:361-369 global shutter: start pose only; :431-455 pinhole) */
            real ip[7][2];
            real im[2] = {0, 0};
            int valid = require_all_valid ? 1 : 0;
            int bail = 0;
            const real MX = (real)width * margin_factor, MY = (real)height * margin_factor;
            for (int i = 0; i < 7 && !bail; ++i) {
                real pc[3];
                ORC(quat_rotate)(pose.q, pts[i], pc);
                for (int k = 0; k < 3; ++k)
                    pc[k] += pose.t[k];
                real u = 0, v = 0;
                int pv = 0;
                if (pc[2] > 0) {
                    u = (pc[0] / pc[2]) * fx + cx;
                    v = (pc[1] / pc[2]) * fy + cy;
                    pv = (-MX <= u && u < (real)width + MX) && (-MY <= v && v < (real)height + MY);
                }
                if (require_all_valid) {
                    valid &= pv;
                    if (!pv) {
                        bail = 1;
                        break;
                    }
                } else {
                    valid |= pv;
                }
                ip[i][0] = u;
                ip[i][1] = v;
                im[0] += w_mean[i] * u;
                im[1] += w_mean[i] * v;
            }
            if (bail || !valid)
                continue;
            real cov[2][2] = {{0, 0}, {0, 0}};
            for (int i = 0; i < 7; ++i) {
                real dx = ip[i][0] - im[0], dy = ip[i][1] - im[1];
                cov[0][0] += w_cov[i] * dx * dx;
                cov[0][1] += w_cov[i] * dx * dy;
                cov[1][0] += w_cov[i] * dy * dx;
                cov[1][1] += w_cov[i] * dy * dy;
            }
            /* add_blur (Utils.cuh:171-179) */
            real det_orig = cov[0][0] * cov[1][1] - cov[0][1] * cov[1][0];
            cov[0][0] += eps2d;
            cov[1][1] += eps2d;
            real det = cov[0][0] * cov[1][1] - cov[0][1] * cov[1][0];
            real comp_arg = det_orig / det;
            real compensation = R_SQRT(comp_arg > 0 ? comp_arg : 0);
            if (det <= 0)
                continue;
            real ood = (real)1 / det;
            real inv00 = cov[1][1] * ood, inv01 = -cov[0][1] * ood, inv11 = cov[0][0] * ood;

            real extend = (real)3.33;
            if (opacities) {
                real op = opacities[gid] * compensation; /* ProjectionUT3DGSFused.cu:156-157 */
                if (op < ALPHA_THRESHOLD)
                    continue;
                real e2 = R_SQRT(2 * R_LOG(op / ALPHA_THRESHOLD));
                if (e2 < extend)
                    extend = e2;
            }
            real b = (real)0.5 * (cov[0][0] + cov[1][1]);
            real t0 = b * b - det;
            real tmp = R_SQRT(t0 > (real)0.01 ? t0 : (real)0.01);
            real v1 = b + tmp;
            real r1 = extend * R_SQRT(v1);
            real ex = extend * R_SQRT(cov[0][0]), ey = extend * R_SQRT(cov[1][1]);
            real radius_x = (real)ceil((double)(ex < r1 ? ex : r1));
            real radius_y = (real)ceil((double)(ey < r1 ? ey : r1));
            if (radius_x <= radius_clip && radius_y <= radius_clip)
                continue;
            if (im[0] + radius_x <= 0 || im[0] - radius_x >= (real)width || im[1] + radius_y <= 0 ||
                im[1] - radius_y >= (real)height)
                continue;
            radii[idx * 2] = (int32_t)radius_x;
            radii[idx * 2 + 1] = (int32_t)radius_y;
            means2d[idx * 2] = im[0];
            means2d[idx * 2 + 1] = im[1];
            depths[idx] = mean_c[2];
            conics[idx * 3] = inv00;
            conics[idx * 3 + 1] = inv01;
            conics[idx * 3 + 2] = inv11;
            if (compensations)
                compensations[idx] = compensation;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * (a4) Spherical harmonics, degree 0..4 (Sloan 2013 recurrences).
 * Follows gsplat/SphericalHarmonicsCUDA.cu:21-110 (fwd) and :113-371 (vjp); cross-checked against
 * tests/torch_impl.cpp:221-321 via the golden vectors.
 * ---------------------------------------------------------------------------------------------- */
static void ORC(sh_bases)(int degree, real x, real y, real z, real* B /* 25 */) {
    B[0] = (real)0.2820947917738781;
    if (degree < 1)
        return;
    B[1] = (real)-0.48860251190292 * y;
    B[2] = (real)0.48860251190292 * z;
    B[3] = (real)-0.48860251190292 * x;
    if (degree < 2)
        return;
    real z2 = z * z;
    real fTmp0B = (real)-1.092548430592079 * z;
    real fC1 = x * x - y * y;
    real fS1 = 2 * x * y;
    B[6] = ((real)0.9461746957575601 * z2 - (real)0.3153915652525201);
    B[7] = fTmp0B * x;
    B[5] = fTmp0B * y;
    B[8] = (real)0.5462742152960395 * fC1;
    B[4] = (real)0.5462742152960395 * fS1;
    if (degree < 3)
        return;
    real fTmp0C = (real)-2.285228997322329 * z2 + (real)0.4570457994644658;
    real fTmp1B = (real)1.445305721320277 * z;
    real fC2 = x * fC1 - y * fS1;
    real fS2 = x * fS1 + y * fC1;
    B[12] = z * ((real)1.865881662950577 * z2 - (real)1.119528997770346);
    B[13] = fTmp0C * x;
    B[11] = fTmp0C * y;
    B[14] = fTmp1B * fC1;
    B[10] = fTmp1B * fS1;
    B[15] = (real)-0.5900435899266435 * fC2;
    B[9] = (real)-0.5900435899266435 * fS2;
    if (degree < 4)
        return;
    real fTmp0D = z * ((real)-4.683325804901025 * z2 + (real)2.007139630671868);
    real fTmp1C = (real)3.31161143515146 * z2 - (real)0.47308734787878;
    real fTmp2B = (real)-1.770130769779931 * z;
    real fC3 = x * fC2 - y * fS2;
    real fS3 = x * fS2 + y * fC2;
    B[20] = ((real)1.984313483298443 * z * B[12] - (real)1.006230589874905 * B[6]);
    B[21] = fTmp0D * x;
    B[19] = fTmp0D * y;
    B[22] = fTmp1C * fC1;
    B[18] = fTmp1C * fS1;
    B[23] = fTmp2B * fC2;
    B[17] = fTmp2B * fS2;
    B[24] = (real)0.6258357354491763 * fC3;
    B[16] = (real)0.6258357354491763 * fS3;
}

/* d(bases)/d(x,y,z) at a unit direction; dB[k][3].  Restates the *_x/_y/_z terms of
 * gsplat/SphericalHarmonicsCUDA.cu:139-361. */
static void ORC(sh_bases_grad)(int degree, real x, real y, real z, real dB[25][3]) {
    for (int k = 0; k < 25; ++k)
        dB[k][0] = dB[k][1] = dB[k][2] = 0;
    if (degree < 1)
        return;
    dB[1][1] = (real)-0.48860251190292;
    dB[2][2] = (real)0.48860251190292;
    dB[3][0] = (real)-0.48860251190292;
    if (degree < 2)
        return;
    real z2 = z * z;
    real fTmp0B = (real)-1.092548430592079 * z;
    real fC1 = x * x - y * y, fS1 = 2 * x * y;
    real fTmp0B_z = (real)-1.092548430592079;
    real fC1_x = 2 * x, fC1_y = -2 * y, fS1_x = 2 * y, fS1_y = 2 * x;
    real pSH6_z = 2 * (real)0.9461746957575601 * z;
    dB[6][2] = pSH6_z;
    dB[7][0] = fTmp0B;
    dB[7][2] = fTmp0B_z * x;
    dB[5][1] = fTmp0B;
    dB[5][2] = fTmp0B_z * y;
    dB[8][0] = (real)0.5462742152960395 * fC1_x;
    dB[8][1] = (real)0.5462742152960395 * fC1_y;
    dB[4][0] = (real)0.5462742152960395 * fS1_x;
    dB[4][1] = (real)0.5462742152960395 * fS1_y;
    if (degree < 3)
        return;
    real fTmp0C = (real)-2.285228997322329 * z2 + (real)0.4570457994644658;
    real fTmp1B = (real)1.445305721320277 * z;
    real fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    real pSH12 = z * ((real)1.865881662950577 * z2 - (real)1.119528997770346);
    real fTmp0C_z = (real)-2.285228997322329 * 2 * z;
    real fTmp1B_z = (real)1.445305721320277;
    real fC2_x = fC1 + x * fC1_x - y * fS1_x;
    real fC2_y = x * fC1_y - fS1 - y * fS1_y;
    real fS2_x = fS1 + x * fS1_x + y * fC1_x;
    real fS2_y = x * fS1_y + fC1 + y * fC1_y;
    real pSH12_z = 3 * (real)1.865881662950577 * z2 - (real)1.119528997770346;
    dB[12][2] = pSH12_z;
    dB[13][0] = fTmp0C;
    dB[13][2] = fTmp0C_z * x;
    dB[11][1] = fTmp0C;
    dB[11][2] = fTmp0C_z * y;
    dB[14][0] = fTmp1B * fC1_x;
    dB[14][1] = fTmp1B * fC1_y;
    dB[14][2] = fTmp1B_z * fC1;
    dB[10][0] = fTmp1B * fS1_x;
    dB[10][1] = fTmp1B * fS1_y;
    dB[10][2] = fTmp1B_z * fS1;
    dB[15][0] = (real)-0.5900435899266435 * fC2_x;
    dB[15][1] = (real)-0.5900435899266435 * fC2_y;
    dB[9][0] = (real)-0.5900435899266435 * fS2_x;
    dB[9][1] = (real)-0.5900435899266435 * fS2_y;
    if (degree < 4)
        return;
    real fTmp0D = z * ((real)-4.683325804901025 * z2 + (real)2.007139630671868);
    real fTmp1C = (real)3.31161143515146 * z2 - (real)0.47308734787878;
    real fTmp2B = (real)-1.770130769779931 * z;
    real fTmp0D_z = 3 * (real)-4.683325804901025 * z2 + (real)2.007139630671868;
    real fTmp1C_z = 2 * (real)3.31161143515146 * z;
    real fTmp2B_z = (real)-1.770130769779931;
    real fC3_x = fC2 + x * fC2_x - y * fS2_x;
    real fC3_y = x * fC2_y - fS2 - y * fS2_y;
    real fS3_x = fS2 + y * fC2_x + x * fS2_x;
    real fS3_y = x * fS2_y + fC2 + y * fC2_y;
    dB[20][2] = (real)1.984313483298443 * (pSH12 + z * pSH12_z) + (real)-1.006230589874905 * pSH6_z;
    dB[21][0] = fTmp0D;
    dB[21][2] = fTmp0D_z * x;
    dB[19][1] = fTmp0D;
    dB[19][2] = fTmp0D_z * y;
    dB[22][0] = fTmp1C * fC1_x;
    dB[22][1] = fTmp1C * fC1_y;
    dB[22][2] = fTmp1C_z * fC1;
    dB[18][0] = fTmp1C * fS1_x;
    dB[18][1] = fTmp1C * fS1_y;
    dB[18][2] = fTmp1C_z * fS1;
    dB[23][0] = fTmp2B * fC2_x;
    dB[23][1] = fTmp2B * fC2_y;
    dB[23][2] = fTmp2B_z * fC2;
    dB[17][0] = fTmp2B * fS2_x;
    dB[17][1] = fTmp2B * fS2_y;
    dB[17][2] = fTmp2B_z * fS2;
    dB[24][0] = (real)0.6258357354491763 * fC3_x;
    dB[24][1] = (real)0.6258357354491763 * fC3_y;
    dB[16][0] = (real)0.6258357354491763 * fS3_x;
    dB[16][1] = (real)0.6258357354491763 * fS3_y;
}

/* colors[n,3]; masked-out rows are written as 0 (the reference leaves them untouched). */
void ORC(sh_fwd)(int degree, int n, int K, const real* dirs, const real* coeffs, const uint8_t* masks, real* colors) {
    const int nb = (degree + 1) * (degree + 1);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < n; ++e) {
        real* out = colors + 3 * (size_t)e;
        out[0] = out[1] = out[2] = 0;
        if (masks && !masks[e])
            continue;
        const real* d = dirs + 3 * (size_t)e;
        real B[25];
        real x = 0, y = 0, z = 0;
        if (degree >= 1) {
            real inorm = (real)1 / R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            x = d[0] * inorm;
            y = d[1] * inorm;
            z = d[2] * inorm;
        }
        ORC(sh_bases)(degree, x, y, z, B);
        const real* c = coeffs + (size_t)e * K * 3;
        for (int ch = 0; ch < 3; ++ch) {
            real r = 0;
            for (int k = 0; k < nb; ++k)
                r += B[k] * c[k * 3 + ch];
            out[ch] = r;
        }
    }
}

/* v_coeffs[n,K,3] (zero outside the active bases / masked rows), v_dirs[n,3] optional. */
void ORC(sh_bwd)(int degree, int n, int K, const real* dirs, const real* coeffs, const uint8_t* masks,
                 const real* v_colors, real* v_coeffs, real* v_dirs /* nullable */) {
    const int nb = (degree + 1) * (degree + 1);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < n; ++e) {
        real* vc = v_coeffs + (size_t)e * K * 3;
        for (int k = 0; k < K * 3; ++k)
            vc[k] = 0;
        if (v_dirs)
            v_dirs[3 * (size_t)e] = v_dirs[3 * (size_t)e + 1] = v_dirs[3 * (size_t)e + 2] = 0;
        if (masks && !masks[e])
            continue;
        const real* d = dirs + 3 * (size_t)e;
        real x = 0, y = 0, z = 0, inorm = 0;
        if (degree >= 1) {
            inorm = (real)1 / R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            x = d[0] * inorm;
            y = d[1] * inorm;
            z = d[2] * inorm;
        }
        real B[25];
        ORC(sh_bases)(degree, x, y, z, B);
        const real* vcol = v_colors + 3 * (size_t)e;
        for (int k = 0; k < nb; ++k)
            for (int ch = 0; ch < 3; ++ch)
                vc[k * 3 + ch] = B[k] * vcol[ch];
        if (v_dirs && degree >= 1) {
            real dB[25][3];
            ORC(sh_bases_grad)(degree, x, y, z, dB);
            const real* c = coeffs + (size_t)e * K * 3;
            real vdn[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ++ch)
                for (int k = 1; k < nb; ++k)
                    for (int a = 0; a < 3; ++a)
                        vdn[a] += vcol[ch] * dB[k][a] * c[k * 3 + ch];
            real dn[3] = {x, y, z};
            real dt = vdn[0] * dn[0] + vdn[1] * dn[1] + vdn[2] * dn[2];
            for (int a = 0; a < 3; ++a)
                v_dirs[3 * (size_t)e + a] = (vdn[a] - dt * dn[a]) * inorm;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * pixel -> world ray, PINHOLE / GLOBAL shutter
 * (gsplat/Cameras.cuh:457-470 image_point_to_camera_ray, :253-266 camera_ray_to_world_ray)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    real Rinv[3][3];
    real org[3];
    real fx, fy, cx, cy;
} ORC(raycam_t);
static void ORC(raycam_init)(const real* vm, const real* K, ORC(raycam_t) * rc) {
    ORC(pose_t) p;
    ORC(pose_from_viewmat)(vm, &p);
    /* glm::inverse(q) = conjugate / dot */
    real d = (p.q[1] * p.q[1] + p.q[2] * p.q[2]) + (p.q[3] * p.q[3] + p.q[0] * p.q[0]);
    real qi[4] = {p.q[0] / d, -p.q[1] / d, -p.q[2] / d, -p.q[3] / d};
    ORC(mat3_cast)(qi, rc->Rinv);
    for (int r = 0; r < 3; ++r)
        rc->org[r] = -(rc->Rinv[r][0] * p.t[0] + rc->Rinv[r][1] * p.t[1] + rc->Rinv[r][2] * p.t[2]);
    rc->fx = K[0];
    rc->fy = K[4];
    rc->cx = K[2];
    rc->cy = K[5];
}
static inline void ORC(pixel_ray)(const ORC(raycam_t) * rc, real px, real py, real d[3]) {
    real u = (px - rc->cx) / rc->fx, v = (py - rc->cy) / rc->fy;
    real len = R_SQRT(u * u + v * v + 1);
    real c[3] = {u / len, v / len, (real)1 / len};
    for (int r = 0; r < 3; ++r)
        d[r] = rc->Rinv[r][0] * c[0] + rc->Rinv[r][1] * c[1] + rc->Rinv[r][2] * c[2];
}

/* per-(pixel, gaussian) response shared by fwd and bwd
 * (gsplat/RasterizeToPixelsFromWorld3DGSFwd.cu:207-239, ...Bwd.cu:252-273) */
typedef struct {
    real R[3][3];
    real Mt[3][3]; /* S^-1 R^T : Mt[a][i] = R[i][a] / s_a */
    real omu[3], gro[3], grd[3], grd_n[3], gcrod[3];
    real power, vis;
} ORC(resp_t);
static inline void ORC(response)(const real* mean, const real* quat, const real* scale, const real o[3],
                                 const real d[3], ORC(resp_t) * r) {
    ORC(quat_to_rotmat)(quat, r->R);
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < 3; ++i)
            r->Mt[a][i] = r->R[i][a] * ((real)1 / scale[a]);
    for (int k = 0; k < 3; ++k)
        r->omu[k] = o[k] - mean[k];
    for (int a = 0; a < 3; ++a) {
        r->gro[a] = r->Mt[a][0] * r->omu[0] + r->Mt[a][1] * r->omu[1] + r->Mt[a][2] * r->omu[2];
        r->grd[a] = r->Mt[a][0] * d[0] + r->Mt[a][1] * d[1] + r->Mt[a][2] * d[2];
    }
    real l = ORC(dot3)(r->grd, r->grd);
    if (l > 0) { /* safe_normalize, Utils.cuh:181-184 */
        real il = (real)1 / R_SQRT(l);
        for (int k = 0; k < 3; ++k)
            r->grd_n[k] = r->grd[k] * il;
    } else {
        for (int k = 0; k < 3; ++k)
            r->grd_n[k] = r->grd[k];
    }
    ORC(cross3)(r->grd_n, r->gro, r->gcrod);
    r->power = (real)-0.5 * ORC(dot3)(r->gcrod, r->gcrod);
    r->vis = R_EXP(r->power);
}

/* ------------------------------------------------------------------------------------------------
 * (a7) from-world alpha blend, forward.  gsplat/RasterizeToPixelsFromWorld3DGSFwd.cu:58-279.
 * colors [C,N,CH], opacities [C,N], backgrounds [C,CH] nullable, tile_masks [C,th,tw] nullable.
 * flatten_ids index [C*N]; gaussian id = g % N (the reference is C=1 only).
 * Outputs renders [C,H,W,CH], alphas [C,H,W], last_ids [C,H,W].
 * ---------------------------------------------------------------------------------------------- */
void ORC(raster_world_fwd)(int C, int N, int CH, const real* means, const real* quats, const real* scales,
                           const real* colors, const real* opacities, const real* backgrounds,
                           const uint8_t* tile_masks, int width, int height, int tile_size, const real* viewmats,
                           const real* Ks, const int32_t* tile_offsets, const int32_t* flatten_ids,
                           int64_t n_isects, real* renders, real* alphas, int32_t* last_ids) {
    const int tw = (width + tile_size - 1) / tile_size, th = (height + tile_size - 1) / tile_size;
    for (int cid = 0; cid < C; ++cid) {
        ORC(raycam_t) rc;
        ORC(raycam_init)(viewmats + 16 * cid, Ks + 9 * cid, &rc);
        const real* bg = backgrounds ? backgrounds + (size_t)cid * CH : NULL;
#pragma omp parallel for schedule(dynamic, 4)
        for (int i = 0; i < height; ++i) {
            for (int j = 0; j < width; ++j) {
                const int tile_id = (i / tile_size) * tw + (j / tile_size);
                const size_t pix = ((size_t)cid * height + i) * width + j;
                real* out = renders + pix * CH;
                if (tile_masks && !tile_masks[(size_t)cid * tw * th + tile_id]) {
                    for (int k = 0; k < CH; ++k)
                        out[k] = bg ? bg[k] : 0;
                    alphas[pix] = 0;
                    last_ids[pix] = 0;
                    continue;
                }
                const size_t tflat = (size_t)cid * tw * th + tile_id;
                int64_t start = tile_offsets[tflat];
                int64_t end = (cid == C - 1 && tile_id == tw * th - 1) ? n_isects : tile_offsets[tflat + 1];
                real d[3];
                ORC(pixel_ray)(&rc, (real)j + (real)0.5, (real)i + (real)0.5, d);
                real T = 1;
                int32_t cur_idx = 0;
                real pix_out[64];
                for (int k = 0; k < CH; ++k)
                    pix_out[k] = 0;
                for (int64_t idx = start; idx < end; ++idx) {
                    const int32_t g = flatten_ids[idx];
                    const int32_t gid = g % N;
                    ORC(resp_t) r;
                    ORC(response)(means + 3 * gid, quats + 4 * gid, scales + 3 * gid, rc.org, d, &r);
                    real alpha = opacities[g] * r.vis;
                    if (alpha > (real)0.999)
                        alpha = (real)0.999;
                    if (alpha < (real)1 / (real)255)
                        continue;
                    const real next_T = T * (1 - alpha);
                    if (next_T <= (real)1e-4)
                        break;
                    const real vis = alpha * T;
                    for (int k = 0; k < CH; ++k)
                        pix_out[k] += colors[(size_t)g * CH + k] * vis;
                    cur_idx = (int32_t)idx;
                    T = next_T;
                }
                alphas[pix] = 1 - T;
                for (int k = 0; k < CH; ++k)
                    out[k] = bg ? pix_out[k] + T * bg[k] : pix_out[k];
                last_ids[pix] = cur_idx;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * (a8) from-world alpha blend, backward.  gsplat/RasterizeToPixelsFromWorld3DGSBwd.cu:63-372 and
 * gsplat/Utils.cuh:104-158,186-194.  Accumulates in double regardless of REAL (the reference's float
 * atomics are order dependent).  Outputs (double): v_means[N,3] v_quats[N,4] v_scales[N,3]
 * v_colors[C,N,3] v_opacities[C,N]; must be zeroed by the caller.
 * ---------------------------------------------------------------------------------------------- */
void ORC(raster_world_bwd)(int C, int N, const real* means, const real* quats, const real* scales,
                           const real* colors, const real* opacities, const real* backgrounds,
                           const uint8_t* tile_masks, int width, int height, int tile_size, const real* viewmats,
                           const real* Ks, const int32_t* tile_offsets, const int32_t* flatten_ids,
                           int64_t n_isects, const real* render_alphas, const int32_t* last_ids,
                           const real* v_render_colors, const real* v_render_alphas, double* v_means,
                           double* v_quats, double* v_scales, double* v_colors, double* v_opacities) {
    const int tw = (width + tile_size - 1) / tile_size, th = (height + tile_size - 1) / tile_size;
    (void)n_isects;
    for (int cid = 0; cid < C; ++cid) {
        ORC(raycam_t) rc;
        ORC(raycam_init)(viewmats + 16 * cid, Ks + 9 * cid, &rc);
        const real* bg = backgrounds ? backgrounds + (size_t)cid * 3 : NULL;
#pragma omp parallel for schedule(dynamic, 4)
        for (int i = 0; i < height; ++i) {
            for (int j = 0; j < width; ++j) {
                const int tile_id = (i / tile_size) * tw + (j / tile_size);
                const size_t tflat = (size_t)cid * tw * th + tile_id;
                if (tile_masks && !tile_masks[tflat])
                    continue;
                const size_t pix = ((size_t)cid * height + i) * width + j;
                const int64_t start = tile_offsets[tflat];
                const int64_t bin_final = last_ids[pix];
                real d[3];
                ORC(pixel_ray)(&rc, (real)j + (real)0.5, (real)i + (real)0.5, d);
                const real T_final = 1 - render_alphas[pix];
                real T = T_final;
                real buffer[3] = {0, 0, 0};
                const real* vrc = v_render_colors + pix * 3;
                const real vra = v_render_alphas[pix];
                for (int64_t idx = bin_final; idx >= start; --idx) {
                    const int32_t g = flatten_ids[idx];
                    const int32_t gid = g % N;
                    const real* quat = quats + 4 * gid;
                    const real* scale = scales + 3 * gid;
                    ORC(resp_t) r;
                    ORC(response)(means + 3 * gid, quat, scale, rc.org, d, &r);
                    const real opac = opacities[g];
                    real alpha = opac * r.vis;
                    if (alpha > (real)0.999)
                        alpha = (real)0.999;
                    if (r.power > 0 || alpha < (real)1 / (real)255)
                        continue;
                    const real ra = (real)1 / (1 - alpha);
                    T *= ra;
                    const real fac = alpha * T;
                    const real* rgb = colors + (size_t)g * 3;
                    real v_alpha = 0;
                    for (int k = 0; k < 3; ++k)
                        v_alpha += (rgb[k] * T - buffer[k] * ra) * vrc[k];
                    v_alpha += T_final * ra * vra;
                    if (bg) {
                        real accum = 0;
                        for (int k = 0; k < 3; ++k)
                            accum += bg[k] * vrc[k];
                        v_alpha += -T_final * ra * accum;
                    }
                    double lm[3] = {0, 0, 0}, ls[3] = {0, 0, 0}, lq[4] = {0, 0, 0, 0}, lo = 0;
                    if (opac * r.vis <= (real)0.999) {
                        const real v_vis = opac * v_alpha;
                        const real v_gradDist = (real)-0.5 * r.vis * v_vis;
                        real v_gcrod[3], v_grd_n[3], v_gro[3], v_grd[3];
                        for (int k = 0; k < 3; ++k)
                            v_gcrod[k] = 2 * v_gradDist * r.gcrod[k];
                        ORC(cross3)(v_gcrod, r.gro, v_grd_n);
                        for (int k = 0; k < 3; ++k)
                            v_grd_n[k] = -v_grd_n[k];
                        ORC(cross3)(v_gcrod, r.grd_n, v_gro);
                        { /* safe_normalize_bw */
                            real l = ORC(dot3)(r.grd, r.grd);
                            if (l > 0) {
                                real il = (real)1 / R_SQRT(l);
                                real il3 = il * il * il;
                                real dt = ORC(dot3)(v_grd_n, r.grd);
                                for (int k = 0; k < 3; ++k)
                                    v_grd[k] = il * v_grd_n[k] - il3 * dt * r.grd[k];
                            } else {
                                for (int k = 0; k < 3; ++k)
                                    v_grd[k] = v_grd_n[k];
                            }
                        }
                        /* G_Mt[a][i] = v_grd[a] d[i] + v_gro[a] omu[i];  v_M = G_Mt^T (M = R S^-1) */
                        real GM[3][3], GR[3][3];
                        for (int a = 0; a < 3; ++a)
                            for (int ii = 0; ii < 3; ++ii)
                                GM[ii][a] = v_grd[a] * d[ii] + v_gro[a] * r.omu[ii];
                        for (int ii = 0; ii < 3; ++ii) {
                            real vo = 0;
                            for (int a = 0; a < 3; ++a)
                                vo += r.Mt[a][ii] * v_gro[a];
                            lm[ii] = -(double)vo;
                        }
                        real vq[4] = {0, 0, 0, 0};
                        for (int a = 0; a < 3; ++a) {
                            real isa = (real)1 / scale[a];
                            for (int ii = 0; ii < 3; ++ii)
                                GR[ii][a] = GM[ii][a] * isa;
                            ls[a] = (double)(-isa * isa *
                                             (r.R[0][a] * GM[0][a] + r.R[1][a] * GM[1][a] + r.R[2][a] * GM[2][a]));
                        }
                        ORC(quat_to_rotmat_vjp)(quat, GR, vq);
                        for (int k = 0; k < 4; ++k)
                            lq[k] = (double)vq[k];
                        lo = (double)(r.vis * v_alpha);
                    }
                    for (int k = 0; k < 3; ++k) {
#pragma omp atomic
                        v_colors[(size_t)g * 3 + k] += (double)(fac * vrc[k]);
#pragma omp atomic
                        v_means[(size_t)gid * 3 + k] += lm[k];
#pragma omp atomic
                        v_scales[(size_t)gid * 3 + k] += ls[k];
                    }
                    for (int k = 0; k < 4; ++k) {
#pragma omp atomic
                        v_quats[(size_t)gid * 4 + k] += lq[k];
                    }
#pragma omp atomic
                    v_opacities[g] += lo;
                    for (int k = 0; k < 3; ++k)
                        buffer[k] += rgb[k] * fac;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * (a11/a12) Adam.  fastgs/optimizer/include/adam_kernels.cuh:13-36; bias corrections computed in
 * double on the host as src/training/optimizers/fused_adam.cpp:66-93 does.
 * ---------------------------------------------------------------------------------------------- */
void ORC(adam_step)(real* param, real* exp_avg, real* exp_avg_sq, const real* grad, int64_t n, real lr, real beta1,
                    real beta2, real eps, real bc1_rcp, real bc2_sqrt_rcp) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const real g = grad[i];
        const real m1 = beta1 * exp_avg[i] + ((real)1 - beta1) * g;
        const real m2 = beta2 * exp_avg_sq[i] + ((real)1 - beta2) * g * g;
        const real denom = R_SQRT(m2) * bc2_sqrt_rcp + eps;
        const real step_size = lr * bc1_rcp;
        param[i] -= step_size * m1 / denom;
        exp_avg[i] = m1;
        exp_avg_sq[i] = m2;
    }
}

#undef R_SQRT
#undef R_EXP
#undef R_LOG
#undef ORC
#undef real
#undef CAT
#undef CAT_
