/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 *
 * Oracle of the legacy 2-D operator surface the reference's gtest files call:
 *   gsplat::quat_scale_to_covar_preci_fwd/bwd      tests/test_basic.cpp:54-81, tests/test_gsplat_ops.cpp:76-96
 *   gsplat::projection_ewa_3dgs_fused_fwd          tests/test_basic.cpp:114-128, tests/test_gsplat_ops.cpp:174-189
 *   gsplat::rasterize_to_pixels_3dgs_fwd/bwd       tests/test_basic.cpp:347-358, launchers gsplat/Rasterization.h:16-63
 * The CUDA sources of these ops are NOT in the reference tree (gsplat/ only holds the from-world kernels), so:
 *   - quat_scale_to_covar_preci and the projection restate the reference's own CPU statement of them,
 *     tests/torch_impl.cpp:38-77 and :80-218 (pinned by tests/golden/torch_impl_golden.npz, generated from the unmodified
 *     file), with the culling tail the reference's surviving projection kernel uses (ProjectionUT3DGSFused.cu:142-199:
 *     opacity-aware extent, radius_clip, frustum test, culled entries left at zero);
 *   - the 2-D blend follows the blend loop of the reference's from-world kernels line by line
 *     (RasterizeToPixelsFromWorld3DGSFwd.cu:193-279, ...Bwd.cu:196-370: alpha clamp 0.999, skip below 1/255, stop at
 *     T <= 1e-4, back-to-front backward with T /= (1 - alpha)) with the 2-D conic response
 *     sigma = 1/2 (a dx^2 + c dy^2) + b dx dy of the upstream gsplat kernel those launchers belong to
 *     (nerfstudio-project/gsplat v1.x, rasterize_to_pixels_3dgs_{fwd,bwd}.cu -- third-party, absent here).
 *     PARITY UNPINNED for the 2-D blend: there is no reference implementation, golden vector or fixture of it in the tree to
 *     pin against.  What holds it: finite differences of its own forward (tests/test_oracle_legacy2d.py) and the identity
 *     "EWA-projected small Gaussians render like the from-world kernel" (whose oracle IS pinned on the reference CUDA build).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define ORC(name) CAT(PFX, name)
#define real REAL
#define R_SQRT(x) ((real)sqrt((double)(x)))
#define R_EXP(x) ((real)exp((double)(x)))
#define R_LOG(x) ((real)log((double)(x)))

/* ---- quat / scale -> covariance, precision (torch_impl.cpp:38-77).  triu: (xx, xy, xz, yy, yz, zz). ---- */
void ORC(quat_scale_to_covar_preci_fwd)(int N, const real* quats, const real* scales, int triu, real* covars /* nullable */,
                                        real* precis /* nullable */) {
    for (int i = 0; i < N; ++i) {
        real R[3][3];
        ORC(quat_to_rotmat)(quats + 4 * i, R);
        for (int which = 0; which < 2; ++which) {
            real* out = which ? precis : covars;
            if (!out)
                continue;
            real M[3][3], S[3][3];
            for (int r = 0; r < 3; ++r)
                for (int a = 0; a < 3; ++a)
                    M[r][a] = R[r][a] * (which ? (real)1 / scales[3 * i + a] : scales[3 * i + a]);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    S[r][c] = M[r][0] * M[c][0] + M[r][1] * M[c][1] + M[r][2] * M[c][2];
            if (triu) {
                real* o = out + 6 * (size_t)i;
                o[0] = S[0][0];
                o[1] = (S[0][1] + S[1][0]) / 2;
                o[2] = (S[0][2] + S[2][0]) / 2;
                o[3] = S[1][1];
                o[4] = (S[1][2] + S[2][1]) / 2;
                o[5] = S[2][2];
            } else {
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        out[9 * (size_t)i + 3 * r + c] = S[r][c];
            }
        }
    }
}

/* VJP of the above (covar = M M^T, M = R S; preci = P P^T, P = R S^-1; v_M = (V + V^T) M; quaternion through
 * gsplat/Utils.cuh:104-126 including the normalisation).  v_quats [N,4], v_scales [N,3] are overwritten. */
void ORC(quat_scale_to_covar_preci_bwd)(int N, const real* quats, const real* scales, int triu,
                                        const real* v_covars /* nullable */, const real* v_precis /* nullable */,
                                        real* v_quats, real* v_scales) {
    for (int i = 0; i < N; ++i) {
        real R[3][3], GR[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, vs[3] = {0, 0, 0};
        ORC(quat_to_rotmat)(quats + 4 * i, R);
        for (int which = 0; which < 2; ++which) {
            const real* vin = which ? v_precis : v_covars;
            if (!vin)
                continue;
            real V[3][3];
            if (triu) {
                const real* v = vin + 6 * (size_t)i;
                V[0][0] = v[0], V[1][1] = v[3], V[2][2] = v[5];
                V[0][1] = V[1][0] = v[1] / 2;
                V[0][2] = V[2][0] = v[2] / 2;
                V[1][2] = V[2][1] = v[4] / 2;
            } else {
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        V[r][c] = vin[9 * (size_t)i + 3 * r + c];
            }
            for (int a = 0; a < 3; ++a) {
                const real s = scales[3 * i + a];
                const real f = which ? (real)1 / s : s; /* M[r][a] = R[r][a] f */
                real acc = 0;
                for (int r = 0; r < 3; ++r) {
                    real vM = 0; /* ((V + V^T) M)[r][a] */
                    for (int c = 0; c < 3; ++c)
                        vM += (V[r][c] + V[c][r]) * R[c][a] * f;
                    GR[r][a] += vM * f;
                    acc += R[r][a] * vM;
                }
                vs[a] += which ? -acc / (s * s) : acc;
            }
        }
        real vq[4] = {0, 0, 0, 0};
        ORC(quat_to_rotmat_vjp)(quats + 4 * i, GR, vq);
        for (int k = 0; k < 4; ++k)
            v_quats[4 * (size_t)i + k] = vq[k];
        for (int k = 0; k < 3; ++k)
            v_scales[3 * (size_t)i + k] = vs[k];
    }
}

/* ---- fused EWA projection, pinhole (torch_impl.cpp:80-218; tail ProjectionUT3DGSFused.cu:142-199) ---- */
void ORC(projection_ewa)(int C, int N, const real* means, const real* covars /* [N,3,3] nullable */, const real* quats,
                         const real* scales, const real* opacities /* [N] nullable */, const real* viewmats,
                         const real* Ks, int width, int height, real eps2d, real near_plane, real far_plane,
                         real radius_clip, int32_t* radii, real* means2d, real* depths, real* conics,
                         real* compensations /* nullable */) {
    const real ALPHA_THRESHOLD = (real)1 / (real)255;
    for (int cid = 0; cid < C; ++cid) {
        const real* vm = viewmats + 16 * cid;
        const real fx = Ks[cid * 9 + 0], fy = Ks[cid * 9 + 4], cx = Ks[cid * 9 + 2], cy = Ks[cid * 9 + 5];
        for (int gid = 0; gid < N; ++gid) {
            const size_t idx = (size_t)cid * N + gid;
            radii[idx * 2] = radii[idx * 2 + 1] = 0;
            means2d[idx * 2] = means2d[idx * 2 + 1] = 0;
            depths[idx] = 0;
            conics[idx * 3] = conics[idx * 3 + 1] = conics[idx * 3 + 2] = 0;
            if (compensations)
                compensations[idx] = 0;
            real mc[3];
            for (int r = 0; r < 3; ++r)
                mc[r] = vm[4 * r] * means[3 * gid] + vm[4 * r + 1] * means[3 * gid + 1] + vm[4 * r + 2] * means[3 * gid + 2] +
                        vm[4 * r + 3];
            if (mc[2] < near_plane || mc[2] > far_plane)
                continue;
            real S[3][3];
            if (covars) {
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        S[r][c] = covars[9 * (size_t)gid + 3 * r + c];
            } else {
                real R[3][3], M[3][3];
                ORC(quat_to_rotmat)(quats + 4 * gid, R);
                for (int r = 0; r < 3; ++r)
                    for (int a = 0; a < 3; ++a)
                        M[r][a] = R[r][a] * scales[3 * gid + a];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        S[r][c] = M[r][0] * M[c][0] + M[r][1] * M[c][1] + M[r][2] * M[c][2];
            }
            /* covar_c = W S W^T (torch_impl.cpp:128-144) */
            real WS[3][3], Sc[3][3];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    WS[r][c] = vm[4 * r] * S[0][c] + vm[4 * r + 1] * S[1][c] + vm[4 * r + 2] * S[2][c];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    Sc[r][c] = WS[r][0] * vm[4 * c] + WS[r][1] * vm[4 * c + 1] + WS[r][2] * vm[4 * c + 2];
            /* persp_proj (torch_impl.cpp:80-126) */
            const real tz = mc[2], tz2 = tz * tz;
            const real tan_fovx = (real)0.5 * width / fx, tan_fovy = (real)0.5 * height / fy;
            const real lim_x_pos = (width - cx) / fx + (real)0.3 * tan_fovx, lim_x_neg = cx / fx + (real)0.3 * tan_fovx;
            const real lim_y_pos = (height - cy) / fy + (real)0.3 * tan_fovy, lim_y_neg = cy / fy + (real)0.3 * tan_fovy;
            real rx = mc[0] / tz, ry = mc[1] / tz;
            rx = rx < -lim_x_neg ? -lim_x_neg : (rx > lim_x_pos ? lim_x_pos : rx);
            ry = ry < -lim_y_neg ? -lim_y_neg : (ry > lim_y_pos ? lim_y_pos : ry);
            const real tx = tz * rx, ty = tz * ry;
            const real J[2][3] = {{fx / tz, 0, -fx * tx / tz2}, {0, fy / tz, -fy * ty / tz2}};
            real JS[2][3], cov[2][2];
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 3; ++c)
                    JS[r][c] = J[r][0] * Sc[0][c] + J[r][1] * Sc[1][c] + J[r][2] * Sc[2][c];
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 2; ++c)
                    cov[r][c] = JS[r][0] * J[c][0] + JS[r][1] * J[c][1] + JS[r][2] * J[c][2];
            const real im[2] = {fx * mc[0] / tz + cx, fy * mc[1] / tz + cy};
            const real det_orig = cov[0][0] * cov[1][1] - cov[0][1] * cov[1][0];
            cov[0][0] += eps2d;
            cov[1][1] += eps2d;
            const real det = cov[0][0] * cov[1][1] - cov[0][1] * cov[1][0];
            if (det <= 0)
                continue;
            const real comp_arg = det_orig / det;
            const real compensation = R_SQRT(comp_arg > 0 ? comp_arg : 0);
            const real ood = (real)1 / det;
            const real inv00 = cov[1][1] * ood, inv01 = -(cov[0][1] + cov[1][0]) / 2 * ood, inv11 = cov[0][0] * ood;
            real extend = (real)3.33;
            if (opacities) {
                real op = opacities[gid];
                if (compensations)
                    op *= compensation;
                if (op < ALPHA_THRESHOLD)
                    continue;
                const real e2 = R_SQRT(2 * R_LOG(op / ALPHA_THRESHOLD));
                if (e2 < extend)
                    extend = e2;
            }
            const real radius_x = (real)ceil((double)(extend * R_SQRT(cov[0][0])));
            const real radius_y = (real)ceil((double)(extend * R_SQRT(cov[1][1])));
            if (radius_x <= radius_clip && radius_y <= radius_clip)
                continue;
            if (im[0] + radius_x <= 0 || im[0] - radius_x >= (real)width || im[1] + radius_y <= 0 ||
                im[1] - radius_y >= (real)height)
                continue;
            radii[idx * 2] = (int32_t)radius_x;
            radii[idx * 2 + 1] = (int32_t)radius_y;
            means2d[idx * 2] = im[0];
            means2d[idx * 2 + 1] = im[1];
            depths[idx] = mc[2];
            conics[idx * 3] = inv00;
            conics[idx * 3 + 1] = inv01;
            conics[idx * 3 + 2] = inv11;
            if (compensations)
                compensations[idx] = compensation;
        }
    }
}

/* ---- 2-D alpha blend, forward.  means2d [C,N,2], conics [C,N,3], colors [C,N,CH], opacities [C,N]. ---- */
void ORC(raster_2d_fwd)(int C, int N, int CH, const real* means2d, const real* conics, const real* colors,
                        const real* opacities, const real* backgrounds, const uint8_t* tile_masks, int width, int height,
                        int tile_size, const int32_t* tile_offsets, const int32_t* flatten_ids, int64_t n_isects,
                        real* renders, real* alphas, int32_t* last_ids) {
    const int tw = (width + tile_size - 1) / tile_size, th = (height + tile_size - 1) / tile_size;
    (void)N;
    for (int cid = 0; cid < C; ++cid) {
        const real* bg = backgrounds ? backgrounds + (size_t)cid * CH : NULL;
#pragma omp parallel for schedule(dynamic, 4)
        for (int i = 0; i < height; ++i) {
            for (int j = 0; j < width; ++j) {
                const int tile_id = (i / tile_size) * tw + (j / tile_size);
                const size_t tflat = (size_t)cid * tw * th + tile_id;
                const size_t pix = ((size_t)cid * height + i) * width + j;
                real* out = renders + pix * CH;
                if (tile_masks && !tile_masks[tflat]) {
                    for (int k = 0; k < CH; ++k)
                        out[k] = bg ? bg[k] : 0;
                    alphas[pix] = 0;
                    last_ids[pix] = 0;
                    continue;
                }
                const int64_t start = tile_offsets[tflat];
                const int64_t end = (cid == C - 1 && tile_id == tw * th - 1) ? n_isects : tile_offsets[tflat + 1];
                const real px = (real)j + (real)0.5, py = (real)i + (real)0.5;
                real T = 1;
                int32_t cur_idx = 0;
                real pix_out[64];
                for (int k = 0; k < CH; ++k)
                    pix_out[k] = 0;
                for (int64_t idx = start; idx < end; ++idx) {
                    const int32_t g = flatten_ids[idx];
                    const real dx = means2d[2 * (size_t)g] - px, dy = means2d[2 * (size_t)g + 1] - py;
                    const real* cn = conics + 3 * (size_t)g;
                    const real sigma = (real)0.5 * (cn[0] * dx * dx + cn[2] * dy * dy) + cn[1] * dx * dy;
                    real alpha = opacities[g] * R_EXP(-sigma);
                    if (alpha > (real)0.999)
                        alpha = (real)0.999;
                    if (sigma < 0 || alpha < (real)1 / (real)255)
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

/* ---- 2-D alpha blend, backward (double accumulation).  Outputs zeroed by the caller:
 * v_means2d [C,N,2], v_means2d_abs [C,N,2] (nullable), v_conics [C,N,3], v_colors [C,N,CH], v_opacities [C,N]. ---- */
void ORC(raster_2d_bwd)(int C, int N, int CH, const real* means2d, const real* conics, const real* colors,
                        const real* opacities, const real* backgrounds, const uint8_t* tile_masks, int width, int height,
                        int tile_size, const int32_t* tile_offsets, const int32_t* flatten_ids, const real* render_alphas,
                        const int32_t* last_ids, const real* v_render_colors, const real* v_render_alphas,
                        double* v_means2d, double* v_means2d_abs, double* v_conics, double* v_colors,
                        double* v_opacities) {
    const int tw = (width + tile_size - 1) / tile_size, th = (height + tile_size - 1) / tile_size;
    (void)N;
    for (int cid = 0; cid < C; ++cid) {
        const real* bg = backgrounds ? backgrounds + (size_t)cid * CH : NULL;
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
                const real px = (real)j + (real)0.5, py = (real)i + (real)0.5;
                const real T_final = 1 - render_alphas[pix];
                real T = T_final;
                real buffer[64];
                for (int k = 0; k < CH; ++k)
                    buffer[k] = 0;
                const real* vrc = v_render_colors + pix * CH;
                const real vra = v_render_alphas[pix];
                for (int64_t idx = bin_final; idx >= start; --idx) {
                    const int32_t g = flatten_ids[idx];
                    const real dx = means2d[2 * (size_t)g] - px, dy = means2d[2 * (size_t)g + 1] - py;
                    const real* cn = conics + 3 * (size_t)g;
                    const real sigma = (real)0.5 * (cn[0] * dx * dx + cn[2] * dy * dy) + cn[1] * dx * dy;
                    const real vis = R_EXP(-sigma);
                    const real opac = opacities[g];
                    real alpha = opac * vis;
                    if (alpha > (real)0.999)
                        alpha = (real)0.999;
                    if (sigma < 0 || alpha < (real)1 / (real)255)
                        continue;
                    const real ra = (real)1 / (1 - alpha);
                    T *= ra;
                    const real fac = alpha * T;
                    const real* rgb = colors + (size_t)g * CH;
                    real v_alpha = 0;
                    for (int k = 0; k < CH; ++k)
                        v_alpha += (rgb[k] * T - buffer[k] * ra) * vrc[k];
                    v_alpha += T_final * ra * vra;
                    if (bg) {
                        real accum = 0;
                        for (int k = 0; k < CH; ++k)
                            accum += bg[k] * vrc[k];
                        v_alpha += -T_final * ra * accum;
                    }
                    for (int k = 0; k < CH; ++k) {
#pragma omp atomic
                        v_colors[(size_t)g * CH + k] += (double)(fac * vrc[k]);
                    }
                    if (opac * vis <= (real)0.999) {
                        const real v_sigma = -opac * vis * v_alpha;
                        const double vc[3] = {(double)((real)0.5 * v_sigma * dx * dx), (double)(v_sigma * dx * dy),
                                              (double)((real)0.5 * v_sigma * dy * dy)};
                        const double vxy[2] = {(double)(v_sigma * (cn[0] * dx + cn[1] * dy)),
                                               (double)(v_sigma * (cn[1] * dx + cn[2] * dy))};
                        for (int k = 0; k < 3; ++k) {
#pragma omp atomic
                            v_conics[3 * (size_t)g + k] += vc[k];
                        }
                        for (int k = 0; k < 2; ++k) {
#pragma omp atomic
                            v_means2d[2 * (size_t)g + k] += vxy[k];
                            if (v_means2d_abs) {
#pragma omp atomic
                                v_means2d_abs[2 * (size_t)g + k] += fabs(vxy[k]);
                            }
                        }
#pragma omp atomic
                        v_opacities[g] += (double)(vis * v_alpha);
                    }
                    for (int k = 0; k < CH; ++k)
                        buffer[k] += rgb[k] * fac;
                }
            }
        }
    }
}

#undef R_SQRT
#undef R_EXP
#undef R_LOG
#undef ORC
#undef real
#undef CAT
#undef CAT_
