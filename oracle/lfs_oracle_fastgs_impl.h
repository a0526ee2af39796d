/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 *
 * Type-generic CPU restatement of the reference's fastgs (EWA) rasterizer, forward and backward
 * (SURVEY §8 rows a9 / a10).  Included by lfs_oracle.c with REAL=float (orc32_) and REAL=double (orc64_).
 * Every block cites the reference file:line it follows (paths relative to /root/reference/fastgs/rasterization).
 *
 * Deviations from the reference that are deliberate and documented:
 *  - depth order of equal-depth primitives: the reference compacts visible primitives with atomicAdd
 *    (include/kernels_forward.cuh:200-203) so ties are ordered non-deterministically; here (and in the CUDA path)
 *    ties are ordered by primitive index;
 *  - transcendental functions are evaluated exactly in `real`; the reference is compiled with --use_fast_math.
 *
 * Parity status: pinned against the unmodified reference CUDA kernels (oracle/_ref/libfastgs_ref.so) through golden
 * vectors recorded on the B200 box: tests/golden/fastgs_ref_golden.npz, made by tests/golden/make_fastgs_golden.py.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define ORC(name) CAT(PFX, name)
#define real REAL
#define R_SQRT(x) ((real)sqrt((double)(x)))
#define R_EXP(x) ((real)exp((double)(x)))
#define R_LOG(x) ((real)log((double)(x)))

#ifndef LFS_ORACLE_FASTGS_CONSTS
#define LFS_ORACLE_FASTGS_CONSTS
#define FG_DILATION 0.3            /* include/rasterization_config.h:16 */
#define FG_MIN_ALPHA_RCP 255.0     /* :17 */
#define FG_MAX_ALPHA 0.999         /* :19 */
#define FG_T_THRESHOLD 1e-4        /* :20 */
#define FG_TILE 16                 /* :28-29 */
#endif

typedef struct {
    int visible;
    real depth;
    uint32_t depth_key;
    real mean2d[2], conic[3], opacity, color[3]; /* colour before the clamp at blend time */
    int x0, x1, y0, y1;                          /* screen bounds in tiles (x_min, x_max, y_min, y_max) */
    int n_touched;
} ORC(fg_prim);

/* include/kernel_utils.cuh:15-39 convert_sh_to_color */
static void ORC(fg_sh_to_color)(const real* sh0, const real* shN, int total_rest, const real pos[3], const real cam[3],
                                int idx, int active, real out[3]) {
    const real* c0 = sh0 + 3 * (size_t)idx;
    const real* c = shN + 3 * (size_t)idx * total_rest;
    real r[3];
    for (int k = 0; k < 3; ++k)
        r[k] = (real)0.5 + (real)0.28209479177387814 * c0[k];
    if (active > 1) {
        real d[3] = {pos[0] - cam[0], pos[1] - cam[1], pos[2] - cam[2]};
        real inv = (real)1 / R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        real x = d[0] * inv, y = d[1] * inv, z = d[2] * inv;
        real b[15];
        int nb = 3;
        b[0] = (real)-0.48860251190291987 * y;
        b[1] = (real)0.48860251190291987 * z;
        b[2] = (real)-0.48860251190291987 * x;
        if (active > 4) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
            b[3] = (real)1.0925484305920792 * xy;
            b[4] = (real)-1.0925484305920792 * yz;
            b[5] = (real)0.94617469575755997 * zz - (real)0.31539156525251999;
            b[6] = (real)-1.0925484305920792 * xz;
            b[7] = (real)0.54627421529603959 * xx - (real)0.54627421529603959 * yy;
            nb = 8;
            if (active > 9) {
                b[8] = (real)0.59004358992664352 * y * ((real)-3 * xx + yy);
                b[9] = (real)2.8906114426405538 * xy * z;
                b[10] = (real)0.45704579946446572 * y * ((real)1 - (real)5 * zz);
                b[11] = (real)0.3731763325901154 * z * ((real)5 * zz - (real)3);
                b[12] = (real)0.45704579946446572 * x * ((real)1 - (real)5 * zz);
                b[13] = (real)1.4453057213202769 * z * (xx - yy);
                b[14] = (real)0.59004358992664352 * x * (-xx + (real)3 * yy);
                nb = 15;
            }
        }
        for (int j = 0; j < nb; ++j)
            for (int k = 0; k < 3; ++k)
                r[k] += b[j] * c[3 * j + k];
    }
    out[0] = r[0], out[1] = r[1], out[2] = r[2];
}

/* include/kernel_utils.cuh:41-105 convert_sh_to_color_backward.
 * grad_color in, writes grad_sh0 [3], grad_shN [total_rest,3] (entries >= active-1 untouched), returns dcolor/dposition. */
static void ORC(fg_sh_backward)(const real* shN, int total_rest, const real pos[3], const real cam[3], int idx, int active,
                                const real gc[3], real* g_sh0, real* g_shN, real dpos[3]) {
    const real* c = shN + 3 * (size_t)idx * total_rest;
    real* g = g_shN + 3 * (size_t)idx * total_rest;
    for (int k = 0; k < 3; ++k)
        g_sh0[3 * (size_t)idx + k] = (real)0.28209479177387814 * gc[k];
    dpos[0] = dpos[1] = dpos[2] = 0;
    if (active <= 1)
        return;
    real xr = pos[0] - cam[0], yr = pos[1] - cam[1], zr = pos[2] - cam[2];
    real inv = (real)1 / R_SQRT(xr * xr + yr * yr + zr * zr);
    real x = xr * inv, y = yr * inv, z = zr * inv;
    real b[15], dx[15], dy[15], dz[15];
    int nb = 3;
    for (int j = 0; j < 15; ++j)
        b[j] = dx[j] = dy[j] = dz[j] = 0;
    b[0] = (real)-0.48860251190291987 * y, dy[0] = (real)-0.48860251190291987;
    b[1] = (real)0.48860251190291987 * z, dz[1] = (real)0.48860251190291987;
    b[2] = (real)-0.48860251190291987 * x, dx[2] = (real)-0.48860251190291987;
    if (active > 4) {
        real xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
        nb = 8;
        b[3] = (real)1.0925484305920792 * xy, dx[3] = (real)1.0925484305920792 * y, dy[3] = (real)1.0925484305920792 * x;
        b[4] = (real)-1.0925484305920792 * yz, dy[4] = (real)-1.0925484305920792 * z, dz[4] = (real)-1.0925484305920792 * y;
        b[5] = (real)0.94617469575755997 * zz - (real)0.31539156525251999, dz[5] = (real)1.8923493915151202 * z;
        b[6] = (real)-1.0925484305920792 * xz, dx[6] = (real)-1.0925484305920792 * z, dz[6] = (real)-1.0925484305920792 * x;
        b[7] = (real)0.54627421529603959 * xx - (real)0.54627421529603959 * yy, dx[7] = (real)1.0925484305920792 * x,
        dy[7] = (real)-1.0925484305920792 * y;
        if (active > 9) {
            nb = 15;
            b[8] = (real)0.59004358992664352 * y * ((real)-3 * xx + yy);
            dx[8] = (real)-3.5402615395598609 * xy, dy[8] = (real)-1.7701307697799304 * xx + (real)1.7701307697799304 * yy;
            b[9] = (real)2.8906114426405538 * xy * z;
            dx[9] = (real)2.8906114426405538 * yz, dy[9] = (real)2.8906114426405538 * xz, dz[9] = (real)2.8906114426405538 * xy;
            b[10] = (real)0.45704579946446572 * y * ((real)1 - (real)5 * zz);
            dy[10] = (real)0.45704579946446572 - (real)2.2852289973223288 * zz, dz[10] = (real)-4.5704579946446566 * yz;
            b[11] = (real)0.3731763325901154 * z * ((real)5 * zz - (real)3);
            dz[11] = (real)5.597644988851731 * zz - (real)1.1195289977703462;
            b[12] = (real)0.45704579946446572 * x * ((real)1 - (real)5 * zz);
            dx[12] = (real)0.45704579946446572 - (real)2.2852289973223288 * zz, dz[12] = (real)-4.5704579946446566 * xz;
            b[13] = (real)1.4453057213202769 * z * (xx - yy);
            dx[13] = (real)2.8906114426405538 * xz, dy[13] = (real)-2.8906114426405538 * yz,
            dz[13] = (real)1.4453057213202769 * xx - (real)1.4453057213202769 * yy;
            b[14] = (real)0.59004358992664352 * x * (-xx + (real)3 * yy);
            dx[14] = (real)-1.7701307697799304 * xx + (real)1.7701307697799304 * yy, dy[14] = (real)3.5402615395598609 * xy;
        }
    }
    real gd[3] = {0, 0, 0}; /* dot(grad_direction_{x,y,z}, grad_color), :93-96 */
    for (int j = 0; j < nb; ++j) {
        real cg = c[3 * j] * gc[0] + c[3 * j + 1] * gc[1] + c[3 * j + 2] * gc[2];
        gd[0] += dx[j] * cg, gd[1] += dy[j] * cg, gd[2] += dz[j] * cg;
        for (int k = 0; k < 3; ++k)
            g[3 * j + k] = b[j] * gc[k];
    }
    /* :97-104 normalisation Jacobian */
    real xx = xr * xr, yy = yr * yr, zz = zr * zr, xy = xr * yr, xz = xr * zr, yz = yr * zr;
    real n2 = xx + yy + zz;
    real s = (real)1 / R_SQRT(n2 * n2 * n2);
    dpos[0] = ((yy + zz) * gd[0] - xy * gd[1] - xz * gd[2]) * s;
    dpos[1] = (-xy * gd[0] + (xx + zz) * gd[1] - yz * gd[2]) * s;
    dpos[2] = (-xz * gd[0] - yz * gd[1] + (xx + yy) * gd[2]) * s;
}

/* include/kernel_utils.cuh:108-143 will_primitive_contribute (mean already shifted by -0.5, :152) */
static int ORC(fg_will_contribute)(const real mean[2], const real conic[3], int tile_x, int tile_y, real power_threshold) {
    real rminx = (real)(tile_x * FG_TILE), rminy = (real)(tile_y * FG_TILE);
    real rmaxx = (real)((tile_x + 1) * FG_TILE - 1), rmaxy = (real)((tile_y + 1) * FG_TILE - 1);
    real x_min_diff = rminx - mean[0];
    real x_left = x_min_diff > 0 ? (real)1 : (real)0;
    real not_in_x = x_left + (mean[0] > rmaxx ? (real)1 : (real)0);
    real y_min_diff = rminy - mean[1];
    real y_above = y_min_diff > 0 ? (real)1 : (real)0;
    real not_in_y = y_above + (mean[1] > rmaxy ? (real)1 : (real)0);
    if (not_in_y + not_in_x == 0)
        return 1;
    real ccx = rmaxx + x_left * (rminx - rmaxx), ccy = rmaxy + y_above * (rminy - rmaxy); /* fast_lerp */
    real diffx = mean[0] - ccx, diffy = mean[1] - ccy;
    real dx = (real)copysign((double)(FG_TILE - 1), (double)x_min_diff), dy = (real)copysign((double)(FG_TILE - 1), (double)y_min_diff);
    real tx = (dx * conic[0] * diffx + dx * conic[1] * diffy) / (dx * conic[0] * dx);
    real ty = (dy * conic[1] * diffx + dy * conic[2] * diffy) / (dy * conic[2] * dy);
    tx = tx < 0 ? 0 : (tx > 1 ? 1 : tx); /* __saturatef (NaN -> 0) */
    ty = ty < 0 ? 0 : (ty > 1 ? 1 : ty);
    if (tx != tx)
        tx = 0;
    if (ty != ty)
        ty = 0;
    tx *= not_in_y, ty *= not_in_x;
    real px = ccx + tx * dx, py = ccy + ty * dy;
    real ex = mean[0] - px, ey = mean[1] - py;
    real max_power = (real)0.5 * (conic[0] * ex * ex + conic[2] * ey * ey) + conic[1] * ex * ey;
    return max_power <= power_threshold;
}

typedef struct {
    real rot[3][3], rs[3][3], cov3d[6], variance[3];
    real qn2, q[4], q2[9]; /* q2 = qxx qyy qzz qxy qxz qyz qrx qry qrz (scaled by 2/|q|^2) */
    real jw1[3], jw2[3], jwc1[3], jwc2[3];
    real x, y, tx, ty, j11, j13, j22, j23, depth;
    real a, b, c; /* dilated 2-D covariance */
} ORC(fg_geom);

/* shared by forward (include/kernels_forward.cuh:60-147) and backward (include/kernels_backward.cuh:56-126) */
static void ORC(fg_geometry)(const real* mean, const real* rs_, const real* rq, const real* w2c, real w, real h, real fx,
                             real fy, real cx, real cy, ORC(fg_geom) * G) {
    const real* r1 = w2c;
    const real* r2 = w2c + 4;
    const real* r3 = w2c + 8;
    G->depth = r3[0] * mean[0] + r3[1] * mean[1] + r3[2] * mean[2] + r3[3];
    for (int k = 0; k < 3; ++k)
        G->variance[k] = R_EXP((real)2 * rs_[k]);
    real qr = rq[0], qx = rq[1], qy = rq[2], qz = rq[3];
    real n2 = qr * qr + qx * qx + qy * qy + qz * qz;
    G->qn2 = n2;
    G->q[0] = qr, G->q[1] = qx, G->q[2] = qy, G->q[3] = qz;
    real qxx = 2 * qx * qx / n2, qyy = 2 * qy * qy / n2, qzz = 2 * qz * qz / n2;
    real qxy = 2 * qx * qy / n2, qxz = 2 * qx * qz / n2, qyz = 2 * qy * qz / n2;
    real qrx = 2 * qr * qx / n2, qry = 2 * qr * qy / n2, qrz = 2 * qr * qz / n2;
    real q2[9] = {qxx, qyy, qzz, qxy, qxz, qyz, qrx, qry, qrz};
    for (int k = 0; k < 9; ++k)
        G->q2[k] = q2[k];
    real R[3][3] = {{1 - (qyy + qzz), qxy - qrz, qry + qxz}, {qrz + qxy, 1 - (qxx + qzz), qyz - qrx}, {qxz - qry, qrx + qyz, 1 - (qxx + qyy)}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            G->rot[i][j] = R[i][j];
            G->rs[i][j] = R[i][j] * G->variance[j];
        }
    int t = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            G->cov3d[t++] = G->rs[i][0] * R[j][0] + G->rs[i][1] * R[j][1] + G->rs[i][2] * R[j][2]; /* m11 m12 m13 m22 m23 m33 */
    G->x = (r1[0] * mean[0] + r1[1] * mean[1] + r1[2] * mean[2] + r1[3]) / G->depth;
    G->y = (r2[0] * mean[0] + r2[1] * mean[1] + r2[2] * mean[2] + r2[3]) / G->depth;
    real cl = ((real)-0.15 * w - cx) / fx, cr = ((real)1.15 * w - cx) / fx;
    real ct = ((real)-0.15 * h - cy) / fy, cb = ((real)1.15 * h - cy) / fy;
    G->tx = G->x < cl ? cl : (G->x > cr ? cr : G->x);
    G->ty = G->y < ct ? ct : (G->y > cb ? cb : G->y);
    G->j11 = fx / G->depth, G->j13 = -G->j11 * G->tx, G->j22 = fy / G->depth, G->j23 = -G->j22 * G->ty;
    for (int k = 0; k < 3; ++k) {
        G->jw1[k] = G->j11 * r1[k] + G->j13 * r3[k];
        G->jw2[k] = G->j22 * r2[k] + G->j23 * r3[k];
    }
    const real* S = G->cov3d;
    G->jwc1[0] = G->jw1[0] * S[0] + G->jw1[1] * S[1] + G->jw1[2] * S[2];
    G->jwc1[1] = G->jw1[0] * S[1] + G->jw1[1] * S[3] + G->jw1[2] * S[4];
    G->jwc1[2] = G->jw1[0] * S[2] + G->jw1[1] * S[4] + G->jw1[2] * S[5];
    G->jwc2[0] = G->jw2[0] * S[0] + G->jw2[1] * S[1] + G->jw2[2] * S[2];
    G->jwc2[1] = G->jw2[0] * S[1] + G->jw2[1] * S[3] + G->jw2[2] * S[4];
    G->jwc2[2] = G->jw2[0] * S[2] + G->jw2[1] * S[4] + G->jw2[2] * S[5];
    G->a = ORC(dot3)(G->jwc1, G->jw1) + (real)FG_DILATION;
    G->b = ORC(dot3)(G->jwc1, G->jw2);
    G->c = ORC(dot3)(G->jwc2, G->jw2) + (real)FG_DILATION;
}

static int ORC(fg_cmp_keys)(const void* pa, const void* pb) {
    const uint64_t a = *(const uint64_t*)pa, b = *(const uint64_t*)pb;
    return a < b ? -1 : (a > b ? 1 : 0);
}

/* Whole pipeline. grad_image == NULL -> forward only.
 * Layouts follow the reference tensors: means [N,3], scales_raw [N,3], rotations_raw [N,4] (w,x,y,z), opacities_raw [N],
 * sh0 [N,3], shN [N,total_rest,3], w2c [4,4] row-major, image [3,H,W], alpha [H,W].
 * densification_info [2,N] is accumulated (+=) like include/kernels_backward.cuh:233-236. */
int64_t ORC(fastgs)(int N, const real* means, const real* scales_raw, const real* rotations_raw, const real* opacities_raw,
                    const real* sh0, const real* shN, int total_rest, const real* w2c, const real* cam_pos, int active,
                    int W, int H, real fx, real fy, real cx, real cy, real near_, real far_, real* image, real* alpha,
                    int32_t* n_touched_out, const real* grad_image, const real* grad_alpha, real* g_means, real* g_scales,
                    real* g_rot, real* g_opac, real* g_sh0, real* g_shN, real* g_w2c, real* densification_info) {
    const int gw = (W + FG_TILE - 1) / FG_TILE, gh = (H + FG_TILE - 1) / FG_TILE, n_tiles = gw * gh;
    ORC(fg_prim)* P = (ORC(fg_prim)*)calloc((size_t)N, sizeof(ORC(fg_prim)));
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
    int n_vis = 0;
    /* ---- preprocess_cu, include/kernels_forward.cuh:18-205 */
    for (int i = 0; i < N; ++i) {
        ORC(fg_prim)* p = &P[i];
        const real* mean = means + 3 * (size_t)i;
        ORC(fg_geom) G;
        ORC(fg_geometry)(mean, scales_raw + 3 * (size_t)i, rotations_raw + 4 * (size_t)i, w2c, (real)W, (real)H, fx, fy, cx, cy, &G);
        int act = 1;
        if (G.depth < near_ || G.depth > far_) /* :61-62 */
            act = 0;
        real opacity = (real)1 / ((real)1 + R_EXP(-opacities_raw[i]));
        if (opacity < (real)(1.0 / FG_MIN_ALPHA_RCP)) /* :75 */
            act = 0;
        if (G.qn2 < (real)1e-8) /* :84 */
            act = 0;
        real det = G.a * G.c - G.b * G.b;
        if (det < (real)1e-8) /* :146 */
            act = 0;
        if (!act || !(det == det))
            continue;
        p->conic[0] = G.c / det, p->conic[1] = -G.b / det, p->conic[2] = G.a / det;
        p->mean2d[0] = G.x * fx + cx, p->mean2d[1] = G.y * fy + cy;
        real pt = R_LOG(opacity * (real)FG_MIN_ALPHA_RCP); /* :159 */
        real ptf = R_SQRT((real)2 * pt);
        real ex = ptf * R_SQRT(G.a) - (real)0.5, ey = ptf * R_SQRT(G.c) - (real)0.5;
        ex = ex > 0 ? ex : 0, ey = ey > 0 ? ey : 0;
        int x0 = (int)floor((double)((p->mean2d[0] - ex) / (real)FG_TILE)), x1 = (int)ceil((double)((p->mean2d[0] + ex) / (real)FG_TILE));
        int y0 = (int)floor((double)((p->mean2d[1] - ey) / (real)FG_TILE)), y1 = (int)ceil((double)((p->mean2d[1] + ey) / (real)FG_TILE));
        x0 = x0 < 0 ? 0 : x0, x1 = x1 < 0 ? 0 : x1, y0 = y0 < 0 ? 0 : y0, y1 = y1 < 0 ? 0 : y1;
        x0 = x0 > gw ? gw : x0, x1 = x1 > gw ? gw : x1, y0 = y0 > gh ? gh : y0, y1 = y1 > gh ? gh : y1;
        if ((x1 - x0) * (y1 - y0) <= 0)
            continue;
        const real ms[2] = {p->mean2d[0] - (real)0.5, p->mean2d[1] - (real)0.5}; /* include/kernel_utils.cuh:152 */
        int nt = 0;
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx)
                nt += ORC(fg_will_contribute)(ms, p->conic, tx, ty, pt);
        if (nt == 0)
            continue;
        p->visible = 1, p->n_touched = nt, p->opacity = opacity, p->depth = G.depth;
        p->x0 = x0, p->x1 = x1, p->y0 = y0, p->y1 = y1;
        ORC(fg_sh_to_color)(sh0, shN, total_rest, mean, cam_pos, i, active, p->color);
        float df = (float)G.depth;
        memcpy(&p->depth_key, &df, 4);
        keys[n_vis++] = ((uint64_t)p->depth_key << 32) | (uint32_t)i;
    }
    if (n_touched_out)
        for (int i = 0; i < N; ++i)
            n_touched_out[i] = P[i].n_touched;
    /* ---- depth sort (src/forward.cu:103-108), ties by index */
    qsort(keys, (size_t)n_vis, sizeof(uint64_t), ORC(fg_cmp_keys));
    /* ---- create_instances_cu + tile sort (include/kernels_forward.cuh:221-320, src/forward.cu:141-147) as per-tile lists */
    int64_t* tcount = (int64_t*)calloc((size_t)n_tiles + 1, sizeof(int64_t));
    for (int s = 0; s < n_vis; ++s) {
        const ORC(fg_prim)* p = &P[(uint32_t)keys[s]];
        const real ms[2] = {p->mean2d[0] - (real)0.5, p->mean2d[1] - (real)0.5};
        const real pt = R_LOG(p->opacity * (real)FG_MIN_ALPHA_RCP);
        for (int ty = p->y0; ty < p->y1; ++ty)
            for (int tx = p->x0; tx < p->x1; ++tx)
                if (ORC(fg_will_contribute)(ms, p->conic, tx, ty, pt))
                    tcount[ty * gw + tx + 1]++;
    }
    for (int t = 0; t < n_tiles; ++t)
        tcount[t + 1] += tcount[t];
    const int64_t n_inst = tcount[n_tiles];
    int32_t* inst = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_inst > 0 ? n_inst : 1));
    int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_tiles + 1));
    memcpy(cursor, tcount, sizeof(int64_t) * (size_t)(n_tiles + 1));
    for (int s = 0; s < n_vis; ++s) {
        const int idx = (int)(uint32_t)keys[s];
        const ORC(fg_prim)* p = &P[idx];
        const real ms[2] = {p->mean2d[0] - (real)0.5, p->mean2d[1] - (real)0.5};
        const real pt = R_LOG(p->opacity * (real)FG_MIN_ALPHA_RCP);
        for (int ty = p->y0; ty < p->y1; ++ty)
            for (int tx = p->x0; tx < p->x1; ++tx)
                if (ORC(fg_will_contribute)(ms, p->conic, tx, ty, pt))
                    inst[cursor[ty * gw + tx]++] = idx;
    }
    /* per-primitive blend gradients (helpers of src/rasterization_api.cu:126-134) */
    const int bwd = grad_image != NULL;
    real* gm2 = NULL;
    real* gcon = NULL;
    real* gcol = NULL;
    real* gop = NULL;
    if (bwd) {
        gm2 = (real*)calloc((size_t)2 * N + 1, sizeof(real));
        gcon = (real*)calloc((size_t)3 * N + 1, sizeof(real));
        gcol = (real*)calloc((size_t)3 * N + 1, sizeof(real));
        gop = (real*)calloc((size_t)N + 1, sizeof(real));
    }
    const size_t npix = (size_t)W * H;
    /* ---- blend_cu (include/kernels_forward.cuh:356-459) and blend_backward_cu (include/kernels_backward.cuh:240-449) */
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < n_tiles; ++tile) {
        const int tyi = tile / gw, txi = tile % gw;
        const int64_t s0 = tcount[tile], s1 = tcount[tile + 1];
        for (int py = tyi * FG_TILE; py < (tyi + 1) * FG_TILE && py < H; ++py)
            for (int px = txi * FG_TILE; px < (txi + 1) * FG_TILE && px < W; ++px) {
                const real pxf = (real)px + (real)0.5, pyf = (real)py + (real)0.5;
                real col[3] = {0, 0, 0}, T = 1;
                int64_t n_contrib = 0;
                for (int64_t j = s0; j < s1; ++j) {
                    const ORC(fg_prim)* p = &P[inst[j]];
                    const real dx = p->mean2d[0] - pxf, dy = p->mean2d[1] - pyf;
                    const real sig = (real)0.5 * (p->conic[0] * dx * dx + p->conic[2] * dy * dy) + p->conic[1] * dx * dy;
                    if (sig < 0)
                        continue;
                    real a = p->opacity * R_EXP(-sig);
                    a = a < (real)FG_MAX_ALPHA ? a : (real)FG_MAX_ALPHA;
                    if (a < (real)(1.0 / FG_MIN_ALPHA_RCP))
                        continue;
                    const real nT = T * ((real)1 - a);
                    if (nT < (real)FG_T_THRESHOLD)
                        break; /* done = true */
                    for (int k = 0; k < 3; ++k) {
                        const real c = p->color[k] > 0 ? p->color[k] : 0; /* fmaxf(color, 0), :413 */
                        col[k] += T * a * c;
                    }
                    T = nT;
                    n_contrib = j - s0 + 1;
                }
                const size_t pi = (size_t)py * W + px;
                image[pi] = col[0], image[npix + pi] = col[1], image[2 * npix + pi] = col[2];
                alpha[pi] = (real)1 - T;
                if (!bwd)
                    continue;
                /* backward over the same list, front to back, up to the last contributor */
                const real gp[3] = {grad_image[pi], grad_image[npix + pi], grad_image[2 * npix + pi]};
                const real ga_common = grad_alpha[pi] * ((real)1 - alpha[pi]); /* :343-345 */
                real after[3] = {col[0], col[1], col[2]}, Tb = 1;
                for (int64_t j = s0; j < s0 + n_contrib; ++j) {
                    const int idx = inst[j];
                    const ORC(fg_prim)* p = &P[idx];
                    const real dx = p->mean2d[0] - pxf, dy = p->mean2d[1] - pyf;
                    const real sig = (real)0.5 * (p->conic[0] * dx * dx + p->conic[2] * dy * dy) + p->conic[1] * dx * dy;
                    if (sig < 0)
                        continue;
                    real a = p->opacity * R_EXP(-sig);
                    a = a < (real)FG_MAX_ALPHA ? a : (real)FG_MAX_ALPHA;
                    if (a < (real)(1.0 / FG_MIN_ALPHA_RCP))
                        continue;
                    const real oma = (real)1 - a, wgt = Tb * a;
                    real c[3], dcol[3];
                    for (int k = 0; k < 3; ++k) {
                        c[k] = p->color[k] > 0 ? p->color[k] : 0;
                        dcol[k] = wgt * gp[k] * (p->color[k] >= 0 ? (real)1 : (real)0); /* :409-410 */
                        after[k] -= wgt * c[k];
                    }
                    const real rcp = (real)1 / oma;
                    real dLda = 0;
                    for (int k = 0; k < 3; ++k)
                        dLda += (Tb * c[k] - after[k] * rcp) * gp[k];
                    dLda += ga_common * rcp;
                    const real helper = -a * dLda; /* :425 */
                    const real dcon[3] = {(real)0.5 * helper * dx * dx, (real)0.5 * helper * dx * dy, (real)0.5 * helper * dy * dy};
                    const real dm[2] = {helper * (p->conic[0] * dx + p->conic[1] * dy), helper * (p->conic[1] * dx + p->conic[2] * dy)};
#pragma omp atomic
                    gop[idx] += a * dLda;
                    for (int k = 0; k < 3; ++k) {
#pragma omp atomic
                        gcol[3 * (size_t)idx + k] += dcol[k];
#pragma omp atomic
                        gcon[3 * (size_t)idx + k] += dcon[k];
                    }
                    for (int k = 0; k < 2; ++k) {
#pragma omp atomic
                        gm2[2 * (size_t)idx + k] += dm[k];
                    }
                    Tb *= oma;
                }
            }
    }
    /* ---- preprocess_backward_cu, include/kernels_backward.cuh:18-237 */
    if (bwd) {
        for (int i = 0; i < N; ++i) {
            if (P[i].n_touched == 0)
                continue;
            const real* mean = means + 3 * (size_t)i;
            real dpos_col[3];
            ORC(fg_sh_backward)(shN, total_rest, mean, cam_pos, i, active, gcol + 3 * (size_t)i, g_sh0, g_shN, dpos_col);
            ORC(fg_geom) G;
            ORC(fg_geometry)(mean, scales_raw + 3 * (size_t)i, rotations_raw + 4 * (size_t)i, w2c, (real)W, (real)H, fx, fy, cx, cy, &G);
            const real a = G.a, b = G.b, c = G.c;
            const real det = a * c - b * b, dr = (real)1 / det, dr2 = dr * dr;
            const real* dc = gcon + 3 * (size_t)i;
            const real dcov[3] = {dr2 * ((real)2 * b * c * dc[1] - c * c * dc[0] - b * b * dc[2]),
                                  dr2 * (b * c * dc[0] - (a * c + b * b) * dc[1] + a * b * dc[2]),
                                  dr2 * ((real)2 * a * b * dc[1] - b * b * dc[0] - a * a * dc[2])};
            const real *u = G.jw1, *v = G.jw2;
            real dS[6];
            {
                int t = 0;
                for (int r = 0; r < 3; ++r)
                    for (int s = r; s < 3; ++s) {
                        dS[t++] = (r == s) ? (u[r] * u[r] * dcov[0] + (real)2 * u[r] * v[r] * dcov[1] + v[r] * v[r] * dcov[2])
                                           : (u[r] * u[s] * dcov[0] + (u[r] * v[s] + u[s] * v[r]) * dcov[1] + v[r] * v[s] * dcov[2]);
                    }
            }
            real djw1[3], djw2[3];
            for (int k = 0; k < 3; ++k) {
                djw1[k] = (real)2 * (G.jwc1[k] * dcov[0] + G.jwc2[k] * dcov[1]);
                djw2[k] = (real)2 * (G.jwc1[k] * dcov[1] + G.jwc2[k] * dcov[2]);
            }
            const real* r1 = w2c;
            const real* r2 = w2c + 4;
            const real* r3 = w2c + 8;
            const real dj11 = ORC(dot3)(r1, djw1), dj22 = ORC(dot3)(r2, djw2), dj13 = ORC(dot3)(r3, djw1), dj23 = ORC(dot3)(r3, djw2);
            const real h1 = dj11 - (real)2 * G.tx * dj13, h2 = dj22 - (real)2 * G.ty * dj23;
            const real* dm2 = gm2 + 2 * (size_t)i;
            const real dcam[3] = {G.j11 * (dm2[0] - dj13 / G.depth), G.j22 * (dm2[1] - dj23 / G.depth),
                                  -G.j11 * (G.x * dm2[0] + h1 / G.depth) - G.j22 * (G.y * dm2[1] + h2 / G.depth)};
            if (g_w2c) {
                for (int r = 0; r < 3; ++r) {
                    g_w2c[4 * r + 3] += dcam[r];
                    for (int k = 0; k < 3; ++k)
                        g_w2c[4 * r + k] += dcam[r] * mean[k];
                }
            }
            for (int k = 0; k < 3; ++k)
                g_means[3 * (size_t)i + k] = r1[k] * dcam[0] + r2[k] * dcam[1] + r3[k] * dcam[2] + dpos_col[k];
            /* scales :199-209; dS order m11 m12 m13 m22 m23 m33 */
            for (int k = 0; k < 3; ++k) {
                const real R0 = G.rot[0][k], R1 = G.rot[1][k], R2 = G.rot[2][k];
                const real dvar = R0 * R0 * dS[0] + R1 * R1 * dS[3] + R2 * R2 * dS[5] + (real)2 * (R0 * R1 * dS[1] + R0 * R2 * dS[2] + R1 * R2 * dS[4]);
                g_scales[3 * (size_t)i + k] = (real)2 * G.variance[k] * dvar;
            }
            /* rotation :212-232 */
            const real Sm[3][3] = {{dS[0], dS[1], dS[2]}, {dS[1], dS[3], dS[4]}, {dS[2], dS[4], dS[5]}};
            real dR[3][3];
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k)
                    dR[r][k] = (real)2 * (G.rs[0][k] * Sm[r][0] + G.rs[1][k] * Sm[r][1] + G.rs[2][k] * Sm[r][2]);
            const real dqxx = -dR[1][1] - dR[2][2], dqyy = -dR[0][0] - dR[2][2], dqzz = -dR[0][0] - dR[1][1];
            const real dqxy = dR[0][1] + dR[1][0], dqxz = dR[0][2] + dR[2][0], dqyz = dR[1][2] + dR[2][1];
            const real dqrx = dR[2][1] - dR[1][2], dqry = dR[0][2] - dR[2][0], dqrz = dR[1][0] - dR[0][1];
            const real* q2 = G.q2;
            const real nh = q2[0] * dqxx + q2[1] * dqyy + q2[2] * dqzz + q2[3] * dqxy + q2[4] * dqxz + q2[5] * dqyz + q2[6] * dqrx + q2[7] * dqry + q2[8] * dqrz;
            const real qr = G.q[0], qx = G.q[1], qy = G.q[2], qz = G.q[3];
            real* gq = g_rot + 4 * (size_t)i;
            gq[0] = (real)2 * (qx * dqrx + qy * dqry + qz * dqrz - qr * nh) / G.qn2;
            gq[1] = (real)2 * ((real)2 * qx * dqxx + qy * dqxy + qz * dqxz + qr * dqrx - qx * nh) / G.qn2;
            gq[2] = (real)2 * ((real)2 * qy * dqyy + qx * dqxy + qz * dqyz + qr * dqry - qy * nh) / G.qn2;
            gq[3] = (real)2 * ((real)2 * qz * dqzz + qx * dqxz + qy * dqyz + qr * dqrz - qz * nh) / G.qn2;
            g_opac[i] = gop[i] * ((real)1 - P[i].opacity); /* include/kernels_backward.cuh:441-442 */
            if (densification_info) {
                densification_info[i] += (real)1;
                const real sx = dm2[0] * (real)0.5 * (real)W, sy = dm2[1] * (real)0.5 * (real)H;
                densification_info[(size_t)N + i] += R_SQRT(sx * sx + sy * sy);
            }
        }
        free(gm2), free(gcon), free(gcol), free(gop);
    }
    free(P), free(keys), free(tcount), free(inst), free(cursor);
    return n_inst;
}

#undef R_SQRT
#undef R_EXP
#undef R_LOG
#undef ORC
#undef real
#undef CAT
#undef CAT_
