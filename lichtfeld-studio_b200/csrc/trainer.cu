// lfs_b200 -- fused training step (one view): raw parameters -> image -> loss gradient -> raw-parameter
// gradients, with no host synchronisation and no intermediate torch ops.  It plays the role of the
// reference's gs::training::rasterize + autograd glue (src/training/rasterization/rasterizer.cpp:46-437,
// rasterizer_autograd.cpp:12-392) for the 3DGUT path, with the activations of SplatData
// (src/core/splat_data.cpp:267-286: sigmoid / exp / normalize / cat) and the SH evaluation fused into the
// per-Gaussian kernels, as the reference's fastgs preprocess does for its EWA path
// (fastgs/rasterization/include/kernels_forward.cuh:18-205).
//
// Parameter / gradient / Adam-state arenas are PLANAR (SoA): plane p holds N_pad floats, planes in the
// reference's parameter-group order (strategies/strategy_utils.cpp:35-40):
//   means x,y,z | sh0 r,g,b | shN (k=1..K-1) x (r,g,b) | scaling x,y,z | rotation w,x,y,z | opacity
// so that one thread per Gaussian reads/writes every plane fully coalesced and the fused Adam / the NCCL
// all-reduce see one flat fp32 buffer of (11 + 3K) * N_pad floats (59 N at SH degree 3).
#include "intersect.cuh"
#include "projection.cuh"
#include "raster.cuh"
#include "sh.cuh"
#include "sort_scan.cuh"

#include <new>

namespace lfs {

struct Planes {
    uint32_t Np, K;
    __host__ __device__ uint32_t mean(int c) const { return (uint32_t)c; }
    __host__ __device__ uint32_t sh(int k, int ch) const { return k == 0 ? 3u + ch : 6u + (k - 1) * 3u + ch; }
    __host__ __device__ uint32_t scaling(int c) const { return 6u + 3u * (K - 1) + c; }
    __host__ __device__ uint32_t rotation(int c) const { return 9u + 3u * (K - 1) + c; }
    __host__ __device__ uint32_t opacity() const { return 13u + 3u * (K - 1); }
    __host__ __device__ uint32_t count() const { return 11u + 3u * K; }
};

struct Trainer {
    lfs_trainer_desc d;
    Planes pl;
    uint32_t tile_w, tile_h, n_tiles;
    uint32_t inst_cap, bucket_cap;
    // device buffers (one cudaMalloc'd blob)
    char* blob = nullptr;
    size_t blob_bytes = 0;
    ViewCam* cam_dev;
    GaussRec* gauss;
    TileRect* rects;
    CullRec* cull;
    int32_t* counts;
    uint32_t *dk_a, *dk_b, *pm_a, *pm_b, *off;
    uint32_t* n_inst; // [4] device counters: n_inst, n_buckets, high-water mark of n_inst since creation
    uint32_t *tk_a, *tk_b, *tv_a, *tv_b;
    int32_t* tile_off;
    uint32_t *bucket_off, *bucket_counts, *bucket_tile, *tile_max, *live;
    float4* ckpt;
    float4* pix_state;
    int32_t* n_contrib;
    float4* v_pix;
    float *v_means, *v_quats, *v_scales, *v_colors, *v_opac;
    float *act_means, *act_quats, *act_scales; // activated AoS copies for the blend-backward epilogue
    float* loss_partials;
    float* ssim_maps; // [3 channels][H*W] float4: dL/dmap * (dm/dmu1, dm/dsigma1^2, dm/dsigma12, -)
    void *scan_scr, *sort_scr;
    const uint32_t* sorted_keys = nullptr;
    const uint32_t* sorted_vals = nullptr;
    ViewCam cam_host;
    float bg[3];
    uint32_t active_degree = 0;
    uint32_t* stats_host = nullptr; // pinned: n_inst, n_buckets
    // optional per-stage CUDA-event timing (bench.py roofline): see lfs_trainer_set_profile
    int profile = 0;
    cudaEvent_t ev[LFS_PROF_STAGES + 2] = {};
    bool ev_ok = false;
    float acc_ms[LFS_PROF_STAGES] = {};
    int acc_n[LFS_PROF_STAGES] = {};
    int pending_from = -1, pending_to = -1;
    void mark(int i, cudaStream_t st) {
        if (profile && ev_ok)
            cudaEventRecord(ev[i], st);
    }
};

static size_t trainer_carve(Trainer& t, void* base) {
    Carver c(base);
    const uint32_t N = t.d.n_gaussians;
    const uint64_t npix = (uint64_t)t.d.width * t.d.height;
    t.cam_dev = c.take<ViewCam>(1);
    t.gauss = c.take<GaussRec>(N);
    t.rects = c.take<TileRect>(N);
    t.cull = c.take<CullRec>(N);
    t.counts = c.take<int32_t>(N);
    t.dk_a = c.take<uint32_t>(N), t.dk_b = c.take<uint32_t>(N);
    t.pm_a = c.take<uint32_t>(N), t.pm_b = c.take<uint32_t>(N);
    t.off = c.take<uint32_t>(N);
    t.n_inst = c.take<uint32_t>(4);
    t.tk_a = c.take<uint32_t>(t.inst_cap), t.tk_b = c.take<uint32_t>(t.inst_cap);
    t.tv_a = c.take<uint32_t>(t.inst_cap), t.tv_b = c.take<uint32_t>(t.inst_cap);
    t.tile_off = c.take<int32_t>(t.n_tiles + 1);
    t.bucket_off = c.take<uint32_t>(t.n_tiles + 1);
    t.bucket_counts = c.take<uint32_t>(t.n_tiles + 1);
    t.bucket_tile = c.take<uint32_t>(t.bucket_cap);
    t.live = c.take<uint32_t>(live_list_words(t.bucket_cap));
    t.tile_max = c.take<uint32_t>(t.n_tiles);
    t.ckpt = c.take<float4>((size_t)t.bucket_cap * kTilePix);
    t.pix_state = c.take<float4>(npix);
    t.n_contrib = c.take<int32_t>(npix);
    t.v_pix = c.take<float4>(npix);
    t.v_means = c.take<float>(3 * (size_t)N);
    t.v_quats = c.take<float>(4 * (size_t)N);
    t.v_scales = c.take<float>(3 * (size_t)N);
    t.v_colors = c.take<float>(3 * (size_t)N);
    t.v_opac = c.take<float>(N);
    t.act_means = c.take<float>(3 * (size_t)N);
    t.act_quats = c.take<float>(4 * (size_t)N);
    t.act_scales = c.take<float>(3 * (size_t)N);
    t.loss_partials = c.take<float>(2 * (size_t)t.n_tiles + 2);
    t.ssim_maps = c.take<float>(12 * (size_t)npix);
    t.scan_scr = c.take<char>(scan_scratch_bytes(N > t.n_tiles + 1 ? N : t.n_tiles + 1));
    t.sort_scr = c.take<char>(radix_scratch_bytes(N > t.inst_cap ? N : t.inst_cap));
    return c.total();
}

// ------------------------------------------------------------------------------------------------------
// AoS (reference SplatData layout) <-> planar arena
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_pack_arena(const float* __restrict__ means, const float* __restrict__ sh0, const float* __restrict__ shN,
                 const float* __restrict__ scaling, const float* __restrict__ rotation,
                 const float* __restrict__ opacity, const uint32_t N, const Planes pl, float* __restrict__ arena,
                 const int to_arena) {
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    if (g >= N)
        return;
    const size_t Np = pl.Np;
#define MOVE(aos_ptr, aos_index, plane)                                   \
    do {                                                                  \
        float* a_ = const_cast<float*>(aos_ptr) + (aos_index);            \
        float* p_ = arena + (size_t)(plane) * Np + g;                     \
        if (to_arena)                                                     \
            *p_ = *a_;                                                    \
        else                                                              \
            *a_ = *p_;                                                    \
    } while (0)
    for (int c = 0; c < 3; ++c)
        MOVE(means, 3 * (size_t)g + c, pl.mean(c));
    for (int c = 0; c < 3; ++c)
        MOVE(sh0, 3 * (size_t)g + c, pl.sh(0, c));
    for (uint32_t k = 1; k < pl.K; ++k)
        for (int c = 0; c < 3; ++c)
            MOVE(shN, ((size_t)g * (pl.K - 1) + (k - 1)) * 3 + c, pl.sh((int)k, c));
    for (int c = 0; c < 3; ++c)
        MOVE(scaling, 3 * (size_t)g + c, pl.scaling(c));
    for (int c = 0; c < 4; ++c)
        MOVE(rotation, 4 * (size_t)g + c, pl.rotation(c));
    MOVE(opacity, (size_t)g, pl.opacity());
#undef MOVE
}

// ------------------------------------------------------------------------------------------------------
// per-Gaussian forward: activations + UT projection + SH colour + GaussRec + tile rectangle + depth key
// ------------------------------------------------------------------------------------------------------
struct PreCfg {
    float eps2d, near_plane, far_plane, radius_clip;
    lfs_ut_params ut;
    int degree;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256)
    k_preprocess_fwd(const float* __restrict__ arena, const Planes pl, const uint32_t N, const ViewCam cam,
                     const PreCfg cfg, GaussRec* __restrict__ gauss, TileRect* __restrict__ rects,
                     CullRec* __restrict__ cull /* null: keep every tile of the AABB */, int32_t* __restrict__ counts, uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ ident,
                     float* __restrict__ act_means, float* __restrict__ act_quats, float* __restrict__ act_scales) {
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    if (g >= N)
        return;
    const size_t Np = pl.Np;
    auto P = [&](uint32_t plane) { return __ldg(arena + (size_t)plane * Np + g); };
    // The SH planes are only needed after the (long) projection and only for visible Gaussians; pull them towards
    // L2 now so that second phase does not pay a full DRAM round trip (costs no registers, unlike early loads).
    {
        const int nb = (cfg.degree + 1) * (cfg.degree + 1);
        for (int k = 0; k < nb; ++k)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(arena + (size_t)pl.sh(k, ch) * Np + g));
    }
    const f3 mean = mk3(P(pl.mean(0)), P(pl.mean(1)), P(pl.mean(2)));
    const f3 scale = mk3(__expf(P(pl.scaling(0))), __expf(P(pl.scaling(1))), __expf(P(pl.scaling(2))));
    float qw = P(pl.rotation(0)), qx = P(pl.rotation(1)), qy = P(pl.rotation(2)), qz = P(pl.rotation(3));
    { // torch::nn::functional::normalize (eps 1e-12), splat_data.cpp:279
        const float nrm = fmaxf(sqrtf(qw * qw + qx * qx + qy * qy + qz * qz), 1e-12f);
        const float inv = 1.0f / nrm;
        qw *= inv, qx *= inv, qy *= inv, qz *= inv;
    }
    const float op = sigmoidf_(P(pl.opacity()));

    const UTOut o = ut_project_pinhole(cam, mean, make_float4(qw, qx, qy, qz), scale, true, op, cfg.eps2d,
                                       cfg.near_plane, cfg.far_plane, cfg.radius_clip, cfg.ut);
    ident[g] = g;
    if (!o.ok) {
        counts[g] = 0;
        rects[g] = TileRect{0, 0, 0, 0};
        depth_keys[g] = 0xFFFFFFFFu;
        return;
    }
    uint32_t x0, y0, x1, y1;
    tile_rect(o.mx, o.my, o.rx, o.ry, (float)kTile, (uint32_t)cam.tile_w, (uint32_t)cam.tile_h, x0, y0, x1, y1);
    int32_t cnt = (int32_t)((y1 - y0) * (x1 - x0));

    // GaussRec geometry (same algebra as k_prep_gaussians)
    float w = qw, x = qx, y = qy, z = qz;
    const float inv_norm = rsqrtf(x * x + y * y + z * z + w * w);
    x *= inv_norm, y *= inv_norm, z *= inv_norm, w *= inv_norm;
    const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y,
                wz = w * z;
    const f3 m0 = mk3(1.f - 2.f * (y2 + z2), 2.f * (xy + wz), 2.f * (xz - wy)) * (1.0f / scale.x);
    const f3 m1 = mk3(2.f * (xy - wz), 1.f - 2.f * (x2 + z2), 2.f * (yz + wx)) * (1.0f / scale.y);
    const f3 m2 = mk3(2.f * (xz + wy), 2.f * (yz - wx), 1.f - 2.f * (x2 + y2)) * (1.0f / scale.z);
    const f3 r0 = mk3(cam.R[0], cam.R[1], cam.R[2]), r1 = mk3(cam.R[3], cam.R[4], cam.R[5]),
             r2 = mk3(cam.R[6], cam.R[7], cam.R[8]);
    const float ifx = 1.0f / cam.fx, ify = 1.0f / cam.fy;
    const f3 omu = mk3(cam.org[0] - mean.x, cam.org[1] - mean.y, cam.org[2] - mean.z);
    const f3 vx = mk3(dot(m0, r0), dot(m1, r0), dot(m2, r0)) * ifx;
    const f3 vy = mk3(dot(m0, r1), dot(m1, r1), dot(m2, r1)) * ify;
    const f3 w2 = mk3(dot(m0, r2), dot(m1, r2), dot(m2, r2));
    const f3 gro = mk3(dot(m0, omu), dot(m1, omu), dot(m2, omu));

    if (cull && cnt > 0) {
        // Q = N - tau D with N = |v x gro|^2, D = |v|^2, v = vx X + vy Y + w0, expanded around the PROJECTED CENTRE
        // (X, Y) = pixel - (mx, my).  Like the tile expansion of the blend kernels this keeps every term O(result):
        // N is built from the cross products (never as |v|^2 |gro|^2 - <v,gro>^2, which cancels catastrophically for
        // |gro| ~ 1e3), and w0 x gro is small because the ray through the projected centre almost hits the Gaussian.
        // (First version expanded around the principal point: bit-exact on 20k Gaussians, NOT on the 1 M scene.)
        const float tau = 2.0f * __logf(255.0f * op) * 1.001f + 0.01f; // margin: never culls a contributing pixel
        const float X0 = o.mx - cam.cx, Y0 = o.my - cam.cy;
        const f3 w0 = w2 + vx * X0 + vy * Y0;
        const f3 cxv = cross(vx, gro), cyv = cross(vy, gro), c0v = cross(w0, gro);
        const float qa = dot(cxv, cxv) - tau * dot(vx, vx), qb = dot(cxv, cyv) - tau * dot(vx, vy);
        const float qc = dot(cyv, cyv) - tau * dot(vy, vy), qd = dot(cxv, c0v) - tau * dot(vx, w0);
        const float qe = dot(cyv, c0v) - tau * dot(vy, w0), qf = dot(c0v, c0v) - tau * dot(w0, w0);
        const float det = qa * qc - qb * qb;
        CullRec cr{1.f, 0.f, 1.f, 0.f, 0.f, __int_as_float(0x7f800000), 1.f, 1.f}; // lim = +inf: keep everything
        if (qa > 0.f && qc > 0.f && det > 1e-12f * qa * qc) {
            const float Xc = (qb * qe - qc * qd) / det, Yc = (qb * qd - qa * qe) / det; // minimiser of Q
            const float qmin = qf + qd * Xc + qe * Yc;
            if (qmin < 0.f && -qmin < 3.0e38f)
                cr = CullRec{qa, qb, qc, Xc + o.mx, Yc + o.my, -qmin, 1.0f / det, 1.0f / qa};
        }
        cull[g] = cr;
        int32_t hit = 0;
        for (uint32_t ty = y0; ty < y1; ++ty) {
            int first, last;
            cull_row_span(cr, ty, x0, x1, first, last);
            hit += last >= first ? last - first + 1 : 0;
        }
        cnt = hit;
    }
    counts[g] = cnt;
    rects[g] = TileRect{(unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1};
    depth_keys[g] = cnt > 0 ? __float_as_uint(o.depth) : 0xFFFFFFFFu;
    if (cnt <= 0)
        return;

    // view-dependent colour: clamp_min(SH(dir) + 0.5, 0)   (rasterizer.cpp:250-266)
    const f3 dir = mk3(mean.x - cam.org[0], mean.y - cam.org[1], mean.z - cam.org[2]);
    f3 col = sh_to_color(cfg.degree, dir, [&](int k) { return mk3(P(pl.sh(k, 0)), P(pl.sh(k, 1)), P(pl.sh(k, 2))); });
    col = mk3(fmaxf(col.x + 0.5f, 0.f), fmaxf(col.y + 0.5f, 0.f), fmaxf(col.z + 0.5f, 0.f));

    float4* og = reinterpret_cast<float4*>(gauss + g);
    og[0] = make_float4(vx.x, vx.y, vx.z, vy.x);
    og[1] = make_float4(vy.y, vy.z, w2.x, w2.y);
    og[2] = make_float4(w2.z, gro.x, gro.y, gro.z);
    og[3] = make_float4(op, col.x, col.y, col.z);
    // activated copies consumed by the blend-backward epilogue (quat / scale chain rule)
    act_means[3 * (size_t)g] = mean.x, act_means[3 * (size_t)g + 1] = mean.y, act_means[3 * (size_t)g + 2] = mean.z;
    reinterpret_cast<float4*>(act_quats)[g] = make_float4(qw, qx, qy, qz);
    act_scales[3 * (size_t)g] = scale.x, act_scales[3 * (size_t)g + 1] = scale.y, act_scales[3 * (size_t)g + 2] = scale.z;
}

// ------------------------------------------------------------------------------------------------------
// per-Gaussian backward: colour clamp mask -> SH VJP -> activation VJPs -> += raw-parameter gradients;
// clears the per-view activated-space accumulators for the next view.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_preprocess_bwd(const float* __restrict__ arena, float* __restrict__ grads, const Planes pl, const uint32_t N,
                     const ViewCam cam, const int degree, const int32_t* __restrict__ counts,
                     float* __restrict__ v_means, float* __restrict__ v_quats, float* __restrict__ v_scales,
                     float* __restrict__ v_colors, float* __restrict__ v_opac) {
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    if (g >= N || counts[g] <= 0)
        return;
    const size_t Np = pl.Np;
    auto P = [&](uint32_t plane) { return __ldg(arena + (size_t)plane * Np + g); };

    f3 vm = mk3(v_means[3 * (size_t)g], v_means[3 * (size_t)g + 1], v_means[3 * (size_t)g + 2]);
    const float4 vq = reinterpret_cast<float4*>(v_quats)[g];
    const f3 vs = mk3(v_scales[3 * (size_t)g], v_scales[3 * (size_t)g + 1], v_scales[3 * (size_t)g + 2]);
    f3 vc = mk3(v_colors[3 * (size_t)g], v_colors[3 * (size_t)g + 1], v_colors[3 * (size_t)g + 2]);
    const float vo = v_opac[g];
    // ---- all loads first (independent, coalesced plane reads): parameters, SH coefficients, old gradients.
    // The read-modify-write of the gradient planes must not be written as `g[i] += v` one after the other:
    // the compiler then serialises 59 dependent DRAM round trips per thread (measured 1.03 ms per view).
    const f3 mean = mk3(P(pl.mean(0)), P(pl.mean(1)), P(pl.mean(2)));
    const float sraw[3] = {P(pl.scaling(0)), P(pl.scaling(1)), P(pl.scaling(2))};
    const float qraw[4] = {P(pl.rotation(0)), P(pl.rotation(1)), P(pl.rotation(2)), P(pl.rotation(3))};
    const float oraw = P(pl.opacity());
    const int nb = (degree + 1) * (degree + 1);
    float cf[25 * 3], go[25 * 3];
#pragma unroll
    for (int k = 0; k < 25; ++k) {
        if (k < nb) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                cf[3 * k + ch] = P(pl.sh(k, ch));
                go[3 * k + ch] = grads[(size_t)pl.sh(k, ch) * Np + g];
            }
        }
    }
    float gm[3], gs[3], gq[4], gop;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        gm[c] = grads[(size_t)pl.mean(c) * Np + g];
        gs[c] = grads[(size_t)pl.scaling(c) * Np + g];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        gq[c] = grads[(size_t)pl.rotation(c) * Np + g];
    gop = grads[(size_t)pl.opacity() * Np + g];

    // ---- SH: colour (for the clamp mask), coefficient gradients, direction gradient
    const f3 dir = mk3(mean.x - cam.org[0], mean.y - cam.org[1], mean.z - cam.org[2]);
    float x = 0.f, y = 0.f, z = 0.f, inorm = 0.f;
    if (degree >= 1) {
        inorm = rsqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        x = dir.x * inorm, y = dir.y * inorm, z = dir.z * inorm;
    }
    ShBasis B;
    sh_bases(degree, x, y, z, B);
    f3 col = mk3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 25; ++k)
        if (k < nb) {
            col.x += B.b[k] * cf[3 * k], col.y += B.b[k] * cf[3 * k + 1], col.z += B.b[k] * cf[3 * k + 2];
        }
    // clamp_min(c + 0.5, 0) passes the gradient where c + 0.5 >= 0
    if (col.x + 0.5f < 0.f)
        vc.x = 0.f;
    if (col.y + 0.5f < 0.f)
        vc.y = 0.f;
    if (col.z + 0.5f < 0.f)
        vc.z = 0.f;
    float sdot[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) {
        sdot[k] = 0.f;
        if (k < nb) {
            go[3 * k] = fmaf(B.b[k], vc.x, go[3 * k]);
            go[3 * k + 1] = fmaf(B.b[k], vc.y, go[3 * k + 1]);
            go[3 * k + 2] = fmaf(B.b[k], vc.z, go[3 * k + 2]);
            if (k >= 1)
                sdot[k] = cf[3 * k] * vc.x + cf[3 * k + 1] * vc.y + cf[3 * k + 2] * vc.z;
        }
    }
    if (degree >= 1) {
        const f3 vdn = sh_bases_vjp(degree, x, y, z, sdot);
        const float dt = vdn.x * x + vdn.y * y + vdn.z * z;
        vm = vm + mk3((vdn.x - dt * x) * inorm, (vdn.y - dt * y) * inorm, (vdn.z - dt * z) * inorm);
    }
    // ---- activation VJPs
    const float op = sigmoidf_(oraw);
    const float nrm = fmaxf(sqrtf(qraw[0] * qraw[0] + qraw[1] * qraw[1] + qraw[2] * qraw[2] + qraw[3] * qraw[3]), 1e-12f);
    const float inv = 1.0f / nrm;
    const float nw = qraw[0] * inv, nx = qraw[1] * inv, ny = qraw[2] * inv, nz = qraw[3] * inv;
    const float dq = vq.x * nw + vq.y * nx + vq.z * ny + vq.w * nz;
    // ---- all stores
#pragma unroll
    for (int k = 0; k < 25; ++k)
        if (k < nb) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                grads[(size_t)pl.sh(k, ch) * Np + g] = go[3 * k + ch];
        }
    grads[(size_t)pl.mean(0) * Np + g] = gm[0] + vm.x;
    grads[(size_t)pl.mean(1) * Np + g] = gm[1] + vm.y;
    grads[(size_t)pl.mean(2) * Np + g] = gm[2] + vm.z;
    grads[(size_t)pl.scaling(0) * Np + g] = fmaf(vs.x, __expf(sraw[0]), gs[0]); // scale = exp(raw)
    grads[(size_t)pl.scaling(1) * Np + g] = fmaf(vs.y, __expf(sraw[1]), gs[1]);
    grads[(size_t)pl.scaling(2) * Np + g] = fmaf(vs.z, __expf(sraw[2]), gs[2]);
    grads[(size_t)pl.opacity() * Np + g] = fmaf(vo, op * (1.0f - op), gop); // opacity = sigmoid(raw)
    grads[(size_t)pl.rotation(0) * Np + g] = gq[0] + (vq.x - dq * nw) * inv; // q_n = q / max(|q|, eps)
    grads[(size_t)pl.rotation(1) * Np + g] = gq[1] + (vq.y - dq * nx) * inv;
    grads[(size_t)pl.rotation(2) * Np + g] = gq[2] + (vq.z - dq * ny) * inv;
    grads[(size_t)pl.rotation(3) * Np + g] = gq[3] + (vq.w - dq * nz) * inv;
    // clear the per-view accumulators LAST: a store issued right behind a load of the same address holds back every
    // later load of the thread for one DRAM round trip (tools/micro/planar_rmw_variants.cu: 0.20 ms vs 0.13 ms)
    v_means[3 * (size_t)g] = v_means[3 * (size_t)g + 1] = v_means[3 * (size_t)g + 2] = 0.f;
    reinterpret_cast<float4*>(v_quats)[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    v_scales[3 * (size_t)g] = v_scales[3 * (size_t)g + 1] = v_scales[3 * (size_t)g + 2] = 0.f;
    v_colors[3 * (size_t)g] = v_colors[3 * (size_t)g + 1] = v_colors[3 * (size_t)g + 2] = 0.f;
    v_opac[g] = 0.f;
}

// Two launches instead of one 255-register kernel (ncu r01c: 12 % occupancy, 2.1 TB/s):
//  k_preprocess_bwd_sh  -- one thread per (Gaussian, colour channel): 16 coefficient planes read-modify-written per
//                          thread, the channel's share of dL/d(dir) goes to the mean-gradient planes with 3 RED atomics;
//  k_preprocess_bwd_geo -- one thread per Gaussian: activation VJPs of mean / scale / quat / opacity, and clears the
//                          per-view activated-space accumulators for the next view.
template <int DEG>
__global__ void __launch_bounds__(256)
    k_preprocess_bwd_sh(const float* __restrict__ arena, float* __restrict__ grads, const Planes pl, const uint32_t N,
                        const ViewCam cam, const int32_t* __restrict__ counts, float* __restrict__ v_colors) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    const int ch = blockIdx.y;
    if (g >= N || counts[g] <= 0)
        return;
    const size_t Np = pl.Np;
    float vc = v_colors[3 * (size_t)g + ch];
    float cf[NB], go[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        cf[k] = __ldg(arena + (size_t)pl.sh(k, ch) * Np + g);
        go[k] = grads[(size_t)pl.sh(k, ch) * Np + g];
    }
    float x = 0.f, y = 0.f, z = 0.f, inorm = 0.f;
    if (DEG >= 1) {
        const f3 dir = mk3(__ldg(arena + (size_t)pl.mean(0) * Np + g) - cam.org[0],
                           __ldg(arena + (size_t)pl.mean(1) * Np + g) - cam.org[1],
                           __ldg(arena + (size_t)pl.mean(2) * Np + g) - cam.org[2]);
        inorm = rsqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        x = dir.x * inorm, y = dir.y * inorm, z = dir.z * inorm;
    }
    ShBasis B;
    sh_bases(DEG, x, y, z, B);
    float col = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k)
        col += B.b[k] * cf[k];
    if (col + 0.5f < 0.f) // clamp_min(c + 0.5, 0) passes the gradient where c + 0.5 >= 0
        vc = 0.f;
    float sdot[25];
#pragma unroll
    for (int k = 0; k < 25; ++k)
        sdot[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        grads[(size_t)pl.sh(k, ch) * Np + g] = fmaf(B.b[k], vc, go[k]);
        if (k >= 1)
            sdot[k] = cf[k] * vc;
    }
    if (DEG >= 1 && vc != 0.f) {
        const f3 vdn = sh_bases_vjp(DEG, x, y, z, sdot);
        const float dt = vdn.x * x + vdn.y * y + vdn.z * z;
        atomicAdd(grads + (size_t)pl.mean(0) * Np + g, (vdn.x - dt * x) * inorm);
        atomicAdd(grads + (size_t)pl.mean(1) * Np + g, (vdn.y - dt * y) * inorm);
        atomicAdd(grads + (size_t)pl.mean(2) * Np + g, (vdn.z - dt * z) * inorm);
    }
    v_colors[3 * (size_t)g + ch] = 0.f; // cleared last, see k_preprocess_bwd
}

__global__ void __launch_bounds__(256)
    k_preprocess_bwd_geo(const float* __restrict__ arena, float* __restrict__ grads, const Planes pl, const uint32_t N,
                         const int32_t* __restrict__ counts, float* __restrict__ v_means, float* __restrict__ v_quats,
                         float* __restrict__ v_scales, float* __restrict__ v_opac) {
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    if (g >= N || counts[g] <= 0)
        return;
    const size_t Np = pl.Np;
    auto P = [&](uint32_t plane) { return __ldg(arena + (size_t)plane * Np + g); };
    const f3 vm = mk3(v_means[3 * (size_t)g], v_means[3 * (size_t)g + 1], v_means[3 * (size_t)g + 2]);
    const float4 vq = reinterpret_cast<float4*>(v_quats)[g];
    const f3 vs = mk3(v_scales[3 * (size_t)g], v_scales[3 * (size_t)g + 1], v_scales[3 * (size_t)g + 2]);
    const float vo = v_opac[g];
    // loads first, stores last (independent DRAM round trips overlap)
    const float sraw[3] = {P(pl.scaling(0)), P(pl.scaling(1)), P(pl.scaling(2))};
    const float qraw[4] = {P(pl.rotation(0)), P(pl.rotation(1)), P(pl.rotation(2)), P(pl.rotation(3))};
    const float oraw = P(pl.opacity());
    float gs[3], gq[4], gop;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        gs[c] = grads[(size_t)pl.scaling(c) * Np + g];
#pragma unroll
    for (int c = 0; c < 4; ++c)
        gq[c] = grads[(size_t)pl.rotation(c) * Np + g];
    gop = grads[(size_t)pl.opacity() * Np + g];

    const float op = sigmoidf_(oraw);
    const float nrm = fmaxf(sqrtf(qraw[0] * qraw[0] + qraw[1] * qraw[1] + qraw[2] * qraw[2] + qraw[3] * qraw[3]), 1e-12f);
    const float inv = 1.0f / nrm;
    const float nw = qraw[0] * inv, nx = qraw[1] * inv, ny = qraw[2] * inv, nz = qraw[3] * inv;
    const float dq = vq.x * nw + vq.y * nx + vq.z * ny + vq.w * nz;
    // the SH kernel adds its dL/d(dir) share to the same planes with atomics -> RED here as well (no ordering needed)
    atomicAdd(grads + (size_t)pl.mean(0) * Np + g, vm.x);
    atomicAdd(grads + (size_t)pl.mean(1) * Np + g, vm.y);
    atomicAdd(grads + (size_t)pl.mean(2) * Np + g, vm.z);
    grads[(size_t)pl.scaling(0) * Np + g] = fmaf(vs.x, __expf(sraw[0]), gs[0]); // scale = exp(raw)
    grads[(size_t)pl.scaling(1) * Np + g] = fmaf(vs.y, __expf(sraw[1]), gs[1]);
    grads[(size_t)pl.scaling(2) * Np + g] = fmaf(vs.z, __expf(sraw[2]), gs[2]);
    grads[(size_t)pl.opacity() * Np + g] = fmaf(vo, op * (1.0f - op), gop); // opacity = sigmoid(raw)
    grads[(size_t)pl.rotation(0) * Np + g] = gq[0] + (vq.x - dq * nw) * inv; // q_n = q / max(|q|, eps)
    grads[(size_t)pl.rotation(1) * Np + g] = gq[1] + (vq.y - dq * nx) * inv;
    grads[(size_t)pl.rotation(2) * Np + g] = gq[2] + (vq.z - dq * ny) * inv;
    grads[(size_t)pl.rotation(3) * Np + g] = gq[3] + (vq.w - dq * nz) * inv;
    v_means[3 * (size_t)g] = v_means[3 * (size_t)g + 1] = v_means[3 * (size_t)g + 2] = 0.f;
    reinterpret_cast<float4*>(v_quats)[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    v_scales[3 * (size_t)g] = v_scales[3 * (size_t)g + 1] = v_scales[3 * (size_t)g + 2] = 0.f;
    v_opac[g] = 0.f;
}

// ------------------------------------------------------------------------------------------------------
// L1 photometric loss on (rgb + T * bg) and its gradient packed for the backward
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_loss_l1(const float4* __restrict__ pix_state, const void* __restrict__ target, const int fmt,
              const uint32_t width, const uint32_t height, const float bg_r, const float bg_g, const float bg_b,
              const float scale, float4* __restrict__ v_pix, float* __restrict__ partials) {
    __shared__ float s_sum[8];
    const uint32_t npix = width * height;
    float acc = 0.f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < npix; i += gridDim.x * 256) {
        const float4 s = pix_state[i];
        float tr, tg, tb;
        if (fmt == LFS_IMG_U8_HWC) {
            const uint8_t* t = static_cast<const uint8_t*>(target) + 3 * (size_t)i;
            tr = t[0] * (1.0f / 255.0f), tg = t[1] * (1.0f / 255.0f), tb = t[2] * (1.0f / 255.0f);
        } else if (fmt == LFS_IMG_F32_HWC) {
            const float* t = static_cast<const float*>(target) + 3 * (size_t)i;
            tr = t[0], tg = t[1], tb = t[2];
        } else { // F32 CHW
            const float* t = static_cast<const float*>(target);
            tr = t[i], tg = t[npix + i], tb = t[2 * (size_t)npix + i];
        }
        // the reference clamps the render to [0,1] before the loss (rasterizer.cpp:401); clamp backward passes inside
        const float rr = fmaf(s.w, bg_r, s.x), rg = fmaf(s.w, bg_g, s.y), rb = fmaf(s.w, bg_b, s.z);
        const float dr = fminf(fmaxf(rr, 0.f), 1.f) - tr, dg = fminf(fmaxf(rg, 0.f), 1.f) - tg,
                    db = fminf(fmaxf(rb, 0.f), 1.f) - tb;
        acc += fabsf(dr) + fabsf(dg) + fabsf(db);
        const float vr = (rr >= 0.f && rr <= 1.f) ? (dr > 0.f ? scale : (dr < 0.f ? -scale : 0.f)) : 0.f;
        const float vg = (rg >= 0.f && rg <= 1.f) ? (dg > 0.f ? scale : (dg < 0.f ? -scale : 0.f)) : 0.f;
        const float vb = (rb >= 0.f && rb <= 1.f) ? (db > 0.f ? scale : (db < 0.f ? -scale : 0.f)) : 0.f;
        v_pix[i] = make_float4(vr, vg, vb, -s.w * (bg_r * vr + bg_g * vg + bg_b * vb));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0)
        s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w)
            t += s_sum[w];
        partials[blockIdx.x] = t * scale;
    }
}
__global__ void __launch_bounds__(256)
    k_loss_finish(const float* __restrict__ partials, const int n, float* __restrict__ loss_accum) {
    __shared__ float s_sum[8];
    float t = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) // fixed order -> deterministic loss value
        t += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0)
        s_sum[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w)
            a += s_sum[w];
        *loss_accum += a;
    }
}

// ------------------------------------------------------------------------------------------------------
// L1 + fused SSIM ("valid" crop) photometric loss, the reference's training loss
// (Trainer::compute_photometric_loss, src/training/trainer.cpp:103-131; fused_ssim, include/kernels/fused_ssim.cuh:27-122;
// kernels src/training/kernels/ssim.cu:64-460).  Same 11-tap separable Gaussian, zero padding and per-pixel formulas;
// re-organised for the fused step: the rendered image is read from the blend state (rgb + T*bg, clamped to [0,1] like
// rasterizer.cpp:401) and the result is written straight into the packed upstream-gradient record of the blend
// backward, so the [3,H,W] image, the SSIM map and dL/dimage never exist as tensors.
// ------------------------------------------------------------------------------------------------------
__constant__ float c_ssim_g[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                   0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                                   0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                                   0.0075987582094967365f, 0.001028380123898387f};
constexpr int kSsimHalo = 5, kSsimS = kTile + 2 * kSsimHalo; // 26

__device__ __forceinline__ float target_at(const void* __restrict__ target, const int fmt, const uint32_t npix,
                                           const uint32_t i, const int c) {
    if (fmt == LFS_IMG_U8_HWC)
        return static_cast<const uint8_t*>(target)[3 * (size_t)i + c] * (1.0f / 255.0f);
    if (fmt == LFS_IMG_F32_HWC)
        return static_cast<const float*>(target)[3 * (size_t)i + c];
    return static_cast<const float*>(target)[(size_t)c * npix + i];
}
__device__ __forceinline__ float chan(const float4 s, const int c) { return c == 0 ? s.x : (c == 1 ? s.y : s.z); }

__global__ void __launch_bounds__(256)
    k_ssim_fwd(const float4* __restrict__ pix_state, const void* __restrict__ target, const int fmt, const int W,
               const int H, const float bg_r, const float bg_g, const float bg_b, const float dmap /* -w*lambda/count */,
               float* __restrict__ maps, float* __restrict__ partials) {
    // One global-load phase for all three channels, then the separable convolutions run out of shared memory.
    // Layouts are chosen for wide shared loads: (X, Y) pairs -> one LDS.64 per tap, the four horizontal sums
    // (mu1, mu2, E[X^2 + Y^2], E[XY]) -> one LDS.128 per tap.  sigma1^2 + sigma2^2 only ever appears as a sum in the
    // SSIM formula, so X^2 and Y^2 share one convolution (4 instead of the reference's 5, ssim.cu:118-247).
    __shared__ float2 sXY[3][kSsimS][kSsimS];
    __shared__ float4 xc[kSsimS][kTile];
    __shared__ float s_red[2][8];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const int px = x0 + tx, py = y0 + ty;
    const uint32_t npix = (uint32_t)W * (uint32_t)H;
    const bool crop_all = !(H > 10 && W > 10); // fused_ssim.cuh:63-70
    for (int t = threadIdx.x; t < kSsimS * kSsimS; t += 256) {
        const int ly = t / kSsimS, lx = t - ly * kSsimS;
        const int gy = y0 + ly - kSsimHalo, gx = x0 + lx - kSsimHalo;
        float X0 = 0.f, X1 = 0.f, X2 = 0.f, Y0 = 0.f, Y1 = 0.f, Y2 = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const uint32_t i = (uint32_t)gy * W + gx;
            const float4 s = pix_state[i];
            X0 = fminf(fmaxf(fmaf(s.w, bg_r, s.x), 0.f), 1.f);
            X1 = fminf(fmaxf(fmaf(s.w, bg_g, s.y), 0.f), 1.f);
            X2 = fminf(fmaxf(fmaf(s.w, bg_b, s.z), 0.f), 1.f);
            Y0 = target_at(target, fmt, npix, i, 0), Y1 = target_at(target, fmt, npix, i, 1);
            Y2 = target_at(target, fmt, npix, i, 2);
        }
        sXY[0][ly][lx] = make_float2(X0, Y0), sXY[1][ly][lx] = make_float2(X1, Y1), sXY[2][ly][lx] = make_float2(X2, Y2);
    }
    __syncthreads();
    float l1 = 0.f, ss = 0.f;
    for (int c = 0; c < 3; ++c) {
        for (int t = threadIdx.x; t < kSsimS * kTile; t += 256) { // horizontal 11x1
            const int ly = t >> 4, lx = (t & 15) + kSsimHalo;
            float2 a01 = make_float2(0.f, 0.f), a23 = a01; // FFMA2 pairs: (E[X], E[Y]) and (E[X^2 + Y^2], E[XY])
#pragma unroll
            for (int d = -kSsimHalo; d <= kSsimHalo; ++d) {
                const float2 w2 = make_float2(c_ssim_g[d + kSsimHalo], c_ssim_g[d + kSsimHalo]);
                const float2 v = sXY[c][ly][lx + d];
                a01 = ffma2(w2, v, a01);
                a23 = ffma2(w2, make_float2(fmaf(v.x, v.x, v.y * v.y), v.x * v.y), a23);
            }
            xc[ly][t & 15] = make_float4(a01.x, a01.y, a23.x, a23.y);
        }
        __syncthreads();
        {
            float2 o01 = make_float2(0.f, 0.f), o23 = o01; // vertical 1x11
#pragma unroll
            for (int d = 0; d < 2 * kSsimHalo + 1; ++d) {
                const float2 w2 = make_float2(c_ssim_g[d], c_ssim_g[d]);
                const float4 r = xc[ty + d][tx];
                o01 = ffma2(w2, make_float2(r.x, r.y), o01);
                o23 = ffma2(w2, make_float2(r.z, r.w), o23);
            }
            const float o0 = o01.x, o1 = o01.y, o2 = o23.x, o3 = o23.y;
            if (px < W && py < H) {
                const float mu1 = o0, mu2 = o1, s12 = o3 - mu1 * mu2;
                const float C1 = 0.0001f, C2 = 0.0009f;
                const float A = mu1 * mu1 + mu2 * mu2 + C1;
                const float B = o2 - mu1 * mu1 - mu2 * mu2 + C2; // sigma1^2 + sigma2^2 + C2
                const float Cc = 2.f * mu1 * mu2 + C1, D = 2.f * s12 + C2;
                const float iAB = 1.0f / (A * B);
                const float val = Cc * D * iAB; // ssim.cu:262
                const bool in_crop = crop_all || (px >= 5 && px < W - 5 && py >= 5 && py < H - 5);
                const float dm = in_crop ? dmap : 0.f;
                const float d_mu1 = (mu2 * 2.f * D) * iAB - (mu2 * 2.f * Cc) * iAB - (mu1 * 2.f * Cc * D) * iAB / A +
                                    (mu1 * 2.f * Cc * D) * iAB / B; // :269
                const uint32_t i = (uint32_t)py * W + px;
                // (dL/dmap) * (dm/dmu1, dm/dsigma1^2, dm/dsigma12) interleaved per channel: one 16-B record per pixel
                reinterpret_cast<float4*>(maps)[(size_t)c * npix + i] =
                    make_float4(dm * d_mu1, dm * (-Cc * D) * iAB / B, dm * (2.f * Cc) * iAB, 0.f); // :269-271
                if (in_crop)
                    ss += val;
                const float2 ctr = sXY[c][ty + kSsimHalo][tx + kSsimHalo];
                l1 += fabsf(ctr.x - ctr.y);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0)
        s_red[0][threadIdx.x >> 5] = l1, s_red[1][threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w)
            a += s_red[0][w], b += s_red[1][w];
        const uint32_t blk = blockIdx.y * gridDim.x + blockIdx.x;
        partials[2 * blk] = a, partials[2 * blk + 1] = b;
    }
}

__global__ void __launch_bounds__(256)
    k_ssim_bwd(const float4* __restrict__ pix_state, const void* __restrict__ target, const int fmt, const int W,
               const int H, const float bg_r, const float bg_g, const float bg_b, const float l1w /* w*(1-lambda)/(3HW) */,
               const float* __restrict__ maps, float4* __restrict__ v_pix) {
    __shared__ float4 sD[3][kSsimS][kSsimS]; // per channel: the three derivative maps of a pixel in one 16-B record
    __shared__ float4 xc[kSsimS][kTile];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const int px = x0 + tx, py = y0 + ty;
    const uint32_t npix = (uint32_t)W * (uint32_t)H;
    const bool inside = px < W && py < H;
    const uint32_t pi = inside ? (uint32_t)py * W + px : 0u;
    const float4 st = inside ? pix_state[pi] : make_float4(0.f, 0.f, 0.f, 0.f);
    float Yt[3] = {0.f, 0.f, 0.f};
    if (inside) {
        Yt[0] = target_at(target, fmt, npix, pi, 0), Yt[1] = target_at(target, fmt, npix, pi, 1);
        Yt[2] = target_at(target, fmt, npix, pi, 2);
    }
    const float4* m4 = reinterpret_cast<const float4*>(maps);
    for (int t = threadIdx.x; t < kSsimS * kSsimS; t += 256) {
        const int ly = t / kSsimS, lx = t - ly * kSsimS;
        const int gy = y0 + ly - kSsimHalo, gx = x0 + lx - kSsimHalo;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const uint32_t i = in ? (uint32_t)gy * W + gx : 0u;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        sD[0][ly][lx] = in ? __ldg(m4 + i) : z;
        sD[1][ly][lx] = in ? __ldg(m4 + npix + i) : z;
        sD[2][ly][lx] = in ? __ldg(m4 + 2 * (size_t)npix + i) : z;
    }
    __syncthreads();
    float g[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < 3; ++c) {
        for (int t = threadIdx.x; t < kSsimS * kTile; t += 256) {
            const int ly = t >> 4, lx = (t & 15) + kSsimHalo;
            float2 a01 = make_float2(0.f, 0.f);
            float a2 = 0.f;
#pragma unroll
            for (int d = -kSsimHalo; d <= kSsimHalo; ++d) {
                const float w = c_ssim_g[d + kSsimHalo];
                const float4 v = sD[c][ly][lx + d];
                a01 = ffma2(make_float2(w, w), make_float2(v.x, v.y), a01);
                a2 = fmaf(w, v.z, a2);
            }
            xc[ly][t & 15] = make_float4(a01.x, a01.y, a2, 0.f);
        }
        __syncthreads();
        if (inside) {
            float2 s01 = make_float2(0.f, 0.f);
            float s2 = 0.f;
#pragma unroll
            for (int d = 0; d < 2 * kSsimHalo + 1; ++d) {
                const float w = c_ssim_g[d];
                const float4 r = xc[ty + d][tx];
                s01 = ffma2(make_float2(w, w), make_float2(r.x, r.y), s01);
                s2 = fmaf(w, r.z, s2);
            }
            const float s0 = s01.x, s1 = s01.y;
            const float bgc = c == 0 ? bg_r : (c == 1 ? bg_g : bg_b);
            const float raw = fmaf(st.w, bgc, chan(st, c));
            const float X = fminf(fmaxf(raw, 0.f), 1.f), Y = Yt[c];
            const float diff = X - Y;
            float gc = s0 + 2.f * X * s1 + Y * s2 + (diff > 0.f ? l1w : (diff < 0.f ? -l1w : 0.f)); // ssim.cu:417 + L1
            if (!(raw >= 0.f && raw <= 1.f)) // torch::clamp backward (rasterizer.cpp:401)
                gc = 0.f;
            g[c] = gc;
        }
        __syncthreads();
    }
    if (inside)
        v_pix[pi] = make_float4(g[0], g[1], g[2], -st.w * (bg_r * g[0] + bg_g * g[1] + bg_b * g[2]));
}

__global__ void __launch_bounds__(256)
    k_ssim_finish(const float* __restrict__ partials, const int n, const float w_l1, const float w_ssim,
                  const float ssim_const, float* __restrict__ loss_accum) {
    __shared__ float s_red[2][8];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) // fixed order -> deterministic loss value
        a += partials[2 * i], b += partials[2 * i + 1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0)
        s_red[0][threadIdx.x >> 5] = a, s_red[1][threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        float x = 0.f, y = 0.f;
        for (int w = 0; w < 8; ++w)
            x += s_red[0][w], y += s_red[1][w];
        *loss_accum += w_l1 * x + ssim_const - w_ssim * y;
    }
}

// external upstream gradients (parity tests / custom losses): v_image [H,W,3], v_alpha [H,W] or null
__global__ void __launch_bounds__(256)
    k_pack_vpix_ext(const float* __restrict__ v_image, const float* __restrict__ v_alpha,
                    const float4* __restrict__ pix_state, const float bg_r, const float bg_g, const float bg_b,
                    const uint32_t npix, float4* __restrict__ v_pix) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npix)
        return;
    const float r = v_image[3 * (size_t)i], g = v_image[3 * (size_t)i + 1], b = v_image[3 * (size_t)i + 2];
    const float va = (v_alpha ? v_alpha[i] : 0.f) - (bg_r * r + bg_g * g + bg_b * b);
    v_pix[i] = make_float4(r, g, b, pix_state[i].w * va);
}

__global__ void __launch_bounds__(256)
    k_export_image(const float4* __restrict__ pix_state, const float bg_r, const float bg_g, const float bg_b,
                   const uint32_t npix, float* __restrict__ image, float* __restrict__ alpha) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npix)
        return;
    const float4 s = pix_state[i];
    if (image) {
        image[3 * (size_t)i] = fmaf(s.w, bg_r, s.x);
        image[3 * (size_t)i + 1] = fmaf(s.w, bg_g, s.y);
        image[3 * (size_t)i + 2] = fmaf(s.w, bg_b, s.z);
    }
    if (alpha)
        alpha[i] = 1.0f - s.w;
}

// n_inst[2] = max over all forwards of the instance count the view WANTED (n_inst[0], may exceed the capacity)
__global__ void k_high_water(uint32_t* __restrict__ n_inst) {
    if (n_inst[0] > n_inst[2])
        n_inst[2] = n_inst[0];
}

} // namespace lfs

// ==========================================================================================================
using namespace lfs;

extern "C" uint64_t lfs_trainer_arena_floats(const lfs_trainer_desc* d) {
    if (!d)
        return 0;
    const uint32_t K = (d->sh_degree_max + 1) * (d->sh_degree_max + 1);
    const uint64_t Np = ((uint64_t)d->n_gaussians + 3) / 4 * 4;
    return (11ull + 3ull * K) * Np;
}

extern "C" void* lfs_trainer_create(const lfs_trainer_desc* d) {
    if (!d || d->n_gaussians == 0 || d->width == 0 || d->height == 0 || d->sh_degree_max > 4) {
        set_error("trainer_create: bad descriptor");
        return nullptr;
    }
    Trainer* t = new (std::nothrow) Trainer();
    if (!t)
        return nullptr;
    t->d = *d;
    t->pl.Np = (d->n_gaussians + 3) / 4 * 4;
    t->pl.K = (d->sh_degree_max + 1) * (d->sh_degree_max + 1);
    t->tile_w = (d->width + kTile - 1) / kTile;
    t->tile_h = (d->height + kTile - 1) / kTile;
    t->n_tiles = t->tile_w * t->tile_h;
    if (t->tile_w >= 65536 || t->tile_h >= 65536) {
        set_error("trainer_create: image too large");
        delete t;
        return nullptr;
    }
    uint64_t cap = d->instance_capacity ? d->instance_capacity : (uint64_t)d->n_gaussians * 8 + 65536;
    if (cap >= (1ull << 31))
        cap = (1ull << 31) - 1;
    t->inst_cap = (uint32_t)cap;
    t->bucket_cap = t->inst_cap / kBucket + t->n_tiles + 1;
    t->blob_bytes = trainer_carve(*t, nullptr);
    if (cudaMalloc(&t->blob, t->blob_bytes) != cudaSuccess) {
        set_error("trainer_create: cudaMalloc(%zu bytes) failed", t->blob_bytes);
        delete t;
        return nullptr;
    }
    trainer_carve(*t, t->blob);
    cudaMemset(t->v_means, 0, sizeof(float) * 3 * (size_t)d->n_gaussians);
    cudaMemset(t->v_quats, 0, sizeof(float) * 4 * (size_t)d->n_gaussians);
    cudaMemset(t->v_scales, 0, sizeof(float) * 3 * (size_t)d->n_gaussians);
    cudaMemset(t->v_colors, 0, sizeof(float) * 3 * (size_t)d->n_gaussians);
    cudaMemset(t->v_opac, 0, sizeof(float) * (size_t)d->n_gaussians);
    cudaMemset(t->n_inst, 0, sizeof(uint32_t) * 4);
    if (cudaMallocHost(&t->stats_host, sizeof(uint32_t) * 4) != cudaSuccess)
        t->stats_host = nullptr;
    else
        t->stats_host[0] = t->stats_host[1] = t->stats_host[2] = t->stats_host[3] = 0;
    return t;
}

extern "C" void lfs_trainer_destroy(void* h) {
    Trainer* t = static_cast<Trainer*>(h);
    if (!t)
        return;
    if (t->blob)
        cudaFree(t->blob);
    if (t->stats_host)
        cudaFreeHost(t->stats_host);
    delete t;
}

extern "C" uint64_t lfs_trainer_scratch_bytes(void* h) { return h ? static_cast<Trainer*>(h)->blob_bytes : 0; }
extern "C" uint64_t lfs_trainer_instance_capacity(void* h) { return h ? static_cast<Trainer*>(h)->inst_cap : 0; }

static int pack_common(void* h, float* means, float* sh0, float* shN, float* scaling, float* rotation, float* opacity,
                       float* arena, int to_arena, void* stream) {
    Trainer* t = static_cast<Trainer*>(h);
    LFS_CHECK_ARG(t && means && sh0 && scaling && rotation && opacity && arena, "trainer_pack: null pointer");
    LFS_CHECK_ARG(t->pl.K == 1 || shN, "trainer_pack: shN is null");
    k_pack_arena<<<div_up(t->d.n_gaussians, 256), 256, 0, (cudaStream_t)stream>>>(
        means, sh0, shN, scaling, rotation, opacity, t->d.n_gaussians, t->pl, arena, to_arena);
    LFS_LAUNCH_OK("k_pack_arena");
    return LFS_OK;
}
extern "C" int lfs_trainer_pack(void* h, const float* means, const float* sh0, const float* shN, const float* scaling,
                                const float* rotation, const float* opacity, float* arena, void* stream) {
    return pack_common(h, const_cast<float*>(means), const_cast<float*>(sh0), const_cast<float*>(shN),
                       const_cast<float*>(scaling), const_cast<float*>(rotation), const_cast<float*>(opacity), arena, 1,
                       stream);
}
extern "C" int lfs_trainer_unpack(void* h, const float* arena, float* means, float* sh0, float* shN, float* scaling,
                                  float* rotation, float* opacity, void* stream) {
    return pack_common(h, means, sh0, shN, scaling, rotation, opacity, const_cast<float*>(arena), 0, stream);
}

extern "C" int lfs_trainer_view_forward(void* h, const float* params_arena, const float* viewmat_host,
                                        const float* K_host, uint32_t active_sh_degree, const float* bg_host,
                                        float* image_out, float* alpha_out, void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(t && params_arena && viewmat_host && K_host, "trainer_view_forward: null pointer");
    LFS_CHECK_ARG(active_sh_degree <= t->d.sh_degree_max, "trainer_view_forward: active degree %u > max %u",
                  active_sh_degree, t->d.sh_degree_max);
    const uint32_t N = t->d.n_gaussians;
    t->cam_host = make_viewcam(viewmat_host, K_host, (int)t->d.width, (int)t->d.height);
    t->active_degree = active_sh_degree;
    for (int k = 0; k < 3; ++k)
        t->bg[k] = bg_host ? bg_host[k] : 0.f;
    LFS_CUDA_OK(cudaMemcpyAsync(t->cam_dev, &t->cam_host, sizeof(ViewCam), cudaMemcpyHostToDevice, stream));

    PreCfg cfg{t->d.eps2d, t->d.near_plane, t->d.far_plane, t->d.radius_clip, t->d.ut, (int)active_sh_degree};
    t->mark(0, stream);
    const bool exact_cull = raster_options().exact_cull != 0;
    k_preprocess_fwd<<<div_up(N, 256), 256, 0, stream>>>(params_arena, t->pl, N, t->cam_host, cfg, t->gauss, t->rects,
                                                         exact_cull ? t->cull : nullptr, t->counts, t->dk_a, t->pm_a, t->act_means, t->act_quats,
                                                         t->act_scales);
    LFS_LAUNCH_OK("k_preprocess_fwd");
    t->mark(1, stream);
    int in_b = 0;
    int rc = radix_sort_pairs(t->dk_a, t->pm_a, t->dk_b, t->pm_b, N, nullptr, 0, 32, t->sort_scr, &in_b, stream);
    if (rc)
        return rc;
    const uint32_t* perm = in_b ? t->pm_b : t->pm_a;
    rc = exclusive_scan_u32(reinterpret_cast<const uint32_t*>(t->counts), perm, t->off, t->n_inst, N, nullptr,
                            t->scan_scr, stream);
    if (rc)
        return rc;
    k_high_water<<<1, 1, 0, stream>>>(t->n_inst);
    LFS_LAUNCH_OK("k_high_water");
    if (exact_cull)
        rc = launch_emit_instances_cull(perm, t->off, N, t->rects, t->counts, t->cull, t->tile_w, t->inst_cap, t->n_inst,
                                        t->tk_a, t->tv_a, stream);
    else
        rc = launch_emit_instances(perm, t->off, N, t->rects, t->tile_w, 0, t->inst_cap, t->n_inst, t->tk_a, t->tv_a,
                                   stream);
    if (rc)
        return rc;
    rc = radix_sort_pairs(t->tk_a, t->tv_a, t->tk_b, t->tv_b, t->inst_cap, t->n_inst, 0, tile_key_bits(t->n_tiles),
                          t->sort_scr, &in_b, stream);
    if (rc)
        return rc;
    t->sorted_keys = in_b ? t->tk_b : t->tk_a;
    t->sorted_vals = in_b ? t->tv_b : t->tv_a;
    rc = launch_tile_offsets(t->sorted_keys, t->inst_cap, t->n_inst, t->n_tiles, t->tile_off, stream);
    if (rc)
        return rc;

    RasterBuffers rb{};
    rb.gauss = t->gauss;
    rb.tile_off = t->tile_off;
    rb.inst_gid = reinterpret_cast<const int32_t*>(t->sorted_vals);
    rb.bucket_off = t->bucket_off;
    rb.bucket_tile = t->bucket_tile;
    rb.ckpt = t->ckpt;
    rb.tile_max_contrib = t->tile_max;
    rb.live = t->live;
    rb.pix_state = t->pix_state;
    rb.n_contrib = t->n_contrib;
    t->mark(2, stream);
    rc = launch_bucket_offsets(rb, t->n_tiles, t->n_inst + 1, t->scan_scr, t->bucket_counts, stream);
    if (rc)
        return rc;
    t->mark(3, stream);
    rc = launch_blend_fwd(rb, t->cam_dev, 1, t->d.width, t->d.height, t->tile_w, t->tile_h, true, nullptr, nullptr, nullptr,
                          nullptr, nullptr, stream);
    if (rc)
        return rc;
    t->mark(4, stream);
    if (image_out || alpha_out) {
        const uint32_t npix = t->d.width * t->d.height;
        k_export_image<<<div_up(npix, 256), 256, 0, stream>>>(t->pix_state, t->bg[0], t->bg[1], t->bg[2], npix,
                                                              image_out, alpha_out);
        LFS_LAUNCH_OK("k_export_image");
    }
    if (t->stats_host)
        LFS_CUDA_OK(cudaMemcpyAsync(t->stats_host, t->n_inst, sizeof(uint32_t) * 3, cudaMemcpyDeviceToHost, stream));
    return LFS_OK;
}

extern "C" int lfs_trainer_view_loss_l1(void* h, const void* target, int target_format, float scale,
                                        float* loss_accum, void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(t && target, "trainer_view_loss_l1: null pointer");
    LFS_CHECK_ARG(target_format >= 0 && target_format <= 2, "trainer_view_loss_l1: bad target format");
    const unsigned grid = t->n_tiles < (unsigned)(num_sms() * 8) ? t->n_tiles : (unsigned)(num_sms() * 8);
    k_loss_l1<<<grid, 256, 0, stream>>>(t->pix_state, target, target_format, t->d.width, t->d.height, t->bg[0], t->bg[1],
                                        t->bg[2], scale, t->v_pix, t->loss_partials);
    LFS_LAUNCH_OK("k_loss_l1");
    if (loss_accum) {
        k_loss_finish<<<1, 256, 0, stream>>>(t->loss_partials, (int)grid, loss_accum);
        LFS_LAUNCH_OK("k_loss_finish");
    }
    return LFS_OK;
}

extern "C" int lfs_trainer_view_loss_ssim_l1(void* h, const void* target, int target_format, float lambda_dssim,
                                             float weight, float* loss_accum, void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(t && target, "trainer_view_loss_ssim_l1: null pointer");
    LFS_CHECK_ARG(target_format >= 0 && target_format <= 2, "trainer_view_loss_ssim_l1: bad target format");
    LFS_CHECK_ARG(lambda_dssim >= 0.f && lambda_dssim <= 1.f, "trainer_view_loss_ssim_l1: lambda_dssim %f not in [0,1]",
                  (double)lambda_dssim);
    const int W = (int)t->d.width, H = (int)t->d.height;
    const double count = 3.0 * ((H > 10 && W > 10) ? (double)(H - 10) * (double)(W - 10) : (double)H * (double)W);
    const double n_l1 = 3.0 * (double)H * (double)W;
    const dim3 grid(t->tile_w, t->tile_h);
    k_ssim_fwd<<<grid, 256, 0, stream>>>(t->pix_state, target, target_format, W, H, t->bg[0], t->bg[1], t->bg[2],
                                         (float)(-(double)weight * lambda_dssim / count), t->ssim_maps, t->loss_partials);
    LFS_LAUNCH_OK("k_ssim_fwd");
    k_ssim_bwd<<<grid, 256, 0, stream>>>(t->pix_state, target, target_format, W, H, t->bg[0], t->bg[1], t->bg[2],
                                         (float)((double)weight * (1.0 - lambda_dssim) / n_l1), t->ssim_maps, t->v_pix);
    LFS_LAUNCH_OK("k_ssim_bwd");
    if (loss_accum) {
        k_ssim_finish<<<1, 256, 0, stream>>>(t->loss_partials, (int)t->n_tiles,
                                             (float)((double)weight * (1.0 - lambda_dssim) / n_l1),
                                             (float)((double)weight * lambda_dssim / count),
                                             (float)((double)weight * lambda_dssim), loss_accum);
        LFS_LAUNCH_OK("k_ssim_finish");
    }
    return LFS_OK;
}

extern "C" int lfs_trainer_view_set_grad(void* h, const float* v_image, const float* v_alpha, void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(t && v_image, "trainer_view_set_grad: null pointer");
    const uint32_t npix = t->d.width * t->d.height;
    k_pack_vpix_ext<<<div_up(npix, 256), 256, 0, stream>>>(v_image, v_alpha, t->pix_state, t->bg[0], t->bg[1], t->bg[2],
                                                           npix, t->v_pix);
    LFS_LAUNCH_OK("k_pack_vpix_ext");
    return LFS_OK;
}

// The backward of a view in two halves, so that a caller running two views on two streams (each view on its own trainer
// handle, both accumulating into one gradient arena) can order just the second halves: the blend backward only touches
// per-handle scratch, the per-Gaussian kernels read-modify-write the shared gradient arena.
extern "C" int lfs_trainer_view_backward_blend(void* h, void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(t, "trainer_view_backward_blend: null handle");
    LFS_CHECK_ARG(t->sorted_vals != nullptr, "trainer_view_backward_blend: no forward has been run");
    const uint32_t N = t->d.n_gaussians;
    RasterBuffers rb{};
    rb.gauss = t->gauss;
    rb.tile_off = t->tile_off;
    rb.inst_gid = reinterpret_cast<const int32_t*>(t->sorted_vals);
    rb.bucket_off = t->bucket_off;
    rb.bucket_tile = t->bucket_tile;
    rb.ckpt = t->ckpt;
    rb.tile_max_contrib = t->tile_max;
    rb.live = t->live;
    rb.pix_state = t->pix_state;
    rb.n_contrib = t->n_contrib;
    t->mark(5, stream);
    int rc = launch_blend_bwd(rb, t->cam_dev, t->v_pix, t->act_quats, t->act_scales, t->act_means, 1, N, t->d.width,
                              t->d.height, t->tile_w, t->tile_h, t->bucket_cap, t->n_inst + 1, t->v_means, t->v_quats,
                              t->v_scales, t->v_colors, t->v_opac, stream);
    if (rc)
        return rc;
    t->mark(6, stream);
    return LFS_OK;
}

extern "C" int lfs_trainer_view_backward_params(void* h, const float* params_arena, float* grads_arena, void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    LFS_CHECK_ARG(t && params_arena && grads_arena, "trainer_view_backward_params: null pointer");
    LFS_CHECK_ARG(t->sorted_vals != nullptr, "trainer_view_backward_params: no forward has been run");
    const uint32_t N = t->d.n_gaussians;
    if (!raster_options().pre_bwd_split) {
        k_preprocess_bwd<<<div_up(N, 256), 256, 0, stream>>>(params_arena, grads_arena, t->pl, N, t->cam_host,
                                                             (int)t->active_degree, t->counts, t->v_means, t->v_quats,
                                                             t->v_scales, t->v_colors, t->v_opac);
        LFS_LAUNCH_OK("k_preprocess_bwd");
    } else {
        const dim3 grid_sh(div_up(N, 256), 3);
#define LFS_SH_BWD(D)                                                                                                  \
    k_preprocess_bwd_sh<D><<<grid_sh, 256, 0, stream>>>(params_arena, grads_arena, t->pl, N, t->cam_host, t->counts,     \
                                                        t->v_colors)
        switch ((int)t->active_degree) {
        case 0: LFS_SH_BWD(0); break;
        case 1: LFS_SH_BWD(1); break;
        case 2: LFS_SH_BWD(2); break;
        case 3: LFS_SH_BWD(3); break;
        default: LFS_SH_BWD(4); break;
        }
#undef LFS_SH_BWD
        LFS_LAUNCH_OK("k_preprocess_bwd_sh");
        k_preprocess_bwd_geo<<<div_up(N, 256), 256, 0, stream>>>(params_arena, grads_arena, t->pl, N, t->counts,
                                                                 t->v_means, t->v_quats, t->v_scales, t->v_opac);
        LFS_LAUNCH_OK("k_preprocess_bwd_geo");
    }
    t->mark(7, stream);
    if (t->profile && t->ev_ok) { // fold this view's stage times into the running sums (blocks: profiling only)
        cudaEventSynchronize(t->ev[7]);
        for (int i = 0; i < LFS_PROF_STAGES; ++i) {
            float ms = 0.f;
            // stage 4 = everything enqueued between the end of the forward and the backward: the loss kernels (and, when
            // targets arrive on a copy stream, the wait for them)
            if (cudaEventElapsedTime(&ms, t->ev[i], t->ev[i + 1]) == cudaSuccess) {
                t->acc_ms[i] += ms;
                t->acc_n[i] += 1;
            }
        }
    }
    return LFS_OK;
}

extern "C" int lfs_trainer_view_backward(void* h, const float* params_arena, float* grads_arena, void* stream) {
    LFS_CHECK_ARG(h && params_arena && grads_arena, "trainer_view_backward: null pointer");
    const int rc = lfs_trainer_view_backward_blend(h, stream);
    return rc ? rc : lfs_trainer_view_backward_params(h, params_arena, grads_arena, stream);
}

extern "C" int lfs_trainer_set_profile(void* h, int enable) {
    Trainer* t = static_cast<Trainer*>(h);
    LFS_CHECK_ARG(t, "trainer_set_profile: null handle");
    if (enable && !t->ev_ok) {
        for (auto& e : t->ev)
            LFS_CUDA_OK(cudaEventCreate(&e));
        t->ev_ok = true;
    }
    t->profile = enable ? 1 : 0;
    for (int i = 0; i < LFS_PROF_STAGES; ++i)
        t->acc_ms[i] = 0.f, t->acc_n[i] = 0;
    return LFS_OK;
}

extern "C" int lfs_trainer_get_profile(void* h, float* mean_ms, int* counts) {
    Trainer* t = static_cast<Trainer*>(h);
    LFS_CHECK_ARG(t && mean_ms, "trainer_get_profile: null pointer");
    for (int i = 0; i < LFS_PROF_STAGES; ++i) {
        mean_ms[i] = t->acc_n[i] ? t->acc_ms[i] / (float)t->acc_n[i] : 0.f;
        if (counts)
            counts[i] = t->acc_n[i];
    }
    return LFS_OK;
}

// Non-blocking: looks at the high-water mark the forwards copy to pinned host memory.  It only ever grows, so a view that
// overflowed is reported by every later call once its copy has completed (at the latest after the next synchronisation).
extern "C" int lfs_trainer_poll_capacity(void* h, uint64_t* max_instances_seen) {
    Trainer* t = static_cast<Trainer*>(h);
    LFS_CHECK_ARG(t, "trainer_poll_capacity: null handle");
    const uint32_t hw = t->stats_host ? *const_cast<volatile uint32_t*>(t->stats_host + 2) : 0u;
    if (max_instances_seen)
        *max_instances_seen = hw;
    if (hw > t->inst_cap) {
        set_error("trainer: a view needed %u tile instances, capacity is %u: its farthest instances were dropped; "
                  "recreate the trainer with a larger instance_capacity", hw, t->inst_cap);
        return LFS_ERR_CAPACITY;
    }
    return LFS_OK;
}

extern "C" int lfs_trainer_debug_copy(void* h, int which, void* dst, uint64_t dst_bytes, uint64_t* full_bytes,
                                      void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    LFS_CHECK_ARG(t, "trainer_debug_copy: null handle");
    const uint64_t npix = (uint64_t)t->d.width * t->d.height;
    uint32_t n_inst = 0;
    if (which == 3 || which == 4) {
        LFS_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
        LFS_CUDA_OK(cudaMemcpy(&n_inst, t->n_inst, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        n_inst = n_inst < t->inst_cap ? n_inst : t->inst_cap;
    }
    const void* src = nullptr;
    uint64_t bytes = 0;
    switch (which) {
    case 0: src = t->tile_off, bytes = 4ull * (t->n_tiles + 1); break;
    case 1: src = t->n_contrib, bytes = 4ull * npix; break;
    case 2: src = t->pix_state, bytes = 16ull * npix; break;
    case 3: src = t->sorted_vals, bytes = 4ull * n_inst; break;
    case 4: src = t->sorted_keys, bytes = 4ull * n_inst; break;
    case 5: src = t->bucket_off, bytes = 4ull * (t->n_tiles + 1); break;
    case 6: src = t->tile_max, bytes = 4ull * t->n_tiles; break;
    default: LFS_CHECK_ARG(false, "trainer_debug_copy: unknown buffer %d", which);
    }
    LFS_CHECK_ARG(src != nullptr, "trainer_debug_copy: no forward has been run");
    if (full_bytes)
        *full_bytes = bytes;
    const uint64_t n = bytes < dst_bytes ? bytes : dst_bytes;
    if (dst && n)
        LFS_CUDA_OK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
    return LFS_OK;
}

extern "C" int lfs_trainer_stats(void* h, uint64_t* n_instances, uint64_t* n_buckets, void* stream_) {
    Trainer* t = static_cast<Trainer*>(h);
    LFS_CHECK_ARG(t, "trainer_stats: null handle");
    LFS_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
    uint32_t v[2] = {0, 0};
    if (t->stats_host) {
        v[0] = t->stats_host[0], v[1] = t->stats_host[1];
    } else {
        LFS_CUDA_OK(cudaMemcpy(v, t->n_inst, sizeof(v), cudaMemcpyDeviceToHost));
    }
    if (n_instances)
        *n_instances = v[0];
    if (n_buckets)
        *n_buckets = v[1];
    if (v[0] > t->inst_cap) {
        set_error("trainer: %u instances exceed the capacity %u; recreate the trainer with a larger "
                  "instance_capacity", v[0], t->inst_cap);
        return LFS_ERR_CAPACITY;
    }
    return LFS_OK;
}
