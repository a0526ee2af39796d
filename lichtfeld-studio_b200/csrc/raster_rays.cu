// lfs_b200 -- from-world alpha blending along ARBITRARY per-pixel rays: the path for every camera that is not a perfect
// pinhole with a global shutter (OpenCV distortion, fisheye, rolling shutter), where the ray direction is not affine in
// the pixel and the tile-local rational-quadratic expansion of raster.cu does not apply.
// Reference behaviour: gsplat/RasterizeToPixelsFromWorld3DGSFwd.cu:58-279, ...Bwd.cu:63-372 (per pixel: world ray through
// the camera model and the shutter pose of that pixel; per pair: gro = M (o - mu), grd = normalize(M d),
// power = -1/2 |grd x gro|^2, M = S^-1 R^T).
//
// Same skeleton as the pinhole kernels -- tile CTA with active-pixel compaction and TMA-staged records in the forward,
// one warp per 32-instance bucket with a software-pipelined lane pipeline in the backward -- with
//   RayRec  (32 B per pixel):  origin, direction (world), validity         written once per view by k_make_rays
//   GenRec  (64 B per (camera, Gaussian)):  the 9 entries of M, M mu, opacity, rgb
// Per (pixel, instance) pair: M d, M o - M mu (18 FMA), cross product, two dot products, rsqrt, ex2.
#include "cameras.cuh"
#include "raster.cuh"

namespace lfs {

struct alignas(16) RayRec {
    float ox, oy, oz, valid;
    float dx, dy, dz, pad;
};
static_assert(sizeof(RayRec) == 32, "RayRec must be 32 bytes");

struct alignas(16) GenRec {
    float m[9];  // M = S^-1 R^T, row-major (m[3a + i] = R[i][a] / s_a)
    float mmu[3]; // M mu
    float opacity;
    float rgb[3];
};
static_assert(sizeof(GenRec) == 64, "GenRec must be 64 bytes");

struct CamArgs { // raw per-camera arguments of the gsplat ops (device pointers, optional ones may be null)
    const float *viewmats0, *viewmats1, *Ks, *radial, *tangential, *prism;
    int model, shutter;
    uint32_t width, height;
};

__device__ __forceinline__ CamModel cam_model_of(const CamArgs& a, const uint32_t cid) {
    const int n_rad = a.model == LFS_FISHEYE ? 4 : 6;
    return make_cam_model(a.viewmats0 + 16 * cid, a.viewmats1 ? a.viewmats1 + 16 * cid : nullptr, a.Ks + 9 * cid, a.width,
                          a.height, a.model, a.shutter, a.radial ? a.radial + n_rad * cid : nullptr,
                          a.tangential ? a.tangential + 2 * cid : nullptr, a.prism ? a.prism + 4 * cid : nullptr);
}

// one thread per pixel: world ray through the pixel centre (px + 0.5, py + 0.5)
__global__ void __launch_bounds__(256) k_make_rays(const CamArgs a, const uint32_t C, RayRec* __restrict__ rays) {
    const uint32_t hw = a.width * a.height;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t cid = blockIdx.y;
    if (i >= hw || cid >= C)
        return;
    const CamModel cam = cam_model_of(a, cid);
    const uint32_t py = i / a.width, px = i - py * a.width;
    f3 o, d;
    const bool ok = pixel_to_world_ray(cam, (float)px + 0.5f, (float)py + 0.5f, o, d);
    RayRec r;
    r.ox = o.x, r.oy = o.y, r.oz = o.z, r.valid = ok ? 1.f : 0.f;
    r.dx = d.x, r.dy = d.y, r.dz = d.z, r.pad = 0.f;
    rays[(size_t)cid * hw + i] = r;
}

__global__ void __launch_bounds__(256)
    k_prep_gaussians_gen(const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
                         const float* __restrict__ colors, const float* __restrict__ opacities, const uint32_t N,
                         const uint32_t total, const uint32_t channels, const uint32_t ch0, GenRec* __restrict__ out) {
    const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= total)
        return;
    const uint32_t gid = idx % N;
    const float4 q = __ldg(reinterpret_cast<const float4*>(quats) + gid);
    float w = q.x, x = q.y, y = q.z, z = q.w;
    const float inv_norm = rsqrtf(x * x + y * y + z * z + w * w); // quat_to_rotmat normalises (Utils.cuh:80-87)
    x *= inv_norm, y *= inv_norm, z *= inv_norm, w *= inv_norm;
    const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    const float R[9] = {1.f - 2.f * (y2 + z2), 2.f * (xy - wz), 2.f * (xz + wy), 2.f * (xy + wz), 1.f - 2.f * (x2 + z2),
                        2.f * (yz - wx),       2.f * (xz - wy), 2.f * (yz + wx), 1.f - 2.f * (x2 + y2)};
    const float is[3] = {1.0f / __ldg(scales + 3 * gid), 1.0f / __ldg(scales + 3 * gid + 1), 1.0f / __ldg(scales + 3 * gid + 2)};
    const float mu[3] = {__ldg(means + 3 * gid), __ldg(means + 3 * gid + 1), __ldg(means + 3 * gid + 2)};
    GenRec r;
#pragma unroll
    for (int a_ = 0; a_ < 3; ++a_) {
#pragma unroll
        for (int i_ = 0; i_ < 3; ++i_)
            r.m[3 * a_ + i_] = R[3 * i_ + a_] * is[a_];
        r.mmu[a_] = r.m[3 * a_] * mu[0] + r.m[3 * a_ + 1] * mu[1] + r.m[3 * a_ + 2] * mu[2];
    }
    r.opacity = __ldg(opacities + idx);
    const float* cp = colors + (size_t)idx * channels + ch0;
    r.rgb[0] = __ldg(cp), r.rgb[1] = ch0 + 1 < channels ? __ldg(cp + 1) : 0.f, r.rgb[2] = ch0 + 2 < channels ? __ldg(cp + 2) : 0.f;
    out[idx] = r;
}

// response of one pair: returns alpha_raw = opacity * exp(power); also the pieces the backward needs
struct PairGeo {
    f3 v, g, n, c; // M d, M o - M mu, normalised v, n x g
    float il, vis, power;
};
__device__ __forceinline__ void pair_geometry(const float4 r0, const float4 r1, const float4 r2, const f3 o, const f3 d,
                                              PairGeo& p) {
    // r0 = (m00 m01 m02 m10) r1 = (m11 m12 m20 m21) r2 = (m22 mmu0 mmu1 mmu2)
    p.v = mk3(fmaf(r0.x, d.x, fmaf(r0.y, d.y, r0.z * d.z)), fmaf(r0.w, d.x, fmaf(r1.x, d.y, r1.y * d.z)),
              fmaf(r1.z, d.x, fmaf(r1.w, d.y, r2.x * d.z)));
    p.g = mk3(fmaf(r0.x, o.x, fmaf(r0.y, o.y, fmaf(r0.z, o.z, -r2.y))), fmaf(r0.w, o.x, fmaf(r1.x, o.y, fmaf(r1.y, o.z, -r2.z))),
              fmaf(r1.z, o.x, fmaf(r1.w, o.y, fmaf(r2.x, o.z, -r2.w))));
    const float l = p.v.x * p.v.x + p.v.y * p.v.y + p.v.z * p.v.z;
    p.il = l > 0.f ? rsqrtf(l) : 1.f; // safe_normalize (Utils.cuh:181-184)
    p.n = p.v * p.il;
    p.c = cross(p.n, p.g);
    p.power = -0.5f * (p.c.x * p.c.x + p.c.y * p.c.y + p.c.z * p.c.z);
    p.vis = ex2_approx(p.power * 1.4426950408889634f);
}

// ------------------------------------------------------------------------------------------------------ forward
constexpr int kGBatch = 64;
constexpr int kGFwdThreads = kTilePix / 4;

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "GWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra GWAIT_DONE;\n\t"
                 "bra GWAIT_LOOP;\n\t"
                 "GWAIT_DONE:\n\t"
                 "}" ::"r"(g_smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void g_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     g_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(g_smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(kGFwdThreads)
    k_blend_fwd_rays(const RasterBuffers rb, const GenRec* __restrict__ gen, const RayRec* __restrict__ rays,
                     const uint32_t width, const uint32_t height, const uint32_t tile_w, const uint32_t tile_h,
                     const bool write_ckpt, const float* __restrict__ backgrounds, const uint8_t* __restrict__ masks,
                     float* __restrict__ renders, float* __restrict__ alphas, int32_t* __restrict__ last_ids) {
    __shared__ __align__(128) float4 s_rec[2][kGBatch * 4]; // GenRec double buffer, written by the TMA engine
    __shared__ __align__(16) float4 s_state[kTilePix];
    __shared__ uint32_t s_ncon[kTilePix];
    __shared__ uint8_t s_list[2][kTilePix];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_warp_tot[kGFwdThreads / 32];
    __shared__ uint32_t s_nact[2];

    const uint32_t tile = blockIdx.x, cam = blockIdx.y;
    const uint32_t n_tiles = tile_w * tile_h;
    const uint32_t ft = cam * n_tiles + tile;
    const uint32_t ty = tile / tile_w, tx = tile - ty * tile_w;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int32_t start = rb.tile_off[ft];
    int32_t end = rb.tile_off[ft + 1];
    const int32_t cnt_raw = end > start ? end - start : 0;
    if (masks && !masks[ft])
        end = start;
    const int32_t cnt = end > start ? end - start : 0;
    const uint32_t boff = rb.bucket_off ? rb.bucket_off[ft] : 0u;
    if (write_ckpt) {
        const uint32_t nb = (uint32_t)(cnt_raw + kBucket - 1) / kBucket;
        for (uint32_t k = tid; k < nb; k += kGFwdThreads)
            rb.bucket_tile[boff + k] = ft;
    }
    float4* ckpt_tile = write_ckpt ? rb.ckpt + (size_t)boff * kTilePix : nullptr;
    if (tid == 0) {
        g_mbar_init(&s_bar[0], kGFwdThreads);
        g_mbar_init(&s_bar[1], kGFwdThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // pixels inside the image whose ray is valid are active (an invalid ray leaves the pixel at background, ...Fwd.cu:138)
    const RayRec* rays_cam = rays + (size_t)cam * width * height;
    uint32_t keep = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t p = tid * 4 + k;
        s_state[p] = make_float4(0.f, 0.f, 0.f, 1.f);
        s_ncon[p] = 0;
        const uint32_t px = tx * kTile + (p & 15u), py = ty * kTile + (p >> 4);
        if (px < width && py < height && rays_cam[(size_t)py * width + px].valid != 0.f)
            keep |= 1u << k;
    }
    int cur = 0;
    auto compact = [&](const uint32_t keep_mask, const uint8_t ids[4], const int dst) {
        const uint32_t c = __popc(keep_mask);
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o)
                x += y;
        }
        if (lane == 31)
            s_warp_tot[warp] = x;
        __syncthreads();
        uint32_t base = x - c;
        if (warp == 1)
            base += s_warp_tot[0];
        if (tid == kGFwdThreads - 1)
            s_nact[dst] = base + c;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((keep_mask >> k) & 1u)
                s_list[dst][base++] = ids[k];
        __syncthreads();
    };
    {
        const uint8_t ids[4] = {(uint8_t)(tid * 4), (uint8_t)(tid * 4 + 1), (uint8_t)(tid * 4 + 2), (uint8_t)(tid * 4 + 3)};
        compact(keep, ids, cur);
    }
    const int nbatch = (cnt + kGBatch - 1) / kGBatch;
    int n_issued[2] = {0, 0}, n_waited[2] = {0, 0};
    auto issue = [&](const int kb_) { // batch kb_ -> buffer kb_ & 1
        const int b = kb_ & 1;
        const int j = kb_ * kGBatch + (int)tid;
        if (j < cnt) {
            const uint32_t gid = (uint32_t)__ldg(rb.inst_gid + start + j);
            g_mbar_expect_tx(&s_bar[b], (uint32_t)sizeof(GenRec));
            g_tma_load_1d(&s_rec[b][4 * tid], gen + gid, (uint32_t)sizeof(GenRec), &s_bar[b]);
        } else {
            g_mbar_arrive(&s_bar[b]);
        }
        ++n_issued[b];
    };
    auto wait = [&](const int b) {
        g_mbar_wait(&s_bar[b], (uint32_t)(n_waited[b] & 1));
        ++n_waited[b];
    };
    if (nbatch > 0)
        issue(0);
    if (nbatch > 1)
        issue(1);

    for (int kb = 0; kb < nbatch; ++kb) {
        const int buf = kb & 1;
        const uint32_t n_act = s_nact[cur];
        if (n_act == 0)
            break;
        uint8_t pid[4];
        float r[4], g[4], b[4], T[4];
        f3 ro[4], rd[4];
        uint32_t ncon[4];
        uint32_t live = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t j = tid * 4 + k;
            pid[k] = 0;
            ro[k] = rd[k] = mk3(0.f, 0.f, 0.f);
            r[k] = g[k] = b[k] = 0.f, T[k] = 1.f, ncon[k] = 0;
            if (j < n_act) {
                pid[k] = s_list[cur][j];
                const float4 st = s_state[pid[k]];
                r[k] = st.x, g[k] = st.y, b[k] = st.z, T[k] = st.w;
                ncon[k] = s_ncon[pid[k]];
                const uint32_t px = tx * kTile + (pid[k] & 15u), py = ty * kTile + (pid[k] >> 4);
                const float4* rp = reinterpret_cast<const float4*>(rays_cam + (size_t)py * width + px);
                const float4 a4 = __ldg(rp), b4 = __ldg(rp + 1);
                ro[k] = mk3(a4.x, a4.y, a4.z), rd[k] = mk3(b4.x, b4.y, b4.z);
                live |= 1u << k;
            }
        }
        const uint32_t mine = live;
        wait(buf);
        if (live) {
            const float4* s = &s_rec[buf][0];
            const int nrec = min(kGBatch, cnt - kb * kGBatch);
            const uint32_t first_li = (uint32_t)(kb * kGBatch);
            for (int t = 0; t < nrec; ++t) {
                const uint32_t li = first_li + t;
                if (write_ckpt && (li & (kBucket - 1)) == 0) {
                    float4* c = ckpt_tile + (size_t)(li >> 5) * kTilePix;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((live >> k) & 1u)
                            c[pid[k]] = make_float4(r[k], g[k], b[k], T[k]);
                }
                const float4 r0 = s[4 * t], r1 = s[4 * t + 1], r2 = s[4 * t + 2], E = s[4 * t + 3]; // E = (opacity, rgb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    PairGeo pg;
                    pair_geometry(r0, r1, r2, ro[k], rd[k], pg);
                    const float alpha = fminf(kAlphaMax, E.x * pg.vis);
                    const bool hit = ((live >> k) & 1u) && alpha >= kAlphaMin;
                    const float next_T = T[k] * (1.0f - alpha);
                    const bool stop = hit && next_T <= kTMin;
                    const bool acc = hit && !stop;
                    const float w = acc ? alpha * T[k] : 0.f;
                    r[k] = fmaf(w, E.y, r[k]);
                    g[k] = fmaf(w, E.z, g[k]);
                    b[k] = fmaf(w, E.w, b[k]);
                    T[k] = acc ? next_T : T[k];
                    ncon[k] = acc ? li + 1 : ncon[k];
                    live &= ~((stop ? 1u : 0u) << k);
                }
                if (live == 0)
                    break;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((mine >> k) & 1u) {
                s_state[pid[k]] = make_float4(r[k], g[k], b[k], T[k]);
                s_ncon[pid[k]] = ncon[k];
            }
        const int changed = __syncthreads_or((int)(live != mine)); // also: every thread is done with s_rec[buf]
        if (changed) {
            compact(live, pid, cur ^ 1);
            cur ^= 1;
        }
        if (kb + 2 < nbatch && s_nact[cur] != 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(kb + 2);
        }
    }
    for (int b = 0; b < 2; ++b) // drain copies still in flight (early exit)
        while (n_waited[b] < n_issued[b])
            wait(b);
    __syncthreads();

    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t p = tid * 4 + k;
        const uint32_t px = tx * kTile + (p & 15u), py = ty * kTile + (p >> 4);
        if (!(px < width && py < height))
            continue;
        const float4 st = s_state[p];
        const uint32_t nc = s_ncon[p];
        m = max(m, nc);
        const size_t pix = ((size_t)cam * height + py) * width + px;
        rb.pix_state[pix] = st;
        rb.n_contrib[pix] = (int32_t)nc;
        if (renders) {
            float br = 0.f, bgc = 0.f, bb = 0.f;
            if (backgrounds)
                br = backgrounds[cam * 3], bgc = backgrounds[cam * 3 + 1], bb = backgrounds[cam * 3 + 2];
            renders[pix * 3] = fmaf(st.w, br, st.x);
            renders[pix * 3 + 1] = fmaf(st.w, bgc, st.y);
            renders[pix * 3 + 2] = fmaf(st.w, bb, st.z);
        }
        if (alphas)
            alphas[pix] = 1.0f - st.w;
        if (last_ids)
            last_ids[pix] = nc > 0 ? start + (int32_t)nc - 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0)
        s_warp_tot[warp] = m;
    __syncthreads();
    if (tid == 0)
        rb.tile_max_contrib[ft] = max(s_warp_tot[0], s_warp_tot[1]);
}

// ------------------------------------------------------------------------------------------------------ backward
// One warp per 32-instance bucket, lane = instance (its GenRec and 16 gradient accumulators live in registers), the
// pixels that reach the bucket stream through the lanes front to back; (T, u) rotate with two shuffles per step
// (see k_blend_bwd_sp in raster.cu for the scheme).  Ring entry per pixel: 3 float4.
constexpr int kGBwdWarps = 4;

__global__ void __launch_bounds__(kGBwdWarps * 32)
    k_blend_bwd_rays(const RasterBuffers rb, const GenRec* __restrict__ gen, const RayRec* __restrict__ rays,
                     const float4* __restrict__ v_pix, const float* __restrict__ quats, const float* __restrict__ scales,
                     const float* __restrict__ means, const uint32_t N, const uint32_t width, const uint32_t height,
                     const uint32_t tile_w, const uint32_t tile_h, const uint32_t n_bucket_cap,
                     const uint32_t* __restrict__ n_buckets_dev, float* __restrict__ v_means, float* __restrict__ v_quats,
                     float* __restrict__ v_scales, float* __restrict__ v_colors, float* __restrict__ v_opacities) {
    __shared__ float4 ringA[kGBwdWarps][64]; // (v_r, v_g, v_b, bits(n_rel))
    __shared__ float4 ringB[kGBwdWarps][64]; // (o.x, o.y, o.z, T0)
    __shared__ float4 ringC[kGBwdWarps][64]; // (d.x, d.y, d.z, u0)
    __shared__ uint8_t s_pix[kGBwdWarps][kTilePix];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * kGBwdWarps + warp;
    uint32_t nbk = *n_buckets_dev;
    nbk = nbk < n_bucket_cap ? nbk : n_bucket_cap;
    if (b >= nbk)
        return;
    const uint32_t n_tiles = tile_w * tile_h;
    const uint32_t ft = rb.bucket_tile[b];
    const uint32_t cam = ft / n_tiles, tile = ft - cam * n_tiles;
    const uint32_t ty = tile / tile_w, tx = tile - ty * tile_w;
    const int32_t tstart = rb.tile_off[ft], tend = rb.tile_off[ft + 1];
    const uint32_t local_b = b - rb.bucket_off[ft];
    if (local_b * kBucket >= rb.tile_max_contrib[ft])
        return;
    const int32_t inst = tstart + (int32_t)(local_b * kBucket + lane);
    const bool valid = inst < tend;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, E = r0;
    uint32_t g = 0;
    if (valid) {
        g = (uint32_t)__ldg(rb.inst_gid + inst);
        const float4* gp = reinterpret_cast<const float4*>(gen + g);
        r0 = __ldg(gp), r1 = __ldg(gp + 1), r2 = __ldg(gp + 2), E = __ldg(gp + 3);
    }
    float vM[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vMmu[3] = {0.f, 0.f, 0.f};
    float acr = 0.f, acg = 0.f, acb = 0.f, aop = 0.f;
    float T = 0.f, u = 0.f;
    {   // lanes ahead of / behind the pixel list read ring slots no stage() has written yet: they run with zero weights,
        // but 0 * NaN is NaN, so the slots must hold finite numbers (shared memory is not cleared between kernels)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        ringA[warp][lane] = ringA[warp][lane + 32] = z4;
        ringB[warp][lane] = ringB[warp][lane + 32] = z4;
        ringC[warp][lane] = ringC[warp][lane + 32] = z4;
    }

    const float4* ck = rb.ckpt + (size_t)b * kTilePix;
    const RayRec* rays_cam = rays + (size_t)cam * width * height;
    int m = 0;
    {
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int rr = 0; rr < kTilePix / 32; ++rr) {
            const int p = rr * 32 + lane;
            const uint32_t px = tx * kTile + (p & 15), py = ty * kTile + (p >> 4);
            bool kp = false;
            if (px < width && py < height)
                kp = (uint32_t)__ldg(rb.n_contrib + ((size_t)cam * height + py) * width + px) > local_b * kBucket;
            const uint32_t mask = __ballot_sync(0xffffffffu, kp);
            if (kp)
                s_pix[warp][m + __popc(mask & lt)] = (uint8_t)p;
            m += __popc(mask);
        }
        __syncwarp();
    }
    auto stage = [&](const int i0) {
        __syncwarp();
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a4, c4 = a4;
        if (i0 + lane < m) {
            const int p = s_pix[warp][i0 + lane];
            const uint32_t px = tx * kTile + (p & 15), py = ty * kTile + (p >> 4);
            const size_t pl = (size_t)py * width + px, pix = (size_t)cam * height * width + pl;
            const int32_t nrel = __ldg(rb.n_contrib + pix) - (int32_t)(local_b * kBucket);
            const float4 k4 = ld_nc4(ck + p);
            const float4 f4 = __ldg(rb.pix_state + pix);
            const float4 v4 = __ldg(v_pix + pix);
            const float4* rp = reinterpret_cast<const float4*>(rays_cam + pl);
            const float4 o4 = __ldg(rp), d4 = __ldg(rp + 1);
            const float u0 = fmaf(f4.x - k4.x, v4.x, fmaf(f4.y - k4.y, v4.y, fmaf(f4.z - k4.z, v4.z, -v4.w)));
            a4 = make_float4(v4.x, v4.y, v4.z, __int_as_float(nrel));
            b4 = make_float4(o4.x, o4.y, o4.z, k4.w);
            c4 = make_float4(d4.x, d4.y, d4.z, u0);
        }
        ringA[warp][(i0 + lane) & 63] = a4;
        ringB[warp][(i0 + lane) & 63] = b4;
        ringC[warp][(i0 + lane) & 63] = c4;
        __syncwarp();
    };
    struct Ev {
        float vr, vg, vb, T0, u0, a_raw;
        f3 o, d;
        PairGeo pg;
        bool pass;
    };
    auto eval = [&](const int i, Ev& e) {
        const int idx = i - lane;
        const float4 ea = ringA[warp][idx & 63], eb = ringB[warp][idx & 63], ec = ringC[warp][idx & 63];
        e.vr = ea.x, e.vg = ea.y, e.vb = ea.z, e.T0 = eb.w, e.u0 = ec.w;
        e.o = mk3(eb.x, eb.y, eb.z), e.d = mk3(ec.x, ec.y, ec.z);
        const bool act = valid && (uint32_t)idx < (uint32_t)m && lane < __float_as_int(ea.w);
        pair_geometry(r0, r1, r2, e.o, e.d, e.pg);
        e.a_raw = E.x * e.pg.vis;
        e.pass = act && fminf(kAlphaMax, e.a_raw) >= kAlphaMin;
    };
    auto chain = [&](const Ev& e) {
        T = __shfl_up_sync(0xffffffffu, T, 1);
        u = __shfl_up_sync(0xffffffffu, u, 1);
        if (lane == 0)
            T = e.T0, u = e.u0;
        const float alpha = e.pass ? fminf(kAlphaMax, e.a_raw) : 0.f;
        const float w = T * alpha;
        acr = fmaf(w, e.vr, acr), acg = fmaf(w, e.vg, acg), acb = fmaf(w, e.vb, acb);
        const float Evv = fmaf(E.y, e.vr, fmaf(E.z, e.vg, E.w * e.vb));
        u = fmaf(-w, Evv, u);
        const float om = 1.0f - alpha;
        const float v_alpha = fmaf(T, Evv, -u * rcp_approx(om));
        T *= om;
        const bool grad = e.pass && e.a_raw <= kAlphaMax; // ...Bwd.cu:318
        aop += grad ? e.pg.vis * v_alpha : 0.f;
        // power = -1/2 |c|^2, alpha = opacity * exp(power):  dL/dc = -(dL/dalpha * alpha) c
        const float vp = grad ? v_alpha * e.a_raw : 0.f;
        const f3 vc = e.pg.c * (-vp);
        const f3 vn = cross(e.pg.g, vc); // c = n x g
        const f3 vg = cross(vc, e.pg.n);
        // n = v * il:  dL/dv = il vn - il^3 (vn . v) v
        const float il = e.pg.il, il3 = il * il * il;
        const float dnv = vn.x * e.pg.v.x + vn.y * e.pg.v.y + vn.z * e.pg.v.z;
        const f3 vv = mk3(il * vn.x - il3 * dnv * e.pg.v.x, il * vn.y - il3 * dnv * e.pg.v.y, il * vn.z - il3 * dnv * e.pg.v.z);
        const float vva[3] = {vv.x, vv.y, vv.z}, vga[3] = {vg.x, vg.y, vg.z};
        const float da[3] = {e.d.x, e.d.y, e.d.z}, oa[3] = {e.o.x, e.o.y, e.o.z};
#pragma unroll
        for (int a_ = 0; a_ < 3; ++a_) {
#pragma unroll
            for (int i_ = 0; i_ < 3; ++i_)
                vM[3 * a_ + i_] = fmaf(vva[a_], da[i_], fmaf(vga[a_], oa[i_], vM[3 * a_ + i_]));
            vMmu[a_] -= vga[a_];
        }
    };
    const int total = m + 31;
    Ev ea_, eb_;
    stage(0);
    eval(0, ea_);
    for (int i = 0; i < total; i += 2) {
        if (((i + 1) & 31) == 0 && i + 1 < m)
            stage(i + 1);
        eval(i + 1, eb_);
        chain(ea_);
        if (i + 1 >= total)
            break;
        if (((i + 2) & 31) == 0 && i + 2 < m)
            stage(i + 2);
        eval(i + 2, ea_);
        chain(eb_);
    }
    if (!valid)
        return;
    // ---- (M, M mu) -> (mean, quat, scale):  M[a][i] = R[i][a] / s_a,  M mu = M * mu
    const uint32_t gid = g % N;
    const f3 mu = mk3(__ldg(means + 3 * gid), __ldg(means + 3 * gid + 1), __ldg(means + 3 * gid + 2));
    const float mua[3] = {mu.x, mu.y, mu.z};
    const float M[9] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
    float v_mean[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int a_ = 0; a_ < 3; ++a_)
#pragma unroll
        for (int i_ = 0; i_ < 3; ++i_) {
            v_mean[i_] = fmaf(M[3 * a_ + i_], vMmu[a_], v_mean[i_]);
            vM[3 * a_ + i_] = fmaf(vMmu[a_], mua[i_], vM[3 * a_ + i_]);
        }
    const float4 q = __ldg(reinterpret_cast<const float4*>(quats) + gid);
    float qw = q.x, qx = q.y, qy = q.z, qz = q.w;
    const float inv_norm = rsqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
    qx *= inv_norm, qy *= inv_norm, qz *= inv_norm, qw *= inv_norm;
    const float R[9] = {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qw * qz), 2.f * (qx * qz + qw * qy),
                        2.f * (qx * qy + qw * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qw * qx),
                        2.f * (qx * qz - qw * qy), 2.f * (qy * qz + qw * qx), 1.f - 2.f * (qx * qx + qy * qy)};
    const float is[3] = {1.0f / __ldg(scales + 3 * gid), 1.0f / __ldg(scales + 3 * gid + 1), 1.0f / __ldg(scales + 3 * gid + 2)};
    float v_scale[3], GR[3][3];
#pragma unroll
    for (int a_ = 0; a_ < 3; ++a_) {
        float s = 0.f;
#pragma unroll
        for (int i_ = 0; i_ < 3; ++i_) {
            GR[i_][a_] = vM[3 * a_ + i_] * is[a_];
            s += R[i_ * 3 + a_] * vM[3 * a_ + i_];
        }
        v_scale[a_] = -is[a_] * is[a_] * s;
    }
    // quat_to_rotmat VJP (algebra of the reference's Utils.cuh:104-126; GR[r][c] = dL/dR[r][c])
    float vq[4];
    vq[0] = 2.f * (qx * (GR[2][1] - GR[1][2]) + qy * (GR[0][2] - GR[2][0]) + qz * (GR[1][0] - GR[0][1]));
    vq[1] = 2.f * (-2.f * qx * (GR[1][1] + GR[2][2]) + qy * (GR[1][0] + GR[0][1]) + qz * (GR[2][0] + GR[0][2]) +
                   qw * (GR[2][1] - GR[1][2]));
    vq[2] = 2.f * (qx * (GR[1][0] + GR[0][1]) - 2.f * qy * (GR[0][0] + GR[2][2]) + qz * (GR[2][1] + GR[1][2]) +
                   qw * (GR[0][2] - GR[2][0]));
    vq[3] = 2.f * (qx * (GR[2][0] + GR[0][2]) + qy * (GR[2][1] + GR[1][2]) - 2.f * qz * (GR[0][0] + GR[1][1]) +
                   qw * (GR[1][0] - GR[0][1]));
    const float qn[4] = {qw, qx, qy, qz};
    const float dq = vq[0] * qn[0] + vq[1] * qn[1] + vq[2] * qn[2] + vq[3] * qn[3];
    atomicAdd(v_means + 3 * gid, v_mean[0]);
    atomicAdd(v_means + 3 * gid + 1, v_mean[1]);
    atomicAdd(v_means + 3 * gid + 2, v_mean[2]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        atomicAdd(v_quats + 4 * gid + k, (vq[k] - dq * qn[k]) * inv_norm);
    atomicAdd(v_scales + 3 * gid, v_scale[0]);
    atomicAdd(v_scales + 3 * gid + 1, v_scale[1]);
    atomicAdd(v_scales + 3 * gid + 2, v_scale[2]);
    atomicAdd(v_colors + 3 * (size_t)g, acr);
    atomicAdd(v_colors + 3 * (size_t)g + 1, acg);
    atomicAdd(v_colors + 3 * (size_t)g + 2, acb);
    atomicAdd(v_opacities + g, aop);
}

// ------------------------------------------------------------------------------------------------------ launchers
size_t rays_scratch_bytes(uint32_t C, uint32_t N, uint64_t n_pix) {
    return align_up(sizeof(RayRec) * n_pix, 256) + align_up(sizeof(GenRec) * (size_t)C * N, 256);
}

int launch_rays_prepare(void* scratch, const float* means, const float* quats, const float* scales, const float* colors,
                        const float* opacities, uint32_t N, uint32_t C, uint32_t channels, uint32_t ch0, uint32_t width,
                        uint32_t height, const float* viewmats0, const float* viewmats1, const float* Ks, int camera_model,
                        int rs_type, const float* radial, const float* tangential, const float* prism, bool make_rays,
                        cudaStream_t stream) {
    RayRec* rays = static_cast<RayRec*>(scratch);
    GenRec* gen = reinterpret_cast<GenRec*>(static_cast<char*>(scratch) + align_up(sizeof(RayRec) * (size_t)C * width * height, 256));
    if (make_rays) {
        CamArgs a{viewmats0, viewmats1, Ks, radial, tangential, prism, camera_model, rs_type, width, height};
        k_make_rays<<<dim3(div_up((uint64_t)width * height, 256), C), 256, 0, stream>>>(a, C, rays);
        LFS_LAUNCH_OK("k_make_rays");
    }
    k_prep_gaussians_gen<<<div_up((uint64_t)C * N, 256), 256, 0, stream>>>(means, quats, scales, colors, opacities, N, C * N,
                                                                          channels, ch0, gen);
    LFS_LAUNCH_OK("k_prep_gaussians_gen");
    return LFS_OK;
}

int launch_blend_fwd_rays(const RasterBuffers& rb, void* scratch, uint32_t C, uint32_t N, uint32_t width, uint32_t height,
                          uint32_t tile_w, uint32_t tile_h, bool write_ckpt, const float* backgrounds, const uint8_t* masks,
                          float* renders, float* alphas, int32_t* last_ids, cudaStream_t stream) {
    (void)N;
    if (C == 0 || tile_w == 0 || tile_h == 0)
        return LFS_OK;
    const RayRec* rays = static_cast<const RayRec*>(scratch);
    const GenRec* gen = reinterpret_cast<const GenRec*>(static_cast<const char*>(scratch) +
                                                        align_up(sizeof(RayRec) * (size_t)C * width * height, 256));
    k_blend_fwd_rays<<<dim3(tile_w * tile_h, C), kGFwdThreads, 0, stream>>>(rb, gen, rays, width, height, tile_w, tile_h,
                                                                           write_ckpt, backgrounds, masks, renders, alphas,
                                                                           last_ids);
    LFS_LAUNCH_OK("k_blend_fwd_rays");
    return LFS_OK;
}

int launch_blend_bwd_rays(const RasterBuffers& rb, void* scratch, const float4* v_pix, const float* quats,
                          const float* scales, const float* means, uint32_t C, uint32_t N, uint32_t width, uint32_t height,
                          uint32_t tile_w, uint32_t tile_h, uint32_t n_bucket_cap, const uint32_t* n_buckets_dev,
                          float* v_means, float* v_quats, float* v_scales, float* v_colors, float* v_opacities,
                          cudaStream_t stream) {
    if (n_bucket_cap == 0)
        return LFS_OK;
    const RayRec* rays = static_cast<const RayRec*>(scratch);
    const GenRec* gen = reinterpret_cast<const GenRec*>(static_cast<const char*>(scratch) +
                                                        align_up(sizeof(RayRec) * (size_t)C * width * height, 256));
    k_blend_bwd_rays<<<div_up(n_bucket_cap, kGBwdWarps), kGBwdWarps * 32, 0, stream>>>(
        rb, gen, rays, v_pix, quats, scales, means, N, width, height, tile_w, tile_h, n_bucket_cap, n_buckets_dev, v_means,
        v_quats, v_scales, v_colors, v_opacities);
    LFS_LAUNCH_OK("k_blend_bwd_rays");
    return LFS_OK;
}

} // namespace lfs
