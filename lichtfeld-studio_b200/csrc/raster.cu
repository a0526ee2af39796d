// lfs_b200 -- from-world alpha-blend rasterizer kernels (see raster.cuh for the formulation and layouts).
// Reference behaviour being reproduced: gsplat/RasterizeToPixelsFromWorld3DGSFwd.cu:58-279 (forward) and
// gsplat/RasterizeToPixelsFromWorld3DGSBwd.cu:63-372 + gsplat/Utils.cuh:104-158 (backward).
#include "raster.cuh"
#include "sort_scan.cuh"

namespace lfs {

RasterOptions& raster_options() {
    static RasterOptions o;
    return o;
}

// Tile-local expansion of one GaussRec into the 12 rational-quadratic coefficients (see raster.cuh), inlined into the
// forward and backward blends, so every operation is spelled with
// explicit fmaf / __fmul_rn: the call sites must produce the same coefficient BITS (a pair has to get the same
// alpha in the forward and in the backward), which rules out leaving FMA contraction to the optimiser.
__device__ __forceinline__ float dot_rn(const f3 a, const f3 b) {
    return fmaf(a.z, b.z, fmaf(a.y, b.y, __fmul_rn(a.x, b.x)));
}
__device__ __forceinline__ f3 cross_rn(const f3 a, const f3 b) {
    return mk3(fmaf(a.y, b.z, -__fmul_rn(a.z, b.y)), fmaf(a.z, b.x, -__fmul_rn(a.x, b.z)),
               fmaf(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ void expand_record(const float4 g0, const float4 g1, const float4 g2, const float Xo,
                                              const float Yo, float4& A, float4& B, float4& Cc) {
    const f3 vx = mk3(g0.x, g0.y, g0.z), vy = mk3(g0.w, g1.x, g1.y), w2 = mk3(g1.z, g1.w, g2.x);
    const f3 gro = mk3(g2.y, g2.z, g2.w);
    const f3 v0 = mk3(fmaf(vy.x, Yo, fmaf(vx.x, Xo, w2.x)), fmaf(vy.y, Yo, fmaf(vx.y, Xo, w2.y)),
                      fmaf(vy.z, Yo, fmaf(vx.z, Xo, w2.z)));
    const f3 c0 = cross_rn(v0, gro), cxv = cross_rn(vx, gro), cyv = cross_rn(vy, gro);
    constexpr float k2 = 2.f * kNScale;
    A.x = __fmul_rn(kNScale, dot_rn(c0, c0));
    A.y = __fmul_rn(k2, dot_rn(c0, cxv));
    A.z = __fmul_rn(k2, dot_rn(c0, cyv));
    A.w = __fmul_rn(kNScale, dot_rn(cxv, cxv));
    B.x = __fmul_rn(k2, dot_rn(cxv, cyv));
    B.y = __fmul_rn(kNScale, dot_rn(cyv, cyv));
    B.z = dot_rn(v0, v0);
    B.w = __fmul_rn(2.f, dot_rn(v0, vx));
    Cc.x = __fmul_rn(2.f, dot_rn(v0, vy));
    Cc.y = dot_rn(vx, vx);
    Cc.z = __fmul_rn(2.f, dot_rn(vx, vy));
    Cc.w = dot_rn(vy, vy);
}

// EWA (fastgs surface) record: g0 = (mean2d.x, mean2d.y, conic.a, conic.b), g1 = (conic.c, opacity, r, g), g2.x = b
// (colour before the clamp).  sigma/2 = 1/2 a ex^2 + 1/2 c ey^2 + b ex ey with e = mean2d - pixel is a plain quadratic
// in the tile-local pixel offset (dx, dy):  N'(dx,dy) = -log2(e) * sigma/2, D == 1.
//   A = (n0, n1x, n1y, n2xx)   B = (n2xy, n2yy, opacity, max(r,0))   C = (max(g,0), max(b,0), -, -)
constexpr float kEScale = -1.4426950408889634f; // -log2(e)
__device__ __forceinline__ void expand_record_ewa(const float4 g0, const float4 g1, const float4 g2, const float tcx,
                                                  const float tcy, float4& A, float4& B, float4& Cc) {
    const float ex = g0.x - tcx, ey = g0.y - tcy, a = g0.z, b = g0.w, c = g1.x;
    const float ax = fmaf(a, ex, __fmul_rn(b, ey)), cy = fmaf(c, ey, __fmul_rn(b, ex)); // (conic e).x, (conic e).y
    A.x = __fmul_rn(kEScale * 0.5f, fmaf(ex, ax, __fmul_rn(ey, cy)));
    A.y = __fmul_rn(-kEScale, ax);
    A.z = __fmul_rn(-kEScale, cy);
    A.w = __fmul_rn(kEScale * 0.5f, a);
    B.x = __fmul_rn(kEScale, b);
    B.y = __fmul_rn(kEScale * 0.5f, c);
    B.z = g1.y;
    B.w = fmaxf(g1.z, 0.f);
    Cc = make_float4(fmaxf(g1.w, 0.f), fmaxf(g2.x, 0.f), 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------------
// bucket bookkeeping
// ------------------------------------------------------------------------------------------------------
__global__ void k_bucket_counts(const int32_t* __restrict__ tile_off, const uint32_t n_tiles,
                                uint32_t* __restrict__ counts) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles)
        return;
    const int32_t c = tile_off[t + 1] - tile_off[t];
    counts[t] = c > 0 ? (uint32_t)((c + kBucket - 1) / kBucket) : 0u;
}
__global__ void k_store_total(const uint32_t* __restrict__ total, uint32_t* __restrict__ off_last) {
    *off_last = *total;
}

int launch_bucket_offsets(const RasterBuffers& rb, uint32_t n_tiles_total, uint32_t* n_buckets_dev, void* scan_scratch,
                          uint32_t* counts_tmp, cudaStream_t stream) {
    k_bucket_counts<<<div_up(n_tiles_total, 256), 256, 0, stream>>>(rb.tile_off, n_tiles_total, counts_tmp);
    LFS_LAUNCH_OK("k_bucket_counts");
    int rc = exclusive_scan_u32(counts_tmp, nullptr, rb.bucket_off, n_buckets_dev, n_tiles_total, nullptr, scan_scratch,
                                stream);
    if (rc)
        return rc;
    k_store_total<<<1, 1, 0, stream>>>(n_buckets_dev, rb.bucket_off + n_tiles_total);
    LFS_LAUNCH_OK("k_store_total");
    return LFS_OK;
}

// ------------------------------------------------------------------------------------------------------
// forward blend
// ------------------------------------------------------------------------------------------------------
constexpr int kBatch = 64; // records per shared-memory stage (4 KB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "WAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra WAIT_DONE;\n\t"
                 "bra WAIT_LOOP;\n\t"
                 "WAIT_DONE:\n\t"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Forward blend with ACTIVE-PIXEL COMPACTION.
// A 16x16 tile is served by 64 threads; every thread works on up to 4 pixels taken from a compact list of the
// tile's still-active pixels that is rebuilt in shared memory after every staged batch of 64 records.  Pixels
// that saturate (T <= 1e-4) or lie outside the image drop out of the list, so the few pixels of a tile that never
// saturate (silhouettes, sky) no longer drag 32-wide warps through the whole instance list with 1-2 live lanes
// (measured before: 2.0 active threads per warp instruction on the C3 scene, profiles/r01_diag_first_gpu_run.json).  One staged record
// (4 x LDS.128) serves 4 pixel evaluations.  The polynomial nesting is mirrored exactly by the backward so a pair
// gets the same alpha bits in both passes.
constexpr int kFwdThreads = kTilePix / 4;

__device__ __forceinline__ float poly2(const float dx, const float dy, const float c0, const float cx,
                                       const float cy, const float cxx, const float cxy, const float cyy) {
    return fmaf(dx, fmaf(dx, cxx, fmaf(dy, cxy, cx)), fmaf(dy, fmaf(dy, cyy, cy), c0));
}

// (N', D)(dx, dy) with coefficient pairs c = {(n0,d0), (n1x,d1x), (n1y,d1y), (n2xx,d2xx), (n2xy,d2xy), (n2yy,d2yy)}:
// the same nesting as poly2, five FFMA2 instead of ten FFMA
__device__ __forceinline__ float2 poly2x2(const float2 DX, const float2 DY, const float2 c0, const float2 cx,
                                          const float2 cy, const float2 cxx, const float2 cxy, const float2 cyy) {
    return ffma2(DX, ffma2(DX, cxx, ffma2(DY, cxy, cx)), ffma2(DY, ffma2(DY, cyy, cy), c0));
}

// staged record of the PK (FFMA2) forward: (N', D) coefficient pairs, colour first
//   [0] = (n0,d0, n1x,d1x)  [1] = (n1y,d1y, n2xx,d2xx)  [2] = (n2xy,d2xy, n2yy,d2yy)  [3] = (r, g, b, opacity)
__device__ __forceinline__ void store_rec_pk(float4* d, const float4 A, const float4 B, const float4 Cc, const float4 E) {
    d[0] = make_float4(A.x, B.z, A.y, B.w);
    d[1] = make_float4(A.z, Cc.x, A.w, Cc.y);
    d[2] = make_float4(B.x, Cc.z, B.y, Cc.w);
    d[3] = make_float4(E.y, E.z, E.w, E.x);
}

// Register-staged variant (round 1's kernel, kept for A/B measurements: lfs_set_option("fwd_variant", 1)): the GaussRec of
// the next batch is gathered with LDG.128 into registers one batch ahead.  EWA: fastgs-surface records (2-D conic, D == 1);
// PK: FFMA2 evaluation of (N', D) (not EWA).
// After the forward: the buckets some pixel reaches (n_contrib > 32 * bucket) go onto the live list of the backward, one
// thread per tile (a few microseconds; as an epilogue of the forward kernel it cost 5 % of that kernel).
__global__ void __launch_bounds__(256) k_live_buckets(const RasterBuffers rb, const uint32_t n_tiles_total) {
    const uint32_t ft = blockIdx.x * 256 + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t n_live = 0, boff = 0;
    if (ft < n_tiles_total) {
        const int32_t start = rb.tile_off[ft], end = rb.tile_off[ft + 1];
        const uint32_t cnt_raw = end > start ? (uint32_t)(end - start) : 0u;
        const uint32_t nb = (cnt_raw + kBucket - 1) / kBucket;
        n_live = min(nb, (rb.tile_max_contrib[ft] + kBucket - 1) / kBucket);
        boff = rb.bucket_off[ft];
    }
    // one atomic per warp (8160 same-address atomics took longer than the rest of the kernel)
    uint32_t incl = n_live;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o)
            incl += y;
    }
    uint32_t base = 0;
    if (lane == 31 && incl)
        base = atomicAdd(rb.live, incl);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n_live;
    for (uint32_t k = 0; k < n_live; ++k)
        rb.live[2 + base + k] = boff + k;
}
static int launch_live_buckets(const RasterBuffers& rb, const uint32_t n_tiles_total, cudaStream_t stream) {
    k_live_buckets<<<div_up(n_tiles_total, 256), 256, 0, stream>>>(rb, n_tiles_total);
    LFS_LAUNCH_OK("k_live_buckets");
    return LFS_OK;
}

template <bool EWA, bool PK = false>
__global__ void __launch_bounds__(kFwdThreads)
    k_blend_fwd(const RasterBuffers rb, const ViewCam* __restrict__ cams, const uint32_t width, const uint32_t height, const uint32_t tile_w,
                const uint32_t tile_h, const bool write_ckpt, const float* __restrict__ backgrounds,
                const uint8_t* __restrict__ masks, float* __restrict__ renders, float* __restrict__ alphas,
                int32_t* __restrict__ last_ids) {
    __shared__ __align__(128) float4 s_rec[2][kBatch * 4];
    __shared__ __align__(16) float4 s_state[kTilePix]; // (r, g, b, T) per pixel
    __shared__ uint32_t s_ncon[kTilePix];
    __shared__ uint8_t s_list[2][kTilePix];
    __shared__ uint32_t s_warp_tot[kFwdThreads / 32];
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
    if (write_ckpt) { // bucket -> tile map for the backward (unmasked count: bucket_off is mask-agnostic)
        const uint32_t nb = (uint32_t)(cnt_raw + kBucket - 1) / kBucket;
        for (uint32_t k = tid; k < nb; k += kFwdThreads)
            rb.bucket_tile[boff + k] = ft;
    }
    float4* ckpt_tile = write_ckpt ? rb.ckpt + (size_t)boff * kTilePix : nullptr;

    // ---- initial state + initial active list (pixels inside the image, row-major order)
    uint32_t keep = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t p = tid * 4 + k;
        s_state[p] = make_float4(0.f, 0.f, 0.f, 1.f);
        s_ncon[p] = 0;
        const uint32_t px = tx * kTile + (p & 15u), py = ty * kTile + (p >> 4);
        keep |= (px < width && py < height ? 1u : 0u) << k;
    }
    int cur = 0;
    auto compact = [&](const uint32_t keep_mask, const uint8_t ids[4], const int dst) {
        // block-wide exclusive scan of popc(keep_mask) over 64 threads, then scatter the kept ids
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
        if (tid == kFwdThreads - 1)
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

    float Xo, Yo;
    uint32_t gid_next = 0; // flattened Gaussian id this thread expands for the NEXT batch
    if (EWA) {
        Xo = (float)(tx * kTile + kTile / 2), Yo = (float)(ty * kTile + kTile / 2); // tile centre in pixel units
    } else {
        Xo = (float)(tx * kTile + kTile / 2) - cams[cam].cx;
        Yo = (float)(ty * kTile + kTile / 2) - cams[cam].cy;
    }
    const int nbatch = (cnt + kBatch - 1) / kBatch;
    if (nbatch > 0) {
        if ((int)tid < min(kBatch, cnt)) {
            const float4* gp = reinterpret_cast<const float4*>(rb.gauss + (uint32_t)__ldg(rb.inst_gid + start + tid));
            float4 A, B, Cc;
            if (EWA) {
                expand_record_ewa(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), Xo, Yo, A, B, Cc);
            } else {
                expand_record(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), Xo, Yo, A, B, Cc);
                if (PK)
                    store_rec_pk(&s_rec[0][4 * tid], A, B, Cc, __ldg(gp + 3));
                else
                    s_rec[0][4 * tid + 3] = __ldg(gp + 3);
            }
            if (!PK)
                s_rec[0][4 * tid] = A, s_rec[0][4 * tid + 1] = B, s_rec[0][4 * tid + 2] = Cc;
        }
        if ((int)tid < cnt - kBatch)
            gid_next = (uint32_t)__ldg(rb.inst_gid + start + kBatch + tid);
        __syncthreads();
    }

    for (int kb = 0; kb < nbatch; ++kb) {
        const int buf = kb & 1;
        const uint32_t n_act = s_nact[cur];
        if (n_act == 0)
            break;
        // this thread's slice of the active list
        uint8_t pid[4];
        float r[4], g[4], b[4], T[4], dx[4], dy[4];
        uint32_t ncon[4];
        uint32_t live = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t j = tid * 4 + k;
            pid[k] = 0;
            if (j < n_act) {
                pid[k] = s_list[cur][j];
                const float4 st = s_state[pid[k]];
                r[k] = st.x, g[k] = st.y, b[k] = st.z, T[k] = st.w;
                ncon[k] = s_ncon[pid[k]];
                dx[k] = (float)(pid[k] & 15u) - 7.5f;
                dy[k] = (float)(pid[k] >> 4) - 7.5f;
                live |= 1u << k;
            } else {
                r[k] = g[k] = b[k] = 0.f, T[k] = 1.f, ncon[k] = 0, dx[k] = dy[k] = 0.f;
            }
        }
        const uint32_t mine = live;

        // gather the next batch's GaussRec (this thread's record) now; it is expanded after the blend below
        float4 pre[4];
        const int next_f4 = (kb + 1 < nbatch) ? min(kBatch, cnt - (kb + 1) * kBatch) : 0;
        if ((int)tid < next_f4) {
            const float4* gp = reinterpret_cast<const float4*>(rb.gauss + gid_next);
            pre[0] = __ldg(gp), pre[1] = __ldg(gp + 1), pre[2] = __ldg(gp + 2);
            if (!EWA)
                pre[3] = __ldg(gp + 3);
        }
        if ((int)tid < cnt - (kb + 2) * kBatch)
            gid_next = (uint32_t)__ldg(rb.inst_gid + start + (kb + 2) * kBatch + tid);

        if (live) {
            const float4* s = &s_rec[buf][0];
            const int nrec = min(kBatch, cnt - kb * kBatch);
            const uint32_t first_li = (uint32_t)(kb * kBatch);
            for (int t = 0; t < nrec; ++t) {
                const uint32_t li = first_li + t;
                if (write_ckpt && (li & (kBucket - 1)) == 0) {
                    float4* c = ckpt_tile + (size_t)(li >> 5) * kTilePix;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((live >> k) & 1u)
                            c[pid[k]] = make_float4(r[k], g[k], b[k], T[k]);
                }
                const float4 A = s[4 * t], B = s[4 * t + 1], Cc = s[4 * t + 2];
                float4 E;
                if (EWA)
                    E = make_float4(B.z, B.w, Cc.x, Cc.y);
                else
                    E = s[4 * t + 3];
                if (PK) { // A, B, Cc hold the three pair-interleaved float4, E = (r, g, b, opacity)
                    const float2 P0 = make_float2(A.x, A.y), P1 = make_float2(A.z, A.w), P2 = make_float2(B.x, B.y),
                                 P3 = make_float2(B.z, B.w), P4 = make_float2(Cc.x, Cc.y), P5 = make_float2(Cc.z, Cc.w);
                    const float2 Erg = make_float2(E.x, E.y);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 ND = poly2x2(make_float2(dx[k], dx[k]), make_float2(dy[k], dy[k]), P0, P1, P2, P3,
                                                  P4, P5);
                        const float vis = ex2_approx(ND.x * rcp_approx(ND.y));
                        const float alpha = fminf(kAlphaMax, E.w * vis);
                        // branch-free (selects): the four pixels of a thread form one basic block the scheduler can
                        // interleave; a pair that does not contribute blends with weight 0
                        const bool hit = ((live >> k) & 1u) && alpha >= kAlphaMin;
                        const float next_T = T[k] * (1.0f - alpha);
                        const bool stop = hit && next_T <= kTMin; // gsplat stops at T <= 1e-4 (...Fwd.cu:244-248)
                        const bool acc = hit && !stop;
                        const float w = acc ? alpha * T[k] : 0.f;
                        const float2 rg = ffma2(make_float2(w, w), Erg, make_float2(r[k], g[k]));
                        r[k] = rg.x, g[k] = rg.y;
                        b[k] = fmaf(w, E.z, b[k]);
                        T[k] = acc ? next_T : T[k];
                        ncon[k] = acc ? li + 1 : ncon[k];
                        live &= ~((stop ? 1u : 0u) << k);
                    }
                    if (live == 0)
                        break;
                    continue;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float Nv = poly2(dx[k], dy[k], A.x, A.y, A.z, A.w, B.x, B.y);
                    float vis;
                    bool ok = true;
                    if (EWA) {
                        vis = ex2_approx(Nv);
                        ok = Nv <= 0.f; // sigma/2 < 0 is skipped (kernels_forward.cuh:423-424)
                    } else {
                        const float Dv = poly2(dx[k], dy[k], B.z, B.w, Cc.x, Cc.y, Cc.z, Cc.w);
                        vis = ex2_approx(Nv * rcp_approx(Dv));
                    }
                    const float alpha = fminf(kAlphaMax, E.x * vis);
                    if (((live >> k) & 1u) && alpha >= kAlphaMin && ok) {
                        const float next_T = T[k] * (1.0f - alpha);
                        // gsplat stops at T <= 1e-4 (RasterizeToPixelsFromWorld3DGSFwd.cu), fastgs at T < 1e-4
                        if (EWA ? (next_T < kTMin) : (next_T <= kTMin)) {
                            live &= ~(1u << k);
                        } else {
                            const float w = alpha * T[k];
                            r[k] = fmaf(w, E.y, r[k]);
                            g[k] = fmaf(w, E.z, g[k]);
                            b[k] = fmaf(w, E.w, b[k]);
                            T[k] = next_T;
                            ncon[k] = li + 1;
                        }
                    }
                }
                if (live == 0)
                    break;
            }
        }
        // write back the pixels this thread owned in this batch
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((mine >> k) & 1u) {
                s_state[pid[k]] = make_float4(r[k], g[k], b[k], T[k]);
                s_ncon[pid[k]] = ncon[k];
            }
        if ((int)tid < next_f4) {
            float4 A, B, Cc;
            float4* d = &s_rec[buf ^ 1][4 * tid];
            if (EWA) {
                expand_record_ewa(pre[0], pre[1], pre[2], Xo, Yo, A, B, Cc);
            } else {
                expand_record(pre[0], pre[1], pre[2], Xo, Yo, A, B, Cc);
                if (PK)
                    store_rec_pk(d, A, B, Cc, pre[3]);
                else
                    d[3] = pre[3];
            }
            if (!PK)
                d[0] = A, d[1] = B, d[2] = Cc;
        }
        // rebuild the active list only when some pixel finished in this batch (compact() syncs the block)
        const int changed = __syncthreads_or((int)(live != mine));
        if (changed) {
            compact(live, pid, cur ^ 1);
            cur ^= 1;
        }
    }
    __syncthreads();

    // ---- outputs: thread t writes its 4 row-adjacent pixels (64 B contiguous per thread)
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
            if (backgrounds) {
                br = backgrounds[cam * 3], bgc = backgrounds[cam * 3 + 1], bb = backgrounds[cam * 3 + 2];
            }
            renders[pix * 3] = fmaf(st.w, br, st.x);
            renders[pix * 3 + 1] = fmaf(st.w, bgc, st.y);
            renders[pix * 3 + 2] = fmaf(st.w, bb, st.z);
        }
        if (alphas)
            alphas[pix] = 1.0f - st.w;
        if (last_ids)
            last_ids[pix] = nc > 0 ? start + (int32_t)nc - 1 : 0;
    }
    // per-tile maximum contributor count (lets the backward skip whole buckets)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0)
        s_warp_tot[warp] = m;
    __syncthreads();
    const uint32_t tile_max = max(s_warp_tot[0], s_warp_tot[1]);
    if (tid == 0)
        rb.tile_max_contrib[ft] = tile_max;
}

// ------------------------------------------------------------------------------------------------------
// forward blend, TMA gather (the default).  Same algorithm and arithmetic as k_blend_fwd<2, ., ., PK> (active-pixel
// compaction, in-kernel expansion, FFMA2 evaluation, branch-free pixel update), but the per-tile Gaussian records of
// the NEXT batch are staged into shared memory by the TMA engine: every thread issues one 64-byte cp.async.bulk of
// "its" GaussRec (gathered through inst_gid) that completes on an mbarrier, blends the current batch while the copies
// are in flight, then expands the raw records from shared memory.  No record ever passes through registers on its way
// in (the register-staged version held 16 registers of prefetched record per thread).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool EWA, int kMinBlocks = 10>
__global__ void __launch_bounds__(kFwdThreads, kMinBlocks)
    k_blend_fwd_tg(const RasterBuffers rb, const ViewCam* __restrict__ cams, const uint32_t width, const uint32_t height,
                   const uint32_t tile_w, const uint32_t tile_h, const bool write_ckpt,
                   const float* __restrict__ backgrounds, const uint8_t* __restrict__ masks, float* __restrict__ renders,
                   float* __restrict__ alphas, int32_t* __restrict__ last_ids) {
    __shared__ __align__(128) float4 s_raw[kBatch * 4]; // GaussRec of the next batch, written by the TMA engine
    __shared__ __align__(16) float4 s_rec[2][kBatch * 4]; // expanded tile-local records (store_rec_pk layout)
    __shared__ __align__(16) float4 s_state[kTilePix];    // (r, g, b, T) per pixel
    __shared__ uint32_t s_ncon[kTilePix];
    __shared__ uint8_t s_list[2][kTilePix];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_warp_tot[kFwdThreads / 32];
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
    if (write_ckpt) { // bucket -> tile map for the backward (unmasked count: bucket_off is mask-agnostic)
        const uint32_t nb = (uint32_t)(cnt_raw + kBucket - 1) / kBucket;
        for (uint32_t k = tid; k < nb; k += kFwdThreads)
            rb.bucket_tile[boff + k] = ft;
    }
    float4* ckpt_tile = write_ckpt ? rb.ckpt + (size_t)boff * kTilePix : nullptr;
    if (tid == 0) {
        mbar_init(&s_bar, kFwdThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- initial state + initial active list (pixels inside the image, row-major order)
    uint32_t keep = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t p = tid * 4 + k;
        s_state[p] = make_float4(0.f, 0.f, 0.f, 1.f);
        s_ncon[p] = 0;
        const uint32_t px = tx * kTile + (p & 15u), py = ty * kTile + (p >> 4);
        keep |= (px < width && py < height ? 1u : 0u) << k;
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
        if (tid == kFwdThreads - 1)
            s_nact[dst] = base + c;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((keep_mask >> k) & 1u)
                s_list[dst][base++] = ids[k];
        __syncthreads();
    };
    {
        const uint8_t ids[4] = {(uint8_t)(tid * 4), (uint8_t)(tid * 4 + 1), (uint8_t)(tid * 4 + 2), (uint8_t)(tid * 4 + 3)};
        compact(keep, ids, cur); // also orders the mbarrier init before its first use
    }

    const float Xo = (float)(tx * kTile + kTile / 2) - (EWA ? 0.f : cams[cam].cx);
    const float Yo = (float)(ty * kTile + kTile / 2) - (EWA ? 0.f : cams[cam].cy);
    const int nbatch = (cnt + kBatch - 1) / kBatch;
    int n_issued = 0, n_waited = 0; // uniform over the CTA: completed phases of s_bar

    // every thread arrives once per batch; those that own a record of the batch add 64 bytes to the transaction count
    // and launch the copy (gathered address: rb.gauss + gid)
    auto issue = [&](const int kb_next, const uint32_t gid) {
        const int nrec = min(kBatch, cnt - kb_next * kBatch);
        if ((int)tid < nrec) {
            mbar_expect_tx(&s_bar, (uint32_t)sizeof(GaussRec));
            tma_load_1d(&s_raw[4 * tid], rb.gauss + gid, (uint32_t)sizeof(GaussRec), &s_bar);
        } else {
            mbar_arrive(&s_bar);
        }
        ++n_issued;
    };
    auto wait_expand = [&](const int kb_next, const int dst) {
        mbar_wait(&s_bar, (uint32_t)(n_waited & 1));
        ++n_waited;
        const int nrec = min(kBatch, cnt - kb_next * kBatch);
        if ((int)tid < nrec) {
            const float4 g0 = s_raw[4 * tid], g1 = s_raw[4 * tid + 1], g2 = s_raw[4 * tid + 2], g3 = s_raw[4 * tid + 3];
            float4 A, B, Cc;
            float4* d = &s_rec[dst][4 * tid];
            if (EWA) {
                expand_record_ewa(g0, g1, g2, Xo, Yo, A, B, Cc);
                d[0] = A, d[1] = B, d[2] = Cc;
            } else {
                expand_record(g0, g1, g2, Xo, Yo, A, B, Cc);
                store_rec_pk(d, A, B, Cc, g3);
            }
        }
    };
    auto load_gid = [&](const int kb_) -> uint32_t {
        const int j = kb_ * kBatch + (int)tid;
        return j < cnt ? (uint32_t)__ldg(rb.inst_gid + start + j) : 0u;
    };

    uint32_t gid_next = 0; // Gaussian id of this thread's record in the batch issued next
    if (nbatch > 0) {
        issue(0, load_gid(0));
        gid_next = load_gid(1);
        wait_expand(0, 0);
        __syncthreads(); // s_rec[0] complete; every thread is done reading s_raw
        if (nbatch > 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(1, gid_next);
            gid_next = load_gid(2);
        }
    }

    for (int kb = 0; kb < nbatch; ++kb) {
        const int buf = kb & 1;
        const uint32_t n_act = s_nact[cur];
        if (n_act == 0)
            break;
        uint8_t pid[4];
        float r[4], g[4], b[4], T[4], dx[4], dy[4];
        uint32_t ncon[4];
        uint32_t live = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t j = tid * 4 + k;
            pid[k] = 0;
            if (j < n_act) {
                pid[k] = s_list[cur][j];
                const float4 st = s_state[pid[k]];
                r[k] = st.x, g[k] = st.y, b[k] = st.z, T[k] = st.w;
                ncon[k] = s_ncon[pid[k]];
                dx[k] = (float)(pid[k] & 15u) - 7.5f;
                dy[k] = (float)(pid[k] >> 4) - 7.5f;
                live |= 1u << k;
            } else {
                r[k] = g[k] = b[k] = 0.f, T[k] = 1.f, ncon[k] = 0, dx[k] = dy[k] = 0.f;
            }
        }
        const uint32_t mine = live;

        if (live) {
            const float4* s = &s_rec[buf][0];
            const int nrec = min(kBatch, cnt - kb * kBatch);
            const uint32_t first_li = (uint32_t)(kb * kBatch);
            for (int t = 0; t < nrec; ++t) {
                const uint32_t li = first_li + t;
                if (write_ckpt && (li & (kBucket - 1)) == 0) {
                    float4* c = ckpt_tile + (size_t)(li >> 5) * kTilePix;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((live >> k) & 1u)
                            c[pid[k]] = make_float4(r[k], g[k], b[k], T[k]);
                }
                const float4 R0 = s[4 * t], R1 = s[4 * t + 1], R2 = s[4 * t + 2];
                if (EWA) { // R0 = (n0, n1x, n1y, n2xx), R1 = (n2xy, n2yy, opacity, max(r,0)), R2 = (max(g,0), max(b,0), -, -)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float Nv = poly2(dx[k], dy[k], R0.x, R0.y, R0.z, R0.w, R1.x, R1.y);
                        const float alpha = fminf(kAlphaMax, R1.z * ex2_approx(Nv));
                        // sigma/2 < 0 is skipped (kernels_forward.cuh:423-424); fastgs stops at T < 1e-4 (strict)
                        const bool hit = ((live >> k) & 1u) && alpha >= kAlphaMin && Nv <= 0.f;
                        const float next_T = T[k] * (1.0f - alpha);
                        const bool stop = hit && next_T < kTMin;
                        const bool acc = hit && !stop;
                        const float w = acc ? alpha * T[k] : 0.f;
                        r[k] = fmaf(w, R1.w, r[k]);
                        g[k] = fmaf(w, R2.x, g[k]);
                        b[k] = fmaf(w, R2.y, b[k]);
                        T[k] = acc ? next_T : T[k];
                        ncon[k] = acc ? li + 1 : ncon[k];
                        live &= ~((stop ? 1u : 0u) << k);
                    }
                } else {
                    const float4 E = s[4 * t + 3]; // (r, g, b, opacity)
                    const float2 P0 = make_float2(R0.x, R0.y), P1 = make_float2(R0.z, R0.w), P2 = make_float2(R1.x, R1.y),
                                 P3 = make_float2(R1.z, R1.w), P4 = make_float2(R2.x, R2.y), P5 = make_float2(R2.z, R2.w);
                    const float2 Erg = make_float2(E.x, E.y);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 ND = poly2x2(make_float2(dx[k], dx[k]), make_float2(dy[k], dy[k]), P0, P1, P2, P3, P4,
                                                  P5);
                        const float vis = ex2_approx(ND.x * rcp_approx(ND.y));
                        const float alpha = fminf(kAlphaMax, E.w * vis);
                        const bool hit = ((live >> k) & 1u) && alpha >= kAlphaMin;
                        const float next_T = T[k] * (1.0f - alpha);
                        const bool stop = hit && next_T <= kTMin; // gsplat stops at T <= 1e-4 (...Fwd.cu:244-248)
                        const bool acc = hit && !stop;
                        const float w = acc ? alpha * T[k] : 0.f;
                        const float2 rg = ffma2(make_float2(w, w), Erg, make_float2(r[k], g[k]));
                        r[k] = rg.x, g[k] = rg.y;
                        b[k] = fmaf(w, E.z, b[k]);
                        T[k] = acc ? next_T : T[k];
                        ncon[k] = acc ? li + 1 : ncon[k];
                        live &= ~((stop ? 1u : 0u) << k);
                    }
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
        if (kb + 1 < nbatch) // the copies of batch kb + 1 have been in flight during the blend above
            wait_expand(kb + 1, buf ^ 1);
        // rebuild the active list only when some pixel finished in this batch; the barrier also means that every thread is
        // done with s_raw and with s_rec[buf]
        const int changed = __syncthreads_or((int)(live != mine));
        if (changed) {
            compact(live, pid, cur ^ 1);
            cur ^= 1;
        }
        if (kb + 2 < nbatch && s_nact[cur] != 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(kb + 2, gid_next);
            gid_next = load_gid(kb + 3);
        }
    }
    // a batch may still be in flight into this CTA's shared memory (early exit, or no active pixel at all): drain it
    while (n_waited < n_issued) {
        mbar_wait(&s_bar, (uint32_t)(n_waited & 1));
        ++n_waited;
    }
    __syncthreads();

    // ---- outputs: thread t writes its 4 row-adjacent pixels (64 B contiguous per thread)
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
            if (backgrounds) {
                br = backgrounds[cam * 3], bgc = backgrounds[cam * 3 + 1], bb = backgrounds[cam * 3 + 2];
            }
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
    const uint32_t tile_max = max(s_warp_tot[0], s_warp_tot[1]);
    if (tid == 0)
        rb.tile_max_contrib[ft] = tile_max;
}

int launch_blend_fwd(const RasterBuffers& rb, const ViewCam* cams_dev, uint32_t C, uint32_t width, uint32_t height,
                     uint32_t tile_w, uint32_t tile_h, bool write_ckpt, const float* backgrounds, const uint8_t* masks,
                     float* renders, float* alphas, int32_t* last_ids, cudaStream_t stream) {
    if (C == 0 || tile_w == 0 || tile_h == 0)
        return LFS_OK;
    if (rb.live && write_ckpt) // live-bucket count and the backward's work counter
        LFS_CUDA_OK(cudaMemsetAsync(rb.live, 0, 2 * sizeof(uint32_t), stream));
    dim3 grid(tile_w * tile_h, C);
    if (raster_options().fwd_variant == 1)
        k_blend_fwd<false, true><<<grid, kFwdThreads, 0, stream>>>(rb, cams_dev, width, height, tile_w, tile_h, write_ckpt,
                                                                  backgrounds, masks, renders, alphas, last_ids);
    else if (raster_options().fwd_variant == 2) // A/B: 12 CTAs per SM at 80 registers
        k_blend_fwd_tg<false, 12><<<grid, kFwdThreads, 0, stream>>>(rb, cams_dev, width, height, tile_w, tile_h, write_ckpt,
                                                                   backgrounds, masks, renders, alphas, last_ids);
    else
        k_blend_fwd_tg<false><<<grid, kFwdThreads, 0, stream>>>(rb, cams_dev, width, height, tile_w, tile_h, write_ckpt,
                                                               backgrounds, masks, renders, alphas, last_ids);
    LFS_LAUNCH_OK("k_blend_fwd");
    if (rb.live && write_ckpt)
        return launch_live_buckets(rb, tile_w * tile_h * C, stream);
    return LFS_OK;
}

int launch_blend_fwd_ewa(const RasterBuffers& rb, uint32_t width, uint32_t height, uint32_t tile_w, uint32_t tile_h,
                         bool write_ckpt, cudaStream_t stream) {
    if (tile_w == 0 || tile_h == 0)
        return LFS_OK;
    if (rb.live && write_ckpt)
        LFS_CUDA_OK(cudaMemsetAsync(rb.live, 0, 2 * sizeof(uint32_t), stream));
    if (raster_options().fwd_variant == 1)
        k_blend_fwd<true><<<dim3(tile_w * tile_h, 1), kFwdThreads, 0, stream>>>(
            rb, nullptr, width, height, tile_w, tile_h, write_ckpt, nullptr, nullptr, nullptr, nullptr, nullptr);
    else
        k_blend_fwd_tg<true><<<dim3(tile_w * tile_h, 1), kFwdThreads, 0, stream>>>(
            rb, nullptr, width, height, tile_w, tile_h, write_ckpt, nullptr, nullptr, nullptr, nullptr, nullptr);
    LFS_LAUNCH_OK("k_blend_fwd<ewa>");
    if (rb.live && write_ckpt)
        return launch_live_buckets(rb, tile_w * tile_h, stream);
    return LFS_OK;
}

// ------------------------------------------------------------------------------------------------------
// backward blend: one warp per 32-instance bucket
// ------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void quat_to_rotmat_dev(const float4 q_wxyz, float R[9], float& inv_norm, float qn[4]) {
    float w = q_wxyz.x, x = q_wxyz.y, y = q_wxyz.z, z = q_wxyz.w;
    inv_norm = rsqrtf(x * x + y * y + z * z + w * w);
    x *= inv_norm, y *= inv_norm, z *= inv_norm, w *= inv_norm;
    qn[0] = w, qn[1] = x, qn[2] = y, qn[3] = z;
    const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y,
                wz = w * z;
    R[0] = 1.f - 2.f * (y2 + z2), R[1] = 2.f * (xy - wz), R[2] = 2.f * (xz + wy);
    R[3] = 2.f * (xy + wz), R[4] = 1.f - 2.f * (x2 + z2), R[5] = 2.f * (yz - wx);
    R[6] = 2.f * (xz - wy), R[7] = 2.f * (yz + wx), R[8] = 1.f - 2.f * (x2 + y2);
}

// per-instance chain rule of the from-world backward: gradients of the 12 tile-local polynomial coefficients ->
// (vx, vy, w2, gro) -> (mean, quat, scale), then one red.global.add per output
__device__ __forceinline__ void bwd_chain_rule(const RasterBuffers& rb, const ViewCam& cm_in, const uint32_t g, const uint32_t N,
                                               const float Xo, const float Yo, const float an0, const float an1,
                                               const float an2, const float an3, const float an4, const float an5,
                                               const float ad0, const float ad1, const float ad2, const float ad3,
                                               const float ad4, const float ad5, const float acr, const float acg,
                                               const float acb, const float aop, const float* __restrict__ quats,
                                               const float* __restrict__ scales, const float* __restrict__ means,
                                               float* __restrict__ v_means, float* __restrict__ v_quats,
                                               float* __restrict__ v_scales, float* __restrict__ v_colors,
                                               float* __restrict__ v_opacities) {
    // ---- per-instance chain rule: polynomial coefficients -> (vx, vy, w2, gro) -> (mean, quat, scale)
    const ViewCam& cm = cm_in;
    const uint32_t gid = g % N;
    const float4* gp = reinterpret_cast<const float4*>(rb.gauss + g);
    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
    const f3 vx = mk3(g0.x, g0.y, g0.z), vy = mk3(g0.w, g1.x, g1.y), w2 = mk3(g1.z, g1.w, g2.x);
    const f3 gro = mk3(g2.y, g2.z, g2.w);
    const f3 v0 = w2 + vx * Xo + vy * Yo;
    const f3 c0 = cross(v0, gro), cxv = cross(vx, gro), cyv = cross(vy, gro);
    // un-scale the N' coefficient gradients
    const float gn0 = kNScale * an0, gn1x = kNScale * an1, gn1y = kNScale * an2, gn2xx = kNScale * an3,
                gn2xy = kNScale * an4, gn2yy = kNScale * an5;
    const f3 v_c0 = (c0 * gn0 + cxv * gn1x + cyv * gn1y) * 2.f;
    const f3 v_cxv = (c0 * gn1x + cxv * gn2xx + cyv * gn2xy) * 2.f;
    const f3 v_cyv = (c0 * gn1y + cxv * gn2xy + cyv * gn2yy) * 2.f;
    f3 v_v0 = (v0 * ad0 + vx * ad1 + vy * ad2) * 2.f;
    f3 v_vx = (v0 * ad1 + vx * ad3 + vy * ad4) * 2.f;
    f3 v_vy = (v0 * ad2 + vx * ad4 + vy * ad5) * 2.f;
    // c = a x gro :  dL/da += gro x v_c ,  dL/dgro += v_c x a
    v_v0 = v_v0 + cross(gro, v_c0);
    v_vx = v_vx + cross(gro, v_cxv);
    v_vy = v_vy + cross(gro, v_cyv);
    const f3 v_gro = cross(v_c0, v0) + cross(v_cxv, vx) + cross(v_cyv, vy);
    // v0 = w2 + Xo vx + Yo vy
    const f3 v_w2 = v_v0;
    v_vx = v_vx + v_v0 * Xo;
    v_vy = v_vy + v_v0 * Yo;

    const float4 q = __ldg(reinterpret_cast<const float4*>(quats) + gid);
    const f3 sc = mk3(__ldg(scales + 3 * gid), __ldg(scales + 3 * gid + 1), __ldg(scales + 3 * gid + 2));
    const f3 mu = mk3(__ldg(means + 3 * gid), __ldg(means + 3 * gid + 1), __ldg(means + 3 * gid + 2));
    float R[9], inv_norm, qn[4];
    quat_to_rotmat_dev(q, R, inv_norm, qn);
    const float is[3] = {1.0f / sc.x, 1.0f / sc.y, 1.0f / sc.z};
    const float omu[3] = {cm.org[0] - mu.x, cm.org[1] - mu.y, cm.org[2] - mu.z};
    const float ifx = 1.0f / cm.fx, ify = 1.0f / cm.fy;
    const float vvx[3] = {v_vx.x * ifx, v_vx.y * ifx, v_vx.z * ifx};
    const float vvy[3] = {v_vy.x * ify, v_vy.y * ify, v_vy.z * ify};
    const float vw2[3] = {v_w2.x, v_w2.y, v_w2.z};
    const float vgro[3] = {v_gro.x, v_gro.y, v_gro.z};
    // v_M[a][i] (M = S^-1 R^T, M[a][i] = R[i][a] / s_a)
    float vM[3][3];
#pragma unroll
    for (int a_ = 0; a_ < 3; ++a_)
#pragma unroll
        for (int i_ = 0; i_ < 3; ++i_)
            vM[a_][i_] = vvx[a_] * cm.R[0 + i_] + vvy[a_] * cm.R[3 + i_] + vw2[a_] * cm.R[6 + i_] + vgro[a_] * omu[i_];
    float v_mean[3], v_scale[3], GR[3][3];
#pragma unroll
    for (int i_ = 0; i_ < 3; ++i_) {
        float s = 0.f;
#pragma unroll
        for (int a_ = 0; a_ < 3; ++a_)
            s += R[i_ * 3 + a_] * is[a_] * vgro[a_];
        v_mean[i_] = -s;
    }
#pragma unroll
    for (int a_ = 0; a_ < 3; ++a_) {
        float s = 0.f;
#pragma unroll
        for (int i_ = 0; i_ < 3; ++i_) {
            GR[i_][a_] = vM[a_][i_] * is[a_];
            s += R[i_ * 3 + a_] * vM[a_][i_];
        }
        v_scale[a_] = -is[a_] * is[a_] * s;
    }
    // quat_to_rotmat VJP (same algebra as the reference's Utils.cuh:104-126; G[r][c] = dL/dR[r][c])
    const float qw = qn[0], qx = qn[1], qy = qn[2], qz = qn[3];
    float vq[4];
    vq[0] = 2.f * (qx * (GR[2][1] - GR[1][2]) + qy * (GR[0][2] - GR[2][0]) + qz * (GR[1][0] - GR[0][1]));
    vq[1] = 2.f * (-2.f * qx * (GR[1][1] + GR[2][2]) + qy * (GR[1][0] + GR[0][1]) + qz * (GR[2][0] + GR[0][2]) +
                   qw * (GR[2][1] - GR[1][2]));
    vq[2] = 2.f * (qx * (GR[1][0] + GR[0][1]) - 2.f * qy * (GR[0][0] + GR[2][2]) + qz * (GR[2][1] + GR[1][2]) +
                   qw * (GR[0][2] - GR[2][0]));
    vq[3] = 2.f * (qx * (GR[2][0] + GR[0][2]) + qy * (GR[2][1] + GR[1][2]) - 2.f * qz * (GR[0][0] + GR[1][1]) +
                   qw * (GR[1][0] - GR[0][1]));
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

// EWA = fastgs surface: records are 2-D conics (D == 1), outputs are the per-primitive helpers of the reference's
// blend_backward_cu (kernels_backward.cuh:240-449) passed in the slots v_means -> grad_mean2d [N,2],
// v_quats -> grad_conic [N,3] (a, b, c; b is the TRUE derivative, twice the reference's stored value),
// v_colors -> grad_color [N,3], v_opacities -> grad_raw_opacity [N]; quats / scales / means / cams are unused.
// Lock-step variant (round 1's kernel, kept for A/B measurements: lfs_set_option("bwd_variant", 1)): alpha and chain of a
// step run back to back.
template <int kBwdWarps, bool EWA = false, bool PK = false>
__global__ void __launch_bounds__(kBwdWarps * 32)
    k_blend_bwd(const RasterBuffers rb, const ViewCam* __restrict__ cams, const float4* __restrict__ v_pix,
                const float* __restrict__ quats, const float* __restrict__ scales, const float* __restrict__ means,
                const uint32_t N, const uint32_t width, const uint32_t height, const uint32_t tile_w,
                const uint32_t tile_h, const uint32_t n_bucket_cap, const uint32_t* __restrict__ n_buckets_dev,
                float* __restrict__ v_means, float* __restrict__ v_quats, float* __restrict__ v_scales,
                float* __restrict__ v_colors, float* __restrict__ v_opacities) {
    // ring of per-pixel records the lanes read directly (no 9-register rotation): two float4 per pixel,
    //   ringA = (v_r, v_g, v_b, dx)   ringB = (dy, bits(n_rel), T0, u0)
    // n_rel = how many instances of THIS bucket the pixel consumed (lane < n_rel <=> the forward evaluated the pair);
    // T0 = transmittance at the bucket start; u0 = <colour accumulated from this bucket on, v_rgb> - v_alpha term.
    __shared__ float4 ringA[kBwdWarps][64];
    __shared__ float4 ringB[kBwdWarps][64];
    __shared__ uint8_t s_pix[kBwdWarps][kTilePix];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * kBwdWarps + warp;
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
    const uint32_t li = local_b * kBucket + lane; // tile-local instance index of this lane
    const int32_t inst = tstart + (int32_t)li;
    const bool valid = inst < tend;
    const float Xo = (float)(tx * kTile + kTile / 2) - (EWA ? 0.f : cams[cam].cx);
    const float Yo = (float)(ty * kTile + kTile / 2) - (EWA ? 0.f : cams[cam].cy);

    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A, Cc = A, E = A;
    uint32_t g = 0;
    if (valid) {
        g = (uint32_t)__ldg(rb.inst_gid + inst);
        if (EWA) {
            const float4* gp = reinterpret_cast<const float4*>(rb.gauss + g);
            expand_record_ewa(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), Xo, Yo, A, B, Cc);
            E = make_float4(B.z, B.w, Cc.x, Cc.y);
        } else {
            const float4* gp = reinterpret_cast<const float4*>(rb.gauss + g);
            expand_record(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), Xo, Yo, A, B, Cc);
            E = __ldg(gp + 3);
        }
    }

    // PK: (N', D) coefficient pairs and (dN', dD) accumulator pairs for the FFMA2 form of the loop body
    const float2 P0 = make_float2(A.x, B.z), P1 = make_float2(A.y, B.w), P2 = make_float2(A.z, Cc.x),
                 P3 = make_float2(A.w, Cc.y), P4 = make_float2(B.x, Cc.z), P5 = make_float2(B.y, Cc.w);
    float2 G0 = make_float2(0.f, 0.f), G1 = G0, G2 = G0, G3 = G0, G4 = G0, G5 = G0, GC = G0; // GC = (acr, acg)

    // state that still rotates through the lanes: transmittance before this lane's instance, and
    // u = <sum_{j >= this instance} c_j alpha_j T_j, v_rgb> - v_alpha_term   (one scalar instead of 3 + 1)
    float T = 0.f, u = 0.f;
    // per-instance accumulators
    float an0 = 0.f, an1 = 0.f, an2 = 0.f, an3 = 0.f, an4 = 0.f, an5 = 0.f;
    float ad0 = 0.f, ad1 = 0.f, ad2 = 0.f, ad3 = 0.f, ad4 = 0.f, ad5 = 0.f;
    float aop = 0.f, acr = 0.f, acg = 0.f, acb = 0.f;

    const float4* ck = rb.ckpt + (size_t)b * kTilePix;

    // compact list of the tile's pixels that reach this bucket (n_contrib > first instance of the bucket):
    // late buckets are reached by few pixels, so the rotation below runs m + 31 instead of 256 + 31 steps
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

    const int ilane = lane;
    for (int i = 0; i < m + 31; ++i) {
        if ((i & 31) == 0 && i < m) { // stage the next 32 pixels of the list into the ring half (i & 32)
            __syncwarp();
            float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a4;
            if (i + lane < m) {
                const int p = s_pix[warp][i + lane];
                const uint32_t px = tx * kTile + (p & 15), py = ty * kTile + (p >> 4);
                const size_t pix = ((size_t)cam * height + py) * width + px;
                const int32_t nrel = __ldg(rb.n_contrib + pix) - (int32_t)(local_b * kBucket);
                const float4 c4 = ld_nc4(ck + p);
                const float4 f4 = __ldg(rb.pix_state + pix);
                const float4 v4 = __ldg(v_pix + pix);
                const float u0 = fmaf(f4.x - c4.x, v4.x, fmaf(f4.y - c4.y, v4.y, fmaf(f4.z - c4.z, v4.z, -v4.w)));
                a4 = make_float4(v4.x, v4.y, v4.z, (float)(p & 15) - 7.5f);
                b4 = make_float4((float)(p >> 4) - 7.5f, __int_as_float(nrel), c4.w, u0);
            }
            ringA[warp][(i + lane) & 63] = a4;
            ringB[warp][(i + lane) & 63] = b4;
            __syncwarp();
        }
        const int idx = i - ilane; // position in the compact list of the pixel currently at this lane
        const float4 ea = ringA[warp][idx & 63], eb = ringB[warp][idx & 63];
        T = __shfl_up_sync(0xffffffffu, T, 1);
        u = __shfl_up_sync(0xffffffffu, u, 1);
        if (lane == 0) {
            T = eb.z;
            u = eb.w;
        }
        const bool active = valid && (uint32_t)idx < (uint32_t)m && ilane < __float_as_int(eb.y);
        if (!active)
            continue;
        const float dx = ea.w, dy = eb.x;
        float Nv, rD = 1.f, p2;
        float2 DX, DY;
        if (PK && !EWA) {
            DX = make_float2(dx, dx), DY = make_float2(dy, dy);
            const float2 ND = poly2x2(DX, DY, P0, P1, P2, P3, P4, P5);
            Nv = ND.x;
            rD = rcp_approx(ND.y);
            p2 = Nv * rD;
        } else {
            Nv = poly2(dx, dy, A.x, A.y, A.z, A.w, B.x, B.y);
            p2 = Nv;
            if (EWA) {
                if (Nv > 0.f) // sigma/2 < 0 is skipped (kernels_backward.cuh:391-392)
                    continue;
            } else {
                const float Dv = poly2(dx, dy, B.z, B.w, Cc.x, Cc.y, Cc.z, Cc.w);
                rD = rcp_approx(Dv);
                p2 = Nv * rD;
            }
        }
        const float vis = ex2_approx(p2);
        const float a_raw = E.x * vis;
        const float alpha = fminf(kAlphaMax, a_raw);
        if (alpha < kAlphaMin)
            continue;
        const float w = T * alpha;
        if (PK && !EWA) {
            GC = ffma2(make_float2(w, w), make_float2(ea.x, ea.y), GC);
        } else {
            acr = fmaf(w, ea.x, acr);
            acg = fmaf(w, ea.y, acg);
        }
        acb = fmaf(w, ea.z, acb);
        const float Ev = fmaf(E.y, ea.x, fmaf(E.z, ea.y, E.w * ea.z)); // <c_i, v_rgb>
        u = fmaf(-w, Ev, u);                                           // now the sum over j > i
        const float om = 1.0f - alpha;                                 // >= 0.001: plain rcp is safe
        const float v_alpha = fmaf(T, Ev, -u * rcp_approx(om));
        if (EWA) {
            // the reference differentiates through the clamped alpha as if it were not clamped (:418-427)
            aop = fmaf(alpha, v_alpha, aop);
            const float vN = v_alpha * alpha * kLn2;
            const float dxx = dx * dx, dxy = dx * dy, dyy = dy * dy;
            an0 += vN;
            an1 = fmaf(vN, dx, an1);
            an2 = fmaf(vN, dy, an2);
            an3 = fmaf(vN, dxx, an3);
            an4 = fmaf(vN, dxy, an4);
            an5 = fmaf(vN, dyy, an5);
        } else if (PK && a_raw <= kAlphaMax) {
            aop = fmaf(vis, v_alpha, aop);
            const float vN = v_alpha * a_raw * kLn2 * rD;
            const float2 V = make_float2(vN, -vN * p2);
            G0.x += V.x, G0.y += V.y;
            G1 = ffma2(V, DX, G1);
            G2 = ffma2(V, DY, G2);
            G3 = ffma2(V, fmul2(DX, DX), G3);
            G4 = ffma2(V, fmul2(DX, DY), G4);
            G5 = ffma2(V, fmul2(DY, DY), G5);
        } else if (a_raw <= kAlphaMax) {
            aop = fmaf(vis, v_alpha, aop);
            const float vN = v_alpha * a_raw * kLn2 * rD;
            const float vD = -vN * p2;
            const float dxx = dx * dx, dxy = dx * dy, dyy = dy * dy;
            an0 += vN;
            an1 = fmaf(vN, dx, an1);
            an2 = fmaf(vN, dy, an2);
            an3 = fmaf(vN, dxx, an3);
            an4 = fmaf(vN, dxy, an4);
            an5 = fmaf(vN, dyy, an5);
            ad0 += vD;
            ad1 = fmaf(vD, dx, ad1);
            ad2 = fmaf(vD, dy, ad2);
            ad3 = fmaf(vD, dxx, ad3);
            ad4 = fmaf(vD, dxy, ad4);
            ad5 = fmaf(vD, dyy, ad5);
        }
        T *= om;
    }

    if (PK && !EWA) {
        an0 = G0.x, ad0 = G0.y, an1 = G1.x, ad1 = G1.y, an2 = G2.x, ad2 = G2.y;
        an3 = G3.x, ad3 = G3.y, an4 = G4.x, ad4 = G4.y, an5 = G5.x, ad5 = G5.y;
        acr = GC.x, acg = GC.y;
    }
    if (!valid)
        return;
    if (EWA) {
        // ---- polynomial-coefficient gradients -> (conic, mean2d); e = mean2d - tile centre
        const float4* gp = reinterpret_cast<const float4*>(rb.gauss + g);
        const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
        const float ex = g0.x - Xo, ey = g0.y - Yo, a = g0.z, b = g0.w, c = g1.x;
        const float v_a = kEScale * (0.5f * ex * ex * an0 - ex * an1 + 0.5f * an3);
        const float v_b = kEScale * (ex * ey * an0 - ey * an1 - ex * an2 + an4);
        const float v_c = kEScale * (0.5f * ey * ey * an0 - ey * an2 + 0.5f * an5);
        const float v_ex = kEScale * ((a * ex + b * ey) * an0 - a * an1 - b * an2);
        const float v_ey = kEScale * ((c * ey + b * ex) * an0 - b * an1 - c * an2);
        atomicAdd(v_means + 2 * (size_t)g, v_ex);
        atomicAdd(v_means + 2 * (size_t)g + 1, v_ey);
        atomicAdd(v_quats + 3 * (size_t)g, v_a);
        atomicAdd(v_quats + 3 * (size_t)g + 1, v_b);
        atomicAdd(v_quats + 3 * (size_t)g + 2, v_c);
        // colour: gradient passes where the unclamped colour is >= 0 (color_grad_factor, :304-309)
        atomicAdd(v_colors + 3 * (size_t)g, g1.z >= 0.f ? acr : 0.f);
        atomicAdd(v_colors + 3 * (size_t)g + 1, g1.w >= 0.f ? acg : 0.f);
        atomicAdd(v_colors + 3 * (size_t)g + 2, g2.x >= 0.f ? acb : 0.f);
        atomicAdd(v_opacities + g, aop * (1.0f - g1.y)); // d alpha / d raw opacity = alpha (1 - opacity), :441
        return;
    }
    bwd_chain_rule(rb, cams[cam], g, N, Xo, Yo, an0, an1, an2, an3, an4, an5, ad0, ad1, ad2, ad3, ad4, ad5, acr, acg, acb, aop,
                   quats, scales, means, v_means, v_quats, v_scales, v_colors, v_opacities);
}

// ------------------------------------------------------------------------------------------------------
// backward blend, software-pipelined (the default): same algorithm, same arithmetic (bit-identical alphas), but
//   * the alpha of step i+1 (ring loads, FFMA2 polynomial chain, rcp, ex2) is issued BEFORE the transmittance / suffix
//     chain of step i: the two dependency chains of a step (about 100 cycles each) overlap inside one warp instead of
//     running back to back (measured before: 0.6 instructions per clock and scheduler at 4 warps each, 142 clocks per
//     step; profiles/r02_bucket_stats_c3.json),
//   * the loop is unrolled by two with alternating register sets, so the hand-over costs no moves,
//   * (N', D) are evaluated and accumulated as FFMA2 pairs.
// ------------------------------------------------------------------------------------------------------
struct BwdEval { // everything the chain / gradient half of a step needs, produced one step ahead
    float vr, vg, vb, dx, dy; // pixel data
    float T0, u0;             // lane 0 only: state entering the bucket
    float a_raw, vis, p2, rD; // opacity * vis (unclamped), vis, N'/D, 1/D
    bool pass;                // the forward evaluated this pair and alpha >= 1/255
};

// one bucket, one warp; ringA / ringB / s_pix are this warp's shared-memory slices
template <bool EWA>
__device__ __forceinline__ void bwd_sp_bucket(const uint32_t b, const int lane, float4* __restrict__ ringA_w,
                                              float4* __restrict__ ringB_w, uint8_t* __restrict__ s_pix_w,
                                              const RasterBuffers& rb, const ViewCam* __restrict__ cams,
                                              const float4* __restrict__ v_pix, const float* __restrict__ quats,
                                              const float* __restrict__ scales, const float* __restrict__ means,
                                              const uint32_t N, const uint32_t width, const uint32_t height,
                                              const uint32_t tile_w, const uint32_t tile_h, float* __restrict__ v_means,
                                              float* __restrict__ v_quats, float* __restrict__ v_scales,
                                              float* __restrict__ v_colors, float* __restrict__ v_opacities) {
    const uint32_t n_tiles = tile_w * tile_h;
    const uint32_t ft = rb.bucket_tile[b];
    const uint32_t cam = ft / n_tiles, tile = ft - cam * n_tiles;
    const uint32_t ty = tile / tile_w, tx = tile - ty * tile_w;
    const int32_t tstart = rb.tile_off[ft], tend = rb.tile_off[ft + 1];
    const uint32_t local_b = b - rb.bucket_off[ft];
    if (local_b * kBucket >= rb.tile_max_contrib[ft])
        return;
    const uint32_t li = local_b * kBucket + lane;
    const int32_t inst = tstart + (int32_t)li;
    const bool valid = inst < tend;
    const float Xo = (float)(tx * kTile + kTile / 2) - (EWA ? 0.f : cams[cam].cx);
    const float Yo = (float)(ty * kTile + kTile / 2) - (EWA ? 0.f : cams[cam].cy);

    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A, Cc = A, E = A;
    uint32_t g = 0;
    if (valid) {
        g = (uint32_t)__ldg(rb.inst_gid + inst);
        const float4* gp = reinterpret_cast<const float4*>(rb.gauss + g);
        if (EWA) {
            expand_record_ewa(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), Xo, Yo, A, B, Cc);
            E = make_float4(B.z, B.w, Cc.x, Cc.y);
        } else {
            expand_record(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), Xo, Yo, A, B, Cc);
            E = __ldg(gp + 3);
        }
    }
    // (N', D) coefficient pairs; EWA: D == 1 (second components unused)
    const float2 P0 = make_float2(A.x, EWA ? 1.f : B.z), P1 = make_float2(A.y, EWA ? 0.f : B.w),
                 P2 = make_float2(A.z, EWA ? 0.f : Cc.x), P3 = make_float2(A.w, EWA ? 0.f : Cc.y),
                 P4 = make_float2(B.x, EWA ? 0.f : Cc.z), P5 = make_float2(B.y, EWA ? 0.f : Cc.w);
    float2 G0 = make_float2(0.f, 0.f), G1 = G0, G2 = G0, G3 = G0, G4 = G0, G5 = G0, GC = G0; // (dN', dD) and (r, g)
    float acb = 0.f, aop = 0.f;
    float T = 0.f, u = 0.f;
    {   // lanes ahead of / behind the pixel list read ring slots no stage() has written yet: they run with zero weights,
        // but 0 * NaN is NaN, so the slots must hold finite numbers (shared memory is not cleared between kernels)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        ringA_w[lane] = ringA_w[lane + 32] = z4;
        ringB_w[lane] = ringB_w[lane + 32] = z4;
    }

    const float4* ck = rb.ckpt + (size_t)b * kTilePix;
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
                s_pix_w[m + __popc(mask & lt)] = (uint8_t)p;
            m += __popc(mask);
        }
        __syncwarp();
    }
    // stage list positions [i0, i0 + 32) into the ring
    auto stage = [&](const int i0) {
        __syncwarp(); // every lane is done reading the entries that are about to be replaced
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a4;
        if (i0 + lane < m) {
            const int p = s_pix_w[i0 + lane];
            const uint32_t px = tx * kTile + (p & 15), py = ty * kTile + (p >> 4);
            const size_t pix = ((size_t)cam * height + py) * width + px;
            const int32_t nrel = __ldg(rb.n_contrib + pix) - (int32_t)(local_b * kBucket);
            const float4 c4 = ld_nc4(ck + p);
            const float4 f4 = __ldg(rb.pix_state + pix);
            const float4 v4 = __ldg(v_pix + pix);
            const float u0 = fmaf(f4.x - c4.x, v4.x, fmaf(f4.y - c4.y, v4.y, fmaf(f4.z - c4.z, v4.z, -v4.w)));
            a4 = make_float4(v4.x, v4.y, v4.z, (float)(p & 15) - 7.5f);
            b4 = make_float4((float)(p >> 4) - 7.5f, __int_as_float(nrel), c4.w, u0);
        }
        ringA_w[(i0 + lane) & 63] = a4;
        ringB_w[(i0 + lane) & 63] = b4;
        __syncwarp();
    };
    // the alpha half of step i for this lane (list position i - lane)
    auto eval = [&](const int i, BwdEval& e) {
        const int idx = i - lane;
        const float4 ea = ringA_w[idx & 63], eb = ringB_w[idx & 63];
        e.vr = ea.x, e.vg = ea.y, e.vb = ea.z, e.dx = ea.w, e.dy = eb.x, e.T0 = eb.z, e.u0 = eb.w;
        const bool act = valid && (uint32_t)idx < (uint32_t)m && lane < __float_as_int(eb.y);
        const float2 ND = poly2x2(make_float2(e.dx, e.dx), make_float2(e.dy, e.dy), P0, P1, P2, P3, P4, P5);
        bool ok = act;
        if (EWA) {
            e.rD = 1.f;
            e.p2 = ND.x;
            ok = ok && ND.x <= 0.f; // sigma/2 < 0 is skipped (kernels_backward.cuh:391-392)
        } else {
            e.rD = rcp_approx(ND.y);
            e.p2 = ND.x * e.rD;
        }
        e.vis = ex2_approx(e.p2);
        e.a_raw = E.x * e.vis;
        e.pass = ok && fminf(kAlphaMax, e.a_raw) >= kAlphaMin;
    };
    // the chain + gradient half.  Branch-free: a pair that does not contribute runs with alpha = 0 and zero gradient
    // weights (selects, not multiplications: its vis / p2 may be garbage), so the whole step is ONE basic block and the
    // scheduler can interleave this half with the alpha half of the next step.
    auto chain = [&](const BwdEval& e) {
        T = __shfl_up_sync(0xffffffffu, T, 1);
        u = __shfl_up_sync(0xffffffffu, u, 1);
        if (lane == 0) {
            T = e.T0;
            u = e.u0;
        }
        const float alpha = e.pass ? fminf(kAlphaMax, e.a_raw) : 0.f;
        const float w = T * alpha;
        GC = ffma2(make_float2(w, w), make_float2(e.vr, e.vg), GC);
        acb = fmaf(w, e.vb, acb);
        const float Ev = fmaf(E.y, e.vr, fmaf(E.z, e.vg, E.w * e.vb)); // <c_i, v_rgb>
        u = fmaf(-w, Ev, u);                                           // now the sum over j > i
        const float om = 1.0f - alpha;                                 // >= 0.001: plain rcp is safe
        const float v_alpha = fmaf(T, Ev, -u * rcp_approx(om));
        T *= om;
        // EWA: the reference differentiates through the clamped alpha as if it were not clamped (:418-427);
        // from-world: no geometry / opacity gradient through a clamped alpha (...Bwd.cu:318)
        const bool grad = EWA ? e.pass : (e.pass && e.a_raw <= kAlphaMax);
        const float vo = (EWA ? alpha : e.vis) * v_alpha;
        aop += grad ? vo : 0.f;
        // (the factor ln 2 of d 2^x / dx is applied once per instance, after the loop)
        const float vN = EWA ? v_alpha * alpha : v_alpha * e.a_raw * e.rD;
        const float2 V = grad ? make_float2(vN, -vN * e.p2) : make_float2(0.f, 0.f);
        const float2 DX = make_float2(e.dx, e.dx), DY = make_float2(e.dy, e.dy);
        G0 = ffma2(V, make_float2(1.f, 1.f), G0);
        G1 = ffma2(V, DX, G1);
        G2 = ffma2(V, DY, G2);
        G3 = ffma2(V, fmul2(DX, DX), G3);
        G4 = ffma2(V, fmul2(DX, DY), G4);
        G5 = ffma2(V, fmul2(DY, DY), G5);
    };

    const int total = m + 31;
    BwdEval ea_, eb_;
    stage(0);
    eval(0, ea_);
    for (int i = 0; i < total; i += 2) {
        // ---- step i: prefetch i+1 into eb_, then chain on ea_
        if (((i + 1) & 31) == 0 && i + 1 < m)
            stage(i + 1);
        eval(i + 1, eb_);
        chain(ea_);
        if (i + 1 >= total)
            break;
        // ---- step i+1: prefetch i+2 into ea_, then chain on eb_
        if (((i + 2) & 31) == 0 && i + 2 < m)
            stage(i + 2);
        eval(i + 2, ea_);
        chain(eb_);
    }

    if (!valid)
        return;
    const float an0 = kLn2 * G0.x, an1 = kLn2 * G1.x, an2 = kLn2 * G2.x, an3 = kLn2 * G3.x, an4 = kLn2 * G4.x, an5 = kLn2 * G5.x;
    const float ad0 = kLn2 * G0.y, ad1 = kLn2 * G1.y, ad2 = kLn2 * G2.y, ad3 = kLn2 * G3.y, ad4 = kLn2 * G4.y, ad5 = kLn2 * G5.y;
    const float acr = GC.x, acg = GC.y;
    if (EWA) {
        // ---- polynomial-coefficient gradients -> (conic, mean2d); e = mean2d - tile centre
        const float4* gp = reinterpret_cast<const float4*>(rb.gauss + g);
        const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
        const float ex = g0.x - Xo, ey = g0.y - Yo, a = g0.z, b_ = g0.w, c = g1.x;
        const float v_a = kEScale * (0.5f * ex * ex * an0 - ex * an1 + 0.5f * an3);
        const float v_b = kEScale * (ex * ey * an0 - ey * an1 - ex * an2 + an4);
        const float v_c = kEScale * (0.5f * ey * ey * an0 - ey * an2 + 0.5f * an5);
        const float v_ex = kEScale * ((a * ex + b_ * ey) * an0 - a * an1 - b_ * an2);
        const float v_ey = kEScale * ((c * ey + b_ * ex) * an0 - b_ * an1 - c * an2);
        atomicAdd(v_means + 2 * (size_t)g, v_ex);
        atomicAdd(v_means + 2 * (size_t)g + 1, v_ey);
        atomicAdd(v_quats + 3 * (size_t)g, v_a);
        atomicAdd(v_quats + 3 * (size_t)g + 1, v_b);
        atomicAdd(v_quats + 3 * (size_t)g + 2, v_c);
        atomicAdd(v_colors + 3 * (size_t)g, g1.z >= 0.f ? acr : 0.f);
        atomicAdd(v_colors + 3 * (size_t)g + 1, g1.w >= 0.f ? acg : 0.f);
        atomicAdd(v_colors + 3 * (size_t)g + 2, g2.x >= 0.f ? acb : 0.f);
        atomicAdd(v_opacities + g, aop * (1.0f - g1.y));
        return;
    }
    bwd_chain_rule(rb, cams[cam], g, N, Xo, Yo, an0, an1, an2, an3, an4, an5, ad0, ad1, ad2, ad3, ad4, ad5, acr, acg, acb, aop,
                   quats, scales, means, v_means, v_quats, v_scales, v_colors, v_opacities);
}


// Driver.  With a live-bucket list (rb.live: written by the forward, [0] = count, [1] = work counter, [2..] = bucket
// ids) the grid is persistent -- kBwdBlocksPerSM CTAs per SM, every warp pulls the next live bucket with one atomic --
// so that the ~86 % of the buckets no pixel reaches (C3: 32 k live of 232 k) cost nothing: with one warp per bucket they
// still cost a CTA slot and three dependent loads each, and a CTA with one live and three dead buckets held the registers
// of four warps (measured: 13.5 of the 20 possible warps per SM active).  Without the list: one warp per bucket.
// kMinBlocks = CTAs per SM the register allocation is bounded for (5: 96 registers, 6: 80 registers, neither spills) and the
// persistent grid is sized with.
template <bool EWA, int kBwdWarps, int kMinBlocks>
__global__ void __launch_bounds__(kBwdWarps * 32, kMinBlocks)
    k_blend_bwd_sp(const RasterBuffers rb, const ViewCam* __restrict__ cams, const float4* __restrict__ v_pix,
                   const float* __restrict__ quats, const float* __restrict__ scales, const float* __restrict__ means,
                   const uint32_t N, const uint32_t width, const uint32_t height, const uint32_t tile_w,
                   const uint32_t tile_h, const uint32_t n_bucket_cap, const uint32_t* __restrict__ n_buckets_dev,
                   float* __restrict__ v_means, float* __restrict__ v_quats, float* __restrict__ v_scales,
                   float* __restrict__ v_colors, float* __restrict__ v_opacities) {
    __shared__ float4 ringA[kBwdWarps][64]; // (v_r, v_g, v_b, dx)
    __shared__ float4 ringB[kBwdWarps][64]; // (dy, bits(n_rel), T0, u0)
    __shared__ uint8_t s_pix[kBwdWarps][kTilePix];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t nbk = *n_buckets_dev;
    nbk = nbk < n_bucket_cap ? nbk : n_bucket_cap;
    if (rb.live == nullptr) {
        const uint32_t b = blockIdx.x * kBwdWarps + warp;
        if (b < nbk)
            bwd_sp_bucket<EWA>(b, lane, ringA[warp], ringB[warp], s_pix[warp], rb, cams, v_pix, quats, scales, means, N,
                               width, height, tile_w, tile_h, v_means, v_quats, v_scales, v_colors, v_opacities);
        return;
    }
    const uint32_t n_live = min(rb.live[0], nbk);
    for (;;) {
        uint32_t w = 0;
        if (lane == 0)
            w = atomicAdd(rb.live + 1, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_live)
            return;
        const uint32_t b = rb.live[2 + w];
        if (b < nbk)
            bwd_sp_bucket<EWA>(b, lane, ringA[warp], ringB[warp], s_pix[warp], rb, cams, v_pix, quats, scales, means, N,
                               width, height, tile_w, tile_h, v_means, v_quats, v_scales, v_colors, v_opacities);
        __syncwarp();
    }
}

static unsigned bwd_sp_grid(const RasterBuffers& rb, const uint32_t n_bucket_cap, const int blocks_per_sm) {
    const unsigned dense = div_up(n_bucket_cap, 4);
    if (rb.live == nullptr)
        return dense;
    const unsigned persistent = (unsigned)(num_sms() * blocks_per_sm);
    return persistent < dense ? persistent : dense;
}

int launch_blend_bwd(const RasterBuffers& rb, const ViewCam* cams_dev, const float4* v_pix, const float* quats,
                     const float* scales, const float* means, uint32_t C, uint32_t N, uint32_t width, uint32_t height,
                     uint32_t tile_w, uint32_t tile_h, uint32_t n_bucket_cap, const uint32_t* n_buckets_dev,
                     float* v_means, float* v_quats, float* v_scales, float* v_colors, float* v_opacities,
                     cudaStream_t stream) {
    (void)C;
    if (n_bucket_cap == 0)
        return LFS_OK;
    if (raster_options().bwd_variant == 1)
        k_blend_bwd<4, false, true><<<div_up(n_bucket_cap, 4), 4 * 32, 0, stream>>>(
            rb, cams_dev, v_pix, quats, scales, means, N, width, height, tile_w, tile_h, n_bucket_cap, n_buckets_dev,
            v_means, v_quats, v_scales, v_colors, v_opacities);
    else {
        RasterBuffers rbl = rb;
        if (raster_options().bwd_variant == 2) // A/B: one warp per bucket, no live list
            rbl.live = nullptr;
        if (rbl.live) // work counter of the persistent grid (a forward may be followed by more than one backward)
            LFS_CUDA_OK(cudaMemsetAsync(rbl.live + 1, 0, sizeof(uint32_t), stream));
        if (raster_options().bwd_variant == 3) // A/B: 6 CTAs per SM at 80 registers
            k_blend_bwd_sp<false, 4, 6><<<bwd_sp_grid(rbl, n_bucket_cap, 6), 4 * 32, 0, stream>>>(
            rbl, cams_dev, v_pix, quats, scales, means, N, width, height, tile_w, tile_h, n_bucket_cap, n_buckets_dev,
            v_means, v_quats, v_scales, v_colors, v_opacities);
        else
            k_blend_bwd_sp<false, 4, 5><<<bwd_sp_grid(rbl, n_bucket_cap, 5), 4 * 32, 0, stream>>>(
            rbl, cams_dev, v_pix, quats, scales, means, N, width, height, tile_w, tile_h, n_bucket_cap, n_buckets_dev,
            v_means, v_quats, v_scales, v_colors, v_opacities);
    }
    LFS_LAUNCH_OK("k_blend_bwd");
    return LFS_OK;
}

int launch_blend_bwd_ewa(const RasterBuffers& rb, const float4* v_pix, uint32_t N, uint32_t width, uint32_t height,
                         uint32_t tile_w, uint32_t tile_h, uint32_t n_bucket_cap, const uint32_t* n_buckets_dev,
                         float* v_mean2d, float* v_conic, float* v_color, float* v_raw_opacity, cudaStream_t stream) {
    if (n_bucket_cap == 0)
        return LFS_OK;
    if (raster_options().bwd_variant == 1)
        k_blend_bwd<4, true><<<div_up(n_bucket_cap, 4), 4 * 32, 0, stream>>>(
            rb, nullptr, v_pix, nullptr, nullptr, nullptr, N, width, height, tile_w, tile_h, n_bucket_cap, n_buckets_dev,
            v_mean2d, v_conic, nullptr, v_color, v_raw_opacity);
    else {
        RasterBuffers rbl = rb;
        if (raster_options().bwd_variant == 2) // A/B: one warp per bucket, no live list
            rbl.live = nullptr;
        if (rbl.live) // work counter of the persistent grid (a forward may be followed by more than one backward)
            LFS_CUDA_OK(cudaMemsetAsync(rbl.live + 1, 0, sizeof(uint32_t), stream));
        if (raster_options().bwd_variant == 3) // A/B: 6 CTAs per SM at 80 registers
            k_blend_bwd_sp<true, 4, 6><<<bwd_sp_grid(rbl, n_bucket_cap, 6), 4 * 32, 0, stream>>>(
            rbl, nullptr, v_pix, nullptr, nullptr, nullptr, N, width, height, tile_w, tile_h, n_bucket_cap, n_buckets_dev,
            v_mean2d, v_conic, nullptr, v_color, v_raw_opacity);
        else
            k_blend_bwd_sp<true, 4, 5><<<bwd_sp_grid(rbl, n_bucket_cap, 5), 4 * 32, 0, stream>>>(
            rbl, nullptr, v_pix, nullptr, nullptr, nullptr, N, width, height, tile_w, tile_h, n_bucket_cap, n_buckets_dev,
            v_mean2d, v_conic, nullptr, v_color, v_raw_opacity);
    }
    LFS_LAUNCH_OK("k_blend_bwd<ewa>");
    return LFS_OK;
}

} // namespace lfs
