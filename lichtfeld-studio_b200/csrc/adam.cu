// lfs_b200 -- fused Adam.  lfs_adam_step is the drop-in for fast_gs::optimizer::adam_step
// (reference fastgs/optimizer/src/adam.cu:11-35, kernel adam_kernels.cuh:13-36); lfs_adam_step_multi updates the
// whole flat parameter arena in ONE launch (the reference launches once per parameter tensor, 6 per step,
// src/training/optimizers/fused_adam.cpp:41-94) and can clear the gradient arena in the same pass.
//
// Pure HBM stream: 16 B read + 12 B written per element (28 B/element; +4 B when the gradient is cleared).
// 128-bit loads/stores, two independent float4 per thread in flight, grid = multiple of 148 SMs.
#include "common.cuh"

namespace lfs {

constexpr int kAdamThreads = 256;
constexpr int kAdamMaxSeg = 16;

struct AdamSegs {
    int n;
    int64_t begin4[kAdamMaxSeg + 1]; // segment boundaries in float4 units
    float step_size[kAdamMaxSeg];    // lr * bias_correction1_rcp
    float bc2[kAdamMaxSeg];          // bias_correction2_sqrt_rcp
    // optional regulariser gradient added before the update (lfs_adam_reg): kind 0 none, 1 coef*exp(p), 2 coef*s(p)(1-s(p))
    int reg_kind[kAdamMaxSeg];
    float reg_coef[kAdamMaxSeg];
    int64_t np4;     // plane length in float4 units (planar arenas: only the first n_valid elements of a plane are real)
    int64_t n_valid;
    int any_reg;
};

__device__ __forceinline__ float reg_grad1(const int kind, const float coef, const float p) {
    if (kind == 1)
        return coef * __expf(p);
    const float sg = 1.0f / (1.0f + __expf(-p));
    return coef * sg * (1.0f - sg);
}
// gradient of the folded regulariser for the float4 at segment-relative index i_rel4 of segment s
__device__ __forceinline__ void add_reg4(const AdamSegs& segs, const int s, const int64_t i_rel4, const float4 p, float4& g) {
    const int kind = segs.reg_kind[s];
    if (kind == 0)
        return;
    const float coef = segs.reg_coef[s];
    const int64_t e = 4 * ((i_rel4 - segs.begin4[s]) % segs.np4); // element index inside its plane
    if (e < segs.n_valid) g.x += reg_grad1(kind, coef, p.x);
    if (e + 1 < segs.n_valid) g.y += reg_grad1(kind, coef, p.y);
    if (e + 2 < segs.n_valid) g.z += reg_grad1(kind, coef, p.z);
    if (e + 3 < segs.n_valid) g.w += reg_grad1(kind, coef, p.w);
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, const float g, const float beta1,
                                            const float beta2, const float eps, const float step_size,
                                            const float bc2) {
    const float m1 = beta1 * m + (1.0f - beta1) * g;
    const float m2 = beta2 * v + (1.0f - beta2) * g * g;
    const float denom = sqrtf(m2) * bc2 + eps;
    p -= step_size * m1 / denom;
    m = m1;
    v = m2;
}

__device__ __forceinline__ void adam_update4(float4& p, float4& m, float4& v, const float4 g, const float beta1,
                                             const float beta2, const float eps, const float step_size,
                                             const float bc2) {
    adam_update(p.x, m.x, v.x, g.x, beta1, beta2, eps, step_size, bc2);
    adam_update(p.y, m.y, v.y, g.y, beta1, beta2, eps, step_size, bc2);
    adam_update(p.z, m.z, v.z, g.z, beta1, beta2, eps, step_size, bc2);
    adam_update(p.w, m.w, v.w, g.w, beta1, beta2, eps, step_size, bc2);
}

__global__ void __launch_bounds__(kAdamThreads)
    k_adam_multi(float* __restrict__ params, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                 float* __restrict__ grads, const AdamSegs segs, const float beta1, const float beta2,
                 const float eps, const int zero_grad) {
    const int64_t n4 = segs.begin4[segs.n];
    float4* p4 = reinterpret_cast<float4*>(params);
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    float4* g4 = reinterpret_cast<float4*>(grads);
    const int64_t stride = (int64_t)gridDim.x * kAdamThreads;
    for (int64_t i = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; i < n4; i += 2 * stride) {
        const int64_t j = i + stride;
        const bool has_j = j < n4;
        // issue all loads first (two float4 x 4 arrays in flight per thread)
        float4 pa = p4[i], ma = m4[i], va = v4[i], ga = g4[i];
        float4 pb, mb, vb, gb;
        if (has_j) {
            pb = p4[j], mb = m4[j], vb = v4[j], gb = g4[j];
        }
        int sa = 0, sb = 0;
#pragma unroll
        for (int s = 1; s < kAdamMaxSeg; ++s) {
            if (s < segs.n) {
                sa += (i >= segs.begin4[s]);
                sb += (j >= segs.begin4[s]);
            }
        }
        if (segs.any_reg)
            add_reg4(segs, sa, i, pa, ga);
        adam_update4(pa, ma, va, ga, beta1, beta2, eps, segs.step_size[sa], segs.bc2[sa]);
        p4[i] = pa, m4[i] = ma, v4[i] = va;
        if (zero_grad)
            g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_j) {
            if (segs.any_reg)
                add_reg4(segs, sb, j, pb, gb);
            adam_update4(pb, mb, vb, gb, beta1, beta2, eps, segs.step_size[sb], segs.bc2[sb]);
            p4[j] = pb, m4[j] = mb, v4[j] = vb;
            if (zero_grad)
                g4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// Multi-GPU step: reduce-scatter + Adam + all-gather in ONE kernel over NVLink peer memory.
// OWNERSHIP is a property of the arena, not of the call: float4 index i (absolute, counted from the start of the arena)
// belongs to rank (i / kOwnChunk4) % world.  A rank keeps Adam moments only for the chunks it owns, so the owner of an
// element must never change -- with this rule it does not depend on which segments a call covers (the shN group joins
// after iteration 1000, fused_adam.cpp:69) and any sub-range of the arena is spread evenly over the ranks.
// The owner sums its chunk of every rank's gradient arena with peer loads in a fixed rank order (every element is reduced
// by exactly one rank, deterministically), applies Adam to its local p / m / v, then stores the new parameters into every
// rank's parameter arena.  Compared with ncclAllReduce(236 MB) + a full Adam on every rank this moves the same bytes over
// NVLink once, does 1/W of the Adam traffic per GPU and needs no separate collective launch.  The caller brackets it with
// two stream-ordered barriers (all backward passes done before; all parameter writes landed after).
constexpr int64_t kOwnChunk4 = 1024; // 16 KB of float4 per ownership chunk
constexpr int kMaxWorld = 16;

struct PeerPtrs {
    float* const* grads;  // device array [world]: peer-mapped gradient arenas
    float* const* params; // device array [world]: peer-mapped parameter arenas
    int world, rank;
};

struct OwnedRange { // the chunks of [lo4, hi4) (absolute float4 indices) owned by `rank`
    int64_t lo4, hi4, c0, n_owned;
    int world;
};
static inline OwnedRange owned_range(int64_t lo4, int64_t hi4, int world, int rank) {
    OwnedRange o{lo4, hi4, 0, 0, world};
    if (hi4 <= lo4)
        return o;
    const int64_t c_first = lo4 / kOwnChunk4, c_last = (hi4 - 1) / kOwnChunk4;
    o.c0 = c_first + (((int64_t)rank - c_first % world) % world + world) % world;
    o.n_owned = o.c0 > c_last ? 0 : (c_last - o.c0) / world + 1;
    return o;
}

__device__ __forceinline__ int seg_of(const AdamSegs& segs, const int64_t i_rel4) {
    int s = 0;
#pragma unroll
    for (int k = 1; k < kAdamMaxSeg; ++k)
        if (k < segs.n)
            s += (i_rel4 >= segs.begin4[k]);
    return s;
}

// W > 0: world size known at compile time, all W peer loads of an element are in flight together; W == 0: generic.
template <int W>
__global__ void __launch_bounds__(kAdamThreads)
    k_adam_multi_p2p(const PeerPtrs pp, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                     const int64_t base4, const AdamSegs segs, const OwnedRange own, const float beta1,
                     const float beta2, const float eps) {
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    float4* p4 = reinterpret_cast<float4*>(pp.params[pp.rank]);
    const int world = W > 0 ? W : pp.world;
    const float4* gsrc[W > 0 ? W : kMaxWorld];
    float4* pdst[W > 0 ? W : kMaxWorld];
#pragma unroll
    for (int r = 0; r < (W > 0 ? W : kMaxWorld); ++r)
        if (r < world) {
            gsrc[r] = reinterpret_cast<const float4*>(pp.grads[r]);
            pdst[r] = reinterpret_cast<float4*>(pp.params[r]);
        }
    for (int64_t k = blockIdx.x; k < own.n_owned; k += gridDim.x) {
        const int64_t c = own.c0 + k * own.world;
        const int64_t lo = max(c * kOwnChunk4, own.lo4), hi = min((c + 1) * kOwnChunk4, own.hi4);
        for (int64_t i = lo + threadIdx.x; i < hi; i += kAdamThreads) {
            float4 x[W > 0 ? W : kMaxWorld];
#pragma unroll
            for (int r = 0; r < (W > 0 ? W : kMaxWorld); ++r)
                if (r < world)
                    x[r] = __ldcv(gsrc[r] + i); // never a stale L1 line
            float4 p = p4[i], m = m4[i], v = v4[i];
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < (W > 0 ? W : kMaxWorld); ++r)
                if (r < world)
                    g.x += x[r].x, g.y += x[r].y, g.z += x[r].z, g.w += x[r].w;
            const int s = seg_of(segs, i - base4);
            if (segs.any_reg)
                add_reg4(segs, s, i - base4, p, g);
            adam_update4(p, m, v, g, beta1, beta2, eps, segs.step_size[s], segs.bc2[s]);
            m4[i] = m, v4[i] = v;
#pragma unroll
            for (int r = 0; r < (W > 0 ? W : kMaxWorld); ++r)
                if (r < world)
                    pdst[r][i] = p;
        }
    }
}

// Same step through the NVSwitch multicast object (NVLS): ONE multimem.ld_reduce returns the sum of the element over all
// ranks, reduced inside the switch (inbound traffic 1x instead of (W-1)x), ONE multimem.st broadcasts the new parameters.
__global__ void __launch_bounds__(kAdamThreads)
    k_adam_multi_mc(float* __restrict__ params_local, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                    const float* __restrict__ grads_mc, float* __restrict__ params_mc, const int64_t base4,
                    const AdamSegs segs, const OwnedRange own, const float beta1, const float beta2, const float eps) {
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    const float4* p4 = reinterpret_cast<const float4*>(params_local);
    for (int64_t k = blockIdx.x; k < own.n_owned; k += gridDim.x) {
        const int64_t c = own.c0 + k * own.world;
        const int64_t lo = max(c * kOwnChunk4, own.lo4), hi = min((c + 1) * kOwnChunk4, own.hi4);
        for (int64_t i = lo + threadIdx.x; i < hi; i += kAdamThreads) {
            float4 g;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w)
                         : "l"(reinterpret_cast<const float4*>(grads_mc) + i)
                         : "memory");
            float4 p = p4[i], m = m4[i], v = v4[i];
            const int s = seg_of(segs, i - base4);
            if (segs.any_reg)
                add_reg4(segs, s, i - base4, p, g);
            adam_update4(p, m, v, g, beta1, beta2, eps, segs.step_size[s], segs.bc2[s]);
            m4[i] = m, v4[i] = v;
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                             reinterpret_cast<float4*>(params_mc) + i),
                         "f"(p.x), "f"(p.y), "f"(p.z), "f"(p.w)
                         : "memory");
        }
    }
}

// single-tensor, arbitrary n / alignment: vector body + scalar head/tail
__global__ void __launch_bounds__(kAdamThreads)
    k_adam_single(float* __restrict__ param, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                  const float* __restrict__ grad, const int64_t n, const float step_size, const float beta1,
                  const float beta2, const float eps, const float bc2, const int vec_ok) {
    const int64_t stride = (int64_t)gridDim.x * kAdamThreads;
    const int64_t tid = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x;
    if (vec_ok) {
        const int64_t n4 = n >> 2;
        float4* p4 = reinterpret_cast<float4*>(param);
        float4* m4 = reinterpret_cast<float4*>(exp_avg);
        float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
        const float4* g4 = reinterpret_cast<const float4*>(grad);
        for (int64_t i = tid; i < n4; i += stride) {
            float4 p = p4[i], m = m4[i], v = v4[i];
            const float4 g = g4[i];
            adam_update4(p, m, v, g, beta1, beta2, eps, step_size, bc2);
            p4[i] = p, m4[i] = m, v4[i] = v;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride)
            adam_update(param[i], exp_avg[i], exp_avg_sq[i], grad[i], beta1, beta2, eps, step_size, bc2);
    } else {
        for (int64_t i = tid; i < n; i += stride)
            adam_update(param[i], exp_avg[i], exp_avg_sq[i], grad[i], beta1, beta2, eps, step_size, bc2);
    }
}

// fills begin4 / step sizes / regulariser fields from the host arrays; returns an error code
static int build_segs(AdamSegs& segs, const char* who, int n_segments, const int64_t* seg_begin_host, const float* lr_host,
                      const float* bc1_rcp_host, const float* bc2_sqrt_rcp_host, const lfs_adam_reg* reg) {
    LFS_CHECK_ARG(n_segments >= 1 && n_segments <= kAdamMaxSeg, "%s: n_segments=%d out of range", who, n_segments);
    LFS_CHECK_ARG(seg_begin_host && lr_host && bc1_rcp_host && bc2_sqrt_rcp_host, "%s: null pointer", who);
    segs.n = n_segments;
    for (int s = 0; s <= n_segments; ++s) {
        LFS_CHECK_ARG((seg_begin_host[s] & 3) == 0, "%s: segment boundary %d not a multiple of 4", who, s);
        LFS_CHECK_ARG(s == 0 || seg_begin_host[s] >= seg_begin_host[s - 1], "%s: unsorted segments", who);
        segs.begin4[s] = (seg_begin_host[s] - seg_begin_host[0]) / 4;
    }
    segs.any_reg = 0;
    segs.np4 = 1, segs.n_valid = 0;
    for (int s = 0; s < kAdamMaxSeg; ++s)
        segs.reg_kind[s] = 0, segs.reg_coef[s] = 0.f;
    for (int s = 0; s < n_segments; ++s) {
        segs.step_size[s] = lr_host[s] * bc1_rcp_host[s];
        segs.bc2[s] = bc2_sqrt_rcp_host[s];
        if (reg && reg->kind[s] != 0 && reg->coef[s] != 0.f) {
            LFS_CHECK_ARG(reg->kind[s] == 1 || reg->kind[s] == 2, "%s: unknown regulariser kind %d", who, reg->kind[s]);
            segs.reg_kind[s] = reg->kind[s], segs.reg_coef[s] = reg->coef[s], segs.any_reg = 1;
        }
    }
    if (segs.any_reg) {
        LFS_CHECK_ARG(reg->plane_elems > 0 && (reg->plane_elems & 3) == 0 && reg->n_valid >= 0 &&
                          reg->n_valid <= reg->plane_elems,
                      "%s: regulariser needs plane_elems (multiple of 4) and n_valid <= plane_elems", who);
        for (int s = 0; s < n_segments; ++s)
            LFS_CHECK_ARG(segs.reg_kind[s] == 0 || ((segs.begin4[s + 1] - segs.begin4[s]) * 4) % reg->plane_elems == 0,
                          "%s: regularised segment %d is not a whole number of planes", who, s);
        segs.np4 = reg->plane_elems / 4, segs.n_valid = reg->n_valid;
    }
    return LFS_OK;
}

static inline unsigned adam_grid(int64_t work_items) {
    const int64_t want = (work_items + kAdamThreads - 1) / kAdamThreads;
    const int64_t cap = (int64_t)num_sms() * 8;
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

} // namespace lfs

extern "C" int lfs_adam_step(float* param, float* exp_avg, float* exp_avg_sq, const float* param_grad,
                             int64_t n_elements, float lr, float beta1, float beta2, float eps,
                             float bias_correction1_rcp, float bias_correction2_sqrt_rcp, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(n_elements >= 0, "adam_step: negative n_elements");
    if (n_elements == 0)
        return LFS_OK;
    LFS_CHECK_ARG(param && exp_avg && exp_avg_sq && param_grad, "adam_step: null pointer");
    const uintptr_t a = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(exp_avg) |
                        reinterpret_cast<uintptr_t>(exp_avg_sq) | reinterpret_cast<uintptr_t>(param_grad);
    const int vec_ok = (a & 15u) == 0;
    k_adam_single<<<adam_grid(vec_ok ? (n_elements + 3) / 4 : n_elements), kAdamThreads, 0, (cudaStream_t)stream>>>(
        param, exp_avg, exp_avg_sq, param_grad, n_elements, lr * bias_correction1_rcp, beta1, beta2, eps,
        bias_correction2_sqrt_rcp, vec_ok);
    LFS_LAUNCH_OK("k_adam_single");
    return LFS_OK;
}

extern "C" int lfs_adam_step_multi(float* params, float* exp_avg, float* exp_avg_sq, float* grads, int n_segments,
                                   const int64_t* seg_begin_host, const float* lr_host, const float* bc1_rcp_host,
                                   const float* bc2_sqrt_rcp_host, float beta1, float beta2, float eps, int zero_grad,
                                   const lfs_adam_reg* reg, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(params && exp_avg && exp_avg_sq && grads, "adam_step_multi: null pointer");
    const uintptr_t a = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(exp_avg) |
                        reinterpret_cast<uintptr_t>(exp_avg_sq) | reinterpret_cast<uintptr_t>(grads);
    LFS_CHECK_ARG((a & 15u) == 0, "adam_step_multi: arenas must be 16-byte aligned");
    AdamSegs segs;
    if (int rc = build_segs(segs, "adam_step_multi", n_segments, seg_begin_host, lr_host, bc1_rcp_host, bc2_sqrt_rcp_host,
                            reg))
        return rc;
    const int64_t n4 = segs.begin4[n_segments];
    if (n4 == 0)
        return LFS_OK;
    const size_t base = (size_t)seg_begin_host[0];
    k_adam_multi<<<adam_grid((n4 + 1) / 2), kAdamThreads, 0, (cudaStream_t)stream>>>(
        params + base, exp_avg + base, exp_avg_sq + base, grads + base, segs, beta1, beta2, eps, zero_grad);
    LFS_LAUNCH_OK("k_adam_multi");
    return LFS_OK;
}

// host-only helpers exposing the ownership rule (tests, checkpoint writers that gather the sharded Adam moments)
extern "C" int lfs_adam_p2p_owner(int64_t float_index, int world) {
    if (world <= 0 || float_index < 0)
        return -1;
    return (int)((float_index / 4 / lfs::kOwnChunk4) % world);
}
extern "C" int lfs_adam_p2p_owned_chunks(int64_t begin_float, int64_t end_float, int world, int rank,
                                         int64_t* first_chunk, int64_t* n_chunks, int64_t* chunk_floats) {
    using namespace lfs;
    LFS_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "adam_p2p_owned_chunks: bad rank %d of %d",
                  rank, world);
    LFS_CHECK_ARG((begin_float & 3) == 0 && (end_float & 3) == 0 && end_float >= begin_float,
                  "adam_p2p_owned_chunks: range must be 16-byte aligned");
    const OwnedRange o = owned_range(begin_float / 4, end_float / 4, world, rank);
    if (first_chunk)
        *first_chunk = o.c0;
    if (n_chunks)
        *n_chunks = o.n_owned;
    if (chunk_floats)
        *chunk_floats = kOwnChunk4 * 4;
    return LFS_OK;
}

extern "C" int lfs_adam_step_multi_p2p(float* exp_avg, float* exp_avg_sq, const void* grads_peers_dev,
                                       const void* params_peers_dev, const float* grads_multicast,
                                       float* params_multicast, float* params_local, int world, int rank, int n_segments,
                                       const int64_t* seg_begin_host, const float* lr_host, const float* bc1_rcp_host,
                                       const float* bc2_sqrt_rcp_host, float beta1, float beta2, float eps,
                                       const lfs_adam_reg* reg, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(exp_avg && exp_avg_sq && grads_peers_dev && params_peers_dev, "adam_step_multi_p2p: null pointer");
    LFS_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "adam_step_multi_p2p: bad rank %d of %d", rank, world);
    AdamSegs segs;
    if (int rc = build_segs(segs, "adam_step_multi_p2p", n_segments, seg_begin_host, lr_host, bc1_rcp_host,
                            bc2_sqrt_rcp_host, reg))
        return rc;
    const int64_t n4 = segs.begin4[n_segments];
    if (n4 == 0)
        return LFS_OK;
    LFS_CHECK_ARG(world <= kMaxWorld, "adam_step_multi_p2p: world %d > %d", world, kMaxWorld);
    const int64_t base4 = seg_begin_host[0] / 4;
    const OwnedRange own = owned_range(base4, base4 + n4, world, rank);
    if (own.n_owned == 0)
        return LFS_OK;
    const int64_t cap = (int64_t)num_sms() * 8;
    const unsigned grid = (unsigned)(own.n_owned < cap ? own.n_owned : cap);
    if (grads_multicast && params_multicast) {
        LFS_CHECK_ARG(params_local != nullptr, "adam_step_multi_p2p: params_local is required with multicast pointers");
        k_adam_multi_mc<<<grid, kAdamThreads, 0, (cudaStream_t)stream>>>(
            params_local, exp_avg, exp_avg_sq, grads_multicast, params_multicast, base4, segs, own, beta1, beta2, eps);
        LFS_LAUNCH_OK("k_adam_multi_mc");
        return LFS_OK;
    }
    const PeerPtrs pp{static_cast<float* const*>(grads_peers_dev), static_cast<float* const*>(params_peers_dev), world,
                      rank};
#define LFS_P2P_LAUNCH(W)                                                                                              \
    k_adam_multi_p2p<W><<<grid, kAdamThreads, 0, (cudaStream_t)stream>>>(pp, exp_avg, exp_avg_sq, base4, segs, own,      \
                                                                        beta1, beta2, eps)
    switch (world) {
    case 2: LFS_P2P_LAUNCH(2); break;
    case 4: LFS_P2P_LAUNCH(4); break;
    case 8: LFS_P2P_LAUNCH(8); break;
    default: LFS_P2P_LAUNCH(0); break;
    }
#undef LFS_P2P_LAUNCH
    LFS_LAUNCH_OK("k_adam_multi_p2p");
    return LFS_OK;
}
