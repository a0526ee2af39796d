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
};

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
        adam_update4(pa, ma, va, ga, beta1, beta2, eps, segs.step_size[sa], segs.bc2[sa]);
        p4[i] = pa, m4[i] = ma, v4[i] = va;
        if (zero_grad)
            g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_j) {
            adam_update4(pb, mb, vb, gb, beta1, beta2, eps, segs.step_size[sb], segs.bc2[sb]);
            p4[j] = pb, m4[j] = mb, v4[j] = vb;
            if (zero_grad)
                g4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// Multi-GPU step: reduce-scatter + Adam + all-gather in ONE kernel over NVLink peer memory.  Rank r owns the slice
// [lo4, hi4) of the flat arena: it sums that slice of every rank's gradient arena with peer loads (fixed rank order ->
// every element is reduced by exactly one rank, deterministically), applies Adam to its local p / m / v, then writes the
// new parameters into every rank's parameter arena with peer stores (every rank clears its own gradients afterwards).
// Compared with ncclAllReduce(236 MB) + a full Adam on every rank this moves the same bytes over NVLink once, does 1/W of
// the Adam traffic per GPU and needs no separate collective launch.  The caller brackets it with two stream-ordered
// barriers (all backward passes done before; all parameter writes landed after).
struct PeerPtrs {
    float* const* grads;  // device array [world]: peer-mapped gradient arenas
    float* const* params; // device array [world]: peer-mapped parameter arenas
    int world, rank;
};
__global__ void __launch_bounds__(kAdamThreads)
    k_adam_multi_p2p(const PeerPtrs pp, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, const size_t base,
                     const AdamSegs segs, const int64_t lo4, const int64_t hi4, const float beta1, const float beta2,
                     const float eps) {
    float4* m4 = reinterpret_cast<float4*>(exp_avg + base);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq + base);
    float4* p4 = reinterpret_cast<float4*>(pp.params[pp.rank] + base);
    const int64_t stride = (int64_t)gridDim.x * kAdamThreads;
    for (int64_t i = lo4 + (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; i < hi4; i += stride) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < pp.world; ++r) {
            const float4 x = __ldcv(reinterpret_cast<const float4*>(pp.grads[r] + base) + i); // never a stale L1 line
            g.x += x.x, g.y += x.y, g.z += x.z, g.w += x.w;
        }
        int s = 0;
#pragma unroll
        for (int k = 1; k < kAdamMaxSeg; ++k)
            if (k < segs.n)
                s += (i >= segs.begin4[k]);
        adam_update4(p, m, v, g, beta1, beta2, eps, segs.step_size[s], segs.bc2[s]);
        m4[i] = m, v4[i] = v;
        for (int r = 0; r < pp.world; ++r)
            reinterpret_cast<float4*>(pp.params[r] + base)[i] = p;
    }
}

// Same step through the NVSwitch multicast object (NVLS): ONE multimem.ld_reduce returns the sum of the slice over all
// ranks, reduced inside the switch (inbound traffic 1x instead of (W-1)x), ONE multimem.st broadcasts the new parameters.
__global__ void __launch_bounds__(kAdamThreads)
    k_adam_multi_mc(float* __restrict__ params_local, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                    const float* __restrict__ grads_mc, float* __restrict__ params_mc, const size_t base,
                    const AdamSegs segs, const int64_t lo4, const int64_t hi4, const float beta1, const float beta2,
                    const float eps) {
    float4* m4 = reinterpret_cast<float4*>(exp_avg + base);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq + base);
    const float4* p4 = reinterpret_cast<const float4*>(params_local + base);
    const int64_t stride = (int64_t)gridDim.x * kAdamThreads;
    for (int64_t i = lo4 + (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; i < hi4; i += stride) {
        float4 p = p4[i], m = m4[i], v = v4[i], g;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w)
                     : "l"(reinterpret_cast<const float4*>(grads_mc + base) + i)
                     : "memory");
        int s = 0;
#pragma unroll
        for (int k = 1; k < kAdamMaxSeg; ++k)
            if (k < segs.n)
                s += (i >= segs.begin4[k]);
        adam_update4(p, m, v, g, beta1, beta2, eps, segs.step_size[s], segs.bc2[s]);
        m4[i] = m, v4[i] = v;
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                         reinterpret_cast<float4*>(params_mc + base) + i),
                     "f"(p.x), "f"(p.y), "f"(p.z), "f"(p.w)
                     : "memory");
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

static inline unsigned adam_grid(int64_t work_items) {
    const int64_t want = (work_items + kAdamThreads - 1) / kAdamThreads;
    const int64_t cap = (int64_t)kNumSMs * 8;
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
                                   void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(n_segments >= 1 && n_segments <= kAdamMaxSeg, "adam_step_multi: n_segments=%d out of range",
                  n_segments);
    LFS_CHECK_ARG(params && exp_avg && exp_avg_sq && grads && seg_begin_host && lr_host && bc1_rcp_host &&
                      bc2_sqrt_rcp_host,
                  "adam_step_multi: null pointer");
    const uintptr_t a = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(exp_avg) |
                        reinterpret_cast<uintptr_t>(exp_avg_sq) | reinterpret_cast<uintptr_t>(grads);
    LFS_CHECK_ARG((a & 15u) == 0, "adam_step_multi: arenas must be 16-byte aligned");
    AdamSegs segs;
    segs.n = n_segments;
    for (int s = 0; s <= n_segments; ++s) {
        LFS_CHECK_ARG((seg_begin_host[s] & 3) == 0, "adam_step_multi: segment boundary %d not a multiple of 4", s);
        LFS_CHECK_ARG(s == 0 || seg_begin_host[s] >= seg_begin_host[s - 1], "adam_step_multi: unsorted segments");
        segs.begin4[s] = (seg_begin_host[s] - seg_begin_host[0]) / 4;
    }
    for (int s = 0; s < n_segments; ++s) {
        segs.step_size[s] = lr_host[s] * bc1_rcp_host[s];
        segs.bc2[s] = bc2_sqrt_rcp_host[s];
    }
    const int64_t n4 = segs.begin4[n_segments];
    if (n4 == 0)
        return LFS_OK;
    const size_t base = (size_t)seg_begin_host[0];
    k_adam_multi<<<adam_grid((n4 + 1) / 2), kAdamThreads, 0, (cudaStream_t)stream>>>(
        params + base, exp_avg + base, exp_avg_sq + base, grads + base, segs, beta1, beta2, eps, zero_grad);
    LFS_LAUNCH_OK("k_adam_multi");
    return LFS_OK;
}

extern "C" int lfs_adam_step_multi_p2p(float* exp_avg, float* exp_avg_sq, const void* grads_peers_dev,
                                       const void* params_peers_dev, const float* grads_multicast,
                                       float* params_multicast, float* params_local, int world, int rank, int n_segments,
                                       const int64_t* seg_begin_host, const float* lr_host, const float* bc1_rcp_host,
                                       const float* bc2_sqrt_rcp_host, float beta1, float beta2, float eps, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(n_segments >= 1 && n_segments <= kAdamMaxSeg, "adam_step_multi_p2p: n_segments=%d out of range",
                  n_segments);
    LFS_CHECK_ARG(exp_avg && exp_avg_sq && grads_peers_dev && params_peers_dev && seg_begin_host && lr_host &&
                      bc1_rcp_host && bc2_sqrt_rcp_host,
                  "adam_step_multi_p2p: null pointer");
    LFS_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "adam_step_multi_p2p: bad rank %d of %d", rank, world);
    AdamSegs segs;
    segs.n = n_segments;
    for (int s = 0; s <= n_segments; ++s) {
        LFS_CHECK_ARG((seg_begin_host[s] & 3) == 0, "adam_step_multi_p2p: segment boundary %d not a multiple of 4", s);
        LFS_CHECK_ARG(s == 0 || seg_begin_host[s] >= seg_begin_host[s - 1], "adam_step_multi_p2p: unsorted segments");
        segs.begin4[s] = (seg_begin_host[s] - seg_begin_host[0]) / 4;
    }
    for (int s = 0; s < n_segments; ++s) {
        segs.step_size[s] = lr_host[s] * bc1_rcp_host[s];
        segs.bc2[s] = bc2_sqrt_rcp_host[s];
    }
    const int64_t n4 = segs.begin4[n_segments];
    if (n4 == 0)
        return LFS_OK;
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo4 = per * rank, hi4 = lo4 + per < n4 ? lo4 + per : n4;
    if (hi4 <= lo4)
        return LFS_OK;
    if (grads_multicast && params_multicast) {
        LFS_CHECK_ARG(params_local != nullptr, "adam_step_multi_p2p: params_local is required with multicast pointers");
        k_adam_multi_mc<<<adam_grid(hi4 - lo4), kAdamThreads, 0, (cudaStream_t)stream>>>(
            params_local, exp_avg, exp_avg_sq, grads_multicast, params_multicast, (size_t)seg_begin_host[0], segs, lo4,
            hi4, beta1, beta2, eps);
        LFS_LAUNCH_OK("k_adam_multi_mc");
        return LFS_OK;
    }
    const PeerPtrs pp{static_cast<float* const*>(grads_peers_dev), static_cast<float* const*>(params_peers_dev), world,
                      rank};
    k_adam_multi_p2p<<<adam_grid(hi4 - lo4), kAdamThreads, 0, (cudaStream_t)stream>>>(
        pp, exp_avg, exp_avg_sq, (size_t)seg_begin_host[0], segs, lo4, hi4, beta1, beta2, eps);
    LFS_LAUNCH_OK("k_adam_multi_p2p");
    return LFS_OK;
}
