// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin extern "C" glue around the UNMODIFIED reference fastgs backend
// (/root/reference/fastgs/rasterization/src/{forward,backward}.cu and
// /root/reference/fastgs/optimizer/src/adam.cu, compiled in place by oracle/Makefile
// into oracle/_ref/libfastgs_ref.so).  It calls the reference's own raw-pointer entry points
//   fast_gs::rasterization::forward   (fastgs/rasterization/include/forward.h:13-38)
//   fast_gs::rasterization::backward  (fastgs/rasterization/include/backward.h:13-50)
//   fast_gs::optimizer::adam_step     (fastgs/optimizer/include/adam.h:9-20)
// so tests / bench can run the reference CUDA path on the GPU box without libtorch.
// Scratch blobs are owned here (cudaMalloc, grow-only) and play the role of the torch byte
// tensors of fastgs/utils/torch_utils.h:10-17.
#include "adam.h"
#include "backward.h"
#include "forward.h"

#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <functional>
#include <tuple>
#include <vector>

namespace {
    struct Blob {
        char* ptr = nullptr;
        size_t cap = 0;
        char* resize(size_t n) {
            if (getenv("REF_FASTGS_DEBUG"))
                fprintf(stderr, "[ref_fastgs] resize request %zu bytes (cap %zu)\n", n, cap);
            if (n > cap) {
                if (ptr)
                    cudaFree(ptr);
                size_t want = n + n / 4 + 256;
                cudaError_t e = cudaMalloc(&ptr, want);
                if (e != cudaSuccess) {
                    fprintf(stderr, "[ref_fastgs] cudaMalloc(%zu) failed: %s\n", want, cudaGetErrorString(e));
                    ptr = nullptr;
                    cap = 0;
                    return nullptr;
                }
                cap = want;
            }
            return ptr;
        }
    };
    struct Ctx {
        Blob prim, tile, inst, bucket;
        int n_visible = 0, n_instances = 0, n_buckets = 0, sel_prim = 0, sel_inst = 0;
        size_t inst_req = 0, inst_n = 0;
    };
} // namespace

extern "C" {

void* ref_fastgs_create() { return new Ctx(); }

void ref_fastgs_destroy(void* h) {
    Ctx* c = static_cast<Ctx*>(h);
    if (!c)
        return;
    for (Blob* b : {&c->prim, &c->tile, &c->inst, &c->bucket})
        if (b->ptr)
            cudaFree(b->ptr);
    delete c;
}

// All pointers are device pointers. w2c: 16 floats row-major; cam_position: 3 floats.
// image [3,H,W], alpha [1,H,W]. out_counts[3] (host) = n_visible, n_instances, n_buckets.
int ref_fastgs_forward(void* h, const float* means, const float* scales_raw, const float* rotations_raw,
                       const float* opacities_raw, const float* sh0, const float* shN, const float* w2c,
                       const float* cam_position, float* image, float* alpha, int n_primitives,
                       int active_sh_bases, int total_bases_sh_rest, int width, int height, float fx, float fy,
                       float cx, float cy, float near_plane, float far_plane, int* out_counts) {
    Ctx* c = static_cast<Ctx*>(h);
    {
        cudaError_t pre = cudaGetLastError(); // do not inherit a stale error from the caller
        if (pre != cudaSuccess)
            fprintf(stderr, "[ref_fastgs] stale CUDA error before forward: %s\n", cudaGetErrorString(pre));
    }
    const bool dbg = getenv("REF_FASTGS_DEBUG") != nullptr;
    const int n_tiles_dbg = ((width + 15) / 16) * ((height + 15) / 16);
    auto stage = [dbg](const char* where) {
        if (!dbg)
            return;
        cudaError_t e1 = cudaDeviceSynchronize();
        cudaError_t e2 = cudaGetLastError();
        fprintf(stderr, "[ref_fastgs] at %s: sync=%s last=%s\n", where, cudaGetErrorString(e1), cudaGetErrorString(e2));
    };
    auto f_prim = [c, stage](size_t n) {
        stage("per_primitive alloc (after per-tile memset)");
        return c->prim.resize(n);
    };
    // The reference zeroes PerTileBuffers::instance_ranges with cudaMemsetAsync on a private static stream
    // (fastgs/rasterization/src/forward.cu:48-55).  On this image (CUDA 12.9 runtime inside a torch process)
    // that call returns cudaErrorInvalidValue (compute-sanitizer, profiles/r01_ref_fastgs_diagnosis.txt), the ranges of
    // empty tiles stay uninitialised and n_buckets becomes garbage.  Zeroing the whole per-tile blob here, right
    // before the reference carves it, gives the state the reference intends without touching its sources.
    auto f_tile = [c](size_t n) {
        char* p = c->tile.resize(n);
        if (p)
            cudaMemset(p, 0, n);
        return p;
    };
    auto f_inst = [c, stage](size_t n) {
        stage("per_instance alloc (after preprocess + depth sort + scan)");
        // On sm_100a the reference's two copies of the exact tile test (preprocess_cu counts, create_instances_cu
        // emits; kernel_utils.cuh:108-210) do not always agree under --use_fast_math, so a few instance slots are
        // never written.  Their uninitialised u16 tile keys can exceed n_tiles and extract_instance_ranges_cu then
        // writes out of bounds (observed: garbage bucket counts at 1080p).  Zero-filling the blob makes such slots
        // harmless duplicates in tile 0; it does not change anything the reference writes itself.
        char* p = c->inst.resize(n);
        if (p)
            cudaMemset(p, 0, n);
        c->inst_req = n;
        return p;
    };
    auto f_bucket = [c, stage, dbg, n_tiles_dbg](size_t n) {
        stage("per_bucket alloc (after create_instances + tile sort + ranges + bucket scan)");
        if (dbg && c->tile.ptr) { // PerTileBuffers layout (buffer_utils.h:114-137): ranges | n_buckets | bucket_offsets
            const size_t o_nb = ((size_t)n_tiles_dbg * 8 + 127) / 128 * 128;
            const size_t o_bo = (o_nb + (size_t)n_tiles_dbg * 4 + 127) / 128 * 128;
            std::vector<unsigned> rng(2 * (size_t)n_tiles_dbg), nb(n_tiles_dbg), bo(n_tiles_dbg);
            cudaMemcpy(rng.data(), c->tile.ptr, rng.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(nb.data(), c->tile.ptr + o_nb, nb.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(bo.data(), c->tile.ptr + o_bo, bo.size() * 4, cudaMemcpyDeviceToHost);
            unsigned long long snb = 0, maxy = 0, bad = 0;
            for (int t = 0; t < n_tiles_dbg; ++t) {
                snb += nb[t];
                if (rng[2 * t + 1] > maxy)
                    maxy = rng[2 * t + 1];
                if (rng[2 * t + 1] < rng[2 * t])
                    ++bad;
            }
            c->inst_n = (size_t)maxy;
            fprintf(stderr, "[ref_fastgs] tiles=%d sum(n_buckets)=%llu last(bucket_offsets)=%u max(range.y)=%llu "
                            "ranges with y<x=%llu first ranges (%u,%u) (%u,%u)\n",
                    n_tiles_dbg, snb, bo[n_tiles_dbg - 1], maxy, bad, rng[0], rng[1], rng[2], rng[3]);
        }
        if (dbg && c->inst.ptr && c->inst_n > 0) { // PerInstanceBuffers: keys cur | keys alt | idx cur | idx alt
            const size_t ni = c->inst_n;
            const size_t o_alt = (ni * 2 + 127) / 128 * 128;
            for (int which = 0; which < 2; ++which) {
                std::vector<unsigned short> k(ni);
                cudaMemcpy(k.data(), c->inst.ptr + (which ? o_alt : 0), ni * 2, cudaMemcpyDeviceToHost);
                size_t inv = 0;
                unsigned mx = 0;
                for (size_t i = 0; i < ni; ++i) {
                    if (i && k[i] < k[i - 1])
                        ++inv;
                    if (k[i] > mx)
                        mx = k[i];
                }
                fprintf(stderr, "[ref_fastgs] key buffer %d: n=%zu inversions=%zu max_key=%u first=%u last=%u\n", which,
                        ni, inv, mx, (unsigned)k[0], (unsigned)k[ni - 1]);
            }
        }
        return c->bucket.resize(n);
    };
    auto r = fast_gs::rasterization::forward(
        f_prim, f_tile, f_inst, f_bucket, reinterpret_cast<const float3*>(means),
        reinterpret_cast<const float3*>(scales_raw), reinterpret_cast<const float4*>(rotations_raw), opacities_raw,
        reinterpret_cast<const float3*>(sh0), reinterpret_cast<const float3*>(shN),
        reinterpret_cast<const float4*>(w2c), reinterpret_cast<const float3*>(cam_position), image, alpha,
        n_primitives, active_sh_bases, total_bases_sh_rest, width, height, fx, fy, cx, cy, near_plane, far_plane);
    {
        cudaError_t e = cudaGetLastError(); // swallow the reference's failed cudaMemsetAsync (see above)
        if (e != cudaSuccess && e != cudaErrorInvalidValue)
            return (int)e;
    }
    c->n_visible = std::get<0>(r);
    c->n_instances = std::get<1>(r);
    c->n_buckets = std::get<2>(r);
    c->sel_prim = std::get<3>(r);
    c->sel_inst = std::get<4>(r);
    if (out_counts) {
        out_counts[0] = c->n_visible;
        out_counts[1] = c->n_instances;
        out_counts[2] = c->n_buckets;
    }
    return (int)cudaDeviceSynchronize();
}

// Gradient outputs must be zero-initialised by the caller exactly as backward_wrapper does
// (fastgs/rasterization/src/rasterization_api.cu:125-132). grad_mean2d_helper [N,2],
// grad_conic_helper [3,N]. densification_info may be null.
int ref_fastgs_backward(void* h, const float* grad_image, const float* grad_alpha, const float* image,
                        const float* alpha, const float* means, const float* scales_raw, const float* rotations_raw,
                        const float* shN, const float* w2c, const float* cam_position, float* grad_means,
                        float* grad_scales_raw, float* grad_rotations_raw, float* grad_opacities_raw, float* grad_sh0,
                        float* grad_shN, float* grad_mean2d_helper, float* grad_conic_helper,
                        float* densification_info, int n_primitives, int active_sh_bases, int total_bases_sh_rest,
                        int width, int height, float fx, float fy, float cx, float cy) {
    Ctx* c = static_cast<Ctx*>(h);
    fast_gs::rasterization::backward(
        grad_image, grad_alpha, image, alpha, reinterpret_cast<const float3*>(means),
        reinterpret_cast<const float3*>(scales_raw), reinterpret_cast<const float4*>(rotations_raw),
        reinterpret_cast<const float3*>(shN), reinterpret_cast<const float4*>(w2c),
        reinterpret_cast<const float3*>(cam_position), c->prim.ptr, c->tile.ptr, c->inst.ptr, c->bucket.ptr,
        reinterpret_cast<float3*>(grad_means), reinterpret_cast<float3*>(grad_scales_raw),
        reinterpret_cast<float4*>(grad_rotations_raw), grad_opacities_raw, reinterpret_cast<float3*>(grad_sh0),
        reinterpret_cast<float3*>(grad_shN), reinterpret_cast<float2*>(grad_mean2d_helper), grad_conic_helper,
        nullptr, densification_info, n_primitives, c->n_visible, c->n_instances, c->n_buckets, c->sel_prim,
        c->sel_inst, active_sh_bases, total_bases_sh_rest, width, height, fx, fy, cx, cy);
    return (int)cudaGetLastError();
}

int ref_fastgs_adam_step(float* param, float* exp_avg, float* exp_avg_sq, const float* grad, int n, float lr,
                         float beta1, float beta2, float eps, float bc1_rcp, float bc2_sqrt_rcp) {
    fast_gs::optimizer::adam_step(param, exp_avg, exp_avg_sq, grad, n, lr, beta1, beta2, eps, bc1_rcp, bc2_sqrt_rcp);
    return (int)cudaGetLastError();
}

} // extern "C"
