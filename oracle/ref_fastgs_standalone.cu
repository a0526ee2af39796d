// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Standalone driver of the UNMODIFIED reference fastgs backend: no torch, no Python, no glue library in the
// process.  It links the reference's own objects (fastgs/rasterization/src/{forward,backward}.cu,
// fastgs/optimizer/src/adam.cu, compiled in place by oracle/Makefile) and calls
//   fast_gs::rasterization::forward   (fastgs/rasterization/include/forward.h:13-38)
//   fast_gs::rasterization::backward  (fastgs/rasterization/include/backward.h:13-50)
//   fast_gs::optimizer::adam_step     (fastgs/optimizer/include/adam.h:9-20)
// exactly as fastgs/rasterization/src/rasterization_api.cu:15-199 does: blobs handed out by resize callbacks and
// NOT zero-filled, gradient outputs zero-filled like backward_wrapper (:125-132), nothing swallowed.
//
//   ref_fastgs_standalone <scene.bin> [--views V] [--steps K] [--warmup W] [--check] [--train]
//
// scene.bin (written by tools/dump_scene.py): int32 header {magic 0x4c465331, N, K_rest, W, H, V, active_bases},
// then fp32 means[N,3] scales_raw[N,3] rotations_raw[N,4] opacities_raw[N] sh0[N,3] shN[N,K_rest,3], then per view
// w2c[16] cam_position[3] fx fy cx cy (23 floats).
//
// --check  after every forward: decode the reference's PerInstanceBuffers / PerTileBuffers (buffer_utils.h:87-137,
//          included from the reference) and verify that the tile keys are sorted, every tile range is consistent
//          and the bucket count equals sum(ceil(range/32)); print one line per view.
// --train  time K steps of {for each view: forward, L1 gradient against a constant target, backward, accumulate}
//          + the 6 Adam launches of FusedAdam::step (src/training/optimizers/fused_adam.cpp:66-93), CUDA events on
//          the legacy default stream (the stream the reference launches on, forward.cu:65).
#include "adam.h"
#include "backward.h"
#include "buffer_utils.h"
#include "forward.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <functional>
#include <string>
#include <tuple>
#include <vector>

#define CK(x)                                                                                       \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) {                                                                    \
            fprintf(stderr, "CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, #x); \
            exit(3);                                                                                \
        }                                                                                           \
    } while (0)

namespace {
    using namespace fast_gs::rasterization;

    struct AbsurdRequest { // thrown out of a resize callback (through the reference's forward) instead of allocating
        size_t bytes;
    };
    struct Blob { // plays the torch byte tensor of fastgs/utils/torch_utils.h:10-17 (resize_ keeps or regrows storage)
        char* ptr = nullptr;
        size_t cap = 0, req = 0;
        char* resize(size_t n) {
            req = n;
            if (n > (64ull << 30))
                throw AbsurdRequest{n};
            if (getenv("REF_FASTGS_TRACE"))
                fprintf(stderr, "[trace] blob request %zu bytes (capacity %zu, ptr %p)\n", n, cap, (void*)ptr);
            if (n > cap) {
                if (ptr)
                    CK(cudaFree(ptr));
                CK(cudaMalloc(&ptr, n));
                cap = n;
            }
            return ptr;
        }
    };

    struct View {
        float w2c[16], campos[3], fx, fy, cx, cy;
    };

    __global__ void k_l1_grad(const float* img, float* g, float target, float scale, int n) {
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) {
            float d = img[i] - target;
            g[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
        }
    }
    __global__ void k_axpy(float* acc, const float* g, int n) {
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n)
            acc[i] += g[i];
    }
} // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s scene.bin [--views V] [--steps K] [--warmup W] [--check] [--train]\n", argv[0]);
        return 2;
    }
    int views_arg = 0, steps = 3, warmup = 1;
    bool check = false, train = false;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--views" && i + 1 < argc) views_arg = atoi(argv[++i]);
        else if (a == "--steps" && i + 1 < argc) steps = atoi(argv[++i]);
        else if (a == "--warmup" && i + 1 < argc) warmup = atoi(argv[++i]);
        else if (a == "--check") check = true;
        else if (a == "--train") train = true;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) {
        perror("scene");
        return 2;
    }
    int hdr[7];
    if (fread(hdr, 4, 7, f) != 7 || hdr[0] != 0x4c465331) {
        fprintf(stderr, "bad scene header\n");
        return 2;
    }
    const int N = hdr[1], KR = hdr[2], W = hdr[3], H = hdr[4], V = hdr[5], bases = hdr[6];
    const int nv = views_arg > 0 && views_arg < V ? views_arg : V;
    auto rd = [&](size_t n) {
        std::vector<float> v(n);
        if (fread(v.data(), 4, n, f) != n) {
            fprintf(stderr, "short scene file\n");
            exit(2);
        }
        return v;
    };
    std::vector<float> h_means = rd((size_t)N * 3), h_scales = rd((size_t)N * 3), h_rot = rd((size_t)N * 4),
                       h_op = rd(N), h_sh0 = rd((size_t)N * 3), h_shN = rd((size_t)N * KR * 3);
    std::vector<View> cams(V);
    for (int v = 0; v < V; ++v) {
        std::vector<float> c = rd(23);
        memcpy(cams[v].w2c, c.data(), 64);
        memcpy(cams[v].campos, c.data() + 16, 12);
        cams[v].fx = c[19], cams[v].fy = c[20], cams[v].cx = c[21], cams[v].cy = c[22];
    }
    fclose(f);

    int rt = 0, drv = 0;
    cudaRuntimeGetVersion(&rt);
    cudaDriverGetVersion(&drv);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("{\"event\":\"start\",\"device\":\"%s\",\"cc\":%d%d,\"cudart\":%d,\"driver\":%d,\"N\":%d,\"W\":%d,\"H\":%d,"
           "\"views\":%d,\"active_sh_bases\":%d}\n", prop.name, prop.major, prop.minor, rt, drv, N, W, H, nv, bases);

    auto up = [](const std::vector<float>& h) {
        float* d;
        CK(cudaMalloc(&d, h.size() * 4 + 16));
        CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        return d;
    };
    float* P[6] = {up(h_means), up(h_sh0), up(h_shN), up(h_scales), up(h_rot), up(h_op)}; // FusedAdam group order
    const size_t PN[6] = {(size_t)N * 3, (size_t)N * 3, (size_t)N * KR * 3, (size_t)N * 3, (size_t)N * 4, (size_t)N};
    const float LR[6] = {0.00016f, 0.0025f, 0.0025f / 20.f, 0.005f, 0.001f, 0.05f}; // eval/default_optimization_params.json
    float *G[6], *A[6], *M1[6], *M2[6];
    for (int k = 0; k < 6; ++k) {
        CK(cudaMalloc(&G[k], PN[k] * 4 + 16));
        CK(cudaMalloc(&A[k], PN[k] * 4 + 16));
        CK(cudaMalloc(&M1[k], PN[k] * 4 + 16));
        CK(cudaMalloc(&M2[k], PN[k] * 4 + 16));
        CK(cudaMemset(M1[k], 0, PN[k] * 4));
        CK(cudaMemset(M2[k], 0, PN[k] * 4));
    }
    float *d_w2c, *d_campos, *image, *alpha, *g_image, *g_alpha, *g_mean2d, *g_conic, *dens;
    CK(cudaMalloc(&d_w2c, V * 64));
    CK(cudaMalloc(&d_campos, V * 16));
    for (int v = 0; v < V; ++v) {
        CK(cudaMemcpy(d_w2c + 16 * v, cams[v].w2c, 64, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_campos + 4 * v, cams[v].campos, 12, cudaMemcpyHostToDevice));
    }
    const size_t PX = (size_t)W * H;
    CK(cudaMalloc(&image, PX * 12));
    CK(cudaMalloc(&alpha, PX * 4));
    CK(cudaMalloc(&g_image, PX * 12));
    CK(cudaMalloc(&g_alpha, PX * 4));
    CK(cudaMemset(g_alpha, 0, PX * 4));
    CK(cudaMalloc(&g_mean2d, (size_t)N * 8));
    CK(cudaMalloc(&g_conic, (size_t)N * 12));
    CK(cudaMalloc(&dens, (size_t)N * 8));
    CK(cudaMemset(dens, 0, (size_t)N * 8));

    Blob b_prim, b_tile, b_inst, b_bucket;
    int bad_views = 0;
    std::tuple<int, int, int, int, int> last;

    auto forward = [&](int v) {
        const View& c = cams[v];
        last = fast_gs::rasterization::forward(
            [&](size_t n) { return b_prim.resize(n); }, [&](size_t n) { return b_tile.resize(n); },
            [&](size_t n) { return b_inst.resize(n); }, [&](size_t n) { return b_bucket.resize(n); },
            reinterpret_cast<const float3*>(P[0]), reinterpret_cast<const float3*>(P[3]),
            reinterpret_cast<const float4*>(P[4]), P[5], reinterpret_cast<const float3*>(P[1]),
            reinterpret_cast<const float3*>(P[2]), reinterpret_cast<const float4*>(d_w2c + 16 * v),
            reinterpret_cast<const float3*>(d_campos + 4 * v), image, alpha, N, bases, KR, W, H, c.fx, c.fy, c.cx, c.cy,
            0.01f, 1e10f); // near / far of src/training/rasterization/fast_rasterizer.cpp:32-33
    };
    auto backward = [&](int v) {
        // zero-fill exactly what backward_wrapper zero-fills (rasterization_api.cu:125-132)
        CK(cudaMemsetAsync(G[0], 0, PN[0] * 4));
        CK(cudaMemsetAsync(G[3], 0, PN[3] * 4));
        CK(cudaMemsetAsync(G[4], 0, PN[4] * 4));
        CK(cudaMemsetAsync(G[5], 0, PN[5] * 4));
        CK(cudaMemsetAsync(G[1], 0, PN[1] * 4));
        CK(cudaMemsetAsync(G[2], 0, PN[2] * 4));
        CK(cudaMemsetAsync(g_mean2d, 0, (size_t)N * 8));
        CK(cudaMemsetAsync(g_conic, 0, (size_t)N * 12));
        fast_gs::rasterization::backward(
            g_image, g_alpha, image, alpha, reinterpret_cast<const float3*>(P[0]),
            reinterpret_cast<const float3*>(P[3]), reinterpret_cast<const float4*>(P[4]),
            reinterpret_cast<const float3*>(P[2]), reinterpret_cast<const float4*>(d_w2c + 16 * v),
            reinterpret_cast<const float3*>(d_campos + 4 * v), b_prim.ptr, b_tile.ptr, b_inst.ptr, b_bucket.ptr,
            reinterpret_cast<float3*>(G[0]), reinterpret_cast<float3*>(G[3]), reinterpret_cast<float4*>(G[4]), G[5],
            reinterpret_cast<float3*>(G[1]), reinterpret_cast<float3*>(G[2]), reinterpret_cast<float2*>(g_mean2d),
            g_conic, nullptr, dens, N, std::get<0>(last), std::get<1>(last), std::get<2>(last), std::get<3>(last),
            std::get<4>(last), bases, KR, W, H, cams[v].fx, cams[v].fy, cams[v].cx, cams[v].cy);
    };

    auto check_view = [&](int v) {
        CK(cudaDeviceSynchronize());
        cudaError_t le = cudaGetLastError();
        const int n_vis = std::get<0>(last), n_inst = std::get<1>(last), n_buckets = std::get<2>(last);
        const int n_tiles = ((W + 15) / 16) * ((H + 15) / 16);
        long long inv = -1, bad_rng = -1, sum_b = -1, over = -1, max_key = -1;
        double img_sum = 0;
        int finite = 1;
        if (n_inst > 0 && n_inst < (1 << 30)) {
            char* p = b_inst.ptr;
            PerInstanceBuffers ib = PerInstanceBuffers::from_blob(p, (size_t)n_inst);
            ib.keys.selector = std::get<4>(last); // keys and values flip together inside cub
            std::vector<unsigned short> k(n_inst);
            CK(cudaMemcpy(k.data(), ib.keys.Current(), (size_t)n_inst * 2, cudaMemcpyDeviceToHost));
            inv = 0, over = 0, max_key = 0;
            for (int i = 0; i < n_inst; ++i) {
                if (i && k[i] < k[i - 1]) ++inv;
                if (k[i] >= n_tiles) ++over;
                if (k[i] > max_key) max_key = k[i];
            }
            char* q = b_tile.ptr;
            PerTileBuffers tb = PerTileBuffers::from_blob(q, (size_t)n_tiles);
            std::vector<unsigned> rng(2 * (size_t)n_tiles);
            CK(cudaMemcpy(rng.data(), tb.instance_ranges, rng.size() * 4, cudaMemcpyDeviceToHost));
            bad_rng = 0, sum_b = 0;
            for (int t = 0; t < n_tiles; ++t) {
                if (rng[2 * t + 1] < rng[2 * t] || rng[2 * t + 1] > (unsigned)n_inst) ++bad_rng;
                else sum_b += (rng[2 * t + 1] - rng[2 * t] + 31) / 32;
            }
        }
        {
            std::vector<float> h(PX * 3);
            CK(cudaMemcpy(h.data(), image, PX * 12, cudaMemcpyDeviceToHost));
            for (float x : h) {
                if (!std::isfinite(x)) finite = 0;
                img_sum += x;
            }
        }
        const bool ok = inv == 0 && bad_rng == 0 && over == 0 && sum_b == n_buckets && finite && le == cudaSuccess;
        if (!ok) ++bad_views;
        printf("{\"event\":\"check\",\"view\":%d,\"n_visible\":%d,\"n_instances\":%d,\"n_buckets\":%d,"
               "\"key_inversions\":%lld,\"keys_out_of_range\":%lld,\"max_key\":%lld,\"n_tiles\":%d,\"bad_ranges\":%lld,"
               "\"sum_ceil_range_32\":%lld,\"image_finite\":%d,\"image_mean\":%.6f,\"last_error\":\"%s\",\"ok\":%s}\n",
               v, n_vis, n_inst, n_buckets, inv, over, max_key, n_tiles, bad_rng, sum_b, finite,
               img_sum / (double)(PX * 3), cudaGetErrorString(le), ok ? "true" : "false");
        fflush(stdout);
    };

    // after the reference asked for an absurd blob: which intermediate result is the first wrong one?
    auto diagnose = [&](int v, size_t bytes) {
        CK(cudaDeviceSynchronize());
        cudaError_t le = cudaGetLastError();
        const int n_tiles = ((W + 15) / 16) * ((H + 15) / 16);
        char* p = b_prim.ptr;
        PerPrimitiveBuffers pb = PerPrimitiveBuffers::from_blob(p, (size_t)N);
        unsigned n_vis = 0, n_inst = 0;
        CK(cudaMemcpy(&n_vis, pb.n_visible_primitives, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&n_inst, pb.n_instances, 4, cudaMemcpyDeviceToHost));
        auto inversions_u32 = [&](const unsigned* d, size_t n) {
            std::vector<unsigned> h(n);
            CK(cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost));
            long long inv = 0;
            for (size_t i = 1; i < n; ++i) inv += h[i] < h[i - 1];
            return inv;
        };
        long long dk0 = -1, dk1 = -1, off_bad = -1, sum_touched = -1;
        if (n_vis > 0 && n_vis <= (unsigned)N) {
            dk0 = inversions_u32(pb.depth_keys.d_buffers[0], n_vis);
            dk1 = inversions_u32(pb.depth_keys.d_buffers[1], n_vis);
            std::vector<unsigned> off(n_vis), touched(N), idx0(n_vis), idx1(n_vis);
            CK(cudaMemcpy(off.data(), pb.offset, (size_t)n_vis * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(touched.data(), pb.n_touched_tiles, (size_t)N * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(idx0.data(), pb.primitive_indices.d_buffers[0], (size_t)n_vis * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(idx1.data(), pb.primitive_indices.d_buffers[1], (size_t)n_vis * 4, cudaMemcpyDeviceToHost));
            // which index buffer is consistent with the offsets?  offset[i+1] - offset[i] == n_touched[idx[i]]
            for (int which = 0; which < 2; ++which) {
                const std::vector<unsigned>& idx = which ? idx1 : idx0;
                long long bad = 0, sum = 0;
                for (unsigned i = 0; i < n_vis; ++i) {
                    const unsigned g = idx[i];
                    const unsigned t = g < (unsigned)N ? touched[g] : 0xffffffffu;
                    sum += t;
                    const unsigned next = i + 1 < n_vis ? off[i + 1] : n_inst;
                    bad += (next - off[i]) != t;
                }
                if (which == 0 || bad < off_bad) off_bad = bad, sum_touched = sum;
            }
        }
        long long k0 = -1, k1 = -1, over0 = -1, over1 = -1;
        if (b_inst.ptr && n_inst > 0 && n_inst < (1u << 30)) {
            char* q = b_inst.ptr;
            PerInstanceBuffers ib = PerInstanceBuffers::from_blob(q, (size_t)n_inst);
            for (int which = 0; which < 2; ++which) {
                std::vector<unsigned short> k(n_inst);
                CK(cudaMemcpy(k.data(), ib.keys.d_buffers[which], (size_t)n_inst * 2, cudaMemcpyDeviceToHost));
                long long inv = 0, over = 0;
                for (unsigned i = 0; i < n_inst; ++i) {
                    if (i && k[i] < k[i - 1]) ++inv;
                    over += k[i] >= n_tiles;
                }
                (which ? k1 : k0) = inv, (which ? over1 : over0) = over;
            }
        }
        long long bad_rng = -1;
        if (b_tile.ptr) {
            char* q = b_tile.ptr;
            PerTileBuffers tb = PerTileBuffers::from_blob(q, (size_t)n_tiles);
            std::vector<unsigned> rng(2 * (size_t)n_tiles);
            CK(cudaMemcpy(rng.data(), tb.instance_ranges, rng.size() * 4, cudaMemcpyDeviceToHost));
            bad_rng = 0;
            for (int t = 0; t < n_tiles; ++t) bad_rng += rng[2 * t + 1] < rng[2 * t] || rng[2 * t + 1] > n_inst;
        }
        printf("{\"event\":\"absurd_request\",\"view\":%d,\"bytes\":%zu,\"last_error\":\"%s\",\"n_visible\":%u,\"n_instances\":%u,"
               "\"depth_key_inversions\":[%lld,%lld],\"offsets_inconsistent_with_touched\":%lld,\"sum_touched\":%lld,"
               "\"tile_key_inversions\":[%lld,%lld],\"tile_keys_out_of_range\":[%lld,%lld],\"bad_tile_ranges\":%lld}\n",
               v, bytes, cudaGetErrorString(le), n_vis, n_inst, dk0, dk1, off_bad, sum_touched, k0, k1, over0, over1, bad_rng);
        fflush(stdout);
    };

    if (check) {
        for (int v = 0; v < nv; ++v) {
            try {
                forward(v);
            } catch (const AbsurdRequest& a) {
                diagnose(v, a.bytes);
                ++bad_views;
                continue;
            }
            check_view(v);
        }
        if (bad_views) {
            printf("{\"event\":\"verdict\",\"ok\":false,\"bad_views\":%d}\n", bad_views);
            return 1;
        }
        printf("{\"event\":\"verdict\",\"ok\":true,\"views\":%d}\n", nv);
    }

    if (train) {
        const float scale = 1.0f / (3.0f * (float)PX);
        const int T = 256;
        auto step = [&](int t) {
            for (int v = 0; v < nv; ++v) {
                forward(v);
                k_l1_grad<<<(unsigned)((PX * 3 + T - 1) / T), T>>>(image, g_image, 0.5f, scale, (int)(PX * 3));
                backward(v);
                for (int k = 0; k < 6; ++k) {
                    if (v == 0)
                        CK(cudaMemcpyAsync(A[k], G[k], PN[k] * 4, cudaMemcpyDeviceToDevice));
                    else
                        k_axpy<<<(unsigned)((PN[k] + T - 1) / T), T>>>(A[k], G[k], (int)PN[k]);
                }
            }
            const float bc1 = 1.0f / (1.0f - powf(0.9f, (float)t)), bc2 = 1.0f / sqrtf(1.0f - powf(0.999f, (float)t));
            for (int k = 0; k < 6; ++k)
                fast_gs::optimizer::adam_step(P[k], M1[k], M2[k], A[k], (int)PN[k], LR[k], 0.9f, 0.999f, 1e-15f, bc1, bc2);
        };
        for (int i = 0; i < warmup; ++i) step(i + 1);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, 0));
        for (int i = 0; i < steps; ++i) step(warmup + i + 1);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= (float)steps;
        // forward-only and backward-only splits (one view, averaged)
        float ms_f = 0, ms_b = 0;
        {
            CK(cudaEventRecord(e0, 0));
            for (int i = 0; i < 4; ++i) forward(i % nv);
            CK(cudaEventRecord(e1, 0));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms_f, e0, e1));
            ms_f /= 4.f;
            forward(0);
            k_l1_grad<<<(unsigned)((PX * 3 + T - 1) / T), T>>>(image, g_image, 0.5f, scale, (int)(PX * 3));
            CK(cudaEventRecord(e0, 0));
            for (int i = 0; i < 4; ++i) backward(0);
            CK(cudaEventRecord(e1, 0));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms_b, e0, e1));
            ms_b /= 4.f;
        }
        printf("{\"event\":\"train\",\"impl\":\"reference fastgs (EWA) CUDA build, unmodified, standalone (no torch)\","
               "\"value\":%.4f,\"unit\":\"views/s\",\"ms_per_step\":%.4f,\"views_per_step\":%d,\"steps\":%d,\"warmup\":%d,"
               "\"forward_ms_per_view\":%.4f,\"backward_ms_per_view\":%.4f,\"n_instances_last\":%d,\"n_buckets_last\":%d,"
               "\"loss\":\"L1 gradient against a constant target (cheaper than the reference's L1+SSIM)\"}\n",
               nv / ms * 1e3, ms, nv, steps, warmup, ms_f, ms_b, std::get<1>(last), std::get<2>(last));
    }
    return 0;
}
