// Microbenchmark 2: which ingredient of the per-Gaussian backward kernel costs the bandwidth?
// base = 59-plane planar RMW (5.8 TB/s measured). Flags add: C = dependent `counts` early-exit load,
// A = AoS stride-12 accumulator loads + zeroing stores, R = 3 RED atomics into planes, X = ~1000 extra FMAs.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int P = 59;
template <bool C, int A, bool R, bool X>
__global__ void __launch_bounds__(256) k(const float* __restrict__ a, float* __restrict__ g, const int* __restrict__ counts,
                                         float* __restrict__ vm, float* __restrict__ vq, unsigned n, size_t Np) {
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (C && counts[i] <= 0) return;
    float extra = 0.f;
    if (A == 2) { // same accumulators, planar (SoA)
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) { extra += vm[(size_t)c2 * n + i]; }
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) { extra += vq[(size_t)c2 * n + i]; }
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) vm[(size_t)c2 * n + i] = 0.f;
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) vq[(size_t)c2 * n + i] = 0.f;
    }
    if (A == 3) { // AoS loads only, no zeroing
        extra = vm[3 * (size_t)i] + vm[3 * (size_t)i + 1] + vm[3 * (size_t)i + 2];
        const float4 q = reinterpret_cast<float4*>(vq)[i];
        extra += q.x + q.y + q.z + q.w;
    }
    if (A == 4) { // AoS zeroing only
        vm[3 * (size_t)i] = vm[3 * (size_t)i + 1] = vm[3 * (size_t)i + 2] = 0.f;
        reinterpret_cast<float4*>(vq)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (A == 1) {
        extra = vm[3 * (size_t)i] + vm[3 * (size_t)i + 1] + vm[3 * (size_t)i + 2];
        const float4 q = reinterpret_cast<float4*>(vq)[i];
        extra += q.x + q.y + q.z + q.w;
        vm[3 * (size_t)i] = vm[3 * (size_t)i + 1] = vm[3 * (size_t)i + 2] = 0.f;
        reinterpret_cast<float4*>(vq)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float va[P], vg[P];
#pragma unroll
    for (int p = 0; p < P; ++p) va[p] = __ldg(a + (size_t)p * Np + i), vg[p] = g[(size_t)p * Np + i];
    if (X) {
#pragma unroll 1
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) va[p] = fmaf(va[p], 1.0001f, extra);
    }
#pragma unroll
    for (int p = R ? 3 : 0; p < P; ++p) g[(size_t)p * Np + i] = fmaf(va[p], 0.5f, vg[p] + extra);
    if (R) {
        atomicAdd(g + i, va[0]); atomicAdd(g + Np + i, va[1]); atomicAdd(g + 2 * Np + i, va[2]);
    }
}
template <bool C, int A, bool R, bool X> void run(const char* name, const float* a, float* g, const int* c, float* vm, float* vq, unsigned n, size_t Np, float* flush, size_t fl) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
        cudaMemsetAsync(flush, r, fl);
        cudaEventRecord(e0);
        k<C, A, R, X><<<(n + 255) / 256, 256>>>(a, g, c, vm, vq, n, Np);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    printf("%-28s %.3f ms  %.2f TB/s (plane bytes only)\n", name, best, (double)n * P * 12 / best / 1e9);
}
int main() {
    const unsigned n = 1000000; const size_t Np = 1000000;
    float *a, *g, *flush, *vm, *vq; int* c; const size_t fl = 512u << 20;
    cudaMalloc(&a, Np * P * 4); cudaMalloc(&g, Np * P * 4); cudaMalloc(&flush, fl); cudaMalloc(&vm, n * 12); cudaMalloc(&vq, n * 16); cudaMalloc(&c, n * 4);
    cudaMemset(a, 0, Np * P * 4); cudaMemset(g, 0, Np * P * 4); cudaMemset(vm, 0, n * 12); cudaMemset(vq, 0, n * 16); cudaMemset(c, 1, n * 4);
    run<false, 0, false, false>("base", a, g, c, vm, vq, n, Np, flush, fl);
    run<true, 0, false, false>("+counts", a, g, c, vm, vq, n, Np, flush, fl);
    run<false, 1, false, false>("+aos", a, g, c, vm, vq, n, Np, flush, fl);
    run<false, 0, true, false>("+red", a, g, c, vm, vq, n, Np, flush, fl);
    run<false, 0, false, true>("+alu", a, g, c, vm, vq, n, Np, flush, fl);
    run<false, 2, false, false>("+soa accumulators", a, g, c, vm, vq, n, Np, flush, fl);
    run<false, 3, false, false>("+aos loads only", a, g, c, vm, vq, n, Np, flush, fl);
    run<false, 4, false, false>("+aos zeroing only", a, g, c, vm, vq, n, Np, flush, fl);
    run<true, 1, true, true>("+all", a, g, c, vm, vq, n, Np, flush, fl);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
