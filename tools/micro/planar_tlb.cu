// Microbenchmark: 59-plane read-modify-write, fully planar [plane][N] vs tiled planar [N/T][plane][T].
// Question: is the ~2 TB/s ceiling of the per-Gaussian trainer kernels a TLB / page-locality effect?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int P = 59;
template <int T> // T == 0: fully planar
__global__ void __launch_bounds__(256) k_rmw(const float* __restrict__ a, float* __restrict__ g, unsigned n, size_t Np) {
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float va[P], vg[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const size_t idx = T == 0 ? (size_t)p * Np + i : (size_t)(i / (T ? T : 1)) * (P * (T ? T : 1)) + (size_t)p * (T ? T : 1) + i % (T ? T : 1);
        va[p] = __ldg(a + idx), vg[p] = g[idx];
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const size_t idx = T == 0 ? (size_t)p * Np + i : (size_t)(i / (T ? T : 1)) * (P * (T ? T : 1)) + (size_t)p * (T ? T : 1) + i % (T ? T : 1);
        g[idx] = fmaf(va[p], 0.5f, vg[p]);
    }
}
template <int T> float run(const float* a, float* g, unsigned n, size_t Np, float* flush, size_t flush_n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
        cudaMemsetAsync(flush, r, flush_n); // evict L2
        cudaEventRecord(e0);
        k_rmw<T><<<(n + 255) / 256, 256>>>(a, g, n, Np);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}
int main() {
    const unsigned n = 1000000; const size_t Np = 1000448; // multiple of 1024
    float *a, *g, *flush; const size_t fl = 512u << 20;
    cudaMalloc(&a, Np * P * 4); cudaMalloc(&g, Np * P * 4); cudaMalloc(&flush, fl);
    cudaMemset(a, 0, Np * P * 4); cudaMemset(g, 0, Np * P * 4);
    const double bytes = (double)n * P * 12;
    float t0 = run<0>(a, g, n, Np, flush, fl), t1 = run<256>(a, g, n, Np, flush, fl), t2 = run<1024>(a, g, n, Np, flush, fl), t3 = run<32>(a, g, n, Np, flush, fl);
    printf("planar   %.3f ms %.2f TB/s\n", t0, bytes / t0 / 1e9);
    printf("tile256  %.3f ms %.2f TB/s\n", t1, bytes / t1 / 1e9);
    printf("tile1024 %.3f ms %.2f TB/s\n", t2, bytes / t2 / 1e9);
    printf("tile32   %.3f ms %.2f TB/s\n", t3, bytes / t3 / 1e9);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
