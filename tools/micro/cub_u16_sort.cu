// Does cub::DeviceRadixSort::SortPairs(DoubleBuffer<ushort>, DoubleBuffer<uint>, size_t n) -- the exact call of the
// reference's fastgs/rasterization/src/forward.cu:141 -- work on this toolchain / GPU at n ~ 7e5?
#include <cstdio>
#include <cub/cub.cuh>
#include <vector>
template <typename K, typename N> void run(const char* name, N n) {
    std::vector<K> hk(n); std::vector<unsigned> hv(n);
    for (size_t i = 0; i < (size_t)n; ++i) hk[i] = (K)((i * 2654435761u) % 8160u), hv[i] = (unsigned)i;
    K *k0, *k1; unsigned *v0, *v1;
    cudaMalloc(&k0, n * sizeof(K)); cudaMalloc(&k1, n * sizeof(K)); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
    cudaMemcpy(k0, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice); cudaMemcpy(v0, hv.data(), n * 4, cudaMemcpyHostToDevice);
    cub::DoubleBuffer<K> dk(k0, k1); cub::DoubleBuffer<unsigned> dv(v0, v1);
    size_t ws = 0; void* w = nullptr;
    cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, ws, dk, dv, n);
    cudaMalloc(&w, ws);
    cudaError_t e2 = cub::DeviceRadixSort::SortPairs(w, ws, dk, dv, n);
    cudaError_t e3 = cudaDeviceSynchronize();
    cudaMemcpy(hk.data(), dk.Current(), n * sizeof(K), cudaMemcpyDeviceToHost);
    size_t inv = 0;
    for (size_t i = 1; i < (size_t)n; ++i) inv += hk[i] < hk[i - 1];
    printf("%-28s n=%zu ws=%zu query=%s sort=%s sync=%s selector=%d inversions=%zu last=%s\n", name, (size_t)n, ws,
           cudaGetErrorString(e1), cudaGetErrorString(e2), cudaGetErrorString(e3), dk.selector, inv, cudaGetErrorString(cudaGetLastError()));
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(w);
}
int main() {
    run<unsigned short, size_t>("u16 keys, size_t n", (size_t)734704);
    run<unsigned short, int>("u16 keys, int n", 734704);
    run<unsigned short, size_t>("u16 keys, size_t n small", (size_t)20000);
    run<unsigned, size_t>("u32 keys, size_t n", (size_t)734704);
    return 0;
}
