// Issue / pipe rates that decide how the blend kernels are written for sm_100a:
//   FFMA (3 register operands) vs FFMA2 (fma.rn.f32x2, two fp32 FMAs per instruction), alone and mixed with MUFU.EX2 and
//   LDS.128, as warp-instructions per clock per SM at several occupancies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* clk) {
    __shared__ float4 sm[256];
    sm[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    __syncthreads();
    float a[8], m = 1.0001f + threadIdx.x * 1e-7f, c = 0.5f;
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = i + threadIdx.x;
        p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f);
    }
    const unsigned long long pm = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
    const unsigned long long pc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    float e = 0.25f;
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { // 16 FFMA
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], m, c);
        } else if (MODE == 1) { // 16 FFMA2
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pm, pc);
        } else if (MODE == 2) { // 16 FFMA + 2 MUFU
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], m, c);
                e = ex2(e * 0.5f);
            }
        } else if (MODE == 3) { // 16 FFMA2 + 2 MUFU
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pm, pc);
                e = ex2(e * 0.5f);
            }
        } else if (MODE == 4) { // 16 FFMA + 4 LDS.128
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], m, c);
                float4 v = sm[(threadIdx.x + it + r) & 255], w = sm[(threadIdx.x * 3 + it + r) & 255];
                acc.x += v.x + w.y;
            }
        } else if (MODE == 5) { // 16 FFMA2 + 4 LDS.128
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pm, pc);
                float4 v = sm[(threadIdx.x + it + r) & 255], w = sm[(threadIdx.x * 3 + it + r) & 255];
                acc.x += v.x + w.y;
            }
        }
    }
    long long t1 = clock64();
    float s = e + acc.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int fma_per_iter, int lanes_per_fma) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    long long* clk;
    cudaMalloc(&out, sizeof(float) * 256 * sms * 8);
    cudaMalloc(&clk, sizeof(long long) * sms * 8);
    const int iters = 4096;
    for (int cta_per_sm : {1, 2, 4, 8}) {
        k<MODE><<<sms * cta_per_sm, 256>>>(out, iters, clk);
        k<MODE><<<sms * cta_per_sm, 256>>>(out, iters, clk);
        cudaDeviceSynchronize();
        long long h[2048];
        cudaMemcpy(h, clk, sizeof(long long) * sms * cta_per_sm, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms * cta_per_sm; ++i) mean += h[i];
        mean /= sms * cta_per_sm;
        const double warps = 8.0 * cta_per_sm;
        const double fma_warp_instr = (double)iters * fma_per_iter * warps;
        printf("{\"mode\":\"%s\",\"warps_per_sm\":%d,\"clk\":%.0f,\"fma_warp_instr_per_clk_per_sm\":%.3f,"
               "\"fp32_fma_lanes_per_clk_per_sm\":%.1f}\n",
               name, (int)warps, mean, fma_warp_instr / mean, fma_warp_instr / mean * 32 * lanes_per_fma);
    }
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0>("16 FFMA", 16, 1);
    run<1>("16 FFMA2", 16, 2);
    run<2>("16 FFMA + 2 MUFU + 2 FMUL", 16, 1);
    run<3>("16 FFMA2 + 2 MUFU + 2 FMUL", 16, 2);
    run<4>("16 FFMA + 4 LDS.128", 16, 1);
    run<5>("16 FFMA2 + 4 LDS.128", 16, 2);
    return 0;
}
