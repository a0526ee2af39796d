// lfs_b200 -- lfs_spherical_harmonics_fwd / _bwd: drop-ins for gsplat::spherical_harmonics_fwd/bwd
// (reference gsplat/SphericalHarmonics.cpp:15-75, kernels gsplat/SphericalHarmonicsCUDA.cu:374-481).
//
// One thread per element handles all three channels (the reference spends three threads per element,
// each walking coeffs with stride 3).  A block first stages its K*3 coefficients per element with
// 128-bit coalesced loads into shared memory when the tile fits (K <= 16), then each thread reads its own
// row (row pitch K*3+1 floats -> conflict-free).  Algorithmic bytes: fwd (12 + 12K + 1) in, 12 out;
// bwd (12 + 12K + 12 + 1) in, 12K + 12 out per element.
#include "sh.cuh"

namespace lfs {

constexpr int kShThreads = 128;
constexpr int kShMaxStageK = 16;

// Cooperative, coalesced copy of `count` rows of `row_floats` floats into smem with pitch row_floats + 1.
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, uint32_t count, uint32_t row_floats,
                                           float* smem) {
    const uint32_t total = count * row_floats;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    const uint32_t pitch = row_floats + 1;
    if (aligned) {
        const uint32_t n_vec = total >> 2;
        for (uint32_t i = threadIdx.x; i < n_vec; i += blockDim.x) {
            const float4 v = ldg4(src + 4 * i);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t e = 4 * i + j;
                smem[(e / row_floats) * pitch + (e % row_floats)] = vv[j];
            }
        }
        for (uint32_t e = (n_vec << 2) + threadIdx.x; e < total; e += blockDim.x)
            smem[(e / row_floats) * pitch + (e % row_floats)] = __ldg(src + e);
    } else {
        for (uint32_t e = threadIdx.x; e < total; e += blockDim.x)
            smem[(e / row_floats) * pitch + (e % row_floats)] = __ldg(src + e);
    }
}

__global__ void __launch_bounds__(kShThreads)
    k_sh_fwd(const uint32_t n, const uint32_t K, const int degree, const float* __restrict__ dirs,
             const float* __restrict__ coeffs, const uint8_t* __restrict__ masks, float* __restrict__ colors,
             const int staged) {
    extern __shared__ float s_coef[];
    const uint32_t first = blockIdx.x * kShThreads;
    const uint32_t count = min((uint32_t)kShThreads, n - first);
    const uint32_t row = K * 3;
    if (staged) {
        stage_rows(coeffs + (size_t)first * row, count, row, s_coef);
        __syncthreads();
    }
    if (threadIdx.x >= count)
        return;
    const uint32_t e = first + threadIdx.x;
    if (masks && !masks[e])
        return;
    const f3 dir = mk3(__ldg(dirs + 3 * (size_t)e), __ldg(dirs + 3 * (size_t)e + 1), __ldg(dirs + 3 * (size_t)e + 2));
    f3 c;
    if (staged) {
        const float* r = s_coef + threadIdx.x * (row + 1);
        c = sh_to_color(degree, dir, [&](int k) { return mk3(r[3 * k], r[3 * k + 1], r[3 * k + 2]); });
    } else {
        const float* r = coeffs + (size_t)e * row;
        c = sh_to_color(degree, dir, [&](int k) { return mk3(__ldg(r + 3 * k), __ldg(r + 3 * k + 1), __ldg(r + 3 * k + 2)); });
    }
    colors[3 * (size_t)e] = c.x;
    colors[3 * (size_t)e + 1] = c.y;
    colors[3 * (size_t)e + 2] = c.z;
}

__global__ void __launch_bounds__(kShThreads)
    k_sh_bwd(const uint32_t n, const uint32_t K, const int degree, const float* __restrict__ dirs,
             const float* __restrict__ coeffs, const uint8_t* __restrict__ masks, const float* __restrict__ v_colors,
             float* __restrict__ v_coeffs, float* __restrict__ v_dirs) {
    const uint32_t e = blockIdx.x * kShThreads + threadIdx.x;
    if (e >= n)
        return;
    const uint32_t row = K * 3;
    float* vr = v_coeffs + (size_t)e * row;
    const int nb = (degree + 1) * (degree + 1);
    const bool live = !(masks && !masks[e]);
    // zero the inactive tail (and everything for masked rows): the reference relies on at::zeros_like
    for (uint32_t k = live ? (uint32_t)nb * 3 : 0; k < row; ++k)
        vr[k] = 0.f;
    if (!live) {
        if (v_dirs)
            v_dirs[3 * (size_t)e] = v_dirs[3 * (size_t)e + 1] = v_dirs[3 * (size_t)e + 2] = 0.f;
        return;
    }
    const f3 dir = mk3(__ldg(dirs + 3 * (size_t)e), __ldg(dirs + 3 * (size_t)e + 1), __ldg(dirs + 3 * (size_t)e + 2));
    const f3 vc = mk3(__ldg(v_colors + 3 * (size_t)e), __ldg(v_colors + 3 * (size_t)e + 1),
                      __ldg(v_colors + 3 * (size_t)e + 2));
    const float* r = coeffs + (size_t)e * row;
    const f3 vd = sh_vjp(
        degree, dir, vc, v_dirs != nullptr,
        [&](int k) { return mk3(__ldg(r + 3 * k), __ldg(r + 3 * k + 1), __ldg(r + 3 * k + 2)); },
        [&](int k, f3 g) {
            vr[3 * k] = g.x;
            vr[3 * k + 1] = g.y;
            vr[3 * k + 2] = g.z;
        });
    if (v_dirs) {
        v_dirs[3 * (size_t)e] = vd.x;
        v_dirs[3 * (size_t)e + 1] = vd.y;
        v_dirs[3 * (size_t)e + 2] = vd.z;
    }
}

} // namespace lfs

extern "C" int lfs_spherical_harmonics_fwd(uint32_t degrees_to_use, const float* dirs, const float* coeffs,
                                           const uint8_t* masks, uint32_t n, uint32_t K, float* colors,
                                           void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(dirs && coeffs && colors, "spherical_harmonics_fwd: null pointer");
    LFS_CHECK_ARG(degrees_to_use <= 4, "spherical_harmonics_fwd: degree %u > 4", degrees_to_use);
    LFS_CHECK_ARG((degrees_to_use + 1) * (degrees_to_use + 1) <= K, "spherical_harmonics_fwd: K=%u too small", K);
    if (n == 0)
        return LFS_OK;
    const int staged = K <= kShMaxStageK;
    const size_t smem = staged ? sizeof(float) * kShThreads * (K * 3 + 1) : 0;
    k_sh_fwd<<<div_up(n, kShThreads), kShThreads, smem, (cudaStream_t)stream>>>(n, K, (int)degrees_to_use, dirs,
                                                                                coeffs, masks, colors, staged);
    LFS_LAUNCH_OK("k_sh_fwd");
    return LFS_OK;
}

extern "C" int lfs_spherical_harmonics_bwd(uint32_t K, uint32_t degrees_to_use, const float* dirs,
                                           const float* coeffs, const uint8_t* masks, const float* v_colors,
                                           uint32_t n, float* v_coeffs, float* v_dirs, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(dirs && coeffs && v_colors && v_coeffs, "spherical_harmonics_bwd: null pointer");
    LFS_CHECK_ARG(degrees_to_use <= 4, "spherical_harmonics_bwd: degree %u > 4", degrees_to_use);
    LFS_CHECK_ARG((degrees_to_use + 1) * (degrees_to_use + 1) <= K, "spherical_harmonics_bwd: K=%u too small", K);
    if (n == 0)
        return LFS_OK;
    k_sh_bwd<<<div_up(n, kShThreads), kShThreads, 0, (cudaStream_t)stream>>>(n, K, (int)degrees_to_use, dirs, coeffs,
                                                                             masks, v_colors, v_coeffs, v_dirs);
    LFS_LAUNCH_OK("k_sh_bwd");
    return LFS_OK;
}
