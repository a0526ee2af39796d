// lfs_b200 -- the three small per-Gaussian ops the reference's densification strategies call through gsplat/Ops.h
// (off the steady-state hot path, kept so that the whole operator surface links):
//   lfs_quats_to_rotmats  <- gsplat::quats_to_rotmats (gsplat/QuatToRotmat.cpp, kernel QuatToRotmatCUDA.cu:14-45)
//   lfs_relocation        <- gsplat::relocation       (gsplat/Relocation.cpp:12-29, kernel RelocationCUDA.cu:12-40)
//   lfs_add_noise         <- gsplat::add_noise        (gsplat/Relocation.cpp:31-50, kernel RelocationCUDA.cu:113-145)
#include "common.cuh"

namespace lfs {

__device__ __forceinline__ void rot_from_quat(float w, float x, float y, float z, float inv_norm, float R[9]) {
    x *= inv_norm, y *= inv_norm, z *= inv_norm, w *= inv_norm;
    const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y,
                wz = w * z;
    R[0] = 1.f - 2.f * (y2 + z2), R[1] = 2.f * (xy - wz), R[2] = 2.f * (xz + wy);
    R[3] = 2.f * (xy + wz), R[4] = 1.f - 2.f * (x2 + z2), R[5] = 2.f * (yz - wx);
    R[6] = 2.f * (xz - wy), R[7] = 2.f * (yz + wx), R[8] = 1.f - 2.f * (x2 + y2);
}

__global__ void __launch_bounds__(256)
    k_quats_to_rotmats(const uint32_t n, const float* __restrict__ quats, float* __restrict__ rotmats) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n)
        return;
    const float4 q = ldg4(quats + 4 * (size_t)i);
    float R[9];
    rot_from_quat(q.x, q.y, q.z, q.w, rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), R);
#pragma unroll
    for (int k = 0; k < 9; ++k)
        rotmats[9 * (size_t)i + k] = R[k]; // row-major [N,3,3]
}

// "3D Gaussian Splatting as Markov Chain Monte Carlo", eq. (9)
__global__ void __launch_bounds__(256)
    k_relocation(const uint32_t n, const float* __restrict__ opacities, const float* __restrict__ scales,
                 const int32_t* __restrict__ ratios, const float* __restrict__ binoms, const int n_max,
                 float* __restrict__ new_opacities, float* __restrict__ new_scales) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n)
        return;
    const int n_idx = ratios[i];
    const float op = opacities[i];
    const float nop = 1.0f - powf(1.0f - op, 1.0f / (float)n_idx);
    new_opacities[i] = nop;
    float denom = 0.f;
    for (int a = 1; a <= n_idx; ++a) {
        float pw = nop; // nop^(k+1)
        for (int k = 0; k <= a - 1; ++k) {
            const float sign = (k & 1) ? -1.f : 1.f;
            denom += binoms[(a - 1) * n_max + k] * (sign * rsqrtf((float)(k + 1))) * pw;
            pw *= nop;
        }
    }
    const float coeff = op / denom;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        new_scales[3 * (size_t)i + c] = coeff * scales[3 * (size_t)i + c];
}

__global__ void __launch_bounds__(256)
    k_add_noise(const uint32_t n, const float* __restrict__ raw_opacities, const float* __restrict__ raw_scales,
                const float* __restrict__ raw_quats, const float* __restrict__ noise, float* __restrict__ means,
                const float current_lr) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n)
        return;
    const float4 q = ldg4(raw_quats + 4 * (size_t)i);
    float R[9];
    // torch-normalize semantics of the reference: 1/|q| capped at 1e12 (RelocationCUDA.cu:84-86)
    rot_from_quat(q.x, q.y, q.z, q.w, fminf(rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e12f), R);
    const float s2[3] = {__expf(2.f * raw_scales[3 * (size_t)i]), __expf(2.f * raw_scales[3 * (size_t)i + 1]),
                         __expf(2.f * raw_scales[3 * (size_t)i + 2])};
    const float nz[3] = {noise[3 * (size_t)i], noise[3 * (size_t)i + 1], noise[3 * (size_t)i + 2]};
    // covariance * noise = R S^2 R^T noise
    float t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        t[a] = s2[a] * (R[0 * 3 + a] * nz[0] + R[1 * 3 + a] * nz[1] + R[2 * 3 + a] * nz[2]);
    const float opacity = 1.0f / (1.0f + __expf(-raw_opacities[i]));
    const float factor = current_lr * (1.0f / (1.0f + __expf(100.f * opacity - 0.5f)));
#pragma unroll
    for (int r = 0; r < 3; ++r)
        means[3 * (size_t)i + r] += factor * (R[r * 3 + 0] * t[0] + R[r * 3 + 1] * t[1] + R[r * 3 + 2] * t[2]);
}

} // namespace lfs

extern "C" int lfs_quats_to_rotmats(const float* quats, uint32_t n, float* rotmats, void* stream) {
    using namespace lfs;
    if (n == 0)
        return LFS_OK;
    LFS_CHECK_ARG(quats && rotmats, "quats_to_rotmats: null pointer");
    k_quats_to_rotmats<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, quats, rotmats);
    LFS_LAUNCH_OK("k_quats_to_rotmats");
    return LFS_OK;
}

extern "C" int lfs_relocation(const float* opacities, const float* scales, const int32_t* ratios, const float* binoms,
                              int n_max, uint32_t n, float* new_opacities, float* new_scales, void* stream) {
    using namespace lfs;
    if (n == 0)
        return LFS_OK;
    LFS_CHECK_ARG(opacities && scales && ratios && binoms && new_opacities && new_scales && n_max > 0,
                  "relocation: null pointer / bad n_max");
    k_relocation<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, opacities, scales, ratios, binoms, n_max,
                                                                   new_opacities, new_scales);
    LFS_LAUNCH_OK("k_relocation");
    return LFS_OK;
}

extern "C" int lfs_add_noise(const float* raw_opacities, const float* raw_scales, const float* raw_quats,
                             const float* noise, float* means, float current_lr, uint32_t n, void* stream) {
    using namespace lfs;
    if (n == 0)
        return LFS_OK;
    LFS_CHECK_ARG(raw_opacities && raw_scales && raw_quats && noise && means, "add_noise: null pointer");
    k_add_noise<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, raw_opacities, raw_scales, raw_quats, noise, means,
                                                                  current_lr);
    LFS_LAUNCH_OK("k_add_noise");
    return LFS_OK;
}
