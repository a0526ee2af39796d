// lfs_b200 -- real spherical harmonics (degree 0..4), evaluation and VJP for one direction.
// Same basis convention and recurrences (Sloan, JCGT 2013) as the reference's
// gsplat/SphericalHarmonicsCUDA.cu:21-371, re-organised: the three colour channels of one Gaussian are
// handled by ONE thread (the reference uses one thread per channel and re-derives the basis three
// times), and coefficients are reached through an accessor so AoS [n,K,3] and the trainer's planar SoA
// layout share the code.
#pragma once
#include "common.cuh"

namespace lfs {

struct ShBasis {
    float b[25];
};

// bases at unit direction (x,y,z); entries >= (degree+1)^2 are left untouched
__device__ __forceinline__ void sh_bases(const int degree, const float x, const float y, const float z, ShBasis& B) {
    B.b[0] = 0.2820947917738781f;
    if (degree < 1)
        return;
    B.b[1] = -0.48860251190292f * y;
    B.b[2] = 0.48860251190292f * z;
    B.b[3] = -0.48860251190292f * x;
    if (degree < 2)
        return;
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y;
    const float fS1 = 2.f * x * y;
    B.b[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    B.b[7] = fTmp0B * x;
    B.b[5] = fTmp0B * y;
    B.b[8] = 0.5462742152960395f * fC1;
    B.b[4] = 0.5462742152960395f * fS1;
    if (degree < 3)
        return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1;
    const float fS2 = x * fS1 + y * fC1;
    B.b[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    B.b[13] = fTmp0C * x;
    B.b[11] = fTmp0C * y;
    B.b[14] = fTmp1B * fC1;
    B.b[10] = fTmp1B * fS1;
    B.b[15] = -0.5900435899266435f * fC2;
    B.b[9] = -0.5900435899266435f * fS2;
    if (degree < 4)
        return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fC3 = x * fC2 - y * fS2;
    const float fS3 = x * fS2 + y * fC2;
    B.b[20] = 1.984313483298443f * z * B.b[12] - 1.006230589874905f * B.b[6];
    B.b[21] = fTmp0D * x;
    B.b[19] = fTmp0D * y;
    B.b[22] = fTmp1C * fC1;
    B.b[18] = fTmp1C * fS1;
    B.b[23] = fTmp2B * fC2;
    B.b[17] = fTmp2B * fS2;
    B.b[24] = 0.6258357354491763f * fC3;
    B.b[16] = 0.6258357354491763f * fS3;
}

// Gradient of sum_k s[k] * B_k(x,y,z) with respect to the unit direction (s[k] = <coeff_k, v_color>).
__device__ __forceinline__ f3 sh_bases_vjp(const int degree, const float x, const float y, const float z,
                                           const float* s /* [25] */) {
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (degree < 1)
        return mk3(0.f, 0.f, 0.f);
    vx += -0.48860251190292f * s[3];
    vy += -0.48860251190292f * s[1];
    vz += 0.48860251190292f * s[2];
    if (degree < 2)
        return mk3(vx, vy, vz);
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    const float fTmp0B_z = -1.092548430592079f;
    const float fC1_x = 2.f * x, fC1_y = -2.f * y, fS1_x = 2.f * y, fS1_y = 2.f * x;
    const float pSH6_z = 2.f * 0.9461746957575601f * z;
    vx += (0.5462742152960395f * fS1_x) * s[4] + (0.5462742152960395f * fC1_x) * s[8] + fTmp0B * s[7];
    vy += (0.5462742152960395f * fS1_y) * s[4] + (0.5462742152960395f * fC1_y) * s[8] + fTmp0B * s[5];
    vz += pSH6_z * s[6] + (fTmp0B_z * x) * s[7] + (fTmp0B_z * y) * s[5];
    if (degree < 3)
        return mk3(vx, vy, vz);
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    const float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    const float fTmp0C_z = -2.285228997322329f * 2.f * z;
    const float fTmp1B_z = 1.445305721320277f;
    const float fC2_x = fC1 + x * fC1_x - y * fS1_x;
    const float fC2_y = x * fC1_y - fS1 - y * fS1_y;
    const float fS2_x = fS1 + x * fS1_x + y * fC1_x;
    const float fS2_y = x * fS1_y + fC1 + y * fC1_y;
    const float pSH12_z = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    vx += (-0.5900435899266435f * fS2_x) * s[9] + (-0.5900435899266435f * fC2_x) * s[15] +
          (fTmp1B * fS1_x) * s[10] + (fTmp1B * fC1_x) * s[14] + fTmp0C * s[13];
    vy += (-0.5900435899266435f * fS2_y) * s[9] + (-0.5900435899266435f * fC2_y) * s[15] +
          (fTmp1B * fS1_y) * s[10] + (fTmp1B * fC1_y) * s[14] + fTmp0C * s[11];
    vz += pSH12_z * s[12] + (fTmp0C_z * x) * s[13] + (fTmp0C_z * y) * s[11] + (fTmp1B_z * fC1) * s[14] +
          (fTmp1B_z * fS1) * s[10];
    if (degree < 4)
        return mk3(vx, vy, vz);
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fTmp0D_z = 3.f * -4.683325804901025f * z2 + 2.007139630671868f;
    const float fTmp1C_z = 2.f * 3.31161143515146f * z;
    const float fTmp2B_z = -1.770130769779931f;
    const float fC3_x = fC2 + x * fC2_x - y * fS2_x;
    const float fC3_y = x * fC2_y - fS2 - y * fS2_y;
    const float fS3_x = fS2 + y * fC2_x + x * fS2_x;
    const float fS3_y = x * fS2_y + fC2 + y * fC2_y;
    const float pSH20_z = 1.984313483298443f * (pSH12 + z * pSH12_z) + -1.006230589874905f * pSH6_z;
    vx += (0.6258357354491763f * fS3_x) * s[16] + (0.6258357354491763f * fC3_x) * s[24] +
          (fTmp2B * fS2_x) * s[17] + (fTmp2B * fC2_x) * s[23] + (fTmp1C * fS1_x) * s[18] +
          (fTmp1C * fC1_x) * s[22] + fTmp0D * s[21];
    vy += (0.6258357354491763f * fS3_y) * s[16] + (0.6258357354491763f * fC3_y) * s[24] +
          (fTmp2B * fS2_y) * s[17] + (fTmp2B * fC2_y) * s[23] + (fTmp1C * fS1_y) * s[18] +
          (fTmp1C * fC1_y) * s[22] + fTmp0D * s[19];
    vz += pSH20_z * s[20] + (fTmp0D_z * x) * s[21] + (fTmp0D_z * y) * s[19] + (fTmp1C_z * fC1) * s[22] +
          (fTmp1C_z * fS1) * s[18] + (fTmp2B_z * fC2) * s[23] + (fTmp2B_z * fS2) * s[17];
    return mk3(vx, vy, vz);
}

// colour = sum_k B_k(dir/|dir|) * coef(k)
template <class CoefF>
__device__ __forceinline__ f3 sh_to_color(const int degree, const f3 dir, CoefF coef) {
    float x = 0.f, y = 0.f, z = 0.f;
    if (degree >= 1) {
        const float inorm = rsqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        x = dir.x * inorm, y = dir.y * inorm, z = dir.z * inorm;
    }
    ShBasis B;
    sh_bases(degree, x, y, z, B);
    const int nb = (degree + 1) * (degree + 1);
    f3 acc = mk3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 25; ++k) {
        if (k < nb) {
            const f3 c = coef(k);
            acc.x += B.b[k] * c.x;
            acc.y += B.b[k] * c.y;
            acc.z += B.b[k] * c.z;
        }
    }
    return acc;
}

// VJP: emits v_coeff(k) = B_k * v_color through `emit(k, f3)` for k < (degree+1)^2 and returns dL/d(dir)
// (zero when !need_vdir or degree == 0).
template <class CoefF, class EmitF>
__device__ __forceinline__ f3 sh_vjp(const int degree, const f3 dir, const f3 v_color, const bool need_vdir,
                                     CoefF coef, EmitF emit) {
    float x = 0.f, y = 0.f, z = 0.f, inorm = 0.f;
    if (degree >= 1) {
        inorm = rsqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        x = dir.x * inorm, y = dir.y * inorm, z = dir.z * inorm;
    }
    ShBasis B;
    sh_bases(degree, x, y, z, B);
    const int nb = (degree + 1) * (degree + 1);
    float s[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) {
        s[k] = 0.f;
        if (k < nb) {
            emit(k, mk3(B.b[k] * v_color.x, B.b[k] * v_color.y, B.b[k] * v_color.z));
            if (need_vdir && k >= 1) {
                const f3 c = coef(k);
                s[k] = c.x * v_color.x + c.y * v_color.y + c.z * v_color.z;
            }
        }
    }
    if (!need_vdir || degree < 1)
        return mk3(0.f, 0.f, 0.f);
    const f3 vdn = sh_bases_vjp(degree, x, y, z, s);
    const float dt = vdn.x * x + vdn.y * y + vdn.z * z;
    return mk3((vdn.x - dt * x) * inorm, (vdn.y - dt * y) * inorm, (vdn.z - dt * z) * inorm);
}

} // namespace lfs
