// lfs_b200 -- lfs_projection_ut_3dgs_fused: drop-in for gsplat::projection_ut_3dgs_fused
// (reference gsplat/Projection.cpp:22-110, kernel gsplat/ProjectionUT3DGSFused.cu:17-203).
//
// HBM plan (per camera-Gaussian pair): 44 B in (mean 12, quat 16, scale 12, opacity 4), 32 B out for
// visible pairs.  means / scales are AoS float3: a block stages its 256 records with 128-bit coalesced
// loads into shared memory (3072 B per array) and each thread then reads its own 3 floats (stride 3 is
// conflict-free); quats are loaded as float4 directly.
#include "projection.cuh"

namespace lfs {

constexpr int kProjThreads = 256;

// Stage `count` float3 records starting at record index `first` into smem (count <= blockDim.x).
// `base` must be 16-byte aligned; first*3 floats is a multiple of 4 when first % 4 == 0 (true: first is a
// multiple of the block size).
__device__ __forceinline__ void stage_float3(const float* __restrict__ base, uint32_t first, uint32_t count,
                                             uint32_t total, float* smem) {
    const uint32_t n_floats = count * 3;
    const float* src = base + (size_t)first * 3;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    if (aligned) {
        const uint32_t n_vec = n_floats >> 2;
        for (uint32_t i = threadIdx.x; i < n_vec; i += blockDim.x)
            reinterpret_cast<float4*>(smem)[i] = ldg4(src + 4 * i);
        for (uint32_t i = (n_vec << 2) + threadIdx.x; i < n_floats; i += blockDim.x)
            smem[i] = __ldg(src + i);
    } else {
        for (uint32_t i = threadIdx.x; i < n_floats; i += blockDim.x)
            smem[i] = __ldg(src + i);
    }
    (void)total;
}

__global__ void __launch_bounds__(kProjThreads)
    k_projection_ut(const uint32_t C, const uint32_t N, const float* __restrict__ means,
                    const float* __restrict__ quats, const float* __restrict__ scales,
                    const float* __restrict__ opacities, const float* __restrict__ viewmats,
                    const float* __restrict__ Ks, const uint32_t width, const uint32_t height, const float eps2d,
                    const float near_plane, const float far_plane, const float radius_clip,
                    const lfs_ut_params ut, const int camera_model, const int rs_type,
                    const float* __restrict__ viewmats1, const float* __restrict__ radial,
                    const float* __restrict__ tangential, const float* __restrict__ prism,
                    int32_t* __restrict__ radii, float* __restrict__ means2d, float* __restrict__ depths,
                    float* __restrict__ conics, float* __restrict__ compensations) {
    __shared__ __align__(16) float s_mean[kProjThreads * 3];
    __shared__ __align__(16) float s_scale[kProjThreads * 3];
    // grid: x over Gaussian blocks, y over cameras
    const uint32_t cid = blockIdx.y;
    const uint32_t first = blockIdx.x * kProjThreads;
    const uint32_t count = min((uint32_t)kProjThreads, N - first);
    stage_float3(means, first, count, N, s_mean);
    stage_float3(scales, first, count, N, s_scale);
    __syncthreads();
    if (threadIdx.x >= count)
        return;
    const uint32_t gid = first + threadIdx.x;
    const size_t idx = (size_t)cid * N + gid;

    const ViewCam cam = make_viewcam(viewmats + 16 * cid, Ks + 9 * cid, (int)width, (int)height);
    const f3 mean = mk3(s_mean[3 * threadIdx.x], s_mean[3 * threadIdx.x + 1], s_mean[3 * threadIdx.x + 2]);
    const f3 scale = mk3(s_scale[3 * threadIdx.x], s_scale[3 * threadIdx.x + 1], s_scale[3 * threadIdx.x + 2]);
    const float4 q = ldg4(quats + 4 * (size_t)gid);
    const float op = opacities ? __ldg(opacities + gid) : 0.f;

    // perfect pinhole + global shutter keeps the specialised routine; everything else goes through the camera model
    const bool general = camera_model != LFS_PINHOLE || rs_type != kRsGlobal || viewmats1 || radial || tangential || prism;
    UTOut o;
    if (general) {
        const int n_rad = camera_model == LFS_FISHEYE ? 4 : 6;
        const CamModel cmod = make_cam_model(viewmats + 16 * cid, viewmats1 ? viewmats1 + 16 * cid : nullptr, Ks + 9 * cid,
                                             width, height, camera_model, rs_type, radial ? radial + n_rad * cid : nullptr,
                                             tangential ? tangential + 2 * cid : nullptr, prism ? prism + 4 * cid : nullptr);
        o = ut_project_general(cmod, mean, q, scale, opacities != nullptr, op, eps2d, near_plane, far_plane, radius_clip, ut);
    } else {
        o = ut_project_pinhole(cam, mean, q, scale, opacities != nullptr, op, eps2d, near_plane, far_plane, radius_clip, ut);
    }
    int2 r = make_int2(0, 0);
    if (o.ok) {
        r = make_int2((int32_t)o.rx, (int32_t)o.ry);
        reinterpret_cast<float2*>(means2d)[idx] = make_float2(o.mx, o.my);
        depths[idx] = o.depth;
        conics[idx * 3 + 0] = o.c00;
        conics[idx * 3 + 1] = o.c01;
        conics[idx * 3 + 2] = o.c11;
        if (compensations)
            compensations[idx] = o.comp;
    }
    reinterpret_cast<int2*>(radii)[idx] = r;
}

} // namespace lfs

extern "C" int lfs_projection_ut_3dgs_fused(const float* means, const float* quats, const float* scales,
                                            const float* opacities, const float* viewmats0,
                                            const float* viewmats1, const float* Ks, uint32_t N, uint32_t C,
                                            uint32_t image_width, uint32_t image_height, float eps2d,
                                            float near_plane, float far_plane, float radius_clip, int camera_model,
                                            const lfs_ut_params* ut_params, int rs_type, const float* radial_coeffs,
                                            const float* tangential_coeffs, const float* thin_prism_coeffs,
                                            int32_t* radii, float* means2d, float* depths, float* conics,
                                            float* compensations, void* stream) {
    using namespace lfs;
    LFS_CHECK_ARG(means && quats && scales && viewmats0 && Ks && radii && means2d && depths && conics,
                  "projection_ut: null required pointer");
    LFS_CHECK_ARG(ut_params != nullptr, "projection_ut: ut_params is null");
    // the reference's UT projection has no ORTHO branch either (ProjectionUT3DGSFused.cu:84-140 asserts)
    LFS_UNSUPPORTED(camera_model != LFS_PINHOLE && camera_model != LFS_FISHEYE,
                    "projection_ut: camera model %d is not supported by the unscented-transform projection", camera_model);
    LFS_CHECK_ARG(rs_type >= 0 && rs_type <= LFS_GLOBAL, "projection_ut: bad shutter type %d", rs_type);
    LFS_CHECK_ARG(camera_model != LFS_FISHEYE || (!tangential_coeffs && !thin_prism_coeffs),
                  "projection_ut: the fisheye model takes radial coefficients only");
    if (N == 0 || C == 0)
        return LFS_OK;
    LFS_CHECK_ARG(C <= 65535, "projection_ut: C too large");
    dim3 grid(div_up(N, kProjThreads), C);
    k_projection_ut<<<grid, kProjThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, quats, scales, opacities, viewmats0, Ks, image_width, image_height, eps2d, near_plane, far_plane,
        radius_clip, *ut_params, camera_model, rs_type, viewmats1, radial_coeffs, tangential_coeffs, thin_prism_coeffs, radii,
        means2d, depths, conics, compensations);
    LFS_LAUNCH_OK("k_projection_ut");
    return LFS_OK;
}
