// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" glue around the UNMODIFIED reference gsplat backend (/root/reference/gsplat/*.cu,*.cpp,
// compiled in place against oracle/glm_shim by oracle/Makefile into oracle/_ref/libgsplat_ref.so).
// Raw device pointers are wrapped with at::from_blob and passed to the reference's public operators
// (gsplat/Ops.h:12-166); results are copied into caller-provided device buffers.
// Only runs where a GPU is present (the reference ops CHECK_CUDA their inputs).
#include "Ops.h"

#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

namespace {
    at::Tensor wrapf(const float* p, at::IntArrayRef sizes) {
        return at::from_blob(const_cast<float*>(p), sizes, at::TensorOptions().dtype(at::kFloat).device(at::kCUDA));
    }
    at::Tensor wrapi(const int32_t* p, at::IntArrayRef sizes) {
        return at::from_blob(const_cast<int32_t*>(p), sizes, at::TensorOptions().dtype(at::kInt).device(at::kCUDA));
    }
    at::Tensor wrapl(const int64_t* p, at::IntArrayRef sizes) {
        return at::from_blob(const_cast<int64_t*>(p), sizes, at::TensorOptions().dtype(at::kLong).device(at::kCUDA));
    }
    at::Tensor wrapb(const bool* p, at::IntArrayRef sizes) {
        return at::from_blob(const_cast<bool*>(p), sizes, at::TensorOptions().dtype(at::kBool).device(at::kCUDA));
    }
    void copy_out(void* dst, const at::Tensor& t, int64_t max_elems = -1) {
        if (!dst || !t.defined() || t.numel() == 0)
            return;
        at::Tensor c = t.contiguous();
        int64_t n = c.numel();
        if (max_elems >= 0 && n > max_elems)
            n = max_elems;
        cudaMemcpy(dst, c.data_ptr(), n * c.element_size(), cudaMemcpyDeviceToDevice);
    }
} // namespace

extern "C" {

// gsplat::projection_ut_3dgs_fused (gsplat/Ops.h:69-98), PINHOLE / GLOBAL shutter, no distortion.
int ref_gsplat_projection_ut(const float* means, const float* quats, const float* scales, const float* opacities,
                             const float* viewmats, const float* Ks, int N, int C, int width, int height, float eps2d,
                             float near_plane, float far_plane, float radius_clip, int calc_compensations,
                             int32_t* radii, float* means2d, float* depths, float* conics, float* compensations) {
    try {
        auto r = gsplat::projection_ut_3dgs_fused(
            wrapf(means, {N, 3}), wrapf(quats, {N, 4}), wrapf(scales, {N, 3}),
            opacities ? at::optional<at::Tensor>(wrapf(opacities, {N})) : at::nullopt, wrapf(viewmats, {C, 4, 4}),
            at::nullopt, wrapf(Ks, {C, 3, 3}), width, height, eps2d, near_plane, far_plane, radius_clip,
            calc_compensations != 0, gsplat::CameraModelType::PINHOLE, UnscentedTransformParameters{},
            ShutterType::GLOBAL, at::nullopt, at::nullopt, at::nullopt);
        copy_out(radii, std::get<0>(r));
        copy_out(means2d, std::get<1>(r));
        copy_out(depths, std::get<2>(r));
        copy_out(conics, std::get<3>(r));
        if (calc_compensations)
            copy_out(compensations, std::get<4>(r));
        cudaDeviceSynchronize();
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_gsplat] %s\n", e.what());
        return -1;
    }
}

int ref_gsplat_sh_fwd(int degree, const float* dirs, const float* coeffs, const bool* masks, int n, int K,
                      float* colors) {
    try {
        // the reference leaves masked-out rows uninitialised (at::empty_like): zero the output first
        cudaMemset(colors, 0, sizeof(float) * 3 * (size_t)n);
        auto r = gsplat::spherical_harmonics_fwd(degree, wrapf(dirs, {n, 3}), wrapf(coeffs, {n, K, 3}),
                                                 masks ? at::optional<at::Tensor>(wrapb(masks, {n})) : at::nullopt);
        if (masks) {
            at::Tensor m = wrapb(masks, {n}).unsqueeze(-1);
            r = at::where(m, r, at::zeros_like(r));
        }
        copy_out(colors, r);
        cudaDeviceSynchronize();
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_gsplat] %s\n", e.what());
        return -1;
    }
}

int ref_gsplat_sh_bwd(int degree, const float* dirs, const float* coeffs, const bool* masks, const float* v_colors,
                      int n, int K, int compute_v_dirs, float* v_coeffs, float* v_dirs) {
    try {
        auto r = gsplat::spherical_harmonics_bwd(K, degree, wrapf(dirs, {n, 3}), wrapf(coeffs, {n, K, 3}),
                                                 masks ? at::optional<at::Tensor>(wrapb(masks, {n})) : at::nullopt,
                                                 wrapf(v_colors, {n, 3}), compute_v_dirs != 0);
        copy_out(v_coeffs, std::get<0>(r));
        if (compute_v_dirs)
            copy_out(v_dirs, std::get<1>(r));
        cudaDeviceSynchronize();
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_gsplat] %s\n", e.what());
        return -1;
    }
}

// Returns n_isects (or <0 on error). isect_ids / flatten_ids receive at most `capacity` entries.
long long ref_gsplat_intersect_tile(const float* means2d, const int32_t* radii, const float* depths, int C, int N,
                                    int tile_size, int tile_width, int tile_height, int sort,
                                    int32_t* tiles_per_gauss, int64_t* isect_ids, int32_t* flatten_ids,
                                    long long capacity) {
    try {
        auto r = gsplat::intersect_tile(wrapf(means2d, {C, N, 2}), wrapi(radii, {C, N, 2}), wrapf(depths, {C, N}),
                                        at::nullopt, at::nullopt, C, tile_size, tile_width, tile_height, sort != 0);
        copy_out(tiles_per_gauss, std::get<0>(r));
        copy_out(isect_ids, std::get<1>(r), capacity);
        copy_out(flatten_ids, std::get<2>(r), capacity);
        cudaDeviceSynchronize();
        return std::get<1>(r).numel();
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_gsplat] %s\n", e.what());
        return -1;
    }
}

int ref_gsplat_intersect_offset(const int64_t* isect_ids, long long n_isects, int C, int tile_width, int tile_height,
                                int32_t* offsets) {
    try {
        auto r = gsplat::intersect_offset(wrapl(isect_ids, {n_isects}), C, tile_width, tile_height);
        copy_out(offsets, r);
        cudaDeviceSynchronize();
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_gsplat] %s\n", e.what());
        return -1;
    }
}

int ref_gsplat_raster_fwd(const float* means, const float* quats, const float* scales, const float* colors,
                          const float* opacities, const float* backgrounds, int N, int C, int channels, int width,
                          int height, int tile_size, const float* viewmats, const float* Ks,
                          const int32_t* tile_offsets, const int32_t* flatten_ids, long long n_isects,
                          float* renders, float* alphas, int32_t* last_ids) {
    try {
        int th = (height + tile_size - 1) / tile_size, tw = (width + tile_size - 1) / tile_size;
        auto r = gsplat::rasterize_to_pixels_from_world_3dgs_fwd(
            wrapf(means, {N, 3}), wrapf(quats, {N, 4}), wrapf(scales, {N, 3}), wrapf(colors, {C, N, channels}),
            wrapf(opacities, {C, N}),
            backgrounds ? at::optional<at::Tensor>(wrapf(backgrounds, {C, channels})) : at::nullopt, at::nullopt,
            width, height, tile_size, wrapf(viewmats, {C, 4, 4}), at::nullopt, wrapf(Ks, {C, 3, 3}),
            gsplat::CameraModelType::PINHOLE, UnscentedTransformParameters{}, ShutterType::GLOBAL, at::nullopt,
            at::nullopt, at::nullopt, wrapi(tile_offsets, {C, th, tw}), wrapi(flatten_ids, {n_isects}));
        copy_out(renders, std::get<0>(r));
        copy_out(alphas, std::get<1>(r));
        copy_out(last_ids, std::get<2>(r));
        cudaDeviceSynchronize();
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_gsplat] %s\n", e.what());
        return -1;
    }
}

int ref_gsplat_raster_bwd(const float* means, const float* quats, const float* scales, const float* colors,
                          const float* opacities, const float* backgrounds, int N, int C, int width, int height,
                          int tile_size, const float* viewmats, const float* Ks, const int32_t* tile_offsets,
                          const int32_t* flatten_ids, long long n_isects, const float* render_alphas,
                          const int32_t* last_ids, const float* v_render_colors, const float* v_render_alphas,
                          float* v_means, float* v_quats, float* v_scales, float* v_colors, float* v_opacities) {
    try {
        int th = (height + tile_size - 1) / tile_size, tw = (width + tile_size - 1) / tile_size;
        auto r = gsplat::rasterize_to_pixels_from_world_3dgs_bwd(
            wrapf(means, {N, 3}), wrapf(quats, {N, 4}), wrapf(scales, {N, 3}), wrapf(colors, {C, N, 3}),
            wrapf(opacities, {C, N}), backgrounds ? at::optional<at::Tensor>(wrapf(backgrounds, {C, 3})) : at::nullopt,
            at::nullopt, width, height, tile_size, wrapf(viewmats, {C, 4, 4}), at::nullopt, wrapf(Ks, {C, 3, 3}),
            gsplat::CameraModelType::PINHOLE, UnscentedTransformParameters{}, ShutterType::GLOBAL, at::nullopt,
            at::nullopt, at::nullopt, wrapi(tile_offsets, {C, th, tw}), wrapi(flatten_ids, {n_isects}),
            wrapf(render_alphas, {C, height, width, 1}), wrapi(last_ids, {C, height, width}),
            wrapf(v_render_colors, {C, height, width, 3}), wrapf(v_render_alphas, {C, height, width, 1}));
        copy_out(v_means, std::get<0>(r));
        copy_out(v_quats, std::get<1>(r));
        copy_out(v_scales, std::get<2>(r));
        copy_out(v_colors, std::get<3>(r));
        copy_out(v_opacities, std::get<4>(r));
        cudaDeviceSynchronize();
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_gsplat] %s\n", e.what());
        return -1;
    }
}

} // extern "C"
