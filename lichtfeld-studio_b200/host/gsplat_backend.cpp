// lfs_b200 -- drop-in `gsplat_backend`: the reference's C++ operator surface (namespace gsplat, gsplat/Ops.h:9-168)
// re-exported with IDENTICAL signatures on top of the C ABI of liblfs_b200.so (include/lfs_b200.h).
//
// Build inside the reference tree in place of gsplat/*.cu,*.cpp: this TU includes the reference's
// own "Ops.h" (declarations only: Ops.h -> Cameras.h, Common.h) so the mangled names are the reference's by
// construction, and src/training/rasterization/*.cpp, strategies/*.cpp and the tests link against it unchanged.
// It is thin host dispatch: CHECK_INPUT exactly like the reference (gsplat/Common.h:12-17), allocate outputs with
// ATen, forward raw pointers + the current CUDA stream to the C ABI.  No kernels here.
#include "Ops.h"

#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <vector>

#include "lfs_b200.h"

namespace {

void lfs_ok(int rc, const char* what) {
    TORCH_CHECK(rc == LFS_OK, what, " failed: ", lfs_last_error(), " (lfs status ", rc, ")");
}
void* cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
const float* fp(const at::Tensor& t) { return t.data_ptr<float>(); }
const float* fpo(const at::optional<at::Tensor>& t) { return t.has_value() ? t.value().data_ptr<float>() : nullptr; }
const uint8_t* maskp(const at::optional<at::Tensor>& t) {
    return t.has_value() ? reinterpret_cast<const uint8_t*>(t.value().data_ptr<bool>()) : nullptr;
}
lfs_ut_params to_ut(const UnscentedTransformParameters& p) {
    return lfs_ut_params{p.alpha, p.beta, p.kappa, p.in_image_margin_factor, p.require_all_sigma_points_valid ? 1 : 0};
}
// lfs_alloc_fn backed by the CUDA caching allocator: scratch lives until the op returns, outputs are handed back
struct Alloc {
    at::Device dev;
    std::vector<at::Tensor> keep;
    at::Tensor tagged[3];
    static void* fn(void* ctx, int tag, size_t bytes) {
        auto* self = static_cast<Alloc*>(ctx);
        try {
            at::Tensor t = at::empty({(int64_t)bytes + 256}, at::TensorOptions().dtype(at::kByte).device(self->dev));
            self->keep.push_back(t);
            const size_t off = (256 - reinterpret_cast<uintptr_t>(t.data_ptr()) % 256) % 256;
            if (tag >= 0 && tag < 3)
                self->tagged[tag] = t.narrow(0, (int64_t)off, (int64_t)bytes);
            return static_cast<char*>(t.data_ptr()) + off;
        } catch (...) {
            return nullptr;
        }
    }
};

} // namespace

namespace gsplat {

at::Tensor spherical_harmonics_fwd(const uint32_t degrees_to_use, const at::Tensor dirs, const at::Tensor coeffs,
                                   const at::optional<at::Tensor> masks) {
    DEVICE_GUARD(dirs);
    CHECK_INPUT(dirs);
    CHECK_INPUT(coeffs);
    if (masks.has_value()) {
        CHECK_INPUT(masks.value());
    }
    TORCH_CHECK(coeffs.size(-1) == 3, "coeffs must have last dimension 3");
    TORCH_CHECK(dirs.size(-1) == 3, "dirs must have last dimension 3");
    at::Tensor colors = at::empty_like(dirs);
    lfs_ok(lfs_spherical_harmonics_fwd(degrees_to_use, fp(dirs), fp(coeffs), maskp(masks), (uint32_t)(dirs.numel() / 3),
                                       (uint32_t)coeffs.size(-2), colors.data_ptr<float>(), cur_stream()),
           "spherical_harmonics_fwd");
    return colors;
}

std::tuple<at::Tensor, at::Tensor> spherical_harmonics_bwd(const uint32_t K, const uint32_t degrees_to_use,
                                                           const at::Tensor dirs, const at::Tensor coeffs,
                                                           const at::optional<at::Tensor> masks,
                                                           const at::Tensor v_colors, bool compute_v_dirs) {
    DEVICE_GUARD(dirs);
    CHECK_INPUT(dirs);
    CHECK_INPUT(coeffs);
    CHECK_INPUT(v_colors);
    if (masks.has_value()) {
        CHECK_INPUT(masks.value());
    }
    at::Tensor v_coeffs = at::empty_like(coeffs);
    at::Tensor v_dirs;
    if (compute_v_dirs)
        v_dirs = at::empty_like(dirs);
    lfs_ok(lfs_spherical_harmonics_bwd(K, degrees_to_use, fp(dirs), fp(coeffs), maskp(masks), fp(v_colors),
                                       (uint32_t)(dirs.numel() / 3), v_coeffs.data_ptr<float>(),
                                       compute_v_dirs ? v_dirs.data_ptr<float>() : nullptr, cur_stream()),
           "spherical_harmonics_bwd");
    return std::make_tuple(v_coeffs, v_dirs);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> intersect_tile(const at::Tensor means2d, const at::Tensor radii,
                                                              const at::Tensor depths,
                                                              const at::optional<at::Tensor> camera_ids,
                                                              const at::optional<at::Tensor> gaussian_ids,
                                                              const uint32_t C, const uint32_t tile_size,
                                                              const uint32_t tile_width, const uint32_t tile_height,
                                                              const bool sort) {
    DEVICE_GUARD(means2d);
    CHECK_INPUT(means2d);
    CHECK_INPUT(radii);
    CHECK_INPUT(depths);
    const bool packed = means2d.dim() == 2; // gsplat/Intersect.cpp:32-39
    if (packed) {
        TORCH_CHECK(camera_ids.has_value() && gaussian_ids.has_value(),
                    "When packed is set, camera_ids and gaussian_ids must be provided.");
        CHECK_INPUT(camera_ids.value());
        CHECK_INPUT(gaussian_ids.value());
    }
    at::Tensor tiles_per_gauss = at::empty_like(depths, depths.options().dtype(at::kInt));
    Alloc al{means2d.device(), {}, {}};
    int64_t* ids = nullptr;
    int32_t* flat = nullptr;
    int64_t n_isects = 0;
    if (packed)
        lfs_ok(lfs_intersect_tile_packed(fp(means2d), radii.data_ptr<int32_t>(), fp(depths),
                                         camera_ids.value().data_ptr<int64_t>(), (uint32_t)means2d.size(0), C, tile_size,
                                         tile_width, tile_height, sort ? 1 : 0, tiles_per_gauss.data_ptr<int32_t>(),
                                         &Alloc::fn, &al, &ids, &flat, &n_isects, cur_stream()),
               "intersect_tile (packed)");
    else
        lfs_ok(lfs_intersect_tile(fp(means2d), radii.data_ptr<int32_t>(), fp(depths), C, (uint32_t)means2d.size(1),
                                  tile_size, tile_width, tile_height, sort ? 1 : 0, tiles_per_gauss.data_ptr<int32_t>(),
                                  &Alloc::fn, &al, &ids, &flat, &n_isects, cur_stream()),
               "intersect_tile");
    at::Tensor isect_ids, flatten_ids;
    if (n_isects > 0) {
        isect_ids = al.tagged[LFS_TAG_ISECT_IDS].view(at::kLong);
        flatten_ids = al.tagged[LFS_TAG_FLATTEN_IDS].view(at::kInt);
    } else {
        isect_ids = at::empty({0}, depths.options().dtype(at::kLong));
        flatten_ids = at::empty({0}, depths.options().dtype(at::kInt));
    }
    return std::make_tuple(tiles_per_gauss, isect_ids, flatten_ids);
}

at::Tensor intersect_offset(const at::Tensor isect_ids, const uint32_t C, const uint32_t tile_width,
                            const uint32_t tile_height) {
    DEVICE_GUARD(isect_ids);
    CHECK_INPUT(isect_ids);
    at::Tensor offsets = at::empty({C, tile_height, tile_width}, isect_ids.options().dtype(at::kInt));
    lfs_ok(lfs_intersect_offset(isect_ids.data_ptr<int64_t>(), isect_ids.numel(), C, tile_width, tile_height,
                                offsets.data_ptr<int32_t>(), cur_stream()),
           "intersect_offset");
    return offsets;
}

at::Tensor quats_to_rotmats(const at::Tensor quats) {
    DEVICE_GUARD(quats);
    CHECK_INPUT(quats);
    at::Tensor rotmats = at::empty({quats.size(0), 3, 3}, quats.options());
    lfs_ok(lfs_quats_to_rotmats(fp(quats), (uint32_t)quats.size(0), rotmats.data_ptr<float>(), cur_stream()),
           "quats_to_rotmats");
    return rotmats;
}

std::tuple<at::Tensor, at::Tensor> relocation(at::Tensor opacities, at::Tensor scales, at::Tensor ratios,
                                              at::Tensor binoms, const int n_max) {
    DEVICE_GUARD(opacities);
    CHECK_INPUT(opacities);
    CHECK_INPUT(scales);
    CHECK_INPUT(ratios);
    CHECK_INPUT(binoms);
    at::Tensor new_opacities = at::empty_like(opacities);
    at::Tensor new_scales = at::empty_like(scales);
    lfs_ok(lfs_relocation(fp(opacities), fp(scales), ratios.data_ptr<int32_t>(), fp(binoms), n_max,
                          (uint32_t)opacities.size(0), new_opacities.data_ptr<float>(), new_scales.data_ptr<float>(),
                          cur_stream()),
           "relocation");
    return std::make_tuple(new_opacities, new_scales);
}

void add_noise(at::Tensor raw_opacities, at::Tensor raw_scales, at::Tensor raw_quats, at::Tensor noise,
               at::Tensor means, const float current_lr) {
    DEVICE_GUARD(raw_opacities);
    CHECK_INPUT(raw_opacities);
    CHECK_INPUT(raw_scales);
    CHECK_INPUT(raw_quats);
    CHECK_INPUT(noise);
    CHECK_INPUT(means);
    lfs_ok(lfs_add_noise(fp(raw_opacities), fp(raw_scales), fp(raw_quats), fp(noise), means.data_ptr<float>(),
                         current_lr, (uint32_t)raw_opacities.size(0), cur_stream()),
           "add_noise");
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor> projection_ut_3dgs_fused(
    const at::Tensor means, const at::Tensor quats, const at::Tensor scales, const at::optional<at::Tensor> opacities,
    const at::Tensor viewmats0, const at::optional<at::Tensor> viewmats1, const at::Tensor Ks,
    const uint32_t image_width, const uint32_t image_height, const float eps2d, const float near_plane,
    const float far_plane, const float radius_clip, const bool calc_compensations, const CameraModelType camera_model,
    const UnscentedTransformParameters ut_params, ShutterType rs_type, const at::optional<at::Tensor> radial_coeffs,
    const at::optional<at::Tensor> tangential_coeffs, const at::optional<at::Tensor> thin_prism_coeffs) {
    DEVICE_GUARD(means);
    CHECK_INPUT(means);
    CHECK_INPUT(quats);
    CHECK_INPUT(scales);
    CHECK_INPUT(viewmats0);
    CHECK_INPUT(Ks);
    if (opacities.has_value()) {
        CHECK_INPUT(opacities.value());
    }
    const uint32_t N = means.size(0), C = Ks.size(0);
    at::Tensor radii = at::empty({C, N, 2}, means.options().dtype(at::kInt));
    at::Tensor means2d = at::empty({C, N, 2}, means.options());
    at::Tensor depths = at::empty({C, N}, means.options());
    at::Tensor conics = at::empty({C, N, 3}, means.options());
    at::Tensor compensations;
    if (calc_compensations)
        compensations = at::zeros({C, N}, means.options());
    const lfs_ut_params ut = to_ut(ut_params);
    lfs_ok(lfs_projection_ut_3dgs_fused(fp(means), fp(quats), fp(scales), fpo(opacities), fp(viewmats0), fpo(viewmats1),
                                        fp(Ks), N, C, image_width, image_height, eps2d, near_plane, far_plane,
                                        radius_clip, (int)camera_model, &ut, (int)rs_type, fpo(radial_coeffs),
                                        fpo(tangential_coeffs), fpo(thin_prism_coeffs), radii.data_ptr<int32_t>(),
                                        means2d.data_ptr<float>(), depths.data_ptr<float>(), conics.data_ptr<float>(),
                                        calc_compensations ? compensations.data_ptr<float>() : nullptr, cur_stream()),
           "projection_ut_3dgs_fused");
    return std::make_tuple(radii, means2d, depths, conics, compensations);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> rasterize_to_pixels_from_world_3dgs_fwd(
    const at::Tensor means, const at::Tensor quats, const at::Tensor scales, const at::Tensor colors,
    const at::Tensor opacities, const at::optional<at::Tensor> backgrounds, const at::optional<at::Tensor> masks,
    const uint32_t image_width, const uint32_t image_height, const uint32_t tile_size, const at::Tensor viewmats0,
    const at::optional<at::Tensor> viewmats1, const at::Tensor Ks, const CameraModelType camera_model,
    const UnscentedTransformParameters ut_params, ShutterType rs_type, const at::optional<at::Tensor> radial_coeffs,
    const at::optional<at::Tensor> tangential_coeffs, const at::optional<at::Tensor> thin_prism_coeffs,
    const at::Tensor tile_offsets, const at::Tensor flatten_ids) {
    DEVICE_GUARD(means);
    CHECK_INPUT(means);
    CHECK_INPUT(quats);
    CHECK_INPUT(scales);
    CHECK_INPUT(colors);
    CHECK_INPUT(opacities);
    CHECK_INPUT(tile_offsets);
    CHECK_INPUT(flatten_ids);
    if (backgrounds.has_value() && backgrounds.value().numel() > 0) {
        CHECK_INPUT(backgrounds.value());
    }
    const uint32_t C = tile_offsets.size(0), N = means.size(0), channels = colors.size(-1);
    at::Tensor renders = at::empty({C, image_height, image_width, channels}, means.options());
    at::Tensor alphas = at::empty({C, image_height, image_width, 1}, means.options());
    at::Tensor last_ids = at::empty({C, image_height, image_width}, means.options().dtype(at::kInt));
    const bool has_bg = backgrounds.has_value() && backgrounds.value().numel() > 0; // rasterizer.cpp:300-303
    const lfs_ut_params ut = to_ut(ut_params);
    Alloc al{means.device(), {}, {}};
    lfs_ok(lfs_rasterize_to_pixels_from_world_3dgs_fwd(
               fp(means), fp(quats), fp(scales), fp(colors), fp(opacities), has_bg ? fp(backgrounds.value()) : nullptr,
               maskp(masks), N, C, channels, image_width, image_height, tile_size, fp(viewmats0), fpo(viewmats1), fp(Ks),
               (int)camera_model, &ut, (int)rs_type, fpo(radial_coeffs), fpo(tangential_coeffs), fpo(thin_prism_coeffs),
               tile_offsets.data_ptr<int32_t>(), flatten_ids.data_ptr<int32_t>(), flatten_ids.numel(), &Alloc::fn, &al,
               renders.data_ptr<float>(), alphas.data_ptr<float>(), last_ids.data_ptr<int32_t>(), cur_stream()),
           "rasterize_to_pixels_from_world_3dgs_fwd");
    return std::make_tuple(renders, alphas, last_ids);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor> rasterize_to_pixels_from_world_3dgs_bwd(
    const at::Tensor means, const at::Tensor quats, const at::Tensor scales, const at::Tensor colors,
    const at::Tensor opacities, const at::optional<at::Tensor> backgrounds, const at::optional<at::Tensor> masks,
    const uint32_t image_width, const uint32_t image_height, const uint32_t tile_size, const at::Tensor viewmats0,
    const at::optional<at::Tensor> viewmats1, const at::Tensor Ks, const CameraModelType camera_model,
    const UnscentedTransformParameters ut_params, ShutterType rs_type, const at::optional<at::Tensor> radial_coeffs,
    const at::optional<at::Tensor> tangential_coeffs, const at::optional<at::Tensor> thin_prism_coeffs,
    const at::Tensor tile_offsets, const at::Tensor flatten_ids, const at::Tensor render_alphas,
    const at::Tensor last_ids, const at::Tensor v_render_colors, const at::Tensor v_render_alphas) {
    DEVICE_GUARD(means);
    CHECK_INPUT(means);
    CHECK_INPUT(quats);
    CHECK_INPUT(scales);
    CHECK_INPUT(colors);
    CHECK_INPUT(opacities);
    CHECK_INPUT(tile_offsets);
    CHECK_INPUT(flatten_ids);
    CHECK_INPUT(render_alphas);
    CHECK_INPUT(last_ids);
    CHECK_INPUT(v_render_colors);
    CHECK_INPUT(v_render_alphas);
    const uint32_t C = tile_offsets.size(0), N = means.size(0), channels = colors.size(-1);
    at::Tensor v_means = at::empty_like(means), v_quats = at::empty_like(quats), v_scales = at::empty_like(scales);
    at::Tensor v_colors = at::empty_like(colors), v_opacities = at::empty_like(opacities);
    const bool has_bg = backgrounds.has_value() && backgrounds.value().numel() > 0;
    const lfs_ut_params ut = to_ut(ut_params);
    Alloc al{means.device(), {}, {}};
    lfs_ok(lfs_rasterize_to_pixels_from_world_3dgs_bwd(
               fp(means), fp(quats), fp(scales), fp(colors), fp(opacities), has_bg ? fp(backgrounds.value()) : nullptr,
               maskp(masks), N, C, channels, image_width, image_height, tile_size, fp(viewmats0), fpo(viewmats1), fp(Ks),
               (int)camera_model, &ut, (int)rs_type, fpo(radial_coeffs), fpo(tangential_coeffs), fpo(thin_prism_coeffs),
               tile_offsets.data_ptr<int32_t>(), flatten_ids.data_ptr<int32_t>(), flatten_ids.numel(), fp(render_alphas),
               last_ids.data_ptr<int32_t>(), fp(v_render_colors), fp(v_render_alphas), &Alloc::fn, &al,
               v_means.data_ptr<float>(), v_quats.data_ptr<float>(), v_scales.data_ptr<float>(),
               v_colors.data_ptr<float>(), v_opacities.data_ptr<float>(), cur_stream()),
           "rasterize_to_pixels_from_world_3dgs_bwd");
    return std::make_tuple(v_means, v_quats, v_scales, v_colors, v_opacities);
}

} // namespace gsplat
