// lfs_b200 -- the legacy 2-D gsplat ops (gsplat_legacy_ops.h) on top of the C ABI of liblfs_b200.so: thin host dispatch
// like gsplat_backend.cpp (CHECK_INPUT as gsplat/Common.h:12-17, ATen outputs, raw pointers + the current stream).
#include "gsplat_legacy_ops.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "Common.h"
#include "lfs_b200.h"

namespace {

void lfs_ok(int rc, const char* what) {
    TORCH_CHECK(rc == LFS_OK, what, " failed: ", lfs_last_error(), " (lfs status ", rc, ")");
}
void* cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
bool present(const at::optional<at::Tensor>& t) { return t.has_value() && t.value().defined() && t.value().numel() > 0; }
const float* fpo(const at::optional<at::Tensor>& t) { return present(t) ? t.value().data_ptr<float>() : nullptr; }
const uint8_t* maskp(const at::optional<at::Tensor>& t) {
    return present(t) ? reinterpret_cast<const uint8_t*>(t.value().data_ptr<bool>()) : nullptr;
}

} // namespace

namespace gsplat {

std::tuple<at::Tensor, at::Tensor> quat_scale_to_covar_preci_fwd(const at::Tensor quats, const at::Tensor scales,
                                                                 const bool compute_covar, const bool compute_preci,
                                                                 const bool triu) {
    DEVICE_GUARD(quats);
    CHECK_INPUT(quats);
    CHECK_INPUT(scales);
    const int64_t N = quats.size(0);
    auto opt = quats.options();
    at::Tensor covars = compute_covar ? (triu ? at::empty({N, 6}, opt) : at::empty({N, 3, 3}, opt)) : at::Tensor();
    at::Tensor precis = compute_preci ? (triu ? at::empty({N, 6}, opt) : at::empty({N, 3, 3}, opt)) : at::Tensor();
    if (compute_covar || compute_preci)
        lfs_ok(lfs_quat_scale_to_covar_preci_fwd(quats.data_ptr<float>(), scales.data_ptr<float>(), (uint32_t)N, triu ? 1 : 0,
                                                 compute_covar ? covars.data_ptr<float>() : nullptr,
                                                 compute_preci ? precis.data_ptr<float>() : nullptr, cur_stream()),
               "quat_scale_to_covar_preci_fwd");
    return std::make_tuple(covars, precis);
}

std::tuple<at::Tensor, at::Tensor> quat_scale_to_covar_preci_bwd(const at::Tensor quats, const at::Tensor scales,
                                                                 const bool triu,
                                                                 const at::optional<at::Tensor> v_covars,
                                                                 const at::optional<at::Tensor> v_precis) {
    DEVICE_GUARD(quats);
    CHECK_INPUT(quats);
    CHECK_INPUT(scales);
    if (present(v_covars)) { // CHECK_INPUT is two statements
        CHECK_INPUT(v_covars.value());
    }
    if (present(v_precis)) { // CHECK_INPUT is two statements
        CHECK_INPUT(v_precis.value());
    }
    at::Tensor v_quats = at::zeros_like(quats), v_scales = at::zeros_like(scales);
    if (present(v_covars) || present(v_precis))
        lfs_ok(lfs_quat_scale_to_covar_preci_bwd(quats.data_ptr<float>(), scales.data_ptr<float>(), (uint32_t)quats.size(0),
                                                 triu ? 1 : 0, fpo(v_covars), fpo(v_precis), v_quats.data_ptr<float>(),
                                                 v_scales.data_ptr<float>(), cur_stream()),
               "quat_scale_to_covar_preci_bwd");
    return std::make_tuple(v_quats, v_scales);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>
projection_ewa_3dgs_fused_fwd(const at::Tensor means, const at::optional<at::Tensor> covars,
                              const at::optional<at::Tensor> quats, const at::optional<at::Tensor> scales,
                              const at::optional<at::Tensor> opacities, const at::Tensor viewmats, const at::Tensor Ks,
                              const uint32_t image_width, const uint32_t image_height, const float eps2d,
                              const float near_plane, const float far_plane, const float radius_clip,
                              const bool calc_compensations, const CameraModelType camera_model) {
    DEVICE_GUARD(means);
    CHECK_INPUT(means);
    CHECK_INPUT(viewmats);
    CHECK_INPUT(Ks);
    if (present(covars)) {
        CHECK_INPUT(covars.value());
    } else {
        TORCH_CHECK(present(quats) && present(scales), "projection_ewa_3dgs_fused_fwd: covars or quats + scales required");
        CHECK_INPUT(quats.value());
        CHECK_INPUT(scales.value());
    }
    if (present(opacities)) { // CHECK_INPUT is two statements
        CHECK_INPUT(opacities.value());
    }
    const int64_t N = means.size(0), C = viewmats.size(0);
    auto opt = means.options();
    at::Tensor radii = at::empty({C, N, 2}, opt.dtype(at::kInt));
    at::Tensor means2d = at::zeros({C, N, 2}, opt), depths = at::zeros({C, N}, opt), conics = at::zeros({C, N, 3}, opt);
    at::Tensor compensations = calc_compensations ? at::zeros({C, N}, opt) : at::Tensor();
    lfs_ok(lfs_projection_ewa_3dgs_fused_fwd(
               means.data_ptr<float>(), fpo(covars), present(covars) ? nullptr : fpo(quats),
               present(covars) ? nullptr : fpo(scales), fpo(opacities), viewmats.data_ptr<float>(), Ks.data_ptr<float>(),
               (uint32_t)N, (uint32_t)C, image_width, image_height, eps2d, near_plane, far_plane, radius_clip,
               (int)camera_model, radii.data_ptr<int32_t>(), means2d.data_ptr<float>(), depths.data_ptr<float>(),
               conics.data_ptr<float>(), calc_compensations ? compensations.data_ptr<float>() : nullptr, cur_stream()),
           "projection_ewa_3dgs_fused_fwd");
    return std::make_tuple(radii, means2d, depths, conics, compensations);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor>
rasterize_to_pixels_3dgs_fwd(const at::Tensor means2d, const at::Tensor conics, const at::Tensor colors,
                             const at::Tensor opacities, const at::optional<at::Tensor> backgrounds,
                             const at::optional<at::Tensor> masks, const uint32_t image_width,
                             const uint32_t image_height, const uint32_t tile_size, const at::Tensor tile_offsets,
                             const at::Tensor flatten_ids) {
    DEVICE_GUARD(means2d);
    CHECK_INPUT(means2d);
    CHECK_INPUT(conics);
    CHECK_INPUT(colors);
    CHECK_INPUT(opacities);
    CHECK_INPUT(tile_offsets);
    CHECK_INPUT(flatten_ids);
    if (present(backgrounds)) { // CHECK_INPUT is two statements
        CHECK_INPUT(backgrounds.value());
    }
    if (present(masks)) { // CHECK_INPUT is two statements
        CHECK_INPUT(masks.value());
    }
    const int64_t C = tile_offsets.size(0), N = means2d.size(-2), channels = colors.size(-1);
    auto opt = means2d.options();
    at::Tensor renders = at::empty({C, (int64_t)image_height, (int64_t)image_width, channels}, opt);
    at::Tensor alphas = at::empty({C, (int64_t)image_height, (int64_t)image_width, 1}, opt);
    at::Tensor last_ids = at::empty({C, (int64_t)image_height, (int64_t)image_width}, opt.dtype(at::kInt));
    lfs_ok(lfs_rasterize_to_pixels_3dgs_fwd(means2d.data_ptr<float>(), conics.data_ptr<float>(), colors.data_ptr<float>(),
                                            opacities.data_ptr<float>(), fpo(backgrounds), maskp(masks), (uint32_t)C,
                                            (uint32_t)N, (uint32_t)channels, image_width, image_height, tile_size,
                                            tile_offsets.data_ptr<int32_t>(), flatten_ids.data_ptr<int32_t>(),
                                            flatten_ids.numel(), renders.data_ptr<float>(), alphas.data_ptr<float>(),
                                            last_ids.data_ptr<int32_t>(), cur_stream()),
           "rasterize_to_pixels_3dgs_fwd");
    return std::make_tuple(renders, alphas, last_ids);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>
rasterize_to_pixels_3dgs_bwd(const at::Tensor means2d, const at::Tensor conics, const at::Tensor colors,
                             const at::Tensor opacities, const at::optional<at::Tensor> backgrounds,
                             const at::optional<at::Tensor> masks, const uint32_t image_width,
                             const uint32_t image_height, const uint32_t tile_size, const at::Tensor tile_offsets,
                             const at::Tensor flatten_ids, const at::Tensor render_alphas, const at::Tensor last_ids,
                             const at::Tensor v_render_colors, const at::Tensor v_render_alphas, const bool absgrad) {
    DEVICE_GUARD(means2d);
    CHECK_INPUT(means2d);
    CHECK_INPUT(conics);
    CHECK_INPUT(colors);
    CHECK_INPUT(opacities);
    CHECK_INPUT(tile_offsets);
    CHECK_INPUT(flatten_ids);
    CHECK_INPUT(render_alphas);
    CHECK_INPUT(last_ids);
    CHECK_INPUT(v_render_colors);
    CHECK_INPUT(v_render_alphas);
    if (present(backgrounds)) { // CHECK_INPUT is two statements
        CHECK_INPUT(backgrounds.value());
    }
    if (present(masks)) { // CHECK_INPUT is two statements
        CHECK_INPUT(masks.value());
    }
    const int64_t C = tile_offsets.size(0), N = means2d.size(-2), channels = colors.size(-1);
    at::Tensor v_means2d = at::empty_like(means2d), v_conics = at::empty_like(conics);
    at::Tensor v_colors = at::empty_like(colors), v_opacities = at::empty_like(opacities);
    at::Tensor v_means2d_abs = absgrad ? at::empty_like(means2d) : at::Tensor();
    lfs_ok(lfs_rasterize_to_pixels_3dgs_bwd(
               means2d.data_ptr<float>(), conics.data_ptr<float>(), colors.data_ptr<float>(), opacities.data_ptr<float>(),
               fpo(backgrounds), maskp(masks), (uint32_t)C, (uint32_t)N, (uint32_t)channels, image_width, image_height,
               tile_size, tile_offsets.data_ptr<int32_t>(), flatten_ids.data_ptr<int32_t>(), flatten_ids.numel(),
               render_alphas.data_ptr<float>(), last_ids.data_ptr<int32_t>(), v_render_colors.data_ptr<float>(),
               v_render_alphas.data_ptr<float>(), absgrad ? v_means2d_abs.data_ptr<float>() : nullptr,
               v_means2d.data_ptr<float>(), v_conics.data_ptr<float>(), v_colors.data_ptr<float>(),
               v_opacities.data_ptr<float>(), cur_stream()),
           "rasterize_to_pixels_3dgs_bwd");
    return std::make_tuple(v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities);
}

} // namespace gsplat
