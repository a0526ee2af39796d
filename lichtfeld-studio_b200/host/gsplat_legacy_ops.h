// lfs_b200 -- declarations of the legacy 2-D gsplat ops the reference's gtest files call but gsplat/Ops.h no longer
// declares: include this next to "Ops.h" to build tests/test_basic.cpp, tests/test_gsplat_ops.cpp and
// tests/test_rasterization.cpp against libgsplat_backend_b200.  Argument order and meaning are the call sites':
//   quat_scale_to_covar_preci_fwd/bwd   tests/test_basic.cpp:54-81, tests/test_gsplat_ops.cpp:76-96
//   projection_ewa_3dgs_fused_fwd       tests/test_basic.cpp:114-128 (an EMPTY covars tensor selects quats + scales)
//   rasterize_to_pixels_3dgs_fwd        tests/test_basic.cpp:347-358 (an EMPTY masks tensor means no masks)
//   rasterize_to_pixels_3dgs_bwd        the launcher's argument list, gsplat/Rasterization.h:38-63
#pragma once
#include <ATen/ATen.h>
#include <tuple>

#include "Common.h" // gsplat::CameraModelType

namespace gsplat {

std::tuple<at::Tensor, at::Tensor> quat_scale_to_covar_preci_fwd(const at::Tensor quats,  // [N, 4]
                                                                 const at::Tensor scales, // [N, 3]
                                                                 const bool compute_covar, const bool compute_preci,
                                                                 const bool triu);

std::tuple<at::Tensor, at::Tensor> quat_scale_to_covar_preci_bwd(const at::Tensor quats, const at::Tensor scales,
                                                                 const bool triu,
                                                                 const at::optional<at::Tensor> v_covars,
                                                                 const at::optional<at::Tensor> v_precis);

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>
projection_ewa_3dgs_fused_fwd(const at::Tensor means,                    // [N, 3]
                              const at::optional<at::Tensor> covars,     // [N, 3, 3]; absent or empty: quats + scales
                              const at::optional<at::Tensor> quats,      // [N, 4]
                              const at::optional<at::Tensor> scales,     // [N, 3]
                              const at::optional<at::Tensor> opacities,  // [N]
                              const at::Tensor viewmats,                 // [C, 4, 4]
                              const at::Tensor Ks,                       // [C, 3, 3]
                              const uint32_t image_width, const uint32_t image_height, const float eps2d,
                              const float near_plane, const float far_plane, const float radius_clip,
                              const bool calc_compensations, const CameraModelType camera_model);

std::tuple<at::Tensor, at::Tensor, at::Tensor>
rasterize_to_pixels_3dgs_fwd(const at::Tensor means2d,                   // [C, N, 2]
                             const at::Tensor conics,                    // [C, N, 3]
                             const at::Tensor colors,                    // [C, N, channels]
                             const at::Tensor opacities,                 // [C, N]
                             const at::optional<at::Tensor> backgrounds, // [C, channels]
                             const at::optional<at::Tensor> masks,       // [C, tile_height, tile_width]; empty: none
                             const uint32_t image_width, const uint32_t image_height, const uint32_t tile_size,
                             const at::Tensor tile_offsets, // [C, tile_height, tile_width]
                             const at::Tensor flatten_ids   // [n_isects]
);

// -> (v_means2d_abs (empty unless absgrad), v_means2d, v_conics, v_colors, v_opacities)
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>
rasterize_to_pixels_3dgs_bwd(const at::Tensor means2d, const at::Tensor conics, const at::Tensor colors,
                             const at::Tensor opacities, const at::optional<at::Tensor> backgrounds,
                             const at::optional<at::Tensor> masks, const uint32_t image_width,
                             const uint32_t image_height, const uint32_t tile_size, const at::Tensor tile_offsets,
                             const at::Tensor flatten_ids, const at::Tensor render_alphas, const at::Tensor last_ids,
                             const at::Tensor v_render_colors, const at::Tensor v_render_alphas, const bool absgrad);

} // namespace gsplat
