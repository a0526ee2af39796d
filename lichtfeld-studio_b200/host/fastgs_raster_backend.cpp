// lfs_b200 -- drop-in for the rasterizer half of `fastgs_backend`: fast_gs::rasterization::forward_wrapper /
// backward_wrapper (reference fastgs/rasterization/include/rasterization_api.h:26-75, implementation
// src/rasterization_api.cu:15-181) on top of lfs_fastgs_forward / lfs_fastgs_backward.
// Includes the reference's own header for the declarations.  The four byte tensors of the tuple
// carry this library's state instead of the reference's buffer_utils.h structs; callers only hand them back
// (fast_rasterizer_autograd.cpp:61-72,131-150).
#include "rasterization_api.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "lfs_b200.h"

namespace {
struct BlobAlloc {
    at::Device dev;
    std::vector<torch::Tensor> keep;
    torch::Tensor blob[4]; // per_primitive, per_tile, per_instance, per_bucket
    static void* fn(void* ctx, int tag, size_t bytes) {
        auto* self = static_cast<BlobAlloc*>(ctx);
        try {
            // caching-allocator blocks are 512-byte aligned, which satisfies the 256-byte contract of lfs_alloc_fn and
            // lets backward() use data_ptr() of the returned tensor directly
            torch::Tensor t = torch::empty({(int64_t)std::max<size_t>(bytes, 1)},
                                           torch::TensorOptions().dtype(torch::kByte).device(self->dev));
            if (reinterpret_cast<uintptr_t>(t.data_ptr()) % 256 != 0)
                return nullptr;
            if (tag >= LFS_TAG_FG_PER_PRIMITIVE && tag <= LFS_TAG_FG_PER_BUCKET)
                self->blob[tag - LFS_TAG_FG_PER_PRIMITIVE] = t;
            else
                self->keep.push_back(t);
            return t.data_ptr();
        } catch (...) {
            return nullptr;
        }
    }
};
void check_cuda_float(const torch::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kFloat, name,
                " must be a contiguous CUDA float tensor");
}
} // namespace

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, int, int, int, int, int>
fast_gs::rasterization::forward_wrapper(const torch::Tensor& means, const torch::Tensor& scales_raw,
                                        const torch::Tensor& rotations_raw, const torch::Tensor& opacities_raw,
                                        const torch::Tensor& sh_coefficients_0, const torch::Tensor& sh_coefficients_rest,
                                        const torch::Tensor& w2c, const torch::Tensor& cam_position,
                                        const int active_sh_bases, const int width, const int height,
                                        const float focal_x, const float focal_y, const float center_x,
                                        const float center_y, const float near_plane, const float far_plane) {
    // the reference only checks under config::debug and otherwise fails asynchronously; checking is cheap
    check_cuda_float(means, "means"), check_cuda_float(scales_raw, "scales_raw");
    check_cuda_float(rotations_raw, "rotations_raw"), check_cuda_float(opacities_raw, "opacities_raw");
    check_cuda_float(sh_coefficients_0, "sh_coefficients_0"), check_cuda_float(sh_coefficients_rest, "sh_coefficients_rest");
    const c10::cuda::OptionalCUDAGuard guard(means.device());
    const int n_primitives = (int)means.size(0);
    const int total_bases_sh_rest = (int)sh_coefficients_rest.size(1);
    const auto fopt = torch::TensorOptions().dtype(torch::kFloat).device(means.device());
    torch::Tensor image = torch::empty({3, height, width}, fopt);
    torch::Tensor alpha = torch::empty({1, height, width}, fopt);
    const torch::Tensor w2c_c = w2c.contiguous(), cam_c = cam_position.contiguous();
    BlobAlloc al{means.device()};
    int n_visible = 0, n_instances = 0, n_buckets = 0, selector = 0;
    const int rc = lfs_fastgs_forward(
        means.data_ptr<float>(), scales_raw.data_ptr<float>(), rotations_raw.data_ptr<float>(),
        opacities_raw.data_ptr<float>(), sh_coefficients_0.data_ptr<float>(),
        total_bases_sh_rest ? sh_coefficients_rest.data_ptr<float>() : nullptr, w2c_c.data_ptr<float>(),
        cam_c.data_ptr<float>(), (uint32_t)n_primitives, active_sh_bases, total_bases_sh_rest, width, height, focal_x,
        focal_y, center_x, center_y, near_plane, far_plane, image.data_ptr<float>(), alpha.data_ptr<float>(),
        &BlobAlloc::fn, &al, &n_visible, &n_instances, &n_buckets, &selector, at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == LFS_OK, "fastgs forward failed: ", lfs_last_error());
    // the reference's primitive-indices selector has no counterpart here (one depth sort buffer pair, order kept in
    // per_primitive_buffers); 0 is returned and ignored on the way back
    return {image, alpha, al.blob[0], al.blob[1], al.blob[2], al.blob[3], n_visible, n_instances, n_buckets, 0, selector};
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
fast_gs::rasterization::backward_wrapper(
    torch::Tensor& densification_info, const torch::Tensor& grad_image, const torch::Tensor& grad_alpha,
    const torch::Tensor& image, const torch::Tensor& alpha, const torch::Tensor& means, const torch::Tensor& scales_raw,
    const torch::Tensor& rotations_raw, const torch::Tensor& sh_coefficients_rest,
    const torch::Tensor& per_primitive_buffers, const torch::Tensor& per_tile_buffers,
    const torch::Tensor& per_instance_buffers, const torch::Tensor& per_bucket_buffers, const torch::Tensor& w2c,
    const torch::Tensor& cam_position, const int active_sh_bases, const int width, const int height, const float focal_x,
    const float focal_y, const float center_x, const float center_y, const float near_plane, const float far_plane,
    const int n_visible_primitives, const int n_instances, const int n_buckets,
    const int primitive_primitive_indices_selector, const int instance_primitive_indices_selector) {
    (void)image, (void)alpha, (void)near_plane, (void)far_plane, (void)n_visible_primitives,
        (void)primitive_primitive_indices_selector; // the forward state already holds what these carried
    check_cuda_float(grad_image, "grad_image"), check_cuda_float(grad_alpha, "grad_alpha");
    const c10::cuda::OptionalCUDAGuard guard(means.device());
    const int n_primitives = (int)means.size(0);
    const int total_bases_sh_rest = (int)sh_coefficients_rest.size(1);
    const auto fopt = torch::TensorOptions().dtype(torch::kFloat).device(means.device());
    torch::Tensor grad_means = torch::empty({n_primitives, 3}, fopt);
    torch::Tensor grad_scales_raw = torch::empty({n_primitives, 3}, fopt);
    torch::Tensor grad_rotations_raw = torch::empty({n_primitives, 4}, fopt);
    torch::Tensor grad_opacities_raw = torch::empty({n_primitives, 1}, fopt);
    torch::Tensor grad_sh_coefficients_0 = torch::empty({n_primitives, 1, 3}, fopt);
    torch::Tensor grad_sh_coefficients_rest = torch::empty({n_primitives, total_bases_sh_rest, 3}, fopt);
    torch::Tensor grad_w2c = torch::Tensor();
    if (w2c.requires_grad())
        grad_w2c = torch::zeros_like(w2c, fopt);
    const bool update_densification_info = densification_info.size(0) > 0;
    const torch::Tensor w2c_c = w2c.contiguous(), cam_c = cam_position.contiguous();
    BlobAlloc al{means.device()};
    const int rc = lfs_fastgs_backward(
        grad_image.data_ptr<float>(), grad_alpha.data_ptr<float>(), means.data_ptr<float>(), scales_raw.data_ptr<float>(),
        rotations_raw.data_ptr<float>(), total_bases_sh_rest ? sh_coefficients_rest.data_ptr<float>() : nullptr,
        w2c_c.data_ptr<float>(), cam_c.data_ptr<float>(), per_primitive_buffers.data_ptr(), per_tile_buffers.data_ptr(),
        per_instance_buffers.data_ptr(), per_bucket_buffers.data_ptr(), grad_means.data_ptr<float>(),
        grad_scales_raw.data_ptr<float>(), grad_rotations_raw.data_ptr<float>(), grad_opacities_raw.data_ptr<float>(),
        grad_sh_coefficients_0.data_ptr<float>(), total_bases_sh_rest ? grad_sh_coefficients_rest.data_ptr<float>() : nullptr,
        grad_w2c.defined() ? grad_w2c.data_ptr<float>() : nullptr,
        update_densification_info ? densification_info.data_ptr<float>() : nullptr, (uint32_t)n_primitives, n_instances,
        n_buckets, instance_primitive_indices_selector, active_sh_bases, total_bases_sh_rest, width, height, focal_x,
        focal_y, center_x, center_y, &BlobAlloc::fn, &al, at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == LFS_OK, "fastgs backward failed: ", lfs_last_error());
    return {grad_means, grad_scales_raw, grad_rotations_raw, grad_opacities_raw, grad_sh_coefficients_0,
            grad_sh_coefficients_rest, grad_w2c};
}
