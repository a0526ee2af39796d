// lfs_b200 -- drop-in for the optimizer half of `fastgs_backend`: fast_gs::optimizer::adam_step_wrapper /
// adam_step (reference fastgs/optimizer/include/adam_api.h:11-21, adam.h:9-20) on top of lfs_adam_step.
// Includes the reference's own headers for the declarations.
#include "adam.h"
#include "adam_api.h"

#include <ATen/cuda/CUDAContext.h>

#include "lfs_b200.h"

void fast_gs::optimizer::adam_step(float* param, float* exp_avg, float* exp_avg_sq, const float* param_grad,
                                   const int n_elements, const float lr, const float beta1, const float beta2,
                                   const float eps, const float bias_correction1_rcp,
                                   const float bias_correction2_sqrt_rcp) {
    const int rc = lfs_adam_step(param, exp_avg, exp_avg_sq, param_grad, n_elements, lr, beta1, beta2, eps,
                                 bias_correction1_rcp, bias_correction2_sqrt_rcp,
                                 at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == LFS_OK, "adam_step failed: ", lfs_last_error());
}

void fast_gs::optimizer::adam_step_wrapper(torch::Tensor& param, torch::Tensor& exp_avg, torch::Tensor& exp_avg_sq,
                                           const torch::Tensor& param_grad, const float lr, const float beta1,
                                           const float beta2, const float eps, const float bias_correction1_rcp,
                                           const float bias_correction2_sqrt_rcp) {
    adam_step(param.data_ptr<float>(), exp_avg.data_ptr<float>(), exp_avg_sq.data_ptr<float>(),
              param_grad.data_ptr<float>(), (int)param.numel(), lr, beta1, beta2, eps, bias_correction1_rcp,
              bias_correction2_sqrt_rcp);
}
