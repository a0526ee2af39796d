// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Python module that runs the reference's OWN training-step callers, compiled UNCHANGED from /root/reference:
//   src/training/rasterization/fast_rasterizer_autograd.cpp   gs::training::FastGSRasterize  (C++ autograd Function)
//   src/training/optimizers/fused_adam.cpp                    gs::training::FusedAdam
//   src/training/kernels/ssim.cu + include/kernels/fused_ssim.cuh   fused_ssim (autograd)
// on top of a fastgs backend chosen at LINK time (oracle/Makefile):
//   ref_fastgs_torch   <- the reference's own fastgs objects (forward.cu, backward.cu, rasterization_api.cu, adam*.cu)
//   b200_fastgs_torch  <- lichtfeld-studio_b200/libgsplat_backend_b200.so (this project's host layer, which exports the
//                         same mangled fast_gs::rasterization::{forward,backward}_wrapper / fast_gs::optimizer symbols)
// Both modules are the same object code above the backend boundary, so (a) the first is the "reference fastgs build"
// training step timed beside ours, (b) the second proves that the reference's callers link and run against our
// boundary unchanged.
//
// What is restated here (20 lines, because gs::Camera / gs::SplatData need the whole application):
//   fast_rasterize()            src/training/rasterization/fast_rasterizer.cpp:12-67  -> Harness::render
//   compute_photometric_loss()  src/training/trainer.cpp:103-131                      -> Harness::photometric_loss
//   create_optimizer()          src/training/strategies/strategy_utils.cpp:21-49      -> Harness::Harness
//   the step structure          src/training/trainer.cpp:579-757 (render, loss, backward, optimizer step, zero_grad)
#include "adam_api.h"
#include "fast_rasterizer_autograd.hpp"
#include "fused_adam.hpp"
#include "kernels/fused_ssim.cuh"
#include "rasterization_api.h"

#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

namespace {
    using torch::Tensor;

    struct Harness {
        Tensor means, sh0, shN, scales, rot, op, densification_info;
        std::unique_ptr<gs::training::FusedAdam> opt;

        Harness(Tensor means_, Tensor sh0_, Tensor shN_, Tensor scales_, Tensor rot_, Tensor op_, std::vector<double> lrs) {
            auto prep = [](Tensor t) { return t.to(torch::kCUDA).contiguous().clone().set_requires_grad(true); };
            means = prep(means_), sh0 = prep(sh0_), shN = prep(shN_), scales = prep(scales_), rot = prep(rot_), op = prep(op_);
            densification_info = torch::zeros({2, means.size(0)}, means.options()).set_requires_grad(false);
            using Options = gs::training::FusedAdam::Options;
            std::vector<torch::optim::OptimizerParamGroup> groups;
            auto add = [&groups](const Tensor& p, double lr) {
                auto o = std::make_unique<Options>(lr);
                o->eps(1e-15).betas(std::make_tuple(0.9, 0.999));
                groups.emplace_back(std::vector<Tensor>{p}, std::unique_ptr<torch::optim::OptimizerOptions>(std::move(o)));
            };
            add(means, lrs[0]), add(sh0, lrs[1]), add(shN, lrs[2]), add(scales, lrs[3]), add(rot, lrs[4]), add(op, lrs[5]);
            auto g = std::make_unique<Options>(0.f);
            g->eps(1e-15);
            opt = std::make_unique<gs::training::FusedAdam>(std::move(groups), std::move(g));
        }

        // fast_rasterize (fast_rasterizer.cpp:12-67)
        std::pair<Tensor, Tensor> render(const Tensor& w2c, const Tensor& cam_position, int active_sh_bases, int width,
                                         int height, float fx, float fy, float cx, float cy, Tensor& bg) {
            fast_gs::rasterization::FastGSSettings s;
            s.cam_position = cam_position;
            s.active_sh_bases = active_sh_bases;
            s.width = width, s.height = height;
            s.focal_x = fx, s.focal_y = fy, s.center_x = cx, s.center_y = cy;
            s.near_plane = 0.01f, s.far_plane = 1e10f;
            auto out = gs::training::FastGSRasterize::apply(means, scales, rot, op, sh0, shN, w2c, densification_info, s);
            Tensor image = out[0] + (1.0f - out[1]) * bg.unsqueeze(-1).unsqueeze(-1);
            return {image, out[1]};
        }

        // compute_photometric_loss (trainer.cpp:103-131)
        static Tensor photometric_loss(Tensor rendered, Tensor gt, float lambda_dssim) {
            rendered = rendered.dim() == 3 ? rendered.unsqueeze(0) : rendered;
            gt = gt.dim() == 3 ? gt.unsqueeze(0) : gt;
            auto l1 = torch::l1_loss(rendered, gt);
            auto ssim_loss = 1.f - fused_ssim(rendered, gt, "valid", true);
            return (1.f - lambda_dssim) * l1 + lambda_dssim * ssim_loss;
        }

        // one batch of views -> one optimizer step; returns the summed loss when read_loss
        double train_step(int iteration, std::vector<Tensor> w2cs, std::vector<Tensor> cam_positions,
                          std::vector<std::vector<double>> intrinsics, std::vector<Tensor> gts, Tensor bg,
                          double lambda_dssim, int active_sh_bases, int width, int height, bool read_loss) {
            Tensor total;
            for (size_t v = 0; v < w2cs.size(); ++v) {
                const auto& k = intrinsics[v];
                auto [image, alpha] = render(w2cs[v], cam_positions[v], active_sh_bases, width, height, (float)k[0],
                                             (float)k[1], (float)k[2], (float)k[3], bg);
                Tensor loss = photometric_loss(image, gts[v], (float)lambda_dssim);
                loss.backward();
                if (read_loss)
                    total = total.defined() ? total + loss.detach() : loss.detach();
            }
            opt->step(iteration);
            opt->zero_grad(true, iteration);
            return read_loss ? total.item<double>() : 0.0;
        }

        // forward + loss + backward of ONE view without the optimizer: returns (image, alpha, loss, grads[6]) for parity
        std::vector<Tensor> view_grads(Tensor w2c, Tensor cam_position, std::vector<double> k, Tensor gt, Tensor bg,
                                       double lambda_dssim, int active_sh_bases, int width, int height) {
            opt->zero_grad(true, 0);
            auto [image, alpha] = render(w2c, cam_position, active_sh_bases, width, height, (float)k[0], (float)k[1],
                                         (float)k[2], (float)k[3], bg);
            Tensor loss = photometric_loss(image, gt, (float)lambda_dssim);
            loss.backward();
            return {image.detach(), alpha.detach(), loss.detach(), means.grad(), sh0.grad(), shN.grad(), scales.grad(),
                    rot.grad(), op.grad()};
        }

        std::vector<Tensor> params() { return {means.detach(), sh0.detach(), shN.detach(), scales.detach(), rot.detach(), op.detach()}; }
    };
} // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("forward_wrapper", &fast_gs::rasterization::forward_wrapper);
    m.def("backward_wrapper", &fast_gs::rasterization::backward_wrapper);
    m.def("adam_step_wrapper", &fast_gs::optimizer::adam_step_wrapper);
    m.def("fused_ssim", [](Tensor a, Tensor b, std::string padding, bool train) { return fused_ssim(a, b, padding, train); });
    py::class_<Harness>(m, "Harness")
        .def(py::init<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, std::vector<double>>())
        .def("train_step", &Harness::train_step)
        .def("view_grads", &Harness::view_grads)
        .def("params", &Harness::params)
        .def_readwrite("densification_info", &Harness::densification_info);
}
