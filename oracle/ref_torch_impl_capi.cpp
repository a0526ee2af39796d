// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" glue around the UNMODIFIED reference CPU restatement /root/reference/tests/torch_impl.cpp
// (namespace reference, declared in tests/torch_impl.hpp:9-65), compiled in place by oracle/Makefile into
// oracle/_ref/libtorch_impl_ref.so against the CPU libtorch of this image.  Host pointers in, host
// pointers out.  Used (a) here, in this container, by tests/golden/make_golden.py to generate the
// committed golden vectors that pin oracle/lfs_oracle.c, and (b) by bench.py --impl reference as the
// reference's own CPU path for the stages it covers (SH, tile intersection, quat->covar).
#include "torch_impl.hpp"

#include <cstring>

namespace {
    torch::Tensor wf(const float* p, at::IntArrayRef s) {
        return torch::from_blob(const_cast<float*>(p), s, torch::kFloat32).clone();
    }
    torch::Tensor wi(const int32_t* p, at::IntArrayRef s) {
        return torch::from_blob(const_cast<int32_t*>(p), s, torch::kInt32).clone();
    }
    void out(void* dst, const torch::Tensor& t, int64_t max_elems = -1) {
        if (!dst || !t.defined())
            return;
        auto c = t.contiguous();
        int64_t n = c.numel();
        if (max_elems >= 0 && n > max_elems)
            n = max_elems;
        std::memcpy(dst, c.data_ptr(), n * c.element_size());
    }
} // namespace

extern "C" {

void ref_ti_set_num_threads(int n) { at::set_num_threads(n); }
int ref_ti_get_num_threads() { return at::get_num_threads(); }

// reference::spherical_harmonics (tests/torch_impl.cpp:296-321)
int ref_ti_spherical_harmonics(int degree, const float* dirs, const float* coeffs, int n, int K, float* colors) {
    auto r = reference::spherical_harmonics(degree, wf(dirs, {n, 3}), wf(coeffs, {n, K, 3}));
    out(colors, r);
    return 0;
}

// reference::quat_scale_to_covar_preci (tests/torch_impl.cpp:38-77), full 3x3 outputs
int ref_ti_quat_scale_to_covar_preci(const float* quats, const float* scales, int n, float* covars, float* precis) {
    auto [c, p] = reference::quat_scale_to_covar_preci(wf(quats, {n, 4}), wf(scales, {n, 3}), true, true, false);
    out(covars, c);
    out(precis, p);
    return 0;
}

// reference::fully_fused_projection (tests/torch_impl.cpp:146-218): EWA pinhole projection
int ref_ti_fully_fused_projection(const float* means, const float* covars, const float* viewmats, const float* Ks,
                                  int N, int C, int width, int height, float eps2d, float near_plane, float far_plane,
                                  int32_t* radii, float* means2d, float* depths, float* conics) {
    auto [r, m, d, c, comp] = reference::fully_fused_projection(wf(means, {N, 3}), wf(covars, {N, 3, 3}),
                                                                wf(viewmats, {C, 4, 4}), wf(Ks, {C, 3, 3}), width,
                                                                height, eps2d, near_plane, far_plane, false, "pinhole");
    out(radii, r);
    out(means2d, m);
    out(depths, d);
    out(conics, c);
    return 0;
}

// reference::isect_tiles (tests/torch_impl.cpp:324-419). Returns n_isects.
long long ref_ti_isect_tiles(const float* means2d, const int32_t* radii, const float* depths, int C, int N,
                             int tile_size, int tile_width, int tile_height, int sort, int32_t* tiles_per_gauss,
                             int64_t* isect_ids, int32_t* flatten_ids, long long capacity) {
    auto [t, ids, flat] = reference::isect_tiles(wf(means2d, {C, N, 2}), wi(radii, {C, N, 2}), wf(depths, {C, N}),
                                                 tile_size, tile_width, tile_height, sort != 0);
    out(tiles_per_gauss, t.to(torch::kInt32));
    out(isect_ids, ids, capacity);
    out(flatten_ids, flat, capacity);
    return ids.numel();
}

} // extern "C"
