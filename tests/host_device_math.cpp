// TEST INFRASTRUCTURE ONLY.  The per-Gaussian DEVICE math of the product -- unscented-transform projection
// (csrc/projection.cuh), spherical harmonics (csrc/sh.cuh), the AABB tile rectangle and the exact tile culling span
// (csrc/intersect.cuh) -- compiled for the HOST with plain g++ so that tests/test_host_device_math.py can check the very
// source the kernels compile against the CPU oracle without a GPU.  The product headers are included unmodified; this file
// only supplies host meanings for the device intrinsics they use (each with the intrinsic's documented semantics).  Host
// arithmetic is not bit-identical to the device's (no FMA contraction here, libm instead of MUFU), so the tests compare at
// fp32 rounding, not bit-exactly; bit-exact device parity is the job of the `-m gpu` tests.
//     g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC -I/usr/local/cuda/include -o tests/_build/libhost_device_math.so
//         tests/host_device_math.cpp
#include <cuda_runtime.h> // for a host compiler: vector types, empty __host__ / __device__
#include <algorithm>
#include <math.h>
#include <stdint.h>

using std::max;
using std::min;

// ---- host meanings of the device intrinsics the headers use ---------------------------------------------------------
#define __logf(x) logf(x)                   // lg2.approx * ln2: ~1e-6 relative; logf is the exact counterpart
#define __sinf(x) sinf(x)                   // sin.approx (the FAST slerp is device-only anyway, cameras.cuh)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frsqrt_rn(float x) { return 1.0f / sqrtf(x); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline unsigned __float2uint_rz(float x) { // cvt.rzi.u32.f32: saturating, NaN -> 0
    if (!(x > 0.f))
        return 0u;
    return x >= 4294967296.f ? 0xffffffffu : (unsigned)x;
}
template <class T>
static inline T __ldg(const T* p) {
    return *p;
}

#include "../lichtfeld-studio_b200/csrc/intersect.cuh"
#include "../lichtfeld-studio_b200/csrc/projection.cuh"
#include "../lichtfeld-studio_b200/csrc/sh.cuh"

using namespace lfs;

namespace {
void store(const UTOut& o, int i, int32_t* radii, float* means2d, float* depths, float* conics, float* comp) {
    depths[i] = o.depth; // written before the near / far test, like the kernel's local; only meaningful where radii > 0
    if (!o.ok) {
        radii[2 * i] = radii[2 * i + 1] = 0;
        return;
    }
    radii[2 * i] = (int32_t)o.rx, radii[2 * i + 1] = (int32_t)o.ry;
    means2d[2 * i] = o.mx, means2d[2 * i + 1] = o.my;
    conics[3 * i] = o.c00, conics[3 * i + 1] = o.c01, conics[3 * i + 2] = o.c11;
    if (comp)
        comp[i] = o.comp;
}
} // namespace

extern "C" {

// one camera; means [n,3], quats [n,4] wxyz, scales [n,3], opacities [n] or null
void hd_ut_project_pinhole(int n, const float* means, const float* quats, const float* scales, const float* opac, const float* vm,
                           const float* K, int w, int h, float eps2d, float near_plane, float far_plane, float radius_clip,
                           const lfs_ut_params* ut, int32_t* radii, float* means2d, float* depths, float* conics, float* comp) {
    const ViewCam cam = make_viewcam(vm, K, w, h);
    for (int i = 0; i < n; ++i) {
        const UTOut o = ut_project_pinhole(cam, f3{means[3 * i], means[3 * i + 1], means[3 * i + 2]},
                                           make_float4(quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]),
                                           f3{scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]}, opac != nullptr,
                                           opac ? opac[i] : 1.f, eps2d, near_plane, far_plane, radius_clip, *ut);
        store(o, i, radii, means2d, depths, conics, comp);
    }
}

void hd_ut_project_general(int n, const float* means, const float* quats, const float* scales, const float* opac, const float* vm0,
                           const float* vm1, const float* K, int w, int h, int model, int shutter, const float* radial,
                           const float* tangential, const float* prism, float eps2d, float near_plane, float far_plane,
                           float radius_clip, const lfs_ut_params* ut, int32_t* radii, float* means2d, float* depths,
                           float* conics, float* comp) {
    const CamModel cam = make_cam_model(vm0, vm1, K, (uint32_t)w, (uint32_t)h, model, shutter, radial, tangential, prism);
    for (int i = 0; i < n; ++i) {
        const UTOut o = ut_project_general(cam, f3{means[3 * i], means[3 * i + 1], means[3 * i + 2]},
                                           make_float4(quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]),
                                           f3{scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]}, opac != nullptr,
                                           opac ? opac[i] : 1.f, eps2d, near_plane, far_plane, radius_clip, *ut);
        store(o, i, radii, means2d, depths, conics, comp);
    }
}

// dirs [n,3], coeffs [n,K,3] -> colors [n,3]
void hd_sh_fwd(int degree, int n, int K, const float* dirs, const float* coeffs, float* colors) {
    for (int i = 0; i < n; ++i) {
        const float* c = coeffs + (size_t)i * K * 3;
        const f3 col = sh_to_color(degree, f3{dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]},
                                   [&](int k) { return f3{c[3 * k], c[3 * k + 1], c[3 * k + 2]}; });
        colors[3 * i] = col.x, colors[3 * i + 1] = col.y, colors[3 * i + 2] = col.z;
    }
}

// -> v_coeffs [n,K,3] (entries >= (degree+1)^2 left untouched), v_dirs [n,3] or null
void hd_sh_bwd(int degree, int n, int K, const float* dirs, const float* coeffs, const float* v_colors, float* v_coeffs,
               float* v_dirs) {
    for (int i = 0; i < n; ++i) {
        const float* c = coeffs + (size_t)i * K * 3;
        float* vc = v_coeffs + (size_t)i * K * 3;
        const f3 vd = sh_vjp(
            degree, f3{dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]}, f3{v_colors[3 * i], v_colors[3 * i + 1], v_colors[3 * i + 2]},
            v_dirs != nullptr, [&](int k) { return f3{c[3 * k], c[3 * k + 1], c[3 * k + 2]}; },
            [&](int k, f3 v) { vc[3 * k] = v.x, vc[3 * k + 1] = v.y, vc[3 * k + 2] = v.z; });
        if (v_dirs)
            v_dirs[3 * i] = vd.x, v_dirs[3 * i + 1] = vd.y, v_dirs[3 * i + 2] = vd.z;
    }
}

// means2d [n,2], radii [n,2] (as floats) -> rects [n,4] = x0, y0, x1, y1
void hd_tile_rect(int n, const float* means2d, const float* radii, float tile_size, uint32_t tile_w, uint32_t tile_h, uint32_t* rects) {
    for (int i = 0; i < n; ++i)
        tile_rect(means2d[2 * i], means2d[2 * i + 1], radii[2 * i], radii[2 * i + 1], tile_size, tile_w, tile_h, rects[4 * i],
                  rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
}

// one culling record (a, b, c, xc, yc, lim; idet and ia derived as the kernels derive them) -> [first, last] per tile row
void hd_cull_row_spans(float a, float b, float c, float xc, float yc, float lim, uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                       int* first, int* last) {
    CullRec r;
    r.a = a, r.b = b, r.c = c, r.xc = xc, r.yc = yc, r.lim = lim;
    r.idet = 1.0f / (a * c - b * b), r.ia = 1.0f / a;
    for (uint32_t ty = y0; ty < y1; ++ty) cull_row_span(r, ty, x0, x1, first[ty - y0], last[ty - y0]);
}

int hd_tile_key_bits(uint32_t n_tiles) { return tile_key_bits(n_tiles); }
uint32_t hd_ref_tile_n_bits(uint32_t n_tiles) { return ref_tile_n_bits(n_tiles); }
}
