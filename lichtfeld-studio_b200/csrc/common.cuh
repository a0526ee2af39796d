// lfs_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lfs_b200.h"

namespace lfs {

// ---- error plumbing (host) -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
#define LFS_CHECK_ARG(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            lfs::set_error(__VA_ARGS__);  \
            return LFS_ERR_INVALID_ARG;   \
        }                                 \
    } while (0)
#define LFS_UNSUPPORTED(cond, ...)        \
    do {                                  \
        if (cond) {                       \
            lfs::set_error(__VA_ARGS__);  \
            return LFS_ERR_UNSUPPORTED;   \
        }                                 \
    } while (0)
#define LFS_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            lfs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return LFS_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)
#define LFS_LAUNCH_OK(name)                                                                    \
    do {                                                                                       \
        lfs::count_launch();                                                                   \
        cudaError_t e__ = cudaPeekAtLastError();                                               \
        if (e__ != cudaSuccess) {                                                              \
            lfs::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));          \
            return LFS_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

static inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
int num_sms();                   // SM count of the current device (queried once per device; 148 on B200)
constexpr int kTile = 16;        // tile edge in pixels (reference constant, rasterizer.cpp:180)
constexpr int kTilePix = kTile * kTile;
constexpr int kBucket = 32;      // gaussians per backward bucket (one warp)

// bump allocator over one scratch blob
struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base ? base + off : nullptr);
        off += count * sizeof(T);
        return p;
    }
    size_t total() const { return align_up(off, 256); }
};

// ---- small device math ------------------------------------------------------------------------------------
struct f3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__host__ __device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ f3 operator*(float s, f3 a) { return f3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return f3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}

// quaternion (w,x,y,z) helpers following the glm routes the reference takes for the camera pose
// (gsplat/Cameras.cuh:39-56, :253-280) so that depths agree with the reference to rounding.
struct quat4 {
    float w, x, y, z;
};
__host__ __device__ __forceinline__ quat4 quat_from_rowmajor_rot(const float* m /* viewmat [4,4] */) {
    // glm::quat_cast on the column-major copy of the row-major rotation: M(c,r) = m[r*4+c]
    const float m00 = m[0], m01 = m[4], m02 = m[8];  // glm m[0][*]
    const float m10 = m[1], m11 = m[5], m12 = m[9];  // glm m[1][*]
    const float m20 = m[2], m21 = m[6], m22 = m[10]; // glm m[2][*]
    const float fx = m00 - m11 - m22, fy = m11 - m00 - m22, fz = m22 - m00 - m11, fw = m00 + m11 + m22;
    int bi = 0;
    float fb = fw;
    if (fx > fb) {
        fb = fx;
        bi = 1;
    }
    if (fy > fb) {
        fb = fy;
        bi = 2;
    }
    if (fz > fb) {
        fb = fz;
        bi = 3;
    }
    const float bv = sqrtf(fb + 1.0f) * 0.5f;
    const float mult = 0.25f / bv;
    quat4 q;
    switch (bi) {
    case 0:
        q = quat4{bv, (m12 - m21) * mult, (m20 - m02) * mult, (m01 - m10) * mult};
        break;
    case 1:
        q = quat4{(m12 - m21) * mult, bv, (m01 + m10) * mult, (m20 + m02) * mult};
        break;
    case 2:
        q = quat4{(m20 - m02) * mult, (m01 + m10) * mult, bv, (m12 + m21) * mult};
        break;
    default:
        q = quat4{(m01 - m10) * mult, (m20 + m02) * mult, (m12 + m21) * mult, bv};
        break;
    }
    return q;
}
__host__ __device__ __forceinline__ f3 quat_rotate(quat4 q, f3 v) {
    const f3 u = mk3(q.x, q.y, q.z);
    const f3 uv = cross(u, v);
    const f3 uuv = cross(u, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}

// Camera block shared by all kernels of one view (filled on the host, passed by value).
struct ViewCam {
    float R[9];  // row-major world->camera rotation
    float t[3];  // translation
    float q[4];  // glm::quat_cast(R) (w,x,y,z): the reference's per-Gaussian pose route
    float org[3]; // camera centre in world space (-R^T t)
    float fx, fy, cx, cy;
    int width, height, tile_w, tile_h;
};

// streaming 128-bit loads
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_nc4(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed fp32 pairs (sm_100a FFMA2 / FMUL2: two IEEE fp32 operations per instruction, each component rounded exactly like
// the scalar fmaf / __fmul_rn): used wherever two accumulations share a multiplier
// (the N' / D polynomials of a pair, the SSIM convolutions).
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
          "l"(*reinterpret_cast<const unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

} // namespace lfs
