// lfs_b200 -- from-world (3DGUT) alpha-blend rasterizer: data layout + launch API shared by the
// gsplat-surface ops and the fused trainer.
//
// B200-first reformulation (NOT the reference's per-pixel ray algebra):
// for a PINHOLE / GLOBAL-shutter camera every pixel ray leaves the same origin o and its direction is
// affine in the pixel coordinate, d(px,py) = R^T((px-cx)/fx, (py-cy)/fy, 1).  The reference's response
//   power = -1/2 |normalize(M d) x M(o-mu)|^2          (RasterizeToPixelsFromWorld3DGSFwd.cu:233-237)
// is therefore exactly   power = -1/2 |v x gro|^2 / |v|^2   with  v = M d  affine in (px,py)  and
// gro = M(o-mu) constant per Gaussian, i.e. a RATIO OF TWO QUADRATIC POLYNOMIALS in the pixel coordinate.
// Expanding both polynomials around the centre of a 16x16 tile keeps every term O(result) (no catastrophic
// cancellation; the cross product v0 x gro is evaluated once per (tile, Gaussian) with the same
// conditioning as the reference) and turns the per-(pixel, Gaussian) work into 10 FMA + rcp + ex2.
//
//   GaussRec (64 B, one per camera-Gaussian pair): vx, vy, w2 (columns of M R^T scaled by 1/fx, 1/fy, 1),
//                                                  gro, opacity, rgb
//   per sorted (tile, Gaussian) instance, computed inside the blend kernels from a gather of the GaussRec:
//        N'(dx,dy) = n0 + n1x dx + n1y dy + n2xx dx^2 + n2xy dx dy + n2yy dy^2   (already x -1/2 log2 e)
//        D (dx,dy) = d0 + d1x dx + ...                                             (> 0)
//        alpha = min(0.999, opacity * 2^(N'/D)),  dx,dy = pixel centre - tile centre in [-7.5, 7.5]
//   Ckpt     (16 B per pixel per 32-instance bucket): (rgb accumulated so far, T) at the bucket start,
//            written by the forward, consumed by the backward (one warp per bucket, lane = instance, pixel
//            state rotating through the lanes: per-Gaussian gradients live in registers, no reduction).
#pragma once
#include "common.cuh"

namespace lfs {

struct alignas(16) GaussRec {
    float vx[3], vy[3], w2[3], gro[3];
    float opacity;
    float rgb[3];
};
static_assert(sizeof(GaussRec) == 64, "GaussRec must be 64 bytes");

constexpr float kNScale = -0.72134752044448170368f; // -0.5 * log2(e)
constexpr float kLn2 = 0.69314718055994530942f;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTMin = 1e-4f;

struct RasterBuffers {
    // inputs
    const GaussRec* gauss;     // [C*N]
    const int32_t* tile_off;   // [n_tiles_total + 1] (start of every (camera, tile); last = n_inst)
    const int32_t* inst_gid;   // [n_inst] flattened (camera*N + gaussian) id per sorted instance
    // per-bucket scratch
    uint32_t* bucket_off;      // [n_tiles_total + 1] exclusive scan of ceil(count/32)
    uint32_t* bucket_tile;     // [n_bucket_cap]
    float4* ckpt;              // [n_bucket_cap * 256]
    uint32_t* tile_max_contrib; // [n_tiles_total]
    uint32_t* live;            // [2 + n_bucket_cap] or null: [0] number of buckets some pixel reaches, [1] work counter of the
                               // backward, [2..] their ids (appended by the forward, tile by tile)
    // per-pixel state
    float4* pix_state;         // [C*H*W] (rgb before background, T_final)
    int32_t* n_contrib;        // [C*H*W] tile-local index + 1 of the last contributor (0 = none)
};

// Tunables (process-wide; set through lfs_set_option).  The *_variant switches select kernels that compute the same
// thing (A/B measurements); they may be changed at any time.
struct RasterOptions {
    int fwd_variant = 0; // forward blend: 0 TMA-gathered records (default), 1 register-staged gather (round 1), 2 = 0 at 80 regs
    int bwd_variant = 0; // backward blend: 0 software-pipelined, persistent over live buckets (default), 1 lock-step (round 1),
                         // 2 = 0 with one warp per bucket, 3 = 0 at 80 registers / 6 CTAs per SM
    int pre_bwd_split = 1; // trainer: SH / geometry halves of the per-Gaussian backward as two launches (A/B switch)
    int fg_variant = 0; // fastgs surface: 0 one thread per primitive (default), 1 warp-cooperative tile tests + emission (A/B)
    int exact_cull = 1; // trainer: drop (tile, Gaussian) instances that provably hold no alpha >= 1/255 (intersect.cuh CullRec)
};
RasterOptions& raster_options();
static inline size_t live_list_words(uint32_t n_bucket_cap) { return 2 + (size_t)n_bucket_cap; }

int launch_bucket_offsets(const RasterBuffers& rb, uint32_t n_tiles_total, uint32_t* n_buckets_dev, void* scan_scratch,
                          uint32_t* counts_tmp, cudaStream_t stream);

// forward blend. renders/alphas/last_ids (gsplat layouts) are optional; backgrounds [C,3] / masks optional.
int launch_blend_fwd(const RasterBuffers& rb, const ViewCam* cams_dev, uint32_t C, uint32_t width, uint32_t height,
                     uint32_t tile_w, uint32_t tile_h, bool write_ckpt, const float* backgrounds, const uint8_t* masks, float* renders,
                     float* alphas, int32_t* last_ids, cudaStream_t stream);

// backward blend: v_pix [C*H*W] = (dL/d rgb, T_final * (dL/d alpha - <bg, dL/d rgb>)).
// Accumulates (atomics) into v_means [N,3], v_quats [N,4], v_scales [N,3], v_colors [C*N,3], v_opacities [C*N].
int launch_blend_bwd(const RasterBuffers& rb, const ViewCam* cams_dev, const float4* v_pix, const float* quats,
                     const float* scales, const float* means, uint32_t C, uint32_t N, uint32_t width, uint32_t height,
                     uint32_t tile_w, uint32_t tile_h, uint32_t n_bucket_cap, const uint32_t* n_buckets_dev,
                     float* v_means, float* v_quats, float* v_scales, float* v_colors, float* v_opacities,
                     cudaStream_t stream);

// fastgs (EWA) surface: rb.gauss holds EWA records (see expand_record_ewa), one camera, outputs stay in rb.pix_state /
// rb.n_contrib / rb.ckpt.  Backward accumulates grad_mean2d [N,2], grad_conic [N,3] (true d/db), grad_color [N,3]
// (already masked by the clamp) and grad_raw_opacity [N].
int launch_blend_fwd_ewa(const RasterBuffers& rb, uint32_t width, uint32_t height, uint32_t tile_w, uint32_t tile_h,
                         bool write_ckpt, cudaStream_t stream);
int launch_blend_bwd_ewa(const RasterBuffers& rb, const float4* v_pix, uint32_t N, uint32_t width, uint32_t height,
                         uint32_t tile_w, uint32_t tile_h, uint32_t n_bucket_cap, const uint32_t* n_buckets_dev,
                         float* v_mean2d, float* v_conic, float* v_color, float* v_raw_opacity, cudaStream_t stream);

// ---- arbitrary per-pixel rays (raster_rays.cu): OpenCV distortion, fisheye, rolling shutter --------------------------
// `scratch` (rays_scratch_bytes) holds one RayRec per pixel and one GenRec per (camera, Gaussian)
size_t rays_scratch_bytes(uint32_t C, uint32_t N, uint64_t n_pix);
int launch_rays_prepare(void* scratch, const float* means, const float* quats, const float* scales, const float* colors,
                        const float* opacities, uint32_t N, uint32_t C, uint32_t channels, uint32_t ch0, uint32_t width,
                        uint32_t height, const float* viewmats0, const float* viewmats1, const float* Ks, int camera_model,
                        int rs_type, const float* radial, const float* tangential, const float* prism, bool make_rays,
                        cudaStream_t stream);
int launch_blend_fwd_rays(const RasterBuffers& rb, void* scratch, uint32_t C, uint32_t N, uint32_t width, uint32_t height,
                          uint32_t tile_w, uint32_t tile_h, bool write_ckpt, const float* backgrounds, const uint8_t* masks,
                          float* renders, float* alphas, int32_t* last_ids, cudaStream_t stream);
int launch_blend_bwd_rays(const RasterBuffers& rb, void* scratch, const float4* v_pix, const float* quats,
                          const float* scales, const float* means, uint32_t C, uint32_t N, uint32_t width, uint32_t height,
                          uint32_t tile_w, uint32_t tile_h, uint32_t n_bucket_cap, const uint32_t* n_buckets_dev,
                          float* v_means, float* v_quats, float* v_scales, float* v_colors, float* v_opacities,
                          cudaStream_t stream);

} // namespace lfs
