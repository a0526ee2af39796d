/* lfs_b200 -- C ABI of the B200-native 3DGS training rasterizer (sm_100a).
 *
 * This header is the drop-in boundary.  The reference (MrNeRF/LichtFeld-Studio) has no C ABI of its own:
 * its boundary is the C++ operator surface of two static archives (gsplat_backend / fastgs_backend).
 * Every entry point below names the reference interface it replaces (file:line, paths relative to the
 * reference repo root).  The thin C++/libtorch dispatch that re-exports the reference's exact symbols
 * (namespace gsplat / fast_gs) on top of this ABI lives in lichtfeld-studio_b200/host/ and the binding a
 * maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every data pointer is a DEVICE pointer unless the name
 * ends in _host; all tensors are contiguous, fp32 unless noted; `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  Functions return 0 on success or a negative lfs_status and never
 * fall back to a CPU path.  lfs_last_error() gives a human readable message for the calling thread.
 */
#ifndef LFS_B200_H
#define LFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFS_ABI_VERSION 2

typedef enum lfs_status {
    LFS_OK = 0,
    LFS_ERR_INVALID_ARG = -1,
    LFS_ERR_UNSUPPORTED = -2, /* valid in the reference, not implemented here: fails loudly */
    LFS_ERR_CUDA = -3,
    LFS_ERR_ALLOC = -4,
    LFS_ERR_CAPACITY = -5 /* a capacity-bounded buffer overflowed; grow and retry */
} lfs_status;

/* gsplat/Common.h:46-50 */
typedef enum lfs_camera_model { LFS_PINHOLE = 0, LFS_ORTHO = 1, LFS_FISHEYE = 2 } lfs_camera_model;
/* gsplat/Cameras.h:16-22 */
typedef enum lfs_shutter_type {
    LFS_ROLLING_TOP_TO_BOTTOM = 0,
    LFS_ROLLING_LEFT_TO_RIGHT = 1,
    LFS_ROLLING_BOTTOM_TO_TOP = 2,
    LFS_ROLLING_RIGHT_TO_LEFT = 3,
    LFS_GLOBAL = 4
} lfs_shutter_type;
/* gsplat/Cameras.h:27-44 UnscentedTransformParameters (same defaults) */
typedef struct lfs_ut_params {
    float alpha;                  /* 0.1 */
    float beta;                   /* 2.0 */
    float kappa;                  /* 0.0 */
    float in_image_margin_factor; /* 0.1 */
    int32_t require_all_sigma_points_valid; /* 1 */
} lfs_ut_params;

/* Output / scratch allocation callback.  Plays the role of the at::empty calls inside the reference ops
 * (gsplat/Intersect.cpp:82-85) and of fastgs' resize callbacks std::function<char*(size_t)>
 * (fastgs/rasterization/include/forward.h:14-17).  Must return a device pointer aligned to 256 bytes that
 * stays valid until the caller releases it; `tag` identifies the buffer (see each function). */
typedef void* (*lfs_alloc_fn)(void* ctx, int tag, size_t bytes);

const char* lfs_last_error(void);
int lfs_abi_version(void);
/* number of kernels launched by this library on behalf of the calling process (bench.py gpu_launches) */
uint64_t lfs_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * gsplat operator surface (3DGUT path)
 * ------------------------------------------------------------------------------------------------------- */

/* Replaces gsplat::projection_ut_3dgs_fused / launch_projection_ut_3dgs_fused_kernel
 * (gsplat/Ops.h:69-98, gsplat/Projection.h:12-40, kernel gsplat/ProjectionUT3DGSFused.cu:17-203).
 * means [N,3], quats [N,4] wxyz, scales [N,3], opacities [N] or NULL, viewmats0 [C,4,4] row-major w2c,
 * Ks [C,3,3].  Outputs radii [C,N,2] i32, means2d [C,N,2], depths [C,N], conics [C,N,3],
 * compensations [C,N] or NULL.  Culled entries: radii = 0, other outputs untouched (as the reference).
 * Camera models: LFS_PINHOLE with optional OpenCV distortion (radial_coeffs [C,6], tangential_coeffs [C,2],
 * thin_prism_coeffs [C,4], any of them NULL), LFS_FISHEYE (radial_coeffs [C,4]; tangential / thin prism must be
 * NULL as in the reference's fisheye model), and every rolling-shutter type with viewmats1 [C,4,4] as the pose
 * at the end of the exposure (gsplat/Cameras.cuh semantics).  LFS_ORTHO returns LFS_ERR_UNSUPPORTED (the
 * reference's UT projection has no orthographic model either). */
int lfs_projection_ut_3dgs_fused(const float* means, const float* quats, const float* scales,
                                 const float* opacities, const float* viewmats0, const float* viewmats1,
                                 const float* Ks, uint32_t N, uint32_t C, uint32_t image_width,
                                 uint32_t image_height, float eps2d, float near_plane, float far_plane,
                                 float radius_clip, int camera_model, const lfs_ut_params* ut_params, int rs_type,
                                 const float* radial_coeffs, const float* tangential_coeffs,
                                 const float* thin_prism_coeffs, int32_t* radii, float* means2d, float* depths,
                                 float* conics, float* compensations, void* stream);

/* Replaces gsplat::spherical_harmonics_fwd / launch_spherical_harmonics_fwd_kernel
 * (gsplat/Ops.h:12-17, gsplat/SphericalHarmonics.h:11-19, kernel gsplat/SphericalHarmonicsCUDA.cu:374-399).
 * dirs [n,3], coeffs [n,K,3], masks [n] (bool bytes) or NULL -> colors [n,3] (masked rows untouched). */
int lfs_spherical_harmonics_fwd(uint32_t degrees_to_use, const float* dirs, const float* coeffs,
                                const uint8_t* masks, uint32_t n, uint32_t K, float* colors, void* stream);

/* Replaces gsplat::spherical_harmonics_bwd (gsplat/Ops.h:18-25, kernel gsplat/SphericalHarmonicsCUDA.cu:445-481).
 * Writes v_coeffs [n,K,3] completely (zeros for inactive bases and masked rows, i.e. the reference's
 * at::zeros_like + kernel); v_dirs [n,3] or NULL likewise. */
int lfs_spherical_harmonics_bwd(uint32_t K, uint32_t degrees_to_use, const float* dirs, const float* coeffs,
                                const uint8_t* masks, const float* v_colors, uint32_t n, float* v_coeffs,
                                float* v_dirs, void* stream);

/* Replaces gsplat::intersect_tile (gsplat/Ops.h:28-38, gsplat/Intersect.cpp:15-122; kernels
 * gsplat/IntersectTile.cu:24-114 and the CUB sort :290-328).  [C,N] layout; the packed layout is lfs_intersect_tile_packed.
 * means2d [C,N,2], radii [C,N,2] i32, depths [C,N] -> tiles_per_gauss [C,N] i32 (caller allocated),
 * and, through `alloc`, isect_ids [n_isects] i64 (tag 1) and flatten_ids [n_isects] i32 (tag 2); scratch is
 * requested with tag 0 and may be released when the call returns.  Like the reference this call blocks
 * once to read n_isects.  Outputs are bit-identical to the reference for identical inputs. */
#define LFS_TAG_SCRATCH 0
#define LFS_TAG_ISECT_IDS 1
#define LFS_TAG_FLATTEN_IDS 2
int lfs_intersect_tile(const float* means2d, const int32_t* radii, const float* depths, uint32_t C, uint32_t N,
                       uint32_t tile_size, uint32_t tile_width, uint32_t tile_height, int sort,
                       int32_t* tiles_per_gauss, lfs_alloc_fn alloc, void* alloc_ctx, int64_t** isect_ids,
                       int32_t** flatten_ids, int64_t* n_isects_host, void* stream);

/* The packed layout of gsplat::intersect_tile (gsplat/Intersect.cpp:32-39, kernel IntersectTile.cu:83-90): means2d [nnz,2],
 * radii [nnz,2], depths [nnz], camera_ids [nnz] i64 (the camera of every element; gaussian_ids is not needed, as in the
 * reference kernel); flatten_ids index [nnz].  Same outputs, same key layout (camera | tile | depth bits), same order as
 * the reference's 64-bit sort.  floor(log2 C) + floor(log2 n_tiles) + 2 <= 32 bits (the reference asserts the same). */
int lfs_intersect_tile_packed(const float* means2d, const int32_t* radii, const float* depths, const int64_t* camera_ids,
                              uint32_t nnz, uint32_t C, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                              int sort, int32_t* tiles_per_gauss, lfs_alloc_fn alloc, void* alloc_ctx, int64_t** isect_ids,
                              int32_t** flatten_ids, int64_t* n_isects_host, void* stream);

/* Replaces gsplat::intersect_offset (gsplat/Ops.h:39-43, kernel gsplat/IntersectTile.cu:206-252).
 * offsets [C, tile_height, tile_width] i32. */
int lfs_intersect_offset(const int64_t* isect_ids, int64_t n_isects, uint32_t C, uint32_t tile_width,
                         uint32_t tile_height, int32_t* offsets, void* stream);

/* Replaces gsplat::rasterize_to_pixels_from_world_3dgs_fwd / launch_..._fwd_kernel<CDIM>
 * (gsplat/Ops.h:100-129, gsplat/Rasterization.h:100-135, kernel gsplat/RasterizeToPixelsFromWorld3DGSFwd.cu:20-279).
 * means [N,3], quats [N,4] (normalised), scales [N,3], colors [C,N,channels], opacities [C,N],
 * backgrounds [C,channels] or NULL, masks [C,th,tw] bool bytes or NULL, tile_offsets [C,th,tw] i32,
 * flatten_ids [n_isects] i32 -> renders [C,H,W,channels], alphas [C,H,W,1], last_ids [C,H,W] i32.
 * tile_size == 16.  Any `channels` >= 1 (processed in groups of three through the same kernels; the
 * reference instantiates CDIM 1..5, 8, 9, 16, 17, 32, 33, ...).  Undistorted global-shutter pinhole cameras take
 * the per-tile polynomial kernels; distortion, fisheye and rolling shutter take the per-pixel-ray kernels
 * (csrc/raster_rays.cu); LFS_ORTHO returns LFS_ERR_UNSUPPORTED.  For C > 1 the gaussian id of a flattened
 * index g is g % N (the reference kernel is single-camera).
 * Scratch (per-gaussian records, pixel rays) is requested through `alloc` with tag 0. */
int lfs_rasterize_to_pixels_from_world_3dgs_fwd(
    const float* means, const float* quats, const float* scales, const float* colors, const float* opacities,
    const float* backgrounds, const uint8_t* masks, uint32_t N, uint32_t C, uint32_t channels,
    uint32_t image_width, uint32_t image_height, uint32_t tile_size, const float* viewmats0,
    const float* viewmats1, const float* Ks, int camera_model, const lfs_ut_params* ut_params, int rs_type,
    const float* radial_coeffs, const float* tangential_coeffs, const float* thin_prism_coeffs,
    const int32_t* tile_offsets, const int32_t* flatten_ids, int64_t n_isects, lfs_alloc_fn alloc,
    void* alloc_ctx, float* renders, float* alphas, int32_t* last_ids, void* stream);

/* Replaces gsplat::rasterize_to_pixels_from_world_3dgs_bwd / launch_..._bwd_kernel<CDIM>
 * (gsplat/Ops.h:131-166, gsplat/Rasterization.h:137-174, kernel gsplat/RasterizeToPixelsFromWorld3DGSBwd.cu:17-372).
 * Outputs v_means [N,3], v_quats [N,4], v_scales [N,3], v_colors [C,N,channels], v_opacities [C,N] are fully
 * written (zero-filled then accumulated, like the reference's at::zeros_like + atomics). */
int lfs_rasterize_to_pixels_from_world_3dgs_bwd(
    const float* means, const float* quats, const float* scales, const float* colors, const float* opacities,
    const float* backgrounds, const uint8_t* masks, uint32_t N, uint32_t C, uint32_t channels,
    uint32_t image_width,
    uint32_t image_height, uint32_t tile_size, const float* viewmats0, const float* viewmats1, const float* Ks,
    int camera_model, const lfs_ut_params* ut_params, int rs_type, const float* radial_coeffs,
    const float* tangential_coeffs, const float* thin_prism_coeffs, const int32_t* tile_offsets,
    const int32_t* flatten_ids, int64_t n_isects, const float* render_alphas, const int32_t* last_ids,
    const float* v_render_colors, const float* v_render_alphas, lfs_alloc_fn alloc, void* alloc_ctx,
    float* v_means, float* v_quats, float* v_scales, float* v_colors, float* v_opacities, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * fused Adam
 * ------------------------------------------------------------------------------------------------------- */

/* Replaces fast_gs::optimizer::adam_step / adam_step_wrapper
 * (fastgs/optimizer/include/adam.h:9-20, adam_api.h:11-21, kernel adam_kernels.cuh:13-36).  Same formula,
 * same argument meaning (bias corrections passed as reciprocals computed in double by the caller,
 * src/training/optimizers/fused_adam.cpp:78-79). */
int lfs_adam_step(float* param, float* exp_avg, float* exp_avg_sq, const float* param_grad, int64_t n_elements,
                  float lr, float beta1, float beta2, float eps, float bias_correction1_rcp,
                  float bias_correction2_sqrt_rcp, void* stream);

/* Multi-tensor form: one launch over a flat arena made of `n_segments` consecutive segments
 * (segment s covers elements [seg_begin[s], seg_begin[s+1]) and uses lr[s], bc1_rcp[s], bc2_sqrt_rcp[s]);
 * replaces the 6 per-group launches of FusedAdam::step (src/training/optimizers/fused_adam.cpp:22-95).
 * Segment boundaries must be multiples of 4 elements.  If zero_grad != 0 the gradient arena is cleared in
 * the same pass (the reference's zero_grad(set_to_none) + torch::zeros of the next backward).
 * seg_begin_host [n_segments+1], lr_host/bc1_rcp_host/bc2_sqrt_rcp_host [n_segments] are HOST arrays,
 * n_segments <= 16. */
/* Optional regularisers folded into the update (no extra pass over the parameters): the mcmc configuration adds
 * scale_reg * mean(exp(scaling_raw)) and opacity_reg * mean(sigmoid(opacity_raw)) to the loss of every iteration
 * (src/training/trainer.cpp:132-158).  Their gradients depend on the parameter alone, so segment s adds
 *   kind 1: coef * exp(p)        kind 2: coef * sigmoid(p) * (1 - sigmoid(p))        kind 0: nothing
 * to the (summed) gradient of every element before the Adam update; coef = weight * n_views / n_elements_of_the_mean.
 * Planar arenas pad every plane to plane_elems floats: only the first n_valid elements of a plane are regularised. */
typedef struct lfs_adam_reg {
    int kind[16];
    float coef[16];
    int64_t plane_elems;
    int64_t n_valid;
} lfs_adam_reg;
int lfs_adam_step_multi(float* params, float* exp_avg, float* exp_avg_sq, float* grads, int n_segments,
                        const int64_t* seg_begin_host, const float* lr_host, const float* bc1_rcp_host,
                        const float* bc2_sqrt_rcp_host, float beta1, float beta2, float eps, int zero_grad,
                        const lfs_adam_reg* reg /* or NULL */, void* stream);

/* Multi-GPU form of lfs_adam_step_multi: reduce-scatter + Adam + all-gather fused in one kernel over
 * NVLink peer memory, replacing ncclAllReduce(gradient arena) + a full Adam on every rank.  grads_peers_dev /
 * params_peers_dev are DEVICE arrays of `world` pointers to every rank's (peer-mapped) gradient / parameter arena, e.g.
 * torch.distributed._symmetric_memory buffer_ptrs_dev; exp_avg / exp_avg_sq are this rank's.  Ownership is fixed per
 * arena element: 16-byte unit i (absolute, from the start of the arena) belongs to rank (i / 1024) % world, whatever
 * segments a call covers -- a rank holds valid Adam moments only for the elements it owns, and it reduces, updates and
 * broadcasts exactly those of [seg_begin[0], seg_begin[n]).  When grads_multicast / params_multicast (NVSwitch multicast addresses of the same buffers, e.g. symmetric
 * memory multicast_ptr; NULL if unsupported) are given, the reduction happens inside the switch (multimem.ld_reduce) and
 * the parameters are broadcast with one multimem.st; params_local is then this rank's own arena.  The caller issues a
 * stream-ordered barrier over all ranks before (all backward passes finished) and after (all parameter writes landed)
 * this call and clears its own gradient arena after the second barrier. */
int lfs_adam_step_multi_p2p(float* exp_avg, float* exp_avg_sq, const void* grads_peers_dev, const void* params_peers_dev,
                            const float* grads_multicast, float* params_multicast, float* params_local, int world,
                            int rank, int n_segments, const int64_t* seg_begin_host, const float* lr_host,
                            const float* bc1_rcp_host, const float* bc2_sqrt_rcp_host, float beta1, float beta2,
                            float eps, const lfs_adam_reg* reg /* or NULL */, void* stream);
/* Ownership rule of lfs_adam_step_multi_p2p, host-only (no CUDA call): the rank that keeps the Adam moments of the arena
 * element `float_index`, and the chunks (each *chunk_floats floats long, chunk c covering floats [c*chunk_floats, ...))
 * of [begin_float, end_float) that `rank` owns: first_chunk, first_chunk + world, ... (n_chunks of them). */
int lfs_adam_p2p_owner(int64_t float_index, int world);
int lfs_adam_p2p_owned_chunks(int64_t begin_float, int64_t end_float, int world, int rank, int64_t* first_chunk,
                              int64_t* n_chunks, int64_t* chunk_floats);

/* ---------------------------------------------------------------------------------------------------------
 * small per-Gaussian ops of the densification strategies (off the steady-state hot path)
 * ------------------------------------------------------------------------------------------------------- */
/* gsplat::quats_to_rotmats (gsplat/Ops.h:46-48): quats [N,4] wxyz -> rotmats [N,3,3] row-major */
int lfs_quats_to_rotmats(const float* quats, uint32_t n, float* rotmats, void* stream);
/* gsplat::relocation (gsplat/Ops.h:52-57): opacities [N], scales [N,3], ratios [N] i32, binoms [n_max,n_max] */
int lfs_relocation(const float* opacities, const float* scales, const int32_t* ratios, const float* binoms,
                   int n_max, uint32_t n, float* new_opacities, float* new_scales, void* stream);
/* gsplat::add_noise (gsplat/Ops.h:59-65): means [N,3] updated in place */
int lfs_add_noise(const float* raw_opacities, const float* raw_scales, const float* raw_quats, const float* noise,
                  float* means, float current_lr, uint32_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * legacy 2-D operator surface: the ops the reference's gtest files call
 * (tests/test_basic.cpp:54-128,347-358; tests/test_gsplat_ops.cpp:76-96,174-189,276-287; tests/test_rasterization.cpp:
 * 110-121) and whose launchers gsplat/Rasterization.h:16-63 still declares, but whose kernels are no longer in the
 * reference tree.  Semantics: tests/torch_impl.cpp:38-218 (the reference's CPU statement), the culling tail of
 * ProjectionUT3DGSFused.cu:142-199, and the blend loop of RasterizeToPixelsFromWorld3DGS{Fwd,Bwd}.cu with the 2-D conic
 * response sigma = 1/2 (a dx^2 + c dy^2) + b dx dy.
 * ------------------------------------------------------------------------------------------------------- */

/* quats [N,4] wxyz (normalised inside), scales [N,3] -> covars / precis [N,3,3], or [N,6] (xx,xy,xz,yy,yz,zz) if triu.
 * Either output may be NULL (the reference's compute_covar / compute_preci flags). */
int lfs_quat_scale_to_covar_preci_fwd(const float* quats, const float* scales, uint32_t N, int triu, float* covars,
                                      float* precis, void* stream);
/* v_covars / v_precis in the forward's layout (either may be NULL) -> v_quats [N,4], v_scales [N,3] (overwritten). */
int lfs_quat_scale_to_covar_preci_bwd(const float* quats, const float* scales, uint32_t N, int triu,
                                      const float* v_covars, const float* v_precis, float* v_quats, float* v_scales,
                                      void* stream);

/* Fused pinhole EWA projection.  means [N,3]; covars [N,3,3] or NULL (then quats [N,4] + scales [N,3]); opacities [N] or
 * NULL (NULL: fixed 3.33 sigma extent); viewmats [C,4,4]; Ks [C,3,3].  Outputs as lfs_projection_ut_3dgs_fused:
 * radii [C,N,2] i32 (always written, 0 = culled), means2d [C,N,2], depths [C,N], conics [C,N,3], compensations [C,N] or
 * NULL; culled entries of the float outputs are left untouched.  camera_model must be LFS_PINHOLE. */
int lfs_projection_ewa_3dgs_fused_fwd(const float* means, const float* covars, const float* quats, const float* scales,
                                      const float* opacities, const float* viewmats, const float* Ks, uint32_t N,
                                      uint32_t C, uint32_t image_width, uint32_t image_height, float eps2d,
                                      float near_plane, float far_plane, float radius_clip, int camera_model,
                                      int32_t* radii, float* means2d, float* depths, float* conics, float* compensations,
                                      void* stream);

/* 2-D alpha blend (gsplat/Rasterization.h:16-36).  means2d [C,N,2], conics [C,N,3], colors [C,N,channels],
 * opacities [C,N], backgrounds [C,channels] or NULL, masks [C,th,tw] bool bytes or NULL, tile_offsets [C,th,tw],
 * flatten_ids [n_isects] (indices into [C*N]) -> renders [C,H,W,channels], alphas [C,H,W,1], last_ids [C,H,W].
 * tile_size == 16, 1 <= channels <= 40. */
int lfs_rasterize_to_pixels_3dgs_fwd(const float* means2d, const float* conics, const float* colors,
                                     const float* opacities, const float* backgrounds, const uint8_t* masks, uint32_t C,
                                     uint32_t N, uint32_t channels, uint32_t image_width, uint32_t image_height,
                                     uint32_t tile_size, const int32_t* tile_offsets, const int32_t* flatten_ids,
                                     int64_t n_isects, float* renders, float* alphas, int32_t* last_ids, void* stream);
/* Backward (gsplat/Rasterization.h:38-63).  Outputs are fully written (zero-filled, then accumulated):
 * v_means2d_abs [C,N,2] or NULL (the absgrad variant), v_means2d [C,N,2], v_conics [C,N,3], v_colors [C,N,channels],
 * v_opacities [C,N]. */
int lfs_rasterize_to_pixels_3dgs_bwd(const float* means2d, const float* conics, const float* colors,
                                     const float* opacities, const float* backgrounds, const uint8_t* masks, uint32_t C,
                                     uint32_t N, uint32_t channels, uint32_t image_width, uint32_t image_height,
                                     uint32_t tile_size, const int32_t* tile_offsets, const int32_t* flatten_ids,
                                     int64_t n_isects, const float* render_alphas, const int32_t* last_ids,
                                     const float* v_render_colors, const float* v_render_alphas, float* v_means2d_abs,
                                     float* v_means2d, float* v_conics, float* v_colors, float* v_opacities, void* stream);

/* ---- fastgs (EWA) rasterizer surface ----------------------------------------------------------------------
 * Replaces fast_gs::rasterization::forward (fastgs/rasterization/include/forward.h:13-38, src/forward.cu:15-199;
 * called by forward_wrapper, src/rasterization_api.cu:15-89, which fast_rasterize() binds,
 * src/training/rasterization/fast_rasterizer.cpp:12-60).  RAW parameters, reference layouts:
 * means [N,3], scales_raw [N,3] (log), rotations_raw [N,4] (w,x,y,z, not normalised), opacities_raw [N,1] (logit),
 * sh_coefficients_0 [N,1,3], sh_coefficients_rest [N,total_bases_sh_rest,3], w2c [4,4] row-major on the DEVICE,
 * cam_position [3] on the DEVICE; active_sh_bases in {1,4,9,16}.  Outputs image [3,H,W] (no background), alpha [1,H,W].
 * The four state buffers of the reference (buffer_utils.h:38-151) are requested through `alloc` with the tags below,
 * must stay valid and unmodified until the matching backward (fast_rasterizer_autograd.cpp:61-72), and their contents
 * are private to this library.  Like the reference the call blocks twice to read n_instances and n_buckets.
 * Equal-depth primitives are ordered by index (the reference's order there is non-deterministic).  Launches on
 * `stream` (the reference uses the legacy default stream and a process-global memset stream; this is re-entrant). */
#define LFS_TAG_FG_PER_PRIMITIVE 10
#define LFS_TAG_FG_PER_TILE 11
#define LFS_TAG_FG_PER_INSTANCE 12
#define LFS_TAG_FG_PER_BUCKET 13
int lfs_fastgs_forward(const float* means, const float* scales_raw, const float* rotations_raw,
                       const float* opacities_raw, const float* sh_coefficients_0, const float* sh_coefficients_rest,
                       const float* w2c, const float* cam_position, uint32_t n_primitives, int active_sh_bases,
                       int total_bases_sh_rest, int width, int height, float focal_x, float focal_y, float center_x,
                       float center_y, float near_plane, float far_plane, float* image, float* alpha,
                       lfs_alloc_fn alloc, void* alloc_ctx, int* n_visible_primitives, int* n_instances,
                       int* n_buckets, int* instance_selector, void* stream);

/* Replaces fast_gs::rasterization::backward (include/backward.h:13-52, src/backward.cu:14-116; called by
 * backward_wrapper, src/rasterization_api.cu:91-181).  grad_image [3,H,W], grad_alpha [1,H,W]; the four buffers and
 * the three counters come from the matching lfs_fastgs_forward.  Writes ALL N rows of grad_means [N,3],
 * grad_scales_raw [N,3], grad_rotations_raw [N,4], grad_opacities_raw [N,1], grad_sh_coefficients_0 [N,1,3],
 * grad_sh_coefficients_rest [N,total,3] (zeros for primitives that touched no tile; the reference relies on
 * torch::zeros for those).  grad_w2c [4,4] (nullable) is ACCUMULATED into rows 0-2 exactly as the reference does
 * (kernels_backward.cuh:162-175); densification_info [2,N] (nullable) is accumulated (:233-236).
 * Scratch is requested through `alloc` with LFS_TAG_SCRATCH. */
int lfs_fastgs_backward(const float* grad_image, const float* grad_alpha, const float* means, const float* scales_raw,
                        const float* rotations_raw, const float* sh_coefficients_rest, const float* w2c,
                        const float* cam_position, const void* per_primitive_buffers, const void* per_tile_buffers,
                        const void* per_instance_buffers, const void* per_bucket_buffers, float* grad_means,
                        float* grad_scales_raw, float* grad_rotations_raw, float* grad_opacities_raw,
                        float* grad_sh_coefficients_0, float* grad_sh_coefficients_rest, float* grad_w2c,
                        float* densification_info, uint32_t n_primitives, int n_instances, int n_buckets,
                        int instance_selector, int active_sh_bases, int total_bases_sh_rest, int width, int height,
                        float focal_x, float focal_y, float center_x, float center_y, lfs_alloc_fn alloc,
                        void* alloc_ctx, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * fused training step (the fast path; what bench.py times)
 *
 * Plays the role of gs::training::rasterize + GUTRasterizationFunction / SphericalHarmonicsFunction forward
 * and backward (src/training/rasterization/rasterizer.cpp:46-437, rasterizer_autograd.cpp:12-392) with the
 * SplatData activations (src/core/splat_data.cpp:267-286) fused in; one call per view, no host sync.
 * Parameter / gradient / Adam arenas are PLANAR fp32 buffers owned by the caller (so torch.distributed /
 * NCCL can all-reduce the gradient arena in place):
 *   plane order  means xyz | sh0 rgb | shN[(k-1)*3+ch], k=1..K-1 | scaling xyz | rotation wxyz | opacity
 *   plane pitch  N_pad = round_up(N, 4) floats;  total (11 + 3K) * N_pad floats, K = (sh_degree_max+1)^2.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct lfs_trainer_desc {
    uint32_t n_gaussians;
    uint32_t sh_degree_max; /* 0..4; K = (d+1)^2 coefficients per channel */
    uint32_t width, height;
    float eps2d;       /* 0.3   (rasterizer.cpp:176) */
    float near_plane;  /* 0.01  (rasterizer.cpp:177) */
    float far_plane;   /* 1e4   (rasterizer.cpp:178) */
    float radius_clip; /* 0     (rasterizer.cpp:179) */
    lfs_ut_params ut;
    uint64_t instance_capacity; /* max tile-Gaussian instances per view; 0 = 8*N + 64K */
} lfs_trainer_desc;

typedef enum lfs_image_format { LFS_IMG_U8_HWC = 0, LFS_IMG_F32_HWC = 1, LFS_IMG_F32_CHW = 2 } lfs_image_format;

uint64_t lfs_trainer_arena_floats(const lfs_trainer_desc* desc);
void* lfs_trainer_create(const lfs_trainer_desc* desc); /* NULL on failure (see lfs_last_error) */
void lfs_trainer_destroy(void* trainer);
uint64_t lfs_trainer_scratch_bytes(void* trainer);
uint64_t lfs_trainer_instance_capacity(void* trainer);

/* reference SplatData layout (means [N,3], sh0 [N,1,3], shN [N,K-1,3], scaling [N,3], rotation [N,4],
 * opacity [N,1]; include/core/splat_data.hpp:104-109)  <->  planar arena */
int lfs_trainer_pack(void* trainer, const float* means, const float* sh0, const float* shN, const float* scaling,
                     const float* rotation, const float* opacity, float* arena, void* stream);
int lfs_trainer_unpack(void* trainer, const float* arena, float* means, float* sh0, float* shN, float* scaling,
                       float* rotation, float* opacity, void* stream);

/* forward of one view.  viewmat_host [16] row-major w2c and K_host [9] are HOST arrays (the reference keeps
 * cameras on the host side too, include/core/camera.hpp:53-65); bg_host [3] or NULL.
 * image_out [H,W,3] / alpha_out [H,W] are optional device outputs (image = rgb + (1-alpha)*bg). */
int lfs_trainer_view_forward(void* trainer, const float* params_arena, const float* viewmat_host,
                             const float* K_host, uint32_t active_sh_degree, const float* bg_host, float* image_out,
                             float* alpha_out, void* stream);
/* L1 photometric loss of the last forward against `target` (device); loss_accum (device float, may be NULL)
 * += scale * sum|clamp(render,0,1) - target|;  sets the upstream gradient for lfs_trainer_view_backward. */
int lfs_trainer_view_loss_l1(void* trainer, const void* target, int target_format, float scale, float* loss_accum,
                             void* stream);
/* L1 + fused-SSIM photometric loss of the last forward, the reference's training loss
 * (Trainer::compute_photometric_loss, src/training/trainer.cpp:103-131: (1-lambda) * l1_loss + lambda * (1 - fused_ssim(
 * rendered, gt, "valid")); include/kernels/fused_ssim.cuh:27-122; kernels src/training/kernels/ssim.cu:64-460).
 * The render is clamped to [0,1] first (rasterizer.cpp:401).  loss_accum (device float, may be NULL) += weight * loss;
 * sets the upstream gradient (scaled by weight) for lfs_trainer_view_backward. */
int lfs_trainer_view_loss_ssim_l1(void* trainer, const void* target, int target_format, float lambda_dssim, float weight,
                                  float* loss_accum, void* stream);
/* alternative: caller-provided upstream gradients v_image [H,W,3], v_alpha [H,W] or NULL (device) */
int lfs_trainer_view_set_grad(void* trainer, const float* v_image, const float* v_alpha, void* stream);
/* backward of the last forward: grads_arena += d loss / d raw parameters */
int lfs_trainer_view_backward(void* trainer, const float* params_arena, float* grads_arena, void* stream);
/* The same in two halves.  _blend (blend backward) only touches the handle's own scratch; _params (SH / projection /
 * activation VJPs) read-modify-writes grads_arena.  A caller that runs consecutive views on two streams, one trainer
 * handle per stream and both on the same arenas, orders only the _params halves with events (trainer.py train_step). */
int lfs_trainer_view_backward_blend(void* trainer, void* stream);
int lfs_trainer_view_backward_params(void* trainer, const float* params_arena, float* grads_arena, void* stream);
/* blocks on `stream`; returns LFS_ERR_CAPACITY if the last forward overflowed instance_capacity */
int lfs_trainer_stats(void* trainer, uint64_t* n_instances, uint64_t* n_buckets, void* stream);
/* Non-blocking overflow check for training loops: every forward records the largest instance count any view has needed
 * so far and copies it to pinned host memory; this call only reads that word.  Returns LFS_ERR_CAPACITY as soon as a view
 * whose instances did not fit (its farthest instances were dropped) has been observed -- call it once per step. */
int lfs_trainer_poll_capacity(void* trainer, uint64_t* max_instances_seen /* or NULL */);
/* Diagnostics / tests: copies one internal per-view buffer of the last forward into a caller DEVICE buffer (async on
 * `stream`; at most dst_bytes).  which: 0 tile_off int32[n_tiles+1] | 1 n_contrib int32[H*W] | 2 pix_state float4[H*W]
 * (rgb before background, T) | 3 sorted instance -> Gaussian id uint32[n_instances] | 4 sorted instance tile keys
 * uint32[n_instances] | 5 bucket_off uint32[n_tiles+1] | 6 tile_max_contrib uint32[n_tiles].  Returns the number of bytes
 * of the full buffer through *full_bytes (may be NULL). */
int lfs_trainer_debug_copy(void* trainer, int which, void* dst, uint64_t dst_bytes, uint64_t* full_bytes, void* stream);

/* ---- densification-side state surgery on the planar arenas -----------------------------------------
 * The reference's strategies edit its six parameter tensors and the Adam moments with index_select + cat per tensor
 * (src/training/strategies/default_strategy.cpp:49-230 duplicate / split / prune, mcmc.cpp:112-347 relocate / add,
 * strategy_utils.cpp:57-131).  On the planar layout every such edit is one ROW GATHER:
 *   output Gaussian i <- every plane of source Gaussian index[i]   (index[i] < 0: leave the destination slot as it is)
 * applied to the parameter arena and, when given, to both Adam moment arenas in the same launch; zero_state[i] != 0
 * zeroes the moments of slot i (a new Gaussian starts with fresh optimiser state, default_strategy.cpp:62-75).
 * Source and destination may be the same arenas only for in-place relocation (every written slot is read by nobody).
 * plane_elems = padded plane length (N_pad) of the respective arena.  All pointers are device pointers. */
int lfs_arena_gather(const float* src_params, const float* src_m, const float* src_v, float* dst_params, float* dst_m,
                     float* dst_v, const int32_t* index, const uint8_t* zero_state, uint32_t n_out, uint32_t n_src,
                     uint32_t planes, uint64_t src_plane_elems, uint64_t dst_plane_elems, void* stream);
/* AoS rows [n, n_planes] <-> slots index[n] (NULL: 0..n-1) of planes [first_plane, first_plane + n_planes): how a strategy
 * reads the activated inputs of a split and writes the new means / scales / opacities (default_strategy.cpp:107-128). */
int lfs_arena_set_rows(float* arena, uint64_t plane_elems, uint32_t first_plane, uint32_t n_planes, const int32_t* index,
                       const float* rows, uint32_t n, uint32_t n_slots, void* stream);
int lfs_arena_get_rows(const float* arena, uint64_t plane_elems, uint32_t first_plane, uint32_t n_planes,
                       const int32_t* index, float* rows, uint32_t n, uint32_t n_slots, void* stream);
/* Multi-GPU: lfs_adam_step_multi_p2p keeps the Adam moments of an element on its owner only.  Before a surgery every
 * rank zeroes what it does not own; an all-reduce(sum) of the result is the complete moment arena on every rank. */
int lfs_adam_p2p_zero_unowned(float* arena, int64_t n_floats, int world, int rank, void* stream);

/* Per-stage device timing of the view step with CUDA events on the launching stream (bench.py roofline).
 * Stages: 0 preprocess_fwd | 1 depth sort + scan + emit + tile sort + offsets | 2 bucket offsets + expand |
 *         3 blend_fwd | 4 (loss: caller side, always 0) | 5 blend_bwd | 6 preprocess_bwd.
 * Enabling it makes lfs_trainer_view_backward block on the stream; never enable it inside a timed region. */
#define LFS_PROF_STAGES 7
int lfs_trainer_set_profile(void* trainer, int enable);
int lfs_trainer_get_profile(void* trainer, float* mean_ms /* [LFS_PROF_STAGES] */, int* counts /* or NULL */);

/* library options (A/B switches between kernels that compute the same thing; the defaults are the measured best):
 *   "fwd_variant"  0 forward blend with TMA-gathered records (cp.async.bulk + mbarrier), 1 register-staged gather (round 1),
 *                  2 = 0 bounded to 80 registers (12 CTAs per SM)
 *   "bwd_variant"  0 software-pipelined backward blend on a persistent grid over the live-bucket list, 1 lock-step backward
 *                  blend (round 1), 2 software-pipelined with one warp per bucket (no list), 3 = 0 bounded to 80 registers
 *                  (6 CTAs per SM)
 *   "sort_variant" 0 histogram / scan / scatter radix passes, 1 onesweep (decoupled look-back) passes, 2 onesweep for keys
 *                  of <= 16 bits only, 3 the passes of 0 with ballot ranking, 4 = 2 + 3 (A/B switches, all bit-identical
 *                  results)
 *   "fg_variant"   fastgs surface: 0 one thread per primitive, 1 warp-cooperative exact tile tests and emission (identical
 *                  results, measured slower)
 *   "exact_cull"   1 trainer drops tile instances that provably hold no alpha >= 1/255, 0 the reference's AABB rule
 *   "pre_bwd_split" 1 per-Gaussian backward as two launches (SH, geometry), 0 one launch */
int lfs_set_option(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* LFS_B200_H */
