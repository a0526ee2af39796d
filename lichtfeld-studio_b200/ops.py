"""Host-side mirror of the reference's gsplat operator surface (gsplat/Ops.h:12-166) over torch tensors.

Same function names, argument meaning and error behaviour as the reference ops (inputs must be CUDA +
contiguous, else ValueError like the reference's CHECK_INPUT -> c10::Error, gsplat/Common.h:12-17); every
function forwards to the C ABI in include/lfs_b200.h.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ALLOC_FN, PINHOLE, SHUTTER_GLOBAL, UTParams, check, load


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    """Current torch stream of the CURRENT device; every op runs under _on_tensor_device, which makes the device of its
    first tensor argument current (the reference's ops do the same with a DEVICE_GUARD, gsplat/Common.h:18-19)."""
    return torch.cuda.current_stream().cuda_stream


def _on_tensor_device(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = next((a.device for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor) and a.is_cuda),
                   None)
        if dev is None:
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


class _Alloc:
    """lfs_alloc_fn backed by torch's caching allocator; keeps buffers alive until released."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.keep = []
        self.cb = ALLOC_FN(self._alloc)

    def _alloc(self, _ctx, tag, nbytes):
        try:
            t = torch.empty(max(int(nbytes), 1) + 256, dtype=torch.uint8, device=self.device)
            off = (-t.data_ptr()) % 256
            self.keep.append(t)
            self.bufs[int(tag)] = (t, off, int(nbytes))
            return t.data_ptr() + off
        except Exception:  # out of memory -> NULL -> LFS_ERR_ALLOC
            return None

    def tensor(self, tag: int, dtype: torch.dtype, numel: int) -> torch.Tensor:
        t, off, nbytes = self.bufs[tag]
        return t[off:off + nbytes].view(dtype)[:numel]


@_on_tensor_device
def projection_ut_3dgs_fused(means, quats, scales, opacities, viewmats0, viewmats1, Ks, image_width, image_height,
                             eps2d, near_plane, far_plane, radius_clip, calc_compensations, camera_model=PINHOLE,
                             ut_params: Optional[UTParams] = None, rs_type=SHUTTER_GLOBAL, radial_coeffs=None,
                             tangential_coeffs=None, thin_prism_coeffs=None):
    """gsplat::projection_ut_3dgs_fused (gsplat/Ops.h:69-98) -> (radii, means2d, depths, conics, compensations)."""
    lib = load()
    _chk(means, "means"), _chk(quats, "quats"), _chk(scales, "scales"), _chk(viewmats0, "viewmats0"), _chk(Ks, "Ks")
    if opacities is not None:
        _chk(opacities, "opacities")
    N, Cc = means.shape[0], Ks.shape[0]
    dev = means.device
    radii = torch.empty((Cc, N, 2), dtype=torch.int32, device=dev)
    means2d = torch.empty((Cc, N, 2), dtype=torch.float32, device=dev)
    depths = torch.empty((Cc, N), dtype=torch.float32, device=dev)
    conics = torch.empty((Cc, N, 3), dtype=torch.float32, device=dev)
    comp = torch.zeros((Cc, N), dtype=torch.float32, device=dev) if calc_compensations else None
    ut = ut_params or UTParams.default()
    check(lib.lfs_projection_ut_3dgs_fused(
        _p(means), _p(quats), _p(scales), _p(opacities), _p(viewmats0), _p(viewmats1), _p(Ks), N, Cc, image_width,
        image_height, eps2d, near_plane, far_plane, radius_clip, camera_model, C.byref(ut), rs_type,
        _p(radial_coeffs), _p(tangential_coeffs), _p(thin_prism_coeffs), _p(radii), _p(means2d), _p(depths),
        _p(conics), _p(comp), _stream()))
    return radii, means2d, depths, conics, comp


@_on_tensor_device
def spherical_harmonics_fwd(degrees_to_use: int, dirs, coeffs, masks=None):
    """gsplat::spherical_harmonics_fwd (gsplat/Ops.h:12-17). dirs [...,3], coeffs [...,K,3] -> colors [...,3]."""
    lib = load()
    _chk(dirs, "dirs"), _chk(coeffs, "coeffs")
    if coeffs.shape[-1] != 3 or dirs.shape[-1] != 3:
        raise ValueError("coeffs / dirs must have last dimension 3")
    m8 = None
    if masks is not None:
        _chk(masks, "masks", torch.bool)
        m8 = masks.view(torch.uint8)
    n, K = dirs.numel() // 3, coeffs.shape[-2]
    colors = torch.empty_like(dirs)
    check(lib.lfs_spherical_harmonics_fwd(degrees_to_use, _p(dirs), _p(coeffs), _p(m8), n, K, _p(colors), _stream()))
    return colors


@_on_tensor_device
def spherical_harmonics_bwd(K: int, degrees_to_use: int, dirs, coeffs, masks, v_colors, compute_v_dirs: bool):
    """gsplat::spherical_harmonics_bwd (gsplat/Ops.h:18-25) -> (v_coeffs, v_dirs or None)."""
    lib = load()
    _chk(dirs, "dirs"), _chk(coeffs, "coeffs"), _chk(v_colors, "v_colors")
    m8 = None
    if masks is not None:
        _chk(masks, "masks", torch.bool)
        m8 = masks.view(torch.uint8)
    n = dirs.numel() // 3
    v_coeffs = torch.empty_like(coeffs)
    v_dirs = torch.empty_like(dirs) if compute_v_dirs else None
    check(lib.lfs_spherical_harmonics_bwd(K, degrees_to_use, _p(dirs), _p(coeffs), _p(m8), _p(v_colors), n,
                                          _p(v_coeffs), _p(v_dirs), _stream()))
    return v_coeffs, v_dirs


@_on_tensor_device
def intersect_tile(means2d, radii, depths, camera_ids, gaussian_ids, C_: int, tile_size: int, tile_width: int,
                   tile_height: int, sort: bool = True):
    """gsplat::intersect_tile (gsplat/Ops.h:28-38) -> (tiles_per_gauss, isect_ids, flatten_ids)."""
    lib = load()
    _chk(means2d, "means2d"), _chk(radii, "radii", torch.int32), _chk(depths, "depths")
    packed = means2d.dim() == 2  # [nnz, 2]: every element carries its camera id (gsplat/Intersect.cpp:32-39)
    if packed:
        if camera_ids is None or gaussian_ids is None:
            raise ValueError("When packed is set, camera_ids and gaussian_ids must be provided.")
        _chk(camera_ids, "camera_ids", torch.int64), _chk(gaussian_ids, "gaussian_ids", torch.int64)
    dev = means2d.device
    tiles_per_gauss = torch.empty(depths.shape, dtype=torch.int32, device=dev)
    al = _Alloc(dev)
    ids_p, flat_p, n_is = C.c_void_p(), C.c_void_p(), C.c_int64(0)
    if packed:
        check(lib.lfs_intersect_tile_packed(_p(means2d), _p(radii), _p(depths), _p(camera_ids), means2d.shape[0], C_,
                                            tile_size, tile_width, tile_height, 1 if sort else 0, _p(tiles_per_gauss),
                                            al.cb, None, C.byref(ids_p), C.byref(flat_p), C.byref(n_is), _stream()))
    else:
        check(lib.lfs_intersect_tile(_p(means2d), _p(radii), _p(depths), C_, means2d.shape[1], tile_size, tile_width,
                                     tile_height, 1 if sort else 0, _p(tiles_per_gauss), al.cb, None, C.byref(ids_p),
                                     C.byref(flat_p), C.byref(n_is), _stream()))
    n = int(n_is.value)
    if n == 0:
        return (tiles_per_gauss, torch.empty((0,), dtype=torch.int64, device=dev),
                torch.empty((0,), dtype=torch.int32, device=dev))
    isect_ids = al.tensor(_lib_tag("ISECT"), torch.int64, n)
    flatten_ids = al.tensor(_lib_tag("FLAT"), torch.int32, n)
    # scratch (tag 0) is dropped when `al` dies: the caching allocator frees it in stream order, no host sync needed
    return tiles_per_gauss, isect_ids, flatten_ids


def _lib_tag(name: str) -> int:
    return {"SCRATCH": 0, "ISECT": 1, "FLAT": 2}[name]


@_on_tensor_device
def intersect_offset(isect_ids, C_: int, tile_width: int, tile_height: int):
    """gsplat::intersect_offset (gsplat/Ops.h:39-43) -> offsets [C, tile_height, tile_width] int32."""
    lib = load()
    _chk(isect_ids, "isect_ids", torch.int64)
    offsets = torch.empty((C_, tile_height, tile_width), dtype=torch.int32, device=isect_ids.device)
    check(lib.lfs_intersect_offset(_p(isect_ids), isect_ids.numel(), C_, tile_width, tile_height, _p(offsets),
                                   _stream()))
    return offsets


@_on_tensor_device
def rasterize_to_pixels_from_world_3dgs_fwd(means, quats, scales, colors, opacities, backgrounds, masks,
                                            image_width, image_height, tile_size, viewmats0, viewmats1, Ks,
                                            camera_model=PINHOLE, ut_params: Optional[UTParams] = None,
                                            rs_type=SHUTTER_GLOBAL, radial_coeffs=None, tangential_coeffs=None,
                                            thin_prism_coeffs=None, tile_offsets=None, flatten_ids=None):
    """gsplat::rasterize_to_pixels_from_world_3dgs_fwd (gsplat/Ops.h:100-129) -> (renders, alphas, last_ids)."""
    lib = load()
    for t, nme in ((means, "means"), (quats, "quats"), (scales, "scales"), (colors, "colors"),
                   (opacities, "opacities"), (viewmats0, "viewmats0"), (Ks, "Ks")):
        _chk(t, nme)
    _chk(tile_offsets, "tile_offsets", torch.int32), _chk(flatten_ids, "flatten_ids", torch.int32)
    if backgrounds is not None:
        _chk(backgrounds, "backgrounds")
    m8 = None
    if masks is not None:
        _chk(masks, "masks", torch.bool)
        m8 = masks.view(torch.uint8)
    Cc, N, ch = tile_offsets.shape[0], means.shape[0], colors.shape[-1]
    dev = means.device
    renders = torch.empty((Cc, image_height, image_width, ch), dtype=torch.float32, device=dev)
    alphas = torch.empty((Cc, image_height, image_width, 1), dtype=torch.float32, device=dev)
    last_ids = torch.empty((Cc, image_height, image_width), dtype=torch.int32, device=dev)
    ut = ut_params or UTParams.default()
    al = _Alloc(dev)
    check(lib.lfs_rasterize_to_pixels_from_world_3dgs_fwd(
        _p(means), _p(quats), _p(scales), _p(colors), _p(opacities), _p(backgrounds), _p(m8), N, Cc, ch, image_width,
        image_height, tile_size, _p(viewmats0), _p(viewmats1), _p(Ks), camera_model, C.byref(ut), rs_type,
        _p(radial_coeffs), _p(tangential_coeffs), _p(thin_prism_coeffs), _p(tile_offsets), _p(flatten_ids),
        flatten_ids.numel(), al.cb, None, _p(renders), _p(alphas), _p(last_ids), _stream()))
    return renders, alphas, last_ids


@_on_tensor_device
def rasterize_to_pixels_from_world_3dgs_bwd(means, quats, scales, colors, opacities, backgrounds, masks,
                                            image_width, image_height, tile_size, viewmats0, viewmats1, Ks,
                                            camera_model=PINHOLE, ut_params: Optional[UTParams] = None,
                                            rs_type=SHUTTER_GLOBAL, radial_coeffs=None, tangential_coeffs=None,
                                            thin_prism_coeffs=None, tile_offsets=None, flatten_ids=None,
                                            render_alphas=None, last_ids=None, v_render_colors=None,
                                            v_render_alphas=None):
    """gsplat::rasterize_to_pixels_from_world_3dgs_bwd (gsplat/Ops.h:131-166)
    -> (v_means, v_quats, v_scales, v_colors, v_opacities)."""
    lib = load()
    for t, nme in ((means, "means"), (quats, "quats"), (scales, "scales"), (colors, "colors"),
                   (opacities, "opacities"), (viewmats0, "viewmats0"), (Ks, "Ks"), (render_alphas, "render_alphas"),
                   (v_render_colors, "v_render_colors"), (v_render_alphas, "v_render_alphas")):
        _chk(t, nme)
    _chk(tile_offsets, "tile_offsets", torch.int32), _chk(flatten_ids, "flatten_ids", torch.int32)
    _chk(last_ids, "last_ids", torch.int32)
    if backgrounds is not None:
        _chk(backgrounds, "backgrounds")
    m8 = None
    if masks is not None:
        _chk(masks, "masks", torch.bool)
        m8 = masks.view(torch.uint8)
    Cc, N = tile_offsets.shape[0], means.shape[0]
    dev = means.device
    v_means = torch.empty_like(means)
    v_quats = torch.empty_like(quats)
    v_scales = torch.empty_like(scales)
    v_colors = torch.empty_like(colors)
    v_opacities = torch.empty_like(opacities)
    ut = ut_params or UTParams.default()
    al = _Alloc(dev)
    check(lib.lfs_rasterize_to_pixels_from_world_3dgs_bwd(
        _p(means), _p(quats), _p(scales), _p(colors), _p(opacities), _p(backgrounds), _p(m8), N, Cc, colors.shape[-1],
        image_width, image_height, tile_size, _p(viewmats0), _p(viewmats1), _p(Ks), camera_model, C.byref(ut), rs_type,
        _p(radial_coeffs), _p(tangential_coeffs), _p(thin_prism_coeffs), _p(tile_offsets), _p(flatten_ids),
        flatten_ids.numel(), _p(render_alphas), _p(last_ids), _p(v_render_colors), _p(v_render_alphas), al.cb, None,
        _p(v_means), _p(v_quats), _p(v_scales), _p(v_colors), _p(v_opacities), _stream()))
    return v_means, v_quats, v_scales, v_colors, v_opacities


# ---- legacy 2-D op surface the reference's gtest files call ------------------------------
@_on_tensor_device
def quat_scale_to_covar_preci_fwd(quats, scales, compute_covar: bool = True, compute_preci: bool = True,
                                  triu: bool = False):
    """gsplat::quat_scale_to_covar_preci_fwd as tests/test_basic.cpp:54-67 calls it -> (covars, precis); an output that was
    not requested is an empty tensor."""
    lib = load()
    _chk(quats, "quats"), _chk(scales, "scales")
    N, dev = quats.shape[0], quats.device
    shape = (N, 6) if triu else (N, 3, 3)
    mk = lambda want: torch.empty(shape if want else (0,), dtype=torch.float32, device=dev)  # noqa: E731
    covars, precis = mk(compute_covar), mk(compute_preci)
    if compute_covar or compute_preci:
        check(lib.lfs_quat_scale_to_covar_preci_fwd(_p(quats), _p(scales), N, int(triu),
                                                    _p(covars) if compute_covar else None,
                                                    _p(precis) if compute_preci else None, _stream()))
    return covars, precis


@_on_tensor_device
def quat_scale_to_covar_preci_bwd(quats, scales, triu: bool, v_covars=None, v_precis=None):
    """gsplat::quat_scale_to_covar_preci_bwd (tests/test_basic.cpp:80-81) -> (v_quats, v_scales)."""
    lib = load()
    _chk(quats, "quats"), _chk(scales, "scales")
    vc = v_covars if (v_covars is not None and v_covars.numel()) else None
    vp = v_precis if (v_precis is not None and v_precis.numel()) else None
    for t, nme in ((vc, "v_covars"), (vp, "v_precis")):
        if t is not None:
            _chk(t, nme)
    v_quats, v_scales = torch.zeros_like(quats), torch.zeros_like(scales)
    if vc is not None or vp is not None:
        check(lib.lfs_quat_scale_to_covar_preci_bwd(_p(quats), _p(scales), quats.shape[0], int(triu), _p(vc), _p(vp),
                                                    _p(v_quats), _p(v_scales), _stream()))
    return v_quats, v_scales


@_on_tensor_device
def projection_ewa_3dgs_fused_fwd(means, covars, quats, scales, opacities, viewmats, Ks, image_width: int,
                                  image_height: int, eps2d: float, near_plane: float, far_plane: float,
                                  radius_clip: float, calc_compensations: bool, camera_model=PINHOLE):
    """gsplat::projection_ewa_3dgs_fused_fwd (tests/test_basic.cpp:114-128)
    -> (radii [C,N,2], means2d, depths, conics, compensations or empty).  An empty `covars` selects quats + scales."""
    lib = load()
    _chk(means, "means"), _chk(viewmats, "viewmats"), _chk(Ks, "Ks")
    cov = covars if (covars is not None and covars.numel()) else None
    if cov is not None:
        _chk(cov, "covars")
    else:
        _chk(quats, "quats"), _chk(scales, "scales")
    op = opacities if (opacities is not None and opacities.numel()) else None
    if op is not None:
        _chk(op, "opacities")
    N, Cc, dev = means.shape[0], viewmats.shape[0], means.device
    radii = torch.empty((Cc, N, 2), dtype=torch.int32, device=dev)
    means2d = torch.zeros((Cc, N, 2), dtype=torch.float32, device=dev)
    depths = torch.zeros((Cc, N), dtype=torch.float32, device=dev)
    conics = torch.zeros((Cc, N, 3), dtype=torch.float32, device=dev)
    comp = torch.zeros((Cc, N), dtype=torch.float32, device=dev) if calc_compensations else torch.empty(0, device=dev)
    check(lib.lfs_projection_ewa_3dgs_fused_fwd(
        _p(means), _p(cov), _p(quats) if cov is None else None, _p(scales) if cov is None else None, _p(op), _p(viewmats),
        _p(Ks), N, Cc, image_width, image_height, eps2d, near_plane, far_plane, radius_clip, camera_model, _p(radii),
        _p(means2d), _p(depths), _p(conics), _p(comp) if calc_compensations else None, _stream()))
    return radii, means2d, depths, conics, comp


def _masks_u8(masks):
    if masks is None or masks.numel() == 0:
        return None
    _chk(masks, "masks", torch.bool)
    return masks.view(torch.uint8)


@_on_tensor_device
def rasterize_to_pixels_3dgs_fwd(means2d, conics, colors, opacities, backgrounds, masks, image_width: int,
                                 image_height: int, tile_size: int, tile_offsets, flatten_ids):
    """gsplat::rasterize_to_pixels_3dgs_fwd (tests/test_basic.cpp:347-358) -> (renders, alphas, last_ids)."""
    lib = load()
    for t, nme in ((means2d, "means2d"), (conics, "conics"), (colors, "colors"), (opacities, "opacities")):
        _chk(t, nme)
    _chk(tile_offsets, "tile_offsets", torch.int32), _chk(flatten_ids, "flatten_ids", torch.int32)
    bg = backgrounds if (backgrounds is not None and backgrounds.numel()) else None
    if bg is not None:
        _chk(bg, "backgrounds")
    Cc, N, ch, dev = tile_offsets.shape[0], means2d.shape[-2], colors.shape[-1], means2d.device
    renders = torch.empty((Cc, image_height, image_width, ch), dtype=torch.float32, device=dev)
    alphas = torch.empty((Cc, image_height, image_width, 1), dtype=torch.float32, device=dev)
    last_ids = torch.empty((Cc, image_height, image_width), dtype=torch.int32, device=dev)
    check(lib.lfs_rasterize_to_pixels_3dgs_fwd(
        _p(means2d), _p(conics), _p(colors), _p(opacities), _p(bg), _p(_masks_u8(masks)), Cc, N, ch, image_width,
        image_height, tile_size, _p(tile_offsets), _p(flatten_ids), flatten_ids.numel(), _p(renders), _p(alphas),
        _p(last_ids), _stream()))
    return renders, alphas, last_ids


@_on_tensor_device
def rasterize_to_pixels_3dgs_bwd(means2d, conics, colors, opacities, backgrounds, masks, image_width: int,
                                 image_height: int, tile_size: int, tile_offsets, flatten_ids, render_alphas, last_ids,
                                 v_render_colors, v_render_alphas, absgrad: bool = False):
    """gsplat::rasterize_to_pixels_3dgs_bwd (launcher gsplat/Rasterization.h:38-63)
    -> (v_means2d_abs or empty, v_means2d, v_conics, v_colors, v_opacities)."""
    lib = load()
    for t, nme in ((means2d, "means2d"), (conics, "conics"), (colors, "colors"), (opacities, "opacities"),
                   (render_alphas, "render_alphas"), (v_render_colors, "v_render_colors"),
                   (v_render_alphas, "v_render_alphas")):
        _chk(t, nme)
    _chk(tile_offsets, "tile_offsets", torch.int32), _chk(flatten_ids, "flatten_ids", torch.int32)
    _chk(last_ids, "last_ids", torch.int32)
    bg = backgrounds if (backgrounds is not None and backgrounds.numel()) else None
    if bg is not None:
        _chk(bg, "backgrounds")
    Cc, N, ch, dev = tile_offsets.shape[0], means2d.shape[-2], colors.shape[-1], means2d.device
    v_abs = torch.empty_like(means2d) if absgrad else torch.empty(0, device=dev)
    v_means2d, v_conics = torch.empty_like(means2d), torch.empty_like(conics)
    v_colors, v_opacities = torch.empty_like(colors), torch.empty_like(opacities)
    check(lib.lfs_rasterize_to_pixels_3dgs_bwd(
        _p(means2d), _p(conics), _p(colors), _p(opacities), _p(bg), _p(_masks_u8(masks)), Cc, N, ch, image_width,
        image_height, tile_size, _p(tile_offsets), _p(flatten_ids), flatten_ids.numel(), _p(render_alphas), _p(last_ids),
        _p(v_render_colors), _p(v_render_alphas), _p(v_abs) if absgrad else None, _p(v_means2d), _p(v_conics),
        _p(v_colors), _p(v_opacities), _stream()))
    return v_abs, v_means2d, v_conics, v_colors, v_opacities


@_on_tensor_device
def adam_step(param, exp_avg, exp_avg_sq, param_grad, lr, beta1, beta2, eps, bias_correction1_rcp,
              bias_correction2_sqrt_rcp):
    """fast_gs::optimizer::adam_step_wrapper (fastgs/optimizer/include/adam_api.h:11-21); in place."""
    lib = load()
    for t, nme in ((param, "param"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq"), (param_grad, "param_grad")):
        _chk(t, nme)
    check(lib.lfs_adam_step(_p(param), _p(exp_avg), _p(exp_avg_sq), _p(param_grad), param.numel(), lr, beta1, beta2,
                            eps, bias_correction1_rcp, bias_correction2_sqrt_rcp, _stream()))


@_on_tensor_device
def quats_to_rotmats(quats):
    """gsplat::quats_to_rotmats (gsplat/Ops.h:46-48): [N,4] -> [N,3,3]."""
    lib = load()
    _chk(quats, "quats")
    out = torch.empty((quats.shape[0], 3, 3), dtype=torch.float32, device=quats.device)
    check(lib.lfs_quats_to_rotmats(_p(quats), quats.shape[0], _p(out), _stream()))
    return out


@_on_tensor_device
def relocation(opacities, scales, ratios, binoms, n_max: int):
    """gsplat::relocation (gsplat/Ops.h:52-57) -> (new_opacities [N], new_scales [N,3])."""
    lib = load()
    _chk(opacities, "opacities"), _chk(scales, "scales"), _chk(ratios, "ratios", torch.int32), _chk(binoms, "binoms")
    new_op, new_sc = torch.empty_like(opacities), torch.empty_like(scales)
    check(lib.lfs_relocation(_p(opacities), _p(scales), _p(ratios), _p(binoms), n_max, opacities.shape[0], _p(new_op),
                             _p(new_sc), _stream()))
    return new_op, new_sc


@_on_tensor_device
def add_noise(raw_opacities, raw_scales, raw_quats, noise, means, current_lr: float) -> None:
    """gsplat::add_noise (gsplat/Ops.h:59-65); means updated in place."""
    lib = load()
    for t, nme in ((raw_opacities, "raw_opacities"), (raw_scales, "raw_scales"), (raw_quats, "raw_quats"),
                   (noise, "noise"), (means, "means")):
        _chk(t, nme)
    check(lib.lfs_add_noise(_p(raw_opacities), _p(raw_scales), _p(raw_quats), _p(noise), _p(means), current_lr,
                            raw_opacities.shape[0], _stream()))


class FastGSContext:
    """State the reference keeps between forward and backward: the four opaque byte buffers, the three counters and
    the selector (fast_rasterizer_autograd.cpp:61-72)."""

    def __init__(self, alloc, n_visible, n_instances, n_buckets, selector, dims):
        self.alloc = alloc  # keeps the buffers alive
        self.n_visible_primitives, self.n_instances, self.n_buckets = n_visible, n_instances, n_buckets
        self.instance_selector = selector
        self.dims = dims

    def buffer(self, tag: int) -> int:
        t, off, _ = self.alloc.bufs[tag]
        return t.data_ptr() + off


@_on_tensor_device
def fastgs_forward(means, scales_raw, rotations_raw, opacities_raw, sh_coefficients_0, sh_coefficients_rest, w2c,
                   cam_position, active_sh_bases: int, width: int, height: int, focal_x: float, focal_y: float,
                   center_x: float, center_y: float, near_plane: float, far_plane: float):
    """fast_gs::rasterization::forward_wrapper (fastgs/rasterization/include/rasterization_api.h:26-44)
    -> (image [3,H,W], alpha [1,H,W], FastGSContext)."""
    lib = load()
    for t, n in ((means, "means"), (scales_raw, "scales_raw"), (rotations_raw, "rotations_raw"),
                 (opacities_raw, "opacities_raw"), (sh_coefficients_0, "sh_coefficients_0"),
                 (sh_coefficients_rest, "sh_coefficients_rest")):
        _chk(t, n)
    w2c, cam_position = w2c.contiguous(), cam_position.contiguous()
    _chk(w2c, "w2c"), _chk(cam_position, "cam_position")
    dev = means.device
    N, total_rest = means.shape[0], sh_coefficients_rest.shape[1]
    image = torch.empty((3, height, width), dtype=torch.float32, device=dev)
    alpha = torch.empty((1, height, width), dtype=torch.float32, device=dev)
    al = _Alloc(dev)
    cnt = [C.c_int(0) for _ in range(4)]
    check(lib.lfs_fastgs_forward(_p(means), _p(scales_raw), _p(rotations_raw), _p(opacities_raw), _p(sh_coefficients_0),
                                 _p(sh_coefficients_rest), _p(w2c), _p(cam_position), N, active_sh_bases, total_rest,
                                 width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane, _p(image),
                                 _p(alpha), al.cb, None, C.byref(cnt[0]), C.byref(cnt[1]), C.byref(cnt[2]),
                                 C.byref(cnt[3]), _stream()))
    ctx = FastGSContext(al, cnt[0].value, cnt[1].value, cnt[2].value, cnt[3].value,
                        (N, total_rest, active_sh_bases, width, height, focal_x, focal_y, center_x, center_y))
    return image, alpha, ctx


@_on_tensor_device
def fastgs_backward(ctx: FastGSContext, grad_image, grad_alpha, means, scales_raw, rotations_raw, sh_coefficients_rest,
                    w2c, cam_position, densification_info=None, want_w2c_grad: bool = False):
    """fast_gs::rasterization::backward_wrapper (rasterization_api.h:46-75) -> (grad_means, grad_scales_raw,
    grad_rotations_raw, grad_opacities_raw, grad_sh_coefficients_0, grad_sh_coefficients_rest, grad_w2c | None)."""
    lib = load()
    N, total_rest, active, width, height, fx, fy, cx, cy = ctx.dims
    _chk(grad_image, "grad_image"), _chk(grad_alpha, "grad_alpha")
    dev = means.device
    g_means = torch.empty((N, 3), dtype=torch.float32, device=dev)
    g_scales = torch.empty((N, 3), dtype=torch.float32, device=dev)
    g_rot = torch.empty((N, 4), dtype=torch.float32, device=dev)
    g_op = torch.empty((N, 1), dtype=torch.float32, device=dev)
    g_sh0 = torch.empty((N, 1, 3), dtype=torch.float32, device=dev)
    g_shN = torch.empty((N, total_rest, 3), dtype=torch.float32, device=dev)
    g_w2c = torch.zeros((4, 4), dtype=torch.float32, device=dev) if want_w2c_grad else None
    if densification_info is not None and densification_info.numel() == 0:
        densification_info = None
    al = _Alloc(dev)
    check(lib.lfs_fastgs_backward(_p(grad_image), _p(grad_alpha), _p(means), _p(scales_raw), _p(rotations_raw),
                                  _p(sh_coefficients_rest), _p(w2c.contiguous()), _p(cam_position.contiguous()),
                                  ctx.buffer(_lib.LFS_TAG_FG_PER_PRIMITIVE), ctx.buffer(_lib.LFS_TAG_FG_PER_TILE),
                                  ctx.buffer(_lib.LFS_TAG_FG_PER_INSTANCE), ctx.buffer(_lib.LFS_TAG_FG_PER_BUCKET),
                                  _p(g_means), _p(g_scales), _p(g_rot), _p(g_op), _p(g_sh0), _p(g_shN), _p(g_w2c),
                                  _p(densification_info), N, ctx.n_instances, ctx.n_buckets, ctx.instance_selector,
                                  active, total_rest, width, height, fx, fy, cx, cy, al.cb, None, _stream()))
    return g_means, g_scales, g_rot, g_op, g_sh0, g_shN, g_w2c
