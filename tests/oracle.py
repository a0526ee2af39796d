"""ctypes front-end of the CPU oracle (oracle/liblfs_oracle.so).  TEST INFRASTRUCTURE ONLY.

numpy in, numpy out.  `prec` selects the float (orc32_) or double (orc64_) variant of the floating-point
stages; integer stages (tile intersection) have a single bit-exact implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "liblfs_oracle.so")
_lib = None


def build():
    r = subprocess.run(["make", "-C", _ORACLE_DIR, "oracle"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_intersect_count.restype = C.c_int64
    return _lib


def _dt(prec):
    return (np.float64, "orc64_", C.c_double) if prec == 64 else (np.float32, "orc32_", C.c_float)


def _a(x, dtype):
    return None if x is None else np.ascontiguousarray(x, dtype=dtype)


def _p(x):
    return None if x is None else x.ctypes.data_as(C.c_void_p)


def projection_ut(means, quats, scales, opacities, viewmats, Ks, width, height, eps2d=0.3, near=0.01, far=1e4,
                  radius_clip=0.0, ut=(0.1, 2.0, 0.0, 0.1, 1), calc_compensations=False, prec=64):
    dt, pfx, cf = _dt(prec)
    means, quats, scales = _a(means, dt), _a(quats, dt), _a(scales, dt)
    opacities, viewmats, Ks = _a(opacities, dt), _a(viewmats, dt), _a(Ks, dt)
    N, Cc = means.shape[0], Ks.shape[0]
    radii = np.zeros((Cc, N, 2), np.int32)
    means2d = np.zeros((Cc, N, 2), dt)
    depths = np.zeros((Cc, N), dt)
    conics = np.zeros((Cc, N, 3), dt)
    comp = np.zeros((Cc, N), dt) if calc_compensations else None
    getattr(lib(), pfx + "projection_ut")(
        C.c_int(Cc), C.c_int(N), _p(means), _p(quats), _p(scales), _p(opacities), _p(viewmats), _p(Ks),
        C.c_int(width), C.c_int(height), cf(eps2d), cf(near), cf(far), cf(radius_clip), cf(ut[0]), cf(ut[1]),
        cf(ut[2]), cf(ut[3]), C.c_int(ut[4]), _p(radii), _p(means2d), _p(depths), _p(conics), _p(comp))
    return radii, means2d, depths, conics, comp


def sh_fwd(degree, dirs, coeffs, masks=None, prec=64):
    dt, pfx, _ = _dt(prec)
    dirs, coeffs = _a(dirs, dt), _a(coeffs, dt)
    n, K = dirs.reshape(-1, 3).shape[0], coeffs.shape[-2]
    m = None if masks is None else _a(masks, np.uint8)
    colors = np.zeros((n, 3), dt)
    getattr(lib(), pfx + "sh_fwd")(C.c_int(degree), C.c_int(n), C.c_int(K), _p(dirs), _p(coeffs), _p(m), _p(colors))
    return colors.reshape(dirs.shape)


def sh_bwd(degree, dirs, coeffs, v_colors, masks=None, compute_v_dirs=True, prec=64):
    dt, pfx, _ = _dt(prec)
    dirs, coeffs, v_colors = _a(dirs, dt), _a(coeffs, dt), _a(v_colors, dt)
    n, K = dirs.reshape(-1, 3).shape[0], coeffs.shape[-2]
    m = None if masks is None else _a(masks, np.uint8)
    v_coeffs = np.zeros(coeffs.shape, dt)
    v_dirs = np.zeros(dirs.shape, dt) if compute_v_dirs else None
    getattr(lib(), pfx + "sh_bwd")(C.c_int(degree), C.c_int(n), C.c_int(K), _p(dirs), _p(coeffs), _p(m),
                                   _p(v_colors), _p(v_coeffs), _p(v_dirs))
    return v_coeffs, v_dirs


def intersect_tile(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True):
    """-> (tiles_per_gauss [C,N] i32, isect_ids [n] i64, flatten_ids [n] i32); float32 arithmetic, bit exact."""
    means2d, radii, depths = _a(means2d, np.float32), _a(radii, np.int32), _a(depths, np.float32)
    Cc, N = depths.shape
    tpg = np.zeros((Cc, N), np.int32)
    n = lib().orc_intersect_count(C.c_int(Cc), C.c_int(N), _p(means2d), _p(radii), C.c_uint32(tile_size),
                                  C.c_uint32(tile_width), C.c_uint32(tile_height), _p(tpg))
    ids = np.zeros((n,), np.int64)
    flat = np.zeros((n,), np.int32)
    lib().orc_intersect_emit(C.c_int(Cc), C.c_int(N), _p(means2d), _p(radii), _p(depths), C.c_uint32(tile_size),
                             C.c_uint32(tile_width), C.c_uint32(tile_height), C.c_int(1 if sort else 0), _p(ids),
                             _p(flat))
    return tpg, ids, flat


def intersect_offset(isect_ids, Cc, tile_width, tile_height):
    isect_ids = _a(isect_ids, np.int64)
    off = np.zeros((Cc, tile_height, tile_width), np.int32)
    lib().orc_intersect_offset(C.c_int64(isect_ids.shape[0]), _p(isect_ids), C.c_int(Cc), C.c_uint32(tile_width),
                               C.c_uint32(tile_height), _p(off))
    return off


def raster_world_fwd(means, quats, scales, colors, opacities, backgrounds, tile_masks, width, height, tile_size,
                     viewmats, Ks, tile_offsets, flatten_ids, prec=64):
    dt, pfx, _ = _dt(prec)
    means, quats, scales = _a(means, dt), _a(quats, dt), _a(scales, dt)
    colors, opacities, viewmats, Ks = _a(colors, dt), _a(opacities, dt), _a(viewmats, dt), _a(Ks, dt)
    backgrounds = _a(backgrounds, dt)
    tile_masks = None if tile_masks is None else _a(tile_masks, np.uint8)
    tile_offsets, flatten_ids = _a(tile_offsets, np.int32), _a(flatten_ids, np.int32)
    Cc, N, CH = viewmats.shape[0], means.shape[0], colors.shape[-1]
    renders = np.zeros((Cc, height, width, CH), dt)
    alphas = np.zeros((Cc, height, width, 1), dt)
    last_ids = np.zeros((Cc, height, width), np.int32)
    getattr(lib(), pfx + "raster_world_fwd")(
        C.c_int(Cc), C.c_int(N), C.c_int(CH), _p(means), _p(quats), _p(scales), _p(colors), _p(opacities),
        _p(backgrounds), _p(tile_masks), C.c_int(width), C.c_int(height), C.c_int(tile_size), _p(viewmats), _p(Ks),
        _p(tile_offsets), _p(flatten_ids), C.c_int64(flatten_ids.shape[0]), _p(renders), _p(alphas), _p(last_ids))
    return renders, alphas, last_ids


def raster_world_bwd(means, quats, scales, colors, opacities, backgrounds, tile_masks, width, height, tile_size,
                     viewmats, Ks, tile_offsets, flatten_ids, render_alphas, last_ids, v_render_colors,
                     v_render_alphas, prec=64):
    dt, pfx, _ = _dt(prec)
    means, quats, scales = _a(means, dt), _a(quats, dt), _a(scales, dt)
    colors, opacities, viewmats, Ks = _a(colors, dt), _a(opacities, dt), _a(viewmats, dt), _a(Ks, dt)
    backgrounds = _a(backgrounds, dt)
    tile_masks = None if tile_masks is None else _a(tile_masks, np.uint8)
    tile_offsets, flatten_ids = _a(tile_offsets, np.int32), _a(flatten_ids, np.int32)
    render_alphas, last_ids = _a(render_alphas, dt), _a(last_ids, np.int32)
    v_render_colors, v_render_alphas = _a(v_render_colors, dt), _a(v_render_alphas, dt)
    Cc, N = viewmats.shape[0], means.shape[0]
    v_means = np.zeros((N, 3), np.float64)
    v_quats = np.zeros((N, 4), np.float64)
    v_scales = np.zeros((N, 3), np.float64)
    v_colors = np.zeros((Cc, N, 3), np.float64)
    v_opac = np.zeros((Cc, N), np.float64)
    getattr(lib(), pfx + "raster_world_bwd")(
        C.c_int(Cc), C.c_int(N), _p(means), _p(quats), _p(scales), _p(colors), _p(opacities), _p(backgrounds),
        _p(tile_masks), C.c_int(width), C.c_int(height), C.c_int(tile_size), _p(viewmats), _p(Ks), _p(tile_offsets),
        _p(flatten_ids), C.c_int64(flatten_ids.shape[0]), _p(render_alphas), _p(last_ids), _p(v_render_colors),
        _p(v_render_alphas), _p(v_means), _p(v_quats), _p(v_scales), _p(v_colors), _p(v_opac))
    return v_means, v_quats, v_scales, v_colors, v_opac


# ---- legacy 2-D op surface (oracle/lfs_oracle_legacy2d_impl.h) ---------------------------------
def quat_scale_to_covar_preci_fwd(quats, scales, compute_covar=True, compute_preci=True, triu=False, prec=64):
    dt, pfx, _ = _dt(prec)
    quats, scales = _a(quats, dt), _a(scales, dt)
    N = quats.shape[0]
    shape = (N, 6) if triu else (N, 3, 3)
    cov = np.zeros(shape, dt) if compute_covar else None
    pre = np.zeros(shape, dt) if compute_preci else None
    getattr(lib(), pfx + "quat_scale_to_covar_preci_fwd")(C.c_int(N), _p(quats), _p(scales), C.c_int(int(triu)), _p(cov),
                                                          _p(pre))
    return cov, pre


def quat_scale_to_covar_preci_bwd(quats, scales, triu, v_covars, v_precis, prec=64):
    dt, pfx, _ = _dt(prec)
    quats, scales, v_covars, v_precis = _a(quats, dt), _a(scales, dt), _a(v_covars, dt), _a(v_precis, dt)
    N = quats.shape[0]
    v_quats, v_scales = np.zeros((N, 4), dt), np.zeros((N, 3), dt)
    getattr(lib(), pfx + "quat_scale_to_covar_preci_bwd")(C.c_int(N), _p(quats), _p(scales), C.c_int(int(triu)),
                                                          _p(v_covars), _p(v_precis), _p(v_quats), _p(v_scales))
    return v_quats, v_scales


def projection_ewa(means, covars, quats, scales, opacities, viewmats, Ks, width, height, eps2d=0.3, near=0.01, far=1e4,
                   radius_clip=0.0, calc_compensations=False, prec=64):
    dt, pfx, cf = _dt(prec)
    means, covars, quats, scales = _a(means, dt), _a(covars, dt), _a(quats, dt), _a(scales, dt)
    opacities, viewmats, Ks = _a(opacities, dt), _a(viewmats, dt), _a(Ks, dt)
    N, Cc = means.shape[0], Ks.shape[0]
    radii = np.zeros((Cc, N, 2), np.int32)
    means2d, depths, conics = np.zeros((Cc, N, 2), dt), np.zeros((Cc, N), dt), np.zeros((Cc, N, 3), dt)
    comp = np.zeros((Cc, N), dt) if calc_compensations else None
    getattr(lib(), pfx + "projection_ewa")(
        C.c_int(Cc), C.c_int(N), _p(means), _p(covars), _p(quats), _p(scales), _p(opacities), _p(viewmats), _p(Ks),
        C.c_int(width), C.c_int(height), cf(eps2d), cf(near), cf(far), cf(radius_clip), _p(radii), _p(means2d), _p(depths),
        _p(conics), _p(comp))
    return radii, means2d, depths, conics, comp


def raster_2d_fwd(means2d, conics, colors, opacities, backgrounds, tile_masks, width, height, tile_size, tile_offsets,
                  flatten_ids, prec=64):
    dt, pfx, _ = _dt(prec)
    means2d, conics, colors, opacities = _a(means2d, dt), _a(conics, dt), _a(colors, dt), _a(opacities, dt)
    backgrounds = _a(backgrounds, dt)
    tile_masks = None if tile_masks is None else _a(tile_masks, np.uint8)
    tile_offsets, flatten_ids = _a(tile_offsets, np.int32), _a(flatten_ids, np.int32)
    Cc, N, CH = means2d.shape[0], means2d.shape[1], colors.shape[-1]
    renders = np.zeros((Cc, height, width, CH), dt)
    alphas = np.zeros((Cc, height, width, 1), dt)
    last_ids = np.zeros((Cc, height, width), np.int32)
    getattr(lib(), pfx + "raster_2d_fwd")(
        C.c_int(Cc), C.c_int(N), C.c_int(CH), _p(means2d), _p(conics), _p(colors), _p(opacities), _p(backgrounds),
        _p(tile_masks), C.c_int(width), C.c_int(height), C.c_int(tile_size), _p(tile_offsets), _p(flatten_ids),
        C.c_int64(flatten_ids.shape[0]), _p(renders), _p(alphas), _p(last_ids))
    return renders, alphas, last_ids


def raster_2d_bwd(means2d, conics, colors, opacities, backgrounds, tile_masks, width, height, tile_size, tile_offsets,
                  flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas, absgrad=False, prec=64):
    dt, pfx, _ = _dt(prec)
    means2d, conics, colors, opacities = _a(means2d, dt), _a(conics, dt), _a(colors, dt), _a(opacities, dt)
    backgrounds = _a(backgrounds, dt)
    tile_masks = None if tile_masks is None else _a(tile_masks, np.uint8)
    tile_offsets, flatten_ids = _a(tile_offsets, np.int32), _a(flatten_ids, np.int32)
    render_alphas, last_ids = _a(render_alphas, dt), _a(last_ids, np.int32)
    v_render_colors, v_render_alphas = _a(v_render_colors, dt), _a(v_render_alphas, dt)
    Cc, N, CH = means2d.shape[0], means2d.shape[1], colors.shape[-1]
    v_m = np.zeros((Cc, N, 2), np.float64)
    v_abs = np.zeros((Cc, N, 2), np.float64) if absgrad else None
    v_con, v_col, v_op = np.zeros((Cc, N, 3), np.float64), np.zeros((Cc, N, CH), np.float64), np.zeros((Cc, N), np.float64)
    getattr(lib(), pfx + "raster_2d_bwd")(
        C.c_int(Cc), C.c_int(N), C.c_int(CH), _p(means2d), _p(conics), _p(colors), _p(opacities), _p(backgrounds),
        _p(tile_masks), C.c_int(width), C.c_int(height), C.c_int(tile_size), _p(tile_offsets), _p(flatten_ids),
        _p(render_alphas), _p(last_ids), _p(v_render_colors), _p(v_render_alphas), _p(v_m), _p(v_abs), _p(v_con), _p(v_col),
        _p(v_op))
    return v_m, v_abs, v_con, v_col, v_op


def adam_step(param, exp_avg, exp_avg_sq, grad, lr, beta1, beta2, eps, bc1_rcp, bc2_sqrt_rcp, prec=64):
    """Returns updated copies (param, exp_avg, exp_avg_sq)."""
    dt, pfx, cf = _dt(prec)
    p, m, v, g = (np.array(x, dtype=dt, copy=True).reshape(-1) for x in (param, exp_avg, exp_avg_sq, grad))
    getattr(lib(), pfx + "adam_step")(_p(p), _p(m), _p(v), _p(g), C.c_int64(p.shape[0]), cf(lr), cf(beta1),
                                      cf(beta2), cf(eps), cf(bc1_rcp), cf(bc2_sqrt_rcp))
    return p, m, v


# ------------------------------------------------------------------------------------------------------------
# composition of the oracle stages into the reference's per-view pipeline
# (src/training/rasterization/rasterizer.cpp:208-360): activated params -> image (+ everything in between)
# ------------------------------------------------------------------------------------------------------------
def render_view(means, quats, scales, opacities, shs, sh_degree, viewmat, K, width, height, bg=None, tile_size=16,
                prec=64):
    vm, Kk = np.asarray(viewmat)[None], np.asarray(K)[None]
    radii, means2d, depths, conics, _ = projection_ut(means, quats, scales, opacities, vm, Kk, width, height,
                                                      prec=prec)
    campos = np.linalg.inv(np.asarray(viewmat, np.float64))[:3, 3]
    dirs = np.asarray(means, np.float64) - campos[None]
    masks = (radii[0] > 0).all(-1)
    cols = sh_fwd(sh_degree, dirs, shs, masks.astype(np.uint8), prec=prec)
    colors = np.maximum(cols + 0.5, 0.0)
    tw, th = (width + tile_size - 1) // tile_size, (height + tile_size - 1) // tile_size
    tpg, ids, flat = intersect_tile(means2d, radii, depths, tile_size, tw, th, True)
    offs = intersect_offset(ids, 1, tw, th)
    bgs = None if bg is None else np.asarray(bg, np.float64)[None]
    renders, alphas, last_ids = raster_world_fwd(means, quats, scales, colors[None], np.asarray(opacities)[None], bgs,
                                                 None, width, height, tile_size, vm, Kk, offs, flat, prec=prec)
    return dict(radii=radii, means2d=means2d, depths=depths, conics=conics, dirs=dirs, masks=masks, sh_colors=cols,
                colors=colors, tiles_per_gauss=tpg, isect_ids=ids, flatten_ids=flat, offsets=offs, renders=renders,
                alphas=alphas, last_ids=last_ids)


def view_loss_grads(scene_raw, viewmat, K, width, height, sh_degree, bg, v_image=None, v_alpha=None, target=None,
                    l1_scale=None, prec=64, lambda_dssim=None):
    """Full per-view chain on raw parameters, as the reference composes it with torch autograd
    (rasterizer.cpp:72-81 activations, :250-266 SH + clamp, rasterizer_autograd.cpp:329-392 blend bwd,
    :84-130 SH bwd): returns (outputs dict, grads dict in the reference's AoS layout).
    Either upstream grads (v_image [H,W,3], v_alpha [H,W]) or an L1 target (uint8/float [H,W,3]) is given."""
    means = np.asarray(scene_raw["means"], np.float64)
    rot = np.asarray(scene_raw["rotation"], np.float64)
    nrm = np.maximum(np.linalg.norm(rot, axis=-1, keepdims=True), 1e-12)
    q = rot / nrm
    scales = np.exp(np.asarray(scene_raw["scaling"], np.float64))
    op_raw = np.asarray(scene_raw["opacity"], np.float64)[:, 0]
    op = 1.0 / (1.0 + np.exp(-op_raw))
    sh0, shN = np.asarray(scene_raw["sh0"], np.float64), np.asarray(scene_raw["shN"], np.float64)
    shs = np.concatenate([sh0, shN], axis=1)
    r = render_view(means, q, scales, op, shs, sh_degree, viewmat, K, width, height, bg=bg, prec=prec)
    image = r["renders"][0]
    loss = None
    if target is not None:
        tgt = np.asarray(target, np.float64) / (255.0 if np.asarray(target).dtype == np.uint8 else 1.0)
        if lambda_dssim is not None:  # the reference's L1 + SSIM loss (src/training/trainer.cpp:103-131)
            loss, v_image, _ = photometric_loss(image, tgt, lambda_dssim)
        else:
            s = 1.0 / (3.0 * width * height) if l1_scale is None else l1_scale
            diff = np.clip(image, 0.0, 1.0) - tgt  # render clamped to [0,1] (rasterizer.cpp:401)
            loss = s * np.abs(diff).sum()
            v_image = s * np.sign(diff) * ((image >= 0.0) & (image <= 1.0))
        v_alpha = np.zeros((height, width))
    v_alpha = np.zeros((height, width)) if v_alpha is None else v_alpha
    vm, Kk = np.asarray(viewmat)[None], np.asarray(K)[None]
    bgs = None if bg is None else np.asarray(bg, np.float64)[None]
    g = raster_world_bwd(means, q, scales, r["colors"][None], op[None], bgs, None, width, height, 16, vm, Kk,
                         r["offsets"], r["flatten_ids"], r["alphas"], r["last_ids"], np.asarray(v_image)[None],
                         np.asarray(v_alpha)[None, :, :, None], prec=prec)
    v_means, v_quats, v_scales, v_colors, v_opac = g[0], g[1], g[2], g[3][0], g[4][0]
    v_colors = v_colors * (r["sh_colors"] + 0.5 >= 0.0)  # clamp_min backward
    v_coeffs, v_dirs = sh_bwd(sh_degree, r["dirs"], shs, v_colors, r["masks"].astype(np.uint8), True, prec=prec)
    v_means = v_means + v_dirs
    dq = (v_quats * q).sum(-1, keepdims=True)
    grads = dict(means=v_means, sh0=v_coeffs[:, :1], shN=v_coeffs[:, 1:], scaling=v_scales * scales,
                 rotation=(v_quats - dq * q) / nrm, opacity=(v_opac * op * (1.0 - op))[:, None])
    r["loss"] = loss
    return r, grads


def fastgs(means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, cam_pos, active_sh_bases, width, height,
           fx, fy, cx, cy, near=0.01, far=1e10, grad_image=None, grad_alpha=None, want_w2c_grad=False,
           densification_info=None, prec=64):
    """Reference fastgs (EWA) forward [+ backward when grad_image is given]; reference tensor layouts
    (fastgs/rasterization/include/rasterization_api.h:26-75). Returns a dict."""
    dt, pfx, cf = _dt(prec)
    means, scales_raw, rotations_raw = _a(means, dt), _a(scales_raw, dt), _a(rotations_raw, dt)
    opacities_raw = _a(np.asarray(opacities_raw).reshape(-1), dt)
    N = means.shape[0]
    sh0 = _a(np.asarray(sh0).reshape(N, 3), dt)
    shN = _a(np.asarray(shN).reshape(N, -1, 3), dt)
    total_rest = shN.shape[1]
    w2c, cam_pos = _a(np.asarray(w2c).reshape(4, 4), dt), _a(np.asarray(cam_pos).reshape(3), dt)
    image = np.zeros((3, height, width), dt)
    alpha = np.zeros((1, height, width), dt)
    n_touched = np.zeros(N, np.int32)
    out = {}
    bwd = grad_image is not None
    if bwd:
        gi, ga = _a(grad_image, dt), _a(np.asarray(grad_alpha).reshape(height, width), dt)
        g = dict(means=np.zeros((N, 3), dt), scales_raw=np.zeros((N, 3), dt), rotations_raw=np.zeros((N, 4), dt),
                 opacities_raw=np.zeros((N, 1), dt), sh0=np.zeros((N, 1, 3), dt), shN=np.zeros((N, total_rest, 3), dt))
        gw = np.zeros((4, 4), dt) if want_w2c_grad else None
        dens = None if densification_info is None else _a(densification_info, dt).copy()
    else:
        gi = ga = gw = dens = None
        g = dict(means=None, scales_raw=None, rotations_raw=None, opacities_raw=None, sh0=None, shN=None)
    fn = getattr(lib(), pfx + "fastgs")
    fn.restype = C.c_int64
    n_inst = fn(C.c_int(N), _p(means), _p(scales_raw), _p(rotations_raw), _p(opacities_raw), _p(sh0), _p(shN),
                C.c_int(total_rest), _p(w2c), _p(cam_pos), C.c_int(active_sh_bases), C.c_int(width), C.c_int(height),
                cf(fx), cf(fy), cf(cx), cf(cy), cf(near), cf(far), _p(image), _p(alpha), _p(n_touched), _p(gi), _p(ga),
                _p(g["means"]), _p(g["scales_raw"]), _p(g["rotations_raw"]), _p(g["opacities_raw"]), _p(g["sh0"]),
                _p(g["shN"]), _p(gw), _p(dens))
    out.update(image=image, alpha=alpha, n_touched=n_touched, n_instances=int(n_inst))
    if bwd:
        out.update(grads=g, grad_w2c=gw, densification_info=dens)
    return out


def fastgs_inputs(sc, view=0):
    """(w2c [4,4], cam_position [3], fx, fy, cx, cy) of a scene view, as fast_rasterizer.cpp:20-45 derives them."""
    vm = np.asarray(sc.viewmats[view], np.float64)
    K = np.asarray(sc.Ks[view], np.float64)
    cam_pos = -vm[:3, :3].T @ vm[:3, 3]
    return vm, cam_pos, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


# ---- photometric loss: L1 + fused SSIM ("valid" crop), the reference's training loss ------------------------
# Trainer::compute_photometric_loss (src/training/trainer.cpp:103-131), fused_ssim (include/kernels/fused_ssim.cuh:27-122),
# kernels src/training/kernels/ssim.cu:64-460.  11-tap separable Gaussian (ssim.cu:17-28), zero padding (:45-53).
_SSIM_G = np.array([0.001028380123898387, 0.0075987582094967365, 0.036000773310661316, 0.10936068743467331,
                    0.21300552785396576, 0.26601171493530273, 0.21300552785396576, 0.10936068743467331,
                    0.036000773310661316, 0.0075987582094967365, 0.001028380123898387], np.float32).astype(np.float64)


def _gconv(a):
    """zero-padded separable 11x11 Gaussian of a [H,W] map (horizontal pass then vertical, ssim.cu:118-236)."""
    H, W = a.shape
    p = np.pad(a, ((0, 0), (5, 5)))
    h = sum(_SSIM_G[k] * p[:, k:k + W] for k in range(11))
    p = np.pad(h, ((5, 5), (0, 0)))
    return sum(_SSIM_G[k] * p[k:k + H, :] for k in range(11))


def photometric_loss(image, target, lambda_dssim, clamp=True):
    """image [H,W,3] (render incl. background, before the [0,1] clamp of rasterizer.cpp:401), target [H,W,3] in [0,1].
    Returns (loss, dL/dimage [H,W,3], parts dict)."""
    image, Y = np.asarray(image, np.float64), np.asarray(target, np.float64)
    H, W, _ = image.shape
    X = np.clip(image, 0.0, 1.0) if clamp else image
    passes = ((image >= 0.0) & (image <= 1.0)) if clamp else np.ones_like(image, bool)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    crop = np.zeros((H, W))
    if H > 10 and W > 10:  # fused_ssim.cuh:63-70 "valid"
        crop[5:H - 5, 5:W - 5] = 1.0
    else:
        crop[:] = 1.0
    count = 3.0 * crop.sum()
    ssim_sum, g = 0.0, np.zeros_like(X)
    for c in range(3):
        x, y = X[..., c], Y[..., c]
        mu1, mu2 = _gconv(x), _gconv(y)
        s1, s2, s12 = _gconv(x * x) - mu1 * mu1, _gconv(y * y) - mu2 * mu2, _gconv(x * y) - mu1 * mu2
        A, B = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
        Cc, D = 2.0 * mu1 * mu2 + C1, 2.0 * s12 + C2
        m = (Cc * D) / (A * B)  # ssim.cu:262
        d_mu1 = (mu2 * 2.0 * D) / (A * B) - (mu2 * 2.0 * Cc) / (A * B) - (mu1 * 2.0 * Cc * D) / (A * A * B) + (
            mu1 * 2.0 * Cc * D) / (A * B * B)  # :269
        d_s1, d_s12 = (-Cc * D) / (A * B * B), (2.0 * Cc) / (A * B)  # :270-271
        ssim_sum += (m * crop).sum()
        dmap = -lambda_dssim / count * crop  # d(lambda (1 - mean))/dmap
        g[..., c] = _gconv(dmap * d_mu1) + 2.0 * x * _gconv(dmap * d_s1) + y * _gconv(dmap * d_s12)  # :417
    l1 = np.abs(X - Y).mean()
    g += (1.0 - lambda_dssim) * np.sign(X - Y) / X.size
    loss = (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim_sum / count)
    return loss, g * passes, dict(l1=l1, ssim=ssim_sum / count)
