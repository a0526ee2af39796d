"""Stage-by-stage GPU diagnostics: runs every op of liblfs_b200.so against the CPU oracle (double) and, when
oracle/_ref/libgsplat_ref.so is present, against the UNMODIFIED reference CUDA kernels, and prints/returns error
metrics without stopping at the first mismatch.  Used by the -m gpu tests (which assert on the same metrics) and
directly:  python tests/gpu_diag.py [--json gpurun_out/diag.json]
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import lichtfeld_studio_b200 as L  # noqa: E402
import oracle as O  # noqa: E402
import ref_libs as R  # noqa: E402
from lichtfeld_studio_b200 import ops, scene  # noqa: E402
from lichtfeld_studio_b200.trainer import SplatTrainer  # noqa: E402

DEV = "cuda:0"


def T(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dtype).to(DEV).contiguous()


def relerr(a, b):
    """max |a-b| / max(max|b|, tiny): per-tensor relative error (the north_star's 1e-4 / 1e-3 gates)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(a - b).max() if a.size else 0.0
    return float(d / max(np.abs(b).max() if b.size else 0.0, 1e-30))


def make_inputs(n, views, w, h, deg, seed=5, sigma_px=3.5):
    sc = scene.make_scene(n, views, w, h, deg, seed=seed, sigma_px=sigma_px)
    means, q, s, op, shs = sc.activated()
    return sc, means, q, s, op, shs


def diag_projection(n=4000, views=2, w=320, h=240):
    sc, means, q, s, op, shs = make_inputs(n, views, w, h, 0)
    out = {}
    got = ops.projection_ut_3dgs_fused(T(means), T(q), T(s), T(op), T(sc.viewmats), None, T(sc.Ks), w, h, 0.3, 0.01,
                                       1e4, 0.0, True)
    radii, m2d, dep, con, comp = [x.cpu().numpy() for x in got]
    o_radii, o_m2d, o_dep, o_con, o_comp = O.projection_ut(means, q, s, op, sc.viewmats, sc.Ks, w, h,
                                                           calc_compensations=True)
    vis_g, vis_o = (radii > 0).all(-1), (o_radii > 0).all(-1)
    both = vis_g & vis_o
    out["n_visible_oracle"] = int(vis_o.sum())
    out["visibility_mismatch"] = int((vis_g != vis_o).sum())
    out["radii_max_diff"] = int(np.abs(radii[both] - o_radii[both]).max())
    out["radii_n_diff"] = int((radii[both] != o_radii[both]).any(-1).sum())
    out["means2d_rel"] = relerr(m2d[both], o_m2d[both])
    out["depths_rel"] = relerr(dep[both], o_dep[both])
    out["conics_rel"] = relerr(con[both], o_con[both])
    out["comp_rel"] = relerr(comp[both], o_comp[both])
    if R.have_gsplat():
        rr, rm, rd, rc_, rcomp = [x.cpu().numpy() for x in R.projection_ut(T(means), T(q), T(s), T(op),
                                                                            T(sc.viewmats), T(sc.Ks), w, h,
                                                                            calc_comp=True)]
        vis_r = (rr > 0).all(-1)
        b2 = vis_g & vis_r
        out["ref_visibility_mismatch"] = int((vis_g != vis_r).sum())
        out["ref_radii_max_diff"] = int(np.abs(radii[b2] - rr[b2]).max())
        out["ref_radii_n_diff"] = int((radii[b2] != rr[b2]).any(-1).sum())
        out["ref_means2d_rel"] = relerr(m2d[b2], rm[b2])
        out["ref_depths_rel"] = relerr(dep[b2], rd[b2])
        out["ref_depth_bits_equal_frac"] = float((dep[b2].view(np.uint32) == rd[b2].view(np.uint32)).mean())
        out["ref_conics_rel"] = relerr(con[b2], rc_[b2])
        # how far is the reference itself from the double oracle (fp32 noise floor of the algorithm)
        b3 = vis_r & vis_o
        out["ref_vs_oracle_means2d_rel"] = relerr(rm[b3], o_m2d[b3])
        out["ref_vs_oracle_conics_rel"] = relerr(rc_[b3], o_con[b3])
    return out


def diag_sh(n=5000):
    rng = np.random.RandomState(3)
    out = {}
    dirs = rng.normal(size=(n, 3)).astype(np.float32)
    masks = rng.uniform(size=n) > 0.2
    for deg in range(5):
        for K in sorted({(deg + 1) ** 2, 16 if deg <= 3 else 25}):
            coeffs = rng.normal(size=(n, K, 3)).astype(np.float32)
            vcol = rng.normal(size=(n, 3)).astype(np.float32)
            col = ops.spherical_harmonics_fwd(deg, T(dirs), T(coeffs), T(masks, torch.bool))
            col = torch.where(T(masks, torch.bool)[:, None], col, torch.zeros_like(col)).cpu().numpy()
            o_col = O.sh_fwd(deg, dirs, coeffs, masks.astype(np.uint8))
            vco, vdi = ops.spherical_harmonics_bwd(K, deg, T(dirs), T(coeffs), T(masks, torch.bool), T(vcol), True)
            o_vco, o_vdi = O.sh_bwd(deg, dirs, coeffs, vcol, masks.astype(np.uint8))
            key = f"d{deg}_k{K}"
            out[key + "_fwd_rel"] = relerr(col, o_col)
            out[key + "_vcoef_rel"] = relerr(vco.cpu().numpy(), o_vco)
            out[key + "_vdir_rel"] = relerr(vdi.cpu().numpy(), o_vdi)
            if R.have_gsplat():
                r_col = R.sh_fwd(deg, T(dirs), T(coeffs), T(masks, torch.bool)).cpu().numpy()
                r_vco, r_vdi = R.sh_bwd(deg, T(dirs), T(coeffs), T(vcol), T(masks, torch.bool))
                out[key + "_ref_fwd_rel"] = relerr(col, r_col)
                out[key + "_ref_vcoef_rel"] = relerr(vco.cpu().numpy(), r_vco.cpu().numpy())
                out[key + "_ref_vdir_rel"] = relerr(vdi.cpu().numpy(), r_vdi.cpu().numpy())
    return out


def diag_intersect():
    out = {}
    rng = np.random.RandomState(11)
    cases = [("c1", 1, 3000, 320, 240, 40), ("c3", 3, 700, 130, 70, 25), ("big", 1, 60000, 1920, 1080, 60),
             ("empty", 1, 50, 64, 64, 1), ("pow2", 2, 100, 64, 64, 20)]
    for tag, Cc, N, W, H, rmax in cases:
        tw, th = (W + 15) // 16, (H + 15) // 16
        m2d = (rng.uniform(-0.2, 1.2, size=(Cc, N, 2)) * np.array([W, H])).astype(np.float32)
        radii = rng.randint(0, rmax, size=(Cc, N, 2)).astype(np.int32)
        dep = rng.uniform(0.05, 20.0, size=(Cc, N)).astype(np.float32)
        dep[:, ::7] = dep[:, 3:4]  # force depth ties: order must fall back to the Gaussian index
        for srt in (True, False):
            tpg, ids, flat = ops.intersect_tile(T(m2d), T(radii, torch.int32), T(dep), None, None, Cc, 16, tw, th, srt)
            o_tpg, o_ids, o_flat = O.intersect_tile(m2d, radii, dep, 16, tw, th, srt)
            k = f"{tag}_{'sorted' if srt else 'unsorted'}"
            out[k + "_n"] = int(len(o_ids))
            out[k + "_tpg_equal"] = bool(np.array_equal(tpg.cpu().numpy(), o_tpg))
            out[k + "_ids_equal"] = bool(ids.numel() == len(o_ids) and np.array_equal(ids.cpu().numpy(), o_ids))
            out[k + "_flat_equal"] = bool(flat.numel() == len(o_flat) and np.array_equal(flat.cpu().numpy(), o_flat))
            if srt:
                off = ops.intersect_offset(ids, Cc, tw, th).cpu().numpy()
                out[k + "_offsets_equal"] = bool(np.array_equal(off, O.intersect_offset(o_ids, Cc, tw, th)))
            # packed layout (gsplat/Intersect.cpp:32-39): the visible elements only, each with its camera id; keys and order
            # must be those of the [C,N] call, flatten_ids index the packed list
            sel = np.flatnonzero((radii.reshape(-1, 2) > 0).all(-1))
            if tag != "empty" and len(sel):
                cam_ids = (sel // N).astype(np.int64)
                p_tpg, p_ids, p_flat = ops.intersect_tile(T(m2d.reshape(-1, 2)[sel]), T(radii.reshape(-1, 2)[sel], torch.int32),
                                                          T(dep.reshape(-1)[sel]), T(cam_ids, torch.int64),
                                                          T((sel % N).astype(np.int64), torch.int64), Cc, 16, tw, th, srt)
                out[k + "_packed_tpg_equal"] = bool(np.array_equal(p_tpg.cpu().numpy(), o_tpg.reshape(-1)[sel]))
                out[k + "_packed_ids_equal"] = bool(p_ids.numel() == len(o_ids) and np.array_equal(p_ids.cpu().numpy(), o_ids))
                out[k + "_packed_flat_equal"] = bool(p_flat.numel() == len(o_flat)
                                                     and np.array_equal(sel[p_flat.cpu().numpy()], o_flat))
            if R.have_gsplat() and tag != "empty":
                r_tpg, r_ids, r_flat = R.intersect_tile(T(m2d), T(radii, torch.int32), T(dep), 16, tw, th, srt)
                out[k + "_ref_tpg_equal"] = bool(torch.equal(tpg, r_tpg))
                out[k + "_ref_ids_equal"] = bool(ids.shape == r_ids.shape and torch.equal(ids, r_ids))
                out[k + "_ref_flat_equal"] = bool(flat.shape == r_flat.shape and torch.equal(flat, r_flat))
                if srt and ids.numel():
                    out[k + "_ref_offsets_equal"] = bool(torch.equal(ops.intersect_offset(ids, Cc, tw, th),
                                                                     R.intersect_offset(r_ids, Cc, tw, th)))
    return out


def _raster_case(n, w, h, seed, sigma_px, with_bg):
    sc, means, q, s, op, shs = make_inputs(n, 1, w, h, 1, seed=seed, sigma_px=sigma_px)
    r = O.render_view(means, q, s, op, shs, 1, sc.viewmats[0], sc.Ks[0], w, h, bg=[0.3, 0.2, 0.1] if with_bg else None)
    return sc, means, q, s, op, r


def diag_raster(n=1500, w=200, h=136, seed=9, sigma_px=4.0, with_bg=True, fwd_variant=0, bwd_variant=0):
    L.load().lfs_set_option(b"fwd_variant", fwd_variant)
    L.load().lfs_set_option(b"bwd_variant", bwd_variant)
    out = {}
    sc, means, q, s, op, r = _raster_case(n, w, h, seed, sigma_px, with_bg)
    bg = T(np.array([[0.3, 0.2, 0.1]])) if with_bg else None
    tm, tq, ts, tc, to = T(means), T(q), T(s), T(r["colors"][None]), T(op[None])
    tvm, tK = T(sc.viewmats[:1]), T(sc.Ks[:1])
    toff, tflat = T(r["offsets"], torch.int32), T(r["flatten_ids"], torch.int32)
    out["n_isects"] = int(len(r["flatten_ids"]))
    ren, al, li = ops.rasterize_to_pixels_from_world_3dgs_fwd(tm, tq, ts, tc, to, bg, None, w, h, 16, tvm, None, tK,
                                                              tile_offsets=toff, flatten_ids=tflat)
    out["fwd_rgb_rel"] = relerr(ren.cpu().numpy(), r["renders"])
    out["fwd_alpha_rel"] = relerr(al.cpu().numpy(), r["alphas"])
    out["fwd_last_ids_mismatch"] = int((li.cpu().numpy() != r["last_ids"]).sum())
    out["n_pixels"] = int(w * h)
    rng = np.random.RandomState(1)
    vC = rng.normal(size=(1, h, w, 3)).astype(np.float32)
    vA = rng.normal(size=(1, h, w, 1)).astype(np.float32)
    g = ops.rasterize_to_pixels_from_world_3dgs_bwd(tm, tq, ts, tc, to, bg, None, w, h, 16, tvm, None, tK,
                                                    tile_offsets=toff, flatten_ids=tflat, render_alphas=al,
                                                    last_ids=li, v_render_colors=T(vC), v_render_alphas=T(vA))
    og = O.raster_world_bwd(means, q, s, r["colors"][None], op[None], None if bg is None else bg.cpu().numpy(), None, w,
                            h, 16, sc.viewmats[:1], sc.Ks[:1], r["offsets"], r["flatten_ids"], r["alphas"],
                            r["last_ids"], vC, vA)
    for nm, a, b in zip(("v_means", "v_quats", "v_scales", "v_colors", "v_opacities"), g, og):
        out["bwd_" + nm + "_rel"] = relerr(a.cpu().numpy(), b)
    if R.have_gsplat():
        rren, ral, rli = R.raster_fwd(tm, tq, ts, tc, to, bg, w, h, 16, tvm, tK, toff, tflat)
        out["ref_fwd_rgb_rel"] = relerr(ren.cpu().numpy(), rren.cpu().numpy())
        out["ref_fwd_alpha_rel"] = relerr(al.cpu().numpy(), ral.cpu().numpy())
        out["ref_fwd_last_ids_mismatch"] = int((li != rli).sum())
        out["ref_vs_oracle_fwd_rgb_rel"] = relerr(rren.cpu().numpy(), r["renders"])
        rg = R.raster_bwd(tm, tq, ts, tc, to, bg, w, h, 16, tvm, tK, toff, tflat, ral, rli, T(vC), T(vA))
        for nm, a, b, c in zip(("v_means", "v_quats", "v_scales", "v_colors", "v_opacities"), g, rg, og):
            out["ref_bwd_" + nm + "_rel"] = relerr(a.cpu().numpy(), b.cpu().numpy())
            out["ref_vs_oracle_bwd_" + nm + "_rel"] = relerr(b.cpu().numpy(), c)
    return out


def diag_adam(n=100003):
    rng = np.random.RandomState(2)
    p, g = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
    m, v = (0.1 * rng.normal(size=n)).astype(np.float32), (0.01 * rng.uniform(size=n)).astype(np.float32)
    tp, tm, tv, tg = T(p), T(m), T(v), T(g)
    args = (1e-3, 0.9, 0.999, 1e-15, 1.0 / (1 - 0.9 ** 3), 1.0 / np.sqrt(1 - 0.999 ** 3))
    ops.adam_step(tp, tm, tv, tg, *args)
    f32 = lambda x: float(np.float32(x))  # the kernel (like the reference's) receives fp32-rounded scalars
    op_, om, ov = O.adam_step(p, m, v, g, *[f32(x) for x in args])
    out = {"param_rel": relerr(tp.cpu().numpy(), op_), "m_rel": relerr(tm.cpu().numpy(), om),
           "v_rel": relerr(tv.cpu().numpy(), ov),
           "update_rel": relerr(tp.cpu().numpy() - p, op_ - p)}
    if R.have_fastgs():
        rp, rm, rv = T(p), T(m), T(v)
        R.FastGS().adam_step(rp, rm, rv, T(g), *args)
        torch.cuda.synchronize()
        out["ref_update_rel"] = relerr(tp.cpu().numpy() - p, rp.cpu().numpy() - p)
        out["ref_max_ulp"] = float((np.abs(tp.cpu().numpy().view(np.int32).astype(np.int64)
                                           - rp.cpu().numpy().view(np.int32).astype(np.int64))).max())
    return out


def diag_trainer(n=3000, w=240, h=160, deg=3, views=2, seed=21, fwd_variant=0, bwd_variant=0, lambda_dssim=None, cull=1):
    L.load().lfs_set_option(b"exact_cull", cull)
    L.load().lfs_set_option(b"fwd_variant", fwd_variant)
    L.load().lfs_set_option(b"bwd_variant", bwd_variant)
    out = {}
    sc = scene.make_scene(n, views, w, h, deg, seed=seed, sigma_px=4.0)
    tr = SplatTrainer(n, w, h, deg, DEV)
    tr.load_scene(sc)
    raw = dict(means=sc.means, sh0=sc.sh0, shN=sc.shN, scaling=sc.scaling, rotation=sc.rotation, opacity=sc.opacity)
    back = tr.export_params()
    out["pack_roundtrip_exact"] = bool(all(np.array_equal(back[k].cpu().numpy(), raw[k]) for k in raw))
    bg = (0.1, 0.3, 0.2)
    tot = {k: np.zeros_like(v, dtype=np.float64) for k, v in raw.items()}
    loss_o = 0.0
    for v in range(views):
        tgt = scene.make_target(v, w, h)
        img, alpha = tr.forward(sc.viewmats[v], sc.Ks[v], deg, bg, want_image=True)
        if lambda_dssim is None:
            tr.loss_l1(T(tgt, torch.uint8))
        else:
            tr.loss_ssim_l1(T(tgt, torch.uint8), lambda_dssim)
        tr.backward()
        n_inst, n_b = tr.stats()
        r, g = O.view_loss_grads(raw, sc.viewmats[v], sc.Ks[v], w, h, deg, bg, target=tgt, lambda_dssim=lambda_dssim)
        loss_o += r["loss"]
        out[f"v{v}_n_inst"] = n_inst
        out[f"v{v}_n_inst_oracle"] = int(len(r["flatten_ids"]))
        out[f"v{v}_image_rel"] = relerr(img.cpu().numpy(), r["renders"][0])
        out[f"v{v}_alpha_rel"] = relerr(alpha.cpu().numpy(), r["alphas"][0, :, :, 0])
        # the same view through the gsplat-surface ops (same device code, composed by the caller as
        # rasterizer.cpp:208-360 does) must agree with the fused step to rounding
        pm, pq, ps, po, pshs = sc.activated()
        tm_, tq_, ts_, to_ = T(pm), T(pq), T(ps), T(po)
        tvm, tK = T(sc.viewmats[v:v + 1]), T(sc.Ks[v:v + 1])
        rad, m2d, dep, con, _ = ops.projection_ut_3dgs_fused(tm_, tq_, ts_, to_, tvm, None, tK, w, h, 0.3, 0.01, 1e4,
                                                            0.0, False)
        campos = torch.linalg.inv(tvm[0])[:3, 3]
        dirs = (tm_ - campos[None]).contiguous()
        msk = (rad[0] > 0).all(-1).contiguous()
        cols = ops.spherical_harmonics_fwd(deg, dirs, T(pshs), msk)
        cols = torch.where(msk[:, None], torch.clamp_min(cols + 0.5, 0.0), torch.zeros_like(cols))
        tw_, th_ = (w + 15) // 16, (h + 15) // 16
        _, ids_, flat_ = ops.intersect_tile(m2d, rad, dep, None, None, 1, 16, tw_, th_, True)
        offs_ = ops.intersect_offset(ids_, 1, tw_, th_)
        ren_, al_, _ = ops.rasterize_to_pixels_from_world_3dgs_fwd(
            tm_, tq_, ts_, cols[None].contiguous(), to_[None].contiguous(), T(np.array([bg])), None, w, h, 16, tvm,
            None, tK, tile_offsets=offs_, flatten_ids=flat_)
        out[f"v{v}_n_inst_ops"] = int(flat_.numel())
        out[f"v{v}_image_vs_ops_rel"] = relerr(img.cpu().numpy(), ren_[0].cpu().numpy())
        out[f"v{v}_alpha_vs_ops_rel"] = relerr(alpha.cpu().numpy(), al_[0, :, :, 0].cpu().numpy())
        for k in tot:
            tot[k] += g[k]
    grads = tr.export_grads()
    for k in tot:
        out["grad_" + k + "_rel"] = relerr(grads[k].cpu().numpy(), tot[k])
    out["loss_rel"] = abs(float(tr.loss_dev.item()) - loss_o) / max(abs(loss_o), 1e-30)
    # Adam (all six groups; iteration > 1000 so that shN is stepped) against the oracle on the exported tensors
    p0 = {k: v.cpu().numpy().astype(np.float64) for k, v in tr.export_params().items()}
    g0 = {k: v.cpu().numpy().astype(np.float64) for k, v in grads.items()}
    tr.iteration = 1000
    lrs = dict(tr.lrs)
    tr.adam_step()
    torch.cuda.synchronize()
    p1 = tr.export_params()
    worst = 0.0
    for k in p0:
        f32 = lambda x: float(np.float32(x))
        want, _, _ = O.adam_step(p0[k], np.zeros_like(p0[k]), np.zeros_like(p0[k]), g0[k], f32(lrs[k]), f32(0.9),
                                 f32(0.999), f32(1e-15), f32(1.0 / (1 - 0.9)), f32(1.0 / np.sqrt(1 - 0.999)))
        got = p1[k].cpu().numpy().reshape(-1).astype(np.float64)
        # the update (~lr) is far below one ulp of a parameter of magnitude ~1, so the only meaningful gate is
        # "the stored fp32 parameter is within 1 ulp of the exactly rounded result"
        ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
        tol = ulp + 1e-6 * np.abs(want - p0[k].reshape(-1))  # 1 ulp of the parameter + 1e-6 of the update
        worst = max(worst, float((np.abs(got - want) / tol).max()))
    out["adam_param_ulp_max"] = worst
    out["grads_cleared"] = bool(float(tr.grads.abs().max().item()) == 0.0)
    return out


def diag_ref_fastgs(n=5000, w=320, h=240, deg=3):
    """Runs the UNMODIFIED reference fastgs CUDA path (EWA) on a small scene: sanity of oracle/_ref/libfastgs_ref.so
    and of its glue, and a coarse cross-check of the two algorithms (EWA vs 3DGUT render the same scene alike)."""
    out = {}
    if not R.have_fastgs():
        return {"unavailable": True}
    sc = scene.make_scene(n, 1, w, h, deg, seed=8, sigma_px=3.5)
    fg = R.FastGS()
    w2c = T(sc.viewmats[0])
    campos = T(np.linalg.inv(sc.viewmats[0].astype(np.float64))[:3, 3])
    fx, fy, cx, cy = [float(x) for x in (sc.Ks[0, 0, 0], sc.Ks[0, 1, 1], sc.Ks[0, 0, 2], sc.Ks[0, 1, 2])]
    P = dict(means=T(sc.means), scales=T(sc.scaling), rot=T(sc.rotation), op=T(sc.opacity), sh0=T(sc.sh0), shN=T(sc.shN))
    img, alpha, counts = fg.forward(P["means"], P["scales"], P["rot"], P["op"], P["sh0"], P["shN"], w2c, campos,
                                    (deg + 1) ** 2, w, h, fx, fy, cx, cy)
    torch.cuda.synchronize()
    out["counts"] = list(counts)
    out["image_finite"] = bool(torch.isfinite(img).all())
    gimg = torch.randn_like(img)
    g = fg.backward(gimg, torch.zeros_like(alpha), img, alpha, P["means"], P["scales"], P["rot"], P["shN"], w2c, campos,
                    (deg + 1) ** 2, w, h, fx, fy, cx, cy)
    torch.cuda.synchronize()
    out["grads_finite"] = bool(all(torch.isfinite(v).all() for v in g.values()))
    out["grad_means_absmax"] = float(g["means"].abs().max())
    tr = SplatTrainer(n, w, h, deg, DEV)
    tr.load_scene(sc)
    ours, oa = tr.forward(sc.viewmats[0], sc.Ks[0], deg, (0, 0, 0), want_image=True)
    out["ewa_vs_gut_image_mean_abs_diff"] = float((ours.permute(2, 0, 1) - img).abs().mean())
    out["ewa_vs_gut_alpha_mean_abs_diff"] = float((oa - alpha[0]).abs().mean())
    return out


def run_all(fast=False):
    rep = {}
    t0 = time.time()
    for name, fn in (("projection", diag_projection), ("sh", diag_sh), ("intersect", diag_intersect),
                     ("raster_default", lambda: diag_raster()), ("raster_round1_kernels", lambda: diag_raster(fwd_variant=1, bwd_variant=1)),
                     ("raster_nobg_dense", lambda: diag_raster(n=4000, w=96, h=80, seed=4, sigma_px=7.0, with_bg=False)),
                     ("adam", diag_adam), ("trainer", diag_trainer), ("ref_fastgs", diag_ref_fastgs),
                     ("trainer_round1_kernels", lambda: diag_trainer(n=1200, w=100, h=84, deg=2, views=1, seed=2, fwd_variant=1,
                                                                     bwd_variant=1))):
        try:
            t = time.time()
            rep[name] = fn()
            rep[name]["_seconds"] = round(time.time() - t, 2)
        except Exception as e:  # keep going: one report per GPU call
            import traceback
            rep[name] = {"_error": repr(e), "_trace": traceback.format_exc()[-1500:]}
        torch.cuda.synchronize()
    rep["_total_seconds"] = round(time.time() - t0, 2)
    return rep


if __name__ == "__main__":
    rep = run_all()
    print(json.dumps(rep, indent=1, default=str))
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        json.dump(rep, open(path, "w"), indent=1, default=str)


def fastgs_case(n=3000, w=200, h=136, deg=3, seed=11, sigma_px=4.0):
    """Seeded inputs of the fastgs surface (raw parameters) for one view: numpy dict."""
    sc = scene.make_scene(n, 1, w, h, deg, seed=seed, sigma_px=sigma_px)
    w2c, cam, fx, fy, cx, cy = O.fastgs_inputs(sc)
    rng = np.random.RandomState(seed + 1)
    return dict(means=sc.means, scales_raw=sc.scaling, rotations_raw=sc.rotation, opacities_raw=sc.opacity.reshape(-1, 1),
                sh0=sc.sh0, shN=sc.shN, w2c=w2c.astype(np.float32), cam_pos=cam.astype(np.float32),
                active_sh_bases=(deg + 1) ** 2, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy,
                grad_image=rng.normal(size=(3, h, w)).astype(np.float32),
                grad_alpha=rng.normal(size=(1, h, w)).astype(np.float32))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def diag_fastgs(n=3000, w=200, h=136, deg=3, seed=11, sigma_px=4.0, with_ref=True):
    """lfs_fastgs_forward / backward vs the double-precision oracle (and vs the unmodified reference CUDA build)."""
    from lichtfeld_studio_b200 import ops
    c = fastgs_case(n, w, h, deg, seed, sigma_px)
    out = {}
    t = {k: T(c[k]) for k in ("means", "scales_raw", "rotations_raw", "opacities_raw", "sh0", "shN", "w2c", "cam_pos",
                              "grad_image", "grad_alpha")}
    img, alpha, ctx = ops.fastgs_forward(t["means"], t["scales_raw"], t["rotations_raw"], t["opacities_raw"], t["sh0"],
                                         t["shN"], t["w2c"], t["cam_pos"], c["active_sh_bases"], w, h, c["fx"], c["fy"],
                                         c["cx"], c["cy"], 0.01, 1e10)
    dens = torch.zeros((2, n), device=DEV)
    g = ops.fastgs_backward(ctx, t["grad_image"], t["grad_alpha"], t["means"], t["scales_raw"], t["rotations_raw"],
                            t["shN"], t["w2c"], t["cam_pos"], densification_info=dens, want_w2c_grad=True)
    torch.cuda.synchronize()
    names = ("means", "scales_raw", "rotations_raw", "opacities_raw", "sh0", "shN")
    orc = O.fastgs(c["means"], c["scales_raw"], c["rotations_raw"], c["opacities_raw"], c["sh0"], c["shN"], c["w2c"],
                   c["cam_pos"], c["active_sh_bases"], w, h, c["fx"], c["fy"], c["cx"], c["cy"], 0.01, 1e10,
                   grad_image=c["grad_image"], grad_alpha=c["grad_alpha"], want_w2c_grad=True,
                   densification_info=np.zeros((2, n)), prec=64)
    out["n_instances"], out["n_instances_oracle"] = ctx.n_instances, orc["n_instances"]
    out["n_visible"], out["n_visible_oracle"] = ctx.n_visible_primitives, int((orc["n_touched"] > 0).sum())
    out["n_buckets"] = ctx.n_buckets
    out["image_rel"], out["alpha_rel"] = _rel(img.cpu().numpy(), orc["image"]), _rel(alpha.cpu().numpy(), orc["alpha"])
    for k, gg in zip(names, g[:6]):
        out[f"grad_{k}_rel"] = _rel(gg.cpu().numpy().reshape(orc["grads"][k].shape), orc["grads"][k])
    out["grad_w2c_rel"] = _rel(g[6].cpu().numpy()[:3], orc["grad_w2c"][:3])
    out["dens_count_equal"] = bool((dens[0].cpu().numpy() == orc["densification_info"][0]).all())
    out["dens_norm_rel"] = _rel(dens[1].cpu().numpy(), orc["densification_info"][1])
    if with_ref and R.have_fastgs():
        fg = R.FastGS()
        rimg, ralpha, counts = fg.forward(t["means"], t["scales_raw"], t["rotations_raw"], t["opacities_raw"], t["sh0"],
                                          t["shN"], t["w2c"], t["cam_pos"], c["active_sh_bases"], w, h, c["fx"], c["fy"],
                                          c["cx"], c["cy"])
        rg = fg.backward(t["grad_image"], t["grad_alpha"], rimg, ralpha, t["means"], t["scales_raw"], t["rotations_raw"],
                         t["shN"], t["w2c"], t["cam_pos"], c["active_sh_bases"], w, h, c["fx"], c["fy"], c["cx"], c["cy"])
        torch.cuda.synchronize()
        out["ref_counts"] = list(counts)
        out["ref_image_rel"], out["ref_alpha_rel"] = _rel(img.cpu().numpy(), rimg.cpu().numpy()), _rel(
            alpha.cpu().numpy(), ralpha.cpu().numpy())
        out["ref_oracle_image_rel"] = _rel(rimg.cpu().numpy(), orc["image"])
        for k, rk, gg in zip(names, ("means", "scales", "rot", "op", "sh0", "shN"), g[:6]):
            out[f"ref_grad_{k}_rel"] = _rel(gg.cpu().numpy(), rg[rk].cpu().numpy().reshape(gg.shape))
            out[f"ref_oracle_grad_{k}_rel"] = _rel(rg[rk].cpu().numpy().reshape(orc["grads"][k].shape), orc["grads"][k])
    return out


def diag_cull_lossless(n=20000, w=640, h=360, deg=1, seed=31, sigma_px=5.0, capacity=None):
    """Exact tile culling must not change a single bit of the forward: a culled (tile, Gaussian) instance holds no
    pixel with alpha >= 1/255, and contributing pairs keep their order."""
    sc = scene.make_scene(n, 1, w, h, deg, seed=seed, sigma_px=sigma_px)
    out = {}
    imgs = []
    for cull in (0, 1):
        L.load().lfs_set_option(b"exact_cull", cull)
        tr = SplatTrainer(n, w, h, deg, DEV, instance_capacity=capacity or 40 * n)
        tr.load_scene(sc)
        img, alpha = tr.forward(sc.viewmats[0], sc.Ks[0], deg, (0.2, 0.1, 0.0), want_image=True)
        img2, _ = tr.forward(sc.viewmats[0], sc.Ks[0], deg, (0.2, 0.1, 0.0), want_image=True)
        out[f"deterministic_cull{cull}"] = bool(torch.equal(img, img2))
        out[f"finite_cull{cull}"] = bool(torch.isfinite(img).all() and (alpha >= 0).all() and (alpha <= 1).all())
        out[f"n_inst_cull{cull}"] = tr.stats()[0]
        imgs.append((img, alpha))
        del tr
        torch.cuda.empty_cache()
    L.load().lfs_set_option(b"exact_cull", 1)
    out["image_bit_identical"] = bool(torch.equal(imgs[0][0], imgs[1][0]) and torch.equal(imgs[0][1], imgs[1][1]))
    out["image_max_abs_diff"] = float((imgs[0][0] - imgs[1][0]).abs().max())
    return out
