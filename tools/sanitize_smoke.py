"""A small pass over every kernel family of liblfs_b200.so, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
    compute-sanitizer --tool initcheck python tools/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tools/sanitize_smoke.py

Sizes are tiny (the tools slow kernels down 10-100 x); results are only checked for being finite -- parity is tests/."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lichtfeld_studio_b200 import ops, scene as S  # noqa: E402
from lichtfeld_studio_b200.trainer import SplatTrainer  # noqa: E402

dev = "cuda:0"
T = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev).contiguous()  # noqa: E731
n, w, h, deg, views = 1500, 160, 112, 3, 3
sc = S.make_scene(n, views, w, h, deg, seed=3, sigma_px=4.0)


def finite(*ts):
    for t in ts:
        assert torch.isfinite(t.float()).all(), "non-finite output"


# ---- fused trainer: single-view calls, then a step with two views in flight, then state surgery
tr = SplatTrainer(n, w, h, deg, dev, view_streams=2)
tr.load_scene(sc)
img, al = tr.forward(sc.viewmats[0], sc.Ks[0], deg, (0.1, 0.2, 0.3), want_image=True)
tr.loss_ssim_l1(T(S.make_target(0, w, h), torch.uint8), 0.2)
tr.backward()
finite(img, al, tr.grads)
tr.iteration = 1000
targets = [torch.as_tensor(S.make_target(v, w, h)).pin_memory() for v in range(views)]
for _ in range(2):
    loss = tr.train_step(sc.viewmats, sc.Ks, targets, (0.0, 0.0, 0.0), deg)
torch.cuda.synchronize()
assert np.isfinite(float(loss))
finite(tr.params)
keep = torch.arange(0, n, 2, device=dev, dtype=torch.int32)
tr.restructure(keep)
torch.cuda.synchronize()

# ---- gsplat surface: projection (pinhole + fisheye + rolling shutter), SH, intersect (both layouts), blend fwd / bwd
means, q, s, op, shs = sc.activated()
tm, tq, ts_, to = T(means), T(q), T(s), T(op)
tw, th = (w + 15) // 16, (h + 15) // 16
for model, rs, vm1, radial in ((0, 4, None, None), (2, 4, None, [[-0.02, 0.004, 0.0, 0.0]]), (0, 0, sc.viewmats[1:2], None)):
    kw = dict(camera_model=model, rs_type=rs, radial_coeffs=None if radial is None else T(radial))
    vm0, tK = T(sc.viewmats[:1]), T(sc.Ks[:1])
    tvm1 = None if vm1 is None else T(vm1)
    rad, m2d, dep, con, _ = ops.projection_ut_3dgs_fused(tm, tq, ts_, to, vm0, tvm1, tK, w, h, 0.3, 0.01, 1e4, 0.0, False, **kw)
    _, ids, flat = ops.intersect_tile(m2d, rad, dep, None, None, 1, 16, tw, th, True)
    offs = ops.intersect_offset(ids, 1, tw, th)
    for ch in (3, 4):
        colors = torch.rand((1, n, ch), device=dev)
        bg = torch.rand((1, ch), device=dev)
        r, a, li = ops.rasterize_to_pixels_from_world_3dgs_fwd(tm, tq, ts_, colors, to[None].contiguous(), bg, None, w, h, 16, vm0,
                                                               tvm1, tK, tile_offsets=offs, flatten_ids=flat, **kw)
        g = ops.rasterize_to_pixels_from_world_3dgs_bwd(tm, tq, ts_, colors, to[None].contiguous(), bg, None, w, h, 16, vm0, tvm1,
                                                        tK, tile_offsets=offs, flatten_ids=flat, render_alphas=a, last_ids=li,
                                                        v_render_colors=torch.randn_like(r), v_render_alphas=torch.randn_like(a),
                                                        **kw)
        finite(r, a, *g)
sel = torch.nonzero((rad[0] > 0).all(-1)).reshape(-1)
p = ops.intersect_tile(m2d[0][sel].contiguous(), rad[0][sel].contiguous(), dep[0][sel].contiguous(), torch.zeros_like(sel),
                       sel.contiguous(), 1, 16, tw, th, True)
dirs = (tm - torch.linalg.inv(T(sc.viewmats[0]))[:3, 3][None]).contiguous()
col = ops.spherical_harmonics_fwd(deg, dirs, T(shs), None)
vco, vdi = ops.spherical_harmonics_bwd(shs.shape[1], deg, dirs, T(shs), None, torch.randn_like(col), True)
finite(col, vco, vdi)

# ---- fastgs surface
vm = sc.viewmats[0].astype(np.float64)
w2c, cam = T(sc.viewmats[0]), T(-vm[:3, :3].T @ vm[:3, 3])
P = dict(means=T(sc.means), scales=T(sc.scaling), rot=T(sc.rotation), op=T(sc.opacity.reshape(-1, 1)), sh0=T(sc.sh0), shN=T(sc.shN))
fx, fy, cx, cy = (float(sc.Ks[0, 0, 0]), float(sc.Ks[0, 1, 1]), float(sc.Ks[0, 0, 2]), float(sc.Ks[0, 1, 2]))
fimg, falpha, ctx = ops.fastgs_forward(P["means"], P["scales"], P["rot"], P["op"], P["sh0"], P["shN"], w2c, cam, 16, w, h, fx, fy,
                                       cx, cy, 0.01, 1e10)
fg = ops.fastgs_backward(ctx, torch.randn_like(fimg), torch.randn_like(falpha), P["means"], P["scales"], P["rot"], P["shN"], w2c,
                         cam, densification_info=torch.zeros((2, n), device=dev))
finite(fimg, falpha, *[x for x in fg[:6]])

# ---- legacy 2-D ops
cov, pre = ops.quat_scale_to_covar_preci_fwd(tq, ts_, True, True, True)
vq, vs = ops.quat_scale_to_covar_preci_bwd(tq, ts_, True, torch.ones_like(cov), torch.ones_like(pre) * 1e-6)
rad2, m2, d2, c2, _ = ops.projection_ewa_3dgs_fused_fwd(tm, None, tq, ts_, to, T(sc.viewmats[:1]), T(sc.Ks[:1]), w, h, 0.3, 0.01, 1e4,
                                                        0.0, False)
_, ids2, flat2 = ops.intersect_tile(m2, rad2, d2, None, None, 1, 16, tw, th, True)
off2 = ops.intersect_offset(ids2, 1, tw, th)
colors = torch.rand((1, n, 3), device=dev)
r2, a2, l2 = ops.rasterize_to_pixels_3dgs_fwd(m2, c2, colors, to[None].contiguous(), None, None, w, h, 16, off2, flat2)
g2 = ops.rasterize_to_pixels_3dgs_bwd(m2, c2, colors, to[None].contiguous(), None, None, w, h, 16, off2, flat2, a2, l2,
                                      torch.randn_like(r2), torch.randn_like(a2), absgrad=True)
finite(cov, pre, vq, vs, r2, a2, *g2)
torch.cuda.synchronize()
print("sanitize_smoke: ok")
