"""Pins the oracle of the legacy 2-D op surface (oracle/lfs_oracle_legacy2d_impl.h):
  * quat/scale -> covariance, precision and the pinhole EWA projection against golden vectors produced by the reference's own
    CPU statement tests/torch_impl.cpp (tests/golden/make_golden.py, run on the unmodified file);
  * the analytic backward passes (quat/scale VJP, 2-D blend backward) by central finite differences of the oracle's own
    forward in double precision -- the reference tree has no implementation of these to run;
  * the 2-D blend forward against the from-world blend on small Gaussians far from the camera, where the EWA projection
    and the ray response agree (tolerance from the linearisation error, not from arithmetic)."""
import numpy as np
import pytest

import oracle as O


def test_quat_scale_to_covar_preci_matches_torch_impl(golden):
    q, s = golden["qs_quats"], golden["qs_scales"]
    for prec, tol in ((32, 2e-5), (64, 2e-6)):  # the golden vectors are fp32
        cov, pre = O.quat_scale_to_covar_preci_fwd(q, s, True, True, False, prec=prec)
        np.testing.assert_allclose(cov, golden["qs_covars"], rtol=tol, atol=tol * np.abs(golden["qs_covars"]).max())
        np.testing.assert_allclose(pre, golden["qs_precis"], rtol=tol * 10, atol=tol * np.abs(golden["qs_precis"]).max())
    cov6, pre6 = O.quat_scale_to_covar_preci_fwd(q, s, True, True, True)
    full, _ = O.quat_scale_to_covar_preci_fwd(q, s, True, False, False)
    np.testing.assert_allclose(cov6, full.reshape(-1, 9)[:, [0, 1, 2, 4, 5, 8]], rtol=1e-12)
    only_pre = O.quat_scale_to_covar_preci_fwd(q, s, False, True, True)
    assert only_pre[0] is None and np.allclose(only_pre[1], pre6)


@pytest.mark.parametrize("triu", [False, True])
def test_quat_scale_to_covar_preci_backward_fd(triu):
    rng = np.random.default_rng(3)
    n = 6
    q = rng.normal(size=(n, 4))
    s = np.exp(rng.normal(size=(n, 3)) * 0.3)
    shape = (n, 6) if triu else (n, 3, 3)
    vc, vp = rng.normal(size=shape), rng.normal(size=shape) * 0.1

    def loss(q_, s_):
        c, p = O.quat_scale_to_covar_preci_fwd(q_, s_, True, True, triu)
        return float((c * vc).sum() + (p * vp).sum())

    vq, vs = O.quat_scale_to_covar_preci_bwd(q, s, triu, vc, vp)
    h = 1e-6
    for arr, grad in ((q, vq), (s, vs)):
        for idx in np.ndindex(arr.shape):
            a0 = arr[idx]
            arr[idx] = a0 + h
            lp = loss(q, s)
            arr[idx] = a0 - h
            lm = loss(q, s)
            arr[idx] = a0
            fd = (lp - lm) / (2 * h)
            assert abs(fd - grad[idx]) <= 1e-6 * max(1.0, abs(fd)), (idx, fd, grad[idx])
    # covariance only / precision only
    vq_c, vs_c = O.quat_scale_to_covar_preci_bwd(q, s, triu, vc, None)
    vq_p, vs_p = O.quat_scale_to_covar_preci_bwd(q, s, triu, None, vp)
    np.testing.assert_allclose(vq_c + vq_p, vq, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(vs_c + vs_p, vs, rtol=1e-9, atol=1e-12)


def test_projection_ewa_matches_torch_impl(golden):
    N, W, H = [int(x) for x in golden["ewa_geom"]]
    for prec, tol in ((32, 5e-5), (64, 5e-6)):
        radii, m2d, dep, con, _ = O.projection_ewa(golden["ewa_means"], None, golden["ewa_quats"], golden["ewa_scales"], None,
                                                   golden["ewa_viewmat"], golden["ewa_K"], W, H, 0.3, 0.01, 1e4, prec=prec)
        want_r = golden["ewa_radii"]
        vis = (want_r > 0).all(-1)
        assert vis.sum() > N // 2
        np.testing.assert_array_equal(radii, want_r)  # fixed 3.33 sigma extent without opacities, as torch_impl.cpp:196-199
        np.testing.assert_allclose(m2d[vis], golden["ewa_means2d"][vis], rtol=tol, atol=tol * 100)
        np.testing.assert_allclose(dep[vis], golden["ewa_depths"][vis], rtol=tol)
        np.testing.assert_allclose(con[vis], golden["ewa_conics"][vis], rtol=tol * 4, atol=tol)
    # explicit covariances == quats + scales
    cov, _ = O.quat_scale_to_covar_preci_fwd(golden["ewa_quats"], golden["ewa_scales"], True, False, False)
    r2 = O.projection_ewa(golden["ewa_means"], cov, None, None, None, golden["ewa_viewmat"], golden["ewa_K"], W, H)
    r1 = O.projection_ewa(golden["ewa_means"], None, golden["ewa_quats"], golden["ewa_scales"], None, golden["ewa_viewmat"],
                          golden["ewa_K"], W, H)
    for a, b in zip(r1[:4], r2[:4]):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)


def test_projection_ewa_culling_tail():
    """opacity-aware extent, radius_clip, near/far, off-screen (ProjectionUT3DGSFused.cu:142-199)"""
    means = np.array([[0, 0, 5.0], [0, 0, 5.0], [0, 0, 0.001], [50.0, 0, 5.0], [0, 0, 5.0]])
    quats = np.tile([1.0, 0, 0, 0], (5, 1))
    scales = np.full((5, 3), 0.05)
    op = np.array([0.9, 0.003, 0.9, 0.9, 0.02])
    vm, K = np.eye(4)[None], np.array([[[200.0, 0, 64], [0, 200.0, 64], [0, 0, 1]]])
    radii, m2d, dep, con, comp = O.projection_ewa(means, None, quats, scales, op, vm, K, 128, 128, calc_compensations=True)
    assert (radii[0, 0] > 0).all() and (radii[0, 1] == 0).all() and (radii[0, 2] == 0).all() and (radii[0, 3] == 0).all()
    assert 0 < radii[0, 4, 0] < radii[0, 0, 0]  # low opacity -> tighter extent
    assert 0 < comp[0, 0] < 1 and m2d[0, 0, 0] == pytest.approx(64.0)
    clipped = O.projection_ewa(means, None, quats, scales, op, vm, K, 128, 128, radius_clip=100.0)
    assert not clipped[0].any()


def _scene2d(rng, C, N, W, H, CH):
    means2d = rng.uniform(-4, [W + 4, H + 4], size=(C, N, 2))
    a, c = rng.uniform(0.02, 0.3, size=(2, C, N))
    b = rng.uniform(-0.9, 0.9, size=(C, N)) * np.sqrt(a * c)
    conics = np.stack([a, b, c], -1)
    colors = rng.uniform(0, 1, size=(C, N, CH))
    opac = rng.uniform(0.05, 1.0, size=(C, N))
    radii = np.ceil(3.33 / np.sqrt(np.minimum(a, c) * (1 - 0.81)))[..., None].repeat(2, -1).astype(np.int32)
    depths = rng.uniform(1, 10, size=(C, N))
    tw, th = (W + 15) // 16, (H + 15) // 16
    _, ids, flat = O.intersect_tile(means2d.astype(np.float32), radii, depths.astype(np.float32), 16, tw, th)
    off = O.intersect_offset(ids, C, tw, th)
    return means2d, conics, colors, opac, off, flat


@pytest.mark.parametrize("CH,with_bg", [(3, True), (1, False), (5, True)])
def test_raster_2d_backward_fd(CH, with_bg):
    rng = np.random.default_rng(7 + CH)
    C, N, W, H = 2, 40, 40, 24
    means2d, conics, colors, opac, off, flat = _scene2d(rng, C, N, W, H, CH)
    bg = rng.uniform(0, 1, size=(C, CH)) if with_bg else None
    v_rc, v_ra = rng.normal(size=(C, H, W, CH)), rng.normal(size=(C, H, W, 1))

    def fwd(m, cn, col, op):
        return O.raster_2d_fwd(m, cn, col, op, bg, None, W, H, 16, off, flat)

    r, al, last = fwd(means2d, conics, colors, opac)
    assert last.max() > 0 and al.max() > 0.5
    g = O.raster_2d_bwd(means2d, conics, colors, opac, bg, None, W, H, 16, off, flat, al, last, v_rc, v_ra, absgrad=True)
    v_m, v_abs, v_con, v_col, v_op = g
    assert (v_abs >= np.abs(v_m) - 1e-12).all() and v_abs.max() > 0

    def loss(m, cn, col, op):
        rr, aa, _ = fwd(m, cn, col, op)
        return float((rr * v_rc).sum() + (aa * v_ra).sum())

    # the forward is piecewise smooth (1/255 cut-off, 0.999 clamp, T stop): probe entries whose +-h neighbourhood keeps
    # the contributor sets unchanged, i.e. where the two one-sided differences agree
    h, checked = 1e-6, 0
    args = [means2d, conics, colors, opac]
    for ai, grad in ((0, v_m), (1, v_con), (2, v_col), (3, v_op)):
        arr = args[ai]
        picks = rng.choice(arr.size, size=min(arr.size, 24), replace=False)
        for flat_i in picks:
            idx = np.unravel_index(flat_i, arr.shape)
            a0 = arr[idx]
            l0 = loss(*args)
            arr[idx] = a0 + h
            lp = loss(*args)
            arr[idx] = a0 - h
            lm = loss(*args)
            arr[idx] = a0
            fp, fm = (lp - l0) / h, (l0 - lm) / h
            if abs(fp - fm) > 1e-4 * max(1.0, abs(fp)):
                continue  # a contributor set changed inside the stencil
            fd = 0.5 * (fp + fm)
            assert abs(fd - grad[idx]) <= 2e-5 * max(1.0, abs(fd)), (ai, idx, fd, grad[idx])
            checked += 1
    assert checked >= 60


def test_raster_2d_masks_and_empty():
    rng = np.random.default_rng(1)
    C, N, W, H, CH = 1, 30, 36, 20, 3
    means2d, conics, colors, opac, off, flat = _scene2d(rng, C, N, W, H, CH)
    bg = np.array([[0.2, 0.4, 0.6]])
    masks = np.ones((C, 2, 3), np.uint8)
    masks[0, 0, 1] = 0
    r, al, last = O.raster_2d_fwd(means2d, conics, colors, opac, bg, masks, W, H, 16, off, flat)
    np.testing.assert_allclose(r[0, :16, 16:32], np.broadcast_to(bg[0], (16, 16, 3)))
    assert not al[0, :16, 16:32].any() and not last[0, :16, 16:32].any()
    g = O.raster_2d_bwd(means2d, conics, colors, opac, bg, masks, W, H, 16, off, flat, al, last, np.ones_like(r),
                        np.ones_like(al))
    r0, al0, _ = O.raster_2d_fwd(means2d, conics, colors, opac, bg, None, W, H, 16, off, flat)
    assert np.abs(r - r0)[0, 16:, :].max() == 0 and g[3].any()
    # no intersections at all: background, zero gradients
    e_off, e_flat = np.zeros((C, 2, 3), np.int32), np.zeros(0, np.int32)
    r, al, last = O.raster_2d_fwd(means2d, conics, colors, opac, bg, None, W, H, 16, e_off, e_flat)
    np.testing.assert_allclose(r, np.broadcast_to(bg[0], r.shape))
    assert not al.any()


def test_raster_2d_agrees_with_from_world_on_small_gaussians():
    """EWA projection + 2-D blend == from-world blend up to the linearisation error of the projection."""
    rng = np.random.default_rng(5)
    N, W, H = 60, 64, 48
    means = np.stack([rng.uniform(-0.5, 0.5, N), rng.uniform(-0.4, 0.4, N), rng.uniform(4, 6, N)], -1)
    quats = rng.normal(size=(N, 4))
    quats /= np.linalg.norm(quats, axis=-1, keepdims=True)
    scales = np.exp(rng.normal(size=(N, 3)) * 0.2 - 3.6)
    opac = rng.uniform(0.2, 0.9, size=N)
    colors = rng.uniform(0, 1, size=(1, N, 3))
    vm, K = np.eye(4)[None], np.array([[[120.0, 0, 32], [0, 120.0, 24], [0, 0, 1]]])
    radii, m2d, dep, con, _ = O.projection_ewa(means, None, quats, scales, opac, vm, K, W, H, eps2d=0.0)
    tw, th = (W + 15) // 16, (H + 15) // 16
    _, ids, flat = O.intersect_tile(m2d.astype(np.float32), radii, dep.astype(np.float32), 16, tw, th)
    off = O.intersect_offset(ids, 1, tw, th)
    r2, a2, _ = O.raster_2d_fwd(m2d, con, colors, opac[None], None, None, W, H, 16, off, flat)
    rw, aw, _ = O.raster_world_fwd(means, quats, scales, colors, opac[None], None, None, W, H, 16, vm, K, off, flat)
    assert a2.max() > 0.3
    assert np.abs(r2 - rw).max() <= 2e-2 and np.abs(a2 - aw).max() <= 2e-2
    assert np.abs(r2 - rw).mean() <= 1e-3
