"""The product's per-Gaussian device math checked on the HOST against the CPU oracle, no GPU.

tests/host_device_math.cpp includes csrc/projection.cuh, csrc/sh.cuh and csrc/intersect.cuh UNMODIFIED and compiles them with
g++ (host meanings for the handful of device intrinsics).  What the `-m gpu` tests prove bit-for-bit on the device is proven
here at fp32 rounding for the same source lines: unscented-transform projection (pinhole fast path and the general camera
path), spherical harmonics forward / VJP, the AABB tile rectangle, and that the exact tile culling never drops a tile that
holds a contributing pixel.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
from lichtfeld_studio_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_device_math.cpp")
CSRC = os.path.join(ROOT, "lichtfeld-studio_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build", "libhost_device_math.so")
CUDA_INC = "/usr/local/cuda/include"

PINHOLE, FISHEYE, RS_GLOBAL = 0, 2, 4


class UtParams(C.Structure):
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("kappa", C.c_float), ("margin", C.c_float), ("require_all", C.c_int32)]


@pytest.fixture(scope="module")
def lib():
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("projection.cuh", "cameras.cuh", "common.cuh", "sh.cuh", "intersect.cuh",
                                                    "sort_scan.cuh")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        r = subprocess.run([gxx, "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + CUDA_INC, "-o", OUT, SRC],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    lb = C.CDLL(OUT)
    lb.hd_ref_tile_n_bits.restype = C.c_uint32
    return lb


def _f(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _outs(n):
    return (np.zeros((n, 2), np.int32), np.zeros((n, 2), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32),
            np.zeros(n, np.float32))


def ut_pinhole(lib, means, q, s, op, vm, K, w, h, eps2d=0.3, near=0.01, far=1e4, clip=0.0, ut=(0.1, 2.0, 0.0, 0.1, 1)):
    means, q, s, op, vm, K = _f(means), _f(q), _f(s), _f(op), _f(vm), _f(K)
    n = len(means)
    radii, m2d, dep, con, comp = _outs(n)
    lib.hd_ut_project_pinhole(C.c_int(n), _p(means), _p(q), _p(s), _p(op), _p(vm), _p(K), C.c_int(w), C.c_int(h), C.c_float(eps2d),
                              C.c_float(near), C.c_float(far), C.c_float(clip), C.byref(UtParams(*ut)), _p(radii), _p(m2d), _p(dep),
                              _p(con), _p(comp))
    return radii, m2d, dep, con, comp


def ut_general(lib, means, q, s, op, vm0, vm1, K, w, h, model=PINHOLE, shutter=RS_GLOBAL, radial=None, tangential=None, prism=None,
               eps2d=0.3, near=0.01, far=1e4, clip=0.0, ut=(0.1, 2.0, 0.0, 0.1, 1)):
    means, q, s, op, vm0, vm1, K = _f(means), _f(q), _f(s), _f(op), _f(vm0), _f(vm1), _f(K)
    radial, tangential, prism = _f(radial), _f(tangential), _f(prism)
    n = len(means)
    radii, m2d, dep, con, comp = _outs(n)
    lib.hd_ut_project_general(C.c_int(n), _p(means), _p(q), _p(s), _p(op), _p(vm0), _p(vm1), _p(K), C.c_int(w), C.c_int(h),
                              C.c_int(model), C.c_int(shutter), _p(radial), _p(tangential), _p(prism), C.c_float(eps2d),
                              C.c_float(near), C.c_float(far), C.c_float(clip), C.byref(UtParams(*ut)), _p(radii), _p(m2d), _p(dep),
                              _p(con), _p(comp))
    return radii, m2d, dep, con, comp


def compare_projection(got, want, min_visible):
    radii, m2d, dep, con, comp = got
    o_radii, o_m2d, o_dep, o_con, o_comp = want
    vis_g, vis_o = (radii > 0).all(-1), (o_radii > 0).all(-1)
    both = vis_g & vis_o
    assert vis_o.sum() >= min_visible
    assert (vis_g != vis_o).sum() <= max(3, vis_o.sum() // 500)  # the gates of tests/test_gpu_parity.py::test_projection_ut
    assert np.abs(radii[both] - o_radii[both]).max() <= 1
    assert (radii[both] != o_radii[both]).any(-1).mean() < 0.01
    assert relerr(m2d[both], o_m2d[both]) <= 1e-4
    assert relerr(dep[both], o_dep[both]) <= 1e-5
    assert relerr(con[both], o_con[both]) <= 1e-3
    if o_comp is not None:
        assert relerr(comp[both], o_comp[both]) <= 1e-3


# ---- unscented-transform projection ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", [64, 32])
def test_ut_projection_pinhole_matches_oracle(lib, prec):
    n, w, h = 4000, 320, 240
    sc = S.make_scene(n, 2, w, h, 0, seed=5, sigma_px=3.5)
    means, q, s, op, _ = sc.activated()
    want = O.projection_ut(means, q, s, op, sc.viewmats, sc.Ks, w, h, calc_compensations=True, prec=prec)
    for v in range(2):
        got = ut_pinhole(lib, means, q, s, op, sc.viewmats[v], sc.Ks[v], w, h)
        compare_projection(got, [x[v] for x in want], 500)


@pytest.mark.parametrize("ut", [(0.1, 2.0, 0.0, 0.1, 0), (0.5, 2.0, 1.0, 0.3, 1), (1.0, 0.0, 0.0, 0.05, 1)])
def test_ut_projection_parameters(lib, ut):
    """alpha / beta / kappa / margin / require_all_sigma_points_valid reach the sums the way the oracle has them."""
    n, w, h = 3000, 256, 192
    sc = S.make_scene(n, 1, w, h, 0, seed=9, sigma_px=5.0)
    means, q, s, op, _ = sc.activated()
    means = (means * 3.0).astype(np.float32)  # spread past the frustum: behind the camera, beyond `far`, outside the image
    op = (op * np.random.RandomState(2).uniform(0.0, 1.0, n) ** 4).astype(np.float32)  # some below 1/255
    want = O.projection_ut(means, q, s, op, sc.viewmats, sc.Ks, w, h, eps2d=0.1, near=0.2, far=5.0, radius_clip=2.0, ut=ut,
                           calc_compensations=True)
    got = ut_pinhole(lib, means, q, s, op, sc.viewmats[0], sc.Ks[0], w, h, eps2d=0.1, near=0.2, far=5.0, clip=2.0, ut=ut)
    compare_projection(got, [x[0] for x in want], 300)
    culled = 1.0 - (want[0][0] > 0).all(-1).mean()
    assert 0.3 < culled < 0.95, culled  # every culling rule of ProjectionUT3DGSFused.cu:142-199 is in play


def test_ut_projection_without_opacities_and_degenerate_inputs(lib):
    n, w, h = 2000, 256, 192
    sc = S.make_scene(n, 1, w, h, 0, seed=11)
    means, q, s, op, _ = sc.activated()
    want = O.projection_ut(means, q, s, None, sc.viewmats, sc.Ks, w, h, calc_compensations=True)
    got = ut_pinhole(lib, means, q, s, None, sc.viewmats[0], sc.Ks[0], w, h)
    compare_projection(got, [x[0] for x in want], 300)
    # zero quaternion -> identity rotation (glm::normalize guard), opacity below 1/255 -> culled, behind the camera -> culled
    vm, K = sc.viewmats[0], sc.Ks[0]
    centre = -vm[:3, :3].T @ vm[:3, 3]
    fwd = vm[2, :3]
    m = np.stack([centre + 4 * fwd, centre + 4 * fwd, centre - 4 * fwd]).astype(np.float32)
    qq = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]], np.float32)
    ss = np.full((3, 3), 0.05, np.float32)
    r0 = ut_pinhole(lib, m, qq, ss, np.array([0.9, 0.9, 0.9], np.float32), vm, K, w, h)
    assert (r0[0][0] > 0).all() and (r0[0][0] == r0[0][1]).all() and np.allclose(r0[1][0], r0[1][1])
    assert (r0[0][2] == 0).all()
    r1 = ut_pinhole(lib, m, qq, ss, np.array([0.9, 1.0 / 256.0, 0.9], np.float32), vm, K, w, h)
    assert (r1[0][0] > 0).all() and (r1[0][1] == 0).all()


def test_ut_general_path_equals_pinhole_path_for_a_plain_camera(lib):
    """ut_project_general with no distortion and a global shutter is the same algorithm as the pinhole fast path."""
    n, w, h = 3000, 320, 240
    sc = S.make_scene(n, 1, w, h, 0, seed=13)
    means, q, s, op, _ = sc.activated()
    a = ut_pinhole(lib, means, q, s, op, sc.viewmats[0], sc.Ks[0], w, h)
    b = ut_general(lib, means, q, s, op, sc.viewmats[0], None, sc.Ks[0], w, h)
    compare_projection(b, a, 300)
    # ... and an all-zero distortion is the plain camera
    c = ut_general(lib, means, q, s, op, sc.viewmats[0], None, sc.Ks[0], w, h, radial=[0] * 6, tangential=[0, 0], prism=[0] * 4)
    compare_projection(c, a, 300)
    # ... and a rolling shutter whose two poses coincide is the global shutter
    d = ut_general(lib, means, q, s, op, sc.viewmats[0], sc.viewmats[0], sc.Ks[0], w, h, shutter=0)
    compare_projection(d, a, 300)


def _ut_numpy64(project, mean, R, scale, ut=(0.1, 2.0, 0.0), eps2d=0.3):
    """float64 unscented transform of one Gaussian through an arbitrary point projection (ProjectionUT3DGSFused.cu:80-140)."""
    alpha, beta, kappa = ut
    D = 3.0
    lam = alpha * alpha * (D + kappa) - D
    pts = [mean] + [mean + np.sqrt(D + lam) * scale[i] * R[:, i] for i in range(3)] + [mean - np.sqrt(D + lam) * scale[i] * R[:, i]
                                                                                         for i in range(3)]
    uv = np.array([project(p) for p in pts])
    wm = np.array([lam / (D + lam)] + [1 / (2 * (D + lam))] * 6)
    wc = wm.copy()
    wc[0] += 1 - alpha * alpha + beta
    mu = (wm[:, None] * uv).sum(0)
    d = uv - mu
    cov = (wc[:, None, None] * d[:, :, None] * d[:, None, :]).sum(0) + eps2d * np.eye(2)
    return mu, np.linalg.inv(cov)


def test_ut_general_distorted_and_fisheye_against_float64(lib):
    from test_host_cameras import FISH_K, RADIAL, TANGENTIAL, PRISM, fisheye_project64, opencv_project64, quat_mat64

    n, w, h = 1200, 320, 240
    sc = S.make_scene(n, 1, w, h, 0, seed=17, sigma_px=2.5)
    means, q, s, op, _ = sc.activated()
    vm, K = sc.viewmats[0].astype(np.float64), sc.Ks[0].astype(np.float64)
    for kw, proj in ((dict(radial=RADIAL, tangential=TANGENTIAL, prism=PRISM),
                      lambda pc: opencv_project64(pc[None], K, RADIAL, TANGENTIAL, PRISM)[0][0]),
                     (dict(model=FISHEYE, radial=FISH_K), lambda pc: fisheye_project64(pc[None], K, FISH_K)[0][0])):
        radii, m2d, dep, con, comp = ut_general(lib, means, q, s, op, sc.viewmats[0], None, sc.Ks[0], w, h, **kw)
        vis = np.nonzero((radii > 0).all(-1))[0]
        assert len(vis) > 200
        worst_m, worst_c = 0.0, 0.0
        for i in vis[:300]:
            mu, conic = _ut_numpy64(lambda p: proj(vm[:3, :3] @ p + vm[:3, 3]), means[i].astype(np.float64),
                                    quat_mat64(q[i].astype(np.float64)), s[i].astype(np.float64))
            worst_m = max(worst_m, np.abs(mu - m2d[i]).max())
            worst_c = max(worst_c, np.abs(np.array([conic[0, 0], conic[0, 1], conic[1, 1]]) - con[i]).max() / np.abs(conic).max())
        assert worst_m < 2e-2, worst_m   # pixels; the +-99 sigma-point weights amplify fp32 rounding of the projections
        assert worst_c < 2e-2, worst_c


# ---- spherical harmonics ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("degree", [0, 1, 2, 3, 4])
def test_spherical_harmonics_match_oracle(lib, degree):
    rng = np.random.RandomState(3 + degree)
    n = 3000
    dirs = rng.normal(size=(n, 3)).astype(np.float32) * rng.uniform(0.1, 10, size=(n, 1)).astype(np.float32)  # not unit
    for K in sorted({(degree + 1) ** 2, 16 if degree <= 3 else 25}):
        coeffs = rng.normal(size=(n, K, 3)).astype(np.float32)
        colors = np.zeros((n, 3), np.float32)
        lib.hd_sh_fwd(C.c_int(degree), C.c_int(n), C.c_int(K), _p(dirs), _p(coeffs), _p(colors))
        assert relerr(colors, O.sh_fwd(degree, dirs, coeffs)) <= 1e-5
        v_colors = rng.normal(size=(n, 3)).astype(np.float32)
        v_coeffs = np.zeros((n, K, 3), np.float32)
        v_dirs = np.zeros((n, 3), np.float32)
        lib.hd_sh_bwd(C.c_int(degree), C.c_int(n), C.c_int(K), _p(dirs), _p(coeffs), _p(v_colors), _p(v_coeffs), _p(v_dirs))
        o_vc, o_vd = O.sh_bwd(degree, dirs, coeffs, v_colors)
        assert relerr(v_coeffs, o_vc) <= 1e-5
        assert (v_coeffs[:, (degree + 1) ** 2:] == 0).all()  # bands above the degree in use get no gradient
        if degree >= 1:
            assert relerr(v_dirs, o_vd) <= 2e-5
        else:
            assert (v_dirs == 0).all()
        # v_dirs not requested: the coefficient gradients do not change
        v2 = np.zeros((n, K, 3), np.float32)
        lib.hd_sh_bwd(C.c_int(degree), C.c_int(n), C.c_int(K), _p(dirs), _p(coeffs), _p(v_colors), _p(v2), None)
        assert (v2 == v_coeffs).all()


def test_spherical_harmonics_direction_gradient_is_tangential(lib):
    """colour depends on dir / |dir| only, so dL/d(dir) is orthogonal to dir, and finite differences agree."""
    rng = np.random.RandomState(7)
    n, degree, K = 500, 3, 16
    dirs = rng.normal(size=(n, 3)).astype(np.float32)
    coeffs = rng.normal(size=(n, K, 3)).astype(np.float32)
    v_colors = rng.normal(size=(n, 3)).astype(np.float32)
    v_coeffs, v_dirs = np.zeros((n, K, 3), np.float32), np.zeros((n, 3), np.float32)
    lib.hd_sh_bwd(C.c_int(degree), C.c_int(n), C.c_int(K), _p(dirs), _p(coeffs), _p(v_colors), _p(v_coeffs), _p(v_dirs))
    along = np.abs((v_dirs * dirs).sum(-1)) / (np.linalg.norm(v_dirs, axis=-1) * np.linalg.norm(dirs, axis=-1))
    assert along.max() < 1e-5
    eps = 1e-3
    fd = np.zeros((n, 3))
    for a in range(3):
        dp, dm = dirs.astype(np.float64).copy(), dirs.astype(np.float64).copy()
        dp[:, a] += eps
        dm[:, a] -= eps
        fd[:, a] = ((O.sh_fwd(degree, dp, coeffs) - O.sh_fwd(degree, dm, coeffs)) * v_colors).sum(-1) / (2 * eps)
    assert relerr(v_dirs, fd) < 1e-4


# ---- tile rectangle, key widths -----------------------------------------------------------------------------------------------
def test_tile_rect_counts_match_the_oracle(lib):
    rng = np.random.RandomState(1)
    n, w, h = 20000, 1920, 1080
    tw, th = (w + 15) // 16, (h + 15) // 16
    m2d = np.stack([rng.uniform(-200, w + 200, n), rng.uniform(-200, h + 200, n)], -1).astype(np.float32)
    radii = rng.randint(0, 90, size=(n, 2)).astype(np.int32)
    radii[rng.uniform(size=n) < 0.1] = 0
    # extremes: far outside on every side, radius larger than the image, exactly on tile boundaries
    m2d[:6] = [[-1e6, 5], [1e6, 5], [5, -1e6], [5, 1e6], [960, 540], [16.0, 32.0]]
    radii[:6] = [[10, 10], [10, 10], [10, 10], [10, 10], [5000, 5000], [16, 16]]
    rects = np.zeros((n, 4), np.uint32)
    lib.hd_tile_rect(C.c_int(n), _p(m2d), _p(radii.astype(np.float32)), C.c_float(16.0), C.c_uint32(tw), C.c_uint32(th), _p(rects))
    area = ((rects[:, 2] - rects[:, 0]) * (rects[:, 3] - rects[:, 1])).astype(np.int64)
    area[(radii <= 0).any(-1)] = 0  # the kernels skip culled Gaussians before the rectangle (radii == 0)
    tpg, ids, flat = O.intersect_tile(m2d[None], radii[None], np.ones((1, n), np.float32), 16, tw, th, sort=False)
    assert (area == tpg[0]).all()
    assert (rects[:, 2] <= tw).all() and (rects[:, 3] <= th).all() and (rects[:, 0] <= rects[:, 2]).all()
    assert area[4] == tw * th and area[:4].sum() == 0
    assert list(rects[5]) == [0, 1, 2, 3]
    # the oracle's instances of each Gaussian are exactly the tiles of its rectangle: same count, all inside, none twice
    assert len(flat) == area.sum()
    tile = (ids >> 32).astype(np.int64)  # one camera: no camera bits above the tile id (IntersectTile.cu:95-108)
    tx, ty = tile % tw, tile // tw
    r = rects[flat].astype(np.int64)
    assert ((r[:, 0] <= tx) & (tx < r[:, 2]) & (r[:, 1] <= ty) & (ty < r[:, 3])).all()
    assert len(np.unique(tile * n + flat)) == len(flat)


def test_tile_key_widths(lib):
    for n_tiles in (1, 2, 3, 4, 255, 256, 257, 8160, 8161, 65535, 65536, 1 << 20):
        b = lib.hd_tile_key_bits(C.c_uint32(n_tiles))
        assert (1 << b) >= n_tiles and (b == 1 or (1 << (b - 1)) < n_tiles)
        assert lib.hd_ref_tile_n_bits(C.c_uint32(n_tiles)) == int(np.floor(np.log2(n_tiles))) + 1  # IntersectTile.cu:150


# ---- exact tile culling -------------------------------------------------------------------------------------------------------
def test_cull_row_span_never_drops_a_contributing_tile(lib):
    """Q(u, v) = a u^2 + 2 b u v + c v^2 - lim <= 0 is the region where a pixel centre can reach alpha >= 1/255.  Whatever
    cull_row_span removes from the AABB row must hold no such pixel centre (lossless), and what it keeps must be tight."""
    rng = np.random.RandomState(21)
    tw, th = 40, 30
    kept_total, kept_needed, dropped = 0, 0, 0
    for trial in range(400):
        ang = rng.uniform(0, np.pi)
        l1, l2 = rng.uniform(3, 120), rng.uniform(3, 120)  # semi-axes in pixels
        Rm = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        A = Rm @ np.diag([1 / l1 ** 2, 1 / l2 ** 2]) @ Rm.T
        scale = rng.uniform(0.5, 50.0)  # the record is scale-free: Q and lim share it
        a, b, c, lim = A[0, 0] * scale, A[0, 1] * scale, A[1, 1] * scale, scale
        xc, yc = rng.uniform(-40, tw * 16 + 40), rng.uniform(-40, th * 16 + 40)
        x0, x1, y0, y1 = 0, tw, 0, th
        first, last = np.zeros(th, np.int32), np.zeros(th, np.int32)
        lib.hd_cull_row_spans(C.c_float(a), C.c_float(b), C.c_float(c), C.c_float(xc), C.c_float(yc), C.c_float(lim),
                              C.c_uint32(x0), C.c_uint32(x1), C.c_uint32(y0), C.c_uint32(y1), _p(first), _p(last))
        px = np.arange(tw * 16) + 0.5 - xc
        py = np.arange(th * 16) + 0.5 - yc
        Q = a * px[None, :] ** 2 + 2 * b * px[None, :] * py[:, None] + c * py[:, None] ** 2 - lim
        need = (Q <= 0).reshape(th, 16, tw, 16).any(axis=(1, 3))  # [th, tw]: tile holds a contributing pixel centre
        keep = np.zeros((th, tw), bool)
        for ty in range(th):
            if last[ty] >= first[ty]:
                keep[ty, first[ty]:last[ty] + 1] = True
        assert not (need & ~keep).any(), (trial, a, b, c, xc, yc, lim)
        kept_total += keep.sum()
        kept_needed += (keep & need).sum()
        dropped += (~keep).sum()
    assert dropped > 0.5 * 400 * tw * th * 0.5          # it does cull
    assert kept_needed / kept_total > 0.97, kept_needed / kept_total  # and what it keeps is nearly all needed


def test_cull_row_span_keeps_everything_for_an_unbounded_region(lib):
    first, last = np.zeros(5, np.int32), np.zeros(5, np.int32)
    lib.hd_cull_row_spans(C.c_float(1), C.c_float(0), C.c_float(1), C.c_float(0), C.c_float(0), C.c_float(np.inf), C.c_uint32(3),
                          C.c_uint32(17), C.c_uint32(2), C.c_uint32(7), _p(first), _p(last))
    assert (first == 3).all() and (last == 16).all()
