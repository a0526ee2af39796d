"""GPU parity of the legacy 2-D op surface (csrc/legacy2d.cu) through the C ABI (ops.py ->
liblfs_b200.so) against the CPU oracle (oracle/lfs_oracle_legacy2d_impl.h) and the torch_impl golden vectors, on the fixtures
of the reference's own gtests (tests/test_basic.cpp:40-128,280-372, tests/test_gsplat_ops.cpp:60-96,150-300).
Tolerances: forward 1e-4, backward 1e-3 (north_star), strict metric of test_gpu_baseline_configs.py."""
import numpy as np
import pytest
import torch

import oracle as O
from test_gpu_baseline_configs import gate, gate_render

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def T(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV).to(dtype).contiguous()


def N_(t):
    return t.detach().cpu().numpy()


def test_quat_scale_to_covar_preci(golden):
    from lichtfeld_studio_b200 import ops
    rep = {}
    q, s = golden["qs_quats"], golden["qs_scales"]
    cov, pre = ops.quat_scale_to_covar_preci_fwd(T(q), T(s), True, True, False)
    gate(rep, "covars_vs_torch_impl", N_(cov), golden["qs_covars"], 1e-4, 1.0)
    gate(rep, "precis_vs_torch_impl", N_(pre), golden["qs_precis"], 1e-4, 1.0)
    # the fixture of tests/test_basic.cpp:40-95: N = 100, randn quats, rand scales * 0.1
    rng = np.random.default_rng(42)
    n = 100
    q = rng.normal(size=(n, 4)).astype(np.float32)
    s = (rng.uniform(size=(n, 3)) * 0.1 + 1e-3).astype(np.float32)
    for triu in (False, True):
        cov, pre = ops.quat_scale_to_covar_preci_fwd(T(q), T(s), True, True, triu)
        assert tuple(cov.shape) == ((n, 6) if triu else (n, 3, 3))
        oc, op_ = O.quat_scale_to_covar_preci_fwd(q, s, True, True, triu)
        gate(rep, f"covars_triu{int(triu)}", N_(cov), oc, 1e-4, 1.0)
        gate(rep, f"precis_triu{int(triu)}", N_(pre), op_, 1e-4, 1.0)
        vc = rng.normal(size=oc.shape).astype(np.float32)
        vp = (rng.normal(size=oc.shape) * 0.01 * s.min() ** 2).astype(np.float32)
        vq, vs = ops.quat_scale_to_covar_preci_bwd(T(q), T(s), triu, T(vc), T(vp))
        oq, os_ = O.quat_scale_to_covar_preci_bwd(q, s, triu, vc, vp)
        gate(rep, f"v_quats_triu{int(triu)}", N_(vq), oq, 1e-3, 0.999)
        gate(rep, f"v_scales_triu{int(triu)}", N_(vs), os_, 1e-3, 0.999)
        vq1, vs1 = ops.quat_scale_to_covar_preci_bwd(T(q), T(s), triu, T(vc), None)
        oq1, os1 = O.quat_scale_to_covar_preci_bwd(q, s, triu, vc, None)
        gate(rep, f"v_quats_covar_only_triu{int(triu)}", N_(vq1), oq1, 1e-3, 0.999)
        gate(rep, f"v_scales_covar_only_triu{int(triu)}", N_(vs1), os1, 1e-3, 0.999)
    only_c, none_p = ops.quat_scale_to_covar_preci_fwd(T(q), T(s), True, False, False)
    assert none_p.numel() == 0 and only_c.shape == (n, 3, 3)


def _projection_case(rng, N, C):
    means = rng.normal(size=(N, 3)).astype(np.float32) * 2.0
    means[:, 2] = np.abs(means[:, 2]) + 2.0
    quats = rng.normal(size=(N, 4)).astype(np.float32)
    scales = (rng.uniform(size=(N, 3)) * 0.1 + 0.005).astype(np.float32)
    op = rng.uniform(size=N).astype(np.float32)
    vm = np.tile(np.eye(4, dtype=np.float32), (C, 1, 1))
    for c in range(C):
        vm[c, 0, 3] = 0.2 * c
    K = np.tile(np.array([[300.0, 0, 320.0], [0, 300.0, 240.0], [0, 0, 1]], np.float32), (C, 1, 1))
    return means, quats, scales, op, vm, K


@pytest.mark.parametrize("with_opacities,comp", [(True, False), (True, True), (False, False)])
def test_projection_ewa(golden, with_opacities, comp):
    from lichtfeld_studio_b200 import ops
    rep = {}
    rng = np.random.default_rng(42)
    N, C, W, H = 1000, 2, 640, 480  # tests/test_gsplat_ops.cpp:150-190
    means, quats, scales, op, vm, K = _projection_case(rng, N, C)
    empty = torch.empty((0, 3, 3), device=DEV)
    got = ops.projection_ewa_3dgs_fused_fwd(T(means), empty, T(quats), T(scales), T(op) if with_opacities else None, T(vm),
                                            T(K), W, H, 0.3, 0.01, 10000.0, 0.0, comp)
    want = O.projection_ewa(means, None, quats, scales, op if with_opacities else None, vm, K, W, H, 0.3, 0.01, 1e4, 0.0, comp)
    radii, wr = N_(got[0]), want[0]
    vis = (wr > 0).all(-1)
    assert vis.sum() > N // 4
    both = vis & (radii > 0).all(-1)
    # a radius is ceil() of an fp32 product: the oracle (double) may round the other way on a handful of entries
    assert (radii != wr).any(-1).sum() <= max(2, N * C // 500), int((radii != wr).any(-1).sum())
    gate(rep, "means2d", N_(got[1])[both], want[1][both], 1e-4, 0.999)
    gate(rep, "depths", N_(got[2])[both], want[2][both], 1e-4, 1.0)
    gate(rep, "conics", N_(got[3])[both], want[3][both], 1e-4, 0.995)
    if comp:
        gate(rep, "compensations", N_(got[4])[both], want[4][both], 1e-4, 0.999)
    else:
        assert got[4].numel() == 0
    # explicit covariances
    cov, _ = ops.quat_scale_to_covar_preci_fwd(T(quats), T(scales), True, False, False)
    got2 = ops.projection_ewa_3dgs_fused_fwd(T(means), cov, T(quats), T(scales), T(op) if with_opacities else None, T(vm),
                                             T(K), W, H, 0.3, 0.01, 10000.0, 0.0, comp)
    assert (N_(got2[0]) != radii).any(-1).sum() <= 2
    gate(rep, "conics_from_covars", N_(got2[3])[both], want[3][both], 1e-4, 0.995)
    # torch_impl golden (no opacities)
    gN, gW, gH = [int(x) for x in golden["ewa_geom"]]
    g = ops.projection_ewa_3dgs_fused_fwd(T(golden["ewa_means"]), empty, T(golden["ewa_quats"]), T(golden["ewa_scales"]),
                                          None, T(golden["ewa_viewmat"]), T(golden["ewa_K"]), gW, gH, 0.3, 0.01, 1e4, 0.0,
                                          False)
    np.testing.assert_array_equal(N_(g[0]), golden["ewa_radii"])
    gv = (golden["ewa_radii"] > 0).all(-1)
    gate(rep, "golden_means2d", N_(g[1])[gv], golden["ewa_means2d"][gv], 1e-4, 1.0)
    gate(rep, "golden_conics", N_(g[3])[gv], golden["ewa_conics"][gv], 1e-4, 1.0)


def test_projection_ewa_refuses_other_camera_models():
    from lichtfeld_studio_b200 import _lib, ops
    rng = np.random.default_rng(0)
    means, quats, scales, op, vm, K = _projection_case(rng, 8, 1)
    with pytest.raises(_lib.LfsUnsupported):
        ops.projection_ewa_3dgs_fused_fwd(T(means), None, T(quats), T(scales), T(op), T(vm), T(K), 64, 64, 0.3, 0.01, 1e4,
                                          0.0, False, camera_model=_lib.FISHEYE)


def _pipeline(rng, N, W, H, CH, C=1):
    """tests/test_gsplat_ops.cpp:210-300 RasterizationPipelineTest: project -> intersect -> offsets -> rasterize"""
    from lichtfeld_studio_b200 import ops
    means = rng.normal(size=(N, 3)).astype(np.float32) * 2.0
    means[:, 2] = np.abs(means[:, 2]) + 1.5
    quats = rng.normal(size=(N, 4)).astype(np.float32)
    quats /= np.linalg.norm(quats, axis=-1, keepdims=True)
    scales = (rng.uniform(size=(N, 3)) * 0.5 + 0.01).astype(np.float32)
    op = rng.uniform(size=N).astype(np.float32)
    colors = rng.uniform(size=(C, N, CH)).astype(np.float32)
    vm = np.tile(np.eye(4, dtype=np.float32), (C, 1, 1))
    K = np.tile(np.array([[200.0, 0, W / 2], [0, 200.0, H / 2], [0, 0, 1]], np.float32), (C, 1, 1))
    radii, m2d, dep, con, _ = ops.projection_ewa_3dgs_fused_fwd(T(means), None, T(quats), T(scales), T(op), T(vm), T(K), W, H,
                                                                0.3, 0.01, 1000.0, 0.0, False)
    tw, th = (W + 15) // 16, (H + 15) // 16
    _, ids, flat = ops.intersect_tile(m2d, radii, dep, None, None, C, 16, tw, th, True)
    off = ops.intersect_offset(ids, C, tw, th)
    opac = T(np.tile(op, (C, 1)))
    return m2d, con, T(colors), opac, off, flat


@pytest.mark.parametrize("CH,with_bg,W,H", [(3, True, 256, 256), (1, False, 200, 120), (4, True, 100, 70), (7, True, 64, 64),
                                            (17, False, 48, 48)])
def test_rasterize_to_pixels_3dgs(CH, with_bg, W, H):
    from lichtfeld_studio_b200 import ops
    rep = {}
    rng = np.random.default_rng(42 + CH)
    C = 2 if CH == 4 else 1
    m2d, con, colors, opac, off, flat = _pipeline(rng, 100 if W < 200 else 400, W, H, CH, C)
    bg = T(rng.uniform(size=(C, CH)).astype(np.float32)) if with_bg else None
    r, al, last = ops.rasterize_to_pixels_3dgs_fwd(m2d, con, colors, opac, bg, None, W, H, 16, off, flat)
    assert tuple(r.shape) == (C, H, W, CH) and tuple(al.shape) == (C, H, W, 1)
    a = [N_(x) for x in (m2d, con, colors, opac)]
    nbg = None if bg is None else N_(bg)
    wr, wa, wl = O.raster_2d_fwd(*a, nbg, None, W, H, 16, N_(off), N_(flat), prec=32)
    assert float(wa.max()) > 0.5
    gate_render(rep, "renders", N_(r), wr)
    gate_render(rep, "alphas", N_(al), wa)
    assert (N_(last) != wl).mean() <= 2e-3  # __expf vs exp at the 1/255 and 1e-4 thresholds
    # backward on the forward's own state
    v_rc = T(rng.normal(size=(C, H, W, CH)).astype(np.float32))
    v_ra = T(rng.normal(size=(C, H, W, 1)).astype(np.float32))
    g = ops.rasterize_to_pixels_3dgs_bwd(m2d, con, colors, opac, bg, None, W, H, 16, off, flat, al, last, v_rc, v_ra,
                                         absgrad=True)
    w = O.raster_2d_bwd(*a, nbg, None, W, H, 16, N_(off), N_(flat), N_(al), N_(last), N_(v_rc), N_(v_ra), absgrad=True,
                        prec=64)
    for name, got, want in zip(("v_means2d_abs", "v_means2d", "v_conics", "v_colors", "v_opacities"), g,
                               (w[1], w[0], w[2], w[3], w[4])):
        gate(rep, name, N_(got), want, 1e-3, 0.995)
    g2 = ops.rasterize_to_pixels_3dgs_bwd(m2d, con, colors, opac, bg, None, W, H, 16, off, flat, al, last, v_rc, v_ra)
    assert g2[0].numel() == 0
    gate(rep, "v_means2d_noabs", N_(g2[1]), w[0], 1e-3, 0.995)
    # values in range for colours in [0,1] and a background in [0,1] (tests/test_basic.cpp:364-371)
    assert float(al.min()) >= 0 and float(al.max()) <= 1 and float(r.min()) >= -1e-6 and float(r.max()) <= 1 + 1e-5


def test_rasterize_to_pixels_3dgs_masks_empty_and_errors():
    from lichtfeld_studio_b200 import _lib, ops
    rng = np.random.default_rng(3)
    W, H, CH = 80, 48, 3
    m2d, con, colors, opac, off, flat = _pipeline(rng, 120, W, H, CH)
    bg = T(np.array([[0.2, 0.4, 0.6]], np.float32))
    masks = torch.ones((1, 3, 5), dtype=torch.bool, device=DEV)
    masks[0, 1, 2] = False
    r, al, last = ops.rasterize_to_pixels_3dgs_fwd(m2d, con, colors, opac, bg, masks, W, H, 16, off, flat)
    blk = r[0, 16:32, 32:48]
    assert torch.allclose(blk, bg[0].expand_as(blk)) and not al[0, 16:32, 32:48].any()
    a = [N_(x) for x in (m2d, con, colors, opac)]
    wr, wa, wl = O.raster_2d_fwd(*a, N_(bg), N_(masks).astype(np.uint8), W, H, 16, N_(off), N_(flat), prec=32)
    assert np.abs(N_(r) - wr).max() <= 1e-4
    v_rc, v_ra = torch.ones_like(r), torch.ones_like(al)
    g = ops.rasterize_to_pixels_3dgs_bwd(m2d, con, colors, opac, bg, masks, W, H, 16, off, flat, al, last, v_rc, v_ra)
    w = O.raster_2d_bwd(*a, N_(bg), N_(masks).astype(np.uint8), W, H, 16, N_(off), N_(flat), N_(al), N_(last), N_(v_rc),
                        N_(v_ra))
    assert np.abs(N_(g[3]) - w[3]).max() <= 1e-3 * np.abs(w[3]).max()
    # no intersections
    e_off = torch.zeros((1, 3, 5), dtype=torch.int32, device=DEV)
    e_flat = torch.zeros(0, dtype=torch.int32, device=DEV)
    r, al, last = ops.rasterize_to_pixels_3dgs_fwd(m2d, con, colors, opac, bg, None, W, H, 16, e_off, e_flat)
    assert torch.allclose(r, bg[0].expand_as(r)) and not al.any()
    g = ops.rasterize_to_pixels_3dgs_bwd(m2d, con, colors, opac, bg, None, W, H, 16, e_off, e_flat, al, last, v_rc, v_ra)
    assert all(not x.any() for x in g[1:])
    with pytest.raises(_lib.LfsUnsupported):
        ops.rasterize_to_pixels_3dgs_fwd(m2d, con, colors, opac, bg, None, W, H, 8, off, flat)
    with pytest.raises(ValueError):
        ops.rasterize_to_pixels_3dgs_fwd(m2d.cpu(), con, colors, opac, bg, None, W, H, 16, off, flat)


def test_legacy_ops_through_the_host_layer():
    """The same ops through libgsplat_backend_b200.so (host/gsplat_legacy_backend.cpp, declared in gsplat_legacy_ops.h with the
    argument lists of the reference's gtest call sites): identical results to the ctypes path, i.e. the boundary executes."""
    import ref_libs as R
    from lichtfeld_studio_b200 import ops
    b200 = R.fastgs_torch_module("b200")
    rng = np.random.default_rng(9)
    N, C, W, H = 300, 1, 128, 96
    means, quats, scales, op, vm, K = _projection_case(rng, N, C)
    K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    tq, ts = T(quats), T(scales)
    c1, p1 = b200.quat_scale_to_covar_preci_fwd(tq, ts, True, True, True)
    c2, p2 = ops.quat_scale_to_covar_preci_fwd(tq, ts, True, True, True)
    assert torch.equal(c1, c2) and torch.equal(p1, p2)
    vq1, vs1 = b200.quat_scale_to_covar_preci_bwd(tq, ts, True, torch.ones_like(c1), None)
    vq2, vs2 = ops.quat_scale_to_covar_preci_bwd(tq, ts, True, torch.ones_like(c1), None)
    assert torch.equal(vq1, vq2) and torch.equal(vs1, vs2)
    empty = torch.empty((0, 3, 3), device=DEV)  # the gtests pass an empty covars tensor (tests/test_basic.cpp:110-116)
    a = b200.projection_ewa_3dgs_fused_fwd(T(means), empty, tq, ts, T(op), T(vm), T(K), W, H, 0.3, 0.01, 1e4, 0.0, False)
    b = ops.projection_ewa_3dgs_fused_fwd(T(means), empty, tq, ts, T(op), T(vm), T(K), W, H, 0.3, 0.01, 1e4, 0.0, False)
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    radii, m2d, dep, con = a[:4]
    tw, th = (W + 15) // 16, (H + 15) // 16
    _, ids, flat = ops.intersect_tile(m2d, radii, dep, None, None, C, 16, tw, th, True)
    off = ops.intersect_offset(ids, C, tw, th)
    colors, opac = T(rng.uniform(size=(C, N, 3)).astype(np.float32)), T(op[None])
    bg = torch.zeros((1, 3), device=DEV)
    no_masks = torch.empty(0, dtype=torch.bool, device=DEV)  # tests/test_basic.cpp:344
    f1 = b200.rasterize_to_pixels_3dgs_fwd(m2d, con, colors, opac, bg, no_masks, W, H, 16, off, flat)
    f2 = ops.rasterize_to_pixels_3dgs_fwd(m2d, con, colors, opac, bg, None, W, H, 16, off, flat)
    assert all(torch.equal(x, y) for x, y in zip(f1, f2)) and float(f1[1].max()) > 0.3
    v_rc, v_ra = torch.ones_like(f1[0]), torch.zeros_like(f1[1])
    g1 = b200.rasterize_to_pixels_3dgs_bwd(m2d, con, colors, opac, bg, no_masks, W, H, 16, off, flat, f1[1], f1[2], v_rc, v_ra,
                                           False)
    g2 = ops.rasterize_to_pixels_3dgs_bwd(m2d, con, colors, opac, bg, None, W, H, 16, off, flat, f1[1], f1[2], v_rc, v_ra)
    for x, y in zip(g1[1:], g2[1:]):  # float atomics: same values up to the order of the adds
        assert float((x - y).abs().max()) <= 1e-5 * max(float(y.abs().max()), 1e-30)
