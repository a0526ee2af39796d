"""Generates tests/golden/torch_impl_golden.npz by running the UNMODIFIED reference CPU restatement
/root/reference/tests/torch_impl.cpp (compiled in place into oracle/_ref/libtorch_impl_ref.so by
`make -C oracle torch_impl_ref`) on seeded inputs.  Run in the build container only (needs /root/reference);
the .npz is committed so the GPU box and CI never touch the reference tree.

Fixtures follow the reference's own tests: seed 42, N=100, 64x64, tile 16 (tests/test_rasterization.cpp:127-353),
SH degrees 0..4 (tests/test_numerical_gradients.cpp:186-225), C=3 N=5 32x32 (tests/test_intersect_debug.cpp:62-107).
"""
import ctypes as C
import os

import numpy as np
import torch  # noqa: F401  (loads libtorch so the reference .so resolves its symbols)

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtorch_impl_ref.so"))
lib.ref_ti_isect_tiles.restype = C.c_longlong


def p(a):
    return a.ctypes.data_as(C.c_void_p)


out = {}
rng = np.random.RandomState(42)

# ---- spherical harmonics (reference::spherical_harmonics, torch_impl.cpp:296-321)
n = 64
dirs = rng.normal(size=(n, 3)).astype(np.float32)
for deg in range(5):
    K = (deg + 1) ** 2
    coeffs = rng.normal(size=(n, K, 3)).astype(np.float32)
    colors = np.zeros((n, 3), np.float32)
    lib.ref_ti_spherical_harmonics(C.c_int(deg), p(dirs), p(coeffs), C.c_int(n), C.c_int(K), p(colors))
    out[f"sh_coeffs_{deg}"] = coeffs
    out[f"sh_colors_{deg}"] = colors
out["sh_dirs"] = dirs
# K larger than the active degree (K=16, degree 1): inactive bases must be ignored
coeffs = rng.normal(size=(n, 16, 3)).astype(np.float32)
colors = np.zeros((n, 3), np.float32)
lib.ref_ti_spherical_harmonics(C.c_int(1), p(dirs), p(coeffs), C.c_int(n), C.c_int(16), p(colors))
out["sh_coeffs_k16_d1"], out["sh_colors_k16_d1"] = coeffs, colors

# ---- tile intersection (reference::isect_tiles, torch_impl.cpp:324-419)
def isect_case(tag, Cc, N, W, H, tile, rmax):
    tw, th = (W + tile - 1) // tile, (H + tile - 1) // tile
    means2d = (rng.uniform(-0.2, 1.2, size=(Cc, N, 2)) * np.array([W, H])).astype(np.float32)
    radii = rng.randint(0, rmax, size=(Cc, N, 2)).astype(np.int32)
    depths = rng.uniform(0.1, 10.0, size=(Cc, N)).astype(np.float32)
    cap = Cc * N * tw * th
    tpg = np.zeros((Cc, N), np.int32)
    ids = np.zeros(cap, np.int64)
    flat = np.zeros(cap, np.int32)
    k = lib.ref_ti_isect_tiles(p(means2d), p(radii), p(depths), C.c_int(Cc), C.c_int(N), C.c_int(tile), C.c_int(tw),
                               C.c_int(th), C.c_int(1), p(tpg), p(ids), p(flat), C.c_longlong(cap))
    for nm, v in (("means2d", means2d), ("radii", radii), ("depths", depths), ("tpg", tpg), ("ids", ids[:k]),
                  ("flat", flat[:k])):
        out[f"isect_{tag}_{nm}"] = v
    out[f"isect_{tag}_geom"] = np.array([Cc, N, W, H, tile, tw, th], np.int64)


isect_case("a", 1, 100, 64, 64, 16, 12)    # StepByStepComparison-like
isect_case("b", 1, 400, 300, 200, 16, 40)  # ragged right/bottom tiles, Gaussians partly outside
isect_case("c", 3, 5, 32, 32, 16, 6)       # tests/test_intersect_debug.cpp:62-107 (C=3; n_tiles=4 is a power of two)
isect_case("d", 2, 50, 80, 48, 16, 10)     # C=2, n_tiles=15 (not a power of two)

# ---- quat/scale -> covariance & precision (reference::quat_scale_to_covar_preci, torch_impl.cpp:38-77)
n = 32
quats = rng.normal(size=(n, 4)).astype(np.float32)
scales = np.exp(rng.normal(size=(n, 3)) * 0.5).astype(np.float32)
cov = np.zeros((n, 3, 3), np.float32)
pre = np.zeros((n, 3, 3), np.float32)
lib.ref_ti_quat_scale_to_covar_preci(p(quats), p(scales), C.c_int(n), p(cov), p(pre))
out["qs_quats"], out["qs_scales"], out["qs_covars"], out["qs_precis"] = quats, scales, cov, pre

# ---- EWA pinhole projection (reference::fully_fused_projection, torch_impl.cpp:146-218): used to cross-check the
# unscented-transform projection on small Gaussians, where both must agree
N, W, H = 64, 128, 96
means = rng.uniform(-1, 1, size=(N, 3)).astype(np.float32)
means[:, 2] += 4.0
q = rng.normal(size=(N, 4)).astype(np.float32)
s = np.exp(rng.normal(size=(N, 3)) * 0.3 - 4.0).astype(np.float32)
cov3 = np.zeros((N, 3, 3), np.float32)
pre3 = np.zeros((N, 3, 3), np.float32)
lib.ref_ti_quat_scale_to_covar_preci(p(q), p(s), C.c_int(N), p(cov3), p(pre3))
vm = np.eye(4, dtype=np.float32)[None].copy()
Kk = np.array([[[100.0, 0, 64.0], [0, 100.0, 48.0], [0, 0, 1]]], np.float32)
radii = np.zeros((1, N, 2), np.int32)
m2d = np.zeros((1, N, 2), np.float32)
dep = np.zeros((1, N), np.float32)
con = np.zeros((1, N, 3), np.float32)
lib.ref_ti_fully_fused_projection(p(means), p(cov3), p(vm), p(Kk), C.c_int(N), C.c_int(1), C.c_int(W), C.c_int(H),
                                  C.c_float(0.3), C.c_float(0.01), C.c_float(1e4), p(radii), p(m2d), p(dep), p(con))
for nm, v in (("means", means), ("quats", q), ("scales", s), ("viewmat", vm), ("K", Kk), ("radii", radii),
              ("means2d", m2d), ("depths", dep), ("conics", con), ("geom", np.array([N, W, H], np.int64))):
    out[f"ewa_{nm}"] = v

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torch_impl_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
