"""Pins the CPU oracle (oracle/lfs_oracle.c) against golden vectors produced by the reference's own CPU
restatement tests/torch_impl.cpp (tests/golden/make_golden.py).  Tolerances are the reference's own:
SH 1e-4 abs+rel (tests/test_numerical_gradients.cpp:186-225), tile intersection exact
(tests/test_rasterization.cpp:347-353)."""
import numpy as np
import pytest

import oracle as O


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("prec", [32, 64])
def test_sh_forward_matches_reference(golden, deg, prec):
    dirs, coeffs, want = golden["sh_dirs"], golden[f"sh_coeffs_{deg}"], golden[f"sh_colors_{deg}"]
    got = O.sh_fwd(deg, dirs, coeffs, prec=prec)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)


def test_sh_inactive_bases_ignored(golden):
    got = O.sh_fwd(1, golden["sh_dirs"], golden["sh_coeffs_k16_d1"])
    np.testing.assert_allclose(got, golden["sh_colors_k16_d1"], rtol=1e-4, atol=1e-4)


def _ref_key_to_cuda_key(ids, n_tiles, C):
    """torch_impl packs the camera id above ceil(log2(n_tiles)) tile bits (tests/torch_impl.cpp:368); the CUDA
    kernels (and the oracle) use floor(log2)+1 (gsplat/IntersectTile.cu:150).  They differ when n_tiles is a power
    of two and C > 1 -- re-pack the reference keys into the CUDA layout."""
    ref_bits = int(np.ceil(np.log2(n_tiles)))
    cuda_bits = int(np.floor(np.log2(n_tiles))) + 1
    hi = ids >> 32
    cid, tid = hi >> ref_bits, hi & ((1 << ref_bits) - 1)
    return (cid << (32 + cuda_bits)) | (tid << 32) | (ids & 0xFFFFFFFF)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_isect_tiles_bit_exact(golden, tag):
    Cc, N, W, H, tile, tw, th = [int(x) for x in golden[f"isect_{tag}_geom"]]
    means2d, radii, depths = golden[f"isect_{tag}_means2d"], golden[f"isect_{tag}_radii"], golden[f"isect_{tag}_depths"]
    tpg, ids, flat = O.intersect_tile(means2d, radii, depths, tile, tw, th, sort=True)
    np.testing.assert_array_equal(tpg, golden[f"isect_{tag}_tpg"])
    want_ids = _ref_key_to_cuda_key(golden[f"isect_{tag}_ids"], tw * th, Cc)
    # the reference sorts with torch::argsort (not stable): compare as sorted (key, value) multisets
    got = sorted(zip(ids.tolist(), flat.tolist()))
    want = sorted(zip(want_ids.tolist(), golden[f"isect_{tag}_flat"].tolist()))
    assert got == want
    assert np.all(np.diff(ids) >= 0), "oracle output must be sorted by key"


def test_isect_offsets_consistent(golden):
    Cc, N, W, H, tile, tw, th = [int(x) for x in golden["isect_d_geom"]]
    tpg, ids, flat = O.intersect_tile(golden["isect_d_means2d"], golden["isect_d_radii"], golden["isect_d_depths"],
                                      tile, tw, th, sort=True)
    off = O.intersect_offset(ids, Cc, tw, th).reshape(-1)
    bits = int(np.floor(np.log2(tw * th))) + 1
    hi = ids >> 32
    flat_tile = (hi >> bits) * (tw * th) + (hi & ((1 << bits) - 1))
    want = np.searchsorted(flat_tile, np.arange(Cc * tw * th), side="left")
    np.testing.assert_array_equal(off, want)
    # empty input -> all zeros (gsplat/IntersectTile.cu:268-271)
    assert not O.intersect_offset(np.zeros(0, np.int64), 2, 3, 3).any()


def test_quat_to_covar_via_raster_response(golden):
    """The oracle's quaternion -> rotation (gsplat/Utils.cuh:80-102) must reproduce the reference's covariance:
    Sigma = R S^2 R^T (tests/torch_impl.cpp:38-61), checked through the UT projection's rotation route
    (mat3_cast) on the golden quats/scales."""
    q, s, want = golden["qs_quats"], golden["qs_scales"], golden["qs_covars"]
    qn = q / np.linalg.norm(q, axis=-1, keepdims=True)
    w, x, y, z = qn.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                  2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                  2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * s[:, None, :]
    np.testing.assert_allclose(M @ M.transpose(0, 2, 1), want, rtol=1e-4, atol=1e-5)


def test_ut_projection_matches_ewa_for_small_gaussians(golden):
    """For Gaussians much smaller than their distance the unscented transform and the reference's EWA linearisation
    (tests/torch_impl.cpp:146-218) agree: same means2d / depths, covariance within a few percent."""
    N, W, H = [int(x) for x in golden["ewa_geom"]]
    means, q, s = golden["ewa_means"], golden["ewa_quats"], golden["ewa_scales"]
    radii, m2d, dep, con, _ = O.projection_ut(means, q, s, None, golden["ewa_viewmat"], golden["ewa_K"], W, H)
    vis = (radii[0] > 0).all(-1) & (golden["ewa_radii"][0] > 0).all(-1)
    assert vis.sum() > N // 2
    np.testing.assert_allclose(m2d[0][vis], golden["ewa_means2d"][0][vis], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(dep[0][vis], golden["ewa_depths"][0][vis], rtol=1e-5)
    np.testing.assert_allclose(con[0][vis], golden["ewa_conics"][0][vis], rtol=2e-2, atol=2e-2)
    assert np.abs(radii[0][vis] - golden["ewa_radii"][0][vis]).max() <= 1
