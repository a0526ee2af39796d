"""oracle/glm_shim held to glm's documented conventions, on the CPU.

The UNMODIFIED reference gsplat kernels are compiled against oracle/glm_shim because glm itself (a vcpkg dependency of the
reference, version unpinned in its tree) is not in this image; every `-m gpu` parity test against oracle/_ref/libgsplat_ref.so
therefore leans on the shim meaning what glm means.  glm is not here to compare with, so this file checks each symbol the
reference uses (`grep -o 'glm::[a-z_0-9A-Z]*' gsplat/*`: fvec2/3, fquat, mat, fmat2/3, dot, cross, normalize, transpose,
outerProduct, inverse, rotate, quat_cast, mat3_cast, slerp, make_vec*) against the convention glm documents and numpy can
state independently: column-major storage m[col][row] with column-major scalar constructors, M * v and M * N as linear algebra,
quaternions constructed (w, x, y, z), q * v = R(q) v with R = mat3_cast(q), quat_cast the inverse of mat3_cast up to sign,
slerp along the shorter arc with glm's lerp fall-back near cos = 1 and no renormalisation.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_glm_shim.cpp")
SHIM = os.path.join(ROOT, "oracle", "glm_shim")
OUT = os.path.join(ROOT, "tests", "_build", "libhost_glm_shim.so")


@pytest.fixture(scope="module")
def lib():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    deps = [SRC, os.path.join(SHIM, "glm", "glm.hpp")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        r = subprocess.run([gxx, "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + SHIM, "-o", OUT, SRC],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    lb = C.CDLL(OUT)
    lb.gs_mix.restype = C.c_float
    lb.gs_mix.argtypes = [C.c_float, C.c_float, C.c_float]
    return lb


def f(a):
    return np.ascontiguousarray(a, np.float32)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def out(*shape):
    return np.zeros(shape, np.float32)


def rot(axis, ang):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    x, y, z = axis
    Kx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def quat(axis, ang):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])


RNG = np.random.RandomState(0)


def test_matrix_products_are_linear_algebra(lib):
    for _ in range(20):
        A, B, v = f(RNG.normal(size=(3, 3))), f(RNG.normal(size=(3, 3))), f(RNG.normal(size=3))
        o3, o33 = out(3), out(3, 3)
        lib.gs_mat3_mul_vec(p(A), p(v), p(o3))
        np.testing.assert_allclose(o3, A.astype(np.float64) @ v, rtol=1e-5, atol=1e-6)
        lib.gs_mat3_mul_mat3(p(A), p(B), p(o33))
        np.testing.assert_allclose(o33, A.astype(np.float64) @ B, rtol=1e-5, atol=2e-6)
        lib.gs_transpose3(p(A), p(o33))
        assert (o33 == A.T).all()
        # non-square: mat<3,2> (3 columns, 2 rows) is a 2x3 matrix; J S J^T is 2x2 (the EWA covariance shape)
        J, o22 = f(RNG.normal(size=(2, 3))), out(2, 2)
        S = f(B @ B.T)
        lib.gs_mat3x2_chain(p(J), p(S), p(o22))
        np.testing.assert_allclose(o22, J.astype(np.float64) @ S @ J.T, rtol=1e-5, atol=1e-5)


def test_outer_product_and_inverse(lib):
    c, r = f([1, 2, 3]), f([5, 7, 11])
    o = out(3, 3)
    lib.gs_outer3(p(c), p(r), p(o))
    assert (o == np.outer(c, r)).all()  # outerProduct(c, r) = c r^T: element [row i][col j] = c_i r_j
    for _ in range(10):
        A, o2 = f(RNG.normal(size=(2, 2)) + 2 * np.eye(2)), out(2, 2)
        lib.gs_inverse2(p(A), p(o2))
        np.testing.assert_allclose(o2, np.linalg.inv(A.astype(np.float64)), rtol=2e-5, atol=1e-6)


def test_matrix_arithmetic_and_constructors(lib):
    A, B = f(RNG.normal(size=(3, 3))), f(RNG.normal(size=(3, 3)))
    s, d, n, sc, acc = out(3, 3), out(3, 3), out(3, 3), out(3, 3), out(3, 3)
    lib.gs_mat3_arith(p(A), p(B), C.c_float(2.5), p(s), p(d), p(n), p(sc), p(acc))
    assert (s == A + B).all() and (d == A - B).all() and (n == -A).all() and (sc == np.float32(2.5) * A).all() and (acc == s).all()
    # mat3(a0..a8) takes its scalars COLUMN by column: a0 a1 a2 is the first column
    s9 = f(np.arange(1, 10))
    fs, dg, fc = out(3, 3), out(3, 3), out(3, 3)
    c0, c1, c2 = f([1, 2, 3]), f([4, 5, 6]), f([7, 8, 9])
    lib.gs_mat3_ctors(p(s9), C.c_float(3.0), p(c0), p(c1), p(c2), p(fs), p(dg), p(fc))
    want = np.array([[1, 4, 7], [2, 5, 8], [3, 6, 9]], np.float32)
    assert (fs == want).all() and (fc == want).all()
    assert (dg == 3 * np.eye(3)).all()


def test_vector_functions(lib):
    for _ in range(10):
        a, b = f(RNG.normal(size=3)), f(RNG.normal(size=3))
        cr, dt, ln, nr = out(3), out(1), out(1), out(3)
        lib.gs_vec3_ops(p(a), p(b), p(cr), p(dt), p(ln), p(nr))
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        np.testing.assert_allclose(cr, np.cross(a64, b64), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dt[0], a64 @ b64, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ln[0], np.linalg.norm(a64), rtol=1e-6)
        np.testing.assert_allclose(nr, a64 / np.linalg.norm(a64), rtol=1e-6, atol=1e-7)
    assert lib.gs_mix(2.0, 10.0, 0.25) == 4.0  # x (1 - a) + y a


def test_quaternion_is_wxyz_and_rotates_like_its_matrix(lib):
    dq = out(4)
    lib.gs_default_quat(p(dq))
    assert list(dq) == [1, 0, 0, 0]
    # a quarter turn about z, built (w, x, y, z), takes +x to +y
    q = f(quat([0, 0, 1], np.pi / 2))
    a, b = out(3), out(3)
    lib.gs_quat_rotate(p(q), p(f([1, 0, 0])), p(a), p(b))
    np.testing.assert_allclose(a, [0, 1, 0], atol=1e-6)
    assert (a == b).all()  # rotate(q, v) is q * v
    for _ in range(30):
        axis, ang = RNG.normal(size=3), RNG.uniform(-np.pi, np.pi)
        q, v = f(quat(axis, ang)), f(RNG.normal(size=3))
        M = out(3, 3)
        lib.gs_mat3_cast(p(q), p(M))
        np.testing.assert_allclose(M, rot(axis, ang), atol=3e-6)     # mat3_cast(q) is THE rotation matrix (as linear algebra)
        lib.gs_quat_rotate(p(q), p(v), p(a), p(b))
        np.testing.assert_allclose(a, rot(axis, ang) @ v.astype(np.float64), atol=5e-6)


def test_quat_cast_inverts_mat3_cast_on_every_branch(lib):
    hit = set()
    for _ in range(300):
        axis, ang = RNG.normal(size=3), RNG.uniform(-np.pi, np.pi)
        R = f(rot(axis, ang))
        q = out(4)
        lib.gs_quat_cast(p(R), p(q))
        want = quat(axis, ang)
        if np.dot(want, q) < 0:
            want = -want
        np.testing.assert_allclose(q, want, atol=3e-6)
        hit.add(int(np.argmax(np.abs(q))))  # glm picks the biggest of w, x, y, z: all four branches must occur
    assert hit == {0, 1, 2, 3}
    # the branch component comes out positive (glm's "biggestVal")
    q = out(4)
    lib.gs_quat_cast(p(f(rot([1, 0, 0], np.pi * 0.99))), p(q))
    assert np.argmax(np.abs(q)) == 1 and q[1] > 0


def test_quaternion_helpers(lib):
    q = f([0.5, -1.5, 2.0, 0.25])
    nq, cq, iq, ln = out(4), out(4), out(4), out(1)
    lib.gs_quat_misc(p(q), p(nq), p(cq), p(iq), p(ln))
    n2 = float((q.astype(np.float64) ** 2).sum())
    np.testing.assert_allclose(ln[0], np.sqrt(n2), rtol=1e-6)
    np.testing.assert_allclose(nq, q / np.sqrt(n2), rtol=1e-6)
    assert (cq == q * np.array([1, -1, -1, -1], np.float32)).all()
    np.testing.assert_allclose(iq, cq / n2, rtol=1e-6)  # inverse = conjugate / dot(q, q): NOT assumed unit
    # normalize of the zero quaternion is the identity (glm's guard), not NaN
    lib.gs_quat_misc(p(f([0, 0, 0, 0])), p(nq), p(cq), p(iq), p(ln))
    assert list(nq) == [1, 0, 0, 0]


def test_slerp(lib):
    o = out(4)
    x, y = f(quat([0.2, 1, 0.1], 0.3)), f(quat([0.1, 0.9, -0.2], 1.4))
    for a in (0.0, 0.25, 0.5, 1.0):
        lib.gs_slerp(p(x), p(y), C.c_float(a), p(o))
        x64, y64 = x.astype(np.float64), y.astype(np.float64)
        th = np.arccos(x64 @ y64)
        want = (np.sin((1 - a) * th) * x64 + np.sin(a * th) * y64) / np.sin(th)
        np.testing.assert_allclose(o, want, atol=3e-7)
        assert abs(np.linalg.norm(o) - 1) < 1e-6
    # shorter arc: slerp(x, -y) is slerp(x, y)
    o2 = out(4)
    lib.gs_slerp(p(x), p(y), C.c_float(0.3), p(o))
    lib.gs_slerp(p(x), p(f(-y)), C.c_float(0.3), p(o2))
    np.testing.assert_allclose(o, o2, atol=1e-7)
    # nearly parallel (cos > 1 - eps): component-wise mix, NOT renormalised
    z = f(x * np.float32(1.0))
    lib.gs_slerp(p(x), p(z), C.c_float(0.4), p(o))
    np.testing.assert_allclose(o, x, atol=1e-7)
    big = f(2 * x)  # not unit: dot = 2 > 1 - eps -> lerp branch, result 1.4 x
    lib.gs_slerp(p(x), p(big), C.c_float(0.4), p(o))
    np.testing.assert_allclose(o, 1.4 * x, rtol=1e-6)


def test_shim_and_product_agree_on_the_pose_route(lib):
    """cameras.cuh restates quat_cast / slerp / mat3_cast for the product; the shim restates them for the reference build.
    Two independent restatements of glm must agree to rounding (exact-sin variant; the device-only fast-math variant of the
    product is the subject of tests/test_gpu_boundary.py)."""
    import test_host_cameras as hc

    cam_lib = hc.build_lib()
    for _ in range(50):
        R0, R1 = rot(RNG.normal(size=3), RNG.uniform(-np.pi, np.pi)), rot(RNG.normal(size=3), RNG.uniform(-0.3, 0.3))
        vm0, vm1 = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
        vm0[:3, :3], vm1[:3, :3] = R0, R1 @ R0
        q_s, q_p, Rb = out(4), out(4), out(9)
        # the reference copies the row-major pose into a glm mat3 element by element (Cameras.cuh:39-56), so glm holds the
        # rotation R itself; the product's quat_from_rowmajor_rot reads m[c][r] = R[r][c] out of the same memory (common.cuh)
        lib.gs_quat_cast(p(f(vm0[:3, :3])), p(q_s))
        cam_lib.hc_quat_roundtrip(p(vm0), p(q_p), p(Rb))
        np.testing.assert_allclose(q_p, q_s, atol=1e-7)
        cam = hc.Cam(vm0, vm1, shutter=hc.RS_TB)
        qs, _ = hc.shutter_pose(cam_lib, cam, [0.0, 0.3, 1.0])
        q1_s = out(4)
        lib.gs_quat_cast(p(f(vm1[:3, :3])), p(q1_s))
        for i, a in enumerate((0.0, 0.3, 1.0)):
            o = out(4)
            lib.gs_slerp(p(q_s), p(q1_s), C.c_float(a), p(o))
            np.testing.assert_allclose(qs[i], o, atol=2e-7)
