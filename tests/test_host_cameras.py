"""Camera models of the 3DGUT path (lichtfeld-studio_b200/csrc/cameras.cuh) checked on the HOST, no GPU.

cameras.cuh is `__host__ __device__` throughout: tests/host_cameras.cu compiles the very functions the projection and ray
kernels use into a small host library, and this file checks them against float64 numpy restatements of the published models
the reference follows (gsplat/Cameras.cuh: perfect pinhole :416-470, OpenCV pinhole :473-757, OpenCV fisheye :760-1024,
shutter pose / rolling-shutter fixed point :253-413) and against the properties those models must have (project / unproject
round trips, rays through the point they came from, the fixed point of the rolling shutter).  The device-only `FAST` slerp
(sin.approx) is covered by the GPU tests; here the exact variant runs.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_cameras.cu")
HDR = os.path.join(ROOT, "lichtfeld-studio_b200", "csrc", "cameras.cuh")
OUT = os.path.join(ROOT, "tests", "_build", "libhost_cameras.so")

PINHOLE, FISHEYE = 0, 2  # LFS_PINHOLE / LFS_FISHEYE of include/lfs_b200.h (checked in test_enum_values)
RS_TB, RS_LR, RS_BT, RS_RL, RS_GLOBAL = 0, 1, 2, 3, 4
W, H = 640, 480
K = np.array([[420.0, 0, 318.5], [0, 415.0, 241.25], [0, 0, 1]], np.float32)


def build_lib():
    """Builds (if stale) and loads the host library; also used by tests/test_glm_shim.py."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    stale = (not os.path.exists(OUT)) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if stale:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        r = subprocess.run([nvcc, "-O2", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT,
                            SRC], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(OUT)


@pytest.fixture(scope="module")
def lib():
    return build_lib()


def _f(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Cam:
    """Argument pack of one camera: the CAM_ARGS of host_cameras.cu."""

    def __init__(self, vm0=None, vm1=None, model=PINHOLE, shutter=RS_GLOBAL, radial=None, tangential=None, prism=None, Kmat=K,
                 w=W, h=H):
        self.keep = [_f(np.eye(4) if vm0 is None else vm0), _f(vm1), _f(Kmat), _f(radial), _f(tangential), _f(prism)]
        self.model, self.shutter, self.w, self.h = model, shutter, w, h

    def args(self):
        vm0, vm1, Kmat, radial, tangential, prism = self.keep
        return [_p(vm0), _p(vm1), _p(Kmat), C.c_int(self.w), C.c_int(self.h), C.c_int(self.model), C.c_int(self.shutter),
                _p(radial), _p(tangential), _p(prism)]


def cam_project(lib, cam, pc, margin=0.15):
    pc = _f(pc)
    uv, ok = np.zeros((len(pc), 2), np.float32), np.zeros(len(pc), np.int32)
    lib.hc_cam_project(*cam.args(), C.c_int(len(pc)), _p(pc), C.c_float(margin), _p(uv), _p(ok))
    return uv, ok.astype(bool)


def cam_unproject(lib, cam, uv):
    uv = _f(uv)
    rays, ok = np.zeros((len(uv), 3), np.float32), np.zeros(len(uv), np.int32)
    lib.hc_cam_unproject(*cam.args(), C.c_int(len(uv)), _p(uv), _p(rays), _p(ok))
    return rays, ok.astype(bool)


def world_to_image(lib, cam, pw, margin=0.15):
    pw = _f(pw)
    uv, ok = np.zeros((len(pw), 2), np.float32), np.zeros(len(pw), np.int32)
    lib.hc_world_to_image(*cam.args(), C.c_int(len(pw)), _p(pw), C.c_float(margin), _p(uv), _p(ok))
    return uv, ok.astype(bool)


def pixel_to_world_ray(lib, cam, uv):
    uv = _f(uv)
    org, dirs, ok = np.zeros((len(uv), 3), np.float32), np.zeros((len(uv), 3), np.float32), np.zeros(len(uv), np.int32)
    lib.hc_pixel_to_world_ray(*cam.args(), C.c_int(len(uv)), _p(uv), _p(org), _p(dirs), _p(ok))
    return org, dirs, ok.astype(bool)


def shutter_time(lib, cam, uv):
    uv = _f(uv)
    t = np.zeros(len(uv), np.float32)
    lib.hc_shutter_time(*cam.args(), C.c_int(len(uv)), _p(uv), _p(t))
    return t


def shutter_pose(lib, cam, t):
    t = _f(t)
    q, tr = np.zeros((len(t), 4), np.float32), np.zeros((len(t), 3), np.float32)
    lib.hc_shutter_pose(*cam.args(), C.c_int(len(t)), _p(t), _p(q), _p(tr))
    return q, tr


# ---- float64 restatements ---------------------------------------------------------------------------------------------
def rot(axis, ang):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    x, y, z = axis
    Kx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def pose(axis, ang, t):
    m = np.eye(4)
    m[:3, :3] = rot(axis, ang)
    m[:3, 3] = t
    return m


def opencv_project64(pc, Kmat, k6, p2, s4):
    pc = np.asarray(pc, np.float64)
    x, y = pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2]
    r2 = x * x + y * y
    icd = (1 + r2 * (k6[0] + r2 * (k6[1] + r2 * k6[2]))) / (1 + r2 * (k6[3] + r2 * (k6[4] + r2 * k6[5])))
    dx = 2 * p2[0] * x * y + p2[1] * (r2 + 2 * x * x) + r2 * (s4[0] + r2 * s4[1])
    dy = p2[0] * (r2 + 2 * y * y) + 2 * p2[1] * x * y + r2 * (s4[2] + r2 * s4[3])
    u = (icd * x + dx) * Kmat[0, 0] + Kmat[0, 2]
    v = (icd * y + dy) * Kmat[1, 1] + Kmat[1, 2]
    return np.stack([u, v], -1), icd


def fisheye_project64(pc, Kmat, k4):
    pc = np.asarray(pc, np.float64)
    nrm = np.hypot(pc[:, 0], pc[:, 1])
    th = np.arctan2(nrm, pc[:, 2])
    t2 = th * th
    d = th * (1 + t2 * (k4[0] + t2 * (k4[1] + t2 * (k4[2] + t2 * k4[3])))) / np.maximum(nrm, 1e-30)
    return np.stack([Kmat[0, 0] * d * pc[:, 0] + Kmat[0, 2], Kmat[1, 1] * d * pc[:, 1] + Kmat[1, 2]], -1), th


def quat_of(R):
    """Unit quaternion (w,x,y,z) of a rotation matrix, float64 (Shepperd's branch on the trace only: the tests keep angles small)."""
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])


def slerp64(q0, q1, t):
    c = float(np.dot(q0, q1))
    if c < 0:
        q1, c = -q1, -c
    a = np.arccos(min(c, 1.0))
    if a < 1e-9:
        return (1 - t) * q0 + t * q1
    return (np.sin((1 - t) * a) * q0 + np.sin(t * a) * q1) / np.sin(a)


def quat_mat64(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def cone_points(rng, n, max_tan, zlo=1.0, zhi=8.0):
    z = rng.uniform(zlo, zhi, n)
    xy = rng.uniform(-max_tan, max_tan, (n, 2)) * z[:, None]
    return np.concatenate([xy, z[:, None]], -1)


RADIAL = [-0.12, 0.035, -0.004, 0.01, -0.002, 0.0005]
TANGENTIAL = [0.0012, -0.0008]
PRISM = [0.0006, -0.0001, -0.0004, 0.00005]
FISH_K = [-0.02, 0.004, -0.0006, 0.00003]


# ---- tests --------------------------------------------------------------------------------------------------------------
def test_enum_values():
    text = open(os.path.join(ROOT, "include", "lfs_b200.h")).read()
    import re

    vals = dict(re.findall(r"(LFS_PINHOLE|LFS_FISHEYE)\s*=\s*(\d+)", text))
    assert int(vals["LFS_PINHOLE"]) == PINHOLE and int(vals["LFS_FISHEYE"]) == FISHEYE


def test_quat_cast_roundtrip(lib):
    rng = np.random.default_rng(0)
    for _ in range(200):  # all four branches of glm::quat_cast get hit by random rotations up to pi
        R = rot(rng.normal(size=3), rng.uniform(-np.pi, np.pi))
        vm = np.eye(4, dtype=np.float32)
        vm[:3, :3] = R
        q, back = np.zeros(4, np.float32), np.zeros(9, np.float32)
        lib.hc_quat_roundtrip(_p(vm), _p(q), _p(back))
        assert abs(np.linalg.norm(q) - 1) < 2e-6
        np.testing.assert_allclose(back.reshape(3, 3), R, atol=3e-6)


def test_perfect_pinhole_project_unproject(lib):
    rng = np.random.default_rng(1)
    cam = Cam()
    pc = cone_points(rng, 4000, 0.9)
    uv, ok = cam_project(lib, cam, pc)
    ref = np.stack([pc[:, 0] / pc[:, 2] * K[0, 0] + K[0, 2], pc[:, 1] / pc[:, 2] * K[1, 1] + K[1, 2]], -1)
    np.testing.assert_allclose(uv, ref, rtol=2e-6, atol=2e-4)
    inside = (ref[:, 0] >= -0.15 * W) & (ref[:, 0] < 1.15 * W) & (ref[:, 1] >= -0.15 * H) & (ref[:, 1] < 1.15 * H)
    edge = (np.abs(ref[:, 0] + 0.15 * W) < 1e-2) | (np.abs(ref[:, 0] - 1.15 * W) < 1e-2) | (np.abs(ref[:, 1] + 0.15 * H) < 1e-2) | (
        np.abs(ref[:, 1] - 1.15 * H) < 1e-2)
    assert (ok == inside)[~edge].all() and inside.sum() > 1000 and (~inside).sum() > 100
    rays, rok = cam_unproject(lib, cam, uv)
    assert rok.all()
    np.testing.assert_allclose(rays, pc / np.linalg.norm(pc, axis=1, keepdims=True), atol=3e-6)
    # behind the camera / on the plane: invalid, uv zeroed
    uvb, okb = cam_project(lib, cam, [[0.1, 0.2, -1.0], [0.0, 0.0, 0.0]])
    assert not okb.any() and (uvb == 0).all()


def test_opencv_pinhole_distortion(lib):
    rng = np.random.default_rng(2)
    cam = Cam(radial=RADIAL, tangential=TANGENTIAL, prism=PRISM)
    pc = cone_points(rng, 4000, 0.6)
    uv, ok = cam_project(lib, cam, pc, margin=10.0)
    ref, icd = opencv_project64(pc, K, RADIAL, TANGENTIAL, PRISM)
    np.testing.assert_allclose(uv, ref, rtol=3e-6, atol=3e-4)
    clear = np.abs(icd - 0.8) > 1e-4
    assert (ok == (icd > 0.8))[clear].all() and ok.sum() > 3000
    # Newton inverse: back to the ray of the point wherever the inverse reports convergence
    rays, rok = cam_unproject(lib, cam, uv[ok])
    assert rok.mean() > 0.99
    want = pc[ok] / np.linalg.norm(pc[ok], axis=1, keepdims=True)
    np.testing.assert_allclose(rays[rok], want[rok], atol=2e-5)


def test_opencv_radial_only_and_partial_coefficients(lib):
    """Any subset of radial / tangential / thin-prism may be given (ut.cuh passes null pointers for the others)."""
    rng = np.random.default_rng(3)
    pc = cone_points(rng, 1000, 0.5)
    z6, z2, z4 = [0.0] * 6, [0.0] * 2, [0.0] * 4
    for radial, tang, prism in ((RADIAL, None, None), (None, TANGENTIAL, None), (None, None, PRISM), (RADIAL, None, PRISM)):
        uv, ok = cam_project(lib, Cam(radial=radial, tangential=tang, prism=prism), pc, margin=10.0)
        ref, _ = opencv_project64(pc, K, radial or z6, tang or z2, prism or z4)
        np.testing.assert_allclose(uv[ok], ref[ok], rtol=3e-6, atol=3e-4)
        assert ok.mean() > 0.95
    # all-zero coefficient arrays are the perfect pinhole
    uv0, _ = cam_project(lib, Cam(), pc)
    uvz, _ = cam_project(lib, Cam(radial=z6, tangential=z2, prism=z4), pc)
    np.testing.assert_allclose(uvz, uv0, atol=1e-4)


def test_opencv_flipped_distortion_is_rejected(lib):
    """A radial polynomial that folds back (icD <= 0.8) must flag the point invalid (Cameras.cuh:536-541)."""
    cam = Cam(radial=[-0.9, 0, 0, 0, 0, 0])
    pc = np.array([[0.05, 0.0, 1.0], [0.5, 0.0, 1.0], [0.8, 0.3, 1.0]])
    _, icd = opencv_project64(pc, K, [-0.9, 0, 0, 0, 0, 0], [0, 0], [0, 0, 0, 0])
    _, ok = cam_project(lib, cam, pc, margin=10.0)
    assert list(ok) == list(icd > 0.8) == [True, False, False]


def test_fisheye_limits(lib):
    out = np.zeros(2, np.float32)
    # no distortion: forward polynomial is monotonic, the limit is the image corner angle
    lib.hc_fisheye_limits(*Cam(model=FISHEYE, radial=[0, 0, 0, 0]).args(), _p(out))
    mdx, mdy = max(W - K[0, 2], K[0, 2]), max(H - K[1, 2], K[1, 2])
    rmax = np.hypot(mdx, mdy)
    corner = max(rmax / K[0, 0], rmax / K[1, 1])
    assert abs(out[0] - corner) < 1e-5
    assert abs(out[1] - corner / max(W / 2 / K[0, 0], H / 2 / K[1, 1])) < 1e-5
    # k1 < 0 alone: d/dtheta (theta + k1 theta^3) = 0 at theta = sqrt(-1 / (3 k1)); make it smaller than the corner angle
    k1 = -1.0
    lib.hc_fisheye_limits(*Cam(model=FISHEYE, radial=[k1, 0, 0, 0]).args(), _p(out))
    assert abs(out[0] - np.sqrt(-1 / (3 * k1))) < 1e-5
    # cubic branch (k3 != 0, k4 == 0) and Newton branch (k4 != 0): the limit is the first stationary point of the polynomial
    for k4 in ([-0.6, 0.05, -0.02, 0.0], [-0.6, 0.05, -0.02, 0.004], [-0.3, -0.05, 0.0, 0.0]):
        lib.hc_fisheye_limits(*Cam(model=FISHEYE, radial=k4).args(), _p(out))
        th = np.linspace(1e-4, 1.6, 400001)
        t2 = th * th
        d = 1 + t2 * (3 * k4[0] + t2 * (5 * k4[1] + t2 * (7 * k4[2] + t2 * 9 * k4[3])))
        first = th[np.argmax(d <= 0)] if (d <= 0).any() else np.inf
        assert abs(out[0] - min(first, corner)) < 2e-5, (k4, out[0], first)


def test_fisheye_project_unproject(lib):
    rng = np.random.default_rng(4)
    cam = Cam(model=FISHEYE, radial=FISH_K)
    lim = np.zeros(2, np.float32)
    lib.hc_fisheye_limits(*cam.args(), _p(lim))
    pc = cone_points(rng, 4000, 1.2)
    uv, ok = cam_project(lib, cam, pc, margin=0.15)
    ref, th = fisheye_project64(pc, K, FISH_K)
    infov = th < lim[0] - 1e-5
    np.testing.assert_allclose(uv[infov], ref[infov], rtol=3e-6, atol=3e-4)
    inside = (ref[:, 0] >= -0.15 * W) & (ref[:, 0] < 1.15 * W) & (ref[:, 1] >= -0.15 * H) & (ref[:, 1] < 1.15 * H)
    near_edge = (np.abs(ref[:, 0] + 0.15 * W) < 1e-2) | (np.abs(ref[:, 0] - 1.15 * W) < 1e-2) | (
        np.abs(ref[:, 1] + 0.15 * H) < 1e-2) | (np.abs(ref[:, 1] - 1.15 * H) < 1e-2)
    assert (ok == inside)[infov & ~near_edge].all()
    # beyond the FOV limit the angle is clamped: the point lands on the limit circle, not further out
    beyond = th > lim[0] + 1e-4
    if beyond.any():
        t2 = float(lim[0]) ** 2
        rlim = lim[0] * (1 + t2 * (FISH_K[0] + t2 * (FISH_K[1] + t2 * (FISH_K[2] + t2 * FISH_K[3]))))
        rr = np.hypot((uv[beyond, 0] - K[0, 2]) / K[0, 0], (uv[beyond, 1] - K[1, 2]) / K[1, 1])
        np.testing.assert_allclose(rr, rlim, rtol=1e-5)
    # inverse (Newton on the forward polynomial)
    sel = infov & ok
    rays, rok = cam_unproject(lib, cam, uv[sel])
    assert rok.all()
    np.testing.assert_allclose(rays, (pc / np.linalg.norm(pc, axis=1, keepdims=True))[sel], atol=1e-5)
    # the principal point maps to the optical axis; a pixel far outside the FOV circle is rejected with the axis as ray
    rays, rok = cam_unproject(lib, cam, [[K[0, 2], K[1, 2]], [K[0, 2] + 50 * W, K[1, 2]]])
    assert rok[0] and not rok[1]
    np.testing.assert_allclose(rays, [[0, 0, 1], [0, 0, 1]], atol=1e-7)


def test_fisheye_point_on_axis(lib):
    cam = Cam(model=FISHEYE, radial=FISH_K)
    uv, ok = cam_project(lib, cam, [[0.0, 0.0, 2.0]])
    assert ok[0]
    np.testing.assert_allclose(uv[0], [K[0, 2], K[1, 2]], atol=1e-4)


def test_shutter_time(lib):
    uv = np.array([[0.5, 0.5], [W - 0.5, H - 0.5], [100.5, 200.5], [320.0, 240.0]], np.float32)
    fl, ce = np.floor(uv), np.ceil(uv)
    want = {RS_TB: fl[:, 1] / (H - 1), RS_LR: fl[:, 0] / (W - 1), RS_BT: (H - ce[:, 1]) / (H - 1), RS_RL: (W - ce[:, 0]) / (W - 1),
            RS_GLOBAL: np.zeros(len(uv))}
    for sh, t in want.items():
        got = shutter_time(lib, Cam(shutter=sh), uv)
        np.testing.assert_allclose(got, t, atol=1e-7)
    # first / last row of a top-to-bottom shutter are exactly the two poses
    t = shutter_time(lib, Cam(shutter=RS_TB), [[3.5, 0.5], [3.5, H - 0.5]])
    assert t[0] == 0.0 and t[1] == 1.0


def test_shutter_pose_is_slerp_and_lerp(lib):
    vm0 = pose([0.2, 1.0, 0.1], 0.05, [0.1, -0.2, 0.3])
    vm1 = pose([0.25, 0.9, 0.0], 0.09, [0.16, -0.17, 0.36])
    cam = Cam(vm0, vm1, shutter=RS_TB)
    ts = np.linspace(0, 1, 11)
    q, tr = shutter_pose(lib, cam, ts)
    q0, q1 = quat_of(vm0[:3, :3]), quat_of(vm1[:3, :3])
    for i, t in enumerate(ts):
        np.testing.assert_allclose(q[i], slerp64(q0, q1, t), atol=5e-7)
        np.testing.assert_allclose(tr[i], (1 - t) * vm0[:3, 3] + t * vm1[:3, 3], atol=1e-7)
        assert abs(np.linalg.norm(q[i]) - 1) < 1e-6  # the exact variant stays unit (the FAST device variant does not)
    # identical poses: the nearly-parallel branch (component-wise mix) returns the pose itself
    q, tr = shutter_pose(lib, Cam(vm0, vm0, shutter=RS_TB), [0.0, 0.37, 1.0])
    np.testing.assert_allclose(q, np.tile(quat_of(vm0[:3, :3]), (3, 1)), atol=3e-7)
    # antipodal representation of the second pose: the short way round is taken (no 360 degree swing)
    big = pose([0, 0, 1], np.pi * 0.98, [0, 0, 0])
    q, _ = shutter_pose(lib, Cam(np.eye(4), big, shutter=RS_TB), [0.5])
    R = quat_mat64(q[0].astype(np.float64))
    np.testing.assert_allclose(R, rot([0, 0, 1], np.pi * 0.49), atol=2e-6)


def test_global_shutter_world_to_image(lib):
    rng = np.random.default_rng(5)
    vm0 = pose([0.3, -1.0, 0.2], 0.4, [0.2, -0.1, 0.5])
    vm1 = pose([0.3, -1.0, 0.2], 0.6, [0.9, 0.9, 0.9])  # must be ignored under a global shutter
    pw = rng.uniform(-2, 2, (2000, 3)) + [0, 0, 5]
    pc = pw @ vm0[:3, :3].T + vm0[:3, 3]
    for kw in (dict(), dict(radial=RADIAL, tangential=TANGENTIAL), dict(model=FISHEYE, radial=FISH_K)):
        uv, ok = world_to_image(lib, Cam(vm0, vm1, **kw), pw)
        uv1, ok1 = world_to_image(lib, Cam(vm0, None, **kw), pw)
        uvc, okc = cam_project(lib, Cam(**kw), pc)
        assert (ok == ok1).all() and (uv == uv1).all()
        assert (ok == okc).mean() > 0.999
        both = ok & okc
        np.testing.assert_allclose(uv[both], uvc[both], atol=2e-3)  # float32 pose route (quaternion) vs float64 matrix
        assert both.sum() > 1000


@pytest.mark.parametrize("shutter", [RS_TB, RS_LR, RS_BT, RS_RL])
@pytest.mark.parametrize("kind", ["pinhole", "opencv", "fisheye"])
def test_rolling_shutter_fixed_point_and_rays(lib, shutter, kind):
    """world_to_image lands on a pixel whose own shutter pose projects the point back onto it, and the world ray of that
    pixel passes through the point (the two directions of Cameras.cuh:322-413 agree)."""
    rng = np.random.default_rng(6 + shutter)
    kw = dict(pinhole={}, opencv=dict(radial=RADIAL, tangential=TANGENTIAL, prism=PRISM), fisheye=dict(model=FISHEYE, radial=FISH_K))[kind]
    vm0 = pose([0.1, 1.0, 0.0], 0.02, [0.0, 0.0, 0.0])
    vm1 = pose([0.1, 1.0, 0.1], 0.05, [0.05, -0.02, 0.03])
    cam = Cam(vm0, vm1, shutter=shutter, **kw)
    pw = cone_points(rng, 1500, 0.45, 2.0, 9.0)
    uv, ok = world_to_image(lib, cam, pw, margin=0.0)
    assert ok.mean() > 0.7
    # it moved: the rolling shutter result differs from both end poses for most points
    uv0, _ = world_to_image(lib, Cam(vm0, None, **kw), pw, margin=0.0)
    assert np.median(np.abs(uv - uv0).max(-1)[ok]) > 0.5
    # fixed point, in float64: pose at time(u, v) projects the point to (u, v).  The time is a step function of the pixel
    # row / column, so an iterate sitting on a row boundary may alternate between two rows: allow one row's worth of motion.
    t = shutter_time(lib, cam, uv).astype(np.float64)
    q0, q1 = quat_of(vm0[:3, :3]), quat_of(vm1[:3, :3])
    res = np.zeros(len(pw))
    hit = np.zeros(len(pw))
    org, dirs, rok = pixel_to_world_ray(lib, cam, uv)
    for i in np.nonzero(ok)[0]:
        R = quat_mat64(slerp64(q0, q1, t[i]))
        tr = (1 - t[i]) * vm0[:3, 3] + t[i] * vm1[:3, 3]
        pc = (R @ pw[i] + tr)[None]
        if kind == "fisheye":
            ref, _ = fisheye_project64(pc, K, FISH_K)
        elif kind == "opencv":
            ref, _ = opencv_project64(pc, K, RADIAL, TANGENTIAL, PRISM)
        else:
            ref, _ = opencv_project64(pc, K, [0] * 6, [0] * 2, [0] * 4)
        res[i] = np.abs(ref[0] - uv[i]).max()
        if rok[i]:
            d = dirs[i].astype(np.float64)
            w = pw[i] - org[i]
            hit[i] = np.linalg.norm(w - d * (w @ d) / (d @ d)) / np.linalg.norm(w)
    sel = ok
    assert np.percentile(res[sel], 99) < 0.2, np.percentile(res[sel], [50, 90, 99, 100])
    assert np.median(res[sel]) < 2e-3
    assert rok[sel].mean() > 0.99
    # the ray of the pixel passes the point at a relative distance of the same order as the pixel residual
    assert np.median(hit[sel & rok]) < 1e-5 and np.percentile(hit[sel & rok], 99) < 5e-4
    assert np.allclose(np.linalg.norm(dirs[sel & rok], axis=1), 1, atol=1e-5)


def test_rolling_shutter_uses_end_pose_when_start_is_invalid(lib):
    """Cameras.cuh:360-378: the iteration starts from whichever end pose sees the point; neither -> invalid."""
    vm0 = pose([0, 1, 0], 0.0, [0, 0, 0])
    vm1 = pose([0, 1, 0], -0.5, [0, 0, 0])  # end pose has turned towards +x
    cam = Cam(vm0, vm1, shutter=RS_LR)
    tanx = (W - K[0, 2]) / K[0, 0]
    only_end = np.array([[tanx * 1.2 * 4, 0.0, 4.0]])  # right of the image at the start pose, visible at the end pose
    _, ok0 = world_to_image(lib, Cam(vm0, None), only_end, margin=0.0)
    _, ok1 = world_to_image(lib, Cam(vm1, None), only_end, margin=0.0)
    assert (not ok0[0]) and ok1[0]
    uv, ok = world_to_image(lib, cam, only_end, margin=0.0)
    assert ok[0] and np.isfinite(uv).all()
    neither = np.array([[0.0, 0.0, -3.0]])
    uv, ok = world_to_image(lib, cam, neither, margin=0.0)
    assert not ok[0]


def test_unproject_failure_zeroes_the_world_ray(lib):
    cam = Cam(model=FISHEYE, radial=FISH_K)
    org, dirs, ok = pixel_to_world_ray(lib, cam, [[K[0, 2] + 50 * W, K[1, 2]]])
    assert not ok[0] and (org == 0).all() and (dirs == 0).all()


def test_world_ray_origin_is_camera_centre(lib):
    vm0 = pose([0.3, -1.0, 0.2], 0.4, [0.2, -0.1, 0.5])
    cam = Cam(vm0, None)
    org, dirs, ok = pixel_to_world_ray(lib, cam, [[10.5, 20.5], [600.5, 400.5]])
    centre = -vm0[:3, :3].T @ vm0[:3, 3]
    assert ok.all()
    np.testing.assert_allclose(org, np.tile(centre, (2, 1)), atol=2e-6)
    # and the direction is R^T K^-1 (u, v, 1), normalised
    for i, (u, v) in enumerate([[10.5, 20.5], [600.5, 400.5]]):
        d = vm0[:3, :3].T @ np.array([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], 1.0])
        np.testing.assert_allclose(dirs[i], d / np.linalg.norm(d), atol=2e-6)
