"""Seeded synthetic scenes for the BASELINE.json configurations (numpy only -> identical on every host).

Distributions follow the reference's own fixtures (tests/test_rasterization.cpp:456-469): means uniform in a
cube, quaternions normal -> normalised (w,x,y,z), log-scales log-normal tuned so that a Gaussian covers a few
16x16 tiles, opacity sigma(o) ~ U(0.3, 0.8), SH ~ 0.3 (U - 0.5); pinhole cameras on a ring looking at the
origin, fx = fy = 0.9 W, centred principal point, global shutter, no distortion.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CONFIGS = {
    # name: (N, n_views, W, H, sh_degree)
    "C1": (10_000, 1, 256, 256, 0),
    "C2": (100_000, 1, 800, 800, 3),
    "C3": (1_000_000, 8, 1920, 1080, 3),
    "C4": (1_000_000, 32, 1920, 1080, 3),
    "C5": (5_000_000, 8, 4096, 2160, 3),
}


@dataclass
class Scene:
    means: np.ndarray      # [N,3]
    sh0: np.ndarray        # [N,1,3]
    shN: np.ndarray        # [N,K-1,3]
    scaling: np.ndarray    # [N,3] raw (log)
    rotation: np.ndarray   # [N,4] raw (w,x,y,z), not normalised
    opacity: np.ndarray    # [N,1] raw (logit)
    viewmats: np.ndarray   # [V,4,4] world->camera, row-major
    Ks: np.ndarray         # [V,3,3]
    width: int
    height: int
    sh_degree: int

    @property
    def n(self) -> int:
        return self.means.shape[0]

    def activated(self):
        """(means, quats normalised, scales, opacities, shs[N,K,3]) as the reference's SplatData getters return
        them (src/core/splat_data.cpp:267-286)."""
        q = self.rotation / np.maximum(np.linalg.norm(self.rotation, axis=-1, keepdims=True), 1e-12)
        shs = np.concatenate([self.sh0, self.shN], axis=1)
        return (self.means, q.astype(np.float32), np.exp(self.scaling).astype(np.float32),
                (1.0 / (1.0 + np.exp(-self.opacity[:, 0]))).astype(np.float32), shs.astype(np.float32))


def look_at(eye: np.ndarray, target: np.ndarray, up=np.array([0.0, 1.0, 0.0])) -> np.ndarray:
    """world->camera [4,4] (OpenCV convention: +z forward, +x right, +y down)."""
    f = target - eye
    f = f / np.linalg.norm(f)
    r = np.cross(f, up)
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f], axis=0)
    m = np.eye(4)
    m[:3, :3] = R
    m[:3, 3] = -R @ eye
    return m


def make_cameras(n_views: int, width: int, height: int, radius: float = 3.5, seed: int = 0):
    rng = np.random.RandomState(1000 + seed)
    vms, Ks = [], []
    for v in range(n_views):
        ang = 2.0 * np.pi * v / max(n_views, 1) + 0.1
        elev = 0.15 * np.sin(3.0 * ang) + 0.05 * rng.uniform(-1, 1)
        eye = radius * np.array([np.cos(ang) * np.cos(elev), np.sin(elev), np.sin(ang) * np.cos(elev)])
        vms.append(look_at(eye, np.zeros(3)))
        f = 0.9 * width
        Ks.append(np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]]))
    return np.stack(vms).astype(np.float32), np.stack(Ks).astype(np.float32)


def make_scene(n: int, n_views: int, width: int, height: int, sh_degree: int, seed: int = 42,
               sigma_px: float = 3.7, radius: float = 3.5) -> Scene:
    rng = np.random.RandomState(seed)
    K = (sh_degree + 1) ** 2
    means = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    rotation = rng.normal(size=(n, 4)).astype(np.float32)
    base = sigma_px * radius / (0.9 * width)
    scaling = (np.log(base) + 0.5 * rng.normal(size=(n, 1)) + 0.3 * rng.normal(size=(n, 3))).astype(np.float32)
    op = rng.uniform(0.3, 0.8, size=(n, 1))
    opacity = np.log(op / (1.0 - op)).astype(np.float32)
    sh0 = (0.3 * (rng.uniform(size=(n, 1, 3)) - 0.5) + 0.5 * rng.uniform(size=(n, 1, 3))).astype(np.float32)
    band = np.concatenate([np.full(2 * l + 1, float(l)) for l in range(1, sh_degree + 1)]) if K > 1 else np.zeros(0)
    shN = (0.3 * (rng.uniform(size=(n, K - 1, 3)) - 0.5) / np.maximum(band, 1.0)[None, :, None]).astype(np.float32)
    viewmats, Ks = make_cameras(n_views, width, height, radius, seed)
    return Scene(means, sh0, shN, scaling, rotation, opacity, viewmats, Ks, width, height, sh_degree)


def make_config(name: str, seed: int = 42, n_override: int | None = None, views_override: int | None = None) -> Scene:
    n, v, w, h, deg = CONFIGS[name]
    return make_scene(n_override or n, views_override or v, w, h, deg, seed)


def make_target(view: int, width: int, height: int, seed: int = 7) -> np.ndarray:
    """Deterministic uint8 [H,W,3] ground-truth image (smooth gradients + stripes; content is irrelevant to
    throughput, only its size and format matter)."""
    ys, xs = np.mgrid[0:height, 0:width]
    ph = 0.37 * view + 0.01 * seed
    r = 0.5 + 0.5 * np.sin(xs * 0.013 + ph)
    g = 0.5 + 0.5 * np.sin(ys * 0.017 + 2.0 * ph)
    b = 0.5 + 0.5 * np.sin((xs + ys) * 0.007 + 3.0 * ph)
    return (np.stack([r, g, b], axis=-1) * 255.0).astype(np.uint8)
