"""Writes a seeded BASELINE config as the flat binary scene file read by oracle/ref_fastgs_standalone.cu:
    python tools/dump_scene.py C3 out.bin [n_override] [views_override]
int32 header {0x4c465331, N, K_rest, W, H, V, active_sh_bases}; fp32 means, scales_raw, rotations_raw, opacities_raw,
sh0, shN; per view w2c[16] cam_position[3] fx fy cx cy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lichtfeld_studio_b200 import scene as S  # noqa: E402


def dump(cfg: str, path: str, n_override: int = 0, views_override: int = 0) -> None:
    n, V, W, H, deg = S.CONFIGS[cfg]
    sc = S.make_scene(n_override or n, views_override or V, W, H, deg, seed=42)
    V = sc.viewmats.shape[0]
    with open(path, "wb") as f:
        np.array([0x4C465331, sc.n, sc.shN.shape[1], W, H, V, (deg + 1) ** 2], np.int32).tofile(f)
        for a in (sc.means, sc.scaling, sc.rotation, sc.opacity, sc.sh0, sc.shN):
            np.ascontiguousarray(a, np.float32).tofile(f)
        for v in range(V):
            vm = sc.viewmats[v].astype(np.float64)
            cam = -vm[:3, :3].T @ vm[:3, 3]
            rec = np.concatenate([sc.viewmats[v].reshape(-1), cam, [sc.Ks[v, 0, 0], sc.Ks[v, 1, 1], sc.Ks[v, 0, 2],
                                                                       sc.Ks[v, 1, 2]]]).astype(np.float32)
            rec.tofile(f)


if __name__ == "__main__":
    dump(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0,
         int(sys.argv[4]) if len(sys.argv) > 4 else 0)
