"""Runs the UNMODIFIED reference fastgs forward on a synthetic scene of a given size (debug aid for the reference
leg of bench.py):  python tools/ref_fastgs_probe.py N W H [views]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_libs as R  # noqa: E402
from lichtfeld_studio_b200 import scene as S  # noqa: E402

n, W, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
views = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sc = S.make_scene(n, views, W, H, 3, seed=42)
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device="cuda:0")
fg = R.FastGS()
P = [T(sc.means), T(sc.scaling), T(sc.rotation), T(sc.opacity), T(sc.sh0), T(sc.shN)]
for v in range(views):
    w2c = T(sc.viewmats[v])
    campos = T(np.linalg.inv(sc.viewmats[v].astype(np.float64))[:3, 3])
    img, alpha, counts = fg.forward(*P, w2c, campos, 16, W, H, float(sc.Ks[v, 0, 0]), float(sc.Ks[v, 1, 1]),
                                    float(sc.Ks[v, 0, 2]), float(sc.Ks[v, 1, 2]))
    torch.cuda.synchronize()
    print("view", v, "counts", counts, "finite", bool(torch.isfinite(img).all()), flush=True)
