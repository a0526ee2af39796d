"""Times the fastgs (EWA) OPERATOR SURFACE -- forward_wrapper + backward_wrapper of one view, the calls FastGSRasterize makes
(fastgs/rasterization/include/rasterization_api.h) -- on the same seeded scene through

    b200   oracle/_ref/b200_fastgs_torch*.so   the reference's symbols exported by this project's host layer
    ref    oracle/_ref/ref_fastgs_torch*.so    the unmodified reference fastgs build (compute_90 PTX objects, see
                                               profiles/r02_ref_fastgs_root_cause.txt)

    python tools/bench_fastgs.py C3 [views] [reps]

Both sides allocate their outputs and state blobs per call and block on the instance / bucket counts, as the reference does.
Prints one JSON line.  TEST / BENCH INFRASTRUCTURE ONLY."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_libs as R  # noqa: E402
import lichtfeld_studio_b200 as L  # noqa: E402
from lichtfeld_studio_b200 import scene as S  # noqa: E402

L.load()  # applies LFS_OPTIONS (A/B switches) to the one liblfs_b200.so of this process, the host layer's included

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
n, V, W, H, deg = S.CONFIGS[cfg]
views = int(sys.argv[2]) if len(sys.argv) > 2 else min(V, 4)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc = S.make_scene(n, views, W, H, deg, seed=42)
dev = "cuda:0"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).to(dev)  # noqa: E731
P = dict(means=T(sc.means), scales=T(sc.scaling), rot=T(sc.rotation), op=T(sc.opacity.reshape(-1, 1)), sh0=T(sc.sh0), shN=T(sc.shN))
cams = []
for v in range(views):
    vm = sc.viewmats[v].astype(np.float64)
    cams.append((T(vm), T(-vm[:3, :3].T @ vm[:3, 3]), float(sc.Ks[v, 0, 0]), float(sc.Ks[v, 1, 1]), float(sc.Ks[v, 0, 2]),
                 float(sc.Ks[v, 1, 2])))
gi = torch.randn((3, H, W), device=dev)
ga = torch.zeros((1, H, W), device=dev)
nb = (deg + 1) ** 2


def make_step(mod):
    dens = torch.zeros((2, n), device=dev)

    def step(v):
        w2c, cp, fx, fy, cx, cy = cams[v]
        r = mod.forward_wrapper(P["means"], P["scales"], P["rot"], P["op"], P["sh0"], P["shN"], w2c, cp, nb, W, H, fx, fy, cx,
                                cy, 0.01, 1e10)
        g = mod.backward_wrapper(dens, gi, ga, r[0], r[1], P["means"], P["scales"], P["rot"], P["shN"], r[2], r[3], r[4], r[5],
                                 w2c, cp, nb, W, H, fx, fy, cx, cy, 0.01, 1e10, r[6], r[7], r[8], r[9], r[10])
        return (int(r[6]), int(r[7]), int(r[8])), r[0], g

    return step


def timeit(fn):
    for v in range(views):
        fn(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        for v in range(views):
            fn(v)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / views)
    return best


out = {"config": cfg, "gaussians": n, "width": W, "height": H, "views": views,
       "what": "forward_wrapper + backward_wrapper per view (op surface, allocations and blocking count reads included)"}
ours = make_step(R.fastgs_torch_module("b200"))
counts, img, g = ours(0)
out["counts_view0"] = list(counts)
out["ours_ms_per_view"] = timeit(ours)
try:
    ref = make_step(R.fastgs_torch_module("ref"))
    rcounts, rimg, rg = ref(0)
    torch.cuda.synchronize()
    out["ref_counts_view0"] = list(rcounts)
    out["ref_vs_ours_image_rel"] = float((rimg - img).abs().max() / img.abs().max())
    out["ref_vs_ours_grad_means_rel"] = float((rg[0] - g[0]).abs().max() / g[0].abs().max())
    out["ref_ms_per_view"] = timeit(ref)
    out["speedup"] = out["ref_ms_per_view"] / out["ours_ms_per_view"]
except Exception as e:
    out["ref_error"] = repr(e)[:300]
print(json.dumps(out))
