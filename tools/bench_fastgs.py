"""Times the fastgs (EWA) SURFACE, forward + backward per view, ours (lfs_fastgs_forward/backward through ops.py) and
the unmodified reference build (oracle/_ref/libfastgs_ref.so) on the same seeded scene:
    python tools/bench_fastgs.py C2 [views] [reps]
Prints one JSON line.  Both sides allocate per call and block on the instance / bucket counts, as the reference does."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lichtfeld_studio_b200 import ops, scene as S  # noqa: E402

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


def ours(v):
    w2c, cp, fx, fy, cx, cy = cams[v]
    img, al, ctx = ops.fastgs_forward(P["means"], P["scales"], P["rot"], P["op"], P["sh0"], P["shN"], w2c, cp, nb, W, H, fx,
                                      fy, cx, cy, 0.01, 1e10)
    g = ops.fastgs_backward(ctx, gi, ga, P["means"], P["scales"], P["rot"], P["shN"], w2c, cp)
    return ctx.n_instances, img, g


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


out = {"config": cfg, "gaussians": n, "width": W, "height": H, "views": views}
n_inst, img, g = ours(0)
out["instances_view0"] = n_inst
out["ours_ms_per_view"] = timeit(ours)
try:
    import ref_libs as R
    if R.have_fastgs():
        fg = R.FastGS()

        def ref(v):
            w2c, cp, fx, fy, cx, cy = cams[v]
            rimg, ral, counts = fg.forward(P["means"], P["scales"], P["rot"], P["op"], P["sh0"], P["shN"], w2c, cp, nb, W, H,
                                           fx, fy, cx, cy)
            rg = fg.backward(gi, ga, rimg, ral, P["means"], P["scales"], P["rot"], P["shN"], w2c, cp, nb, W, H, fx, fy, cx, cy)
            return counts, rimg, rg

        counts, rimg, rg = ref(0)
        torch.cuda.synchronize()
        out["ref_counts_view0"] = list(counts)
        out["ref_vs_ours_image_rel"] = float((rimg - img).abs().max() / img.abs().max())
        out["ref_vs_ours_grad_means_rel"] = float((rg["means"] - g[0]).abs().max() / g[0].abs().max())
        out["ref_ms_per_view"] = timeit(ref)
        out["speedup"] = out["ref_ms_per_view"] / out["ours_ms_per_view"]
except Exception as e:  # the reference build is known to fail at 1080p on this image (profiles/r01_ref_fastgs_diagnosis.txt)
    out["ref_error"] = repr(e)[:300]
print(json.dumps(out))
