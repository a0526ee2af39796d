"""Runs a few fused view steps of a workload (for ncu captures):
   ncu --set full --clock-control none --import-source on -k regex:k_blend -s 4 -c 4 -o gpurun_out/prof python tools/profile_view.py C3 2
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lichtfeld_studio_b200 import scene as S  # noqa: E402
from lichtfeld_studio_b200.trainer import SplatTrainer  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
views = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, _, W, H, deg = S.CONFIGS[cfg]
sc = S.make_scene(n, views, W, H, deg, seed=42)
tr = SplatTrainer(n, W, H, deg, "cuda:0", instance_capacity=int(os.environ.get("LFS_INST_CAP", "12000000")))
tr.load_scene(sc)
tr.iteration = 1000  # steady state: all six Adam groups active (shN is skipped for iterations <= 1000)
tg = [torch.as_tensor(S.make_target(v, W, H)).cuda() for v in range(views)]
for it in range(2):
    for v in range(views):
        tr.forward(sc.viewmats[v], sc.Ks[v], deg)
        tr.loss_ssim_l1(tg[v], 0.2)
        tr.backward()
    tr.adam_step()
torch.cuda.synchronize()
print("instances", tr.stats())
