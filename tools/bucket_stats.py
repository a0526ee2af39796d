"""Shape of the blend-backward work on a BASELINE config: per live 32-instance bucket, m = number of tile pixels that
reach it.  The backward warp runs m + 31 lock-step iterations per bucket; this prints the distribution of m, the
ramp share 31 / (m + 31) and the (pixel, instance) pair counts.   python tools/bucket_stats.py C3 [view]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lichtfeld_studio_b200 import scene as S  # noqa: E402
from lichtfeld_studio_b200._lib import check  # noqa: E402
from lichtfeld_studio_b200.trainer import SplatTrainer  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
view = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n, V, W, H, deg = S.CONFIGS[cfg]
sc = S.make_scene(n, max(view + 1, 1), W, H, deg, seed=42)
tr = SplatTrainer(n, W, H, deg, "cuda:0", instance_capacity=int(n * 12 + (1 << 20)))
tr.load_scene(sc)
tr.forward(sc.viewmats[view], sc.Ks[view], deg)
n_inst, n_buckets = tr.stats()


def grab(which, dtype, count):
    t = torch.empty(count, dtype=dtype, device="cuda:0")
    check(tr.lib.lfs_trainer_debug_copy(tr.h, which, t.data_ptr(), t.numel() * t.element_size(), None, tr._stream()))
    torch.cuda.synchronize()
    return t.cpu().numpy()


tw, th = (W + 15) // 16, (H + 15) // 16
toff = grab(0, torch.int32, tw * th + 1)
nc = grab(1, torch.int32, W * H).reshape(H, W)
pad = np.zeros((th * 16, tw * 16), np.int32)
pad[:H, :W] = nc
tiles = pad.reshape(th, 16, tw, 16).transpose(0, 2, 1, 3).reshape(th * tw, 256)  # n_contrib of every tile's pixels
cnt = toff[1:] - toff[:-1]
mx = tiles.max(axis=1)
live_b = (mx + 31) // 32
ms = []
for b in range(int(live_b.max()) + 1):
    sel = live_b > b
    if not sel.any():
        break
    ms.append((tiles[sel] > 32 * b).sum(axis=1))
m = np.concatenate(ms).astype(np.int64)
steps = m + 31
hist = np.bincount(np.minimum(m // 32, 8), minlength=9)
pairs_eval = int((np.minimum(np.maximum(tiles[:, :, None] - 32 * np.arange(int(live_b.max()) + 1)[None, None, :], 0), 32)
                  ).sum()) if live_b.max() < 64 else -1
out = {"config": cfg, "view": view, "instances": n_inst, "buckets_total": int(((cnt + 31) // 32).sum()),
       "buckets_live": int(m.size), "instances_in_live_buckets": int(np.minimum(cnt, live_b * 32).sum()),
       "mean_m": float(m.mean()), "median_m": float(np.median(m)), "sum_m": int(m.sum()), "sum_steps": int(steps.sum()),
       "ramp_share_of_steps": float(31.0 * m.size / steps.sum()),
       "m_hist_by_32": hist.tolist(), "pairs_evaluated": pairs_eval,
       "lane_utilisation_upper_bound": float(pairs_eval / (32.0 * steps.sum())) if pairs_eval > 0 else None,
       "pixels_mean_contrib": float(nc.mean()), "tile_inst_mean": float(cnt.mean()), "tile_inst_max": int(cnt.max()),
       "tile_live_buckets_mean": float(live_b.mean())}
print(json.dumps(out))
