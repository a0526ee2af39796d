"""torchrun --nproc-per-node W tools/check_p2p.py : on the SAME per-rank gradients the fused P2P step must give the
same parameters as ncclAllReduce + local Adam (bit-identical for W = 2, where a + b has one summation order), every
rank must end with identical parameters, and the gradient arenas must come back cleared."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lichtfeld_studio_b200 import dp, scene as S  # noqa: E402
from lichtfeld_studio_b200.trainer import SplatTrainer  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n, W, H, deg, V = 20000, 320, 200, 3, 2 * world
sc = S.make_scene(n, V, W, H, deg, seed=5)
tr = SplatTrainer(n, W, H, deg, dev)
tr.load_scene(sc)
tr.iteration = 1000
for v in dp.shard_views(V, world, rank):  # this rank's gradients (atomics: not bit-reproducible, so they are computed ONCE)
    tr.forward(sc.viewmats[v], sc.Ks[v], deg, (0.1, 0.2, 0.3))
    tr.loss_ssim_l1(torch.as_tensor(S.make_target(v, W, H)).to(dev), 0.2)
    tr.backward()
torch.cuda.synchronize()
G, P0 = tr.grads.clone(), tr.params.clone()


START_ITER = 998  # three updates: iterations 999, 1000 (two runs, shN skipped) and 1001 (ONE run over the whole arena):
                  # the shN group joins the run and slice ownership must not move


def reset():
    tr.exp_avg.zero_(), tr.exp_avg_sq.zero_()
    tr.step_count = [0] * 6
    tr.iteration = START_ITER
    tr.lrs = dict(lrs0)


lrs0 = dict(tr.lrs)
res = {}
for mode in ("nccl", "p2p"):
    if mode == "p2p":
        tr.enable_p2p()
    tr.params.copy_(P0)
    reset()
    for _ in range(3):  # three updates on the same gradients, crossing iteration 1000 -> 1001
        tr.grads.copy_(G)
        torch.cuda.synchronize()
        dist.barrier()
        if tr.p2p:
            tr._h_grads.barrier(channel=0)
        else:
            dist.all_reduce(tr.grads)
        tr.adam_step()
    torch.cuda.synchronize()
    res[mode] = (tr.params.clone(), bool((tr.grads == 0).all()))
diff = (res["nccl"][0] - res["p2p"][0]).abs()
frac_equal = float((diff == 0).float().mean())
same = [torch.empty_like(res["p2p"][0]) for _ in range(world)]
dist.all_gather(same, res["p2p"][0])
ident = all(torch.equal(same[0], t) for t in same)
moved = float((res["p2p"][0] - P0).abs().max())
if rank == 0:
    print({"world": world, "multicast": bool(tr._mc_grads), "max_abs_diff_p2p_vs_nccl": float(diff.max()), "fraction_bit_equal": frac_equal,
           "max_param_change": moved, "ranks_bit_identical_after_p2p": ident,
           "grads_cleared": [res["nccl"][1], res["p2p"][1]]})
    assert moved > 0 and ident and res["nccl"][1] and res["p2p"][1]
    if world == 2:
        assert float(diff.max()) == 0.0
    else:  # other summation order of W partial sums: only elements with a vanishing gradient can flip Adam's sign-like step
        assert frac_equal > 0.98
dist.destroy_process_group()
