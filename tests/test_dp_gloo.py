"""N > 1 path on CPU: world_size-2 gloo run of the view-sharded data-parallel logic.  Each rank computes the
per-view gradients of ITS views with the CPU oracle (the test's stand-in for the device step), packs them into the
planar arena, all-reduces, applies Adam; the result must equal the single-process run over all views and be
bit-identical across the ranks."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, W, H, DEG, VIEWS = 120, 48, 32, 1, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _views_grads(views):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import lichtfeld_studio_b200  # noqa: F401
    import oracle as O
    from lichtfeld_studio_b200 import dp, scene
    sc = scene.make_scene(N, VIEWS, W, H, DEG, seed=5, sigma_px=3.0)
    raw = dict(means=sc.means, sh0=sc.sh0, shN=sc.shN, scaling=sc.scaling, rotation=sc.rotation, opacity=sc.opacity)
    K = (DEG + 1) ** 2
    arena = np.zeros_like(dp.pack_planar(raw, N, K), dtype=np.float64)
    for v in views:
        _, g = O.view_loss_grads(raw, sc.viewmats[v], sc.Ks[v], W, H, DEG, (0.1, 0.1, 0.1),
                                 target=scene.make_target(v, W, H))
        arena += dp.pack_planar(g, N, K).astype(np.float64)
    return dp.pack_planar(raw, N, K).astype(np.float64), arena


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import lichtfeld_studio_b200  # noqa: F401
    import oracle as O
    from lichtfeld_studio_b200 import dp
    params, grads = _views_grads(dp.shard_views(VIEWS, world, rank))
    t = torch.from_numpy(grads)
    dp.allreduce_sum_(t)
    p, m, v = O.adam_step(params, np.zeros_like(params), np.zeros_like(params), t.numpy(), 1e-2, 0.9, 0.999, 1e-15,
                          1.0 / (1 - 0.9), 1.0 / np.sqrt(1 - 0.999))
    q.put((rank, t.numpy().copy(), p))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharding_partition():
    sys.path.insert(0, ROOT)
    import lichtfeld_studio_b200  # noqa: F401
    from lichtfeld_studio_b200 import dp
    for world in (1, 2, 3, 8):
        seen = sorted(v for r in range(world) for v in dp.shard_views(13, world, r))
        assert seen == list(range(13))
    with pytest.raises(ValueError):
        dp.shard_views(4, 2, 2)


def test_planar_pack_roundtrip():
    sys.path.insert(0, ROOT)
    import lichtfeld_studio_b200  # noqa: F401
    from lichtfeld_studio_b200 import dp, scene
    sc = scene.make_scene(37, 1, 32, 32, 3, seed=1)
    raw = dict(means=sc.means, sh0=sc.sh0, shN=sc.shN, scaling=sc.scaling, rotation=sc.rotation, opacity=sc.opacity)
    a = dp.pack_planar(raw, 37, 16)
    assert a.shape == (59 * 40,)
    back = dp.unpack_planar(a, 37, 16)
    for k in raw:
        np.testing.assert_array_equal(back[k], raw[k])
    assert dp.plane_table(16)["_total"] == 59 and dp.plane_table(1)["_total"] == 14


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params, grads_all = _views_grads(range(VIEWS))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    want_p, _, _ = O.adam_step(params, np.zeros_like(params), np.zeros_like(params), grads_all, 1e-2, 0.9, 0.999,
                               1e-15, 1.0 / (1 - 0.9), 1.0 / np.sqrt(1 - 0.999))
    np.testing.assert_array_equal(res[0][1], res[1][1])  # identical gradient arenas after the all-reduce
    np.testing.assert_array_equal(res[0][2], res[1][2])  # hence identical parameters, no broadcast needed
    np.testing.assert_allclose(res[0][1], grads_all, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(res[0][2], want_p, rtol=1e-12, atol=1e-15)


def test_cost_balanced_view_sharding():
    """shard_views(costs=...): a partition (every view exactly once), equal view counts per rank like round-robin, and
    a smaller maximum rank load than round-robin on skewed costs; identical on every rank by construction."""
    import lichtfeld_studio_b200  # noqa: F401
    from lichtfeld_studio_b200 import dp
    rng = np.random.RandomState(0)
    for world, n in ((2, 7), (4, 32), (8, 64)):
        costs = rng.lognormal(0.0, 0.6, size=n)
        parts = [dp.shard_views(n, world, r, costs) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert [len(p) for p in parts] == [len(dp.shard_views(n, world, r)) for r in range(world)]
        load = max(costs[p].sum() for p in parts)
        rr = max(costs[dp.shard_views(n, world, r)].sum() for r in range(world))
        assert load <= rr + 1e-12
        assert load <= 1.15 * costs.sum() / world
    with pytest.raises(ValueError):
        dp.shard_views(4, 2, 0, [1.0, 2.0])
