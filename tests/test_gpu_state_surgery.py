"""-m gpu: densification-side state surgery on the planar arenas against the reference's semantics
restated with torch index_select / cat on the exported AoS tensors (default_strategy.cpp:49-230): parameters AND both
Adam moments of every kept / duplicated / split Gaussian, bit-exact (it is a gather), fresh moments for new Gaussians,
and the trainer keeps training afterwards."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from lichtfeld_studio_b200 import scene  # noqa: E402
from lichtfeld_studio_b200.trainer import GROUPS, SplatTrainer  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _trained(n=5003, w=160, h=112, deg=2):
    sc = scene.make_scene(n, 2, w, h, deg, seed=3, sigma_px=4.0)
    tr = SplatTrainer(n, w, h, deg, "cuda:0")
    tr.load_scene(sc)
    tr.iteration = 1000
    for v in range(2):  # non-trivial Adam moments
        tr.forward(sc.viewmats[v], sc.Ks[v], deg)
        tr.loss_ssim_l1(torch.as_tensor(scene.make_target(v, w, h)).cuda(), 0.2)
        tr.backward()
    tr.adam_step()
    torch.cuda.synchronize()
    return sc, tr


def _state(tr):
    return tr.export_params(), tr._unpack(tr.exp_avg), tr._unpack(tr.exp_avg_sq)


def test_remove_duplicate_split_match_the_reference_semantics():
    sc, tr = _trained()
    g = torch.Generator(device="cuda").manual_seed(0)
    p0, m0, v0 = _state(tr)
    n0 = tr.N
    # ---- remove
    prune = torch.rand(n0, device="cuda", generator=g) < 0.3
    keep = torch.nonzero(~prune).squeeze(-1)
    tr.remove(prune)
    p1, m1, v1 = _state(tr)
    assert tr.N == keep.numel()
    for k in GROUPS:
        assert torch.equal(p1[k], p0[k].index_select(0, keep)), k
        assert torch.equal(m1[k], m0[k].index_select(0, keep)) and torch.equal(v1[k], v0[k].index_select(0, keep)), k
    # ---- duplicate
    dup = torch.rand(tr.N, device="cuda", generator=g) < 0.2
    sel = torch.nonzero(dup).squeeze(-1)
    tr.duplicate(dup)
    p2, m2, v2 = _state(tr)
    for k in GROUPS:
        assert torch.equal(p2[k], torch.cat([p1[k], p1[k].index_select(0, sel)])), k
        assert torch.equal(m2[k][: p1[k].shape[0]], m1[k]) and float(m2[k][p1[k].shape[0]:].abs().max()) == 0.0, k
        assert torch.equal(v2[k][: p1[k].shape[0]], v1[k]) and float(v2[k][p1[k].shape[0]:].abs().max()) == 0.0, k
    # ---- split
    spl = torch.rand(tr.N, device="cuda", generator=g) < 0.25
    ssel, rest = torch.nonzero(spl).squeeze(-1), torch.nonzero(~spl).squeeze(-1)
    gen = torch.Generator(device="cuda").manual_seed(7)
    tr.split(spl, revised_opacity=True, generator=gen)
    p3, m3, v3 = _state(tr)
    nr, ns = rest.numel(), ssel.numel()
    assert tr.N == nr + 2 * ns
    # the reference's split (default_strategy.cpp:86-160) restated with torch on the AoS tensors, same noise
    gen = torch.Generator(device="cuda").manual_seed(7)
    scales = torch.exp(p2["scaling"].index_select(0, ssel))
    quats = torch.nn.functional.normalize(p2["rotation"].index_select(0, ssel), dim=-1)
    w, x, y, z = quats.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    eps = torch.randn((2, ns, 3), device="cuda", generator=gen)
    samples = torch.einsum("nij,nj,bnj->bni", R, scales, eps)
    want_means = torch.cat([p2["means"].index_select(0, rest),
                            (p2["means"].index_select(0, ssel).unsqueeze(0) + samples).reshape(-1, 3)])
    assert torch.allclose(p3["means"], want_means, rtol=1e-5, atol=1e-6)
    assert torch.allclose(p3["scaling"][nr:], torch.log(scales / 1.6).repeat(2, 1), rtol=1e-6, atol=1e-6)
    new_op = torch.logit(1.0 - torch.sqrt(1.0 - torch.sigmoid(p2["opacity"].index_select(0, ssel))))
    assert torch.allclose(p3["opacity"][nr:], new_op.repeat(2, 1), rtol=1e-5, atol=1e-6)
    for k in ("sh0", "shN", "rotation"):
        assert torch.equal(p3[k], torch.cat([p2[k].index_select(0, rest), p2[k].index_select(0, ssel).repeat(
            2, *([1] * (p2[k].dim() - 1)))])), k
    for k in GROUPS:
        assert torch.equal(m3[k][:nr], m2[k].index_select(0, rest)) and float(m3[k][nr:].abs().max()) == 0.0, k
        assert torch.equal(v3[k][:nr], v2[k].index_select(0, rest)) and float(v3[k][nr:].abs().max()) == 0.0, k
    # ---- and the trainer goes on training on the restructured model
    tr.forward(sc.viewmats[0], sc.Ks[0], 2)
    tr.loss_ssim_l1(torch.as_tensor(scene.make_target(0, sc.width, sc.height)).cuda(), 0.2)
    tr.backward()
    tr.adam_step()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(tr.params).all()) and tr.stats()[0] > 0
