"""Static properties of the compiled kernels (cuobjdump of the in-tree liblfs_b200.so; no GPU).

What the profiles claim about the hot kernels must be true of the binary that ships: the default forward blend stages its
records with TMA bulk copies behind an mbarrier, the backward blend and the SSIM kernels run on packed fp32 (FFMA2), neither
spills, both fit the occupancy their launch bounds promise, the multicast Adam kernel reduces through multimem.  Guards
against a refactor or a flag change silently losing one of them (a register bump that costs a CTA per SM, a gather falling
back to LDG, a spill)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_evidence as SE  # noqa: E402

REGS_PER_SM = 65536
SMEM_PER_SM = 228 * 1024             # shared-memory carve-out per SM (227 KB usable by one CTA)
CTA_SMEM_OVERHEAD = 1024             # reserved per resident CTA


@pytest.fixture(scope="module")
def info():
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(SE.LIB):
        import __graft_entry__ as G

        G.build()
    return SE.collect()


def fits(k, threads, ctas_per_sm):
    regs = (k["REG"] + 7) // 8 * 8  # allocation granularity
    return regs * threads * ctas_per_sm <= REGS_PER_SM and (k["SHARED"] + CTA_SMEM_OVERHEAD) * ctas_per_sm <= SMEM_PER_SM


def no_local_memory(k):
    return k["STACK"] == 0 and k["LOCAL"] == 0 and k["counts"]["LDL"] == 0 and k["counts"]["STL"] == 0


def test_default_kernel_list_matches_the_binary(info):
    for prefix in SE.DEFAULT_KERNELS:
        SE.select(info, prefix)  # raises unless exactly one kernel matches
    assert len(info) >= 90


@pytest.mark.parametrize("ewa", ["false", "true"])
def test_forward_blend_uses_tma_and_fits_ten_ctas(info, ewa):
    k = SE.select(info, "lfs::k_blend_fwd_tg<%s, 10>" % ewa)
    c = k["counts"]
    assert c["UBLKCP"] >= 1 and c["SYNCS"] >= 2, c       # cp.async.bulk + mbarrier arrive / try_wait
    assert c["LDG.E.128"] == 0, c                        # no register-staged gather of the records left
    assert c["MUFU.EX2"] >= 1 and c["LDS.128"] >= 8, c
    assert no_local_memory(k), k
    assert fits(k, 64, 10), k                            # __launch_bounds__(64, 10)
    if ewa == "false":
        assert c["FFMA2"] >= 16, c                       # the N' / D polynomial pairs


@pytest.mark.parametrize("ewa", ["false", "true"])
def test_backward_blend_is_packed_fp32_and_fits_five_ctas(info, ewa):
    k = SE.select(info, "lfs::k_blend_bwd_sp<%s, 4, 5>" % ewa)
    c = k["counts"]
    assert c["FFMA2"] >= 40 and c["FMUL2"] >= 4, c
    assert c["REDG"] >= 10, c                            # per-Gaussian gradients leave as fire-and-forget reductions
    assert c["ATOMG"] <= 1, c                            # the only atomic with a return value is the work counter
    assert no_local_memory(k), k
    assert fits(k, 128, 5), k


def test_ab_variants_keep_their_register_targets(info):
    assert SE.select(info, "lfs::k_blend_bwd_sp<false, 4, 6>")["REG"] <= 80   # bwd_variant 3: 6 CTAs per SM
    assert SE.select(info, "lfs::k_blend_fwd_tg<false, 12>")["REG"] <= 80     # fwd_variant 2: 12 CTAs per SM


def test_loss_kernels_are_packed_and_spill_free(info):
    for name in ("lfs::k_ssim_fwd(", "lfs::k_ssim_bwd("):
        k = SE.select(info, name)
        assert k["counts"]["FFMA2"] >= 30 and no_local_memory(k), k


def test_step_kernels_do_not_spill(info):
    """Every kernel of the default C3 training step (profiles/r02_launch_shares.txt) is free of local memory."""
    for name in ("lfs::k_preprocess_fwd(", "lfs::k_emit_instances_cull(", "lfs::k_rs_hist(", "lfs::k_rs_scatter<0>",
                 "lfs::k_tile_offsets(", "lfs::k_bucket_counts(", "lfs::k_live_buckets(", "lfs::k_preprocess_bwd_sh<3>",
                 "lfs::k_preprocess_bwd_geo(", "lfs::k_adam_multi(", "lfs::k_fg_preprocess(", "lfs::k_fg_emit("):
        assert no_local_memory(SE.select(info, name)), name


def test_exchange_kernels(info):
    mc = SE.select(info, "lfs::k_adam_multi_mc(")
    assert mc["counts"]["LDGMC"] >= 1, mc["counts"]                       # multimem.ld_reduce
    assert mc["counts"]["STG.E.128.STRONG.SYS"] >= 1, mc["counts"]        # multimem.st (system-scope store to the multicast VA)
    p2p = SE.select(info, "lfs::k_adam_multi_p2p<8>")
    assert p2p["counts"]["LDG.E.128"] >= 8, p2p["counts"]                 # one 128-bit load per peer, unrolled over the world
    assert no_local_memory(mc) and no_local_memory(p2p)


def test_no_tensor_core_or_library_kernels(info):
    """The path is HBM / issue bound integer and fp32 work: nothing in the library pretends otherwise (no MMA), and every
    kernel is this project's (namespace lfs), i.e. no CUB / cuBLAS instantiation was linked in."""
    for k in info.values():
        assert "lfs::" in k["name"], k["name"]
