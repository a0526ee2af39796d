"""Ownership rule of the fused reduce-scatter + Adam + all-gather step (lfs_adam_step_multi_p2p), host side only.
A rank keeps Adam moments ONLY for the arena elements it owns, so ownership must be a function of the element alone:
the same whatever run of segments a call covers (the shN group joins the run after iteration 1000,
src/training/optimizers/fused_adam.cpp:69, and the old per-run split then moved every slice boundary)."""
import ctypes as C

import numpy as np
import pytest


def owned_chunks(lib, lo, hi, world, rank):
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    assert lib.lfs_adam_p2p_owned_chunks(lo, hi, world, rank, C.byref(a), C.byref(b), C.byref(c)) == 0
    return [a.value + k * world for k in range(b.value)], c.value


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_every_chunk_has_exactly_one_owner_and_it_does_not_depend_on_the_run(world):
    import lichtfeld_studio_b200 as L
    lib = L.load()
    Np = 1_000_004  # padded Gaussian count of a 1 M scene
    planes = [3, 3, 45, 3, 4, 1]
    seg = np.concatenate([[0], np.cumsum(planes)]) * Np
    runs_early = [(seg[0], seg[2]), (seg[3], seg[6])]  # iteration <= 1000: [means, sh0] and [scaling, rotation, opacity]
    runs_late = [(seg[0], seg[6])]                     # afterwards: the whole arena
    owner_of = {}
    for runs in (runs_early, runs_late):
        for lo, hi in runs:
            seen = {}
            chunk = None
            for r in range(world):
                cs, chunk = owned_chunks(lib, int(lo), int(hi), world, r)
                for c in cs:
                    assert c not in seen, f"chunk {c} owned by ranks {seen[c]} and {r}"
                    seen[c] = r
                    assert owner_of.setdefault(c, r) == r, f"chunk {c}: owner changed between runs"
                    assert lib.lfs_adam_p2p_owner(c * chunk, world) == r
                    assert lib.lfs_adam_p2p_owner(c * chunk + chunk - 1, world) == r
            first, last = int(lo) // chunk, (int(hi) - 1) // chunk
            assert sorted(seen) == list(range(first, last + 1)), "the chunks of the run are not covered exactly once"
            # any sub-range is spread evenly: per-rank chunk counts differ by at most one
            counts = np.bincount(list(seen.values()), minlength=world)
            assert counts.max() - counts.min() <= 1


def test_bad_arguments_are_rejected():
    import lichtfeld_studio_b200 as L
    lib = L.load()
    z = C.c_int64(0)
    assert lib.lfs_adam_p2p_owned_chunks(0, 1024, 4, 4, C.byref(z), C.byref(z), C.byref(z)) < 0   # rank out of range
    assert lib.lfs_adam_p2p_owned_chunks(2, 1024, 4, 0, C.byref(z), C.byref(z), C.byref(z)) < 0   # unaligned
    assert lib.lfs_adam_p2p_owner(-1, 4) == -1 and lib.lfs_adam_p2p_owner(0, 0) == -1
