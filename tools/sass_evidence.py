"""Static evidence about the kernels in the in-tree liblfs_b200.so: per kernel, counts of the SASS mnemonics that prove TMA /
mbarrier / packed-fp32 / multimem use, and registers / stack / shared memory (cuobjdump -sass, -res-usage).  No GPU needed.

    python tools/sass_evidence.py                 # report for the kernels the launchers select by default
    python tools/sass_evidence.py --all           # every kernel
    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt

`collect()` is also what tests/test_sass_properties.py asserts on."""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lichtfeld-studio_b200", "liblfs_b200.so")

# mnemonic (prefix match on the opcode field) -> what it proves
MNEMONICS = [
    ("UBLKCP", "cp.async.bulk global->shared (TMA bulk copy)"),
    ("UTMALDG", "cp.async.bulk.tensor (TMA tensor tile)"),
    ("SYNCS", "mbarrier operations"),
    ("FFMA2", "fma.rn.f32x2"),
    ("FMUL2", "mul.rn.f32x2"),
    ("FADD2", "add.rn.f32x2"),
    ("MUFU.EX2", "ex2.approx"),
    ("MUFU.RCP", "rcp.approx"),
    ("MUFU.RSQ", "rsqrt.approx"),
    ("SHFL", "warp shuffles"),
    ("LDS.128", "128-bit shared loads"),
    ("LDG.E.128", "128-bit global loads"),
    ("STG.E.128", "128-bit global stores"),
    ("REDG", "global reductions (fire-and-forget atomics)"),
    ("ATOMG", "global atomics with return"),
    ("MATCH", "match.any"),
    ("LDGMC", "multimem.ld_reduce (NVSwitch multicast reduce)"),
    ("STG.E.128.STRONG.SYS", "system-scope 128-bit store (what multimem.st to a multicast address lowers to)"),
    ("LDL", "local-memory loads (spills / indexed private arrays)"),
    ("STL", "local-memory stores"),
]

# the kernels launch_* picks with every option at its default (csrc/*.cu), by demangled-name prefix
DEFAULT_KERNELS = [
    "lfs::k_preprocess_fwd(", "lfs::k_emit_instances_cull(", "lfs::k_rs_hist(", "lfs::k_rs_scatter<0>", "lfs::k_tile_offsets(",
    "lfs::k_bucket_counts(", "lfs::k_blend_fwd_tg<false, 10>", "lfs::k_live_buckets(", "lfs::k_ssim_fwd(", "lfs::k_ssim_bwd(",
    "lfs::k_blend_bwd_sp<false, 4, 5>", "lfs::k_preprocess_bwd_sh<3>", "lfs::k_preprocess_bwd_geo(", "lfs::k_adam_multi(",
    "lfs::k_adam_multi_mc(", "lfs::k_adam_multi_p2p<8>", "lfs::k_blend_fwd_tg<true, 10>", "lfs::k_blend_bwd_sp<true, 4, 5>",
    "lfs::k_fg_preprocess(", "lfs::k_fg_emit(", "lfs::k_fg_preprocess_bwd(", "lfs::k_blend_fwd_rays(", "lfs::k_blend_bwd_rays(",
    "lfs::k_projection_ut(",
]


def _run(args):
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(" ".join(args) + " failed:\n" + r.stderr[-2000:])
    return r.stdout


def _demangle(names):
    filt = shutil.which("c++filt") or shutil.which("cu++filt")
    if not filt or not names:
        return {n: n for n in names}
    out = subprocess.run([filt], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def collect(lib=LIB):
    """-> {mangled: {"name": demangled, "insts": n, "counts": {mnemonic: n}, "REG": r, "STACK": s, "SHARED": b, "LOCAL": l}}"""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(lib):
        raise FileNotFoundError(lib + " (run __graft_entry__.build() first)")
    info = {}
    cur = None
    op_re = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for line in _run([cuobjdump, "-sass", lib]).splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = info.setdefault(m.group(1), {"insts": 0, "counts": {k: 0 for k, _ in MNEMONICS}})
            continue
        if cur is None:
            continue
        m = op_re.match(line)
        if not m:
            continue
        op = m.group(1)
        cur["insts"] += 1
        for k, _ in MNEMONICS:
            if op.startswith(k):
                cur["counts"][k] += 1
    fn = None
    for line in _run([cuobjdump, "-res-usage", lib]).splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn in info and "REG:" in line:
            for key, val in re.findall(r"([A-Z]+)(?:\[\d+\])?:(\d+)", line):
                info[fn].setdefault(key, int(val))
            fn = None
    names = _demangle(list(info))
    for k in info:
        info[k]["name"] = names[k]
    return info


def select(info, prefix):
    hits = [v for v in info.values() if v["name"].startswith(prefix) or v["name"].startswith("void " + prefix)]
    if len(hits) != 1:
        raise KeyError("%r matches %d kernels" % (prefix, len(hits)))
    return hits[0]


def main():
    info = collect()
    everything = "--all" in sys.argv
    print("SASS evidence (cuobjdump -sass / -res-usage of the in-tree lichtfeld-studio_b200/liblfs_b200.so, sm_100a).")
    print("Counts of the mnemonics that prove TMA / mbarrier / packed-fp32 / multimem use, per kernel" +
          (" (every kernel)." if everything else " that the launchers select BY DEFAULT (tools/sass_evidence.py DEFAULT_KERNELS)."))
    for k, what in MNEMONICS:
        print("   %-10s %s" % (k, what))
    rows = sorted(info.values(), key=lambda v: v["name"]) if everything else [select(info, p) for p in DEFAULT_KERNELS]
    for v in rows:
        name = v["name"]
        print("\n== " + (name if len(name) < 150 else name[:147] + "..."))
        print("   %d instructions, REG %s, STACK %s, SHARED %s, LOCAL %s" % (v["insts"], v.get("REG"), v.get("STACK"), v.get("SHARED"),
                                                                            v.get("LOCAL")))
        print("   " + "  ".join("%s %d" % (k, n) for k, n in v["counts"].items() if n))


if __name__ == "__main__":
    main()
