"""CPU checks of the measurement helpers: bench.py's per-stage roofline arithmetic and the launch-share tool on the committed
ncu launch list (profiles/r02_launches_bench_steps2.csv)."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_stage_roofline_arithmetic():
    b = _bench()
    prof = {"preprocess_fwd": 0.14, "sort_intersect": 0.44, "blend_fwd": 0.41, "loss": 0.24, "blend_bwd": 0.67,
            "preprocess_bwd": 0.17, "expand": 0.02}
    n, I, P, deg, peak = 1_000_000, 7.34e6, 1920.0 * 1080.0, 3, 6573.5
    r = b.stage_roofline(prof, n, I, P, deg, peak)
    assert set(r) == {"preprocess_fwd", "sort_intersect", "blend_fwd", "loss", "blend_bwd", "preprocess_bwd"}
    bwd = r["blend_bwd"]
    assert bwd["algorithmic_bytes"] == 172 * I + 24 * P
    assert abs(bwd["GBps"] - bwd["algorithmic_bytes"] / 0.67e-3 / 1e9) < 1e-6
    assert abs(bwd["frac_of_hbm_peak"] - bwd["GBps"] / peak) < 1e-12
    assert b.stage_roofline({"blend_bwd": 0.0}, n, I, P, deg, peak) == {}  # a stage that was not timed is left out


def test_launch_shares_of_the_committed_launch_list():
    csv = os.path.join(ROOT, "profiles", "r02_launches_bench_steps2.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), csv], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    head = out.stdout.splitlines()[0]
    assert "8 views" in head, head
    lines = [l for l in out.stdout.splitlines() if " x  " in l]
    shares = {l.split(" x  ")[1].strip(): float(l.split("%")[0]) for l in lines}
    top = max(shares, key=shares.get)
    assert top.startswith("k_blend_bwd_sp"), top  # the backward blend is the dominant kernel of a step
    assert abs(sum(shares.values()) - 100.0) < 0.5
