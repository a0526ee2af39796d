#!/usr/bin/env python
"""bench.py -- training views/sec (fwd + bwd + Adam) of the B200-native 3DGS rasterizer.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic views: every view is rendered (projection +
SH + tile sort + blend forward), its L1 + fused-SSIM photometric gradient (the reference's loss) is back-propagated (blend backward + per-Gaussian
backward), gradients of the batch are summed (NCCL all-reduce across ranks) and ONE fused Adam step is taken.
Workload = BASELINE.json configs[2] ("C3": 1M Gaussians, 8 views of 1920x1080 per GPU, SH degree 3); with N GPUs
the global batch is 8 N views (weak scaling; the views of a step are sharded round-robin over the ranks).

One JSON line is printed by rank 0 (see the contract in the task statement): `value` = views/s with the target
images already resident in HBM; `e2e` = the same through SplatTrainer.train_step with pinned HOST uint8 targets
(H2D per view, D2H of the loss per step); `roofline` = blend-backward kernel, algorithmic bytes (172 I + 24 P per
view) over its CUDA-event time against MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle
port timed on a bounded crop of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "training views/sec (fwd+bwd+Adam) @1M Gaussians 1080p"
UNIT = "views/s"
LAMBDA_DSSIM = 0.2  # eval/default_optimization_params.json; loss = (1-l) L1 + l (1 - SSIM), src/training/trainer.cpp:122-125


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--n-gaussians", type=int, default=0, help="override N (testing only; invalidates the metric)")
    ap.add_argument("--views-per-gpu", type=int, default=0)
    ap.add_argument("--dp", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N > 1: fused reduce-scatter+Adam+all-gather over NVLink peer memory (p2p) or ncclAllReduce + "
                         "local Adam (nccl); auto = p2p when symmetric memory can be set up")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / reference_gpu legs")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--ref-gpu-leg", default="", choices=["", "fastgs", "gsplat"], help=argparse.SUPPRESS)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(scene_obj, seconds: float):
    """The reference's path restated on the CPU (oracle port, OpenMP over all host cores) on a bounded crop of the
    same workload: one view of the full Gaussian set, a centred crop window of the 1080p image, fwd + bwd."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np

    import oracle as O

    sc = scene_obj
    cores = os.cpu_count() or 1
    W, H = sc.width, sc.height
    cw, ch = min(W, 256), min(H, 144)
    x0, y0 = (W - cw) // 2 // 16 * 16, (H - ch) // 2 // 16 * 16
    K = sc.Ks[0].copy()
    K[0, 2] -= x0
    K[1, 2] -= y0
    raw = dict(means=sc.means, sh0=sc.sh0, shN=sc.shN, scaling=sc.scaling, rotation=sc.rotation, opacity=sc.opacity)
    tgt = np.zeros((ch, cw, 3), np.uint8)
    t0 = time.time()
    reps = 0
    while True:
        O.view_loss_grads(raw, sc.viewmats[0], K, cw, ch, sc.sh_degree, (0.0, 0.0, 0.0), target=tgt, prec=32,
                          lambda_dssim=LAMBDA_DSSIM)
        reps += 1
        if time.time() - t0 > seconds or reps >= 3:
            break
    dt = (time.time() - t0) / reps
    frac = (cw * ch) / float(W * H)
    # per-Gaussian work (projection / SH) is paid in full by the crop; pixel work scales with the crop area:
    # views/s is reported for the crop as if it were the image (an upper bound for the CPU path)
    return {"value": frac / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 view, all {sc.n} Gaussians, centred {cw}x{ch} crop of {W}x{H} ({frac:.4f} of the pixels), "
                      f"fwd + L1/SSIM loss + bwd, {reps} reps, {dt:.2f} s each; value = crop fraction / time"}


def reference_gpu_leg(sc, steps: int, warmup: int, device):
    """UNMODIFIED reference fastgs CUDA path (oracle/_ref/libfastgs_ref.so) on the same scene, cameras and loss:
    per view forward -> L1 gradient (torch) -> backward, then its 6 Adam launches.  Reported beside our number as
    the north_star asks; it is a baseline, never part of the product path."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch

    import ref_libs as R
    if not R.have_fastgs():
        return {"unavailable": "oracle/_ref/libfastgs_ref.so not present"}
    fg = R.FastGS()
    T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=device)
    P = dict(means=T(sc.means), scales=T(sc.scaling), rot=T(sc.rotation), op=T(sc.opacity), sh0=T(sc.sh0),
             shN=T(sc.shN))
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}
    lrs = dict(means=0.00016, sh0=0.0025, shN=0.0025 / 20, scales=0.005, rot=0.001, op=0.05)
    W, H = sc.width, sc.height
    V = sc.viewmats.shape[0]
    from lichtfeld_studio_b200 import scene as S
    tg = [torch.as_tensor(S.make_target(v, W, H)).to(device).permute(2, 0, 1).float().div_(255.0).contiguous()
          for v in range(min(V, 8))]
    cams = []
    for v in range(V):
        w2c = T(sc.viewmats[v])
        campos = T(np.linalg.inv(sc.viewmats[v].astype(np.float64))[:3, 3])
        cams.append((w2c, campos, float(sc.Ks[v, 0, 0]), float(sc.Ks[v, 1, 1]), float(sc.Ks[v, 0, 2]),
                     float(sc.Ks[v, 1, 2])))
    nb = (sc.sh_degree + 1) ** 2
    scale = 1.0 / (3.0 * W * H)

    def step(t):
        acc = None
        for v in range(V):
            w2c, campos, fx, fy, cx, cy = cams[v]
            img, alpha, _ = fg.forward(P["means"], P["scales"], P["rot"], P["op"], P["sh0"], P["shN"], w2c, campos, nb,
                                       W, H, fx, fy, cx, cy)
            gimg = torch.sign(img - tg[v % len(tg)]) * scale
            galpha = torch.zeros_like(alpha)
            g = fg.backward(gimg, galpha, img, alpha, P["means"], P["scales"], P["rot"], P["shN"], w2c, campos, nb, W,
                            H, fx, fy, cx, cy)
            if acc is None:
                acc = g
            else:
                for k in ("means", "scales", "rot", "op", "sh0", "shN"):
                    acc[k] += g[k]
        bc1, bc2 = 1.0 / (1.0 - 0.9 ** t), 1.0 / (1.0 - 0.999 ** t) ** 0.5
        for k in ("means", "sh0", "shN", "scales", "rot", "op"):
            fg.adam_step(P[k], state[k][0], state[k][1], acc[k], lrs[k], 0.9, 0.999, 1e-15, bc1, bc2)

    for i in range(warmup):
        step(i + 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i + 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "reference fastgs (EWA) CUDA build, unmodified, sm_100a, --use_fast_math", "value": V / ms * 1e3,
            "unit": UNIT, "ms_per_step": ms, "views_per_step": V,
            "note": "targets resident in HBM; loss gradient by torch elementwise ops; includes its 3 blocking D2H reads"}


def reference_gsplat_gpu_leg(sc, steps: int, warmup: int, device):
    """UNMODIFIED reference gsplat (3DGUT) CUDA path (oracle/_ref/libgsplat_ref.so, built against oracle/glm_shim):
    per view projection_ut -> SH fwd -> intersect_tile (+CUB sort) -> intersect_offset -> rasterize fwd -> L1
    gradient -> rasterize bwd -> SH bwd, i.e. the kernel sequence of src/training/rasterization/rasterizer.cpp:208-360
    and its autograd backward, with the SplatData activations done by torch as the reference does; then the
    reference's 6 fused-Adam launches.  The torch autograd bookkeeping of the reference is NOT included (this is a
    lower bound of the reference's step time)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch

    import ref_libs as R
    if not R.have_gsplat():
        return {"unavailable": "oracle/_ref/libgsplat_ref.so not present"}
    T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=device)
    P = dict(means=T(sc.means), sh0=T(sc.sh0), shN=T(sc.shN), scales=T(sc.scaling), rot=T(sc.rotation), op=T(sc.opacity))
    W, H, deg = sc.width, sc.height, sc.sh_degree
    V = sc.viewmats.shape[0]
    tw, th = (W + 15) // 16, (H + 15) // 16
    from lichtfeld_studio_b200 import scene as S
    tg = [torch.as_tensor(S.make_target(v, W, H)).to(device).float().div_(255.0).contiguous() for v in range(min(V, 8))]
    cams = [(T(sc.viewmats[v:v + 1]), T(sc.Ks[v:v + 1]),
             T(np.linalg.inv(sc.viewmats[v].astype(np.float64))[:3, 3])) for v in range(V)]
    scale = 1.0 / (3.0 * W * H)
    fg = R.FastGS() if R.have_fastgs() else None
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}
    lrs = dict(means=0.00016, sh0=0.0025, shN=0.0025 / 20, scales=0.005, rot=0.001, op=0.05)
    n_isects = []

    def step(t):
        acc = {k: torch.zeros_like(v) for k, v in P.items()}
        for v in range(V):
            vm, K, campos = cams[v]
            # SplatData getters (src/core/splat_data.cpp:267-286)
            opac = torch.sigmoid(P["op"]).squeeze(-1)
            scl = torch.exp(P["scales"])
            quat = torch.nn.functional.normalize(P["rot"], dim=-1)
            shs = torch.cat([P["sh0"], P["shN"]], dim=1).contiguous()
            radii, m2d, dep, con, _ = R.projection_ut(P["means"], quat, scl, opac, vm, K, W, H)
            dirs = (P["means"] - campos[None]).contiguous()
            masks = (radii[0] > 0).all(-1).contiguous()
            cols = R.sh_fwd(deg, dirs, shs, masks)
            colors = torch.clamp_min(cols + 0.5, 0.0)
            _, ids, flat = R.intersect_tile(m2d, radii, dep, 16, tw, th, True)
            offs = R.intersect_offset(ids, 1, tw, th)
            if len(n_isects) < V:
                n_isects.append(int(flat.numel()))
            ren, al, li = R.raster_fwd(P["means"], quat, scl, colors[None].contiguous(), opac[None].contiguous(), None,
                                       W, H, 16, vm, K, offs, flat)
            v_ren = (torch.sign(ren[0] - tg[v % len(tg)]) * scale)[None].contiguous()
            v_al = torch.zeros_like(al)
            vmn, vq, vs, vc, vo = R.raster_bwd(P["means"], quat, scl, colors[None].contiguous(), opac[None].contiguous(),
                                               None, W, H, 16, vm, K, offs, flat, al, li, v_ren, v_al)
            vcol = vc[0] * (cols + 0.5 >= 0)
            v_shs, v_dirs = R.sh_bwd(deg, dirs, shs, vcol.contiguous(), masks, True)
            # activation VJPs (what torch autograd does for the reference)
            acc["means"] += vmn + v_dirs
            acc["sh0"] += v_shs[:, :1]
            acc["shN"] += v_shs[:, 1:]
            acc["scales"] += vs * scl
            dq = (vq * quat).sum(-1, keepdim=True)
            acc["rot"] += (vq - dq * quat) / P["rot"].norm(dim=-1, keepdim=True).clamp_min(1e-12)
            acc["op"] += (vo[0] * opac * (1 - opac))[:, None]
        bc1, bc2 = 1.0 / (1.0 - 0.9 ** t), 1.0 / (1.0 - 0.999 ** t) ** 0.5
        for k in ("means", "sh0", "shN", "scales", "rot", "op"):
            if fg is not None:
                fg.adam_step(P[k], state[k][0], state[k][1], acc[k].contiguous(), lrs[k], 0.9, 0.999, 1e-15, bc1, bc2)

    for i in range(warmup):
        step(i + 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i + 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "reference gsplat (3DGUT) CUDA kernels, unmodified, built against oracle/glm_shim, sm_100a, "
                    "--use_fast_math; kernel sequence of rasterizer.cpp without the autograd bookkeeping",
            "value": V / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "views_per_step": V,
            "instances_per_view": float(np.mean(n_isects)) if n_isects else None}


# ------------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    import numpy as np
    from lichtfeld_studio_b200 import scene as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_n, cfg_v, W, H, deg = S.CONFIGS[a.config]
    n = a.n_gaussians or cfg_n
    vpg = a.views_per_gpu or (8 if a.config in ("C3", "C4") else cfg_v)
    V = vpg * max(world, 1)

    if a.ref_gpu_leg:  # child process: time the unmodified reference CUDA build and print one JSON object
        import torch
        sc = S.make_scene(n, vpg, W, H, deg, seed=42)
        fn = reference_gpu_leg if a.ref_gpu_leg == "fastgs" else reference_gsplat_gpu_leg
        print("REFGPU " + json.dumps(fn(sc, max(2, min(a.steps, 4)), 1, torch.device("cuda:0"))))
        return 0

    if a.impl == "reference":
        # reference arm (tier rule): the reference's path on the box's host cores -- rank 0 only
        if rank != 0:
            return 0
        sc = S.make_scene(n, 1, W, H, deg, seed=42)
        t_all = time.time()
        vals = []
        for _ in range(max(1, min(a.steps, 3))):
            vals.append(cpu_reference_leg(sc, max(5.0, min(a.cpu_seconds, 30.0)) / 3.0))
        cb = dict(vals[-1])
        cb["value"] = statistics.median(v["value"] for v in vals)
        line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus,
                "steps": len(vals), "warmup": 0, "ms_per_step": 1e3 / cb["value"] if cb["value"] else None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{a.config}: {n} Gaussians, {vpg} views/GPU x {max(world, 1)} GPU of {W}x{H}, "
                                       f"SH degree {deg}, 3DGUT from-world rasterizer, L1 + fused-SSIM loss (lambda_dssim "
                                       f"{LAMBDA_DSSIM}); CPU port of the reference path (no CPU implementation of the "
                                       "blend exists in the reference: tests/test_rasterization.cpp:96-98)",
                           "gaussians": n, "views_per_step": V, "width": W, "height": H, "sh_degree": deg},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": round(time.time() - t_all, 1)}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from lichtfeld_studio_b200.trainer import SplatTrainer

    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    sc = S.make_scene(n, V, W, H, deg, seed=42)
    my_views = list(range(rank, V, world))

    tr = SplatTrainer(n, W, H, deg, device)
    tr.load_scene(sc)
    tr.iteration = 1000  # steady state: the reference skips the shN group only for iteration <= 1000
    # capacity calibration (one forward per local view, with sync) -- outside every timed region
    need = 0
    for v in my_views:
        tr.forward(sc.viewmats[v], sc.Ks[v], deg)
        try:
            ni, _ = tr.stats()
        except Exception:
            ni = int(tr.lib.lfs_trainer_instance_capacity(tr.h)) * 2
        need = max(need, ni)
    cap = int(tr.lib.lfs_trainer_instance_capacity(tr.h))
    if need > cap or cap > 2 * need + (1 << 20):
        new_cap = int(need * 1.25) + (1 << 16)
        del tr
        torch.cuda.empty_cache()
        tr = SplatTrainer(n, W, H, deg, device, instance_capacity=new_cap)
        tr.load_scene(sc)
        tr.iteration = 1000
    n_inst_per_view = []
    for v in my_views:
        tr.forward(sc.viewmats[v], sc.Ks[v], deg)
        n_inst_per_view.append(tr.stats()[0])

    dp_mode = "single"
    if world > 1:
        dp_mode = "nccl all-reduce of the flat gradient arena + local Adam"
        if a.dp in ("auto", "p2p"):
            try:
                tr.enable_p2p()
                dp_mode = ("fused reduce-scatter + Adam + all-gather in one kernel (lfs_adam_step_multi_p2p), " +
                           ("NVSwitch multicast: multimem.ld_reduce / multimem.st" if tr._mc_grads
                            else "NVLink peer loads / stores"))
            except Exception as e:  # symmetric memory unavailable on this box: the NCCL path is the validated default
                if a.dp == "p2p":
                    raise
                dp_mode += f" (p2p unavailable: {type(e).__name__})"
        flag = torch.tensor([1 if tr.p2p else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks or none
        if int(flag.item()) == 0:
            tr.p2p = False
    targets_host = [torch.as_tensor(S.make_target(v, W, H)).pin_memory() for v in range(V)]
    targets_dev = {v: targets_host[v].to(device) for v in my_views}
    bg = (0.0, 0.0, 0.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        tr.loss_dev.zero_()
        for v in my_views:
            tr.forward(sc.viewmats[v], sc.Ks[v], deg, bg)
            tr.loss_ssim_l1(targets_dev[v], LAMBDA_DSSIM)
            tr.backward()
        if world > 1:
            if tr.p2p:
                tr._h_grads.barrier(channel=0)
            else:
                dist.all_reduce(tr.grads, op=dist.ReduceOp.SUM)
        tr.adam_step()

    def step_e2e():
        tr.train_step(sc.viewmats, sc.Ks, targets_host, bg, deg, world, rank, read_loss=True)
        torch.cuda.current_stream().synchronize()  # the step's result (loss) is read on the host

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = tr.lib.lfs_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        launches = tr.lib.lfs_launch_count() - l0
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item()), int(launches)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_step, launches = timed(step_resident, a.steps, a.warmup)
    clk = clocks.stop() if rank == 0 else {}
    ms_e2e, _ = timed(step_e2e, a.steps, max(1, a.warmup))

    # per-kernel timing for the roofline (profiling pass, outside the timed regions)
    tr.set_profile(True)
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    prof = tr.get_profile()
    tr.set_profile(False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    I = float(np.mean(n_inst_per_view)) if n_inst_per_view else 0.0
    P = float(W * H)
    bwd_bytes = 172.0 * I + 24.0 * P
    fwd_bytes = 60.0 * I + 20.0 * P
    bwd_ms = prof.get("blend_bwd", 0.0)
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9 if bwd_ms > 0 else 0.0
    h2d = sum(int(targets_host[v].numel()) for v in my_views) + len(my_views) * (16 + 9) * 4
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture of tools/profile_view.py on this
    # workload (profiles/r01_ncu_full_final.json; ncu cannot run inside a timed bench)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_full_final.json")) as f:
            for k in json.load(f):
                if k["kernel"].startswith("k_blend_bwd"):
                    traffic = (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) * 1e6  # ncu reports Mbyte
                    traffic_src = "profiles/r01_ncu_full_final.json (dram__bytes_read.sum + dram__bytes_write.sum)"
                    break
    except (OSError, KeyError, ValueError):
        pass
    line = {
        "metric": METRIC, "value": V / ms_step * 1e3, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{a.config}: {n} Gaussians, {vpg} views/GPU x {world} GPU of {W}x{H}, SH degree {deg}, "
                               "3DGUT from-world rasterizer, L1 + fused-SSIM loss (lambda_dssim 0.2), eval/default_optimization_params.json lrs",
                   "gaussians": n, "views_per_step": V, "width": W, "height": H, "sh_degree": deg,
                   "parallelism": f"view-sharded dp{world}: {dp_mode}",
                   "instances_per_view": I, "l2": "per-step working set (params+grads+Adam state+records) >= 1 GB, "
                                                  "larger than the 126 MB L2: no flush needed"},
        "clocks": clk,
        "e2e": {"value": V / ms_e2e * 1e3, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "note": "SplatTrainer.train_step: pinned uint8 HWC targets copied per view on a side stream, "
                        "loss scalar read back per step"},
        "gpu_launches": launches,
        "roofline": {"kernel": "k_blend_bwd", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                     "algorithmic_bytes": bwd_bytes, "launch_ms": bwd_ms,
                     "note": "algorithmic bytes = 172*I + 24*P per view (BASELINE.md s5); the kernel is FP32/SFU "
                             "bound, not HBM bound (SURVEY s7 'roofline honesty')"},
        "stage_ms_per_view": prof,
        "stage_roofline": {"blend_fwd_GBps": fwd_bytes / (prof.get("blend_fwd", 0) * 1e-3) / 1e9
                           if prof.get("blend_fwd", 0) > 0 else None},
    }
    if not a.no_extras and world == 1:
        try:
            line["cpu_baseline"] = cpu_reference_leg(S.make_scene(n, 1, W, H, deg, seed=42), a.cpu_seconds)
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
        try:  # separate process: a fault inside the reference build must not take the bench down
            del tr
            torch.cuda.empty_cache()
            line["reference_gpu"] = {}
            for which in ("gsplat", "fastgs"):
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-gpu-leg", which, "--config",
                                    a.config, "--n-gaussians", str(a.n_gaussians), "--views-per-gpu", str(vpg),
                                    "--steps", str(a.steps)], capture_output=True, text=True, timeout=600)
                got = [l for l in r.stdout.splitlines() if l.startswith("REFGPU ")]
                line["reference_gpu"][which] = (json.loads(got[-1][7:]) if got
                                                else {"error": (r.stderr or r.stdout).strip().splitlines()[-1][:300]})
            if "error" in line["reference_gpu"].get("fastgs", {}):
                line["reference_gpu"]["fastgs"]["note"] = (
                    "the unmodified reference fastgs build returns garbage bucket counts above ~1e5 primitives on this "
                    "image (profiles/r01_ref_fastgs_diagnosis.txt); it runs, and matches ours, on the small parity cases")
        except Exception as e:
            line["reference_gpu"] = {"error": repr(e)}
        try:  # the EWA (fastgs) surface of this library on the same workload, op level (forward + backward per view)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_fastgs.py"), a.config, "2", "2"],
                               capture_output=True, text=True, timeout=300)
            got = [l for l in r.stdout.splitlines() if l.startswith("{")]
            line["fastgs_surface"] = (json.loads(got[-1]) if got
                                      else {"error": (r.stderr or r.stdout).strip().splitlines()[-1][:300]})
        except Exception as e:
            line["fastgs_surface"] = {"error": repr(e)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
