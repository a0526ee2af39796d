"""Training step of the B200-native rasterizer: view-sharded data parallel + fused Adam.

Mirrors the reference's hot loop (src/training/trainer.cpp:579-757: render -> photometric loss -> backward ->
strategy step -> FusedAdam::step, src/training/optimizers/fused_adam.cpp:22-95) for the steady state (no
densification), generalised from the reference's batch of ONE view on ONE GPU to a batch of B views:
gradients of the views are summed, then one Adam step is taken.  With world_size > 1 the views of a step are
partitioned round-robin over the ranks (one process per GPU), the flat planar gradient arena is summed with a
single NCCL all-reduce over NVLink, and every rank applies the identical Adam update (parameters stay
replicated; no broadcast).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, dp
from ._lib import IMG_U8_HWC, TrainerDesc, UTParams, check, load

GROUPS = ("means", "sh0", "shN", "scaling", "rotation", "opacity")  # strategies/strategy_utils.cpp:35-40


def default_lrs(scene_scale: float = 1.0) -> dict:
    """eval/default_optimization_params.json (eval/, not parameter/)."""
    return {"means": 0.00016 * scene_scale, "sh0": 0.0025, "shN": 0.0025 / 20.0, "scaling": 0.005,
            "rotation": 0.001, "opacity": 0.05}


class SplatTrainer:
    def __init__(self, n_gaussians: int, width: int, height: int, sh_degree_max: int = 3, device="cuda:0",
                 instance_capacity: int = 0, lrs: Optional[dict] = None, betas=(0.9, 0.999), eps: float = 1e-15,
                 iterations: int = 30000, eps2d=0.3, near_plane=0.01, far_plane=1e4, radius_clip=0.0,
                 scale_reg: float = 0.0, opacity_reg: float = 0.0, view_streams: Optional[int] = None):
        """view_streams: how many views train_step keeps in flight (each on its own stream and its own per-view scratch;
        default 2, LFS_VIEW_STREAMS overrides).  The single-view calls (forward / loss_* / backward) use lane 0."""
        self.lib = load()
        self.device = torch.device(device)
        self.N, self.W, self.H, self.deg_max = n_gaussians, width, height, sh_degree_max
        self.K = (sh_degree_max + 1) ** 2
        self.Np = (n_gaussians + 3) // 4 * 4
        self.desc = TrainerDesc(n_gaussians, sh_degree_max, width, height, eps2d, near_plane, far_plane, radius_clip,
                                UTParams.default(), instance_capacity)
        if view_streams is None:
            view_streams = int(os.environ.get("LFS_VIEW_STREAMS", "2"))
        self.n_lanes = max(1, min(4, int(view_streams)))
        self._hs = [None] * self.n_lanes
        self._cur = 0
        self.lanes_enabled = True  # False: train_step runs its views one after the other on the current stream
        self._create_handle()
        n_floats = int(self.lib.lfs_trainer_arena_floats(C.byref(self.desc)))
        assert n_floats == (11 + 3 * self.K) * self.Np
        mk = lambda: torch.zeros(n_floats, dtype=torch.float32, device=self.device)
        self.params, self.grads, self.exp_avg, self.exp_avg_sq = mk(), mk(), mk(), mk()
        self._loss_lanes = torch.zeros(self.n_lanes, dtype=torch.float32, device=self.device)
        self.loss_dev = self._loss_lanes[0:1]  # lane 0 is what the single-view calls accumulate into
        # planar segment table in the reference's group order
        planes = [3, 3, 3 * (self.K - 1), 3, 4, 1]
        self.seg_begin = [0]
        for p in planes:
            self.seg_begin.append(self.seg_begin[-1] + p * self.Np)
        self.lrs = dict(lrs or default_lrs())
        self.betas, self.eps = betas, eps
        # mcmc regularisers (eval/mcmc_optimization_params.json: 0.01 each; src/training/trainer.cpp:132-158), folded into
        # the Adam step; _views_since_step counts the views whose loss they belong to (the reference adds them per view)
        self.scale_reg, self.opacity_reg = float(scale_reg), float(opacity_reg)
        self._views_since_step = 0
        self.step_count = [0] * 6
        self.iteration = 0
        self.means_gamma = math.pow(0.01, 1.0 / iterations)  # strategy_utils.cpp:47-55 (means group only)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._lane_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_lanes)] if self.n_lanes > 1 else []
        n_slots = 2 * self.n_lanes
        self._tgt = [torch.empty((height, width, 3), dtype=torch.uint8, device=self.device) for _ in range(n_slots)]
        self._tgt_ready = [torch.cuda.Event() for _ in range(n_slots)]
        self._tgt_free = [torch.cuda.Event() for _ in range(n_slots)]
        self._ev_start = torch.cuda.Event()
        self._ev_grads = [torch.cuda.Event() for _ in range(self.n_lanes)]  # "this lane's last gradient RMW is enqueued"
        self._loss_pinned = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.p2p = False

    @property
    def h(self):
        return self._hs[self._cur]

    def _create_handle(self) -> None:
        """(Re-)creates the per-view scratch of every lane from self.desc."""
        for k in range(self.n_lanes):
            old, self._hs[k] = self._hs[k], None
            if old:
                self.lib.lfs_trainer_destroy(old)
        with torch.cuda.device(self.device):
            for k in range(self.n_lanes):
                self._hs[k] = self.lib.lfs_trainer_create(C.byref(self.desc))
                if not self._hs[k]:
                    raise _lib.LfsError(-4, self.lib.lfs_last_error().decode())

    # ---- instance capacity (tile-Gaussian pairs per view) ---------------------------------------------------------
    @property
    def instance_capacity(self) -> int:
        return int(self.lib.lfs_trainer_instance_capacity(self.h))

    def poll_capacity(self) -> int:
        """Non-blocking: raises LfsError(LFS_ERR_CAPACITY) once a view whose tile instances did not fit has been seen
        (lfs_trainer_poll_capacity).  forward() / train_step() call it, so an overflow never goes unnoticed for more than
        one step.  Returns the largest per-view instance count observed so far."""
        worst = 0
        for h in self._hs:
            hw = C.c_uint64(0)
            check(self.lib.lfs_trainer_poll_capacity(h, C.byref(hw)))
            worst = max(worst, int(hw.value))
        return worst

    def ensure_capacity(self, viewmats, Ks, active_sh_degree: Optional[int] = None, headroom: float = 1.25) -> int:
        """Blocking calibration (outside any timed region): renders the given views once, and if one of them needs more
        tile instances than the scratch holds -- or the scratch is more than twice too large -- re-creates the per-view
        scratch with `headroom` x the largest need.  Parameters / optimiser state live in torch tensors and survive."""
        need = 0
        for vm, K in zip(viewmats, Ks):
            self._forward_unchecked(vm, K, active_sh_degree)
            n_inst, n_b = C.c_uint64(0), C.c_uint64(0)
            self.lib.lfs_trainer_stats(self.h, C.byref(n_inst), C.byref(n_b), self._stream())  # overflow is expected here
            need = max(need, int(n_inst.value))
        cap = self.instance_capacity
        if need > cap or cap > 2 * need + (1 << 20):
            self.desc.instance_capacity = int(need * headroom) + (1 << 16)
            torch.cuda.current_stream(self.device).synchronize()
            self._create_handle()
            torch.cuda.empty_cache()
        return need

    # ---- multi-GPU: peer-mapped arenas for the fused reduce-scatter + Adam + all-gather step -----------------------
    def enable_p2p(self, group=None) -> None:
        """Re-homes the parameter and gradient arenas in torch symmetric memory (CUDA backend: cuMem + fabric/IPC
        handles, every rank maps every peer's buffer over NVLink) so that adam_step() can run
        lfs_adam_step_multi_p2p instead of ncclAllReduce + a full local Adam.  Collective: call on every rank."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self._p2p_world, self._p2p_rank = dist.get_world_size(group), dist.get_rank(group)
        n = self.params.numel()
        new_p = symm_mem.empty(n, dtype=torch.float32, device=self.device)
        new_g = symm_mem.empty(n, dtype=torch.float32, device=self.device)
        new_p.copy_(self.params)
        new_g.copy_(self.grads)
        self._h_params = symm_mem.rendezvous(new_p, group)
        self._h_grads = symm_mem.rendezvous(new_g, group)
        self.params, self.grads = new_p, new_g
        # NVSwitch multicast (NVLS) addresses of the same buffers, 0 when the fabric has none.  The in-switch variant
        # (multimem.ld_reduce / multimem.st: the reduction happens inside the switch, 1/W of the inbound traffic) is used
        # from 4 ranks on when the fabric offers it, plain peer loads / stores otherwise (tools/check_p2p.py validates both
        # against ncclAllReduce + Adam)
        self._mc_grads = int(getattr(self._h_grads, "multicast_ptr", 0) or 0)
        self._mc_params = int(getattr(self._h_params, "multicast_ptr", 0) or 0)
        # measured (profiles/r02_bench_{2,8}gpu_*.json): at 8 GPUs the in-switch reduction wins (15.58 vs 15.68 ms per step) and
        # is bit-identical to ncclAllReduce + Adam; at 2 GPUs a rank's own half would cross NVLink twice and plain peer
        # loads are faster (15.23 vs 15.66 ms).  LFS_P2P_MULTICAST=1 / 0 forces either.
        want = os.environ.get("LFS_P2P_MULTICAST", "auto")
        use_mc = want == "1" or (want != "0" and self._p2p_world >= 4)
        if not use_mc or not (self._mc_grads and self._mc_params):
            self._mc_grads = self._mc_params = 0
        self.p2p = True

    def __del__(self):
        hs, self._hs = getattr(self, "_hs", []), []
        for h in hs:
            if h:
                self.lib.lfs_trainer_destroy(h)

    # ---- parameters ------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def load_params(self, means, sh0, shN, scaling, rotation, opacity) -> None:
        """AoS tensors in the reference's SplatData layout (include/core/splat_data.hpp:104-109)."""
        ts = [torch.as_tensor(t, dtype=torch.float32).to(self.device).contiguous()
              for t in (means, sh0, shN, scaling, rotation, opacity)]
        if self.K == 1:
            ts[2] = torch.zeros(1, device=self.device)
        check(self.lib.lfs_trainer_pack(self.h, *[t.data_ptr() for t in ts], self.params.data_ptr(), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def load_scene(self, scene) -> None:
        self.load_params(scene.means, scene.sh0, scene.shN, scene.scaling, scene.rotation, scene.opacity)

    def _unpack(self, arena: torch.Tensor) -> dict:
        N, K, dev = self.N, self.K, self.device
        out = {"means": torch.empty((N, 3), device=dev), "sh0": torch.empty((N, 1, 3), device=dev),
               "shN": torch.empty((N, max(K - 1, 1), 3), device=dev), "scaling": torch.empty((N, 3), device=dev),
               "rotation": torch.empty((N, 4), device=dev), "opacity": torch.empty((N, 1), device=dev)}
        check(self.lib.lfs_trainer_unpack(self.h, arena.data_ptr(), *[out[k].data_ptr() for k in GROUPS],
                                          self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        if K == 1:
            out["shN"] = out["shN"][:, :0]
        return out

    def export_params(self) -> dict:
        return self._unpack(self.params)

    def export_grads(self) -> dict:
        return self._unpack(self.grads)

    # ---- densification-side state surgery -----------------------------------------------------
    def gather_moments(self) -> None:
        """Multi-GPU (p2p): the fused step keeps the Adam moments of an element on its owner only.  Before the arenas
        are restructured every rank zeroes what it does not own and the moments are summed over the ranks, so that all
        ranks hold the complete state and apply the same surgery.  Ownership is positional (lfs_adam_p2p_owner), so
        afterwards every rank simply goes on updating the chunks it owns.  Collective; once per densification step."""
        if not self.p2p:
            return
        for arena in (self.exp_avg, self.exp_avg_sq):
            check(self.lib.lfs_adam_p2p_zero_unowned(arena.data_ptr(), arena.numel(), self._p2p_world, self._p2p_rank,
                                                     self._stream()))
            dp.allreduce_sum_(arena)

    def restructure(self, index: torch.Tensor, zero_state: Optional[torch.Tensor] = None) -> None:
        """One row gather over every plane of the parameter arena and both Adam moment arenas (lfs_arena_gather):
        Gaussian i of the new model = Gaussian index[i] of the old one; zero_state[i] starts it with fresh moments.
        Every strategy edit of the reference (duplicate / split / remove, default_strategy.cpp:49-230) is this call.
        The per-view scratch is re-created for the new count; gradients are dropped (the reference's zero_grad)."""
        if self.p2p:
            raise _lib.LfsError(_lib.LFS_ERR_UNSUPPORTED, "restructure: call gather_moments(), disable p2p, restructure, "
                                "enable_p2p() again (the symmetric arenas have a fixed size)")
        index = index.to(device=self.device, dtype=torch.int32).contiguous()
        n_new = int(index.numel())
        zs = None if zero_state is None else zero_state.to(device=self.device, dtype=torch.uint8).contiguous()
        np_new = (n_new + 3) // 4 * 4
        planes = 11 + 3 * self.K
        mk = lambda: torch.zeros(planes * np_new, dtype=torch.float32, device=self.device)  # noqa: E731
        new_p, new_m, new_v = mk(), mk(), mk()
        check(self.lib.lfs_arena_gather(self.params.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                        new_p.data_ptr(), new_m.data_ptr(), new_v.data_ptr(), index.data_ptr(),
                                        None if zs is None else zs.data_ptr(), n_new, self.N, planes, self.Np, np_new,
                                        self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        self.params, self.exp_avg, self.exp_avg_sq, self.grads = new_p, new_m, new_v, mk()
        cap_per_gauss = self.instance_capacity / max(self.N, 1)
        self.N, self.Np = n_new, np_new
        self.desc.n_gaussians = n_new
        self.desc.instance_capacity = int(cap_per_gauss * n_new) + (1 << 16)
        self._create_handle()
        planes_per_group = [3, 3, 3 * (self.K - 1), 3, 4, 1]
        self.seg_begin = [0]
        for p in planes_per_group:
            self.seg_begin.append(self.seg_begin[-1] + p * self.Np)

    def _rows(self, group: str, index: Optional[torch.Tensor] = None) -> torch.Tensor:
        first, cnt = dp.plane_table(self.K)[group]
        n = self.N if index is None else int(index.numel())
        out = torch.empty((n, cnt), dtype=torch.float32, device=self.device)
        check(self.lib.lfs_arena_get_rows(self.params.data_ptr(), self.Np, first, cnt,
                                          None if index is None else index.data_ptr(), out.data_ptr(), n, self.N,
                                          self._stream()))
        return out

    def _set_rows(self, group: str, index: torch.Tensor, rows: torch.Tensor) -> None:
        first, cnt = dp.plane_table(self.K)[group]
        rows = rows.to(torch.float32).contiguous()
        check(self.lib.lfs_arena_set_rows(self.params.data_ptr(), self.Np, first, cnt, index.data_ptr(), rows.data_ptr(),
                                          int(index.numel()), self.N, self._stream()))

    def remove(self, is_prune: torch.Tensor) -> None:
        """DefaultStrategy::remove (default_strategy.cpp:188-218)."""
        keep = torch.nonzero(~is_prune.to(self.device).bool()).squeeze(-1)
        self.restructure(keep)

    def duplicate(self, is_duplicated: torch.Tensor) -> None:
        """DefaultStrategy::duplicate (default_strategy.cpp:49-84): selected Gaussians are appended, fresh moments."""
        sel = torch.nonzero(is_duplicated.to(self.device).bool()).squeeze(-1)
        n0 = self.N
        index = torch.cat([torch.arange(n0, device=self.device), sel])
        zero = torch.zeros(index.numel(), dtype=torch.uint8, device=self.device)
        zero[n0:] = 1
        self.restructure(index, zero)

    def split(self, is_split: torch.Tensor, revised_opacity: bool = False, generator=None) -> None:
        """DefaultStrategy::split (default_strategy.cpp:86-160): a selected Gaussian is replaced by two samples of
        itself (means + R S eps, scales / 1.6, optionally revised opacity), appended after the unselected ones."""
        from . import ops
        is_split = is_split.to(self.device).bool()
        sel = torch.nonzero(is_split).squeeze(-1).to(torch.int32)
        rest = torch.nonzero(~is_split).squeeze(-1)
        ns = int(sel.numel())
        scales = torch.exp(self._rows("scaling", sel))
        quats = torch.nn.functional.normalize(self._rows("rotation", sel), dim=-1)
        means = self._rows("means", sel)
        op_raw = self._rows("opacity", sel)
        rot = ops.quats_to_rotmats(quats.contiguous())
        eps = torch.randn((2, ns, 3), device=self.device, generator=generator)
        samples = torch.einsum("nij,nj,bnj->bni", rot, scales, eps)
        index = torch.cat([rest, sel.long(), sel.long()])
        zero = torch.zeros(index.numel(), dtype=torch.uint8, device=self.device)
        zero[rest.numel():] = 1
        self.restructure(index, zero)
        new = torch.arange(rest.numel(), rest.numel() + 2 * ns, device=self.device, dtype=torch.int32)
        self._set_rows("means", new, (means.unsqueeze(0) + samples).reshape(-1, 3))
        self._set_rows("scaling", new, torch.log(scales / 1.6).repeat(2, 1))
        if revised_opacity:
            new_op = 1.0 - torch.sqrt(1.0 - torch.sigmoid(op_raw))
            self._set_rows("opacity", new, torch.logit(new_op).repeat(2, 1))
        torch.cuda.current_stream(self.device).synchronize()

    # ---- one view ----------------------------------------------------------------------------------------
    def forward(self, viewmat: np.ndarray, K: np.ndarray, active_sh_degree: Optional[int] = None,
                bg: Sequence[float] = (0.0, 0.0, 0.0), want_image: bool = False):
        self.poll_capacity()
        return self._forward_unchecked(viewmat, K, active_sh_degree, bg, want_image)

    def _forward_unchecked(self, viewmat: np.ndarray, K: np.ndarray, active_sh_degree: Optional[int] = None,
                           bg: Sequence[float] = (0.0, 0.0, 0.0), want_image: bool = False):
        vm = np.ascontiguousarray(viewmat, dtype=np.float32).reshape(16)
        kk = np.ascontiguousarray(K, dtype=np.float32).reshape(9)
        bgc = (C.c_float * 3)(*[float(b) for b in bg])
        deg = self.deg_max if active_sh_degree is None else active_sh_degree
        image = alpha = None
        if want_image:
            image = torch.empty((self.H, self.W, 3), dtype=torch.float32, device=self.device)
            alpha = torch.empty((self.H, self.W), dtype=torch.float32, device=self.device)
        check(self.lib.lfs_trainer_view_forward(
            self.h, self.params.data_ptr(), vm.ctypes.data_as(C.POINTER(C.c_float)),
            kk.ctypes.data_as(C.POINTER(C.c_float)), deg, bgc, image.data_ptr() if want_image else None,
            alpha.data_ptr() if want_image else None, self._stream()))
        return image, alpha

    def loss_l1(self, target: torch.Tensor, scale: Optional[float] = None, fmt: int = IMG_U8_HWC) -> None:
        s = 1.0 / (3.0 * self.W * self.H) if scale is None else scale
        check(self.lib.lfs_trainer_view_loss_l1(self.h, target.data_ptr(), fmt, s, self._loss_ptr(), self._stream()))

    def loss_ssim_l1(self, target: torch.Tensor, lambda_dssim: float = 0.2, weight: float = 1.0,
                     fmt: int = IMG_U8_HWC) -> None:
        """The reference's photometric loss (src/training/trainer.cpp:103-131): (1 - l) * L1 + l * (1 - SSIM_valid);
        lambda_dssim = 0.2 is eval/default_optimization_params.json."""
        check(self.lib.lfs_trainer_view_loss_ssim_l1(self.h, target.data_ptr(), fmt, lambda_dssim, weight,
                                                     self._loss_ptr(), self._stream()))

    def set_grad(self, v_image: torch.Tensor, v_alpha: Optional[torch.Tensor] = None) -> None:
        check(self.lib.lfs_trainer_view_set_grad(self.h, v_image.data_ptr(),
                                                 None if v_alpha is None else v_alpha.data_ptr(), self._stream()))

    def _loss_ptr(self) -> int:
        return self._loss_lanes.data_ptr() + 4 * self._cur

    def backward(self) -> None:
        check(self.lib.lfs_trainer_view_backward(self.h, self.params.data_ptr(), self.grads.data_ptr(),
                                                 self._stream()))
        self._views_since_step += 1

    def stats(self):
        n_inst, n_b = C.c_uint64(0), C.c_uint64(0)
        check(self.lib.lfs_trainer_stats(self.h, C.byref(n_inst), C.byref(n_b), self._stream()))
        return int(n_inst.value), int(n_b.value)

    PROF_STAGES = ("preprocess_fwd", "sort_intersect", "expand", "blend_fwd", "loss", "blend_bwd", "preprocess_bwd")

    def set_profile(self, enable: bool) -> None:
        check(self.lib.lfs_trainer_set_profile(self.h, 1 if enable else 0))

    def get_profile(self) -> dict:
        ms = (C.c_float * 7)()
        cnt = (C.c_int * 7)()
        check(self.lib.lfs_trainer_get_profile(self.h, ms, cnt))
        return {k: float(ms[i]) for i, k in enumerate(self.PROF_STAGES)}

    # ---- optimiser (FusedAdam::step mirror) -----------------------------------------------------------------
    def _adam_reg(self, run, n_views: int):
        """lfs_adam_reg for one run of consecutive groups: d/dp of scale_reg * mean(exp(scaling_raw)) and
        opacity_reg * mean(sigmoid(opacity_raw)), once per view of the (global) batch."""
        if (self.scale_reg == 0.0 and self.opacity_reg == 0.0) or n_views == 0:
            return None
        reg = _lib.AdamReg()
        reg.plane_elems, reg.n_valid = self.Np, self.N
        used = False
        for k, g in enumerate(run):
            name = GROUPS[g[0]]
            if name == "scaling" and self.scale_reg != 0.0:
                reg.kind[k], reg.coef[k], used = 1, self.scale_reg * n_views / (3.0 * self.N), True
            elif name == "opacity" and self.opacity_reg != 0.0:
                reg.kind[k], reg.coef[k], used = 2, self.opacity_reg * n_views / float(self.N), True
        return C.byref(reg) if used else None

    def adam_step(self, n_views: Optional[int] = None) -> None:
        """n_views: views of the GLOBAL batch this step closes (default: the views this rank rendered since the last
        step times the p2p / NCCL world size, i.e. an evenly sharded batch)."""
        self.iteration += 1
        if n_views is None:
            n_views = self._views_since_step * (self._p2p_world if self.p2p else getattr(self, "_dp_world", 1))
        self._views_since_step = 0
        b1, b2 = self.betas
        groups = []
        for gi, name in enumerate(GROUPS):
            self.step_count[gi] += 1
            if name == "shN" and (self.iteration <= 1000 or self.K == 1):
                continue  # fused_adam.cpp:68-70 (the counter is still incremented)
            t = self.step_count[gi]
            groups.append((gi, self.lrs[name], 1.0 / (1.0 - b1 ** t), 1.0 / math.sqrt(1.0 - b2 ** t)))
        # consecutive groups share one launch
        runs, cur = [], []
        for g in groups:
            if cur and g[0] != cur[-1][0] + 1:
                runs.append(cur)
                cur = []
            cur.append(g)
        if cur:
            runs.append(cur)
        for run in runs:
            n = len(run)
            seg = (C.c_int64 * (n + 1))(*([self.seg_begin[g[0]] for g in run] + [self.seg_begin[run[-1][0] + 1]]))
            lr = (C.c_float * n)(*[g[1] for g in run])
            bc1 = (C.c_float * n)(*[g[2] for g in run])
            bc2 = (C.c_float * n)(*[g[3] for g in run])
            if self.p2p:
                check(self.lib.lfs_adam_step_multi_p2p(self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                                       self._h_grads.buffer_ptrs_dev, self._h_params.buffer_ptrs_dev,
                                                       self._mc_grads or None, self._mc_params or None,
                                                       self.params.data_ptr(), self._p2p_world, self._p2p_rank, n, seg,
                                                       lr, bc1, bc2, b1, b2, self.eps, self._adam_reg(run, n_views),
                                                       self._stream()))
            else:
                check(self.lib.lfs_adam_step_multi(self.params.data_ptr(), self.exp_avg.data_ptr(),
                                                   self.exp_avg_sq.data_ptr(), self.grads.data_ptr(), n, seg, lr, bc1,
                                                   bc2, b1, b2, self.eps, 1, self._adam_reg(run, n_views), self._stream()))
        if self.p2p:
            self._h_params.barrier(channel=1)  # every rank's parameter writes have landed, nobody reads gradients any more
            self.grads.zero_()
        elif len(runs) != 1 or len(runs[0]) != 6:
            self.grads.zero_()  # skipped segments still need their gradients cleared
        self.lrs["means"] *= self.means_gamma  # ExponentialLR on group 0 only

    # ---- full step: B views -> one Adam update --------------------------------------------------------------
    def train_step(self, viewmats: np.ndarray, Ks: np.ndarray, targets_pinned: Sequence[torch.Tensor],
                   bg=(0.0, 0.0, 0.0), active_sh_degree: Optional[int] = None, world_size: int = 1, rank: int = 0,
                   read_loss: bool = True, lambda_dssim: Optional[float] = 0.2,
                   view_costs: Optional[Sequence[float]] = None):
        """targets_pinned: pinned host uint8 [H,W,3] tensors (or tensors already on the device), one per view of the GLOBAL
        batch (entries of views other ranks render are not touched); this rank renders
        views rank, rank + world_size, ... (or the cost-balanced partition of dp.shard_views when view_costs is given).
        Host->device copies run on a side stream, double buffered."""
        main = torch.cuda.current_stream(self.device)
        self._loss_lanes.zero_()
        my_views = dp.shard_views(len(targets_pinned), world_size, rank, view_costs)
        L = self.n_lanes if (len(my_views) > 1 and self.lanes_enabled) else 1
        n_slots = 2 * L
        if L > 1:
            self._ev_start.record(main)  # the lanes start behind the previous optimiser step
        last_grads = None
        for i, v in enumerate(my_views):
            slot, lane = i % n_slots, i % L
            st = self._lane_streams[lane] if L > 1 else main
            resident = targets_pinned[v].is_cuda  # already in HBM (bench.py's device-resident leg): no staging copy
            if not resident:
                with torch.cuda.stream(self.copy_stream):
                    if i >= n_slots:
                        self.copy_stream.wait_event(self._tgt_free[slot])
                    self._tgt[slot].copy_(targets_pinned[v], non_blocking=True)
                    self._tgt_ready[slot].record(self.copy_stream)
            target = targets_pinned[v] if resident else self._tgt[slot]
            self._cur = lane
            with torch.cuda.stream(st):
                if L > 1 and i < L:
                    st.wait_event(self._ev_start)
                self.forward(viewmats[v], Ks[v], active_sh_degree, bg)
                if not resident:
                    st.wait_event(self._tgt_ready[slot])
                if lambda_dssim is None:
                    self.loss_l1(target)
                else:
                    self.loss_ssim_l1(target, lambda_dssim)
                if not resident:
                    self._tgt_free[slot].record(st)
                if L == 1:
                    self.backward()
                else:
                    # the blend backward works on this lane's scratch; only the per-Gaussian kernels, which read-modify-write
                    # the shared gradient arena, are ordered behind the previous view's
                    check(self.lib.lfs_trainer_view_backward_blend(self.h, self._stream()))
                    if last_grads is not None:
                        st.wait_event(last_grads)
                    check(self.lib.lfs_trainer_view_backward_params(self.h, self.params.data_ptr(),
                                                                    self.grads.data_ptr(), self._stream()))
                    self._views_since_step += 1
                    last_grads = self._ev_grads[lane]
                    last_grads.record(st)
        self._cur = 0
        if L > 1:
            for lane in range(min(L, len(my_views))):
                main.wait_event(self._ev_grads[lane])
            self.loss_dev += self._loss_lanes[1:].sum()
        if world_size > 1:
            if self.p2p:
                self._h_grads.barrier(channel=0)  # all ranks finished their backward passes; then the fused step
            else:
                dp.allreduce_sum_(self.grads)  # ONE collective over the flat planar gradient arena
            dp.allreduce_sum_(self.loss_dev)
        self.adam_step(n_views=len(targets_pinned))
        if read_loss:
            self._loss_pinned.copy_(self.loss_dev, non_blocking=True)
        self.poll_capacity()
        return self._loss_pinned
