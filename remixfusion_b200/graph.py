"""CUDA-graph capture of one mapping iteration.

At the batch sizes the reference trains with (``mapping.sample`` = 2048 rays x 59 samples, mp_slam/mapper.py:394-423) the
kernels of this library take ~0.15 ms while the Python / launch path around them takes ~0.9 ms: the iteration is
launch-bound.  ``GraphedMappingStep`` captures  mapping() -> total loss -> backward() -> fused Adam step + zero_grad  once for
a fixed ray count and replays it with one graph launch per iteration.  Everything inside is already on the device
(jitter is drawn with the device generator while capturing, the optimiser reads its step count from a device tensor).
With ``ray_grads=True`` the captured iteration is the bundle-adjustment one (mp_slam/mapper.py:456-495): the ray origins /
directions require gradients (``mapping(..., clamp=True)``), and the gradients w.r.t. them come back with the loss, for the caller's
pose optimiser.
"""
from __future__ import annotations

import torch

from . import abi
from .optim import Adam


class GraphedMappingStep:
    def __init__(self, model, optimizer: Adam, n_rays: int, loss_fn, eager_steps: int = 3, ray_grads: bool = False):
        """model: JointEncoding in train mode; optimizer: remixfusion_b200.optim.Adam(capturable=True) over its parameters;
        loss_fn(ret) -> scalar (e.g. ``lambda r: configs.total_loss(cfg, r)``, mp_slam/slam.py:162-169).  The first
        ``eager_steps`` calls run eagerly (they are real optimisation steps: lazy initialisation happens there), the next
        call captures the graph and every call from then on replays it."""
        if int(eager_steps) < 1:
            raise abi.RfError("GraphedMappingStep: eager_steps must be >= 1 (lazy initialisation — sampling tables, optimiser "
                              "state, per-camera caches — runs in the eager steps and cannot be captured)")
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise abi.RfError("GraphedMappingStep needs remixfusion_b200.optim.Adam(..., capturable=True)")
        dev = next(model.parameters()).device
        self.model, self.opt, self.loss_fn, self.n = model, optimizer, loss_fn, int(n_rays)
        self.ro = torch.zeros(self.n, 3, device=dev); self.rd = torch.zeros(self.n, 3, device=dev)
        self.tc = torch.zeros(self.n, 3, device=dev); self.td = torch.zeros(self.n, 1, device=dev)
        self.eager_left, self.graph, self.out = int(eager_steps), None, None
        self.ray_grads = bool(ray_grads)
        if self.ray_grads:                                  # static leaves with static gradient buffers (accumulated in place)
            self.ro.requires_grad_(True); self.rd.requires_grad_(True)
            self.ro.grad = torch.zeros_like(self.ro); self.rd.grad = torch.zeros_like(self.rd)

    def _iteration(self):
        if self.ray_grads:
            self.ro.grad.zero_(); self.rd.grad.zero_()
            ret = self.model.mapping(self.ro, self.rd, self.tc, self.td, clamp=True)
        else:
            ret = self.model.mapping(self.ro, self.rd, self.tc, self.td)
        loss = self.loss_fn(ret)
        loss.backward()
        self.opt.step(zero_grad=True)
        out = {k: v.detach() for k, v in ret.items()}
        if self.ray_grads:
            out["g_rays_o"], out["g_rays_d"] = self.ro.grad, self.rd.grad
        return loss.detach(), out

    def __call__(self, rays_o, rays_d, target_rgb, target_d):
        """Returns (loss, ret) — tensors that are overwritten by the next call once the graph is live."""
        if rays_o.shape[0] != self.n:
            raise abi.RfError(f"captured for {self.n} rays, got {rays_o.shape[0]}")
        with torch.no_grad():
            self.ro.copy_(rays_o); self.rd.copy_(rays_d); self.tc.copy_(target_rgb); self.td.copy_(target_d.reshape(self.n, 1))
        if self.eager_left > 0:
            self.eager_left -= 1
            return self._iteration()
        if self.graph is None:
            for g in self.opt.param_groups:                 # gradients must exist (and stay the same tensors) across replays
                for p in g["params"]:
                    if p.requires_grad and p.grad is None:
                        p.grad = torch.zeros_like(p)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._iteration()
        self.graph.replay()
        return self.out
