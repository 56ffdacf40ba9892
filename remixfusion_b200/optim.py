"""Fused Adam — drop-in for the ``torch.optim.Adam`` the reference builds in
``SLAM.create_optimizer`` (mp_slam/slam.py:271-286): same constructor arguments and parameter groups, same
``state_dict`` layout (``step``, ``exp_avg``, ``exp_avg_sq``), dense torch semantics.  ``step()`` runs one kernel per
parameter (``rf_adam_step``); ``step(zero_grad=True)`` also clears the gradients in the same pass, which is what
``map_optimizer.step(); map_optimizer.zero_grad()`` (mp_slam/mapper.py:417-423) amounts to.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import torch

from . import abi


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=False):
        """capturable=True keeps the step count in a device tensor read by the kernel (as torch's capturable Adam does), so
        that ``step()`` contains no host-side state and can be captured in a CUDA graph (remixfusion_b200.graph)."""
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=capturable))

    @torch.no_grad()
    def step(self, closure=None, zero_grad=False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = abi.lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise abi.RfError("fused Adam: contiguous float32 CUDA parameters only")
                st = self.state[p]
                cap = bool(group.get("capturable", False))
                if len(st) == 0:
                    # torch keeps the step as a tensor: on the host, or on the device when capturable
                    st["step"] = torch.tensor(0.0, dtype=torch.float32, device=p.device if cap else "cpu")
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                rc = lib.rf_adam_step(abi.dptr(p), abi.dptr(p.grad), abi.dptr(st["exp_avg"]), abi.dptr(st["exp_avg_sq"]),
                                      C.c_int64(p.numel()), C.c_double(group["lr"]), C.c_double(b1), C.c_double(b2),
                                      C.c_double(group["eps"]), C.c_double(group["weight_decay"]),
                                      C.c_int64(0 if cap else int(st["step"].item())), abi.dptr(st["step"]) if cap else None,
                                      C.c_int(1 if zero_grad else 0), abi.stream_ptr())
                abi.check(rc, "rf_adam_step")
        return loss
