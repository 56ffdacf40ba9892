"""TEST INFRASTRUCTURE ONLY — pure-PyTorch (CPU, autograd) stand-ins for the three ``tcnn.Encoding`` types the
reference instantiates (model/encodings.py:33-51 HashGrid, :65-76 OneBlob; model/scene_rep.py:60-93 Dense grids).

PARITY UNPINNED for this file: tiny-cuda-nn is an un-vendored, un-pinned dependency of the reference
(requirements.txt:21, `git+https://github.com/NVlabs/tiny-cuda-nn/`), it is not installed here, and the reference
ships no test or golden vector at that boundary.  The arithmetic below restates the published upstream algorithm
(include/tiny-cuda-nn/encodings/grid.h `kernel_grid`, common_device.h `grid_scale`/`grid_resolution`/`pos_fract`/
`grid_index`/`coherent_prime_hash`, encodings/oneblob.h `quartic_cdf`), summarised in SURVEY.md Appendix B.
What IS pinned: how the reference *calls* these modules (constructor configs, call sites, layouts) — the harness
in oracle/ref_import.py runs the reference's own model/scene_rep.py on top of this stand-in.

The module also mimics the small part of the ``tinycudann`` API the reference touches: ``Encoding(n_input_dims,
encoding_config, dtype)`` with ``.params``, ``.n_output_dims`` and ``forward``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

PRIMES = (1, 2654435761, 805459861)
U32 = 0xFFFFFFFF


def grid_levels(n_levels, n_features, is_hash, log2_hashmap_size, base_resolution, per_level_scale):
    """Appendix B1.  Returns lists scale (np.float32), resolution, size, offset (entries)."""
    pls = np.float32(per_level_scale)                      # JSON double narrowed to float by tcnn
    # log2 / exp2 evaluated in double and rounded once to float: reproducible on any host (tcnn: std::log2(float) on
    # the host and exp2f on the device, which may differ from this in the last bit — part of "parity unpinned")
    log2s = np.float32(np.log2(np.float64(pls)))
    scale, res, size, offset = [], [], [], [0]
    for l in range(n_levels):
        e = np.float32(np.exp2(np.float64(np.float32(l) * log2s)))
        s = np.float32(e * np.float32(base_resolution) - np.float32(1.0))
        r = int(np.ceil(s)) + 1
        max_params = (2 ** 32 - 1) // 2
        n = max_params if float(np.float32(r) ** 3) > float(max_params) else r ** 3
        n = (n + 7) // 8 * 8
        if is_hash:
            n = min(n, 1 << log2_hashmap_size)
        scale.append(s); res.append(r); size.append(n); offset.append(offset[-1] + n)
    return scale, res, size, offset


def _fma32(a: torch.Tensor, b: float, c: float) -> torch.Tensor:
    """fmaf(b, a, c) for fp32 tensors via an exact fp64 product (value only, no grad)."""
    return (a.detach().double() * float(b) + float(c)).float()


def grid_indices(x: torch.Tensor, scale, res, size, is_hash):
    """Corner indices [N,8] (int64, within the level), weights [N,8] (fp32, differentiable in x) — Appendix B2-B3."""
    pos_val = _fma32(x, float(scale), 0.5)
    cell = torch.floor(pos_val)
    lin = x * float(scale)                                 # carries d pos / d x = scale
    frac = (pos_val - cell) + (lin - lin.detach())
    cell_u = cell.to(torch.int64) & U32                    # (uint32_t)(int)floorf(pos)
    idxs, ws = [], []
    for corner in range(8):
        w = torch.ones_like(frac[:, 0])
        c = []
        for d in range(3):
            if (corner >> d) & 1:
                w = w * frac[:, d]
                c.append((cell_u[:, d] + 1) & U32)
            else:
                w = w * (1 - frac[:, d])
                c.append(cell_u[:, d])
        stride, index, d = 1, torch.zeros_like(c[0]), 0
        while d < 3 and stride <= size:
            index = (index + c[d] * stride) & U32
            stride = (stride * res) & U32
            d += 1
        if is_hash and size < stride:
            index = torch.zeros_like(c[0])
            for d in range(3):
                index = index ^ ((c[d] * PRIMES[d]) & U32)
        idxs.append(index % size)
        ws.append(w)
    return torch.stack(idxs, 1), torch.stack(ws, 1)


class GridStandIn(nn.Module):
    def __init__(self, n_levels, n_features, is_hash, log2_hashmap_size, base_resolution, per_level_scale):
        super().__init__()
        self.n_levels, self.n_features, self.is_hash = n_levels, n_features, is_hash
        self.scale, self.res, self.size, self.offset = grid_levels(
            n_levels, n_features, is_hash, log2_hashmap_size, base_resolution, per_level_scale)
        self.n_output_dims = n_levels * n_features
        g = torch.Generator().manual_seed(1337)
        self.params = nn.Parameter((torch.rand(self.offset[-1] * n_features, generator=g) * 2 - 1) * 1e-4)

    def forward(self, x):
        x = x.to(torch.float)
        table = self.params.view(-1, self.n_features)
        outs = []
        for l in range(self.n_levels):
            idx, w = grid_indices(x, self.scale[l], self.res[l], self.size[l], self.is_hash)
            acc = torch.zeros(x.shape[0], self.n_features, dtype=torch.float32)
            for corner in range(8):
                acc = acc + w[:, corner, None] * table[self.offset[l] + idx[:, corner]]
            outs.append(acc)
        return torch.cat(outs, 1)

    def level_indices(self, x):
        """Absolute table indices [N, L, 8] int64 — for the bit-exact hash-index parity test."""
        return torch.stack([self.offset[l] + grid_indices(x.float(), self.scale[l], self.res[l], self.size[l], self.is_hash)[0]
                            for l in range(self.n_levels)], 1)


def quartic_cdf(t, n_bins):
    u = t * float(n_bins)
    u2 = u * u
    u4 = u2 * u2
    return torch.clamp((15.0 / 16.0) * u * (1 - (2.0 / 3.0) * u2 + (1.0 / 5.0) * u4) + 0.5, 0.0, 1.0)


class OneBlobStandIn(nn.Module):
    """Appendix B7: out[:, d*n_bins + k] for coordinate d, bin k."""

    def __init__(self, n_bins, n_input_dims=3):
        super().__init__()
        self.n_bins, self.n_input_dims = n_bins, n_input_dims
        self.n_output_dims = n_bins * n_input_dims
        self.params = nn.Parameter(torch.zeros(0))

    def forward(self, x):
        x = x.to(torch.float)
        k = torch.arange(self.n_bins, dtype=torch.float32) / self.n_bins            # left boundaries
        d = k[None, None, :] - x[:, :, None]                                          # [N,3,bins]
        left = quartic_cdf(d, self.n_bins) + quartic_cdf(d - 1.0, self.n_bins) + quartic_cdf(d + 1.0, self.n_bins)
        right = torch.cat([left[:, :, 1:], left[:, :, :1] + 1.0], dim=2)              # wrap: L_16 := L_0 + 1
        return (right - left).reshape(x.shape[0], -1)


def Encoding(n_input_dims, encoding_config, dtype=torch.float, seed=1337):
    """Drop-in for ``tcnn.Encoding`` as called at model/encodings.py:39-50,67-74 and model/scene_rep.py:60-93."""
    assert n_input_dims == 3 and dtype == torch.float
    ot = encoding_config["otype"]
    if ot == "HashGrid" or (ot == "Grid" and encoding_config.get("type", "Hash") == "Hash"):
        return GridStandIn(encoding_config["n_levels"], encoding_config["n_features_per_level"], True,
                           encoding_config["log2_hashmap_size"], encoding_config["base_resolution"],
                           encoding_config["per_level_scale"])
    if ot == "Grid" and encoding_config["type"] == "Dense":
        return GridStandIn(encoding_config["n_levels"], encoding_config["n_features_per_level"], False, 0,
                           encoding_config["base_resolution"], encoding_config["per_level_scale"])
    if ot == "OneBlob":
        return OneBlobStandIn(encoding_config["n_bins"], n_input_dims)
    raise NotImplementedError(f"stand-in for tcnn encoding {ot!r} (never selected by a shipped config)")


class _NetworkStub:
    def __init__(self, *a, **k):
        raise NotImplementedError("tcnn.Network is disabled in every shipped config (decoder.tcnn_network: False)")


def as_module():
    """A fake ``tinycudann`` module object for sys.modules (used only by oracle/ref_import.py)."""
    import types
    m = types.ModuleType("tinycudann")
    m.Encoding = Encoding
    m.Network = _NetworkStub
    return m
