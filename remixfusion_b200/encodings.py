"""Encoders — drop-in for the reference's ``get_encoder`` (model/encodings.py:6-102) and for the ``tcnn.Encoding``
objects it returned.  Same call signature and return value ``(module, out_dim)``; modules expose what the reference
touches: ``forward(x[N,3]) -> [N,out] fp32``, ``.params`` (one flat fp32 ``nn.Parameter``; tiny-cuda-nn's layout:
levels concatenated, entry-major then feature), ``.n_output_dims``.  The arithmetic runs in
librf_b200.so (``rf_grid_encode_*`` / ``rf_oneblob_*``); there is no tiny-cuda-nn and no CPU fallback.

Only the encodings the reference's configs select are built: HashGrid / Dense grid (`grid.enc`), OneBlob (`pos.enc`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import abi


def make_grid_desc(n_levels, n_features, is_hash, log2_hashmap_size, base_resolution, per_level_scale) -> abi.GridDesc:
    d = abi.GridDesc()
    rc = abi.lib().rf_grid_desc_init(C.byref(d), C.c_int(n_levels), C.c_int(n_features), C.c_int(1 if is_hash else 0),
                                     C.c_int(log2_hashmap_size), C.c_int(base_resolution), C.c_double(per_level_scale))
    abi.check(rc, "rf_grid_desc_init")
    return d


class _GridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, desc):
        x32 = x.detach().to(torch.float32).contiguous()
        n = x32.shape[0]
        out = torch.empty(n, desc.n_output_dims, dtype=torch.float32, device=x32.device)
        rc = abi.lib().rf_grid_encode_forward(C.byref(desc), abi.dptr(params.detach()), abi.dptr(x32), C.c_int64(n),
                                              abi.dptr(out), abi.stream_ptr())
        abi.check(rc, "rf_grid_encode_forward")
        ctx.desc = desc
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(x32, params)
        return out

    @staticmethod
    def backward(ctx, dout):
        x32, params = ctx.saved_tensors
        need_x, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_p):
            return None, None, None
        dout = dout.contiguous().to(torch.float32)
        gp = torch.zeros_like(params) if need_p else None
        dx = torch.empty_like(x32) if need_x else None
        rc = abi.lib().rf_grid_encode_backward(C.byref(ctx.desc), abi.dptr(params.detach()), abi.dptr(x32),
                                               C.c_int64(x32.shape[0]), abi.dptr(dout), abi.dptr(gp), abi.dptr(dx),
                                               abi.stream_ptr())
        abi.check(rc, "rf_grid_encode_backward")
        return (dx.to(ctx.x_dtype) if need_x else None), gp, None


class GridEncoding(nn.Module):
    """tcnn ``Grid`` encoding (``HashGrid`` or ``Dense``), Linear interpolation, fp32 params and outputs."""

    def __init__(self, n_levels, n_features_per_level, base_resolution, per_level_scale, log2_hashmap_size=19,
                 is_hash=True, device=None):
        super().__init__()
        self.desc = make_grid_desc(n_levels, n_features_per_level, is_hash, log2_hashmap_size, base_resolution,
                                   per_level_scale)
        self.n_input_dims = 3
        self.n_output_dims = self.desc.n_output_dims
        dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        # tiny-cuda-nn initialises U(-1e-4, 1e-4) from its own pcg32 stream; values are not reproducible
        self.params = nn.Parameter((torch.rand(self.desc.n_params, dtype=torch.float32, device=dev) * 2 - 1) * 1e-4)

    def forward(self, x):
        return _GridFn.apply(x.reshape(-1, 3), self.params, self.desc)


class _OneBlobFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_bins):
        x32 = x.detach().to(torch.float32).contiguous()
        n = x32.shape[0]
        out = torch.empty(n, 3 * n_bins, dtype=torch.float32, device=x32.device)
        abi.check(abi.lib().rf_oneblob_forward(abi.dptr(x32), C.c_int64(n), C.c_int(n_bins), abi.dptr(out), abi.stream_ptr()),
                  "rf_oneblob_forward")
        ctx.n_bins, ctx.x_dtype = n_bins, x.dtype
        ctx.save_for_backward(x32)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x32,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        dx = torch.empty_like(x32)
        dout32 = dout.contiguous().to(torch.float32)          # keep the copy referenced until the kernel is enqueued
        abi.check(abi.lib().rf_oneblob_backward(abi.dptr(x32), C.c_int64(x32.shape[0]), C.c_int(ctx.n_bins),
                                                abi.dptr(dout32), abi.dptr(dx), abi.stream_ptr()),
                  "rf_oneblob_backward")
        return dx.to(ctx.x_dtype), None


class OneBlobEncoding(nn.Module):
    def __init__(self, n_bins=16, device=None):
        super().__init__()
        self.n_bins = n_bins
        self.n_input_dims = 3
        self.n_output_dims = 3 * n_bins
        dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        self.params = nn.Parameter(torch.zeros(0, dtype=torch.float32, device=dev))     # tcnn registers it regardless

    def forward(self, x):
        return _OneBlobFn.apply(x.reshape(-1, 3), self.n_bins)


def Encoding(n_input_dims, encoding_config, dtype=torch.float, device=None):
    """Constructor with the signature of ``tcnn.Encoding`` for the configurations the reference uses
    (model/encodings.py:39-50,67-74; model/scene_rep.py:60-93)."""
    if n_input_dims != 3 or dtype not in (torch.float, torch.float32):
        raise abi.RfError("Encoding: only 3-D inputs and dtype=torch.float are built (the reference uses nothing else)")
    ot = encoding_config["otype"]
    if ot == "HashGrid" or (ot == "Grid" and encoding_config.get("type", "Hash") == "Hash"):
        return GridEncoding(encoding_config["n_levels"], encoding_config["n_features_per_level"],
                            encoding_config["base_resolution"], encoding_config["per_level_scale"],
                            encoding_config["log2_hashmap_size"], True, device)
    if ot == "Grid" and encoding_config.get("type") == "Dense":
        return GridEncoding(encoding_config["n_levels"], encoding_config["n_features_per_level"],
                            encoding_config["base_resolution"], encoding_config["per_level_scale"], 0, False, device)
    if ot == "OneBlob":
        return OneBlobEncoding(encoding_config["n_bins"], device)
    raise NotImplementedError(f"encoding {ot!r}: never selected by a shipped config (SURVEY.md §2.1 row 4)")


def get_encoder(encoding, input_dim=3, degree=4, n_bins=16, n_frequencies=12, n_levels=16, level_dim=2,
                base_resolution=16, log2_hashmap_size=19, desired_resolution=512):
    """model/encodings.py:6-102.  Returns (module, out_dim)."""
    enc = encoding.lower()
    if "dense" in enc:                                                  # :14-30
        n_levels = 4
        per_level_scale = np.exp2(np.log2(desired_resolution / n_levels) / (n_levels - 1))
        embed = Encoding(input_dim, {"otype": "Grid", "type": "Dense", "n_levels": n_levels,
                                     "n_features_per_level": level_dim, "base_resolution": base_resolution,
                                     "per_level_scale": per_level_scale, "interpolation": "Linear"}, torch.float)
    elif "hash" in enc or "tiled" in enc:                               # :33-51
        per_level_scale = np.exp2(np.log2(desired_resolution / n_levels) / (n_levels - 1)) if n_levels > 1 else 1.0
        embed = Encoding(input_dim, {"otype": "HashGrid", "n_levels": n_levels, "n_features_per_level": level_dim,
                                     "log2_hashmap_size": log2_hashmap_size, "base_resolution": base_resolution,
                                     "per_level_scale": per_level_scale}, torch.float)
    elif "blob" in enc:                                                 # :65-76
        embed = Encoding(input_dim, {"otype": "OneBlob", "n_bins": n_bins}, torch.float)
    else:
        raise NotImplementedError(f"get_encoder({encoding!r}): SphericalHarmonics / Frequency / Identity are never "
                                  "selected by the reference's configs and are not built")
    return embed, embed.n_output_dims
