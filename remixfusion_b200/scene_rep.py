"""Mixed scene representation — drop-in for ``JointEncoding`` (model/scene_rep.py:12-529).

Same constructor, sub-module names (``embed_res_fn``, ``embedpos_fn``, ``GBV``, ``GBW``, ``decoder_res``, ``sdf_net_res``,
``color_net_res``, ``rba``), method names, arguments and return dictionaries as the reference, so mp_slam/mapper.py and
mp_slam/slam.py call it unchanged and checkpoints interchange (SURVEY.md §5 "Checkpoint / resume").

What differs is where the work happens: ``render_rays`` / ``mapping`` run the fused sm_100a kernels of librf_b200.so
(ray sampling -> hash + OneBlob + GBV trilerp -> decoder -> SDF-to-weight compositing -> loss sums, and the whole
backward) through one ``torch.autograd.Function``; the ``query_*`` point queries use the same kernels.  There is no
tiny-cuda-nn and no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import abi
from .decoder import ColorSDFNet
from .encodings import Encoding, get_encoder


def batchify(fn, chunk=1024 * 64):
    """model/utils.py:149-167."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def _chunk_rays(cfg, hdesc, n, S, dev):
    """Rays per launch.  The tensor-core path keeps 160 B of feature planes per sample between forward and backward and the
    backward needs about as much scratch; when that does not fit into the free device memory the batch is processed in
    chunks of rays (the backward then recomputes each chunk's planes).  RF_RAY_CHUNK forces a chunk size (tests)."""
    forced = int(os.environ.get("RF_RAY_CHUNK", "0"))
    if forced > 0:
        return max(128, forced // 128 * 128)
    if dev.type != "cuda" or n * S < (1 << 26) or torch.cuda.is_current_stream_capturing():
        return n
    L = abi.lib()
    probe = 1 << 16
    per_ray = 4.0 * (int(L.rf_ray_workspace_floats(C.byref(cfg), C.byref(hdesc), C.c_int64(probe))) +
                     int(L.rf_ray_scratch_floats(C.byref(cfg), C.byref(hdesc), C.c_int64(probe), C.c_int(1)))) / probe
    free, _ = torch.cuda.mem_get_info(dev)
    free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)          # blocks the caching allocator can reuse
    budget = 0.6 * free
    if per_ray * n <= budget:
        return n
    return max(128, int(budget / per_ray) // 128 * 128)


class _RayQueryFn(torch.autograd.Function):
    """rays (+ optional targets) -> rgb_map, depth_map, raw, losses[4].  Backward recomputes activations in-kernel."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, hash_params, w_sdf0, w_sdf1, w_col0, w_col1, gbv_params, z_vals, target_d,
                target_rgb, meta):
        ctx.set_materialize_grads(False)
        cfg, hdesc, gdesc, with_losses, group = meta["cfg"], meta["hash_desc"], meta["gbv_desc"], meta["with_losses"], meta["group"]
        dev = rays_o.device
        ro = rays_o.detach().to(torch.float32).contiguous(); rd = rays_d.detach().to(torch.float32).contiguous()
        n, S = ro.shape[0], z_vals.shape[1]
        raw = torch.empty(n, S, 4, dtype=torch.float32, device=dev)
        rgb_map = torch.empty(n, 3, dtype=torch.float32, device=dev)
        depth_map = torch.empty(n, dtype=torch.float32, device=dev)
        partials = torch.zeros(8, dtype=torch.float64, device=dev) if with_losses else None
        p = abi.RayParams(abi.dptr(hash_params.detach()), abi.dptr(gbv_params.detach()), abi.dptr(w_sdf0.detach()),
                          abi.dptr(w_sdf1.detach()), abi.dptr(w_col0.detach()), abi.dptr(w_col1.detach()))
        td = target_d.detach().reshape(-1).to(torch.float32).contiguous() if target_d is not None else None
        tc = target_rgb.detach().to(torch.float32).contiguous() if target_rgb is not None else None
        n_total = n
        if group is not None:
            import torch.distributed as dist
            if meta.get("equal_shards", False):
                n_total = n * dist.get_world_size(group)
            else:
                cnt = torch.tensor([n], dtype=torch.int64, device=dev)
                dist.all_reduce(cnt, group=group)
                n_total = int(cnt.item())
        cfg.n_rays_total = n_total
        chunk = _chunk_rays(cfg, hdesc, n, S, dev)
        nws = int(abi.lib().rf_ray_workspace_floats(C.byref(cfg), C.byref(hdesc), C.c_int64(min(n, chunk))))
        # feature planes of the tensor-core path: written by the forward, re-read by the backward (one chunk of rays at a time
        # when the whole batch's planes do not fit: the backward then recomputes them chunk by chunk)
        ws = torch.empty(nws, dtype=torch.float32, device=dev) if nws > 0 else None
        sl = lambda t, a, b: None if t is None else t[a:b]
        for a in range(0, max(n, 1), max(chunk, 1)):
            b = min(n, a + chunk)
            rc = abi.lib().rf_ray_query_forward(C.byref(cfg), C.byref(hdesc), C.byref(gdesc), C.byref(p), abi.dptr(ro[a:b]), abi.dptr(rd[a:b]),
                                                abi.dptr(sl(td, a, b)), abi.dptr(sl(tc, a, b)), abi.dptr(z_vals[a:b]), C.c_int64(b - a), abi.dptr(raw[a:b]),
                                                abi.dptr(rgb_map[a:b]), abi.dptr(depth_map[a:b]), abi.dptr(partials), abi.dptr(ws),
                                                abi.stream_ptr())
            abi.check(rc, "rf_ray_query_forward")
        if chunk < n:
            ws = None                                            # holds the last chunk only
        losses = torch.zeros(4, dtype=torch.float32, device=dev)
        if with_losses:
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(partials, group=group)          # loss means and mask counts are batch-global
            # rgb: mse over N*3 (scene_rep.py:501); depth: mean over valid rays (:504-507); sdf / fs with the balance
            # weights of get_masks (model/utils.py:190-196, :242-245) — one tiny kernel instead of a chain of scalar torch ops
            abi.check(abi.lib().rf_ray_loss_finalize(abi.dptr(partials), C.c_int64(n_total), C.c_int(S), abi.dptr(losses),
                                                     abi.stream_ptr()), "rf_ray_loss_finalize")
        ctx.meta = meta
        ctx.n_total = n_total
        ctx.chunk = chunk
        ctx.ws = ws if any(ctx.needs_input_grad) else None
        ctx.save_for_backward(ro, rd, hash_params, w_sdf0, w_sdf1, w_col0, w_col1, gbv_params, z_vals, td, tc, raw,
                              rgb_map, depth_map, partials)
        return rgb_map, depth_map, raw, losses

    @staticmethod
    def backward(ctx, d_rgb_map, d_depth_map, d_raw, d_losses):
        (ro, rd, hash_params, w_sdf0, w_sdf1, w_col0, w_col1, gbv_params, z_vals, td, tc, raw, rgb_map, depth_map,
         partials) = ctx.saved_tensors
        meta = ctx.meta
        cfg, hdesc, gdesc = meta["cfg"], meta["hash_desc"], meta["gbv_desc"]
        cfg.n_rays_total = ctx.n_total
        need = ctx.needs_input_grad
        dev = ro.device
        n, S = z_vals.shape
        f32 = lambda t: None if t is None else t.contiguous().to(torch.float32)
        g_hash = torch.zeros_like(hash_params) if need[2] else None
        g_w = [torch.zeros_like(w) if need[3 + i] else None for i, w in enumerate((w_sdf0, w_sdf1, w_col0, w_col1))]
        ba = need[0] or need[1]
        g_o = torch.empty_like(ro) if ba else None
        g_d = torch.empty_like(rd) if ba else None
        chunk = min(ctx.chunk, n) if n > 0 else 0
        nsc = int(abi.lib().rf_ray_scratch_floats(C.byref(cfg), C.byref(hdesc), C.c_int64(chunk), C.c_int(1 if ba else 0)))
        scratch = torch.empty(nsc, dtype=torch.float32, device=dev)
        p = abi.RayParams(abi.dptr(hash_params.detach()), abi.dptr(gbv_params.detach()), abi.dptr(w_sdf0.detach()),
                          abi.dptr(w_sdf1.detach()), abi.dptr(w_col0.detach()), abi.dptr(w_col1.detach()))
        use_loss = d_losses is not None and partials is not None
        # contiguous fp32 copies of the upstream gradients must stay referenced until the kernels are enqueued: a temporary
        # dropped right after dptr() hands its block back to the caching allocator, and the next temporary may reuse it
        up = [f32(d_rgb_map), f32(d_depth_map), f32(d_raw), f32(d_losses) if use_loss else None]
        ws = ctx.ws
        if chunk < n:                                            # the planes are recomputed chunk by chunk (see forward)
            nws = int(abi.lib().rf_ray_workspace_floats(C.byref(cfg), C.byref(hdesc), C.c_int64(chunk)))
            ws = torch.empty(nws, dtype=torch.float32, device=dev) if nws > 0 else None
        sl = lambda t, a, b: None if t is None else t[a:b]
        for a in range(0, max(n, 1), max(chunk, 1)):
            b = min(n, a + chunk)
            if chunk < n:
                # raw / rgb_map / depth_map come out bit-identical to the first pass; no loss sums this time (partials = NULL)
                rc = abi.lib().rf_ray_query_forward(C.byref(cfg), C.byref(hdesc), C.byref(gdesc), C.byref(p), abi.dptr(ro[a:b]), abi.dptr(rd[a:b]),
                                                    abi.dptr(None), abi.dptr(None), abi.dptr(z_vals[a:b]), C.c_int64(b - a), abi.dptr(raw[a:b]),
                                                    abi.dptr(rgb_map[a:b]), abi.dptr(depth_map[a:b]), abi.dptr(None), abi.dptr(ws), abi.stream_ptr())
                abi.check(rc, "rf_ray_query_forward (recompute)")
            grads = abi.RayGrads(abi.dptr(g_hash), abi.dptr(g_w[0]), abi.dptr(g_w[1]), abi.dptr(g_w[2]), abi.dptr(g_w[3]),
                                 abi.dptr(sl(g_o, a, b)), abi.dptr(sl(g_d, a, b)))
            rc = abi.lib().rf_ray_query_backward(
                C.byref(cfg), C.byref(hdesc), C.byref(gdesc), C.byref(p), abi.dptr(ro[a:b]), abi.dptr(rd[a:b]), abi.dptr(sl(td, a, b)), abi.dptr(sl(tc, a, b)),
                C.c_int64(b - a), abi.dptr(z_vals[a:b]), abi.dptr(raw[a:b]), abi.dptr(rgb_map[a:b]), abi.dptr(depth_map[a:b]),
                abi.dptr(sl(up[0], a, b)), abi.dptr(sl(up[1], a, b)), abi.dptr(sl(up[2], a, b)),
                abi.dptr(up[3]), abi.dptr(partials if use_loss else None),
                C.byref(grads), abi.dptr(ws), abi.dptr(scratch), abi.stream_ptr())
            abi.check(rc, "rf_ray_query_backward")
        return (g_o if need[0] else None, g_d if need[1] else None, g_hash, g_w[0], g_w[1], g_w[2], g_w[3],
                None, None, None, None, None)


class JointEncoding(nn.Module):
    def __init__(self, config, bound_box, num_kf=None, rba_factory=None, process_group=None, equal_shards=False):
        super().__init__()
        self.config = config
        self.bounding_box = bound_box                # float64 tensor [3,2] in the reference (run.py:90)
        self.num_kf = num_kf
        self.process_group = process_group           # ray batches sharded over ranks: loss sums are all-reduced
        # every rank passes the same number of rays to mapping(): the global ray count is then n * world_size and needs
        # neither a collective nor the host synchronisation of reading it back
        self.equal_shards = bool(equal_shards)
        self.get_resolution()
        self.get_encoding(config)
        self.get_decoder(config, rba_factory)
        self.count = 0
        self.clamp = False
        self._z_tables = None

    # ---- model/scene_rep.py:23-39 ---------------------------------------------------------------------------
    def get_resolution(self):
        bb = self.bounding_box
        dim_max = float((bb[:, 1] - bb[:, 0]).max())
        g = self.config["grid"]
        self.resolution_sdf = g["voxel_sdf"] if g["voxel_sdf"] > 10 else int(dim_max / g["voxel_sdf"])
        self.resolution_color = g["voxel_color"] if g["voxel_color"] > 10 else int(dim_max / g["voxel_color"])

    # ---- model/scene_rep.py:41-93 ---------------------------------------------------------------------------
    def get_encoding(self, config, GBV=True):
        self.embedpos_fn, self.input_ch_pos = get_encoder(config["pos"]["enc"], n_bins=config["pos"]["n_bins"])
        self.embed_res_fn, embed_out = get_encoder(config["grid"]["enc"], log2_hashmap_size=config["grid"]["hash_size"],
                                                   desired_resolution=self.resolution_sdf)
        self.input_ch = embed_out
        if GBV:
            self.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
            gv = config["globalV"]
            def dense(nf):
                e = Encoding(3, {"otype": "Grid", "type": "Dense", "n_levels": gv["n_levels"], "n_features_per_level": nf,
                                 "base_resolution": gv["base_resolution"], "per_level_scale": gv["per_level_scale"],
                                 "interpolation": "Linear"}, torch.float)
                e.requires_grad_(False)
                return e
            self.GBV = dense(gv["n_features_per_level"])
            self.GBW = dense(1)
            with torch.no_grad():
                self.GBW.params[:] = 0.0

    # ---- model/scene_rep.py:95-105 --------------------------------------------------------------------------
    def get_decoder(self, config, rba_factory=None):
        self.decoder_res = ColorSDFNet(config, input_ch=self.input_ch, input_ch_pos=self.input_ch_pos)
        if self.embed_res_fn.params.is_cuda:
            self.decoder_res.to(self.embed_res_fn.params.device)
        self.color_net_res = batchify(self.decoder_res.color_net, None)
        self.sdf_net_res = batchify(self.decoder_res.sdf_net, None)
        # The pose-residual MLP (model/rba.py) is outside the hot path; a caller that needs it passes its constructor.
        self.rba = rba_factory(self.num_kf, scale=config["mapping"]["pose_scale"]) if rba_factory is not None else None

    # ---- kernel-side configuration --------------------------------------------------------------------------
    def _ray_cfg(self, clamp=None) -> abi.RayCfg:
        c, t, cam = self.config, self.config["training"], self.config["cam"]
        cfg = abi.RayCfg()
        cfg.range_d, cfg.n_range_d, cfg.n_samples_d = t["range_d"], t["n_range_d"], t["n_samples_d"]
        cfg.near_, cfg.far_, cfg.perturb = cam["near"], cam["far"], 1 if t["perturb"] > 0. else 0
        cfg.c_trunc, cfg.trunc = t["c_trunc"], t["trunc"]
        cfg.clamp_mode = 1 if (self.clamp if clamp is None else clamp) else 0
        cfg.clamp_thr = c["mapping"].get("clamp", 1.0)
        cfg.sc_factor, cfg.depth_trunc, cfg.rgb_missing = c["data"]["sc_factor"], cam["depth_trunc"], t["rgb_missing"]
        d = c["decoder"]
        if d["hidden_dim"] != d["hidden_dim_color"] or d["num_layers"] != 2 or d["num_layers_color"] != 2:
            raise abi.RfError("fused decoder: built for num_layers = num_layers_color = 2 and equal hidden widths "
                              "(every shipped config; model/decoder.py defaults differ only in width)")
        cfg.hidden, cfg.n_bins, cfg.geo_feat = d["hidden_dim"], c["pos"]["n_bins"], d["geo_feat_dim"]
        # 1 (default): tcgen05 tensor-core decoder fed by feature planes; 0: fp32 SIMT decoder (accuracy anchor)
        cfg.mlp_precision = int(c.get("b200", {}).get("mlp_precision", 1))
        cfg.n_rays_total = 0
        bb = self.bounding_box.detach().cpu().numpy().astype(np.float64) if isinstance(self.bounding_box, torch.Tensor) \
            else np.asarray(self.bounding_box, dtype=np.float64)
        for a in range(3):
            cfg.bbox[2 * a], cfg.bbox[2 * a + 1] = float(bb[a, 0]), float(bb[a, 1])
        return cfg

    def _tables(self, device):
        """The three linspace tables of render_rays (:422-427), produced by torch.linspace on the CPU like the reference."""
        t, cam = self.config["training"], self.config["cam"]
        key = (t["range_d"], t["n_range_d"], t["n_samples_d"], cam["near"], cam["far"], str(device))
        if self._z_tables is None or self._z_tables[0] != key:
            parts = [torch.linspace(-t["range_d"], t["range_d"], steps=t["n_range_d"]),
                     torch.linspace(cam["near"], cam["far"], steps=t["n_range_d"])]
            if t["n_samples_d"] > 0:
                parts.append(torch.linspace(cam["near"], cam["far"], t["n_samples_d"]))
            self._z_tables = (key, torch.cat(parts).to(torch.float32).to(device))
        return self._z_tables[1]

    def _meta(self, with_losses):
        return {"cfg": self._ray_cfg(), "hash_desc": self.embed_res_fn.desc, "gbv_desc": self.GBV.desc,
                "with_losses": with_losses, "group": self.process_group, "equal_shards": self.equal_shards}

    def sample_z(self, target_d, n_rays, u=None):
        """z_vals [N,S] (model/scene_rep.py:417-441).  `u` injects the jitter (tests); otherwise it is drawn like the
        reference does — torch.rand on the CPU default generator (:441) — up to 2^20 samples, on the device beyond."""
        t = self.config["training"]
        dev = target_d.device
        S = t["n_range_d"] + t["n_samples_d"]
        cfg = self._ray_cfg()
        if cfg.perturb and u is None:
            # a pageable host-to-device copy cannot be captured in a CUDA graph: draw on the device while capturing
            on_host = n_rays * S <= (1 << 20) and not (dev.type == "cuda" and torch.cuda.is_current_stream_capturing())
            u = torch.rand(n_rays, S).to(dev) if on_host else torch.rand(n_rays, S, device=dev)
        if u is not None:
            u = u.to(dev, torch.float32).contiguous()
        z_vals = torch.empty(n_rays, S, dtype=torch.float32, device=dev)
        td = target_d.detach().reshape(-1).to(torch.float32).contiguous()
        rc = abi.lib().rf_ray_sample_z(C.byref(cfg), abi.dptr(td), abi.dptr(u), abi.dptr(self._tables(dev)), C.c_int64(n_rays),
                                       abi.dptr(z_vals), abi.stream_ptr())
        abi.check(rc, "rf_ray_sample_z")
        return z_vals

    def _weights(self):
        w = self.decoder_res.fused_weights()
        if w is None:
            raise abi.RfError("fused decoder needs the 2-layer SDF / colour nets")
        return w

    # batches at least this large are rendered valid-depth rays first (see _render)
    GROUP_RAYS_MIN = 1 << 16

    def _render(self, rays_o, rays_d, target_d, target_rgb, with_losses, u=None):
        n = rays_o.shape[0]
        if target_d is None:
            # no depth prior: training.n_samples depths evenly spaced over [near, far] (model/scene_rep.py:431-433), jittered
            # like the others (:437-441); the rest of the pipeline sees them as one group of S samples
            t, cam = self.config["training"], self.config["cam"]
            ns = int(t["n_samples"])
            z_vals = torch.linspace(cam["near"], cam["far"], ns).to(rays_o.device, torch.float32)[None, :].repeat(n, 1)
            if t["perturb"] > 0.:
                mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
                upper = torch.cat([mids, z_vals[..., -1:]], -1)
                lower = torch.cat([z_vals[..., :1], mids], -1)
                uu = torch.rand(z_vals.shape).to(z_vals) if u is None else u.to(z_vals)
                z_vals = lower + (upper - lower) * uu
            meta = self._meta(False)
            meta["cfg"].n_range_d, meta["cfg"].n_samples_d = ns, 0
            w = self._weights()
            rgb_map, depth_map, raw, losses = _RayQueryFn.apply(
                rays_o, rays_d, self.embed_res_fn.params, w[0], w[1], w[2], w[3], self.GBV.params, z_vals.contiguous(), None, None, meta)
            return rgb_map, depth_map, raw, z_vals, losses
        # Training batches are processed with the rays that have a depth measurement first, the others (sensor holes, no
        # return) last.  Every per-ray result is independent of the order; what changes is how the backward's 128-ray tiles
        # are composed: rays with a measurement stop contributing a few centimetres behind the surface (n_live, see
        # csrc/ray_encode.cu) at nearly the same sample index, rays without one carry gradient along their whole length, and
        # one such ray in a tile keeps the whole tile alive.  The permutation is undone on the per-ray outputs.
        inv = None
        if (with_losses and n >= self.GROUP_RAYS_MIN and rays_o.is_cuda and not torch.cuda.is_current_stream_capturing()):
            perm = torch.sort((target_d.detach().reshape(-1) <= 0).to(torch.uint8), stable=True).indices
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(n, device=perm.device)
            rays_o, rays_d = rays_o.index_select(0, perm), rays_d.index_select(0, perm)
            target_d, target_rgb = target_d.index_select(0, perm), target_rgb.index_select(0, perm)
            if u is not None:
                u = u.to(perm.device).index_select(0, perm)          # row r of the jitter belongs to ray r
        z_vals = self.sample_z(target_d, n, u)
        w = self._weights()
        rgb_map, depth_map, raw, losses = _RayQueryFn.apply(
            rays_o, rays_d, self.embed_res_fn.params, w[0], w[1], w[2], w[3], self.GBV.params, z_vals, target_d, target_rgb,
            self._meta(with_losses))
        if inv is not None:                                           # raw / z_vals are not handed out by mapping()
            rgb_map, depth_map = rgb_map.index_select(0, inv), depth_map.index_select(0, inv)
        return rgb_map, depth_map, raw, z_vals, losses

    # ---- model/scene_rep.py:407-456 -------------------------------------------------------------------------
    def render_rays(self, rays_o, rays_d, target_d=None, tracking=False, frameid=None, render_flag=False, u=None):
        rgb_map, depth_map, raw, z_vals, _ = self._render(rays_o, rays_d, target_d, None, False, u)
        return {"rgb_res_map": rgb_map, "depth_res_map": depth_map, "z_vals": z_vals, "raw": raw}

    # ---- model/scene_rep.py:460-529 -------------------------------------------------------------------------
    def mapping(self, rays_o, rays_d, target_rgb, target_d, tracking=False, render_flag=False, clamp=False, u=None):
        self.clamp = clamp
        if not self.training:
            return self.render_rays(rays_o, rays_d, target_d=target_d, tracking=tracking, render_flag=render_flag, u=u)
        rgb_map, depth_map, raw, z_vals, losses = self._render(rays_o, rays_d, target_d, target_rgb, True, u)
        return {"rgb_res_loss": losses[0], "depth_res_loss": losses[1], "sdf_res_loss": losses[2], "fs_res_loss": losses[3],
                "rgb_res": rgb_map, "depth_res": depth_map}

    # ---- point queries (model/scene_rep.py:212-349, 370-402) ------------------------------------------------
    def _point_query(self, inputs_flat, variant):
        x = inputs_flat.detach().to(torch.float32).contiguous()
        n = x.shape[0]
        raw = torch.empty(n, 4, dtype=torch.float32, device=x.device)
        w = self._weights()
        p = abi.RayParams(abi.dptr(self.embed_res_fn.params.detach()), abi.dptr(self.GBV.params.detach()),
                          abi.dptr(w[0].detach()), abi.dptr(w[1].detach()), abi.dptr(w[2].detach()), abi.dptr(w[3].detach()))
        cfg = self._ray_cfg()
        nws = int(abi.lib().rf_point_workspace_floats(C.byref(cfg), C.byref(self.embed_res_fn.desc), C.c_int64(n)))
        ws = torch.empty(nws, dtype=torch.float32, device=x.device) if nws > 0 else None
        rc = abi.lib().rf_point_query_forward(C.byref(cfg), C.byref(self.embed_res_fn.desc), C.byref(self.GBV.desc), C.byref(p),
                                              abi.dptr(x), C.c_int64(n), C.c_int(variant), abi.dptr(raw), abi.dptr(ws),
                                              abi.stream_ptr())
        abi.check(rc, "rf_point_query_forward")
        return raw

    def query_color_sdf(self, query_points, ranged_mask=None):
        """[.., 3] normalised coords -> raw [N,4] (model/scene_rep.py:314-349).  Forward only; gradients flow through
        render_rays / mapping."""
        return self._point_query(torch.reshape(query_points, [-1, query_points.shape[-1]]), 0)

    def run_network(self, inputs, flat=False):
        """model/scene_rep.py:370-402 (world-space points; float64 normalisation by the bounding box)."""
        inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
        if self.config["grid"]["tcnn_encoding"]:
            bb = self.bounding_box.to(inputs_flat.device)
            inputs_flat = (inputs_flat - bb[:, 0]) / (bb[:, 1] - bb[:, 0])
        out = self.query_color_sdf(inputs_flat)
        return out if flat else torch.reshape(out, list(inputs.shape[:-1]) + [out.shape[-1]])

    def query_sdf_res(self, query_points, return_geo=False, embed=False):
        """model/scene_rep.py:212-248.  embed=True returns the raw hash features (differentiable: SLAM.smoothness)."""
        inputs_flat = torch.reshape(query_points, [-1, query_points.shape[-1]])
        if embed:
            embedded = self.embed_res_fn(inputs_flat)
            return torch.reshape(embedded, list(query_points.shape[:-1]) + [embedded.shape[-1]])
        if return_geo:
            embedded = self.embed_res_fn(inputs_flat)
            embedded_pos = self.embedpos_fn(inputs_flat)
            ex = self.GBV(inputs_flat)
            t = torch.clamp(ex[..., 0] * self.config["training"]["c_trunc"] / self.config["training"]["trunc"], -1, 1)
            out = self.sdf_net_res(torch.cat([embedded, embedded_pos, t.unsqueeze(-1)], dim=-1))
            sdf = torch.reshape(out[..., 0] + t, list(query_points.shape[:-1]))
            return sdf, torch.reshape(out[..., 1:], list(query_points.shape[:-1]) + [out.shape[-1] - 1])
        raw = self._point_query(inputs_flat, 1)
        return torch.reshape(raw[..., 3], list(query_points.shape[:-1]))

    def query_sdf_ex(self, query_points, return_geo=False, embed=False):
        inputs_flat = torch.reshape(query_points, [-1, query_points.shape[-1]])
        return torch.reshape(self.GBV(inputs_flat)[..., 0], list(query_points.shape[:-1]))

    def query_w_res(self, query_points, return_geo=False, embed=False):
        inputs_flat = torch.reshape(query_points, [-1, query_points.shape[-1]])
        return torch.reshape(self.GBW(inputs_flat), list(query_points.shape[:-1]))

    def query_color_residual(self, query_points):
        return self._point_query(torch.reshape(query_points, [-1, query_points.shape[-1]]), 2)[..., :3]

    def query_color_ex(self, query_points):
        inputs_flat = torch.reshape(query_points, [-1, query_points.shape[-1]])
        return self.GBV(inputs_flat)[..., 1:]

    # ---- model/scene_rep.py:156-179 (and :107-127 inside the kernel) ------------------------------------------
    def raw2outputs(self, raw, z_vals):
        """raw [N,S,4], z_vals [N,S] -> (rgb_map [N,3], depth_map [N]): the compositing kernel alone (forward only; gradients
        flow through render_rays / mapping)."""
        raw = raw.detach().to(torch.float32).contiguous(); z = z_vals.detach().to(torch.float32).contiguous()
        n = raw.shape[0]
        rgb_map = torch.empty(n, 3, dtype=torch.float32, device=raw.device)
        depth_map = torch.empty(n, dtype=torch.float32, device=raw.device)
        cfg = self._ray_cfg()
        cfg.n_range_d, cfg.n_samples_d = int(z.shape[1]), 0          # samples per ray as given (one group)
        abi.check(abi.lib().rf_ray_composite(C.byref(cfg), abi.dptr(raw), abi.dptr(z), C.c_int64(n), abi.dptr(rgb_map), abi.dptr(depth_map),
                                             abi.stream_ptr()), "rf_ray_composite")
        return rgb_map, depth_map
