"""Device-resident keyframe ray store and sampler.

What it replaces: the reference's keyframe database (model/keyframe.py:5-96) keeps the per-keyframe ray subsets in host
memory, draws indices with Python's ``random.sample`` and ships the sampled rays to the GPU on every mapping iteration
(mp_slam/mapper.py:394-409).  Here one preallocated device buffer holds ``[capacity, rays_per_keyframe, 7]`` rows
(direction xyz, colour rgb, depth), frame ids live next to it on the device, and draws are device-side permutations — so an
iteration does no host work and can feed ``GraphedMappingStep`` directly.

Public surface kept for the callers in mp_slam/mapper.py: ``KeyFrameDatabase(config, H, W, num_kf, num_rays_to_save, device)``,
``add_keyframe(batch, filter_depth)``, ``sample_global_rays(bs)``, ``sample_single_keyframe_rays``, ``attach_ids``,
``get_length`` / ``len()``, attributes ``rays``, ``frame_ids``, ``kf_poses``, ``kf_fuse_poses``, ``kf_error``,
``kf_error_cnt``, ``all_fuse_pose``, ``num_rays_to_save``.  Every draw can be overridden with ``idxs=`` (parity tests feed the
indices the reference drew and require identical rays).
"""
from __future__ import annotations

import torch

_ROW = 7    # direction (3) | colour (3) | depth (1)


def _distinct(n: int, k: int, device) -> torch.Tensor:
    """k distinct integers from [0, n), uniformly — the distribution of random.sample(range(n), k) — drawn on `device`."""
    return torch.randperm(int(n), device=device)[: int(k)]


class KeyFrameDatabase:
    def __init__(self, config, H, W, num_kf, num_rays_to_save, device, num_frame=None):
        dev = torch.device(device)
        self.config, self.H, self.W, self.device = config, int(H), int(W), dev
        self.num_rays_to_save = int(num_rays_to_save)
        self.keyframes = {}
        # ray rows of all keyframes, one slab per keyframe, and the id of the frame each slab came from
        self.rays = torch.zeros(int(num_kf), self.num_rays_to_save, _ROW, device=dev)
        self._ids = torch.zeros(int(num_kf), dtype=torch.long, device=dev)
        self._count = 0
        # bookkeeping the mapper reads / writes (host poses as in the reference, error statistics on the device)
        self.kf_poses = torch.zeros(int(num_kf), 4, 4)
        self.kf_fuse_poses = torch.zeros(int(num_kf), 4, 4)
        self.kf_error = torch.zeros(int(num_kf), device=dev)
        self.kf_error_cnt = torch.zeros(int(num_kf), device=dev)
        self.all_fuse_pose = torch.zeros(int(num_frame), 4, 4, device=dev) if num_frame is not None else None

    # ---- size / ids -------------------------------------------------------------------------------------
    @property
    def frame_ids(self):
        return self._ids[: self._count] if self._count else None

    def __len__(self):
        return self._count

    def get_length(self):
        return self._count

    def attach_ids(self, frame_ids):
        ids = torch.as_tensor(frame_ids, dtype=torch.long, device=self.device).reshape(-1)
        self._ids[self._count: self._count + ids.numel()] = ids
        self._count += int(ids.numel())

    # ---- choosing the pixels of one keyframe (model/keyframe.py:28-49) -----------------------------------
    def _valid_depth(self, rows):
        d = rows[..., -1]
        return (d > 0.0) & (d <= self.config["cam"]["depth_trunc"])

    def sample_single_keyframe_rays(self, rays, option="random", first=False, idxs=None):
        """rays: [1, H*W, 7].  'random': num_rays_to_save pixels of the frame; 'filter_depth': of its valid-depth pixels
        when there are more than num_rays_to_save of them (otherwise pixel indices are drawn over the whole frame, as the
        reference does).  The first keyframe always indexes the whole frame."""
        rows = rays.to(self.device)
        take = lambda n: torch.as_tensor(idxs, dtype=torch.long, device=self.device) if idxs is not None \
            else _distinct(n, self.num_rays_to_save, self.device)
        if option == "random":
            return rows[:, take(self.H * self.W)]
        if option != "filter_depth":
            raise NotImplementedError(option)
        good = rows[self._valid_depth(rows), :]
        chosen = take(good.shape[0] if good.shape[0] > self.num_rays_to_save else self.H * self.W)
        return rows[:, chosen] if first else good[chosen, :]

    def add_keyframe(self, batch, filter_depth=False, idxs=None):
        """batch: 'direction' [1,H,W,3], 'rgb' [1,H,W,3], 'depth' [1,H,W], 'frame_id' (model/keyframe.py:60-82)."""
        fid = batch["frame_id"]
        fid = fid.reshape(-1) if isinstance(fid, torch.Tensor) else torch.tensor([int(fid)])
        rows = torch.cat([batch["direction"].to(self.device), batch["rgb"].to(self.device),
                          batch["depth"].to(self.device).unsqueeze(-1)], dim=-1).reshape(1, -1, _ROW)
        slab = self.sample_single_keyframe_rays(rows, "filter_depth" if filter_depth else "random",
                                                first=bool(int(fid[0]) == 0), idxs=idxs)
        self.attach_ids(fid)
        self.rays[self._count - 1] = slab.reshape(self.num_rays_to_save, _ROW)

    # ---- drawing a training batch over all keyframes (model/keyframe.py:84-96) ---------------------------
    def sample_global_rays(self, bs, idxs=None):
        """-> (rays [bs,7], frame id of each ray [bs])."""
        total = self._count * self.num_rays_to_save
        pick = torch.as_tensor(idxs, dtype=torch.long, device=self.device) if idxs is not None else _distinct(total, bs, self.device)
        flat = self.rays[: self._count].reshape(total, _ROW)
        return flat[pick], self._ids[torch.div(pick, self.num_rays_to_save, rounding_mode="floor")]
