"""GPU-resident keyframe ray store + sampler — drop-in for ``KeyFrameDatabase`` (model/keyframe.py:5-96).

The reference keeps ``rays [num_kf, num_rays_to_save, 7]`` (direction 3, rgb 3, depth 1) in host memory, draws indices with
Python's ``random.sample`` and copies the sampled rays to the GPU in every mapping iteration (mp_slam/mapper.py:394-409).
Here the store lives on the device and indices are drawn on the device (uniform, without replacement, like
``random.sample``), so an iteration needs no host work and can feed ``GraphedMappingStep`` directly.  Same constructor,
attributes and methods; ``idxs=`` injects the indices (parity tests: the same indices give the same rays as the reference).
"""
from __future__ import annotations

import torch


class KeyFrameDatabase(object):
    def __init__(self, config, H, W, num_kf, num_rays_to_save, device, num_frame=None) -> None:
        self.config = config
        self.keyframes = {}
        self.device = torch.device(device)
        self.rays = torch.zeros((num_kf, num_rays_to_save, 7), device=self.device)      # model/keyframe.py:10 (host there)
        self.num_rays_to_save = num_rays_to_save
        self.frame_ids = None
        self.H = H
        self.W = W
        self.kf_poses = torch.zeros((num_kf, 4, 4))
        self.kf_fuse_poses = torch.zeros((num_kf, 4, 4))
        self.kf_error = torch.zeros((num_kf), device=self.device)
        self.kf_error_cnt = torch.zeros((num_kf), device=self.device)
        if num_frame is not None:
            self.all_fuse_pose = torch.zeros((num_frame, 4, 4), device=self.device)

    def __len__(self):
        return len(self.frame_ids)

    def get_length(self):
        return self.__len__()

    def _draw(self, n, k):
        """k distinct indices in [0, n), uniform — what random.sample(range(n), k) returns, drawn on the device."""
        return torch.randperm(n, device=self.device)[:k]

    def sample_single_keyframe_rays(self, rays, option="random", first=False, idxs=None):
        """model/keyframe.py:28-49.  rays: [1, H*W, 7]."""
        rays = rays.to(self.device)
        if option == "random":
            if idxs is None:
                idxs = self._draw(self.H * self.W, self.num_rays_to_save)
        elif option == "filter_depth":
            valid = (rays[..., -1] > 0.0) & (rays[..., -1] <= self.config["cam"]["depth_trunc"])
            rays_valid = rays[valid, :]
            if len(rays_valid) > self.num_rays_to_save:
                if idxs is None:
                    idxs = self._draw(len(rays_valid), self.num_rays_to_save)
            else:
                if idxs is None:
                    idxs = self._draw(self.H * self.W, self.num_rays_to_save)
                # the reference writes `option == "random"` here (a comparison, not an assignment): option stays
                # 'filter_depth' and the indices are applied to rays_valid below unless `first`
        else:
            raise NotImplementedError()
        idxs = torch.as_tensor(idxs, device=self.device, dtype=torch.long)
        if option == "random" or first:
            return rays[:, idxs]
        return rays_valid[idxs, :]

    def attach_ids(self, frame_ids):
        frame_ids = frame_ids.to(self.device)
        self.frame_ids = frame_ids if self.frame_ids is None else torch.cat([self.frame_ids, frame_ids], dim=0)

    def add_keyframe(self, batch, filter_depth=False, idxs=None):
        """model/keyframe.py:60-82.  batch: 'direction' [1,H,W,3], 'rgb' [1,H,W,3], 'depth' [1,H,W], 'frame_id'."""
        first = bool(batch["frame_id"] == 0)
        rays = torch.cat([batch["direction"].to(self.device), batch["rgb"].to(self.device),
                          batch["depth"].to(self.device)[..., None]], dim=-1)
        rays = rays.reshape(1, -1, rays.shape[-1])
        rays = self.sample_single_keyframe_rays(rays, "filter_depth" if filter_depth else "random", first=first, idxs=idxs)
        fid = batch["frame_id"]
        if not isinstance(fid, torch.Tensor):
            fid = torch.tensor([fid])
        self.attach_ids(fid.reshape(-1))
        self.rays[len(self.frame_ids) - 1] = rays

    def sample_global_rays(self, bs, idxs=None):
        """model/keyframe.py:84-96: bs rays over all stored keyframes and the frame id of each."""
        num_kf = self.__len__()
        if idxs is None:
            idxs = self._draw(num_kf * self.num_rays_to_save, bs)
        idxs = torch.as_tensor(idxs, device=self.device, dtype=torch.long)
        sample_rays = self.rays[:num_kf].reshape(-1, 7)[idxs]
        frame_ids = self.frame_ids[torch.div(idxs, self.num_rays_to_save, rounding_mode="floor")]
        return sample_rays, frame_ids
