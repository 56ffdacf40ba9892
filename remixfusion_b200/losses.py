"""Loss terms of the mapping iteration that live outside ``JointEncoding.mapping`` in the reference: the total-loss mix and the
feature-grid smoothness term of ``SLAM`` (mp_slam/slam.py:145-217).  Everything runs on the device with no host
synchronisation, so an iteration that includes the smoothness term can be captured in a CUDA graph
(``remixfusion_b200.graph.GraphedMappingStep``): while a stream is capturing, the two random offsets are drawn with the device
generator (a pageable host-to-device copy of ``torch.rand`` on the CPU, which is what the reference does, cannot be captured)."""
from __future__ import annotations

import torch


class Smoothness:
    """``SLAM.smoothness(sample_points, voxel_size, margin)`` (mp_slam/slam.py:193-217): total variation of the hash features
    over a randomly placed (sample_points - 1)^3 lattice of spacing ``voxel_size``; gradients flow into the hash table through
    ``query_sdf_res(embed=True)`` (the differentiable grid encoder, csrc/encoders.cu)."""

    def __init__(self, model, sample_points=256, voxel_size=0.1, margin=0.05):
        self.model, self.sample_points, self.voxel_size, self.margin = model, int(sample_points), float(voxel_size), float(margin)
        dev = model.embed_res_fn.params.device
        n = self.sample_points - 1
        ax = torch.arange(0, n, dtype=torch.long, device=dev)
        self.coords = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).float()        # utils.coordinates(n, flatten=False)
        bb = model.bounding_box.to(dev) if isinstance(model.bounding_box, torch.Tensor) else torch.as_tensor(model.bounding_box, device=dev)
        self.bb = bb.to(torch.float64)

    def __call__(self, rand_offset=None, rand_shift=None):
        """rand_offset [3], rand_shift [1,1,1,3] in [0,1): the two ``torch.rand`` draws of the reference (tests inject them)."""
        bb = self.bb
        volume = bb[:, 1] - bb[:, 0]
        grid_size = (self.sample_points - 1) * self.voxel_size
        offset_max = volume - grid_size - 2 * self.margin
        capturing = bb.is_cuda and torch.cuda.is_current_stream_capturing()
        if rand_offset is None:
            rand_offset = torch.rand(3, device=bb.device) if capturing else torch.rand(3).to(bb.device)
        if rand_shift is None:
            rand_shift = torch.rand((1, 1, 1, 3), device=bb.device) if capturing else torch.rand((1, 1, 1, 3)).to(bb.device)
        offset = rand_offset.to(offset_max) * offset_max + self.margin
        pts = (self.coords.to(volume) + rand_shift.to(volume)) * self.voxel_size + bb[:, 0] + offset
        if self.model.config["grid"]["tcnn_encoding"]:
            pts = (pts - bb[:, 0]) / (bb[:, 1] - bb[:, 0])
        feat = self.model.query_sdf_res(pts, embed=True)
        tv_x = torch.pow(feat[1:, ...] - feat[:-1, ...], 2).sum()
        tv_y = torch.pow(feat[:, 1:, ...] - feat[:, :-1, ...], 2).sum()
        tv_z = torch.pow(feat[:, :, 1:, ...] - feat[:, :, :-1, ...], 2).sum()
        return (tv_x + tv_y + tv_z) / (self.sample_points ** 3)


def make_loss_fn(config, model, smooth=True):
    """``SLAM.get_loss_from_ret(ret, smooth=...)`` (mp_slam/slam.py:145-190) as a closure over a model: the weighted sum of the
    four mapping losses, plus ``smooth_weight`` x smoothness when the config asks for it."""
    t = config["training"]
    sm = None
    if smooth and t.get("smooth_weight", 0) > 0:
        sm = Smoothness(model, t["smooth_pts"], t["smooth_vox"], t["smooth_margin"])

    def loss_fn(ret):
        loss = (t["rgb_weight"] * ret["rgb_res_loss"] + t["depth_weight"] * ret["depth_res_loss"]
                + t["sdf_weight"] * ret["sdf_res_loss"] + t["fs_weight"] * ret["fs_res_loss"])
        if sm is not None:
            loss = loss + t["smooth_weight"] * sm()
        return loss
    loss_fn.smoothness = sm
    return loss_fn
