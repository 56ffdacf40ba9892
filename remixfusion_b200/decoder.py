"""Decoder MLPs — same classes, constructor arguments, parameter names and shapes as model/decoder.py (torch branch;
`decoder.tcnn_network` is False in every shipped config), so checkpoints interchange: ``sdf_net.model.{0,2}.weight``,
``color_net.model.{0,2}.weight``, bias-free Linear + ReLU (model/decoder.py:37-53, :92-110).

These modules own the weights.  In the fused ray path (scene_rep.JointEncoding.render_rays / mapping) the kernels
read the weight tensors directly; ``forward`` here is the stand-alone module call the reference also allows
(e.g. ``sdf_net_res(torch.cat([...]))``, model/scene_rep.py:235) and runs as plain torch ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _mlp(in_dim, hidden, out_dim, num_layers):
    layers = []
    for l in range(num_layers):
        i = in_dim if l == 0 else hidden
        o = out_dim if l == num_layers - 1 else hidden
        layers.append(nn.Linear(i, o, bias=False))
        if l != num_layers - 1:
            layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*nn.ModuleList(layers))


class ColorNet(nn.Module):
    def __init__(self, config, input_ch=4, geo_feat_dim=15, hidden_dim_color=64, num_layers_color=3):
        super().__init__()
        self.config = config
        self.input_ch, self.geo_feat_dim = input_ch, geo_feat_dim
        self.hidden_dim_color, self.num_layers_color = hidden_dim_color, num_layers_color
        if config["decoder"].get("tcnn_network", False):
            raise NotImplementedError("decoder.tcnn_network: True (tcnn FullyFusedMLP) is disabled in every shipped config")
        self.model = _mlp(input_ch + geo_feat_dim, hidden_dim_color, 3, num_layers_color)

    def forward(self, input_feat):
        return self.model(input_feat)


class SDFNet(nn.Module):
    def __init__(self, config, input_ch=3, geo_feat_dim=15, hidden_dim=64, num_layers=2):
        super().__init__()
        self.config = config
        self.input_ch, self.geo_feat_dim, self.hidden_dim, self.num_layers = input_ch, geo_feat_dim, hidden_dim, num_layers
        if config["decoder"].get("tcnn_network", False):
            raise NotImplementedError("decoder.tcnn_network: True (tcnn FullyFusedMLP) is disabled in every shipped config")
        self.model = _mlp(input_ch, hidden_dim, 1 + geo_feat_dim, num_layers)

    def forward(self, x, return_geo=True):
        out = self.model(x)
        return out if return_geo else out[..., :1]


class ColorSDFNet(nn.Module):
    """model/decoder.py:116-146."""

    def __init__(self, config, input_ch=3, input_ch_pos=12):
        super().__init__()
        self.config = config
        d = config["decoder"]
        self.color_net = ColorNet(config, input_ch=input_ch_pos + 3, geo_feat_dim=d["geo_feat_dim"],
                                  hidden_dim_color=d["hidden_dim_color"], num_layers_color=d["num_layers_color"])
        self.sdf_net = SDFNet(config, input_ch=input_ch + input_ch_pos + 1, geo_feat_dim=d["geo_feat_dim"],
                              hidden_dim=d["hidden_dim"], num_layers=d["num_layers"])

    def forward(self, embed, embed_pos, ex_tsdf, ex_rgb):
        if embed_pos is not None:
            h = self.sdf_net(torch.cat([embed, embed_pos, ex_tsdf], dim=-1), return_geo=True)
        else:
            h = self.sdf_net(embed, return_geo=True)
        sdf, geo_feat = h[..., :1], h[..., 1:]
        if embed_pos is not None:
            rgb = self.color_net(torch.cat([embed_pos, geo_feat, ex_rgb], dim=-1))
        else:
            rgb = self.color_net(torch.cat([geo_feat], dim=-1))
        return torch.cat([rgb, sdf], -1)

    def fused_weights(self):
        """The four weight matrices the fused kernels consume, or None if the architecture is not the 2-layer one."""
        s, c = self.sdf_net.model, self.color_net.model
        if len(s) != 3 or len(c) != 3:
            return None
        return s[0].weight, s[2].weight, c[0].weight, c[2].weight
