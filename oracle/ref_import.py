"""TEST INFRASTRUCTURE ONLY, CONTAINER ONLY — runs the reference's OWN Stage-2 code on CPU.

Imports model/scene_rep.py, model/decoder.py, model/utils.py from /root/reference (read-only, never copied) with
the two absent third-party modules stubbed in ``sys.modules``:
  * ``tinycudann``  -> oracle/tcnn_standin.py (pure-PyTorch HashGrid / Dense grid / OneBlob; parity unpinned there)
  * ``kornia.geometry.conversions`` -> placeholders (only model/rba.py, outside the hot path, uses them)
and builds a ``JointEncoding`` without the pose-residual MLP (``RBA.__init__`` hard-codes ``.cuda()``,
model/rba.py:45-47; it is not on the hot path).  Everything else — get_resolution, get_encoding, the decoder,
render_rays, run_network, query_color_sdf, raw2outputs, sdf2weights, mapping and the losses — is the reference's code.

/root/reference does not exist on the GPU box, so nothing that runs there imports this file.  It is used by
tests/golden/make_ray_golden.py (fixtures) and by tests that are skipped when the reference tree is absent.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF_ROOT = os.environ.get("RF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


_loaded = {}


def load():
    """Returns the reference modules dict {scene_rep, decoder, utils, encodings}."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"{REF_ROOT} is not present")
    from oracle import tcnn_standin
    sys.modules.setdefault("tinycudann", tcnn_standin.as_module())
    if "kornia" not in sys.modules:
        k = types.ModuleType("kornia"); kg = types.ModuleType("kornia.geometry"); kc = types.ModuleType("kornia.geometry.conversions")
        kc.angle_axis_to_rotation_matrix = lambda *a, **kw: (_ for _ in ()).throw(NotImplementedError("kornia stub"))
        kc.rotation_matrix_to_angle_axis = lambda *a, **kw: (_ for _ in ()).throw(NotImplementedError("kornia stub"))
        k.geometry = kg; kg.conversions = kc
        sys.modules["kornia"] = k; sys.modules["kornia.geometry"] = kg; sys.modules["kornia.geometry.conversions"] = kc
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    _loaded["scene_rep"] = importlib.import_module("model.scene_rep")
    _loaded["decoder"] = importlib.import_module("model.decoder")
    _loaded["utils"] = importlib.import_module("model.utils")
    _loaded["encodings"] = importlib.import_module("model.encodings")
    return _loaded


def make_reference_model(config, bound_box):
    """``JointEncoding(config, bound_box)`` minus RBA, on CPU, using the reference's own methods."""
    m = load()
    JE = m["scene_rep"].JointEncoding
    obj = JE.__new__(JE)
    nn.Module.__init__(obj)
    obj.config = config
    obj.bounding_box = bound_box              # float64 tensor in the reference (run.py:90)
    obj.num_kf = None
    obj.get_resolution()                      # model/scene_rep.py:23-39
    obj.get_encoding(config)                  # model/scene_rep.py:41-93
    # model/scene_rep.py:95-105 without `self.rba = RBA(...)`
    obj.decoder_res = m["decoder"].ColorSDFNet(config, input_ch=obj.input_ch, input_ch_pos=obj.input_ch_pos)
    obj.color_net_res = m["utils"].batchify(obj.decoder_res.color_net, None)
    obj.sdf_net_res = m["utils"].batchify(obj.decoder_res.sdf_net, None)
    obj.count = 0
    obj.clamp = False
    return obj
