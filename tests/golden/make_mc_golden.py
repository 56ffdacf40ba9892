"""Generates tests/golden/mc_golden.npz: outputs of the REFERENCE's own C++ marching cubes (oracle/_ref/libmc_ref.so, compiled
from /root/reference/thirdparty/NumpyMarchingCubes by oracle/build_ref.py) on seeded volumes.
Run:  python tests/golden/make_mc_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import mc_oracle                 # noqa: E402
from tests._mc_common import mc_volumes      # noqa: E402

out = {}
for name, (vol, iso, trunc) in mc_volumes().items():
    V, F = mc_oracle.marching_cubes(vol, iso, trunc)
    out[f"{name}_V"] = V; out[f"{name}_F"] = F.astype(np.int64)
    print(name, vol.shape, V.shape, F.shape)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_golden.npz"), **out)
