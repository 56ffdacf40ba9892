"""N4 (second half), GPU: ``remixfusion_b200.lattice.marching_cubes`` (csrc/marching_cubes.cu) against the reference's own C++
marching cubes — its golden outputs (tests/golden/mc_golden.npz) and, live, the compiled reference (oracle/_ref/libmc_ref.so)."""
import os

import numpy as np
import pytest
import torch

from oracle import mc_oracle
from tests._mc_common import mc_volumes, same_surface

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "mc_golden.npz"))
VOLS = mc_volumes()


@pytest.mark.parametrize("name", list(VOLS))
def test_marching_cubes_matches_reference(cuda, rf_lib, name):
    from remixfusion_b200.lattice import marching_cubes
    vol, iso, trunc = VOLS[name]
    V, F = marching_cubes(torch.from_numpy(vol).to(cuda), iso, trunc)
    V, F = V.cpu().numpy(), F.cpu().numpy()
    ok, msg = same_surface(V, F, G[f"{name}_V"], G[f"{name}_F"])
    assert ok, f"{name}: {msg}"
    if F.size:
        assert F.max() < V.shape[0] and F.min() == 0 and len(np.unique(F)) == V.shape[0]      # every vertex is referenced
        # the welded mesh is as compact as the reference's tolerance-merged one (closed surfaces: V - E + F = 2 per component)
        assert V.shape[0] <= G[f"{name}_V"].shape[0] * 1.02 + 2, (V.shape[0], G[f"{name}_V"].shape[0])
    else:
        assert G[f"{name}_F"].shape[0] == 0


def test_mask_and_larger_volume_vs_compiled_reference(cuda, rf_lib):
    """A 96 x 80 x 72 volume with a weight mask (utils.py:161-170), against the compiled reference run live."""
    from remixfusion_b200.lattice import marching_cubes
    assert mc_oracle.available(), "oracle/_ref/libmc_ref.so must travel with the snapshot (python -c 'import __graft_entry__ as g; g.build()')"
    g = np.stack(np.meshgrid(np.arange(96), np.arange(80), np.arange(72), indexing="ij"), -1).astype(np.float64)
    f = np.minimum(np.linalg.norm(g - [40, 41, 33], axis=-1) - 21.5, np.linalg.norm(g - [62, 30, 40], axis=-1) - 17.25)
    f += 1.5 * np.sin(g[..., 0] * 0.31) * np.cos(g[..., 1] * 0.23 + g[..., 2] * 0.4)
    vol = np.clip(f / 4.0, -1.0, 1.0).astype(np.float32)
    mask = (g[..., 2] > 10) & ~((g[..., 0] > 70) & (g[..., 1] < 20))
    V, F = marching_cubes(torch.from_numpy(vol).to(cuda), 0.0, 3.0, mask=torch.from_numpy(mask).to(cuda))
    ref_vol = np.where(mask, vol, np.nan).astype(np.float32)
    Vr, Fr = mc_oracle.marching_cubes(ref_vol, 0.0, 3.0)
    ok, msg = same_surface(V.cpu().numpy(), F.cpu().numpy(), Vr, Fr)
    assert ok and Fr.shape[0] > 10000, msg
