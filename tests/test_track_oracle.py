"""CPU: the NumPy tracker oracle (oracle/track_oracle.py) against golden outputs of the literal reference kernels
(tests/golden/track_golden.npz, generated on a B200 by tests/golden/make_track_golden.py)."""
import os

import numpy as np
import pytest

from oracle import track_oracle as TO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "track_golden.npz")


@pytest.fixture(scope="module")
def G():
    if not os.path.exists(GOLD):
        pytest.skip("track_golden.npz not generated yet")
    return np.load(GOLD)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_vertex_and_normal_maps(G, tag):
    v = TO.vertex_map(G["depth"], G["K"], 6.0, 0.06, G[f"{tag}_row_sample"], float(G[f"{tag}_sample_range"]))
    # back-projection: (pj - cx) * z / fx — products / quotients rounded one by one in both: bit-exact
    np.testing.assert_array_equal(v, G[f"{tag}_vertex"])
    n = TO.normal_map(G[f"{tag}_vertex"])
    ref = G[f"{tag}_normal"]
    assert np.array_equal(n == 0, ref == 0)
    np.testing.assert_allclose(n, ref, rtol=0, atol=2e-6)       # emulated FMA: <= 1 ulp of a unit vector component


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("level,li", [(8, 3), (4, 1)])
def test_fitness(G, tag, level, li):
    val, cnt = TO.fitness(G["tsdf"], G["dims"], G["origin"], float(G["voxel"]), G[f"{tag}_vertex"], G[f"{tag}_normal"], G["K"],
                          G["R"], G["T"], G["cand"], G["ss"], level, li)
    ref_v, ref_c = G[f"{tag}_value_l{level}"], G[f"{tag}_count_l{level}"]
    assert float(ref_c.sum()) > 1000
    # double-rounded FMA emulation can move a vertex across a voxel / pixel boundary for a handful of pairs
    assert (cnt != ref_c).mean() <= 0.01 and np.abs(cnt - ref_c).max() <= 2
    close = np.isclose(val, ref_v, rtol=2e-5, atol=1e-5)
    assert close.mean() >= 0.97, f"{(~close).sum()} of {close.size} candidates differ"
    assert np.abs(val - ref_v).max() <= 4.0                      # a flipped voxel changes one term by < 2 (|tsdf - gt| <= 2)
