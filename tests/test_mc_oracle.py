"""CPU suite for N4 (second half): the marching-cubes formulation of csrc/marching_cubes.cu — restated in NumPy by
oracle/mc_port.py from the constants inside that .cu file — against the reference's own compiled C++ (oracle/_ref/libmc_ref.so,
built from thirdparty/NumpyMarchingCubes by oracle/build_ref.py) and the golden outputs it produced (tests/golden/mc_golden.npz)."""
import os
import re

import numpy as np
import pytest

from oracle import mc_oracle, mc_port
from tests._mc_common import mc_volumes, same_surface

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "mc_golden.npz"))
VOLS = mc_volumes()


@pytest.mark.parametrize("name", list(VOLS))
def test_port_matches_reference_golden(name):
    vol, iso, trunc = VOLS[name]
    soup = mc_port.triangle_soup(vol, iso, trunc)
    V = soup.reshape(-1, 3); F = np.arange(V.shape[0]).reshape(-1, 3)
    ok, msg = same_surface(V, F, G[f"{name}_V"], G[f"{name}_F"])
    assert ok, f"{name}: {msg}"
    assert (name == "empty") == (soup.shape[0] == 0)


def test_compiled_reference_reproduces_its_golden():
    """Where oracle/_ref/libmc_ref.so is present (the build container, and the GPU box: it travels with the snapshot)."""
    if not mc_oracle.available():
        pytest.skip("oracle/_ref/libmc_ref.so not built")
    for name, (vol, iso, trunc) in VOLS.items():
        V, F = mc_oracle.marching_cubes(vol, iso, trunc)
        assert np.array_equal(V, G[f"{name}_V"]) and np.array_equal(F.astype(np.int64), G[f"{name}_F"]), name


def test_case_table_is_the_reference_table():
    """The nibble-packed case table in csrc/marching_cubes.cu equals triTable of the reference's tables.h, and the reference's
    edgeTable is the union of the edges each case uses (which is how the kernel derives it).  Only where /root/reference exists."""
    path = "/root/reference/thirdparty/NumpyMarchingCubes/marching_cubes/src/tables.h"
    if not os.path.exists(path):
        pytest.skip("reference tree absent")
    src = open(path).read()
    rows = re.findall(r"\{([^{}]*)\}", re.search(r"triTable\s*\[256\]\[16\]\s*=\s*\{(.*?)\};", src, re.S).group(1))
    tri_ref = [[int(x) for x in r.split(",") if x.strip()] for r in rows]
    edge_ref = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", re.search(r"edgeTable\s*\[256\]\s*=\s*\{(.*?)\};", src, re.S).group(1))]
    tri, *_ = mc_port.kernel_constants()
    for c in range(256):
        unpacked = [(tri[c] >> (4 * i)) & 0xF for i in range(16)]
        assert [(-1 if e == 0xF else e) for e in unpacked] == tri_ref[c], c
        assert edge_ref[c] == sum(1 << e for e in set(x for x in tri_ref[c] if x >= 0)), c
