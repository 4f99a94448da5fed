"""OBJloader::Terrain::height: CPU properties of the oracle restatement and GPU parity of the
batched query (cell row / column / triangle bit exact, height to fp64 round-off)."""
import os

import numpy as np
import pytest

import helpers
from oracle import tl_terrain as tt


@pytest.fixture(scope="module")
def field():
    d = np.load(os.path.join(helpers.GOLDEN, "cdm_terrain_heights.npz"))
    return d["heights"], float(d["size_x"]), float(d["size_z"]), d["terrain_pos"]


def test_oracle_height_properties(field):
    H, sx, sz, tp = field
    T = tt.OracleTerrain(H, sx, sz, True)
    nsx, nsz = H.shape[1] - 1, H.shape[0] - 1
    for i, j in ((0, 0), (3, 7), (62, 62), (10, 40)):              # the surface interpolates the grid vertices
        x, z = j * sx / nsx, i * sz / nsz
        assert abs(T.height(x + 1e-9, z + 1e-9)[0] - H[i, j]) < 1e-6
    rng = np.random.default_rng(0)
    for _ in range(200):
        x, z = rng.uniform(0, sx * 0.999), rng.uniform(-2 * sz, 3 * sz)
        h, i, j, tri = T.height(x, z)
        assert 0 <= i < nsz and 0 <= j < nsx and tri in (0, 1)
        cell = [H[i, j], H[i + 1, j], H[i, j + 1], H[i + 1, j + 1]]
        assert min(cell) - 1e-9 <= h <= max(cell) + 1e-9         # planar patch inside the cell's range
        assert abs(T.height(x, z + sz)[0] - h) < 1e-9             # tiles along Z
    # continuity across the cell diagonal (dx == dz): both triangles agree there
    x0, z0 = 5 * sx / nsx, 9 * sz / nsz
    e = 0.3 * sx / nsx
    up = T.height(x0 + e + 1e-7, z0 + e * sz / sx * (nsx / nsz))
    lo = T.height(x0 + e - 1e-7, z0 + e * sz / sx * (nsx / nsz))
    assert up[3] != lo[3] and abs(up[0] - lo[0]) < 1e-4
    assert T.height(-1.0, 10.0) == (0.0, -1, -1, -1) and T.height(sx, 10.0)[1] == -1        # outside in x
    flat = tt.OracleTerrain(H, sx, sz, False)
    assert flat.height(10.0, -1.0)[1] == -1 and flat.height(10.0, sz)[1] == -1               # no tiling: outside in z
    p = tt.get_terrain_pos(T, np.array([1.0, 5.0, 2.0]), tp)
    assert p[0] == 1.0 and p[2] == 2.0 and abs(p[1] - (T.height(2500.0, 2600.0)[0] / 100 + tp[1])) < 1e-12


@pytest.mark.gpu
def test_gpu_height_matches_oracle(field):
    import torch
    from ipcdnnwalk_b200.tl_terrain import Terrain
    H, sx, sz, tp = field
    rng = np.random.default_rng(1)
    n = 20000
    xz = np.stack([rng.uniform(-0.05 * sx, 1.05 * sx, n), rng.uniform(-2 * sz, 3 * sz, n)], 1)
    xz[:50, 0] = np.arange(50) * sx / (H.shape[1] - 1)                  # exactly on grid lines
    xz[50:100, 1] = np.arange(50) * sz / (H.shape[0] - 1)
    for tile in (True, False):
        T = Terrain(H, sx, sz, tile)
        O = tt.OracleTerrain(H, sx, sz, tile)
        h, hit = T.height(xz, with_hit=True)
        torch.cuda.synchronize()
        h, hit = h.cpu().numpy(), hit.cpu().numpy()
        ref = [O.height(a, b) for a, b in xz]
        assert np.array_equal(hit, np.array([r[1:] for r in ref], dtype=np.int32))      # cell index and triangle: bit exact
        assert np.abs(h - np.array([r[0] for r in ref])).max() < 1e-9
        T.close()
    T = Terrain(H, sx, sz, True)
    pos = np.stack([rng.uniform(-20, 20, 100), np.zeros(100), rng.uniform(-60, 60, 100)], 1)
    got = T.get_terrain_pos(pos, tp).cpu().numpy()
    O = tt.OracleTerrain(H, sx, sz, True)
    want = np.stack([tt.get_terrain_pos(O, p, tp) for p in pos])
    assert np.abs(got - want).max() < 1e-11
