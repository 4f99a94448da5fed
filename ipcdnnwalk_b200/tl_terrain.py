"""taesooLib height-field terrain (`OBJloader.Terrain`) on the device: host-side mirror of the python-bound class
(work/taesooLib/MainLib/python/MainlibPython.hpp:2689, `Terrain.height(vector2)`) and of gym_cdm2's
`OBJloader.Terrain:getTerrainPos` (gym_cdm2/deepmimiccdm-v1.lua:415-424), batched over query points."""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import check


def load_raw(path, size_x, size_y):
    """2-byte little-endian height image (Terrain.cpp:479-485)."""
    return np.fromfile(path, dtype="<u2", count=size_x * size_y).reshape(size_y, size_x)


class Terrain:
    """heights [rows, cols] float64 (mHeight), size_x / size_z (mSize), tile_along_z; queries run on the GPU."""

    def __init__(self, heights, size_x, size_z, tile_along_z=True, device=0):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("Terrain needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        h = np.ascontiguousarray(heights, dtype=np.float64)
        self._h = C.c_void_p()
        check(self.lib.tlterrain_create(int(device), h.shape[0], h.shape[1], float(size_x), float(size_z), int(bool(tile_along_z)),
                                        h.ctypes.data_as(C.c_void_p), C.byref(self._h)))

    def height(self, xz, with_hit=False):
        """xz float64 [n,2] (terrain-local units) -> height [n] (and hit [n,3] = cell row, cell column, triangle)."""
        t = self.torch
        xz = t.as_tensor(xz, dtype=t.float64, device=self.device).contiguous()
        n = xz.shape[0]
        out = t.empty(n, dtype=t.float64, device=self.device)
        hit = t.empty(n, 3, dtype=t.int32, device=self.device) if with_hit else None
        check(self.lib.tlterrain_height(self._h, n, C.c_void_p(xz.data_ptr()), C.c_void_p(out.data_ptr()),
                                        C.c_void_p(hit.data_ptr()) if with_hit else None,
                                        C.c_void_p(t.cuda.current_stream(self.device).cuda_stream)))
        return (out, hit) if with_hit else out

    def get_terrain_pos(self, pos, terrain_pos):
        """getTerrainPos: pos [n,3] metres -> [n,3] with y = terrain height (cm table, `terrain_pos` = g_terrainPos)."""
        t = self.torch
        pos = t.as_tensor(pos, dtype=t.float64, device=self.device)
        tp = t.as_tensor(terrain_pos, dtype=t.float64, device=self.device)
        loc = (pos - tp) * 100.0
        h = self.height(loc[:, [0, 2]].contiguous())
        out = pos.clone()
        out[:, 1] = h / 100.0 + tp[1]
        return out

    def close(self):
        if self._h:
            self.lib.tlterrain_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
