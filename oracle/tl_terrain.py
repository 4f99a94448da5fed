"""ORACLE (test infrastructure only).  taesooLib's height-field terrain as gym_cdm2 uses it:

  work/taesooLib/BaseLib/motion/Terrain.cpp:305-393   OBJloader::Terrain::height  (cell by floor, triangle by dx > dz,
                                                      plane / vertical-ray intersection, tiling along Z)
  work/taesooLib/BaseLib/motion/Terrain.cpp:395-507   loader: 2-byte raw image, BAD_NOTEBOOK decimation to 64 x 64
                                                      (every (size/64)-th sample), height = heightMax * v / 65536
  work/taesooLib/BaseLib/motion/intersectionTest.cpp:96-104,243-260   Plane::setPlane / Ray::intersects
  gym_cdm2/deepmimiccdm-v1.lua:368-423                g_terrain = Terrain(heightmap1_256_256_2bytes.raw, 256, 256, 4800,
                                                      4800, 100 h, 1, 1, tileAlongZ = true); getTerrainPos (metres <-> cm)

NOT restated: with tileAlongZ the loader blends the first and last sizeY/8 image rows with `m::c1stitch` (taesooLib C++,
needs Eigen; Terrain.cpp:398-428).  `load_heights` applies a stated stand-in (linear cross-fade of the same rows) so the
table tiles; the QUERY -- the part on the per-step path -- is restated exactly and works on any height table.
PARITY UNPINNED (no recorded terrain query exists; refTraj_walkterrain-v1.dat was recorded with terrainHeight = 0)."""
import math
import numpy as np


def load_raw(path, size_x, size_y):
    return np.fromfile(path, dtype="<u2", count=size_x * size_y).reshape(size_y, size_x)


def load_heights(image, height_max, tile_along_z=True):
    """Terrain::Terrain(filename, ...) -> mHeight [rows, cols] (same units as height_max)."""
    sy0, sx0 = image.shape
    sx, sy = min(64, sx0), min(64, sy0)                       # BAD_NOTEBOOK decimation
    img = image[::sy0 // sy, ::sx0 // sx][:sy, :sx].astype(np.float64)
    if tile_along_z:                                          # STAND-IN for m::c1stitch (see header)
        nw = sy // 8
        a, b = img[:nw].copy(), img[sy - nw:].copy()
        for y in range(nw):
            w = (y + 1.0) / (nw + 1.0)
            img[sy - nw + y] = np.rint(b[y] * (1 - w) + a[0] * w)
        img[sy - 1] = img[0]
    return height_max * img / 65536.0


class OracleTerrain:
    def __init__(self, heights, size_x, size_z, tile_along_z=True):
        self.h = np.asarray(heights, dtype=np.float64)
        self.size_x, self.size_z, self.tile = float(size_x), float(size_z), bool(tile_along_z)

    def height(self, x0, x1):
        """-> (height, cell row i, cell column j, triangle 0 upper / 1 lower); (0, -1, -1, -1) outside."""
        nsx, nsz = self.h.shape[1] - 1, self.h.shape[0] - 1
        nx = (x0 / self.size_x) * float(nsx)
        nz = (x1 / self.size_z) * float(nsz)
        j, i = int(math.floor(nx)), int(math.floor(nz))
        dx, dz = nx - math.floor(nx), nz - math.floor(nz)
        if j < 0 or j >= nsx:
            return 0.0, -1, -1, -1
        if self.tile:
            while i < 0:
                i += nsz
            while i >= nsz:
                i -= nsz
            nz = float(i) + dz
            x1 = nz / float(nsz) * self.size_z
        elif i < 0 or i >= nsz:
            return 0.0, -1, -1, -1
        h = self.h
        if dx > dz:
            v0 = np.array([float(j) * self.size_x / float(nsx), h[i, j], float(i) * self.size_z / float(nsz)])
            v1 = np.array([float(j + 1) * self.size_x / float(nsx), h[i + 1, j + 1], float(i + 1) * self.size_z / float(nsz)])
            v2 = np.array([v1[0], h[i, j + 1], v0[2]])
            tri = 0
        else:
            v0 = np.array([float(j) * self.size_x / float(nsx), h[i, j], float(i) * self.size_z / float(nsz)])
            v2 = np.array([float(j + 1) * self.size_x / float(nsx), h[i + 1, j + 1], float(i + 1) * self.size_z / float(nsz)])
            v1 = np.array([v0[0], h[i + 1, j], v2[2]])
            tri = 1
        n = np.cross(v1 - v0, v2 - v0)
        n = n / math.sqrt(n.dot(n))
        d = -n.dot(v0)
        org = np.array([x0, 1000.0, x1])
        denom = n.dot(np.array([0.0, -1.0, 0.0]))
        t = -((n.dot(org) + d) / denom)
        return org[1] + t * -1.0, i, j, tri


def get_terrain_pos(terrain, pos, terrain_pos):
    """OBJloader.Terrain:getTerrainPos (deepmimiccdm-v1.lua:415-424): metres in, metres out."""
    xz = (np.asarray(pos, float) - np.asarray(terrain_pos, float)) * 100.0
    h = terrain.height(xz[0], xz[2])[0]
    return np.array([pos[0], h / 100.0 + terrain_pos[1], pos[2]])
