"""Terrain description handed to `srb_upload_terrain` (terrain variant of the SRBTrack env).

Reference: the group-0 geoms of gym_trackSRB/Characters/SRB5_playground.xml that `terrain_height_ray`
(env/testSRB_mujoco_original_viewer_terrain.py:741-752) ray-casts against -- one plane, PNG height fields, and the
pyramid box stacks.  `TerrainModel.from_npz` reads the compact container written by tests/tools/make_golden.py
(8-bit pixel arrays as stored in the PNG + geometry); pixels are converted the way mujoco stores hfield_data:
image rows bottom-up, elevations normalised to [0,1] in float32."""
import ctypes as C

import numpy as np

from ._lib import SrbHField, SrbBox, SrbTerrainBody


def hfield_data_from_pixels(pix):
    data = np.ascontiguousarray(pix[::-1, :].astype(np.float32))
    lo, hi = np.float32(data.min()), np.float32(data.max())
    return np.ascontiguousarray(((data - lo) / (hi - lo)).astype(np.float32))


class TerrainModel:
    def __init__(self, plane_size, plane_geom, hfields, box_pos, box_size, box_geom, box_body, body_id, body_xpos, body_snap):
        self.plane_size = np.ascontiguousarray(plane_size, dtype=np.float64)
        self.plane_geom = int(plane_geom)
        self.hfields = hfields                      # list of dict(data float32 [nrow, ncol], size[4], pos[3], geom_id)
        self.box_pos = np.asarray(box_pos, dtype=np.float64).reshape(-1, 3)
        self.box_size = np.asarray(box_size, dtype=np.float64).reshape(-1, 3)
        self.box_geom = np.asarray(box_geom, dtype=np.int32)
        self.box_body = np.asarray(box_body, dtype=np.int32)
        self.body_id = np.asarray(body_id, dtype=np.int32)
        self.body_xpos = np.asarray(body_xpos, dtype=np.float64).reshape(-1, 3)
        self.body_snap = np.asarray(body_snap, dtype=np.int32)

    @staticmethod
    def from_npz(path):
        d = np.load(path)
        hfs = []
        for i in range(int(d["n_hf"])):
            hfs.append(dict(data=hfield_data_from_pixels(d[f"hf{i}_pix"]), size=np.asarray(d[f"hf{i}_size"], dtype=np.float64),
                            pos=np.asarray(d[f"hf{i}_pos"], dtype=np.float64), geom_id=int(d[f"hf{i}_geom"])))
        return TerrainModel(d["plane_size"], d["plane_geom"], hfs, d["box_pos"], d["box_size"], d["box_geom"], d["box_body"],
                            d["body_id"], d["body_xpos"], d["body_snap"])

    def c_tables(self):
        """ctypes arrays for srb_upload_terrain (keeps the numpy buffers alive through the returned tuple)."""
        hf = (SrbHField * max(len(self.hfields), 1))()
        for i, h in enumerate(self.hfields):
            hf[i].data_host = h["data"].ctypes.data
            hf[i].nrow, hf[i].ncol = h["data"].shape
            hf[i].size[:] = list(h["size"]); hf[i].pos[:] = list(h["pos"]); hf[i].geom_id = h["geom_id"]
        nb = self.box_geom.shape[0]
        bx = (SrbBox * max(nb, 1))()
        for k in range(nb):
            bx[k].pos[:] = list(self.box_pos[k]); bx[k].size[:] = list(self.box_size[k])
            bx[k].geom_id = int(self.box_geom[k]); bx[k].body_id = int(self.box_body[k])
        nbd = self.body_id.shape[0]
        bd = (SrbTerrainBody * max(nbd, 1))()
        for b in range(nbd):
            bd[b].xpos[:] = list(self.body_xpos[b]); bd[b].body_id = int(self.body_id[b]); bd[b].snap = int(self.body_snap[b])
        return hf, len(self.hfields), bx, nb, bd, nbd
