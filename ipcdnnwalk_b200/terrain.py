"""Terrain description handed to `srb_upload_terrain` (terrain variant of the SRBTrack env).

Reference: the group-0 geoms of gym_trackSRB/Characters/SRB5_playground.xml that `terrain_height_ray`
(env/testSRB_mujoco_original_viewer_terrain.py:741-752) ray-casts against -- one plane, PNG height fields, and the
pyramid box stacks.  `TerrainModel.from_mjcf` is the product-side loader of that MJCF file and its PNG height fields
(what `mujoco.MjModel.from_xml_path(model_path)` does for the reference, `SRBEnv5_mocap_list_terrain_window.py:126`);
`TerrainModel.from_npz` reads the same content from a compact container (8-bit pixel arrays as stored in the PNG +
geometry; `python -m ipcdnnwalk_b200.terrain <xml> <out.npz>` writes one).  Pixels are converted the way mujoco stores
hfield_data: image rows bottom-up, elevations normalised to [0,1] in float32.  Geoms are numbered the way mujoco compiles
a model: body by body in depth-first order, the world body's own geoms first."""
import ctypes as C
import os
import struct
import xml.etree.ElementTree as ET
import zlib

import numpy as np

from ._lib import SrbHField, SrbBox, SrbTerrainBody


def hfield_data_from_pixels(pix):
    data = np.ascontiguousarray(pix[::-1, :].astype(np.float32))
    lo, hi = np.float32(data.min()), np.float32(data.max())
    return np.ascontiguousarray(((data - lo) / (hi - lo)).astype(np.float32))


def read_png_gray8(path):
    """Minimal PNG reader for the height-field images (non-interlaced, 8 or 16 bit, grey / RGB / with alpha) -> uint8 [rows, cols]
    grey values, what mjCHField::LoadPNG asks its decoder for (8-bit grey)."""
    b = open(path, "rb").read()
    if b[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(b):
        n, tag = struct.unpack(">I4s", b[pos:pos + 8])
        body = b[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat.append(body)
        elif tag == b"IEND":
            break
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    nch = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if nch is None or interlace or depth not in (8, 16):
        raise ValueError(f"{path}: unsupported PNG layout (colour type {ctype}, depth {depth}, interlace {interlace})")
    bpp = nch * depth // 8
    raw = zlib.decompress(b"".join(idat))
    stride = w * bpp
    img = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    for r in range(h):
        f = raw[r * (stride + 1)]
        line = np.frombuffer(raw, dtype=np.uint8, count=stride, offset=r * (stride + 1) + 1).astype(np.int32)
        if f == 2:
            line = (line + prev) & 255
        elif f in (1, 3, 4):                                  # filters that depend on the pixel to the left: byte serial
            out = np.zeros(stride, dtype=np.int32)
            for i in range(stride):
                a = out[i - bpp] if i >= bpp else 0
                c = prev[i - bpp] if i >= bpp else 0
                up = prev[i]
                if f == 1:
                    pr = a
                elif f == 3:
                    pr = (a + up) >> 1
                else:
                    pa, pb, pc = abs(up - c), abs(a - c), abs(a + up - 2 * c)
                    pr = a if (pa <= pb and pa <= pc) else (up if pb <= pc else c)
                out[i] = (line[i] + pr) & 255
            line = out
        elif f != 0:
            raise ValueError(f"{path}: bad PNG filter {f}")
        img[r] = line
        prev = line
    px = img.reshape(h, w, nch, depth // 8)[..., 0].astype(np.float64)      # most significant byte of every channel
    if nch >= 3:                                                             # grey = the decoder's RGB -> luma conversion
        grey = np.floor(0.299 * px[..., 0] + 0.587 * px[..., 1] + 0.114 * px[..., 2] + 0.5)
    else:
        grey = px[..., 0]
    return grey.astype(np.uint8)


def _floats(text, default=None):
    if text is None:
        return None if default is None else np.asarray(default, dtype=np.float64)
    return np.array([float(x) for x in text.split()], dtype=np.float64)


class TerrainModel:
    def __init__(self, plane_size, plane_geom, hfields, box_pos, box_size, box_geom, box_body, body_id, body_xpos, body_snap, body_names=None):
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
        self.body_names = list(body_names) if body_names is not None else [f"pyramid_{i + 1}" for i in range(self.body_id.shape[0])]

    @staticmethod
    def from_mjcf(xml_path):
        """Parse the MJCF file of the terrain env (SRB5_playground.xml): the plane, the PNG height fields and the `pyramid*` box
        stacks, numbered as mujoco numbers bodies and geoms.  Only what the vertical ray of terrain_height_ray can hit is kept
        (geom group 0; the SRB box is group 1)."""
        root = ET.parse(xml_path).getroot()
        base = os.path.dirname(os.path.abspath(xml_path))
        assets = {}
        for hf in root.iter("hfield"):
            if hf.get("file"):
                assets[hf.get("name")] = (_floats(hf.get("size")), os.path.normpath(os.path.join(base, hf.get("file"))))
        world = root.find("worldbody")
        plane_size, plane_geom = None, -1
        hfields, box_pos, box_size, box_geom, box_body = [], [], [], [], []
        bodies = []
        counters = dict(geom=0, body=0)

        def visit(body, xpos, body_id, name):
            nonlocal plane_size, plane_geom
            for g in body.findall("geom"):
                gid = counters["geom"]; counters["geom"] += 1
                if int(g.get("group", "0")) != 0:
                    continue
                gtype = g.get("type", "sphere")
                if gtype == "plane" and body_id == 0:
                    plane_size, plane_geom = _floats(g.get("size"))[:2], gid
                elif gtype == "hfield" and g.get("hfield") in assets:
                    size, png = assets[g.get("hfield")]
                    pix = read_png_gray8(png)
                    hfields.append(dict(pix=pix, data=hfield_data_from_pixels(pix), size=size, pos=xpos + _floats(g.get("pos"), (0, 0, 0)), geom_id=gid))
                elif gtype == "box" and name.startswith("pyramid"):
                    box_pos.append(xpos + _floats(g.get("pos"), (0, 0, 0))); box_size.append(_floats(g.get("size")))
                    box_geom.append(gid); box_body.append(body_id)
            for child in body.findall("body"):
                counters["body"] += 1
                cid, cname = counters["body"], child.get("name", "")
                cpos = xpos + _floats(child.get("pos"), (0, 0, 0))
                if cname.startswith("pyramid"):
                    bodies.append(dict(body_id=cid, name=cname, xpos=cpos))
                visit(child, cpos, cid, cname)
        visit(world, np.zeros(3), 0, "world")
        ids = {b["body_id"] for b in bodies}
        snap = [1 if (b["body_id"] + 1) in ids else 0 for b in bodies]     # a pyramid body followed by a pyramid body (:349, :410)
        return TerrainModel(plane_size, plane_geom, hfields, box_pos, box_size, box_geom, box_body, [b["body_id"] for b in bodies],
                            [b["xpos"] for b in bodies], snap, body_names=[b["name"] for b in bodies])

    def save_npz(self, path):
        d = dict(plane_size=self.plane_size, plane_geom=self.plane_geom, n_hf=len(self.hfields), box_pos=self.box_pos, box_size=self.box_size,
                 box_geom=self.box_geom, box_body=self.box_body, body_id=self.body_id, body_xpos=self.body_xpos, body_snap=self.body_snap,
                 body_name=np.array(self.body_names))
        for i, hf in enumerate(self.hfields):
            d[f"hf{i}_pix"] = hf["pix"]; d[f"hf{i}_size"] = hf["size"]; d[f"hf{i}_pos"] = hf["pos"]; d[f"hf{i}_geom"] = hf["geom_id"]
        np.savez_compressed(path, **d)

    @staticmethod
    def builtin_playground():
        """The terrain of gym_trackSRB/Characters/SRB5_playground.xml as parsed by `from_mjcf` (constant table shipped with the package)."""
        return TerrainModel.from_npz(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "playground_terrain.npz"))

    @staticmethod
    def from_npz(path):
        d = np.load(path)
        hfs = []
        for i in range(int(d["n_hf"])):
            hfs.append(dict(pix=d[f"hf{i}_pix"], data=hfield_data_from_pixels(d[f"hf{i}_pix"]), size=np.asarray(d[f"hf{i}_size"], dtype=np.float64),
                            pos=np.asarray(d[f"hf{i}_pos"], dtype=np.float64), geom_id=int(d[f"hf{i}_geom"])))
        names = [str(n) for n in d["body_name"]] if "body_name" in d else None
        return TerrainModel(d["plane_size"], d["plane_geom"], hfs, d["box_pos"], d["box_size"], d["box_geom"], d["box_body"],
                            d["body_id"], d["body_xpos"], d["body_snap"], body_names=names)

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


if __name__ == "__main__":                                 # python -m ipcdnnwalk_b200.terrain <SRB5_playground.xml> <out.npz>
    import sys
    TerrainModel.from_mjcf(sys.argv[1]).save_npz(sys.argv[2])
