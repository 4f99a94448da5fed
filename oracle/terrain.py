"""ORACLE (test infrastructure only).  Terrain model of gym_trackSRB/Characters/SRB5_playground.xml and the
vertical `mj_ray` query the terrain env makes:

  env/testSRB_mujoco_original_viewer_terrain.py:741-752   terrain_height_ray (mj_ray from z=100 straight down,
                                                          geomgroup 0, bodyexclude=cdm)
  gym_trackSRB/Characters/SRB5_playground.xml:23-52,...   plane (size 100 100), hfields hf_hill (512^2 PNG, size
                                                          40 40 6 0.1 at (-42,42)) and hf_rough (128^2 PNG, size
                                                          40 40 1.6 0.1 at (42,-42)), 4 pyramids x 81 boxes

mujoco is third party, un-vendored and unpinned; restated from its published algorithm (engine_ray.c:
mj_ray / mju_rayGeom / mj_rayHfield, user_objects.cc: mjCHField::LoadPNG / Compile):
  * mj_ray keeps the geom with the strictly smallest distance (first geom wins ties), geoms in model order
    (body by body, the world body's geoms first: floor 0, hf_hill 1, hf_rough 2, cdm box 3 [group 1: excluded], pyramid boxes 4..327);
  * plane: hit only inside its rendered size; box: a vertical ray hits the top face iff |dx|<=hx and |dy|<=hy;
  * hfield: PNG rows are stored bottom-up (image row nrow-1-r -> data row r), elevations are normalised to
    [0,1] by (v-min)/(max-min) in float32, vertex (c,r) sits at (2*sx*c/(ncol-1)-sx, 2*sy*r/(nrow-1)-sy),
    every cell is split along the diagonal (c,r)-(c+1,r+1): triangle 0 contains (c+1,r), triangle 1 (c,r+1).
PARITY UNPINNED -- the row order and the triangulation are version specific and can only
be pinned against a real mujoco wheel; the integer outputs (geom id, cell, triangle) are what tests compare."""
import re
import numpy as np


class Terrain:
    def __init__(self, plane_size, plane_geom, hfields, boxes, bodies):
        self.plane_size = np.asarray(plane_size, dtype=float)
        self.plane_geom = int(plane_geom)
        self.hfields = hfields      # list of dict(data float32 [nrow,ncol], size[4], pos[3], geom_id)
        self.boxes = boxes          # list of dict(pos[3] world, size[3], geom_id, body_id) in model order
        self.bodies = bodies        # list of dict(body_id, name, xpos[3], snap)

    # ---- the query ---------------------------------------------------------------------------------
    def hfield_cell(self, hf, x, y):
        """-> (c, r, triangle, height) or None when (x, y) is outside the footprint."""
        sx, sy, sz, _ = hf["size"]
        lx, ly = x - hf["pos"][0], y - hf["pos"][1]
        if abs(lx) > sx or abs(ly) > sy:
            return None
        data = hf["data"]
        nrow, ncol = data.shape
        dx, dy = 2.0 * sx / (ncol - 1), 2.0 * sy / (nrow - 1)
        u, v = (lx + sx) / dx, (ly + sy) / dy
        c = min(max(int(np.floor(u)), 0), ncol - 2)
        r = min(max(int(np.floor(v)), 0), nrow - 2)
        fu, fv = u - c, v - r
        z00, z10 = float(data[r, c]) * sz, float(data[r, c + 1]) * sz
        z01, z11 = float(data[r + 1, c]) * sz, float(data[r + 1, c + 1]) * sz
        if fu >= fv:
            tri, z = 0, z00 + fu * (z10 - z00) + fv * (z11 - z10)
        else:
            tri, z = 1, z00 + fv * (z01 - z00) + fu * (z11 - z01)
        return c, r, tri, hf["pos"][2] + z

    def height_ray(self, x, y):
        """terrain_height_ray: -> (height, geom_id); (None, -1) when nothing is hit."""
        best, gid = None, -1
        if abs(x) <= self.plane_size[0] and abs(y) <= self.plane_size[1]:
            best, gid = 0.0, self.plane_geom
        for hf in self.hfields:
            hit = self.hfield_cell(hf, x, y)
            if hit is not None and (best is None or hit[3] > best):
                best, gid = hit[3], hf["geom_id"]
        for b in self.boxes:
            if abs(x - b["pos"][0]) <= b["size"][0] and abs(y - b["pos"][1]) <= b["size"][1]:
                top = b["pos"][2] + b["size"][2]
                if best is None or top > best:
                    best, gid = top, b["geom_id"]
        return best, gid

    def box_by_geom(self, gid):
        g0 = self.boxes[0]["geom_id"]
        i = gid - g0
        return self.boxes[i] if 0 <= i < len(self.boxes) else None

    def body(self, body_id):
        for b in self.bodies:
            if b["body_id"] == body_id:
                return b
        return None

    # ---- serialisation (tests/golden/terrain_playground.npz) --------------------------------------------
    def save(self, path):
        d = dict(plane_size=self.plane_size, plane_geom=self.plane_geom, n_hf=len(self.hfields))
        for i, hf in enumerate(self.hfields):
            d[f"hf{i}_pix"] = hf["pix"]
            d[f"hf{i}_size"] = np.asarray(hf["size"], dtype=float)
            d[f"hf{i}_pos"] = np.asarray(hf["pos"], dtype=float)
            d[f"hf{i}_geom"] = hf["geom_id"]
        d["box_pos"] = np.array([b["pos"] for b in self.boxes]); d["box_size"] = np.array([b["size"] for b in self.boxes])
        d["box_geom"] = np.array([b["geom_id"] for b in self.boxes]); d["box_body"] = np.array([b["body_id"] for b in self.boxes])
        d["body_id"] = np.array([b["body_id"] for b in self.bodies]); d["body_xpos"] = np.array([b["xpos"] for b in self.bodies])
        d["body_snap"] = np.array([b["snap"] for b in self.bodies]); d["body_name"] = np.array([b["name"] for b in self.bodies])
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path):
        d = np.load(path)
        hfs = []
        for i in range(int(d["n_hf"])):
            pix = d[f"hf{i}_pix"]
            hfs.append(dict(pix=pix, data=normalise_png(pix), size=d[f"hf{i}_size"], pos=d[f"hf{i}_pos"], geom_id=int(d[f"hf{i}_geom"])))
        boxes = [dict(pos=p, size=s, geom_id=int(g), body_id=int(b)) for p, s, g, b in zip(d["box_pos"], d["box_size"], d["box_geom"], d["box_body"])]
        bodies = [dict(body_id=int(i), xpos=x, snap=int(s), name=str(n)) for i, x, s, n in zip(d["body_id"], d["body_xpos"], d["body_snap"], d["body_name"])]
        return Terrain(d["plane_size"], int(d["plane_geom"]), hfs, boxes, bodies)


def normalise_png(pix):
    """mjCHField::LoadPNG (rows reversed) + Compile (normalise to [0,1]) in float32."""
    data = np.ascontiguousarray(pix[::-1, :].astype(np.float32))
    emin, emax = np.float32(data.min()), np.float32(data.max())
    return ((data - emin) / (emax - emin)).astype(np.float32)


def load_playground(xml_path, terrain_dir):
    """Parse SRB5_playground.xml (needs PIL for the PNGs; used only by tests/tools/make_golden.py)."""
    from PIL import Image
    t = open(xml_path).read()
    hf_assets = {}
    for m in re.finditer(r'<hfield name="(\w+)" size="([^"]+)" file="([^"]+)"', t):
        hf_assets[m.group(1)] = (np.array([float(x) for x in m.group(2).split()]), m.group(3).split("/")[-1])
    wb = t[t.index("<worldbody>"):]
    # mujoco numbers geoms body by body (mjCModel::MakeLists): the world body's own geoms first, in document order, then each child
    # body's.  Two passes over the flat element stream: pass 0 takes the geoms outside any <body>, pass 1 the geoms inside bodies.
    geom_id = 0
    hfields, boxes, bodies = [], [], []
    plane_size, plane_geom = None, -1
    for want_inside in (False, True):
        body_id = 0
        cur_body = None
        for m in re.finditer(r'<(/?)(geom|body)\b([^>]*?)(/?)>', wb):
            closing, tag, attrs = m.group(1), m.group(2), m.group(3)
            a = dict(re.findall(r'(\w+)="([^"]*)"', attrs))
            if tag == "body":
                if closing:
                    cur_body = None
                else:
                    body_id += 1
                    cur_body = dict(body_id=body_id, name=a.get("name", ""), xpos=np.array([float(x) for x in a.get("pos", "0 0 0").split()]))
                    if want_inside and cur_body["name"].startswith("pyramid"):
                        bodies.append(cur_body)
                continue
            if closing or (cur_body is not None) != want_inside:
                continue
            gtype = a.get("type", "sphere")
            if gtype == "plane":
                plane_size = np.array([float(x) for x in a["size"].split()][:2]); plane_geom = geom_id
            elif gtype == "hfield":
                size, fname = hf_assets[a["hfield"]]
                pix = np.array(Image.open(f"{terrain_dir}/{fname}").convert("L"), dtype=np.uint8)
                hfields.append(dict(pix=pix, data=normalise_png(pix), size=size, pos=np.array([float(x) for x in a["pos"].split()]), geom_id=geom_id))
            elif gtype == "box" and cur_body is not None and cur_body["name"].startswith("pyramid"):
                lp = np.array([float(x) for x in a["pos"].split()])
                boxes.append(dict(pos=cur_body["xpos"] + lp, size=np.array([float(x) for x in a["size"].split()]), geom_id=geom_id, body_id=cur_body["body_id"]))
            geom_id += 1
    for k, b in enumerate(bodies):
        nxt = [c for c in bodies if c["body_id"] == b["body_id"] + 1]
        b["snap"] = 1 if nxt else 0          # pyramid followed by a pyramid (the reference crashes on the last one)
    return Terrain(plane_size, plane_geom, hfields, boxes, bodies)
