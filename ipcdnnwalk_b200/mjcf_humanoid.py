"""Reader for the MuJoCo humanoid the reference converts mocap with
(gym_trackSRB/Characters/humanoid_deepmimic_withpalm_Tpose.xml): a free root joint, axis-aligned hinge joints at the
body origins, one massive geom per body.  Produces the same skeleton dict as vrml_skeleton (bodies = bones in
document order, hinge joints = Euler channels in document order) so that the batched FK kernel evaluates
mj_kinematics + the mass-weighted CoM of the whole tree (data.subtree_com[0]) for every mocap frame.

Extra keys: `sign` (per channel, -1 for a negative axis such as the knee's "0 -1 0"), `geoms` (name -> (body, local pos))."""
import json
import os
import xml.etree.ElementTree as ET


def _vec(s, n=3):
    v = [float(x) for x in s.split()]
    assert len(v) == n
    return v


def parse_humanoid(xml_path):
    root = ET.parse(xml_path).getroot()
    bones = []
    geoms = {}

    def add(el, parent):
        b = dict(name=el.get("name"), parent=parent, offset=_vec(el.get("pos", "0 0 0")), channels="", sign=[], mass=0.0, com=[0.0, 0.0, 0.0], free=False)
        idx = len(bones)
        bones.append(b)
        msum, csum = 0.0, [0.0, 0.0, 0.0]
        for ch in el:
            if ch.tag == "joint":
                if ch.get("type") == "free":
                    b["free"] = True
                    continue
                assert ch.get("type", "hinge") == "hinge" and _vec(ch.get("pos", "0 0 0")) == [0, 0, 0], "hinge joints at the body origin only"
                ax = _vec(ch.get("axis"))
                nz = [i for i, a in enumerate(ax) if a != 0]
                assert len(nz) == 1 and abs(ax[nz[0]]) == 1, "axis-aligned hinge joints only"
                b["channels"] += "XYZ"[nz[0]]
                b["sign"].append(int(ax[nz[0]]))
            elif ch.tag == "geom":
                if ch.get("fromto") is not None:
                    ft = _vec(ch.get("fromto"), 6)
                    gp = [0.5 * (ft[i] + ft[i + 3]) for i in range(3)]
                else:
                    gp = _vec(ch.get("pos", "0 0 0"))
                gm = float(ch.get("mass"))
                geoms[ch.get("name")] = (idx, gp)
                msum += gm
                csum = [c + gm * g for c, g in zip(csum, gp)]
        b["mass"] = msum
        b["com"] = [c / msum for c in csum] if msum > 0 else [0.0, 0.0, 0.0]
        for ch in el.findall("body"):
            add(ch, idx)

    for b in root.find("worldbody").findall("body"):
        add(b, -1)
    assert bones[0]["free"] and len(bones[0]["channels"]) == 0
    depth = []
    for b in bones:
        depth.append(0 if b["parent"] < 0 else depth[b["parent"]] + 1)
    nang = sum(len(b["channels"]) for b in bones)
    return dict(names=[b["name"] for b in bones], parent=[b["parent"] for b in bones], depth=depth,
                channels=[b["channels"] for b in bones], sign=[b["sign"] for b in bones], offset=[b["offset"] for b in bones],
                mass=[b["mass"] for b in bones], com=[b["com"] for b in bones], inertia=[[0.0] * 9 for _ in bones],
                n_pose=7 + nang, n_dq=6 + nang, geoms={k: [v[0], v[1]] for k, v in geoms.items()})


def builtin(name="humanoid_deepmimic_withpalm_Tpose"):
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", f"{name}.json")))
