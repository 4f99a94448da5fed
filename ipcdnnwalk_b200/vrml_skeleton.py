"""Reader for taesooLib's VRML-like skeleton files (`*.wrl`), e.g.
gym_trackSRB/taesooLib/hanyang_lowdof_T_sh.wrl:

  DEF <name> Joint { jointType "free"|"rotate"  jointAxis "<channels>"  translation x y z
                     children [ Segment { centerOfMass..  mass..  momentsOfInertia [9] ... }  DEF ... Joint {...} ] }

Bones are numbered depth-first as written (= taesooLib treeIndex-1).  The result is a plain dict of lists
(JSON-serialisable) consumed by `ipcdnnwalk_b200.bone_fk` and by the oracle."""
import json
import re


def _tokens(text):
    text = re.sub(r"#[^\n]*", "", text)
    return re.findall(r'"[^"]*"|[\[\]{}]|[^\s\[\]{}]+', text)


def parse_wrl(path):
    tok = _tokens(open(path).read())
    bones = []
    pos = 0

    def skip_block(i):
        """i points at '{' or '['; returns index after the matching close."""
        open_, close = tok[i], "}" if tok[i] == "{" else "]"
        depth = 0
        while True:
            if tok[i] == open_:
                depth += 1
            elif tok[i] == close:
                depth -= 1
                if depth == 0:
                    return i + 1
            i += 1

    def parse_segment(i, bone):
        assert tok[i] == "{"
        i += 1
        while tok[i] != "}":
            t = tok[i]
            if t == "centerOfMass":
                bone["com"] = [float(x) for x in tok[i + 1:i + 4]]; i += 4
            elif t == "mass":
                bone["mass"] = float(tok[i + 1]); i += 2
            elif t == "momentsOfInertia":
                assert tok[i + 1] == "["
                bone["inertia"] = [float(x) for x in tok[i + 2:i + 11]]; i += 12
            elif tok[i] in "{[":
                i = skip_block(i)
            else:
                i += 1
        return i + 1

    def parse_joint(i, parent):
        # tok[i] == 'DEF', name, 'Joint', '{'
        name = tok[i + 1]
        assert tok[i + 2] == "Joint" and tok[i + 3] == "{"
        bone = dict(name=name, parent=parent, jointType="rotate", channels="", offset=[0.0, 0.0, 0.0],
                    mass=0.0, com=[0.0, 0.0, 0.0], inertia=[0.0] * 9)
        idx = len(bones)
        bones.append(bone)
        i += 4
        while tok[i] != "}":
            t = tok[i]
            if t == "jointType":
                bone["jointType"] = tok[i + 1].strip('"'); i += 2
            elif t == "jointAxis":
                bone["channels"] = tok[i + 1].strip('"'); i += 2
            elif t == "translation":
                bone["offset"] = [float(x) for x in tok[i + 1:i + 4]]; i += 4
            elif t == "children":
                assert tok[i + 1] == "["
                i += 2
                while tok[i] != "]":
                    if tok[i] == "Segment":
                        i = parse_segment(i + 1, bone)
                    elif tok[i] == "DEF" and tok[i + 2] == "Joint":
                        i = parse_joint(i, idx)
                    elif tok[i] in "{[":
                        i = skip_block(i)
                    else:
                        i += 1
                i += 1
            elif tok[i] in "{[":
                i = skip_block(i)
            else:
                i += 1
        return i + 1

    while pos < len(tok):
        if tok[pos] == "DEF" and pos + 2 < len(tok) and tok[pos + 2] == "Joint":
            pos = parse_joint(pos, -1)
        else:
            pos += 1
    return skeleton_from_bones(bones)


def skeleton_from_bones(bones):
    assert bones and bones[0]["jointType"] == "free" and bones[0]["parent"] == -1
    for b in bones[1:]:
        assert b["jointType"] == "rotate" and set(b["channels"]) <= set("XYZ") and 1 <= len(b["channels"]) <= 3
        assert b["parent"] < bones.index(b), "bones must be in depth-first order"
    depth = []
    for b in bones:
        depth.append(0 if b["parent"] < 0 else depth[b["parent"]] + 1)
    nang = sum(len(b["channels"]) for b in bones[1:])
    return dict(names=[b["name"] for b in bones], parent=[b["parent"] for b in bones], depth=depth,
                channels=["" if i == 0 else b["channels"] for i, b in enumerate(bones)],
                offset=[b["offset"] for b in bones], mass=[b["mass"] for b in bones], com=[b["com"] for b in bones],
                inertia=[b["inertia"] for b in bones], n_pose=7 + nang, n_dq=6 + nang)


def load_json(path):
    return json.load(open(path))


def builtin(name="hanyang_lowdof_T"):
    import os
    return load_json(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", f"{name}.json"))
