"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the
public header declares; the host layer fails loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import helpers

HEADER = os.path.join(helpers.ROOT, "include", "srbtrack_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(srb_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported(built_lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/srbtrack_b200.h but not exported"


def test_binding_table_matches_header(built_lib):
    from ipcdnnwalk_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared()


def test_fk_header_symbols_exported(built_lib):
    from ipcdnnwalk_b200 import _lib
    txt = open(os.path.join(helpers.ROOT, "include", "bonefk_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b(fk_[a-z_0-9]+)\s*\(", txt)))
    assert names == sorted(_lib.FK_EXPORTED_SYMBOLS) and len(names) == 8
    for n in names:
        assert hasattr(built_lib, n)


def test_cdm_header_symbols_exported(built_lib):
    from ipcdnnwalk_b200 import _lib
    txt = open(os.path.join(helpers.ROOT, "include", "cdmenv_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b(cdm_[a-z_0-9]+)\s*\(", txt)))
    assert names == sorted(_lib.CDM_EXPORTED_SYMBOLS) and len(names) == 17
    for n in names:
        assert hasattr(built_lib, n)


def test_terrain_header_symbols_exported(built_lib):
    from ipcdnnwalk_b200 import _lib
    txt = open(os.path.join(helpers.ROOT, "include", "cdmenv_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b(tlterrain_[a-z_0-9]+)\s*\(", txt)))
    assert names == sorted(_lib.TLT_EXPORTED_SYMBOLS) and len(names) == 3
    for n in names:
        assert hasattr(built_lib, n)


def test_mmsto_header_symbols_exported(built_lib):
    from ipcdnnwalk_b200 import _lib
    txt = open(os.path.join(helpers.ROOT, "include", "mmsto_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b((?:mmsto|fullbody)_[a-z_0-9]+)\s*\(", txt)))
    assert names == sorted(_lib.MMSTO_EXPORTED_SYMBOLS) and len(names) == 14
    for n in names:
        assert hasattr(built_lib, n)
    import torch
    if not torch.cuda.is_available():
        import helpers_mmsto as hm
        from ipcdnnwalk_b200 import ShortTrajOpt
        with pytest.raises(RuntimeError):
            ShortTrajOpt(hm.skeleton(), hm.NF, hm.MARKERS)


def test_cdm_errors_without_gpu(built_lib):
    import ctypes as C
    from ipcdnnwalk_b200 import _lib
    d = [C.c_int32() for _ in range(4)]
    assert built_lib.cdm_state_dims(*[C.byref(x) for x in d]) == 0
    assert [x.value for x in d] == [61, 4, 8, 8]
    h = C.c_void_p()
    for cfg in (_lib.CdmConfig(0, 64, 4, 1, 0, 0), _lib.CdmConfig(4, 16, 4, 1, 0, 0), _lib.CdmConfig(4, 64, 3, 1, 0, 0)):
        assert built_lib.cdm_create(C.byref(cfg), C.byref(h)) != 0
    import torch
    if not torch.cuda.is_available():
        cfg = _lib.CdmConfig(16, 64, 4, 1, 0, 0)
        assert built_lib.cdm_create(C.byref(cfg), C.byref(h)) != 0 and b"no CPU fallback" in built_lib.srb_last_error()
        import helpers_cdm as hc
        from ipcdnnwalk_b200 import CDMVecEnv
        tab, skel = hc.load_tables()
        with pytest.raises(RuntimeError):
            CDMVecEnv(tab, skel, 4)


def test_cdm_skeleton_struct_and_table_packing():
    import helpers_cdm as hc
    from ipcdnnwalk_b200 import cdm_tables, cdm_vec_env
    tab, skel = hc.load_tables()
    s = cdm_vec_env.skeleton_struct(skel)
    assert list(s.leg_nchan) == [3, 1, 3, 3, 1, 3]
    assert list(s.leg_axes) == [1 | (2 << 2) | (0 << 4), 0, 1 | (2 << 2) | (0 << 4)] * 2            # YZX, X, YZX
    rows = cdm_tables.pack_rows(tab)
    assert rows.shape == (tab["cdm"].shape[0], 48) and rows[0, 47] in (0, 1, 2, 3)
    assert (rows[:, 43:47] == np.round(rows[:, 43:47])).all()


def test_skeleton_parser_matches_builtin_tables():
    from ipcdnnwalk_b200 import vrml_skeleton
    s = vrml_skeleton.builtin()
    assert len(s["names"]) == 20 and s["n_pose"] == 60 and s["n_dq"] == 59 and max(s["depth"]) == 7
    assert s["channels"][2] == "X" and s["channels"][11] == "XYZ" and s["channels"][12] == "ZXY"
    wrl = "/root/reference/gym_trackSRB/taesooLib/hanyang_lowdof_T_sh.wrl"
    if os.path.exists(wrl):
        assert vrml_skeleton.parse_wrl(wrl) == s


def test_state_dims_and_errors_without_gpu(built_lib):
    import ctypes as C
    d = [C.c_int32() for _ in range(4)]
    assert built_lib.srb_state_dims(*[C.byref(x) for x in d]) == 0
    assert [x.value for x in d] == [46, 7, 32, 8]
    import torch
    if not torch.cuda.is_available():
        from ipcdnnwalk_b200 import _lib
        cfg = _lib.SrbConfig(16, 0, 64, 8, 1, 0, 0)
        h = C.c_void_p()
        rc = built_lib.srb_create(C.byref(cfg), C.byref(h))
        assert rc != 0 and b"no CPU fallback" in built_lib.srb_last_error()
        from ipcdnnwalk_b200 import SRBVecEnv
        with pytest.raises(RuntimeError):
            SRBVecEnv(helpers.fresh_clips(), 4)


def test_bad_config_rejected(built_lib):
    import ctypes as C
    from ipcdnnwalk_b200 import _lib
    h = C.c_void_p()
    for cfg in (_lib.SrbConfig(0, 0, 64, 8, 1, 0, 0), _lib.SrbConfig(4, 0, 16, 8, 1, 0, 0), _lib.SrbConfig(4, 0, 64, 3, 1, 0, 0),
                _lib.SrbConfig(4, 7, 64, 8, 1, 0, 0)):
        assert built_lib.srb_create(C.byref(cfg), C.byref(h)) != 0


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the product package (nor the product arm of bench.py) may import it, and importing the
    whole package must not pull it in as a side effect."""
    import ast
    import subprocess
    import sys
    root = helpers.ROOT
    pkg = os.path.join(root, "ipcdnnwalk_b200")
    for fn in sorted(os.listdir(pkg)):
        if fn.endswith(".py"):
            tree = ast.parse(open(os.path.join(pkg, fn)).read())
            for node in ast.walk(tree):
                names = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""] if isinstance(node, ast.ImportFrom) else []
                assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f"{fn} imports the oracle"
    code = ("import sys; sys.path.insert(0, %r); import ipcdnnwalk_b200 as p; "
            "from ipcdnnwalk_b200 import SRBVecEnv, SRBEnv, SB3VecEnv, CDMVecEnv, LuaEnv, Skeleton, ShortTrajOpt, TerrainModel; "
            "import ipcdnnwalk_b200.clip_builder, ipcdnnwalk_b200.tl_binary, ipcdnnwalk_b200.mjcf_humanoid; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported by the product package'" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_product_entry_points_fail_loudly_without_a_gpu():
    """No CPU fallback: constructing an environment / optimiser on a box without CUDA raises instead of computing something else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    from ipcdnnwalk_b200 import SRBVecEnv
    from ipcdnnwalk_b200.short_traj_opt import extractSingleFrameFullbody
    with pytest.raises(Exception):
        SRBVecEnv(helpers.fresh_clips(), 4, variant="test", precision=64)
    with pytest.raises(RuntimeError):
        extractSingleFrameFullbody(np.zeros((1, 29)), np.zeros((1, 214)), 60.0, np.eye(3))
