"""ctypes binding of the C-ABI in include/srbtrack_b200.h (libsrbtrack_b200.so, built in-tree).

There is no CPU fallback: if the shared library has not been built the import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPCDNNWALK_B200_LIB", os.path.join(_HERE, "libsrbtrack_b200.so"))   # override: A/B runs of two builds

OBS_DIM, ACT_DIM, RINFO_DIM, PLAN_DIM, DBG_DIM = 150, 22, 8, 9, 256
VARIANT_TEST, VARIANT_TRAINING, VARIANT_TERRAIN = 0, 1, 2


class SrbConfig(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("variant", C.c_int32), ("precision", C.c_int32), ("lanes_per_env", C.c_int32),
                ("auto_reset", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64)]


class SrbHField(C.Structure):
    _fields_ = [("data_host", C.c_void_p), ("nrow", C.c_int32), ("ncol", C.c_int32), ("size", C.c_double * 4), ("pos", C.c_double * 3), ("geom_id", C.c_int32)]


class SrbBox(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("size", C.c_double * 3), ("geom_id", C.c_int32), ("body_id", C.c_int32)]


class SrbTerrainBody(C.Structure):
    _fields_ = [("xpos", C.c_double * 3), ("body_id", C.c_int32), ("snap", C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    "srb_create": (C.c_int, [C.POINTER(SrbConfig), C.POINTER(_P)]),
    "srb_destroy": (C.c_int, [_P]),
    "srb_last_error": (C.c_char_p, []),
    "srb_upload_clip": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "srb_reset": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "srb_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "srb_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "srb_pinned_buffers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "srb_set_host_mode": (C.c_int, [_P, C.c_int32]),
    "srb_host_timeline": (C.c_int, [_P, C.c_int32, _P]),
    "srb_set_aux": (C.c_int, [_P, _P, _P, _P, _P]),
    "srb_set_reward_weights": (C.c_int, [_P, _P]),
    "srb_set_impulse": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32]),
    "srb_state_dims": (C.c_int, [C.POINTER(C.c_int32)] * 4),
    "srb_get_state": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "srb_set_state": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "srb_obs_stats": (C.c_int, [_P, _P, _P, _P]),
    "srb_pack_obs_extra": (C.c_int, [_P, _P, _P]),
    "srb_episode_stats": (C.c_int, [_P, _P, C.c_int32]),
    "srb_episode_stats_device": (C.c_int, [_P, _P, _P]),
    "srb_xchg_create": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "srb_xchg_connect": (C.c_int, [_P, _P]),
    "srb_rollout_stats_allreduce": (C.c_int, [_P, _P, _P, _P]),
    "srb_xchg_status": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "srb_tracking_action": (C.c_int, [_P, C.c_float, C.c_uint64, C.c_uint32, _P, _P]),
    "srb_upload_terrain": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P]),
    "srb_set_planar_offsets": (C.c_int, [_P, _P]),
    "srb_terrain_height": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P]),
    "srb_terrain_device": (_P, [_P]),
    "srb_launch_count": (C.c_int64, [_P]),
}
_FK_SIGNATURES = {          # include/bonefk_b200.h
    "fk_create": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "fk_destroy": (C.c_int, [_P]),
    "fk_set_generic": (C.c_int, [_P, C.c_int32]),
    "fk_is_specialised": (C.c_int, [_P]),
    "fk_set_tma": (C.c_int, [_P, C.c_int32]),
    "fk_dims": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "fk_forward_f32": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "fk_forward_f64": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P]),
}
class CdmConfig(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("precision", C.c_int32), ("lanes_per_env", C.c_int32), ("auto_reset", C.c_int32),
                ("device", C.c_int32), ("seed", C.c_uint64)]


class CdmSkeleton(C.Structure):
    _fields_ = [("root_off", C.c_double * 3), ("leg_off", (C.c_double * 3) * 6), ("leg_nchan", C.c_int32 * 6), ("leg_axes", C.c_int32 * 6),
                ("toe", C.c_double * 3), ("heel", C.c_double * 3)]


CDM_OBS_DIM, CDM_ACT_DIM, CDM_TAB_W, CDM_PLAN_DIM, CDM_DBG_DIM = 27, 10, 48, 16, 96
_CDM_SIGNATURES = {         # include/cdmenv_b200.h
    "cdm_create": (C.c_int, [C.POINTER(CdmConfig), C.POINTER(_P)]),
    "cdm_destroy": (C.c_int, [_P]),
    "cdm_upload_table": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.POINTER(CdmSkeleton)]),
    "cdm_set_param": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "cdm_get_param": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double)]),
    "cdm_reset": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "cdm_step": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "cdm_step_host": (C.c_int, [_P, _P, _P, _P, _P]),
    "cdm_pinned_buffers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "cdm_set_aux": (C.c_int, [_P, _P, _P, _P]),
    "cdm_state_dims": (C.c_int, [C.POINTER(C.c_int32)] * 4),
    "cdm_get_state": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "cdm_set_state": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "cdm_episode_stats": (C.c_int, [_P, _P, C.c_int32]),
    "cdm_launch_count": (C.c_int64, [_P]),
    "cdm_set_terrain": (C.c_int, [_P, _P, _P]),
    "cdm_obs_dim": (C.c_int, [_P]),
}
_TLT_SIGNATURES = {
    "tlterrain_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, _P, C.POINTER(_P)]),
    "tlterrain_destroy": (C.c_int, [_P]),
    "tlterrain_height": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
}
_MMSTO_SIGNATURES = {       # include/mmsto_b200.h
    "mmsto_create": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, C.POINTER(_P)]),
    "mmsto_destroy": (C.c_int, [_P]),
    "mmsto_ndof": (C.c_int, [_P]),
    "mmsto_set_param": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "mmsto_get_param": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double)]),
    "mmsto_set_dof_weights": (C.c_int, [_P, _P]),
    "mmsto_solve": (C.c_int, [_P, C.c_int32] + [_P] * 11 + [_P]),
    "mmsto_evaluate": (C.c_int, [_P, C.c_int32] + [_P] * 11 + [_P]),
    "mmsto_remove_com_sliding": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "mmsto_window_com": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P]),
    "mmsto_launch_count": (C.c_int64, [_P]),
    "mmsto_remove_foot_sliding": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_double, _P, _P, _P]),
    "fullbody_extract": (C.c_int, [C.c_int32, C.c_int32, _P, _P, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mmsto_last_error": (C.c_char_p, []),
}
MMSTO_EXPORTED_SYMBOLS = tuple(_MMSTO_SIGNATURES)
EXPORTED_SYMBOLS = tuple(_SIGNATURES)
CDM_EXPORTED_SYMBOLS = tuple(_CDM_SIGNATURES)
TLT_EXPORTED_SYMBOLS = tuple(_TLT_SIGNATURES)
FK_EXPORTED_SYMBOLS = tuple(_FK_SIGNATURES)

_lib = None


def load():
    """Load the shared library (no GPU needed to load it; compute entry points need one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  ipcdnnwalk_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(_SIGNATURES.items()) + list(_FK_SIGNATURES.items()) + list(_CDM_SIGNATURES.items()) + list(_TLT_SIGNATURES.items()) + list(_MMSTO_SIGNATURES.items()):
        if "IPCDNNWALK_B200_LIB" in os.environ and not hasattr(lib, name):
            continue                       # an older build loaded for an A/B timing run
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SrbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise SrbError(f"libsrbtrack_b200 error {rc}: {load().srb_last_error().decode()}")
