"""ipcdnnwalk_b200 -- B200-native batched simulator for the SRB character-tracking hot path of
taesoobear/IPCDNNwalk.  Hand-written sm_100a CUDA behind a C-ABI; no CPU fallback."""
from ._lib import OBS_DIM, ACT_DIM, RINFO_DIM, PLAN_DIM, DBG_DIM, LIB_PATH, SrbError  # noqa: F401
from .srb_vec_env import SRBVecEnv, SB3VecEnv, REWARD_INFO_KEYS  # noqa: F401
from . import vrml_skeleton  # noqa: F401
from .bone_fk import Skeleton, BoneForwardKinematics, LoaderToTree  # noqa: F401
from .terrain import TerrainModel  # noqa: F401
from .srb_env import SRBEnv  # noqa: F401
from .cdm_vec_env import CDMVecEnv, LuaEnv  # noqa: F401
from . import cdm_tables  # noqa: F401
from .short_traj_opt import ShortTrajOpt  # noqa: F401
