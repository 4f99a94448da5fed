/* cdmenv_b200.h -- C-ABI of the B200-native batched AdaptiveSRB environment (`deepmimiccdm-v1`, gym_cdm2),
 * SURVEY 8a rows a-16 ... a-20, BASELINE config 4.  Same library (libsrbtrack_b200.so) and conventions as
 * srbtrack_b200.h: 0 / negative return codes, message from srb_last_error(), caller-owned device buffers,
 * device entry points asynchronous on `stream`, one handle per GPU, not thread-safe per handle.
 *
 * What it replaces (paths relative to the reference root): the Lua globals a `LuaEnv` calls through `lua.F`
 * (gym_cdm2/luaEnvGymnasium.py:57-62,71,80-125):
 *     get_dim()          -> CDM_OBS_DIM / CDM_ACT_DIM            gym_cdm2/deepmimiccdm-v1.lua:434
 *     init_env()         -> cdm_create + cdm_upload_table        :323-432 (RagdollSim ctor + prepareMotionForSim)
 *     reset(0)           -> cdm_reset                            :903, RagdollSim:reset :1578-1722
 *     step(0, action)    -> cdm_step / cdm_step_host             :853-897, RagdollSim:step :1014-1574
 * for num_envs environments at once (the reference runs them in separate processes, gym_cdm2/train_gym.py:165-176).
 */
#ifndef CDMENV_B200_H
#define CDMENV_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cdm_env cdm_env;

#define CDM_OBS_DIM 27    /* 11 + 6*numLegs + 2 + 2   deepmimiccdm-v1.lua:435 (flat ground) */
#define CDM_ACT_DIM 10    /* R foot xz, L foot xz, d(lin vel) 3, d(ang vel) 3; scaled by 0.2 (defaultRLsetting.lua:26-28) */
#define CDM_TAB_W 48      /* reference-table row, see below */
#define CDM_PLAN_DIM 16   /* reset draws: quaternion noise 4, position noise 3, dq noise 6 (angular, linear), extra linear velocity 3 */
#define CDM_DBG_DIM 96    /* qdd 2x6 | a_des 2x6 | lambda 2x20 | pose_dif com_lvdif ref_comvel timing_mod | theta_d 7 | iters 2 | #contacts */

typedef struct cdm_config {
  int32_t num_envs;
  int32_t precision;      /* 64: fp64 state + arithmetic (reference precision); 32: fp32 state, QP still fp64 (ridge 1e-8) */
  int32_t lanes_per_env;  /* 0 = default (1: one thread per environment, structured QP); 2,4,8 lanes of a warp cooperate on one environment */
  int32_t auto_reset;     /* 1: a terminated env is reset inside the step */
  int32_t device;
  uint64_t seed;          /* stream of the device-side reset draws (CT.randomUniform / math.random in RagdollSim:reset) */
} cdm_config;

/* The part of the skeleton the step needs (ankle markers of the reference pose, getJointStateFromRefTree :996-1012):
 * joint translations of the two leg chains of work/taesooLib/Resource/motion/MOB1/hanyang_lowdof_T.wrl and the
 * marker offsets on the ankles (:984-985).  Chain order: L hip, L knee, L ankle, R hip, R knee, R ankle.
 * leg_axes packs the Euler channels 2 bits each, first channel lowest (0 X, 1 Y, 2 Z). */
typedef struct cdm_skeleton {
  double root_off[3];
  double leg_off[6][3];
  int32_t leg_nchan[6];
  int32_t leg_axes[6];
  double toe[3], heel[3];
} cdm_skeleton;

int cdm_create(const cdm_config* cfg, cdm_env** out);
int cdm_destroy(cdm_env* env);

/* Reference tables of one clip (RagdollSim:prepareMotionForSim gym_cdm2/RagdollSim.lua:18-95 -> fc.prepareMotionsForSim
 * module/CDMTraj.lua:1400-1610, lazily built integer tables defaultRLsetting.lua:291-346,395-418).
 * rows_host[nframes][CDM_TAB_W] (fp64, HOST): CDM pose 7 | d/dt CDM body twist v 3, w 3 | mocapFeetOffset R 3, L 3 |
 * leg-chain pose 21 (root xyz, root quat wxyz, L hip/knee/ankle, R hip/knee/ankle angles) | motionDOF_iframe 1 |
 * swingDuration R,L | rttTouchDown R,L | rttTouchOff R,L | contact bits (R=1, L=2).
 * nf_cycle = numFrames of the un-looped clip - 1. */
int cdm_upload_table(cdm_env* env, int32_t nframes, int32_t nf_cycle, const double* rows_host, const cdm_skeleton* skel);

/* model / QPparam overrides of a spec file (gym_cdm2/spec/walk2-v1.lua, defaultRLsetting.lua:13-45).  Names:
 * k_p k_d acc_clamp max_leg_len max_foot_vel stride_thr max_contact_len min_swing_phase max_contact_y healthy_min
 * healthy_max fall_angle timing_mod_coef facing_weight ef_scale limit_length max_turning filter_end_phase logQ
 * w_ddq w_lam inv_fric mass inertia_x inertia_y inertia_z grav noise_q noise_pos noise_dq noise_vx noise_vy noise_vz
 * use_lqr_cache filter_K0 filter_K1 (contact-filter LQR gain; default = the gain the shipped reference build always uses, Q15) */
int cdm_set_param(cdm_env* env, const char* name, double value);
int cdm_get_param(cdm_env* env, const char* name, double* value);

/* RagdollSim:reset for `count` environments; env_idx_host NULL = 0..count-1; plan_host[count][CDM_PLAN_DIM] explicit
 * (already scaled) draws or NULL = draw on the device; obs_dev = full [num_envs][CDM_OBS_DIM] array or NULL. */
int cdm_reset(cdm_env* env, const int32_t* env_idx_host, int32_t count, const double* plan_host, float* obs_dev, void* stream);

/* RagdollSim:step for all environments: act_dev[N][10] -> obs_dev[N][27], rew_dev[N], done_dev[N] (0/1).
 * The reference's NaN convention is kept per environment (NaN observation -> zeros, reward 0, done; :1557-1561). */
int cdm_step(cdm_env* env, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, void* stream);
/* Same through HOST buffers: H2D actions, kernel, D2H results on the handle's stream, synchronised.  That stream is ordered after
 * every reset / step the handle launched on a caller stream (recorded event), so reset(); step_host() needs no caller-side sync. */
int cdm_step_host(cdm_env* env, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host);
int cdm_pinned_buffers(cdm_env* env, float** act, float** obs, float** rew, uint8_t** done);

/* Side channels (NULL = detach): terminal observations under auto-reset, explicit draws for the next auto-reset,
 * per-env QP / reward intermediates (the tensors the reference's disabled checkpoint comparator looks at,
 * module/QPservo_v3.lua:830-843,1059,1226). */
int cdm_set_aux(cdm_env* env, float* term_obs_dev, const double* reset_plan_dev, double* dbg_dev);

/* RagdollSim:saveFullState / restoreFullState (:1730-1776) of one environment. */
int cdm_state_dims(int32_t* n_real, int32_t* n_int, int32_t* hist_slots, int32_t* hist_width);
int cdm_get_state(cdm_env* env, int32_t env_index, double* reals, int32_t* ints, double* hist);
int cdm_set_state(cdm_env* env, int32_t env_index, const double* reals, const int32_t* ints, const double* hist);

/* out[3] = sum of episode returns, sum of episode lengths, episode count accumulated by auto-reset. */
int cdm_episode_stats(cdm_env* env, double out[3], int32_t clear);
int64_t cdm_launch_count(cdm_env* env);

/* ---- OBJloader::Terrain (work/taesooLib/BaseLib/motion/Terrain.cpp:305-470), the height field of the gym_cdm2 terrain spec
 * (gym_cdm2/deepmimiccdm-v1.lua:368-423, spec/walkterrain-v1.lua).  heights_host[rows][cols] = mHeight (already scaled,
 * same unit as the query coordinates); size_x / size_z = mSize.x / mSize.z; tile_along_z as the constructor flag.
 * tlterrain_height: xz_dev[n][2] (terrain-local) -> height_dev[n]; hit_dev[n][3] (may be NULL) = cell row i, cell column j,
 * triangle (0 upper, 1 lower), all -1 outside the field (height 0). */
typedef struct tl_terrain tl_terrain;
#define CDM_OBS_DIM_TERRAIN 31   /* + 4 terrain height samples ahead of the body (gym_cdm2/deepmimiccdmterrain-v1.lua:3-16) */
int tlterrain_create(int32_t device, int32_t rows, int32_t cols, double size_x, double size_z, int32_t tile_along_z,
                     const double* heights_host, tl_terrain** out);
int tlterrain_destroy(tl_terrain* t);
int tlterrain_height(tl_terrain* t, int64_t n, const double* xz_dev, double* height_dev, int32_t* hit_dev, void* stream);
/* The terrain variant of the environment (gym_cdm2/spec/walkterrain-v1.lua -> deepmimiccdmterrain-v1.lua): every
 * RagdollSim:projectToGround goes to this height field (terrain_pos = g_terrainPos in metres, heights / sizes in cm,
 * deepmimiccdm-v1.lua:368-424), the targets follow the terrain (:1205-1238), the observation grows to CDM_OBS_DIM_TERRAIN.
 * t == NULL returns to flat ground.  The terrain must outlive the environment.  Spec values that differ from walk2-v1
 * (healthy_min 0.7) are set with cdm_set_param.  Call before cdm_reset.  cdm_step launches the terrain instantiation of the step kernel
 * while a terrain is attached and the flat-ground one (no terrain code at all) otherwise; results on flat ground are the same either way. */
int cdm_set_terrain(cdm_env* env, tl_terrain* t, const double terrain_pos[3]);
int cdm_obs_dim(cdm_env* env);   /* CDM_OBS_DIM or CDM_OBS_DIM_TERRAIN: row length of obs_dev / term_obs_dev */

#ifdef __cplusplus
}
#endif
#endif
