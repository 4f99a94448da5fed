/* srbtrack_b200.h -- C-ABI of the B200-native batched SRBTrack environment step.
 *
 * This is the drop-in boundary for the data-parallel hot path of the reference
 * (taesoobear/IPCDNNwalk).  Each entry point names the reference interface it replaces
 * (paths relative to the reference repository root).  All functions return 0 on success and a
 * negative code on failure; the message is available from srb_last_error().  No exceptions cross
 * the ABI.  One handle per GPU; a handle is not thread-safe.
 *
 * Pointers suffixed _dev are CUDA device pointers (caller-owned, e.g. torch tensors' data_ptr());
 * pointers suffixed _host are host pointers.  `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream); device entry points are asynchronous on that stream.
 */
#ifndef SRBTRACK_B200_H
#define SRBTRACK_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct srb_env srb_env; /* opaque: N batched `SRBEnv` instances on one device */

#define SRB_OBS_DIM 150   /* observation_space Box(25*6)   env/SRBEnv5_mocap_list_window.py:175 */
#define SRB_ACT_DIM 22    /* action_space Box(22)          env/SRBEnv5_mocap_list_window.py:174 */
#define SRB_RINFO_DIM 8   /* info['reward_info'] entries   env/SRBEnv5_mocap_list_window.py:792-801 */
#define SRB_PLAN_DIM 9    /* reset draws: clip, start frame, vel noise(3), noise quaternion wxyz(4) */
#define SRB_DBG_DIM 256   /* per-env debug record: 4 sub-steps x [qdd(6) nu(6) qfrc(6) iters(1) pad(5) lambda(40)] */

enum { SRB_VARIANT_TEST = 0,      /* env/SRBEnv5_mocap_list_window.py          (SDS foot filter, force clamp) */
       SRB_VARIANT_TRAINING = 1,  /* env/SRBEnv5_mocap_list_window_training.py (what PPO trains on)           */
       SRB_VARIANT_TERRAIN = 2    /* env/SRBEnv5_mocap_list_terrain_window.py  (height-field / pyramid contacts) */ };

typedef struct srb_config {
  int32_t num_envs;       /* replaces make_vec_env(SRBEnv, n_envs=...)  walk_SRBTrack2025_train_delta.py:94-100 */
  int32_t variant;        /* SRB_VARIANT_*                                                                    */
  int32_t precision;      /* 64: fp64 state+arithmetic (reference precision); 32: fp32                        */
  int32_t lanes_per_env;  /* 0 = by batch size (8 up to 6144 envs, 4 up to 16384, else 2); 1,2,4,8,16,32 lanes cooperate on one environment           */
  int32_t auto_reset;     /* 1: a terminated env is reset inside the step (SubprocVecEnv semantics)            */
  int32_t device;         /* CUDA device ordinal                                                               */
  uint64_t seed;          /* stream for device-side reset draws (random.randrange / np.random in SRBEnv.reset) */
} srb_config;

/* SRBEnv.__init__ (env/SRBEnv5_mocap_list_window.py:123-189) for num_envs environments. */
int srb_create(const srb_config* cfg, srb_env** out);
int srb_destroy(srb_env* env);
const char* srb_last_error(void);

/* SRBEnv.loadMocap (env/SRBEnv5_mocap_list_window.py:1171-1197): install the reference tables of one
 * clip.  All arrays are HOST, row-major, nframes = 2*cycle_length-1 rows:
 *   pos_com[n][3], ori_com[n][9], vel_com[n][6], root_quat[n][4] (wxyz), feet_pos[n][2][3] (z==0 <=> contact),
 *   feet_yaw[n][2][2] = (cos, sin) of the pure-yaw feet_ori_list matrices.
 * The 25 look-ahead observation features per frame (get_obs_window :1096-1161) are computed on the device here. */
int srb_upload_clip(srb_env* env, int32_t clip_id, int32_t nframes, int32_t cycle_length, int32_t is_stand,
                    const double* pos_com, const double* ori_com, const double* vel_com, const double* root_quat,
                    const double* feet_pos, const double* feet_yaw);

/* SRBEnv.reset / set_state (env/SRBEnv5_mocap_list_window.py:805-911) for `count` environments.
 * env_idx_host NULL = environments 0..count-1.  plan_host[count][SRB_PLAN_DIM] gives the draws explicitly
 * (noise quaternion w = NaN reproduces testing_mode: no noise); NULL = draw on the device.
 * obs_dev (may be NULL) is the full [num_envs][SRB_OBS_DIM] observation array; rows of the reset envs are written. */
int srb_reset(srb_env* env, const int32_t* env_idx_host, int32_t count, const double* plan_host, float* obs_dev, void* stream);

/* SRBEnv.step (env/SRBEnv5_mocap_list_window.py:208-494) for all environments:
 *   act_dev[N][22] -> obs_dev[N][150], rew_dev[N], done_dev[N] (0/1), rinfo_dev[N][8] (may be NULL). */
int srb_step(srb_env* env, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, float* rinfo_dev, void* stream);

/* Same step through HOST buffers (what a per-process gym caller holds): copies actions host->device and
 * results device->host on the handle's own stream and synchronises.  The handle orders that stream after every reset / step it
 * was asked to launch on a caller stream (an event is recorded there), so reset(); step_host() needs no synchronisation by the caller.  NULL pointers = use the handle's
 * pinned staging buffers in place (see srb_pinned_buffers). */
int srb_step_host(srb_env* env, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host, float* rinfo_host);
int srb_pinned_buffers(srb_env* env, float** act, float** obs, float** rew, uint8_t** done, float** rinfo);
/* How srb_step_host moves data: 1 (default) = zero copy, the step kernel reads the actions from and writes all its results to the
 * mapped pinned buffers directly (contiguous 16-byte-vectorised blocks per warp); 0 = staged cudaMemcpyAsync copies; 2 = split: the
 * look-ahead windows of the observation rows (known before the step) are gathered first and the observation array is copied by a
 * copy engine while the step kernel runs, a patch launch then writes the self features over it; 3 = as 2, but only the window
 * columns travel by copy engine (one pitched copy) and the kernel stores the self features / reward / done / info to the mapped
 * buffers itself.  Measured on one B200 (4096 envs, tools/exp_e2e_modes.py): mode 1 88 us per step, 3: 107 us, 0: 115 us, 2: 131 us --
 * copy-engine traffic and SM-issued stores to host memory slow each other down, so the split forms do not pay; kept selectable. */
int srb_set_host_mode(srb_env* env, int32_t mode);
/* Device timeline of the last mode-2 host step, milliseconds after its first launch was enqueued: out_ms[1] window gather done,
 * [2] observation copy done, [3] step kernel done, [4] patch launch done.  enable = 1 switches the timing events on. */
int srb_host_timeline(srb_env* env, int32_t enable, float out_ms[5]);

/* Optional side channels (any may be NULL = detach):
 *   contact_prob_dev[N][2]  Policy_prior_posterior_MoE_window._last_contact_prob hysteresis input
 *                           (env/SRBEnv5_mocap_list_window.py:239-275); NaN = None
 *   term_obs_dev[N][150]    observation of the terminal step when auto_reset replaces it
 *   reset_plan_dev[N][9]    explicit draws for the next auto-reset of each env (parity tests)
 *   dbg_dev[N][SRB_DBG_DIM] per-sub-step QP intermediates (qdd, nu, qfrc, iterations, lambda) */
int srb_set_aux(srb_env* env, const float* contact_prob_dev, float* term_obs_dev, const double* reset_plan_dev, double* dbg_dev);

/* Reward weights (w_s w_m w_p w_e w_p1 w_p2 w_p3 w_p4 w_c; :561-569) and the external-push schedule
 * (impulse_interval / impulse_duration / force; :182-206).  force[force_len], force_len <= 6, is what update_impulse copies into
 * data.xfrc_applied[1] (`f[0:len(self.force)] = self.force`, :199): world-frame force at the body CoM and, for a 6-vector, a
 * world-frame torque (the push of walk_SRBTrack2025_terrain_singlethreaded.py:117-119: torque = impulse_lpos x force). */
int srb_set_reward_weights(srb_env* env, const double w[9]);
int srb_set_impulse(srb_env* env, int32_t interval, int32_t duration, const double* force, int32_t force_len);

/* SRBEnv.getFullState / restoreFullState (env/SRBEnv5_mocap_list_window.py:1005-1053) of one environment.
 * reals[srb_state_dims real count], ints[int count], hist[slots*width]; layout in srb_state_dims(). */
int srb_state_dims(int32_t* n_real, int32_t* n_int, int32_t* hist_slots, int32_t* hist_width);
int srb_get_state(srb_env* env, int32_t env_index, double* reals, int32_t* ints, double* hist);
int srb_set_state(srb_env* env, int32_t env_index, const double* reals, const int32_t* ints, const double* hist);

/* Running statistics for the one collective of the path (VecNormalize obs_rms, gym_cdm2/train_gym.py:98):
 * acc_dev[1 + 2*150] += (count, sum x, sum x^2) over the N rows of obs_dev.  The caller all-reduces acc. */
int srb_obs_stats(srb_env* env, const float* obs_dev, double* acc_dev, void* stream);

/* The same statistics AND their all-reduce over the ranks of one node in ONE kernel over NVLink / NVSwitch peer memory (replaces
 * srb_obs_stats + srb_episode_stats_device + an NCCL all-reduce, the per-rollout exchange of SURVEY 8e):
 *   srb_xchg_create   allocates this rank's exchange block and returns its 64-byte CUDA IPC handle;
 *   srb_xchg_connect  opens the blocks of all ranks (handles[world][64], gathered by the caller, e.g. with torch.distributed);
 *   srb_rollout_stats_allreduce  acc_dev[1 + 2*150 + 3] = sum over ranks of (count, sum x, sum x^2 | episode return, length, count sums);
 *                     a collective: every rank calls it the same number of times; results are bit-identical on all ranks;
 *   srb_xchg_status   err = 1 if a peer's contribution never arrived (the kernel gives up after ~1 s instead of hanging). */
int srb_xchg_create(srb_env* env, int32_t rank, int32_t world, void* handle_out);
int srb_xchg_connect(srb_env* env, const void* handles);
int srb_rollout_stats_allreduce(srb_env* env, const float* obs_dev, double* acc_dev, void* stream);
int srb_xchg_status(srb_env* env, int32_t* err_out);

/* pack_obs_extra (gym_trackSRB/taesooLib/DeltaExtractor.py:1015-1036): out_dev[N][29] doubles = the SRB state in taesooLib's
 * Y-up convention: pose_tl 7 (translation, quaternion w x y z) | dq 6 (world angular, world linear velocity) | left / right
 * foot frame 7 each (translation, quaternion) | left / right contact flag.  Input of fullbody_extract (include/mmsto_b200.h). */
int srb_pack_obs_extra(srb_env* env, double* out_dev, void* stream);
/* Episode statistics accumulated by auto-reset: out[3] = sum of returns, sum of lengths, count. */
int srb_episode_stats(srb_env* env, double out[3], int32_t clear);
/* The same three sums copied to DEVICE memory on `stream` (no synchronisation): the tail of the vector the per-rollout all-reduce
 * carries next to the observation statistics (episode-return / length counters). */
int srb_episode_stats_device(srb_env* env, double* out_dev, void* stream);

/* ---- terrain variant (env/SRBEnv5_mocap_list_terrain_window.py, model gym_trackSRB/Characters/SRB5_playground.xml) ----
 * The group-0 geoms the reference ray-casts against (terrain_height_ray, env/testSRB_mujoco_original_viewer_terrain.py:741-752):
 * one plane, height fields (elevation data already normalised to [0,1], row 0 at -y, as mujoco stores hfield_data) and
 * axis-aligned boxes (consecutive geoms in model order, grouped by body, tops ascending inside a body). */
typedef struct srb_hfield { const float* data_host; int32_t nrow, ncol; double size[4]; double pos[3]; int32_t geom_id; } srb_hfield;
typedef struct srb_box { double pos[3]; double size[3]; int32_t geom_id; int32_t body_id; } srb_box;
typedef struct srb_terrain_body { double xpos[3]; int32_t body_id; int32_t snap; /* 1: pyramid body followed by a pyramid body (:349,:410) */ } srb_terrain_body;
int srb_upload_terrain(srb_env* env, const double plane_size[2], int32_t plane_geom, int32_t n_hf, const srb_hfield* hfs,
                       int32_t n_box, const srb_box* boxes, int32_t n_body, const srb_terrain_body* bodies);
/* initial_pos_planar of every environment's reference motion (SRBEnv.reset(..., initial_pos_planar), :528,:636); xy_host[N][2].
 * Takes effect at the next reset of each environment. */
int srb_set_planar_offsets(srb_env* env, const double* xy_host);
/* Batched terrain_height_ray: xy_dev[n][2] -> height_dev[n]; hit_dev[n][4] (may be NULL) = geom id, hfield cell column, row, triangle. */
int srb_terrain_height(srb_env* env, int32_t n, const double* xy_dev, double* height_dev, int32_t* hit_dev, void* stream);

/* Device-side terrain description of this handle (NULL before srb_upload_terrain): lets the full-body mapping kernels of
 * include/mmsto_b200.h evaluate projectToGround on the same playground (mmsto_remove_foot_sliding, ground_mode 2). */
const void* srb_terrain_device(void* env);

/* Benchmark / roll-out driver (not in the reference; stands in for the policy, walk_SRBTrack2025.py:424):
 * act_dev[N][22] = reference-tracking action for the next reference frame + N(0, sigma^2) noise. */
int srb_tracking_action(srb_env* env, float sigma, uint64_t seed, uint32_t step, float* act_dev, void* stream);

/* Number of kernels this handle has launched so far (for launch accounting). */
int64_t srb_launch_count(srb_env* env);

#ifdef __cplusplus
}
#endif
#endif /* SRBTRACK_B200_H */
