// Shared layout constants of the AdaptiveSRB (gym_cdm2 `deepmimiccdm-v1`) batched step: SURVEY 8a rows a-16 ... a-20.
#pragma once
#include <stdint.h>

namespace cdm {

enum {
  ACT_W = 10,          // [R foot xz, L foot xz, d(lin vel) 3, d(ang vel) 3]          (deepmimiccdm-v1.lua:434)
  OBS_W = 27,          // 11 + 6*numLegs + 2 + 2                                        (:435)
  OBS_W_TERRAIN = 31,  // + 4 terrain height samples                                    (deepmimiccdmterrain-v1.lua:14-16)
  TAB_W = 48,          // reference-table row, see ipcdnnwalk_b200/cdm_tables.py: pack_rows
  HIST_SLOTS = 8, HIST_W = 8,   // replay buffer ring (RagdollSim.lua:128-143): root pose of the last steps
  DBG_W = 96,          // qdd 2x6 | a_des 2x6 | lambda 2x20 | pose_dif com_lvdif ref_comvel timing_mod | theta_d 7 | iters 2 | ncontacts
  PLAN_W = 16,         // reset draws: quaternion noise 4, position noise 3, dq noise 6, extra linear-velocity noise 3
  NQCOL = 20,          // unique friction-basis columns: 2 legs x 2 points x 5 directions (each is pushed twice, Q7)
};

// table row columns
enum { T_CDM = 0, T_DV = 7, T_DW = 10, T_OFFR = 13, T_OFFL = 16, T_LEG = 19, T_IFR = 40, T_SWDUR = 41, T_RTTD = 43, T_RTTO = 45, T_CON = 47 };

// structure-of-arrays environment state: st[field][Npad]
enum {
  C_P = 0, C_Q = 3, C_V = 7, C_WB = 10, C_IFRAME = 13, C_GTIME = 14, C_TGT = 15, C_PREVCOM = 19, C_EPRET = 22,
  C_LEG = 23, LEG_W = 19,      // per leg: contactPos 3, contactPreFilter 3, particle x 3, particle v 3, yaw SDS (x, v) 2,
                               //          lqrFilterQorigin 4, swingPhase 1
  L_CPOS = 0, L_PRE = 3, L_BX = 6, L_BV = 9, L_FQ = 12, L_ORG = 14, L_SWPH = 18,
  NF_REAL = C_LEG + 2 * LEG_W,
};
enum { I_FLAGS = 0, I_NREP = 1, I_EPISODE = 2, I_STEPS = 3, NF_INT = 4 };
enum { FLAG_SW_R = 1, FLAG_SW_L = 2 };

template <typename real> struct Params {
  real k_p, k_d, acc_clamp, max_leg_len, max_foot_vel, stride_thr, max_contact_len, min_swing_phase, max_contact_y;
  real contact_thr, touch_down_height; int unlimited_leglen;   // deepmimiccdm-v1.lua:702 (late touch-down when the leg length is limited: spec/run2-v1.lua); stride_thr < 0 = no strideThr
  real healthy_min, healthy_max, fall_angle, timing_mod_coef, facing_weight, ef_scale, limit_length, max_turning, filter_end_phase;
  real K0, K1;                 // LQR gain of the contact filter (M=30, b=1e-3, k=1, Q=10^logQ): constant, see cdm_api.cu
  double inv_ridge2;           // 2 * w_ddq / w_lam : the ridge of the de-duplicated QP (Q7), always fp64
  real mass, inertia[3], grav;
  real basis[5][3];            // friction basis directions (QPservo_v3.lua:956-968)
  real root_off[3];            // translation of the root joint in the skeleton file
  real leg_off[6][3];          // joint offsets: L hip, L knee, L ankle, R hip, R knee, R ankle
  real toe[3], heel[3];
  int leg_nchan[6], leg_axes[6];   // channels per joint and their axes packed 2 bits each (0 X, 1 Y, 2 Z), first channel lowest
  real noise_q, noise_pos, noise_dq, noise_v[3];
};

// constants derived from the table at upload (reset state = frame `start` of the clip)
template <typename real> struct ResetConst {
  real cdm0[7], dcdm0[7];
  real contact0[2][3];         // mid-ankle markers of the reference at the start frame, relative to the root, y = 0
  int start_swing[2];
};

// height field of the gym_cdm2 terrain spec (OBJloader::Terrain); h == nullptr: flat ground
struct TLTerrainDev { const double* h; int rows, cols; double size_x, size_z; int tile; };
struct Ground { TLTerrainDev T; double px, py, pz; };

template <typename real> struct StepArgs {
  int N, Npad, nframes, nf_cycle, obs_w;
  Ground ground;
  real* st; int* ist; real* hist;
  const real* tab;
  const float* act; float* obs; float* rew; uint8_t* done;
  float* term_obs; double* ep_stats; const double* reset_plan; double* dbg;
  int auto_reset;
  int gpw;                      // lane groups (= environments) per warp actually used, <= 32 / G; the remaining lanes shadow group 0
  unsigned long long seed;
  Params<real> p;
  ResetConst<real> rc;
};

template <typename real> struct ResetArgs {
  int N, Npad, count, nframes, nf_cycle, obs_w;
  Ground ground;
  const int* idx;
  real* st; int* ist; real* hist;
  const real* tab;
  const double* plan;
  unsigned long long seed;
  float* obs;
  Params<real> p;
  ResetConst<real> rc;
};

}  // namespace cdm
