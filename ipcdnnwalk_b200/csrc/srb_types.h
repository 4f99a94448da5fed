// Shared host/device declarations of the SRBTrack batched step (not part of the public C-ABI).
#pragma once
#include <stdint.h>

namespace srb {

// ---- per-environment state, structure-of-arrays: st[field * Npad + env] ------------------------
enum RealField {
  F_POS = 0,      // 3  data.qpos[0:3]
  F_QUAT = 3,     // 4  data.qpos[3:7]   (w,x,y,z) as integrated by mj_step2
  F_QVEL = 7,     // 6  data.qvel        (world linear, body-frame angular)
  F_QLAG = 13,    // 4  quaternion data.xmat was last computed from (one sub-step stale)
  F_FOOTL = 17,   // 3  left_foot_contact_pos_global
  F_FOOTR = 20,   // 3  right_foot_contact_pos_global
  F_FYAWL = 23,   // 2  left_foot_contact_ori as (cos, sin) of its yaw
  F_FYAWR = 25,   // 2
  F_FILTYAW = 27, // 2  contactFilter[0:2]
  F_BALLX = 29,   // 4  SDS position state  (left x, left y, right x, right y)
  F_BALLV = 33,   // 4  SDS velocity state
  F_EPRET = 37,   // 1  running episode return
  F_NU = 38,      // 6  QP dual of the last sub-step (warm start)
  F_PLANAR = 44,  // 2  initial_pos_planar of this environment's reference (terrain variant)
  NF_REAL = 46
};
enum IntField {
  I_FLAGS = 0,    // bit0 left_foot_in_contact, bit1 right_foot_in_contact, bit2 filter initialised, bit3 is_stand
  I_FRAME = 1,    // current_frame
  I_RIF = 2,      // random_initial_frame
  I_CLIP = 3,
  I_FCOUNT = 4,   // frame_counter (impulse schedule)
  I_IMPULSE = 5,  // remaining impulse sub-steps
  I_EPISODE = 6,  // episodes started (RNG stream position)
  NF_INT = 7
};
enum { FLAG_L = 1, FLAG_R = 2, FLAG_FILT = 4, FLAG_STAND = 8, FLAG_NU = 16 };

// history ring: hist[(env * HIST_SLOTS + slot) * HIST_W + k] ; entry e lives in slot e % HIST_SLOTS
enum { HIST_SLOTS = 32, HIST_W = 8 };   // pos(3) quat_lag(4) pad(1)

// ---- reference clip tables ----------------------------------------------------------------------
// one record per frame, CLIP_W reals:
//   0:3 pos_com | 3:12 ori (row-major) | 12:18 vel_com | 18:21 left foot pos (z==0 <=> contact)
//   21:24 right foot pos | 24:26 left foot yaw (cos,sin) | 26:28 right foot yaw | 28:32 root quat (wxyz)
enum { CLIP_W = 32, FEAT_W = 25, OBS_W = 150, ACT_W = 22, RINFO_W = 8 };

template <typename real> struct ClipDesc {
  const real* frames;   // [nframes][CLIP_W]
  const float* feat;    // [nframes][FEAT_W] look-ahead observation features, precomputed on device
  int nframes;          // 2*cycle_length-1
  int cycle_length;
  int is_stand;
  int pad;
};

enum Variant { VARIANT_TEST = 0, VARIANT_TRAINING = 1, VARIANT_TERRAIN = 2 };

// ---- terrain of SRB5_playground.xml: plane + height fields + axis-aligned box stacks (pyramids) -------
struct HFieldDev { const float* data; int nrow, ncol; double sx, sy, sz, px, py, pz; int geom_id; int pad; };
struct TerrainBodyDev { double cx, cy; double x0, x1, y0, y1; int first, count, snap, body_id; int nested, pad; };   // nested: concentric boxes, half sizes shrink with the index
enum { MAX_HF = 4, MAX_TBODY = 16 };
struct TerrainDev {
  int enabled, n_hf, n_body, n_box;
  int plane_geom, box_geom0;
  double plane_hx, plane_hy;
  HFieldDev hf[MAX_HF];
  TerrainBodyDev body[MAX_TBODY];
  const double* box;   // [n_box][5] cx cy hx hy top, model order (ascending tops inside a body)
};

template <typename real> struct Params {
  real h;               // model.opt.timestep
  int nsub;             // round(mocap_dt / timestep)
  real mass, inertia[3], armature, damping, grav;
  real mu_b;            // 1.2 * self.mu = 1.2 * (1.2 * floor friction)
  real kp, kd, w_lambda;
  real inv_w_lambda, inv_M[6], inv_Mh[6];   // 1/w, 1/(M+armature), 1/(M+armature+h*damping)
  real force_thr;
  real clamp_bound;     // 1.01 * sqrt(1 + mu_b^2): |F_foot| <= clamp_bound * F_total,z (see the force clamp in srb_step_kernel)
  real w[9];            // w_s w_m w_p w_e w_p1 w_p2 w_p3 w_p4 w_c
  int use_filter, use_clamp, use_impulse, obs_after_inc, use_terrain;
  int impulse_interval, impulse_duration, impulse_trigger_at_end;
  real impulse_force[6];   // xfrc_applied[1]: world-frame force, world-frame torque (f[0:len(self.force)] = self.force, :199)
  int max_frame;        // 512
  int aset_warm;        // 1: a sub-step's QP starts from the active set the previous sub-step ended with (same optimum, fewer systems)
};

// debug record per env (doubles), only written when a debug buffer is attached
enum { DBG_SUB = 64, DBG_W = 4 * DBG_SUB };   // per sub-step: qdd(6) nu(6) qfrc(6) iters(1) pad(5) lam(40)

template <typename real> struct StepArgs {
  int N, Npad;
  real* st; int* ist; real* hist;
  const ClipDesc<real>* clips; int nclips;
  const float* act; float* obs; float* rew; uint8_t* done; float* rinfo; float* term_obs;
  const float* contact_prob;     // [N][2] or null (NaN entry == None)
  const double* reset_plan;      // [N][9] or null: clip, frame, vel_noise(3), noise_quat(4) for the next auto-reset
  double* dbg;                   // [N][DBG_W] or null
  double* ep_stats;              // [3] sum of episode returns, sum of episode lengths, episode count (atomicAdd)
  unsigned long long seed;
  int auto_reset;
  const TerrainDev* terrain;     // device pointer (terrain variant) or null
  int split_obs;                 // coalesced-I/O kernel only: the look-ahead window of the observation rows is delivered separately
                                 // (srb_obs_window_kernel + copy engine); the kernel writes the 25 self features of every row only
  uint8_t* done_dev;             // optional device copy of the done flags (the host path's window fix-up reads it) or null
  Params<real> p;
};

template <typename real> struct ResetArgs {
  int N, Npad, count;
  const int* idx;                // [count] env indices or null (= all)
  real* st; int* ist; real* hist;
  const ClipDesc<real>* clips; int nclips;
  const double* plan;            // [count][9] explicit draws or null (=> device RNG)
  float* obs;                    // [N][OBS_W] rows of the reset envs are written, may be null
  unsigned long long seed;
  const TerrainDev* terrain;
  Params<real> p;
};

}  // namespace srb
