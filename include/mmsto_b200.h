/* C-ABI of the batched momentum-mapped space-time optimiser (MMSTO): SURVEY 8f rank 1.
 *
 * Replaces, for a batch of independent characters, one call each of
 *   gym_cdm2/shortTermOptimizer.lua:349-367,368-860   ShortTrajOpt:solveEndEffectors / solveEndEffectors_euler
 * in the online configuration of walk_SRBTrack2025.py:529-587 (onlineOpt = true, frames 0 and 1 constant, no collision
 * half-spaces, weightEEvel = 0), including the L-BFGS loop it drives:
 *   BaseLib/math/optimize.h:426-470 (LBFGS_METHOD, backtracking line search), BaseLib/motion/IK_lbfgs/lbfgs.cpp:246-738.
 * The kinematic tree is LoaderToTree in the Euler-root mode (BaseLib/motion/IK_sdls/NodeWrap.cpp:496-573,1262-1560):
 *   q = [root xyz, root Euler angles YZX, joint angles in tree order], ndof = 6 + #joint angles.
 * All arrays are DEVICE pointers to double unless stated otherwise; one launch solves all problems.  fp64 only (the
 * reference solves in double; weights of 1e5 on squared metres leave no room for fp32). */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmsto_solver mmsto_solver;   /* opaque: skeleton + weights + workspace on one device */

enum { MMSTO_MAX_FRAMES = 8, MMSTO_MARKERS = 4, MMSTO_MAX_VELCON = 4, MMSTO_VELCON_W = 10, MMSTO_INFO_W = 4 };

/* Skeleton tables (HOST) exactly as fk_create takes them (include/bonefk_b200.h); the root bone gets the translational
 * channels XYZ and the rotational channels YZX (VRMLloader.cpp:16,1887).  Limits: nb <= 24, ndof <= 64.
 * markers (shortTermOptimizer.lua `markers`, walk_SRBTrack2025.py:336-347): bone index, local position, weight.
 * nf = number of frames of the window (planningHorizon, <= 8). */
int mmsto_create(int32_t device, int32_t nbones, const int32_t* parent, const double* offset, const int32_t* nchan,
                 const int32_t* axes, const double* mass, const double* local_com, const double* inertia,
                 int32_t nf, const int32_t* marker_bone, const double* marker_lpos, const double* marker_weight,
                 mmsto_solver** out);
int mmsto_destroy(mmsto_solver* s);
int mmsto_ndof(mmsto_solver* s);
/* named scalars: weightDDX weightDX weightAngleOriginal weightEE weightEEcon weightVelHalfSpace weightMomentum weightCOM
 * clampThr (degrees) epsilon max_iterations (0 = unlimited as in libLBFGS; capped by iteration_cap) iteration_cap */
int mmsto_set_param(mmsto_solver* s, const char* name, double value);
int mmsto_get_param(mmsto_solver* s, const char* name, double* value);
/* dof_weights[ndof] (HOST), ShortTrajOpt.dof_weights */
int mmsto_set_dof_weights(mmsto_solver* s, const double* w);

/* One solveEndEffectors_euler per problem.
 *   x_original [n][nf][ndof], dx [n][nf-1][ndof], ddx [n][nf-2][ndof]   desired Euler motion and its derivatives
 *   euler_mot  [n][nf][ndof]    the input motion (only frames 0, 1 are read: the constants)
 *   gpos_target[n][nf][3*4]     marker targets; com [n][nf][3]; momentum [n][nf-1][6] ([M, F] about the CoM)
 *   con        [n] int32        contact bits, bit (marker*8 + frame); used when weightEEcon > 0
 *   velcon     [n][4][10]       velocity half-space constraints: frame (or -1), bone, min speed, normal 3, local pos 3, pad
 *   x_out      [n][nf][ndof]    solution; info [n][4] = final objective, libLBFGS return code, iterations, evaluations
 *                               (a window whose objective turns NaN stops with LBFGSERR_UNKNOWNERROR = -1024 and its last
 *                               finite point; libLBFGS itself has no such check)
 * stream = cudaStream_t as void*.  Returns 0, or a negative error code (mmsto_last_error has the text).
 * A solver owns one workspace: calls on the same handle must be ordered (same stream, or synchronised by the caller); use one
 * handle per concurrent stream. */
int mmsto_solve(mmsto_solver* s, int32_t n, const double* x_original, const double* dx, const double* ddx,
                const double* euler_mot, const double* gpos_target, const double* com, const double* momentum,
                const int32_t* con, const double* velcon, double* x_out, double* info, void* stream);
/* objective and gradient at x [n][nf][ndof] (one evaluation of LBFGS_opta's function, all variables): obj [n],
 * grad [n][nf][ndof]; the momentum Jacobian cache is rebuilt at x as on the first evaluation of a solve. */
int mmsto_evaluate(mmsto_solver* s, int32_t n, const double* x, const double* x_original, const double* dx, const double* ddx,
                   const double* gpos_target, const double* com, const double* momentum, const int32_t* con,
                   const double* velcon, double* obj, double* grad, void* stream);
/* removeCOMsliding_online (gym_trackSRB/taesooLib/DeltaExtractor.py:421-512; called right after the solve,
 * walk_SRBTrack2025.py:617, walk_SRBTrack2025_terrain_singlethreaded.py:610): for every window of nvar (<= 8) poses
 *   poses [n][nvar][ndof + 1]  MotionDOF rows (root xyz, root quaternion w x y z, joint angles), updated IN PLACE
 *   com_srb, desired_com [n][nvar][3]; fix_y_also as in the reference; ground_h [n][nvar] = terrain height under each
 *   current CoM (what projectToGround returns there; get the CoMs with mmsto_window_com) or NULL for no projection
 *   curr_com_out / opt_com_out [n][nvar][3]: optional outputs (currCOM, optCOM). */
int mmsto_remove_com_sliding(mmsto_solver* s, int32_t n, int32_t nvar, double* poses, const double* com_srb, const double* desired_com,
                             int32_t fix_y_also, const double* ground_h, double* curr_com_out, double* opt_com_out, void* stream);
/* CoM of every pose of every window (mLoader_full.setPoseDOF + calcCOM, DeltaExtractor.py:428-430): com_out [n][nvar][3] */
int mmsto_window_com(mmsto_solver* s, int32_t n, int32_t nvar, const double* poses, double* com_out, void* stream);
/* removeFootSliding_online / removeFootSliding_online_stair (gym_trackSRB/taesooLib/DeltaExtractor.py:650-760 / :763-893; called after
 * the CoM step, walk_SRBTrack2025.py:631, walk_SRBTrack2025_terrain_singlethreaded.py:624) for n planning windows of nvar (<= 8) frames:
 *   con_all [n][nvar][4]      contact targets toe L, heel L, toe R, heel R (> 0.5 = contact)
 *   initial_feet [n][12]      the four marker positions the window must start from
 *   marker_traj [n][nvar][12] smooth marker trajectories of the current full-body plan (fc.getMarkerTraj)
 *   marker [n][nvar][12]      the jerky marker predictions (markerTarget columns 0:12)
 *   desired_vel [n][nvar][12] marker velocity targets (markerTarget columns 12:24) or NULL = forward differences of marker_traj * 30
 *   foot_center [n][2][nvar][3]  foot-centre trajectories: non-NULL selects the _stair variant (contact positions from them, hull pass)
 *   ground_mode 0: no projectToGround; 1: flat ground at height ground_y (walk_SRBTrack2025.py:53-54); 2: the playground terrain of the
 *               SRBTrack handle `srb_terrain_env` (an srb_env* of the terrain variant; Y-up query through the Z-up height fields)
 *   out [n][nvar][12]         optMarkerTraj.
 * fc.getContactPositions_v2 / _stair are not in the reference tree; the contact positions follow the stated stand-ins W5 / W6
 * (oracle/foot_sliding.py). */
int mmsto_remove_foot_sliding(int32_t device, int32_t n, int32_t nvar, double foot_len, const double* con_all, const double* initial_feet,
                              const double* marker_traj, const double* marker, const double* desired_vel, const double* foot_center,
                              int32_t ground_mode, double ground_y, void* srb_terrain_env, double* out, void* stream);
/* _extractSingleFrameFullbody (gym_trackSRB/taesooLib/DeltaExtractor.py:577-643): n items, each one decoded full-body row
 * (DeltaExtractor.py:17-22: toPelvis 6 | toFeet 24 | toMomentum 9 | poses 53 | dx 59 | ddx 59 | con 4 = 214 doubles) plus the
 * pack_obs_extra record of its frame (pose_tl 7 | dq 6 | two foot frames 7 | contacts 2 = 29 doubles).  srb_* describe the SRB
 * segment of the caller's loader (mass, 3x3 inertia row-major, local CoM).  Outputs per item: human_pose 60 (root xyz, root
 * quaternion, angles), dx_ddx 118, feetpos 24 (4 marker positions, 4 marker velocities), com 3, momentum 6 ([M, F] about the
 * CoM), con_bin 4, con 4.  Hanyang_lowdof_T layout only (59 dofs), as mmsto_config fixes it. */
int fullbody_extract(int32_t device, int32_t n, const double* extra, const double* row, double srb_mass, const double* srb_inertia9,
                     const double* srb_local_com, double* human_pose, double* dx_ddx, double* feetpos, double* com, double* momentum,
                     double* con_bin, double* con, void* stream);
long long mmsto_launch_count(mmsto_solver* s);
const char* mmsto_last_error(void);

#ifdef __cplusplus
}
#endif
