/* bonefk_b200.h -- C-ABI of the batched full-body forward kinematics + CoM + centroidal momentum pass
 * (BASELINE config 5).  Same library (libsrbtrack_b200.so), same conventions as
 * srbtrack_b200.h: 0 / negative return codes, srb_last_error(), caller-owned device buffers, asynchronous on
 * `stream`.  Replaces, for N poses at once, the taesooLib python-bound classes
 *   BoneForwardKinematics(loader).setPoseDOF(pose); .forwardKinematics(); .globalFrame(i)
 *       work/taesooLib/MainLib/python/MainlibPython.hpp:3171-3197, BaseLib/motion/BoneKinematics.cpp:18-47,97-136
 *   LoaderToTree(loader,..).setPoseDOF(dofInfo,pose); .setDQ(dq); .calcCOM(loader); .calcMomentumCOM(loader)
 *       MainlibPython.hpp:3496-3550, BaseLib/motion/IK_sdls/NodeWrap.cpp:325-386,496-600
 * as used by gym_trackSRB/taesooLib/DeltaExtractor.py:577-643.
 */
#ifndef BONEFK_B200_H
#define BONEFK_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fk_skeleton fk_skeleton;   /* opaque: one skeleton (VRMLloader) resident on one device */

/* Skeleton tables (HOST), bones in depth-first order as written in the .wrl (bone 0 = free root joint):
 *   parent[nb] (-1 for the root), offset[nb][3] (Joint translation), nchan[nb] (0 for the root), axes[nb][3]
 *   (0=X,1=Y,2=Z per Euler channel, jointAxis string), mass[nb], local_com[nb][3] (Segment centerOfMass),
 *   inertia[nb][9] (Segment momentsOfInertia, row-major).  Limits: nb <= 64, tree depth <= 16. */
int fk_create(int32_t device, int32_t nbones, const int32_t* parent, const double* offset, const int32_t* nchan,
              const int32_t* axes, const double* mass, const double* local_com, const double* inertia, fk_skeleton** out);
int fk_destroy(fk_skeleton* sk);
/* Skeletons whose tree structure equals hanyang_lowdof_T(_sh).wrl (20 bones, 53 channels: the skeleton of the SRBTrack
 * full-body mapping) run a skeleton-specialised kernel; fk_set_generic(sk, 1) forces the structure-agnostic kernel (A/B runs,
 * tests).  fk_is_specialised reports which one fk_forward_* will launch. */
int fk_set_generic(fk_skeleton* sk, int32_t force_generic);
int fk_is_specialised(fk_skeleton* sk);
/* The specialised path moves full 64-pose tiles with TMA bulk copies through shared memory.  mode -1 (default) = measured
 * best variant per precision / outputs; 0 = rows accessed directly by the owning thread; 1 = rolled bone loop on TMA tiles;
 * 3 = unrolled sweep on TMA tiles (A/B runs, tests). */
int fk_set_tma(fk_skeleton* sk, int32_t use_tma);
/* pose / dq row lengths: 7 + #angles (root xyz, root quaternion wxyz, angles) and 6 + #angles
 * (root angular velocity world, root linear velocity world, joint rates). */
int fk_dims(fk_skeleton* sk, int32_t* nbones, int32_t* n_pose, int32_t* n_dq);

/* n poses: pose_dev[n][n_pose], dq_dev[n][n_dq] (NULL: no velocities) ->
 *   xform_dev[n][nbones][7] global (quaternion wxyz, translation) per bone (may be NULL),
 *   com_dev[n][3] (may be NULL), mom_dev[n][6] centroidal momentum (angular about the CoM, linear) (may be NULL). */
int fk_forward_f32(fk_skeleton* sk, int64_t n, const float* pose_dev, const float* dq_dev, float* xform_dev, float* com_dev, float* mom_dev, void* stream);
int fk_forward_f64(fk_skeleton* sk, int64_t n, const double* pose_dev, const double* dq_dev, double* xform_dev, double* com_dev, double* mom_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
