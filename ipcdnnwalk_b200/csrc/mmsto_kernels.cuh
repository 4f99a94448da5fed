// Batched momentum-mapped space-time optimiser (MMSTO): gym_cdm2/shortTermOptimizer.lua:349-860 in the
// online configuration of walk_SRBTrack2025.py:529-587, with the L-BFGS loop of BaseLib/motion/IK_lbfgs/lbfgs.cpp:246-738
// inside the kernel.
//
// Mapping: 8 lanes per problem, lane = frame of the window (nf <= 8), 4 problems per warp, one warp per CTA.  A lane owns
// the kinematic tree of its frame (LoaderToTree in the Euler-root mode: one slide / hinge node per dof); everything that
// couples frames -- finite-difference velocities, the Gaussian-filtered marker / CoM residuals, the momentum Jacobian
// averaged over the window, the L-BFGS dot products -- is an 8-lane shuffle.  The L-BFGS vectors (x, g, d, xp, gp, 6 s,
// 6 y, the targets) live in a per-warp workspace laid out [vector][dof][32 lanes]: every access of the warp is one
// contiguous 256-byte row.  Shared memory per warp (53 kB, 4 warps per SM): bone frames as quaternion + origin
// [bone][7][32 lanes], a depth-indexed stack [8][6][32 lanes] for parent velocities (forward sweep) and child force /
// torque sums (reverse sweep), and the per-problem Gram matrices of the L-BFGS recursion.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <math.h>

namespace mmsto {

enum { MAXB = 24, MAXD = 64, NFMAX = 8, NMARK = 4, MAXHS = 4, HS_W = 10, LB_M = 6, GRP = 8, PPW = 4,
       V_X = 0, V_G, V_D, V_XP, V_GP, V_XO, V_DX, V_DDX, V_SUMO, V_S, V_Y = V_S + LB_M, NV = V_Y + LB_M,
       SM_W = 7, SM_Q = 0, SM_T = 4, STK_LEVELS = 8, STK_W = 6,
       GM_CG = 144, GM_W = 152 };   // per-problem L-BFGS scalars in shared memory: Gram matrices 3 x 36, g.S, g.Y, coefficients, alpha, ys

enum { LBFGS_SUCCESS = 0, LBFGS_ALREADY_MINIMIZED = 2, LBFGSERR_MINIMUMSTEP = -1000, LBFGSERR_MAXIMUMSTEP = -999,
       LBFGSERR_MAXIMUMLINESEARCH = -998, LBFGSERR_MAXIMUMITERATION = -997, LBFGSERR_INCREASEGRADIENT = -994,
       LBFGSERR_UNKNOWNERROR = -1024 };

struct Skel {
  int nb, ndof;
  int parent[MAXB], nchan[MAXB], axis[MAXB][3], dof0[MAXB], sub_end[MAXB], depth[MAXB];
  double off[MAXB][3], mass[MAXB], com[MAXB][3], inertia[MAXB][6];   // inertia: xx yy zz xy xz yz
  double total_mass;
};

struct Prm {
  int nf, max_iter, iter_cap, max_ls;
  double w_ddx, w_dx, w_ao, w_ee, w_eecon, w_velhs, w_mom, w_com;
  double s_ddx[MAXD], s_dx[MAXD], s_ao[MAXD];      // sqrt(weight * dof_weight)  (OperatorStitch.cpp:286)
  int mbone[NMARK];
  double mlpos[NMARK][3], mweight[NMARK];
  double filt[NFMAX];                              // Filter::GetGaussFilter(nf)
  double clamp, eps, ftol, wolfe, min_step, max_step;
};

struct Args {
  int n;
  const double *x_original, *dx, *ddx, *euler_mot, *gpos_target, *com, *momentum, *velcon, *x_in;
  const int* con;
  double *x_out, *info, *obj, *grad;
  double *ws, *mc;
  int* queue;                     // next unsolved problem (solve only): the 8-lane groups of the persistent warps pull from it
  const Skel* sk;
  const Prm* prm;
};

__device__ __forceinline__ double grp_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void mat_vec(const double* R, const double* v, double* o) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2]; o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2]; o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
__device__ __forceinline__ void matT_vec(const double* R, const double* v, double* o) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2]; o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2]; o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
// R <- R * Rot(axis a, angle): columns u = a+1, v = a+2 (cyclic) mix
__device__ __forceinline__ void rot_axis(double* R, int a, double s, double c) {
  const int u = a == 2 ? 0 : a + 1, v = u == 2 ? 0 : u + 1;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double cu = R[r * 3 + u], cv = R[r * 3 + v];
    R[r * 3 + u] = cu * c + cv * s;
    R[r * 3 + v] = cv * c - cu * s;
  }
}

struct Ctx {
  const Skel* sk; const Prm* P;
  double* V;       // this warp's vector workspace
  double* MC;      // this warp's momentum-Jacobian cache [dof][6][PPW]
  double* sm;      // shared memory: bone frames
  double* gm;      // shared memory: this problem's L-BFGS scalars
  int lane, f, pg; bool valid;
  double tgt[3 * NMARK], comt[3], momt[6];
  int con;
  const double* velcon;   // this problem's [MAXHS][HS_W]
};

#define VEC(ctx, v, j) (ctx).V[((size_t)(v) * MAXD + (j)) * 32 + (ctx).lane]
#define SMB(ctx, b, k) (ctx).sm[((b) * SM_W + (k)) * 32 + (ctx).lane]
#define STK(ctx, d, k) (ctx).sm[(((ctx).sk->nb * SM_W) + (d) * STK_W + (k)) * 32 + (ctx).lane]

// unit quaternion (w x y z) -> rotation matrix, row major
__device__ __forceinline__ void quat_to_mat(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0 - 2.0 * (yy + zz); R[1] = 2.0 * (xy - wz); R[2] = 2.0 * (xz + wy);
  R[3] = 2.0 * (xy + wz); R[4] = 1.0 - 2.0 * (xx + zz); R[5] = 2.0 * (yz - wx);
  R[6] = 2.0 * (xz - wy); R[7] = 2.0 * (yz + wx); R[8] = 1.0 - 2.0 * (xx + yy);
}
// column a of the rotation matrix of q: the world direction of the local axis a
__device__ __forceinline__ void quat_col(const double* q, int a, double* W) {
  const int u = a == 2 ? 0 : a + 1, v = u == 2 ? 0 : u + 1;
  const double w = q[0], qa = q[1 + a], qu = q[1 + u], qv = q[1 + v];
  W[a] = 1.0 - 2.0 * (qu * qu + qv * qv); W[u] = 2.0 * (qa * qu + w * qv); W[v] = 2.0 * (qa * qv - w * qu);
}
// q <- q * (cos, sin e_a): right-multiplied rotation about the local axis a by the angle whose HALF has sine s, cosine c
__device__ __forceinline__ void quat_rot_axis(double* q, int a, double s, double c) {
  const int u = a == 2 ? 0 : a + 1, v = u == 2 ? 0 : u + 1;
  const double w = q[0], qa = q[1 + a], qu = q[1 + u], qv = q[1 + v];
  q[0] = w * c - qa * s; q[1 + a] = qa * c + w * s; q[1 + u] = qu * c + qv * s; q[1 + v] = qv * c - qu * s;
}
__device__ __forceinline__ void load_quat(const Ctx& c, int b, double* q) {
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = SMB(c, b, SM_Q + k);
}
__device__ __forceinline__ void load_frame(const Ctx& c, int b, double* R, double* t) {
  double q[4];
  load_quat(c, b, q);
  quat_to_mat(q, R);
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = SMB(c, b, SM_T + k);
}
// world axis of channel ch of bone b (HingeJoint::ComputeW).  Only the bone frames are kept (as quaternions): the last
// channel's axis is a column of the bone frame, the first one a column of the parent frame, and a middle one is rebuilt
// from the parent frame and the preceding angle(s)
__device__ __forceinline__ void hinge_axis(const Ctx& c, int b, int ch, double* W) {
  const Skel& sk = *c.sk;
  const int a = sk.axis[b][ch], nc = sk.nchan[b];
  double q[4];
  if (ch == nc - 1) load_quat(c, b, q);
  else {
    const int p = sk.parent[b];
    if (p < 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else load_quat(c, p, q);
    for (int k = 0; k < ch; ++k) {
      double sn, cs;
      sincos(0.5 * (c.valid ? VEC(c, V_X, sk.dof0[b] + k) : 0.0), &sn, &cs);
      quat_rot_axis(q, sk.axis[b][k], sn, cs);
    }
  }
  quat_col(q, a, W);
}

// momentum about `C` of bone b moving with world angular velocity w and origin velocity v (Liegroup.cpp:66-74 Inertia * se3,
// then dAd to the world frame): adds to Hm (about the WORLD origin) and Hf
__device__ __forceinline__ void bone_momentum(const Skel& sk, int b, const double* R, const double* t, const double* w, const double* v,
                                               double* Hm, double* Hf) {
  double wb[3], vb[3], r[3], M[3], F[3], tmp[3];
  matT_vec(R, w, wb); matT_vec(R, v, vb);
  const double m = sk.mass[b];
  r[0] = m * sk.com[b][0]; r[1] = m * sk.com[b][1]; r[2] = m * sk.com[b][2];
  const double* I = sk.inertia[b];
  cross3(r, vb, tmp);
  M[0] = I[0] * wb[0] + I[3] * wb[1] + I[4] * wb[2] + tmp[0];
  M[1] = I[3] * wb[0] + I[1] * wb[1] + I[5] * wb[2] + tmp[1];
  M[2] = I[4] * wb[0] + I[5] * wb[1] + I[2] * wb[2] + tmp[2];
  cross3(wb, r, tmp);
  F[0] = tmp[0] + m * vb[0]; F[1] = tmp[1] + m * vb[1]; F[2] = tmp[2] + m * vb[2];
  double Fw[3], Mw[3];
  mat_vec(R, F, Fw); mat_vec(R, M, Mw);
  cross3(t, Fw, tmp);
#pragma unroll
  for (int k = 0; k < 3; ++k) { Hm[k] += Mw[k] + tmp[k]; Hf[k] += Fw[k]; }
}

// One evaluation of LBFGS_opta's function at V_X: objective (same value on the 8 lanes of a problem) and, in V_G, the
// gradient of this lane's frame.  `first` (per problem; `any_first` = some problem of the warp): rebuild the momentum-Jacobian
// cache (shortTermOptimizer.lua:669-684).
__device__ double evaluate(Ctx& c, bool any_first, bool first) {
  const Skel& sk = *c.sk; const Prm& P = *c.P;
  const int nf = P.nf, nd = sk.ndof, f = c.f;
  const bool valid = c.valid, has_next = valid && f + 1 < nf;
  const double INV = 30.0, INVS = 900.0;
  double o = 0.0;

  // ---------------- forward sweep: Tree::Compute + ComputeDS + calcCOM + calcMomentumCOM of this frame
  double comacc[3] = {0, 0, 0}, Hm[3] = {0, 0, 0}, Hf[3] = {0, 0, 0};
  double qn[3], dqn[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const int j = sk.dof0[0] + ch;
    qn[ch] = valid ? VEC(c, V_X, j) : 0.0;
    dqn[ch] = has_next ? c.V[((size_t)V_X * MAXD + j) * 32 + c.lane + 1] : 0.0;
  }
  for (int b = 0; b < sk.nb; ++b) {
    double q[4], R[9], t[3], w[3], v[3];
    const int p = sk.parent[b], dep = sk.depth[b];
    if (p < 0) {
      q[0] = 1; q[1] = q[2] = q[3] = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double x0 = valid ? VEC(c, V_X, k) : 0.0;
        t[k] = sk.off[b][k] + x0;
        v[k] = has_next ? (c.V[((size_t)V_X * MAXD + k) * 32 + c.lane + 1] - x0) * INV : 0.0;
        w[k] = 0.0;
      }
    } else {
      double tp[3], wp[3], vp[3], d[3], x[3];
      load_quat(c, p, q);
      quat_to_mat(q, R);
#pragma unroll
      for (int k = 0; k < 3; ++k) { tp[k] = SMB(c, p, SM_T + k); wp[k] = STK(c, dep - 1, k); vp[k] = STK(c, dep - 1, 3 + k); }
      mat_vec(R, sk.off[b], d);
      cross3(wp, d, x);
#pragma unroll
      for (int k = 0; k < 3; ++k) { t[k] = tp[k] + d[k]; w[k] = wp[k]; v[k] = vp[k] + x[k]; }
    }
    const int nc = sk.nchan[b];
    double qv[3], dqv[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { qv[ch] = qn[ch]; dqv[ch] = dqn[ch]; }
    if (b + 1 < sk.nb) {                                   // the angles of the next bone are fetched while this one is computed
      const int jb = sk.dof0[b + 1], ncn = sk.nchan[b + 1];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const int j = jb + (ch < ncn ? ch : 0);
        qn[ch] = valid ? VEC(c, V_X, j) : 0.0;
        dqn[ch] = has_next ? c.V[((size_t)V_X * MAXD + j) * 32 + c.lane + 1] : 0.0;
      }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (ch < nc) {
        const int a = sk.axis[b][ch];
        double W[3];
        quat_col(q, a, W);
        const double th = qv[ch];
        const double dq = has_next ? (dqv[ch] - th) * INV : 0.0;
        double sn, cs;
        sincos(0.5 * th, &sn, &cs);
        quat_rot_axis(q, a, sn, cs);
        w[0] += W[0] * dq; w[1] += W[1] * dq; w[2] += W[2] * dq;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) SMB(c, b, SM_Q + k) = q[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { SMB(c, b, SM_T + k) = t[k]; STK(c, dep, k) = w[k]; STK(c, dep, 3 + k) = v[k]; }
    quat_to_mat(q, R);
    double cb[3];
    mat_vec(R, sk.com[b], cb);
#pragma unroll
    for (int k = 0; k < 3; ++k) comacc[k] += (cb[k] + t[k]) * sk.mass[b];
    bone_momentum(sk, b, R, t, w, v, Hm, Hf);
  }
  double C[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) C[k] = comacc[k] * (1.0 / sk.total_mass);

  // ---------------- momentum Jacobian cache: mean over frames 0 .. nf-2 of calcMomentumJacobianTranspose
  if (any_first && P.w_mom > 0) {                      // warp-uniform: every lane takes part in the reductions
    const bool use = valid && f < nf - 1;
    for (int j = 0; j < nd; ++j) {
      double acc[6] = {0, 0, 0, 0, 0, 0};
      int b0 = 0, ch = 0;
      bool slide = j < 3;
      if (!slide) {
        for (b0 = 0; b0 < sk.nb; ++b0) if (j >= sk.dof0[b0] && j < sk.dof0[b0] + sk.nchan[b0]) break;
        ch = j - sk.dof0[b0];
      }
      double W[3], sj[3];
      if (slide) { W[0] = j == 0; W[1] = j == 1; W[2] = j == 2; sj[0] = sj[1] = sj[2] = 0; }
      else {
        hinge_axis(c, b0, ch, W);
#pragma unroll
        for (int k = 0; k < 3; ++k) sj[k] = SMB(c, b0, SM_T + k);
      }
      for (int b = b0; b < sk.sub_end[b0]; ++b) {
        double R[9], t[3], w[3], v[3], d[3], hm[3] = {0, 0, 0}, hf[3] = {0, 0, 0};
        load_frame(c, b, R, t);
        if (slide) { w[0] = w[1] = w[2] = 0; v[0] = W[0]; v[1] = W[1]; v[2] = W[2]; }
        else {
#pragma unroll
          for (int k = 0; k < 3; ++k) { w[k] = W[k]; d[k] = t[k] - sj[k]; }
          cross3(W, d, v);
        }
        bone_momentum(sk, b, R, t, w, v, hm, hf);
        double x[3];
        cross3(C, hf, x);
#pragma unroll
        for (int k = 0; k < 3; ++k) { acc[k] += hm[k] - x[k]; acc[3 + k] += hf[k]; }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double s = grp_sum(use ? acc[k] : 0.0) * (1.0 / (nf - 1));
        if (first && f == 0) c.MC[((size_t)j * 6 + k) * PPW + c.pg] = s;
      }
    }
    __syncwarp();
  }

  // ---------------- residuals that couple the frames
  double mom_c[6] = {0, 0, 0, 0, 0, 0};      // (-2 w filt[f] invDT) (desired - actual momentum) of interval f
  if (P.w_mom > 0) {
    double x[3];
    cross3(C, Hf, x);
    const bool use = valid && f < nf - 1;
    const double fl = use ? P.filt[f] : 0.0;
    double ss = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dm = c.momt[k] - (Hm[k] - x[k]), df = c.momt[3 + k] - Hf[k];
      ss += dm * dm + df * df;
      mom_c[k] = (-2.0 * P.w_mom * fl * INV) * dm; mom_c[3 + k] = (-2.0 * P.w_mom * fl * INV) * df;
    }
    if (use) o += ss * P.w_mom * fl;
  }
  // point forces of this frame, applied when the reverse sweep visits their bone
  double Fm[NMARK][3], cfh[MAXHS], Fcom[3] = {0, 0, 0};
  bool com_on = false;
#pragma unroll
  for (int h = 0; h < MAXHS; ++h) cfh[h] = 0.0;

  // marker terms (:713-784)
#pragma unroll
  for (int m = 0; m < NMARK; ++m) {
    const int b = P.mbone[m];
    double R[9], t[3], pm[3], F[3] = {0, 0, 0};
    load_frame(c, b, R, t);
    mat_vec(R, P.mlpos[m], pm);
#pragma unroll
    for (int k = 0; k < 3; ++k) pm[k] += t[k];
    const double fl = valid ? P.filt[f] : 0.0;
    const bool cn = valid && ((c.con >> (m * 8 + f)) & 1);
    double dS[3], dC[3];
    const double cnt = grp_sum(cn ? 1.0 : 0.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double e = c.tgt[3 * m + k] - pm[k];
      dS[k] = grp_sum(valid ? e * fl : 0.0);
      dC[k] = grp_sum(cn ? e : 0.0);
    }
    if (P.w_ee > 0) {
      const double w = P.w_ee * P.mweight[m];
      if (f == 0) o += dot3(dS, dS) * w;
#pragma unroll
      for (int k = 0; k < 3; ++k) F[k] += -(2.0 * w * fl) * dS[k];
    }
    if (P.w_eecon > 0 && cnt > 0) {
      const double w = P.w_eecon * P.mweight[m];
#pragma unroll
      for (int k = 0; k < 3; ++k) dC[k] *= 1.0 / cnt;
      if (f == 0) o += dot3(dC, dC) * w;
      if (cn) {
#pragma unroll
        for (int k = 0; k < 3; ++k) F[k] += -(2.0 * w / cnt) * dC[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) Fm[m][k] = valid ? F[k] : 0.0;
  }
  // velocity half-space terms (:466-503): frame fc against frame fc - 1, fc >= 2
  if (P.w_velhs != 0) {
#pragma unroll
    for (int h = 0; h < MAXHS; ++h) {
      const double* hc = c.velcon + h * HS_W;
      const int fc = (int)hc[0];
      const bool act = fc >= 2 && fc < nf;               // uniform over the problem's lanes, not over the warp
      const int b = min(max((int)hc[1], 0), sk.nb - 1);
      double R[9], t[3], pm[3];
      load_frame(c, b, R, t);
      mat_vec(R, hc + 6, pm);
#pragma unroll
      for (int k = 0; k < 3; ++k) pm[k] += t[k];
      double gv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double prev = __shfl_up_sync(0xffffffffu, pm[k], 1, GRP), next = __shfl_down_sync(0xffffffffu, pm[k], 1, GRP);
        gv[k] = f == fc ? (pm[k] - prev) * INV : (next - pm[k]) * INV;
      }
      const double* nrm = hc + 3;
      const double d = (nrm[0] * hc[2] - gv[0]) * nrm[0] + (nrm[1] * hc[2] - gv[1]) * nrm[1] + (nrm[2] * hc[2] - gv[2]) * nrm[2];
      if (act && valid && d > 0 && (f == fc || f == fc - 1)) {
        cfh[h] = (f == fc ? -1.0 : 1.0) * (2.0 * P.w_velhs * INV * d);      // force = cfh * normal at the constrained point
        if (f == fc) o += d * d * P.w_velhs;
      }
    }
  }
  // CoM term (:786-809)
  if (P.w_com > 0) {
    const int fs = max(nf / 2, 2);
    const bool use = valid && f < fs;
    double dS[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dS[k] = grp_sum(use ? c.comt[k] - C[k] : 0.0) * (1.0 / fs);
    if (f == 0) o += dot3(dS, dS) * P.w_com;
    if (use) {
      com_on = true;
#pragma unroll
      for (int k = 0; k < 3; ++k) Fcom[k] = (-2.0 * P.w_com / fs) * dS[k] * (1.0 / sk.total_mass);      // times the bone mass, at the bone CoM
    }
  }

  // ---------------- reverse sweep: g_j = W_j . (N - s_j x F) for hinges, W_j . F for slides (Node.cpp:386-400,466-469).
  // Bones are stored depth first, so when bone b is reached level depth[b] of the stack holds exactly the sum over its children.
  for (int d = 0; d < STK_LEVELS; ++d)
#pragma unroll
    for (int k = 0; k < STK_W; ++k) STK(c, d, k) = 0.0;
  for (int b = sk.nb - 1; b >= 0; --b) {
    const int dep = sk.depth[b];
    double F[3], N[3], R[9], t[3], x[3], tau[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { F[k] = STK(c, dep, k); N[k] = STK(c, dep, 3 + k); STK(c, dep, k) = 0.0; STK(c, dep, 3 + k) = 0.0; }
    load_frame(c, b, R, t);
    auto push = [&](const double* lp, const double* Fp) {
      double pw[3], n[3];
      mat_vec(R, lp, pw);
#pragma unroll
      for (int k = 0; k < 3; ++k) pw[k] += t[k];
      cross3(pw, Fp, n);
#pragma unroll
      for (int k = 0; k < 3; ++k) { F[k] += Fp[k]; N[k] += n[k]; }
    };
#pragma unroll
    for (int m = 0; m < NMARK; ++m) if (P.mbone[m] == b) push(P.mlpos[m], Fm[m]);
    if (P.w_velhs != 0) {
#pragma unroll
      for (int h = 0; h < MAXHS; ++h) {
        const double* hc = c.velcon + h * HS_W;
        if (cfh[h] != 0.0 && (int)hc[1] == b) {
          const double Fp[3] = {cfh[h] * hc[3], cfh[h] * hc[4], cfh[h] * hc[5]};
          push(hc + 6, Fp);
        }
      }
    }
    if (com_on) {
      const double Fp[3] = {Fcom[0] * sk.mass[b], Fcom[1] * sk.mass[b], Fcom[2] * sk.mass[b]};
      push(sk.com[b], Fp);
    }
    cross3(t, F, x);
#pragma unroll
    for (int k = 0; k < 3; ++k) tau[k] = N[k] - x[k];
    for (int ch = sk.nchan[b] - 1; ch >= 0; --ch) {
      double W[3];
      hinge_axis(c, b, ch, W);
      if (valid) VEC(c, V_G, sk.dof0[b] + ch) = dot3(W, tau);
    }
    if (sk.parent[b] >= 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { STK(c, dep - 1, k) += F[k]; STK(c, dep - 1, 3 + k) += N[k]; }
    } else if (valid) {
      VEC(c, V_G, 0) = F[0]; VEC(c, V_G, 1) = F[1]; VEC(c, V_G, 2) = F[2];
    }
  }

  // ---------------- quadratic default terms (:164-265) and the momentum gradient through the cached Jacobian
  {
    double mom_p[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double pv = __shfl_up_sync(0xffffffffu, mom_c[k], 1, GRP);
      mom_p[k] = (f == 0 ? 0.0 : pv) - mom_c[k];           // G[f] += mcache (c_{f-1} - c_f)
    }
    const bool ctr = valid && f >= 1 && f <= nf - 2;       // this frame is the centre i of a ddx / dx term
    const bool ctr_n = valid && f + 1 >= 1 && f + 1 <= nf - 2, ctr_p = valid && f - 1 >= 1 && f - 1 <= nf - 2;
    double oq = 0.0;
#pragma unroll 4
    for (int j = 0; j < nd; ++j) {
      // the frames of a problem sit in neighbouring lanes: neighbours' values by shuffle instead of reloading them
      const double gold = VEC(c, V_G, j);
      const double x0 = valid ? VEC(c, V_X, j) : 0.0;
      const double dd0 = VEC(c, V_DDX, j), d0 = VEC(c, V_DX, j);            // rows f hold ddx[f-1] / dx[f-1] (see the prologue)
      const double sumo = VEC(c, V_SUMO, j);
      const double xm1 = __shfl_up_sync(0xffffffffu, x0, 1, GRP), xm2 = __shfl_up_sync(0xffffffffu, x0, 2, GRP);
      const double xp1 = __shfl_down_sync(0xffffffffu, x0, 1, GRP), xp2 = __shfl_down_sync(0xffffffffu, x0, 2, GRP);
      const double ddn = __shfl_down_sync(0xffffffffu, dd0, 1, GRP), ddp = __shfl_up_sync(0xffffffffu, dd0, 1, GRP);
      const double dn = __shfl_down_sync(0xffffffffu, d0, 1, GRP);
      double sum = x0;                                                       // frames in order 0..7 regardless of the lane
      sum += __shfl_xor_sync(0xffffffffu, sum, 1); sum += __shfl_xor_sync(0xffffffffu, sum, 2); sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      double g = 0.0;
      if (P.w_ddx != 0) {
        const double s = P.s_ddx[j], a = s * INVS, bq = s * (-2.0 * INVS);
        if (ctr) { const double e = a * xm1 + bq * x0 + a * xp1 + s * (-dd0); oq += e * e; g += 2.0 * e * bq; }
        if (ctr_n) { const double e = a * x0 + bq * xp1 + a * xp2 + s * (-ddn); g += 2.0 * e * a; }
        if (ctr_p) { const double e = a * xm2 + bq * xm1 + a * x0 + s * (-ddp); g += 2.0 * e * a; }
      }
      if (P.w_dx != 0) {
        const double s = P.s_dx[j], a = s * INV;
        if (ctr) { const double e = (-a) * xm1 + a * x0 + s * (-d0); oq += e * e; g += 2.0 * e * a; }
        if (ctr_n) { const double e = (-a) * x0 + a * xp1 + s * (-dn); g += 2.0 * e * (-a); }
      }
      if (P.w_ao != 0) {
        const double s = P.s_ao[j];
        const double e = s * sum + s * (-sumo);
        if (f == 0) oq += e * e;
        g += 2.0 * e * s;
      }
      if (P.w_mom > 0) {
        const double* mc = c.MC + (size_t)j * 6 * PPW + c.pg;
#pragma unroll
        for (int k = 0; k < 6; ++k) g += mc[k * PPW] * mom_p[k];
      }
      if (valid) VEC(c, V_G, j) = gold + g;
    }
    if (valid) o += oq;
  }
  __syncwarp();
  return grp_sum(valid ? o : 0.0);
}

// three dot products over the free variables (frames >= 2) of this problem
__device__ __forceinline__ void dots3(const Ctx& c, bool fr, int nd, int a0, int b0, int a1, int b1, int a2, int b2, double* r0, double* r1, double* r2) {
  double s0 = 0, s1 = 0, s2 = 0;
  if (fr)
#pragma unroll 4
    for (int j = 0; j < nd; ++j) {
      s0 += VEC(c, a0, j) * VEC(c, b0, j); s1 += VEC(c, a1, j) * VEC(c, b1, j); s2 += VEC(c, a2, j) * VEC(c, b2, j);
    }
  *r0 = grp_sum(s0); *r1 = grp_sum(s1); *r2 = grp_sum(s2);
}
__device__ __forceinline__ double dot1(const Ctx& c, bool fr, int nd, int a, int b) {
  double s = 0;
  if (fr) {
#pragma unroll 4
    for (int j = 0; j < nd; ++j) s += VEC(c, a, j) * VEC(c, b, j);
  }
  return grp_sum(s);
}

template <bool EVAL_ONLY>
__global__ void __launch_bounds__(32, 4) mmsto_kernel(const Args A) {
  extern __shared__ double smem[];
  const Skel& sk = *A.sk; const Prm& P = *A.prm;
  const int lane = threadIdx.x, f = lane & 7, pg = lane >> 3, nf = P.nf, nd = sk.ndof;
  Ctx c;
  c.sk = &sk; c.P = &P; c.sm = smem; c.gm = smem + ((size_t)sk.nb * SM_W + STK_LEVELS * STK_W) * 32 + pg * GM_W; c.lane = lane; c.f = f; c.pg = pg;
  c.V = A.ws + (size_t)blockIdx.x * NV * MAXD * 32;
  c.MC = A.mc + (size_t)blockIdx.x * MAXD * 6 * PPW;
  const double thr = P.clamp * (3.14159265358979323846 / 180.0);
  int prob = -1;
  bool pvalid = false, valid = false;
  size_t p64 = 0;
  // problem -> registers and workspace for the groups with `loading` set (called by the whole warp: the frame sums are
  // shuffles).  Row f of V_DX holds dx[f-1], row f of V_DDX holds ddx[f-1] (the rows the terms centred at f use).
  auto load_problem = [&](bool loading) {
    if (loading) {
      pvalid = prob >= 0 && prob < A.n; valid = pvalid && f < nf; c.valid = valid;
      p64 = pvalid ? prob : 0;
      c.velcon = A.velcon + p64 * MAXHS * HS_W;
      c.con = A.con ? A.con[p64] : 0;
#pragma unroll
      for (int k = 0; k < 3 * NMARK; ++k) c.tgt[k] = valid ? A.gpos_target[(p64 * nf + f) * 3 * NMARK + k] : 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) c.comt[k] = valid ? A.com[(p64 * nf + f) * 3 + k] : 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) c.momt[k] = (valid && f < nf - 1) ? A.momentum[(p64 * (nf - 1) + f) * 6 + k] : 0.0;
    }
    for (int j = 0; j < MAXD; ++j) {
      double xo = 0, dxv = 0, ddxv = 0, x = 0;
      if (loading && valid && j < nd) {
        xo = A.x_original[(p64 * nf + f) * nd + j];
        if (f >= 1 && f - 1 < nf - 1) dxv = A.dx[(p64 * (nf - 1) + f - 1) * nd + j];
        if (f >= 1 && f - 1 < nf - 2) ddxv = A.ddx[(p64 * (nf - 2) + f - 1) * nd + j];
        if (EVAL_ONLY) x = A.x_in[(p64 * nf + f) * nd + j];
        else if (f < 2) x = fmin(fmax(A.euler_mot[(p64 * nf + f) * nd + j], xo - thr), xo + thr);     // :829-838
        else x = xo;
      }
      const double s = grp_sum(xo);                        // sum over the frames (:231-235); fixed order inside the shuffle tree
      if (loading) {
        if (!EVAL_ONLY) {
#pragma unroll
          for (int q = 0; q < 2 * LB_M; ++q) VEC(c, V_S + q, j) = 0.0;               // stale history of an earlier solve
        }
        VEC(c, V_XO, j) = xo; VEC(c, V_DX, j) = dxv; VEC(c, V_DDX, j) = ddxv; VEC(c, V_X, j) = x; VEC(c, V_G, j) = 0.0; VEC(c, V_D, j) = 0.0;
        VEC(c, V_SUMO, j) = s;
      }
    }
    __syncwarp();
  };

  if (EVAL_ONLY) {
    prob = blockIdx.x * PPW + pg;
    load_problem(true);
    const double fx = evaluate(c, true, true);
    if (valid) {
      for (int j = 0; j < nd; ++j) A.grad[(p64 * nf + f) * nd + j] = VEC(c, V_G, j);
      if (f == 0) A.obj[p64] = fx;
    }
    return;
  }

  // ---- lbfgs() (lbfgs.cpp:246-640) as a per-problem state machine; one evaluation per trip of the loop
  // The warps are persistent: an 8-lane group that has finished its window pulls the next unsolved one from A.queue, so a
  // warp never idles behind the slowest of four windows and a launch has no partly filled last wave.
  bool fr = false;                                         // the free variables (W2 of oracle/mmsto.py)
  int phase = 2, k = 1, end = 0, count = 0, ret = 0, n_eval = 0, iters = 0;
  double step = 0, finit = 0, dginit = 0, dgtest = 0, fx = 0;
  bool firstev = false, need_load = true;
  for (;;) {
    if (__any_sync(0xffffffffu, need_load)) {
      int idx = -1;
      if (need_load && f == 0) idx = atomicAdd(A.queue, 1);
      idx = __shfl_sync(0xffffffffu, idx, lane & ~7);
      if (need_load) prob = idx;
      load_problem(need_load);
      if (need_load) {
        fr = valid && f >= 2;
        phase = pvalid ? 0 : 2; k = 1; end = 0; count = 0; ret = 0; n_eval = 0; iters = 0;
        step = finit = dginit = dgtest = fx = 0;
        firstev = pvalid; need_load = false;
      }
    }
    if (__all_sync(0xffffffffu, phase == 2)) break;        // every group found the queue empty
    const double fv = evaluate(c, __any_sync(0xffffffffu, firstev), firstev);
    firstev = false;
    if (phase != 2) ++n_eval;
    double dg, gg, xx;
    dots3(c, fr, nd, V_G, V_D, V_G, V_G, V_X, V_X, &dg, &gg, &xx);
    bool begin_ls = false, accepted = false;
    if (phase != 2 && !(fv == fv)) {
      // NaN containment (not in libLBFGS, which would keep iterating on NaNs): a window with non-finite inputs stops here with
      // LBFGSERR_UNKNOWNERROR and the last finite point, it does not hold the other windows of its warp until the iteration cap
      if (phase == 1 && fr) for (int j = 0; j < nd; ++j) { VEC(c, V_X, j) = VEC(c, V_XP, j); VEC(c, V_G, j) = VEC(c, V_GP, j); }
      fx = fv; ret = LBFGSERR_UNKNOWNERROR; phase = 2; iters = k;
    }
    if (phase == 0) {
      fx = fv;
      const double xnorm = fmax(sqrt(xx), 1.0), gnorm = sqrt(gg);
      if (gnorm / xnorm <= P.eps) { ret = LBFGS_ALREADY_MINIMIZED; phase = 2; }
      else { step = 1.0 / sqrt(gg); begin_ls = true; }
    } else if (phase == 1) {
      ++count;
      fx = fv;
      double width = 0;
      if (fv > finit + step * dgtest) width = 0.5;
      else if (dg < P.wolfe * dginit) width = 2.1;
      else accepted = true;
      if (!accepted) {
        int err = 0;
        if (step < P.min_step) err = LBFGSERR_MINIMUMSTEP;
        else if (step > P.max_step) err = LBFGSERR_MAXIMUMSTEP;
        else if (P.max_ls <= count) err = LBFGSERR_MAXIMUMLINESEARCH;
        if (err) {                                         // revert to the previous point (:437-443)
          if (fr) for (int j = 0; j < nd; ++j) { VEC(c, V_X, j) = VEC(c, V_XP, j); VEC(c, V_G, j) = VEC(c, V_GP, j); }
          ret = err; phase = 2; iters = k;
        } else {
          step *= width;
          if (fr) for (int j = 0; j < nd; ++j) VEC(c, V_X, j) = VEC(c, V_XP, j) + VEC(c, V_D, j) * step;
        }
      }
    }
    if (phase == 0 && begin_ls) {                          // d = -g
      if (fr) for (int j = 0; j < nd; ++j) VEC(c, V_D, j) = -VEC(c, V_G, j);
    }
    bool upd = false;
    if (accepted) {
      const double xnorm = fmax(sqrt(xx), 1.0), gnorm = sqrt(gg);
      iters = k;
      if (gnorm / xnorm <= P.eps) { ret = LBFGS_SUCCESS; phase = 2; }
      else if ((P.max_iter != 0 && P.max_iter < k + 1) || P.iter_cap < k + 1) { ret = LBFGSERR_MAXIMUMITERATION; phase = 2; }
      else upd = true;
    }
    if (__any_sync(0xffffffffu, upd)) {
      // s = x - xp, y = g - gp and the two-loop recursion (:470-530).  The recursion runs on COEFFICIENTS: the direction is
      // d = cg g + sum_i cs_i s_i + cy_i y_i, and every inner product the recursion needs is an entry of the Gram matrix of
      // {s_i, y_i, g}, of which only the rows of the new pair and of g change per iteration.  That replaces 2 m dependent
      // dot-product / axpy passes over the workspace by m independent passes of inner products and one pass that forms d.
      double *SS = c.gm, *SY = c.gm + 36, *YY = c.gm + 72, *gS = c.gm + 108, *gY = c.gm + 114, *cs = c.gm + 120, *cy = c.gm + 126,
             *al = c.gm + 132, *ysv = c.gm + 138;
      const int e = end;
      if (fr && upd) {
#pragma unroll 4
        for (int j = 0; j < nd; ++j) {
          const double sv = VEC(c, V_X, j) - VEC(c, V_XP, j), yv = VEC(c, V_G, j) - VEC(c, V_GP, j);
          VEC(c, V_S + e, j) = sv; VEC(c, V_Y + e, j) = yv;
        }
      }
      for (int i = 0; i < LB_M; ++i) {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
        if (fr && upd) {
#pragma unroll 8
          for (int j = 0; j < nd; ++j) {
            const double Si = VEC(c, V_S + i, j), Yi = VEC(c, V_Y + i, j), sj = VEC(c, V_S + e, j), yj = VEC(c, V_Y + e, j), gj = VEC(c, V_G, j);
            a0 += sj * Si; a1 += sj * Yi; a2 += yj * Si; a3 += yj * Yi; a4 += gj * Si; a5 += gj * Yi;
          }
        }
        a0 = grp_sum(a0); a1 = grp_sum(a1); a2 = grp_sum(a2); a3 = grp_sum(a3); a4 = grp_sum(a4); a5 = grp_sum(a5);
        if (upd && f == 0) {
          SS[e * 6 + i] = a0; SS[i * 6 + e] = a0; SY[e * 6 + i] = a1; SY[i * 6 + e] = a2; YY[e * 6 + i] = a3; YY[i * 6 + e] = a3;
          gS[i] = a4; gY[i] = a5;
        }
      }
      __syncwarp();
      if (upd) {
        const int bound = LB_M <= k ? LB_M : k;
        ++k;
        end = (end + 1) % LB_M;
        if (f == 0) {                                      // one lane per problem runs the recursion on the coefficients
          const double ys = SY[e * 6 + e], yy = YY[e * 6 + e];
          ysv[e] = ys;
          for (int q = 0; q < LB_M; ++q) { cs[q] = 0; cy[q] = 0; }
          double cg = -1.0;
          int jn = end;
          for (int i = 0; i < bound; ++i) {
            jn = (jn + LB_M - 1) % LB_M;
            double sq = cg * gS[jn];
            for (int q = 0; q < LB_M; ++q) sq += cs[q] * SS[jn * 6 + q] + cy[q] * SY[jn * 6 + q];
            const double a = sq / ysv[jn];
            al[jn] = a; cy[jn] -= a;
          }
          const double sc = ys / yy;
          cg *= sc;
          for (int q = 0; q < LB_M; ++q) { cs[q] *= sc; cy[q] *= sc; }
          for (int i = 0; i < bound; ++i) {
            double yq = cg * gY[jn];
            for (int q = 0; q < LB_M; ++q) yq += cs[q] * SY[q * 6 + jn] + cy[q] * YY[jn * 6 + q];
            cs[jn] += al[jn] - yq / ysv[jn];
            jn = (jn + 1) % LB_M;
          }
          c.gm[GM_CG] = cg;
        }
      }
      __syncwarp();
      if (fr && upd) {
        double rs[LB_M], ry[LB_M];
#pragma unroll
        for (int q = 0; q < LB_M; ++q) { rs[q] = cs[q]; ry[q] = cy[q]; }
        const double cg = c.gm[GM_CG];
#pragma unroll 2
        for (int j = 0; j < nd; ++j) {
          double d = cg * VEC(c, V_G, j);
#pragma unroll
          for (int q = 0; q < LB_M; ++q) d += rs[q] * VEC(c, V_S + q, j) + ry[q] * VEC(c, V_Y + q, j);
          VEC(c, V_D, j) = d;
        }
      }
      if (upd) { step = 1.0; begin_ls = true; }
    }
    if (__any_sync(0xffffffffu, begin_ls)) {
      const double d0 = dot1(c, fr && begin_ls, nd, V_G, V_D);
      if (begin_ls) {
        if (0 < d0) { ret = LBFGSERR_INCREASEGRADIENT; phase = 2; iters = k; }
        else {
          dginit = d0; finit = fx; dgtest = P.ftol * dginit; count = 0; phase = 1;
          if (fr) for (int j = 0; j < nd; ++j) {
            const double xv = VEC(c, V_X, j);
            VEC(c, V_XP, j) = xv; VEC(c, V_GP, j) = VEC(c, V_G, j);
            VEC(c, V_X, j) = xv + VEC(c, V_D, j) * step;
          }
        }
      }
    }
    if (phase == 2 && pvalid) {                            // this window is done: results out, next one at the top of the loop
      if (valid) {
        for (int j = 0; j < nd; ++j) A.x_out[(p64 * nf + f) * nd + j] = VEC(c, V_X, j);
        if (f == 0) {
          A.info[p64 * 4 + 0] = fx; A.info[p64 * 4 + 1] = ret; A.info[p64 * 4 + 2] = iters; A.info[p64 * 4 + 3] = n_eval;
        }
      }
      pvalid = false; valid = false; c.valid = false; fr = false; need_load = true;
    }
    __syncwarp();
  }
}

}  // namespace mmsto

// ---------------------------------------------------------------------------------------------------------------------
// removeCOMsliding_online (gym_trackSRB/taesooLib/DeltaExtractor.py:421-512), SURVEY 8f rank 2: the step after the MMSTO
// solve in the online loop.  8 lanes per window (lane = frame of the nvar <= 8 poses), 16 windows per 128-thread CTA.
// Per lane: CoM of its pose (quaternion-root FK, parent frames on a depth-indexed stack); lanes 0..2 of a window then
// solve the three 1-D hard-constrained smoothing problems (QuadraticFunctionHardCon: KKT system of nvar + 3 unknowns,
// Gaussian elimination with partial pivoting as math.LUsolve); every lane finally moves its root so that the CoM lands
// on the optimised one, leaning about the ground point (plusCOM * R * minusCOM).
namespace mmsto {

struct ComSlideArgs {
  int n, nvar, npose, fix_y, com_only;
  double* poses;                 // [n][nvar][npose] in/out
  const double *com_srb, *desired_com, *ground_h, *kinv;
  double *curr_out, *opt_out;
  const Skel* sk;
};

__device__ __forceinline__ void quat_mul(const double* a, const double* b, double* o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

// inverse of the KKT matrix of removeCOMsliding_online's QuadraticFunctionHardCon (OperatorStitch.cpp:531-620): H = 2 sum c c^T
// over the (-1, 2, -1) sqrt(1.5) stencils, constraints x0, x1, sum_{v>=1} x_v; Gauss-Jordan with partial pivoting, one thread
__global__ void comslide_kkt_inverse_kernel(int nvar, double* kinv) {
  const int N = nvar + 3;
  double K[11][22];
  for (int r = 0; r < N; ++r) for (int c = 0; c < 2 * N; ++c) K[r][c] = c == N + r ? 1.0 : 0.0;
  const double s = sqrt(1.5);
  for (int v = 1; v < nvar - 1; ++v) {
    const double cf[3] = {s * -1.0, s * 2.0, s * -1.0};
    for (int a = 0; a < 3; ++a) {
      K[v - 1 + a][v - 1 + a] += 2.0 * cf[a] * cf[a];
      for (int b = a + 1; b < 3; ++b) { K[v - 1 + a][v - 1 + b] += 2.0 * cf[a] * cf[b]; K[v - 1 + b][v - 1 + a] = K[v - 1 + a][v - 1 + b]; }
    }
  }
  K[nvar][0] = 1.0; K[0][nvar] = 1.0; K[nvar + 1][1] = 1.0; K[1][nvar + 1] = 1.0;
  for (int v = 1; v < nvar; ++v) { K[nvar + 2][v] = 1.0; K[v][nvar + 2] = 1.0; }
  for (int col = 0; col < N; ++col) {
    int piv = col; double best = fabs(K[col][col]);
    for (int r = col + 1; r < N; ++r) if (fabs(K[r][col]) > best) { best = fabs(K[r][col]); piv = r; }
    if (piv != col) for (int c = 0; c < 2 * N; ++c) { const double t = K[col][c]; K[col][c] = K[piv][c]; K[piv][c] = t; }
    const double inv = 1.0 / K[col][col];
    for (int c = 0; c < 2 * N; ++c) K[col][c] *= inv;
    for (int r = 0; r < N; ++r) {
      if (r == col) continue;
      const double m = K[r][col];
      if (m != 0.0) for (int c = 0; c < 2 * N; ++c) K[r][c] -= m * K[col][c];
    }
  }
  for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) kinv[r * N + c] = K[r][N + c];
}

__global__ void __launch_bounds__(128) comslide_kernel(const ComSlideArgs A) {
  __shared__ double xs[16][3][NFMAX];
  const Skel& sk = *A.sk;
  const int lane = threadIdx.x & 31, f = lane & 7, wl = threadIdx.x >> 3;           // wl: window within the CTA
  const int win = blockIdx.x * 16 + wl, nvar = A.nvar;
  const bool pvalid = win < A.n, valid = pvalid && f < nvar;
  const size_t w64 = pvalid ? win : 0;
  double* pose = A.poses + (w64 * nvar + (valid ? f : 0)) * A.npose;

  // ---- CoM of this pose (BoneForwardKinematics::setPoseDOF + calcCOM; root quaternion normalised as quater::normalize)
  double sq[STK_LEVELS][4], st[STK_LEVELS][3];
  double com[3] = {0, 0, 0}, rootq[4], roott[3];
  {
    double q0[4] = {pose[3], pose[4], pose[5], pose[6]};
#pragma unroll
    for (int k = 0; k < 4; ++k) rootq[k] = q0[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) roott[k] = pose[k];
    int ja = 7;
    for (int b = 0; b < sk.nb; ++b) {
      const int dep = sk.depth[b];
      double q[4], t[3], R[9];
      if (b == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = q0[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = pose[k] + sk.off[0][k];
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = sq[dep - 1][k];
        quat_to_mat(q, R);
        double d[3];
        mat_vec(R, sk.off[b], d);
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = st[dep - 1][k] + d[k];
        for (int ch = 0; ch < sk.nchan[b]; ++ch) {
          double sn, cs;
          sincos(0.5 * pose[ja + ch], &sn, &cs);
          quat_rot_axis(q, sk.axis[b][ch], sn, cs);
        }
        ja += sk.nchan[b];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) sq[dep][k] = q[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) st[dep][k] = t[k];
      quat_to_mat(q, R);
      double cb[3];
      mat_vec(R, sk.com[b], cb);
#pragma unroll
      for (int k = 0; k < 3; ++k) com[k] += (cb[k] + t[k]) * sk.mass[b];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) com[k] *= 1.0 / sk.total_mass;
  }
  if (valid && A.curr_out) {
#pragma unroll
    for (int k = 0; k < 3; ++k) A.curr_out[(w64 * nvar + f) * 3 + k] = com[k];
  }
  if (A.com_only) return;

  // ---- three hard-constrained 1-D problems per window: lane `dim` of the window solves dimension dim
  double c0[3], c1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { c0[k] = __shfl_sync(0xffffffffu, com[k], (lane & ~7) + 0); c1[k] = __shfl_sync(0xffffffffu, com[k], (lane & ~7) + 1); }
  if (pvalid && f < 3) {
    // The KKT matrix depends on nvar only: its inverse is built once (comslide_kkt_inverse_kernel) and each problem is the
    // product of its first nvar rows with the right-hand side [-2 sum v c | curr0, curr1, sum desired].
    const int dim = f, N = nvar + 3;
    double d[11];
    for (int r = 0; r < N; ++r) d[r] = 0.0;
    const double* cs = A.com_srb + w64 * nvar * 3;
    const double s = sqrt(1.5);
    for (int v = 1; v < nvar - 1; ++v) {
      const double cst = s * (-1 * (-1 * cs[(v - 1) * 3 + dim] + 2 * cs[v * 3 + dim] - cs[(v + 1) * 3 + dim]));
      d[v - 1] -= 2.0 * cst * (s * -1.0); d[v] -= 2.0 * cst * (s * 2.0); d[v + 1] -= 2.0 * cst * (s * -1.0);
    }
    d[nvar] = dim == 0 ? c0[0] : dim == 1 ? c0[1] : c0[2];
    d[nvar + 1] = dim == 0 ? c1[0] : dim == 1 ? c1[1] : c1[2];
    double sum = 0.0;
    for (int v = 0; v < nvar - 1; ++v) sum += A.desired_com[(w64 * nvar + v) * 3 + dim];
    d[nvar + 2] = sum;
    for (int v = 0; v < nvar; ++v) {
      double acc = 0.0;
      for (int cc = 0; cc < N; ++cc) acc += A.kinv[v * N + cc] * d[cc];
      xs[wl][dim][v] = acc;
    }
  }
  __syncthreads();
  if (!valid) return;
  double opt[3] = {xs[wl][0][f], xs[wl][1][f], xs[wl][2][f]};
  if (!A.fix_y) opt[1] = com[1];
  if (A.opt_out) {
#pragma unroll
    for (int k = 0; k < 3; ++k) A.opt_out[(w64 * nvar + f) * 3 + k] = opt[k];
  }
  // ---- root <- T(opt) * R * T(-curr) * root, R = rotation by -|delta| * height * 0.5 about delta x up
  const double dx = opt[0] - com[0], dz = opt[2] - com[2];
  double axis[3] = {-dz, 0.0, dx};                          // (dx, 0, dz) x (0, 1, 0)
  const double ln = sqrt(axis[0] * axis[0] + axis[2] * axis[2]);
  const double il = ln == 0 ? 0.0 : 1.0 / ln;
  axis[0] *= il; axis[2] *= il;
  double cy = com[1];
  if (A.ground_h) cy -= A.ground_h[w64 * nvar + f];
  const double ang = -sqrt(dx * dx + dz * dz) * cy * 0.5;
  double sn, cs;
  sincos(0.5 * ang, &sn, &cs);
  const double Rq[4] = {cs, sn * axis[0], sn * axis[1], sn * axis[2]};
  double nq[4], R[9], d[3], nt[3];
  quat_mul(Rq, rootq, nq);
  quat_to_mat(Rq, R);
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = roott[k] - com[k];
  mat_vec(R, d, nt);
#pragma unroll
  for (int k = 0; k < 3; ++k) pose[k] = nt[k] + opt[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) pose[3 + k] = nq[k];
}

}  // namespace mmsto

// ---------------------------------------------------------------------------------------------------------------------
// _extractSingleFrameFullbody (gym_trackSRB/taesooLib/DeltaExtractor.py:577-643), SURVEY 8f rank 2: decoded full-body row +
// SRB state of its frame -> the inputs of the MMSTO solve.  One warp per item: the row is staged in shared memory with
// coalesced loads, the copy-and-scale parts (joint angles, dx, ddx, contacts) are written by all lanes, the handful of
// small transforms (pelvis exp(se3), one-body momentum, foot markers, root velocity) by one lane each.
namespace mmsto {

enum { FB_TO_PELVIS = 6, FB_TO_FEET = 24, FB_TO_MOM = 9, FB_POSES = 53, FB_NDOF = 59, FB_ROW = 214, FB_EXTRA = 29,
       FB_C_FEET = 6, FB_C_MOM = 30, FB_C_POSES = 39, FB_C_DX = 92, FB_C_DDX = 151, FB_C_CON = 210 };

struct FullbodyArgs {
  int n;
  const double *extra, *row;           // [n][29], [n][214]
  double mass, inertia[6], r[3];       // SRB segment: mass, Ixx Iyy Izz Ixy Ixz Iyz, mass * local CoM
  double lcom[3];
  double *human, *dx_ddx, *feetpos, *com, *momentum, *con_bin, *con;
};

// vector3::rotate(quater, v): (2 u.v) u + (s^2 - u.u) v + 2 s u x v   (no normalisation, as the reference)
__device__ __forceinline__ void q_rotate(const double* q, const double* v, double* o) {
  const double s = q[0], ux = q[1], uy = q[2], uz = q[3];
  const double a = 2.0 * (ux * v[0] + uy * v[1] + uz * v[2]), b = s * s - (ux * ux + uy * uy + uz * uz), c2 = 2.0 * s;
  o[0] = a * ux + b * v[0] + c2 * (uy * v[2] - uz * v[1]);
  o[1] = a * uy + b * v[1] + c2 * (uz * v[0] - ux * v[2]);
  o[2] = a * uz + b * v[2] + c2 * (ux * v[1] - uy * v[0]);
}

struct FullbodyOut { double *human, *dx_ddx, *feetpos, *com, *momentum, *con_bin, *con; };

// task `task` of item `item` of the output arrays O (S: its staged row, E: its SRB record)
__device__ __forceinline__ void fullbody_task(const FullbodyArgs& A, const FullbodyOut& O, int task, size_t item, const double* S, const double* E) {
  // E: pose_tl 0..6 | dq 7..12 | foot0 13..19 | foot1 20..26 | con 27..28
  const double* rq = E + 3;
  double* human = O.human + item * (7 + FB_POSES);
  double* dxo = O.dx_ddx + item * 2 * FB_NDOF;
  if (task == 0) {                                    // humanPose root = SRBroot * exp(toPelvis)
    const double w0 = S[0], w1 = S[1], w2 = S[2];
    const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    double q2[4], p[3];
    if (fabs(theta) < 1e-9) { q2[0] = 1.0; q2[1] = q2[2] = q2[3] = 0.0; p[0] = S[3]; p[1] = S[4]; p[2] = S[5];
      // toBaseR of the identity: tr = 3 > 0 -> w = sqrt(4) / 2 = 1
    } else {
      const double it = 1.0 / theta;
      const double s[6] = {it * w0, it * w1, it * w2, it * S[3], it * S[4], it * S[5]};
      double st, ct;
      sincos(theta, &st, &ct);
      const double vt = 1.0 - ct, ut = (theta - st) * (s[0] * s[3] + s[1] * s[4] + s[2] * s[5]), t0 = s[2] * st, t1 = s[1] * st, t2 = s[0] * st;
      double m[3][3];
      m[0][0] = s[0] * s[0] * vt + ct; m[1][0] = s[0] * s[1] * vt + t0; m[2][0] = s[0] * s[2] * vt - t1;
      m[0][1] = s[0] * s[1] * vt - t0; m[1][1] = s[1] * s[1] * vt + ct; m[2][1] = s[1] * s[2] * vt + t2;
      m[0][2] = s[0] * s[2] * vt + t1; m[1][2] = s[1] * s[2] * vt - t2; m[2][2] = s[2] * s[2] * vt + ct;
      p[0] = st * s[3] + ut * s[0] + vt * (s[1] * s[5] - s[2] * s[4]);
      p[1] = st * s[4] + ut * s[1] + vt * (s[2] * s[3] - s[0] * s[5]);
      p[2] = st * s[5] + ut * s[2] + vt * (s[0] * s[4] - s[1] * s[3]);
      const double tr = m[0][0] + m[1][1] + m[2][2];   // toBaseR
      if (tr > 0.0) {
        double sq = sqrt(tr + 1.0);
        q2[0] = sq * 0.5; sq = 0.5 / sq;
        q2[1] = (m[2][1] - m[1][2]) * sq; q2[2] = (m[0][2] - m[2][0]) * sq; q2[3] = (m[1][0] - m[0][1]) * sq;
      } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double sq = sqrt((m[i][i] - (m[j][j] + m[k][k])) + 1.0);
        q2[i + 1] = sq * 0.5; sq = 0.5 / sq;
        q2[0] = (m[k][j] - m[j][k]) * sq; q2[j + 1] = (m[j][i] + m[i][j]) * sq; q2[k + 1] = (m[k][i] + m[i][k]) * sq;
      }
    }
    double hq[4], ht[3];
    quat_mul(rq, q2, hq);
    q_rotate(rq, p, ht);
#pragma unroll
    for (int k = 0; k < 3; ++k) human[k] = ht[k] + E[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) human[3 + k] = hq[k];
  } else if (task == 1) {                             // SRB momentum about its CoM + decoded corrections; CoM target
    double inv[4], wb[3], vb[3], M[3], F[3], tmp[3];
    const double n2 = rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3], il = 1.0 / sqrt(n2);
    inv[0] = rq[0] * il; inv[1] = rq[1] * -1.0 * il; inv[2] = rq[2] * -1.0 * il; inv[3] = rq[3] * -1.0 * il;   // quater::inverse
    q_rotate(inv, E + 7, wb); q_rotate(inv, E + 10, vb);
    const double* I = A.inertia;
    cross3(A.r, vb, tmp);
    M[0] = I[0] * wb[0] + I[3] * wb[1] + I[4] * wb[2] + tmp[0];
    M[1] = I[3] * wb[0] + I[1] * wb[1] + I[5] * wb[2] + tmp[1];
    M[2] = I[4] * wb[0] + I[5] * wb[1] + I[2] * wb[2] + tmp[2];
    cross3(wb, A.r, tmp);
#pragma unroll
    for (int k = 0; k < 3; ++k) F[k] = tmp[k] + A.mass * vb[k];
    // inv_dAd(T) then dAd(T(com)): world frame, about the CoM = R lcom + t
    double inv2[4], tinv[3], rt[3], Fw[3], Mw[3], comw[3], mt[3];
    q_rotate(inv, E, rt);
#pragma unroll
    for (int k = 0; k < 3; ++k) tinv[k] = -rt[k];
    cross3(tinv, F, tmp);
#pragma unroll
    for (int k = 0; k < 3; ++k) mt[k] = M[k] - tmp[k];
    const double n3 = inv[0] * inv[0] + inv[1] * inv[1] + inv[2] * inv[2] + inv[3] * inv[3], il2 = 1.0 / sqrt(n3);
    inv2[0] = inv[0] * il2; inv2[1] = inv[1] * -1.0 * il2; inv2[2] = inv[2] * -1.0 * il2; inv2[3] = inv[3] * -1.0 * il2;
    q_rotate(inv2, mt, Mw); q_rotate(inv2, F, Fw);
    q_rotate(rq, A.lcom, comw);
#pragma unroll
    for (int k = 0; k < 3; ++k) comw[k] += E[k];
    cross3(comw, Fw, tmp);
    double eM[3], eF[3], eC[3];
    q_rotate(rq, S + FB_C_MOM, eC); q_rotate(rq, S + FB_C_MOM + 3, eM); q_rotate(rq, S + FB_C_MOM + 6, eF);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      O.momentum[item * 6 + k] = (Mw[k] - tmp[k]) + eM[k];
      O.momentum[item * 6 + 3 + k] = Fw[k] + eF[k];
      O.com[item * 3 + k] = E[k] + eC[k];
    }
  } else if (task < 6) {                              // foot marker positions in the foot frames
    const int ic = task - 2;
    const double* ft = E + 13 + 7 * (ic >> 1);
    double g[3];
    q_rotate(ft + 3, S + FB_C_FEET + ic * 3, g);
#pragma unroll
    for (int k = 0; k < 3; ++k) O.feetpos[item * 24 + ic * 3 + k] = g[k] + ft[k];
  } else if (task < 10) {                             // foot marker velocities
    const int ic = task - 6;
    double g[3];
    q_rotate(rq, S + FB_C_FEET + 12 + ic * 3, g);
#pragma unroll
    for (int k = 0; k < 3; ++k) O.feetpos[item * 24 + 12 + ic * 3 + k] = g[k] * 10.0 + E[10 + k];
  } else {                                            // root translational velocity / acceleration: set_dx_ddx_v3
    const bool vel = task == 10;
    const int c0 = vel ? FB_C_DX : FB_C_DDX;
    const double v[3] = {S[c0] * 10.0, S[c0 + 1] * 10.0, S[c0 + 2] * 10.0};
    double g[3];
    q_rotate(rq, v, g);
#pragma unroll
    for (int k = 0; k < 3; ++k) dxo[(vel ? 0 : FB_NDOF) + k] = g[k] + (vel ? E[10 + k] : 0.0);
  }
}

// FB_ITEMS = 8 items per 128-thread CTA (29 kB of shared memory, 7 CTAs per SM; 16 items: fewer CTAs and 23 % slower, 4: too few
// bytes per bulk copy).  (1) the 8 decoder rows (13.7 kB) and SRB records (1.9 kB) are contiguous in global memory and
// arrive by two TMA bulk copies on an mbarrier; (2) the 12 small transforms of every item run as (task, item) pairs, 8
// consecutive threads on the same task (a warp diverges four ways at most), and the copy-and-scale columns (joint angles, dx,
// ddx, contacts) are produced by all threads -- everything into shared-memory output tiles; (3) the seven output tiles leave
// by bulk stores.  The last, partly filled CTA of a launch uses plain loads / stores.
enum { FB_ITEMS = 8, FB_TASKS = 12 };
struct FullbodyTile {
  double row[FB_ITEMS][FB_ROW];
  double ext[FB_ITEMS][FB_EXTRA];
  double human[FB_ITEMS][7 + FB_POSES];
  double dx_ddx[FB_ITEMS][2 * FB_NDOF];
  double feetpos[FB_ITEMS][24];
  double com[FB_ITEMS][3];
  double momentum[FB_ITEMS][6];
  double con_bin[FB_ITEMS][4];
  double con[FB_ITEMS][4];
};
static_assert(sizeof(double) * FB_ITEMS * FB_ROW % 16 == 0 && sizeof(double) * FB_ITEMS * FB_EXTRA % 16 == 0 &&
              sizeof(double) * FB_ITEMS * 3 % 16 == 0 && offsetof(FullbodyTile, ext) % 16 == 0 && offsetof(FullbodyTile, human) % 16 == 0 &&
              offsetof(FullbodyTile, dx_ddx) % 16 == 0 && offsetof(FullbodyTile, feetpos) % 16 == 0 && offsetof(FullbodyTile, com) % 16 == 0 &&
              offsetof(FullbodyTile, momentum) % 16 == 0 && offsetof(FullbodyTile, con_bin) % 16 == 0 && offsetof(FullbodyTile, con) % 16 == 0,
              "bulk copies need 16-byte multiples");

__device__ __forceinline__ unsigned fb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) fullbody_extract_kernel(const FullbodyArgs A) {
  extern __shared__ __align__(128) unsigned char fb_raw[];
  FullbodyTile& T = *reinterpret_cast<FullbodyTile*>(fb_raw);
  __shared__ __align__(8) unsigned long long mbar;
  const size_t item0 = (size_t)blockIdx.x * FB_ITEMS;
  const int cnt = (int)min((size_t)FB_ITEMS, (size_t)A.n - item0);
  const bool full = cnt == FB_ITEMS;
  const int tid = threadIdx.x;
  if (full) {
    constexpr unsigned row_bytes = sizeof(double) * FB_ITEMS * FB_ROW, ext_bytes = sizeof(double) * FB_ITEMS * FB_EXTRA;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fb_smem_u32(&mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb_smem_u32(&mbar)), "r"(row_bytes + ext_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(fb_smem_u32(T.row)), "l"(A.row + item0 * FB_ROW), "r"(row_bytes), "r"(fb_smem_u32(&mbar)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(fb_smem_u32(T.ext)), "l"(A.extra + item0 * FB_EXTRA), "r"(ext_bytes), "r"(fb_smem_u32(&mbar)) : "memory");
    }
    __syncthreads();
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0, 1, 0, P; }" : "=r"(ok) : "r"(fb_smem_u32(&mbar)) : "memory");
  } else {
    for (int k = tid; k < cnt * FB_ROW; k += 128) T.row[k / FB_ROW][k % FB_ROW] = A.row[item0 * FB_ROW + k];
    for (int k = tid; k < cnt * FB_EXTRA; k += 128) T.ext[k / FB_EXTRA][k % FB_EXTRA] = A.extra[item0 * FB_EXTRA + k];
    __syncthreads();
  }
  FullbodyOut O;
  size_t base;
  if (full) { O = {&T.human[0][0], &T.dx_ddx[0][0], &T.feetpos[0][0], &T.com[0][0], &T.momentum[0][0], &T.con_bin[0][0], &T.con[0][0]}; base = 0; }
  else { O = {A.human, A.dx_ddx, A.feetpos, A.com, A.momentum, A.con_bin, A.con}; base = item0; }
  for (int p = tid; p < FB_TASKS * FB_ITEMS; p += 128) {
    const int task = p / FB_ITEMS, it = p % FB_ITEMS;
    if (it < cnt) fullbody_task(A, O, task, base + it, T.row[it], T.ext[it]);
  }
  for (int k = tid; k < cnt * FB_POSES; k += 128) {
    const int it = k / FB_POSES, c = k % FB_POSES;
    O.human[(base + it) * (7 + FB_POSES) + 7 + c] = T.row[it][FB_C_POSES + c];
  }
  for (int k = tid; k < cnt * 2 * FB_NDOF; k += 128) {
    const int it = k / (2 * FB_NDOF), c = k % (2 * FB_NDOF);
    const int col = c < FB_NDOF ? c : c - FB_NDOF;
    if (col >= 3) O.dx_ddx[(base + it) * 2 * FB_NDOF + c] = T.row[it][(c < FB_NDOF ? FB_C_DX : FB_C_DDX) + col] * 10.0;
  }
  for (int k = tid; k < cnt * 4; k += 128) {
    const double cv = T.row[k / 4][FB_C_CON + k % 4];
    O.con[base * 4 + k] = cv; O.con_bin[base * 4 + k] = cv > 0 ? 1.0 : 0.0;
  }
  if (full) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
#define FB_STORE(dst, src, bytes) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(fb_smem_u32(src)), "r"((unsigned)(bytes)) : "memory")
      FB_STORE(A.human + item0 * (7 + FB_POSES), T.human, sizeof(T.human));
      FB_STORE(A.dx_ddx + item0 * 2 * FB_NDOF, T.dx_ddx, sizeof(T.dx_ddx));
      FB_STORE(A.feetpos + item0 * 24, T.feetpos, sizeof(T.feetpos));
      FB_STORE(A.com + item0 * 3, T.com, sizeof(T.com));
      FB_STORE(A.momentum + item0 * 6, T.momentum, sizeof(T.momentum));
      FB_STORE(A.con_bin + item0 * 4, T.con_bin, sizeof(T.con_bin));
      FB_STORE(A.con + item0 * 4, T.con, sizeof(T.con));
#undef FB_STORE
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}

}  // namespace mmsto
