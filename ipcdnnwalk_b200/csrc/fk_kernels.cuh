// Batched full-body forward kinematics + CoM + centroidal momentum.
//
// Reference semantics (taesooLib): BoneForwardKinematics::setPoseDOF/forwardKinematics
// (BaseLib/motion/BoneKinematics.cpp:18-47,97-136), Euler channels composed by right multiplication
// (math/quater.cpp:428-470), LoaderToTree velocity propagation (IK_sdls/Node.cpp:68-99) and
// calcCOM / calcMomentumCOM (IK_sdls/NodeWrap.cpp:325-386) with the Liegroup spatial inertia
// Inertia(mass, momentsOfInertia, mass*localCOM) (Liegroup.h:141-147, no parallel-axis term: Q13).
//
// Mapping: one thread per pose.  Bones are stored depth-first, so the parent of a bone is always the entry one
// level up on the current root-to-bone path: each thread keeps a stack indexed by tree depth (<= MAXD levels x 13
// reals: global quaternion 4, translation 3, body angular 3 and linear 3 velocity) in shared memory laid out
// [slot][thread] (bank-conflict free), i.e. a level-by-level sweep whose live state is one path, not the whole
// skeleton.  Pose rows (60 reals) are read and bone transforms (7 reals) written directly by the owning thread:
// consecutive accesses of a thread hit the same 32-byte sectors, so L1 absorbs the row stride and DRAM traffic stays
// at the algorithmic 836 B (fp32) / 1672 B (fp64) per pose.
#pragma once
#include "srb_math.cuh"

namespace fk {
using namespace srb;

enum { MAXB = 64, MAXD = 16 };

struct SkelConst {
  int nb, n_pose, n_dq, maxdepth;
  int depth[MAXB];
  int nchan[MAXB];
  int axes[MAXB];          // 2 bits per channel: 0 X, 1 Y, 2 Z
  double offset[MAXB][3];
  double mass[MAXB];
  double com[MAXB][3];
  double inertia[MAXB][6]; // Ixx Iyy Izz Ixy Ixz Iyz
  double inv_total_mass;
};

template <typename T> __device__ __forceinline__ V3<T> qrot(Q4<T> q, V3<T> v) {
  // vector3::rotate(quater, v) (math/vector3.cpp:344-373)
  V3<T> u = {q.x, q.y, q.z};
  T s = q.w;
  T a = T(2) * dot(u, v), b = s * s - dot(u, u), c = T(2) * s;
  V3<T> x = cross(u, v);
  return {a * u.x + b * v.x + c * x.x, a * u.y + b * v.y + c * x.y, a * u.z + b * v.z + c * x.z};
}
template <typename T> __device__ __forceinline__ Q4<T> qconj_unit(Q4<T> q) {
  // quater::inverse: conjugate then normalize()
  T inv = num<T>::rsqrt_pos(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w * inv, -q.x * inv, -q.y * inv, -q.z * inv};
}

template <typename real, int BLOCK>
__global__ void __launch_bounds__(BLOCK) fk_forward_kernel(const SkelConst* __restrict__ skp, int n, const real* __restrict__ pose,
                                                           const real* __restrict__ dq, real* __restrict__ xform,
                                                           real* __restrict__ com_out, real* __restrict__ mom_out) {
  extern __shared__ __align__(16) unsigned char fk_smem[];
  __shared__ SkelConst sk;
  {
    const int* src = reinterpret_cast<const int*>(skp);
    int* dst = reinterpret_cast<int*>(&sk);
    for (int i = threadIdx.x; i < (int)(sizeof(SkelConst) / 4); i += BLOCK) dst[i] = src[i];
  }
  __syncthreads();
  real* stk = reinterpret_cast<real*>(fk_smem);           // [MAXD-slot][13][BLOCK]
  const int tid = threadIdx.x;
  const int i = blockIdx.x * BLOCK + tid;
  if (i >= n) return;
#define STK(d, k) stk[((d) * 13 + (k)) * BLOCK + tid]
  const real* p = pose + (size_t)i * sk.n_pose;
  const real* v = dq ? dq + (size_t)i * sk.n_dq : nullptr;
  real* xo = xform ? xform + (size_t)i * sk.nb * 7 : nullptr;
  const bool vel = (v != nullptr);

  V3<real> csum = {0, 0, 0};
  real HM[3] = {0, 0, 0}, HF[3] = {0, 0, 0};
  int start = 7;
  // software pipeline: the angles / rates of the NEXT bone are loaded while the current bone is computed
  // (a thread walks its own row, so without this every bone would expose a DRAM round trip)
  real ang_n[3] = {0, 0, 0}, rate_n[3] = {0, 0, 0};
  const real rp[7] = {p[0], p[1], p[2], p[3], p[4], p[5], p[6]};
  real rv[6] = {0, 0, 0, 0, 0, 0};
  if (vel) {
#pragma unroll
    for (int k = 0; k < 6; k++) rv[k] = v[k];
  }
  if (sk.nb > 1) {
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (k < sk.nchan[1]) { ang_n[k] = p[7 + k]; if (vel) rate_n[k] = v[6 + k]; }
  }
  // transforms are written 4 bones (28 reals) at a time with 128-bit stores: a thread owns its whole output row, so
  // one warp-level store touches 32 rows either way -- wide stores cut the L1 sector transactions 4x (fp32) / 2x (fp64)
  const bool vec_ok = xo != nullptr && ((sk.nb * 7 * (int)sizeof(real)) % 16 == 0);   // row base 16-byte aligned
  constexpr int GB = sizeof(real) == 4 ? 4 : 2;          // bones per store group: 28 floats = 7 float4, 14 doubles = 7 double2
  for (int b0 = 0; b0 < sk.nb; b0 += GB) {
  real outv[7 * GB];
#pragma unroll
  for (int jb = 0; jb < GB; jb++) {
    const int b = b0 + jb;
    if (b >= sk.nb) break;
    const int d = sk.depth[b];
    real ang[3] = {ang_n[0], ang_n[1], ang_n[2]}, rate[3] = {rate_n[0], rate_n[1], rate_n[2]};
    const int nc = sk.nchan[b];
    {
      const int nb1 = b + 1;
      if (nb1 < sk.nb && b >= 1) {
        const int s1 = start + nc;
#pragma unroll
        for (int k = 0; k < 3; k++)
          if (k < sk.nchan[nb1]) { ang_n[k] = p[s1 + k]; if (vel) rate_n[k] = v[6 + (s1 - 7) + k]; }
      }
    }
    Q4<real> q; V3<real> t, w = {0, 0, 0}, lv = {0, 0, 0};
    const V3<real> off = {(real)sk.offset[b][0], (real)sk.offset[b][1], (real)sk.offset[b][2]};
    if (b == 0) {
      q = {rp[3], rp[4], rp[5], rp[6]};
      t = {rp[0] + off.x, rp[1] + off.y, rp[2] + off.z};
      if (vel) {
        Q4<real> qi = qconj_unit(q);
        w = qrot(qi, V3<real>{rv[0], rv[1], rv[2]});
        lv = qrot(qi, V3<real>{rv[3], rv[4], rv[5]});
      }
    } else {
      const Q4<real> pq = {STK(d - 1, 0), STK(d - 1, 1), STK(d - 1, 2), STK(d - 1, 3)};
      const V3<real> pt = {STK(d - 1, 4), STK(d - 1, 5), STK(d - 1, 6)};
      // local rotation: right-multiplied single-axis quaternions; velocity propagated one hinge at a time
      Q4<real> lq = {real(1), real(0), real(0), real(0)};
      if (vel) {
        V3<real> pw = {STK(d - 1, 7), STK(d - 1, 8), STK(d - 1, 9)};
        V3<real> pv = {STK(d - 1, 10), STK(d - 1, 11), STK(d - 1, 12)};
        w = pw;
        lv = pv + cross(pw, off);                 // (v_p + w_p x offset), rotated below by each hinge
      }
      const int ax = sk.axes[b];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (k < nc) {
          const int a = (ax >> (2 * k)) & 3;
          real s, c;
          num<real>::sincos(real(0.5) * ang[k], &s, &c);
          // lq * (c, s e_a): single-axis right multiplication
          Q4<real> r;
          if (a == 0) r = {lq.w * c - lq.x * s, lq.x * c + lq.w * s, lq.y * c + lq.z * s, lq.z * c - lq.y * s};
          else if (a == 1) r = {lq.w * c - lq.y * s, lq.x * c - lq.z * s, lq.y * c + lq.w * s, lq.z * c + lq.x * s};
          else r = {lq.w * c - lq.z * s, lq.x * c + lq.y * s, lq.y * c - lq.x * s, lq.z * c + lq.w * s};
          lq = r;
          if (vel) {
            // rotate by the inverse hinge rotation: planar rotation by -theta about axis a (cos, sin of the full angle)
            const real C = c * c - s * s, S = real(2) * s * c;
            if (a == 0) { w = {w.x, C * w.y + S * w.z, C * w.z - S * w.y}; lv = {lv.x, C * lv.y + S * lv.z, C * lv.z - S * lv.y}; w.x += rate[k]; }
            else if (a == 1) { w = {C * w.x - S * w.z, w.y, C * w.z + S * w.x}; lv = {C * lv.x - S * lv.z, lv.y, C * lv.z + S * lv.x}; w.y += rate[k]; }
            else { w = {C * w.x + S * w.y, C * w.y - S * w.x, w.z}; lv = {C * lv.x + S * lv.y, C * lv.y - S * lv.x, lv.z}; w.z += rate[k]; }
          }
        }
      }
      start += nc;
      q = quat_mul(pq, lq);
      t = qrot(pq, off) + pt;
    }
    STK(d, 0) = q.w; STK(d, 1) = q.x; STK(d, 2) = q.y; STK(d, 3) = q.z;
    STK(d, 4) = t.x; STK(d, 5) = t.y; STK(d, 6) = t.z;
    if (vel) { STK(d, 7) = w.x; STK(d, 8) = w.y; STK(d, 9) = w.z; STK(d, 10) = lv.x; STK(d, 11) = lv.y; STK(d, 12) = lv.z; }
    outv[7 * jb] = q.w; outv[7 * jb + 1] = q.x; outv[7 * jb + 2] = q.y; outv[7 * jb + 3] = q.z;
    outv[7 * jb + 4] = t.x; outv[7 * jb + 5] = t.y; outv[7 * jb + 6] = t.z;
    // CoM and momentum contributions
    const real m = (real)sk.mass[b];
    const V3<real> c = {(real)sk.com[b][0], (real)sk.com[b][1], (real)sk.com[b][2]};
    V3<real> gc = qrot(q, c) + t;
    csum = csum + m * gc;
    if (vel && mom_out) {
      const real Ixx = (real)sk.inertia[b][0], Iyy = (real)sk.inertia[b][1], Izz = (real)sk.inertia[b][2];
      const real Ixy = (real)sk.inertia[b][3], Ixz = (real)sk.inertia[b][4], Iyz = (real)sk.inertia[b][5];
      const V3<real> r = m * c;
      V3<real> M = {Ixx * w.x + Ixy * w.y + Ixz * w.z + r.y * lv.z - r.z * lv.y,
                    Ixy * w.x + Iyy * w.y + Iyz * w.z + r.z * lv.x - r.x * lv.z,
                    Ixz * w.x + Iyz * w.y + Izz * w.z + r.x * lv.y - r.y * lv.x};
      V3<real> F = {w.y * r.z - w.z * r.y + m * lv.x, w.z * r.x - w.x * r.z + m * lv.y, w.x * r.y - w.y * r.x + m * lv.z};
      // inv_dAd(T): world wrench about the world origin  (R M + t x R F, R F)
      V3<real> Fw = qrot(q, F);
      V3<real> Mw = qrot(q, M) + cross(t, Fw);
      HM[0] += Mw.x; HM[1] += Mw.y; HM[2] += Mw.z; HF[0] += Fw.x; HF[1] += Fw.y; HF[2] += Fw.z;
    }
  }
  if (xo) {
    real* o = xo + b0 * 7;
    if (vec_ok && b0 + GB <= sk.nb) {
      if constexpr (sizeof(real) == 4) {
#pragma unroll
        for (int k = 0; k < 7; k++) reinterpret_cast<float4*>(o)[k] = make_float4(outv[4 * k], outv[4 * k + 1], outv[4 * k + 2], outv[4 * k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < 7; k++) reinterpret_cast<double2*>(o)[k] = make_double2(outv[2 * k], outv[2 * k + 1]);
      }
    } else {
      const int cnt = min(GB, sk.nb - b0) * 7;
#pragma unroll
      for (int k = 0; k < 7 * GB; k++) if (k < cnt) o[k] = outv[k];
    }
  }
  }
  const real itm = (real)sk.inv_total_mass;
  V3<real> com = itm * csum;
  if (com_out) { real* o = com_out + (size_t)i * 3; o[0] = com.x; o[1] = com.y; o[2] = com.z; }
  if (vel && mom_out) {
    V3<real> Fv = {HF[0], HF[1], HF[2]};
    V3<real> sh = cross(com, Fv);                 // dAd(transf(identity, com), H)
    real* o = mom_out + (size_t)i * 6;
    o[0] = HM[0] - sh.x; o[1] = HM[1] - sh.y; o[2] = HM[2] - sh.z; o[3] = HF[0]; o[4] = HF[1]; o[5] = HF[2];
  }
#undef STK
}

}  // namespace fk
