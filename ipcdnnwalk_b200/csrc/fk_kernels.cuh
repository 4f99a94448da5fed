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


// ------------------------------------------------------------------------------------------------------------------
// Skeleton-specialised instantiation for the skeleton the SRBTrack full-body mapping uses
// (gym_trackSRB/taesooLib/hanyang_lowdof_T_sh.wrl and work/taesooLib/Resource/motion/MOB1/hanyang_lowdof_T.wrl: 20 bones,
// 53 Euler channels).  The tree structure (depth, channel count, axes, first pose column of every bone) is a compile-time
// table, the bone loop is fully unrolled, so the pose row (60 reals, loaded up front with 128-bit loads -- every byte of
// the row is requested before the first use, which is what hides the DRAM latency of the thread-per-row layout), the
// root-to-bone stack and the 140 output reals all live in registers: no shared-memory stack (the occupancy limiter of
// the generic kernel), no structure look-ups, no branches.  Numeric constants (offsets, masses, inertias) stay run-time.
struct Hanyang {
  static constexpr int NB = 20, NPOSE = 60, NDQ = 59, MAXDEPTH = 8;
  __host__ __device__ static constexpr int depth(int b) { constexpr int t[NB] = {0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 4, 5, 6, 7, 4, 5, 6, 7, 4, 5}; return t[b]; }
  __host__ __device__ static constexpr int nchan(int b) { constexpr int t[NB] = {0, 3, 1, 3, 3, 1, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3}; return t[b]; }
  __host__ __device__ static constexpr int axes(int b) { constexpr int t[NB] = {0, 9, 0, 9, 9, 0, 9, 9, 9, 9, 9, 36, 18, 18, 9, 36, 18, 18, 9, 9}; return t[b]; }
  __host__ __device__ static constexpr int start(int b) { constexpr int t[NB] = {0, 7, 10, 11, 14, 17, 18, 21, 24, 27, 30, 33, 36, 39, 42, 45, 48, 51, 54, 57}; return t[b]; }
  // parent of bone b: 0 = the previous bone, 1 = Hips, 2 = Chest2 (the two branch points of this tree)
  __host__ __device__ static constexpr int psel(int b) { constexpr int t[NB] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 2, 0}; return t[b]; }
  static constexpr int CHEST2 = 9;
  // the same tables packed into immediates for the rolled loop (a table indexed at run time would live in local memory)
  template <typename F> __host__ __device__ static constexpr unsigned long long pack(F f, int bits, int first, int count) {
    unsigned long long r = 0;
    for (int i = 0; i < count; i++) r |= (unsigned long long)f(first + i) << (bits * i);
    return r;
  }
  __device__ static __forceinline__ int nchan_rt(int b) { constexpr unsigned long long m = pack(nchan, 2, 0, NB); return (int)((m >> (2 * b)) & 3); }
  __device__ static __forceinline__ int psel_rt(int b) { constexpr unsigned long long m = pack(psel, 2, 0, NB); return (int)((m >> (2 * b)) & 3); }
  __device__ static __forceinline__ int axes_rt(int b) {
    constexpr unsigned long long lo = pack(axes, 6, 0, 10), hi = pack(axes, 6, 10, 10);
    return (int)(((b < 10 ? lo >> (6 * b) : hi >> (6 * (b - 10)))) & 63);
  }
  __device__ static __forceinline__ int start_rt(int b) {
    constexpr unsigned long long lo = pack(start, 6, 0, 10), hi = pack(start, 6, 10, 10);
    return (int)(((b < 10 ? lo >> (6 * b) : hi >> (6 * (b - 10)))) & 63);
  }
  static bool matches(const SkelConst& k) {
    if (k.nb != NB || k.n_pose != NPOSE || k.n_dq != NDQ) return false;
    for (int b = 0; b < NB; b++)
      if (k.depth[b] != depth(b) || k.nchan[b] != nchan(b) || (b > 0 && k.axes[b] != axes(b))) return false;
    return true;
  }
};

template <typename real> struct BoneNum { real off[Hanyang::NB][3], mass[Hanyang::NB], com[Hanyang::NB][3], inr[Hanyang::NB][6], itm; };

// The unrolled sweep over the 20 bones.  p / v: the pose and dq rows in registers; StoreX(k, a, b, c, d): sink of the k-th
// group of output reals (128 bits: 4 floats or 2 doubles), called as soon as its words are final.
// SYNC: a CTA-wide barrier in front of every bone keeps the warps of a CTA inside the same ~5-7 kB of this 100-150 kB
// straight-line code, so an instruction-cache line fetched for one warp serves all of them.
template <typename real, bool MOM, bool SYNC, typename StoreX>
__device__ __forceinline__ void hanyang_sweep(const BoneNum<real>& bn, const real (&p)[Hanyang::NPOSE], const real* v, bool want_x, StoreX&& store_x,
                                              V3<real>* com_o, real mom_o[6]) {
  using H = Hanyang;
  real S[H::MAXDEPTH][MOM ? 13 : 7];
  real out[H::NB * 7];
  V3<real> csum = {0, 0, 0};
  real HM[3] = {0, 0, 0}, HF[3] = {0, 0, 0};
#pragma unroll
  for (int b = 0; b < H::NB; b++) {
    if constexpr (SYNC) __syncthreads();
    const int d = H::depth(b), nc = H::nchan(b), ax = H::axes(b), st = H::start(b);
    Q4<real> q; V3<real> t, w = {0, 0, 0}, lv = {0, 0, 0};
    const V3<real> off = {bn.off[b][0], bn.off[b][1], bn.off[b][2]};
    if (b == 0) {
      q = {p[3], p[4], p[5], p[6]};
      t = {p[0] + off.x, p[1] + off.y, p[2] + off.z};
      if constexpr (MOM) {
        Q4<real> qi = qconj_unit(q);
        w = qrot(qi, V3<real>{v[0], v[1], v[2]});
        lv = qrot(qi, V3<real>{v[3], v[4], v[5]});
      }
    } else {
      const int pd = d > 0 ? d - 1 : 0;
      const Q4<real> pq = {S[pd][0], S[pd][1], S[pd][2], S[pd][3]};
      const V3<real> pt = {S[pd][4], S[pd][5], S[pd][6]};
      Q4<real> lq = {real(1), real(0), real(0), real(0)};
      if constexpr (MOM) {
        V3<real> pw = {S[pd][7], S[pd][8], S[pd][9]};
        V3<real> pv = {S[pd][10], S[pd][11], S[pd][12]};
        w = pw;
        lv = pv + cross(pw, off);
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (k < nc) {
          const int a = (ax >> (2 * k)) & 3;
          real s, c;
          num<real>::sincos(real(0.5) * p[st + k], &s, &c);
          Q4<real> r;
          if (a == 0) r = {lq.w * c - lq.x * s, lq.x * c + lq.w * s, lq.y * c + lq.z * s, lq.z * c - lq.y * s};
          else if (a == 1) r = {lq.w * c - lq.y * s, lq.x * c - lq.z * s, lq.y * c + lq.w * s, lq.z * c + lq.x * s};
          else r = {lq.w * c - lq.z * s, lq.x * c + lq.y * s, lq.y * c - lq.x * s, lq.z * c + lq.w * s};
          lq = r;
          if constexpr (MOM) {
            const real C = c * c - s * s, Sn = real(2) * s * c;
            const real rate = v[6 + (st - 7) + k];
            if (a == 0) { w = {w.x, C * w.y + Sn * w.z, C * w.z - Sn * w.y}; lv = {lv.x, C * lv.y + Sn * lv.z, C * lv.z - Sn * lv.y}; w.x += rate; }
            else if (a == 1) { w = {C * w.x - Sn * w.z, w.y, C * w.z + Sn * w.x}; lv = {C * lv.x - Sn * lv.z, lv.y, C * lv.z + Sn * lv.x}; w.y += rate; }
            else { w = {C * w.x + Sn * w.y, C * w.y - Sn * w.x, w.z}; lv = {C * lv.x + Sn * lv.y, C * lv.y - Sn * lv.x, lv.z}; w.z += rate; }
          }
        }
      }
      q = quat_mul(pq, lq);
      t = qrot(pq, off) + pt;
    }
    S[d][0] = q.w; S[d][1] = q.x; S[d][2] = q.y; S[d][3] = q.z; S[d][4] = t.x; S[d][5] = t.y; S[d][6] = t.z;
    if constexpr (MOM) { S[d][7] = w.x; S[d][8] = w.y; S[d][9] = w.z; S[d][10] = lv.x; S[d][11] = lv.y; S[d][12] = lv.z; }
    out[7 * b] = q.w; out[7 * b + 1] = q.x; out[7 * b + 2] = q.y; out[7 * b + 3] = q.z; out[7 * b + 4] = t.x; out[7 * b + 5] = t.y; out[7 * b + 6] = t.z;
    const real m = bn.mass[b];
    const V3<real> c = {bn.com[b][0], bn.com[b][1], bn.com[b][2]};
    csum = csum + m * (qrot(q, c) + t);
    if constexpr (MOM) {
      const real Ixx = bn.inr[b][0], Iyy = bn.inr[b][1], Izz = bn.inr[b][2], Ixy = bn.inr[b][3], Ixz = bn.inr[b][4], Iyz = bn.inr[b][5];
      const V3<real> r = m * c;
      V3<real> M = {Ixx * w.x + Ixy * w.y + Ixz * w.z + r.y * lv.z - r.z * lv.y,
                    Ixy * w.x + Iyy * w.y + Iyz * w.z + r.z * lv.x - r.x * lv.z,
                    Ixz * w.x + Iyz * w.y + Izz * w.z + r.x * lv.y - r.y * lv.x};
      V3<real> F = {w.y * r.z - w.z * r.y + m * lv.x, w.z * r.x - w.x * r.z + m * lv.y, w.x * r.y - w.y * r.x + m * lv.z};
      V3<real> Fw = qrot(q, F);
      V3<real> Mw = qrot(q, M) + cross(t, Fw);
      HM[0] += Mw.x; HM[1] += Mw.y; HM[2] += Mw.z; HF[0] += Fw.x; HF[1] += Fw.y; HF[2] += Fw.z;
    }
    if (want_x) {
      constexpr int W = sizeof(real) == 4 ? 4 : 2;
      const int done_words = 7 * (b + 1), prev_words = 7 * b;
#pragma unroll
      for (int k = prev_words / W; k < done_words / W; k++) {
        if constexpr (sizeof(real) == 4) store_x(k, out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
        else store_x(k, out[2 * k], out[2 * k + 1], real(0), real(0));
      }
    }
  }
  const V3<real> com = bn.itm * csum;
  *com_o = com;
  if constexpr (MOM) {
    V3<real> Fv = {HF[0], HF[1], HF[2]};
    V3<real> sh = cross(com, Fv);
    mom_o[0] = HM[0] - sh.x; mom_o[1] = HM[1] - sh.y; mom_o[2] = HM[2] - sh.z; mom_o[3] = HF[0]; mom_o[4] = HF[1]; mom_o[5] = HF[2];
  }
}

template <typename real, int BLOCK> __device__ __forceinline__ void load_bone_numbers(BoneNum<real>& bn, const SkelConst* __restrict__ skp) {
  for (int k = threadIdx.x; k < Hanyang::NB; k += BLOCK) {
    for (int c = 0; c < 3; c++) { bn.off[k][c] = (real)skp->offset[k][c]; bn.com[k][c] = (real)skp->com[k][c]; }
    for (int c = 0; c < 6; c++) bn.inr[k][c] = (real)skp->inertia[k][c];
    bn.mass[k] = (real)skp->mass[k];
  }
  if (threadIdx.x == 0) bn.itm = (real)skp->inv_total_mass;
}

// Variant A: rows read / written directly by the owning thread (used for the ragged tail and unaligned buffers).
template <typename real, bool MOM, int BLOCK>
__global__ void __launch_bounds__(BLOCK) fk_hanyang_kernel(const SkelConst* __restrict__ skp, int n, const real* __restrict__ pose,
                                                           const real* __restrict__ dq, real* __restrict__ xform,
                                                           real* __restrict__ com_out, real* __restrict__ mom_out) {
  using H = Hanyang;
  __shared__ BoneNum<real> bn;
  load_bone_numbers<real, BLOCK>(bn, skp);
  __syncthreads();
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  real p[H::NPOSE];
  if constexpr (sizeof(real) == 4) {
    const float4* r = reinterpret_cast<const float4*>(pose + (size_t)i * H::NPOSE);
#pragma unroll
    for (int k = 0; k < H::NPOSE / 4; k++) { float4 t = __ldg(r + k); p[4 * k] = t.x; p[4 * k + 1] = t.y; p[4 * k + 2] = t.z; p[4 * k + 3] = t.w; }
  } else {
    const double2* r = reinterpret_cast<const double2*>(pose + (size_t)i * H::NPOSE);
#pragma unroll
    for (int k = 0; k < H::NPOSE / 2; k++) { double2 t = __ldg(r + k); p[2 * k] = t.x; p[2 * k + 1] = t.y; }
  }
  real v[MOM ? H::NDQ : 1];
  if constexpr (MOM) {
    const real* r = dq + (size_t)i * H::NDQ;
#pragma unroll
    for (int k = 0; k < H::NDQ; k++) v[k] = __ldg(r + k);
  }
  real* xrow = xform ? xform + (size_t)i * (H::NB * 7) : nullptr;
  V3<real> com; real mom[6];
  hanyang_sweep<real, MOM, false>(bn, p, v, xrow != nullptr, [&](int k, real a, real b, real c, real d) {
    if constexpr (sizeof(real) == 4) reinterpret_cast<float4*>(xrow)[k] = make_float4(a, b, c, d);
    else reinterpret_cast<double2*>(xrow)[k] = make_double2(a, b);
  }, &com, mom);
  if (com_out) { real* o = com_out + (size_t)i * 3; o[0] = com.x; o[1] = com.y; o[2] = com.z; }
  if constexpr (MOM) {
    real* o = mom_out + (size_t)i * 6;
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = mom[k];
  }
}

// Variant B (full tiles of BLOCK poses): TMA-staged rows and a ROLLED bone loop.
//  * I/O.  With rows accessed straight from global memory every warp-level LDG / STG touches 32 different 128-byte lines
//    and the L1TEX wavefront queue (one FIFO per SM, ~2 cycles per line) is the bound: (15 + 35) requests x 32 lines x 2
//    cycles per 32 poses = 0.36 ms per 1M poses against 0.13 ms of HBM time.  Here every thread owns a 140-real slot of
//    shared memory.  Its pose row is dropped into the LAST 60 reals of the slot by a TMA bulk copy (cp.async.bulk, one per
//    row, all completing on one mbarrier); the dq tile is one more bulk copy.  The 140 output reals are written into the
//    slot from its start -- bone b's seven reals end at word 7(b+1) <= 80 + (first pose column of bone b+1), i.e. they
//    never reach an angle that is still unread -- and the whole tile of transforms, then CoM and momentum, leaves by
//    bulk stores.  No LSU traffic to global memory at all.
//  * Code size.  The unrolled sweep is 104 kB (FK) / 146 kB (momentum) of SASS, the instruction caches hold 6 kB (L0) and
//    32 kB (L1.5): ncu showed 26 % / 61 % of the stall samples on instruction fetch.  The tree of this skeleton has two
//    branch points only (Hips, Chest2), so the root-to-bone stack collapses to "previous bone, Hips, Chest2" in registers
//    and the bone loop can stay rolled: one bone body (~10 kB), angles and rates read from shared memory by index.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <typename real> struct BoneState { Q4<real> q; V3<real> t, w, lv; };
template <typename real, bool MOM, int BLOCK>
__global__ void __launch_bounds__(BLOCK) fk_hanyang_tma_rolled_kernel(const SkelConst* __restrict__ skp, const real* __restrict__ pose,
                                                               const real* __restrict__ dq, real* __restrict__ xform,
                                                               real* __restrict__ com_out, real* __restrict__ mom_out) {
  using H = Hanyang;
  constexpr int XW = H::NB * 7;                               // 140 reals per slot
  constexpr int PIN = XW - H::NPOSE;                          // the pose row sits at words [80, 140) of the slot
  extern __shared__ __align__(128) unsigned char fk_tile[];
  real* sx = reinterpret_cast<real*>(fk_tile);                // [BLOCK][140]
  real* sc = sx + BLOCK * XW;                                  // [BLOCK][3]
  real* sm = sc + BLOCK * 3;                                   // [BLOCK][6]
  real* sdq = sm + BLOCK * 6;                                  // [BLOCK][59]  (momentum only)
  __shared__ BoneNum<real> bn;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * BLOCK;
  constexpr unsigned row_bytes = H::NPOSE * sizeof(real), dq_bytes = BLOCK * H::NDQ * sizeof(real);
  static_assert(row_bytes % 16 == 0 && dq_bytes % 16 == 0 && (XW * sizeof(real)) % 16 == 0 && (PIN * sizeof(real)) % 16 == 0 &&
                (BLOCK * 3 * sizeof(real)) % 16 == 0, "bulk copies need 16-byte multiples");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(BLOCK * row_bytes + (MOM ? dq_bytes : 0u)) : "memory");
    if constexpr (MOM)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sdq)), "l"(dq + row0 * H::NDQ), "r"(dq_bytes), "r"(smem_u32(&mbar)) : "memory");
  }
  load_bone_numbers<real, BLOCK>(bn, skp);
  __syncthreads();                                             // barrier armed, bone numbers in place
  real* slot = sx + tid * XW;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(slot + PIN)), "l"(pose + (row0 + tid) * H::NPOSE), "r"(row_bytes), "r"(smem_u32(&mbar)) : "memory");
  {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0, 1, 0, P; }" : "=r"(ok) : "r"(smem_u32(&mbar)) : "memory");
  }
  const real* p = slot + PIN;
  const real* v = sdq + tid * H::NDQ;
  const bool want_x = xform != nullptr;
  BoneState<real> cur, hips, chest2;
  V3<real> csum = {0, 0, 0};
  real HM[3] = {0, 0, 0}, HF[3] = {0, 0, 0};
#pragma unroll 1
  for (int b = 0; b < H::NB; b++) {
    const int nc = H::nchan_rt(b), ax = H::axes_rt(b), st = H::start_rt(b), sel = H::psel_rt(b);
    const V3<real> off = {bn.off[b][0], bn.off[b][1], bn.off[b][2]};
    Q4<real> q; V3<real> t, w = {0, 0, 0}, lv = {0, 0, 0};
    if (b == 0) {
      q = {p[3], p[4], p[5], p[6]};
      t = {p[0] + off.x, p[1] + off.y, p[2] + off.z};
      if constexpr (MOM) {
        Q4<real> qi = qconj_unit(q);
        w = qrot(qi, V3<real>{v[0], v[1], v[2]});
        lv = qrot(qi, V3<real>{v[3], v[4], v[5]});
      }
    } else {
      BoneState<real> par = cur;
      if (sel == 1) par = hips; else if (sel == 2) par = chest2;
      Q4<real> lq = {real(1), real(0), real(0), real(0)};
      if constexpr (MOM) { w = par.w; lv = par.lv + cross(par.w, off); }
#pragma unroll 1
      for (int k = 0; k < nc; k++) {
        const int a = (ax >> (2 * k)) & 3;
        real s, c;
        num<real>::sincos(real(0.5) * p[st + k], &s, &c);
        // lq * (c, s e_a) written once for all axes: e_a picks the component pair that mixes
        const real e0 = a == 0 ? s : real(0), e1 = a == 1 ? s : real(0), e2 = a == 2 ? s : real(0);
        lq = {lq.w * c - lq.x * e0 - lq.y * e1 - lq.z * e2, lq.x * c + lq.w * e0 + lq.y * e2 - lq.z * e1,
              lq.y * c + lq.w * e1 + lq.z * e0 - lq.x * e2, lq.z * c + lq.w * e2 + lq.x * e1 - lq.y * e0};
        if constexpr (MOM) {
          const real C = c * c - s * s, Sn = real(2) * s * c;
          const real rate = v[6 + (st - 7) + k];
          if (a == 0) { w = {w.x, C * w.y + Sn * w.z, C * w.z - Sn * w.y}; lv = {lv.x, C * lv.y + Sn * lv.z, C * lv.z - Sn * lv.y}; w.x += rate; }
          else if (a == 1) { w = {C * w.x - Sn * w.z, w.y, C * w.z + Sn * w.x}; lv = {C * lv.x - Sn * lv.z, lv.y, C * lv.z + Sn * lv.x}; w.y += rate; }
          else { w = {C * w.x + Sn * w.y, C * w.y - Sn * w.x, w.z}; lv = {C * lv.x + Sn * lv.y, C * lv.y - Sn * lv.x, lv.z}; w.z += rate; }
        }
      }
      q = quat_mul(par.q, lq);
      t = qrot(par.q, off) + par.t;
    }
    cur.q = q; cur.t = t; cur.w = w; cur.lv = lv;
    if (b == 0) hips = cur;
    if (b == H::CHEST2) chest2 = cur;
    if (want_x) {                                             // the angles of this bone have been consumed: overwrite from the slot start
      real* o = slot + 7 * b;
      o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z; o[4] = t.x; o[5] = t.y; o[6] = t.z;
    }
    const real m = bn.mass[b];
    const V3<real> c = {bn.com[b][0], bn.com[b][1], bn.com[b][2]};
    csum = csum + m * (qrot(q, c) + t);
    if constexpr (MOM) {
      const real Ixx = bn.inr[b][0], Iyy = bn.inr[b][1], Izz = bn.inr[b][2], Ixy = bn.inr[b][3], Ixz = bn.inr[b][4], Iyz = bn.inr[b][5];
      const V3<real> r = m * c;
      V3<real> M = {Ixx * w.x + Ixy * w.y + Ixz * w.z + r.y * lv.z - r.z * lv.y,
                    Ixy * w.x + Iyy * w.y + Iyz * w.z + r.z * lv.x - r.x * lv.z,
                    Ixz * w.x + Iyz * w.y + Izz * w.z + r.x * lv.y - r.y * lv.x};
      V3<real> F = {w.y * r.z - w.z * r.y + m * lv.x, w.z * r.x - w.x * r.z + m * lv.y, w.x * r.y - w.y * r.x + m * lv.z};
      V3<real> Fw = qrot(q, F);
      V3<real> Mw = qrot(q, M) + cross(t, Fw);
      HM[0] += Mw.x; HM[1] += Mw.y; HM[2] += Mw.z; HF[0] += Fw.x; HF[1] += Fw.y; HF[2] += Fw.z;
    }
  }
  const V3<real> com = bn.itm * csum;
  sc[tid * 3] = com.x; sc[tid * 3 + 1] = com.y; sc[tid * 3 + 2] = com.z;
  if constexpr (MOM) {
    V3<real> Fv = {HF[0], HF[1], HF[2]};
    V3<real> sh = cross(com, Fv);
    real* o = sm + tid * 6;
    o[0] = HM[0] - sh.x; o[1] = HM[1] - sh.y; o[2] = HM[2] - sh.z; o[3] = HF[0]; o[4] = HF[1]; o[5] = HF[2];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk-copy engine
  __syncthreads();
  if (tid == 0) {
    if (want_x)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(xform + row0 * XW), "r"(smem_u32(sx)), "r"((unsigned)(BLOCK * XW * sizeof(real))) : "memory");
    if (com_out != nullptr)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(com_out + row0 * 3), "r"(smem_u32(sc)), "r"((unsigned)(BLOCK * 3 * sizeof(real))) : "memory");
    if constexpr (MOM)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mom_out + row0 * 6), "r"(smem_u32(sm)), "r"((unsigned)(BLOCK * 6 * sizeof(real))) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must stay alive until the engine has read it
  }
}

// Variant D (full tiles): TMA-staged rows as in variant B, CHAIN-WISE unrolled code.  The skeleton is five chains hanging off two
// branch points: two legs (Hip YZX, Knee X, Ankle YZX), the spine (3 x YZX), two arms (Collar YZX, Shoulder XYZ, Elbow ZXY, Wrist
// ZXY) and Neck / Head (2 x YZX).  Left and right limbs have the same channel structure, so ONE unrolled leg body and ONE unrolled
// arm body run twice (side loop), and one YZX bone body serves the spine and Neck / Head (5 trips).  Channel axes are compile-time
// constants inside a body (axis-specialised quaternion and velocity updates: the instruction count of the fully unrolled sweep,
// 200 M warp instructions per 1 M poses instead of the rolled loop's 345 M), while the code is 22 channel bodies + 8 bone
// epilogues instead of 53 + 20: ~40 % of the unrolled kernel's 104 / 146 kB, and every body's second trip hits the
// instruction cache.  Bones are still visited in index order except Neck / Head, whose 14 output reals wait in registers
// until the arms have consumed their angles (the in-place slot overwrite of variant B needs index order).
template <typename real, bool MOM, int NC, int AX>
__device__ __forceinline__ void bone_fixed(const BoneState<real>& par, BoneState<real>& cur, const real* __restrict__ ang, const real* __restrict__ rate, V3<real> off) {
  Q4<real> lq = {real(1), real(0), real(0), real(0)};
  V3<real> w = {0, 0, 0}, lv = {0, 0, 0};
  if constexpr (MOM) { w = par.w; lv = par.lv + cross(par.w, off); }
#pragma unroll
  for (int k = 0; k < NC; k++) {
    constexpr int dummy = 0; (void)dummy;
    const int a = (AX >> (2 * k)) & 3;                       // compile-time after unrolling
    real s, c;
    num<real>::sincos(real(0.5) * ang[k], &s, &c);
    Q4<real> r;
    if (a == 0) r = {lq.w * c - lq.x * s, lq.x * c + lq.w * s, lq.y * c + lq.z * s, lq.z * c - lq.y * s};
    else if (a == 1) r = {lq.w * c - lq.y * s, lq.x * c - lq.z * s, lq.y * c + lq.w * s, lq.z * c + lq.x * s};
    else r = {lq.w * c - lq.z * s, lq.x * c + lq.y * s, lq.y * c - lq.x * s, lq.z * c + lq.w * s};
    lq = r;
    if constexpr (MOM) {
      const real C = c * c - s * s, Sn = real(2) * s * c;
      const real rt = rate[k];
      if (a == 0) { w = {w.x, C * w.y + Sn * w.z, C * w.z - Sn * w.y}; lv = {lv.x, C * lv.y + Sn * lv.z, C * lv.z - Sn * lv.y}; w.x += rt; }
      else if (a == 1) { w = {C * w.x - Sn * w.z, w.y, C * w.z + Sn * w.x}; lv = {C * lv.x - Sn * lv.z, lv.y, C * lv.z + Sn * lv.x}; w.y += rt; }
      else { w = {C * w.x + Sn * w.y, C * w.y - Sn * w.x, w.z}; lv = {C * lv.x + Sn * lv.y, C * lv.y - Sn * lv.x, lv.z}; w.z += rt; }
    }
  }
  cur.q = quat_mul(par.q, lq);
  cur.t = qrot(par.q, off) + par.t;
  cur.w = w; cur.lv = lv;
}

template <typename real, bool MOM> struct FkAccum {
  V3<real> csum = {0, 0, 0};
  real HM[3] = {0, 0, 0}, HF[3] = {0, 0, 0};
  __device__ __forceinline__ void add(const BoneNum<real>& bn, int b, const BoneState<real>& B) {
    const real m = bn.mass[b];
    const V3<real> c = {bn.com[b][0], bn.com[b][1], bn.com[b][2]};
    csum = csum + m * (qrot(B.q, c) + B.t);
    if constexpr (MOM) {
      const real Ixx = bn.inr[b][0], Iyy = bn.inr[b][1], Izz = bn.inr[b][2], Ixy = bn.inr[b][3], Ixz = bn.inr[b][4], Iyz = bn.inr[b][5];
      const V3<real> r = m * c;
      const V3<real> w = B.w, lv = B.lv;
      V3<real> M = {Ixx * w.x + Ixy * w.y + Ixz * w.z + r.y * lv.z - r.z * lv.y,
                    Ixy * w.x + Iyy * w.y + Iyz * w.z + r.z * lv.x - r.x * lv.z,
                    Ixz * w.x + Iyz * w.y + Izz * w.z + r.x * lv.y - r.y * lv.x};
      V3<real> F = {w.y * r.z - w.z * r.y + m * lv.x, w.z * r.x - w.x * r.z + m * lv.y, w.x * r.y - w.y * r.x + m * lv.z};
      V3<real> Fw = qrot(B.q, F);
      V3<real> Mw = qrot(B.q, M) + cross(B.t, Fw);
      HM[0] += Mw.x; HM[1] += Mw.y; HM[2] += Mw.z; HF[0] += Fw.x; HF[1] += Fw.y; HF[2] += Fw.z;
    }
  }
};

template <typename real, bool MOM, int BLOCK>
__global__ void __launch_bounds__(BLOCK) fk_hanyang_tma_chain_kernel(const SkelConst* __restrict__ skp, const real* __restrict__ pose,
                                                              const real* __restrict__ dq, real* __restrict__ xform,
                                                              real* __restrict__ com_out, real* __restrict__ mom_out) {
  using H = Hanyang;
  constexpr int XW = H::NB * 7;
  constexpr int PIN = XW - H::NPOSE;
  constexpr int YZX = 9, XYZ = 36, ZXY = 18, X = 0;
  extern __shared__ __align__(128) unsigned char fk_tile[];
  real* sx = reinterpret_cast<real*>(fk_tile);                // [BLOCK][140]
  real* sc = sx + BLOCK * XW;                                  // [BLOCK][3]
  real* sm = sc + BLOCK * 3;                                   // [BLOCK][6]
  real* sdq = sm + BLOCK * 6;                                  // [BLOCK][59]  (momentum only)
  __shared__ BoneNum<real> bn;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * BLOCK;
  constexpr unsigned row_bytes = H::NPOSE * sizeof(real), dq_bytes = BLOCK * H::NDQ * sizeof(real);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(BLOCK * row_bytes + (MOM ? dq_bytes : 0u)) : "memory");
    if constexpr (MOM)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sdq)), "l"(dq + row0 * H::NDQ), "r"(dq_bytes), "r"(smem_u32(&mbar)) : "memory");
  }
  load_bone_numbers<real, BLOCK>(bn, skp);
  __syncthreads();
  real* slot = sx + tid * XW;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(slot + PIN)), "l"(pose + (row0 + tid) * H::NPOSE), "r"(row_bytes), "r"(smem_u32(&mbar)) : "memory");
  {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0, 1, 0, P; }" : "=r"(ok) : "r"(smem_u32(&mbar)) : "memory");
  }
  const real* p = slot + PIN;
  const real* v = sdq + tid * H::NDQ - 1;                      // rate of pose column c (c >= 7) = v[c]: dq index 6 + (c - 7)
  const bool want_x = xform != nullptr;
  FkAccum<real, MOM> acc;
  auto offv = [&](int b) { return V3<real>{bn.off[b][0], bn.off[b][1], bn.off[b][2]}; };
  auto put = [&](int b, const BoneState<real>& B) {
    if (want_x) { real* o = slot + 7 * b; o[0] = B.q.w; o[1] = B.q.x; o[2] = B.q.y; o[3] = B.q.z; o[4] = B.t.x; o[5] = B.t.y; o[6] = B.t.z; }
  };
  BoneState<real> hips, chest2;
  {                                                            // Hips: the free root
    const V3<real> off = offv(0);
    hips.q = {p[3], p[4], p[5], p[6]};
    hips.t = {p[0] + off.x, p[1] + off.y, p[2] + off.z};
    hips.w = {0, 0, 0}; hips.lv = {0, 0, 0};
    if constexpr (MOM) {
      Q4<real> qi = qconj_unit(hips.q);
      hips.w = qrot(qi, V3<real>{sdq[tid * H::NDQ], sdq[tid * H::NDQ + 1], sdq[tid * H::NDQ + 2]});
      hips.lv = qrot(qi, V3<real>{sdq[tid * H::NDQ + 3], sdq[tid * H::NDQ + 4], sdq[tid * H::NDQ + 5]});
    }
    put(0, hips); acc.add(bn, 0, hips);
  }
#pragma unroll 1
  for (int side = 0; side < 2; side++) {                       // legs: bones 1-3 / 4-6, pose columns 7-13 / 14-20
    const int b0 = 1 + 3 * side, st = 7 + 7 * side;
    BoneState<real> hip, knee, ankle;
    bone_fixed<real, MOM, 3, YZX>(hips, hip, p + st, v + st, offv(b0)); put(b0, hip); acc.add(bn, b0, hip);
    bone_fixed<real, MOM, 1, X>(hip, knee, p + st + 3, v + st + 3, offv(b0 + 1)); put(b0 + 1, knee); acc.add(bn, b0 + 1, knee);
    bone_fixed<real, MOM, 3, YZX>(knee, ankle, p + st + 4, v + st + 4, offv(b0 + 2)); put(b0 + 2, ankle); acc.add(bn, b0 + 2, ankle);
  }
  real pend[14];                                               // Neck / Head outputs, stored after the arms
  {
    BoneState<real> par = hips, cur;
#pragma unroll 1
    for (int k = 0; k < 5; k++) {                              // Chest, Chest1, Chest2 (7-9, columns 21-29), then Neck, Head (18-19, columns 54-59)
      const int b = k < 3 ? 7 + k : 15 + k, st = k < 3 ? 21 + 3 * k : 45 + 3 * k;
      bone_fixed<real, MOM, 3, YZX>(par, cur, p + st, v + st, offv(b));
      acc.add(bn, b, cur);
      if (k < 3) put(b, cur);
      else {
        const int o = 7 * (k - 3);
        if (k == 3) { pend[0] = cur.q.w; pend[1] = cur.q.x; pend[2] = cur.q.y; pend[3] = cur.q.z; pend[4] = cur.t.x; pend[5] = cur.t.y; pend[6] = cur.t.z; }
        else { pend[7] = cur.q.w; pend[8] = cur.q.x; pend[9] = cur.q.y; pend[10] = cur.q.z; pend[11] = cur.t.x; pend[12] = cur.t.y; pend[13] = cur.t.z; }
        (void)o;
      }
      if (k == 2) { chest2 = cur; par = chest2; } else par = cur;
    }
  }
#pragma unroll 1
  for (int side = 0; side < 2; side++) {                       // arms: bones 10-13 / 14-17, pose columns 30-41 / 42-53
    const int b0 = 10 + 4 * side, st = 30 + 12 * side;
    BoneState<real> collar, shoulder, elbow, wrist;
    bone_fixed<real, MOM, 3, YZX>(chest2, collar, p + st, v + st, offv(b0)); put(b0, collar); acc.add(bn, b0, collar);
    bone_fixed<real, MOM, 3, XYZ>(collar, shoulder, p + st + 3, v + st + 3, offv(b0 + 1)); put(b0 + 1, shoulder); acc.add(bn, b0 + 1, shoulder);
    bone_fixed<real, MOM, 3, ZXY>(shoulder, elbow, p + st + 6, v + st + 6, offv(b0 + 2)); put(b0 + 2, elbow); acc.add(bn, b0 + 2, elbow);
    bone_fixed<real, MOM, 3, ZXY>(elbow, wrist, p + st + 9, v + st + 9, offv(b0 + 3)); put(b0 + 3, wrist); acc.add(bn, b0 + 3, wrist);
  }
  if (want_x) {
#pragma unroll
    for (int i = 0; i < 14; i++) slot[7 * 18 + i] = pend[i];
  }
  const V3<real> com = bn.itm * acc.csum;
  sc[tid * 3] = com.x; sc[tid * 3 + 1] = com.y; sc[tid * 3 + 2] = com.z;
  if constexpr (MOM) {
    V3<real> Fv = {acc.HF[0], acc.HF[1], acc.HF[2]};
    V3<real> sh = cross(com, Fv);
    real* o = sm + tid * 6;
    o[0] = acc.HM[0] - sh.x; o[1] = acc.HM[1] - sh.y; o[2] = acc.HM[2] - sh.z; o[3] = acc.HF[0]; o[4] = acc.HF[1]; o[5] = acc.HF[2];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    if (want_x)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(xform + row0 * XW), "r"(smem_u32(sx)), "r"((unsigned)(BLOCK * XW * sizeof(real))) : "memory");
    if (com_out != nullptr)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(com_out + row0 * 3), "r"(smem_u32(sc)), "r"((unsigned)(BLOCK * 3 * sizeof(real))) : "memory");
    if constexpr (MOM)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mom_out + row0 * 6), "r"(smem_u32(sm)), "r"((unsigned)(BLOCK * 6 * sizeof(real))) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// Variant C (full tiles): TMA-staged rows with the UNROLLED sweep.  Pose / dq tiles arrive by two bulk copies, every thread
// pulls its row into registers with conflict-free 128-bit shared loads (pitch 240 / 560 bytes: the eight 16-byte spans of
// a quarter-warp fall in distinct bank groups), the region is then reused for the output tile, which leaves by bulk stores.
// SYNC (see hanyang_sweep) is what makes the 100-150 kB of straight-line code affordable.
template <typename real, bool MOM, bool SYNC, int BLOCK>
__global__ void __launch_bounds__(BLOCK) fk_hanyang_tma_unrolled_kernel(const SkelConst* __restrict__ skp, const real* __restrict__ pose,
                                                                        const real* __restrict__ dq, real* __restrict__ xform,
                                                                        real* __restrict__ com_out, real* __restrict__ mom_out) {
  using H = Hanyang;
  constexpr int XW = H::NB * 7;
  extern __shared__ __align__(128) unsigned char fk_tile[];
  real* sx = reinterpret_cast<real*>(fk_tile);                // [BLOCK][140]   (first: pose rows [BLOCK][60], dq rows [BLOCK][59])
  real* sc = sx + BLOCK * XW;
  real* sm = sc + BLOCK * 3;
  __shared__ BoneNum<real> bn;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * BLOCK;
  constexpr unsigned in_bytes = BLOCK * H::NPOSE * sizeof(real), dq_bytes = BLOCK * H::NDQ * sizeof(real);
  static_assert(in_bytes % 16 == 0 && dq_bytes % 16 == 0 && (BLOCK * 3 * sizeof(real)) % 16 == 0, "tile sizes must be multiples of 16 bytes");
  real* sdq = sx + BLOCK * H::NPOSE;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(in_bytes + (MOM ? dq_bytes : 0u)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sx)), "l"(pose + row0 * H::NPOSE), "r"(in_bytes), "r"(smem_u32(&mbar)) : "memory");
    if constexpr (MOM)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sdq)), "l"(dq + row0 * H::NDQ), "r"(dq_bytes), "r"(smem_u32(&mbar)) : "memory");
  }
  load_bone_numbers<real, BLOCK>(bn, skp);
  __syncthreads();
  {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0, 1, 0, P; }" : "=r"(ok) : "r"(smem_u32(&mbar)) : "memory");
  }
  real p[H::NPOSE];
  if constexpr (sizeof(real) == 4) {
    const float4* r = reinterpret_cast<const float4*>(sx + tid * H::NPOSE);
#pragma unroll
    for (int k = 0; k < H::NPOSE / 4; k++) { float4 t = r[k]; p[4 * k] = t.x; p[4 * k + 1] = t.y; p[4 * k + 2] = t.z; p[4 * k + 3] = t.w; }
  } else {
    const double2* r = reinterpret_cast<const double2*>(sx + tid * H::NPOSE);
#pragma unroll
    for (int k = 0; k < H::NPOSE / 2; k++) { double2 t = r[k]; p[2 * k] = t.x; p[2 * k + 1] = t.y; }
  }
  real v[MOM ? H::NDQ : 1];
  if constexpr (MOM) {
#pragma unroll
    for (int k = 0; k < H::NDQ; k++) v[k] = sdq[tid * H::NDQ + k];
  }
  __syncthreads();                                             // all inputs are in registers: the region becomes the output tile
  real* xrow = sx + tid * XW;
  V3<real> com; real mom[6];
  hanyang_sweep<real, MOM, SYNC>(bn, p, v, xform != nullptr, [&](int k, real a, real b, real c, real d) {
    if constexpr (sizeof(real) == 4) reinterpret_cast<float4*>(xrow)[k] = make_float4(a, b, c, d);
    else reinterpret_cast<double2*>(xrow)[k] = make_double2(a, b);
  }, &com, mom);
  sc[tid * 3] = com.x; sc[tid * 3 + 1] = com.y; sc[tid * 3 + 2] = com.z;
  if constexpr (MOM) {
#pragma unroll
    for (int k = 0; k < 6; k++) sm[tid * 6 + k] = mom[k];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    if (xform != nullptr)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(xform + row0 * XW), "r"(smem_u32(sx)), "r"((unsigned)(BLOCK * XW * sizeof(real))) : "memory");
    if (com_out != nullptr)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(com_out + row0 * 3), "r"(smem_u32(sc)), "r"((unsigned)(BLOCK * 3 * sizeof(real))) : "memory");
    if constexpr (MOM)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mom_out + row0 * 6), "r"(smem_u32(sm)), "r"((unsigned)(BLOCK * 6 * sizeof(real))) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

}  // namespace fk
