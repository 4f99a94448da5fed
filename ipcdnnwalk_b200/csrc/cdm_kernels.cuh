// AdaptiveSRB (gym_cdm2 `deepmimiccdm-v1`, spec walk2-v1, training mode) batched environment step for sm_100a.
// SURVEY 8a rows a-16 ... a-20.  Reference paths are relative to the reference repository root.
//
// Mapping: as in srb_kernels.cuh -- G lanes of a warp cooperate on one environment, one warp per CTA.  The foot
// state machine, contact filter, reward and observation are scalar and computed redundantly by the G lanes; the
// contact QP (QPservo:_QPsolve -> eiquadprog) is spread over the lanes with the dual semismooth Newton of
// srb::QP.  After eliminating (qdd, tau) through the equality constraints (M is diagonal for the single box) the
// problem is   min_{lam>=0} |A lam - y|^2 + (w_lam/w_ddq)|lam|^2,   A = M^-1 J^T V,   y = a_des + M^-1 b.
// Every stance point is pushed twice by the reference (Q7); the duplicated columns carry equal multipliers, so the
// 40-column problem equals the 20-column one with half the ridge.  The ridge is 1e-8: the Newton matrix
// I + 2e8 * A_S A_S^T has condition ~1e8 (rank 5 with a single stance foot), so the QP is evaluated in fp64 in
// both precision modes; everything else follows `real`.
#pragma once
#include "srb_kernels.cuh"
#include "cdm_types.h"

namespace cdm {
using srb::V3; using srb::Q4; using srb::num; using srb::dot; using srb::cross; using srb::norm; using srb::quat_mul;

#define CDM_PI 3.14159265358979323846

// ---- taesooLib quaternion / vector helpers (BaseLib/math/quater.cpp, vector3.cpp; restated in oracle/tl_math.py) ----
template <typename T> __device__ __forceinline__ V3<T> qrot(Q4<T> q, V3<T> v) {                 // vector3::rotate
  V3<T> u = {q.x, q.y, q.z};
  T s = q.w;
  return (T(2) * dot(u, v)) * u + (s * s - dot(u, u)) * v + (T(2) * s) * cross(u, v);
}
template <typename T> __device__ __forceinline__ Q4<T> qnormalize(Q4<T> q) {
  T inv = T(1) / num<T>::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w * inv, q.x * inv, q.y * inv, q.z * inv};
}
template <typename T> __device__ __forceinline__ Q4<T> qinv(Q4<T> q) { return qnormalize(Q4<T>{q.w, -q.x, -q.y, -q.z}); }
template <typename T> __device__ __forceinline__ T qdot(Q4<T> a, Q4<T> b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __device__ __forceinline__ Q4<T> qneg(Q4<T> q) { return {-q.w, -q.x, -q.y, -q.z}; }
template <typename T> __device__ __forceinline__ Q4<T> qalign(Q4<T> q, Q4<T> other) { return qdot(q, other) < T(0) ? qneg(q) : q; }
template <typename T> __device__ __forceinline__ Q4<T> qaxisY(T ang) {
  T s, c;
  num<T>::sincos(T(0.5) * ang, &s, &c);
  return {c, T(0), s, T(0)};
}
template <typename T> __device__ __forceinline__ V3<T> vnormalize(V3<T> v) { return (T(1) / norm(v)) * v; }

template <typename T> __device__ __forceinline__ Q4<T> axis_to_axis_core(V3<T> va, V3<T> vb) {
  V3<T> vh = vnormalize(va + vb);
  V3<T> ax = cross(va, vh);
  return {dot(va, vh), ax.x, ax.y, ax.z};
}
// quater::axisToAxis (quater.cpp:307-325)
template <typename T> __device__ __forceinline__ Q4<T> axis_to_axis(V3<T> vfrom, V3<T> vto) {
  V3<T> va = vnormalize(vfrom), vb = vnormalize(vto);
  if (dot(va, vb) < T(-0.999)) {
    Q4<T> q = axis_to_axis_core(va, T(-1) * vb);
    return quat_mul(Q4<T>{T(6.123233995736766e-17), T(0), T(0), T(1)}, q);       // quater(pi, Z) * q
  }
  return axis_to_axis_core(va, vb);
}
// quater::rotationY(): q = rotY * offset (twist about Y of decomposeNoTwistTimesTwist, quater.cpp:891-929)
template <typename T> __device__ __noinline__ Q4<T> rotation_y(Q4<T> q) {
  const V3<T> Yv = {T(0), T(1), T(0)};
  Q4<T> nt = axis_to_axis(Yv, qrot(q, Yv));
  return quat_mul(qinv(nt), q);
}
template <typename T> __device__ __forceinline__ V3<T> rotation_vector(Q4<T> q) {               // 2 * quater::ln
  T sc = num<T>::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  T theta = num<T>::atan2(sc, q.w);
  sc = sc > T(1e-10) ? theta / sc : T(1);
  return {T(2) * (sc * q.x), T(2) * (sc * q.y), T(2) * (sc * q.z)};
}
template <typename T> __device__ __forceinline__ T rotation_angle(Q4<T> q) { return norm(rotation_vector(q)); }
template <typename T> __device__ __forceinline__ T rotation_angle_about_y(Q4<T> q) {
  V3<T> r = rotation_vector(rotation_y(q));
  T n = norm(r);
  return r.y < T(0) ? -n : n;
}
template <typename T> __device__ __forceinline__ Q4<T> qdifference(Q4<T> q1, Q4<T> q2) {        // delta * q1 = q2
  return qnormalize(quat_mul(qalign(q2, q1), qinv(q1)));
}
template <typename T> __device__ __forceinline__ Q4<T> qlin(Q4<T> a, T wa, Q4<T> b, T wb) {
  return {a.w * wa + b.w * wb, a.x * wa + b.x * wb, a.y * wa + b.y * wb, a.z * wa + b.z * wb};
}
// quater::safeSlerp (quater.cpp:217-272)
template <typename T> __device__ __forceinline__ Q4<T> safe_slerp(Q4<T> a, Q4<T> b, T t) {
  if (qdot(a, b) < T(0)) a = qneg(a);
  T c = qdot(a, b);
  if (T(1) + c > T(1e-10)) {
    if (T(1) - c > T(1e-10)) {
      T theta = num<T>::acos(c);
      T s0, s1, sn, dummy;
      num<T>::sincos((T(1) - t) * theta, &s0, &dummy);
      num<T>::sincos(t * theta, &s1, &dummy);
      num<T>::sincos(theta, &sn, &dummy);
      Q4<T> r = qlin(a, s0, b, s1);
      T inv = T(1) / sn;
      return {r.w * inv, r.x * inv, r.y * inv, r.z * inv};
    }
    return qnormalize(qlin(a, T(1) - t, b, t));
  }
  T s0, s1, dummy;
  num<T>::sincos((T(0.5) - t) * T(CDM_PI), &s0, &dummy);
  num<T>::sincos(t * T(CDM_PI), &s1, &dummy);
  return qlin(a, s0, b, s1);
}
// matrix4::setRotation(quater) rows, no normalisation
template <typename T> __device__ __forceinline__ void q2mat(Q4<T> q, T R[9]) {
  T w = q.w, x = q.x, y = q.y, z = q.z;
  R[0] = T(1) - T(2) * (y * y + z * z); R[1] = T(2) * (x * y - w * z); R[2] = T(2) * (x * z + w * y);
  R[3] = T(2) * (x * y + w * z); R[4] = T(1) - T(2) * (x * x + z * z); R[5] = T(2) * (y * z - w * x);
  R[6] = T(2) * (x * z - w * y); R[7] = T(2) * (y * z + w * x); R[8] = T(1) - T(2) * (x * x + y * y);
}
// transf:twist (module.lua:7531): unskew(T1^-1 (T2 - T1) / dt), body frame of T1
template <typename T> __device__ __forceinline__ void twist(Q4<T> q1, V3<T> p1, Q4<T> q2, V3<T> p2, V3<T>* w, V3<T>* v) {
  T R1[9], R2[9], D[9];
  q2mat(q1, R1); q2mat(q2, R2);
#pragma unroll
  for (int i = 0; i < 9; i++) D[i] = R2[i] - R1[i];
  // M = R1^T D ; M[i][j] = sum_k R1[k][i] D[k][j]
  T m12 = R1[1] * D[2] + R1[4] * D[5] + R1[7] * D[8];
  T m02 = R1[0] * D[2] + R1[3] * D[5] + R1[6] * D[8];
  T m01 = R1[0] * D[1] + R1[3] * D[4] + R1[6] * D[7];
  *w = {-m12, m02, -m01};
  V3<T> d = p2 - p1;
  *v = {R1[0] * d.x + R1[3] * d.y + R1[6] * d.z, R1[1] * d.x + R1[4] * d.y + R1[7] * d.z, R1[2] * d.x + R1[5] * d.y + R1[8] * d.z};
}
template <typename T> __device__ __forceinline__ T sop_map(T t, T lo, T hi, T v1, T v2) {
  t = (t - lo) / (hi - lo);
  return v1 * (T(1) - t) + v2 * t;
}
template <typename T> __device__ __forceinline__ T clamp_map(T t, T lo, T hi, T v1, T v2) {
  t = num<T>::fmin(num<T>::fmax(t, lo), hi);
  return sop_map(t, lo, hi, v1, v2);
}
template <typename T> __device__ __forceinline__ T sin1(T x) { T s, c; num<T>::sincos(x, &s, &c); return s; }
template <typename T> __device__ __forceinline__ T cos1(T x) { T s, c; num<T>::sincos(x, &s, &c); return c; }
template <typename T> __device__ __forceinline__ T map_sin(T i, T a, T b, T c, T d) { return sin1(sop_map(i, a, b, T(0), T(CDM_PI / 2))) * (d - c) + c; }
template <typename T> __device__ __forceinline__ T map_cos(T i, T a, T b, T c, T d) { return sop_map(cos1(sop_map(i, a, b, T(0), T(CDM_PI / 2))), T(1), T(0), c, d); }
// math.smoothClamp (module.lua:2876-2913)
template <typename T> __device__ __forceinline__ T smooth_clamp(T v, T vmax) { return v < T(CDM_PI / 2) * vmax ? sin1(v / vmax) * vmax : vmax; }
template <typename T> __device__ __forceinline__ V3<T> smooth_clamp_vec3(V3<T> v, T max_len) {
  T ln = norm(v);
  if (ln > T(0.001)) return (smooth_clamp(ln, max_len) / ln) * v;
  return v;
}
template <typename T> __device__ __forceinline__ T smooth_clamp_sym(T x, T mag) { return x > T(0) ? smooth_clamp(x, mag) : -smooth_clamp(-x, mag); }

// ---- reference-table sampling ------------------------------------------------------------------------------------
// matrixn::sampleRow (BaseLib/math/matrixn.cpp:781-805): the fractional part is rounded to float32, rows within
// 0.005 snap; MotionDOF::samplePose (MotionDOF.cpp:593-620) keeps the fraction in double.
template <typename T> struct RowMix { int i0, i1; T w0, w1; };
template <typename T> __device__ __forceinline__ RowMix<T> row_mix(T t, int rows, bool f32_fraction) {
  T fl = num<T>::floor(t);
  int a = (int)fl;
  T f = f32_fraction ? (T)(float)(t - (T)(float)fl) : t - fl;
  RowMix<T> m;
  if (f < T(0.005)) { m.i0 = m.i1 = min(max(a, 0), rows - 1); m.w0 = T(1); m.w1 = T(0); return m; }
  if (f > T(0.995)) { m.i0 = m.i1 = min(max(a + 1, 0), rows - 1); m.w0 = T(1); m.w1 = T(0); return m; }
  if (a < 0) { f -= T(1); m.i0 = a + 1; m.i1 = a + 2; }
  else if (a + 1 >= rows) { f += T(1); m.i0 = a - 1; m.i1 = a; }
  else { m.i0 = a; m.i1 = a + 1; }
  m.w0 = T(1) - f; m.w1 = f;
  m.i0 = min(max(m.i0, 0), rows - 1); m.i1 = min(max(m.i1, 0), rows - 1);   // only reachable with a NaN / inf state
  return m;
}
template <typename T, int NCOL> __device__ __forceinline__ void tab_mix(const T* __restrict__ tab, const RowMix<T>& m, int col, T out[NCOL]) {
  const T* r0 = tab + (size_t)m.i0 * TAB_W + col;
  const T* r1 = tab + (size_t)m.i1 * TAB_W + col;
#pragma unroll
  for (int k = 0; k < NCOL; k++) out[k] = __ldg(r0 + k) * m.w0 + __ldg(r1 + k) * m.w1;
}

// ------------------------------------------------------------------------------------------------------------------
// OBJloader::Terrain::height (work/taesooLib/BaseLib/motion/Terrain.cpp:305-393): cell by floor, triangle by dx > dz,
// plane through the triangle (Plane::setPlane, intersectionTest.cpp:96-104) intersected with the vertical ray from y = 1000
// (Ray::intersects :243-260), tiling along Z.  hit = (cell row i, cell column j, triangle 0 upper / 1 lower), -1 outside.
__device__ __forceinline__ double tl_terrain_height(const TLTerrainDev& T, double x0, double x1, int* hit) {
  const int nsx = T.cols - 1, nsz = T.rows - 1;
  const double nx = (x0 / T.size_x) * (double)nsx;
  double nz = (x1 / T.size_z) * (double)nsz;
  int j = (int)::floor(nx), i = (int)::floor(nz);
  const double dx = nx - ::floor(nx), dz = nz - ::floor(nz);
  if (hit) { hit[0] = -1; hit[1] = -1; hit[2] = -1; }
  if (j < 0 || j >= nsx || !(nx == nx) || !(nz == nz)) return 0.0;
  if (T.tile) {
    i %= nsz;
    if (i < 0) i += nsz;
    nz = (double)i + dz;
    x1 = nz / (double)nsz * T.size_z;
  } else if (i < 0 || i >= nsz) {
    return 0.0;
  }
  const double* h = T.h;
  V3<double> v0, v1, v2;
  const bool upper = dx > dz;
  v0 = {(double)j * T.size_x / (double)nsx, __ldg(h + (size_t)i * T.cols + j), (double)i * T.size_z / (double)nsz};
  if (upper) {
    v1 = {(double)(j + 1) * T.size_x / (double)nsx, __ldg(h + (size_t)(i + 1) * T.cols + j + 1), (double)(i + 1) * T.size_z / (double)nsz};
    v2 = {v1.x, __ldg(h + (size_t)i * T.cols + j + 1), v0.z};
  } else {
    v2 = {(double)(j + 1) * T.size_x / (double)nsx, __ldg(h + (size_t)(i + 1) * T.cols + j + 1), (double)(i + 1) * T.size_z / (double)nsz};
    v1 = {v0.x, __ldg(h + (size_t)(i + 1) * T.cols + j), v2.z};
  }
  V3<double> n = cross(v1 - v0, v2 - v0);
  n = (1.0 / ::sqrt(dot(n, n))) * n;
  const double d = -dot(n, v0);
  const V3<double> org = {x0, 1000.0, x1};
  const double denom = dot(n, V3<double>{0.0, -1.0, 0.0});
  const double t = -((dot(n, org) + d) / denom);
  if (hit) { hit[0] = i; hit[1] = j; hit[2] = upper ? 0 : 1; }
  return org.y + t * -1.0;
}

// Ground under a world point: RagdollSim:projectToGround / OBJloader.Terrain:getTerrainPos (deepmimiccdm-v1.lua:415-424,2171-2177).
// `Ground` with T.h == nullptr is flat ground (walk2-v1); otherwise the gym_cdm2 terrain (walkterrain-v1), metres <-> cm.
// One copy of the height query per kernel, CALLED from its 20 sites: inlined, the copies were 45 % of the SASS and every one of them was
// fetched once per step (terrain) or jumped over (flat ground inside the shared helpers); as a function the body stays in the instruction
// cache.  16384 envs, walkterrain-v1 policy: 209.8 -> 163.9 us per step, U(-1,1) actions 348.6 -> 256.2 us (profiles/r02_cdm_terrain_call_ab.txt).
__device__ __noinline__ double tl_terrain_height_call(const TLTerrainDev& T, double x0, double x1) { return tl_terrain_height(T, x0, x1, nullptr); }
template <typename real> __device__ __forceinline__ real ground_y(const Ground* g, real x, real z) {
  if (g == nullptr || g->T.h == nullptr) return real(0);
  return (real)(tl_terrain_height_call(g->T, ((double)x - g->px) * 100.0, ((double)z - g->pz) * 100.0) / 100.0 + g->py);
}

// ---- state in registers ------------------------------------------------------------------------------------------
template <typename real> struct Leg {
  V3<real> cpos, pre, bx, bv; real fq_x, fq_v; Q4<real> org; real swph;
  bool swing; bool is_left;
};
template <typename real> struct CdmState {
  V3<real> p; Q4<real> Q; V3<real> v, wb; real iframe, gtime; Q4<real> tgt; V3<real> prevcom; real ep_ret;
  Leg<real> leg[2];
  int flags, nrep, episode, steps;
};
template <typename real> __device__ __forceinline__ void load_state(CdmState<real>& s, const real* __restrict__ st, const int* __restrict__ ist, int Npad, int e) {
#define LD(f) st[(size_t)(f) * Npad + e]
  s.p = {LD(C_P), LD(C_P + 1), LD(C_P + 2)};
  s.Q = {LD(C_Q), LD(C_Q + 1), LD(C_Q + 2), LD(C_Q + 3)};
  s.v = {LD(C_V), LD(C_V + 1), LD(C_V + 2)};
  s.wb = {LD(C_WB), LD(C_WB + 1), LD(C_WB + 2)};
  s.iframe = LD(C_IFRAME); s.gtime = LD(C_GTIME);
  s.tgt = {LD(C_TGT), LD(C_TGT + 1), LD(C_TGT + 2), LD(C_TGT + 3)};
  s.prevcom = {LD(C_PREVCOM), LD(C_PREVCOM + 1), LD(C_PREVCOM + 2)};
  s.ep_ret = LD(C_EPRET);
#pragma unroll
  for (int l = 0; l < 2; l++) {
    const int b = C_LEG + l * LEG_W;
    Leg<real>& g = s.leg[l];
    g.cpos = {LD(b + L_CPOS), LD(b + L_CPOS + 1), LD(b + L_CPOS + 2)};
    g.pre = {LD(b + L_PRE), LD(b + L_PRE + 1), LD(b + L_PRE + 2)};
    g.bx = {LD(b + L_BX), LD(b + L_BX + 1), LD(b + L_BX + 2)};
    g.bv = {LD(b + L_BV), LD(b + L_BV + 1), LD(b + L_BV + 2)};
    g.fq_x = LD(b + L_FQ); g.fq_v = LD(b + L_FQ + 1);
    g.org = {LD(b + L_ORG), LD(b + L_ORG + 1), LD(b + L_ORG + 2), LD(b + L_ORG + 3)};
    g.swph = LD(b + L_SWPH);
    g.is_left = l == 1;
  }
#undef LD
  s.flags = ist[(size_t)I_FLAGS * Npad + e]; s.nrep = ist[(size_t)I_NREP * Npad + e];
  s.episode = ist[(size_t)I_EPISODE * Npad + e]; s.steps = ist[(size_t)I_STEPS * Npad + e];
  s.leg[0].swing = (s.flags & FLAG_SW_R) != 0; s.leg[1].swing = (s.flags & FLAG_SW_L) != 0;
}
template <typename real> __device__ __forceinline__ void store_state(CdmState<real>& s, real* __restrict__ st, int* __restrict__ ist, int Npad, int e) {
#define ST(f, v) st[(size_t)(f) * Npad + e] = (v)
  ST(C_P, s.p.x); ST(C_P + 1, s.p.y); ST(C_P + 2, s.p.z);
  ST(C_Q, s.Q.w); ST(C_Q + 1, s.Q.x); ST(C_Q + 2, s.Q.y); ST(C_Q + 3, s.Q.z);
  ST(C_V, s.v.x); ST(C_V + 1, s.v.y); ST(C_V + 2, s.v.z);
  ST(C_WB, s.wb.x); ST(C_WB + 1, s.wb.y); ST(C_WB + 2, s.wb.z);
  ST(C_IFRAME, s.iframe); ST(C_GTIME, s.gtime);
  ST(C_TGT, s.tgt.w); ST(C_TGT + 1, s.tgt.x); ST(C_TGT + 2, s.tgt.y); ST(C_TGT + 3, s.tgt.z);
  ST(C_PREVCOM, s.prevcom.x); ST(C_PREVCOM + 1, s.prevcom.y); ST(C_PREVCOM + 2, s.prevcom.z);
  ST(C_EPRET, s.ep_ret);
#pragma unroll
  for (int l = 0; l < 2; l++) {
    const int b = C_LEG + l * LEG_W;
    const Leg<real>& g = s.leg[l];
    ST(b + L_CPOS, g.cpos.x); ST(b + L_CPOS + 1, g.cpos.y); ST(b + L_CPOS + 2, g.cpos.z);
    ST(b + L_PRE, g.pre.x); ST(b + L_PRE + 1, g.pre.y); ST(b + L_PRE + 2, g.pre.z);
    ST(b + L_BX, g.bx.x); ST(b + L_BX + 1, g.bx.y); ST(b + L_BX + 2, g.bx.z);
    ST(b + L_BV, g.bv.x); ST(b + L_BV + 1, g.bv.y); ST(b + L_BV + 2, g.bv.z);
    ST(b + L_FQ, g.fq_x); ST(b + L_FQ + 1, g.fq_v);
    ST(b + L_ORG, g.org.w); ST(b + L_ORG + 1, g.org.x); ST(b + L_ORG + 2, g.org.y); ST(b + L_ORG + 3, g.org.z);
    ST(b + L_SWPH, g.swph);
  }
#undef ST
  s.flags = (s.leg[0].swing ? FLAG_SW_R : 0) | (s.leg[1].swing ? FLAG_SW_L : 0);
  ist[(size_t)I_FLAGS * Npad + e] = s.flags; ist[(size_t)I_NREP * Npad + e] = s.nrep;
  ist[(size_t)I_EPISODE * Npad + e] = s.episode; ist[(size_t)I_STEPS * Npad + e] = s.steps;
}

// ---- contact filter (deepmimiccdm-v1.lua:441-599; SDS / MassParticle3D: Resource/scripts/control/SDRE.lua:470-620) ----
// NonlinearController.oneStepRaw with the LQR gain K: the constant G = k/M is subtracted (SDRE.lua:134)
template <typename real> __device__ __forceinline__ void sds_step2(real& x, real& v, real xd, real vd, real K0, real K1, real dt) {
  const real M = real(30), b = real(0.001), k = real(1);
#pragma unroll
  for (int i = 0; i < 2; i++) {
    real cf = K0 * xd + K1 * vd - (K0 * x + K1 * v);
    real dd = cf * (real(1) / M) - k / M - (b / M) * v;
    v += dd * dt;
    x += v * dt;
  }
}
template <typename real> __device__ __forceinline__ void start_swing(Leg<real>& li) {          // :441-468
  li.swing = true;
  li.swph = real(0);
  li.pre = li.bx;                       // prevContact = contactFilter (= particle position)
  li.bv = {real(0), real(0), real(0)};
  li.fq_x = real(0); li.fq_v = real(0);
}
template <typename real> __device__ __forceinline__ void start_spprt(Leg<real>& li, Q4<real> rotY) {   // :470-497
  li.swing = false;
  li.org = rotY;
}
template <typename real>
__device__ __noinline__ void update_swing_foot(Leg<real>& li, V3<real> contact, Q4<real> Q, V3<real> p, Q4<real> rotY, V3<real> gv, const Params<real>& P) {
  // :602-640 leg-length limit, then filterContact mode 'sw' (:519-599)
  V3<real> local = qrot(qinv(Q), contact - p);
  real lim = P.max_leg_len + real(-0.01);
  real ln = norm(local);
  if (ln > lim) local = lim * vnormalize(local);
  li.cpos = qrot(Q, local) + p;
  real w = li.swing ? li.swph : P.filter_end_phase;
  V3<real> vel = li.cpos - li.pre;
  real alpha2 = map_sin(w, real(0), P.filter_end_phase, real(0), real(1));
  vel = smooth_clamp_vec3(vel, P.max_foot_vel * alpha2);
  li.pre = li.pre + vel;
  const real dt = real(1.0 / 60.0) / real(2);
  sds_step2(li.bx.x, li.bv.x, li.pre.x, gv.x, P.K0, P.K1, dt);
  sds_step2(li.bx.y, li.bv.y, li.pre.y, real(0), P.K0, P.K1, dt);
  sds_step2(li.bx.z, li.bv.z, li.pre.z, gv.z, P.K0, P.K1, dt);
  const Q4<real> QI = {real(1), real(0), real(0), real(0)};
  Q4<real> delta = qalign(qdifference(li.org, rotY), QI);
  real ang = rotation_angle_about_y(delta);
  sds_step2(li.fq_x, li.fq_v, ang, real(0), P.K0, P.K1, dt);
}

// integer / vector tables are indexed with the real-valued frame: (int) truncation (Q8)
template <typename real> __device__ __forceinline__ real tab_at(const real* __restrict__ tab, int F, real iframe, int col) {
  return __ldg(tab + (size_t)min(max((int)iframe, 0), F - 1) * TAB_W + col);
}

// getFilteredFootPos (:1865-1903)
template <typename real>
__device__ __noinline__ void filtered_foot_pos(const CdmState<real>& s, Q4<real> rotY, int l, const real* __restrict__ tab, int F, const Ground* gr, V3<real>* pos_out, Q4<real>* ori_out) {
  const Leg<real>& li = s.leg[l];
  const Leg<real>& oli = s.leg[1 - l];
  V3<real> pos = li.bx;
  Q4<real> ori = li.org;
  if (li.swing) {
    ori = quat_mul(li.org, qaxisY(li.fq_x));
    if (!oli.swing) {
      V3<real> side = qrot(rotY, V3<real>{real(1), real(0), real(0)});
      real depth = dot(pos - oli.bx, side);
      const real thr = real(0.05);
      real ph = clamp_map(tab_at(tab, F, s.iframe, T_RTTD + l), real(0), tab_at(tab, F, s.iframe, T_SWDUR + l), real(0), real(1));
      // math.pow(sin(phase*pi), 0.5); in fp32 sin(float(pi)) is -8.7e-8 where the fp64 reference has +1.2e-16
      real weight = num<real>::sqrt(num<real>::fmax(sin1(ph * real(CDM_PI)), real(0)));
      if (li.is_left) {
        if (depth < thr) pos = pos - ((depth - thr) * weight) * side;
      } else if (-depth < thr) {
        pos = pos + ((-depth - thr) * weight) * side;
      }
    }
  }
  pos.y = ground_y<real>(gr, pos.x, pos.z);
  *pos_out = pos; *ori_out = ori;
}

// model:getPhase (defaultRLsetting.lua:239-283); mutates iframe when it runs off the stitched table
template <typename real> __device__ __forceinline__ real get_phase(CdmState<real>& s, const real* __restrict__ tab, int F, int nf) {
  if (s.iframe > real(nf * 2)) s.iframe -= real(nf);
  if (s.iframe >= real(F - 1)) s.iframe -= real(nf);
  real iframe = s.iframe;
  int ifl = min(max((int)num<real>::floor(iframe), 0), F - 2);
  real t0 = __ldg(tab + (size_t)ifl * TAB_W + T_IFR), t1 = __ldg(tab + (size_t)(ifl + 1) * TAB_W + T_IFR);
  if (t1 - t0 < real(0)) {
    iframe = sop_map(iframe, real(ifl), real(ifl + 1), t0, t1 + real(nf));
    if (iframe > real(nf)) iframe -= real(nf);
  }
  RowMix<real> m = row_mix(iframe, F, true);
  real o[1];
  tab_mix<real, 1>(tab, m, T_IFR, o);
  return o[0] / real(nf);
}

// _get_obs (:1779-1859), flat ground
template <typename real, int G>
__device__ __noinline__ void write_obs(CdmState<real>& s, const real* __restrict__ tab, int F, int nf, float* __restrict__ o, int lane, bool active, bool force_zero,
                                       const Ground* gr) {
  const Q4<real> QI = {real(1), real(0), real(0), real(0)};
  Q4<real> rotY = rotation_y(s.Q);
  Q4<real> iry = qinv(rotY);
  V3<real> ww = qrot(s.Q, s.wb);
  real ob[OBS_W_TERRAIN];
  const bool terr = gr != nullptr && gr->T.h != nullptr;
  const real gy = ground_y<real>(gr, s.p.x, s.p.z);            // getRefCoord: the root projected to the ground
  ob[0] = s.p.y - gy;
  Q4<real> offq = qalign(quat_mul(iry, s.Q), QI);
  ob[1] = offq.w; ob[2] = offq.x; ob[3] = offq.y; ob[4] = offq.z;
  V3<real> a = qrot(iry, ww), b = qrot(iry, s.v);
  ob[5] = a.x; ob[6] = a.y; ob[7] = a.z; ob[8] = b.x; ob[9] = b.y; ob[10] = b.z;
  V3<real> ref_t = {s.p.x, gy, s.p.z};
#pragma unroll
  for (int l = 0; l < 2; l++) {
    const Leg<real>& li = s.leg[l];
    int c = 11 + 6 * l;
    V3<real> cpos = li.bx;
    if (terr) cpos.y = ground_y<real>(gr, cpos.x, cpos.z);     // deepmimiccdmterrain-v1.lua:58-60
    V3<real> lp = qrot(iry, cpos - ref_t);
    ob[c] = lp.x; ob[c + 1] = lp.y; ob[c + 2] = lp.z;
    V3<real> dv = li.swing ? li.bv : V3<real>{real(0), real(0), real(0)};
    V3<real> vv = qrot(iry, dv);
    ob[c + 3] = vv.x; ob[c + 4] = vv.z;
    ob[c + 5] = li.swing ? real(0) : rotation_angle_about_y(qalign(quat_mul(iry, li.org), QI));
  }
  real phase = get_phase(s, tab, F, nf);
  ob[23] = map_cos(phase, real(0), real(0.25), real(1), real(0));
  ob[24] = map_sin(phase, real(0), real(0.25), real(0), real(1));
  V3<real> tv = qrot(quat_mul(iry, s.tgt), V3<real>{real(0), real(0), real(1)});
  ob[25] = tv.x; ob[26] = tv.z;
#pragma unroll
  for (int k = 0; k < 4; k++) ob[OBS_W + k] = real(0);
  if (terr) {                                                  // terrain_sample_loc (deepmimiccdmterrain-v1.lua:3-8,95-99)
    const real sx[4] = {real(0), real(0.5), real(-0.5), real(1)};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      V3<real> pos = qrot(rotY, V3<real>{sx[k], real(0), real(0.5)}) + ref_t;
      ob[OBS_W + k] = ground_y<real>(gr, pos.x, pos.z) - ref_t.y;
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < OBS_W_TERRAIN; i++)
      if ((i % G) == lane && (i < OBS_W || terr)) o[i] = force_zero ? 0.f : (float)ob[i];
  }
}
template <typename real> __device__ __forceinline__ bool obs_has_nan(const CdmState<real>& s) {
  real acc = s.p.x + s.p.y + s.p.z + s.Q.w + s.Q.x + s.Q.y + s.Q.z + s.v.x + s.v.y + s.v.z + s.wb.x + s.wb.y + s.wb.z + s.iframe;
#pragma unroll
  for (int l = 0; l < 2; l++) acc += s.leg[l].bx.x + s.leg[l].bx.y + s.leg[l].bx.z + s.leg[l].bv.x + s.leg[l].bv.y + s.leg[l].bv.z + s.leg[l].org.w + s.leg[l].org.y;
  return !(acc == acc) || !(num<real>::abs(acc) < real(1e30));
}

// ---- reference markers: FK of the two leg chains (getJointStateFromRefTree :996-1012, BoneForwardKinematics) -------
template <typename real>
__device__ __noinline__ void ref_markers(const real* __restrict__ tab, int F, real f, const Params<real>& P, V3<real> mk[4], V3<real>* cdm_pos, Q4<real>* cdm_q,
                                         const Ground* gr = nullptr) {
  RowMix<real> m = row_mix(f, F, true);
  real c[7];
  tab_mix<real, 7>(tab, m, T_CDM, c);
  *cdm_pos = {c[0], c[1], c[2]};
  *cdm_q = qnormalize(Q4<real>{c[3], c[4], c[5], c[6]});
  if (mk == nullptr) return;
  real r[21];
  tab_mix<real, 21>(tab, m, T_LEG, r);
  Q4<real> rq = qnormalize(Q4<real>{r[3], r[4], r[5], r[6]});
  V3<real> rt = {r[0] + P.root_off[0], r[1] + P.root_off[1], r[2] + P.root_off[2]};
  int col = 7;
#pragma unroll
  for (int side = 0; side < 2; side++) {                // table order: left chain, then right chain
    Q4<real> q = rq;
    V3<real> t = rt;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int jj = side * 3 + j;
      t = qrot(q, V3<real>{P.leg_off[jj][0], P.leg_off[jj][1], P.leg_off[jj][2]}) + t;
      Q4<real> lq = {real(1), real(0), real(0), real(0)};
      for (int ch = 0; ch < P.leg_nchan[jj]; ch++) {
        int ax = (P.leg_axes[jj] >> (2 * ch)) & 3;
        real sn, cs;
        num<real>::sincos(real(0.5) * r[col], &sn, &cs);
        col++;
        Q4<real> qa = {cs, ax == 0 ? sn : real(0), ax == 1 ? sn : real(0), ax == 2 ? sn : real(0)};
        lq = quat_mul(lq, qa);
      }
      q = quat_mul(q, lq);
    }
    V3<real> toe = qrot(q, V3<real>{P.toe[0], P.toe[1], P.toe[2]}) + t;
    V3<real> heel = qrot(q, V3<real>{P.heel[0], P.heel[1], P.heel[2]}) + t;
    toe.y = ground_y<real>(gr, toe.x, toe.z); heel.y = ground_y<real>(gr, heel.x, heel.z);       // projectToGround
    const int o = side == 0 ? 2 : 0;                    // marker order: R toe, R heel, L toe, L heel
    mk[o] = toe; mk[o + 1] = heel;
  }
}

// ---- reset (RagdollSim:reset :1578-1722) -------------------------------------------------------------------------
template <typename real> __device__ inline void draw_reset_plan(unsigned long long seed, int env, int episode, const Params<real>& P, double plan[PLAN_W]) {
  double u[16];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    uint32_t c0[4] = {(uint32_t)env, (uint32_t)episode, (uint32_t)(2 * b), 0xCD3u}, c1[4] = {(uint32_t)env, (uint32_t)episode, (uint32_t)(2 * b + 1), 0xCD3u};
    srb::philox4x32(c0, (uint32_t)seed, (uint32_t)(seed >> 32));
    srb::philox4x32(c1, (uint32_t)seed, (uint32_t)(seed >> 32));
    u[4 * b] = srb::u01(c0[0], c0[1]); u[4 * b + 1] = srb::u01(c0[2], c0[3]);
    u[4 * b + 2] = srb::u01(c1[0], c1[1]); u[4 * b + 3] = srb::u01(c1[2], c1[3]);
  }
  for (int i = 0; i < 4; i++) plan[i] = (2.0 * u[i] - 1.0) * (double)P.noise_q;           // CT.randomUniform{-s, s}
  for (int i = 0; i < 3; i++) plan[4 + i] = (2.0 * u[4 + i] - 1.0) * (double)P.noise_pos;
  for (int i = 0; i < 6; i++) plan[7 + i] = (2.0 * u[7 + i] - 1.0) * (double)P.noise_dq;
  for (int i = 0; i < 3; i++) plan[13 + i] = (u[13 + i] - 0.5) * (double)P.noise_v[i];       // (math.random()-0.5)*scale
}
template <typename real>
__device__ inline void apply_reset(CdmState<real>& s, const Params<real>& P, const ResetConst<real>& rc, const double plan[PLAN_W], real hist0[HIST_W],
                                   const Ground* gr = nullptr) {
  s.tgt = {real(1), real(0), real(0), real(0)};
  s.gtime = real(0);
  s.iframe = real(0);
  s.ep_ret = real(0);
  s.steps = 0;
  s.episode += 1;
  V3<real> pos = {real(0) + (real)plan[4], rc.cdm0[1] + ground_y<real>(gr, real(0), real(0)) + (real)plan[5], real(0) + (real)plan[6]};   // :1595-1596
  Q4<real> q = {rc.cdm0[3] + (real)plan[0], rc.cdm0[4] + (real)plan[1], rc.cdm0[5] + (real)plan[2], rc.cdm0[6] + (real)plan[3]};
  q = qnormalize(q);
  s.p = pos; s.Q = q;
  // MotionDOF.dposeToDQ: [R w_body, R v_body] + noise ; setDQ keeps the angular part in the body frame
  V3<real> wv = qrot(q, V3<real>{rc.dcdm0[4], rc.dcdm0[5], rc.dcdm0[6]});
  V3<real> lv = qrot(q, V3<real>{rc.dcdm0[0], rc.dcdm0[1], rc.dcdm0[2]});
  wv = wv + V3<real>{(real)plan[7], (real)plan[8], (real)plan[9]};
  lv = lv + V3<real>{(real)plan[10], (real)plan[11], (real)plan[12]};
  lv = lv + V3<real>{(real)plan[13], (real)plan[14], (real)plan[15]};
  s.v = lv;
  s.wb = qrot(qinv(q), wv);
  hist0[0] = pos.x; hist0[1] = pos.y; hist0[2] = pos.z; hist0[3] = q.w; hist0[4] = q.x; hist0[5] = q.y; hist0[6] = q.z; hist0[7] = real(0);
  s.nrep = 1;
  Q4<real> rotY = rotation_y(q);
#pragma unroll
  for (int l = 0; l < 2; l++) {
    Leg<real>& li = s.leg[l];
    li.is_left = l == 1;
    li.cpos = {rc.contact0[l][0], rc.contact0[l][1], rc.contact0[l][2]};
    li.pre = li.cpos; li.bx = li.cpos;
    li.bv = {real(0), real(0), real(0)};
    li.fq_x = real(0); li.fq_v = real(0);
    li.org = rotY;
    li.swph = real(0);
    li.swing = false;
    if (rc.start_swing[l]) start_swing(li); else start_spprt(li, rotY);
  }
  s.prevcom = s.p;
}

// ------------------------------------------------------------------------------------------------------------------
// Structured contact QP for ONE thread (lanes_per_env = 1).  The 20 columns are never stored: with the contact arms
// r_c = R^T (c - p) and the friction directions n_d = R^T basis_d in the body frame,
//     a_{c,d} = D [r_c x n_d ; n_d],  D = diag(1/I, 1/m),     t_{c,d} = a_{c,d} . nu = n_d . (V + W x r_c),  W = D_w nu_w, V = D_v nu_v,
//     sum_{d in S_c} a a^T = D [[ [r]x N [r]x^T, [r]x N ], [ N [r]x^T, N ]] D,   N = sum_{d in S_c} n_d n_d^T,
// so a Newton system costs ~460 flops and the line search evaluates t on the fly (the active set travels as a 20-bit
// mask).  Same algorithm as srb::QP::solve_ls (exact line search on the dual), fp64.
struct SQP {
  double r[4][3], n[5][3], Di[6];
  unsigned valid;                                              // bit c: contact c present

  // t(nu) for all columns; returns the mask of t < 0 (invalid contacts contribute no bits); optionally sum_{t<0} t * t2 etc.
  __device__ __forceinline__ unsigned tmask(const double nu[6]) const {
    const double W[3] = {nu[0] * Di[0], nu[1] * Di[1], nu[2] * Di[2]}, V[3] = {nu[3] * Di[3], nu[4] * Di[4], nu[5] * Di[5]};
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const double u0 = V[0] + W[1] * r[c][2] - W[2] * r[c][1], u1 = V[1] + W[2] * r[c][0] - W[0] * r[c][2], u2 = V[2] + W[0] * r[c][1] - W[1] * r[c][0];
#pragma unroll
      for (int d = 0; d < 5; d++) {
        const double t = n[d][0] * u0 + n[d][1] * u1 + n[d][2] * u2;
        if (t < 0 && ((valid >> c) & 1)) m |= 1u << (5 * c + d);
      }
    }
    return m;
  }
  // H (lower triangle, 21) = I + invw * sum_{j in mask} a_j a_j^T
  __device__ __forceinline__ void hessian(unsigned mask, double invw, double H[21]) const {
    double TT[6] = {0, 0, 0, 0, 0, 0};      // xx yy zz xy xz yz of the torque block
    double TF[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // torque row i, force column j
    double FF[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; c++) {
      double N[6] = {0, 0, 0, 0, 0, 0};     // xx yy zz xy xz yz
#pragma unroll
      for (int d = 0; d < 5; d++) {
        const double m = ((mask >> (5 * c + d)) & 1) ? 1.0 : 0.0;
        const double a0 = m * n[d][0], a1 = m * n[d][1], a2 = m * n[d][2];
        N[0] += a0 * n[d][0]; N[1] += a1 * n[d][1]; N[2] += a2 * n[d][2]; N[3] += a0 * n[d][1]; N[4] += a0 * n[d][2]; N[5] += a1 * n[d][2];
      }
      const double rx = r[c][0], ry = r[c][1], rz = r[c][2];
      // X = [r]x N : column j of X = r x (column j of N)
      const double Nc[3][3] = {{N[0], N[3], N[4]}, {N[3], N[1], N[5]}, {N[4], N[5], N[2]}};   // Nc[j] = column j (symmetric)
      double X[3][3];                       // X[i][j]
#pragma unroll
      for (int j = 0; j < 3; j++) {
        X[0][j] = ry * Nc[j][2] - rz * Nc[j][1];
        X[1][j] = rz * Nc[j][0] - rx * Nc[j][2];
        X[2][j] = rx * Nc[j][1] - ry * Nc[j][0];
      }
      // Y = X [r]x^T : row i of Y = r x (row i of X)
      double Y[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        Y[i][0] = ry * X[i][2] - rz * X[i][1];
        Y[i][1] = rz * X[i][0] - rx * X[i][2];
        Y[i][2] = rx * X[i][1] - ry * X[i][0];
      }
      TT[0] += Y[0][0]; TT[1] += Y[1][1]; TT[2] += Y[2][2]; TT[3] += Y[1][0]; TT[4] += Y[2][0]; TT[5] += Y[2][1];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) TF[3 * i + j] += X[i][j];
#pragma unroll
      for (int k = 0; k < 6; k++) FF[k] += N[k];
    }
    // lower triangle, index i*(i+1)/2+j, rows 0..2 torque, 3..5 force; scaled by D on both sides
    const double d0 = Di[0], d1 = Di[1], d2 = Di[2], d3 = Di[3], d4 = Di[4], d5 = Di[5];
    H[0] = 1.0 + invw * d0 * d0 * TT[0];
    H[1] = invw * d1 * d0 * TT[3]; H[2] = 1.0 + invw * d1 * d1 * TT[1];
    H[3] = invw * d2 * d0 * TT[4]; H[4] = invw * d2 * d1 * TT[5]; H[5] = 1.0 + invw * d2 * d2 * TT[2];
    // force rows: H[3+i][j] = (TF^T)[i][j] = TF[j][i]  (torque column j)
    H[6] = invw * d3 * d0 * TF[0]; H[7] = invw * d3 * d1 * TF[3]; H[8] = invw * d3 * d2 * TF[6]; H[9] = 1.0 + invw * d3 * d3 * FF[0];
    H[10] = invw * d4 * d0 * TF[1]; H[11] = invw * d4 * d1 * TF[4]; H[12] = invw * d4 * d2 * TF[7]; H[13] = invw * d4 * d3 * FF[3]; H[14] = 1.0 + invw * d4 * d4 * FF[1];
    H[15] = invw * d5 * d0 * TF[2]; H[16] = invw * d5 * d1 * TF[5]; H[17] = invw * d5 * d2 * TF[8]; H[18] = invw * d5 * d3 * FF[4]; H[19] = invw * d5 * d4 * FF[5];
    H[20] = 1.0 + invw * d5 * d5 * FF[2];
  }
  // returns the number of Newton systems; nu in: initial guess, out: solution
  // ---- column values kept in shared memory (one slot per lane: element j of a lane at [j * 32]) ----------------------------------
  // T0_j = t_j(nu) and Sd_j = t_j(direction) make every evaluation along the search line r_j(alpha) = T0_j + alpha Sd_j one FMA per
  // column, so the columns (3 + 3 FMAs each plus the contact-point velocities) are formed ONCE per Newton system instead of once per
  // probe: t(nn) doubles as line(1), line(0) and the probes read the two arrays, and T0 += alpha Sd replaces the re-evaluation of
  // t(nu) after the step.  (Registers cannot hold the 40 values: that variant spilled and lost)
  // T0_j = t_j(v); returns the mask of t < 0.  Columns of absent contacts are stored as +1 (never active).
  __device__ __forceinline__ unsigned store_t(const double v[6], double* __restrict__ T0) const {
    const double W[3] = {v[0] * Di[0], v[1] * Di[1], v[2] * Di[2]}, V[3] = {v[3] * Di[3], v[4] * Di[4], v[5] * Di[5]};
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const double u0 = V[0] + W[1] * r[c][2] - W[2] * r[c][1], u1 = V[1] + W[2] * r[c][0] - W[0] * r[c][2], u2 = V[2] + W[0] * r[c][1] - W[1] * r[c][0];
      const bool vc = (valid >> c) & 1;
#pragma unroll
      for (int d = 0; d < 5; d++) {
        const double t = n[d][0] * u0 + n[d][1] * u1 + n[d][2] * u2;
        T0[(5 * c + d) * 32] = vc ? t : 1.0;
        if (t < 0 && vc) m |= 1u << (5 * c + d);
      }
    }
    return m;
  }
  // Sd_j = t_j(v) - T0_j; returns the mask of t_j(v) < 0
  __device__ __forceinline__ unsigned store_s(const double v[6], const double* __restrict__ T0, double* __restrict__ Sd) const {
    const double W[3] = {v[0] * Di[0], v[1] * Di[1], v[2] * Di[2]}, V[3] = {v[3] * Di[3], v[4] * Di[4], v[5] * Di[5]};
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const double u0 = V[0] + W[1] * r[c][2] - W[2] * r[c][1], u1 = V[1] + W[2] * r[c][0] - W[0] * r[c][2], u2 = V[2] + W[0] * r[c][1] - W[1] * r[c][0];
      const bool vc = (valid >> c) & 1;
#pragma unroll
      for (int d = 0; d < 5; d++) {
        const double t = n[d][0] * u0 + n[d][1] * u1 + n[d][2] * u2;
        Sd[(5 * c + d) * 32] = vc ? t - T0[(5 * c + d) * 32] : 0.0;
        if (t < 0 && vc) m |= 1u << (5 * c + d);
      }
    }
    return m;
  }
  // along the line: g-sum = sum_{r_j<0} r_j s_j, slope-sum = sum_{r_j<0} s_j^2, mask of r_j < 0, r_j = T0_j + alpha Sd_j
  static __device__ __forceinline__ unsigned line_s(const double* __restrict__ T0, const double* __restrict__ Sd, double alpha, double* gsum, double* ssum) {
    unsigned m = 0;
    double g = 0, sl = 0;
#pragma unroll
    for (int j = 0; j < 20; j++) {
      const double sd = Sd[j * 32];
      const double rr = fma(alpha, sd, T0[j * 32]);
      if (rr < 0) { m |= 1u << j; g += rr * sd; sl += sd * sd; }
    }
    *gsum = g; *ssum = sl;
    return m;
  }
  // T0 += alpha Sd (the new nu's column values); returns their mask of t < 0
  static __device__ __forceinline__ unsigned advance(double* __restrict__ T0, const double* __restrict__ Sd, double alpha) {
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 20; j++) {
      const double t = fma(alpha, Sd[j * 32], T0[j * 32]);
      T0[j * 32] = t;
      if (t < 0) m |= 1u << j;
    }
    return m;
  }

  // T0, Sd: this lane's slots of two [20][32] shared-memory arrays
  __device__ __forceinline__ int solve_ls(const double y[6], double invw, double nu[6], double* __restrict__ T0, double* __restrict__ Sd) const {
    unsigned S = store_t(nu, T0);
    int iters = 0;
    double rhs[6];
#pragma unroll
    for (int i = 0; i < 6; i++) rhs[i] = -y[i];
#pragma unroll 1
    for (int it = 0; it < 32; it++) {
      double H[21], nn[6];
      hessian(S, invw, H);
      srb::ldl_solve6<double>(H, rhs, nn);
      iters++;
      const unsigned Sn = store_s(nn, T0, Sd);
      if (Sn == S) {
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = nn[i];
        break;
      }
      double dir[6], dd = 0, b0 = 0, nmax = 0;
#pragma unroll
      for (int i = 0; i < 6; i++) {
        dir[i] = nn[i] - nu[i];
        dd += dir[i] * dir[i]; b0 += dir[i] * (nu[i] + y[i]);
        nmax = fmax(nmax, fabs(nn[i]));
      }
      double gs, ss;
      line_s(T0, Sd, 1.0, &gs, &ss);
      double alpha = 1.0;
      if (b0 + dd + invw * gs > 0) {                           // the full step overshoots: exact line search in (0, 1)
        unsigned ml = line_s(T0, Sd, 0.0, &gs, &ss);
        double gl = b0 + invw * gs, sl = dd + invw * ss, al = 0.0, lo = 0.0, hi = 1.0;
        if (!(gl < 0)) alpha = 0.0;
        else {
#pragma unroll 1
          for (int k2 = 0; k2 < 48; k2++) {
            double cand = al - gl / sl;
            const bool newton = cand > lo && cand < hi;
            if (!newton) cand = 0.5 * (lo + hi);
            const unsigned mc = line_s(T0, Sd, cand, &gs, &ss);
            const double gc = b0 + cand * dd + invw * gs, sc = dd + invw * ss;
            al = cand;
            if (newton && mc == ml) break;
            if (gc < 0) lo = cand; else hi = cand;
            gl = gc; sl = sc; ml = mc;
            if (hi - lo <= 4e-13 * hi) break;
          }
          alpha = al;
        }
      }
      double stepn = 0;
#pragma unroll
      for (int i = 0; i < 6; i++) { const double dlt = alpha * dir[i]; nu[i] += dlt; stepn = fmax(stepn, fabs(dlt)); }
      if (stepn <= 1e-13 * (1.0 + nmax)) break;
      S = advance(T0, Sd, alpha);
    }
    return iters;
  }
};

// ------------------------------------------------------------------------------------------------------------------
// the step kernel (RagdollSim:step :1014-1574)
// TERR: a terrain is attached (cdm_set_terrain).  The flat instantiation carries no terrain code at all: `gr` is a compile-time null and
// the 20 inlined height queries (45 % of the SASS of the one-kernel-for-both build) fold away.
template <typename real, int G, bool TERR>
__global__ void __launch_bounds__(32) cdm_step_kernel(const StepArgs<real> A) {
  constexpr int EPW = 32 / G;
  using QPT = srb::QP<double, G, NQCOL>;
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 % G;
  // Partially filled warps (experiment switch, cdm_set_param "envs_per_warp"): with gpw < EPW only the first gpw lane groups of a
  // warp carry an environment; the others shadow the warp's first environment (same branch decisions, no stores).
  const int grp = lane32 / G;
  const bool in_use = grp < A.gpw;
  const int e_raw = blockIdx.x * A.gpw + (in_use ? grp : 0);
  const bool active = in_use && e_raw < A.N;
  const int e = e_raw < A.N ? e_raw : A.N - 1;
  const Params<real>& P = A.p;
  const real* __restrict__ tab = A.tab;
  const int F = A.nframes, nf = A.nf_cycle;
  __shared__ __align__(16) double hsm[srb::HReduce<double, G>::SMEM_WORDS];
  __shared__ double sq_t0[G == 1 ? 20 * 32 : 1], sq_sd[G == 1 ? 20 * 32 : 1];      // SQP column values along the search line (one slot per lane)

  CdmState<real> s;
  load_state(s, A.st, A.ist, A.Npad, e);
  real* hist_env = A.hist + (size_t)e * HIST_SLOTS * HIST_W;
  real act[ACT_W];
#pragma unroll
  for (int i = 0; i < ACT_W; i++) act[i] = (real)__ldg(A.act + (size_t)e * ACT_W + i) * real(0.2);    // model.actionScaling
  {                                                            // step(): a NaN action is replaced by math.random()-0.5 (:858-863)
    bool anynan = false;
#pragma unroll
    for (int i = 0; i < ACT_W; i++) anynan = anynan || !(act[i] == act[i]);
    if (anynan) {
#pragma unroll
      for (int b = 0; b < 3; b++) {
        uint32_t c[4] = {(uint32_t)e, (uint32_t)s.steps, (uint32_t)(s.episode * 4 + b), 0xAC7u};
        srb::philox4x32(c, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (4 * b + j < ACT_W) act[4 * b + j] = ((real)((double)c[j] * (1.0 / 4294967296.0)) - real(0.5)) * real(0.2);
      }
    }
  }
  double* dbg = (A.dbg != nullptr && active && lane == 0) ? A.dbg + (size_t)e * DBG_W : nullptr;
  float* obs = A.obs + (size_t)e * A.obs_w;
  const Ground* gr = TERR ? &A.ground : nullptr;
  const real RL_STEP = real(1.0 / 60.0), FRAME_RATE = real(30);
  const Q4<real> QI = {real(1), real(0), real(0), real(0)};
  const V3<real> Zv = {real(0), real(0), real(1)};

  real reward = real(0);
  bool done = false;
  if (s.iframe >= real(F)) {                                   // :1017-1025 ran past the reference table
    done = true;
    write_obs<real, G>(s, tab, F, nf, obs, lane, active, false, gr);
  } else {
    const Q4<real> Q0 = s.Q;
    const V3<real> p0 = s.p;
    const Q4<real> rotY = rotation_y(Q0);
    const V3<real> gv = s.v;
    // ---- getDesiredFinalVel with target_Y_UI = identity (:1907-1936)
    {
      real vfront = num<real>::fmax(real(0), dot(gv, qrot(rotY, Zv)));
      V3<real> front = qrot(rotY, Zv);
      real ang = real(0);
      if (vfront != real(0)) {
        ang = num<real>::atan2(real(0), vfront) - num<real>::atan2(front.x, front.z);
        real sa, ca;
        num<real>::sincos(ang, &sa, &ca);
        ang = num<real>::atan2(sa, ca);
      }
      const real kt[5] = {real(0), real(1), real(2.5), real(5.5), real(10.5)};
      const real ka[5] = {real(CDM_PI), real(CDM_PI), real(120.0 * CDM_PI / 180.0), real(50.0 * CDM_PI / 180.0), real(30.0 * CDM_PI / 180.0)};
      real mx = ka[4];
#pragma unroll
      for (int i = 3; i >= 0; i--)
        if (kt[i + 1] >= vfront - real(1e-10)) mx = sop_map(sop_map(vfront, kt[i], kt[i + 1], real(0), real(1)), real(0), real(1), ka[i], ka[i + 1]);
      mx *= P.max_turning;
      ang = num<real>::fmin(num<real>::fmax(ang, -mx), mx);
      s.tgt = quat_mul(rotY, qaxisY(ang));
    }
    // ---- per leg: calcContactPos (:602-754), then the filtered contact points (:1077-1093)
    V3<real> cpt[2][2];
    real early[2] = {real(-1), real(-1)};
    bool has_early[2] = {false, false};
    bool late = false;                                         // lateTouchDown (:702-709): only with a leg-length limit (run2-v1)
    real ref_root_y = real(0);
    if (!P.unlimited_leglen) {                                 // info.loader_ref:bone(1) height = motionDOF_cdm:sample(iframe) (setRefTree, :1043)
      const RowMix<real> m0 = row_mix(s.iframe, F, false);
      ref_root_y = m0.i0 == m0.i1 ? __ldg(tab + (size_t)m0.i0 * TAB_W + T_CDM + 1)
                                  : m0.w0 * __ldg(tab + (size_t)m0.i0 * TAB_W + T_CDM + 1) + m0.w1 * __ldg(tab + (size_t)m0.i1 * TAB_W + T_CDM + 1);
    }
#pragma unroll
    for (int l = 0; l < 2; l++) {
      Leg<real>& li = s.leg[l];
      const bool was_swing = li.swing;
      if (!was_swing) {
        if ((int)tab_at(tab, F, s.iframe, T_RTTO + l) <= 0) {
          start_swing(li);
          update_swing_foot(li, li.cpos, Q0, p0, rotY, gv, P);
        }
      } else {
        const int oc = l == 0 ? T_OFFR : T_OFFL;
        V3<real> off = {tab_at(tab, F, s.iframe, oc), tab_at(tab, F, s.iframe, oc + 1), tab_at(tab, F, s.iframe, oc + 2)};
        V3<real> desired = p0 + qrot(rotY, V3<real>{act[2 * l], real(0), act[2 * l + 1] + real(0)} + off);
        desired.y = ground_y<real>(gr, desired.x, desired.z);     // projectToGround
        li.swph += RL_STEP;
        update_swing_foot(li, desired, Q0, p0, rotY, gv, P);
        const int rt = (int)tab_at(tab, F, s.iframe, T_RTTD + l);
        real com_height = p0.y;
        if (gr != nullptr) {                             // :686-692 height relative to the terrain between body and landing spot
          const real ph = clamp_map(tab_at(tab, F, s.iframe, T_RTTD + l), real(0), tab_at(tab, F, s.iframe, T_SWDUR + l), real(0), real(1));
          com_height -= map_cos(ph, real(0), real(1), ground_y<real>(gr, p0.x, p0.z), ground_y<real>(gr, li.cpos.x, li.cpos.z));
        }
        if (rt <= 0) {
          if (P.unlimited_leglen || com_height < ref_root_y + P.contact_thr) {
            li.cpos.y = ground_y<real>(gr, li.cpos.x, li.cpos.z);     // projectToGround
            start_spprt(li, rotY);
          } else {
            late = true;
          }
        } else if (com_height < P.touch_down_height && li.swph > P.min_swing_phase) {     // touchDownHeight: 0 in training
          has_early[l] = true; early[l] = real(rt);
        } else {
          const Leg<real>& oli = s.leg[1 - l];
          V3<real> a = li.bx, b = oli.bx;
          a.y = real(0); b.y = real(0);
          V3<real> d = a - (b + qrot(rotY, V3<real>{real(0), real(0), P.stride_thr}));
          if (P.stride_thr >= real(0) && !oli.swing && norm(d) > P.stride_thr && li.swph > P.min_swing_phase) { has_early[l] = true; early[l] = real(rt); }
        }
      }
      V3<real> pos; Q4<real> ori;
      filtered_foot_pos(s, rotY, l, tab, F, gr, &pos, &ori);
      cpt[l][0] = pos + qrot(ori, V3<real>{real(0), real(0), real(0.12)});
      cpt[l][1] = pos + qrot(ori, V3<real>{real(0), real(0), real(-0.12)});
    }
    // ---- late touch-down: the reference frame does not advance this step; else early touch-down -> advanceTime (:1150-1200)
    if (late) {
      s.iframe += -(FRAME_RATE * RL_STEP);
    } else {
      const real mspc = FRAME_RATE * RL_STEP;
      if (has_early[0] && has_early[1]) { if (early[0] < early[1]) has_early[1] = false; else has_early[0] = false; }
      bool have = false;
      real etd = real(0);
#pragma unroll
      for (int l = 0; l < 2; l++) {
        if (!have && has_early[l]) {
          real ev = early[l];
          if (ev > mspc) ev = mspc;
          else {                                                           // enforceTouchDown (:767-774)
            start_spprt(s.leg[l], rotY);
            s.leg[l].cpos.y = ground_y<real>(gr, s.leg[l].cpos.x, s.leg[l].cpos.z);
          }
          etd = ev; have = true;
        }
      }
      if (have) s.iframe += etd;
    }
    // ---- sampleTargetPose (RagdollSim.lua:145-176): desired CDM pose aligned to the simulated past
    Q4<real> thq; V3<real> thp; V3<real> dth_v, dth_w;
    {
      RowMix<real> mp = row_mix(s.iframe, F, false);
      real r0[7], r1[7];
#pragma unroll
      for (int k = 0; k < 7; k++) { r0[k] = __ldg(tab + (size_t)mp.i0 * TAB_W + T_CDM + k); r1[k] = __ldg(tab + (size_t)mp.i1 * TAB_W + T_CDM + k); }
      if (mp.i0 == mp.i1) {
        thp = {r0[0], r0[1], r0[2]}; thq = {r0[3], r0[4], r0[5], r0[6]};
      } else {
        thp = {mp.w0 * r0[0] + mp.w1 * r1[0], mp.w0 * r0[1] + mp.w1 * r1[1], mp.w0 * r0[2] + mp.w1 * r1[2]};
        thq = safe_slerp(Q4<real>{r0[3], r0[4], r0[5], r0[6]}, Q4<real>{r1[3], r1[4], r1[5], r1[6]}, mp.w1);
      }
      RowMix<real> md = row_mix(s.iframe, F, true);
      real dv[6];
      tab_mix<real, 6>(tab, md, T_DV, dv);
      dth_v = {dv[0], dv[1], dv[2]}; dth_w = {dv[3], dv[4], dv[5]};
      const real* hs = hist_env + (size_t)(max(0, s.nrep - 5) & (HIST_SLOTS - 1)) * HIST_W;     // sampleSimPose(-2)
      Q4<real> Tq = rotation_y(Q4<real>{hs[3], hs[4], hs[5], hs[6]});
      V3<real> Tp = {hs[0], hs[1], hs[2]};
      RowMix<real> mr = row_mix(num<real>::fmax(s.iframe + real(-2), real(0)), F, true);
      real rr[7];
      tab_mix<real, 7>(tab, mr, T_CDM, rr);
      Q4<real> Rq = rotation_y(Q4<real>{rr[3], rr[4], rr[5], rr[6]});
      V3<real> Rp = {rr[0], rr[1], rr[2]};
      Q4<real> dq_ = quat_mul(Rq, qinv(Tq));
      V3<real> dp_ = Rp - qrot(dq_, Tp);
      dp_.y = real(0);
      dq_ = rotation_y(dq_);
      Q4<real> iq = qinv(dq_);
      thp = qrot(iq, thp - dp_);
      thq = quat_mul(iq, thq);
    }
    real terrain_y = real(0);
    if (gr != nullptr) {                                       // :1205-1238 targets follow the terrain
      terrain_y = ground_y<real>(gr, thp.x, thp.z);
      thp.y += terrain_y;
      const real nxt = ground_y<real>(gr, p0.x + gv.x * real(0.5), p0.z + gv.z * real(0.5));
      real dy = (nxt - ground_y<real>(gr, p0.x, p0.z)) / real(0.5);
      if (dy < real(0)) {
        V3<real> v_d = qrot(Q0, dth_v);
        v_d.y += num<real>::fmax(dy, real(-5));
        dth_v = qrot(qinv(Q0), v_d);
      }
    }
    // ---- contact list (:1242-1261)
    int cmask = 0;
#pragma unroll
    for (int l = 0; l < 2; l++) {
      if (!s.leg[l].swing) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
          cpt[l][j].y = ground_y<real>(gr, cpt[l][j].x, cpt[l][j].z);
          if (norm(cpt[l][j] - p0) < P.max_contact_len) cmask |= 1 << (2 * l + j);
        }
      }
    }
    if (dbg != nullptr) {
      dbg[68] = (double)thp.x; dbg[69] = (double)thp.y; dbg[70] = (double)thp.z;
      dbg[71] = (double)thq.w; dbg[72] = (double)thq.x; dbg[73] = (double)thq.y; dbg[74] = (double)thq.z;
      dbg[77] = (double)(2 * __popc(cmask));
    }
    // ---- QPsim:frameMove, niter = 2 (QPservo_v3.lua:1720-1866) on flat ground
    const double Md[6] = {(double)P.inertia[0], (double)P.inertia[1], (double)P.inertia[2], (double)P.mass, (double)P.mass, (double)P.mass};
    const real h = real(1.0 / 120.0);
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
      dth_v = dth_v + V3<real>{act[4], act[5], act[6]};
      dth_w = dth_w + V3<real>{act[7], act[8], act[9]};
      // _calcDesiredAcceleration (:689-737)
      Q4<real> q2 = qalign(thq, s.Q);
      V3<real> w_tw, v_tw;
      twist(s.Q, s.p, q2, thp, &w_tw, &v_tw);
      Q4<real> Qi = qinv(s.Q);
      V3<real> vb = qrot(Qi, s.v);
      V3<real> lin = P.k_p * v_tw + P.k_d * (dth_v - vb);
      V3<real> ang = P.k_p * w_tw + P.k_d * (dth_w - s.wb);
      real ades[6] = {smooth_clamp_sym(ang.x, P.acc_clamp), smooth_clamp_sym(ang.y, P.acc_clamp), smooth_clamp_sym(ang.z, P.acc_clamp),
                      smooth_clamp_sym(lin.x, P.acc_clamp), smooth_clamp_sym(lin.y, P.acc_clamp), smooth_clamp_sym(lin.z, P.acc_clamp)};
      // calcMassMatrix3 (:205-304): body-frame bias  [w x I w, R^T (m g) + m w x v_b]
      V3<real> Iw = {P.inertia[0] * s.wb.x, P.inertia[1] * s.wb.y, P.inertia[2] * s.wb.z};
      V3<real> bw = cross(s.wb, Iw);
      V3<real> bl = qrot(Qi, V3<real>{real(0), P.mass * P.grav, real(0)}) + P.mass * cross(s.wb, vb);
      real bias[6] = {bw.x, bw.y, bw.z, bl.x, bl.y, bl.z};
      const int use = (s.p.y - terrain_y > real(0) && s.p.y - terrain_y < P.max_contact_y) ? cmask : 0;   // refCoord = vertical shift by terrain_y
      double y[6], nu[6];
#pragma unroll
      for (int i = 0; i < 6; i++) { y[i] = (double)ades[i] + (double)bias[i] / Md[i]; nu[i] = -y[i]; }
      int iters = 0;
      if constexpr (G == 1) {
        // one thread per environment: structured QP, nothing stored per column
        unsigned lam_mask = 0;
        SQP sq;
        if (use != 0) {
          const Q4<double> Qid = {(double)Qi.w, (double)Qi.x, (double)Qi.y, (double)Qi.z};
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const V3<real> cp = cpt[c >> 1][c & 1];
            V3<double> rb = qrot(Qid, V3<double>{(double)cp.x - (double)s.p.x, (double)cp.y - (double)s.p.y, (double)cp.z - (double)s.p.z});
            sq.r[c][0] = rb.x; sq.r[c][1] = rb.y; sq.r[c][2] = rb.z;
          }
#pragma unroll
          for (int d = 0; d < 5; d++) {
            V3<double> nb = qrot(Qid, V3<double>{(double)P.basis[d][0], (double)P.basis[d][1], (double)P.basis[d][2]});
            sq.n[d][0] = nb.x; sq.n[d][1] = nb.y; sq.n[d][2] = nb.z;
          }
#pragma unroll
          for (int i = 0; i < 6; i++) sq.Di[i] = 1.0 / Md[i];
          sq.valid = (unsigned)use;
          iters = sq.solve_ls(y, P.inv_ridge2, nu, sq_t0 + lane32, sq_sd + lane32);
          lam_mask = sq.tmask(nu);
        }
        if (A.dbg != nullptr && active) {
          double* dl = A.dbg + (size_t)e * DBG_W + 24 + 20 * it;
          for (int c = 0; c < NQCOL; c++) dl[c] = 0.0;
          if (use != 0) {
            const double W[3] = {nu[0] * sq.Di[0], nu[1] * sq.Di[1], nu[2] * sq.Di[2]}, V[3] = {nu[3] * sq.Di[3], nu[4] * sq.Di[4], nu[5] * sq.Di[5]};
            for (int c = 0; c < 4; c++) {
              const double u0 = V[0] + W[1] * sq.r[c][2] - W[2] * sq.r[c][1], u1 = V[1] + W[2] * sq.r[c][0] - W[0] * sq.r[c][2], u2 = V[2] + W[0] * sq.r[c][1] - W[1] * sq.r[c][0];
              for (int d = 0; d < 5; d++) {
                const double t = sq.n[d][0] * u0 + sq.n[d][1] * u1 + sq.n[d][2] * u2;
                if ((lam_mask >> (5 * c + d)) & 1) dl[5 * c + d] = -t * P.inv_ridge2 * 0.5;
              }
            }
          }
        }
      } else {
      double t[QPT::NC];
      QPT qp;
      if (__any_sync(0xffffffffu, use != 0)) {
        // columns J^T V / M in the body frame (:1088-1170): [R^T ((c - p) x n), R^T n]
        const Q4<double> Qid = {(double)Qi.w, (double)Qi.x, (double)Qi.y, (double)Qi.z};
#pragma unroll
        for (int k = 0; k < QPT::NC; k++) {
          const int c = lane + k * G;
          const int ci = c / 5, d = c - 5 * ci;
          const bool valid = c < NQCOL && ((use >> ci) & 1);
          const int cl = (ci >> 1) & 1, cj = ci & 1;
          V3<real> cp = cl ? (cj ? cpt[1][1] : cpt[1][0]) : (cj ? cpt[0][1] : cpt[0][0]);
          V3<double> rel = {(double)cp.x - (double)s.p.x, (double)cp.y - (double)s.p.y, (double)cp.z - (double)s.p.z};
          const int dd = d < 5 ? d : 0;
          V3<double> n = {(double)P.basis[dd][0], (double)P.basis[dd][1], (double)P.basis[dd][2]};
          V3<double> tq = qrot(Qid, cross(rel, n)), fn = qrot(Qid, n);
          const double m = valid ? 1.0 : 0.0;
          qp.a[k][0] = m * tq.x / Md[0]; qp.a[k][1] = m * tq.y / Md[1]; qp.a[k][2] = m * tq.z / Md[2];
          qp.a[k][3] = m * fn.x / Md[3]; qp.a[k][4] = m * fn.y / Md[4]; qp.a[k][5] = m * fn.z / Md[5];
        }
        iters = qp.solve_ls(lane, y, P.inv_ridge2, nu, t, hsm);
      }
      if (A.dbg != nullptr && active) {
        double* dl = A.dbg + (size_t)e * DBG_W + 24 + 20 * it;
#pragma unroll
        for (int k = 0; k < QPT::NC; k++) {
          const int c = lane + k * G;
          if (c < NQCOL) dl[c] = (use != 0 && ((use >> (c / 5)) & 1) && t[k] < 0) ? -t[k] * P.inv_ridge2 * 0.5 : 0.0;
        }
      }
      }
      real qdd[6];
#pragma unroll
      for (int i = 0; i < 6; i++) qdd[i] = (real)(nu[i] + (double)ades[i]);
      if (dbg != nullptr) {
#pragma unroll
        for (int i = 0; i < 6; i++) { dbg[6 * it + i] = (double)qdd[i]; dbg[12 + 6 * it + i] = (double)ades[i]; }
        dbg[75 + it] = (double)iters;
      }
      // stepKinematic (QPservo_v3.lua:305-321 -> Trbdl integrate :758-763, integrateQuater :211-250)
      V3<real> lin_acc = qrot(s.Q, V3<real>{qdd[3], qdd[4], qdd[5]} + cross(s.wb, vb));
      s.v = s.v + h * lin_acc;
      s.wb = s.wb + h * V3<real>{qdd[0], qdd[1], qdd[2]};
      s.p = s.p + h * s.v;
      const Q4<real> q = s.Q;
      const V3<real> w = s.wb;
      Q4<real> qd = {real(0.5) * (-q.x * w.x - q.y * w.y - q.z * w.z), real(0.5) * (q.w * w.x + q.y * w.z - q.z * w.y),
                     real(0.5) * (q.w * w.y - q.x * w.z + q.z * w.x), real(0.5) * (q.w * w.z + q.x * w.y - q.y * w.x)};
      s.Q = qnormalize(Q4<real>{q.w + qd.w * h, q.x + qd.x * h, q.y + qd.y * h, q.z + qd.z * h});
    }
    // ---- calcRefDif2 (:1946-2170)
    real pose_dif, com_lvdif, ref_comvel, ref_y;
    {
      const real delta_frame = FRAME_RATE * RL_STEP;
      const real prev_delta = delta_frame + real(-2.0);
      const real* hs = hist_env + (size_t)(max(0, s.nrep - 4) & (HIST_SLOTS - 1)) * HIST_W;     // sampleSimPose(-1.5), before the push
      const V3<real> sp = {hs[0], hs[1], hs[2]};
      const Q4<real> sq = {hs[3], hs[4], hs[5], hs[6]};
      __syncwarp();
      if (active && lane == 0) {                               // pushSimPose
        real* hw = hist_env + (size_t)(s.nrep & (HIST_SLOTS - 1)) * HIST_W;
        hw[0] = s.p.x; hw[1] = s.p.y; hw[2] = s.p.z; hw[3] = s.Q.w; hw[4] = s.Q.x; hw[5] = s.Q.y; hw[6] = s.Q.z;
      }
      s.nrep += 1;
      V3<real> c0p, c1p; Q4<real> c0q, c1q;
      V3<real> mk[4];
      ref_markers<real>(tab, F, num<real>::fmin(s.iframe + real(0), real(F - 1)), P, nullptr, &c0p, &c0q);
      ref_markers<real>(tab, F, num<real>::fmin(s.iframe + delta_frame, real(F - 1)), P, mk, &c1p, &c1q, gr);
      ref_y = c1p.y;
      const Q4<real> sim_rotY = rotation_y(s.Q);
      V3<real> smk[4];
#pragma unroll
      for (int l = 0; l < 2; l++) {
        V3<real> pos; Q4<real> ori;
        filtered_foot_pos(s, sim_rotY, l, tab, F, gr, &pos, &ori);
        smk[2 * l] = pos + qrot(ori, V3<real>{real(0), real(0), real(0.12)});
        smk[2 * l + 1] = pos + qrot(ori, V3<real>{real(0), real(0), real(-0.12)});
        smk[2 * l].y = ground_y<real>(gr, smk[2 * l].x, smk[2 * l].z); smk[2 * l + 1].y = ground_y<real>(gr, smk[2 * l + 1].x, smk[2 * l + 1].z);
      }
      RowMix<real> ms = row_mix(num<real>::fmax(s.iframe + prev_delta, real(0)), F, true);
      real sr[7];
      tab_mix<real, 7>(tab, ms, T_CDM, sr);
      const V3<real> srp = {sr[0], sr[1], sr[2]};
      const Q4<real> prev_sim_ori = rotation_y(sq), prev_ref_ori = rotation_y(Q4<real>{sr[3], sr[4], sr[5], sr[6]});
      const Q4<real> ips = qinv(prev_sim_ori), ipr = qinv(prev_ref_ori);
      Q4<real> d_sim = quat_mul(ips, s.Q), d_ref = quat_mul(ipr, c1q);
      pose_dif = rotation_angle(quat_mul(qinv(d_sim), d_ref)) * real(5);
      V3<real> v_sim = qrot(ips, s.p - sp), v_ref = qrot(ipr, c1p - srp);
      pose_dif += norm(v_sim - v_ref);
      const V3<real> rcm_t = {srp.x, real(0), srp.z}, rcs_t = {sp.x, ground_y<real>(gr, sp.x, sp.z), sp.z};
      real ef = real(0);
      const real com_d = norm(qrot(ipr, c1p - rcm_t) - qrot(ips, s.p - rcs_t));
#pragma unroll
      for (int l = 0; l < 2; l++) {
        if (!s.leg[l].swing) {
#pragma unroll
          for (int j = 0; j < 2; j++) {
            V3<real> d = qrot(ipr, mk[2 * l + j] - rcm_t) - qrot(ips, smk[2 * l + j] - rcs_t);
            d.y = real(0);
            ef += norm(d) * real(3);
          }
          ef += com_d * real(3);
        } else {
          ef += com_d;
        }
      }
      if (!s.leg[0].swing && !s.leg[1].swing) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
          V3<real> d = qrot(ipr, mk[j] - mk[2 + j]) - qrot(ips, smk[j] - smk[2 + j]);
          d.y = real(0);
          ef += norm(d) * real(3);
        }
      }
      pose_dif += ef * P.ef_scale;
      pose_dif *= real(10);
      const Q4<real> ref_rotY = rotation_y(c1q);
      const Q4<real> isr = qinv(sim_rotY);
      const Q4<real> target_delta = quat_mul(isr, s.tgt);
      V3<real> v1 = qrot(isr, s.p - s.prevcom);
      V3<real> v2 = qrot(quat_mul(target_delta, qinv(ref_rotY)), c1p - c0p);
      V3<real> dxz3 = v1 - v2;
      real dxz = num<real>::sqrt(dxz3.x * dxz3.x + dxz3.z * dxz3.z);
      real dy = num<real>::abs(v1.y - v2.y) * real(1.0);
      com_lvdif = num<real>::sqrt(dxz * dxz + dy * dy) / RL_STEP;
      s.prevcom = s.p;
      pose_dif = pose_dif / real(10);
      ref_comvel = norm(c1p - c0p) / RL_STEP;
    }
    const real timing_mod = P.timing_mod_coef * (num<real>::fmax(com_lvdif, real(0.7)) / num<real>::fmax(ref_comvel, real(0.7)));
    // ---- is_healthy / reward (:1350-1461)
    const real min_y = num<real>::fmin(P.healthy_min, ref_y * real(0.7));
    const real curr_y = s.p.y - ground_y<real>(gr, s.p.x, s.p.z);          // :1361-1365 height above the terrain
    bool healthy = (min_y < curr_y) && (curr_y < P.healthy_max);
    if (rotation_angle(quat_mul(qinv(rotation_y(s.Q)), s.Q)) > P.fall_angle) healthy = false;
    const real delta_y = rotation_angle_about_y(quat_mul(qinv(rotY), s.tgt));
    const real facing = cos1(delta_y) - real(1);
    reward = exp((double)(-pose_dif * real(2))) * exp((double)(-com_lvdif * real(2))) * exp((double)(facing * P.facing_weight));
    done = !healthy;
    const bool bad = obs_has_nan(s) || !(reward == reward);
    write_obs<real, G>(s, tab, F, nf, obs, lane, active, bad, gr);
    if (s.gtime * FRAME_RATE > P.limit_length) done = true;
    if (bad) { reward = real(0); done = true; s.iframe = real(0); }
    s.iframe += FRAME_RATE * RL_STEP * (real(1) + timing_mod);
    s.gtime += RL_STEP;
    if (dbg != nullptr) { dbg[64] = (double)pose_dif; dbg[65] = (double)com_lvdif; dbg[66] = (double)ref_comvel; dbg[67] = (double)timing_mod; }
  }
  s.steps += 1;
  s.ep_ret += reward;
  if (active && lane == 0) { A.rew[e] = (float)reward; A.done[e] = done ? 1 : 0; }
  // ---- fused auto-reset (what the vectorised trainer does after a terminal step, gym_cdm2/train_gym.py)
  if (A.auto_reset && __any_sync(0xffffffffu, done)) {
    __syncwarp();
    if (done && active && A.term_obs != nullptr) {
      float* to = A.term_obs + (size_t)e * A.obs_w;
      for (int i = lane; i < A.obs_w; i += G) to[i] = obs[i];
    }
    __syncwarp();
    if (done) {
      if (active && lane == 0 && A.ep_stats != nullptr) {
        atomicAdd(A.ep_stats + 0, (double)s.ep_ret);
        atomicAdd(A.ep_stats + 1, (double)s.steps);
        atomicAdd(A.ep_stats + 2, 1.0);
      }
      double plan[PLAN_W];
      if (A.reset_plan != nullptr) {
#pragma unroll
        for (int i = 0; i < PLAN_W; i++) plan[i] = A.reset_plan[(size_t)e * PLAN_W + i];
      } else {
        draw_reset_plan(A.seed, e, s.episode, P, plan);
      }
      real h0[HIST_W];
      apply_reset(s, P, A.rc, plan, h0, gr);
      if (active && lane == 0) {
#pragma unroll
        for (int i = 0; i < HIST_W; i++) hist_env[i] = h0[i];
      }
      write_obs<real, G>(s, tab, F, nf, obs, lane, active, false, gr);
    }
  }
  if (active && lane == 0) store_state(s, A.st, A.ist, A.Npad, e);
}

static __global__ void tl_terrain_height_kernel(TLTerrainDev T, long long n, const double* __restrict__ xz, double* __restrict__ out, int* __restrict__ hit) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int hh[3];
  out[k] = tl_terrain_height(T, xz[2 * k], xz[2 * k + 1], hh);
  if (hit) { hit[3 * k] = hh[0]; hit[3 * k + 1] = hh[1]; hit[3 * k + 2] = hh[2]; }
}

// explicit / random reset: one thread per environment
template <typename real> __global__ void cdm_reset_kernel(const ResetArgs<real> A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.count) return;
  int e = A.idx ? A.idx[i] : i;
  if (e < 0 || e >= A.N) return;
  CdmState<real> s;
  load_state(s, A.st, A.ist, A.Npad, e);
  double plan[PLAN_W];
  if (A.plan != nullptr) {
    for (int k = 0; k < PLAN_W; k++) plan[k] = A.plan[(size_t)i * PLAN_W + k];
  } else {
    draw_reset_plan(A.seed, e, s.episode, A.p, plan);
  }
  real h0[HIST_W];
  const Ground* gr = A.ground.T.h != nullptr ? &A.ground : nullptr;
  apply_reset(s, A.p, A.rc, plan, h0, gr);
  real* hist_env = A.hist + (size_t)e * HIST_SLOTS * HIST_W;
  for (int k = 0; k < HIST_W; k++) hist_env[k] = h0[k];
  if (A.obs != nullptr) write_obs<real, 1>(s, A.tab, A.nframes, A.nf_cycle, A.obs + (size_t)e * A.obs_w, 0, true, false, gr);
  store_state(s, A.st, A.ist, A.Npad, e);
}

}  // namespace cdm
