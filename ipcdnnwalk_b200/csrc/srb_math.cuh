// Small fixed-size math used by the SRB step kernels.  Everything is templated on the scalar type
// `real` (float or double) and lives in registers: 3-vectors, row-major 3x3 matrices, wxyz quaternions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace srb {

template <typename T> struct num;
template <> struct num<float> {
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float rsqrt(float x) { return rsqrtf(x); }
#ifdef SRB_LIBM_ATAN2
  static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
#else
  // atan2f without libm's special-case handling (42 -> ~25 instructions; the step kernels call it ~8 times per environment step):
  // a = min(|x|,|y|) / max(|x|,|y|) in [0,1], atan(a) = a * P(a^2) with a degree-8 fit of atan(sqrt(s))/sqrt(s) on [0,1]
  // (max abs error 1.1e-7 evaluated in float, i.e. rounding level of results up to pi), quadrants by reflection.  NaN in -> NaN out;
  // atan2(0, 0) = 0.
  static __device__ __forceinline__ float atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.f ? mn / mx : 0.f;
    const float s = a * a;
    float p = 0.0028340642f;
    p = fmaf(p, s, -0.016005030f); p = fmaf(p, s, 0.042587608f); p = fmaf(p, s, -0.074954458f); p = fmaf(p, s, 0.10636754f);
    p = fmaf(p, s, -0.14202571f); p = fmaf(p, s, 0.19992484f); p = fmaf(p, s, -0.33333066f); p = fmaf(p, s, 1.0f);
    float r = p * a;
    r = ay > ax ? 1.57079637f - r : r;
    r = x < 0.f ? 3.14159274f - r : r;
    r = copysignf(r, y);
    const float n = x + y;
    return n != n ? n : r;
  }
#endif
  static __device__ __forceinline__ float acos(float x) { return acosf(x); }
  static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
  static __device__ __forceinline__ float pow(float a, float b) { return powf(a, b); }
  static __device__ __forceinline__ float floor(float x) { return floorf(x); }
  static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
  static __device__ __forceinline__ float fmax(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float fmin(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float tiny() { return 1e-4f; }   // series threshold for small angles
  static __device__ __forceinline__ float series_below() { return 0.3f; }
  static __device__ __forceinline__ float eps_conv() { return 2e-6f; }
  // MUFU-based approximations (<= 2 ulp): the fp32 mode's tolerance is 1e-3
  static __device__ __forceinline__ float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
  static __device__ __forceinline__ float rsqrt_pos(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
};
template <> struct num<double> {
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
  static __device__ __forceinline__ double acos(double x) { return ::acos(x); }
  static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
  static __device__ __forceinline__ double pow(double a, double b) { return ::pow(a, b); }
  static __device__ __forceinline__ double floor(double x) { return ::floor(x); }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  static __device__ __forceinline__ double fmax(double a, double b) { return ::fmax(a, b); }
  static __device__ __forceinline__ double fmin(double a, double b) { return ::fmin(a, b); }
  static __device__ __forceinline__ double tiny() { return 1e-6; }
  static __device__ __forceinline__ double series_below() { return 0.02; }
  static __device__ __forceinline__ double eps_conv() { return 1e-13; }
  // MUFU seed (2^-23) + two Newton steps: relative error ~1e-16 without the IEEE-division slow path
  static __device__ __forceinline__ double rcp(double d) {
    double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0); x = fma(x, e, x);
    e = fma(-d, x, 1.0); x = fma(x, e, x);
    return x;
  }
  static __device__ __forceinline__ double rsqrt_pos(double d) {
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double hd = 0.5 * d;
    double e = fma(-hd * y, y, 0.5); y = fma(y, e, y);
    e = fma(-hd * y, y, 0.5); y = fma(y, e, y);
    e = fma(-hd * y, y, 0.5); y = fma(y, e, y);
    return y;
  }
};

template <typename T> struct V3 { T x, y, z; };
template <typename T> struct M3 { T m[9]; };   // row-major
template <typename T> struct Q4 { T w, x, y, z; };

template <typename T> __device__ __forceinline__ V3<T> v3(T x, T y, T z) { return V3<T>{x, y, z}; }
template <typename T> __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> __device__ __forceinline__ V3<T> operator*(T s, V3<T> a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T> __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> __device__ __forceinline__ T norm(V3<T> a) { return num<T>::sqrt(dot(a, a)); }

// R v and R^T v
template <typename T> __device__ __forceinline__ V3<T> mul(const M3<T>& R, V3<T> v) {
  return {R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
          R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z};
}
template <typename T> __device__ __forceinline__ V3<T> mulT(const M3<T>& R, V3<T> v) {
  return {R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
          R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z};
}
template <typename T> __device__ __forceinline__ M3<T> matmul(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
// A^T B
template <typename T> __device__ __forceinline__ M3<T> matmulTN(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[i] * B.m[j] + A.m[3 + i] * B.m[3 + j] + A.m[6 + i] * B.m[6 + j];
  return C;
}

// mju_quat2Mat of the normalised quaternion (what mj_kinematics stores in xmat).
template <typename T> __device__ __forceinline__ M3<T> quat2mat_normalized(Q4<T> q) {
  T inv = num<T>::rsqrt_pos(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w *= inv; q.x *= inv; q.y *= inv; q.z *= inv;
  T q00 = q.w * q.w, q01 = q.w * q.x, q02 = q.w * q.y, q03 = q.w * q.z;
  T q11 = q.x * q.x, q12 = q.x * q.y, q13 = q.x * q.z, q22 = q.y * q.y, q23 = q.y * q.z, q33 = q.z * q.z;
  M3<T> R;
  R.m[0] = q00 + q11 - q22 - q33; R.m[1] = 2 * (q12 - q03); R.m[2] = 2 * (q13 + q02);
  R.m[3] = 2 * (q12 + q03); R.m[4] = q00 - q11 + q22 - q33; R.m[5] = 2 * (q23 - q01);
  R.m[6] = 2 * (q13 - q02); R.m[7] = 2 * (q23 + q01); R.m[8] = q00 - q11 - q22 + q33;
  return R;
}

template <typename T> __device__ __forceinline__ Q4<T> quat_mul(Q4<T> a, Q4<T> b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}

// mju_quatIntegrate: q <- normalize(q) * quat(axis = w/|w|, angle = h|w|)
template <typename T> __device__ __forceinline__ Q4<T> quat_integrate(Q4<T> q, V3<T> w, T h) {
  T n = norm(w);
  Q4<T> r;
  if (n < T(1e-15)) {
    r = {T(1), T(0), T(0), T(0)};
  } else {
    T s, c;
    num<T>::sincos(T(0.5) * h * n, &s, &c);
    T k = s * num<T>::rcp(n);
    r = {c, k * w.x, k * w.y, k * w.z};
  }
  T inv = num<T>::rsqrt_pos(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w *= inv; q.x *= inv; q.y *= inv; q.z *= inv;
  return quat_mul(q, r);
}

// Rodrigues formula, reference: axis_angle_to_rotation_matrix (testSRB_mujoco_original_viewer.py:310-327)
template <typename T> __device__ __forceinline__ M3<T> exp_so3(V3<T> aa) {
  T angle = norm(aa);
  M3<T> R;
  if (angle < T(1e-6)) {   // caller-side threshold of SRBEnv.step (:224)
    R.m[0] = 1; R.m[1] = 0; R.m[2] = 0; R.m[3] = 0; R.m[4] = 1; R.m[5] = 0; R.m[6] = 0; R.m[7] = 0; R.m[8] = 1;
    return R;
  }
  T inv = T(1) / angle;
  T ax = aa.x * inv, ay = aa.y * inv, az = aa.z * inv;
  T s, c;
  num<T>::sincos(angle, &s, &c);
  T v = T(1) - c;
  // I + s K + v K^2 with K = [a]x
  R.m[0] = T(1) - v * (ay * ay + az * az); R.m[1] = -s * az + v * ax * ay; R.m[2] = s * ay + v * ax * az;
  R.m[3] = s * az + v * ax * ay; R.m[4] = T(1) - v * (ax * ax + az * az); R.m[5] = -s * ax + v * ay * az;
  R.m[6] = -s * ay + v * ax * az; R.m[7] = s * ax + v * ay * az; R.m[8] = T(1) - v * (ax * ax + ay * ay);
  return R;
}

// Closed-form SE(3) logarithm of T_rel = [Rr, t; 0 1]  (replaces scipy.linalg.logm of the 4x4,
// testSRB_mujoco_original_viewer.py:175-201).  Returns v (translation part) and w (rotation part).
template <typename T> __device__ __forceinline__ void log_se3(const M3<T>& Rr, V3<T> t, V3<T>* v_out, V3<T>* w_out) {
  const T PI = T(3.14159265358979323846);
  V3<T> a = {T(0.5) * (Rr.m[7] - Rr.m[5]), T(0.5) * (Rr.m[2] - Rr.m[6]), T(0.5) * (Rr.m[3] - Rr.m[1])};  // sin(th) * axis
  T sn = norm(a);
  T cs = T(0.5) * (Rr.m[0] + Rr.m[4] + Rr.m[8] - T(1));
  T th = num<T>::atan2(sn, cs);
  V3<T> w;
  if (th < num<T>::tiny()) {
    T k = T(1) + th * th * T(1.0 / 6.0);
    w = k * a;
  } else if (th < PI - T(1e-3)) {
    w = (th / sn) * a;
  } else {
    // near pi: axis from the symmetric part, sign from `a`
    T xx = num<T>::sqrt(num<T>::fmax(T(0), (Rr.m[0] - cs) / (T(1) - cs)));
    T yy = num<T>::sqrt(num<T>::fmax(T(0), (Rr.m[4] - cs) / (T(1) - cs)));
    T zz = num<T>::sqrt(num<T>::fmax(T(0), (Rr.m[8] - cs) / (T(1) - cs)));
    // fix relative signs using the largest component
    T sxy = Rr.m[1] + Rr.m[3], sxz = Rr.m[2] + Rr.m[6], syz = Rr.m[5] + Rr.m[7];
    if (xx >= yy && xx >= zz) { if (sxy < 0) yy = -yy; if (sxz < 0) zz = -zz; }
    else if (yy >= zz) { if (sxy < 0) xx = -xx; if (syz < 0) zz = -zz; }
    else { if (sxz < 0) xx = -xx; if (syz < 0) yy = -yy; }
    V3<T> ax = {xx, yy, zz};
    if (dot(ax, a) < 0) ax = T(-1) * ax;
    T n = norm(ax);
    w = (th / n) * ax;
  }
  // V^-1 = I - 1/2 [w]x + c2 [w]x^2,  c2 = (1/th^2)(1 - th sin th / (2 (1 - cos th)))
  //       = (1 - (th/2) cot(th/2)) / th^2 ; series below the cancellation threshold
  T c2;
  T th2 = th * th;
  if (th < num<T>::series_below()) {
    c2 = T(1.0 / 12.0) + th2 * (T(1.0 / 720.0) + th2 * T(1.0 / 30240.0));
  } else {
    T sh, ch;
    num<T>::sincos(T(0.5) * th, &sh, &ch);
    c2 = (T(1) - T(0.5) * th * ch / sh) / th2;
  }
  V3<T> wt = cross(w, t);
  V3<T> wwt = cross(w, wt);
  *v_out = t - T(0.5) * wt + c2 * wwt;
  *w_out = w;
}

// yaw-only frame from the body x axis projected to the ground (global2projected_SRB / forward_facing,
// testSRB_mujoco_original_viewer.py:83-114): columns x=(fx,fy,0), y=(-fy,fx,0), z=(0,0,1).
template <typename T> __device__ __forceinline__ void yaw_frame(const M3<T>& R, T* fx, T* fy) {
  T x = R.m[0], y = R.m[3];
  T n2 = x * x + y * y;
  if (n2 > T(1e-30)) { T inv = num<T>::rsqrt_pos(n2); x = x * inv; y = y * inv; }
  *fx = x; *fy = y;
}
// ff^T v and ff v for the yaw frame (fx, fy)
template <typename T> __device__ __forceinline__ V3<T> yawT(T fx, T fy, V3<T> v) { return {fx * v.x + fy * v.y, -fy * v.x + fx * v.y, v.z}; }
template <typename T> __device__ __forceinline__ V3<T> yawR(T fx, T fy, V3<T> v) { return {fx * v.x - fy * v.y, fy * v.x + fx * v.y, v.z}; }

}  // namespace srb
