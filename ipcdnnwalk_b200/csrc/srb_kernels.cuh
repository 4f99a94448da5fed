// SRBTrack batched environment step for sm_100a.
//
// Mapping: G lanes (a power of two, 1..32) cooperate on one environment, 32/G environments per warp,
// one warp per CTA (no CTA-level sharing is needed, and 1-warp CTAs give the finest SM load balance at
// the 4096-environment batch the headline config uses).  The scalar part of the step (kinematics,
// action decoding, swing-foot filter, integration, observation, reward, termination) is computed
// redundantly by the G lanes (same instruction stream, no extra issue slots); the contact-force QP --
// the dominant cost -- is spread over the lanes: each lane owns 40/G friction-basis columns, the 6x6
// semismooth-Newton Hessian is all-reduced with xor-shuffles and factorised (LDL^T) in registers.
//
// Reference for every block is cited inline (paths relative to the reference repository root).
#pragma once
#include "srb_math.cuh"
#include "srb_types.h"

namespace srb {

// ------------------------------------------------------------------------------------------------
// state in registers
template <typename real> struct EnvState {
  V3<real> pos; Q4<real> quat; V3<real> vlin, vang; Q4<real> qlag;
  V3<real> foot[2]; real yc[2], ys[2];
  real filt_yaw[2]; real ballx[4]; real ballv[4];
  real ep_ret;
  real nu[6];        // QP dual from the last sub-step (warm start only; the optimum is unique)
  real planar[2];    // initial_pos_planar (terrain variant)
  int flags, frame, rif, clip, fcount, impulse, episode;
};

template <typename real> __device__ __forceinline__ void load_state(EnvState<real>& s, const real* __restrict__ st,
                                                                   const int* __restrict__ ist, int Npad, int e) {
#define LD(f) st[(size_t)(f) * Npad + e]
  s.pos = {LD(F_POS), LD(F_POS + 1), LD(F_POS + 2)};
  s.quat = {LD(F_QUAT), LD(F_QUAT + 1), LD(F_QUAT + 2), LD(F_QUAT + 3)};
  s.vlin = {LD(F_QVEL), LD(F_QVEL + 1), LD(F_QVEL + 2)};
  s.vang = {LD(F_QVEL + 3), LD(F_QVEL + 4), LD(F_QVEL + 5)};
  s.qlag = {LD(F_QLAG), LD(F_QLAG + 1), LD(F_QLAG + 2), LD(F_QLAG + 3)};
  s.foot[0] = {LD(F_FOOTL), LD(F_FOOTL + 1), LD(F_FOOTL + 2)};
  s.foot[1] = {LD(F_FOOTR), LD(F_FOOTR + 1), LD(F_FOOTR + 2)};
  s.yc[0] = LD(F_FYAWL); s.ys[0] = LD(F_FYAWL + 1); s.yc[1] = LD(F_FYAWR); s.ys[1] = LD(F_FYAWR + 1);
  s.filt_yaw[0] = LD(F_FILTYAW); s.filt_yaw[1] = LD(F_FILTYAW + 1);
#pragma unroll
  for (int i = 0; i < 4; i++) { s.ballx[i] = LD(F_BALLX + i); s.ballv[i] = LD(F_BALLV + i); }
  s.ep_ret = LD(F_EPRET);
#pragma unroll
  for (int i = 0; i < 6; i++) s.nu[i] = LD(F_NU + i);
  s.planar[0] = LD(F_PLANAR); s.planar[1] = LD(F_PLANAR + 1);
#undef LD
#define LDI(f) ist[(size_t)(f) * Npad + e]
  s.flags = LDI(I_FLAGS); s.frame = LDI(I_FRAME); s.rif = LDI(I_RIF); s.clip = LDI(I_CLIP);
  s.fcount = LDI(I_FCOUNT); s.impulse = LDI(I_IMPULSE); s.episode = LDI(I_EPISODE);
#undef LDI
}

template <typename real> __device__ __forceinline__ void store_state(const EnvState<real>& s, real* __restrict__ st,
                                                                    int* __restrict__ ist, int Npad, int e) {
#define ST(f, v) st[(size_t)(f) * Npad + e] = (v)
  ST(F_POS, s.pos.x); ST(F_POS + 1, s.pos.y); ST(F_POS + 2, s.pos.z);
  ST(F_QUAT, s.quat.w); ST(F_QUAT + 1, s.quat.x); ST(F_QUAT + 2, s.quat.y); ST(F_QUAT + 3, s.quat.z);
  ST(F_QVEL, s.vlin.x); ST(F_QVEL + 1, s.vlin.y); ST(F_QVEL + 2, s.vlin.z);
  ST(F_QVEL + 3, s.vang.x); ST(F_QVEL + 4, s.vang.y); ST(F_QVEL + 5, s.vang.z);
  ST(F_QLAG, s.qlag.w); ST(F_QLAG + 1, s.qlag.x); ST(F_QLAG + 2, s.qlag.y); ST(F_QLAG + 3, s.qlag.z);
  ST(F_FOOTL, s.foot[0].x); ST(F_FOOTL + 1, s.foot[0].y); ST(F_FOOTL + 2, s.foot[0].z);
  ST(F_FOOTR, s.foot[1].x); ST(F_FOOTR + 1, s.foot[1].y); ST(F_FOOTR + 2, s.foot[1].z);
  ST(F_FYAWL, s.yc[0]); ST(F_FYAWL + 1, s.ys[0]); ST(F_FYAWR, s.yc[1]); ST(F_FYAWR + 1, s.ys[1]);
  ST(F_FILTYAW, s.filt_yaw[0]); ST(F_FILTYAW + 1, s.filt_yaw[1]);
#pragma unroll
  for (int i = 0; i < 4; i++) { ST(F_BALLX + i, s.ballx[i]); ST(F_BALLV + i, s.ballv[i]); }
  ST(F_EPRET, s.ep_ret);
#pragma unroll
  for (int i = 0; i < 6; i++) ST(F_NU + i, s.nu[i]);
  ST(F_PLANAR, s.planar[0]); ST(F_PLANAR + 1, s.planar[1]);
#undef ST
#define STI(f, v) ist[(size_t)(f) * Npad + e] = (v)
  STI(I_FLAGS, s.flags); STI(I_FRAME, s.frame); STI(I_RIF, s.rif); STI(I_CLIP, s.clip);
  STI(I_FCOUNT, s.fcount); STI(I_IMPULSE, s.impulse); STI(I_EPISODE, s.episode);
#undef STI
}

// ------------------------------------------------------------------------------------------------
// Terrain query = the vertical mj_ray of terrain_height_ray (env/testSRB_mujoco_original_viewer_terrain.py:741-752)
// over the group-0 geoms of SRB5_playground.xml: plane, height fields, pyramid boxes.  The highest surface wins,
// the first geom in model order wins ties (mj_ray keeps the strictly smallest distance).  Integer outputs: geom id,
// and for height fields the cell (c, r) and the triangle of the (c,r)-(c+1,r+1) split.
struct TerrainHit { int geom; int cell_c, cell_r, tri; };
template <typename real>
__device__ __forceinline__ real terrain_query(const TerrainDev& T, real x, real y, TerrainHit* hit) {
  real best = real(0);
  int gid = -1, cc = -1, cr = -1, tri = -1;
  bool have = false;
  if (num<real>::abs(x) <= (real)T.plane_hx && num<real>::abs(y) <= (real)T.plane_hy) { best = real(0); gid = T.plane_geom; have = true; }
  for (int i = 0; i < T.n_hf; i++) {
    const HFieldDev& h = T.hf[i];
    real lx = x - (real)h.px, ly = y - (real)h.py;
    if (num<real>::abs(lx) > (real)h.sx || num<real>::abs(ly) > (real)h.sy) continue;
    real dx = real(2) * (real)h.sx / (real)(h.ncol - 1), dy = real(2) * (real)h.sy / (real)(h.nrow - 1);
    real u = (lx + (real)h.sx) / dx, v = (ly + (real)h.sy) / dy;
    int c = min(max((int)num<real>::floor(u), 0), h.ncol - 2), r = min(max((int)num<real>::floor(v), 0), h.nrow - 2);
    real fu = u - (real)c, fv = v - (real)r;
    const float* d0 = h.data + (size_t)r * h.ncol + c;
    real z00 = (real)__ldg(d0) * (real)h.sz, z10 = (real)__ldg(d0 + 1) * (real)h.sz;
    real z01 = (real)__ldg(d0 + h.ncol) * (real)h.sz, z11 = (real)__ldg(d0 + h.ncol + 1) * (real)h.sz;
    int t; real z;
    if (fu >= fv) { t = 0; z = z00 + fu * (z10 - z00) + fv * (z11 - z10); }
    else { t = 1; z = z00 + fv * (z01 - z00) + fu * (z11 - z01); }
    z += (real)h.pz;
    if (!have || z > best) { best = z; gid = h.geom_id; cc = c; cr = r; tri = t; have = true; }
  }
  for (int b = 0; b < T.n_body; b++) {
    const TerrainBodyDev& B = T.body[b];
    if ((double)x < B.x0 || (double)x > B.x1 || (double)y < B.y0 || (double)y > B.y1) continue;
    if (B.nested) {
      // concentric layers whose half sizes shrink with the index (a pyramid): "inside layer k" is monotone in k, so the
      // highest layer under the point is found by bisection (7 probes for 81 layers) with the same comparisons as the scan
      const double* b0 = T.box + (size_t)B.first * 5;
      const real ax = num<real>::abs(x - (real)b0[0]), ay = num<real>::abs(y - (real)b0[1]);
      int lo = B.first - 1, hi = B.first + B.count - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const double* bx = T.box + (size_t)mid * 5;
        if (ax <= (real)bx[2] && ay <= (real)bx[3]) lo = mid; else hi = mid - 1;
      }
      if (lo >= B.first) {
        real top = (real)T.box[(size_t)lo * 5 + 4];
        if (!have || top > best) { best = top; gid = T.box_geom0 + lo; cc = cr = tri = -1; have = true; }
      }
      continue;
    }
    for (int k = B.first + B.count - 1; k >= B.first; k--) {           // tops ascend in model order: highest first
      const double* bx = T.box + (size_t)k * 5;
      if (num<real>::abs(x - (real)bx[0]) <= (real)bx[2] && num<real>::abs(y - (real)bx[1]) <= (real)bx[3]) {
        real top = (real)bx[4];
        if (!have || top > best) { best = top; gid = T.box_geom0 + k; cc = cr = tri = -1; have = true; }
        break;
      }
    }
  }
  if (hit) { hit->geom = gid; hit->cell_c = cc; hit->cell_r = cr; hit->tri = tri; }
  return best;
}

// Reference tables of the terrain variant (viewer_terrain.py:1043-1065 after the planar shift of
// SRBEnv5_mocap_list_terrain_window.py:636): positions are the flat tables shifted by initial_pos_planar and lifted
// by the terrain height underneath; "original" foot z (== 0 <=> reference contact) is the flat table's.
template <typename real> struct LiftedFrame { V3<real> pos; real hc; V3<real> foot[2]; bool contact[2]; };
template <typename real>
__device__ __forceinline__ LiftedFrame<real> lifted_frame(const real* __restrict__ fr, const real planar[2], const TerrainDev& T) {
  LiftedFrame<real> L;
  real x = fr[0] + planar[0], y = fr[1] + planar[1];
  L.hc = terrain_query<real>(T, x, y, nullptr);
  L.pos = {x, y, fr[2] + L.hc};
#pragma unroll
  for (int i = 0; i < 2; i++) {
    real fx = fr[18 + 3 * i] + planar[0], fy = fr[19 + 3 * i] + planar[1];
    L.foot[i] = {fx, fy, real(0) + terrain_query<real>(T, fx, fy, nullptr)};
    L.contact[i] = fr[20 + 3 * i] == real(0);
  }
  return L;
}

// ------------------------------------------------------------------------------------------------
// counter-based RNG (Philox-4x32-10) for device-side episode resets
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {   // 53-bit uniform in [0,1)
  return (double)((((unsigned long long)a << 32 | b) >> 11)) * (1.0 / 9007199254740992.0);
}

// draws of SRBEnv.reset / set_state (SRBEnv5_mocap_list_window.py:810,823,860-867, viewer.py:719-731):
// plan = clip, start frame, uniform(-.05,.05)^3 velocity noise, small random quaternion.
template <typename real>
__device__ inline void draw_reset_plan(unsigned long long seed, int env, int episode, const ClipDesc<real>* clips, int nclips, double plan[9], bool terrain = false) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c[4] = {(uint32_t)env, (uint32_t)episode, 0x5EB0u, 0u};
  philox4x32(c, k0, k1);
  uint32_t d[4] = {(uint32_t)env, (uint32_t)episode, 0x5EB1u, 0u};
  philox4x32(d, k0, k1);
  uint32_t g[4] = {(uint32_t)env, (uint32_t)episode, 0x5EB2u, 0u};
  philox4x32(g, k0, k1);
  const float U32 = 2.3283064365386963e-10f;                         // 2^-32
  int clip = min((int)(__umulhi(c[0], (uint32_t)nclips)), nclips - 1);   // random.randrange(n)
  int cyc = clips[clip].cycle_length;
  int hi = terrain ? max(cyc - 2, 0) : max(cyc - 2, 512);            // random.randint(0, hi) inclusive (Q12)
  int frame = min((int)__umulhi(c[1], (uint32_t)(hi + 1)), hi);
  plan[0] = clip; plan[1] = frame;
  plan[2] = (double)((((float)c[2] + 0.5f) * U32 * 2.0f - 1.0f) * 5e-2f);
  plan[3] = (double)((((float)c[3] + 0.5f) * U32 * 2.0f - 1.0f) * 5e-2f);
  plan[4] = (double)((((float)d[0] + 0.5f) * U32 * 2.0f - 1.0f) * 5e-2f);
  float ang = (((float)d[1] + 0.5f) * U32 * 2.0f - 1.0f) * 5e-2f;
  // axis ~ normalised N(0,I) (Box-Muller), angle ~ U(-.05,.05)   (viewer.py:719-731)
  float u1 = ((float)(d[2] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(d[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.2831853f * u2, &sn, &cs);
  float n0 = r * cs, n1 = r * sn;
  float u3 = ((float)(g[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u4 = ((float)(g[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float n2 = sqrtf(-2.0f * __logf(u3)) * __cosf(6.2831853f * u4);
  float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
  if (nn < 1e-20f) { n0 = 1.f; n1 = 0.f; n2 = 0.f; nn = 1.f; }
  float sh, ch;
  __sincosf(0.5f * ang, &sh, &ch);
  sh /= nn;
  plan[5] = (double)ch; plan[6] = (double)(n0 * sh); plan[7] = (double)(n1 * sh); plan[8] = (double)(n2 * sh);
}

// ------------------------------------------------------------------------------------------------
// SRBEnv.set_state (SRBEnv5_mocap_list_window.py:847-911) from an explicit plan.
// plan[5] (noise quat w) == NaN selects testing_mode (no noise).
template <typename real>
__device__ inline void apply_reset(EnvState<real>& s, const ClipDesc<real>* clips, int nclips, const double plan[9],
                                   real* __restrict__ hist_env, const TerrainDev* T = nullptr) {
  int clip = min(max((int)plan[0], 0), nclips - 1);
  const ClipDesc<real> cd = clips[clip];
  int f = min(max((int)plan[1], 0), cd.nframes - 1);
  const real* fr = cd.frames + (size_t)f * CLIP_W;
  s.clip = clip; s.rif = f; s.frame = 0;
  s.pos = {fr[0], fr[1], fr[2]};
  Q4<real> mq = {fr[28], fr[29], fr[30], fr[31]};
  s.vlin = {fr[12], fr[13], fr[14]};
  s.vang = {fr[15], fr[16], fr[17]};
  if (plan[5] == plan[5]) {
    Q4<real> nq = {(real)plan[5], (real)plan[6], (real)plan[7], (real)plan[8]};
    Q4<real> pq = quat_mul(nq, mq);
    real inv = real(1) / num<real>::sqrt(pq.w * pq.w + pq.x * pq.x + pq.y * pq.y + pq.z * pq.z);
    s.quat = {pq.w * inv, pq.x * inv, pq.y * inv, pq.z * inv};
    s.vlin = {s.vlin.x + (real)plan[2], s.vlin.y + (real)plan[3], s.vlin.z + (real)plan[4]};
  } else {
    s.quat = mq;
  }
  s.qlag = s.quat;                                 // mj_forward refreshes xmat (:892)
  bool lc = (fr[20] == real(0)), rc = (fr[23] == real(0));
  s.foot[0] = {fr[18], fr[19], real(0)};
  s.foot[1] = {fr[21], fr[22], real(0)};
  if (T != nullptr) {                              // terrain variant (..._terrain_window.py:417-443): lifted tables
    LiftedFrame<real> L = lifted_frame<real>(fr, s.planar, *T);
    s.pos = L.pos; s.foot[0] = L.foot[0]; s.foot[1] = L.foot[1];
  }
  s.yc[0] = fr[24]; s.ys[0] = fr[25]; s.yc[1] = fr[26]; s.ys[1] = fr[27];
  // legacy slices [:3] / [3:] of the 10-entry site list (Q2): right = left OR right
  s.flags = (lc ? FLAG_L : 0) | ((lc || rc) ? FLAG_R : 0) | (cd.is_stand ? FLAG_STAND : 0);   // clears FLAG_FILT / FLAG_NU
  s.filt_yaw[0] = s.filt_yaw[1] = real(0);
#pragma unroll
  for (int i = 0; i < 4; i++) { s.ballx[i] = real(0); s.ballv[i] = real(0); }
  s.ep_ret = real(0);
#pragma unroll
  for (int i = 0; i < 6; i++) s.nu[i] = real(0);
  s.episode += 1;     // frame_counter / impulse persist across resets, as in the reference
  // history entry 0 (:895-911)
  real* h0 = hist_env;
  h0[0] = s.pos.x; h0[1] = s.pos.y; h0[2] = s.pos.z;
  h0[3] = s.qlag.w; h0[4] = s.qlag.x; h0[5] = s.qlag.y; h0[6] = s.qlag.z; h0[7] = real(0);
}

// ------------------------------------------------------------------------------------------------
// observation (SRBEnv.get_obs_window, SRBEnv5_mocap_list_window.py:1055-1168).
// Self part (25) from the state; look-ahead part = 5 precomputed 25-float records of the clip table
// (they depend on the reference frame only).  `lane`/`G`: cooperative copy of the window.
template <typename real, int G>
__device__ __forceinline__ void write_obs(const EnvState<real>& s, const ClipDesc<real>& cd, float* __restrict__ o, int lane, bool active,
                                          const TerrainDev* T = nullptr, bool window = true) {
  M3<real> R = quat2mat_normalized(s.qlag);
  real fx, fy;
  yaw_frame(R, &fx, &fy);
  if (active && lane == 0) {
    real h0 = s.pos.z;
    if (T != nullptr) h0 -= terrain_query<real>(*T, s.pos.x, s.pos.y, nullptr);     // terrain-relative height (:549-552)
    o[0] = (float)h0;
    // (proj^T R)[:, :2] flattened column-major
    o[1] = (float)(fx * R.m[0] + fy * R.m[3]); o[2] = (float)(-fy * R.m[0] + fx * R.m[3]); o[3] = (float)R.m[6];
    o[4] = (float)(fx * R.m[1] + fy * R.m[4]); o[5] = (float)(-fy * R.m[1] + fx * R.m[4]); o[6] = (float)R.m[7];
    V3<real> vl = yawT(fx, fy, s.vlin);
    V3<real> va = yawT(fx, fy, mul(R, s.vang));
    o[7] = (float)vl.x; o[8] = (float)vl.y; o[9] = (float)vl.z;
    o[10] = (float)va.x; o[11] = (float)va.y; o[12] = (float)va.z;
    V3<real> lf = yawT(fx, fy, s.foot[0] - s.pos);
    V3<real> rf = yawT(fx, fy, s.foot[1] - s.pos);
    real l00 = fx * s.yc[0] + fy * s.ys[0], l10 = -fy * s.yc[0] + fx * s.ys[0];
    real r10 = -fy * s.yc[1] + fx * s.ys[1];
    o[13] = (float)lf.x; o[14] = (float)lf.y; o[15] = (float)lf.z; o[16] = (float)num<real>::atan2(l10, l00);
    o[17] = (float)rf.x; o[18] = (float)rf.y; o[19] = (float)rf.z; o[20] = (float)num<real>::atan2(r10, l00);   // Q1
    int c = s.flags & 3;
    o[21] = c == 0 ? 1.f : 0.f; o[22] = c == 1 ? 1.f : 0.f; o[23] = c == 2 ? 1.f : 0.f; o[24] = c == 3 ? 1.f : 0.f;
  }
  if (active && T != nullptr) {
    // terrain variant: the look-ahead features depend on the terrain under the shifted reference (:574-623);
    // one frame per lane
    const int f0 = s.frame + s.rif + 1;
    for (int wi = lane; wi < 5; wi += G) {
      const int f = min(f0 + wi, cd.nframes - 1);
      const real* fr = cd.frames + (size_t)f * CLIP_W;
      LiftedFrame<real> L = lifted_frame<real>(fr, s.planar, *T);
      M3<real> Rr;
#pragma unroll
      for (int i = 0; i < 9; i++) Rr.m[i] = fr[3 + i];
      real rx, ry;
      yaw_frame(Rr, &rx, &ry);
      float* w = o + FEAT_W * (1 + wi);
      w[0] = (float)(L.pos.z - L.hc);
      w[1] = (float)(rx * Rr.m[0] + ry * Rr.m[3]); w[2] = (float)(-ry * Rr.m[0] + rx * Rr.m[3]); w[3] = (float)Rr.m[6];
      w[4] = (float)(rx * Rr.m[1] + ry * Rr.m[4]); w[5] = (float)(-ry * Rr.m[1] + rx * Rr.m[4]); w[6] = (float)Rr.m[7];
      V3<real> vl = yawT(rx, ry, V3<real>{fr[12], fr[13], fr[14]});
      V3<real> va = yawT(rx, ry, mul(Rr, V3<real>{fr[15], fr[16], fr[17]}));
      w[7] = (float)vl.x; w[8] = (float)vl.y; w[9] = (float)vl.z; w[10] = (float)va.x; w[11] = (float)va.y; w[12] = (float)va.z;
      int c = 0;
#pragma unroll
      for (int i = 0; i < 2; i++) {
        V3<real> l = yawT(rx, ry, L.foot[i] - L.pos);
        real yc = fr[24 + 2 * i], ys = fr[25 + 2 * i];
        w[13 + 4 * i] = (float)l.x; w[14 + 4 * i] = (float)l.y; w[15 + 4 * i] = (float)l.z;
        w[16 + 4 * i] = (float)num<real>::atan2(-ry * yc + rx * ys, rx * yc + ry * ys);
        if (L.contact[i]) c |= (1 << i);
      }
      w[21] = c == 0 ? 1.f : 0.f; w[22] = c == 1 ? 1.f : 0.f; w[23] = c == 2 ? 1.f : 0.f; w[24] = c == 3 ? 1.f : 0.f;
    }
  } else if (active && window) {
    // 5 consecutive frame records are contiguous in the feature table; at the table end every frame index is clamped on its own
    // (min(f, nframes - 1)), the rule the terrain variant, the reward and the termination test use too.  The reference has no clamp
    // (a clip shorter than start frame + 512 + 6 would raise IndexError there), so only consistency matters.
    const int f0 = s.frame + s.rif + 1;
    if (f0 + 4 > cd.nframes - 1) {
      for (int i = lane; i < 5 * FEAT_W; i += G) {
        const int wi = i / FEAT_W, c = i - wi * FEAT_W;
        o[FEAT_W + i] = __ldg(cd.feat + (size_t)min(max(f0 + wi, 0), cd.nframes - 1) * FEAT_W + c);
      }
      return;
    }
    const float* __restrict__ src = cd.feat + (size_t)f0 * FEAT_W;
    constexpr int PER = (5 * FEAT_W + G - 1) / G;
    if constexpr (PER <= 32) {
      float v[PER];
#pragma unroll
      for (int k = 0; k < PER; k++) { int i = lane + k * G; v[k] = (i < 5 * FEAT_W) ? __ldg(src + i) : 0.f; }
#pragma unroll
      for (int k = 0; k < PER; k++) { int i = lane + k * G; if (i < 5 * FEAT_W) o[FEAT_W + i] = v[k]; }
    } else {
#pragma unroll 5
      for (int i = lane; i < 5 * FEAT_W; i += G) o[FEAT_W + i] = __ldg(src + i);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// group helpers
template <int G> __device__ __forceinline__ float gsum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int G> __device__ __forceinline__ double gsum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// All-reduce of the 21 lower-triangle Hessian partials over the G lanes of a group.
// Generic: xor-shuffle butterfly (21 * log2 G shuffles + adds).
// G == 8: staged through shared memory with 128-bit accesses -- every lane stores its 21 partials (padded row,
// stride chosen so a quarter-warp's 16-byte chunks fall in distinct bank groups), lanes 0..5 of a group each sum
// 4 entries over the 8 rows, the 24 sums are read back by all lanes: ~20 (fp32) / ~41 (fp64) LSU instructions
// + 28 adds instead of 63 / 126 SHFL + 63 adds.  (The same staging for G == 4 was measured 4-6 % SLOWER than its 42-shuffle butterfly:
// profiles/r02_ab_g4_smem_reduce_*.txt.)
template <typename real, int G> struct HReduce {
  static constexpr int SMEM_WORDS = 1;
  static __device__ __forceinline__ void run(real H[21], real* sm, int lane32) {
#pragma unroll
    for (int i = 0; i < 21; i++) H[i] = gsum<G>(H[i]);
  }
};
template <> struct HReduce<float, 8> {
  static constexpr int ROW = 28, RES = 32 * ROW, SMEM_WORDS = RES + 4 * 24;
  static __device__ __forceinline__ void run(float H[21], float* sm, int lane32) {
    float4* row = reinterpret_cast<float4*>(sm + lane32 * ROW);
    row[0] = make_float4(H[0], H[1], H[2], H[3]); row[1] = make_float4(H[4], H[5], H[6], H[7]);
    row[2] = make_float4(H[8], H[9], H[10], H[11]); row[3] = make_float4(H[12], H[13], H[14], H[15]);
    row[4] = make_float4(H[16], H[17], H[18], H[19]); row[5] = make_float4(H[20], 0.f, 0.f, 0.f);
    __syncwarp();
    const int j = lane32 & 7, base = lane32 & ~7, grp = lane32 >> 3;
    if (j < 6) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 8; r++) {
        float4 v = *reinterpret_cast<const float4*>(sm + (base + r) * ROW + 4 * j);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(sm + RES + grp * 24 + 4 * j) = acc;
    }
    __syncwarp();
    const float4* res = reinterpret_cast<const float4*>(sm + RES + grp * 24);
#pragma unroll
    for (int q = 0; q < 5; q++) { float4 v = res[q]; H[4 * q] = v.x; H[4 * q + 1] = v.y; H[4 * q + 2] = v.z; H[4 * q + 3] = v.w; }
    H[20] = res[5].x;
  }
};
template <> struct HReduce<double, 8> {
  static constexpr int ROW = 26, RES = 32 * ROW, SMEM_WORDS = RES + 4 * 24;
  static __device__ __forceinline__ void run(double H[21], double* sm, int lane32) {
    double2* row = reinterpret_cast<double2*>(sm + lane32 * ROW);
#pragma unroll
    for (int q = 0; q < 10; q++) row[q] = make_double2(H[2 * q], H[2 * q + 1]);
    row[10] = make_double2(H[20], 0.0); row[11] = make_double2(0.0, 0.0);
    __syncwarp();
    const int j = lane32 & 7, base = lane32 & ~7, grp = lane32 >> 3;
    if (j < 6) {
      double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const double2* src = reinterpret_cast<const double2*>(sm + (base + r) * ROW + 4 * j);
        double2 v0 = src[0], v1 = src[1];
        a0.x += v0.x; a0.y += v0.y; a1.x += v1.x; a1.y += v1.y;
      }
      double2* dst = reinterpret_cast<double2*>(sm + RES + grp * 24 + 4 * j);
      dst[0] = a0; dst[1] = a1;
    }
    __syncwarp();
    const double2* res = reinterpret_cast<const double2*>(sm + RES + grp * 24);
#pragma unroll
    for (int q = 0; q < 10; q++) { double2 v = res[q]; H[2 * q] = v.x; H[2 * q + 1] = v.y; }
    H[20] = res[10].x;
  }
};

// 6x6 SPD solve H x = rhs by LDL^T, fully unrolled in registers.  H: lower triangle, index i*(i+1)/2+j.
template <typename real> __device__ __forceinline__ void ldl_solve6(const real* Hl, const real* rhs, real* x) {
  real L[6][6];      // unit lower factor
  real W[6][6];      // W[i][k] = L[i][k] * d[k]
  real dinv[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    real d = Hl[j * (j + 1) / 2 + j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[j][k] * W[j][k];
    dinv[j] = num<real>::rcp(d);
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      real v = Hl[i * (i + 1) / 2 + j];
#pragma unroll
      for (int k = 0; k < j; k++) v -= L[i][k] * W[j][k];
      W[i][j] = v;
      L[i][j] = v * dinv[j];
    }
  }
  real z[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    real v = rhs[i];
#pragma unroll
    for (int k = 0; k < i; k++) v -= L[i][k] * z[k];
    z[i] = v;
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    real v = z[i] * dinv[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) v -= L[k][i] * x[k];
    x[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Contact-force QP (solve_qp_for_contact_forces, testSRB_mujoco_original_viewer.py:515-599):
//     min |qdd - qdd_d|^2 + w |lam|^2   s.t.  M qdd - C lam = -b,  lam >= 0.
// qdd is eliminated (M is diagonal and constant for the free body):  with A = M^-1 C, y = qdd_d + M^-1 b
//     min_{lam>=0} |A lam - y|^2 + w |lam|^2.
// Its dual is an unconstrained, strictly convex piecewise-quadratic in nu = A lam - y (= qdd - qdd_d):
//     phi(nu) = 1/2|nu|^2 + nu.y + 1/(2w) |max(0, -A^T nu)|^2 ,   lam = max(0, -A^T nu)/w,
// minimised by a semismooth Newton iteration on 6 unknowns: on the active set S={j: a_j.nu<0},
// (I + A_S A_S^T / w) nu = -y.  A full step is accepted when S is reproduced (exact optimum) or phi
// decreases, otherwise it is halved (globally convergent, finite termination).
// Columns: c = s*8 + f*4 + d  (site s of 5, foot f, friction direction d), lane owns c = lane + k*G.
template <typename real, int G, int NCOLS = 40> struct QP {
  static constexpr int NC = (NCOLS + G - 1) / G;
  real a[NC][6];
  // G == 4 with the SRBTrack column order: a lane's ten columns are ONE friction direction over both feet (k even: left, k odd: right)
  // and the five sites (k >> 1): their linear parts are bl (= M^-1 B, the same for both feet) times the foot's contact flag mf.
  static constexpr bool SHARED_LIN4 = (G == 4 && NCOLS == 40);
  real bl[3], mf[2];

  __device__ __forceinline__ void build(int lane, const M3<real>& R, V3<real> p, const V3<real> foot[2], const real yc[2],
                                        const real ys[2], int contact_mask, real fx, real fy, real mu_b, const real Minv[6]) {
    if constexpr (G == 8 && NCOLS == 40) {
      // A lane's five columns share the foot f and the friction direction d and differ only in the site offset o_k in the foot's
      // yaw frame (e1, e2): site_k = foot + ox_k e1 + oy_k e2, so the torque rows are affine in (ox_k, oy_k):
      //   R^T((site_k - p) x B) = T0 + ox_k T1 + oy_k T2,  T0 = R^T((foot - p) x B), T1 = R^T(e1 x B), T2 = R^T(e2 x B)
      // -- three cross products and rotations per lane instead of five.
      const int f = (lane >> 2) & 1, d = lane & 3;
      const real bx = d == 0 ? mu_b : (d == 1 ? -mu_b : real(0));
      const real by = d == 2 ? mu_b : (d == 3 ? -mu_b : real(0));
      const V3<real> B = {fx * bx - fy * by, fy * bx + fx * by, real(1)};
      const V3<real> ft = f ? foot[1] : foot[0];
      const real c_ = f ? yc[1] : yc[0], s_ = f ? ys[1] : ys[0];
      const real m = ((contact_mask >> f) & 1) ? real(1) : real(0);
      const real w3 = m * Minv[3], w4 = m * Minv[4], w5 = m * Minv[5];
      V3<real> T0 = mulT(R, cross(ft - p, B));
      V3<real> T1 = mulT(R, V3<real>{s_, -c_, c_ * B.y - s_ * B.x});           // e1 = (c, s, 0):  e1 x B
      V3<real> T2 = mulT(R, V3<real>{c_, s_, -s_ * B.y - c_ * B.x});           // e2 = (-s, c, 0): e2 x B
      T0 = {w3 * T0.x, w4 * T0.y, w5 * T0.z}; T1 = {w3 * T1.x, w4 * T1.y, w5 * T1.z}; T2 = {w3 * T2.x, w4 * T2.y, w5 * T2.z};
      const real l0 = m * Minv[0] * B.x, l1 = m * Minv[1] * B.y, l2 = m * Minv[2] * B.z;
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const real ox = k == 0 ? real(0) : ((k == 1 || k == 2) ? real(0.12) : real(-0.12));
        const real oy = k == 0 ? real(0) : ((k == 1 || k == 3) ? real(0.06) : real(-0.06));
        a[k][0] = l0; a[k][1] = l1; a[k][2] = l2;
        a[k][3] = T0.x + ox * T1.x + oy * T2.x; a[k][4] = T0.y + ox * T1.y + oy * T2.y; a[k][5] = T0.z + ox * T1.z + oy * T2.z;
      }
      return;
    }
    if constexpr (SHARED_LIN4) {
      // same affine form as above, per foot:  tau_k = T0_f + ox_k T1_f + oy_k T2_f,  f = k & 1, site = k >> 1
      const int d = lane & 3;
      const real bx = d == 0 ? mu_b : (d == 1 ? -mu_b : real(0));
      const real by = d == 2 ? mu_b : (d == 3 ? -mu_b : real(0));
      const V3<real> B = {fx * bx - fy * by, fy * bx + fx * by, real(1)};
      bl[0] = Minv[0] * B.x; bl[1] = Minv[1] * B.y; bl[2] = Minv[2] * B.z;
      V3<real> T0[2], T1[2], T2[2];
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const real m = ((contact_mask >> f) & 1) ? real(1) : real(0);
        mf[f] = m;
        const real w3 = m * Minv[3], w4 = m * Minv[4], w5 = m * Minv[5];
        const real c_ = yc[f], s_ = ys[f];
        V3<real> t0 = mulT(R, cross(foot[f] - p, B));
        V3<real> t1 = mulT(R, V3<real>{s_, -c_, c_ * B.y - s_ * B.x});
        V3<real> t2 = mulT(R, V3<real>{c_, s_, -s_ * B.y - c_ * B.x});
        T0[f] = {w3 * t0.x, w4 * t0.y, w5 * t0.z}; T1[f] = {w3 * t1.x, w4 * t1.y, w5 * t1.z}; T2[f] = {w3 * t2.x, w4 * t2.y, w5 * t2.z};
      }
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const int f = k & 1, site = k >> 1;
        const real ox = site == 0 ? real(0) : ((site == 1 || site == 2) ? real(0.12) : real(-0.12));
        const real oy = site == 0 ? real(0) : ((site == 1 || site == 3) ? real(0.06) : real(-0.06));
        a[k][0] = mf[f] * bl[0]; a[k][1] = mf[f] * bl[1]; a[k][2] = mf[f] * bl[2];
        a[k][3] = T0[f].x + ox * T1[f].x + oy * T2[f].x; a[k][4] = T0[f].y + ox * T1[f].y + oy * T2[f].y; a[k][5] = T0[f].z + ox * T1[f].z + oy * T2[f].z;
      }
      return;
    }
#pragma unroll
    for (int k = 0; k < NC; k++) {
      int c = lane + k * G;
      int sidx = c >> 3, f = (c >> 2) & 1, d = c & 3;
      bool valid = (c < 40) && ((contact_mask >> f) & 1);
      real bx = d == 0 ? mu_b : (d == 1 ? -mu_b : real(0));
      real by = d == 2 ? mu_b : (d == 3 ? -mu_b : real(0));
      V3<real> B = {fx * bx - fy * by, fy * bx + fx * by, real(1)};     // curr_projSRB_ori @ b_d
      real ox = sidx == 0 ? real(0) : ((sidx == 1 || sidx == 2) ? real(0.12) : real(-0.12));
      real oy = sidx == 0 ? real(0) : ((sidx == 1 || sidx == 3) ? real(0.06) : real(-0.06));
      V3<real> ft = f ? foot[1] : foot[0];
      real c_ = f ? yc[1] : yc[0], s_ = f ? ys[1] : ys[0];
      V3<real> site = {ft.x + (c_ * ox - s_ * oy), ft.y + (s_ * ox + c_ * oy), ft.z};   // set_feet_site_this_step_env
      V3<real> r = site - p;
      V3<real> tau = mulT(R, cross(r, B));                              // jacp^T B, rotational rows (body frame)
      real m = valid ? real(1) : real(0);
      a[k][0] = m * Minv[0] * B.x; a[k][1] = m * Minv[1] * B.y; a[k][2] = m * Minv[2] * B.z;
      a[k][3] = m * Minv[3] * tau.x; a[k][4] = m * Minv[4] * tau.y; a[k][5] = m * Minv[5] * tau.z;
    }
  }

  // G == 8 with the SRBTrack column order c = site*8 + foot*4 + dir: a lane's five columns are the five sites of ONE foot
  // and ONE friction direction, so their linear parts (Minv B, rows 0..2) are identical and only the torque parts differ.
  static constexpr bool SHARED_LIN = (G == 8 && NCOLS == 40);
  __device__ __forceinline__ void tvals(const real nu[6], real t[NC]) const {
    if constexpr (SHARED_LIN) {
      const real base = a[0][0] * nu[0] + a[0][1] * nu[1] + a[0][2] * nu[2];
#pragma unroll
      for (int k = 0; k < NC; k++) t[k] = base + a[k][3] * nu[3] + a[k][4] * nu[4] + a[k][5] * nu[5];
    } else if constexpr (SHARED_LIN4) {
      const real base = bl[0] * nu[0] + bl[1] * nu[1] + bl[2] * nu[2];
      const real bf[2] = {mf[0] * base, mf[1] * base};
#pragma unroll
      for (int k = 0; k < NC; k++) t[k] = bf[k & 1] + a[k][3] * nu[3] + a[k][4] * nu[4] + a[k][5] * nu[5];
    } else {
#pragma unroll
      for (int k = 0; k < NC; k++)
        t[k] = a[k][0] * nu[0] + a[k][1] * nu[1] + a[k][2] * nu[2] + a[k][3] * nu[3] + a[k][4] * nu[4] + a[k][5] * nu[5];
    }
  }
  __device__ __forceinline__ real phi(const real nu[6], const real y[6], const real t[NC], real inv2w) const {
    real ps = 0;
#pragma unroll
    for (int k = 0; k < NC; k++) { real m = t[k] < 0 ? t[k] : real(0); ps += m * m; }
    ps = gsum<G>(ps);
    real q = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) q += nu[i] * (real(0.5) * nu[i] + y[i]);
    return q + inv2w * ps;
  }

  // H (lower triangle, 21) = I + (1/w) sum_{k: t_k<0} a_k a_k^T, all-reduced over the group
  __device__ __forceinline__ void hessian(const real t[NC], real invw, real H[21], real* sm) const {
    if constexpr (SHARED_LIN || SHARED_LIN4) {
      // sum_k m_k [b; tau_k][b; tau_k]^T = [[n b b^T, b tsum^T], [tsum b^T, sum m_k tau_k tau_k^T]]   (80 instead of 135 ops)
      // (G == 4: a column of a foot that is not in contact has t = 0 and is never active, so the unmasked b serves both feet)
      real n = 0, ts0 = 0, ts1 = 0, ts2 = 0, aa[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const real m = t[k] < 0 ? real(1) : real(0);
        const real t0 = m * a[k][3], t1 = m * a[k][4], t2 = m * a[k][5];
        n += m; ts0 += t0; ts1 += t1; ts2 += t2;
        aa[0] += t0 * a[k][3]; aa[1] += t1 * a[k][3]; aa[2] += t1 * a[k][4]; aa[3] += t2 * a[k][3]; aa[4] += t2 * a[k][4]; aa[5] += t2 * a[k][5];
      }
      const real b0 = SHARED_LIN4 ? bl[0] : a[0][0], b1 = SHARED_LIN4 ? bl[1] : a[0][1], b2 = SHARED_LIN4 ? bl[2] : a[0][2];
      const real nb0 = n * b0, nb1 = n * b1, nb2 = n * b2;
      H[0] = nb0 * b0; H[1] = nb1 * b0; H[2] = nb1 * b1; H[3] = nb2 * b0; H[4] = nb2 * b1; H[5] = nb2 * b2;
      H[6] = ts0 * b0; H[7] = ts0 * b1; H[8] = ts0 * b2; H[9] = aa[0];
      H[10] = ts1 * b0; H[11] = ts1 * b1; H[12] = ts1 * b2; H[13] = aa[1]; H[14] = aa[2];
      H[15] = ts2 * b0; H[16] = ts2 * b1; H[17] = ts2 * b2; H[18] = aa[3]; H[19] = aa[4]; H[20] = aa[5];
    } else {
#pragma unroll
    for (int i = 0; i < 21; i++) H[i] = 0;
#pragma unroll
    for (int k = 0; k < NC; k++) {
      real m = t[k] < 0 ? real(1) : real(0);
#pragma unroll
      for (int i = 0; i < 6; i++) {
        real ai = m * a[k][i];
#pragma unroll
        for (int j = 0; j <= i; j++) H[i * (i + 1) / 2 + j] += ai * a[k][j];
      }
    }
    }
    HReduce<real, G>::run(H, sm, lane_id());
#pragma unroll
    for (int i = 0; i < 21; i++) H[i] *= invw;
#pragma unroll
    for (int i = 0; i < 6; i++) H[i * (i + 1) / 2 + i] += real(1);
  }

  // returns number of Newton systems solved; nu in: initial guess, out: solution; t out: A^T nu.
  // Semismooth Newton with full steps; a step is accepted when it reproduces the active set (exact optimum of
  // the piece), is below rounding level, or decreases the dual objective phi; otherwise it is halved until phi
  // decreases (globally convergent; also keeps the iteration-count tail short, which matters because the
  // environments of a warp iterate in lock-step).
  // S0 (optional): guess of the optimal active set (bit k = column k of this lane active), e.g. the active set the previous
  // physics sub-step ended with.  The first Newton system is then built from S0 instead of from sign(A^T nu): when the guess
  // is right the solve ends after ONE system (its solution reproduces S0 = exact optimum of that piece), otherwise the
  // iteration continues from the usual state.  Acceptance tests are unchanged, so the result is the same exact optimum.
  __device__ __forceinline__ int solve(int lane, const real y[6], real inv_w, real nu[6], real t[NC], real* sm, const unsigned* S0 = nullptr) {
    constexpr int MAX_IT = 48;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane_id() / G) * G));
    const real invw = inv_w, inv2w = real(0.5) * invw;
    tvals(nu, t);
    real ph = phi(nu, y, t, inv2w);
    bool done = false;
    int iters = 0;
    real rhs[6];
#pragma unroll
    for (int i = 0; i < 6; i++) rhs[i] = -y[i];
    real ts[NC];                                              // signs the next Hessian is built from / the new set is compared with
#pragma unroll
    for (int k = 0; k < NC; k++) ts[k] = S0 != nullptr ? (((*S0 >> k) & 1u) ? real(-1) : real(1)) : t[k];
    for (int it = 0; it < MAX_IT; it++) {
      if (!__any_sync(0xffffffffu, !done)) break;
      real H[21];
      hessian(ts, invw, H, sm);
      real nn[6], tn[NC];
      ldl_solve6(H, rhs, nn);
      tvals(nn, tn);
      bool same = true;
#pragma unroll
      for (int k = 0; k < NC; k++) same = same && ((tn[k] < 0) == (ts[k] < 0));
      unsigned bal = __ballot_sync(0xffffffffu, same);
      same = (bal & gmask) == gmask;
      // The acceptance tests (phi of the new point, rounding-level step, step halving) matter only where the active set changed;
      // the last system of a solve -- every unfinished environment of the warp reproduced its set -- skips them.
      real phn = ph;
      if (__any_sync(0xffffffffu, !done && !same)) {
        phn = phi(nn, y, tn, inv2w);
        // a step below rounding level is a fixed point (needed in fp32, where the sign of a ~0 t_j can flicker): |step|_2 <= eps (1 + |nu|_2)
        real d2 = 0, n2 = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) { const real dd = nn[i] - nu[i]; d2 += dd * dd; n2 += nn[i] * nn[i]; }
        // (with a guessed set the fixed-point shortcut is only valid once the Hessian came from sign(A^T nu) itself)
        if (S0 == nullptr || it > 0) same = same || (d2 <= num<real>::eps_conv() * num<real>::eps_conv() * (real(1) + n2));
        const real ph_tol = real(4) * num<real>::eps_conv() * (real(1) + (ph < 0 ? -ph : ph));   // phi resolution
        bool need_bt = !done && !same && !(phn < ph + ph_tol);
        if (__any_sync(0xffffffffu, need_bt)) {                 // rare
          real alpha = real(0.5);
          while (__any_sync(0xffffffffu, need_bt)) {
            real nt[6], tt[NC];
#pragma unroll
            for (int i = 0; i < 6; i++) nt[i] = nu[i] + alpha * (nn[i] - nu[i]);
            tvals(nt, tt);
            real pht = phi(nt, y, tt, inv2w);
            if (need_bt && (pht < ph + ph_tol || alpha < real(1e-3))) {
#pragma unroll
              for (int i = 0; i < 6; i++) nn[i] = nt[i];
#pragma unroll
              for (int k = 0; k < NC; k++) tn[k] = tt[k];
              phn = pht;
              need_bt = false;
            }
            alpha *= real(0.5);
          }
        }
      }
      if (!done) {
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = nn[i];
#pragma unroll
        for (int k = 0; k < NC; k++) ts[k] = tn[k];
        ph = phn;
        iters++;
        if (same) done = true;
      }
    }
#pragma unroll
    for (int k = 0; k < NC; k++) t[k] = ts[k];                // A^T nu of the accepted iterate
    return iters;
  }
  // Same Newton iteration, globalised by an EXACT line search instead of step halving (the finite Newton method of
  // Mangasarian / Keerthi-DeCoste for this class of piecewise-quadratic problems): along d = nn - nu the derivative
  //   g(alpha) = d.(nu + y) + alpha |d|^2 + (1/w) sum_j min(0, t_j + alpha s_j) s_j ,   s = A^T d,
  // is increasing and piecewise linear with one kink per column; its root in (0,1) is found by a bracketed Newton
  // iteration that stops as soon as a Newton step stays on one linear piece.  phi decreases monotonically, so the
  // active sets cannot cycle.  Needed when 1/w is huge (gym_cdm2: 2e8), where step halving never reaches the
  // descent region; typical cost: <3 Newton systems and ~8 evaluations of g per solve.
  __device__ __forceinline__ int solve_ls(int lane, const real y[6], real inv_w, real nu[6], real t[NC], real* sm) {
    constexpr int MAX_IT = 32;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane_id() / G) * G));
    const real invw = inv_w;
    tvals(nu, t);
    bool done = false;
    int iters = 0;
    real rhs[6];
#pragma unroll
    for (int i = 0; i < 6; i++) rhs[i] = -y[i];
    for (int it = 0; it < MAX_IT; it++) {
      if (!__any_sync(0xffffffffu, !done)) break;
      real H[21];
      hessian(t, invw, H, sm);
      real nn[6], tn[NC], s[NC];
      ldl_solve6(H, rhs, nn);
      tvals(nn, tn);
      bool same = true;
      real a1 = 0;
#pragma unroll
      for (int k = 0; k < NC; k++) {
        same = same && ((tn[k] < 0) == (t[k] < 0));
        s[k] = tn[k] - t[k];
        a1 += tn[k] < 0 ? tn[k] * s[k] : real(0);
      }
      same = (__ballot_sync(0xffffffffu, same) & gmask) == gmask;
      real dd = 0, b0 = 0, nmax = 0;
#pragma unroll
      for (int i = 0; i < 6; i++) {
        real d = nn[i] - nu[i];
        dd += d * d; b0 += d * (nu[i] + y[i]);
        real av = nn[i] < 0 ? -nn[i] : nn[i];
        nmax = av > nmax ? av : nmax;
      }
      const real g1 = b0 + dd + invw * gsum<G>(a1);
      real alpha = real(1);
      bool ls = !done && !same && g1 > 0;
      if (__any_sync(0xffffffffu, ls)) {
        real al = 0, lo = 0, hi = 1;
        real p1 = 0, p2 = 0;
        unsigned ml = 0;
#pragma unroll
        for (int k = 0; k < NC; k++) if (t[k] < 0) { p1 += t[k] * s[k]; p2 += s[k] * s[k]; ml |= 1u << k; }
        real gl = b0 + invw * gsum<G>(p1), sl = dd + invw * gsum<G>(p2);
        if (ls && !(gl < 0)) { ls = false; alpha = 0; }                    // no descent left: optimum to rounding
        for (int k2 = 0; k2 < 48; k2++) {
          if (!__any_sync(0xffffffffu, ls)) break;
          real cand = al - gl / sl;
          const bool newton = cand > lo && cand < hi;
          if (!newton) cand = real(0.5) * (lo + hi);
          real q1 = 0, q2 = 0;
          unsigned mc = 0;
#pragma unroll
          for (int k = 0; k < NC; k++) {
            real r = t[k] + cand * s[k];
            if (r < 0) { q1 += r * s[k]; q2 += s[k] * s[k]; mc |= 1u << k; }
          }
          const real gc = b0 + cand * dd + invw * gsum<G>(q1), sc = dd + invw * gsum<G>(q2);
          const bool piece = (__ballot_sync(0xffffffffu, mc == ml) & gmask) == gmask;
          if (ls) {
            if (newton && piece) { al = cand; ls = false; }
            else {
              if (gc < 0) lo = cand; else hi = cand;
              al = cand; gl = gc; sl = sc; ml = mc;
              if (hi - lo <= real(4) * num<real>::eps_conv() * hi) ls = false;
            }
            alpha = al;
          }
        }
      }
      if (!done) {
        real stepn = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) { real d = alpha * (nn[i] - nu[i]); nu[i] = same ? nn[i] : nu[i] + d; d = d < 0 ? -d : d; stepn = d > stepn ? d : stepn; }
#pragma unroll
        for (int k = 0; k < NC; k++) t[k] = same ? tn[k] : t[k] + alpha * s[k];
        iters++;
        if (same || stepn <= num<real>::eps_conv() * (real(1) + nmax)) done = true;
      }
    }
    return iters;
  }
  static __device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
};

// ------------------------------------------------------------------------------------------------
// SDS swing-foot LQR filter (work/controlmodule.py:962-1086, gains: useKpreset :1041-1057 with
// sop.piecewiseLinearMap (module.lua:5203) and matrixn::sampleRow (matrixn.cpp:781, float32 `t`)).
__constant__ double SDS_KP[7][2] = {{262.685, 127.410}, {942.739, 281.657}, {3103.828, 553.238}, {9941.174, 1033.866},
                                    {31563.832, 1887.117}, {99941.017, 3403.601}, {316168.772, 6099.859}};
__constant__ double SDS_KT[8] = {1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12};
template <typename real> __device__ __noinline__ void sds_gain_slow(real speed, real* K0, real* K1);
template <typename real> __device__ __forceinline__ void sds_gain(real speed, real* K0, real* K1) {
  if (speed < real(3)) { *K0 = real(31563.832); *K1 = real(1887.117); return; }     // logQ = 9 -> row 4 exactly
  sds_gain_slow(speed, K0, K1);
}
template <typename real> __device__ __noinline__ void sds_gain_slow(real speed, real* K0, real* K1) {
  const double (*KP)[2] = SDS_KP;
  double logQ = speed > real(6) ? 11.0 : 9.0 + (((double)speed - 3.0) / 3.0) * (11.0 - 9.0);
  logQ = fmax(9.0, logQ);
  double Q = ::pow(10.0, logQ);
  // piecewiseLinearMap(Q, 10^5..10^12, 0..7)
  const double* KT = SDS_KT;
  double index = 7.0;
  for (int i = 0; i < 7; i++) {
    double t1 = KT[i + 1];
    double d = t1 - KT[i];
    if (t1 >= Q - 1e-10) {
      double lo = t1 - d;
      double wgt = (Q - lo) / (t1 - lo);
      wgt = 0.0 * (1.0 - wgt) + 1.0 * wgt;
      index = (double)i * (1.0 - wgt) + (double)(i + 1) * wgt;
      break;
    }
  }
  int a = (int)::floor(index);
  float t = (float)(index - (double)(float)a);
  double k0, k1;
  if (t < 0.005) { k0 = KP[a][0]; k1 = KP[a][1]; }
  else if (t > 0.995) { k0 = KP[min(a + 1, 6)][0]; k1 = KP[min(a + 1, 6)][1]; }
  else if (a + 1 >= 7) { double tt = (double)t + 1.0; k0 = KP[a - 1][0] * (1.0 - tt) + KP[a][0] * tt; k1 = KP[a - 1][1] * (1.0 - tt) + KP[a][1] * tt; }
  else { double tt = (double)t; k0 = KP[a][0] * (1.0 - tt) + KP[a + 1][0] * tt; k1 = KP[a][1] * (1.0 - tt) + KP[a + 1][1] * tt; }
  *K0 = (real)k0; *K1 = (real)k1;
}

template <typename real> __device__ __forceinline__ void sds_4steps(real& x, real& v, real xd, real K0, real K1) {
  const real invM = real(1) / real(60.0), dt = real(1) / real(120.0);
  const real C = real(0.001) * invM, Gc = real(1) * invM, H = real(1) * invM;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    real U = K0 * xd + K1 * real(0);
    real cf = U - (K0 * x + K1 * v);
    real dd = H * cf - Gc - C * v;
    v += dd * dt;
    x += v * dt;
  }
}

// ------------------------------------------------------------------------------------------------
// reward (SRBEnv._get_reward, SRBEnv5_mocap_list_window.py:496-803); returns r_t and the 8 info terms
template <typename real> struct RefFrame { V3<real> pos; M3<real> R; };
template <typename real> __device__ __forceinline__ RefFrame<real> load_ref(const ClipDesc<real>& cd, int f) {
  f = min(max(f, 0), cd.nframes - 1);
  const real* fr = cd.frames + (size_t)f * CLIP_W;
  RefFrame<real> o;
  o.pos = {fr[0], fr[1], fr[2]};
#pragma unroll
  for (int i = 0; i < 9; i++) o.R.m[i] = fr[3 + i];
  return o;
}
// L1 norm of the 6-D rotation difference  |(Pa^T Ra)[:, :2] - (Pb^T Rb)[:, :2]|_1 for yaw frames Pa, Pb
template <typename real> __device__ __forceinline__ real rot6_l1_yaw(real ax, real ay, const M3<real>& Ra, real bx, real by, const M3<real>& Rb) {
  real s = 0;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    real a0 = ax * Ra.m[c] + ay * Ra.m[3 + c], a1 = -ay * Ra.m[c] + ax * Ra.m[3 + c], a2 = Ra.m[6 + c];
    real b0 = bx * Rb.m[c] + by * Rb.m[3 + c], b1 = -by * Rb.m[c] + bx * Rb.m[3 + c], b2 = Rb.m[6 + c];
    s += num<real>::abs(a0 - b0) + num<real>::abs(a1 - b1) + num<real>::abs(a2 - b2);
  }
  return s;
}
template <typename real> __device__ __forceinline__ real rot6_l1_full(const M3<real>& Pa, const M3<real>& Ra, const M3<real>& Pb, const M3<real>& Rb) {
  M3<real> A = matmulTN(Pa, Ra), B = matmulTN(Pb, Rb);
  real s = 0;
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) s += num<real>::abs(A.m[3 * r + c] - B.m[3 * r + c]);
  return s;
}

template <typename real, int G>
__device__ __forceinline__ real compute_reward(const EnvState<real>& s, const ClipDesc<real>& cd, const real* __restrict__ hist_env,
                                               int pred_contact, const Params<real>& p, float info[8], int lane) {
  const int cf = s.frame, f0 = cf + s.rif;
  M3<real> X = quat2mat_normalized(s.qlag);            // data.xmat (stale by one sub-step)
  real cx, cy;
  yaw_frame(X, &cx, &cy);
  RefFrame<real> cur = load_ref(cd, f0), nxt = load_ref(cd, f0 + 1);
  real mcx, mcy, mnx, mny;
  yaw_frame(cur.R, &mcx, &mcy);
  yaw_frame(nxt.R, &mnx, &mny);
  const real* nfr = cd.frames + (size_t)min(f0 + 1, cd.nframes - 1) * CLIP_W;

  // contact timing
  int mocap_next = (nfr[20] == real(0) ? 1 : 0) + (nfr[23] == real(0) ? 2 : 0);
  real r_c = (pred_contact == mocap_next) ? real(0) : real(10);

  // end-effector term
  real r_e = 0;
  {
    V3<real> origin = {s.pos.x, s.pos.y, real(0)};
    V3<real> morigin = {nxt.pos.x, nxt.pos.y, real(0)};
#pragma unroll
    for (int i = 0; i < 2; i++) {
      V3<real> ref_xy = {nfr[18 + 3 * i], nfr[19 + 3 * i], real(0)};
      V3<real> dv = yawT(cx, cy, s.foot[i] - origin) - yawT(mnx, mny, ref_xy - morigin);
      bool on = (pred_contact == 3) || (pred_contact == (i == 0 ? 1 : 2));
      real nrm = norm(dv);
      r_e += (on ? real(1.9) : real(0.1)) * (nrm * nrm);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
      // (proj^T foot_ori)[:, :2] vs (mproj^T ref_foot_ori)[:, :2]; both foot matrices are pure yaw
      real c_ = s.yc[i], s_ = s.ys[i], rc = nfr[24 + 2 * i], rs = nfr[25 + 2 * i];
      real a00 = cx * c_ + cy * s_, a10 = -cy * c_ + cx * s_, a01 = cx * (-s_) + cy * c_, a11 = -cy * (-s_) + cx * c_;
      real b00 = mnx * rc + mny * rs, b10 = -mny * rc + mnx * rs, b01 = mnx * (-rs) + mny * rc, b11 = -mny * (-rs) + mnx * rc;
      r_e += num<real>::abs(a00 - b00) + num<real>::abs(a10 - b10) + num<real>::abs(a01 - b01) + num<real>::abs(a11 - b11);
    }
  }

  // delta terms vs history at -1, -5, -15 steps
  real dp = 0, dR = 0;
  auto hist_entry = [&](int e, V3<real>* hp, M3<real>* hR) {
    const real* h = hist_env + (size_t)(e & (HIST_SLOTS - 1)) * HIST_W;
    *hp = {h[0], h[1], h[2]};
    *hR = quat2mat_normalized(Q4<real>{h[3], h[4], h[5], h[6]});
  };
  if constexpr (G >= 4) {
    // The three history terms are the same computation on different data (entry cf - {0, 5, 15}): lanes 0..2 of the group take one
    // each (the other lanes repeat lane 0's), the sums are gathered by shuffles.  The -15 term projects with the full history
    // matrices, the other two with their yaw frames; a yaw frame written as a matrix makes that one code path too.
    const int j = lane < 3 ? lane : 0;
    const int back = j == 0 ? 0 : (j == 1 ? 5 : 15);
    const bool on = cf >= back;
    V3<real> hp; M3<real> hR;
    hist_entry(on ? cf - back : cf, &hp, &hR);
    RefFrame<real> m = load_ref(cd, on ? f0 - back : f0);
    real hx, hy, mx, my;
    yaw_frame(hR, &hx, &hy);
    yaw_frame(m.R, &mx, &my);
    real dpj = norm(yawT(hx, hy, s.pos - hp) - yawT(mx, my, nxt.pos - m.pos));
    M3<real> Pa = hR, Pb = m.R;
    if (back != 15) {
      Pa.m[0] = hx; Pa.m[1] = -hy; Pa.m[2] = real(0); Pa.m[3] = hy; Pa.m[4] = hx; Pa.m[5] = real(0); Pa.m[6] = real(0); Pa.m[7] = real(0); Pa.m[8] = real(1);
      Pb.m[0] = mx; Pb.m[1] = -my; Pb.m[2] = real(0); Pb.m[3] = my; Pb.m[4] = mx; Pb.m[5] = real(0); Pb.m[6] = real(0); Pb.m[7] = real(0); Pb.m[8] = real(1);
    }
    real dRj = rot6_l1_full(Pa, X, Pb, nxt.R);
    if (!on) { dpj = real(0); dRj = real(0); }
    const int base = (threadIdx.x & 31) & ~(G - 1);
    dp = __shfl_sync(0xffffffffu, dpj, base + 2) + __shfl_sync(0xffffffffu, dpj, base + 1) + __shfl_sync(0xffffffffu, dpj, base);
    dR = __shfl_sync(0xffffffffu, dRj, base + 2) + __shfl_sync(0xffffffffu, dRj, base + 1) + __shfl_sync(0xffffffffu, dRj, base);
  } else {
    if (cf >= 15) {
      V3<real> hp; M3<real> hR;
      hist_entry(cf - 15, &hp, &hR);
      RefFrame<real> m15 = load_ref(cd, f0 - 15);
      real hx, hy, mx, my;
      yaw_frame(hR, &hx, &hy);
      yaw_frame(m15.R, &mx, &my);
      dp += norm(yawT(hx, hy, s.pos - hp) - yawT(mx, my, nxt.pos - m15.pos));
      dR += rot6_l1_full(hR, X, m15.R, nxt.R);
    }
    if (cf >= 5) {
      V3<real> hp; M3<real> hR;
      hist_entry(cf - 5, &hp, &hR);
      RefFrame<real> m5 = load_ref(cd, f0 - 5);
      real hx, hy, mx, my;
      yaw_frame(hR, &hx, &hy);
      yaw_frame(m5.R, &mx, &my);
      dp += norm(yawT(hx, hy, s.pos - hp) - yawT(mx, my, nxt.pos - m5.pos));
      dR += rot6_l1_yaw(hx, hy, X, mx, my, nxt.R);
    }
    {
      V3<real> hp; M3<real> hR;
      hist_entry(cf, &hp, &hR);
      real hx, hy;
      yaw_frame(hR, &hx, &hy);
      dp += norm(yawT(hx, hy, s.pos - hp) - yawT(mcx, mcy, nxt.pos - cur.pos));
      dR += rot6_l1_yaw(hx, hy, X, mcx, mcy, nxt.R);
    }
  }
  // absolute terms
  V3<real> pa = {real(0), real(0), s.pos.z}, pb = {real(0), real(0), nxt.pos.z};
  real p_term = norm(yawT(cx, cy, pa) - yawT(mnx, mny, pb));
  real R_term = rot6_l1_yaw(cx, cy, X, mnx, mny, nxt.R);

  const real w_s = p.w[0], w_m = p.w[1], w_p = p.w[2], w_e = p.w[3], w_p1 = p.w[4], w_p2 = p.w[5], w_p3 = p.w[6], w_p4 = p.w[7], w_c = p.w[8];
  real r_p = w_p1 * dp + w_p2 * dR + w_p3 * p_term + w_p4 * R_term;
  real r_m = w_p * r_p + w_e * r_e + w_c * r_c;
  real r_t = w_s * real(1) - w_m * r_m;
  info[0] = (float)(w_s * real(1));
  info[1] = (float)(-w_m * w_c * r_c);
  info[2] = (float)(-w_m * w_e * r_e);
  info[3] = (float)(-w_m * w_p * w_p1 * dp);
  info[4] = (float)(-w_m * w_p * w_p2 * dR);
  info[5] = (float)(-w_m * w_p * w_p3 * p_term);
  info[6] = (float)(-w_m * w_p * w_p4 * R_term);
  info[7] = (float)r_t;
  return r_t;
}

// SRBEnv._get_done (:916-1003)
template <typename real> __device__ __forceinline__ bool compute_done(const EnvState<real>& s, const ClipDesc<real>& cd, const Params<real>& p,
                                                                      const TerrainDev* T = nullptr) {
  int f1 = min(s.frame + 1 + s.rif, cd.nframes - 1);
  real ref_h = cd.frames[(size_t)f1 * CLIP_W + 2];
  if (T != nullptr) {
    const real* fr = cd.frames + (size_t)f1 * CLIP_W;
    ref_h += terrain_query<real>(*T, fr[0] + s.planar[0], fr[1] + s.planar[1], nullptr);
  }
  real h = s.pos.z;
  if (ref_h - h > real(0.2) || h - ref_h > real(0.2) || h < real(0.50) || h > real(1.2)) return true;
  M3<real> X = quat2mat_normalized(s.qlag);
  real zn = num<real>::sqrt(X.m[2] * X.m[2] + X.m[5] * X.m[5] + X.m[8] * X.m[8]);
  real ca = X.m[8] / (zn * real(1));
  ca = num<real>::fmin(real(1), num<real>::fmax(real(-1), ca));
  real deg = num<real>::acos(ca) * real(57.29577951308232);
  if (deg > real(65)) return true;
  return s.frame >= p.max_frame - 1;
}

// ------------------------------------------------------------------------------------------------
// the step kernel
// CIO (coalesced I/O): actions in and observations / reward / done / reward_info out are staged through shared memory so
// that every access of the I/O arrays is one contiguous, 16-byte-vectorised block per warp (the rows of a warp's
// environments are adjacent).  Used by the host-buffer path (srb_step_host), where those arrays live in mapped pinned
// host memory and 32-byte strided stores would crawl over PCIe (measured: +72 us per 4096-env step vs +34 us).
// TERRAIN: the terrain variant (heightfield / pyramid queries, lifted references) is a separate instantiation, so the flat
// kernels carry none of that code: 192 kB -> 104 kB of SASS, -5 % step time (the instruction caches hold 6 + 32 kB).
template <typename real, int G, bool CIO = false, bool TERRAIN = false>
__global__ void __launch_bounds__(32) srb_step_kernel(const StepArgs<real> A) {
  constexpr int EPW = 32 / G;
  long long tk_entry = 0;                                  // phase clocks: only with a debug buffer attached (tools/phase_cycles.py)
  unsigned long long gt_entry = 0;
  if (A.dbg != nullptr) {
    tk_entry = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
  }
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 % G;
  const int e_raw = (blockIdx.x * EPW) + lane32 / G;
  const bool active = e_raw < A.N;
  const int e = active ? e_raw : A.N - 1;
  const Params<real>& P = A.p;

  __shared__ __align__(16) real hsm[HReduce<real, G>::SMEM_WORDS];
  float af[ACT_W];                                         // issued before the state loads: one DRAM round trip for both
  if (!CIO) {
    const float2* act2 = reinterpret_cast<const float2*>(A.act + (size_t)e * ACT_W);
#pragma unroll
    for (int i = 0; i < ACT_W / 2; i++) { const float2 v = __ldg(act2 + i); af[2 * i] = v.x; af[2 * i + 1] = v.y; }
  }
  EnvState<real> s;
  load_state(s, A.st, A.ist, A.Npad, e);
  const ClipDesc<real> cd = A.clips[s.clip];
  real* hist_env = A.hist + (size_t)e * HIST_SLOTS * HIST_W;
  __shared__ __align__(16) float sio[CIO ? EPW * OBS_W : 4];
  const int env0 = blockIdx.x * EPW;
  const int nrows = min(EPW, A.N - env0);
  if (CIO) {
    for (int i = lane32; i < nrows * ACT_W; i += 32) sio[i] = __ldg(A.act + (size_t)env0 * ACT_W + i);
    __syncwarp();
  }
  // Everything the post-step phase reads (history entries, reference frames, look-ahead features) is known now:
  // prefetch those lines into L1 so their DRAM latency overlaps the four sub-steps instead of stalling the reward.
  {
    const int f0 = s.frame + s.rif;
    for (int id = lane; id < 16; id += G) {
      const char* ptr;
      if (id < 3) {
        int back = id == 0 ? 0 : (id == 1 ? 5 : 15);
        ptr = (const char*)(hist_env + (size_t)((s.frame - back) & (HIST_SLOTS - 1)) * HIST_W);
      } else if (id < 11) {
        int k = (id - 3) >> 1;                              // two 128-byte lines per record (fp64 records are 256 B)
        int fr = k == 0 ? f0 : (k == 1 ? f0 + 1 : (k == 2 ? f0 - 5 : f0 - 15));
        fr = min(max(fr, 0), cd.nframes - 1);
        ptr = (const char*)(cd.frames + (size_t)fr * CLIP_W) + ((id - 3) & 1) * 128;
        if (sizeof(real) == 4 && ((id - 3) & 1)) ptr -= 128;
      } else {
        int fw = min(f0 + 1 + (P.obs_after_inc ? 1 : 0), max(cd.nframes - 5, 0));      // (a prefetch hint only)
        ptr = (const char*)(cd.feat + (size_t)fw * FEAT_W) + (id - 11) * 128;
      }
      asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
    }
  }
  real a[ACT_W];
  if (CIO) {
    const float* sa = sio + (active ? lane32 / G : nrows - 1) * ACT_W;
#pragma unroll
    for (int i = 0; i < ACT_W; i++) a[i] = (real)sa[i];
    __syncwarp();                                        // sio is reused for the observation rows below
  } else {
#pragma unroll
    for (int i = 0; i < ACT_W; i++) a[i] = (real)af[i];
  }
  // ---- frames from the (stale) xmat, action decoding (SRBEnv.step :209-231)
  const M3<real> R0 = quat2mat_normalized(s.qlag);
  real fx, fy;
  yaw_frame(R0, &fx, &fy);
  const V3<real> origin = {s.pos.x, s.pos.y, real(0)};
  const V3<real> tv_lin = mul(R0, V3<real>{a[4], a[5], a[6]});
  const V3<real> tv_ang = {a[7], a[8], a[9]};
  const V3<real> p_hat = mul(R0, V3<real>{a[10], a[11], a[12]}) + s.pos;
  const M3<real> R_hat = matmul(R0, exp_so3(V3<real>{a[13], a[14], a[15]}));

  // ---- contact decision (:234-311): first arg-max of the one-hot, or policy-probability hysteresis
  int sig = 0;
  {
    real best = a[18];
#pragma unroll
    for (int i = 1; i < 4; i++) if (a[18 + i] > best) { best = a[18 + i]; sig = i; }
  }
  const int pred_contact = sig;                         // np.argmax(self.next_contact_one_hot) in _get_reward
  if (!P.use_filter && (s.flags & FLAG_STAND)) sig = 3;  // training variant, stand clip
  bool lc = (sig & 1) != 0, rc = (sig & 2) != 0;
  if (P.use_filter && A.contact_prob != nullptr) {
    float p0 = __ldg(A.contact_prob + 2 * (size_t)e), p1 = __ldg(A.contact_prob + 2 * (size_t)e + 1);
    if (p0 == p0) {
      bool l0 = (s.flags & FLAG_L) != 0, r0 = (s.flags & FLAG_R) != 0;
      const double conf = 0.7;
      lc = (l0 && (double)p0 < (1.0 - conf)) ? false : ((!l0 && (double)p0 > conf) ? true : l0);
      rc = (r0 && (double)p1 < (1.0 - conf)) ? false : ((!r0 && (double)p1 > conf) ? true : r0);
    }
  }
  const bool last_lc = (s.flags & FLAG_L) != 0, last_rc = (s.flags & FLAG_R) != 0;
  s.flags = (s.flags & ~(FLAG_L | FLAG_R)) | (lc ? FLAG_L : 0) | (rc ? FLAG_R : 0);
  const TerrainDev* T = TERRAIN ? (P.use_terrain ? A.terrain : nullptr) : nullptr;

  // ---- feet (:320-371; training: ..._training.py:231-258)
  if (P.use_filter) {
    if (!(s.flags & FLAG_FILT)) {                       // initFilter (:1209-1234)
#pragma unroll
      for (int i = 0; i < 2; i++) {
        s.filt_yaw[i] = a[16 + i];
        V3<real> g = yawR(fx, fy, V3<real>{a[2 * i], a[2 * i + 1], real(0)}) + origin;
        s.ballx[2 * i] = g.x; s.ballx[2 * i + 1] = g.y;
        s.ballv[2 * i] = real(0); s.ballv[2 * i + 1] = real(0);
      }
      s.flags |= FLAG_FILT;
    }
    real speed = num<real>::sqrt(s.vlin.x * s.vlin.x + s.vlin.y * s.vlin.y);
    real K0, K1;
    sds_gain(speed, &K0, &K1);
#pragma unroll
    for (int i = 0; i < 2; i++) {
      bool in_contact = i == 0 ? lc : rc;
      if (!in_contact) {                                // filterFoot (:1238-1285)
        s.filt_yaw[i] = s.filt_yaw[i] * real(0.5) + a[16 + i] * (real(1.0) - real(0.5));
        V3<real> g = yawR(fx, fy, V3<real>{a[2 * i], a[2 * i + 1], real(0)}) + origin;
        sds_4steps(s.ballx[2 * i], s.ballv[2 * i], g.x, K0, K1);
        sds_4steps(s.ballx[2 * i + 1], s.ballv[2 * i + 1], g.y, K0, K1);
        s.foot[i] = {s.ballx[2 * i], s.ballx[2 * i + 1], real(0)};
        real yaw = s.filt_yaw[i];
        if (T != nullptr) {                             // terrain variant (..._terrain_window.py:340-342,401-403)
          s.foot[i].z = terrain_query<real>(*T, s.foot[i].x, s.foot[i].y, nullptr);
          if (i == 1) yaw = a[17];                      // Q9: the right swing foot uses the raw action yaw
        }
        real sa, ca;
        num<real>::sincos(yaw, &sa, &ca);
        s.yc[i] = fx * ca - fy * sa; s.ys[i] = fy * ca + fx * sa;     // ffSRB_ori @ Rz(filter yaw)
      } else {                                          // updateFilter (:1289-1299): rows 0 of both matrices
        real v1x = s.yc[i], v1y = -s.ys[i], v2x = fx, v2y = -fy;
        real n1 = num<real>::sqrt(v1x * v1x + v1y * v1y), n2 = num<real>::sqrt(v2x * v2x + v2y * v2y);
        v1x /= n1; v1y /= n1; v2x /= n2; v2y /= n2;
        s.filt_yaw[i] = num<real>::atan2(v1x * v2y - v1y * v2x, v1x * v2x + v1y * v2y);
        s.ballv[2 * i] = real(0); s.ballv[2 * i + 1] = real(0);
        if (T != nullptr && !(i == 0 ? last_lc : last_rc)) {
          // touchdown on a pyramid layer: snap the foot to the nearest edge of the mid-size between this layer and
          // the next geom (..._terrain_window.py:346-380, 407-442); the height sampled while swinging is kept (Q10)
          TerrainHit hit;
          terrain_query<real>(*T, s.foot[i].x, s.foot[i].y, &hit);
          int k = hit.geom - T->box_geom0;
          if (hit.geom >= 0 && k >= 0 && k + 1 < T->n_box) {
            int bi = -1;
            for (int b = 0; b < T->n_body; b++) if (k >= T->body[b].first && k < T->body[b].first + T->body[b].count) bi = b;
            if (bi >= 0 && T->body[bi].snap) {
              const double* g0 = T->box + (size_t)k * 5; const double* g1 = g0 + 5;
              real hx = ((real)g0[2] + (real)g1[2]) / real(2), hy = ((real)g0[3] + (real)g1[3]) / real(2);
              real x_min = (real)T->body[bi].cx - hx, x_max = (real)T->body[bi].cx + hx;
              real y_min = (real)T->body[bi].cy - hy, y_max = (real)T->body[bi].cy + hy;
              real px = s.foot[i].x, py = s.foot[i].y;
              real dx_min = num<real>::abs(px - x_min), dx_max = num<real>::abs(px - x_max);
              real dy_min = num<real>::abs(py - y_min), dy_max = num<real>::abs(py - y_max);
              real cx = px < x_min ? x_min : (px > x_max ? x_max : px), cy = py < y_min ? y_min : (py > y_max ? y_max : py);
              real md = dx_min; md = dx_max < md ? dx_max : md; md = dy_min < md ? dy_min : md; md = dy_max < md ? dy_max : md;
              if (md == dx_min) { px = x_min; py = cy; }
              else if (md == dx_max) { px = x_max; py = cy; }
              else if (md == dy_min) { px = cx; py = y_min; }
              else { px = cx; py = y_max; }
              s.foot[i].x = px; s.foot[i].y = py;
            }
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      bool in_contact = i == 0 ? lc : rc;
      if (!in_contact) {
        s.foot[i] = yawR(fx, fy, V3<real>{a[2 * i], a[2 * i + 1], real(0)}) + origin;
        real sa, ca;
        num<real>::sincos(a[16 + i], &sa, &ca);
        s.yc[i] = fx * ca - fy * sa; s.ys[i] = fy * ca + fx * sa;
      }
    }
  }
  const int contact_mask = (lc ? 1 : 0) | (rc ? 2 : 0);

  // ---- 4 physics sub-steps (:382-450)
  real Mdiag[6] = {P.mass + P.armature, P.mass + P.armature, P.mass + P.armature,
                   P.inertia[0] + P.armature, P.inertia[1] + P.armature, P.inertia[2] + P.armature};
  real Minv[6];
#pragma unroll
  for (int i = 0; i < 6; i++) Minv[i] = P.inv_M[i];
  real nu[6];
#pragma unroll
  for (int i = 0; i < 6; i++) nu[i] = s.nu[i];
  bool warm = (s.flags & FLAG_NU) != 0;
  unsigned aset = 0; bool have_aset = false;              // active set of the previous sub-step's optimum (this lane's columns)
  double* dbg = (A.dbg != nullptr && active) ? A.dbg + (size_t)e * DBG_W : nullptr;
  long long tk0 = 0, tk_pre = 0;
  if (A.dbg != nullptr) { tk_pre = clock64(); tk0 = tk_pre; }   // phase clocks land in the debug record's pad slots

  int fc_mod = P.use_impulse ? s.fcount % P.impulse_interval : 0;
  for (int sub = 0; sub < P.nsub; sub++) {
    long long tk_sub = 0;
    if (A.dbg != nullptr) tk_sub = clock64();
    // mj_step1: kinematics + bias
    const M3<real> R = quat2mat_normalized(s.quat);
    s.qlag = s.quat;
    V3<real> Iw = {P.inertia[0] * s.vang.x, P.inertia[1] * s.vang.y, P.inertia[2] * s.vang.z};
    V3<real> bw = cross(s.vang, Iw);
    real bias[6] = {real(0), real(0), P.mass * P.grav, bw.x, bw.y, bw.z};
    // update_impulse (:191-206)
    real xf[6] = {0, 0, 0, 0, 0, 0};     // xfrc_applied: world force, world torque -> J^T: torque onto the body axes (mj_xfrcAccumulate)
    if (P.use_impulse) {
      const bool trig = fc_mod == (P.impulse_trigger_at_end ? P.impulse_interval - 1 : 0);   // frame_counter % interval, kept incrementally
      fc_mod = fc_mod + 1 == P.impulse_interval ? 0 : fc_mod + 1;
      if (trig) s.impulse = P.impulse_duration;
      if (s.impulse > 0) {
        s.impulse -= 1;
        xf[0] = P.impulse_force[0]; xf[1] = P.impulse_force[1]; xf[2] = P.impulse_force[2];
        const V3<real> tq = mulT(R, V3<real>{P.impulse_force[3], P.impulse_force[4], P.impulse_force[5]});
        xf[3] = tq.x; xf[4] = tq.y; xf[5] = tq.z;
      }
      s.fcount += 1;
    }
    real qfrc[6] = {0, 0, 0, 0, 0, 0};
    if (__any_sync(0xffffffffu, contact_mask != 0)) {
      // desired acceleration (viewer.py:534-538)
      real pfx, pfy;
      yaw_frame(R, &pfx, &pfy);
      V3<real> v_tw, w_tw;
      log_se3(matmulTN(R, R_hat), mulT(R, p_hat - s.pos), &v_tw, &w_tw);
      V3<real> vw = mul(R, v_tw);
      real y[6];
      y[0] = P.kp * vw.x + P.kd * (tv_lin.x - s.vlin.x) + Minv[0] * bias[0];
      y[1] = P.kp * vw.y + P.kd * (tv_lin.y - s.vlin.y) + Minv[1] * bias[1];
      y[2] = P.kp * vw.z + P.kd * (tv_lin.z - s.vlin.z) + Minv[2] * bias[2];
      y[3] = P.kp * w_tw.x + P.kd * (tv_ang.x - s.vang.x) + Minv[3] * bias[3];
      y[4] = P.kp * w_tw.y + P.kd * (tv_ang.y - s.vang.y) + Minv[4] * bias[4];
      y[5] = P.kp * w_tw.z + P.kd * (tv_ang.z - s.vang.z) + Minv[5] * bias[5];
      QP<real, G> qp;
      qp.build(lane, R, s.pos, s.foot, s.yc, s.ys, contact_mask, pfx, pfy, P.mu_b, Minv);
      if (!warm) {
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = -y[i];
      }
      real t[QP<real, G>::NC];
      int iters = qp.solve(lane, y, P.inv_w_lambda, nu, t, hsm, (P.aset_warm && have_aset) ? &aset : nullptr);
      warm = true;
      aset = 0;
#pragma unroll
      for (int k = 0; k < QP<real, G>::NC; k++) aset |= (t[k] < 0 ? 1u : 0u) << k;
      have_aset = true;
      if (contact_mask != 0) {
        s.flags |= FLAG_NU;
#pragma unroll
        for (int i = 0; i < 6; i++) s.nu[i] = nu[i];
      }
      const real invw = P.inv_w_lambda;
      bool clamped = false;
      // Every friction-basis vector is (mu_b * unit horizontal, 1): a foot's force obeys |F_foot| <= sqrt(1 + mu_b^2) F_foot,z and
      // F_foot,z <= F_total,z = M (nu + y)_z (all multipliers >= 0), so the per-foot sums are only needed when that bound can
      // reach the threshold at all (never in a normal gait: 1e4 N).  Same decisions, three group reductions fewer per sub-step.
      const bool may_clamp = Mdiag[2] * (nu[2] + y[2]) * P.clamp_bound > P.force_thr;
      if (P.use_clamp && __any_sync(0xffffffffu, may_clamp && contact_mask != 0)) {
        // per-foot force sums and the 1e4 N clamp (:402-443, Q3)
        real FL[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < QP<real, G>::NC; k++) {
          int c = lane + k * G;
          bool left = ((c >> 2) & 1) == 0;
          real lam = (t[k] < 0 ? -t[k] : real(0)) * invw;
          real m = left ? lam : real(0);
          FL[0] += m * qp.a[k][0] * Mdiag[0]; FL[1] += m * qp.a[k][1] * Mdiag[1]; FL[2] += m * qp.a[k][2] * Mdiag[2];
        }
#pragma unroll
        for (int i = 0; i < 3; i++) FL[i] = gsum<G>(FL[i]);
        real FT[3] = {Mdiag[0] * (nu[0] + y[0]), Mdiag[1] * (nu[1] + y[1]), Mdiag[2] * (nu[2] + y[2])};
        real nL = num<real>::sqrt(FL[0] * FL[0] + FL[1] * FL[1] + FL[2] * FL[2]);
        real nR = num<real>::sqrt((FT[0] - FL[0]) * (FT[0] - FL[0]) + (FT[1] - FL[1]) * (FT[1] - FL[1]) + (FT[2] - FL[2]) * (FT[2] - FL[2]));
        clamped = (nL > P.force_thr) || (nR > P.force_thr);
        if (__any_sync(0xffffffffu, clamped)) {
          real sL = nL > P.force_thr ? P.force_thr / nL : real(1), sR = nR > P.force_thr ? P.force_thr / nR : real(1);
          real W[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int k = 0; k < QP<real, G>::NC; k++) {
            int c = lane + k * G;
            int site = ((c >> 2) & 1) * 5 + (c >> 3);
            real lam = (t[k] < 0 ? -t[k] : real(0)) * invw * (site > 5 ? sR : sL);
#pragma unroll
            for (int i = 0; i < 6; i++) W[i] += lam * qp.a[k][i] * Mdiag[i];
          }
#pragma unroll
          for (int i = 0; i < 6; i++) W[i] = gsum<G>(W[i]);
          if (clamped) {
#pragma unroll
            for (int i = 0; i < 6; i++) qfrc[i] = W[i];
          }
        }
      }
      if (!clamped && contact_mask != 0) {
        // sum_i J_i^T f_i = C lam = M (nu + y) at the optimum
#pragma unroll
        for (int i = 0; i < 6; i++) qfrc[i] = Mdiag[i] * (nu[i] + y[i]);
      }
      if (dbg != nullptr && contact_mask != 0) {
        double* d = dbg + sub * DBG_SUB;
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 6; i++) { d[i] = (double)(nu[i] + y[i] - Minv[i] * bias[i]); d[6 + i] = (double)nu[i]; d[12 + i] = (double)qfrc[i]; }
          d[18] = (double)iters;
          d[19] = (double)(clock64() - tk_sub);           // cycles: sub-step start .. QP solved
        }
#pragma unroll
        for (int k = 0; k < QP<real, G>::NC; k++) {
          int c = lane + k * G;
          if (c < 40) {
            int site = ((c >> 2) & 1) * 5 + (c >> 3);
            d[24 + site * 4 + (c & 3)] = (double)((t[k] < 0 ? -t[k] : real(0)) * invw);
          }
        }
      }
    }
    // mj_step2: acceleration + semi-implicit Euler with implicit joint damping
    real qv[6] = {s.vlin.x, s.vlin.y, s.vlin.z, s.vang.x, s.vang.y, s.vang.z};
#pragma unroll
    for (int i = 0; i < 6; i++) {
      real f = (-P.damping * qv[i]) - bias[i] + qfrc[i] + xf[i];
      qv[i] += P.h * (f * P.inv_Mh[i]);
    }
    s.vlin = {qv[0], qv[1], qv[2]};
    s.vang = {qv[3], qv[4], qv[5]};
    s.pos = {s.pos.x + P.h * s.vlin.x, s.pos.y + P.h * s.vlin.y, s.pos.z + P.h * s.vlin.z};
    s.quat = quat_integrate(s.quat, s.vang, P.h);
  }

  if (dbg != nullptr && lane == 0) { dbg[20] = (double)(clock64() - tk_pre); }    // cycles: 4 sub-steps
  // ---- history push (:453-474): entry frame+1 = (qpos[:3], stale xmat)
  if (active && lane == 0) {
    real* h = hist_env + (size_t)((s.frame + 1) & (HIST_SLOTS - 1)) * HIST_W;
    h[0] = s.pos.x; h[1] = s.pos.y; h[2] = s.pos.z; h[3] = s.qlag.w; h[4] = s.qlag.x; h[5] = s.qlag.y; h[6] = s.qlag.z;
  }

  // ---- obs / reward / done (:477-494; training variant takes the obs after the frame increment)
  float* obs = CIO ? sio + (lane32 / G) * OBS_W : A.obs + (size_t)e * OBS_W;
  const bool want_window = !(CIO && A.split_obs);        // split host path: the window travels by copy engine (srb_obs_window_kernel)
  if (!P.obs_after_inc) write_obs<real, G>(s, cd, obs, lane, active, T, want_window);
  float info[8];
  real r = real(0);
  if (T == nullptr) {
    r = compute_reward<real, G>(s, cd, hist_env, pred_contact, P, info, lane);
  } else {                                               // terrain env: _get_reward returns (0, {}) (:525-526)
#pragma unroll
    for (int i = 0; i < 8; i++) info[i] = 0.f;
  }
  bool done = compute_done(s, cd, P, T);
  s.frame += 1;
  if (P.obs_after_inc) write_obs<real, G>(s, cd, obs, lane, active, T, want_window);
  s.ep_ret += r;
  if (!CIO) {
    if (active && lane == 0) {
      A.rew[e] = (float)r;
      A.done[e] = done ? 1 : 0;
      if (A.rinfo != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; i++) A.rinfo[(size_t)e * RINFO_W + i] = info[i];
      }
    }
  } else {                                               // reward / done / reward_info of the warp's environments, coalesced
    __shared__ float srew[CIO ? EPW : 1];
    __shared__ __align__(16) float sinfo[CIO ? EPW * RINFO_W : 4];
    if (lane == 0) {
      srew[lane32 / G] = (float)r;
#pragma unroll
      for (int i = 0; i < 8; i++) sinfo[(lane32 / G) * RINFO_W + i] = info[i];
    }
    const unsigned dbits = __ballot_sync(0xffffffffu, done);
    __syncwarp();
    if (lane32 < nrows) {
      A.rew[env0 + lane32] = srew[lane32]; A.done[env0 + lane32] = (dbits >> (lane32 * G)) & 1u;
      if (A.done_dev != nullptr) A.done_dev[env0 + lane32] = (dbits >> (lane32 * G)) & 1u;
    }
    if (A.rinfo != nullptr)
      for (int i = lane32; i < nrows * RINFO_W; i += 32) A.rinfo[(size_t)env0 * RINFO_W + i] = sinfo[i];
  }
  // ---- fused auto-reset (what SB3's VecEnv does after a terminal step)
  const bool any_reset = A.auto_reset && __any_sync(0xffffffffu, done);
  long long tk_r0 = 0;
  if (A.dbg != nullptr) tk_r0 = clock64();                // reset-path clocks land in the pad slots of sub-step 1
  if (any_reset) {
    __syncwarp();                                        // obs row of this env was written by sibling lanes
    if (done && active && A.term_obs != nullptr) {
      float* to = A.term_obs + (size_t)e * OBS_W;
      for (int i = lane; i < OBS_W; i += G) to[i] = obs[i];
    }
    __syncwarp();                                        // ... and is overwritten by the reset observation below
    if (done) {
      if (active && lane == 0 && A.ep_stats != nullptr) {
        atomicAdd(A.ep_stats + 0, (double)s.ep_ret);
        atomicAdd(A.ep_stats + 1, (double)s.frame);
        atomicAdd(A.ep_stats + 2, 1.0);
      }
      double plan[9];
      if (A.reset_plan != nullptr) {
#pragma unroll
        for (int i = 0; i < 9; i++) plan[i] = A.reset_plan[(size_t)e * 9 + i];
      } else {
        draw_reset_plan(A.seed, e, s.episode, A.clips, A.nclips, plan, T != nullptr);
      }
      if (dbg != nullptr && lane == 0) dbg[DBG_SUB + 19] = (double)(clock64() - tk_r0);
      EnvState<real> ns = s;
      real h0[HIST_W];
      apply_reset(ns, A.clips, A.nclips, plan, h0, T);
      s = ns;
      if (dbg != nullptr && lane == 0) dbg[DBG_SUB + 20] = (double)(clock64() - tk_r0);
      if (active && lane == 0) {
#pragma unroll
        for (int i = 0; i < HIST_W; i++) hist_env[i] = h0[i];
      }
      const ClipDesc<real> ncd = A.clips[s.clip];
      write_obs<real, G>(s, ncd, obs, lane, active, T, want_window);
      if (dbg != nullptr && lane == 0) dbg[DBG_SUB + 21] = (double)(clock64() - tk_r0);
    }
  }
  if (active && lane == 0) store_state(s, A.st, A.ist, A.Npad, e);
  if (dbg != nullptr && lane == 0) dbg[DBG_SUB + 22] = (double)(clock64() - tk_r0);
  if (CIO) {                                             // the warp's observation rows: one contiguous block
    __syncwarp();
    float* dst = A.obs + (size_t)env0 * OBS_W;
    if (A.split_obs) {                                    // self part of every row only (25 floats = 100 bytes per row)
      for (int i = lane32; i < nrows * FEAT_W; i += 32) { const int r = i / FEAT_W, c = i - r * FEAT_W; dst[r * OBS_W + c] = sio[r * OBS_W + c]; }
    } else if ((EPW % 2 == 0) && nrows == EPW) {
      const float4* s4 = reinterpret_cast<const float4*>(sio);
      float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 2
      for (int i = lane32; i < EPW * OBS_W / 4; i += 32) d4[i] = s4[i];
    } else {
      for (int i = lane32; i < nrows * OBS_W; i += 32) dst[i] = sio[i];
    }
  }
  if (dbg != nullptr && lane == 0) { dbg[21] = (double)(clock64() - tk0); dbg[22] = (double)(tk_pre - tk_entry); { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); dbg[23] = (double)gt; dbg[64 + 23] = (double)gt_entry; } }
}

// ------------------------------------------------------------------------------------------------
// Look-ahead window of the observation rows (get_obs_window :1096-1161: 5 precomputed 25-float records of the clip table), gathered
// on its own: it depends on (clip, frame index) only, i.e. it is known BEFORE the step runs.  The host-buffer path (srb_step_host,
// mode 2) launches this first into the device copy of the observation array and lets a copy engine move the whole array to the host
// as ONE contiguous transfer while the step kernel computes; a second launch after that copy patches what the step produced:
// `self_src` != null (patch launch): the 25 self features of every row are copied from there (device rows) to dst, and the window
// is rewritten only for environments flagged in `only_done` (auto-reset in the step just finished: their window changed; null =
// no environment, i.e. auto-reset is off).
// `patch_only` (mode 3, where the step kernel itself stores the self features to the host rows and only the window columns travel by
// copy engine): windows of the `only_done` environments, no self features.
// `after_inc`: the training variant takes the observation after current_frame += 1.  One 160-thread block per environment.
template <typename real>
__global__ void srb_obs_window_kernel(int N, int Npad, const int* __restrict__ ist, const ClipDesc<real>* __restrict__ clips, int after_inc,
                                      const uint8_t* __restrict__ only_done, const float* __restrict__ self_src, float* __restrict__ dst, int patch_only) {
  const int e = blockIdx.x;
  if (e >= N) return;
  const int i = threadIdx.x;
  if (self_src != nullptr && i >= 5 * FEAT_W && i < 6 * FEAT_W) {
    const int c = i - 5 * FEAT_W;
    dst[(size_t)e * OBS_W + c] = self_src[(size_t)e * OBS_W + c];
  }
  if ((self_src != nullptr || patch_only) && (only_done == nullptr || !only_done[e])) return;      // patch launch: windows of auto-reset environments only
  if (i >= 5 * FEAT_W) return;
  const int frame = ist[(size_t)I_FRAME * Npad + e], rif = ist[(size_t)I_RIF * Npad + e], clip = ist[(size_t)I_CLIP * Npad + e];
  const ClipDesc<real> cd = clips[clip];
  const int f0 = frame + rif + 1 + (after_inc ? 1 : 0);
  const int wi = i / FEAT_W, c = i - wi * FEAT_W;
  dst[(size_t)e * OBS_W + FEAT_W + i] = __ldg(cd.feat + (size_t)min(max(f0 + wi, 0), cd.nframes - 1) * FEAT_W + c);   // per-frame clamp, as write_obs
}

// ------------------------------------------------------------------------------------------------
// explicit / random reset kernel: one thread per environment
template <typename real> __global__ void srb_reset_kernel(const ResetArgs<real> A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.count) return;
  int e = A.idx ? A.idx[i] : i;
  if (e < 0 || e >= A.N) return;
  EnvState<real> s;
  load_state(s, A.st, A.ist, A.Npad, e);
  double plan[9];
  if (A.plan != nullptr) {
    for (int k = 0; k < 9; k++) plan[k] = A.plan[(size_t)i * 9 + k];
  } else {
    draw_reset_plan(A.seed, e, s.episode, A.clips, A.nclips, plan, A.p.use_terrain != 0);
  }
  real* hist_env = A.hist + (size_t)e * HIST_SLOTS * HIST_W;
  const TerrainDev* T = A.p.use_terrain ? A.terrain : nullptr;
  apply_reset(s, A.clips, A.nclips, plan, hist_env, T);
  store_state(s, A.st, A.ist, A.Npad, e);
  if (A.obs != nullptr) {
    const ClipDesc<real> cd = A.clips[s.clip];
    write_obs<real, 1>(s, cd, A.obs + (size_t)e * OBS_W, 0, true, T);
  }
}


// ------------------------------------------------------------------------------------------------
// Reference-tracking controller used to drive benchmarks and roll-out tests on the device: the action
// that asks the QP to reach the next reference frame (feet, yaw and contact from the clip, pose and
// velocity targets = next reference frame expressed in the current body frame) plus N(0, sigma^2) noise.
// Not part of the reference; it plays the role of the policy (walk_SRBTrack2025.py:424 `ppo predict`).
template <typename real>
__global__ void srb_tracking_action_kernel(int N, int Npad, const real* __restrict__ st, const int* __restrict__ ist,
                                           const ClipDesc<real>* __restrict__ clips, float sigma, unsigned long long seed,
                                           unsigned int step, float* __restrict__ act, const TerrainDev* __restrict__ T) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  EnvState<real> s;
  load_state(s, st, ist, Npad, e);
  const ClipDesc<real> cd = clips[s.clip];
  int f = min(s.frame + s.rif + 1, cd.nframes - 1);
  const real* fr = cd.frames + (size_t)f * CLIP_W;
  M3<real> R = quat2mat_normalized(s.qlag);
  real fx, fy;
  yaw_frame(R, &fx, &fy);
  real a[ACT_W];
#pragma unroll
  for (int i = 0; i < ACT_W; i++) a[i] = real(0);
  V3<real> origin = {s.pos.x, s.pos.y, real(0)};
  V3<real> rpos = {fr[0], fr[1], fr[2]};
  real px = real(0), py = real(0);
  if (T != nullptr) {                               // terrain variant: shifted + lifted reference
    px = s.planar[0]; py = s.planar[1];
    rpos = lifted_frame<real>(fr, s.planar, *T).pos;
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    V3<real> l = yawT(fx, fy, V3<real>{fr[18 + 3 * i] + px, fr[19 + 3 * i] + py, real(0)} - origin);
    a[2 * i] = l.x; a[2 * i + 1] = l.y;
    real yc = fr[24 + 2 * i], ys = fr[25 + 2 * i];
    a[16 + i] = num<real>::atan2(-fy * yc + fx * ys, fx * yc + fy * ys);
  }
  M3<real> Rr;
#pragma unroll
  for (int i = 0; i < 9; i++) Rr.m[i] = fr[3 + i];
  V3<real> vl = mulT(R, V3<real>{fr[12], fr[13], fr[14]});
  a[4] = vl.x; a[5] = vl.y; a[6] = vl.z; a[7] = fr[15]; a[8] = fr[16]; a[9] = fr[17];
  V3<real> v_tw, w_tw;
  V3<real> dp = mulT(R, rpos - s.pos);
  log_se3(matmulTN(R, Rr), dp, &v_tw, &w_tw);
  a[10] = dp.x; a[11] = dp.y; a[12] = dp.z; a[13] = w_tw.x; a[14] = w_tw.y; a[15] = w_tw.z;
  int c = (fr[20] == real(0) ? 1 : 0) + (fr[23] == real(0) ? 2 : 0);
  a[18 + c] = real(1);
  float* o = act + (size_t)e * ACT_W;
  for (int i = 0; i < ACT_W; i += 2) {
    uint32_t cnt[4] = {(uint32_t)e, step, (uint32_t)i, 0xAC7u};
    philox4x32(cnt, (uint32_t)seed, (uint32_t)(seed >> 32));
    float u1 = 1.0f - (float)(cnt[0] >> 8) * (1.0f / 16777216.0f), u2 = (float)(cnt[1] >> 8) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.2831853f * u2, &sn, &cs);
    o[i] = (float)a[i] + sigma * r * cs;
    o[i + 1] = (float)a[i + 1] + sigma * r * sn;
  }
}

// ------------------------------------------------------------------------------------------------
// batched terrain_height_ray: xy[n][2] -> height[n], hit[n][4] = (geom id, hfield cell c, cell r, triangle)
static __global__ void srb_terrain_height_kernel(const TerrainDev* __restrict__ T, int n, const double* __restrict__ xy, double* __restrict__ height,
                                          int* __restrict__ hit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  TerrainHit h;
  height[i] = terrain_query<double>(*T, xy[2 * i], xy[2 * i + 1], &h);
  if (hit) { hit[4 * i] = h.geom; hit[4 * i + 1] = h.cell_c; hit[4 * i + 2] = h.cell_r; hit[4 * i + 3] = h.tri; }
}

// ------------------------------------------------------------------------------------------------
// look-ahead feature precompute: 25 floats per reference frame (get_obs_window :1096-1161), fp64.
static __global__ void srb_clip_features_kernel(const double* __restrict__ frames, int nframes, float* __restrict__ feat) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  const double* fr = frames + (size_t)f * CLIP_W;
  M3<double> R;
  for (int i = 0; i < 9; i++) R.m[i] = fr[3 + i];
  V3<double> pos = {fr[0], fr[1], fr[2]};
  double fx, fy;
  yaw_frame(R, &fx, &fy);
  float* o = feat + (size_t)f * FEAT_W;
  o[0] = (float)pos.z;
  o[1] = (float)(fx * R.m[0] + fy * R.m[3]); o[2] = (float)(-fy * R.m[0] + fx * R.m[3]); o[3] = (float)R.m[6];
  o[4] = (float)(fx * R.m[1] + fy * R.m[4]); o[5] = (float)(-fy * R.m[1] + fx * R.m[4]); o[6] = (float)R.m[7];
  V3<double> vl = yawT(fx, fy, V3<double>{fr[12], fr[13], fr[14]});
  V3<double> va = yawT(fx, fy, mul(R, V3<double>{fr[15], fr[16], fr[17]}));
  o[7] = (float)vl.x; o[8] = (float)vl.y; o[9] = (float)vl.z;
  o[10] = (float)va.x; o[11] = (float)va.y; o[12] = (float)va.z;
  int c = 0;
  for (int i = 0; i < 2; i++) {
    V3<double> fp = {fr[18 + 3 * i], fr[19 + 3 * i], 0.0};
    if (fr[20 + 3 * i] == 0.0) c |= (1 << i);
    V3<double> l = yawT(fx, fy, fp - pos);
    double yc = fr[24 + 2 * i], ys = fr[25 + 2 * i];
    double m00 = fx * yc + fy * ys, m10 = -fy * yc + fx * ys;
    o[13 + 4 * i] = (float)l.x; o[14 + 4 * i] = (float)l.y; o[15 + 4 * i] = (float)l.z; o[16 + 4 * i] = (float)atan2(m10, m00);
  }
  o[21] = c == 0 ? 1.f : 0.f; o[22] = c == 1 ? 1.f : 0.f; o[23] = c == 2 ? 1.f : 0.f; o[24] = c == 3 ? 1.f : 0.f;
}

}  // namespace srb
