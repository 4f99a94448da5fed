// removeFootSliding_online / removeFootSliding_online_stair for batches of planning windows.
//
// Reference: gym_trackSRB/taesooLib/DeltaExtractor.py:650-893 (call sites walk_SRBTrack2025.py:631,
// walk_SRBTrack2025_terrain_singlethreaded.py:624); QuadraticFunctionHardCon: Resource/scripts/module.lua:2108-2207;
// math.projectTrajectoryToItsHull: gym_cdm2/ConvexHull2D.lua:362-455 over GrahamScan (PhysicsLib/convexhull/graham.cpp:81-209).
// fc.getContactPositions_v2 / _stair are not in the reference tree: restated as the stand-ins W5 / W6 of oracle/foot_sliding.py.
//
// Every hard constraint of these programmes pins ONE variable (x_f = value), so the KKT system of QuadraticFunctionHardCon reduces
// to a symmetric positive definite system in the free frames: A_ff x_f = b_f - A_fp x_p with A = sum_t w_t c_t c_t^T (velocity terms:
// tridiagonal, acceleration terms: pentadiagonal, the mean term: rank one), b = -sum_t w_t k_t c_t.  At most 7 free frames
// (nvar <= 8, frame 0 is always pinned): Gaussian elimination in registers / local memory, one thread per (window, marker), three
// coordinates each.  The stair variant's hull pass couples toe and heel of a foot and runs as a second small kernel.
#pragma once
#include "srb_kernels.cuh"

namespace footslide {

enum { NV = 8, MAXI = 4 };                                    // frames per window, contact intervals per foot (runs of a <= 8 bit mask)

struct Args {
  int n, nvar, stair, ground_mode;                            // ground_mode: 0 none, 1 constant height, 2 SRB playground terrain (Z up)
  double foot_len, ground_y;
  const double* con_all;                                      // [n][nvar][4]  toe L, heel L, toe R, heel R (> 0.5 = contact)
  const double* initial_feet;                                 // [n][12]
  const double* marker_traj;                                  // [n][nvar][12] smooth trajectories
  const double* marker;                                       // [n][nvar][12] jerky predictions
  const double* desired_vel;                                  // [n][nvar][12] or null (forward differences * 30 of marker_traj)
  const double* foot_center;                                  // [n][2][nvar][3] (stair) or null
  const srb::TerrainDev* terrain;                             // ground_mode 2
  double* out;                                                // [n][nvar][12]
};

// projectToGround: height under a Y-up point.  ZtoY maps (x, y, z)_zup -> (y, z, x)_yup, so the Z-up ground coordinates are (p.z, p.x).
__device__ __forceinline__ double ground_height(const Args& A, double px, double pz) {
  if (A.ground_mode == 2) return srb::terrain_query<double>(*A.terrain, pz, px, nullptr);
  return A.ground_y;
}

__device__ __forceinline__ void hnorm(double& x, double& y, double& z) {
  y = 0.0;
  const double n = sqrt(x * x + z * z);
  if (n > 0.0) { x /= n; z /= n; }
}

__global__ void footslide_kernel(const Args A) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= A.n * 4) return;
  const int w = tid >> 2, imarker = tid & 3, ifoot = imarker >> 1, i = imarker & 1;
  const int nvar = A.nvar;
  const double* con = A.con_all + (size_t)w * nvar * 4;
  const double* mt = A.marker_traj + (size_t)w * nvar * 12;
  const double* mk = A.marker + (size_t)w * nvar * 12;
  const double* init = A.initial_feet + (size_t)w * 12;
  const double fl = A.foot_len;
  unsigned toe_m = 0, heel_m = 0;
  for (int f = 0; f < nvar; f++) {
    if (con[f * 4 + ifoot * 2] > 0.5) toe_m |= 1u << f;
    if (con[f * 4 + ifoot * 2 + 1] > 0.5) heel_m |= 1u << f;
  }
  const unsigned any_m = toe_m | heel_m, mine_m = i == 0 ? toe_m : heel_m;
  // ---- contact intervals: position (W5 / W6) and direction of the foot on each
  int is[MAXI], ie[MAXI], ni = 0;
  double cp[MAXI][3], cd[MAXI][3];
  for (int f = 0; f < nvar;) {
    if (!((any_m >> f) & 1)) { f++; continue; }
    int e = f;
    while (e < nvar && ((any_m >> e) & 1)) e++;
    double ax = 0, ay = 0, az = 0, dx = 0, dz = 0;
    for (int k = f; k < e; k++) {
      double tx = mk[k * 12 + ifoot * 6], ty = mk[k * 12 + ifoot * 6 + 1], tz = mk[k * 12 + ifoot * 6 + 2];
      double hx = mk[k * 12 + ifoot * 6 + 3], hy = mk[k * 12 + ifoot * 6 + 4], hz = mk[k * 12 + ifoot * 6 + 5];
      dx += tx - hx; dz += tz - hz;
      const bool ct = (toe_m >> k) & 1, ch = (heel_m >> k) & 1;
      if (ct && !ch) { double ux = hx - tx, uy = 0, uz = hz - tz; hnorm(ux, uy, uz); hx = tx + ux * fl; hy = ty + uy * fl; hz = tz + uz * fl; }
      else if (ch && !ct) { double ux = tx - hx, uy = 0, uz = tz - hz; hnorm(ux, uy, uz); tx = hx + ux * fl; ty = hy + uy * fl; tz = hz + uz * fl; }
      if (A.stair) {
        const double* fc = A.foot_center + ((size_t)(w * 2 + ifoot) * nvar + k) * 3;
        ax += fc[0]; ay += fc[1]; az += fc[2];
      } else {
        ax += tx * 0.5 + hx * 0.5; ay += ty * 0.5 + hy * 0.5; az += tz * 0.5 + hz * 0.5;
      }
    }
    const double inv = 1.0 / (double)(e - f);
    if (A.stair) { cp[ni][0] = ax / (double)(e - f); cp[ni][1] = ay / (double)(e - f); cp[ni][2] = az / (double)(e - f); }   // numpy mean
    else { cp[ni][0] = ax * inv; cp[ni][1] = ay * inv; cp[ni][2] = az * inv; }
    double uy = 0;
    hnorm(dx, uy, dz);
    cd[ni][0] = dx; cd[ni][1] = 0.0; cd[ni][2] = dz;
    is[ni] = f; ie[ni] = e; ni++;
    f = e;
  }
  if (ni > 0) {                                                // the first contact is anchored at where the foot already is (:673-678)
    if (toe_m & 1u) { for (int d = 0; d < 3; d++) cp[0][d] = init[ifoot * 6 + d] - cd[0][d] * fl * 0.5; }
    else if (heel_m & 1u) { for (int d = 0; d < 3; d++) cp[0][d] = init[ifoot * 6 + 3 + d] + cd[0][d] * fl * 0.5; }
  }
  // ---- pinned frames of this marker
  unsigned pin_m = 1u;                                          // x_0 = initialFeet
  double pv[NV][3];
  for (int d = 0; d < 3; d++) pv[0][d] = init[imarker * 3 + d];
  const double sgn = i == 0 ? 0.5 : -0.5;
  for (int c = 0; c < ni; c++) {
    double mc[3], oc[3];
    for (int d = 0; d < 3; d++) { mc[d] = cp[c][d] + cd[c][d] * fl * sgn; oc[d] = cp[c][d] - cd[c][d] * fl * sgn; }
    if (A.ground_mode != 0 && !A.stair) { mc[1] = ground_height(A, mc[0], mc[2]); oc[1] = ground_height(A, oc[0], oc[2]); }
    for (int f = is[c] == 0 ? 1 : is[c]; f < ie[c]; f++) {
      pin_m |= 1u << f;
      if ((mine_m >> f) & 1) { for (int d = 0; d < 3; d++) pv[f][d] = mc[d]; }
      else {                                                   // only the other marker is down: rotate about it at distance footLen
        const int other = ifoot * 2 + (1 - i);
        double ux = mt[f * 12 + imarker * 3] - mt[f * 12 + other * 3], uy = mt[f * 12 + imarker * 3 + 1] - mt[f * 12 + other * 3 + 1],
               uz = mt[f * 12 + imarker * 3 + 2] - mt[f * 12 + other * 3 + 2];
        const double nn = sqrt(ux * ux + uy * uy + uz * uz);
        ux /= nn; uy /= nn; uz /= nn;
        pv[f][0] = oc[0] + fl * ux; pv[f][1] = oc[1] + fl * uy; pv[f][2] = oc[2] + fl * uz;
      }
    }
  }
  // ---- three quadratic programmes
  double* out = A.out + (size_t)w * nvar * 12;
  for (int dim = 0; dim < 3; dim++) {
    double M[NV][NV + 1];                                      // [A | b]
    for (int r = 0; r < NV; r++) for (int c = 0; c <= NV; c++) M[r][c] = 0.0;
    for (int v = 1; v < nvar; v++) {                           // (x_v - x_{v-1} - vel/30)^2
      double dv;
      if (A.desired_vel != nullptr) dv = A.desired_vel[((size_t)w * nvar + (v - 1)) * 12 + imarker * 3 + dim];
      else dv = (mt[v * 12 + imarker * 3 + dim] - mt[(v - 1) * 12 + imarker * 3 + dim]) * 30.0;
      const double k = -dv / 30.0;
      M[v][v] += 1.0; M[v - 1][v - 1] += 1.0; M[v][v - 1] -= 1.0; M[v - 1][v] -= 1.0;
      M[v][NV] -= k; M[v - 1][NV] += k;
    }
    for (int v = 1; v < nvar - 1; v++) {                       // 1.5 (-x_{v-1} + 2 x_v - x_{v+1})^2
      const double c3[3] = {-1.0, 2.0, -1.0};
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) M[v - 1 + a][v - 1 + b] += 1.5 * c3[a] * c3[b];
    }
    {                                                          // (c sum_{f>=2} x_f - c sum_{f>=2} marker_f)^2
      const double c = dim == 1 ? 1.0 : 0.1;
      double sm = 0.0;
      for (int f = 2; f < nvar; f++) sm += mk[f * 12 + imarker * 3 + dim];
      const double k = c * (-1.0 * sm);
      for (int a = 2; a < nvar; a++) { for (int b = 2; b < nvar; b++) M[a][b] += c * c; M[a][NV] -= k * c; }
    }
    // pinned frames -> right-hand side; their rows / columns become identity
    for (int p = 0; p < nvar; p++) {
      if (!((pin_m >> p) & 1)) continue;
      const double val = pv[p][dim];
      for (int r = 0; r < nvar; r++) if (!((pin_m >> r) & 1)) M[r][NV] -= M[r][p] * val;
    }
    for (int p = 0; p < nvar; p++) {
      if (!((pin_m >> p) & 1)) continue;
      for (int r = 0; r < nvar; r++) { M[r][p] = 0.0; M[p][r] = 0.0; }
      M[p][p] = 1.0; M[p][NV] = pv[p][dim];
    }
    // Gaussian elimination with partial pivoting (SPD on the free block, identity on the pinned one)
    for (int c = 0; c < nvar; c++) {
      int piv = c;
      double best = fabs(M[c][c]);
      for (int r = c + 1; r < nvar; r++) if (fabs(M[r][c]) > best) { best = fabs(M[r][c]); piv = r; }
      if (piv != c) for (int k = c; k <= NV; k++) { const double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
      const double inv = 1.0 / M[c][c];
      for (int r = c + 1; r < nvar; r++) {
        const double f = M[r][c] * inv;
        if (f != 0.0) { for (int k = c; k < nvar; k++) M[r][k] -= f * M[c][k]; M[r][NV] -= f * M[c][NV]; }
      }
    }
    double x[NV];
    for (int r = nvar - 1; r >= 0; r--) {
      double s = M[r][NV];
      for (int k = r + 1; k < nvar; k++) s -= M[r][k] * x[k];
      x[r] = s / M[r][r];
    }
    for (int f = 0; f < nvar; f++) out[f * 12 + imarker * 3 + dim] = x[f];
  }
  // ---- flat / terrain variant: nothing may end up below the ground (:752-758)
  if (A.ground_mode != 0 && !A.stair) {
    for (int f = 0; f < nvar; f++) {
      double* p = out + f * 12 + imarker * 3;
      const double g = ground_height(A, p[0], p[2]);
      if (p[1] < g) p[1] = g;
    }
  }
}

// math.projectTrajectoryToItsHull on y[0..n): upper side of the convex hull of the points (i, y_i), applied only when the hull has
// more than three edges.  GrahamScan semantics: points strictly above the chord first-last form the upper partition, the rest the
// lower one; each half hull drops collinear points.
__device__ inline void project_to_hull(double* y, int n) {
  auto dir = [](double x0, double y0, double x1, double y1, double x2, double y2) { return (x0 - x1) * (y2 - y1) - (x2 - x1) * (y0 - y1); };
  int up[NV], lo[NV], nu = 0, nl = 0;
  up[nu++] = 0; lo[nl++] = 0;
  for (int pass = 0; pass < 2; pass++) {
    int* h = pass == 0 ? lo : up;
    int& m = pass == 0 ? nl : nu;
    const double factor = pass == 0 ? 1.0 : -1.0;
    for (int k = 1; k < n; k++) {
      if (k < n - 1) {
        const bool above = dir(0.0, y[0], (double)(n - 1), y[n - 1], (double)k, y[k]) < 0;
        if (above != (pass == 1)) continue;
      }
      h[m++] = k;
      while (m >= 3 && factor * dir((double)h[m - 3], y[h[m - 3]], (double)h[m - 1], y[h[m - 1]], (double)h[m - 2], y[h[m - 2]]) <= 0) { h[m - 2] = h[m - 1]; m--; }
    }
  }
  if ((nl - 1) + (nu - 1) <= 3) return;
  double o[NV];
  for (int i = 0; i < n; i++) {
    double ymax = -10000.0;
    for (int pass = 0; pass < 2; pass++) {
      const int* h = pass == 0 ? lo : up;
      const int m = pass == 0 ? nl : nu;
      for (int k = 0; k + 1 < m; k++) {
        if (h[k] <= i && i <= h[k + 1]) {
          const double t = (double)(i - h[k]) / (double)(h[k + 1] - h[k]);
          const double v = y[h[k]] * (1.0 - t) + y[h[k + 1]] * t;
          ymax = v > ymax ? v : ymax;
        }
      }
    }
    o[i] = y[i] > ymax ? y[i] : ymax;
  }
  for (int i = 0; i < n; i++) y[i] = o[i];
}

// avoidHull x 3 (toe weights 0.5, 0.75, 0.25; DeltaExtractor.py:863-893): one thread per (window, foot)
__global__ void footslide_hull_kernel(const Args A) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= A.n * 2) return;
  const int w = tid >> 1, ifoot = tid & 1, nvar = A.nvar;
  double* out = A.out + (size_t)w * nvar * 12 + ifoot * 6;
  const double wts[3] = {0.5, 0.75, 0.25};
  for (int pass = 0; pass < 3; pass++) {
    const double tw = wts[pass];
    double ty[NV];
    for (int f = 0; f < nvar; f++) {
      const double* p = out + f * 12;
      ty[f] = ground_height(A, p[0] * tw + p[3] * (1.0 - tw), p[2] * tw + p[5] * (1.0 - tw));
    }
    project_to_hull(ty, nvar);
    for (int f = 0; f < nvar; f++) {
      double* p = out + f * 12;
      const double py = p[1] * tw + p[4] * (1.0 - tw);
      if (py < ty[f]) { const double d = ty[f] - py; p[1] += d; p[4] += d; }
    }
  }
}

}  // namespace footslide
