// C-ABI of the batched MMSTO solver (include/mmsto_b200.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/mmsto_b200.h"
#include "mmsto_kernels.cuh"
#include "footslide_kernels.cuh"

extern "C" const void* srb_terrain_device(void* env);   // srb_api.cu: the TerrainDev of an SRBTrack handle (terrain variant)

namespace {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, const char* a = "") { snprintf(g_err, sizeof(g_err), fmt, a); return code; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(-2, "CUDA error: %s", cudaGetErrorString(e_)); } while (0)
}  // namespace

struct mmsto_solver {
  int device = 0;
  mmsto::Skel sk;
  mmsto::Prm prm;
  std::map<std::string, double> named;
  std::vector<double> dof_w;
  mmsto::Skel* d_sk = nullptr;
  mmsto::Prm* d_prm = nullptr;
  double* ws = nullptr; double* mc = nullptr; int ws_warps = 0;
  int* d_queue = nullptr; int max_warps = 592;   // solve: persistent warps (4 per SM) pulling windows from a counter
  double* d_kinv = nullptr; int kinv_nvar = 0;    // inverse KKT matrix of removeCOMsliding_online for kinv_nvar frames
  double* d_zero = nullptr; size_t zero_n = 0;       // all-zero velocity constraints / contacts when the caller passes none
  bool dirty = true;
  long long launches = 0;
};

static void gauss_filter(int kernelsize, double* out) {           // Filter::GetGaussFilter (BaseLib/math/Filter.cpp:257-277)
  if (kernelsize % 2 == 0) kernelsize = kernelsize / 2 * 2 + 1;
  const int k = kernelsize / 2;
  const double gamma = 1.0 / kernelsize;
  double sum = 0;
  for (int i = -k; i <= k; ++i) { out[i + k] = exp(gamma * -1 * i * i); sum += out[i + k]; }
  for (int i = 0; i < kernelsize; ++i) out[i] /= sum;
}

static int sync_params(mmsto_solver* s) {
  if (!s->dirty) return 0;
  mmsto::Prm& P = s->prm;
  auto& m = s->named;
  P.w_ddx = m["weightDDX"]; P.w_dx = m["weightDX"]; P.w_ao = m["weightAngleOriginal"]; P.w_ee = m["weightEE"];
  P.w_eecon = m["weightEEcon"]; P.w_velhs = m["weightVelHalfSpace"]; P.w_mom = m["weightMomentum"]; P.w_com = m["weightCOM"];
  P.clamp = m["clampThr"]; P.eps = m["epsilon"]; P.max_iter = (int)m["max_iterations"]; P.iter_cap = (int)m["iteration_cap"];
  P.max_ls = 40; P.ftol = 1e-4; P.wolfe = 0.9; P.min_step = 1e-20; P.max_step = 1e20;          // lbfgs.cpp:114-119
  for (int j = 0; j < s->sk.ndof; ++j) {
    P.s_ddx[j] = sqrt(P.w_ddx * s->dof_w[j]); P.s_dx[j] = sqrt(P.w_dx * s->dof_w[j]); P.s_ao[j] = sqrt(P.w_ao * s->dof_w[j]);
  }
  CK(cudaMemcpy(s->d_prm, &P, sizeof(P), cudaMemcpyHostToDevice));
  s->dirty = false;
  return 0;
}

static int prepare(mmsto_solver* s, int n, bool solve) {
  CK(cudaSetDevice(s->device));
  if (int r = sync_params(s)) return r;
  if (!s->d_queue) {
    CK(cudaMalloc(&s->d_queue, sizeof(int)));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    s->max_warps = 4 * sms;
  }
  int warps = (n + mmsto::PPW - 1) / mmsto::PPW;
  if (solve && warps > s->max_warps) warps = s->max_warps;
  if (warps > s->ws_warps) {
    cudaFree(s->ws); cudaFree(s->mc);
    const size_t wb = (size_t)warps * mmsto::NV * mmsto::MAXD * 32 * sizeof(double), mb = (size_t)warps * mmsto::MAXD * 6 * mmsto::PPW * sizeof(double);
    CK(cudaMalloc(&s->ws, wb)); CK(cudaMalloc(&s->mc, mb));
    CK(cudaMemset(s->ws, 0, wb)); CK(cudaMemset(s->mc, 0, mb));
    s->ws_warps = warps;
  }
  const size_t zn = (size_t)n * mmsto::MAXHS * mmsto::HS_W;
  if (zn > s->zero_n) {
    cudaFree(s->d_zero);
    CK(cudaMalloc(&s->d_zero, zn * sizeof(double)));
    std::vector<double> z(zn, 0.0);
    for (size_t i = 0; i < (size_t)n * mmsto::MAXHS; ++i) z[i * mmsto::HS_W] = -1.0;      // frame -1: no constraint
    CK(cudaMemcpy(s->d_zero, z.data(), zn * sizeof(double), cudaMemcpyHostToDevice));
    s->zero_n = zn;
  }
  return 0;
}

template <bool EVAL_ONLY> static int launch(mmsto_solver* s, mmsto::Args& A, void* stream) {
  int warps = (A.n + mmsto::PPW - 1) / mmsto::PPW;
  if (!EVAL_ONLY) {
    if (warps > s->max_warps) warps = s->max_warps;
    CK(cudaMemsetAsync(s->d_queue, 0, sizeof(int), (cudaStream_t)stream));
    A.queue = s->d_queue;
  }
  const size_t smem = (((size_t)s->sk.nb * mmsto::SM_W + mmsto::STK_LEVELS * mmsto::STK_W) * 32 + mmsto::PPW * mmsto::GM_W) * sizeof(double);
  auto kern = mmsto::mmsto_kernel<EVAL_ONLY>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<warps, 32, smem, (cudaStream_t)stream>>>(A);
  CK(cudaGetLastError());
  ++s->launches;
  return 0;
}

extern "C" {

const char* mmsto_last_error(void) { return g_err; }

int mmsto_create(int32_t device, int32_t nb, const int32_t* parent, const double* offset, const int32_t* nchan, const int32_t* axes,
                 const double* mass, const double* local_com, const double* inertia, int32_t nf, const int32_t* marker_bone,
                 const double* marker_lpos, const double* marker_weight, mmsto_solver** out) {
  if (!out) return fail(-1, "mmsto_create: out is null");
  if (nb < 1 || nb > mmsto::MAXB) return fail(-1, "mmsto_create: 1 <= nbones <= 24 required");
  if (nf < 3 || nf > mmsto::NFMAX) return fail(-1, "mmsto_create: 3 <= nf <= 8 required");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(-3, "mmsto_create: no CUDA device (there is no CPU fallback)");
  CK(cudaSetDevice(device));
  auto* s = new mmsto_solver();
  s->device = device;
  mmsto::Skel& sk = s->sk;
  memset(&sk, 0, sizeof(sk));
  sk.nb = nb;
  int d = 3;                                       // dofs 0..2: root slides X Y Z
  for (int b = 0; b < nb; ++b) {
    sk.parent[b] = parent[b];
    if (b == 0) {
      if (parent[0] >= 0) { delete s; return fail(-1, "mmsto_create: bone 0 must be the root"); }
      sk.nchan[0] = 3; sk.axis[0][0] = 1; sk.axis[0][1] = 2; sk.axis[0][2] = 0;              // YZX (VRMLloader.cpp:16)
    } else {
      if (parent[b] < 0 || parent[b] >= b) { delete s; return fail(-1, "mmsto_create: bones must be stored parents first"); }
      sk.nchan[b] = nchan[b];
      for (int k = 0; k < 3; ++k) sk.axis[b][k] = k < nchan[b] ? axes[b * 3 + k] : 0;
    }
    sk.dof0[b] = d; d += sk.nchan[b];
    sk.depth[b] = b == 0 ? 0 : sk.depth[parent[b]] + 1;
    if (sk.depth[b] >= mmsto::STK_LEVELS) { delete s; return fail(-1, "mmsto_create: tree depth <= 8 bones required"); }
    for (int k = 0; k < 3; ++k) { sk.off[b][k] = offset[b * 3 + k]; sk.com[b][k] = local_com[b * 3 + k]; }
    sk.mass[b] = mass[b]; sk.total_mass += mass[b];
    const double* I = inertia + b * 9;
    sk.inertia[b][0] = I[0]; sk.inertia[b][1] = I[4]; sk.inertia[b][2] = I[8]; sk.inertia[b][3] = I[1]; sk.inertia[b][4] = I[2]; sk.inertia[b][5] = I[5];
  }
  sk.ndof = d;
  if (d > mmsto::MAXD) { delete s; return fail(-1, "mmsto_create: ndof <= 64 required"); }
  for (int b = 0; b < nb; ++b) {                   // depth-first storage: a subtree is a contiguous range of bones
    int e = b + 1;
    while (e < nb) { int p = parent[e]; while (p > b) p = parent[p]; if (p != b) break; ++e; }
    sk.sub_end[b] = e;
  }
  mmsto::Prm& P = s->prm;
  memset(&P, 0, sizeof(P));
  P.nf = nf;
  for (int m = 0; m < mmsto::NMARK; ++m) {
    if (marker_bone[m] < 0 || marker_bone[m] >= nb) { delete s; return fail(-1, "mmsto_create: marker bone out of range"); }
    P.mbone[m] = marker_bone[m]; P.mweight[m] = marker_weight[m];
    for (int k = 0; k < 3; ++k) P.mlpos[m][k] = marker_lpos[m * 3 + k];
  }
  gauss_filter(nf, P.filt);
  s->dof_w.assign(d, 1.0);
  // walk_SRBTrack2025.py:529-537 (the values the online loop sets every frame); clampThr shortTermOptimizer.lua:118
  s->named = {{"weightDDX", 10}, {"weightDX", 100}, {"weightAngleOriginal", 100000}, {"weightEE", 100000}, {"weightEEcon", 0},
              {"weightVelHalfSpace", 100000}, {"weightMomentum", 1000}, {"weightCOM", 100000}, {"clampThr", 25}, {"epsilon", 1e-9},
              {"max_iterations", 0}, {"iteration_cap", 2000}};
  CK(cudaMalloc(&s->d_sk, sizeof(mmsto::Skel)));
  CK(cudaMalloc(&s->d_prm, sizeof(mmsto::Prm)));
  CK(cudaMemcpy(s->d_sk, &sk, sizeof(sk), cudaMemcpyHostToDevice));
  *out = s;
  return 0;
}

int mmsto_destroy(mmsto_solver* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaFree(s->d_sk); cudaFree(s->d_prm); cudaFree(s->ws); cudaFree(s->mc); cudaFree(s->d_zero); cudaFree(s->d_kinv); cudaFree(s->d_queue);
  delete s;
  return 0;
}

int mmsto_ndof(mmsto_solver* s) { return s ? s->sk.ndof : fail(-1, "mmsto_ndof: null solver"); }

int mmsto_set_param(mmsto_solver* s, const char* name, double value) {
  if (!s || !name) return fail(-1, "mmsto_set_param: null argument");
  auto it = s->named.find(name);
  if (it == s->named.end()) return fail(-1, "mmsto_set_param: unknown parameter '%s'", name);
  it->second = value; s->dirty = true;
  return 0;
}
int mmsto_get_param(mmsto_solver* s, const char* name, double* value) {
  if (!s || !name || !value) return fail(-1, "mmsto_get_param: null argument");
  auto it = s->named.find(name);
  if (it == s->named.end()) return fail(-1, "mmsto_get_param: unknown parameter '%s'", name);
  *value = it->second;
  return 0;
}
int mmsto_set_dof_weights(mmsto_solver* s, const double* w) {
  if (!s || !w) return fail(-1, "mmsto_set_dof_weights: null argument");
  s->dof_w.assign(w, w + s->sk.ndof); s->dirty = true;
  return 0;
}

int mmsto_solve(mmsto_solver* s, int32_t n, const double* x_original, const double* dx, const double* ddx, const double* euler_mot,
                const double* gpos_target, const double* com, const double* momentum, const int32_t* con, const double* velcon,
                double* x_out, double* info, void* stream) {
  if (!s) return fail(-1, "mmsto_solve: null solver");
  if (n <= 0) return 0;
  if (!x_original || !dx || !ddx || !euler_mot || !gpos_target || !com || !momentum || !x_out || !info) return fail(-1, "mmsto_solve: null array");
  if (int r = prepare(s, n, true)) return r;
  mmsto::Args A{};
  A.n = n; A.x_original = x_original; A.dx = dx; A.ddx = ddx; A.euler_mot = euler_mot; A.gpos_target = gpos_target; A.com = com;
  A.momentum = momentum; A.con = con; A.velcon = velcon ? velcon : s->d_zero; A.x_out = x_out; A.info = info;
  A.ws = s->ws; A.mc = s->mc; A.sk = s->d_sk; A.prm = s->d_prm;
  return launch<false>(s, A, stream);
}

int mmsto_evaluate(mmsto_solver* s, int32_t n, const double* x, const double* x_original, const double* dx, const double* ddx,
                   const double* gpos_target, const double* com, const double* momentum, const int32_t* con, const double* velcon,
                   double* obj, double* grad, void* stream) {
  if (!s) return fail(-1, "mmsto_evaluate: null solver");
  if (n <= 0) return 0;
  if (!x || !x_original || !dx || !ddx || !gpos_target || !com || !momentum || !obj || !grad) return fail(-1, "mmsto_evaluate: null array");
  if (int r = prepare(s, n, false)) return r;
  mmsto::Args A{};
  A.n = n; A.x_in = x; A.x_original = x_original; A.dx = dx; A.ddx = ddx; A.euler_mot = x; A.gpos_target = gpos_target; A.com = com;
  A.momentum = momentum; A.con = con; A.velcon = velcon ? velcon : s->d_zero; A.obj = obj; A.grad = grad;
  A.ws = s->ws; A.mc = s->mc; A.sk = s->d_sk; A.prm = s->d_prm;
  return launch<true>(s, A, stream);
}

int mmsto_remove_com_sliding(mmsto_solver* s, int32_t n, int32_t nvar, double* poses, const double* com_srb, const double* desired_com,
                             int32_t fix_y_also, const double* ground_h, double* curr_com_out, double* opt_com_out, void* stream) {
  if (!s) return fail(-1, "mmsto_remove_com_sliding: null solver");
  if (n <= 0) return 0;
  if (nvar < 4 || nvar > mmsto::NFMAX) return fail(-1, "mmsto_remove_com_sliding: 4 <= nvar <= 8 required");
  if (!poses || !com_srb || !desired_com) return fail(-1, "mmsto_remove_com_sliding: null array");
  CK(cudaSetDevice(s->device));
  mmsto::ComSlideArgs A{};
  A.n = n; A.nvar = nvar; A.npose = s->sk.ndof + 1; A.fix_y = fix_y_also; A.com_only = 0; A.poses = poses; A.com_srb = com_srb;
  A.desired_com = desired_com; A.ground_h = ground_h; A.curr_out = curr_com_out; A.opt_out = opt_com_out; A.sk = s->d_sk;
  if (!s->d_kinv) CK(cudaMalloc(&s->d_kinv, 11 * 11 * sizeof(double)));
  if (s->kinv_nvar != nvar) {
    CK(cudaStreamSynchronize((cudaStream_t)stream));       // an earlier launch may still read the previous inverse
    mmsto::comslide_kkt_inverse_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(nvar, s->d_kinv);
    CK(cudaGetLastError());
    s->kinv_nvar = nvar; ++s->launches;
  }
  A.kinv = s->d_kinv;
  mmsto::comslide_kernel<<<(n + 15) / 16, 128, 0, (cudaStream_t)stream>>>(A);
  CK(cudaGetLastError());
  ++s->launches;
  return 0;
}

int mmsto_window_com(mmsto_solver* s, int32_t n, int32_t nvar, const double* poses, double* com_out, void* stream) {
  if (!s) return fail(-1, "mmsto_window_com: null solver");
  if (n <= 0) return 0;
  if (nvar < 1 || nvar > mmsto::NFMAX) return fail(-1, "mmsto_window_com: 1 <= nvar <= 8 required");
  if (!poses || !com_out) return fail(-1, "mmsto_window_com: null array");
  CK(cudaSetDevice(s->device));
  mmsto::ComSlideArgs A{};
  A.n = n; A.nvar = nvar; A.npose = s->sk.ndof + 1; A.com_only = 1; A.poses = const_cast<double*>(poses); A.curr_out = com_out; A.sk = s->d_sk;
  mmsto::comslide_kernel<<<(n + 15) / 16, 128, 0, (cudaStream_t)stream>>>(A);
  CK(cudaGetLastError());
  ++s->launches;
  return 0;
}

int mmsto_remove_foot_sliding(int32_t device, int32_t n, int32_t nvar, double foot_len, const double* con_all, const double* initial_feet,
                              const double* marker_traj, const double* marker, const double* desired_vel, const double* foot_center,
                              int32_t ground_mode, double ground_y, void* srb_terrain_env, double* out, void* stream) {
  if (n <= 0) return 0;
  if (nvar < 3 || nvar > footslide::NV) return fail(-1, "mmsto_remove_foot_sliding: 3 <= nvar <= 8 required");
  if (!con_all || !initial_feet || !marker_traj || !marker || !out) return fail(-1, "mmsto_remove_foot_sliding: null array");
  if (ground_mode < 0 || ground_mode > 2) return fail(-1, "mmsto_remove_foot_sliding: ground_mode must be 0 (none), 1 (constant) or 2 (SRB terrain)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(-3, "mmsto_remove_foot_sliding: no CUDA device (there is no CPU fallback)");
  CK(cudaSetDevice(device));
  footslide::Args A{};
  A.n = n; A.nvar = nvar; A.stair = foot_center != nullptr; A.ground_mode = ground_mode; A.foot_len = foot_len; A.ground_y = ground_y;
  A.con_all = con_all; A.initial_feet = initial_feet; A.marker_traj = marker_traj; A.marker = marker; A.desired_vel = desired_vel;
  A.foot_center = foot_center; A.out = out; A.terrain = nullptr;
  if (ground_mode == 2) {
    A.terrain = (const srb::TerrainDev*)(srb_terrain_env ? srb_terrain_device(srb_terrain_env) : nullptr);
    if (!A.terrain) return fail(-1, "mmsto_remove_foot_sliding: ground_mode 2 needs an SRBTrack handle with an uploaded terrain");
  }
  if (A.stair && ground_mode == 0) return fail(-1, "mmsto_remove_foot_sliding: the stair variant needs a ground (projectToGround is not optional there)");
  footslide::footslide_kernel<<<(n * 4 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(A);
  CK(cudaGetLastError());
  if (A.stair) {
    footslide::footslide_hull_kernel<<<(n * 2 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(A);
    CK(cudaGetLastError());
  }
  return 0;
}

int fullbody_extract(int32_t device, int32_t n, const double* extra, const double* row, double srb_mass, const double* srb_inertia9,
                     const double* srb_local_com, double* human_pose, double* dx_ddx, double* feetpos, double* com, double* momentum,
                     double* con_bin, double* con, void* stream) {
  if (n <= 0) return 0;
  if (!extra || !row || !srb_inertia9 || !srb_local_com || !human_pose || !dx_ddx || !feetpos || !com || !momentum || !con_bin || !con)
    return fail(-1, "fullbody_extract: null array");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(-3, "fullbody_extract: no CUDA device (there is no CPU fallback)");
  CK(cudaSetDevice(device));
  mmsto::FullbodyArgs A{};
  A.n = n; A.extra = extra; A.row = row; A.mass = srb_mass;
  const double* I = srb_inertia9;
  A.inertia[0] = I[0]; A.inertia[1] = I[4]; A.inertia[2] = I[8]; A.inertia[3] = I[1]; A.inertia[4] = I[2]; A.inertia[5] = I[5];
  for (int k = 0; k < 3; ++k) { A.lcom[k] = srb_local_com[k]; A.r[k] = srb_mass * srb_local_com[k]; }
  A.human = human_pose; A.dx_ddx = dx_ddx; A.feetpos = feetpos; A.com = com; A.momentum = momentum; A.con_bin = con_bin; A.con = con;
  // bulk copies need 16-byte aligned global addresses: every array base must be (cudaMalloc / torch allocations are 256-byte aligned)
  const void* ptrs[] = {extra, row, human_pose, dx_ddx, feetpos, com, momentum, con_bin, con};
  for (const void* q : ptrs) if (((uintptr_t)q & 15) != 0) return fail(-1, "fullbody_extract: arrays must be 16-byte aligned");
  CK(cudaFuncSetAttribute(mmsto::fullbody_extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(mmsto::FullbodyTile)));
  mmsto::fullbody_extract_kernel<<<(n + mmsto::FB_ITEMS - 1) / mmsto::FB_ITEMS, 128, sizeof(mmsto::FullbodyTile), (cudaStream_t)stream>>>(A);
  CK(cudaGetLastError());
  return 0;
}

long long mmsto_launch_count(mmsto_solver* s) { return s ? s->launches : 0; }

}  // extern "C"
