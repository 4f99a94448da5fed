// C-ABI of the batched AdaptiveSRB step (declared in include/cdmenv_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/cdmenv_b200.h"
#include "cdm_kernels.cuh"

using namespace cdm;

int srb_set_error(int code, const std::string& msg);       // defined in srb_api.cu
#define CK(call)                                                                                                   \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess) return srb_set_error(-2, std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)

struct cdm_env {
  cdm_config cfg;
  int N = 0, Npad = 0, G = 4;
  size_t rsz = 8;
  void* st = nullptr; int* ist = nullptr; void* hist = nullptr;
  void* tab = nullptr; int nframes = 0, nf_cycle = 0;
  double* ep_stats = nullptr;
  float* term_obs = nullptr; const double* reset_plan = nullptr; double* dbg = nullptr;
  std::map<std::string, double> hp;          // named parameters
  cdm_skeleton skel; bool have_skel = false;
  ResetConst<double> rc;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_state = nullptr;        // recorded after every state-mutating launch on a caller stream; cdm_step_host's stream waits on it
  float *h_act = nullptr, *h_obs = nullptr, *h_rew = nullptr; uint8_t* h_done = nullptr;
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr; uint8_t* d_done = nullptr;
  int64_t launches = 0;
  Ground ground;                             // T.h == nullptr: flat ground (walk2-v1); else the walkterrain-v1 height field
  int obs_w = OBS_W;
};
struct tl_terrain { int device; double* h = nullptr; TLTerrainDev T; };

static void default_params(std::map<std::string, double>& p) {
  // gym_cdm2/spec/walk2-v1.lua over defaultRLsetting.lua('walk4') in training mode (no GUI)
  p["k_p"] = 120.0; p["k_d"] = 35.0;                  // QPparam.k_p_ID / k_d_ID
  p["w_ddq"] = 10000.0; p["w_lam"] = 0.0001;          // ddqObjWeight / lambdaObjWeight
  p["acc_clamp"] = 400.0;                             // QPservo:_calcDesiredAcceleration smoothClamp
  p["max_leg_len"] = 1.10; p["max_foot_vel"] = 3.0; p["logQ"] = 7.0;
  // Q15: the shipped reference build (NoLapack => EXCLUDE_CLAPACK, src/CMakeLists.txt:41) makes Physics.LQR return the gain
  // cached in work/_LQR_CACHE_2.DAT for every 2x2 system (PhysicsLib/SDRE.cpp:136-154); the recorded reference roll-outs
  // reproduce only with it.  use_lqr_cache = 0 switches to the real ARE solution for (M=30, b=1e-3, k=1, Q=10^logQ).
  p["use_lqr_cache"] = 1.0; p["filter_K0"] = 9999.00005; p["filter_K1"] = 1095.39024412;
  p["unlimited_leglen"] = 1.0; p["contact_thr"] = 0.0; p["touch_down_height"] = 0.0;   // training mode (defaultRLsetting.lua:183-187); contactThr 0 = motionInfos.walk4.contactThr (defaultRLsetting.lua:82)
  p["stride_thr"] = 0.45; p["max_contact_len"] = 0.9; p["min_swing_phase"] = 0.25; p["max_contact_y"] = 1.1;
  p["healthy_min"] = 0.8; p["healthy_max"] = 2.0; p["fall_angle"] = 20.0 * M_PI / 180.0;
  p["timing_mod_coef"] = 0.4; p["facing_weight"] = 1.5; p["ef_scale"] = 1.0; p["limit_length"] = 300.0;
  p["max_turning"] = 0.5; p["filter_end_phase"] = 0.4; p["inv_fric"] = 1.0;
  p["mass"] = 60.0; p["inertia_x"] = 9.683; p["inertia_y"] = 2.056; p["inertia_z"] = 10.209; p["grav"] = 9.8;
  p["envs_per_warp"] = 0.0;     // launch shape, not physics: 0 = full warps (see launch_step), else 1 .. 32 / lanes_per_env
  p["noise_q"] = 0.5; p["noise_pos"] = 0.05; p["noise_dq"] = 0.5; p["noise_vx"] = 0.05; p["noise_vy"] = 0.5; p["noise_vz"] = 0.05;
}

// Physics.LQR for A=[[0,1],[-k/M,-b/M]], B=[0,1/M]^T, Q=diag(q,0), R=1 (PhysicsLib/SDRE.cpp:100-245): the stabilising
// solution of the 2x2 continuous ARE in closed form
static void lqr_gain(double M, double b, double k, double q, double* K0, double* K1) {
  const double a0 = k / M, a1 = b / M, be = 1.0 / M;
  const double p2 = (-a0 + std::sqrt(a0 * a0 + be * be * q)) / (be * be);
  const double p3 = (-a1 + std::sqrt(a1 * a1 + be * be * (2.0 * p2))) / (be * be);
  *K0 = be * p2; *K1 = be * p3;
}

static void qrot_h(const double q[4], const double v[3], double o[3]) {
  const double u[3] = {q[1], q[2], q[3]}, s = q[0];
  const double uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2], uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  const double c[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
  for (int i = 0; i < 3; i++) o[i] = 2.0 * uv * u[i] + (s * s - uu) * v[i] + 2.0 * s * c[i];
}

template <typename real> static Params<real> make_params(cdm_env* h) {
  auto& p = h->hp;
  Params<real> q;
  memset(&q, 0, sizeof(q));
  q.k_p = (real)p["k_p"]; q.k_d = (real)p["k_d"]; q.acc_clamp = (real)p["acc_clamp"]; q.max_leg_len = (real)p["max_leg_len"];
  q.unlimited_leglen = p["unlimited_leglen"] != 0.0; q.contact_thr = (real)p["contact_thr"]; q.touch_down_height = (real)p["touch_down_height"];
  q.max_foot_vel = (real)p["max_foot_vel"]; q.stride_thr = (real)p["stride_thr"]; q.max_contact_len = (real)p["max_contact_len"];
  q.min_swing_phase = (real)p["min_swing_phase"]; q.max_contact_y = (real)p["max_contact_y"]; q.healthy_min = (real)p["healthy_min"];
  q.healthy_max = (real)p["healthy_max"]; q.fall_angle = (real)p["fall_angle"]; q.timing_mod_coef = (real)p["timing_mod_coef"];
  q.facing_weight = (real)p["facing_weight"]; q.ef_scale = (real)p["ef_scale"]; q.limit_length = (real)p["limit_length"];
  q.max_turning = (real)p["max_turning"]; q.filter_end_phase = (real)p["filter_end_phase"];
  double K0 = p["filter_K0"], K1 = p["filter_K1"];
  if (p["use_lqr_cache"] == 0.0) lqr_gain(30.0, 0.001, 1.0, std::pow(10.0, p["logQ"]), &K0, &K1);   // MassParticle3D(30, B, k, 10^logQ, 0, RL_step/2), :501-515
  q.K0 = (real)K0; q.K1 = (real)K1;
  q.inv_ridge2 = 2.0 * p["w_ddq"] / p["w_lam"];
  q.mass = (real)p["mass"]; q.inertia[0] = (real)p["inertia_x"]; q.inertia[1] = (real)p["inertia_y"]; q.inertia[2] = (real)p["inertia_z"];
  q.grav = (real)p["grav"];
  // friction basis (QPservo:_buildProblem, module/QPservo_v3.lua:956-968): the normal and four directions tilted by atan(1/invFricCoef)
  const double ang = std::atan(1.0 / p["inv_fric"]);
  const double Y[3] = {0, 1, 0};
  const double qs[5][4] = {{1, 0, 0, 0}, {std::cos(-ang / 2), std::sin(-ang / 2), 0, 0}, {std::cos(-ang / 2), 0, 0, std::sin(-ang / 2)},
                           {std::cos(ang / 2), std::sin(ang / 2), 0, 0}, {std::cos(ang / 2), 0, 0, std::sin(ang / 2)}};
  for (int d = 0; d < 5; d++) {
    double o[3];
    qrot_h(qs[d], Y, o);
    if (d == 0) { o[0] = 0; o[1] = 1; o[2] = 0; }
    for (int i = 0; i < 3; i++) q.basis[d][i] = (real)o[i];
  }
  for (int i = 0; i < 3; i++) { q.root_off[i] = (real)h->skel.root_off[i]; q.toe[i] = (real)h->skel.toe[i]; q.heel[i] = (real)h->skel.heel[i]; }
  for (int j = 0; j < 6; j++) {
    for (int i = 0; i < 3; i++) q.leg_off[j][i] = (real)h->skel.leg_off[j][i];
    q.leg_nchan[j] = h->skel.leg_nchan[j]; q.leg_axes[j] = h->skel.leg_axes[j];
  }
  q.noise_q = (real)p["noise_q"]; q.noise_pos = (real)p["noise_pos"]; q.noise_dq = (real)p["noise_dq"];
  q.noise_v[0] = (real)p["noise_vx"]; q.noise_v[1] = (real)p["noise_vy"]; q.noise_v[2] = (real)p["noise_vz"];
  return q;
}
template <typename real> static ResetConst<real> make_rc(cdm_env* h) {
  ResetConst<real> r;
  for (int i = 0; i < 7; i++) { r.cdm0[i] = (real)h->rc.cdm0[i]; r.dcdm0[i] = (real)h->rc.dcdm0[i]; }
  for (int l = 0; l < 2; l++) { for (int i = 0; i < 3; i++) r.contact0[l][i] = (real)h->rc.contact0[l][i]; r.start_swing[l] = h->rc.start_swing[l]; }
  return r;
}

// reset constants from the table: reference markers of the start frame (model:setRefTree + getJointStateFromRefTree)
__global__ void cdm_reset_const_kernel(const double* tab, int F, Params<double> P, ResetConst<double>* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ResetConst<double> r;
  V3<double> mk[4], cp; Q4<double> cq;
  ref_markers<double>(tab, F, 0.0, P, mk, &cp, &cq);
  for (int i = 0; i < 7; i++) r.cdm0[i] = tab[T_CDM + i];
  // DMotionDOF row 0: body twist [v(0:3), 0, w(4:7)]
  r.dcdm0[0] = tab[T_DV]; r.dcdm0[1] = tab[T_DV + 1]; r.dcdm0[2] = tab[T_DV + 2]; r.dcdm0[3] = 0.0;
  r.dcdm0[4] = tab[T_DW]; r.dcdm0[5] = tab[T_DW + 1]; r.dcdm0[6] = tab[T_DW + 2];
  const int con = (int)tab[T_CON];
  for (int l = 0; l < 2; l++) {
    V3<double> c = 0.5 * mk[2 * l] + 0.5 * mk[2 * l + 1];
    r.contact0[l][0] = c.x - r.cdm0[0]; r.contact0[l][1] = 0.0; r.contact0[l][2] = c.z - r.cdm0[2];
    r.start_swing[l] = ((con >> l) & 1) ? 0 : 1;
  }
  *out = r;
}

static int cdm_create_device_state(cdm_env* h);
extern "C" int cdm_destroy(cdm_env* h);

extern "C" int cdm_create(const cdm_config* cfg, cdm_env** out) {
  if (!cfg || !out) return srb_set_error(-1, "cdm_create: null argument");
  if (cfg->num_envs <= 0) return srb_set_error(-1, "cdm_create: num_envs must be positive");
  if (cfg->precision != 32 && cfg->precision != 64) return srb_set_error(-1, "cdm_create: precision must be 32 or 64");
  int G = cfg->lanes_per_env == 0 ? 1 : cfg->lanes_per_env;   // thread per environment + structured QP: measured best at 16384 envs (tools/bench_cdm.py)
  if (G != 1 && G != 2 && G != 4 && G != 8) return srb_set_error(-1, "cdm_create: lanes_per_env must be 1,2,4 or 8");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return srb_set_error(-3, std::string("cdm_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return srb_set_error(-1, "cdm_create: bad device ordinal");
  CK(cudaSetDevice(cfg->device));
  cdm_env* h = new cdm_env();
  h->cfg = *cfg; h->N = cfg->num_envs; h->G = G;
  h->Npad = (h->N + 31) / 32 * 32;
  h->rsz = cfg->precision == 64 ? 8 : 4;
  default_params(h->hp);
  memset(&h->skel, 0, sizeof(h->skel));
  memset(&h->ground, 0, sizeof(h->ground));
  const int rc = cdm_create_device_state(h);
  if (rc) { cdm_destroy(h); return rc; }                   // nothing half-built survives a failed allocation
  *out = h;
  return 0;
}

static int cdm_create_device_state(cdm_env* h) {
  CK(cudaMalloc(&h->st, h->rsz * NF_REAL * h->Npad));
  CK(cudaMemset(h->st, 0, h->rsz * NF_REAL * h->Npad));
  CK(cudaMalloc(&h->ist, sizeof(int) * NF_INT * h->Npad));
  CK(cudaMemset(h->ist, 0, sizeof(int) * NF_INT * h->Npad));
  CK(cudaMalloc(&h->hist, h->rsz * (size_t)h->N * HIST_SLOTS * HIST_W));
  CK(cudaMemset(h->hist, 0, h->rsz * (size_t)h->N * HIST_SLOTS * HIST_W));
  CK(cudaMalloc(&h->ep_stats, sizeof(double) * 3));
  CK(cudaMemset(h->ep_stats, 0, sizeof(double) * 3));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_state, cudaEventDisableTiming));
  return 0;
}

extern "C" int cdm_destroy(cdm_env* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->st); cudaFree(h->ist); cudaFree(h->hist); cudaFree(h->tab); cudaFree(h->ep_stats);
  cudaFreeHost(h->h_act); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_rew); cudaFreeHost(h->h_done);
  cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_rew); cudaFree(h->d_done);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev_state) cudaEventDestroy(h->ev_state);
  delete h;
  return 0;
}

extern "C" int cdm_set_param(cdm_env* h, const char* name, double value) {
  if (!h || !name) return srb_set_error(-1, "cdm_set_param: null argument");
  auto it = h->hp.find(name);
  if (it == h->hp.end()) return srb_set_error(-1, std::string("cdm_set_param: unknown parameter '") + name + "'");
  it->second = value;
  return 0;
}
extern "C" int cdm_get_param(cdm_env* h, const char* name, double* value) {
  if (!h || !name || !value) return srb_set_error(-1, "cdm_get_param: null argument");
  auto it = h->hp.find(name);
  if (it == h->hp.end()) return srb_set_error(-1, std::string("cdm_get_param: unknown parameter '") + name + "'");
  *value = it->second;
  return 0;
}

extern "C" int cdm_upload_table(cdm_env* h, int32_t nframes, int32_t nf_cycle, const double* rows, const cdm_skeleton* skel) {
  if (!h || !rows || !skel) return srb_set_error(-1, "cdm_upload_table: null argument");
  if (nframes < 8 || nf_cycle < 1 || nf_cycle * 2 + 2 >= nframes) return srb_set_error(-1, "cdm_upload_table: the stitched table must hold more than two cycles");
  for (int j = 0; j < 6; j++) if (skel->leg_nchan[j] < 1 || skel->leg_nchan[j] > 3) return srb_set_error(-1, "cdm_upload_table: leg joints need 1..3 Euler channels");
  CK(cudaSetDevice(h->cfg.device));
  h->skel = *skel; h->have_skel = true;
  h->nframes = nframes; h->nf_cycle = nf_cycle;
  const size_t n = (size_t)nframes * TAB_W;
  double* d64 = nullptr;
  CK(cudaMalloc(&d64, sizeof(double) * n));
  CK(cudaMemcpy(d64, rows, sizeof(double) * n, cudaMemcpyHostToDevice));
  ResetConst<double>* rc_dev = nullptr;
  CK(cudaMalloc(&rc_dev, sizeof(ResetConst<double>)));
  cdm_reset_const_kernel<<<1, 32, 0, h->stream>>>(d64, nframes, make_params<double>(h), rc_dev);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(&h->rc, rc_dev, sizeof(h->rc), cudaMemcpyDeviceToHost));
  cudaFree(rc_dev);
  cudaFree(h->tab); h->tab = nullptr;
  if (h->cfg.precision == 64) {
    h->tab = d64;
  } else {
    std::vector<float> f(n);
    for (size_t i = 0; i < n; i++) f[i] = (float)rows[i];
    CK(cudaMalloc(&h->tab, sizeof(float) * n));
    CK(cudaMemcpy(h->tab, f.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
    cudaFree(d64);
  }
  return 0;
}

template <typename real> static void do_reset(cdm_env* h, const int* idx_dev, int count, const double* plan_dev, float* obs_dev, cudaStream_t s) {
  ResetArgs<real> A;
  A.N = h->N; A.Npad = h->Npad; A.count = count; A.nframes = h->nframes; A.nf_cycle = h->nf_cycle; A.idx = idx_dev;
  A.obs_w = h->obs_w; A.ground = h->ground;
  A.st = (real*)h->st; A.ist = h->ist; A.hist = (real*)h->hist; A.tab = (const real*)h->tab;
  A.plan = plan_dev; A.seed = h->cfg.seed; A.obs = obs_dev; A.p = make_params<real>(h); A.rc = make_rc<real>(h);
  cdm_reset_kernel<real><<<(count + 127) / 128, 128, 0, s>>>(A);
  h->launches++;
}

extern "C" int cdm_reset(cdm_env* h, const int32_t* env_idx_host, int32_t count, const double* plan_host, float* obs_dev, void* stream) {
  if (!h) return srb_set_error(-1, "cdm_reset: null handle");
  if (!h->tab) return srb_set_error(-1, "cdm_reset: no reference table uploaded");
  if (count <= 0 || count > h->N) return srb_set_error(-1, "cdm_reset: bad count");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int* idx_dev = nullptr; double* plan_dev = nullptr;
  // temporaries are released on every path, also when a later allocation or copy fails
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaFree(idx_dev); cudaFree(plan_dev); return srb_set_error(-2, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  if (env_idx_host) {
    for (int i = 0; i < count; i++) if (env_idx_host[i] < 0 || env_idx_host[i] >= h->N) return srb_set_error(-1, "cdm_reset: env index out of range");
    CKR(cudaMalloc(&idx_dev, sizeof(int) * count));
    CKR(cudaMemcpyAsync(idx_dev, env_idx_host, sizeof(int) * count, cudaMemcpyHostToDevice, s));
  }
  if (plan_host) {
    CKR(cudaMalloc(&plan_dev, sizeof(double) * PLAN_W * count));
    CKR(cudaMemcpyAsync(plan_dev, plan_host, sizeof(double) * PLAN_W * count, cudaMemcpyHostToDevice, s));
  }
  if (h->cfg.precision == 64) do_reset<double>(h, idx_dev, count, plan_dev, obs_dev, s);
  else do_reset<float>(h, idx_dev, count, plan_dev, obs_dev, s);
  CKR(cudaGetLastError());
  CKR(cudaEventRecord(h->ev_state, s));
  if (idx_dev || plan_dev) {
    CKR(cudaStreamSynchronize(s));
    cudaFree(idx_dev); cudaFree(plan_dev);
  }
#undef CKR
  return 0;
}

template <typename real, int G> static void launch_step(cdm_env* h, const float* act, float* obs, float* rew, uint8_t* done, cudaStream_t s) {
  StepArgs<real> A;
  A.N = h->N; A.Npad = h->Npad; A.nframes = h->nframes; A.nf_cycle = h->nf_cycle;
  A.obs_w = h->obs_w; A.ground = h->ground;
  A.st = (real*)h->st; A.ist = h->ist; A.hist = (real*)h->hist; A.tab = (const real*)h->tab;
  A.act = act; A.obs = obs; A.rew = rew; A.done = done;
  A.term_obs = h->term_obs; A.ep_stats = h->ep_stats; A.reset_plan = h->reset_plan; A.dbg = h->dbg;
  A.auto_reset = h->cfg.auto_reset; A.seed = h->cfg.seed; A.p = make_params<real>(h); A.rc = make_rc<real>(h);
  constexpr int EPW = 32 / G;
  // environments per warp: full warps unless envs_per_warp says otherwise.  Partially filled warps were measured as an answer to
  // branch divergence (a warp runs the union of its environments' code paths) and LOSE: 16384 envs, policy workload, fp32:
  // 32 per warp 131 us, 16: 169 us, 12: 236 us, 8: 292 us (profiles/r02_cdm_envs_per_warp.jsonl) -- the step is bound by
  // instruction fetch / issue over all warps, not by the length of one warp's stream.
  int gpw = (int)h->hp["envs_per_warp"];
  if (gpw <= 0) gpw = EPW;
  A.gpw = gpw < EPW ? gpw : EPW;
  if (h->ground.T.h != nullptr) cdm_step_kernel<real, G, true><<<(h->N + A.gpw - 1) / A.gpw, 32, 0, s>>>(A);
  else cdm_step_kernel<real, G, false><<<(h->N + A.gpw - 1) / A.gpw, 32, 0, s>>>(A);
  h->launches++;
}
template <typename real> static int dispatch_step(cdm_env* h, const float* act, float* obs, float* rew, uint8_t* done, cudaStream_t s) {
  switch (h->G) {
    case 1: launch_step<real, 1>(h, act, obs, rew, done, s); break;
    case 2: launch_step<real, 2>(h, act, obs, rew, done, s); break;
    case 4: launch_step<real, 4>(h, act, obs, rew, done, s); break;
    case 8: launch_step<real, 8>(h, act, obs, rew, done, s); break;
    default: return srb_set_error(-1, "bad lanes_per_env");
  }
  return 0;
}

extern "C" int cdm_step(cdm_env* h, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, void* stream) {
  if (!h) return srb_set_error(-1, "cdm_step: null handle");
  if (!act_dev || !obs_dev || !rew_dev || !done_dev) return srb_set_error(-1, "cdm_step: null buffer");
  if (!h->tab) return srb_set_error(-1, "cdm_step: no reference table uploaded");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int rc = h->cfg.precision == 64 ? dispatch_step<double>(h, act_dev, obs_dev, rew_dev, done_dev, s) : dispatch_step<float>(h, act_dev, obs_dev, rew_dev, done_dev, s);
  if (rc) return rc;
  CK(cudaGetLastError());
  if (s != h->stream) CK(cudaEventRecord(h->ev_state, s));
  return 0;
}

static int ensure_host_buffers(cdm_env* h) {
  if (h->h_act) return 0;
  size_t N = h->N;
  CK(cudaMallocHost(&h->h_act, sizeof(float) * N * ACT_W));
  CK(cudaMallocHost(&h->h_obs, sizeof(float) * N * OBS_W_TERRAIN));
  CK(cudaMallocHost(&h->h_rew, sizeof(float) * N));
  CK(cudaMallocHost(&h->h_done, N));
  CK(cudaMalloc(&h->d_act, sizeof(float) * N * ACT_W));
  CK(cudaMalloc(&h->d_obs, sizeof(float) * N * OBS_W_TERRAIN));
  CK(cudaMalloc(&h->d_rew, sizeof(float) * N));
  CK(cudaMalloc(&h->d_done, N));
  return 0;
}
extern "C" int cdm_pinned_buffers(cdm_env* h, float** act, float** obs, float** rew, uint8_t** done) {
  if (!h) return srb_set_error(-1, "cdm_pinned_buffers: null handle");
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_host_buffers(h);
  if (rc) return rc;
  if (act) *act = h->h_act; if (obs) *obs = h->h_obs; if (rew) *rew = h->h_rew; if (done) *done = h->h_done;
  return 0;
}
extern "C" int cdm_step_host(cdm_env* h, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host) {
  if (!h) return srb_set_error(-1, "cdm_step_host: null handle");
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_host_buffers(h);
  if (rc) return rc;
  size_t N = h->N;
  if (act_host && act_host != h->h_act) memcpy(h->h_act, act_host, sizeof(float) * N * ACT_W);
  cudaStream_t s = h->stream;
  CK(cudaStreamWaitEvent(s, h->ev_state, 0));            // resets / device steps issued on caller streams come first
  CK(cudaMemcpyAsync(h->d_act, h->h_act, sizeof(float) * N * ACT_W, cudaMemcpyHostToDevice, s));
  rc = cdm_step(h, h->d_act, h->d_obs, h->d_rew, h->d_done, s);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_obs, h->d_obs, sizeof(float) * N * h->obs_w, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->h_rew, h->d_rew, sizeof(float) * N, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->h_done, h->d_done, N, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (obs_host && obs_host != h->h_obs) memcpy(obs_host, h->h_obs, sizeof(float) * N * h->obs_w);
  if (rew_host && rew_host != h->h_rew) memcpy(rew_host, h->h_rew, sizeof(float) * N);
  if (done_host && done_host != h->h_done) memcpy(done_host, h->h_done, N);
  return 0;
}

extern "C" int cdm_set_aux(cdm_env* h, float* term_obs_dev, const double* reset_plan_dev, double* dbg_dev) {
  if (!h) return srb_set_error(-1, "cdm_set_aux: null handle");
  h->term_obs = term_obs_dev; h->reset_plan = reset_plan_dev; h->dbg = dbg_dev;
  return 0;
}

extern "C" int cdm_set_terrain(cdm_env* h, tl_terrain* t, const double terrain_pos[3]) {
  if (!h) return srb_set_error(-1, "cdm_set_terrain: null handle");
  if (t == nullptr) { memset(&h->ground, 0, sizeof(h->ground)); h->obs_w = OBS_W; return 0; }
  if (!terrain_pos) return srb_set_error(-1, "cdm_set_terrain: null terrain_pos");
  if (t->device != h->cfg.device) return srb_set_error(-1, "cdm_set_terrain: terrain lives on another device");
  h->ground.T = t->T; h->ground.px = terrain_pos[0]; h->ground.py = terrain_pos[1]; h->ground.pz = terrain_pos[2];
  h->obs_w = OBS_W_TERRAIN;
  return 0;
}
extern "C" int cdm_obs_dim(cdm_env* h) { return h ? h->obs_w : 0; }

extern "C" int cdm_state_dims(int32_t* n_real, int32_t* n_int, int32_t* hist_slots, int32_t* hist_width) {
  if (n_real) *n_real = NF_REAL; if (n_int) *n_int = NF_INT; if (hist_slots) *hist_slots = HIST_SLOTS; if (hist_width) *hist_width = HIST_W;
  return 0;
}
template <typename real> static int get_state_t(cdm_env* h, int e, double* reals, int32_t* ints, double* hist) {
  std::vector<real> col(NF_REAL);
  CK(cudaMemcpy2D(col.data(), sizeof(real), (const real*)h->st + e, sizeof(real) * h->Npad, sizeof(real), NF_REAL, cudaMemcpyDeviceToHost));
  for (int i = 0; i < NF_REAL; i++) reals[i] = (double)col[i];
  CK(cudaMemcpy2D(ints, sizeof(int), h->ist + e, sizeof(int) * h->Npad, sizeof(int), NF_INT, cudaMemcpyDeviceToHost));
  if (hist) {
    std::vector<real> hb(HIST_SLOTS * HIST_W);
    CK(cudaMemcpy(hb.data(), (const real*)h->hist + (size_t)e * HIST_SLOTS * HIST_W, sizeof(real) * hb.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < hb.size(); i++) hist[i] = (double)hb[i];
  }
  return 0;
}
template <typename real> static int set_state_t(cdm_env* h, int e, const double* reals, const int32_t* ints, const double* hist) {
  std::vector<real> col(NF_REAL);
  for (int i = 0; i < NF_REAL; i++) col[i] = (real)reals[i];
  CK(cudaMemcpy2D((real*)h->st + e, sizeof(real) * h->Npad, col.data(), sizeof(real), sizeof(real), NF_REAL, cudaMemcpyHostToDevice));
  CK(cudaMemcpy2D(h->ist + e, sizeof(int) * h->Npad, ints, sizeof(int), sizeof(int), NF_INT, cudaMemcpyHostToDevice));
  if (hist) {
    std::vector<real> hb(HIST_SLOTS * HIST_W);
    for (size_t i = 0; i < hb.size(); i++) hb[i] = (real)hist[i];
    CK(cudaMemcpy((real*)h->hist + (size_t)e * HIST_SLOTS * HIST_W, hb.data(), sizeof(real) * hb.size(), cudaMemcpyHostToDevice));
  }
  return 0;
}
extern "C" int cdm_get_state(cdm_env* h, int32_t e, double* reals, int32_t* ints, double* hist) {
  if (!h || !reals || !ints) return srb_set_error(-1, "cdm_get_state: null argument");
  if (e < 0 || e >= h->N) return srb_set_error(-1, "cdm_get_state: env index out of range");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  return h->cfg.precision == 64 ? get_state_t<double>(h, e, reals, ints, hist) : get_state_t<float>(h, e, reals, ints, hist);
}
extern "C" int cdm_set_state(cdm_env* h, int32_t e, const double* reals, const int32_t* ints, const double* hist) {
  if (!h || !reals || !ints) return srb_set_error(-1, "cdm_set_state: null argument");
  if (e < 0 || e >= h->N) return srb_set_error(-1, "cdm_set_state: env index out of range");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  return h->cfg.precision == 64 ? set_state_t<double>(h, e, reals, ints, hist) : set_state_t<float>(h, e, reals, ints, hist);
}

extern "C" int cdm_episode_stats(cdm_env* h, double out[3], int32_t clear) {
  if (!h || !out) return srb_set_error(-1, "cdm_episode_stats: null argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, h->ep_stats, sizeof(double) * 3, cudaMemcpyDeviceToHost));
  if (clear) CK(cudaMemset(h->ep_stats, 0, sizeof(double) * 3));
  return 0;
}
extern "C" int64_t cdm_launch_count(cdm_env* h) { return h ? h->launches : 0; }

// ---- OBJloader::Terrain (gym_cdm2 terrain spec) -------------------------------------------------------
extern "C" int tlterrain_create(int32_t device, int32_t rows, int32_t cols, double size_x, double size_z, int32_t tile_along_z,
                                const double* heights_host, tl_terrain** out) {
  if (!heights_host || !out) return srb_set_error(-1, "tlterrain_create: null argument");
  if (rows < 2 || cols < 2 || !(size_x > 0) || !(size_z > 0)) return srb_set_error(-1, "tlterrain_create: bad dimensions");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return srb_set_error(-3, std::string("tlterrain_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if (device < 0 || device >= ndev) return srb_set_error(-1, "tlterrain_create: bad device ordinal");
  CK(cudaSetDevice(device));
  tl_terrain* t = new tl_terrain();
  t->device = device;
  {
    cudaError_t ce = cudaMalloc(&t->h, sizeof(double) * (size_t)rows * cols);
    if (ce == cudaSuccess) ce = cudaMemcpy(t->h, heights_host, sizeof(double) * (size_t)rows * cols, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(t->h); delete t; return srb_set_error(-2, std::string("tlterrain_create: ") + cudaGetErrorString(ce)); }
  }
  t->T.h = t->h; t->T.rows = rows; t->T.cols = cols; t->T.size_x = size_x; t->T.size_z = size_z; t->T.tile = tile_along_z ? 1 : 0;
  *out = t;
  return 0;
}
extern "C" int tlterrain_destroy(tl_terrain* t) {
  if (!t) return 0;
  cudaSetDevice(t->device);
  cudaFree(t->h);
  delete t;
  return 0;
}
extern "C" int tlterrain_height(tl_terrain* t, int64_t n, const double* xz_dev, double* height_dev, int32_t* hit_dev, void* stream) {
  if (!t || !xz_dev || !height_dev) return srb_set_error(-1, "tlterrain_height: null argument");
  if (n <= 0) return 0;
  CK(cudaSetDevice(t->device));
  tl_terrain_height_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(t->T, n, xz_dev, height_dev, hit_dev);
  CK(cudaGetLastError());
  return 0;
}
