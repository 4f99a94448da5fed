// C-ABI of the batched SRBTrack step (declared in include/srbtrack_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <cmath>
#include <algorithm>

#include "../../include/srbtrack_b200.h"
#include "srb_kernels.cuh"

using namespace srb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int srb_set_error(int code, const std::string& msg) { return fail(code, msg); }   // shared with fk_api.cu
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// ---- exchange block of one rank: IPC-shared device memory, two slots used alternately (epoch parity) ----------------------------
enum { XCHG_W = 1 + 2 * OBS_W + 3, XCHG_MAX_RANKS = 16 };
struct XchgSlot { double v[XCHG_W]; unsigned long long flag; unsigned long long pad[7]; };
struct XchgBlock { XchgSlot slot[2]; };
struct XchgPeers { XchgBlock* p[XCHG_MAX_RANKS]; };

struct ClipStore { void* frames = nullptr; float* feat = nullptr; int nframes = 0, cycle = 0, is_stand = 0; };

struct srb_env {
  srb_config cfg;
  int N = 0, Npad = 0, G = 8;
  size_t rsz = 8;
  void* st = nullptr; int* ist = nullptr; void* hist = nullptr;
  void* clips_dev = nullptr;            // ClipDesc<real>[MAX_CLIPS]
  std::vector<ClipStore> clips;
  int nclips = 0;
  double* ep_stats = nullptr;
  const float* contact_prob = nullptr; float* term_obs = nullptr; const double* reset_plan = nullptr; double* dbg = nullptr;
  Params<double> p;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_state = nullptr;        // recorded after every state-mutating launch on a caller stream; srb_step_host's stream waits on it
  // pinned host staging + device I/O for the host-buffer path
  float *h_act = nullptr, *h_obs = nullptr, *h_rew = nullptr, *h_rinfo = nullptr; uint8_t* h_done = nullptr;
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr, *d_rinfo = nullptr; uint8_t* d_done = nullptr;
  cudaStream_t stream2 = nullptr;       // second stream of the host path: window gather + its D2H copy
  cudaEvent_t ev_win = nullptr, ev_dma = nullptr;
  cudaEvent_t tl[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // timeline of the last split host step (timing events)
  int tl_on = 0; float tl_ms[6] = {0, 0, 0, 0, 0, 0};
  int split_obs = 0; uint8_t* done_dev_arg = nullptr;   // per-launch arguments of the split host path
  int64_t launches = 0;
  int host_mode = 1;                    // srb_step_host: 1 = zero copy (default), 2 = split (windows by copy engine), 0 = staged copies
  // terrain (variant 2)
  TerrainDev* terrain_dev = nullptr; std::vector<float*> hf_data; double* box_dev = nullptr;
  // fused statistics + all-reduce over NVLink peer memory (srb_rollout_stats_allreduce)
  int x_rank = -1, x_world = 0; unsigned long long x_epoch = 0;
  XchgBlock* x_mine = nullptr; XchgBlock* x_peer[XCHG_MAX_RANKS] = {nullptr}; double* x_scratch = nullptr; int* x_err = nullptr;
};
static const int MAX_CLIPS = 64;

template <typename real> static Params<real> convert_params(const Params<double>& p) {
  Params<real> q;
  q.h = (real)p.h; q.nsub = p.nsub; q.mass = (real)p.mass;
  for (int i = 0; i < 3; i++) q.inertia[i] = (real)p.inertia[i];
  for (int i = 0; i < 6; i++) q.impulse_force[i] = (real)p.impulse_force[i];
  q.armature = (real)p.armature; q.damping = (real)p.damping; q.grav = (real)p.grav; q.mu_b = (real)p.mu_b;
  q.kp = (real)p.kp; q.kd = (real)p.kd; q.w_lambda = (real)p.w_lambda; q.force_thr = (real)p.force_thr;
  for (int i = 0; i < 9; i++) q.w[i] = (real)p.w[i];
  q.use_filter = p.use_filter; q.use_clamp = p.use_clamp; q.use_impulse = p.use_impulse; q.obs_after_inc = p.obs_after_inc; q.use_terrain = p.use_terrain;
  q.impulse_interval = p.impulse_interval; q.impulse_duration = p.impulse_duration; q.impulse_trigger_at_end = p.impulse_trigger_at_end;
  q.max_frame = p.max_frame; q.aset_warm = p.aset_warm;
  q.inv_w_lambda = (real)(1.0 / p.w_lambda);
  q.clamp_bound = (real)(1.01 * std::sqrt(1.0 + (double)q.mu_b * (double)q.mu_b));
  const double Md[6] = {p.mass + p.armature, p.mass + p.armature, p.mass + p.armature,
                        p.inertia[0] + p.armature, p.inertia[1] + p.armature, p.inertia[2] + p.armature};
  for (int i = 0; i < 6; i++) { q.inv_M[i] = (real)(1.0 / Md[i]); q.inv_Mh[i] = (real)(1.0 / (Md[i] + p.h * p.damping)); }
  return q;
}

static void default_params(Params<double>& p, int variant) {
  // gym_trackSRB/Characters/SRB5.xml:1-37
  const double hx = 0.19557607215607947, hy = 0.25406692031825, hz = 0.667757440991862;
  p.h = 0.0083332;
  p.nsub = (int)std::lround(0.033333 / 0.0083332);     // round(self.mocap_dt / self.model.opt.timestep)
  p.mass = 60.0;
  p.inertia[0] = p.mass / 3.0 * (hy * hy + hz * hz);
  p.inertia[1] = p.mass / 3.0 * (hx * hx + hz * hz);
  p.inertia[2] = p.mass / 3.0 * (hx * hx + hy * hy);
  p.armature = 1.0; p.damping = 1.0; p.grav = 9.81;
  p.mu_b = 1.2 * (1.2 * 1.0);                           // b = (1.2*mu, 0, 1), mu = 1.2 * floor friction
  p.kp = 120.0; p.kd = 35.0; p.w_lambda = 0.001; p.force_thr = 10000.0;
  p.max_frame = 512;
  p.aset_warm = getenv("SRB_ASET_WARM") ? atoi(getenv("SRB_ASET_WARM")) : 1;
  p.use_terrain = 0;
  for (int i = 0; i < 6; i++) p.impulse_force[i] = 0.0;
  if (variant == SRB_VARIANT_TEST || variant == SRB_VARIANT_TERRAIN) {
    const double w[9] = {5, 0.1, 0.5, 0, 50, 20, 700, 50, 1};
    memcpy(p.w, w, sizeof(w));
    p.use_filter = 1; p.use_clamp = 1; p.use_impulse = 1; p.obs_after_inc = 0;
    p.impulse_interval = 1000000000; p.impulse_duration = 24; p.impulse_trigger_at_end = 1;
    p.impulse_force[0] = 700; p.impulse_force[1] = 700; p.impulse_force[2] = -350;
    if (variant == SRB_VARIANT_TERRAIN) { p.use_clamp = 0; p.use_terrain = 1; }     // env/SRBEnv5_mocap_list_terrain_window.py
  } else {
    const double w[9] = {5, 1, 0.5, 0.01, 5, 1.5, 25, 1.5, 0.1};
    memcpy(p.w, w, sizeof(w));
    p.use_filter = 0; p.use_clamp = 0; p.use_impulse = 0; p.obs_after_inc = 1;
    p.impulse_interval = 100000000; p.impulse_duration = 24; p.impulse_trigger_at_end = 0;
    p.impulse_force[0] = 1000 / std::sqrt(2.0) * 0.5; p.impulse_force[1] = -1000 / std::sqrt(2.0) * 0.5; p.impulse_force[2] = 0;
  }
}

extern "C" const char* srb_last_error(void) { return g_err.c_str(); }

static int create_device_state(srb_env* h);
extern "C" int srb_destroy(srb_env* h);

extern "C" int srb_create(const srb_config* cfg, srb_env** out) {
  if (!cfg || !out) return fail(-1, "srb_create: null argument");
  if (cfg->num_envs <= 0) return fail(-1, "srb_create: num_envs must be positive");
  if (cfg->precision != 32 && cfg->precision != 64) return fail(-1, "srb_create: precision must be 32 or 64");
  // default lanes per environment by batch size (measured, DESIGN 5 / profiles/r02_lanes_sweep.txt): 8 lanes keep the dependent chain
  // of a warp short while the batch leaves schedulers idle; from ~3-4 k environments on, 4 lanes (structured QP: a lane's ten columns
  // share their linear part) issue fewer instructions per environment and win at every larger size measured (up to 65536; 2 lanes
  // never do).  fp32 switches after 4096 because the host-buffer path (coalesced-I/O kernel) still prefers 8 lanes there.
  int G = cfg->lanes_per_env;
  if (G == 0) G = cfg->num_envs <= (cfg->precision == 64 ? 3072 : 4096) ? 8 : 4;
  if (G != 1 && G != 2 && G != 4 && G != 8 && G != 16 && G != 32) return fail(-1, "srb_create: lanes_per_env must be 1,2,4,8,16 or 32");
  if (cfg->variant != SRB_VARIANT_TEST && cfg->variant != SRB_VARIANT_TRAINING && cfg->variant != SRB_VARIANT_TERRAIN) return fail(-1, "srb_create: unknown variant");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(-3, std::string("srb_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(-1, "srb_create: bad device ordinal");
  CK(cudaSetDevice(cfg->device));
  srb_env* h = new srb_env();
  h->cfg = *cfg; h->N = cfg->num_envs; h->G = G;
  h->Npad = (h->N + 31) / 32 * 32;
  h->rsz = cfg->precision == 64 ? 8 : 4;
  default_params(h->p, cfg->variant);
  const int rc = create_device_state(h);
  if (rc) { srb_destroy(h); return rc; }                  // nothing half-built survives a failed allocation
  *out = h;
  return 0;
}

static int create_device_state(srb_env* h) {
  CK(cudaMalloc(&h->st, h->rsz * NF_REAL * h->Npad));
  CK(cudaMemset(h->st, 0, h->rsz * NF_REAL * h->Npad));
  CK(cudaMalloc(&h->ist, sizeof(int) * NF_INT * h->Npad));
  CK(cudaMemset(h->ist, 0, sizeof(int) * NF_INT * h->Npad));
  CK(cudaMalloc(&h->hist, h->rsz * (size_t)h->N * HIST_SLOTS * HIST_W));
  CK(cudaMemset(h->hist, 0, h->rsz * (size_t)h->N * HIST_SLOTS * HIST_W));
  CK(cudaMalloc(&h->clips_dev, sizeof(ClipDesc<double>) * MAX_CLIPS));
  CK(cudaMemset(h->clips_dev, 0, sizeof(ClipDesc<double>) * MAX_CLIPS));
  h->clips.resize(MAX_CLIPS);
  CK(cudaMalloc(&h->ep_stats, sizeof(double) * 3));
  CK(cudaMemset(h->ep_stats, 0, sizeof(double) * 3));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_state, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_win, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_dma, cudaEventDisableTiming));
  for (int i = 0; i < 5; i++) CK(cudaEventCreate(&h->tl[i]));
  return 0;
}

extern "C" int srb_destroy(srb_env* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->st); cudaFree(h->ist); cudaFree(h->hist); cudaFree(h->clips_dev); cudaFree(h->ep_stats);
  for (auto& c : h->clips) { cudaFree(c.frames); cudaFree(c.feat); }
  for (auto p : h->hf_data) cudaFree(p);
  cudaFree(h->box_dev); cudaFree(h->terrain_dev);
  cudaFreeHost(h->h_act); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_rew); cudaFreeHost(h->h_rinfo); cudaFreeHost(h->h_done);
  cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_rew); cudaFree(h->d_rinfo); cudaFree(h->d_done);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev_state) cudaEventDestroy(h->ev_state);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->ev_win) cudaEventDestroy(h->ev_win);
  if (h->ev_dma) cudaEventDestroy(h->ev_dma);
  for (int i = 0; i < 5; i++) if (h->tl[i]) cudaEventDestroy(h->tl[i]);
  for (int r = 0; r < h->x_world; r++) if (r != h->x_rank && h->x_peer[r]) cudaIpcCloseMemHandle(h->x_peer[r]);
  cudaFree(h->x_mine); cudaFree(h->x_scratch); cudaFree(h->x_err);
  delete h;
  return 0;
}

extern "C" int srb_upload_clip(srb_env* h, int32_t clip_id, int32_t nframes, int32_t cycle_length, int32_t is_stand,
                               const double* pos_com, const double* ori_com, const double* vel_com, const double* root_quat,
                               const double* feet_pos, const double* feet_yaw) {
  if (!h) return fail(-1, "srb_upload_clip: null handle");
  if (clip_id < 0 || clip_id >= MAX_CLIPS) return fail(-1, "srb_upload_clip: clip_id out of range");
  if (nframes <= 0 || !pos_com || !ori_com || !vel_com || !root_quat || !feet_pos || !feet_yaw) return fail(-1, "srb_upload_clip: bad arguments");
  CK(cudaSetDevice(h->cfg.device));
  std::vector<double> rec((size_t)nframes * CLIP_W);
  for (int f = 0; f < nframes; f++) {
    double* r = &rec[(size_t)f * CLIP_W];
    for (int i = 0; i < 3; i++) r[i] = pos_com[3 * f + i];
    for (int i = 0; i < 9; i++) r[3 + i] = ori_com[9 * f + i];
    for (int i = 0; i < 6; i++) r[12 + i] = vel_com[6 * f + i];
    for (int i = 0; i < 6; i++) r[18 + i] = feet_pos[6 * f + i];
    for (int i = 0; i < 4; i++) r[24 + i] = feet_yaw[4 * f + i];
    for (int i = 0; i < 4; i++) r[28 + i] = root_quat[4 * f + i];
  }
  ClipStore& cs = h->clips[clip_id];
  cudaFree(cs.frames); cudaFree(cs.feat);
  cs.frames = nullptr; cs.feat = nullptr;
  // features are always computed from the fp64 table
  double* dfr = nullptr;
  CK(cudaMalloc(&dfr, sizeof(double) * rec.size()));
  CK(cudaMemcpy(dfr, rec.data(), sizeof(double) * rec.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cs.feat, sizeof(float) * (size_t)nframes * FEAT_W));
  srb_clip_features_kernel<<<(nframes + 127) / 128, 128, 0, h->stream>>>(dfr, nframes, cs.feat);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (h->cfg.precision == 64) {
    cs.frames = dfr;
  } else {
    std::vector<float> recf(rec.size());
    for (size_t i = 0; i < rec.size(); i++) recf[i] = (float)rec[i];
    CK(cudaMalloc(&cs.frames, sizeof(float) * recf.size()));
    CK(cudaMemcpy(cs.frames, recf.data(), sizeof(float) * recf.size(), cudaMemcpyHostToDevice));
    cudaFree(dfr);
  }
  cs.nframes = nframes; cs.cycle = cycle_length; cs.is_stand = is_stand;
  ClipDesc<double> d;   // same layout for float/double instantiations (pointer, pointer, 4 ints)
  d.frames = (const double*)cs.frames; d.feat = cs.feat; d.nframes = nframes; d.cycle_length = cycle_length; d.is_stand = is_stand; d.pad = 0;
  CK(cudaMemcpy((char*)h->clips_dev + sizeof(ClipDesc<double>) * clip_id, &d, sizeof(d), cudaMemcpyHostToDevice));
  if (clip_id + 1 > h->nclips) h->nclips = clip_id + 1;
  return 0;
}

template <typename real> static int do_reset(srb_env* h, const int* idx_dev, int count, const double* plan_dev, float* obs_dev, cudaStream_t s) {
  ResetArgs<real> A;
  A.N = h->N; A.Npad = h->Npad; A.count = count; A.idx = idx_dev;
  A.st = (real*)h->st; A.ist = h->ist; A.hist = (real*)h->hist;
  A.clips = (const ClipDesc<real>*)h->clips_dev; A.nclips = h->nclips;
  A.plan = plan_dev; A.obs = obs_dev; A.seed = h->cfg.seed; A.p = convert_params<real>(h->p); A.terrain = h->terrain_dev;
  srb_reset_kernel<real><<<(count + 127) / 128, 128, 0, s>>>(A);
  h->launches++;
  return 0;
}

extern "C" int srb_reset(srb_env* h, const int32_t* env_idx_host, int32_t count, const double* plan_host, float* obs_dev, void* stream) {
  if (!h) return fail(-1, "srb_reset: null handle");
  if (h->nclips == 0) return fail(-1, "srb_reset: no clip uploaded");
  if (h->p.use_terrain && !h->terrain_dev) return fail(-1, "srb_reset: terrain variant needs srb_upload_terrain first");
  if (count <= 0 || count > h->N) return fail(-1, "srb_reset: bad count");
  for (int c = 0; c < h->nclips; c++) if (h->clips[c].frames == nullptr) return fail(-1, "srb_reset: clip ids must be contiguous from 0");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int* idx_dev = nullptr; double* plan_dev = nullptr;
  // temporaries are released on every path (CKR), also when a later allocation or copy fails
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaFree(idx_dev); cudaFree(plan_dev); return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  if (env_idx_host) {
    for (int i = 0; i < count; i++) if (env_idx_host[i] < 0 || env_idx_host[i] >= h->N) return fail(-1, "srb_reset: env index out of range");
    CKR(cudaMalloc(&idx_dev, sizeof(int) * count));
    CKR(cudaMemcpyAsync(idx_dev, env_idx_host, sizeof(int) * count, cudaMemcpyHostToDevice, s));
  }
  if (plan_host) {
    CKR(cudaMalloc(&plan_dev, sizeof(double) * 9 * count));
    CKR(cudaMemcpyAsync(plan_dev, plan_host, sizeof(double) * 9 * count, cudaMemcpyHostToDevice, s));
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

template <typename real, int G, bool CIO = false, bool TERRAIN = false> static void launch_step(srb_env* h, const float* act, float* obs, float* rew, uint8_t* done, float* rinfo, cudaStream_t s) {
  StepArgs<real> A;
  A.N = h->N; A.Npad = h->Npad;
  A.st = (real*)h->st; A.ist = h->ist; A.hist = (real*)h->hist;
  A.clips = (const ClipDesc<real>*)h->clips_dev; A.nclips = h->nclips;
  A.act = act; A.obs = obs; A.rew = rew; A.done = done; A.rinfo = rinfo; A.term_obs = h->term_obs;
  A.contact_prob = h->contact_prob; A.reset_plan = h->reset_plan; A.dbg = h->dbg; A.ep_stats = h->ep_stats;
  A.seed = h->cfg.seed; A.auto_reset = h->cfg.auto_reset; A.p = convert_params<real>(h->p); A.terrain = h->terrain_dev;
  A.split_obs = CIO ? h->split_obs : 0; A.done_dev = CIO ? h->done_dev_arg : nullptr;
  constexpr int EPW = 32 / G;
  int blocks = (h->N + EPW - 1) / EPW;
  srb_step_kernel<real, G, CIO, TERRAIN><<<blocks, 32, 0, s>>>(A);
  h->launches++;
}

static bool cio_supported(int G) { return G == 4 || G == 8; }
template <typename real> static int dispatch_step(srb_env* h, const float* act, float* obs, float* rew, uint8_t* done, float* rinfo, cudaStream_t s, bool cio = false) {
  if (h->p.use_terrain) {                                // terrain variant: separate instantiations
    if (h->G == 8) { launch_step<real, 8, false, true>(h, act, obs, rew, done, rinfo, s); return 0; }
    if (h->G == 4) { launch_step<real, 4, false, true>(h, act, obs, rew, done, rinfo, s); return 0; }
    return fail(-1, "the terrain variant is built for lanes_per_env 4 and 8");
  }
  if (cio) {
    if (h->G == 8) { launch_step<real, 8, true>(h, act, obs, rew, done, rinfo, s); return 0; }
    if (h->G == 4) { launch_step<real, 4, true>(h, act, obs, rew, done, rinfo, s); return 0; }
    return fail(-1, "coalesced-I/O step is built for lanes_per_env 4 and 8");
  }
  switch (h->G) {
    case 1: launch_step<real, 1>(h, act, obs, rew, done, rinfo, s); break;
    case 2: launch_step<real, 2>(h, act, obs, rew, done, rinfo, s); break;
    case 4: launch_step<real, 4>(h, act, obs, rew, done, rinfo, s); break;
    case 8: launch_step<real, 8>(h, act, obs, rew, done, rinfo, s); break;
    case 16: launch_step<real, 16>(h, act, obs, rew, done, rinfo, s); break;
    case 32: launch_step<real, 32>(h, act, obs, rew, done, rinfo, s); break;
    default: return fail(-1, "bad lanes_per_env");
  }
  return 0;
}

static int step_impl(srb_env* h, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, float* rinfo_dev, void* stream, bool cio);
extern "C" int srb_step(srb_env* h, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, float* rinfo_dev, void* stream) {
  return step_impl(h, act_dev, obs_dev, rew_dev, done_dev, rinfo_dev, stream, false);
}
static int step_impl(srb_env* h, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, float* rinfo_dev, void* stream, bool cio) {
  if (!h) return fail(-1, "srb_step: null handle");
  if (!act_dev || !obs_dev || !rew_dev || !done_dev) return fail(-1, "srb_step: null buffer");
  if (h->nclips == 0) return fail(-1, "srb_step: no clip uploaded");
  if (h->p.use_terrain && !h->terrain_dev) return fail(-1, "srb_step: terrain variant needs srb_upload_terrain first");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int rc = h->cfg.precision == 64 ? dispatch_step<double>(h, act_dev, obs_dev, rew_dev, done_dev, rinfo_dev, s, cio)
                                  : dispatch_step<float>(h, act_dev, obs_dev, rew_dev, done_dev, rinfo_dev, s, cio);
  if (rc) return rc;
  CK(cudaGetLastError());
  if (s != h->stream) CK(cudaEventRecord(h->ev_state, s));
  return 0;
}

static int ensure_host_buffers(srb_env* h) {
  if (h->h_act) return 0;
  size_t N = h->N;
  CK(cudaHostAlloc(&h->h_act, sizeof(float) * N * ACT_W, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h->h_obs, sizeof(float) * N * OBS_W, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h->h_rew, sizeof(float) * N, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h->h_rinfo, sizeof(float) * N * RINFO_W, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h->h_done, N, cudaHostAllocMapped));
  CK(cudaMalloc(&h->d_act, sizeof(float) * N * ACT_W));
  CK(cudaMalloc(&h->d_obs, sizeof(float) * N * OBS_W));
  CK(cudaMalloc(&h->d_rew, sizeof(float) * N));
  CK(cudaMalloc(&h->d_rinfo, sizeof(float) * N * RINFO_W));
  CK(cudaMalloc(&h->d_done, N));
  return 0;
}

template <typename real> static void launch_window(srb_env* h, int after_inc, const uint8_t* only_done, const float* self_src, float* dst, cudaStream_t s, int patch_only = 0) {
  srb_obs_window_kernel<real><<<h->N, 160, 0, s>>>(h->N, h->Npad, h->ist, (const ClipDesc<real>*)h->clips_dev, after_inc, only_done, self_src, dst, patch_only);
  h->launches++;
}

extern "C" int srb_pinned_buffers(srb_env* h, float** act, float** obs, float** rew, uint8_t** done, float** rinfo) {
  if (!h) return fail(-1, "srb_pinned_buffers: null handle");
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_host_buffers(h);
  if (rc) return rc;
  if (act) *act = h->h_act; if (obs) *obs = h->h_obs; if (rew) *rew = h->h_rew; if (done) *done = h->h_done; if (rinfo) *rinfo = h->h_rinfo;
  return 0;
}

extern "C" int srb_step_host(srb_env* h, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host, float* rinfo_host) {
  if (!h) return fail(-1, "srb_step_host: null handle");
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_host_buffers(h);
  if (rc) return rc;
  size_t N = h->N;
  if (act_host && act_host != h->h_act) memcpy(h->h_act, act_host, sizeof(float) * N * ACT_W);
  cudaStream_t s = h->stream;
  CK(cudaStreamWaitEvent(s, h->ev_state, 0));            // resets / device steps issued on caller streams come first
  if (h->host_mode == 2 && cio_supported(h->G) && !h->p.use_terrain) {
    // split: 83 % of an observation row is the look-ahead window, which depends on (clip, frame) only.  It is gathered first
    // (3 us) into the device observation array, and that array travels to the host as ONE contiguous copy-engine transfer on a
    // second stream WHILE the step kernel computes (SM-issued stores to mapped memory reach ~35 GB/s, the copy engine ~52 GB/s).
    // The step kernel reads the actions from and writes reward / done / reward_info to the mapped pinned buffers (zero copy) and
    // leaves the 25 self features of every row in the device array; a patch launch ordered after the big copy writes those self
    // features, and the new windows of the environments that were auto-reset, over the rows in host memory.
    float *a = nullptr, *o = nullptr, *r = nullptr, *ri = nullptr; uint8_t* d = nullptr;
    CK(cudaHostGetDevicePointer((void**)&a, h->h_act, 0)); CK(cudaHostGetDevicePointer((void**)&o, h->h_obs, 0));
    CK(cudaHostGetDevicePointer((void**)&r, h->h_rew, 0)); CK(cudaHostGetDevicePointer((void**)&d, h->h_done, 0));
    CK(cudaHostGetDevicePointer((void**)&ri, h->h_rinfo, 0));
    CK(cudaStreamWaitEvent(h->stream2, h->ev_state, 0));
    if (h->tl_on) CK(cudaEventRecord(h->tl[0], h->stream2));
    if (h->cfg.precision == 64) launch_window<double>(h, h->p.obs_after_inc, nullptr, nullptr, h->d_obs, h->stream2);
    else launch_window<float>(h, h->p.obs_after_inc, nullptr, nullptr, h->d_obs, h->stream2);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_win, h->stream2));
    if (h->tl_on) CK(cudaEventRecord(h->tl[1], h->stream2));
    CK(cudaMemcpyAsync(h->h_obs, h->d_obs, sizeof(float) * N * OBS_W, cudaMemcpyDeviceToHost, h->stream2));
    CK(cudaEventRecord(h->ev_dma, h->stream2));
    if (h->tl_on) CK(cudaEventRecord(h->tl[2], h->stream2));
    CK(cudaStreamWaitEvent(s, h->ev_win, 0));            // the step must not advance the frame counters before they were read
    h->split_obs = 1; h->done_dev_arg = h->d_done;
    rc = step_impl(h, a, h->d_obs, r, d, ri, s, true);
    h->split_obs = 0; h->done_dev_arg = nullptr;
    if (rc) return rc;
    if (h->tl_on) CK(cudaEventRecord(h->tl[3], s));
    CK(cudaStreamWaitEvent(s, h->ev_dma, 0));
    const uint8_t* only_done = h->cfg.auto_reset ? h->d_done : nullptr;   // without auto-reset no window changes
    if (h->cfg.precision == 64) launch_window<double>(h, 0, only_done, h->d_obs, o, s);
    else launch_window<float>(h, 0, only_done, h->d_obs, o, s);
    CK(cudaGetLastError());
    if (h->tl_on) CK(cudaEventRecord(h->tl[4], s));
    CK(cudaStreamSynchronize(s));
    if (h->tl_on) for (int i = 1; i < 5; i++) cudaEventElapsedTime(&h->tl_ms[i], h->tl[0], h->tl[i]);
  } else if (h->host_mode == 3 && cio_supported(h->G) && !h->p.use_terrain) {
    // split, second form: only the WINDOW columns (125 of 150 floats per row) travel by copy engine, as one pitched transfer that
    // overlaps the step kernel; the kernel stores the 25 self features, reward / done / reward_info straight into the mapped host
    // arrays (disjoint from what the copy engine writes), so nothing but the new windows of auto-reset environments is left to patch.
    float *a = nullptr, *o = nullptr, *r = nullptr, *ri = nullptr; uint8_t* d = nullptr;
    CK(cudaHostGetDevicePointer((void**)&a, h->h_act, 0)); CK(cudaHostGetDevicePointer((void**)&o, h->h_obs, 0));
    CK(cudaHostGetDevicePointer((void**)&r, h->h_rew, 0)); CK(cudaHostGetDevicePointer((void**)&d, h->h_done, 0));
    CK(cudaHostGetDevicePointer((void**)&ri, h->h_rinfo, 0));
    CK(cudaStreamWaitEvent(h->stream2, h->ev_state, 0));
    if (h->cfg.precision == 64) launch_window<double>(h, h->p.obs_after_inc, nullptr, nullptr, h->d_obs, h->stream2);
    else launch_window<float>(h, h->p.obs_after_inc, nullptr, nullptr, h->d_obs, h->stream2);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_win, h->stream2));
    CK(cudaMemcpy2DAsync(h->h_obs + FEAT_W, sizeof(float) * OBS_W, h->d_obs + FEAT_W, sizeof(float) * OBS_W, sizeof(float) * (OBS_W - FEAT_W), N,
                         cudaMemcpyDeviceToHost, h->stream2));
    CK(cudaEventRecord(h->ev_dma, h->stream2));
    CK(cudaStreamWaitEvent(s, h->ev_win, 0));            // the step must not advance the frame counters before they were read
    h->split_obs = 1; h->done_dev_arg = h->d_done;
    rc = step_impl(h, a, o, r, d, ri, s, true);
    h->split_obs = 0; h->done_dev_arg = nullptr;
    if (rc) return rc;
    if (h->cfg.auto_reset) {
      CK(cudaStreamWaitEvent(s, h->ev_dma, 0));
      if (h->cfg.precision == 64) launch_window<double>(h, 0, h->d_done, nullptr, o, s, 1);
      else launch_window<float>(h, 0, h->d_done, nullptr, o, s, 1);
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(h->stream2));
  } else if (h->host_mode == 1 && cio_supported(h->G) && !h->p.use_terrain) {
    // zero copy: the step kernel loads the actions from and stores its results to the mapped pinned buffers; every such
    // access is a contiguous 16-byte-vectorised block per warp (see srb_step_kernel), i.e. full-size PCIe transactions
    // that overlap the computation of the other warps instead of four copies serialised after it
    float *a = nullptr, *o = nullptr, *r = nullptr, *ri = nullptr; uint8_t* d = nullptr;
    CK(cudaHostGetDevicePointer((void**)&a, h->h_act, 0)); CK(cudaHostGetDevicePointer((void**)&o, h->h_obs, 0));
    CK(cudaHostGetDevicePointer((void**)&r, h->h_rew, 0)); CK(cudaHostGetDevicePointer((void**)&d, h->h_done, 0));
    CK(cudaHostGetDevicePointer((void**)&ri, h->h_rinfo, 0));
    rc = step_impl(h, a, o, r, d, ri, s, true);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s));
  } else {
    CK(cudaMemcpyAsync(h->d_act, h->h_act, sizeof(float) * N * ACT_W, cudaMemcpyHostToDevice, s));
    rc = srb_step(h, h->d_act, h->d_obs, h->d_rew, h->d_done, h->d_rinfo, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->h_obs, h->d_obs, sizeof(float) * N * OBS_W, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->h_rew, h->d_rew, sizeof(float) * N, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->h_done, h->d_done, N, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->h_rinfo, h->d_rinfo, sizeof(float) * N * RINFO_W, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  if (obs_host && obs_host != h->h_obs) memcpy(obs_host, h->h_obs, sizeof(float) * N * OBS_W);
  if (rew_host && rew_host != h->h_rew) memcpy(rew_host, h->h_rew, sizeof(float) * N);
  if (done_host && done_host != h->h_done) memcpy(done_host, h->h_done, N);
  if (rinfo_host && rinfo_host != h->h_rinfo) memcpy(rinfo_host, h->h_rinfo, sizeof(float) * N * RINFO_W);
  return 0;
}

// Timeline of the last split (mode 2) host step, milliseconds after the window gather was enqueued: [1] window gather done,
// [2] observation copy done, [3] step kernel done, [4] patch launch done.  enable = 1 switches the timing events on.
extern "C" int srb_host_timeline(srb_env* h, int32_t enable, float out_ms[5]) {
  if (!h) return fail(-1, "srb_host_timeline: null handle");
  h->tl_on = enable;
  if (out_ms) for (int i = 0; i < 5; i++) out_ms[i] = h->tl_ms[i];
  return 0;
}

extern "C" int srb_set_host_mode(srb_env* h, int32_t mode) {
  if (!h) return fail(-1, "srb_set_host_mode: null handle");
  if (mode < 0 || mode > 3) return fail(-1, "srb_set_host_mode: mode must be 0 (staged copies), 1 (zero copy), 2 (split) or 3 (split, window columns only)");
  h->host_mode = mode;
  return 0;
}

extern "C" int srb_set_aux(srb_env* h, const float* contact_prob_dev, float* term_obs_dev, const double* reset_plan_dev, double* dbg_dev) {
  if (!h) return fail(-1, "srb_set_aux: null handle");
  h->contact_prob = contact_prob_dev; h->term_obs = term_obs_dev; h->reset_plan = reset_plan_dev; h->dbg = dbg_dev;
  return 0;
}

extern "C" int srb_set_reward_weights(srb_env* h, const double w[9]) {
  if (!h || !w) return fail(-1, "srb_set_reward_weights: null argument");
  memcpy(h->p.w, w, sizeof(double) * 9);
  return 0;
}

extern "C" int srb_set_impulse(srb_env* h, int32_t interval, int32_t duration, const double* force, int32_t force_len) {
  if (!h || !force) return fail(-1, "srb_set_impulse: null argument");
  if (interval <= 0) return fail(-1, "srb_set_impulse: interval must be positive");
  if (force_len < 0 || force_len > 6) return fail(-1, "srb_set_impulse: force has at most 6 components (force, torque)");
  h->p.impulse_interval = interval; h->p.impulse_duration = duration;
  for (int i = 0; i < 6; i++) h->p.impulse_force[i] = i < force_len ? force[i] : 0.0;
  h->p.use_impulse = h->cfg.variant != SRB_VARIANT_TRAINING;   // the training env never calls update_impulse (..._training.py:366-367)
  return 0;
}

extern "C" int srb_state_dims(int32_t* n_real, int32_t* n_int, int32_t* hist_slots, int32_t* hist_width) {
  if (n_real) *n_real = NF_REAL; if (n_int) *n_int = NF_INT; if (hist_slots) *hist_slots = HIST_SLOTS; if (hist_width) *hist_width = HIST_W;
  return 0;
}

template <typename real> static int get_state_t(srb_env* h, int e, double* reals, int32_t* ints, double* hist) {
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
template <typename real> static int set_state_t(srb_env* h, int e, const double* reals, const int32_t* ints, const double* hist) {
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

extern "C" int srb_get_state(srb_env* h, int32_t e, double* reals, int32_t* ints, double* hist) {
  if (!h || !reals || !ints) return fail(-1, "srb_get_state: null argument");
  if (e < 0 || e >= h->N) return fail(-1, "srb_get_state: env index out of range");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  return h->cfg.precision == 64 ? get_state_t<double>(h, e, reals, ints, hist) : get_state_t<float>(h, e, reals, ints, hist);
}
extern "C" int srb_set_state(srb_env* h, int32_t e, const double* reals, const int32_t* ints, const double* hist) {
  if (!h || !reals || !ints) return fail(-1, "srb_set_state: null argument");
  if (e < 0 || e >= h->N) return fail(-1, "srb_set_state: env index out of range");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  return h->cfg.precision == 64 ? set_state_t<double>(h, e, reals, ints, hist) : set_state_t<float>(h, e, reals, ints, hist);
}

// ---- observation statistics: acc[0] += N, acc[1+j] += sum_i obs[i][j], acc[1+150+j] += sum_i obs[i][j]^2
__global__ void srb_obs_stats_kernel(const float* __restrict__ obs, int N, double* __restrict__ acc) {
  // one CTA handles a strip of rows; thread j (<150) owns column j -> coalesced row reads
  int j = threadIdx.x;
  int rows_per = (N + gridDim.x - 1) / gridDim.x;
  int r0 = blockIdx.x * rows_per, r1 = min(N, r0 + rows_per);
  if (j < OBS_W) {
    double s = 0, q = 0;
    for (int r = r0; r < r1; r++) { double v = (double)obs[(size_t)r * OBS_W + j]; s += v; q += v * v; }
    if (r1 > r0) { atomicAdd(acc + 1 + j, s); atomicAdd(acc + 1 + OBS_W + j, q); }
  }
  if (j == 0 && r1 > r0) atomicAdd(acc, (double)(r1 - r0));
}

extern "C" int srb_obs_stats(srb_env* h, const float* obs_dev, double* acc_dev, void* stream) {
  if (!h || !obs_dev || !acc_dev) return fail(-1, "srb_obs_stats: null argument");
  CK(cudaSetDevice(h->cfg.device));
  int blocks = h->N < 148 * 4 ? (h->N + 7) / 8 : 148 * 4;
  if (blocks < 1) blocks = 1;
  srb_obs_stats_kernel<<<blocks, 160, 0, (cudaStream_t)stream>>>(obs_dev, h->N, acc_dev);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ---- the path's one collective as ONE kernel over NVLink / NVSwitch peer memory ------------------------------------------------
// acc = sum over ranks of [count, sum x, sum x^2 of this rank's observation rows | episode return / length / count sums].
// Phase 1 (all CTAs): column sums of the rank's obs rows into a scratch vector.  Phase 2 (the CTA that finishes last): publish the
// vector in this rank's exchange slot (release store of the epoch into the slot's flag, system scope), then for every rank in rank
// order: wait for that rank's flag (acquire load through the peer mapping) and add its vector with volatile peer loads -- every rank
// sums in the same order, so all ranks hold bit-identical results.  Slots alternate with the epoch's parity: a rank can be at most
// one collective ahead of the slowest reader of its slot (it needs that reader's NEXT flag to get further), so two slots suffice.
// A flag that never arrives (a peer died) ends the wait after ~1 s and sets the error word instead of hanging the device.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__global__ void srb_stats_allreduce_kernel(const float* __restrict__ obs, int N, const double* __restrict__ ep_stats, double* __restrict__ scratch,
                                           XchgPeers peers, int rank, int world, unsigned long long epoch, double* __restrict__ acc, int* __restrict__ err) {
  const int j = threadIdx.x;
  const int rows_per = (N + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(N, r0 + rows_per);
  if (j < OBS_W) {
    double s = 0, q = 0;
    for (int r = r0; r < r1; r++) { double v = (double)obs[(size_t)r * OBS_W + j]; s += v; q += v * v; }
    if (r1 > r0) { atomicAdd(scratch + 1 + j, s); atomicAdd(scratch + 1 + OBS_W + j, q); }
  }
  if (j == 0 && r1 > r0) atomicAdd(scratch, (double)(r1 - r0));
  __shared__ int last;
  __threadfence();
  __syncthreads();
  if (j == 0) last = atomicAdd(reinterpret_cast<int*>(scratch + XCHG_W), 1) == (int)gridDim.x - 1;     // ticket counter behind the vector
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int sl = (int)(epoch & 1ull);
  XchgSlot* mine = &peers.p[rank]->slot[sl];
  for (int k = j; k < XCHG_W; k += blockDim.x) {
    const double v = k < 1 + 2 * OBS_W ? __ldcg(scratch + k) : ep_stats[k - (1 + 2 * OBS_W)];
    mine->v[k] = v;
    if (k < 1 + 2 * OBS_W) scratch[k] = 0.0;                       // ready for the next call
  }
  if (j == 0) *reinterpret_cast<int*>(scratch + XCHG_W) = 0;
  __threadfence_system();
  __syncthreads();
  if (j == 0) st_release_sys(&mine->flag, epoch);
  double tot[(XCHG_W + 159) / 160];
#pragma unroll
  for (int i = 0; i < (XCHG_W + 159) / 160; i++) tot[i] = 0.0;
  __shared__ int bad;
  if (j == 0) bad = 0;
  for (int r = 0; r < world; r++) {
    const XchgSlot* ps = &peers.p[r]->slot[sl];
    if (j == 0 && r != rank) {
      unsigned long long spins = 0;
      while (ld_acquire_sys(&ps->flag) < epoch) {
        __nanosleep(200);
        if (++spins > (1ull << 22)) { bad = 1; break; }
      }
    }
    __syncthreads();
    if (bad) break;
#pragma unroll
    for (int i = 0; i < (XCHG_W + 159) / 160; i++) {
      const int k = j + i * 160;
      if (k < XCHG_W) tot[i] += *reinterpret_cast<const volatile double*>(&ps->v[k]);
    }
  }
  if (bad) { if (j == 0) atomicExch(err, 1); return; }
#pragma unroll
  for (int i = 0; i < (XCHG_W + 159) / 160; i++) { const int k = j + i * 160; if (k < XCHG_W) acc[k] = tot[i]; }
}

extern "C" int srb_xchg_create(srb_env* h, int32_t rank, int32_t world, void* handle_out) {
  if (!h || !handle_out) return fail(-1, "srb_xchg_create: null argument");
  if (world < 1 || world > XCHG_MAX_RANKS || rank < 0 || rank >= world) return fail(-1, "srb_xchg_create: rank / world out of range");
  if (h->x_mine) return fail(-1, "srb_xchg_create: already created");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMalloc(&h->x_mine, sizeof(XchgBlock)));
  CK(cudaMemset(h->x_mine, 0, sizeof(XchgBlock)));
  CK(cudaMalloc(&h->x_scratch, sizeof(double) * (XCHG_W + 1)));
  CK(cudaMemset(h->x_scratch, 0, sizeof(double) * (XCHG_W + 1)));
  CK(cudaMalloc(&h->x_err, sizeof(int)));
  CK(cudaMemset(h->x_err, 0, sizeof(int)));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mh;
  CK(cudaIpcGetMemHandle(&mh, h->x_mine));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &mh, 64);
  h->x_rank = rank; h->x_world = world; h->x_epoch = 0;
  h->x_peer[rank] = h->x_mine;
  return 0;
}

extern "C" int srb_xchg_connect(srb_env* h, const void* handles) {
  if (!h || !handles || !h->x_mine) return fail(-1, "srb_xchg_connect: create the exchange first");
  CK(cudaSetDevice(h->cfg.device));
  for (int r = 0; r < h->x_world; r++) {
    if (r == h->x_rank) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    h->x_peer[r] = (XchgBlock*)p;
  }
  return 0;
}

extern "C" int srb_rollout_stats_allreduce(srb_env* h, const float* obs_dev, double* acc_dev, void* stream) {
  if (!h || !obs_dev || !acc_dev) return fail(-1, "srb_rollout_stats_allreduce: null argument");
  if (!h->x_mine) return fail(-1, "srb_rollout_stats_allreduce: no exchange (srb_xchg_create / srb_xchg_connect)");
  for (int r = 0; r < h->x_world; r++) if (!h->x_peer[r]) return fail(-1, "srb_rollout_stats_allreduce: exchange not connected");
  CK(cudaSetDevice(h->cfg.device));
  XchgPeers P;
  for (int r = 0; r < XCHG_MAX_RANKS; r++) P.p[r] = r < h->x_world ? h->x_peer[r] : nullptr;
  int blocks = h->N < 148 * 4 ? (h->N + 7) / 8 : 148 * 4;
  if (blocks < 1) blocks = 1;
  h->x_epoch += 1;
  srb_stats_allreduce_kernel<<<blocks, 160, 0, (cudaStream_t)stream>>>(obs_dev, h->N, h->ep_stats, h->x_scratch, P, h->x_rank, h->x_world, h->x_epoch, acc_dev, h->x_err);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int srb_xchg_status(srb_env* h, int32_t* err_out) {
  if (!h || !err_out || !h->x_err) return fail(-1, "srb_xchg_status: no exchange");
  CK(cudaSetDevice(h->cfg.device));
  int e = 0;
  CK(cudaMemcpy(&e, h->x_err, sizeof(int), cudaMemcpyDeviceToHost));
  *err_out = e;
  return 0;
}

// ---- pack_obs_extra (gym_trackSRB/taesooLib/DeltaExtractor.py:1015-1036): the SRB state in taesooLib's Y-up convention for the
// SRB -> full-body mapping.  In the reference the foot-centre sites' xpos / xmat are overwritten with the feet of the action in
// every sub-step (env/SRBEnv5_mocap_list_window.py:58-87), i.e. they ARE left/right_foot_contact_pos_global / _ori.
// ZtoY (work/rendermodule.py:631-636): vector (x, y, z) -> (y, z, x), quaternion (w, x, y, z) -> (w, y, z, x).
template <typename real>
__global__ void srb_pack_obs_extra_kernel(const real* __restrict__ st, const int* __restrict__ ist, int N, int Npad, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
#define LD(f) (double)st[(size_t)(f) * Npad + e]
  double* o = out + (size_t)e * 29;
  // pose_tl: transf(quat, pos).ZtoY() written as translation, rotation
  const double px = LD(F_POS), py = LD(F_POS + 1), pz = LD(F_POS + 2);
  const double qw = LD(F_QUAT), qx = LD(F_QUAT + 1), qy = LD(F_QUAT + 2), qz = LD(F_QUAT + 3);
  o[0] = py; o[1] = pz; o[2] = px; o[3] = qw; o[4] = qy; o[5] = qz; o[6] = qx;
  // dq = [ q_yup * ZtoY(body angular velocity), ZtoY(world linear velocity) ]   (vector3::rotate, no normalisation)
  const double wx = LD(F_QVEL + 4), wy = LD(F_QVEL + 5), wz = LD(F_QVEL + 3);          // ZtoY of qvel[3:6]
  const double s = qw, ux = qy, uy = qz, uz = qx;
  const double a = 2.0 * (ux * wx + uy * wy + uz * wz), b = s * s - (ux * ux + uy * uy + uz * uz), c2 = 2.0 * s;
  o[7] = a * ux + b * wx + c2 * (uy * wz - uz * wy);
  o[8] = a * uy + b * wy + c2 * (uz * wx - ux * wz);
  o[9] = a * uz + b * wz + c2 * (ux * wy - uy * wx);
  o[10] = LD(F_QVEL + 1); o[11] = LD(F_QVEL + 2); o[12] = LD(F_QVEL);
  const int flags = ist[(size_t)I_FLAGS * Npad + e];
#pragma unroll
  for (int f = 0; f < 2; f++) {
    const int fp = f == 0 ? F_FOOTL : F_FOOTR, fy = f == 0 ? F_FYAWL : F_FYAWR;
    const double c = LD(fy), sn = LD(fy + 1);
    // RE.toQuater(site_xmat) = quater::setRotation(matrix) on [[c,-s,0],[s,c,0],[0,0,1]] (quater.cpp:341-386), then ZtoY
    double w, z;
    const double tr = c + c + 1.0;
    if (tr > 0.0) { double r = sqrt(tr + 1.0); w = r * 0.5; r = 0.5 / r; z = (sn - (-sn)) * r; }
    else { double r = sqrt((1.0 - (c + c)) + 1.0); z = r * 0.5; r = 0.5 / r; w = (sn - (-sn)) * r; }
    double* fo = o + 13 + 7 * f;
    fo[0] = LD(fp + 1); fo[1] = LD(fp + 2); fo[2] = LD(fp);
    fo[3] = w; fo[4] = 0.0; fo[5] = z; fo[6] = 0.0;
  }
  o[27] = (flags & FLAG_L) ? 1.0 : 0.0;
  o[28] = (flags & FLAG_R) ? 1.0 : 0.0;
#undef LD
}

extern "C" int srb_pack_obs_extra(srb_env* h, double* out_dev, void* stream) {
  if (!h || !out_dev) return fail(-1, "srb_pack_obs_extra: null argument");
  CK(cudaSetDevice(h->cfg.device));
  const int blocks = (h->N + 127) / 128;
  if (h->cfg.precision == 64) srb_pack_obs_extra_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>((const double*)h->st, h->ist, h->N, h->Npad, out_dev);
  else srb_pack_obs_extra_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>((const float*)h->st, h->ist, h->N, h->Npad, out_dev);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int srb_episode_stats(srb_env* h, double out[3], int32_t clear) {
  if (!h || !out) return fail(-1, "srb_episode_stats: null argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, h->ep_stats, sizeof(double) * 3, cudaMemcpyDeviceToHost));
  if (clear) CK(cudaMemset(h->ep_stats, 0, sizeof(double) * 3));
  return 0;
}

extern "C" int srb_episode_stats_device(srb_env* h, double* out_dev, void* stream) {
  if (!h || !out_dev) return fail(-1, "srb_episode_stats_device: null argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(out_dev, h->ep_stats, sizeof(double) * 3, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int srb_upload_terrain(srb_env* h, const double plane_size[2], int32_t plane_geom, int32_t n_hf, const srb_hfield* hfs,
                                  int32_t n_box, const srb_box* boxes, int32_t n_body, const srb_terrain_body* bodies) {
  if (!h || !plane_size) return fail(-1, "srb_upload_terrain: null argument");
  if (n_hf < 0 || n_hf > MAX_HF || n_body < 0 || n_body > MAX_TBODY || n_box < 0) return fail(-1, "srb_upload_terrain: too many height fields / bodies");
  if ((n_hf && !hfs) || (n_box && !boxes) || (n_body && !bodies)) return fail(-1, "srb_upload_terrain: null table");
  // 1. validate and build everything on the host; the installed terrain is untouched until the new one is complete
  for (int i = 0; i < n_hf; i++)
    if (!hfs[i].data_host || hfs[i].nrow < 2 || hfs[i].ncol < 2) return fail(-1, "srb_upload_terrain: bad height field");
  TerrainDev T;
  memset(&T, 0, sizeof(T));
  T.enabled = 1; T.n_hf = n_hf; T.n_body = n_body; T.n_box = n_box;
  T.plane_geom = plane_geom; T.plane_hx = plane_size[0]; T.plane_hy = plane_size[1];
  std::vector<double> bx((size_t)n_box * 5);
  for (int k = 0; k < n_box; k++) {
    if (k > 0 && boxes[k].geom_id != boxes[k - 1].geom_id + 1) return fail(-1, "srb_upload_terrain: boxes must be consecutive geoms in model order");
    bx[5 * k] = boxes[k].pos[0]; bx[5 * k + 1] = boxes[k].pos[1]; bx[5 * k + 2] = boxes[k].size[0]; bx[5 * k + 3] = boxes[k].size[1];
    bx[5 * k + 4] = boxes[k].pos[2] + boxes[k].size[2];
  }
  T.box_geom0 = n_box ? boxes[0].geom_id : 0;
  for (int b = 0; b < n_body; b++) {
    TerrainBodyDev& o = T.body[b];
    o.cx = bodies[b].xpos[0]; o.cy = bodies[b].xpos[1]; o.snap = bodies[b].snap; o.body_id = bodies[b].body_id;
    o.first = -1; o.count = 0; o.x0 = o.y0 = 1e300; o.x1 = o.y1 = -1e300; o.nested = 1; o.pad = 0;
    for (int k = 0; k < n_box; k++) {
      if (boxes[k].body_id != bodies[b].body_id) continue;
      if (o.first < 0) o.first = k;
      if (k != o.first + o.count) return fail(-1, "srb_upload_terrain: the boxes of a body must be contiguous");
      if (o.count > 0 && bx[5 * k + 4] < bx[5 * (k - 1) + 4]) return fail(-1, "srb_upload_terrain: box tops must ascend inside a body");
      if (o.count > 0 && (boxes[k].pos[0] != boxes[o.first].pos[0] || boxes[k].pos[1] != boxes[o.first].pos[1] ||
                          boxes[k].size[0] > boxes[k - 1].size[0] || boxes[k].size[1] > boxes[k - 1].size[1])) o.nested = 0;
      o.count++;
      o.x0 = std::min(o.x0, boxes[k].pos[0] - boxes[k].size[0]); o.x1 = std::max(o.x1, boxes[k].pos[0] + boxes[k].size[0]);
      o.y0 = std::min(o.y0, boxes[k].pos[1] - boxes[k].size[1]); o.y1 = std::max(o.y1, boxes[k].pos[1] + boxes[k].size[1]);
    }
    if (o.first < 0) o.first = 0;
  }
  // 2. new device buffers; on any failure they are released and the old terrain stays installed
  CK(cudaSetDevice(h->cfg.device));
  std::vector<float*> new_hf;
  double* new_box = nullptr;
  TerrainDev* new_dev = h->terrain_dev;
  auto drop = [&]() { for (auto p : new_hf) cudaFree(p); cudaFree(new_box); if (new_dev != h->terrain_dev) cudaFree(new_dev); };
#define CKT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { drop(); return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  for (int i = 0; i < n_hf; i++) {
    const srb_hfield& s = hfs[i];
    float* d = nullptr;
    CKT(cudaMalloc(&d, sizeof(float) * (size_t)s.nrow * s.ncol));
    new_hf.push_back(d);
    CKT(cudaMemcpy(d, s.data_host, sizeof(float) * (size_t)s.nrow * s.ncol, cudaMemcpyHostToDevice));
    HFieldDev& o = T.hf[i];
    o.data = d; o.nrow = s.nrow; o.ncol = s.ncol; o.sx = s.size[0]; o.sy = s.size[1]; o.sz = s.size[2];
    o.px = s.pos[0]; o.py = s.pos[1]; o.pz = s.pos[2]; o.geom_id = s.geom_id;
  }
  if (n_box) {
    CKT(cudaMalloc(&new_box, sizeof(double) * bx.size()));
    CKT(cudaMemcpy(new_box, bx.data(), sizeof(double) * bx.size(), cudaMemcpyHostToDevice));
  }
  T.box = new_box;
  if (!new_dev) CKT(cudaMalloc(&new_dev, sizeof(TerrainDev)));
  CKT(cudaDeviceSynchronize());                            // no launch may still read the tables that are about to be replaced
  CKT(cudaMemcpy(new_dev, &T, sizeof(T), cudaMemcpyHostToDevice));
#undef CKT
  // 3. swap
  for (auto p : h->hf_data) cudaFree(p);
  cudaFree(h->box_dev);
  h->hf_data = new_hf; h->box_dev = new_box; h->terrain_dev = new_dev;
  return 0;
}

extern "C" const void* srb_terrain_device(void* env) { return env ? ((srb_env*)env)->terrain_dev : nullptr; }

extern "C" int srb_set_planar_offsets(srb_env* h, const double* xy_host) {
  if (!h || !xy_host) return fail(-1, "srb_set_planar_offsets: null argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  for (int c = 0; c < 2; c++) {
    if (h->cfg.precision == 64) {
      std::vector<double> col(h->N);
      for (int i = 0; i < h->N; i++) col[i] = xy_host[2 * i + c];
      CK(cudaMemcpy((double*)h->st + (size_t)(F_PLANAR + c) * h->Npad, col.data(), sizeof(double) * h->N, cudaMemcpyHostToDevice));
    } else {
      std::vector<float> col(h->N);
      for (int i = 0; i < h->N; i++) col[i] = (float)xy_host[2 * i + c];
      CK(cudaMemcpy((float*)h->st + (size_t)(F_PLANAR + c) * h->Npad, col.data(), sizeof(float) * h->N, cudaMemcpyHostToDevice));
    }
  }
  return 0;
}

extern "C" int srb_terrain_height(srb_env* h, int32_t n, const double* xy_dev, double* height_dev, int32_t* hit_dev, void* stream) {
  if (!h || !xy_dev || !height_dev) return fail(-1, "srb_terrain_height: null argument");
  if (!h->terrain_dev) return fail(-1, "srb_terrain_height: no terrain uploaded");
  if (n <= 0) return 0;
  CK(cudaSetDevice(h->cfg.device));
  srb_terrain_height_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->terrain_dev, n, xy_dev, height_dev, hit_dev);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int srb_tracking_action(srb_env* h, float sigma, uint64_t seed, uint32_t step, float* act_dev, void* stream) {
  if (!h || !act_dev) return fail(-1, "srb_tracking_action: null argument");
  if (h->nclips == 0) return fail(-1, "srb_tracking_action: no clip uploaded");
  CK(cudaSetDevice(h->cfg.device));
  int blocks = (h->N + 127) / 128;
  if (h->cfg.precision == 64)
    srb_tracking_action_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>(h->N, h->Npad, (const double*)h->st, h->ist, (const ClipDesc<double>*)h->clips_dev, sigma, seed, step, act_dev, h->p.use_terrain ? h->terrain_dev : nullptr);
  else
    srb_tracking_action_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>(h->N, h->Npad, (const float*)h->st, h->ist, (const ClipDesc<float>*)h->clips_dev, sigma, seed, step, act_dev, h->p.use_terrain ? h->terrain_dev : nullptr);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int64_t srb_launch_count(srb_env* h) { return h ? h->launches : 0; }
