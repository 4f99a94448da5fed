// C-ABI of the batched FK / CoM / momentum pass (include/bonefk_b200.h).
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cmath>
#include <string.h>
#include "../../include/bonefk_b200.h"
#include "fk_kernels.cuh"

extern "C" const char* srb_last_error(void);
int srb_set_error(int code, const std::string& msg);       // defined in srb_api.cu
#define FCK(call)                                                                                     \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return srb_set_error(-2, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct fk_skeleton {
  int device = 0;
  fk::SkelConst host;
  fk::SkelConst* dev = nullptr;
};

extern "C" int fk_create(int32_t device, int32_t nb, const int32_t* parent, const double* offset, const int32_t* nchan,
                         const int32_t* axes, const double* mass, const double* local_com, const double* inertia, fk_skeleton** out) {
  if (!parent || !offset || !nchan || !axes || !mass || !local_com || !inertia || !out) return srb_set_error(-1, "fk_create: null argument");
  if (nb < 1 || nb > fk::MAXB) return srb_set_error(-1, "fk_create: nbones must be in 1..64");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return srb_set_error(-3, std::string("fk_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if (device < 0 || device >= ndev) return srb_set_error(-1, "fk_create: bad device ordinal");
  fk_skeleton* s = new fk_skeleton();
  s->device = device;
  fk::SkelConst& k = s->host;
  memset(&k, 0, sizeof(k));
  k.nb = nb;
  int nang = 0, maxd = 0;
  double total = 0;
  if (parent[0] != -1 || nchan[0] != 0) { delete s; return srb_set_error(-1, "fk_create: bone 0 must be the free root (parent -1, no Euler channels)"); }
  for (int b = 0; b < nb; b++) {
    if (b > 0 && (parent[b] < 0 || parent[b] >= b)) { delete s; return srb_set_error(-1, "fk_create: bones must be depth-first (parent index < bone index)"); }
    k.depth[b] = b == 0 ? 0 : k.depth[parent[b]] + 1;
    if (b > 0 && parent[b] != b - 1) {
      // depth-first order: the parent must lie on the path to the previous bone
      int a = b - 1;
      while (a >= 0 && a != parent[b]) a = parent[a];
      if (a != parent[b]) { delete s; return srb_set_error(-1, "fk_create: bones are not in depth-first order"); }
    }
    if (k.depth[b] >= fk::MAXD) { delete s; return srb_set_error(-1, "fk_create: tree deeper than 16"); }
    maxd = k.depth[b] > maxd ? k.depth[b] : maxd;
    if (nchan[b] < 0 || nchan[b] > 3 || (b > 0 && nchan[b] == 0)) { delete s; return srb_set_error(-1, "fk_create: joints need 1..3 Euler channels"); }
    k.nchan[b] = nchan[b];
    int ax = 0;
    for (int c = 0; c < nchan[b]; c++) {
      int a = axes[3 * b + c];
      if (a < 0 || a > 2) { delete s; return srb_set_error(-1, "fk_create: axis must be 0,1,2"); }
      ax |= a << (2 * c);
    }
    k.axes[b] = ax;
    nang += nchan[b];
    for (int i = 0; i < 3; i++) { k.offset[b][i] = offset[3 * b + i]; k.com[b][i] = local_com[3 * b + i]; }
    k.mass[b] = mass[b];
    total += mass[b];
    const double* I = inertia + 9 * b;
    k.inertia[b][0] = I[0]; k.inertia[b][1] = I[4]; k.inertia[b][2] = I[8]; k.inertia[b][3] = I[1]; k.inertia[b][4] = I[2]; k.inertia[b][5] = I[5];
  }
  k.n_pose = 7 + nang; k.n_dq = 6 + nang; k.maxdepth = maxd;
  k.inv_total_mass = total > 0 ? 1.0 / total : 0.0;
  FCK(cudaSetDevice(device));
  FCK(cudaMalloc(&s->dev, sizeof(fk::SkelConst)));
  FCK(cudaMemcpy(s->dev, &k, sizeof(k), cudaMemcpyHostToDevice));
  *out = s;
  return 0;
}

extern "C" int fk_destroy(fk_skeleton* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaFree(s->dev);
  delete s;
  return 0;
}

extern "C" int fk_dims(fk_skeleton* s, int32_t* nb, int32_t* n_pose, int32_t* n_dq) {
  if (!s) return srb_set_error(-1, "fk_dims: null handle");
  if (nb) *nb = s->host.nb; if (n_pose) *n_pose = s->host.n_pose; if (n_dq) *n_dq = s->host.n_dq;
  return 0;
}

template <typename real, int BLOCK>
static int fk_launch(fk_skeleton* s, int64_t n, const real* pose, const real* dq, real* xform, real* com, real* mom, void* stream) {
  if (!s) return srb_set_error(-1, "fk_forward: null handle");
  if (n <= 0) return 0;
  if (!pose) return srb_set_error(-1, "fk_forward: null pose buffer");
  if (n > 2000000000LL) return srb_set_error(-1, "fk_forward: n too large for one call");
  FCK(cudaSetDevice(s->device));
  size_t smem = (size_t)(s->host.maxdepth + 1) * 13 * BLOCK * sizeof(real);
  auto kern = fk::fk_forward_kernel<real, BLOCK>;
  FCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = (int)((n + BLOCK - 1) / BLOCK);
  kern<<<blocks, BLOCK, smem, (cudaStream_t)stream>>>(s->dev, (int)n, pose, dq, xform, com, mom);
  FCK(cudaGetLastError());
  return 0;
}

extern "C" int fk_forward_f32(fk_skeleton* s, int64_t n, const float* pose, const float* dq, float* xform, float* com, float* mom, void* stream) {
  return fk_launch<float, 64>(s, n, pose, dq, xform, com, mom, stream);
}
extern "C" int fk_forward_f64(fk_skeleton* s, int64_t n, const double* pose, const double* dq, double* xform, double* com, double* mom, void* stream) {
  return fk_launch<double, 64>(s, n, pose, dq, xform, com, mom, stream);
}
