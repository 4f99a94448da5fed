// C-ABI of the batched FK / CoM / momentum pass (include/bonefk_b200.h).
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cmath>
#include <string.h>
#include <stdint.h>
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
  bool hanyang = false;        // tree structure equals the compile-time table of fk::Hanyang -> specialised kernel
  bool force_generic = false;
  int tma_mode = -1;           // see fk_launch
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
  {
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = cudaMalloc(&s->dev, sizeof(fk::SkelConst));
    if (ce == cudaSuccess) ce = cudaMemcpy(s->dev, &k, sizeof(k), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(s->dev); delete s; return srb_set_error(-2, std::string("fk_create: ") + cudaGetErrorString(ce)); }
  }
  s->hanyang = fk::Hanyang::matches(k);
  *out = s;
  return 0;
}

extern "C" int fk_set_generic(fk_skeleton* s, int32_t force_generic) {
  if (!s) return srb_set_error(-1, "fk_set_generic: null handle");
  s->force_generic = force_generic != 0;
  return 0;
}
extern "C" int fk_set_tma(fk_skeleton* s, int32_t use_tma) {
  if (!s) return srb_set_error(-1, "fk_set_tma: null handle");
  s->tma_mode = use_tma;
  return 0;
}
extern "C" int fk_is_specialised(fk_skeleton* s) { return (s && s->hanyang && !s->force_generic) ? 1 : 0; }

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
  if (s->hanyang && !s->force_generic) {               // skeleton-specialised instantiations
    const bool al = ((uintptr_t)pose % 16 == 0) && (!xform || (uintptr_t)xform % 16 == 0) && (!dq || (uintptr_t)dq % 16 == 0) &&
                    (!com || (uintptr_t)com % 16 == 0) && (!mom || (uintptr_t)mom % 16 == 0);
    const bool momentum = dq && mom;
    if (al) {
      cudaStream_t st = (cudaStream_t)stream;
      // tma_mode: -1 auto (measured best per precision / outputs, tools/bench_fk.py), 0 direct rows, 1 rolled loop on TMA tiles,
      //           3 unrolled sweep on TMA tiles, 4 chain-wise unrolled bodies on TMA tiles.  (A per-bone CTA barrier to share instruction-cache lines was tried with
      //           128- and 256-pose tiles: 10-20 % slower than the plain unrolled tile kernel, removed.)
      int mode = s->tma_mode;
      // measured on a B200, 1 M poses (profiles/r02_fk_bench.jsonl): with momentum the chain-wise kernel wins in both precisions
      // (fp32 0.44 ms vs 0.60 rolled / 0.77 unrolled; fp64 0.94 ms vs 1.07 generic); FK + CoM alone: fp32 unrolled tiles, fp64 direct rows
      if (mode < 0) mode = momentum ? 4 : (sizeof(real) == 4 ? 3 : 0);
      if (mode != 0 && mode != 1 && mode != 4) mode = 3;
      const int TB = mode == 0 ? 1 : 64;
      const int64_t full = mode == 0 ? 0 : n / TB;
      if (full > 0) {
#define FK_LAUNCH_TILE(KERN, TBV, EXTRA)                                                                                     \
        do {                                                                                                                  \
          const size_t tile = (size_t)(TBV) * (140 + 3 + 6 + (EXTRA)) * sizeof(real);                                        \
          auto k = KERN;                                                                                                      \
          FCK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile));                               \
          k<<<(unsigned)full, (TBV), tile, st>>>(s->dev, pose, momentum ? dq : nullptr, xform, com, momentum ? mom : nullptr); \
        } while (0)
        if (mode == 1) {
          if (momentum) FK_LAUNCH_TILE((fk::fk_hanyang_tma_rolled_kernel<real, true, 64>), 64, 59);
          else FK_LAUNCH_TILE((fk::fk_hanyang_tma_rolled_kernel<real, false, 64>), 64, 0);
        } else if (mode == 4) {                          // chain-wise unrolled bodies on TMA tiles
          if (momentum) FK_LAUNCH_TILE((fk::fk_hanyang_tma_chain_kernel<real, true, 64>), 64, 59);
          else FK_LAUNCH_TILE((fk::fk_hanyang_tma_chain_kernel<real, false, 64>), 64, 0);
        } else {
          if (momentum) FK_LAUNCH_TILE((fk::fk_hanyang_tma_unrolled_kernel<real, true, false, 64>), 64, 0);
          else FK_LAUNCH_TILE((fk::fk_hanyang_tma_unrolled_kernel<real, false, false, 64>), 64, 0);
        }
#undef FK_LAUNCH_TILE
      }
      const int64_t done = full * TB, rest = n - done;
      if (rest > 0) {                                    // ragged tail (or everything when the TMA path is switched off)
        constexpr int HB = 128;
        const int blocks = (int)((rest + HB - 1) / HB);
        const real* p2 = pose + done * 60; const real* d2 = dq ? dq + done * 59 : nullptr;
        real* x2 = xform ? xform + done * 140 : nullptr; real* c2 = com ? com + done * 3 : nullptr; real* m2 = mom ? mom + done * 6 : nullptr;
        if (momentum) fk::fk_hanyang_kernel<real, true, HB><<<blocks, HB, 0, st>>>(s->dev, (int)rest, p2, d2, x2, c2, m2);
        else fk::fk_hanyang_kernel<real, false, HB><<<blocks, HB, 0, st>>>(s->dev, (int)rest, p2, nullptr, x2, c2, nullptr);
      }
      FCK(cudaGetLastError());
      return 0;
    }
  }
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
