// oracle/_ref/liblbfgs_ref.so = the REFERENCE's libLBFGS (work/taesooLib/BaseLib/motion/IK_lbfgs/lbfgs.cpp, compiled from where it
// lies) + this C-ABI shim, so that a test can drive the reference optimiser with a Python objective through ctypes callbacks.
// Parameters are set the way the reference's wrapper sets them (BaseLib/math/optimize.h:426-470, class LBFGS_METHOD):
// lbfgs_parameter_init, linesearch = LBFGS_LINESEARCH_BACKTRACKING, epsilon = constructor argument.  Test infrastructure only.
#include "lbfgs.h"
using namespace LBFGS;
extern "C" {
typedef double (*ref_eval_t)(void* instance, const double* x, double* g, int n, double step);
typedef int (*ref_progress_t)(void* instance, const double* x, const double* g, double fx, double xnorm, double gnorm, double step, int n, int k, int ls);
// returns the libLBFGS status code; x is updated in place; fx_out receives the final objective.  m <= 0 / max_iterations < 0 keep
// the library defaults (m = 6, max_iterations = 0 = until convergence).
int ref_lbfgs(int n, double* x, double* fx_out, ref_eval_t ev, ref_progress_t pr, void* instance, double epsilon, int m, int max_iterations) {
  lbfgs_parameter_t param;
  lbfgs_parameter_init(&param);
  param.linesearch = LBFGS_LINESEARCH_BACKTRACKING;
  param.epsilon = epsilon;
  if (m > 0) param.m = m;
  if (max_iterations >= 0) param.max_iterations = max_iterations;
  lbfgsfloatval_t* xx = lbfgs_malloc(n);
  for (int i = 0; i < n; i++) xx[i] = x[i];
  lbfgsfloatval_t fx = 0;
  int ret = lbfgs(n, xx, &fx, (lbfgs_evaluate_t)ev, (lbfgs_progress_t)pr, instance, &param);
  for (int i = 0; i < n; i++) x[i] = xx[i];
  if (fx_out) *fx_out = fx;
  lbfgs_free(xx);
  return ret;
}
// the defaults lbfgs_parameter_init installs (for the record in the fixture)
void ref_lbfgs_defaults(double* out) {
  lbfgs_parameter_t p;
  lbfgs_parameter_init(&p);
  out[0] = p.m; out[1] = p.epsilon; out[2] = p.past; out[3] = p.delta; out[4] = p.max_iterations; out[5] = p.linesearch;
  out[6] = p.max_linesearch; out[7] = p.min_step; out[8] = p.max_step; out[9] = p.ftol; out[10] = p.wolfe; out[11] = p.gtol; out[12] = p.xtol;
  out[13] = sizeof(lbfgsfloatval_t);
}
}
