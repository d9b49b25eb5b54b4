// reject_solve.cuh — device-resident loop state and kernel argument blocks shared by the
// match and reject/solve kernels.
#pragma once
#include "common.cuh"

namespace sicp {

// Lives in device memory for the whole run: the iteration loop never needs the host to carry
// a value from one kernel to the next, so iterations can be queued back to back (or replayed
// from a CUDA graph) and the host only reads the small per-iteration record.
struct DevState {
  Rigid T;      // current cumulative transform H(x)
  Rigid Tinv;
  double x[6];
  double x_new[6];
  Rigid T_new;
  double An[36];   // unweighted J^T M J at the last solution (for the sigmas)
  double sigma[6];
  double w;        // distance weight in use (frozen after iteration 0, simpleicp.py:229-234)
  double prev_mean, prev_std;
  long long n_kept;
  int skip;        // fewer than 6 correspondences or solve not requested
  int stop;        // stop rule met: every later queued kernel returns immediately
  int converged;
  int iterations_done;
  int lm_ok;
  int pad;
};

struct RSArgs {
  long long K;
  const double* dist;
  const float4* q_nrm;
  const double* q_xyz;
  const long long* nn_idx;
  const double* mov_xyz;
  uint8_t* keep;
  double* resid;
  const unsigned int* unresolved;
  DevState* state;
  sicp_iter_record* rec;
  double min_planarity;  // compared against the float32 planarity promoted to float64
  double min_change;
  double w_param;
  double obs[6];
  double wobs[6];
  double cm[3];
  int it;
  int do_solve;
  int arm_stop;
};

struct RSWork {
  unsigned int* hist;
  unsigned int* hist_other;
  unsigned long long* cand;
  unsigned int* counters;
  unsigned int* counters_other;
  unsigned long long* minkey;
  unsigned long long* minkey_other;
  double* partials;
  unsigned long long* phase_t;  // 32 x %globaltimer stamps of block 0 (diagnostics)
  unsigned int* barrier;        // {arrival count, generation} of the grid barrier
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

}  // namespace sicp
