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
  // Predictor for the next iteration's order statistics: the match kernel bins every
  // point-to-plane distance of a planarity survivor into a linear histogram centred on the
  // previous median, +-4 previous MADs wide (see lh_bin below).
  double pred_med, pred_mad, pred_minpl;
  int pred_valid;   // pred_* describe a finished reject phase of this run
  int hist_filled;  // the last match filled the linear histogram with the current pred_*
  // Linearised variant (SICP_VARIANT_LINEARIZED*): T_res is the affine map (I + [alpha]x) T + t
  // whose point-to-plane residuals are the reference's "A x - l" (c++/src/corrpts.cpp:155);
  // H_rep is the matrix the reference driver reports (dH * H, or H * dH for the C++ driver).
  // In the default variant T_res == T_new and H_rep == T.
  Rigid T_res;
  Rigid H_rep;
};

// Linear histogram shared by the match kernels (producers) and k_reject_solve (consumer).
// LH_BINS regular bins over [med - 4 mad, med + 4 mad), then one underflow and one overflow
// counter.  The bin function is monotone non-decreasing in d, which is all the exactness of the
// selection needs; both sides evaluate this very expression.
constexpr int LH_BINS = 4096;
__host__ __device__ inline int lh_bin(double d, double med, double mad) {
  const double lo = med - 4.0 * mad;
  const double inv_w = (double)LH_BINS / (8.0 * mad);
  const double t = (d - lo) * inv_w;
  if (!(t >= 0.0)) return LH_BINS;  // underflow (and NaN)
  if (t >= (double)LH_BINS) return LH_BINS + 1;
  return (int)t;
}

struct RSArgs {
  long long K;
  const double* dist;
  const float4* q_nrm;
  const double* q_xyz;
  const long long* nn_idx;
  const double* mov_xyz;
  uint8_t* keep;
  double* resid;
  unsigned int* unresolved;  // [K] = queries the grid left to the brute-force pass; reset by the kernel after reading
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
  int hist_expected;  // the preceding match launch fed lin_hist (if the predictor was valid)
  int variant;        // sicp_variant
  double stat_minpl;  // planarity bound of the set the median/MAD are taken over (-inf: everybody)
  // fused kernel only (k_rs_fused): what the match kernels left per correspondence
  const unsigned int* binstore;  // LH_BINS x bin_cap: correspondence numbers per predictor-histogram bin
  int bin_cap;
  const double* m_xyz;         // matched movable point (caller coordinates), K x 3
  int want_sigma;              // evaluate the parameter sigmas even if the stop rule does not fire
  // q_nrm is the match kernel's per-iteration copy whose .w is the signed effective planarity
  // (nn.cu: effective_planarity): |w| enters both planarity tests, a set sign bit = rejected by
  // the angle between the normals after the distance rejection
  int pl_signed;
};
__device__ __forceinline__ double pl_stat(float w, int pl_signed) { return (double)(pl_signed ? fabsf(w) : w); }
__device__ __forceinline__ bool pl_keep(float w, double min_planarity, int pl_signed) {
  if (pl_signed) return (double)fabsf(w) >= min_planarity && !signbit(w);
  return (double)w >= min_planarity;
}

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
  unsigned int* lin_hist;       // LH_BINS + 2 counters filled by the match kernels
  unsigned int* ticket;         // fused kernel: blocks finished so far (the last one solves)
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

}  // namespace sicp
