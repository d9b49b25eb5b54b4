// reject_solve.cu — one kernel for everything between the match and the next match:
//
//   CorrPts.reject_wrt_planarity                    python/simpleicp/corrpts.py:139-163
//   CorrPts.reject_wrt_point_to_plane_distances     python/simpleicp/corrpts.py:165-188
//   SimpleICPOptimization.estimate_parameters       python/simpleicp/optimization.py:65-124
//   SimpleICPOptimization.estimate_parameter_uncertainties          optimization.py:126-170
//   SimpleICP.__check_convergence_criteria          python/simpleicp/simpleicp.py:355-379
//
// Phases (grid-wide barriers between them; a single block for K <= 4096, a cooperative grid of
// one block per SM otherwise):
//   A  median of d over the planarity survivors   : MSB-first radix select on order-preserving
//      64-bit keys, 11-bit digits, early exit to an in-block bitonic sort once <= 2048 candidates
//      remain.  NumPy semantics: mean of the two middle elements for even n.
//   B  MAD = median(|d - median|) the same way (raw MAD, no 1.4826: corrpts.py:186).
//   C  keep = |d - median| <= 3 MAD; accumulate, over the kept correspondences, the 13 x 13
//      moment matrix M = sum phi phi^T of phi = [n (x) (p', 1), -n.q'] (73 distinct sums).  The
//      residual of EVERY rigid transform is phi . theta(x), so one pass over the data suffices
//      and the non-linear least squares runs on M alone.
//   D  block 0 / warp 0: Levenberg-Marquardt on the exact Euler model to convergence, with the
//      reference's fixed (weight inf) / observed (0 < weight < inf) / free parameter handling.
//   E  residuals at the solution (direct float64 evaluation, reference operation order), their
//      mean / population std, parameter sigmas, stop rule, per-iteration record.
#include <cooperative_groups.h>

#include <algorithm>

#include "ctx.cuh"
#include "reject_solve.cuh"

namespace cg = cooperative_groups;

namespace sicp {

namespace {

constexpr int RS_THREADS = 384;  // 12 warps = 3 accumulation roles x 4 warps
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 2048;
constexpr int RS_LEVELS = 6;
constexpr int RS_CAP = 2048;
constexpr int RS_NACC = 26;           // accumulators per role
constexpr int RS_NPART = 3 * RS_NACC; // partial sums per block in phase C

__device__ __constant__ int kShift[RS_LEVELS] = {53, 42, 31, 20, 9, 0};
__device__ __constant__ int kWidth[RS_LEVELS] = {11, 11, 11, 11, 11, 9};

struct Shared {
  unsigned int hist[RS_BINS];
  unsigned long long sortbuf[RS_CAP];
  double red[RS_WARPS][RS_NACC];
  unsigned int scan_tmp[RS_WARPS];
  // select state (identical in every block)
  unsigned long long prefix;
  unsigned int k;
  unsigned int cnt;
  unsigned int below;
  unsigned int total;
  double bc[8];  // broadcast scalars
  // LM workspace
  double M[13][13];
  double J[13][6];
  double th[13];
  double B[13][7];
  double A[36];
  double g[6];
  double F;
  double tot[RS_NPART];
};

template <bool MULTI>
__device__ __forceinline__ void gsync() {
  if (MULTI)
    cg::this_grid().sync();
  else
    __syncthreads();
}

// Find the bin holding rank k in hist (n_bins <= RS_BINS); writes s.k (rank inside the bin),
// s.cnt (bin count), returns the bin through s.below (count below) and the return value.
__device__ unsigned int find_bin(Shared& s, const unsigned int* __restrict__ ghist, int n_bins,
                                 unsigned int k, unsigned int* total_out) {
  // each thread owns a contiguous chunk of 6 bins (384 * 6 = 2304 >= 2048)
  constexpr int PER = (RS_BINS + RS_THREADS - 1) / RS_THREADS;
  const int b0 = threadIdx.x * PER;
  unsigned int v[PER];
  unsigned int sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = b0 + j;
    v[j] = (b < n_bins) ? ghist[b] : 0u;
    sum += v[j];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s.scan_tmp[w] = incl;
  __syncthreads();
  unsigned int woff = 0, total = 0;
  for (int i = 0; i < RS_WARPS; ++i) {
    if (i < w) woff += s.scan_tmp[i];
    total += s.scan_tmp[i];
  }
  unsigned int run = woff + incl - sum;  // exclusive prefix of this thread's chunk
  if (k >= run && k < run + sum) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (k < run + v[j]) {
        s.below = run;
        s.cnt = v[j];
        s.k = k - run;
        s.bc[7] = (double)(b0 + j);
        break;
      }
      run += v[j];
    }
  }
  __syncthreads();
  if (total_out) *total_out = total;
  return (unsigned int)s.bc[7];
}

__device__ void bitonic_sort(unsigned long long* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n_pow2; t += RS_THREADS) {
        const int p = t ^ j;
        if (p > t) {
          const unsigned long long x = a[t], y = a[p];
          const bool up = ((t & k) == 0);
          if ((x > y) == up) {
            a[t] = y;
            a[p] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

// Median of {keyfn(i) : i in S1} with NumPy semantics.  SEL = 0: keys of d; SEL = 1: keys of
// |d - center|.  Result broadcast through s.bc[0] (lower middle) and s.bc[1] (upper middle).
template <bool MULTI, int SEL>
__device__ void radix_median(Shared& s, const RSArgs& a, RSWork wk, double center,
                             unsigned int& n1_out) {
  const long long K = a.K;
  const double minpl = a.min_planarity;
  unsigned int* ghist_base = wk.hist + (size_t)SEL * RS_LEVELS * RS_BINS;
  unsigned long long prefix = 0;
  unsigned int k = 0, cnt = 0, n1 = 0;
  bool even = false;
  int level = 0;
  for (;; ++level) {
    const int shift = kShift[level], width = kWidth[level];
    const int n_bins = 1 << width;
    for (int b = threadIdx.x; b < RS_BINS; b += RS_THREADS) s.hist[b] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)RS_THREADS + threadIdx.x; i < K;
         i += (long long)gridDim.x * RS_THREADS) {
      if ((double)a.q_nrm[i].w >= minpl) {
        const double v = (SEL == 0) ? a.dist[i] : fabs(a.dist[i] - center);
        const unsigned long long key = f64_to_key(v);
        if (level == 0 || (key >> kShift[level - 1]) == prefix)
          atomicAdd(&s.hist[(unsigned int)(key >> shift) & (n_bins - 1)], 1u);
      }
    }
    __syncthreads();
    unsigned int* gh = ghist_base + (size_t)level * RS_BINS;
    if (MULTI) {
      for (int b = threadIdx.x; b < n_bins; b += RS_THREADS)
        if (s.hist[b]) atomicAdd(&gh[b], s.hist[b]);
      gsync<MULTI>();
    }
    const unsigned int* src = MULTI ? gh : s.hist;
    unsigned int total = 0;
    if (level == 0) {
      // total = |S1| decides the ranks
      find_bin(s, src, n_bins, 0u, &total);
      n1 = total;
      if (n1 == 0) {
        n1_out = 0;
        return;
      }
      k = (n1 - 1) >> 1;
      even = ((n1 & 1u) == 0u);
    }
    const unsigned int bin = find_bin(s, src, n_bins, k, nullptr);
    k = s.k;
    cnt = s.cnt;
    prefix = (prefix << width) | bin;
    __syncthreads();
    if (cnt <= RS_CAP || level == RS_LEVELS - 1) break;
  }
  // ---- gather the candidates of the selected bin; track the smallest key above it
  const int shift = kShift[level];
  unsigned long long* cand = wk.cand + (size_t)SEL * RS_CAP;
  unsigned int* ccount = wk.counters + SEL;
  unsigned long long* gmin = wk.minkey + SEL;
  const bool gather = (cnt <= RS_CAP);
  unsigned long long mymin = ~0ull;
  if (!MULTI) {
    if (threadIdx.x == 0) s.total = 0;
    __syncthreads();
  }
  for (long long i = blockIdx.x * (long long)RS_THREADS + threadIdx.x; i < K;
       i += (long long)gridDim.x * RS_THREADS) {
    if ((double)a.q_nrm[i].w >= minpl) {
      const double v = (SEL == 0) ? a.dist[i] : fabs(a.dist[i] - center);
      const unsigned long long key = f64_to_key(v);
      const unsigned long long top = key >> shift;
      if (top == prefix) {
        if (gather) {
          if (MULTI)
            cand[atomicAdd(ccount, 1u)] = key;
          else
            s.sortbuf[atomicAdd(&s.total, 1u)] = key;
        }
      } else if (top > prefix) {
        mymin = min(mymin, key);
      }
    }
  }
  // block-min of mymin
  for (int o = 16; o > 0; o >>= 1) mymin = min(mymin, __shfl_xor_sync(0xffffffffu, mymin, o));
  __shared__ unsigned long long wmin[RS_WARPS];
  if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = mymin;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long m = wmin[0];
    for (int i = 1; i < RS_WARPS; ++i) m = min(m, wmin[i]);
    if (MULTI) {
      if (m != ~0ull) atomicMin(gmin, m);
    } else {
      wmin[0] = m;
    }
  }
  gsync<MULTI>();
  const unsigned long long above = MULTI ? *gmin : wmin[0];
  unsigned long long klo, khi;
  if (gather) {
    int np2 = 32;
    while (np2 < (int)cnt) np2 <<= 1;
    if (MULTI)
      for (int t = threadIdx.x; t < np2; t += RS_THREADS) s.sortbuf[t] = (t < (int)cnt) ? cand[t] : ~0ull;
    else
      for (int t = threadIdx.x; t < np2; t += RS_THREADS)
        if (t >= (int)cnt) s.sortbuf[t] = ~0ull;
    __syncthreads();
    bitonic_sort(s.sortbuf, np2);
    klo = s.sortbuf[k];
    khi = (k + 1 < cnt) ? s.sortbuf[k + 1] : above;
  } else {
    // only reachable when all 64 key bits are fixed: every candidate equals the prefix
    klo = prefix;
    khi = (k + 1 < cnt) ? prefix : above;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s.bc[0] = key_to_f64(klo);
    s.bc[1] = even ? key_to_f64(khi) : key_to_f64(klo);
  }
  __syncthreads();
  n1_out = n1;
}

// --- Levenberg-Marquardt on the 13 x 13 moment matrix (warp 0 of block 0) -------------------
// theta(x) = [R(alpha) row-wise with t'_a after each row, 1], t' = R c_m + t - c_f.
__device__ void lm_eval(Shared& s, const double* x, const double* cm, const double* cf, int lane) {
  if (lane == 0) {
    double s1, c1, s2, c2, s3, c3;
    sincos(x[0], &s1, &c1);
    sincos(x[1], &s2, &c2);
    sincos(x[2], &s3, &c3);
    double R[9], D[3][9];
    R[0] = c2 * c3; R[1] = -c2 * s3; R[2] = s2;
    R[3] = c1 * s3 + s1 * s2 * c3; R[4] = c1 * c3 - s1 * s2 * s3; R[5] = -s1 * c2;
    R[6] = s1 * s3 - c1 * s2 * c3; R[7] = s1 * c3 + c1 * s2 * s3; R[8] = c1 * c2;
    // d/d alpha1
    D[0][0] = 0; D[0][1] = 0; D[0][2] = 0;
    D[0][3] = -R[6]; D[0][4] = -R[7]; D[0][5] = -R[8];
    D[0][6] = R[3]; D[0][7] = R[4]; D[0][8] = R[5];
    // d/d alpha2
    D[1][0] = -s2 * c3; D[1][1] = s2 * s3; D[1][2] = c2;
    D[1][3] = s1 * c2 * c3; D[1][4] = -s1 * c2 * s3; D[1][5] = s1 * s2;
    D[1][6] = -c1 * c2 * c3; D[1][7] = c1 * c2 * s3; D[1][8] = -c1 * s2;
    // d/d alpha3
    D[2][0] = -c2 * s3; D[2][1] = -c2 * c3; D[2][2] = 0;
    D[2][3] = c1 * c3 - s1 * s2 * s3; D[2][4] = -c1 * s3 - s1 * s2 * c3; D[2][5] = 0;
    D[2][6] = s1 * c3 + c1 * s2 * s3; D[2][7] = -s1 * s3 + c1 * s2 * c3; D[2][8] = 0;
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) {
        s.th[a * 4 + b] = R[a * 3 + b];
        for (int k = 0; k < 3; ++k) s.J[a * 4 + b][k] = D[k][a * 3 + b];
        for (int k = 3; k < 6; ++k) s.J[a * 4 + b][k] = 0.0;
      }
      s.th[a * 4 + 3] = R[a * 3 + 0] * cm[0] + R[a * 3 + 1] * cm[1] + R[a * 3 + 2] * cm[2] + x[3 + a] - cf[a];
      for (int k = 0; k < 3; ++k)
        s.J[a * 4 + 3][k] = D[k][a * 3 + 0] * cm[0] + D[k][a * 3 + 1] * cm[1] + D[k][a * 3 + 2] * cm[2];
      for (int k = 3; k < 6; ++k) s.J[a * 4 + 3][k] = (k - 3 == a) ? 1.0 : 0.0;
    }
    s.th[12] = 1.0;
    for (int k = 0; k < 6; ++k) s.J[12][k] = 0.0;
  }
  __syncwarp();
  for (int e = lane; e < 91; e += 32) {
    const int r = e / 7, c = e % 7;
    double acc = 0.0;
    for (int m = 0; m < 13; ++m) acc = fma(s.M[r][m], (c < 6) ? s.J[m][c] : s.th[m], acc);
    s.B[r][c] = acc;
  }
  __syncwarp();
  for (int e = lane; e < 43; e += 32) {
    double acc = 0.0;
    if (e < 36) {
      const int i = e / 6, j = e % 6;
      for (int m = 0; m < 13; ++m) acc = fma(s.J[m][i], s.B[m][j], acc);
      s.A[e] = acc;
    } else if (e < 42) {
      const int i = e - 36;
      for (int m = 0; m < 13; ++m) acc = fma(s.J[m][i], s.B[m][6], acc);
      s.g[i] = acc;
    } else {
      for (int m = 0; m < 13; ++m) acc = fma(s.th[m], s.B[m][6], acc);
      s.F = acc;
    }
  }
  __syncwarp();
}

// In-place Cholesky solve of the n x n SPD system A x = b (n <= 6), returns false if not SPD.
__device__ bool chol_solve(double* A, double* b, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= A[j * 6 + k] * A[j * 6 + k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    A[j * 6 + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = A[i * 6 + j];
      for (int k = 0; k < j; ++k) v -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = v / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= A[i * 6 + k] * b[k];
    b[i] = v / A[i * 6 + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i];
    for (int k = i + 1; k < n; ++k) v -= A[k * 6 + i] * b[k];
    b[i] = v / A[i * 6 + i];
  }
  return true;
}

struct LmOut {
  double x[6];
  double An[36];  // unweighted J^T M J at the solution
  int iters;
  int ok;
};

__device__ void lm_solve(Shared& s, const RSArgs& a, double w, const double* x0, const double* cm,
                         const double* cf, LmOut& out, int lane) {
  double x[6], xn[6];
  for (int j = 0; j < 6; ++j) x[j] = x0[j];
  int fidx[6], nf = 0;
  for (int j = 0; j < 6; ++j)
    if (isfinite(a.wobs[j])) fidx[nf++] = j;
  const double w2 = w * w;
  auto obs_cost = [&](const double* xx) {
    double c = 0.0;
    for (int j = 0; j < 6; ++j)
      if (a.wobs[j] > 0.0 && isfinite(a.wobs[j])) {
        const double r = a.wobs[j] * (xx[j] - a.obs[j]);
        c += r * r;
      }
    return c;
  };
  lm_eval(s, x, cm, cf, lane);
  double F = w2 * s.F + obs_cost(x);
  double lambda = 0.0;
  int it = 0, ok = 1;
  double Ak[36], gk[6];
  for (int e = 0; e < 36; ++e) Ak[e] = s.A[e];
  for (int e = 0; e < 6; ++e) gk[e] = s.g[e];
  for (it = 0; it < 40 && nf > 0; ++it) {
    // reduced, weighted Gauss-Newton system on the free parameters
    double Ar[36], br[6];
    for (int i = 0; i < nf; ++i) {
      const int pi = fidx[i];
      double gi = w2 * gk[pi];
      if (a.wobs[pi] > 0.0) gi += a.wobs[pi] * a.wobs[pi] * (x[pi] - a.obs[pi]);
      br[i] = -gi;
      for (int j = 0; j < nf; ++j) Ar[i * 6 + j] = w2 * Ak[pi * 6 + fidx[j]];
      if (a.wobs[pi] > 0.0) Ar[i * 6 + i] += a.wobs[pi] * a.wobs[pi];
      Ar[i * 6 + i] *= (1.0 + lambda);
    }
    const bool spd = chol_solve(Ar, br, nf);
    if (!spd) {
      lambda = fmax(lambda * 10.0, 1e-6);
      if (lambda > 1e10) {
        ok = 0;
        break;
      }
      continue;
    }
    double rel = 0.0;
    for (int j = 0; j < 6; ++j) xn[j] = x[j];
    for (int i = 0; i < nf; ++i) {
      xn[fidx[i]] = x[fidx[i]] + br[i];
      rel = fmax(rel, fabs(br[i]) / fmax(fabs(x[fidx[i]]), 1e-3));
    }
    lm_eval(s, xn, cm, cf, lane);
    const double Fn = w2 * s.F + obs_cost(xn);
    if (Fn <= F * (1.0 + 1e-10) + 1e-300 || rel < 1e-12) {
      const bool stalled = (it >= 1 && Fn >= F * (1.0 - 1e-14));
      for (int j = 0; j < 6; ++j) x[j] = xn[j];
      for (int e = 0; e < 36; ++e) Ak[e] = s.A[e];
      for (int e = 0; e < 6; ++e) gk[e] = s.g[e];
      F = Fn;
      lambda = (lambda > 1e-9) ? lambda * 0.1 : 0.0;
      if (rel < 1e-11 || stalled) {
        ++it;
        break;
      }
    } else {
      lambda = fmax(lambda * 10.0, 1e-4);
      if (lambda > 1e10) break;
    }
  }
  for (int j = 0; j < 6; ++j) out.x[j] = x[j];
  for (int e = 0; e < 36; ++e) out.An[e] = Ak[e];
  out.iters = it;
  out.ok = ok;
}

// sigma of the free parameters: Cxx = s0^2 (A^T P A)^-1 in the reference's formulation
// (optimization.py:147-160): N = w * sum a a^T + diag(w_obs), vPv = w sum r^2 + sum w_obs dx^2.
__device__ void uncertainties(const RSArgs& a, const double* An, double w, const double* x,
                              double sum_r2, long long n_kept, double* sigma) {
  int fidx[6], nf = 0, nobs = 0;
  for (int j = 0; j < 6; ++j) {
    sigma[j] = nan("");
    if (isfinite(a.wobs[j])) fidx[nf++] = j;
    if (a.wobs[j] > 0.0 && isfinite(a.wobs[j])) ++nobs;
  }
  if (nf == 0) return;
  double vPv = w * sum_r2;
  for (int j = 0; j < 6; ++j)
    if (a.wobs[j] > 0.0 && isfinite(a.wobs[j])) vPv += a.wobs[j] * (x[j] - a.obs[j]) * (x[j] - a.obs[j]);
  const double dof = (double)(n_kept + nobs - nf);
  const double s02 = vPv / dof;
  for (int c = 0; c < nf; ++c) {
    double N[36], e[6];
    for (int i = 0; i < nf; ++i) {
      for (int j = 0; j < nf; ++j) N[i * 6 + j] = w * An[fidx[i] * 6 + fidx[j]];
      if (a.wobs[fidx[i]] > 0.0) N[i * 6 + i] += a.wobs[fidx[i]];
      e[i] = (i == c) ? 1.0 : 0.0;
    }
    if (chol_solve(N, e, nf)) sigma[fidx[c]] = sqrt(s02 * e[c]);
  }
}

template <bool MULTI>
__global__ void __launch_bounds__(RS_THREADS, 1) k_reject_solve(RSArgs a, RSWork wk) {
  __shared__ Shared s;
  DevState* st = a.state;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long K = a.K;
  const int G = gridDim.x;

  // zero the other parity's workspace for the next launch (nobody is using it)
  {
    unsigned int* oh = wk.hist_other;
    const long long nz = 2ll * RS_LEVELS * RS_BINS;
    for (long long i = blockIdx.x * (long long)RS_THREADS + tid; i < nz; i += (long long)G * RS_THREADS) oh[i] = 0;
    if (blockIdx.x == 0 && tid < 2) {
      wk.counters_other[tid] = 0;
      wk.minkey_other[tid] = ~0ull;
    }
  }
  if (st->stop) return;  // a previous iteration already met the stop rule

  // ---- A: median
  unsigned int n1 = 0;
  radix_median<MULTI, 0>(s, a, wk, 0.0, n1);
  sicp_iter_record* rec = a.rec;
  if (n1 == 0) {
    if (blockIdx.x == 0 && tid == 0) {
      rec->n_kept = 0;
      rec->median = rec->mad = nan("");
      st->n_kept = 0;
    }
    return;
  }
  const double median = 0.5 * (s.bc[0] + s.bc[1]);
  __syncthreads();
  // ---- B: MAD
  unsigned int n1b = 0;
  radix_median<MULTI, 1>(s, a, wk, median, n1b);
  const double mad = 0.5 * (s.bc[0] + s.bc[1]);
  const double lim = 3.0 * mad;
  __syncthreads();

  // ---- C: keep flags + moment accumulation
  const Rigid Tin = st->T;
  double cm[3] = {a.cm[0], a.cm[1], a.cm[2]}, cf[3];
  rigid_apply(Tin, cm[0], cm[1], cm[2], cf[0], cf[1], cf[2]);
  {
    const int role = warp % 3, sub = (warp / 3) * 32 + lane;  // 128 threads per role
    double acc[RS_NACC];
#pragma unroll
    for (int j = 0; j < RS_NACC; ++j) acc[j] = 0.0;
    const long long chunk = (K + G - 1) / G;
    const long long i0 = blockIdx.x * chunk, i1 = min(i0 + chunk, K);
    for (long long i = i0 + sub; i < i1; i += 128) {
      const float4 nr = a.q_nrm[i];
      const double d = a.dist[i];
      const bool kp = ((double)nr.w >= a.min_planarity) && (fabs(d - median) <= lim);
      if (role == 0) a.keep[i] = kp ? 1 : 0;
      if (!kp) continue;
      const long long j = a.nn_idx[i];
      const double u0 = a.mov_xyz[3 * j + 0] - cm[0], u1 = a.mov_xyz[3 * j + 1] - cm[1],
                   u2 = a.mov_xyz[3 * j + 2] - cm[2];
      const double q0 = a.q_xyz[3 * i + 0] - cf[0], q1 = a.q_xyz[3 * i + 1] - cf[1],
                   q2 = a.q_xyz[3 * i + 2] - cf[2];
      const double n0 = (double)nr.x, n1d = (double)nr.y, n2 = (double)nr.z;
      const double sc = -(n0 * q0 + n1d * q1 + n2 * q2);
      double na, nb, nv;
      if (role == 0) {
        na = n0 * n0; nb = n0 * n1d; nv = n0;
      } else if (role == 1) {
        na = n0 * n2; nb = n1d * n1d; nv = n1d;
      } else {
        na = n1d * n2; nb = n2 * n2; nv = n2;
      }
      const double U[10] = {u0 * u0, u0 * u1, u0 * u2, u0, u1 * u1, u1 * u2, u1, u2 * u2, u2, 1.0};
#pragma unroll
      for (int t = 0; t < 10; ++t) {
        acc[t] = fma(na, U[t], acc[t]);
        acc[10 + t] = fma(nb, U[t], acc[10 + t]);
      }
      const double sn = sc * nv;
      acc[20] = fma(sn, u0, acc[20]);
      acc[21] = fma(sn, u1, acc[21]);
      acc[22] = fma(sn, u2, acc[22]);
      acc[23] += sn;
      if (role == 0) {
        acc[24] = fma(sc, sc, acc[24]);
        acc[25] += 1.0;
      } else if (role == 1) {
        acc[24] += d;
        acc[25] = fma(d, d, acc[25]);
      }
    }
#pragma unroll
    for (int j = 0; j < RS_NACC; ++j) {
      const double v = warp_sum(acc[j]);
      if (lane == 0) s.red[warp][j] = v;
    }
    __syncthreads();
    if (tid < RS_NPART) {
      const int r = tid / RS_NACC, j = tid % RS_NACC;
      double v = 0.0;
      for (int ww = r; ww < RS_WARPS; ww += 3) v += s.red[ww][j];
      if (MULTI)
        wk.partials[(size_t)blockIdx.x * RS_NPART + tid] = v;
      else
        s.tot[tid] = v;
    }
  }
  gsync<MULTI>();

  // ---- D: block 0 reduces the partials and solves
  if (blockIdx.x == 0) {
    if (MULTI) {
      if (tid < RS_NPART) {
        double v = 0.0;
        for (int b = 0; b < G; ++b) v += wk.partials[(size_t)b * RS_NPART + tid];
        s.tot[tid] = v;
      }
    }
    __syncthreads();
    // assemble M (13 x 13) from T (6 x 10), V (3 x 4), S
    for (int e = tid; e < 169; e += RS_THREADS) {
      const int r = e / 13, c = e % 13;
      double v;
      if (r == 12 && c == 12) {
        v = s.tot[0 * RS_NACC + 24];
      } else if (r == 12 || c == 12) {
        const int o = (r == 12) ? c : r;
        v = s.tot[(o / 4) * RS_NACC + 20 + (o % 4)];
      } else {
        int aa = r / 4, bb = r % 4, cc = c / 4, dd = c % 4;
        if (aa > cc) { int t = aa; aa = cc; cc = t; }
        if (bb > dd) { int t = bb; bb = dd; dd = t; }
        const int i6 = (aa == 0) ? cc : (aa == 1 ? 2 + cc : 5);          // (0,0)0 (0,1)1 (0,2)2 (1,1)3 (1,2)4 (2,2)5
        const int i10 = (bb == 0) ? dd : (bb == 1 ? 3 + dd : (bb == 2 ? 5 + dd : 9));  // (1,1)4 (1,2)5 (1,3)6 (2,2)7 (2,3)8 (3,3)9
        v = s.tot[(i6 / 2) * RS_NACC + (i6 % 2) * 10 + i10];
      }
      s.M[r][c] = v;
    }
    __syncthreads();
    if (warp == 0) {
      const long long n_kept = (long long)(s.tot[0 * RS_NACC + 25] + 0.5);
      const double sum_d = s.tot[1 * RS_NACC + 24], sum_d2 = s.tot[1 * RS_NACC + 25];
      const double mean_d = (n_kept > 0) ? sum_d / (double)n_kept : nan("");
      const double var_d = (n_kept > 0) ? fmax(sum_d2 / (double)n_kept - mean_d * mean_d, 0.0) : nan("");
      double w = st->w;
      if (a.it == 0 || !(w > 0.0)) {
        w = a.w_param;
        if (!(w > 0.0)) w = 1.0 / var_d;  // distance_weights=None: 1/std(d)^2 (simpleicp.py:233-234)
      }
      int skip = (n_kept < 6) || !a.do_solve;
      LmOut lo;
      lo.iters = 0;
      lo.ok = 1;
      if (!skip) {
        lm_solve(s, a, w, st->x, cm, cf, lo, lane);
      }
      if (lane == 0) {
        rec->n_kept = n_kept;
        rec->median = median;
        rec->mad = mad;
        rec->mean_dist = mean_d;
        rec->std_dist = sqrt(var_d);
        rec->distance_weight = w;
        rec->lm_iterations = lo.iters;
        rec->n_bruteforce = a.unresolved ? (int)a.unresolved[K] : 0;
        st->n_kept = n_kept;
        st->skip = skip;
        st->w = w;
        if (n_kept < 6 && a.do_solve && a.arm_stop) st->stop = 2;  // too few correspondences
        if (!skip) {
          for (int j = 0; j < 6; ++j) st->x_new[j] = lo.x[j];
          for (int e = 0; e < 36; ++e) st->An[e] = lo.An[e];
          st->T_new = rigid_from_x(lo.x);
          st->lm_ok = lo.ok;
        }
      }
    }
  }
  gsync<MULTI>();
  if (st->skip) return;

  // ---- E: residuals at the solution, in the reference's operation order
  {
    const Rigid Tn = st->T_new;
    double sr = 0.0, sr2 = 0.0;
    for (long long i = blockIdx.x * (long long)RS_THREADS + tid; i < K; i += (long long)G * RS_THREADS) {
      if (!a.keep[i]) continue;
      const long long j = a.nn_idx[i];
      double tx, ty, tz;
      rigid_apply(Tn, a.mov_xyz[3 * j + 0], a.mov_xyz[3 * j + 1], a.mov_xyz[3 * j + 2], tx, ty, tz);
      const float4 nr = a.q_nrm[i];
      const double dx = tx - a.q_xyz[3 * i + 0], dy = ty - a.q_xyz[3 * i + 1], dz = tz - a.q_xyz[3 * i + 2];
      const double r = __dadd_rn(__dadd_rn(__dmul_rn(dx, (double)nr.x), __dmul_rn(dy, (double)nr.y)),
                                 __dmul_rn(dz, (double)nr.z));
      a.resid[i] = r;
      sr += r;
      sr2 = fma(r, r, sr2);
    }
    sr = warp_sum(sr);
    sr2 = warp_sum(sr2);
    if (lane == 0) {
      s.red[warp][0] = sr;
      s.red[warp][1] = sr2;
    }
    __syncthreads();
    if (tid == 0) {
      double t0 = 0, t1 = 0;
      for (int ww = 0; ww < RS_WARPS; ++ww) {
        t0 += s.red[ww][0];
        t1 += s.red[ww][1];
      }
      if (MULTI) {
        wk.partials[(size_t)G * RS_NPART + 2 * blockIdx.x + 0] = t0;
        wk.partials[(size_t)G * RS_NPART + 2 * blockIdx.x + 1] = t1;
      } else {
        s.bc[2] = t0;
        s.bc[3] = t1;
      }
    }
  }
  gsync<MULTI>();
  if (blockIdx.x == 0 && tid == 0) {
    double t0 = 0, t1 = 0;
    if (MULTI) {
      for (int b = 0; b < G; ++b) {
        t0 += wk.partials[(size_t)G * RS_NPART + 2 * b + 0];
        t1 += wk.partials[(size_t)G * RS_NPART + 2 * b + 1];
      }
    } else {
      t0 = s.bc[2];
      t1 = s.bc[3];
    }
    const long long n = st->n_kept;
    const double mean = t0 / (double)n;
    const double sd = sqrt(fmax(t1 / (double)n - mean * mean, 0.0));
    rec->mean_res = mean;
    rec->std_res = sd;
    double sigma[6];
    uncertainties(a, st->An, st->w, st->x_new, t1, n, sigma);
    for (int j = 0; j < 6; ++j) {
      rec->x[j] = st->x_new[j];
      st->sigma[j] = sigma[j];
      st->x[j] = st->x_new[j];
    }
    st->T = st->T_new;
    st->Tinv = rigid_inverse(st->T_new);
    // stop rule (simpleicp.py:355-379): relative change in percent of mean and population std
    int stop = 0;
    if (a.it > 0) {
      const double m0 = st->prev_mean, s0 = st->prev_std;
      const double cmn = (m0 == 0.0) ? ((mean == 0.0) ? 0.0 : kInf) : fabs((mean - m0) / m0 * 100.0);
      const double csd = (s0 == 0.0) ? ((sd == 0.0) ? 0.0 : kInf) : fabs((sd - s0) / s0 * 100.0);
      stop = (cmn < a.min_change && csd < a.min_change) ? 1 : 0;
    }
    st->prev_mean = mean;
    st->prev_std = sd;
    st->iterations_done = a.it + 1;
    if (stop && a.arm_stop) st->stop = 1;
    st->converged = stop;
  }
}

// Ordered compaction of the kept residuals (final iteration only): block-level scan, one block
// per 4096 elements, two kernels.
__global__ void __launch_bounds__(256)
    k_compact_count(const uint8_t* __restrict__ keep, long long K, unsigned int* __restrict__ bsum) {
  const long long base = (long long)blockIdx.x * 4096;
  unsigned int c = 0;
  for (int j = threadIdx.x; j < 4096; j += 256)
    if (base + j < K) c += keep[base + j] ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ unsigned int sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    bsum[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
    k_compact_write(const uint8_t* __restrict__ keep, const double* __restrict__ resid, long long K,
                    const unsigned int* __restrict__ bsum, double* __restrict__ out) {
  // offset of this block = sum of previous block counts (few hundred blocks at most)
  __shared__ unsigned int boff;
  __shared__ unsigned int wsum[8];
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (unsigned int b = 0; b < blockIdx.x; ++b) t += bsum[b];
    boff = t;
  }
  __syncthreads();
  const long long base = (long long)blockIdx.x * 4096 + threadIdx.x * 16;
  unsigned int flags = 0, c = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (base + j < K && keep[base + j]) {
      flags |= 1u << j;
      ++c;
    }
  unsigned int incl = c;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  unsigned int woff = 0;
  for (int i = 0; i < w; ++i) woff += wsum[i];
  unsigned int pos = boff + woff + incl - c;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (flags & (1u << j)) out[pos++] = resid[base + j];
}

}  // namespace

void reject_solve_launch(Ctx& c, const sicp_run_params& p, int it, bool do_solve, bool arm_stop,
                         int rec_slot) {
  const long long K = c.K;
  const bool multi = K > 4096;
  int G = 1;
  if (multi) G = (int)std::min<long long>(c.num_sms, (K + RS_THREADS - 1) / RS_THREADS);
  // workspace: two parities
  const size_t hist_n = 2ull * RS_LEVELS * RS_BINS;
  if (c.ws.hist.cap < 2 * hist_n) {
    c.ws.hist.reserve(2 * hist_n);
    SICP_CUDA(cudaMemsetAsync(c.ws.hist.p, 0, 2 * hist_n * sizeof(unsigned int), c.stream));
    c.ws.cand.reserve(2ull * 2 * RS_CAP);
    c.ws.counters.reserve(64);
    SICP_CUDA(cudaMemsetAsync(c.ws.counters.p, 0, 64 * sizeof(unsigned int), c.stream));
    c.ws.minkey.reserve(16);
    SICP_CUDA(cudaMemsetAsync(c.ws.minkey.p, 0xff, 16 * sizeof(unsigned long long), c.stream));
    c.rs_parity = 0;
  }
  c.ws.partials.reserve((size_t)c.num_sms * (RS_NPART + 2) + 16);
  c.keep.reserve(K);
  c.resid.reserve(K);
  const int par = c.rs_parity;
  c.rs_parity ^= 1;

  RSArgs a;
  a.K = K;
  a.dist = c.dist.p;
  a.q_nrm = c.q_nrm.p;
  a.q_xyz = c.q_xyz.p;
  a.nn_idx = c.nn_idx.p;
  a.mov_xyz = c.mov_xyz.p;
  a.keep = c.keep.p;
  a.resid = c.resid.p;
  a.unresolved = (c.nn_engine == SICP_NN_AUTO) ? c.unresolved.p : nullptr;
  a.state = c.dev_state.p;
  a.rec = c.ws.rec.p + rec_slot;
  a.min_planarity = p.min_planarity;
  a.min_change = p.min_change;
  a.w_param = p.lsq.distance_weight;
  for (int j = 0; j < 6; ++j) {
    a.obs[j] = p.lsq.observed[j];
    a.wobs[j] = p.lsq.obs_weight[j];
  }
  for (int j = 0; j < 3; ++j) a.cm[j] = c.mov_center[j];
  a.it = it;
  a.do_solve = do_solve ? 1 : 0;
  a.arm_stop = arm_stop ? 1 : 0;

  RSWork wk;
  wk.hist = c.ws.hist.p + par * hist_n;
  wk.hist_other = c.ws.hist.p + (par ^ 1) * hist_n;
  wk.cand = c.ws.cand.p + par * 2 * RS_CAP;
  wk.counters = c.ws.counters.p + par * 2;
  wk.counters_other = c.ws.counters.p + (par ^ 1) * 2;
  wk.minkey = c.ws.minkey.p + par * 2;
  wk.minkey_other = c.ws.minkey.p + (par ^ 1) * 2;
  wk.partials = c.ws.partials.p;

  if (multi) {
    void* args[] = {&a, &wk};
    SICP_CUDA(cudaLaunchCooperativeKernel((void*)k_reject_solve<true>, dim3(G), dim3(RS_THREADS),
                                          args, 0, c.stream));
  } else {
    k_reject_solve<false><<<1, RS_THREADS, 0, c.stream>>>(a, wk);
    SICP_CUDA(cudaGetLastError());
  }
}

void compact_residuals_launch(Ctx& c) {
  const long long K = c.K;
  const unsigned int nb = (unsigned int)((K + 4095) / 4096);
  c.ws.counters.reserve(64);
  c.compact_sums.reserve(nb + 1);
  c.resid_compact.reserve(std::max<long long>(K, 1));
  k_compact_count<<<nb, 256, 0, c.stream>>>(c.keep.p, K, c.compact_sums.p);
  k_compact_write<<<nb, 256, 0, c.stream>>>(c.keep.p, c.resid.p, K, c.compact_sums.p,
                                           c.resid_compact.p);
  SICP_CUDA(cudaGetLastError());
}

}  // namespace sicp
