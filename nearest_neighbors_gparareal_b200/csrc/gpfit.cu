// Per-slice, per-dimension Gaussian-process fit and prediction.
//
// Replaces NNGP_p.get_preds (models.py:185-226): for every output dimension j the 9*R
// Nelder-Mead searches (models.py:228-260, SciPy _minimize_neldermead) of the negative log
// marginal likelihood (models.py:240-252) of a squared-exponential GP on the m nearest
// neighbours (models.py:86-92, 145-155), the selection rule (models.py:212-215) and the
// posterior mean (models.py:162-168) -- one launch per predict instead of d*9*R pickled
// tasks through pool.map.
//
// Mapping: grid = (d, nq); one CTA per (query, output dimension); one WARP per Nelder-Mead
// search (9 warps; restarts loop).  Inside a warp lane r owns row r of the m x m kernel matrix
// (m <= 32) in registers: right-looking Cholesky with shuffle broadcasts, forward solve fused
// into the factorisation, L transposed through a per-warp shared-memory tile for the
// column-oriented backward solve.  The m x m squared-distance matrix of the neighbours is
// computed once per query (gp_prep_kernel) and shared by all d*9*R searches.  The simplex
// arithmetic uses explicitly rounded (non-fused) operations in SciPy's order so the search
// follows the reference trajectory; a failed factorisation (pivot <= 0 or NaN) makes the
// objective +inf exactly like the reference (NaN -> inf, models.py:250-251).
#include "common.cuh"

#include <cmath>

static constexpr int GP_WARPS = NNGP_N_JITTER;  // one warp per jitter value
static constexpr unsigned FULL = 0xffffffffu;

// 10**jitter for jitter = -20..-12 (models.py:186, :88)
__constant__ double c_jit10[NNGP_N_JITTER] = {1e-20, 1e-19, 1e-18, 1e-17, 1e-16,
                                              1e-15, 1e-14, 1e-13, 1e-12};

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double dnan() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// np.sum over m values held one per lane, in NumPy's pairwise order (8 lanes, then the tail)
__device__ __forceinline__ double numpy_sum_lanes(double v, int m, int lane) {
  if (m < 8) {
    double r = 0.0;
    for (int i = 0; i < m; i++) r = __dadd_rn(r, shfl(v, i));
    return r;
  }
  const int nb = m >> 3;  // full blocks of 8
  double acc = v;         // lane j<8 : r[j]
  for (int b = 1; b < nb; b++) acc = __dadd_rn(acc, __shfl_down_sync(FULL, v, 8 * b));
  double t = __dadd_rn(acc, __shfl_xor_sync(FULL, acc, 1));
  t = __dadd_rn(t, __shfl_xor_sync(FULL, t, 2));
  t = __dadd_rn(t, __shfl_xor_sync(FULL, t, 4));
  double res = shfl(t, 0);
  for (int i = nb * 8; i < m; i++) res = __dadd_rn(res, shfl(v, i));
  return res;
}

// ---------------------------------------------------------------------------------------
// GP core for one warp.  On return (true): alpha = (K^-1 y)_lane, dg = L_lane,lane.
// r2s: shared, symmetric [M*M] squared distances (entry (a,b) at a*M+b); Ls: per-warp shared
// scratch [M*(M+1)].  Follows _fit_gp_jit (models.py:86-92).
// ---------------------------------------------------------------------------------------
struct GpSol {
  double alpha, dg, amp, c;
  bool ok;
};

// One copy per M in the module (noinline): the Nelder-Mead loop, the objective kernel and the
// final refit all call the same code, which keeps the instruction footprint inside the I-cache.
template <int M>
__device__ __noinline__ GpSol gp_factor_solve(double th0, double th1, double jit10,
                                              const double* __restrict__ r2s, double y, int m,
                                              int lane, double* __restrict__ Ls) {
  GpSol o;
  const double amp = exp10(th1);       // 10**sigma_y
  const double inv = 1.0 / exp10(th0);  // 1/(10**sigma_x)
  const double c = -0.5 * inv;
  o.amp = amp;
  o.c = c;
  o.alpha = 0.0;
  o.dg = 1.0;
  o.ok = false;
  const bool rowvalid = lane < m;
  const int col = (lane < M) ? lane : 0;
  double a[M];
  double dd = 1.0;
#pragma unroll
  for (int j = 0; j < M; j++) {
    const double kv = amp * exp(c * r2s[j * M + col]);
    const bool valid = rowvalid && (j < m);
    a[j] = valid ? kv : 0.0;
    if (j == lane) dd = valid ? (kv + jit10) : 1.0;
  }
  double z = rowvalid ? y : 0.0;
  double dg = 1.0;
  // right-looking Cholesky (LAPACK potf2 semantics: fail on pivot <= 0 or NaN), forward solve
  // L z = y fused (x_k = x_k / L_kk ; x_r -= x_k L_rk)
#pragma unroll
  for (int k = 0; k < M; k++) {
    if (k >= m) break;
    const double p = shfl(dd, k);
    if (!(p > 0.0)) return o;
    const double s = sqrt(p);
    const double is = 1.0 / s;
    const double zk = shfl(z, k) / s;
    const double lk = a[k] * is;
    if (lane == k) {
      dg = s;
      z = zk;
    } else if (lane > k) {
      a[k] = lk;
      z = z - zk * lk;
      dd = dd - lk * lk;
    }
#pragma unroll
    for (int j = k + 1; j < M; j++) {
      if (j >= m) break;
      const double ljk = shfl(lk, j);
      a[j] = a[j] - lk * ljk;
    }
  }
  // transpose L through shared memory: lane r writes row r, lane k reads column k
  constexpr int LD = M + 1;
  __syncwarp();
  if (lane < M) {
#pragma unroll
    for (int j = 0; j < M; j++) Ls[lane * LD + j] = a[j];
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < M; r++) a[r] = Ls[r * LD + col];  // a[r] = L[r][lane] for r > lane
  // backward solve L^T alpha = z, column oriented (x_r = x_r / L_rr ; x_k -= x_r L_rk)
  double alpha = 0.0;
#pragma unroll
  for (int r = M - 1; r >= 0; r--) {
    if (r >= m) continue;
    const double ar = shfl(z / dg, r);
    if (lane == r) alpha = ar;
    if (lane < r) z = z - a[r] * ar;
  }
  o.alpha = alpha;
  o.dg = dg;
  o.ok = true;
  return o;
}

// log_lik of models.py:240-252 (warp-uniform result)
template <int M>
__device__ __forceinline__ double gp_nll(double th0, double th1, double jit10,
                                         const double* __restrict__ r2s, double y, int m,
                                         int lane, double* __restrict__ Ls, double half_m_log2pi) {
  const GpSol g = gp_factor_solve<M>(th0, th1, jit10, r2s, y, m, lane, Ls);
  if (!g.ok) return dinf();
  const double ya = warp_sum((lane < m) ? y * g.alpha : 0.0);
  const double sl = numpy_sum_lanes(log(g.dg), m, lane);
  // -(-0.5*y@alpha - sum(log(diag L)) - (N/2) log(2 pi))
  const double res = -(((-0.5 * ya) - sl) - half_m_log2pi);
  return (res != res) ? dinf() : res;
}

// ---------------------------------------------------------------------------------------
// Nelder-Mead, SciPy semantics (non-adaptive, N=2): see oracle/nelder_mead.py for the
// restatement this is tested against.  All lanes carry the same simplex.  Written as a state
// machine around ONE objective call site (every branch of an iteration only chooses the next
// point to evaluate), which keeps the code small.
// ---------------------------------------------------------------------------------------
struct NMOut {
  double x0, x1, f;
  int nfev;
};

__device__ __forceinline__ void sort3(double (&sx)[3][2], double (&sf)[3]) {
  // stable insertion sort, ascending fsim (np.argsort on 3 elements is an insertion sort)
  auto swp = [&](int a, int b) {
    double t = sf[a]; sf[a] = sf[b]; sf[b] = t;
    t = sx[a][0]; sx[a][0] = sx[b][0]; sx[b][0] = t;
    t = sx[a][1]; sx[a][1] = sx[b][1]; sx[b][1] = t;
  };
  if (sf[1] < sf[0]) swp(0, 1);
  if (sf[2] < sf[1]) {
    swp(1, 2);
    if (sf[1] < sf[0]) swp(0, 1);
  }
}

enum { PH_INIT0 = 0, PH_INIT1 = 1, PH_INIT2 = 2, PH_REFLECT, PH_EXPAND, PH_OUTC, PH_INC, PH_SHRINK1, PH_SHRINK2 };

// sim[j] = sim[0] + sigma*(sim[j] - sim[0]), sigma = 1/2
__device__ __forceinline__ double shrink_to(double x0, double xj) {
  return __dadd_rn(x0, __dmul_rn(0.5, __dsub_rn(xj, x0)));
}

template <int M>
__device__ NMOut nelder_mead(double s0, double s1, double jit10, double fatol, double xatol,
                             const double* __restrict__ r2s, double y, int m, int lane,
                             double* __restrict__ Ls, double hml) {
  const int maxfun = 400, maxiter = 400;  // 200 * N
  double sx[3][2], sf[3];
  sx[0][0] = s0; sx[0][1] = s1;
  sx[1][0] = (s0 != 0.0) ? __dmul_rn(1.05, s0) : 0.00025; sx[1][1] = s1;
  sx[2][0] = s0; sx[2][1] = (s1 != 0.0) ? __dmul_rn(1.05, s1) : 0.00025;
  sf[0] = sf[1] = sf[2] = dinf();
  int fcalls = 0, it = 1, phase = PH_INIT0;
  double p0 = s0, p1 = s1;
  double xb0 = 0, xb1 = 0, xr0 = 0, xr1 = 0, fxr = 0;
  for (;;) {
    const double f = gp_nll<M>(p0, p1, jit10, r2s, y, m, lane, Ls, hml);
    fcalls++;
    bool aborted = false, do_shrink = false;
    if (phase < PH_INIT2) {
      sf[phase] = f;
      phase++;
      p0 = sx[phase][0];
      p1 = sx[phase][1];
      continue;
    }
    const bool initial = (phase == PH_INIT2);
    if (initial) {
      sf[2] = f;
    } else if (phase == PH_REFLECT) {
      fxr = f;
      int nphase = -1;
      double q0 = 0, q1 = 0;
      const double w0 = sx[2][0], w1 = sx[2][1];
      if (fxr < sf[0]) {  // expansion: (1+rho*chi)*xbar - rho*chi*worst
        q0 = __dsub_rn(__dmul_rn(3.0, xb0), __dmul_rn(2.0, w0));
        q1 = __dsub_rn(__dmul_rn(3.0, xb1), __dmul_rn(2.0, w1));
        nphase = PH_EXPAND;
      } else if (fxr < sf[1]) {
        sx[2][0] = xr0; sx[2][1] = xr1; sf[2] = fxr;
      } else if (fxr < sf[2]) {  // outside contraction: (1+psi*rho)*xbar - psi*rho*worst
        q0 = __dsub_rn(__dmul_rn(1.5, xb0), __dmul_rn(0.5, w0));
        q1 = __dsub_rn(__dmul_rn(1.5, xb1), __dmul_rn(0.5, w1));
        nphase = PH_OUTC;
      } else {  // inside contraction: (1-psi)*xbar + psi*worst
        q0 = __dadd_rn(__dmul_rn(0.5, xb0), __dmul_rn(0.5, w0));
        q1 = __dadd_rn(__dmul_rn(0.5, xb1), __dmul_rn(0.5, w1));
        nphase = PH_INC;
      }
      if (nphase >= 0) {
        if (fcalls >= maxfun) {
          aborted = true;  // SciPy raises inside func(): the iteration's update is dropped
        } else {
          phase = nphase; p0 = q0; p1 = q1;
          continue;
        }
      }
    } else if (phase == PH_EXPAND) {
      if (f < fxr) { sx[2][0] = p0; sx[2][1] = p1; sf[2] = f; }
      else { sx[2][0] = xr0; sx[2][1] = xr1; sf[2] = fxr; }
    } else if (phase == PH_OUTC) {
      if (f <= fxr) { sx[2][0] = p0; sx[2][1] = p1; sf[2] = f; }
      else do_shrink = true;
    } else if (phase == PH_INC) {
      if (f < sf[2]) { sx[2][0] = p0; sx[2][1] = p1; sf[2] = f; }
      else do_shrink = true;
    } else if (phase == PH_SHRINK1) {
      sf[1] = f;
      sx[2][0] = shrink_to(sx[0][0], sx[2][0]);
      sx[2][1] = shrink_to(sx[0][1], sx[2][1]);
      if (fcalls >= maxfun) {
        aborted = true;  // vertex moved, value kept
      } else {
        phase = PH_SHRINK2; p0 = sx[2][0]; p1 = sx[2][1];
        continue;
      }
    } else {  // PH_SHRINK2
      sf[2] = f;
    }
    if (do_shrink) {
      sx[1][0] = shrink_to(sx[0][0], sx[1][0]);
      sx[1][1] = shrink_to(sx[0][1], sx[1][1]);
      if (fcalls >= maxfun) {
        aborted = true;
      } else {
        phase = PH_SHRINK1; p0 = sx[1][0]; p1 = sx[1][1];
        continue;
      }
    }
    // end of an iteration (or of the three initial evaluations)
    if (!initial && !aborted) it++;
    sort3(sx, sf);
    if (!(fcalls < maxfun && it < maxiter)) break;
    {
      const double d10 = fabs(__dsub_rn(sx[1][0], sx[0][0])), d11 = fabs(__dsub_rn(sx[1][1], sx[0][1]));
      const double d20 = fabs(__dsub_rn(sx[2][0], sx[0][0])), d21 = fabs(__dsub_rn(sx[2][1], sx[0][1]));
      const double f1 = fabs(__dsub_rn(sf[0], sf[1])), f2 = fabs(__dsub_rn(sf[0], sf[2]));
      // np.max(...) <= tol is false as soon as any term is NaN
      if (d10 <= xatol && d11 <= xatol && d20 <= xatol && d21 <= xatol && f1 <= fatol && f2 <= fatol)
        break;
    }
    xb0 = __dmul_rn(__dadd_rn(sx[0][0], sx[1][0]), 0.5);
    xb1 = __dmul_rn(__dadd_rn(sx[0][1], sx[1][1]), 0.5);
    xr0 = __dsub_rn(__dmul_rn(2.0, xb0), sx[2][0]);
    xr1 = __dsub_rn(__dmul_rn(2.0, xb1), sx[2][1]);
    phase = PH_REFLECT; p0 = xr0; p1 = xr1;
  }
  NMOut o;
  o.x0 = sx[0][0];
  o.x1 = sx[0][1];
  o.f = fmin(fmin(sf[0], sf[1]), sf[2]);
  o.nfev = fcalls;
  return o;
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
struct FitArgs {
  const long long* idx;   // [nq,m]
  const double* dist;     // [nq,m]
  const double* r2;       // [nq,m,m]
  const double* Y;        // dataset targets [rows,d]
  const signed char* starts;  // [nq,d,9,R,2]
  const double* add;      // optional [nq,d] added to the prediction (sweep: uG_next)
  double* pred;           // [nq,d]
  double* theta_opt;      // [nq,d,2]
  double* jitter_opt;     // [nq,d]
  double* fval_opt;       // [nq,d]
  int* nfev;              // [nq,d,9,R]
  double* fvals;          // [nq,d,9,R]
  double* thetas;         // [nq,d,9,R,2]
  unsigned long long* counters;  // [0] NM runs, [1] nll evaluations
  int d, m, R;
  long long ld_pred;      // row stride of pred / add
  double fatol, xatol;
};

template <int M>
__device__ __forceinline__ void load_problem(const long long* idx, const double* dist,
                                             const double* r2, const double* Y, int d, int m,
                                             int q, int j, double* r2s, double* ys, double* kqs) {
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int a = e / M, b = e - a * M;
    r2s[e] = (a < m && b < m) ? r2[((long long)q * m + a) * m + b] : 0.0;
  }
  for (int r = threadIdx.x; r < M; r += blockDim.x) {
    ys[r] = (r < m) ? Y[idx[(long long)q * m + r] * d + j] : 0.0;
    kqs[r] = (r < m && dist != nullptr) ? dist[(long long)q * m + r] : 0.0;
  }
}

template <int M>
__device__ __forceinline__ double posterior_mean(double th0, double th1, double jit10,
                                                 const double* r2s, const double* ys,
                                                 const double* kqs, int m, int lane, double* Ls) {
  const int col = (lane < M) ? lane : 0;
  const GpSol g = gp_factor_solve<M>(th0, th1, jit10, r2s, ys[col], m, lane, Ls);
  if (!g.ok) return dnan();
  // K_star = kernel(x, new_x); post_mean = K_star.T @ alph  (models.py:165-167)
  const double ks = g.amp * exp(g.c * kqs[col]);
  return warp_sum((lane < m) ? ks * g.alpha : 0.0);
}

// CTAs per SM the fit kernel is compiled for (register cap): 288 threads x 2 CTAs -> 112 regs
template <int M> struct FitOcc { static constexpr int value = 2; };

template <int M>
__global__ void __launch_bounds__(GP_WARPS * 32, FitOcc<M>::value)
gp_fit_predict_kernel(FitArgs A) {
  extern __shared__ double sm[];
  const int m = A.m, d = A.d, R = A.R, nruns = NNGP_N_JITTER * R;
  double* r2s = sm;                       // M*M
  double* ys = r2s + M * M;               // M
  double* kqs = ys + M;                   // M
  double* Lsall = kqs + M;                // GP_WARPS * M*(M+1)
  double* rf = Lsall + GP_WARPS * M * (M + 1);  // nruns
  double* rt = rf + nruns;                // nruns*2
  const int j = blockIdx.x, q = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  load_problem<M>(A.idx, A.dist, A.r2, A.Y, d, m, q, j, r2s, ys, kqs);
  __syncthreads();
  double* Ls = Lsall + w * M * (M + 1);
  const double y = ys[(lane < M) ? lane : 0];
  const double hml = (m / 2.0) * 1.8378770664093453;  // (N/2)*np.log(2*np.pi)
  const long long task0 = ((long long)q * d + j) * nruns;
  for (int run = w; run < nruns; run += GP_WARPS) {
    const int a = run / R;
    const signed char* st = A.starts + (task0 + run) * 2;
    const NMOut o = nelder_mead<M>((double)st[0], (double)st[1], c_jit10[a], A.fatol, A.xatol,
                                   r2s, y, m, lane, Ls, hml);
    if (lane == 0) {
      rf[run] = o.f;
      rt[2 * run] = o.x0;
      rt[2 * run + 1] = o.x1;
      atomicAdd(A.counters, 1ULL);
      atomicAdd(A.counters + 1, (unsigned long long)o.nfev);
      if (A.nfev) A.nfev[task0 + run] = o.nfev;
      if (A.fvals) A.fvals[task0 + run] = o.f;
      if (A.thetas) {
        A.thetas[(task0 + run) * 2] = o.x0;
        A.thetas[(task0 + run) * 2 + 1] = o.x1;
      }
    }
  }
  __syncthreads();
  if (w == 0) {
    // models.py:212-215: mask = fval < 0.9*min; empty mask -> all; first minimum in task order
    double fmin_all = rf[0];
    for (int r = 1; r < nruns; r++) fmin_all = (rf[r] < fmin_all) ? rf[r] : fmin_all;
    const double thr = fmin_all * 0.9;
    bool any = false;
    for (int r = 0; r < nruns; r++) any |= (rf[r] < thr);
    int best = -1;
    double fb = 0.0;
    for (int r = 0; r < nruns; r++) {
      if (any && !(rf[r] < thr)) continue;
      if (best < 0 || rf[r] < fb) {
        best = r;
        fb = rf[r];
      }
    }
    const int a = best / R;
    const double th0 = rt[2 * best], th1 = rt[2 * best + 1];
    double mean = posterior_mean<M>(th0, th1, c_jit10[a], r2s, ys, kqs, m, lane, Ls);
    if (lane == 0) {
      const long long o = (long long)q * A.ld_pred + j;
      if (A.add) mean = mean + A.add[o];
      A.pred[o] = mean;
      const long long oo = (long long)q * d + j;
      if (A.theta_opt) {
        A.theta_opt[oo * 2] = th0;
        A.theta_opt[oo * 2 + 1] = th1;
      }
      if (A.jitter_opt) A.jitter_opt[oo] = (double)(a - 20);
      if (A.fval_opt) A.fval_opt[oo] = fb;
    }
  }
}

// objective at given hyper-parameters: one warp per (q, j, t)
template <int M>
__global__ void __launch_bounds__(GP_WARPS * 32)
gp_nll_kernel(const long long* idx, const double* r2, const double* Y, int d, int m, int nt,
              const double* theta, const double* jitter10, double* out) {
  extern __shared__ double sm[];
  double* r2s = sm;
  double* ys = r2s + M * M;
  double* kqs = ys + M;
  double* Lsall = kqs + M;
  const int j = blockIdx.x, q = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  load_problem<M>(idx, nullptr, r2, Y, d, m, q, j, r2s, ys, kqs);
  __syncthreads();
  double* Ls = Lsall + w * M * (M + 1);
  const double y = ys[(lane < M) ? lane : 0];
  const double hml = (m / 2.0) * 1.8378770664093453;
  const long long base = ((long long)q * d + j) * nt;
  for (int t = w; t < nt; t += GP_WARPS) {
    const double v = gp_nll<M>(theta[(base + t) * 2], theta[(base + t) * 2 + 1], jitter10[base + t],
                               r2s, y, m, lane, Ls, hml);
    if (lane == 0) out[base + t] = v;
  }
}

// posterior mean at given hyper-parameters: one warp per (q, j)
template <int M>
__global__ void __launch_bounds__(GP_WARPS * 32)
gp_mean_kernel(const long long* idx, const double* dist, const double* r2, const double* Y, int d,
               int m, const double* theta, const double* jitter, double* pred) {
  extern __shared__ double sm[];
  double* r2s = sm;
  double* Lsall = r2s + M * M;            // GP_WARPS * M*(M+1)
  double* ysall = Lsall + GP_WARPS * M * (M + 1);  // GP_WARPS * M
  double* kqs = ysall + GP_WARPS * M;     // M
  const int q = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * GP_WARPS + w;
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int a = e / M, b = e - a * M;
    r2s[e] = (a < m && b < m) ? r2[((long long)q * m + a) * m + b] : 0.0;
  }
  for (int r = threadIdx.x; r < M; r += blockDim.x) kqs[r] = (r < m) ? dist[(long long)q * m + r] : 0.0;
  double* ys = ysall + w * M;
  if (j < d && lane < M) ys[lane] = (lane < m) ? Y[idx[(long long)q * m + lane] * d + j] : 0.0;
  __syncthreads();
  if (j >= d) return;
  const long long o = (long long)q * d + j;
  const double jit10 = exp10(jitter[o]);
  const double mean = posterior_mean<M>(theta[o * 2], theta[o * 2 + 1], jit10, r2s, ys, kqs, m, lane,
                                        Lsall + w * M * (M + 1));
  if (lane == 0) pred[o] = mean;
}

// pairwise squared distances of the m neighbours of each query: r2[q,a,b] = ||x_a - x_b||^2,
// summed strictly left to right (cdist arithmetic; this file allows FMA, so use intrinsics).
__global__ void __launch_bounds__(256)
gp_prep_kernel(const long long* __restrict__ idx, const double* __restrict__ X, int d, int m,
               double* __restrict__ r2) {
  constexpr int JC = 64;
  __shared__ double tile[NNGP_MAX_NEIGHBOURS][JC + 1];
  const int q = blockIdx.x;
  const int npairs = m * (m + 1) / 2;
  // up to 3 pairs per thread (m=32 -> 528 pairs on 256 threads)
  int pa[3], pb[3];
  double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int u = 0; u < 3; u++) {
    const int pidx = threadIdx.x + u * 256;
    int a = 0, b = 0;
    if (pidx < npairs) {
      a = (int)((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5);
      while ((a + 1) * (a + 2) / 2 <= pidx) a++;
      while (a * (a + 1) / 2 > pidx) a--;
      b = pidx - a * (a + 1) / 2;
    }
    pa[u] = a;
    pb[u] = b;
  }
  for (int j0 = 0; j0 < d; j0 += JC) {
    const int jn = min(JC, d - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < m * JC; e += 256) {
      const int r = e / JC, jj = e - r * JC;
      if (jj < jn) tile[r][jj] = X[idx[(long long)q * m + r] * d + j0 + jj];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 3; u++) {
      if (threadIdx.x + u * 256 < npairs) {
        double s = acc[u];
        for (int jj = 0; jj < jn; jj++) {
          const double diff = __dsub_rn(tile[pa[u]][jj], tile[pb[u]][jj]);
          s = __dadd_rn(s, __dmul_rn(diff, diff));
        }
        acc[u] = s;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 3; u++) {
    if (threadIdx.x + u * 256 < npairs) {
      r2[((long long)q * m + pa[u]) * m + pb[u]] = acc[u];
      r2[((long long)q * m + pb[u]) * m + pa[u]] = acc[u];
    }
  }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
size_t gp_prep_bytes(int nq, int m) { return sizeof(double) * (size_t)nq * m * m; }

int gp_prep_launch(nngp_handle_t h, const long long* d_idx, int nq, int m, double* d_r2,
                   cudaStream_t st) {
  if (nq <= 0) return 0;
  ProfScope prof(h, 2, st);
  gp_prep_kernel<<<nq, 256, 0, st>>>(d_idx, h->ds_x, h->ds_d, m, d_r2);
  h->launches++;
  NNGP_CUDA(h, cudaGetLastError());
  return 0;
}

template <int M>
static size_t fit_smem(int R) {
  return sizeof(double) * ((size_t)M * M + 2 * M + (size_t)GP_WARPS * M * (M + 1) + 3 * (size_t)NNGP_N_JITTER * R);
}

template <int M>
static int fit_launch_m(nngp_handle_t h, const FitArgs& A, int nq, cudaStream_t st) {
  const size_t smem = fit_smem<M>(A.R);
  if (smem > 227 * 1024) return nngp_fail(h, "fit: shared memory %zu too large (R=%d)", smem, A.R);
  static bool attr_set = false;
  if (!attr_set) {
    NNGP_CUDA(h, cudaFuncSetAttribute(gp_fit_predict_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  ProfScope prof(h, 3, st);
  gp_fit_predict_kernel<M><<<dim3(A.d, nq), GP_WARPS * 32, smem, st>>>(A);
  h->launches++;
  NNGP_CUDA(h, cudaGetLastError());
  return 0;
}

#define DISPATCH_M(m, CALL)                                                   \
  do {                                                                        \
    if ((m) <= 12) { constexpr int MM = 12; CALL; }                           \
    else if ((m) <= 16) { constexpr int MM = 16; CALL; }                      \
    else if ((m) <= 20) { constexpr int MM = 20; CALL; }                      \
    else if ((m) <= 24) { constexpr int MM = 24; CALL; }                      \
    else { constexpr int MM = 32; CALL; }                                     \
  } while (0)

int gp_fit_predict_launch(nngp_handle_t h, const long long* d_idx, const double* d_dist,
                          const double* d_r2, int nq, int m, int R, const signed char* d_starts,
                          double fatol, double xatol, double* d_pred, const double* d_add,
                          double* d_theta_opt, double* d_jitter_opt, double* d_fval_opt,
                          int* d_nfev, double* d_fvals, double* d_thetas, cudaStream_t st) {
  if (nq <= 0) return 0;
  if (m < 1 || m > NNGP_MAX_NEIGHBOURS) return nngp_fail(h, "fit: m=%d outside [1,%d]", m, NNGP_MAX_NEIGHBOURS);
  if (R < 1) return nngp_fail(h, "fit: n_restarts=%d < 1", R);
  FitArgs A;
  A.idx = d_idx; A.dist = d_dist; A.r2 = d_r2; A.Y = h->ds_y; A.starts = d_starts; A.add = d_add;
  A.pred = d_pred; A.theta_opt = d_theta_opt; A.jitter_opt = d_jitter_opt; A.fval_opt = d_fval_opt;
  A.nfev = d_nfev; A.fvals = d_fvals; A.thetas = d_thetas; A.counters = h->d_counters;
  A.d = h->ds_d; A.m = m; A.R = R; A.ld_pred = h->ds_d; A.fatol = fatol; A.xatol = xatol;
  int rc = 0;
  DISPATCH_M(m, rc = fit_launch_m<MM>(h, A, nq, st));
  return rc;
}

template <int M>
static int nll_launch_m(nngp_handle_t h, const long long* idx, const double* r2, int nq, int m,
                        int nt, const double* theta, const double* j10, double* out,
                        cudaStream_t st) {
  const size_t smem = sizeof(double) * ((size_t)M * M + 2 * M + (size_t)GP_WARPS * M * (M + 1));
  static bool attr_set = false;
  if (!attr_set) {
    NNGP_CUDA(h, cudaFuncSetAttribute(gp_nll_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  gp_nll_kernel<M><<<dim3(h->ds_d, nq), GP_WARPS * 32, smem, st>>>(idx, r2, h->ds_y, h->ds_d, m, nt, theta, j10, out);
  h->launches++;
  NNGP_CUDA(h, cudaGetLastError());
  return 0;
}

int gp_nll_launch(nngp_handle_t h, const long long* d_idx, const double* d_r2, int nq, int m,
                  int nt, const double* d_theta, const double* d_jitter10, double* d_nll,
                  cudaStream_t st) {
  if (nq <= 0 || nt <= 0) return 0;
  if (m < 1 || m > NNGP_MAX_NEIGHBOURS) return nngp_fail(h, "nll: m=%d outside [1,%d]", m, NNGP_MAX_NEIGHBOURS);
  int rc = 0;
  DISPATCH_M(m, rc = nll_launch_m<MM>(h, d_idx, d_r2, nq, m, nt, d_theta, d_jitter10, d_nll, st));
  return rc;
}

template <int M>
static int mean_launch_m(nngp_handle_t h, const long long* idx, const double* dist,
                         const double* r2, int nq, int m, const double* theta,
                         const double* jitter, double* pred, cudaStream_t st) {
  const size_t smem = sizeof(double) * ((size_t)M * M + (size_t)GP_WARPS * M * (M + 1) + (size_t)GP_WARPS * M + M);
  static bool attr_set = false;
  if (!attr_set) {
    NNGP_CUDA(h, cudaFuncSetAttribute(gp_mean_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int d = h->ds_d;
  gp_mean_kernel<M><<<dim3((d + GP_WARPS - 1) / GP_WARPS, nq), GP_WARPS * 32, smem, st>>>(idx, dist, r2, h->ds_y, d, m, theta, jitter, pred);
  h->launches++;
  NNGP_CUDA(h, cudaGetLastError());
  return 0;
}

int gp_mean_launch(nngp_handle_t h, const long long* d_idx, const double* d_dist,
                   const double* d_r2, int nq, int m, const double* d_theta,
                   const double* d_jitter, double* d_pred, cudaStream_t st) {
  if (nq <= 0) return 0;
  if (m < 1 || m > NNGP_MAX_NEIGHBOURS) return nngp_fail(h, "mean: m=%d outside [1,%d]", m, NNGP_MAX_NEIGHBOURS);
  int rc = 0;
  DISPATCH_M(m, rc = mean_launch_m<MM>(h, d_idx, d_dist, d_r2, nq, m, d_theta, d_jitter, d_pred, st));
  return rc;
}
