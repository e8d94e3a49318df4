// GP fit for 32 < m <= NNGP_MAX_NEIGHBOURS_BIG neighbours (included by gpfit.cu).
//
// nn='adaptive' makes m = max(10, k+2) (models.py:172-175): past iteration 30 the neighbour set no longer fits one
// lane per row (the Thomas-labyrinth runs of the reference reach K = 159).  Rare and small (d = 3 there), so this
// path is written for generality, not for the roofline: one CTA per Nelder-Mead search, the kernel matrix in shared
// memory, right-looking square-root-free LDL^T with the forward solve fused, the same failed-pivot rule and the same
// Nelder-Mead step function (nm_step) as the warp kernels; selection and posterior mean in a second kernel, one CTA
// per (query, dimension).
#pragma once

static constexpr int BIG_THREADS = 128;

// neighbour matrix for any m: one thread per pair (a >= b), strict left-to-right sums like gp_prep_kernel
__global__ void __launch_bounds__(256)
gp_prep_big_kernel(const long long* __restrict__ idx, const double* __restrict__ X, int d, int m, int nq,
                   double* __restrict__ r2) {
  const int npairs = m * (m + 1) / 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nq * npairs) return;
  const int q = (int)(t / npairs), pidx = (int)(t - (long long)q * npairs);
  int a = (int)((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= pidx) a++;
  while (a * (a + 1) / 2 > pidx) a--;
  const int b = pidx - a * (a + 1) / 2;
  const double* xa = X + idx[(long long)q * m + a] * d;
  const double* xb = X + idx[(long long)q * m + b] * d;
  double s = 0.0;
  for (int j = 0; j < d; j++) {
    const double diff = __dsub_rn(xa[j], xb[j]);
    s = __dadd_rn(s, __dmul_rn(diff, diff));
  }
  r2[((long long)q * m + a) * m + b] = s;
  r2[((long long)q * m + b) * m + a] = s;
}

// objective (ALPHA = false: returns nll, +inf on failure) or alpha = K^-1 y left in zs[] (ALPHA = true: returns 0 / NaN);
// A: m x m lower triangle in shared memory (row stride ld), zs: m doubles of shared memory, red: >= 2 doubles
template <bool ALPHA>
__device__ double gp_big_eval(double th0, double th1, double jit10, const double* __restrict__ r2, const double* ys,
                              int m, int ld, double* A, double* zs, double* dpiv, double* red, double guard,
                              double& amp_out, double& c_out) {
  const int tid = threadIdx.x;
  double amp, inv;
  exp10_pair(th1, -th0, amp, inv);
  const double c = -0.5 * inv;
  amp_out = amp;
  c_out = c;
  const double dd0 = fma(amp, exp_neg(c * 0.0), jit10);
  const double pmin = dd0 * guard;
  for (int e = tid; e < m * m; e += BIG_THREADS) {
    const int i = e / m, j = e - i * m;
    if (j < i) A[i * ld + j] = amp * exp_neg(c * r2[i * m + j]);
    else if (j == i) A[i * ld + i] = dd0;
  }
  for (int i = tid; i < m; i += BIG_THREADS) zs[i] = ys[i];
  __syncthreads();
  bool ok = (dd0 == dd0);
  double quad = 0.0, logdet = 0.0;
  for (int k = 0; k < m; k++) {
    const double p = A[k * ld + k];
    ok = ok && (p > pmin);
    if (!ok) break;  // uniform: every thread reads the same pivot
    const double ip = 1.0 / p;
    const double zk = zs[k];
    quad = fma(zk * zk, ip, quad);
    logdet += log(p);
    if (tid == 0) dpiv[k] = p;
    // rows i > k: w = a_ik / d_k; z_i -= w z_k; a_ij -= w a_jk for k < j <= i (a_ii is the running pivot of row i)
    for (int i = k + 1 + tid; i < m; i += BIG_THREADS) {
      const double w = A[i * ld + k] * ip;
      zs[i] = fma(-w, zk, zs[i]);
    }
    const int nrem = m - k - 1;
    for (int e = tid; e < nrem * nrem; e += BIG_THREADS) {
      const int ii = e / nrem, jj = e - ii * nrem;
      if (jj <= ii) {
        const int i = k + 1 + ii, j = k + 1 + jj;
        const double w = A[i * ld + k] * ip;
        A[i * ld + j] = fma(-w, A[j * ld + k], A[i * ld + j]);
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (!ALPHA) {
    const double res = 0.5 * quad + 0.5 * logdet + (m / 2.0) * 1.8378770664093453;
    return (ok && res == res) ? res : dinf();
  }
  if (!ok) return dnan();
  // alpha = L'^-T D^-1 z: backward, column oriented (l'_ik = a_ik / d_k still sits unscaled in A)
  for (int i = tid; i < m; i += BIG_THREADS) zs[i] = zs[i] / dpiv[i];
  __syncthreads();
  for (int k = m - 1; k > 0; k--) {
    const double ak = zs[k];
    const double ipk = 1.0 / dpiv[k];
    (void)ipk;
    // row k of L' couples alpha_k into the rows j < k: alpha_j -= l'_kj alpha_k, l'_kj = a_kj / d_j
    for (int j = tid; j < k; j += BIG_THREADS) zs[j] = fma(-(A[k * ld + j] / dpiv[j]), ak, zs[j]);
    __syncthreads();
  }
  return 0.0;
}

__global__ void __launch_bounds__(BIG_THREADS)
gp_fit_big_kernel(FitArgs A_) {
  extern __shared__ double sm[];
  const FitArgs& A = A_;
  const int m = A.m, d = A.d, R = A.R, nruns = NNGP_N_JITTER * R;
  const int ld = m | 1;
  double* Am = sm;
  double* zs = Am + (size_t)m * ld;
  double* ys = zs + m;
  double* dpiv = ys + m;
  double* red = dpiv + m;
  __shared__ int s_task;
  for (;;) {
    if (threadIdx.x == 0) {
      const unsigned pos = atomicAdd(A.queue, 1u);
      s_task = (pos < (unsigned)A.ntasks) ? (A.order ? A.order[pos] : (int)pos) : -1;
    }
    __syncthreads();
    const int task = s_task;
    __syncthreads();
    if (task < 0) break;
    const int qj = task / nruns, run = task - qj * nruns;
    const int q = qj / A.dl, j = A.j0 + (qj - q * A.dl);
    const long long gqj = (long long)q * d + j, gtask = gqj * nruns + run;
    const double jit10 = c_jit10[run / R];
    const double* r2 = A.r2 + (long long)q * m * m;
    for (int i = threadIdx.x; i < m; i += BIG_THREADS) ys[i] = A.Y[A.idx[(long long)q * m + i] * d + j];
    __syncthreads();
    NMState S;
    const signed char* st = A.starts + gtask * 2;
    nm_init(S, (double)st[0], (double)st[1]);
    bool fin = false;
    while (!fin) {
      double amp, c;
      const double f = gp_big_eval<false>(S.p0, S.p1, jit10, r2, ys, m, ld, Am, zs, dpiv, red, A.guard, amp, c);
      fin = nm_step(S, f, A.fatol, A.xatol, false);
    }
    if (threadIdx.x == 0) {
      const double fbest = fmin(fmin(S.sf[0], S.sf[1]), S.sf[2]);
      A.res[(long long)task * 3] = fbest;
      A.res[(long long)task * 3 + 1] = S.sx[0][0];
      A.res[(long long)task * 3 + 2] = S.sx[0][1];
      atomicAdd(A.counters, 1ULL);
      atomicAdd(A.counters + 1, (unsigned long long)S.fcalls);
      if (A.nfev) A.nfev[gtask] = S.fcalls;
      if (A.fvals) A.fvals[gtask] = fbest;
      if (A.thetas) {
        A.thetas[gtask * 2] = S.sx[0][0];
        A.thetas[gtask * 2 + 1] = S.sx[0][1];
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(BIG_THREADS)
gp_select_mean_big_kernel(FitArgs A_, int nqj) {
  extern __shared__ double sm[];
  const FitArgs& A = A_;
  const int m = A.m, d = A.d, R = A.R, nruns = NNGP_N_JITTER * R;
  const int ld = m | 1;
  double* Am = sm;
  double* zs = Am + (size_t)m * ld;
  double* ys = zs + m;
  double* dpiv = ys + m;
  double* red = dpiv + m;
  const int qj = blockIdx.x;
  if (qj >= nqj) return;
  const int q = qj / A.dl, j = A.j0 + (qj - q * A.dl);
  const long long gqj = (long long)q * d + j;
  const double* rf = A.res + (long long)qj * nruns * 3;
  double fmin_all = rf[0];
  for (int r = 1; r < nruns; r++) fmin_all = (rf[3 * r] < fmin_all) ? rf[3 * r] : fmin_all;
  const double thr = fmin_all * 0.9;
  bool any = false;
  for (int r = 0; r < nruns; r++) any |= (rf[3 * r] < thr);
  int best = -1;
  double fb = 0.0;
  for (int r = 0; r < nruns; r++) {
    const double v = rf[3 * r];
    if (any && !(v < thr)) continue;
    if (best < 0 || v < fb) {
      best = r;
      fb = v;
    }
  }
  const int ab = best / R;
  const double th0 = rf[3 * best + 1], th1 = rf[3 * best + 2];
  for (int i = threadIdx.x; i < m; i += BIG_THREADS) ys[i] = A.Y[A.idx[(long long)q * m + i] * d + j];
  __syncthreads();
  double amp, c;
  const double st = gp_big_eval<true>(th0, th1, c_jit10[ab], A.r2 + (long long)q * m * m, ys, m, ld, Am, zs, dpiv, red,
                                      A.guard, amp, c);
  __syncthreads();
  // post_mean = sum_i k(x_i, q) alpha_i
  double part = 0.0;
  for (int i = threadIdx.x; i < m; i += BIG_THREADS) part += amp * exp_neg(c * A.dist[(long long)q * m + i]) * zs[i];
  part = warp_sum(part);
  __shared__ double wsum[BIG_THREADS / 32];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double mean = 0.0;
    for (int w = 0; w < BIG_THREADS / 32; w++) mean += wsum[w];
    if (st != st) mean = dnan();
    const long long op = (long long)q * A.ld_pred + j;
    if (A.add) mean = mean + A.add[op];
    A.pred[op] = mean;
    if (A.theta_opt) {
      A.theta_opt[gqj * 2] = th0;
      A.theta_opt[gqj * 2 + 1] = th1;
    }
    if (A.jitter_opt) A.jitter_opt[gqj] = (double)(ab - 20);
    if (A.fval_opt) A.fval_opt[gqj] = fb;
  }
}
