// extern "C" entry points of libnngpara.so (see include/nngpara.h), handle management,
// host-buffer variants, the fused on-device sweep and the roofline micro-benchmarks.
#include "common.cuh"

#include <cstdarg>
#include <cstdlib>
#include <cstring>

#define NNGP_QUEUE_SLOTS 4096

static std::string g_create_error;

int nngp_fail(nngp_handle_t h, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return -1;
}

static inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

// every entry point makes the handle's device current first: a process may hold handles on several devices
#define NNGP_USE_DEVICE(h)                                                                  \
  do {                                                                                      \
    if ((h) != nullptr) cudaSetDevice((h)->device);                                         \
  } while (0)

ProfScope::ProfScope(nngp_handle_t h_, int cls, cudaStream_t st_) : h(h_), st(st_), on(h_->profiling) {
  if (!on) return;
  nngp_handle_s::ProfRec r;
  r.cls = cls;
  for (cudaEvent_t* e : {&r.a, &r.b}) {
    if (!h->prof_free.empty()) {
      *e = h->prof_free.back();
      h->prof_free.pop_back();
    } else {
      cudaEventCreate(e);
    }
  }
  cudaEventRecord(r.a, st);
  h->prof_recs.push_back(r);
}

ProfScope::~ProfScope() {
  if (on) cudaEventRecord(h->prof_recs.back().b, st);
}

// a zeroed task-queue head for the next fit launch on stream st (slots are re-zeroed in bulk)
// (a fit uses up to three counters: its task queue, the number of parked long searches and their queue)
static unsigned int* next_queue(nngp_handle_t h, cudaStream_t st) {
  if (h->queue_next + 4 > NNGP_QUEUE_SLOTS) {
    // the bulk re-zeroing is ordered after the fits of `st` only: a fit still running on another stream of this
    // handle (host entry points use the handle's own stream, device entry points the caller's) must finish first
    if (h->queue_stream_valid && h->queue_stream != st) cudaStreamSynchronize(h->queue_stream);
    cudaMemsetAsync(h->d_queues, 0, NNGP_QUEUE_SLOTS * sizeof(unsigned int), st);
    h->queue_next = 0;
  }
  if (h->queue_stream_valid && h->queue_stream != st && h->queue_next != 0) {
    // the stream changed in the middle of a ring: the slots handed out so far may still be in use there, and a later
    // wrap on either stream must not overtake them -- cheap (a stream switch is rare) and keeps the invariant simple
    cudaStreamSynchronize(h->queue_stream);
  }
  h->queue_stream = st;
  h->queue_stream_valid = true;
  unsigned int* p = h->d_queues + h->queue_next;
  h->queue_next += 4;
  return p;
}

// the fit kernel leaves its completion counters zero; they only need clearing when they move
static int ensure_done_zero(nngp_handle_t h, void* ws_fit, int nq, int d, int m, int R, cudaStream_t st) {
  const char* ptr = (const char*)ws_fit + gp_fit_done_offset(nq, d, m, R);
  const size_t bytes = sizeof(unsigned int) * (size_t)nq * d;
  if (h->done_ptr != ptr || h->done_bytes != bytes) {
    NNGP_CUDA(h, cudaMemsetAsync((void*)ptr, 0, bytes, st));
    h->done_ptr = ptr;
    h->done_bytes = bytes;
  }
  return 0;
}

void* nngp_workspace(nngp_handle_t h, size_t bytes) {
  // whoever asks for the block may overwrite the (left-zero) completion counters of an earlier fit:
  // the next fit clears them again (ensure_done_zero)
  h->done_ptr = nullptr;
  if (bytes <= h->ws_bytes) return h->ws;
  // grow: work already enqueued may still use the old block
  cudaDeviceSynchronize();
  if (h->ws) cudaFree(h->ws);
  h->ws = nullptr;
  h->ws_bytes = 0;
  size_t want = bytes + bytes / 4 + (1 << 20);
  if (cudaMalloc(&h->ws, want) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  h->ws_bytes = want;
  return h->ws;
}

static void* pinned_buf(nngp_handle_t h, size_t bytes) {
  if (bytes <= h->pinned_bytes) return h->pinned;
  if (h->pinned) cudaFreeHost(h->pinned);
  h->pinned = nullptr;
  h->pinned_bytes = 0;
  size_t want = bytes + bytes / 4 + (1 << 16);
  if (cudaMallocHost(&h->pinned, want) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  h->pinned_bytes = want;
  return h->pinned;
}

static void* stage_buf(nngp_handle_t h, size_t bytes) {
  if (bytes <= h->stage_bytes) return h->stage;
  cudaDeviceSynchronize();
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr;
  h->stage_bytes = 0;
  size_t want = bytes + bytes / 4 + (1 << 16);
  if (cudaMalloc(&h->stage, want) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  h->stage_bytes = want;
  return h->stage;
}

// simple carve-out helper over a byte block (256 B aligned pieces)
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  template <class T> T* take(size_t count) {
    T* p = (T*)(base + off);
    off += ((count * sizeof(T) + 255) / 256) * 256;
    return p;
  }
  static size_t pad(size_t bytes) { return ((bytes + 255) / 256) * 256; }
};

extern "C" {

int nngp_abi_version(void) { return NNGP_ABI_VERSION; }

int nngp_create(int device, nngp_handle_t* out) {
  if (!out) return nngp_fail(nullptr, "nngp_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return nngp_fail(nullptr, "nngp_create: no CUDA device (%s); this library has no CPU fallback",
                     cudaGetErrorString(e));
  if (device < 0 || device >= count) return nngp_fail(nullptr, "nngp_create: device %d of %d", device, count);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return nngp_fail(nullptr, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10)
    return nngp_fail(nullptr, "nngp_create: device sm_%d%d is not Blackwell sm_100a", prop.major, prop.minor);
  nngp_handle_t h = new nngp_handle_s();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  if (const char* fm = getenv("NNGP_FIT_MODE")) h->fit_mode = (strcmp(fm, "warp") == 0) ? 1 : (strcmp(fm, "grouped") == 0) ? 2 : (strcmp(fm, "quad") == 0) ? 3 : 0;
  if (getenv("NNGP_FIT_LEGACY")) h->fit_mode = 1;
  if (const char* b = getenv("NNGP_FIT_BUDGET")) h->fit_budget = atoi(b);
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return nngp_fail(nullptr, "cudaStreamCreate failed");
  }
  if (cudaMalloc(&h->d_counters, 8 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMemset(h->d_counters, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) {
    delete h;
    return nngp_fail(nullptr, "cudaMalloc(counters) failed");
  }
  if (cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc(&h->d_ticket, 8 * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(h->d_ticket, 0, 8 * sizeof(unsigned int)) != cudaSuccess) {
    delete h;
    return nngp_fail(nullptr, "nngp_create: auxiliary stream / events failed");
  }
  if (cudaMalloc(&h->d_queues, NNGP_QUEUE_SLOTS * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(h->d_queues, 0, NNGP_QUEUE_SLOTS * sizeof(unsigned int)) != cudaSuccess) {
    delete h;
    return nngp_fail(nullptr, "cudaMalloc(queues) failed");
  }
  if (const char* g = getenv("NNGP_PIVOT_GUARD_ULPS")) {
    if (nngp_set_pivot_guard(h, atof(g)) != 0) {
      g_create_error = h->err;
      delete h;
      return -1;
    }
  }
  if (rk_set_tableaus(h) != 0) {
    g_create_error = h->err;
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}

int nngp_destroy(nngp_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& s : h->systems) {
    if (s.d_mn) cudaFree(s.d_mn);
    if (s.d_mx) cudaFree(s.d_mx);
  }
  if (h->ds_x) cudaFree(h->ds_x);
  if (h->ds_y) cudaFree(h->ds_y);
  if (h->ds_xt) cudaFree(h->ds_xt);
  if (h->ws) cudaFree(h->ws);
  if (h->stage) cudaFree(h->stage);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->d_ticket) cudaFree(h->d_ticket);
  if (h->d_counters) cudaFree(h->d_counters);
  if (h->d_queues) cudaFree(h->d_queues);
  for (auto& r : h->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->prof_free) cudaEventDestroy(e);
  delete h;
  return 0;
}

const char* nngp_last_error(nngp_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int nngp_synchronize(nngp_handle_t h, void* stream) {
  NNGP_USE_DEVICE(h);
  NNGP_CUDA(h, cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

long long nngp_launch_count(nngp_handle_t h) { return h->launches; }

int nngp_set_pivot_guard(nngp_handle_t h, double ulps) {
  if (!(ulps == ulps) || ulps < 0.0 || ulps > 1e6) return nngp_fail(h, "set_pivot_guard: ulps=%g outside [0, 1e6]", ulps);
  // below 2^-10 ulp the running product of the pivots could leave the normal range: "0" means "as small as possible"
  if (ulps < 9.765625e-4) ulps = 9.765625e-4;
  h->pivot_guard = ulps * 2.220446049250313e-16;
  return 0;
}

int nngp_set_fit_mode(nngp_handle_t h, int mode) {
  if (mode < 0 || mode > 3) return nngp_fail(h, "set_fit_mode: mode=%d outside {0 auto, 1 warp, 2 grouped, 3 quad}", mode);
  h->fit_mode = mode;
  return 0;
}

int nngp_set_fit_budget(nngp_handle_t h, int evaluations) {
  if (evaluations < 0 || evaluations > 400) return nngp_fail(h, "set_fit_budget: %d outside [0, 400]", evaluations);
  h->fit_budget = evaluations;
  return 0;
}

double nngp_get_pivot_guard(nngp_handle_t h) { return h->pivot_guard / 2.220446049250313e-16; }

int nngp_counters(nngp_handle_t h, long long* nm_runs, long long* nll_evals, int reset) {
  NNGP_USE_DEVICE(h);
  unsigned long long v[8];
  NNGP_CUDA(h, cudaDeviceSynchronize());
  NNGP_CUDA(h, cudaMemcpy(v, h->d_counters, sizeof(v), cudaMemcpyDeviceToHost));
  if (getenv("NNGP_FIT_STATS"))
    fprintf(stderr, "fit stats: searches %llu evals %llu | warp-rounds %llu wanted group-evals %llu busy group-slots %llu "
            "fast-forwarded iterations %llu max rounds of a warp (all launches) %llu\n", v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
  if (nm_runs) *nm_runs = (long long)v[0];
  if (nll_evals) *nll_evals = (long long)v[1];
  if (reset) NNGP_CUDA(h, cudaMemset(h->d_counters, 0, sizeof(v)));
  return 0;
}

int nngp_profile_enable(nngp_handle_t h, int on) {
  h->profiling = (on != 0);
  return 0;
}

int nngp_profile_read(nngp_handle_t h, double* ms, long long* counts, int reset) {
  NNGP_USE_DEVICE(h);
  NNGP_CUDA(h, cudaDeviceSynchronize());
  for (auto& r : h->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      h->prof_ms[r.cls] += t;
      h->prof_n[r.cls] += 1;
    }
    h->prof_free.push_back(r.a);
    h->prof_free.push_back(r.b);
  }
  cudaGetLastError();
  h->prof_recs.clear();
  for (int i = 0; i < NNGP_PROF_CLASSES; i++) {
    if (ms) ms[i] = h->prof_ms[i];
    if (counts) counts[i] = h->prof_n[i];
    if (reset) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  }
  return 0;
}

// ---- systems ---------------------------------------------------------------------------
int nngp_system_create(nngp_handle_t h, int system_id, int d, const double* params, int n_params,
                       int normalize, const double* mn, const double* mx, int* sys_out) {
  NNGP_USE_DEVICE(h);
  if (!sys_out) return nngp_fail(h, "system_create: sys_out is NULL");
  if (system_id < NNGP_SYS_FHN_ODE || system_id > NNGP_SYS_BURGERS)
    return nngp_fail(h, "system_create: unknown system id %d", system_id);
  if (n_params < 0 || n_params > NNGP_MAX_PARAMS) return nngp_fail(h, "system_create: n_params=%d", n_params);
  if (d < 1) return nngp_fail(h, "system_create: d=%d", d);
  if (normalize != 0 && normalize != 1) return nngp_fail(h, "Only identity and -11 are implemented");
  if (normalize && (!mn || !mx)) return nngp_fail(h, "system_create: normalize=1 needs mn and mx");
  SystemDesc s;
  s.system_id = system_id;
  s.d = d;
  s.n_params = n_params;
  for (int i = 0; i < n_params; i++) s.params[i] = params[i];
  s.normalize = normalize;
  if (normalize) {
    NNGP_CUDA(h, cudaMalloc(&s.d_mn, sizeof(double) * d));
    NNGP_CUDA(h, cudaMalloc(&s.d_mx, sizeof(double) * d));
    NNGP_CUDA(h, cudaMemcpy(s.d_mn, mn, sizeof(double) * d, cudaMemcpyHostToDevice));
    NNGP_CUDA(h, cudaMemcpy(s.d_mx, mx, sizeof(double) * d, cudaMemcpyHostToDevice));
  }
  h->systems.push_back(s);
  *sys_out = (int)h->systems.size() - 1;
  return 0;
}

static int get_sys(nngp_handle_t h, int sys, const SystemDesc** out) {
  if (sys < 0 || sys >= (int)h->systems.size()) return nngp_fail(h, "bad system handle %d", sys);
  *out = &h->systems[sys];
  return 0;
}

int nngp_rhs_eval(nngp_handle_t h, int sys, int n, const double* d_u, double* d_out, void* stream) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  return rhs_launch(h, *s, n, d_u, d_out, as_stream(stream));
}

int nngp_rhs_eval_host(nngp_handle_t h, int sys, int n, const double* u, double* out) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  const size_t bytes = sizeof(double) * (size_t)n * s->d;
  double* dev = (double*)stage_buf(h, 2 * Carver::pad(bytes));
  if (!dev) return nngp_fail(h, "rhs_eval_host: out of device memory");
  double* dout = (double*)((char*)dev + Carver::pad(bytes));
  NNGP_CUDA(h, cudaMemcpyAsync(dev, u, bytes, cudaMemcpyHostToDevice, h->own_stream));
  if (int rc = rhs_launch(h, *s, n, dev, dout, h->own_stream)) return rc;
  NNGP_CUDA(h, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, h->own_stream));
  NNGP_CUDA(h, cudaStreamSynchronize(h->own_stream));
  return 0;
}

// ---- propagators -----------------------------------------------------------------------
int nngp_rk_batch(nngp_handle_t h, int sys, int method, int h_mode, long long steps, int n_slices,
                  const double* d_t0, const double* d_t1, const double* d_u0, long long ld_u0,
                  double* d_u1, long long ld_u1, void* stream) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  return rk_launch(h, *s, method, h_mode, steps, n_slices, d_t0, d_t1, d_u0, ld_u0, d_u1, ld_u1,
                   as_stream(stream));
}

int nngp_rk_batch_host(nngp_handle_t h, int sys, int method, int h_mode, long long steps,
                       int n_slices, const double* t0, const double* t1, const double* u0,
                       double* u1) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  if (n_slices <= 0) return 0;
  const int d = s->d;
  const size_t nb_t = sizeof(double) * (size_t)n_slices, nb_u = nb_t * d;
  // pinned staging: [t0 | t1 | u0] in, [u1] out
  const size_t in_bytes = 2 * Carver::pad(nb_t) + Carver::pad(nb_u);
  char* pin = (char*)pinned_buf(h, in_bytes + Carver::pad(nb_u));
  char* dev = (char*)stage_buf(h, in_bytes + Carver::pad(nb_u));
  if (!pin || !dev) return nngp_fail(h, "rk_batch_host: out of staging memory");
  Carver cp(pin), cd(dev);
  double* p_t0 = cp.take<double>(n_slices); double* p_t1 = cp.take<double>(n_slices);
  double* p_u0 = cp.take<double>((size_t)n_slices * d); double* p_u1 = cp.take<double>((size_t)n_slices * d);
  double* g_t0 = cd.take<double>(n_slices); double* g_t1 = cd.take<double>(n_slices);
  double* g_u0 = cd.take<double>((size_t)n_slices * d); double* g_u1 = cd.take<double>((size_t)n_slices * d);
  memcpy(p_t0, t0, nb_t);
  memcpy(p_t1, t1, nb_t);
  memcpy(p_u0, u0, nb_u);
  NNGP_CUDA(h, cudaMemcpyAsync(g_t0, p_t0, in_bytes, cudaMemcpyHostToDevice, h->own_stream));
  if (int rc = rk_launch(h, *s, method, h_mode, steps, n_slices, g_t0, g_t1, g_u0, d, g_u1, d, h->own_stream)) return rc;
  NNGP_CUDA(h, cudaMemcpyAsync(p_u1, g_u1, nb_u, cudaMemcpyDeviceToHost, h->own_stream));
  NNGP_CUDA(h, cudaStreamSynchronize(h->own_stream));
  memcpy(u1, p_u1, nb_u);
  return 0;
}

// every step of n_slices solves: d_traj[n_slices][steps+1][d]  (RK.run, RK.py:91-99)
int nngp_rk_full(nngp_handle_t h, int sys, int method, int h_mode, long long steps, int n_slices,
                 const double* d_t0, const double* d_t1, const double* d_u0, long long ld_u0, double* d_traj,
                 void* stream) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  if (n_slices <= 0) return 0;
  if (!d_traj) return nngp_fail(h, "rk_full: d_traj is NULL");
  if (steps < 1) return nngp_fail(h, "steps must be >= 1 (got %lld)", steps);
  // the last state also goes to the last row of each trajectory: use it as the u1 buffer
  double* u1 = d_traj + (long long)steps * s->d;
  return rk_launch(h, *s, method, h_mode, steps, n_slices, d_t0, d_t1, d_u0, ld_u0, u1, (steps + 1) * (long long)s->d,
                   as_stream(stream), d_traj);
}

int nngp_rk_full_host(nngp_handle_t h, int sys, int method, int h_mode, long long steps, double t0, double t1,
                      const double* u0, double* traj) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  if (steps < 1) return nngp_fail(h, "steps must be >= 1 (got %lld)", steps);
  const int d = s->d;
  const size_t nb_u = sizeof(double) * d, nb_tr = sizeof(double) * (size_t)(steps + 1) * d;
  char* dev = (char*)stage_buf(h, 2 * Carver::pad(sizeof(double)) + Carver::pad(nb_u) + Carver::pad(nb_tr));
  if (!dev) return nngp_fail(h, "rk_full_host: out of device memory (%zu bytes)", nb_tr);
  Carver cd(dev);
  double* g_t0 = cd.take<double>(1); double* g_t1 = cd.take<double>(1);
  double* g_u0 = cd.take<double>(d); double* g_tr = cd.take<double>((size_t)(steps + 1) * d);
  NNGP_CUDA(h, cudaMemcpyAsync(g_t0, &t0, sizeof(double), cudaMemcpyHostToDevice, h->own_stream));
  NNGP_CUDA(h, cudaMemcpyAsync(g_t1, &t1, sizeof(double), cudaMemcpyHostToDevice, h->own_stream));
  NNGP_CUDA(h, cudaMemcpyAsync(g_u0, u0, nb_u, cudaMemcpyHostToDevice, h->own_stream));
  if (int rc = nngp_rk_full(h, sys, method, h_mode, steps, 1, g_t0, g_t1, g_u0, d, g_tr, h->own_stream)) return rc;
  NNGP_CUDA(h, cudaMemcpyAsync(traj, g_tr, nb_tr, cudaMemcpyDeviceToHost, h->own_stream));
  NNGP_CUDA(h, cudaStreamSynchronize(h->own_stream));
  return 0;
}

int nngp_get_tableau(int method, int* stages, double* a, double* b, double* c) {
  if (method != 1 && method != 2 && method != 4 && method != 8) return -1;
  rk_host_tableau(method, stages, a, b, c);
  return 0;
}

// ---- dataset ---------------------------------------------------------------------------
int nngp_dataset_reserve(nngp_handle_t h, long long cap_rows, int d) {
  NNGP_USE_DEVICE(h);
  if (cap_rows < 1 || d < 1) return nngp_fail(h, "dataset_reserve: cap_rows=%lld d=%d", cap_rows, d);
  if (h->ds_x && h->ds_d == d && h->ds_cap >= cap_rows) return 0;
  if (h->ds_x && h->ds_d != d && h->ds_rows > 0)
    return nngp_fail(h, "dataset_reserve: dimension change %d -> %d with %lld rows held", h->ds_d, d, h->ds_rows);
  NNGP_CUDA(h, cudaDeviceSynchronize());
  double *nx = nullptr, *ny = nullptr, *nxt = nullptr;
  const size_t bytes = sizeof(double) * (size_t)cap_rows * d;
  NNGP_CUDA(h, cudaMalloc(&nx, bytes));
  NNGP_CUDA(h, cudaMalloc(&ny, bytes));
  NNGP_CUDA(h, cudaMalloc(&nxt, bytes));
  const long long keep = (h->ds_x && h->ds_d == d) ? h->ds_rows : 0;
  double *ox = h->ds_x, *oy = h->ds_y, *oxt = h->ds_xt;
  h->ds_x = nx; h->ds_y = ny; h->ds_xt = nxt;
  h->ds_d = d; h->ds_cap = cap_rows; h->ds_rows = 0;
  if (keep > 0) {
    if (int rc = dataset_append_launch(h, ox, oy, keep, nullptr)) return rc;
    NNGP_CUDA(h, cudaDeviceSynchronize());
  }
  if (ox) cudaFree(ox);
  if (oy) cudaFree(oy);
  if (oxt) cudaFree(oxt);
  return 0;
}

int nngp_dataset_reset(nngp_handle_t h) {
  h->ds_rows = 0;
  return 0;
}

int nngp_dataset_truncate(nngp_handle_t h, long long rows) {
  if (rows < 0 || rows > h->ds_rows) return nngp_fail(h, "dataset_truncate: rows=%lld outside [0,%lld]", rows, h->ds_rows);
  h->ds_rows = rows;
  return 0;
}

long long nngp_dataset_rows(nngp_handle_t h) { return h->ds_rows; }
int nngp_dataset_dim(nngp_handle_t h) { return h->ds_d; }

int nngp_dataset_append(nngp_handle_t h, const double* d_x, const double* d_y, long long rows,
                        void* stream) {
  NNGP_USE_DEVICE(h);
  return dataset_append_launch(h, d_x, d_y, rows, as_stream(stream));
}

int nngp_dataset_append_host(nngp_handle_t h, const double* x, const double* y, long long rows) {
  NNGP_USE_DEVICE(h);
  if (rows <= 0) return 0;
  if (!h->ds_x) return nngp_fail(h, "dataset not reserved (call nngp_dataset_reserve)");
  if (h->ds_rows + rows > h->ds_cap) {
    long long want = h->ds_cap * 2;
    if (want < h->ds_rows + rows) want = h->ds_rows + rows;
    if (int rc = nngp_dataset_reserve(h, want, h->ds_d)) return rc;
  }
  const size_t bytes = sizeof(double) * (size_t)rows * h->ds_d;
  char* dev = (char*)stage_buf(h, 2 * Carver::pad(bytes));
  if (!dev) return nngp_fail(h, "dataset_append_host: out of device memory");
  double* gx = (double*)dev;
  double* gy = (double*)(dev + Carver::pad(bytes));
  NNGP_CUDA(h, cudaMemcpyAsync(gx, x, bytes, cudaMemcpyHostToDevice, h->own_stream));
  NNGP_CUDA(h, cudaMemcpyAsync(gy, y, bytes, cudaMemcpyHostToDevice, h->own_stream));
  if (int rc = dataset_append_launch(h, gx, gy, rows, h->own_stream)) return rc;
  NNGP_CUDA(h, cudaStreamSynchronize(h->own_stream));
  return 0;
}

int nngp_append_iteration(nngp_handle_t h, const double* d_u_cur, const double* d_uF,
                          const double* d_uG_cur, int N, int I, int d, void* stream) {
  NNGP_USE_DEVICE(h);
  return append_iteration_launch(h, d_u_cur, d_uF, d_uG_cur, N, I, d, as_stream(stream));
}

int nngp_rowwise_maxabs_diff(nngp_handle_t h, const double* d_a, const double* d_b, int rows,
                             int d, double* d_err, void* stream) {
  NNGP_USE_DEVICE(h);
  return rowwise_maxabs_launch(h, d_a, d_b, rows, d, d_err, as_stream(stream));
}

// ---- kNN -------------------------------------------------------------------------------
int nngp_knn(nngp_handle_t h, const double* d_q, int nq, int m, long long n_rows,
             long long* d_idx, double* d_dist, void* stream) {
  NNGP_USE_DEVICE(h);
  const long long n = (n_rows > 0) ? n_rows : h->ds_rows;
  void* ws = nngp_workspace(h, knn_workspace_bytes(nq, n, m));
  if (!ws) return nngp_fail(h, "knn: out of device memory for %d x %lld distances", nq, n);
  return knn_launch(h, d_q, nq, m, n_rows, d_idx, d_dist, ws, as_stream(stream));
}

int nngp_knn_host(nngp_handle_t h, const double* q, int nq, int m, long long n_rows,
                  long long* idx, double* dist) {
  NNGP_USE_DEVICE(h);
  if (nq <= 0) return 0;
  const int d = h->ds_d;
  const size_t bq = Carver::pad(sizeof(double) * (size_t)nq * d);
  const size_t bi = Carver::pad(sizeof(long long) * (size_t)nq * m);
  const size_t bd = Carver::pad(sizeof(double) * (size_t)nq * m);
  char* dev = (char*)stage_buf(h, bq + bi + bd);
  if (!dev) return nngp_fail(h, "knn_host: out of device memory");
  double* gq = (double*)dev;
  long long* gi = (long long*)(dev + bq);
  double* gd = (double*)(dev + bq + bi);
  NNGP_CUDA(h, cudaMemcpyAsync(gq, q, sizeof(double) * (size_t)nq * d, cudaMemcpyHostToDevice, h->own_stream));
  if (int rc = nngp_knn(h, gq, nq, m, n_rows, gi, gd, h->own_stream)) return rc;
  NNGP_CUDA(h, cudaMemcpyAsync(idx, gi, sizeof(long long) * (size_t)nq * m, cudaMemcpyDeviceToHost, h->own_stream));
  NNGP_CUDA(h, cudaMemcpyAsync(dist, gd, sizeof(double) * (size_t)nq * m, cudaMemcpyDeviceToHost, h->own_stream));
  NNGP_CUDA(h, cudaStreamSynchronize(h->own_stream));
  return 0;
}

// ---- GP fit / predict ------------------------------------------------------------------
// workspace layout for a fit on nq queries: [knn distances nq*n | fit block (r2 | results | done)]
static int fit_workspace(nngp_handle_t h, int nq, long long n, int m, int R, void** knn_ws, void** fit_ws,
                         int** order = nullptr) {
  const size_t bk = Carver::pad(knn_workspace_bytes(nq, n, m));
  const size_t bf = Carver::pad(gp_fit_ws_bytes(nq, h->ds_d, m, R));
  const size_t bo = Carver::pad(sizeof(int) * (size_t)nq * h->ds_d * NNGP_N_JITTER * R);
  char* ws = (char*)nngp_workspace(h, bk + bf + bo);
  if (!ws) return nngp_fail(h, "fit: out of device memory (workspace %zu bytes)", bk + bf + bo);
  *knn_ws = ws;
  *fit_ws = ws + bk;
  if (order) *order = (int*)(ws + bk + bf);
  return 0;
}

int nngp_fit_predict(nngp_handle_t h, const double* d_q, const long long* d_idx,
                     const double* d_dist, int nq, int m, int n_restarts,
                     const signed char* d_starts, double fatol, double xatol, double* d_pred,
                     double* d_theta_opt, double* d_jitter_opt, double* d_fval_opt, int* d_nfev,
                     double* d_fvals, double* d_thetas, void* stream) {
  NNGP_USE_DEVICE(h);
  (void)d_q;
  if (n_restarts < 1) return nngp_fail(h, "fit: n_restarts=%d < 1", n_restarts);
  void *kws, *fws;
  int* order;
  if (int rc = fit_workspace(h, nq, h->ds_rows, m, n_restarts, &kws, &fws, &order)) return rc;
  cudaStream_t st = as_stream(stream);
  if (int rc = ensure_done_zero(h, fws, nq, h->ds_d, m, n_restarts, st)) return rc;
  if (int rc = gp_prep_launch(h, d_idx, nq, m, (double*)fws, st)) return rc;
  if (int rc = gp_order_launch(h, d_starts, nq, h->ds_d * NNGP_N_JITTER * n_restarts, 1, order, st)) return rc;
  return gp_fit_predict_launch(h, d_idx, d_dist, fws, next_queue(h, st), order, nq, m, n_restarts, d_starts, fatol,
                               xatol, d_pred, nullptr, h->ds_d, d_theta_opt, d_jitter_opt, d_fval_opt, d_nfev,
                               d_fvals, d_thetas, st);
}

int nngp_gp_nll(nngp_handle_t h, const long long* d_idx, int nq, int m, int nt,
                const double* d_theta, const double* d_jitter10, double* d_nll, void* stream) {
  NNGP_USE_DEVICE(h);
  void *kws, *fws;
  if (int rc = fit_workspace(h, nq, h->ds_rows, m, 1, &kws, &fws)) return rc;
  cudaStream_t st = as_stream(stream);
  if (int rc = gp_prep_launch(h, d_idx, nq, m, (double*)fws, st)) return rc;
  return gp_nll_launch(h, d_idx, (const double*)fws, nq, m, nt, d_theta, d_jitter10, d_nll, st);
}

int nngp_gp_mean(nngp_handle_t h, const double* d_q, const long long* d_idx, const double* d_dist,
                 int nq, int m, const double* d_theta, const double* d_jitter, double* d_pred,
                 void* stream) {
  NNGP_USE_DEVICE(h);
  (void)d_q;
  void *kws, *fws;
  if (int rc = fit_workspace(h, nq, h->ds_rows, m, 1, &kws, &fws)) return rc;
  cudaStream_t st = as_stream(stream);
  if (int rc = gp_prep_launch(h, d_idx, nq, m, (double*)fws, st)) return rc;
  return gp_mean_launch(h, d_idx, d_dist, (const double*)fws, nq, m, d_theta, d_jitter, d_pred, st);
}

int nngp_predict_host(nngp_handle_t h, const double* q, int nq, int m, long long n_rows,
                      int n_restarts, const signed char* starts, double fatol, double xatol,
                      double* pred, long long* idx, double* theta_opt, double* jitter_opt,
                      double* fval_opt, int* nfev, double* fvals, double* thetas) {
  return nngp_predict_host_block(h, q, nq, m, n_rows, n_restarts, starts, fatol, xatol, 0, -1, pred, idx, theta_opt,
                                 jitter_opt, fval_opt, nfev, fvals, thetas);
}

int nngp_predict_host_block(nngp_handle_t h, const double* q, int nq, int m, long long n_rows,
                            int n_restarts, const signed char* starts, double fatol, double xatol, int j0, int dl,
                            double* pred, long long* idx, double* theta_opt, double* jitter_opt,
                            double* fval_opt, int* nfev, double* fvals, double* thetas) {
  NNGP_USE_DEVICE(h);
  if (nq <= 0) return 0;
  if (!h->ds_x) return nngp_fail(h, "predict: dataset is empty");
  const int d = h->ds_d, R = n_restarts;
  if (dl < 0) {
    j0 = 0;
    dl = d;
  }
  if (j0 < 0 || dl < 1 || j0 + dl > d) return nngp_fail(h, "predict: dimension block [%d,%d) outside [0,%d)", j0, j0 + dl, d);
  if (dl != d && nq != 1) return nngp_fail(h, "predict: a dimension block needs a single query (got %d)", nq);
  const long long n = (n_rows > 0) ? n_rows : h->ds_rows;
  const size_t nqd = (size_t)nq * d, ntask = nqd * NNGP_N_JITTER * R;
  size_t total = 0;
  auto add = [&](size_t bytes) { size_t o = total; total += Carver::pad(bytes); return o; };
  const size_t o_q = add(sizeof(double) * nqd), o_st = add(ntask * 2), o_idx = add(sizeof(long long) * nq * m),
               o_dist = add(sizeof(double) * nq * m), o_pred = add(sizeof(double) * nqd),
               o_th = add(sizeof(double) * nqd * 2), o_jit = add(sizeof(double) * nqd),
               o_fv = add(sizeof(double) * nqd), o_nfev = add(sizeof(int) * ntask),
               o_fvals = add(sizeof(double) * ntask), o_thetas = add(sizeof(double) * ntask * 2);
  char* dev = (char*)stage_buf(h, total);
  if (!dev) return nngp_fail(h, "predict_host: out of device memory (%zu bytes)", total);
  cudaStream_t st = h->own_stream;
  double* gq = (double*)(dev + o_q);
  signed char* gst = (signed char*)(dev + o_st);
  long long* gidx = (long long*)(dev + o_idx);
  double* gdist = (double*)(dev + o_dist);
  if (R < 1) return nngp_fail(h, "predict: n_restarts=%d < 1", R);
  // query and starts are adjacent in the device block: one copy from the pinned staging buffer instead of two from
  // pageable memory (a single predict of the sweep is latency-bound: every call on this path counts)
  const size_t in_bytes = (o_st - o_q) + ntask * 2;
  if (in_bytes <= (1u << 20)) {
    char* pin = (char*)pinned_buf(h, in_bytes);
    if (!pin) return nngp_fail(h, "predict_host: out of pinned memory");
    memcpy(pin, q, sizeof(double) * nqd);
    memcpy(pin + (o_st - o_q), starts, ntask * 2);
    NNGP_CUDA(h, cudaMemcpyAsync(gq, pin, in_bytes, cudaMemcpyHostToDevice, st));
  } else {
    NNGP_CUDA(h, cudaMemcpyAsync(gq, q, sizeof(double) * nqd, cudaMemcpyHostToDevice, st));
    NNGP_CUDA(h, cudaMemcpyAsync(gst, starts, ntask * 2, cudaMemcpyHostToDevice, st));
  }
  void *kws, *fws;
  int* order;
  if (int rc = fit_workspace(h, nq, n, m, R, &kws, &fws, &order)) return rc;
  if (int rc = ensure_done_zero(h, fws, nq, d, m, R, st)) return rc;
  if (nq == 1 && n >= m && knn_prep_fused_ok(h, n, m)) {  // the sweep's shape: scan + top-m + neighbour matrix in one launch
    if (int rc = knn_prep_fused_launch(h, gq, m, n, gidx, gdist, (double*)fws, kws, h->d_ticket, st)) return rc;
  } else {
    if (int rc = knn_launch(h, gq, nq, m, n, gidx, gdist, kws, st)) return rc;
    if (int rc = gp_prep_launch(h, gidx, nq, m, (double*)fws, st)) return rc;
  }
  if (dl == d) {
    if (int rc = gp_order_launch(h, gst, nq, d * NNGP_N_JITTER * R, 1, order, st)) return rc;
  } else {  // the searches of the block are a contiguous range of the (single) query's task list
    if (int rc = gp_order_launch(h, gst + (size_t)j0 * NNGP_N_JITTER * R * 2, 1, dl * NNGP_N_JITTER * R, 0, order, st)) return rc;
  }
  if (int rc = gp_fit_predict_launch(h, gidx, gdist, fws, next_queue(h, st), order, nq, m, R, gst, fatol, xatol,
                                     (double*)(dev + o_pred), nullptr, d,
                                     theta_opt ? (double*)(dev + o_th) : nullptr,
                                     jitter_opt ? (double*)(dev + o_jit) : nullptr,
                                     fval_opt ? (double*)(dev + o_fv) : nullptr,
                                     nfev ? (int*)(dev + o_nfev) : nullptr,
                                     fvals ? (double*)(dev + o_fvals) : nullptr,
                                     thetas ? (double*)(dev + o_thetas) : nullptr, st, j0, dl))
    return rc;
  // cudaMemcpyDefault: `pred` may be a host buffer or (unified addressing) a device buffer of the caller -- the ranks of a
  // dimension-sharded host predict all-gather their blocks on the device and read the result back once
  NNGP_CUDA(h, cudaMemcpyAsync(pred, dev + o_pred, sizeof(double) * nqd, cudaMemcpyDefault, st));
  if (idx) NNGP_CUDA(h, cudaMemcpyAsync(idx, gidx, sizeof(long long) * nq * m, cudaMemcpyDeviceToHost, st));
  if (theta_opt) NNGP_CUDA(h, cudaMemcpyAsync(theta_opt, dev + o_th, sizeof(double) * nqd * 2, cudaMemcpyDeviceToHost, st));
  if (jitter_opt) NNGP_CUDA(h, cudaMemcpyAsync(jitter_opt, dev + o_jit, sizeof(double) * nqd, cudaMemcpyDeviceToHost, st));
  if (fval_opt) NNGP_CUDA(h, cudaMemcpyAsync(fval_opt, dev + o_fv, sizeof(double) * nqd, cudaMemcpyDeviceToHost, st));
  if (nfev) NNGP_CUDA(h, cudaMemcpyAsync(nfev, dev + o_nfev, sizeof(int) * ntask, cudaMemcpyDeviceToHost, st));
  if (fvals) NNGP_CUDA(h, cudaMemcpyAsync(fvals, dev + o_fvals, sizeof(double) * ntask, cudaMemcpyDeviceToHost, st));
  if (thetas) NNGP_CUDA(h, cudaMemcpyAsync(thetas, dev + o_thetas, sizeof(double) * ntask * 2, cudaMemcpyDeviceToHost, st));
  NNGP_CUDA(h, cudaStreamSynchronize(st));
  return 0;
}

// Everything of a slice that precedes the fits: the coarse step uG_next[i+1] = G(u_next[i]) on the auxiliary stream,
// beside it ONE launch for distance scan + top-m + neighbour matrix (general shapes: three launches); the caller's
// stream joins the auxiliary one before the fit, whose epilogue adds uG_next[i+1] to the prediction.
static int sweep_slice_prologue(nngp_handle_t h, const SystemDesc& s, int method_g, int h_mode, long long steps_g,
                                const double* d_ti, double* ui, double* gn, int d, int m, long long n, long long* idx,
                                double* dist, void* kws, void* fws, cudaStream_t st) {
  const bool fork = getenv("NNGP_SWEEP_NO_FORK") == nullptr;
  cudaStream_t sg = fork ? h->aux_stream : st;
  if (fork) {
    NNGP_CUDA(h, cudaEventRecord(h->ev_fork, st));
    NNGP_CUDA(h, cudaStreamWaitEvent(sg, h->ev_fork, 0));
  }
  if (int rc = rk_launch(h, s, method_g, h_mode, steps_g, 1, d_ti, d_ti + 1, ui, d, gn, d, sg)) return rc;
  if (fork) NNGP_CUDA(h, cudaEventRecord(h->ev_join, sg));
  if (knn_prep_fused_ok(h, n, m) && getenv("NNGP_SWEEP_NO_FUSED_PROLOGUE") == nullptr) {
    if (int rc = knn_prep_fused_launch(h, ui, m, n, idx, dist, (double*)fws, kws, h->d_ticket, st)) return rc;
  } else {
    if (int rc = knn_launch(h, ui, 1, m, n, idx, dist, kws, st)) return rc;
    if (int rc = gp_prep_launch(h, idx, 1, m, (double*)fws, st)) return rc;
  }
  if (fork) NNGP_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
  return 0;
}

// ---- fused on-device sweep (parareal.py:359-382) --------------------------------------
int nngp_sweep(nngp_handle_t h, int sys, int method_g, int h_mode, long long steps_g,
               const double* d_t, int N, int I, int m, int n_restarts,
               const signed char* d_starts, double fatol, double xatol, double* d_u_next,
               double* d_uG_next, int d, void* stream) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  if (d != s->d || d != h->ds_d) return nngp_fail(h, "sweep: d=%d, system d=%d, dataset d=%d", d, s->d, h->ds_d);
  if (I < 0 || I > N) return nngp_fail(h, "sweep: I=%d outside [0,%d]", I, N);
  cudaStream_t st = as_stream(stream);
  if (n_restarts < 1) return nngp_fail(h, "sweep: n_restarts=%d < 1", n_restarts);
  const long long n = h->ds_rows;
  const size_t bk = Carver::pad(knn_workspace_bytes(1, n, m));
  const size_t bf = Carver::pad(gp_fit_ws_bytes(1, d, m, n_restarts));
  const size_t bi = Carver::pad(sizeof(long long) * m), bd = Carver::pad(sizeof(double) * m);
  const int seg_len = d * NNGP_N_JITTER * n_restarts;
  const size_t bo = Carver::pad(sizeof(int) * (size_t)seg_len * (size_t)(N - I));
  char* ws = (char*)nngp_workspace(h, bk + bf + bi + bd + bo);
  if (!ws) return nngp_fail(h, "sweep: out of device memory");
  void* kws = ws;
  void* fws = ws + bk;
  long long* idx = (long long*)(ws + bk + bf);
  double* dist = (double*)(ws + bk + bf + bi);
  int* order = (int*)(ws + bk + bf + bi + bd);
  if (int rc = ensure_done_zero(h, fws, 1, d, m, n_restarts, st)) return rc;
  // queue order of every predict of the sweep in one launch (the starts are all known up front)
  if (int rc = gp_order_launch(h, d_starts, N - I, seg_len, 0, order, st)) return rc;
  const size_t per_predict = (size_t)d * NNGP_N_JITTER * n_restarts * 2;
  for (int i = I; i < N; i++) {
    double* ui = d_u_next + (long long)i * d;
    double* un = d_u_next + (long long)(i + 1) * d;
    double* gn = d_uG_next + (long long)(i + 1) * d;
    if (int rc = sweep_slice_prologue(h, *s, method_g, h_mode, steps_g, d_t + i, ui, gn, d, m, n, idx, dist, kws, fws, st)) return rc;
    if (int rc = gp_fit_predict_launch(h, idx, dist, fws, next_queue(h, st), order + (size_t)(i - I) * seg_len, 1, m,
                                       n_restarts, d_starts + (size_t)(i - I) * per_predict, fatol, xatol, un, gn, d,
                                       nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st))
      return rc;
  }
  return 0;
}

// One rank's share of the sweep for slices i_first .. i_first+i_count-1 (I = first slice of the whole
// sweep: d_starts holds the starts of predicts I, I+1, ... in order): G and the neighbour search are done
// in full (replicated, cheap), the d*9*R searches only for the output dimensions [j0, j0+dl); u_next[i+1]
// is written for those dimensions only -- the caller all-gathers the row before the next slice.
int nngp_sweep_shard(nngp_handle_t h, int sys, int method_g, int h_mode, long long steps_g,
                     const double* d_t, int N, int I, int i_first, int i_count, int m, int n_restarts,
                     const signed char* d_starts, double fatol, double xatol, double* d_u_next,
                     double* d_uG_next, int d, int j0, int dl, void* stream) {
  NNGP_USE_DEVICE(h);
  const SystemDesc* s;
  if (int rc = get_sys(h, sys, &s)) return rc;
  if (d != s->d || d != h->ds_d) return nngp_fail(h, "sweep: d=%d, system d=%d, dataset d=%d", d, s->d, h->ds_d);
  if (I < 0 || I > N || i_first < I || i_count < 0 || i_first + i_count > N)
    return nngp_fail(h, "sweep_shard: slices [%d,%d) outside [%d,%d)", i_first, i_first + i_count, I, N);
  if (j0 < 0 || dl < 1 || j0 + dl > d) return nngp_fail(h, "sweep_shard: dimensions [%d,%d) outside [0,%d)", j0, j0 + dl, d);
  if (n_restarts < 1) return nngp_fail(h, "sweep: n_restarts=%d < 1", n_restarts);
  cudaStream_t st = as_stream(stream);
  const long long n = h->ds_rows;
  const size_t bk = Carver::pad(knn_workspace_bytes(1, n, m));
  const size_t bf = Carver::pad(gp_fit_ws_bytes(1, d, m, n_restarts));
  const size_t bi = Carver::pad(sizeof(long long) * m), bd = Carver::pad(sizeof(double) * m);
  const int nruns = NNGP_N_JITTER * n_restarts;
  const int seg_len = dl * nruns;
  const size_t bo = Carver::pad(sizeof(int) * (size_t)seg_len);
  char* ws = (char*)nngp_workspace(h, bk + bf + bi + bd + bo);
  if (!ws) return nngp_fail(h, "sweep: out of device memory");
  void* kws = ws;
  void* fws = ws + bk;
  long long* idx = (long long*)(ws + bk + bf);
  double* dist = (double*)(ws + bk + bf + bi);
  int* order = (int*)(ws + bk + bf + bi + bd);
  if (int rc = ensure_done_zero(h, fws, 1, d, m, n_restarts, st)) return rc;
  const size_t per_predict = (size_t)d * nruns * 2;
  for (int i = i_first; i < i_first + i_count; i++) {
    double* ui = d_u_next + (long long)i * d;
    double* un = d_u_next + (long long)(i + 1) * d;
    double* gn = d_uG_next + (long long)(i + 1) * d;
    const signed char* starts_i = d_starts + (size_t)(i - I) * per_predict;
    if (int rc = sweep_slice_prologue(h, *s, method_g, h_mode, steps_g, d_t + i, ui, gn, d, m, n, idx, dist, kws, fws, st)) return rc;
    if (int rc = gp_order_launch(h, starts_i + (size_t)j0 * nruns * 2, 1, seg_len, 0, order, st)) return rc;
    if (int rc = gp_fit_predict_launch(h, idx, dist, fws, next_queue(h, st), order, 1, m, n_restarts, starts_i, fatol,
                                       xatol, un, gn, d, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st, j0, dl))
      return rc;
  }
  return 0;
}

int nngp_selftest_math(nngp_handle_t h, const double* d_x, int n, double* d_exp_neg, double* d_rcp,
                       double* d_exp10, void* stream) {
  NNGP_USE_DEVICE(h);
  return selftest_math_launch(h, d_x, n, d_exp_neg, d_rcp, d_exp10, as_stream(stream));
}

// ---- roofline micro-benchmarks ---------------------------------------------------------
__global__ void fp64_fma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int nngp_bench_fp64(nngp_handle_t h, int iters, double* tflops_out) {
  NNGP_USE_DEVICE(h);
  const int blocks = 148 * 8, threads = 256;
  double* out = (double*)nngp_workspace(h, sizeof(double) * blocks * threads);
  if (!out) return nngp_fail(h, "bench_fp64: out of memory");
  cudaEvent_t e0, e1;
  NNGP_CUDA(h, cudaEventCreate(&e0));
  NNGP_CUDA(h, cudaEventCreate(&e1));
  fp64_fma_kernel<<<blocks, threads, 0, h->own_stream>>>(out, iters, 0.999999, 1e-7);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    NNGP_CUDA(h, cudaEventRecord(e0, h->own_stream));
    fp64_fma_kernel<<<blocks, threads, 0, h->own_stream>>>(out, iters, 0.999999, 1e-7);
    NNGP_CUDA(h, cudaEventRecord(e1, h->own_stream));
    NNGP_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  h->launches += 6;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return 0;
}

__global__ void copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

int nngp_bench_copy(nngp_handle_t h, long long bytes, double* gbs_out) {
  NNGP_USE_DEVICE(h);
  const size_t n = (size_t)bytes / sizeof(double2);
  char* ws = (char*)nngp_workspace(h, 2 * n * sizeof(double2));
  if (!ws) return nngp_fail(h, "bench_copy: out of memory");
  double2* a = (double2*)ws;
  double2* b = a + n;
  NNGP_CUDA(h, cudaMemsetAsync(a, 0, n * sizeof(double2), h->own_stream));
  cudaEvent_t e0, e1;
  NNGP_CUDA(h, cudaEventCreate(&e0));
  NNGP_CUDA(h, cudaEventCreate(&e1));
  copy_kernel<<<148 * 16, 512, 0, h->own_stream>>>(a, b, n);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    NNGP_CUDA(h, cudaEventRecord(e0, h->own_stream));
    copy_kernel<<<148 * 16, 512, 0, h->own_stream>>>(a, b, n);
    NNGP_CUDA(h, cudaEventRecord(e1, h->own_stream));
    NNGP_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gbs = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
    if (gbs > best) best = gbs;
  }
  h->launches += 6;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *gbs_out = best;
  return 0;
}

}  // extern "C"
