// Runtime plumbing of libwavelet_sm100a.so: errors, device binding, scratch
// arenas, twiddle tables, and the small host-side pieces of the C ABI.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace wtb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_error("CUDA error %s (%d) at %s:%d: %s", cudaGetErrorName(e), (int)e, file, line, what);
  return WTB_ECUDA;
}

static std::atomic<uint64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }

// ---- device binding -------------------------------------------------------------
static std::mutex g_mu;
static int g_device = -1;  // process-wide (one process per GPU)
static int g_sms = 0;

static int bind_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device visible (%s); libwavelet_sm100a has no CPU fallback",
              e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    cudaGetLastError();
    return WTB_ENODEVICE;
  }
  WTB_REQUIRE(device >= 0 && device < count, WTB_EINVAL, "device %d out of range [0,%d)", device, count);
  cudaDeviceProp p;
  WTB_CUDA(cudaGetDeviceProperties(&p, device));
  WTB_REQUIRE(p.major == 10, WTB_ENODEVICE,
              "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
  WTB_CUDA(cudaSetDevice(device));
  g_device = device;
  g_sms = p.multiProcessorCount;
  return WTB_OK;
}

int ensure_device() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device >= 0) {
    // a new host thread starts on device 0: re-select the bound device
    int cur = -1;
    WTB_CUDA(cudaGetDevice(&cur));
    if (cur != g_device) WTB_CUDA(cudaSetDevice(g_device));
    return WTB_OK;
  }
  int dev = 0;
  if (const char *env = getenv("WTB_DEVICE")) dev = atoi(env);
  return bind_device(dev);
}

int sm_count() { return g_sms > 0 ? g_sms : 148; }

// ---- arenas ------------------------------------------------------------------------
static std::mutex g_arena_mu;
static std::vector<Arena *> g_arenas;  // every thread's arenas, for shutdown

static Arena *new_arena() {
  Arena *a = new Arena();
  std::lock_guard<std::mutex> lk(g_arena_mu);
  g_arenas.push_back(a);
  return a;
}

static int reserve(Arena *a, size_t bytes, void **out) {
  if (bytes == 0) bytes = 256;
  if (a->cap < bytes || a->device != g_device) {
    if (a->ptr) {
      WTB_CUDA(cudaDeviceSynchronize());
      WTB_CUDA(cudaFree(a->ptr));
      a->ptr = nullptr;
      a->cap = 0;
    }
    size_t want = bytes + bytes / 8;
    want = (want + (1u << 20) - 1) & ~size_t((1u << 20) - 1);
    cudaError_t e = cudaMalloc(&a->ptr, want);
    if (e != cudaSuccess) {
      a->ptr = nullptr;
      return cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
    }
    a->cap = want;
    a->device = g_device;
  }
  *out = a->ptr;
  return WTB_OK;
}

int arena_reserve(size_t bytes, void **out) {
  static thread_local Arena *a = new_arena();
  return reserve(a, bytes, out);
}
int staging_reserve(size_t bytes, void **out) {
  static thread_local Arena *a = new_arena();
  return reserve(a, bytes, out);
}
int params_reserve(size_t bytes, void **out) {
  static thread_local Arena *a = new_arena();
  return reserve(a, bytes, out);
}

// ---- twiddles ----------------------------------------------------------------------
template <typename T> struct TwCache {
  std::mutex mu;
  std::map<std::pair<int, int>, void *> tab;
};
template <typename T> static TwCache<T> &tw_cache() {
  static TwCache<T> c;
  return c;
}

template <typename T> int twiddles(int N, const cplx<T> **out) {
  TwCache<T> &c = tw_cache<T>();
  std::lock_guard<std::mutex> lk(c.mu);
  auto key = std::make_pair(g_device, N);
  auto it = c.tab.find(key);
  if (it == c.tab.end()) {
    std::vector<cplx<T>> h(N);
    for (int k = 0; k < N; ++k) {
      // exact octant symmetry is not needed; long double keeps the table
      // correctly rounded in T
      long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)N;
      h[k] = mk<T>((T)cosl(a), (T)sinl(a));
    }
    void *d = nullptr;
    WTB_CUDA(cudaMalloc(&d, sizeof(cplx<T>) * N));
    WTB_CUDA(cudaMemcpy(d, h.data(), sizeof(cplx<T>) * N, cudaMemcpyHostToDevice));
    it = c.tab.emplace(key, d).first;
  }
  *out = (const cplx<T> *)it->second;
  return WTB_OK;
}
template int twiddles<float>(int, const cplx<float> **);
template int twiddles<double>(int, const cplx<double> **);

void wct_fast_release();  // wct_fast.cu: the radix-16 twiddle tables

static void free_tables() {
  {
    TwCache<float> &c = tw_cache<float>();
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto &kv : c.tab) cudaFree(kv.second);
    c.tab.clear();
  }
  {
    TwCache<double> &c = tw_cache<double>();
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto &kv : c.tab) cudaFree(kv.second);
    c.tab.clear();
  }
}

// ---- axes --------------------------------------------------------------------------
int resolve_axes(int n0, double dt, double dj, double s0, int J, double f0, Axes *ax) {
  Mother m;
  m.param = f0;
  return resolve_axes(n0, dt, dj, s0, J, m, ax);
}

int resolve_axes(int n0, double dt, double dj, double s0, int J, const Mother &m, Axes *ax) {
  WTB_REQUIRE(n0 > 0 && dt > 0 && dj > 0, WTB_EINVAL, "n0, dt and dj must be positive");
  WTB_REQUIRE(m.kind == WTB_MORLET || m.kind == WTB_PAUL || m.kind == WTB_DOG, WTB_EINVAL, "unknown mother wavelet %d", m.kind);
  WTB_REQUIRE(m.kind == WTB_MORLET ? m.param > 0 : (m.param >= 1 && m.param <= 40 && m.param == std::floor(m.param)),
              WTB_EINVAL, "mother wavelet parameter %g out of range", m.param);
  const double fl = mother_flambda(m);
  if (s0 == -1) s0 = 2 * dt / fl;
  WTB_REQUIRE(s0 > 0, WTB_EINVAL, "s0 must be positive (or -1)");
  if (J == -1) J = (int)std::nearbyint(std::log2(n0 * dt / s0) / dj);
  WTB_REQUIRE(J >= 0 && J < 4096, WTB_EINVAL, "J=%d out of range", J);
  ax->J = J;
  ax->scales.resize(J + 1);
  ax->freqs.resize(J + 1);
  for (int j = 0; j <= J; ++j) {
    ax->scales[j] = s0 * std::pow(2.0, j * dj);
    ax->freqs[j] = 1.0 / (fl * ax->scales[j]);
  }
  return WTB_OK;
}

}  // namespace wtb

using namespace wtb;

extern "C" {

int wtb_version(void) { return 100; }

uint64_t wtb_kernel_launches(void) { return wtb::launches(); }

int wtb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int wtb_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  return bind_device(device);
}

void wtb_shutdown(void) {
  if (g_device < 0) return;
  cudaDeviceSynchronize();
  {
    std::lock_guard<std::mutex> lk(g_arena_mu);
    for (Arena *a : g_arenas) {
      if (a->ptr) cudaFree(a->ptr);
      a->ptr = nullptr;
      a->cap = 0;
    }
  }
  free_tables();
  wct_fast_release();
}

const char *wtb_last_error(void) { return g_err; }

int wtb_cwt_axes_mother(int n0, double dt, double dj, double s0, int J, int mother, double param, int *J_out,
                        double *scales, double *freqs, double *coi) {
  Axes ax;
  Mother m;
  m.kind = mother;
  m.param = param;
  WTB_TRY(resolve_axes(n0, dt, dj, s0, J, m, &ax));
  if (J_out) *J_out = ax.J;
  for (int j = 0; j <= ax.J; ++j) {
    if (scales) scales[j] = ax.scales[j];
    if (freqs) freqs[j] = ax.freqs[j];
  }
  if (coi) {
    // pycwt.cwt: flambda * coi() * dt * (n0/2 - |t - (n0-1)/2|), coi() = 1/sqrt(2) for Morlet
    const double c = mother_flambda(m) * mother_coi(m) * dt;
    for (int t = 0; t < n0; ++t) coi[t] = c * (n0 / 2.0 - std::fabs(t - (n0 - 1) / 2.0));
  }
  return WTB_OK;
}

int wtb_cwt_axes(int n0, double dt, double dj, double s0, int J, double f0, int *J_out,
                 double *scales, double *freqs, double *coi) {
  return wtb_cwt_axes_mother(n0, dt, dj, s0, J, WTB_MORLET, f0, J_out, scales, freqs, coi);
}

int wtb_wct_mc_geometry(double dt, double dj, double s0, int J, double f0, int *nsurr, int *maxscale) {
  WTB_REQUIRE(J >= 0, WTB_EINVAL, "wct_significance needs a resolved J >= 0");
  const double fl = morlet_flambda(f0);
  if (s0 == -1) s0 = 2 * dt / fl;
  const double ms = s0 * std::pow(2.0, J * dj) / dt;
  const int N = (int)std::ceil(ms * 6);
  WTB_REQUIRE(N > 1, WTB_EINVAL, "degenerate surrogate length %d", N);
  if (nsurr) *nsurr = N;
  if (maxscale) {
    // largest s with any t such that period[s] <= coi[t]; max coi is at the centre
    const double c = fl / std::sqrt(2.0) * dt;
    double coimax = 0;
    for (int t = 0; t < N; ++t) coimax = std::fmax(coimax, c * (N / 2.0 - std::fabs(t - (N - 1) / 2.0)));
    int m = 0;
    for (int j = 0; j <= J; ++j) {
      const double freq = 1.0 / (fl * (s0 * std::pow(2.0, j * dj)));
      const double period = 1.0 / freq;  // same rounding path as pycwt
      if (period <= coimax) m = j;
    }
    *maxscale = m;
  }
  return WTB_OK;
}

int wtb_wct_sig_from_hist(const uint64_t *hist, int S, int maxscale, double level,
                          const uint8_t *row_has_points, double *sig95) {
  WTB_REQUIRE(hist && sig95 && S > 0, WTB_EINVAL, "null argument");
  WTB_REQUIRE(maxscale >= 0 && maxscale <= S, WTB_EINVAL, "maxscale out of range");
  const int nb = WTB_NBINS;
  for (int s = 0; s < S; ++s) sig95[s] = (row_has_points && row_has_points[s]) ? NAN : 0.0;
  std::vector<double> P, Y;
  for (int s = 0; s < maxscale; ++s) {
    P.clear();
    Y.clear();
    double cum = 0;
    for (int b = 0; b < nb; ++b) {
      const uint64_t c = hist[(size_t)s * nb + b];
      if (c == 0) continue;
      cum += (double)c;
      P.push_back(cum);
      Y.push_back((b + 0.5) / nb);
    }
    if (P.empty()) { sig95[s] = NAN; continue; }
    const double tot = P.back();
    for (double &p : P) p = (p - 0.5) / tot;
    // np.interp(level, P, Y): clamps outside, linear inside
    double v;
    if (level <= P.front()) v = Y.front();
    else if (level >= P.back()) v = Y.back();
    else {
      size_t i = 1;
      while (P[i] < level) ++i;
      // numpy picks the interval [i-1, i] with P[i-1] <= level < P[i] (ties -> later)
      while (i + 1 < P.size() && P[i] <= level) ++i;
      const double slope = (Y[i] - Y[i - 1]) / (P[i] - P[i - 1]);
      v = slope * (level - P[i - 1]) + Y[i - 1];
    }
    sig95[s] = v;
  }
  return WTB_OK;
}

int wtb_dwt_max_level(int n, int L) {
  if (L < 2 || n < L - 1) return 0;
  int lev = (int)std::floor(std::log2((double)n / (L - 1.0)));
  return lev < 0 ? 0 : lev;
}

int wtb_dwt_coeff_lens(int n, int L, int level, int *lens) {
  WTB_REQUIRE(n > 0 && L >= 2 && level >= 0 && lens, WTB_EINVAL, "bad dwt_coeff_lens arguments");
  int cur = n;
  for (int l = 0; l < level; ++l) {
    cur = (cur + L - 1) / 2;
    lens[level - l] = cur;  // cD_{l+1}
  }
  lens[0] = cur;  // cA_level (== n when level == 0)
  return WTB_OK;
}

int wtb_waverec_len(const int *lens, int level, int L) {
  if (!lens || level < 0) return WTB_EINVAL;
  int a = lens[0];
  for (int l = 1; l <= level; ++l) {
    const int d = lens[l];
    if (a == d + 1) a -= 1;
    if (a != d) { set_error("waverec: coefficient lengths %d and %d do not match", a, d); return WTB_EINVAL; }
    a = 2 * d - L + 2;
    if (a <= 0) { set_error("waverec: level too short for the filter"); return WTB_EINVAL; }
  }
  return a;
}

}  // extern "C"
