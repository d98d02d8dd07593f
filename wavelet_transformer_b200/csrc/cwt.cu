// Batched CWT (pycwt.cwt, src/cwt.py:110-114 of the reference), Morlet unless stated:
//   X = fft(x, nfft);  W[s,:] = ifft(X * sqrt(2*pi*s/dt) * pi^-1/4 * exp(-(s*w-f0)^2/2));
//   power = |W|^2, truncated to the first n0 samples.
// Generic kernels (any pow2 nfft, float/double) live here; the FP32 fast path
// for the headline shape is in cwt_fast.cu and is tried first.
#include "spectral.cuh"

namespace wtb {

// implemented in cwt_fast.cu; returns 1 when the shape is not covered
int cwt_fast_try(const float *d_x, int64_t batch, int n0, int nfft, double dt, const Axes &ax,
                 double f0, int flags, float *d_power, cudaStream_t st);

// implemented in cwt_fast.cu (two / four warps per row for nfft = 2048 / 4096, spectra in group layout); 1 = not covered
int cwt_dif_try(const float2 *d_layout, int64_t stride, int64_t batch, int n0, int nfft, double dt, const Axes &ax,
                double f0, int flags, float *d_power, cudaStream_t st);

// implemented in cwt_fast.cu: the shape runs those kernels, which read the spectra in group layout
bool cwt_dif_covers(int64_t batch, int n0, int nfft, int S, double f0);

// implemented in wct_fast.cu (register-FFT rows for nfft = 4096); returns 1 when not covered
int cwt_rows_4096_try(const float2 *d_xhat, int64_t batch, int n0, int N, double dt, const Axes &ax, double f0,
                      int flags, float *d_power, float2 *d_coef, cudaStream_t st);

// implemented in wct_fast.cu: forward FFTs with the radix-16 register kernel (FP32, N = 4096); 1 = not covered
int fwd_fft_4096_try(const float *d_y, int64_t nseries, int n0, int N, float2 *d_xhat, cudaStream_t st);

// Forward FFT (N = 2048, FP32) for k_cwt_pair_2048: the positive half of the spectrum leaves in the layout
// that kernel's warps load with one 16-byte access per packed pair -- per series 8 KB at the start of its
// xhat row: float4 [parity h][m < 8][lane] = (Re X^[k0], Re X^[k1], Im X^[k0], Im X^[k1]),
// k0 = 2 (lane + 64 m) + h, k1 = k0 + 64.
__global__ void k_fwd_fft_pair2048(const float *__restrict__ x, int n0, FftPlan<float> plan, float2 *__restrict__ xhat) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *a = reinterpret_cast<float2 *>(smem_raw);
  float2 *b = a + plan.M;
  const int N = plan.n;
  const int64_t row = blockIdx.x;
  const float *xr = x + row * n0;
  for (int t = threadIdx.x; t < N; t += blockDim.x) a[t] = make_float2(t < n0 ? xr[t] : 0.0f, 0.0f);
  __syncthreads();
  const float2 *r = plan_fft<float, -1>(a, b, plan);
  float4 *o = reinterpret_cast<float4 *>(xhat + row * N);
  for (int i = threadIdx.x; i < 2 * 8 * 32; i += blockDim.x) {
    const int lane = i & 31, m = (i >> 5) & 7, h = i >> 8;
    const float2 v0 = r[2 * (lane + 64 * m) + h], v1 = r[2 * (lane + 64 * m) + h + 64];
    o[i] = make_float4(v0.x, v1.x, v0.y, v1.y);
  }
}

// The same layout for N = 4096 (four warps per row: k0 = 4 (lane + 64 m) + r, k1 = k0 + 128, r < 4; 16 KB), built from
// the natural-order spectrum of the radix-16 forward kernel into the unused negative-frequency half of the row.
__global__ void k_dif_layout4096(float2 *__restrict__ xhat) {
  const float2 *r = xhat + (int64_t)blockIdx.x * 4096;
  float4 *o = reinterpret_cast<float4 *>(xhat + (int64_t)blockIdx.x * 4096 + 2048);
  for (int i = threadIdx.x; i < 4 * 8 * 32; i += blockDim.x) {
    const int lane = i & 31, m = (i >> 5) & 7, h = i >> 8;
    const float2 v0 = r[4 * (lane + 64 * m) + h], v1 = r[4 * (lane + 64 * m) + h + 128];
    o[i] = make_float4(v0.x, v1.x, v0.y, v1.y);
  }
}

// One CTA = one (series, chunk of scales).  smem: 2 * N complex.
template <typename T>
__global__ void k_cwt_rows(const cplx<T> *__restrict__ xhat, int n0, FftPlan<T> plan, int S,
                           int chunk, const double *__restrict__ scales, double dt, double f0,
                           T *__restrict__ power,
                           cplx<T> *__restrict__ coef, int coi_mask, double coi_c, double flambda,
                           int mother, int order, T pre_re, T pre_im) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx<T> *a = reinterpret_cast<cplx<T> *>(smem_raw);
  cplx<T> *b = a + plan.M;
  const int N = plan.n;
  const int nchunks = (S + chunk - 1) / chunk;
  const int64_t row = blockIdx.x / nchunks;
  const int c = blockIdx.x % nchunks;
  const cplx<T> *xh = xhat + row * N;
  const int s_end = min(S, (c + 1) * chunk);
  for (int s = c * chunk; s < s_end; ++s) {
    const double sc = scales[s];
    const T s_over_dt = T(sc / dt);
    // (s * ftfreqs[1] * N)^0.5 * pi^-0.25, times 1/N of the inverse transform
    const T norm = T(sqrt(2.0 * kPi * sc / dt) * kPiM14 / double(N));
    if (mother == WTB_MORLET) {
      for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const T d = morlet_daughter<T>(k, N, s_over_dt, norm, T(f0));
        cplx<T> v = xh[k];
        a[k] = mk<T>(v.x * d, v.y * d);
      }
    } else {
      // Paul: (sw)^m exp(-sw) H(sw);  DOG: (sw)^m exp(-(sw)^2/2);  times conj(prefactor)
      const T nrm = T(sqrt(2.0 * kPi * sc / dt) / double(N));
      for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const int kk = (k < (N + 1) / 2) ? k : k - N;
        const T z = s_over_dt * (T(2.0 * kPi) * T(kk) / T(N));
        // z^m e^{-z} (Paul) and z^m e^{-z^2/2} (DOG) evaluated as one exponential: z^m alone
        // overflows FP32 for large orders and scales while the product is tiny (inf * 0 = NaN)
        const T az = fabs(z);
        const T lz = T(order) * log(az);                      // -inf at z = 0: the daughter vanishes (m >= 1)
        const T sgn = (z < T(0) && (order & 1)) ? T(-1) : T(1);
        const T g = mother == WTB_PAUL ? (z > T(0) ? dev_exp<T>(lz - z) : T(0))
                                       : (az > T(0) ? sgn * dev_exp<T>(lz - T(0.5) * z * z) : T(0));
        const T dr = nrm * pre_re * g, di = nrm * pre_im * g;
        cplx<T> v = xh[k];
        a[k] = mk<T>(v.x * dr - v.y * di, v.x * di + v.y * dr);
      }
    }
    __syncthreads();
    cplx<T> *r = plan_fft<T, +1>(a, b, plan);
    const int64_t obase = (row * S + s) * (int64_t)n0;
    const double period = 1.0 / (1.0 / (flambda * sc));
    for (int t = threadIdx.x; t < n0; t += blockDim.x) {
      const cplx<T> w = r[t];
      if (coef) coef[obase + t] = w;
      if (power) {
        T p = w.x * w.x + w.y * w.y;
        if (coi_mask) {
          const double coi = coi_c * (n0 / 2.0 - fabs(t - (n0 - 1) / 2.0));
          if (period > coi) p = T(NAN);
        }
        power[obase + t] = p;
      }
    }
    __syncthreads();  // r may alias a; next iteration overwrites it
  }
}

template <typename T>
static int cwt_device(const T *d_x, int64_t batch, int n0, int N, double dt, const Axes &ax,
                      const Mother &mo, int flags, T *d_power, cplx<T> *d_coef, cudaStream_t st) {
  const int S = ax.J + 1;
  const double f0 = mo.kind == WTB_MORLET ? mo.param : 0.0;
  if (mo.kind == WTB_MORLET && sizeof(T) == 4 && d_coef == nullptr && d_power != nullptr &&
      !(flags & WTB_GENERIC_ONLY)) {
    int rc = cwt_fast_try((const float *)d_x, batch, n0, N, dt, ax, f0, flags, (float *)d_power, st);
    if (rc != 1) return rc;
  }
  FftPlan<T> plan;
  WTB_TRY(make_plan<T>(N, &plan));
  const size_t smem = 2 * sizeof(cplx<T>) * (size_t)plan.M;
  WTB_REQUIRE(smem <= 227 * 1024, WTB_EUNSUPPORTED,
              "nfft=%d needs %zu B of shared memory per CTA (limit 227 KB): a power of two up to %d, any other "
              "length up to %d for %s", N, smem, sizeof(T) == 4 ? 8192 : 4096, sizeof(T) == 4 ? 4096 : 2048,
              sizeof(T) == 4 ? "float" : "double");
  void *scratch = nullptr;
  const size_t sc_bytes = ((sizeof(double) * S + 255) / 256) * 256;
  WTB_TRY(arena_reserve(sc_bytes + sizeof(cplx<T>) * (size_t)batch * N, &scratch));
  double *d_scales = (double *)scratch;
  cplx<T> *d_xhat = (cplx<T> *)((char *)scratch + sc_bytes);
  WTB_CUDA(cudaMemcpyAsync(d_scales, ax.scales.data(), sizeof(double) * S, cudaMemcpyHostToDevice, st));
  const int threads = plan.M >= 1024 ? 256 : (plan.M >= 256 ? 128 : 64);
  WTB_CUDA(cudaFuncSetAttribute(k_fwd_fft<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WTB_CUDA(cudaFuncSetAttribute(k_cwt_rows<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int fwd_rc = 1;
  if constexpr (sizeof(T) == 4) {
    if (mo.kind == WTB_MORLET && !(flags & WTB_GENERIC_ONLY))
      fwd_rc = fwd_fft_4096_try((const float *)d_x, batch, n0, N, (float2 *)d_xhat, st);
    if (fwd_rc < 0) return fwd_rc;
  }
  bool dif = false;            // the two / four-warps-per-row kernels take this shape (spectra in group layout)
  if constexpr (sizeof(T) == 4) {
    dif = mo.kind == WTB_MORLET && !(flags & WTB_GENERIC_ONLY) && d_power && !d_coef && cwt_dif_covers(batch, n0, N, S, f0);
    if (dif && N == 2048) {
      WTB_CUDA(cudaFuncSetAttribute(k_fwd_fft_pair2048, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_fwd_fft_pair2048<<<(unsigned)batch, threads, smem, st>>>((const float *)d_x, n0, plan, (float2 *)d_xhat);
      WTB_LAUNCH_CHECK();
      fwd_rc = 0;
    }
  }
  if (fwd_rc == 1) {
    k_fwd_fft<T><<<(unsigned)batch, threads, smem, st>>>(d_x, n0, plan, d_xhat);
    WTB_LAUNCH_CHECK();
  }
  if constexpr (sizeof(T) == 4) {
    if (dif) {
      if (N == 4096) {
        k_dif_layout4096<<<(unsigned)batch, 256, 0, st>>>((float2 *)d_xhat);
        WTB_LAUNCH_CHECK();
      }
      const int rc = cwt_dif_try((const float2 *)d_xhat + (N == 4096 ? 2048 : 0), N, batch, n0, N, dt, ax, f0, flags,
                                 (float *)d_power, st);
      if (rc != 1) return rc;
    }
    if (mo.kind == WTB_MORLET && !(flags & WTB_GENERIC_ONLY)) {
      const int rc = cwt_rows_4096_try((const float2 *)d_xhat, batch, n0, N, dt, ax, f0, flags, (float *)d_power,
                                       (float2 *)d_coef, st);
      if (rc != 1) return rc;
    }
  }
  // enough CTAs to fill the machine for small batches, few forward-FFT re-reads for large ones
  int chunk = S;
  const int64_t want = 4LL * sm_count();
  if (batch < want) chunk = (int)std::max<int64_t>(1, (int64_t)S * batch / want);
  chunk = std::max(1, std::min(chunk, 16));
  const int nchunks = (S + chunk - 1) / chunk;
  WTB_REQUIRE(batch * nchunks < (1LL << 31), WTB_EUNSUPPORTED, "batch too large for one launch");
  double pre_re, pre_im;
  mother_prefactor(mo, &pre_re, &pre_im);
  k_cwt_rows<T><<<(unsigned)(batch * nchunks), threads, smem, st>>>(
      d_xhat, n0, plan, S, chunk, d_scales, dt, f0, d_power, d_coef,
      (flags & WTB_COI_MASK) ? 1 : 0, mother_flambda(mo) * mother_coi(mo) * dt, mother_flambda(mo), mo.kind,
      (int)mo.param, T(pre_re), T(pre_im));
  WTB_LAUNCH_CHECK();
  return WTB_OK;
}

static thread_local bool g_in_shard = false;   // set on a pool worker while it runs its block

template <typename T>
static int cwt_entry(const void *x, int64_t batch, int n0, int N, double dt, const Axes &ax,
                     const Mother &f0, int flags, void *power_out, void *coef_out, cudaStream_t st) {
  const int S = ax.J + 1;
  if (flags & WTB_DEVICE_PTRS)
    return cwt_device<T>((const T *)x, batch, n0, N, dt, ax, f0, flags, (T *)power_out,
                         (cplx<T> *)coef_out, st);
  // host buffers, several GPUs (wtb_init_multi): contiguous blocks of series, one per device
  if (pool_size() > 1 && batch >= 2 * pool_size() && !g_in_shard) {
    return run_sharded_fn(batch, 2, st, [&](int, int64_t first, int64_t count, cudaStream_t s) {
      g_in_shard = true;
      const int rc = cwt_entry<T>((const T *)x + first * n0, count, n0, N, dt, ax, f0, flags,
                                  power_out ? (T *)power_out + first * S * n0 : nullptr,
                                  coef_out ? (cplx<T> *)coef_out + first * S * n0 : nullptr, s);
      g_in_shard = false;
      return rc;
    });
  }
  // host buffers: stream the batch through a bounded staging arena
  const size_t per_row = sizeof(T) * ((size_t)n0 + (power_out ? (size_t)S * n0 : 0) +
                                      (coef_out ? 2 * (size_t)S * n0 : 0));
  const size_t budget = size_t(1) << 30;
  int64_t rows = std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)(budget / per_row)));
  void *stage = nullptr;
  WTB_TRY(staging_reserve(per_row * rows + 1024, &stage));
  T *d_x = (T *)stage;
  size_t off = (sizeof(T) * (size_t)rows * n0 + 255) / 256 * 256;
  T *d_power = nullptr;
  cplx<T> *d_coef = nullptr;
  if (power_out) { d_power = (T *)((char *)stage + off); off += (sizeof(T) * (size_t)rows * S * n0 + 255) / 256 * 256; }
  if (coef_out) d_coef = (cplx<T> *)((char *)stage + off);
  for (int64_t b0 = 0; b0 < batch; b0 += rows) {
    const int64_t nb = std::min(rows, batch - b0);
    WTB_CUDA(cudaMemcpyAsync(d_x, (const T *)x + b0 * n0, sizeof(T) * nb * n0, cudaMemcpyHostToDevice, st));
    WTB_TRY(cwt_device<T>(d_x, nb, n0, N, dt, ax, f0, flags, d_power, d_coef, st));
    if (power_out)
      WTB_TRY(copy_to_host((T *)power_out + b0 * S * n0, d_power, sizeof(T) * nb * S * n0, st));
    if (coef_out)
      WTB_TRY(copy_to_host((cplx<T> *)coef_out + b0 * S * n0, d_coef, sizeof(cplx<T>) * nb * S * n0, st));
    WTB_CUDA(cudaStreamSynchronize(st));
  }
  return WTB_OK;
}

}  // namespace wtb

using namespace wtb;

extern "C" int wtb_cwt(const void *x, int64_t batch, int n0, int nfft, double dt, double dj, double s0, int J,
                       int mother, double param, int flags, void *power_out, void *coef_out, void *stream) {
  WTB_REQUIRE(x && batch >= 0 && n0 > 0, WTB_EINVAL, "wtb_cwt: bad x/batch/n0");
  WTB_REQUIRE(power_out || coef_out, WTB_EINVAL, "wtb_cwt: no output requested");
  WTB_REQUIRE(nfft >= n0 && nfft >= 2, WTB_EINVAL,
              "nfft=%d must be >= n0=%d (a power of two: pycwt's scipy.fftpack padding rule; n0 itself: "
              "the un-padded mkl_fft rule)", nfft, n0);
  WTB_ENTER(flags, x, stream);
  Mother mo;
  mo.kind = mother;
  mo.param = param;
  Axes ax;
  WTB_TRY(resolve_axes(n0, dt, dj, s0, J, mo, &ax));
  if (batch == 0) return WTB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (flags & WTB_F64) return cwt_entry<double>(x, batch, n0, nfft, dt, ax, mo, flags, power_out, coef_out, st);
  return cwt_entry<float>(x, batch, n0, nfft, dt, ax, mo, flags, power_out, coef_out, st);
}

extern "C" int wtb_cwt_morlet(const void *x, int64_t batch, int n0, int nfft, double dt, double dj,
                              double s0, int J, double f0, int flags, void *power_out,
                              void *coef_out, void *stream) {
  return wtb_cwt(x, batch, n0, nfft, dt, dj, s0, J, WTB_MORLET, f0, flags, power_out, coef_out, stream);
}

// ---- inverse transform (pycwt.icwt): x[t] = factor * sum_s Re(W[s,t]) / sqrt(s_j) ----------
namespace wtb {
template <typename T>
__global__ void k_icwt(const cplx<T> *__restrict__ coef, int S, int n0, const double *__restrict__ scales,
                       double factor, T *__restrict__ out) {
  const int64_t b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n0) return;
  const cplx<T> *w = coef + b * (int64_t)S * n0 + t;
  double acc = 0;
  for (int s = 0; s < S; ++s) acc += (double)w[(int64_t)s * n0].x * rsqrt(scales[s]);
  out[b * (int64_t)n0 + t] = T(factor * acc);
}

template <typename T>
static int icwt_impl(const void *coef, int64_t batch, int S, int n0, const double *scales, double factor, int flags,
                     void *out, cudaStream_t st) {
  const bool dev = flags & WTB_DEVICE_PTRS;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  void *scratch = nullptr;
  WTB_TRY(arena_reserve(al(sizeof(double) * S), &scratch));
  double *d_scales = (double *)scratch;
  WTB_CUDA(cudaMemcpyAsync(d_scales, scales, sizeof(double) * S, cudaMemcpyHostToDevice, st));
  const size_t in_row = sizeof(cplx<T>) * (size_t)S * n0, out_row = sizeof(T) * (size_t)n0;
  const int64_t rows = dev ? batch : std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)((size_t(1) << 30) / in_row)));
  const cplx<T> *d_in = (const cplx<T> *)coef;
  T *d_out = (T *)out;
  if (!dev) {
    void *stage = nullptr;
    WTB_TRY(staging_reserve(al(in_row * rows) + al(out_row * rows), &stage));
    d_in = (const cplx<T> *)stage;
    d_out = (T *)((char *)stage + al(in_row * rows));
  }
  for (int64_t b0 = 0; b0 < batch; b0 += rows) {
    const int64_t nb = std::min(rows, batch - b0);
    WTB_REQUIRE(nb < 65536, WTB_EUNSUPPORTED, "wtb_icwt: at most 65535 series per launch");
    if (!dev) WTB_CUDA(cudaMemcpyAsync((void *)d_in, (const char *)coef + b0 * in_row, in_row * nb, cudaMemcpyHostToDevice, st));
    k_icwt<T><<<dim3((n0 + 255) / 256, (unsigned)nb), 256, 0, st>>>(d_in, S, n0, d_scales, factor, d_out);
    WTB_LAUNCH_CHECK();
    if (!dev) {
      WTB_TRY(copy_to_host((char *)out + b0 * out_row, d_out, out_row * nb, st));
      WTB_CUDA(cudaStreamSynchronize(st));
    }
  }
  return WTB_OK;
}
}  // namespace wtb

extern "C" int wtb_icwt(const void *coef, int64_t batch, int S, int n0, const double *scales, double factor,
                        int flags, void *x_out, void *stream) {
  WTB_REQUIRE(coef && x_out && scales && batch >= 0 && S > 0 && n0 > 0, WTB_EINVAL, "wtb_icwt: bad arguments");
  WTB_ENTER(flags, coef, stream);
  if (batch == 0) return WTB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (flags & WTB_F64) return icwt_impl<double>(coef, batch, S, n0, scales, factor, flags, x_out, st);
  return icwt_impl<float>(coef, batch, S, n0, scales, factor, flags, x_out, st);
}
