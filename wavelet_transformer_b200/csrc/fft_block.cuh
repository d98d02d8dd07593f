// Generic in-shared-memory Stockham autosort FFT, one transform per CTA.
//
// Any power-of-two N >= 2, float or double.  Radix-4 passes (plus one radix-2
// pass when log2 N is odd) ping-pong between two shared buffers; the result
// pointer is returned (it is `a` when the pass count is even, `b` otherwise).
// This is the "any N / any precision" path; the FP32 headline sizes have their
// own register-resident kernels (cwt_fast.cu).
#pragma once

#include "common.cuh"

namespace wtb {

// tw[k] = exp(-2*pi*i*k/N).  SIGN = -1 forward, +1 inverse (unnormalised).
template <typename T, int SIGN>
__device__ __forceinline__ cplx<T> tw_at(const cplx<T> *__restrict__ tw, int idx) {
  cplx<T> w = tw[idx];
  if (SIGN > 0) w.y = -w.y;
  return w;
}

// multiply by -i (forward) or +i (inverse)
template <typename T, int SIGN> __device__ __forceinline__ cplx<T> rot90(cplx<T> v) {
  return SIGN < 0 ? mk<T>(v.y, -v.x) : mk<T>(-v.y, v.x);
}

// All threads of the CTA must call this; `a` holds the input; contains
// __syncthreads() (one after every pass, so the result is visible on return).
template <typename T, int SIGN>
__device__ cplx<T> *block_fft(cplx<T> *a, cplx<T> *b, int N, int log2N,
                              const cplx<T> *__restrict__ tw) {
  using C = cplx<T>;
  C *in = a, *out = b;
  int Ns = 1;
  int log2Ns = 0;
  if (log2N & 1) {
    const int half = N >> 1;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      C v0 = in[j], v1 = in[j + half];
      out[2 * j] = cadd(v0, v1);
      out[2 * j + 1] = csub(v0, v1);
    }
    Ns = 2;
    log2Ns = 1;
    C *t = in; in = out; out = t;
    __syncthreads();
  }
  const int quarter = N >> 2;
  while (Ns < N) {
    const int tstride = N >> (log2Ns + 2);  // table step: N / (4*Ns)
    for (int j = threadIdx.x; j < quarter; j += blockDim.x) {
      const int k = j & (Ns - 1);
      C v0 = in[j];
      C v1 = in[j + quarter];
      C v2 = in[j + 2 * quarter];
      C v3 = in[j + 3 * quarter];
      if (k) {
        const int base = k * tstride;
        v1 = cmul(v1, tw_at<T, SIGN>(tw, base));
        v2 = cmul(v2, tw_at<T, SIGN>(tw, 2 * base));
        v3 = cmul(v3, tw_at<T, SIGN>(tw, 3 * base));
      }
      const C t0 = cadd(v0, v2), t1 = csub(v0, v2);
      const C t2 = cadd(v1, v3), t3 = rot90<T, SIGN>(csub(v1, v3));
      const int j0 = ((j - k) << 2) + k;
      out[j0] = cadd(t0, t2);
      out[j0 + Ns] = cadd(t1, t3);
      out[j0 + 2 * Ns] = csub(t0, t2);
      out[j0 + 3 * Ns] = csub(t1, t3);
    }
    Ns <<= 2;
    log2Ns += 2;
    C *t = in; in = out; out = t;
    __syncthreads();
  }
  return in;
}

// ---- transforms of any length ------------------------------------------------------------
// pycwt pads every transform to a power of two only when mkl_fft is absent (the reference's pip /
// uv install); its conda environment (environment.yml:126, mkl_fft) transforms the series at its
// own length.  A plan covers both: a power-of-two length n is one Stockham FFT; any other n is
// Bluestein's chirp-z identity  t k = (t^2 + k^2 - (k - t)^2) / 2,
//   X[k] = c[k] * sum_t (x[t] c[t]) conj(c)[k - t],   c[t] = exp(-i pi t^2 / n),
// i.e. one circular convolution of length M = 2^m >= 2 n - 1 with a fixed kernel whose transform
// is tabulated (chat, 1/M folded in): two length-M FFTs and three pointwise passes.
template <typename T> struct FftPlan {
  int n = 0;                          // transform length
  int M = 0;                          // work length: n itself when it is a power of two
  int log2M = 0;
  const cplx<T> *tw = nullptr;        // exp(-2 pi i k / M)
  const cplx<T> *chirp = nullptr;     // [n]  exp(-i pi t^2 / n); null for a power of two
  const cplx<T> *chat = nullptr;      // [M]  FFT_M of the wrapped conj chirp, divided by M
};

// `a` holds the n inputs (capacity M, like `b`); all threads call; the result pointer (n valid
// values) is returned after a barrier.  SIGN = -1 forward, +1 inverse (unnormalised).
template <typename T, int SIGN>
__device__ cplx<T> *plan_fft(cplx<T> *a, cplx<T> *b, const FftPlan<T> &p) {
  using C = cplx<T>;
  if (!p.chirp) return block_fft<T, SIGN>(a, b, p.M, p.log2M, p.tw);
  for (int t = threadIdx.x; t < p.M; t += blockDim.x) {
    C v = mk<T>(T(0), T(0));
    if (t < p.n) {
      C c = p.chirp[t];
      if (SIGN > 0) c.y = -c.y;
      v = cmul(a[t], c);
    }
    a[t] = v;
  }
  __syncthreads();
  C *r = block_fft<T, -1>(a, b, p.M, p.log2M, p.tw);
  C *other = (r == a) ? b : a;
  for (int k = threadIdx.x; k < p.M; k += blockDim.x) {
    C h = p.chat[k];
    if (SIGN > 0) h.y = -h.y;           // the kernel is even, so conj(chat) is the conjugate chirp's transform
    r[k] = cmul(r[k], h);
  }
  __syncthreads();
  C *z = block_fft<T, +1>(r, other, p.M, p.log2M, p.tw);
  for (int t = threadIdx.x; t < p.n; t += blockDim.x) {
    C c = p.chirp[t];
    if (SIGN > 0) c.y = -c.y;
    z[t] = cmul(z[t], c);
  }
  __syncthreads();
  return z;
}

}  // namespace wtb
