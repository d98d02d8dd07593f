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

}  // namespace wtb
