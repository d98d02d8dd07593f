// Shared host/device helpers for libwavelet_sm100a.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/wtb.h"

namespace wtb {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define WTB_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return ::wtb::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// after every kernel launch: count it (wtb_kernel_launches) and surface launch errors
void note_launch();
#define WTB_LAUNCH_CHECK()            \
  do {                                \
    ::wtb::note_launch();             \
    WTB_CUDA(cudaGetLastError());     \
  } while (0)

#define WTB_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::wtb::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define WTB_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != WTB_OK) return rc__; \
  } while (0)

// ---- complex arithmetic -----------------------------------------------------
template <typename T> struct cplx_of;
template <> struct cplx_of<float> { using type = float2; };
template <> struct cplx_of<double> { using type = double2; };
template <typename T> using cplx = typename cplx_of<T>::type;

template <typename T> __host__ __device__ __forceinline__ cplx<T> mk(T x, T y) {
  cplx<T> r; r.x = x; r.y = y; return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
  C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}

constexpr double kPi = 3.14159265358979323846264338327950288;

// ---- device scratch / tables (host side) --------------------------------------
// Grow-only device arena owned by the library; one per host thread so that
// concurrent Streamlit sessions never share scratch.  Freed by wtb_shutdown().
struct Arena {
  void *ptr = nullptr;
  size_t cap = 0;
  int device = -1;
};

int arena_reserve(size_t bytes, void **out);  // thread-local arena, 256B aligned
// Second independent arena (I/O staging for host-pointer calls).
int staging_reserve(size_t bytes, void **out);
// small per-thread buffer for kernel parameter tables, separate from the two arenas above so a
// callee can fill it while its caller's arena holds live data; released by wtb_shutdown as well
int params_reserve(size_t bytes, void **out);

// Twiddle table exp(-2*pi*i*k/N), k in [0,N), in precision T; cached per (device,N).
template <typename T> int twiddles(int N, const cplx<T> **out);

int ensure_device();  // lazily binds device 0 (or WTB_DEVICE) and checks sm_100
int sm_count();

inline int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// Morlet constants (pycwt.mothers.Morlet)
inline double morlet_flambda(double f0) { return 4.0 * kPi / (f0 + std::sqrt(2.0 + f0 * f0)); }

// pycwt.mothers: kind = WTB_MORLET / WTB_PAUL / WTB_DOG, param = f0 or the order m
struct Mother {
  int kind = WTB_MORLET;
  double param = 6.0;
};
inline double mother_flambda(const Mother &m) {
  if (m.kind == WTB_PAUL) return 4.0 * kPi / (2.0 * m.param + 1.0);
  if (m.kind == WTB_DOG) return 2.0 * kPi / std::sqrt(m.param + 0.5);
  return morlet_flambda(m.param);
}
// e-folding factor of the cone of influence (Torrence & Compo table 1)
inline double mother_coi(const Mother &m) { return m.kind == WTB_PAUL ? std::sqrt(2.0) : 1.0 / std::sqrt(2.0); }
// conj of the constant complex prefactor of psi_ft (1 for Morlet: pi^-1/4 stays in the kernel)
inline void mother_prefactor(const Mother &m, double *re, double *im) {
  *re = 1.0;
  *im = 0.0;
  const int order = (int)m.param;
  if (m.kind == WTB_PAUL) {
    double fact = 1.0;  // (2m-1)!
    for (int k = 2; k <= 2 * order - 1; ++k) fact *= k;
    *re = std::pow(2.0, order) / std::sqrt(order * fact);
  } else if (m.kind == WTB_DOG) {
    const double c = 1.0 / std::sqrt(std::tgamma(m.param + 0.5));
    // -i^m: m%4 = 0 -> -1, 1 -> -i, 2 -> +1, 3 -> +i; conjugated here
    switch (order & 3) {
      case 0: *re = -c; break;
      case 1: *re = 0; *im = c; break;
      case 2: *re = c; break;
      default: *re = 0; *im = -c; break;
    }
  }
}

struct Axes {
  int J = 0;
  std::vector<double> scales, freqs;
};
int resolve_axes(int n0, double dt, double dj, double s0, int J, double f0, Axes *ax);
int resolve_axes(int n0, double dt, double dj, double s0, int J, const Mother &m, Axes *ax);

}  // namespace wtb
