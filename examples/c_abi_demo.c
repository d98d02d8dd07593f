/* The C ABI of libwavelet_sm100a.so from plain C: what a cgo / JNI / N-API binding would do.
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo \
 *       -Lwavelet_transformer_b200/lib -lwavelet_sm100a -Wl,-rpath,$PWD/wavelet_transformer_b200/lib -lm
 *   /tmp/c_abi_demo            # needs a B200; prints one line of checksums
 *
 * Mirrors src/cwt.py:110-114 (pycwt.cwt + |W|^2 of one series, float64), src/modwt.py:126
 * (MODWT, LA8, J = 4) on a deterministic test signal, and the Monte-Carlo coherence significance
 * that src/wct.py:106-118 reaches through pycwt.wct(sig=True): one call, spread over every GPU
 * named by WTB_GPUS (wtb_init_multi), thresholds independent of the GPU count.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "wtb.h"

#define CHECK(call)                                                         \
  do {                                                                      \
    int rc_ = (call);                                                       \
    if (rc_ != WTB_OK) {                                                    \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, wtb_last_error()); \
      return 1;                                                             \
    }                                                                       \
  } while (0)

int main(void) {
  enum { N0 = 600, NFFT = 1024 };
  const double dt = 1.0 / 12, dj = 1.0 / 12, s0 = 2.0 / 12, f0 = 6.0;
  static double x[N0];
  for (int t = 0; t < N0; ++t) x[t] = sin(2 * M_PI * t / 37.0) + 0.5 * cos(2 * M_PI * t / 90.0);

  CHECK(wtb_init(0));
  int J = -1;
  CHECK(wtb_cwt_axes(N0, dt, dj, s0, -1, f0, &J, NULL, NULL, NULL));
  const int S = J + 1;
  double *scales = malloc(sizeof(double) * S), *power = malloc(sizeof(double) * (size_t)S * N0);
  CHECK(wtb_cwt_axes(N0, dt, dj, s0, J, f0, NULL, scales, NULL, NULL));
  CHECK(wtb_cwt_morlet(x, 1, N0, NFFT, dt, dj, s0, J, f0, WTB_F64, power, NULL, NULL));
  double psum = 0, pmax = 0;
  int smax = 0;
  for (int s = 0; s < S; ++s)
    for (int t = 0; t < N0; ++t) {
      const double p = power[(size_t)s * N0 + t];
      psum += p;
      if (p > pmax) { pmax = p; smax = s; }
    }

  /* LA8 = PyWavelets sym4 decomposition filters (dec_lo; dec_hi is its quadrature mirror) */
  const double g[8] = {-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161,
                       0.29785779560527736,  -0.09921954357684722, -0.012603967262037833, 0.0322231006040427};
  double h[8];
  for (int k = 0; k < 8; ++k) h[k] = ((k + 1) % 2 ? -1.0 : 1.0) * g[7 - k];
  enum { JM = 4 };
  static double w[(JM + 1) * N0], back[N0];
  CHECK(wtb_modwt(x, 1, N0, g, h, 8, JM, WTB_F64, w, NULL));
  CHECK(wtb_imodwt(w, 1, N0, g, h, 8, JM, WTB_F64, back, NULL));
  double err = 0, energy_w = 0, energy_x = 0;
  for (int t = 0; t < N0; ++t) {
    err = fmax(err, fabs(back[t] - x[t]));
    energy_x += x[t] * x[t];
  }
  for (int i = 0; i < (JM + 1) * N0; ++i) energy_w += w[i] * w[i];

  /* 95 % coherence thresholds of two AR(1) processes (0.8, 0.6): 24 realisations, seed 7 */
  enum { JS = 24 };
  double sig95[JS + 1];
  static uint64_t hist[(JS + 1) * WTB_NBINS];
  CHECK(wtb_init_multi(0)); /* WTB_GPUS, else every visible device; one device needs no workers */
  CHECK(wtb_wct_significance(0.8, 0.6, dt, 0.25, s0, JS, f0, 0.95, 24, 7, NULL, 0, sig95, hist));
  unsigned long long hist_total = 0;
  for (int i = 0; i < (JS + 1) * WTB_NBINS; ++i) hist_total += hist[i];

  printf("S=%d peak_scale=%.6f power_sum=%.9e modwt_energy_ratio=%.12f imodwt_err=%.3e sig95_0=%.9f "
         "hist_total=%llu gpus=%d launches=%llu\n", S,
         scales[smax], psum, energy_w / energy_x, err, sig95[0], hist_total, wtb_gpu_count(),
         (unsigned long long)wtb_kernel_launches());
  free(scales);
  free(power);
  wtb_shutdown();
  return 0;
}
