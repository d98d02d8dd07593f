// FP32 fast path of the fused CWT+power kernel for nfft = 1024 (BASELINE cfg4).
//
// One warp owns one series at a time and never synchronises with other warps:
//   forward FFT(1024) -> X^ in shared memory (positive half only)
//   for every scale row s: Y = X^ * daughter_s (band-limited, one-sided)
//                          w = IFFT_1024(Y) as 32 x 32:  k = k1 + 32 k2,  t = t2 + 32 t1
//                          power = |w|^2 -> 128-byte coalesced streaming stores
// A 1024-point transform is two in-register 32-point DFTs per thread with one
// shared-memory transpose between them:
//   step A (lane = k1): A[k1][t2]  = sum_k2 Y[k1+32 k2] w32^(k2 t2), times w1024^(k1 t2)
//   step B (lane = t2): x[t2+32t1] = sum_k1 A'[k1][t2] w32^(k1 t1)
// Lanes map to t2, so for each t1 the warp stores 32 consecutive floats.
//
// Work that the Morlet daughter makes unnecessary is skipped exactly (to 1e-6 of
// the daughter peak, far inside the FP32 tolerance): negative frequencies, and
// bins above k_hi(s) = (f0 + 5.3) N dt / (2 pi s).  Rows with k_hi < 32 need no
// step A and no transpose; rows with fewer non-zero inputs skip butterfly stages.
#include "common.cuh"
#include "fft32_gen.cuh"

namespace wtb {

namespace {

constexpr int kN = 1024;
constexpr int kWarps = 8;             // warps per CTA
constexpr int kTrStride = 33;         // padded row of the transpose buffer (float2 units)
constexpr float kZCut = 5.3f;         // daughter dropped where |s*w - f0| > kZCut  (exp(-14) ~ 8e-7)

struct RowParam {
  float a;      // (s/dt) * 2*pi/N : s*w_k = a*k
  float norm;   // sqrt(2*pi*s/dt) * pi^-1/4 / N
  int L;        // log2(#non-zero inputs) of the first DFT that runs (step A if multi, else step B)
  int multi;    // 1: k_hi >= 32 -> step A + transpose + full step B
};

struct WarpSmem {
  float2 xhat[kN / 2];              // X^[k], k < 512
  float2 tr[32 * kTrStride];        // transpose buffer
  float2 y[32];                     // Y[k] of a single-pass row
};

struct CtaSmem {
  float2 tw[32 * 32];               // tw[a*32+b] = exp(+2*pi*i*a*b/1024)  (symmetric)
  WarpSmem w[kWarps];
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

constexpr int kMaxRows = 512;          // scale rows staged in shared memory

// One dit32 call site serves the forward transform (s = -1), step A and step B of
// every scale row: the hot code stays well inside the 32 KB instruction cache.
__global__ void __launch_bounds__(kWarps * 32, 2)
k_cwt_fast_1024(const float *__restrict__ x, int64_t batch, int n0, int S,
                const RowParam *__restrict__ rows, float f0, float *__restrict__ power) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaSmem &sm = *reinterpret_cast<CtaSmem *>(smem_raw);
  RowParam *srow = reinterpret_cast<RowParam *>(smem_raw + sizeof(CtaSmem));
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)((i >> 5) * (i & 31)) / (float)kN, &s, &c);
    sm.tw[i] = make_float2(c, s);
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) srow[i] = rows[i];
  __syncthreads();
  WarpSmem &ws = sm.w[warp];
  const float2 *tw = sm.tw;
  const int64_t gwarp = (int64_t)blockIdx.x * kWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  const bool full_row = (n0 == kN);
  const float lanef = (float)lane;
  float2 u[32];

  for (int64_t b = gwarp; b < batch; b += nwarps) {
    const float *xr = x + b * n0;
    float *out = power + b * (int64_t)S * n0 + lane;
#pragma unroll 1
    for (int s = -1; s < S; ++s) {
      int L, two_pass;
      if (s < 0) {
        // forward FFT of the real series: X^[t] = conj(sum_k x[k] w^(+k t))
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) {
          const int k = lane + 32 * k2;
          u[fft32::br5(k2)] = make_float2(k < n0 ? __ldg(xr + k) : 0.0f, 0.0f);
        }
        L = 5;
        two_pass = 1;
      } else {
        const RowParam rp = srow[s];
        L = rp.L;
        two_pass = rp.multi;
        const float zl = fmaf(rp.a, lanef, -f0);   // s*w_k - f0 at k = lane
        if (two_pass) {
          const int K2 = 1 << rp.L;
          const float a32 = rp.a * 32.0f;
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            if (k2 < K2) {
              const float z = fmaf(a32, (float)k2, zl);
              const float d = rp.norm * __expf(-0.5f * z * z);
              const float2 v = ws.xhat[lane + 32 * k2];
              u[fft32::br5(k2)] = make_float2(v.x * d, v.y * d);
            }
          }
        } else {
          const float d = rp.norm * __expf(-0.5f * zl * zl);
          const float2 v = ws.xhat[lane];
          ws.y[lane] = make_float2(v.x * d, v.y * d);
          __syncwarp();
          const int K1 = 1 << rp.L;
#pragma unroll
          for (int k1 = 0; k1 < 32; ++k1)
            if (k1 < K1) u[fft32::br5(k1)] = cmulf(ws.y[k1], tw[k1 * 32 + lane]);
          __syncwarp();
        }
      }
#pragma unroll 1
      for (;;) {
        fft32::dit32(u, L);
        if (!two_pass) break;
        // step A done (lane = k1): twiddle by w1024^(k1 t2), transpose, reload with lane = t2
#pragma unroll
        for (int t2 = 0; t2 < 32; ++t2)
          ws.tr[lane * kTrStride + t2] = cmulf(u[t2], tw[t2 * 32 + lane]);
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) u[fft32::br5(k1)] = ws.tr[k1 * kTrStride + lane];
        __syncwarp();
        L = 5;
        two_pass = 0;
      }
      if (s < 0) {
#pragma unroll
        for (int t1 = 0; t1 < 16; ++t1) ws.xhat[lane + 32 * t1] = make_float2(u[t1].x, -u[t1].y);
        __syncwarp();
      } else {
        float *orow = out + (int64_t)s * n0;
        if (full_row) {
#pragma unroll
          for (int t1 = 0; t1 < 32; ++t1)
            __stcs(orow + 32 * t1, fmaf(u[t1].x, u[t1].x, u[t1].y * u[t1].y));
        } else {
#pragma unroll
          for (int t1 = 0; t1 < 32; ++t1)
            if (lane + 32 * t1 < n0) __stcs(orow + 32 * t1, fmaf(u[t1].x, u[t1].x, u[t1].y * u[t1].y));
        }
      }
    }
  }
}

}  // namespace

int cwt_fast_try(const float *d_x, int64_t batch, int n0, int nfft, double dt, const Axes &ax, double f0,
                 int flags, float *d_power, cudaStream_t st) {
  if (nfft != kN || (flags & WTB_COI_MASK) || f0 < 1.0) return 1;
  const int S = ax.J + 1;
  std::vector<RowParam> rows(S);
  for (int s = 0; s < S; ++s) {
    const double a = ax.scales[s] / dt * 2.0 * kPi / kN;
    int khi = (int)std::floor((f0 + kZCut) / a);
    if (khi > kN / 2 - 1) khi = kN / 2 - 1;
    if (khi < 0) khi = 0;
    RowParam &r = rows[s];
    r.a = (float)a;
    r.norm = (float)(std::sqrt(2.0 * kPi * ax.scales[s] / dt) * 0.75112554446494248286 / kN);
    if (khi >= 32) {
      r.multi = 1;
      r.L = ilog2(khi / 32 + 1);       // K2 = 2^L >= ceil((khi+1)/32), at most 16 (one-sided)
      if (r.L > 4) r.L = 4;
    } else {
      r.multi = 0;
      r.L = ilog2(khi + 1);            // K1 = 2^L >= khi+1
    }
  }
  void *scratch = nullptr;
  WTB_TRY(arena_reserve(sizeof(RowParam) * S, &scratch));
  RowParam *d_rows = (RowParam *)scratch;
  WTB_CUDA(cudaMemcpyAsync(d_rows, rows.data(), sizeof(RowParam) * S, cudaMemcpyHostToDevice, st));
  // rows.data() is pageable: the copy is staged before the call returns
  if (S > kMaxRows) return 1;
  const size_t smem = sizeof(CtaSmem) + sizeof(RowParam) * S;
  WTB_CUDA(cudaFuncSetAttribute(k_cwt_fast_1024, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ctas_needed = (batch + kWarps - 1) / kWarps;
  const int grid = (int)std::min<int64_t>(ctas_needed, 2LL * sm_count());
  k_cwt_fast_1024<<<grid, kWarps * 32, smem, st>>>(d_x, batch, n0, S, d_rows, (float)f0, d_power);
  WTB_LAUNCH_CHECK();
  return WTB_OK;
}

}  // namespace wtb
