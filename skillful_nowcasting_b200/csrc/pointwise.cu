// HBM-bound kernels of the DGMR hot path: layout permutes, pooling/upsampling, ConvGRU gate
// arithmetic, BatchNorm (grouped batch statistics), losses, Adam.  All fp32, channels-last.
// Design rule (B200): these are bandwidth kernels -> coalesced along C, float4 where C%4==0,
// grid-stride with grids sized in multiples of the SM count, no shared-memory staging needed.
#include "common.cuh"
#include <mutex>
#include <string.h>
#include <math.h>

namespace dgmr {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

__device__ __forceinline__ float rna_tf32_pw(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------ permute
struct PermuteArgs {
  int ndim;
  int64_t shape[8], sstr[8], dstr[8];
};
// I = uint32_t when all offsets fit 32 bits (up to 8 div/mod pairs per element: 64-bit ones dominate the kernel otherwise)
// T = float4 when the innermost dimension is contiguous on both sides and everything is 16-byte aligned (shape / strides then count float4s)
__device__ __forceinline__ void permute_store(float* d, float v, int acc) { if (acc) *d += v; else *d = v; }
__device__ __forceinline__ void permute_store(float4* d, float4 v, int acc) {
  if (acc) { float4 o = *d; v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w); }
  *d = v;
}
template <typename I, typename T>
__global__ void permute_kernel(const T* __restrict__ src, T* __restrict__ dst, PermuteArgs a, int64_t total_, int acc) {
  const I total = (I)total_;
  for (I i = blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    I rem = i, so = 0, dof = 0;
#pragma unroll
    for (int d = 7; d >= 0; --d) {
      if (d < a.ndim) {
        const I sh = (I)a.shape[d];
        I q = rem / sh;
        I r = rem - q * sh;
        rem = q;
        so += r * (I)a.sstr[d];
        dof += r * (I)a.dstr[d];
      }
    }
    permute_store(dst + dof, src[so], acc);
  }
}

// ------------------------------------------------------------------ simple pointwise
__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, const float* __restrict__ y, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = y ? a * x[i] + b * y[i] : a * x[i];
}
__global__ void axpby4_kernel(float a, const float4* __restrict__ x, float b, const float4* __restrict__ y, float4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 u = x[i], r;
    if (y) { float4 v = y[i]; r = make_float4(a * u.x + b * v.x, a * u.y + b * v.y, a * u.z + b * v.z, a * u.w + b * v.w); }
    else r = make_float4(a * u.x, a * u.y, a * u.z, a * u.w);
    out[i] = r;
  }
}
__global__ void relu_fwd4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 u = x[i];
    y[i] = make_float4(fmaxf(u.x, 0.f), fmaxf(u.y, 0.f), fmaxf(u.z, 0.f), fmaxf(u.w, 0.f));
  }
}
__global__ void relu_bwd4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, float4* __restrict__ dx, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 u = x[i], d = dy[i];
    dx[i] = make_float4(u.x > 0.f ? d.x : 0.f, u.y > 0.f ? d.y : 0.f, u.z > 0.f ? d.z : 0.f, u.w > 0.f ? d.w : 0.f);
  }
}
__global__ void fill_kernel(float* x, float v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = fmaxf(x[i], 0.f);
}
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dx[i] = x[i] > 0.f ? dy[i] : 0.f;
}
// 3xTF32 operand split: hi = nearest tf32 of x, lo = nearest tf32 of (x - hi).  x - hi is exact in fp32, |lo| <= 2^-11 |x| and what
// the pair drops is <= 2^-22 |x|, unbiased (round-to-nearest both times: the MMA itself would truncate).
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    float h = rna_tf32_pw(v);
    hi[i] = h;
    lo[i] = rna_tf32_pw(v - h);
  }
}

// y[a][c] (+)= sum_r x[a][r][c]; grid (c blocks, r chunks, a)
__global__ void reduce_mid_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t R, int64_t C, int64_t chunk) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  int64_t a = blockIdx.z;
  int64_t r0 = (int64_t)blockIdx.y * chunk, r1 = r0 + chunk < R ? r0 + chunk : R;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += x[(a * R + r) * C + c];
  atomicAdd(&y[a * C + c], s);
}

// ------------------------------------------------------------------ pool / upsample
// y[n,do,ho,wo,c] = scale * sum_{window} x ; output dims floor(D/pd) ...
__global__ void pool_sum_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int D, int H, int W, int C,
                                int pd, int ph, int pw, float scale) {
  int Do = D / pd, Ho = H / ph, Wo = W / pw;
  int64_t total = (int64_t)N * Do * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = i % C; int64_t r = i / C;
    int wo = r % Wo; r /= Wo;
    int ho = r % Ho; r /= Ho;
    int dd = r % Do; int n = r / Do;
    float s = 0.f;
    for (int a = 0; a < pd; ++a)
      for (int b = 0; b < ph; ++b)
        for (int e = 0; e < pw; ++e)
          s += x[((((int64_t)n * D + dd * pd + a) * H + ho * ph + b) * W + wo * pw + e) * C + c];
    y[i] = s * scale;
  }
}
// single-channel images (the radar frames both discriminators average-pool first, ref: dgmr/discriminators.py:108,199): window 1x2x2, rows = N*D*H.
// One thread per TWO output pixels: two float4 loads (rows 2h, 2h+1), one float2 store -- the generic kernel's four scalar loads and three 64-bit
// div/mod pairs per pixel ran at a quarter of the memory rate.
__global__ void pool_sum_c1_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows_out, int W, float scale) {
  const int Wq = W >> 2;                       // float4 groups per input row = output pixel pairs per output row
  const int64_t total = rows_out * Wq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Wq; const int q = (int)(i - r * Wq);
    const float4 a = reinterpret_cast<const float4*>(x + (2 * r) * W)[q], b = reinterpret_cast<const float4*>(x + (2 * r + 1) * W)[q];
    reinterpret_cast<float2*>(y + r * (W >> 1))[q] = make_float2((a.x + a.y + b.x + b.y) * scale, (a.z + a.w + b.z + b.w) * scale);
  }
}
// its backward: x [rows, W] -> y [2*rows, 2W], every input pixel replicated 2x2 (times scale)
__global__ void upsample_c1_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows_in, int W, float scale) {
  const int Wh = W >> 1;                       // float2 groups per input row
  const int64_t total = rows_in * Wh;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Wh; const int q = (int)(i - r * Wh);
    const float2 v = reinterpret_cast<const float2*>(x + r * W)[q];
    const float4 o = make_float4(v.x * scale, v.x * scale, v.y * scale, v.y * scale);
    reinterpret_cast<float4*>(y + (2 * r) * (2 * W))[q] = o;
    reinterpret_cast<float4*>(y + (2 * r + 1) * (2 * W))[q] = o;
  }
}
// y[n,do,ho,wo,c] = scale * x[n,do/ud,ho/uh,wo/uw,c] if inside x's replicated extent else 0
__global__ void upsample_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int D, int H, int W, int C,
                                int ud, int uh, int uw, int Do, int Ho, int Wo, float scale) {
  int64_t total = (int64_t)N * Do * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = i % C; int64_t r = i / C;
    int wo = r % Wo; r /= Wo;
    int ho = r % Ho; r /= Ho;
    int dd = r % Do; int n = r / Do;
    int ds = dd / ud, hs = ho / uh, ws = wo / uw;
    float v = 0.f;
    if (ds < D && hs < H && ws < W) v = scale * x[((((int64_t)n * D + ds) * H + hs) * W + ws) * C + c];
    y[i] = v;
  }
}

// I = uint32_t when every index fits 32 bits (a 64-bit div/mod costs ~4x a 32-bit one and these kernels do several per element)
template <typename I>
__global__ void pool_sum4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int D, int H, int W, int C4,
                                 int pd, int ph, int pw, float scale) {
  int Do = D / pd, Ho = H / ph, Wo = W / pw;
  const I total = (I)N * Do * Ho * Wo * C4;
  for (I i = blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    int c = i % C4; I r = i / C4;
    int wo = r % Wo; r /= Wo;
    int ho = r % Ho; r /= Ho;
    int dd = r % Do; int n = r / Do;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < pd; ++a)
      for (int b = 0; b < ph; ++b)
        for (int e = 0; e < pw; ++e) {
          float4 v = x[((((I)n * D + dd * pd + a) * H + ho * ph + b) * W + wo * pw + e) * C4 + c];
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    y[i] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
  }
}
// I = uint32_t when every index fits 32 bits (a 64-bit div/mod costs ~4x a 32-bit one and these kernels do several per element)
template <typename I>
__global__ void upsample4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int D, int H, int W, int C4,
                                 int ud, int uh, int uw, int Do, int Ho, int Wo, float scale) {
  const I total = (I)N * Do * Ho * Wo * C4;
  for (I i = blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    int c = i % C4; I r = i / C4;
    int wo = r % Wo; r /= Wo;
    int ho = r % Ho; r /= Ho;
    int dd = r % Do; int n = r / Do;
    int ds = dd / ud, hs = ho / uh, ws = wo / uw;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ds < D && hs < H && ws < W) {
      float4 u = x[((((I)n * D + ds) * H + hs) * W + ws) * C4 + c];
      v = make_float4(u.x * scale, u.y * scale, u.z * scale, u.w * scale);
    }
    y[i] = v;
  }
}

// ------------------------------------------------------------------ ConvGRU gates (ref: dgmr/layers/ConvGRU.py:72-82)
__device__ __forceinline__ float sigmoid_acc(float v) { return 1.0f / (1.0f + expf(-v)); }   // full-accuracy expf (parity with the oracle)
// xr / xu / xc (nullable, same pitch as the pre-activation they belong to): the input-dependent part of the pre-activation, added HERE and the sum
// written back (the backward reads the complete pre-activation) -- the tap-split convolution then accumulates into a zeroed buffer and the
// per-step copy of the x part (one launch per convolution and step) disappears
__global__ void gru_gate_fwd_kernel(float* __restrict__ pre_r, int ld, const float* __restrict__ xr, const float* __restrict__ h, float* __restrict__ rh, int64_t rows, int Ch, int rnd) {
  int64_t total = rows * Ch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / Ch; int c = i - r * Ch;
    float p = pre_r[r * ld + c];
    if (xr) { p += xr[r * ld + c]; pre_r[r * ld + c] = p; }
    float v = sigmoid_acc(p) * h[i];
    rh[i] = rnd ? rna_tf32_pw(v) : v;
  }
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__global__ void gru_gate_fwd4_kernel(float* __restrict__ pre_r, int ld, const float* __restrict__ xr, const float4* __restrict__ h, float4* __restrict__ rh, int64_t rows, int C4, int rnd) {
  int64_t total = rows * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C4; int c = i - r * C4;
    float4 p = *reinterpret_cast<const float4*>(pre_r + r * ld + 4 * c), hv = h[i];
    if (xr) { p = add4(p, *reinterpret_cast<const float4*>(xr + r * ld + 4 * c)); *reinterpret_cast<float4*>(pre_r + r * ld + 4 * c) = p; }
    float4 v = make_float4(sigmoid_acc(p.x) * hv.x, sigmoid_acc(p.y) * hv.y, sigmoid_acc(p.z) * hv.z, sigmoid_acc(p.w) * hv.w);
    if (rnd) v = make_float4(rna_tf32_pw(v.x), rna_tf32_pw(v.y), rna_tf32_pw(v.z), rna_tf32_pw(v.w));
    rh[i] = v;
  }
}
__global__ void gru_blend_fwd_kernel(float* __restrict__ pre_u, int ld, const float* __restrict__ xu, const float* __restrict__ h, float* __restrict__ c_,
                                     const float* __restrict__ xc, float* __restrict__ hn, float* __restrict__ hn_tf32, int64_t rows, int Ch, int relu_c) {
  int64_t total = rows * Ch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / Ch; int c = i - r * Ch;
    float p = pre_u[r * ld + c];
    if (xu) { p += xu[r * ld + c]; pre_u[r * ld + c] = p; }
    float u = sigmoid_acc(p);
    float cv = c_[i];
    if (xc) { cv += xc[i]; c_[i] = cv; }
    if (relu_c) cv = fmaxf(cv, 0.f);
    float v = u * h[i] + (1.0f - u) * cv;
    hn[i] = v;
    if (hn_tf32) hn_tf32[i] = rna_tf32_pw(v);
  }
}
__global__ void gru_blend_fwd4_kernel(float* __restrict__ pre_u, int ld, const float* __restrict__ xu, const float4* __restrict__ h, float4* __restrict__ c_,
                                      const float4* __restrict__ xc, float4* __restrict__ hn, float4* __restrict__ hn_tf32, int64_t rows, int C4, int relu_c) {
  int64_t total = rows * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C4; int c = i - r * C4;
    float4 p = *reinterpret_cast<const float4*>(pre_u + r * ld + 4 * c), hv = h[i], cv = c_[i];
    if (xu) { p = add4(p, *reinterpret_cast<const float4*>(xu + r * ld + 4 * c)); *reinterpret_cast<float4*>(pre_u + r * ld + 4 * c) = p; }
    if (xc) { cv = add4(cv, xc[i]); c_[i] = cv; }
    if (relu_c) cv = make_float4(fmaxf(cv.x, 0.f), fmaxf(cv.y, 0.f), fmaxf(cv.z, 0.f), fmaxf(cv.w, 0.f));
    float ux = sigmoid_acc(p.x), uy = sigmoid_acc(p.y), uz = sigmoid_acc(p.z), uw = sigmoid_acc(p.w);
    float4 v = make_float4(ux * hv.x + (1.0f - ux) * cv.x, uy * hv.y + (1.0f - uy) * cv.y, uz * hv.z + (1.0f - uz) * cv.z, uw * hv.w + (1.0f - uw) * cv.w);
    hn[i] = v;
    if (hn_tf32) hn_tf32[i] = make_float4(rna_tf32_pw(v.x), rna_tf32_pw(v.y), rna_tf32_pw(v.z), rna_tf32_pw(v.w));
  }
}
// sc / dz (nullable): additionally write dz = d_pre * sc[c] (tf32-rounded if rnd) -- the operand of the recurrent convolution's dgrad / wgrad, so
// that the backward walk needs no separate prologue pass per step (its per-step scale gradients are reduced in one grouped pass afterwards)
__global__ void gru_gate_bwd_kernel(const float* __restrict__ d_rh, const float* __restrict__ pre_r, int ld, const float* __restrict__ h,
                                    float* __restrict__ d_pre_r, int ldd, float* __restrict__ dh, int acc, int64_t rows, int Ch,
                                    const float* __restrict__ sc, float* __restrict__ dz, int rnd) {
  int64_t total = rows * Ch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / Ch; int c = i - r * Ch;
    float g = 1.0f / (1.0f + expf(-pre_r[r * ld + c]));
    float d = d_rh[i];
    const float dp = d * h[i] * g * (1.0f - g);
    d_pre_r[r * ldd + c] = dp;
    if (dz) { const float z = dp * sc[c]; dz[r * ldd + c] = rnd ? rna_tf32_pw(z) : z; }
    float v = d * g;
    if (acc) dh[i] += v; else dh[i] = v;
  }
}
__global__ void gru_blend_bwd_kernel(const float* __restrict__ d_hn, const float* __restrict__ pre_u, int ld, const float* __restrict__ h, const float* __restrict__ c_,
                                     float* __restrict__ d_pre_u, int ldd, float* __restrict__ dc, float* __restrict__ dh, int acc, int64_t rows, int Ch, int relu_c,
                                     const float* __restrict__ sc_u, float* __restrict__ dz_u, const float* __restrict__ sc_c, float* __restrict__ dz_c, int rnd) {
  int64_t total = rows * Ch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / Ch; int c = i - r * Ch;
    float u = 1.0f / (1.0f + expf(-pre_u[r * ld + c]));
    float d = d_hn[i];
    float cp = c_[i];
    float cv = relu_c ? fmaxf(cp, 0.f) : cp;
    const float dpu = d * (h[i] - cv) * u * (1.0f - u);
    d_pre_u[r * ldd + c] = dpu;
    if (dz_u) { const float z = dpu * sc_u[c]; dz_u[r * ldd + c] = rnd ? rna_tf32_pw(z) : z; }
    const float dcv = (relu_c && !(cp > 0.f)) ? 0.f : d * (1.0f - u);
    dc[i] = dcv;
    if (dz_c) { const float z = dcv * sc_c[c]; dz_c[i] = rnd ? rna_tf32_pw(z) : z; }
    float v = d * u;
    if (acc) dh[i] += v; else dh[i] = v;
  }
}

// ------------------------------------------------------------------ BatchNorm
// x: [G*rows, C].  Block = 256 threads: cpl = min(C,256) channel lanes x rp row lanes.
// Each block reduces `chunk` rows of one group and atomically adds (double) into sums[g][c][0..1].
__global__ void bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, int64_t rows, int C, int64_t chunk) {
  extern __shared__ double sh[];  // [rp][cpl][2]
  int g = blockIdx.y;
  int cpl = C < 256 ? C : 256;
  int rp = 256 / cpl;
  int cl = threadIdx.x % cpl, rl = threadIdx.x / cpl;
  int64_t r0 = (int64_t)blockIdx.x * chunk;
  int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
  const float* xg = x + (int64_t)g * rows * C;
  for (int c = cl; c < C; c += cpl) {
    double s = 0.0, q = 0.0;
    if (rl < rp) {
      float fs = 0.f, fq = 0.f;
      int cnt = 0;
      for (int64_t r = r0 + rl; r < r1; r += rp) {
        float v = xg[r * C + c];
        fs += v; fq += v * v;
        if (++cnt == 64) { s += fs; q += fq; fs = fq = 0.f; cnt = 0; }
      }
      s += fs; q += fq;
      sh[(rl * cpl + cl) * 2 + 0] = s;
      sh[(rl * cpl + cl) * 2 + 1] = q;
    }
    __syncthreads();
    if (rl == 0) {
      for (int j = 1; j < rp; ++j) { s += sh[(j * cpl + cl) * 2]; q += sh[(j * cpl + cl) * 2 + 1]; }
      atomicAdd(&sums[((int64_t)g * C + c) * 2 + 0], s);
      atomicAdd(&sums[((int64_t)g * C + c) * 2 + 1], q);
    }
    __syncthreads();
  }
}
// float4 variant (C % 4 == 0): thread = (row lane, 4-channel lane), 4 independent rows in flight per thread.
__global__ void bn_stats4_kernel(const float4* __restrict__ x, double* __restrict__ sums, int64_t rows, int C4, int64_t chunk) {
  extern __shared__ double sh[];  // [rp][cpl][8]
  const int g = blockIdx.y;
  const int cpl = C4 < 256 ? C4 : 256;
  const int rp = 256 / cpl;
  const int cl = threadIdx.x % cpl, rl = threadIdx.x / cpl;
  const int64_t r0 = (int64_t)blockIdx.x * chunk;
  const int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
  const float4* xg = x + (int64_t)g * rows * C4;
  for (int cb = 0; cb < C4; cb += cpl) {   // block-uniform trip count (barriers inside)
    const int c = cb + cl;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rl < rp && c < C4) {
      float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
      int cnt = 0;
      int64_t r = r0 + rl;
      for (; r + 3 * (int64_t)rp < r1; r += 4 * (int64_t)rp) {
        float4 v0 = xg[r * C4 + c], v1 = xg[(r + rp) * C4 + c], v2 = xg[(r + 2 * (int64_t)rp) * C4 + c], v3 = xg[(r + 3 * (int64_t)rp) * C4 + c];
        fs[0] += (v0.x + v1.x) + (v2.x + v3.x); fq[0] += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x);
        fs[1] += (v0.y + v1.y) + (v2.y + v3.y); fq[1] += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
        fs[2] += (v0.z + v1.z) + (v2.z + v3.z); fq[2] += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z);
        fs[3] += (v0.w + v1.w) + (v2.w + v3.w); fq[3] += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
        if (++cnt == 16) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { acc[2 * k] += fs[k]; acc[2 * k + 1] += fq[k]; fs[k] = fq[k] = 0.f; }
          cnt = 0;
        }
      }
      for (; r < r1; r += rp) {
        float4 v = xg[r * C4 + c];
        fs[0] += v.x; fq[0] += v.x * v.x; fs[1] += v.y; fq[1] += v.y * v.y; fs[2] += v.z; fq[2] += v.z * v.z; fs[3] += v.w; fq[3] += v.w * v.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc[2 * k] += fs[k]; acc[2 * k + 1] += fq[k]; }
#pragma unroll
      for (int k = 0; k < 8; ++k) sh[(rl * cpl + cl) * 8 + k] = acc[k];
    }
    __syncthreads();
    // 8 values per 4-channel lane: spread the final cross-row reduction over 8*cpl threads
    for (int t = threadIdx.x; t < cpl * 8; t += blockDim.x) {
      const int lane4 = t >> 3, k = t & 7;
      if (lane4 + cb < C4) {
        double v = 0.0;
        for (int j = 0; j < rp; ++j) v += sh[(j * cpl + lane4) * 8 + k];
        atomicAdd(&sums[((int64_t)g * C4 * 4 + (int64_t)(cb + lane4) * 4 + (k >> 1)) * 2 + (k & 1)], v);
      }
    }
    __syncthreads();
  }
}
// block = 32 channels x 8 group lanes: the per-(group, channel) statistics (fp64 divisions) of all G groups are formed in parallel, then one lane per
// channel walks the running-statistics recurrence in group order (the reference's G successive module calls) over the values parked in shared memory
__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, int64_t rows, int G, int C, float eps, float mom, int training,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ a, float* __restrict__ b) {
  extern __shared__ float fin_sh[];                      // [G][32] batch means, [G][32] unbiased variances
  const int cl = threadIdx.x & 31, gl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool live = c < C;
  const float ga = (live && gamma) ? gamma[c] : 1.f, be = (live && beta) ? beta[c] : 0.f;
  const float rm0 = live ? rmean[c] : 0.f, rv0 = live ? rvar[c] : 1.f;
  if (live) {
    for (int g = gl; g < G; g += 8) {
      float m, is;
      if (training) {
        const double s = sums[((int64_t)g * C + c) * 2], q = sums[((int64_t)g * C + c) * 2 + 1];
        const double md = s / (double)rows;
        double vd = q / (double)rows - md * md;
        if (vd < 0) vd = 0;
        m = (float)md;
        const float var = (float)vd;
        is = 1.0f / sqrtf(var + eps);
        fin_sh[g * 32 + cl] = m;
        fin_sh[(G + g) * 32 + cl] = rows > 1 ? (float)(vd * (double)rows / (double)(rows - 1)) : var;
      } else {
        m = rm0;
        is = 1.0f / sqrtf(rv0 + eps);
      }
      const int64_t o = (int64_t)g * C + c;
      mean[o] = m; invstd[o] = is;
      const float aa = ga * is;
      a[o] = aa; b[o] = be - m * aa;
    }
  }
  if (!training) return;
  __syncthreads();
  if (live && gl == 0) {
    float rm = rm0, rv = rv0;
    for (int g = 0; g < G; ++g) {
      rm = (1.f - mom) * rm + mom * fin_sh[g * 32 + cl];
      rv = (1.f - mom) * rv + mom * fin_sh[(G + g) * 32 + cl];
    }
    rmean[c] = rm; rvar[c] = rv;
  }
}
// y = act(a*x+b), optional nearest x2 upsample on write.  One thread per output element.
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                int64_t rows, int G, int C, int relu, int up2, int H, int W, int rnd, float* __restrict__ xr) {
  int64_t orows = up2 ? rows * 4 : rows;
  int64_t total = (int64_t)G * orows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = i % C; int64_t r = i / C;  // global output row
    int g = r / orows;
    int64_t xrow = r;
    if (up2) {
      int Wo = 2 * W, Ho = 2 * H;
      int wo = r % Wo; int64_t t = r / Wo;
      int ho = t % Ho; int64_t n = t / Ho;
      xrow = (n * H + (ho >> 1)) * W + (wo >> 1);
    }
    const float u = x[xrow * C + c];
    float v = a[(int64_t)g * C + c] * u + b[(int64_t)g * C + c];
    v = relu ? fmaxf(v, 0.f) : v;
    y[i] = rnd ? rna_tf32_pw(v) : v;
    if (xr) xr[i] = rna_tf32_pw(u);      // (never with up2: host-checked)
  }
}
// I = uint32_t when every index fits 32 bits (a 64-bit div/mod costs ~4x a 32-bit one and these kernels do several per element)
template <typename I>
__global__ void bn_apply4_kernel(const float4* __restrict__ x, const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y,
                                 int64_t rows, int G, int C4, int relu, int up2, int H, int W, int rnd, float4* __restrict__ xr4) {
  const I orows = (I)(up2 ? rows * 4 : rows);
  const I total = (I)G * orows * C4;
  for (I i = blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    int c = i % C4; I r = i / C4;
    int g = r / orows;
    I xr = r;
    if (up2) {
      int Wo = 2 * W, Ho = 2 * H;
      int wo = r % Wo; I t = r / Wo;
      int ho = t % Ho; I n = t / Ho;
      xr = (n * H + (ho >> 1)) * W + (wo >> 1);
    }
    float4 aa = a[g * C4 + c], bb = b[g * C4 + c], u = x[xr * C4 + c];
    float4 v = make_float4(aa.x * u.x + bb.x, aa.y * u.y + bb.y, aa.z * u.z + bb.z, aa.w * u.w + bb.w);
    if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    if (rnd) v = make_float4(rna_tf32_pw(v.x), rna_tf32_pw(v.y), rna_tf32_pw(v.z), rna_tf32_pw(v.w));
    y[i] = v;
    if (xr4) xr4[i] = make_float4(rna_tf32_pw(u.x), rna_tf32_pw(u.y), rna_tf32_pw(u.z), rna_tf32_pw(u.w));      // (never with up2: host-checked)
  }
}
// dpre at low-res row r, channel c (sums the 4 replicas if up2, applies relu mask)
__device__ __forceinline__ float bn_dpre(const float* __restrict__ dy, int64_t r, int c, int C, int up2, int H, int W, float yv, int relu) {
  float d;
  if (up2) {
    int w = r % W; int64_t t = r / W; int h = t % H; int64_t n = t / H;
    int64_t base = ((n * 2 * H + 2 * h) * (2 * W) + 2 * w) * C + c;
    d = dy[base] + dy[base + C] + dy[base + (int64_t)2 * W * C] + dy[base + (int64_t)2 * W * C + C];
  } else {
    d = dy[r * C + c];
  }
  if (relu && !(yv > 0.f)) d = 0.f;
  return d;
}
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                     const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ red,
                                     int64_t rows, int C, int64_t chunk, int relu, int up2, int H, int W) {
  extern __shared__ double sh[];
  int g = blockIdx.y;
  int cpl = C < 256 ? C : 256;
  int rp = 256 / cpl;
  int cl = threadIdx.x % cpl, rl = threadIdx.x / cpl;
  int64_t r0 = (int64_t)blockIdx.x * chunk;
  int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
  for (int c = cl; c < C; c += cpl) {
    double s = 0.0, q = 0.0;
    if (rl < rp) {
      float aa = a[(int64_t)g * C + c], bb = b[(int64_t)g * C + c], m = mean[(int64_t)g * C + c], is = invstd[(int64_t)g * C + c];
      float fs = 0.f, fq = 0.f; int cnt = 0;
      for (int64_t r = r0 + rl; r < r1; r += rp) {
        int64_t gr = (int64_t)g * rows + r;
        float xv = x[gr * C + c];
        float d = bn_dpre(dy, gr, c, C, up2, H, W, aa * xv + bb, relu);
        fs += d; fq += d * (xv - m) * is;
        if (++cnt == 64) { s += fs; q += fq; fs = fq = 0.f; cnt = 0; }
      }
      s += fs; q += fq;
      sh[(rl * cpl + cl) * 2] = s; sh[(rl * cpl + cl) * 2 + 1] = q;
    }
    __syncthreads();
    if (rl == 0) {
      for (int j = 1; j < rp; ++j) { s += sh[(j * cpl + cl) * 2]; q += sh[(j * cpl + cl) * 2 + 1]; }
      atomicAdd(&red[((int64_t)g * C + c) * 2], s);
      atomicAdd(&red[((int64_t)g * C + c) * 2 + 1], q);
    }
    __syncthreads();
  }
}
// float4 variant.  Thread = (row lane, 4-channel lane) with its per-channel constants in registers; U rows of loads in flight per thread;
// fp32 partials are flushed into fp64 slots in SHARED memory every 64 rows (keeps the register count at 3-4 CTAs per SM).
template <bool UP2>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce4_kernel(const float* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ a,
                                                                const float4* __restrict__ b, const float4* __restrict__ mean, const float4* __restrict__ invstd,
                                                                double* __restrict__ red, int64_t rows, int C4, int64_t chunk, int relu, int H, int W) {
  extern __shared__ double sh[];  // [256][8]
  const int g = blockIdx.y;
  const int C = C4 * 4;
  const int cpl = C4 < 256 ? C4 : 256;
  const int rp = 256 / cpl;
  const int cl = threadIdx.x % cpl, rl = threadIdx.x / cpl;
  const int64_t r0 = (int64_t)blockIdx.x * chunk;
  const int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
  double* my = sh + (size_t)threadIdx.x * 8;
  for (int cb = 0; cb < C4; cb += cpl) {   // block-uniform trip count (barriers inside)
    const int c = cb + cl;
#pragma unroll
    for (int k = 0; k < 8; ++k) my[k] = 0.0;
    if (rl < rp && c < C4) {
      const int64_t o4 = (int64_t)g * C4 + c;
      const float4 aa = a[o4], bb = b[o4], m = mean[o4], is = invstd[o4];
      float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
      struct Row { float4 xv, d; };
      auto load = [&](int64_t r, Row& w) {
        const int64_t gr = (int64_t)g * rows + r;
        w.xv = x[gr * C4 + c];
        if (UP2) {
          const uint32_t gu = (uint32_t)gr;               // host checks G*rows < 2^31 for the upsampled form
          const uint32_t wq = gu % (uint32_t)W, t = gu / (uint32_t)W, hq = t % (uint32_t)H, n = t / (uint32_t)H;
          const float4* p0 = reinterpret_cast<const float4*>(dy + (((int64_t)n * 2 * H + 2 * hq) * (2 * (int64_t)W) + 2 * wq) * C) + c;
          const float4* p1 = p0 + (int64_t)2 * W * C4;
          const float4 q0 = p0[0], q1 = p0[C4], q2 = p1[0], q3 = p1[C4];
          w.d = make_float4(q0.x + q1.x + q2.x + q3.x, q0.y + q1.y + q2.y + q3.y, q0.z + q1.z + q2.z + q3.z, q0.w + q1.w + q2.w + q3.w);
        } else {
          w.d = reinterpret_cast<const float4*>(dy)[gr * C4 + c];
        }
      };
      auto finish = [&](Row& w) {
        float4 d = w.d; const float4 xv = w.xv;
        if (relu) {
          if (!(aa.x * xv.x + bb.x > 0.f)) d.x = 0.f;
          if (!(aa.y * xv.y + bb.y > 0.f)) d.y = 0.f;
          if (!(aa.z * xv.z + bb.z > 0.f)) d.z = 0.f;
          if (!(aa.w * xv.w + bb.w > 0.f)) d.w = 0.f;
        }
        fs[0] += d.x; fq[0] += d.x * (xv.x - m.x) * is.x;
        fs[1] += d.y; fq[1] += d.y * (xv.y - m.y) * is.y;
        fs[2] += d.z; fq[2] += d.z * (xv.z - m.z) * is.z;
        fs[3] += d.w; fq[3] += d.w * (xv.w - m.w) * is.w;
      };
      auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < 4; ++k) { my[2 * k] += (double)fs[k]; my[2 * k + 1] += (double)fq[k]; fs[k] = fq[k] = 0.f; }
      };
      constexpr int U = UP2 ? 2 : 4;
      int cnt = 0;
      int64_t r = r0 + rl;
      if (!UP2) {
        // pointer-bumped main loop: one 64-bit pointer pair and a 32-bit stride instead of per-row 64-bit index arithmetic
        const float4* xp = x + ((int64_t)g * rows + r) * C4 + c;
        const float4* dp = reinterpret_cast<const float4*>(dy) + ((int64_t)g * rows + r) * C4 + c;
        const int step = rp * C4;
        for (; r + (int64_t)(U - 1) * rp < r1; r += (int64_t)U * rp, xp += (int64_t)U * step, dp += (int64_t)U * step) {
          Row w[U];
#pragma unroll
          for (int u = 0; u < U; ++u) { w[u].xv = xp[u * step]; w[u].d = dp[u * step]; }
#pragma unroll
          for (int u = 0; u < U; ++u) finish(w[u]);
          if (++cnt == 64 / U) { flush(); cnt = 0; }
        }
      }
      if (UP2) {
        for (; r + (int64_t)(U - 1) * rp < r1; r += (int64_t)U * rp) {
          Row w[U];
#pragma unroll
          for (int u = 0; u < U; ++u) load(r + (int64_t)u * rp, w[u]);
#pragma unroll
          for (int u = 0; u < U; ++u) finish(w[u]);
          if (++cnt == 64 / U) { flush(); cnt = 0; }
        }
      }
      for (; r < r1; r += rp) { Row w; load(r, w); finish(w); }
      flush();
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cpl * 8; t += blockDim.x) {
      const int lane4 = t >> 3, k = t & 7;
      if (lane4 + cb < C4) {
        double v = 0.0;
        for (int j = 0; j < rp; ++j) v += sh[(size_t)(j * cpl + lane4) * 8 + k];
        atomicAdd(&red[((int64_t)g * C + (int64_t)(cb + lane4) * 4 + (k >> 1)) * 2 + (k & 1)], v);
      }
    }
    __syncthreads();
  }
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, const double* __restrict__ red, float* __restrict__ dx,
                                    int64_t rows, int G, int C, int relu, int up2, int H, int W, int training, const float* __restrict__ oscale, int rnd,
                                    const float* __restrict__ dx_add) {
  int64_t total = (int64_t)G * rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = i % C; int64_t r = i / C; int g = r / rows;
    int64_t o = (int64_t)g * C + c;
    float aa = a[o], xv = x[i];
    float d = bn_dpre(dy, r, c, C, up2, H, W, aa * xv + b[o], relu);
    float v;
    if (training) {
      float xh = (xv - mean[o]) * invstd[o];
      float m1 = (float)(red[o * 2] / (double)rows), m2 = (float)(red[o * 2 + 1] / (double)rows);
      v = aa * (d - m1 - xh * m2);
    } else {
      v = aa * d;
    }
    if (oscale) v *= oscale[o];
    if (dx_add) v += dx_add[i];
    dx[i] = rnd ? rna_tf32_pw(v) : v;
  }
}
// float4 variant (C % 4 == 0): same thread map as the reduce kernel (grid = row chunks x G, thread = row lane x 4-channel lane), so the per-channel
// constants (a, b, mean, invstd, the two means of the reduce pass, the output scale) are loaded ONCE per thread instead of once per element, and
// U rows of loads are in flight per thread.
template <bool UP2, bool ADD>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply4_kernel(const float* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ a,
                                                               const float4* __restrict__ b, const float4* __restrict__ mean, const float4* __restrict__ invstd,
                                                               const double* __restrict__ red, float4* __restrict__ dx, int64_t rows, int C4, int64_t chunk,
                                                               int relu, int H, int W, int training, const float4* __restrict__ oscale, int rnd,
                                                               const float4* __restrict__ dx_add) {
  const int g = blockIdx.y;
  const int C = C4 * 4;
  const int cpl = C4 < 256 ? C4 : 256;
  const int rp = 256 / cpl;
  const int cl = threadIdx.x % cpl, rl = threadIdx.x / cpl;
  if (rl >= rp) return;
  const int64_t r0 = (int64_t)blockIdx.x * chunk;
  const int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
  for (int c = cl; c < C4; c += cpl) {
    const int64_t o4 = (int64_t)g * C4 + c;
    const float4 aa = a[o4], bb = b[o4];
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), is = m, m1 = m, m2 = m, os = make_float4(1.f, 1.f, 1.f, 1.f);
    if (training) {
      m = mean[o4]; is = invstd[o4];
      const double inv = 1.0 / (double)rows;
      const double* rq = red + o4 * 8;
      m1 = make_float4((float)(rq[0] * inv), (float)(rq[2] * inv), (float)(rq[4] * inv), (float)(rq[6] * inv));
      m2 = make_float4((float)(rq[1] * inv), (float)(rq[3] * inv), (float)(rq[5] * inv), (float)(rq[7] * inv));
    }
    if (oscale) os = oscale[o4];
    struct Row { float4 xv, d, ad; int64_t i; };
    auto load = [&](int64_t r, Row& w) {
      const int64_t gr = (int64_t)g * rows + r;
      w.i = gr * C4 + c;
      w.xv = x[w.i];
      if (UP2) {
        const uint32_t gu = (uint32_t)gr;               // host checks G*rows < 2^31 for the upsampled form
        const uint32_t wq = gu % (uint32_t)W, t = gu / (uint32_t)W, hq = t % (uint32_t)H, n = t / (uint32_t)H;
        const float4* p0 = reinterpret_cast<const float4*>(dy + (((int64_t)n * 2 * H + 2 * hq) * (2 * (int64_t)W) + 2 * wq) * C) + c;
        const float4* p1 = p0 + (int64_t)2 * W * C4;
        const float4 q0 = p0[0], q1 = p0[C4], q2 = p1[0], q3 = p1[C4];
        w.d = make_float4(q0.x + q1.x + q2.x + q3.x, q0.y + q1.y + q2.y + q3.y, q0.z + q1.z + q2.z + q3.z, q0.w + q1.w + q2.w + q3.w);
      } else {
        w.d = reinterpret_cast<const float4*>(dy)[w.i];
      }
      if (ADD) w.ad = dx_add[w.i];
    };
    auto finish = [&](Row& w) {
      float4 d = w.d; const float4 xv = w.xv;
      if (relu) {
        if (!(aa.x * xv.x + bb.x > 0.f)) d.x = 0.f;
        if (!(aa.y * xv.y + bb.y > 0.f)) d.y = 0.f;
        if (!(aa.z * xv.z + bb.z > 0.f)) d.z = 0.f;
        if (!(aa.w * xv.w + bb.w > 0.f)) d.w = 0.f;
      }
      float4 v;
      if (training) {
        v.x = aa.x * (d.x - m1.x - (xv.x - m.x) * is.x * m2.x);
        v.y = aa.y * (d.y - m1.y - (xv.y - m.y) * is.y * m2.y);
        v.z = aa.z * (d.z - m1.z - (xv.z - m.z) * is.z * m2.z);
        v.w = aa.w * (d.w - m1.w - (xv.w - m.w) * is.w * m2.w);
      } else {
        v = make_float4(aa.x * d.x, aa.y * d.y, aa.z * d.z, aa.w * d.w);
      }
      if (oscale) { v.x *= os.x; v.y *= os.y; v.z *= os.z; v.w *= os.w; }
      if (ADD) { v.x += w.ad.x; v.y += w.ad.y; v.z += w.ad.z; v.w += w.ad.w; }
      if (rnd) v = make_float4(rna_tf32_pw(v.x), rna_tf32_pw(v.y), rna_tf32_pw(v.z), rna_tf32_pw(v.w));
      dx[w.i] = v;
    };
    constexpr int U = (UP2 || ADD) ? 2 : 4;
    int64_t r = r0 + rl;
    for (; r + (int64_t)(U - 1) * rp < r1; r += (int64_t)U * rp) {
      Row w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) load(r + (int64_t)u * rp, w[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) finish(w[u]);
    }
    for (; r < r1; r += rp) { Row w; load(r, w); finish(w); }
  }
}
__global__ void bn_bwd_params_kernel(const double* __restrict__ red, float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int C, int acc) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int g = 0; g < G; ++g) { s += red[((int64_t)g * C + c) * 2]; q += red[((int64_t)g * C + c) * 2 + 1]; }
  if (dgamma) { if (acc) dgamma[c] += (float)q; else dgamma[c] = (float)q; }
  if (dbeta) { if (acc) dbeta[c] += (float)s; else dbeta[c] = (float)s; }
}

// ------------------------------------------------------------------ D head
__global__ void sumpool_relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
  int64_t total = (int64_t)N * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = i % C; int64_t n = i / C;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += fmaxf(x[(n * HW + p) * C + c], 0.f);
    y[i] = s;
  }
}
__global__ void sumpool_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int N, int HW, int C) {
  int64_t total = (int64_t)N * HW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = i % C; int64_t n = i / ((int64_t)HW * C);
    dx[i] = x[i] > 0.f ? dy[n * C + c] : 0.f;
  }
}

// ------------------------------------------------------------------ attention (ref quirk: positions = (c,h), features = w)
// Q[l][j] = q[b, h, j, c] with l = c*H + h, channels-last storage q[((b*H+h)*W+j)*C + c]
__device__ __forceinline__ int64_t att_idx(int b, int l, int j, int H, int W, int C) {
  int c = l / H, h = l - c * H;
  return (((int64_t)b * H + h) * W + j) * C + c;
}
__global__ void attention_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, float* __restrict__ out,
                                     float* __restrict__ beta, int H, int W, int C) {
  extern __shared__ float shf[];  // logits[L] + qrow[W] + red[32]
  int L = C * H, l = blockIdx.x, b = blockIdx.y;
  float* logit = shf; float* qrow = shf + L; float* red = qrow + W;
  for (int j = threadIdx.x; j < W; j += blockDim.x) qrow[j] = q[att_idx(b, l, j, H, W, C)];
  __syncthreads();
  float mx = -INFINITY;
  for (int m = threadIdx.x; m < L; m += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < W; ++j) s += qrow[j] * k[att_idx(b, m, j, H, W, C)];
    logit[m] = s; mx = fmaxf(mx, s);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int m = threadIdx.x; m < L; m += blockDim.x) { float e = expf(logit[m] - mx); logit[m] = e; sum += e; }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
  float inv = 1.0f / sum;
  for (int m = threadIdx.x; m < L; m += blockDim.x) { float p = logit[m] * inv; logit[m] = p; beta[((int64_t)b * L + l) * L + m] = p; }
  __syncthreads();
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    float s = 0.f;
    for (int m = 0; m < L; ++m) s += logit[m] * v[att_idx(b, m, j, H, W, C)];
    out[att_idx(b, l, j, H, W, C)] = s;
  }
}
// per row l: dbeta, dlogit -> ws[b][l][:], dq row
__global__ void attention_bwd1_kernel(const float* __restrict__ dout, const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ beta,
                                      float* __restrict__ dq, float* __restrict__ ws, int H, int W, int C) {
  extern __shared__ float shf[];  // dl[L] + dorow[W] + red[32]
  int L = C * H, l = blockIdx.x, b = blockIdx.y;
  float* dl = shf; float* dorow = shf + L; float* red = dorow + W;
  for (int j = threadIdx.x; j < W; j += blockDim.x) dorow[j] = dout[att_idx(b, l, j, H, W, C)];
  __syncthreads();
  float dot = 0.f;
  for (int m = threadIdx.x; m < L; m += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < W; ++j) s += dorow[j] * v[att_idx(b, m, j, H, W, C)];
    dl[m] = s;
    dot += s * beta[((int64_t)b * L + l) * L + m];
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  dot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) dot += red[w];
  for (int m = threadIdx.x; m < L; m += blockDim.x) {
    float d = beta[((int64_t)b * L + l) * L + m] * (dl[m] - dot);
    dl[m] = d; ws[((int64_t)b * L + l) * L + m] = d;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    float s = 0.f;
    for (int m = 0; m < L; ++m) s += dl[m] * k[att_idx(b, m, j, H, W, C)];
    dq[att_idx(b, l, j, H, W, C)] = s;
  }
}
// per column m: dk[m][j] = sum_l dlogit[l][m] q[l][j]; dv[m][j] = sum_l beta[l][m] dout[l][j]
__global__ void attention_bwd2_kernel(const float* __restrict__ dout, const float* __restrict__ q, const float* __restrict__ beta, const float* __restrict__ ws,
                                      float* __restrict__ dk, float* __restrict__ dv, int H, int W, int C) {
  int L = C * H, m = blockIdx.x, b = blockIdx.y;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    float sk = 0.f, sv = 0.f;
    for (int l = 0; l < L; ++l) {
      sk += ws[((int64_t)b * L + l) * L + m] * q[att_idx(b, l, j, H, W, C)];
      sv += beta[((int64_t)b * L + l) * L + m] * dout[att_idx(b, l, j, H, W, C)];
    }
    dk[att_idx(b, m, j, H, W, C)] = sk;
    dv[att_idx(b, m, j, H, W, C)] = sv;
  }
}

// ------------------------------------------------------------------ losses
__global__ void hinge_disc_kernel(const float* __restrict__ s, int B, int cols, float* __restrict__ loss, float* __restrict__ ds) {
  // single block; scores [2B][cols]
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < 2 * B * cols; i += blockDim.x) {
    int row = i / cols;
    float v = s[i], l, g;
    if (row < B) { l = fmaxf(1.f - v, 0.f); g = (1.f - v) > 0.f ? -1.f / B : 0.f; }
    else { l = fmaxf(1.f + v, 0.f); g = (1.f + v) > 0.f ? 1.f / B : 0.f; }
    acc += l / B; ds[i] = g;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; *loss = t; }
}
__global__ void hinge_gen_kernel(const float* __restrict__ s, int n, float* __restrict__ loss, float* __restrict__ ds) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { acc += s[i]; ds[i] = -1.f / n; }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; *loss = -t / n; }
}
__global__ void grid_cell_fwd_kernel(const float* __restrict__ gen, const float* __restrict__ tgt, float cap, double* __restrict__ acc, int64_t n) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t = tgt[i];
    s += (double)fabsf((gen[i] - t) * fmaxf(t + 1.f, cap));
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; atomicAdd(acc, t); }
}
__global__ void grid_cell_final_kernel(const double* __restrict__ acc, float coef, float* __restrict__ loss) { *loss = (float)(*acc) * coef; }
__global__ void grid_cell_bwd_kernel(const float* __restrict__ gen, const float* __restrict__ tgt, float cap, float coef, const float* __restrict__ gout, float* __restrict__ dgen, int64_t n) {
  float go = *gout * coef;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t = tgt[i];
    float w = fmaxf(t + 1.f, cap);
    float d = (gen[i] - t) * w;
    dgen[i] = d > 0.f ? w * go : (d < 0.f ? -w * go : 0.f);
  }
}

// ------------------------------------------------------------------ Adam (torch.optim.Adam semantics)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gs) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gr = g[i] * gs;
    float mi = b1 * m[i] + (1.f - b1) * gr;
    float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

}  // namespace dgmr

using namespace dgmr;

extern "C" {

const char* dgmr_last_error(void) { return g_err; }
int dgmr_abi_version(void) { return 1; }

int dgmr_permute(const float* src, float* dst, int ndim, const int64_t* shape, const int64_t* sstr, const int64_t* dstr, int accumulate, dgmr_stream_t stream) {
  DGMR_REQUIRE(ndim >= 1 && ndim <= 8, "dgmr_permute: ndim %d out of range", ndim);
  PermuteArgs a; a.ndim = ndim;
  int64_t total = 1;
  for (int d = 0; d < 8; ++d) { a.shape[d] = 1; a.sstr[d] = 0; a.dstr[d] = 0; }
  for (int d = 0; d < ndim; ++d) { a.shape[d] = shape[d]; a.sstr[d] = sstr[d]; a.dstr[d] = dstr[d]; total *= shape[d]; }
  if (total == 0) return 0;
  int64_t max_s = 0, max_d = 0; bool nonneg = true;
  for (int d = 0; d < ndim; ++d) { max_s += (shape[d] - 1) * sstr[d]; max_d += (shape[d] - 1) * dstr[d]; nonneg = nonneg && sstr[d] >= 0 && dstr[d] >= 0; }
  const int64_t lim = (int64_t)1 << 31;
  // float4 form: innermost dimension contiguous on both sides, a multiple of 4 long, every other stride and both base pointers 16-byte aligned
  bool vec = ndim >= 1 && sstr[ndim - 1] == 1 && dstr[ndim - 1] == 1 && shape[ndim - 1] % 4 == 0 && al16(src) && al16(dst);
  for (int d = 0; d + 1 < ndim && vec; ++d) vec = sstr[d] % 4 == 0 && dstr[d] % 4 == 0;
  if (vec) {
    a.shape[ndim - 1] /= 4;
    for (int d = 0; d + 1 < ndim; ++d) { a.sstr[d] /= 4; a.dstr[d] /= 4; }
    total /= 4;
    if (nonneg && max_s < lim && max_d < lim)
      permute_kernel<uint32_t, float4><<<ew_grid(total, 256, 1), 256, 0, S(stream)>>>((const float4*)src, (float4*)dst, a, total, accumulate);
    else
      permute_kernel<int64_t, float4><<<ew_grid(total, 256, 1), 256, 0, S(stream)>>>((const float4*)src, (float4*)dst, a, total, accumulate);
  } else if (nonneg && total < lim && max_s < lim && max_d < lim) permute_kernel<uint32_t, float><<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(src, dst, a, total, accumulate);
  else permute_kernel<int64_t, float><<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(src, dst, a, total, accumulate);
  DGMR_CHECK_LAUNCH("dgmr_permute");
  return 0;
}
int dgmr_reduce_mid(const float* x, float* y, int64_t A, int64_t R, int64_t C, int accumulate, dgmr_stream_t stream) {
  if (A * C == 0) return 0;
  if (!accumulate) DGMR_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)(A * C), S(stream)));
  if (R == 0) return 0;
  DGMR_REQUIRE(A <= 65535, "dgmr_reduce_mid: A too large");
  int64_t cb = ceil_div(C, 256);
  int64_t want = ceil_div((int64_t)sm_count() * 4, cb * A);
  if (want < 1) want = 1;
  int64_t chunk = ceil_div(R, want); if (chunk < 8) chunk = 8;
  int64_t rb = ceil_div(R, chunk); if (rb > 65535) { rb = 65535; chunk = ceil_div(R, rb); rb = ceil_div(R, chunk); }
  reduce_mid_kernel<<<dim3((unsigned)cb, (unsigned)rb, (unsigned)A), 256, 0, S(stream)>>>(x, y, R, C, chunk);
  DGMR_CHECK_LAUNCH("dgmr_reduce_mid");
  return 0;
}
int dgmr_axpby(float a, const float* x, float b, const float* y, float* out, int64_t n, dgmr_stream_t stream) {
  if (n == 0) return 0;
  if (n % 4 == 0 && al16(x) && al16(out) && (!y || al16(y)))
    axpby4_kernel<<<ew_grid(n / 4, 256, 2), 256, 0, S(stream)>>>(a, (const float4*)x, b, (const float4*)y, (float4*)out, n / 4);
  else
    axpby_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(a, x, b, y, out, n);
  DGMR_CHECK_LAUNCH("dgmr_axpby");
  return 0;
}
int dgmr_fill(float* x, float value, int64_t n, dgmr_stream_t stream) {
  if (n == 0) return 0;
  fill_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(x, value, n);
  DGMR_CHECK_LAUNCH("dgmr_fill");
  return 0;
}
int dgmr_relu_fwd(const float* x, float* y, int64_t n, dgmr_stream_t stream) {
  if (n == 0) return 0;
  if (n % 4 == 0 && al16(x) && al16(y))
    relu_fwd4_kernel<<<ew_grid(n / 4, 256, 2), 256, 0, S(stream)>>>((const float4*)x, (float4*)y, n / 4);
  else
    relu_fwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(x, y, n);
  DGMR_CHECK_LAUNCH("dgmr_relu_fwd");
  return 0;
}
int dgmr_relu_bwd(const float* dy, const float* x, float* dx, int64_t n, dgmr_stream_t stream) {
  if (n == 0) return 0;
  if (n % 4 == 0 && al16(x) && al16(dy) && al16(dx))
    relu_bwd4_kernel<<<ew_grid(n / 4, 256, 2), 256, 0, S(stream)>>>((const float4*)dy, (const float4*)x, (float4*)dx, n / 4);
  else
    relu_bwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(dy, x, dx, n);
  DGMR_CHECK_LAUNCH("dgmr_relu_bwd");
  return 0;
}
int dgmr_split_tf32(const float* x, float* hi, float* lo, int64_t n, dgmr_stream_t stream) {
  if (n == 0) return 0;
  split_tf32_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(x, hi, lo, n);
  DGMR_CHECK_LAUNCH("dgmr_split_tf32");
  return 0;
}
int dgmr_pool_sum(const float* x, float* y, int N, int D, int H, int W, int C, int pd, int ph, int pw, float scale, dgmr_stream_t stream) {
  DGMR_REQUIRE(pd >= 1 && ph >= 1 && pw >= 1 && pd <= 2 && ph <= 2 && pw <= 2, "dgmr_pool_sum: window must be 1 or 2");
  int64_t total = (int64_t)N * (D / pd) * (H / ph) * (W / pw) * C;
  if (total == 0) return 0;
  if (C == 1 && pd == 1 && ph == 2 && pw == 2 && H % 2 == 0 && W % 4 == 0 && al16(x) && al16(y)) {
    const int64_t rows_out = (int64_t)N * D * (H / 2);
    pool_sum_c1_kernel<<<ew_grid(rows_out * (W / 4), 256, 1), 256, 0, S(stream)>>>(x, y, rows_out, W, scale);
    DGMR_CHECK_LAUNCH("dgmr_pool_sum");
    return 0;
  }
  if (C % 4 == 0 && al16(x) && al16(y))
    if ((int64_t)N * D * H * W * C < (int64_t)1 << 31) pool_sum4_kernel<uint32_t><<<ew_grid(total / 4, 256, 1), 256, 0, S(stream)>>>((const float4*)x, (float4*)y, N, D, H, W, C / 4, pd, ph, pw, scale);
    else pool_sum4_kernel<int64_t><<<ew_grid(total / 4, 256, 1), 256, 0, S(stream)>>>((const float4*)x, (float4*)y, N, D, H, W, C / 4, pd, ph, pw, scale);
  else
    pool_sum_kernel<<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(x, y, N, D, H, W, C, pd, ph, pw, scale);
  DGMR_CHECK_LAUNCH("dgmr_pool_sum");
  return 0;
}
int dgmr_upsample(const float* x, float* y, int N, int D, int H, int W, int C, int ud, int uh, int uw, int Do, int Ho, int Wo, float scale, dgmr_stream_t stream) {
  DGMR_REQUIRE(Do >= D * ud && Ho >= H * uh && Wo >= W * uw, "dgmr_upsample: output smaller than replicated input");
  int64_t total = (int64_t)N * Do * Ho * Wo * C;
  if (total == 0) return 0;
  if (C == 1 && ud == 1 && uh == 2 && uw == 2 && Do == D && Ho == 2 * H && Wo == 2 * W && W % 2 == 0 && al16(x) && al16(y)) {
    const int64_t rows_in = (int64_t)N * D * H;
    upsample_c1_kernel<<<ew_grid(rows_in * (W / 2), 256, 1), 256, 0, S(stream)>>>(x, y, rows_in, W, scale);
    DGMR_CHECK_LAUNCH("dgmr_upsample");
    return 0;
  }
  if (C % 4 == 0 && al16(x) && al16(y))
    if (total < (int64_t)1 << 31) upsample4_kernel<uint32_t><<<ew_grid(total / 4, 256, 1), 256, 0, S(stream)>>>((const float4*)x, (float4*)y, N, D, H, W, C / 4, ud, uh, uw, Do, Ho, Wo, scale);
    else upsample4_kernel<int64_t><<<ew_grid(total / 4, 256, 1), 256, 0, S(stream)>>>((const float4*)x, (float4*)y, N, D, H, W, C / 4, ud, uh, uw, Do, Ho, Wo, scale);
  else
    upsample_kernel<<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(x, y, N, D, H, W, C, ud, uh, uw, Do, Ho, Wo, scale);
  DGMR_CHECK_LAUNCH("dgmr_upsample");
  return 0;
}
int dgmr_gru_gate_fwd(float* pre_r, int ld, const float* x_r, const float* h, float* rh, int64_t rows, int Ch, int flags, dgmr_stream_t stream) {
  int64_t n = rows * Ch; if (n == 0) return 0;
  const int rnd = (flags & DGMR_FLAG_ROUND_TF32) ? 1 : 0;
  if (Ch % 4 == 0 && ld % 4 == 0 && al16(pre_r) && al16(x_r) && al16(h) && al16(rh))
    gru_gate_fwd4_kernel<<<ew_grid(n / 4, 256, 1), 256, 0, S(stream)>>>(pre_r, ld, x_r, (const float4*)h, (float4*)rh, rows, Ch / 4, rnd);
  else
    gru_gate_fwd_kernel<<<ew_grid(n, 256, 2), 256, 0, S(stream)>>>(pre_r, ld, x_r, h, rh, rows, Ch, rnd);
  DGMR_CHECK_LAUNCH("dgmr_gru_gate_fwd");
  return 0;
}
int dgmr_gru_blend_fwd(float* pre_u, int ld, const float* x_u, const float* h, float* c, const float* x_c, float* hnew, float* hnew_tf32, int64_t rows, int Ch,
                       int relu_c, dgmr_stream_t stream) {
  int64_t n = rows * Ch; if (n == 0) return 0;
  if (Ch % 4 == 0 && ld % 4 == 0 && al16(pre_u) && al16(x_u) && al16(h) && al16(c) && al16(x_c) && al16(hnew) && al16(hnew_tf32))
    gru_blend_fwd4_kernel<<<ew_grid(n / 4, 256, 1), 256, 0, S(stream)>>>(pre_u, ld, x_u, (const float4*)h, (float4*)c, (const float4*)x_c, (float4*)hnew,
                                                                         (float4*)hnew_tf32, rows, Ch / 4, relu_c);
  else
    gru_blend_fwd_kernel<<<ew_grid(n, 256, 2), 256, 0, S(stream)>>>(pre_u, ld, x_u, h, c, x_c, hnew, hnew_tf32, rows, Ch, relu_c);
  DGMR_CHECK_LAUNCH("dgmr_gru_blend_fwd");
  return 0;
}
int dgmr_gru_gate_bwd(const float* d_rh, const float* pre_r, int ld, const float* h, float* d_pre_r, int ldd, float* dh, int accumulate, int64_t rows, int Ch,
                      const float* dz_scale, float* dz, int dz_round, dgmr_stream_t stream) {
  int64_t n = rows * Ch; if (n == 0) return 0;
  DGMR_REQUIRE((dz == nullptr) == (dz_scale == nullptr), "dgmr_gru_gate_bwd: dz and dz_scale go together");
  gru_gate_bwd_kernel<<<ew_grid(n, 256, 2), 256, 0, S(stream)>>>(d_rh, pre_r, ld, h, d_pre_r, ldd, dh, accumulate, rows, Ch, dz_scale, dz, dz_round);
  DGMR_CHECK_LAUNCH("dgmr_gru_gate_bwd");
  return 0;
}
int dgmr_gru_blend_bwd(const float* d_hnew, const float* pre_u, int ld, const float* h, const float* c, float* d_pre_u, int ldd, float* dc, float* dh, int accumulate,
                       int64_t rows, int Ch, int relu_c, const float* dz_u_scale, float* dz_u, const float* dz_c_scale, float* dz_c, int dz_round,
                       dgmr_stream_t stream) {
  int64_t n = rows * Ch; if (n == 0) return 0;
  DGMR_REQUIRE((dz_u == nullptr) == (dz_u_scale == nullptr) && (dz_c == nullptr) == (dz_c_scale == nullptr), "dgmr_gru_blend_bwd: dz and its scale go together");
  gru_blend_bwd_kernel<<<ew_grid(n, 256, 2), 256, 0, S(stream)>>>(d_hnew, pre_u, ld, h, c, d_pre_u, ldd, dc, dh, accumulate, rows, Ch, relu_c, dz_u_scale, dz_u,
                                                                   dz_c_scale, dz_c, dz_round);
  DGMR_CHECK_LAUNCH("dgmr_gru_blend_bwd");
  return 0;
}

static int64_t bn_chunk(int64_t rows, int G) {
  int64_t blocks_per_group = (int64_t)sm_count() * 16 / (G > 0 ? G : 1);
  if (blocks_per_group < 1) blocks_per_group = 1;
  int64_t chunk = ceil_div(rows, blocks_per_group);
  if (chunk < 16) chunk = 16;
  return chunk;
}
int dgmr_bn_stats(const float* x, double* sums, int64_t rows, int G, int C, dgmr_stream_t stream) {
  DGMR_REQUIRE(rows > 0 && G > 0 && C > 0, "dgmr_bn_stats: bad dims");
  DGMR_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * G * C, S(stream)));
  int64_t chunk = bn_chunk(rows, G);
  dim3 grid((unsigned)ceil_div(rows, chunk), G);
  if (C % 4 == 0 && al16(x))
    bn_stats4_kernel<<<grid, 256, 256 * 8 * sizeof(double), S(stream)>>>((const float4*)x, sums, rows, C / 4, chunk);
  else
    bn_stats_kernel<<<grid, 256, 256 * 2 * sizeof(double), S(stream)>>>(x, sums, rows, C, chunk);
  DGMR_CHECK_LAUNCH("dgmr_bn_stats");
  return 0;
}
int dgmr_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean, float* running_var, int64_t rows, int G, int C,
                     float eps, float momentum, int training, float* mean, float* invstd, float* a, float* b, dgmr_stream_t stream) {
  DGMR_REQUIRE(G >= 1 && G <= 128, "dgmr_bn_finalize: G out of range (1..128)");
  bn_finalize_kernel<<<(C + 31) / 32, 256, sizeof(float) * 2 * G * 32, S(stream)>>>(sums, gamma, beta, running_mean, running_var, rows, G, C, eps, momentum, training, mean,
                                                                                    invstd, a, b);
  DGMR_CHECK_LAUNCH("dgmr_bn_finalize");
  return 0;
}
int dgmr_bn_apply(const float* x, const float* a, const float* b, float* y, float* x_rounded, int64_t rows, int G, int C, int relu, int up2, int H, int W,
                  dgmr_stream_t stream) {
  const int rnd = (relu & DGMR_FLAG_ROUND_TF32) ? 1 : 0;   // output feeds tensor-core convs only: emit tf32-rounded values directly
  relu &= ~DGMR_FLAG_ROUND_TF32;
  int64_t total = (int64_t)G * rows * C * (up2 ? 4 : 1);
  if (total == 0) return 0;
  DGMR_REQUIRE(!up2 || (rows % ((int64_t)H * W) == 0), "dgmr_bn_apply: rows not a multiple of H*W");
  DGMR_REQUIRE(!(up2 && x_rounded), "dgmr_bn_apply: x_rounded is not available together with up2");
  if (C % 4 == 0 && al16(x) && al16(y) && al16(a) && al16(b) && al16(x_rounded))
    if (total < (int64_t)1 << 31) bn_apply4_kernel<uint32_t><<<ew_grid(total / 4, 256, 1), 256, 0, S(stream)>>>((const float4*)x, (const float4*)a, (const float4*)b, (float4*)y, rows, G, C / 4, relu, up2, H, W, rnd, (float4*)x_rounded);
    else bn_apply4_kernel<int64_t><<<ew_grid(total / 4, 256, 1), 256, 0, S(stream)>>>((const float4*)x, (const float4*)a, (const float4*)b, (float4*)y, rows, G, C / 4, relu, up2, H, W, rnd, (float4*)x_rounded);
  else
    bn_apply_kernel<<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(x, a, b, y, rows, G, C, relu, up2, H, W, rnd, x_rounded);
  DGMR_CHECK_LAUNCH("dgmr_bn_apply");
  return 0;
}
int dgmr_bn_bwd_reduce(const float* dy, const float* x, const float* a, const float* b, const float* mean, const float* invstd, double* red,
                       int64_t rows, int G, int C, int relu, int up2, int H, int W, dgmr_stream_t stream) {
  DGMR_CUDA(cudaMemsetAsync(red, 0, sizeof(double) * 2 * G * C, S(stream)));
  int64_t chunk = bn_chunk(rows, G);
  dim3 grid((unsigned)ceil_div(rows, chunk), G);
  if (C % 4 == 0 && al16(dy) && al16(x) && al16(a) && al16(b) && al16(mean) && al16(invstd) && (!up2 || rows * G < ((int64_t)1 << 31))) {
    if (up2)
      bn_bwd_reduce4_kernel<true><<<grid, 256, 256 * 8 * sizeof(double), S(stream)>>>(dy, (const float4*)x, (const float4*)a, (const float4*)b, (const float4*)mean,
                                                                                      (const float4*)invstd, red, rows, C / 4, chunk, relu, H, W);
    else
      bn_bwd_reduce4_kernel<false><<<grid, 256, 256 * 8 * sizeof(double), S(stream)>>>(dy, (const float4*)x, (const float4*)a, (const float4*)b, (const float4*)mean,
                                                                                       (const float4*)invstd, red, rows, C / 4, chunk, relu, H, W);
  }
  else
    bn_bwd_reduce_kernel<<<grid, 256, 256 * 2 * sizeof(double), S(stream)>>>(dy, x, a, b, mean, invstd, red, rows, C, chunk, relu, up2, H, W);
  DGMR_CHECK_LAUNCH("dgmr_bn_bwd_reduce");
  return 0;
}
int dgmr_bn_bwd_apply(const float* dy, const float* x, const float* a, const float* b, const float* mean, const float* invstd, const float* out_scale,
                      const double* red, float* dx, const float* dx_add, float* dgamma, float* dbeta, int accumulate, int64_t rows, int G, int C,
                      int relu, int up2, int H, int W, int training, dgmr_stream_t stream) {
  const int rnd = (relu & DGMR_FLAG_ROUND_TF32) ? 1 : 0;
  relu &= ~DGMR_FLAG_ROUND_TF32;
  const float* oscale = out_scale;
  int64_t total = (int64_t)G * rows * C;
  if (dx && total) {
    if (C % 4 == 0 && al16(dy) && al16(x) && al16(dx) && al16(a) && al16(b) && al16(mean) && al16(invstd) && al16(oscale) && al16(dx_add) &&
        (!up2 || rows * G < ((int64_t)1 << 31))) {
      const int64_t chunk = bn_chunk(rows, G);
      dim3 grid((unsigned)ceil_div(rows, chunk), G);
#define DGMR_BNA(U_, A_) bn_bwd_apply4_kernel<U_, A_><<<grid, 256, 0, S(stream)>>>(dy, (const float4*)x, (const float4*)a, (const float4*)b, (const float4*)mean, \
        (const float4*)invstd, red, (float4*)dx, rows, C / 4, chunk, relu, H, W, training, (const float4*)oscale, rnd, (const float4*)dx_add)
      if (up2) { if (dx_add) DGMR_BNA(true, true); else DGMR_BNA(true, false); }
      else { if (dx_add) DGMR_BNA(false, true); else DGMR_BNA(false, false); }
#undef DGMR_BNA
    }
    else
      bn_bwd_apply_kernel<<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(dy, x, a, b, mean, invstd, red, dx, rows, G, C, relu, up2, H, W, training, oscale, rnd, dx_add);
    DGMR_CHECK_LAUNCH("dgmr_bn_bwd_apply");
  }
  if (dgamma || dbeta) {
    bn_bwd_params_kernel<<<(C + 127) / 128, 128, 0, S(stream)>>>(red, dgamma, dbeta, G, C, accumulate);
    DGMR_CHECK_LAUNCH("dgmr_bn_bwd_params");
  }
  return 0;
}
int dgmr_sumpool_relu_fwd(const float* x, float* y, int N, int HW, int C, dgmr_stream_t stream) {
  int64_t n = (int64_t)N * C; if (n == 0) return 0;
  sumpool_relu_fwd_kernel<<<ew_grid(n, 256, 1), 256, 0, S(stream)>>>(x, y, N, HW, C);
  DGMR_CHECK_LAUNCH("dgmr_sumpool_relu_fwd");
  return 0;
}
int dgmr_sumpool_relu_bwd(const float* dy, const float* x, float* dx, int N, int HW, int C, dgmr_stream_t stream) {
  int64_t n = (int64_t)N * HW * C; if (n == 0) return 0;
  sumpool_relu_bwd_kernel<<<ew_grid(n, 256, 1), 256, 0, S(stream)>>>(dy, x, dx, N, HW, C);
  DGMR_CHECK_LAUNCH("dgmr_sumpool_relu_bwd");
  return 0;
}
int dgmr_attention_fwd(const float* q, const float* k, const float* v, float* out, float* beta, int B, int H, int W, int C, dgmr_stream_t stream) {
  int L = C * H;
  size_t sh = (size_t)(L + W + 32) * sizeof(float);
  DGMR_REQUIRE(sh <= 200 * 1024, "dgmr_attention_fwd: L=%d too large", L);
  if (sh > 48 * 1024) DGMR_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  attention_fwd_kernel<<<dim3(L, B), 128, sh, S(stream)>>>(q, k, v, out, beta, H, W, C);
  DGMR_CHECK_LAUNCH("dgmr_attention_fwd");
  return 0;
}
int dgmr_attention_bwd(const float* dout, const float* q, const float* k, const float* v, const float* beta, float* dq, float* dk, float* dv, float* ws,
                       int B, int H, int W, int C, dgmr_stream_t stream) {
  int L = C * H;
  size_t sh = (size_t)(L + W + 32) * sizeof(float);
  DGMR_REQUIRE(sh <= 200 * 1024, "dgmr_attention_bwd: L=%d too large", L);
  if (sh > 48 * 1024) DGMR_CUDA(cudaFuncSetAttribute(attention_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  attention_bwd1_kernel<<<dim3(L, B), 128, sh, S(stream)>>>(dout, k, v, beta, dq, ws, H, W, C);
  DGMR_CHECK_LAUNCH("dgmr_attention_bwd1");
  attention_bwd2_kernel<<<dim3(L, B), 32, 0, S(stream)>>>(dout, q, beta, ws, dk, dv, H, W, C);
  DGMR_CHECK_LAUNCH("dgmr_attention_bwd2");
  return 0;
}
int dgmr_hinge_disc(const float* scores, int B, int cols, float* loss, float* dscores, dgmr_stream_t stream) {
  DGMR_REQUIRE(B > 0 && cols > 0, "dgmr_hinge_disc: bad dims");
  hinge_disc_kernel<<<1, 128, 0, S(stream)>>>(scores, B, cols, loss, dscores);
  DGMR_CHECK_LAUNCH("dgmr_hinge_disc");
  return 0;
}
int dgmr_hinge_gen(const float* scores, int n, float* loss, float* dscores, dgmr_stream_t stream) {
  hinge_gen_kernel<<<1, 128, 0, S(stream)>>>(scores, n, loss, dscores);
  DGMR_CHECK_LAUNCH("dgmr_hinge_gen");
  return 0;
}
int dgmr_grid_cell_fwd(const float* gen, const float* target, float cap, float coef, float* loss, double* acc, int64_t n, dgmr_stream_t stream) {
  DGMR_REQUIRE(acc != nullptr, "dgmr_grid_cell_fwd: fp64 scratch (1 double) required");
  DGMR_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), S(stream)));
  grid_cell_fwd_kernel<<<ew_grid(n, 256, 8), 256, 0, S(stream)>>>(gen, target, cap, acc, n);
  DGMR_CHECK_LAUNCH("dgmr_grid_cell_fwd");
  grid_cell_final_kernel<<<1, 1, 0, S(stream)>>>(acc, coef, loss);
  DGMR_CHECK_LAUNCH("dgmr_grid_cell_final");
  return 0;
}
int dgmr_grid_cell_bwd(const float* gen, const float* target, float cap, float coef, const float* gout, float* dgen, int64_t n, dgmr_stream_t stream) {
  if (n == 0) return 0;
  grid_cell_bwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(gen, target, cap, coef, gout, dgen, n);
  DGMR_CHECK_LAUNCH("dgmr_grid_cell_bwd");
  return 0;
}
int dgmr_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, int step, float grad_scale, dgmr_stream_t stream) {
  if (n == 0) return 0;
  DGMR_REQUIRE(step >= 1, "dgmr_adam: step must be >= 1");
  float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2s, grad_scale);
  DGMR_CHECK_LAUNCH("dgmr_adam");
  return 0;
}

}  // extern "C"
