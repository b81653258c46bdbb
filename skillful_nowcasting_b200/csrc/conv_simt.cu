// Generic CUDA-core (fp32 FMA) implicit-GEMM convolution kernels + weight pack/unpack +
// conv backward prologue + spectral-norm power iteration.
// The SIMT conv kernels serve every shape the tensor-core path (conv_umma.cu) does not take
// (tiny channel counts such as the 4-channel space-to-depth inputs, batch-1 latent stack,
// 1x1/2x2 images deep in the discriminators) and are the on-device cross-check for it.
#include "common.cuh"

namespace dgmr {

struct ConvDims {
  int N, D, H, W, Cin, Cout, kd, kh, kw, G;
};

__device__ __forceinline__ float rna_tf32(float x) {  // round to nearest tf32 (10 explicit mantissa bits), low 13 bits zero
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ------------------------------------------------------------------ forward / dgrad
// y[m][co] = act( sum_{tap,ci} x[pix(m)+tap][ci] * wp[tap][co][ci] * scale + bias + res )
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) conv_simt_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                                                            const float* __restrict__ scale, const float* __restrict__ res, float* __restrict__ y,
                                                            ConvDims d, int act) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)d.N * d.D * d.H * d.W;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  // decode my load row's pixel
  const int64_t mrow = m0 + lrow;
  int pn = 0, pd_ = 0, ph_ = 0, pw_ = 0;
  const bool mvalid = mrow < M;
  if (mvalid) {
    int64_t r = mrow;
    pw_ = r % d.W; r /= d.W;
    ph_ = r % d.H; r /= d.H;
    pd_ = r % d.D; pn = r / d.D;
  }
  const int co_l = n0 + lrow;
  const bool vec = (d.Cin & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid >> 4, tx = tid & 15;
  const int taps = d.kd * d.kh * d.kw;
  for (int tap = 0; tap < taps; ++tap) {
    int tw = tap % d.kw, th = (tap / d.kw) % d.kh, td = tap / (d.kw * d.kh);
    int iw = pw_ + tw - d.kw / 2, ih = ph_ + th - d.kh / 2, id = pd_ + td - d.kd / 2;
    bool pvalid = mvalid && iw >= 0 && iw < d.W && ih >= 0 && ih < d.H && id >= 0 && id < d.D;
    const float* xp = x + ((((int64_t)pn * d.D + id) * d.H + ih) * d.W + iw) * d.Cin;
    const float* wrow = wp + ((int64_t)tap * d.Cout + co_l) * d.Cin;
    for (int c0 = 0; c0 < d.Cin; c0 += BK) {
      float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
      int c = c0 + lk;
      if (pvalid) {
        if (vec && c + 3 < d.Cin) {
          float4 t = *reinterpret_cast<const float4*>(xp + c);
          a4[0] = t.x; a4[1] = t.y; a4[2] = t.z; a4[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (c + i < d.Cin) a4[i] = xp[c + i];
        }
      }
      if (co_l < d.Cout) {
        if (vec && c + 3 < d.Cin) {
          float4 t = *reinterpret_cast<const float4*>(wrow + c);
          b4[0] = t.x; b4[1] = t.y; b4[2] = t.z; b4[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (c + i < d.Cin) b4[i] = wrow[c + i];
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) { As[lk + i][lrow] = a4[i]; Bs[lk + i][lrow] = b4[i]; }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float a_[4] = {av.x, av.y, av.z, av.w}, b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
      }
    }
  }
  const int64_t per_img = (int64_t)d.D * d.H * d.W;
  const int imgs_per_group = d.N / d.G;
  const bool round_out = (act & DGMR_FLAG_ROUND_OUT) != 0, res_up2 = (act & DGMR_FLAG_RES_UP2) != 0;
  act &= 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int g = (int)((m / per_img) / imgs_per_group);
    int64_t mr = m;     // row of the residual: the same pixel, or (h/2, w/2) of a half-resolution tensor
    if (res_up2) {
      int64_t r = m;
      const int w_ = (int)(r % d.W); r /= d.W;
      const int h_ = (int)(r % d.H); r /= d.H;    // r = n*D + dd
      mr = (r * (d.H >> 1) + (h_ >> 1)) * (d.W >> 1) + (w_ >> 1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= d.Cout) continue;
      float v = acc[i][j];
      if (scale) v *= scale[(int64_t)g * d.Cout + co];
      if (bias) v += bias[co];
      if (res) v += res[mr * d.Cout + co];
      if (act == DGMR_ACT_RELU) v = fmaxf(v, 0.f);
      y[m * d.Cout + co] = round_out ? rna_tf32(v) : v;
    }
  }
}

// ------------------------------------------------------------------ wgrad
// dwp[tap][co][ci] += sum_{p in chunk} dz[p][co] * x[p+tap][ci]
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) conv_simt_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ dwp,
                                                              ConvDims d, int ci_tiles, int64_t chunk) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int co0 = (blockIdx.x / ci_tiles) * BM, ci0 = (blockIdx.x % ci_tiles) * BN;
  const int tap = blockIdx.y;
  const int tw = tap % d.kw, th = (tap / d.kw) % d.kh, td = tap / (d.kw * d.kh);
  const int64_t M = (int64_t)d.N * d.D * d.H * d.W;
  const int64_t p_begin = (int64_t)blockIdx.z * chunk;
  const int64_t p_end = p_begin + chunk < M ? p_begin + chunk : M;
  const int lk = tid >> 4, lq = (tid & 15) * 4;  // k row 0..15, 4 consecutive channels
  const bool veca = (d.Cout & 3) == 0, vecb = (d.Cin & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid >> 4, tx = tid & 15;
  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    int64_t p = p0 + lk;
    float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < p_end) {
      int co = co0 + lq;
      const float* dp = dz + p * d.Cout;
      if (veca && co + 3 < d.Cout) {
        float4 t = *reinterpret_cast<const float4*>(dp + co);
        a4[0] = t.x; a4[1] = t.y; a4[2] = t.z; a4[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (co + i < d.Cout) a4[i] = dp[co + i];
      }
      int64_t r = p;
      int pw_ = r % d.W; r /= d.W;
      int ph_ = r % d.H; r /= d.H;
      int pd_ = r % d.D; int pn = r / d.D;
      int iw = pw_ + tw - d.kw / 2, ih = ph_ + th - d.kh / 2, id = pd_ + td - d.kd / 2;
      if (iw >= 0 && iw < d.W && ih >= 0 && ih < d.H && id >= 0 && id < d.D) {
        const float* xp = x + ((((int64_t)pn * d.D + id) * d.H + ih) * d.W + iw) * d.Cin;
        int ci = ci0 + lq;
        if (vecb && ci + 3 < d.Cin) {
          float4 t = *reinterpret_cast<const float4*>(xp + ci);
          b4[0] = t.x; b4[1] = t.y; b4[2] = t.z; b4[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (ci + i < d.Cin) b4[i] = xp[ci + i];
        }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lk][lq]) = make_float4(a4[0], a4[1], a4[2], a4[3]);
    *reinterpret_cast<float4*>(&Bs[lk][lq]) = make_float4(b4[0], b4[1], b4[2], b4[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a_[4] = {av.x, av.y, av.z, av.w}, b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= d.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = ci0 + tx * 4 + j;
      if (ci >= d.Cin) continue;
      atomicAdd(&dwp[((int64_t)tap * d.Cout + co) * d.Cin + ci], acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------ pack / unpack

__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ packed, int Cout, int CinTot, int ci0, int Cin, int taps, int mode, int rnd) {
  int64_t total = (int64_t)taps * Cout * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // iterate in destination order for coalesced writes
    float v;
    if (mode == 0) {
      int ci = i % Cin; int64_t r = i / Cin; int co = r % Cout; int tap = r / Cout;
      v = w[((int64_t)co * CinTot + ci0 + ci) * taps + tap];
    } else {
      int co = i % Cout; int64_t r = i / Cout; int ci = r % Cin; int tapf = r / Cin;
      int tap = taps - 1 - tapf;
      v = w[((int64_t)co * CinTot + ci0 + ci) * taps + tap];
    }
    packed[i] = rnd ? rna_tf32(v) : v;
  }
}
// All the packs a network needs after an optimiser step in ONE launch: blockIdx.y = item, blockIdx.x strides over its elements.
// Destination may be wider than the packed slice: CinPad >= Cin (zero-padded input channels: the pad stays as the caller zeroed it) and a
// [co0, co0+Cout) window of CoutTot rows (two gate weights of a ConvGRU side by side along Cout).
constexpr int kPackMaxItems = 64;
struct PackMultiArgs { dgmr_pack_item it[kPackMaxItems]; };
__global__ void pack_weight_multi_kernel(const __grid_constant__ PackMultiArgs a) {
  const dgmr_pack_item& t = a.it[blockIdx.y];
  const int rnd = (t.mode & DGMR_FLAG_ROUND_TF32) ? 1 : 0, mode = t.mode & ~DGMR_FLAG_ROUND_TF32;
  const int64_t total = (int64_t)t.taps * t.Cout * t.Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v; int64_t o;
    if (mode == 0) {
      int ci = i % t.Cin; int64_t r = i / t.Cin; int co = r % t.Cout; int tap = r / t.Cout;
      v = t.w[((int64_t)co * t.CinTot + t.ci0 + ci) * t.taps + tap];
      o = ((int64_t)tap * t.CoutTot + t.co0 + co) * t.CinPad + ci;
    } else {
      int co = i % t.Cout; int64_t r = i / t.Cout; int ci = r % t.Cin; int tapf = r / t.Cin;
      v = t.w[((int64_t)co * t.CinTot + t.ci0 + ci) * t.taps + (t.taps - 1 - tapf)];
      o = ((int64_t)tapf * t.CinPad + ci) * t.CoutTot + t.co0 + co;
    }
    t.packed[o] = rnd ? rna_tf32(v) : v;
  }
}
__global__ void round_tf32_kernel(const float4* x, float4* y, int64_t n4) {   // y may alias x
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    y[i] = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
  }
}
__global__ void round_tf32_tail_kernel(const float* x, float* y, int64_t start, int64_t n) {
  int64_t i = start + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = rna_tf32(x[i]);
}
__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ gw, int Cout, int CinTot, int ci0, int Cin, int taps, int acc) {
  int64_t total = (int64_t)taps * Cout * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // iterate in gw order: (co, ci, tap)
    int tap = i % taps; int64_t r = i / taps; int ci = r % Cin; int co = r / Cin;
    float v = packed[((int64_t)tap * Cout + co) * Cin + ci];
    int64_t o = ((int64_t)co * CinTot + ci0 + ci) * taps + tap;
    if (acc) gw[o] += v; else gw[o] = v;
  }
}

// ------------------------------------------------------------------ backward prologue
// rows = pixels per group; grid (row chunks, G); 256 threads as (cpl channel-vector lanes x rp row lanes).
// V = 4: float4 path (Cout % 4 == 0), V = 1: scalar fallback.  One pass: reads dy (+y, +res), writes dz (+dpre),
// and reduces dbias / dscale per channel -> HBM-bound, 8-20 B per element.  The template flags only REMOVE paths (HASY: y may be
// read, DSR: dscale with a residual, GEO: pooled-gradient / half-resolution-residual index maps), so that the common variants stay
// small enough for 3-4 CTAs per SM with four rows of loads in flight per thread; the fp64 partial sums live in shared memory
// (touched once per 32 iterations), not in registers.
struct PoolGeom { int pd, ph, pw, D, H, W; float inv; int lw, lh; };   // lw / lh: log2 of W / H when they are powers of two (shift instead of divide), else -1
template <int V, bool HASY, bool DSR, bool GEO>
__global__ void __launch_bounds__(256, 3) conv_bwd_prep_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ res,
                                                               const float* __restrict__ bias, const float* __restrict__ scale, float* __restrict__ dz,
                                                               float* __restrict__ dpre, float* __restrict__ dbias, float* __restrict__ dscale,
                                                               int64_t rows, int C, int64_t chunk, int act, int rnd, int up_h, int up_w, PoolGeom pg) {
  // pg.pd > 0: `dy` is the gradient of the AVERAGE-POOLED output (window pd x ph x pw, floor) of this convolution: read at (d/pd, h/ph, w/pw) and
  // divided by the window size -- the pooling backward (a nearest upsample pass) happens inside this load
  // up_h, up_w > 0: `res` is the HALF-resolution tensor of DGMR_FLAG_RES_UP2 (full-resolution image up_h x up_w): read at (h/2, w/2)
  extern __shared__ double sh[];  // [256][2*V]: (sum dpre, sum dpre*(y-b-res)) per thread
  const int g = blockIdx.y;
  const int CV = C / V;
  const int cpl = CV < 256 ? CV : 256;
  const int rp = 256 / cpl;
  const int cl = threadIdx.x % cpl, rl = threadIdx.x / cpl;
  const int64_t r0 = (int64_t)blockIdx.x * chunk;
  const int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
  const bool reduce = dbias != nullptr || dscale != nullptr;
  const bool use_y = HASY && y != nullptr;
  const bool use_res = DSR && res != nullptr && dscale != nullptr;
  const bool relu = HASY && act == DGMR_ACT_RELU;
  double* my = sh + (size_t)threadIdx.x * 2 * V;
  for (int cb = 0; cb < CV; cb += cpl) {     // block-uniform trip count (barriers inside)
    const int cv = cb + cl;
    const int c = cv * V;
    if (reduce) {
#pragma unroll
      for (int i = 0; i < 2 * V; ++i) my[i] = 0.0;
    }
    if (rl < rp && cv < CV) {
      float sc[V], bi[V], fs[V], fq[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { sc[i] = scale ? scale[(int64_t)g * C + c + i] : 1.f; bi[i] = bias ? bias[c + i] : 0.f; fs[i] = 0.f; fq[i] = 0.f; }
      struct Row { float d[V], yv[V], rv[V]; int64_t o; };
      auto load = [&](int64_t r, Row& w) {
        const int64_t R = (int64_t)g * rows + r;
        w.o = R * C + c;
        int64_t ro = w.o, dyo = w.o;
        float dsc = 1.f;
        if (GEO) {
          const uint32_t Ru = (uint32_t)R;                         // host checks G*rows < 2^31 for these maps
          if (up_w > 0 && use_res) {
            const uint32_t hw = (uint32_t)up_h * (uint32_t)up_w;
            const uint32_t img = Ru / hw, rem = Ru - img * hw;
            const uint32_t hh = rem / (uint32_t)up_w, ww = rem - hh * (uint32_t)up_w;
            ro = ((int64_t)(img * (uint32_t)(up_h >> 1) + (hh >> 1)) * (up_w >> 1) + (ww >> 1)) * C + c;
          }
          if (pg.pd > 0) {
            uint32_t t = Ru, ww, hh;
            if (pg.lw >= 0) { ww = t & (uint32_t)(pg.W - 1); t >>= pg.lw; } else { ww = t % (uint32_t)pg.W; t /= (uint32_t)pg.W; }
            if (pg.lh >= 0) { hh = t & (uint32_t)(pg.H - 1); t >>= pg.lh; } else { hh = t % (uint32_t)pg.H; t /= (uint32_t)pg.H; }
            const uint32_t nn = t / (uint32_t)pg.D, dd = t - nn * (uint32_t)pg.D;
            const uint32_t Dp = pg.D / pg.pd, Hp = pg.H / pg.ph, Wp = pg.W / pg.pw;
            const uint32_t dq = dd >> (pg.pd - 1), hq = hh >> (pg.ph - 1), wq = ww >> (pg.pw - 1);       // windows are 1 or 2 wide (host-checked)
            if (dq < Dp && hq < Hp && wq < Wp) { dyo = ((((int64_t)nn * Dp + dq) * Hp + hq) * Wp + wq) * C + c; dsc = pg.inv; }
            else { dyo = 0; dsc = 0.f; }     // rim dropped by the floor: no gradient
          }
        }
        if (V == 4) {
          float4 t = *reinterpret_cast<const float4*>(dy + dyo); w.d[0] = t.x * dsc; w.d[1] = t.y * dsc; w.d[2] = t.z * dsc; w.d[3] = t.w * dsc;
          if (use_y) { float4 u = *reinterpret_cast<const float4*>(y + w.o); w.yv[0] = u.x; w.yv[1] = u.y; w.yv[2] = u.z; w.yv[3] = u.w; }
          if (use_res) { float4 u = *reinterpret_cast<const float4*>(res + ro); w.rv[0] = u.x; w.rv[1] = u.y; w.rv[2] = u.z; w.rv[3] = u.w; }
        } else {
          w.d[0] = dy[dyo] * dsc;
          if (use_y) w.yv[0] = y[w.o];
          if (use_res) w.rv[0] = res[ro];
        }
      };
      auto finish = [&](Row& w) {
        float zo[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float yy = use_y ? w.yv[i] : 0.f;
          if (relu && !(yy > 0.f)) w.d[i] = 0.f;
          zo[i] = w.d[i] * sc[i];
          if (rnd) zo[i] = rna_tf32(zo[i]);
          fs[i] += w.d[i];
          if (HASY && dscale) fq[i] += w.d[i] * (yy - bi[i] - (use_res ? w.rv[i] : 0.f));
        }
        if (V == 4) {
          if (dz) *reinterpret_cast<float4*>(dz + w.o) = make_float4(zo[0], zo[1], zo[2], zo[3]);
          if (dpre) *reinterpret_cast<float4*>(dpre + w.o) = make_float4(w.d[0], w.d[1], w.d[2], w.d[3]);
        } else {
          if (dz) dz[w.o] = zo[0];
          if (dpre) dpre[w.o] = w.d[0];
        }
      };
      auto flush = [&]() {
        if (reduce) {
#pragma unroll
          for (int i = 0; i < V; ++i) { my[2 * i] += (double)fs[i]; my[2 * i + 1] += (double)fq[i]; fs[i] = 0.f; fq[i] = 0.f; }
        }
      };
      constexpr int U = (DSR || V == 1) ? 2 : 4;     // rows of loads in flight per thread
      int cnt = 0;
      int64_t r = r0 + rl;
      for (; r + (int64_t)(U - 1) * rp < r1; r += (int64_t)U * rp) {
        Row w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) load(r + (int64_t)u * rp, w[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) finish(w[u]);
        if (++cnt == 64 / U) { flush(); cnt = 0; }
      }
      for (; r < r1; r += rp) { Row a; load(r, a); finish(a); }
      flush();
    }
    if (reduce) {
      __syncthreads();
      // 2V partial sums per channel lane and row lane: spread the cross-row-lane reduction over cpl * 2V threads
      for (int t = threadIdx.x; t < cpl * 2 * V; t += 256) {
        const int lane = t / (2 * V), k = t - lane * (2 * V);
        if (cb + lane < CV) {
          double v = 0.0;
          for (int j = 0; j < rp; ++j) v += sh[(size_t)(j * cpl + lane) * 2 * V + k];
          const int ch = (cb + lane) * V + (k >> 1);
          if ((k & 1) == 0) { if (dbias) atomicAdd(&dbias[ch], (float)v); }
          else if (dscale) atomicAdd(&dscale[(int64_t)g * C + ch], (float)(v / (double)scale[(int64_t)g * C + ch]));
        }
      }
      __syncthreads();
    }
  }
}

// out[r] = <a[r, off:off+cols], b[r, off:off+cols]> / denom[r]: the spectral-norm scale gradient of a G = 1 convolution from its WEIGHT gradient
// (with Y = s conv(X, W) + b and dW = wgrad(s dY, X): dL/ds[co] = <dY[co], conv(X, W)[co]> = <dW[co], W[co]> / s[co]) -- a few KB of weights
// instead of a pass over the activations.  One CTA per output channel, fp64 accumulation.
__global__ void __launch_bounds__(256) rowdot_div_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ denom,
                                                         float* __restrict__ out, int64_t cols, int64_t ld, int64_t offset) {
  __shared__ double sh[8];
  const int r = blockIdx.x;
  const float* ar = a + (int64_t)r * ld + offset;
  const float* br = b + (int64_t)r * ld + offset;
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < cols; j += 256) acc += (double)ar[j] * (double)br[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh[w];
    out[r] = (float)(denom ? v / (double)denom[r] : v);
  }
}

// ------------------------------------------------------------------ spectral norm
// Column-partitioned persistent kernel: the CTAs of one weight each own a column slab W[:, k0:k1) (kept in shared memory
// when it fits, streamed from L2 otherwise) and run ALL G power iterations of that weight inside one launch, two group
// barriers per iteration.  The multi-weight entry point runs every spectrally normalised layer of a module (the 37 layers of
// the sampler, ...) in ONE launch: CTAs are split over the weights in proportion to their size and each weight's CTA group
// synchronises on its own arrival counter, so the per-iteration latency (a few global round trips) is paid once per
// iteration for all layers together instead of once per layer.
struct SnItem {
  const float* w; float* u; float* v;
  float* inv_sigma; float* u_hist; float* v_hist;
  float* ws;          // zeroed: t[(G+1)][R], nq[G], barrier counter
  int R, K, G, training;
  float eps;
  int cta_begin, cta_count, kw, in_smem;
};
constexpr int kSnMaxItems = 40;
struct SnMultiArgs { SnItem it[kSnMaxItems]; int n; };

// Barrier among `size` co-resident CTAs on a monotonically increasing counter (barrier k completes at k*size arrivals).
__device__ __forceinline__ void sn_group_barrier(unsigned* counter, unsigned& epoch, unsigned size) {
  __syncthreads();
  if (size > 1) {
    if (threadIdx.x == 0) {
      ++epoch;
      const unsigned target = epoch * size;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      } while (seen < target);
      __threadfence();
    }
    __syncthreads();
  }
}
__device__ __forceinline__ float block_sumsq(const float* a, int n, float* red) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { float v = __ldcg(a + i); s += v * v; }
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
  return tot;
}
__device__ void sn_run(const SnItem& a, const int rank, float* smem) {
  unsigned epoch = 0;   // only thread 0's copy is used
  const unsigned size = (unsigned)a.cta_count;
  const int R = a.R, K = a.K, kwmax = a.kw;
  const int k0 = rank * kwmax;
  const int kw = (k0 + kwmax <= K) ? kwmax : (K - k0 > 0 ? K - k0 : 0);
  float* ush = smem;                // [R]
  float* vsh = ush + ((R + 3) & ~3); // [kwmax], 16-byte aligned
  float* qsh = vsh + kwmax;         // [kwmax]
  float* red = qsh + kwmax;         // [32]
  float* wsh = red + 32;            // [R][kwmax] if in_smem
  float* t = a.ws;
  float* nq = a.ws + (size_t)(a.G + 1) * R;
  unsigned* bar = reinterpret_cast<unsigned*>(nq + a.G);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (a.in_smem) {
    for (int r = warp; r < R; r += nwarps)
      for (int k = lane; k < kw; k += 32) wsh[r * kwmax + k] = a.w[(int64_t)r * K + k0 + k];
  }
  for (int k = threadIdx.x; k < kw; k += blockDim.x) vsh[k] = a.v[k0 + k];
  __syncthreads();
  // slab access: shared-memory copy (pitch kwmax) or straight from global/L2 (pitch K).  float4 loads + 4-way unrolling keep
  // enough bytes in flight per SM to stream the slab at memory speed (1024 threads per CTA).
  const float* wb = a.in_smem ? wsh : a.w + k0;
  const int64_t ld = a.in_smem ? kwmax : K;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(wb) & 15u) == 0) && ((kw & 3) == 0);
  // t[slot] += W[:, mine] v[mine]
  auto matvec = [&](int slot) {
    for (int r = warp; r < R; r += nwarps) {
      const float* wr = wb + (int64_t)r * ld;
      float s = 0.f;
      if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
        const float4* v4 = reinterpret_cast<const float4*>(vsh);
        const int n4 = kw >> 2;
        int k = lane;
        for (; k + 96 < n4; k += 128) {
          float4 a0 = w4[k], a1 = w4[k + 32], a2 = w4[k + 64], a3 = w4[k + 96];
          float4 b0 = v4[k], b1 = v4[k + 32], b2 = v4[k + 64], b3 = v4[k + 96];
          s += a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w + a1.x * b1.x + a1.y * b1.y + a1.z * b1.z + a1.w * b1.w +
               a2.x * b2.x + a2.y * b2.y + a2.z * b2.z + a2.w * b2.w + a3.x * b3.x + a3.y * b3.y + a3.z * b3.z + a3.w * b3.w;
        }
        for (; k < n4; k += 32) { float4 a0 = w4[k], b0 = v4[k]; s += a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w; }
      } else {
        for (int k = lane; k < kw; k += 32) s += wr[k] * vsh[k];
      }
      s = warp_sum(s);
      if (lane == 0 && kw > 0) atomicAdd(&t[(int64_t)slot * R + r], s);
    }
  };
  // qsh[mine] += W[:, mine]^T u : warps split the rows, lanes own (groups of 4) columns
  auto colsum = [&]() {
    if (vec) {
      const int n4 = kw >> 2;
      for (int kb = 0; kb < n4; kb += 32) {
        const int k = kb + lane;
        if (k < n4) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          int r = warp;
          for (; r + 3 * nwarps < R; r += 4 * nwarps) {
            float4 a0 = reinterpret_cast<const float4*>(wb + (int64_t)r * ld)[k];
            float4 a1 = reinterpret_cast<const float4*>(wb + (int64_t)(r + nwarps) * ld)[k];
            float4 a2 = reinterpret_cast<const float4*>(wb + (int64_t)(r + 2 * nwarps) * ld)[k];
            float4 a3 = reinterpret_cast<const float4*>(wb + (int64_t)(r + 3 * nwarps) * ld)[k];
            float u0 = ush[r], u1 = ush[r + nwarps], u2 = ush[r + 2 * nwarps], u3 = ush[r + 3 * nwarps];
            acc.x += u0 * a0.x + u1 * a1.x + u2 * a2.x + u3 * a3.x;
            acc.y += u0 * a0.y + u1 * a1.y + u2 * a2.y + u3 * a3.y;
            acc.z += u0 * a0.z + u1 * a1.z + u2 * a2.z + u3 * a3.z;
            acc.w += u0 * a0.w + u1 * a1.w + u2 * a2.w + u3 * a3.w;
          }
          for (; r < R; r += nwarps) {
            float4 a0 = reinterpret_cast<const float4*>(wb + (int64_t)r * ld)[k];
            float u0 = ush[r];
            acc.x += u0 * a0.x; acc.y += u0 * a0.y; acc.z += u0 * a0.z; acc.w += u0 * a0.w;
          }
          atomicAdd(&qsh[4 * k], acc.x); atomicAdd(&qsh[4 * k + 1], acc.y); atomicAdd(&qsh[4 * k + 2], acc.z); atomicAdd(&qsh[4 * k + 3], acc.w);
        }
      }
    } else {
      for (int kb = 0; kb < kw; kb += 32) {
        const int k = kb + lane;
        if (k < kw) {
          float s = 0.f;
          for (int r = warp; r < R; r += nwarps) s += ush[r] * wb[(int64_t)r * ld + k];
          atomicAdd(&qsh[k], s);
        }
      }
    }
  };
  matvec(0);
  sn_group_barrier(bar, epoch, size);
  if (!a.training) {
    // sigma = u0 . (W v0), same for all G calls
    if (rank == 0) {
      float s = 0.f;
      for (int r = threadIdx.x; r < R; r += blockDim.x) s += a.u[r] * __ldcg(t + r);
      s = warp_sum(s);
      if (lane == 0) red[warp] = s;
      __syncthreads();
      if (threadIdx.x == 0) { float tot = 0.f; for (int w = 0; w < nwarps; ++w) tot += red[w]; red[0] = 1.0f / tot; }
      __syncthreads();
      float inv = red[0];
      for (int g = threadIdx.x; g < a.G; g += blockDim.x) a.inv_sigma[g] = inv;
      for (int i = threadIdx.x; i < a.G * R; i += blockDim.x) a.u_hist[i] = a.u[i % R];
    }
    for (int g = 0; g < a.G; ++g)
      for (int k = threadIdx.x; k < kw; k += blockDim.x) a.v_hist[(int64_t)g * K + k0 + k] = vsh[k];
    return;
  }
  for (int g = 0; g < a.G; ++g) {
    // u_g = normalize(t[g])   (every CTA recomputes it; R <= 768)
    const float* tg = t + (int64_t)g * R;
    float nrm = sqrtf(block_sumsq(tg, R, red));
    float inv = 1.0f / fmaxf(nrm, a.eps);
    for (int r = threadIdx.x; r < R; r += blockDim.x) ush[r] = __ldcg(tg + r) * inv;
    for (int k = threadIdx.x; k < kwmax; k += blockDim.x) qsh[k] = 0.f;
    __syncthreads();
    if (rank == 0) for (int r = threadIdx.x; r < R; r += blockDim.x) a.u_hist[(int64_t)g * R + r] = ush[r];
    colsum();
    __syncthreads();
    float part = 0.f;
    for (int k = threadIdx.x; k < kw; k += blockDim.x) part += qsh[k] * qsh[k];
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) { float tot = 0.f; for (int w = 0; w < nwarps; ++w) tot += red[w]; atomicAdd(&nq[g], tot); }
    sn_group_barrier(bar, epoch, size);
    float qn = sqrtf(__ldcg(nq + g));
    float qinv = 1.0f / fmaxf(qn, a.eps);
    for (int k = threadIdx.x; k < kw; k += blockDim.x) { float vv = qsh[k] * qinv; vsh[k] = vv; a.v_hist[(int64_t)g * K + k0 + k] = vv; }
    __syncthreads();
    matvec(g + 1);
    sn_group_barrier(bar, epoch, size);
    if (rank == 0) {
      const float* tn = t + (int64_t)(g + 1) * R;
      float s = 0.f;
      for (int r = threadIdx.x; r < R; r += blockDim.x) s += ush[r] * __ldcg(tn + r);
      s = warp_sum(s);
      __syncthreads();
      if (lane == 0) red[warp] = s;
      __syncthreads();
      if (threadIdx.x == 0) { float tot = 0.f; for (int w = 0; w < nwarps; ++w) tot += red[w]; a.inv_sigma[g] = 1.0f / tot; }
      __syncthreads();
    }
  }
  // persist final u, v
  if (rank == 0) for (int r = threadIdx.x; r < R; r += blockDim.x) a.u[r] = ush[r];
  for (int k = threadIdx.x; k < kw; k += blockDim.x) a.v[k0 + k] = vsh[k];
}
__global__ void __launch_bounds__(1024) sn_multi_kernel(const __grid_constant__ SnMultiArgs args) {
  extern __shared__ float smem[];
  const int bid = blockIdx.x;
  for (int i = 0; i < args.n; ++i) {
    const SnItem& it = args.it[i];
    if (bid >= it.cta_begin && bid < it.cta_begin + it.cta_count) {
      sn_run(it, bid - it.cta_begin, smem);
      return;
    }
  }
}
__global__ void sn_bwd_kernel(const float* __restrict__ dis, const float* __restrict__ is, const float* __restrict__ uh, const float* __restrict__ vh,
                              float* __restrict__ dw, int R, int K, int G, int acc) {
  int64_t total = (int64_t)R * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = i % K; int r = i / K;
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += (-dis[g] * is[g] * is[g]) * uh[(int64_t)g * R + r] * vh[(int64_t)g * K + k];
    if (acc) dw[i] += s; else dw[i] = s;
  }
}
constexpr int kSnBwdMaxItems = 48;
struct SnBwdMultiArgs { dgmr_sn_bwd_item it[kSnBwdMaxItems]; };
// the rank-G corrections of many weights in one launch: blockIdx.y = weight.  dw[r][k] (+)= sum_g coef[g] * u_g[r] * v_g[k]: the per-call
// coefficients sit in shared memory and each thread owns four consecutive k (one 16-byte access of dw and of every v_g) -- the pass should cost
// about one read + one write of dw
constexpr int kSnBwdMaxG = 64;
__global__ void __launch_bounds__(256) sn_bwd_multi_kernel(const __grid_constant__ SnBwdMultiArgs a) {
  const dgmr_sn_bwd_item& t = a.it[blockIdx.y];
  __shared__ float coef[kSnBwdMaxG];
  const bool fast = t.G <= kSnBwdMaxG && (t.K & 3) == 0 && ((reinterpret_cast<uintptr_t>(t.dw) | reinterpret_cast<uintptr_t>(t.v_hist)) & 15) == 0;
  if (fast) {
    for (int g = threadIdx.x; g < t.G; g += blockDim.x) { const float is = t.inv_sigma[g]; coef[g] = -t.d_inv_sigma[g] * is * is; }
    __syncthreads();
    const int K4 = t.K >> 2;
    const int64_t total4 = (int64_t)t.R * K4;
    float4* dw4 = reinterpret_cast<float4*>(t.dw);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / K4), k4 = (int)(i - (int64_t)r * K4);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < t.G; ++g) {
        const float c = coef[g] * t.u_hist[(int64_t)g * t.R + r];
        const float4 v = reinterpret_cast<const float4*>(t.v_hist + (int64_t)g * t.K)[k4];
        s.x += c * v.x; s.y += c * v.y; s.z += c * v.z; s.w += c * v.w;
      }
      if (t.accumulate) { const float4 o = dw4[i]; s.x = o.x + s.x; s.y = o.y + s.y; s.z = o.z + s.z; s.w = o.w + s.w; }
      dw4[i] = s;
    }
    return;
  }
  const int64_t total = (int64_t)t.R * t.K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % t.K), r = (int)(i / t.K);
    float s = 0.f;
    for (int g = 0; g < t.G; ++g) {
      const float is = t.inv_sigma[g];
      s += (-t.d_inv_sigma[g] * is * is) * t.u_hist[(int64_t)g * t.R + r] * t.v_hist[(int64_t)g * t.K + k];
    }
    if (t.accumulate) t.dw[i] += s; else t.dw[i] = s;
  }
}

// exposed to conv_umma.cu / api
int launch_conv_simt_fwd(const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y,
                         int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G, int act, cudaStream_t st) {
  ConvDims d{N, D, H, W, Cin, Cout, kd, kh, kw, G};
  int64_t M = (int64_t)N * D * H * W;
  dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(Cout, 64));
  conv_simt_fwd_kernel<64, 64, 16><<<grid, 256, 0, st>>>(x, wp, bias, scale, res, y, d, act);
  DGMR_CHECK_LAUNCH("conv_simt_fwd");
  return 0;
}
int launch_conv_simt_wgrad(const float* x, const float* dz, float* dwp, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, cudaStream_t st) {
  ConvDims d{N, D, H, W, Cin, Cout, kd, kh, kw, 1};
  int taps = kd * kh * kw;
  int64_t M = (int64_t)N * D * H * W;
  int co_tiles = (int)ceil_div(Cout, 64), ci_tiles = (int)ceil_div(Cin, 64);
  int64_t base = (int64_t)co_tiles * ci_tiles * taps;
  int64_t want = (int64_t)sm_count() * 4;
  int64_t ksplit = ceil_div(want, base);
  int64_t max_split = ceil_div(M, 256);
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > 65535) ksplit = 65535;
  int64_t chunk = ceil_div(ceil_div(M, ksplit), 16) * 16;
  ksplit = ceil_div(M, chunk);
  if (cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)taps * Cout * Cin, st) != cudaSuccess) { set_error("conv_simt_wgrad: memset failed"); return 2; }
  dim3 grid((unsigned)(co_tiles * ci_tiles), (unsigned)taps, (unsigned)ksplit);
  conv_simt_wgrad_kernel<64, 64, 16><<<grid, 256, 0, st>>>(x, dz, dwp, d, ci_tiles, chunk);
  DGMR_CHECK_LAUNCH("conv_simt_wgrad");
  return 0;
}

}  // namespace dgmr

using namespace dgmr;

extern "C" {

int dgmr_pack_weight(const float* w, float* packed, int Cout, int CinTot, int ci0, int Cin, int taps, int mode, dgmr_stream_t stream) {
  const int rnd = (mode & DGMR_FLAG_ROUND_TF32) ? 1 : 0;
  mode &= ~DGMR_FLAG_ROUND_TF32;
  DGMR_REQUIRE(ci0 >= 0 && ci0 + Cin <= CinTot && (mode == 0 || mode == 1), "dgmr_pack_weight: bad slice/mode");
  int64_t total = (int64_t)taps * Cout * Cin;
  if (total == 0) return 0;
  pack_weight_kernel<<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(w, packed, Cout, CinTot, ci0, Cin, taps, mode, rnd);
  DGMR_CHECK_LAUNCH("dgmr_pack_weight");
  return 0;
}
int dgmr_pack_weight_multi(const dgmr_pack_item* items, int n, dgmr_stream_t stream) {
  DGMR_REQUIRE(n >= 1 && items != nullptr, "dgmr_pack_weight_multi: empty");
  for (int base = 0; base < n; base += kPackMaxItems) {
    const int m = n - base < kPackMaxItems ? n - base : kPackMaxItems;
    PackMultiArgs args;
    int64_t biggest = 0;
    for (int i = 0; i < m; ++i) {
      const dgmr_pack_item& t = items[base + i];
      const int mode = t.mode & ~DGMR_FLAG_ROUND_TF32;
      DGMR_REQUIRE(t.w && t.packed && t.Cout > 0 && t.Cin > 0 && t.taps > 0 && t.ci0 >= 0 && t.ci0 + t.Cin <= t.CinTot && t.CinPad >= t.Cin && t.co0 >= 0 &&
                   t.co0 + t.Cout <= t.CoutTot && (mode == 0 || mode == 1), "dgmr_pack_weight_multi: bad item %d", base + i);
      args.it[i] = t;
      const int64_t tot = (int64_t)t.taps * t.Cout * t.Cin;
      if (tot > biggest) biggest = tot;
    }
    for (int i = m; i < kPackMaxItems; ++i) args.it[i] = args.it[0];
    int gx = (int)ceil_div(biggest, (int64_t)256 * 8);
    if (gx < 1) gx = 1;
    if (gx > 4 * sm_count()) gx = 4 * sm_count();
    pack_weight_multi_kernel<<<dim3((unsigned)gx, (unsigned)m), 256, 0, S(stream)>>>(args);
    DGMR_CHECK_LAUNCH("dgmr_pack_weight_multi");
  }
  return 0;
}
int dgmr_round_tf32(const float* x, float* y, int64_t n, dgmr_stream_t stream) {
  if (n == 0) return 0;
  DGMR_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0, "dgmr_round_tf32: pointers must be 16-byte aligned");
  int64_t n4 = n / 4;
  if (n4) {
    round_tf32_kernel<<<ew_grid(n4, 256, 2), 256, 0, S(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n4);
    DGMR_CHECK_LAUNCH("dgmr_round_tf32");
  }
  if (n % 4) { round_tf32_tail_kernel<<<1, 4, 0, S(stream)>>>(x, y, n4 * 4, n); DGMR_CHECK_LAUNCH("dgmr_round_tf32_tail"); }
  return 0;
}
int dgmr_unpack_wgrad(const float* packed, float* gw, int Cout, int CinTot, int ci0, int Cin, int taps, int accumulate, dgmr_stream_t stream) {
  DGMR_REQUIRE(ci0 >= 0 && ci0 + Cin <= CinTot, "dgmr_unpack_wgrad: bad slice");
  int64_t total = (int64_t)taps * Cout * Cin;
  if (total == 0) return 0;
  unpack_wgrad_kernel<<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(packed, gw, Cout, CinTot, ci0, Cin, taps, accumulate);
  DGMR_CHECK_LAUNCH("dgmr_unpack_wgrad");
  return 0;
}
int dgmr_rowdot_div(const float* a, const float* b, const float* denom, float* out, int rows, int64_t cols, int64_t ld, int64_t offset, dgmr_stream_t stream) {
  DGMR_REQUIRE(rows > 0 && cols > 0 && ld >= cols && offset >= 0 && offset + cols <= ld, "dgmr_rowdot_div: bad geometry");
  rowdot_div_kernel<<<rows, 256, 0, S(stream)>>>(a, b, denom, out, cols, ld, offset);
  DGMR_CHECK_LAUNCH("dgmr_rowdot_div");
  return 0;
}
int dgmr_conv_bwd_prep(const float* dy, const float* y, const float* res, const float* bias, const float* scale, float* dz, float* dpre, float* dbias,
                       float* dscale, int64_t rows, int G, int Cout, int act, int accumulate_dbias, int up_h, int up_w, int pool_d, int pool_h, int pool_w,
                       int D, int H, int W, dgmr_stream_t stream) {
  const int rnd = (act & DGMR_FLAG_ROUND_TF32) ? 1 : 0;
  act &= ~DGMR_FLAG_ROUND_TF32;
  PoolGeom pg{pool_d, pool_h, pool_w, D, H, W, 0.f, -1, -1};
  if (W > 0 && (W & (W - 1)) == 0) { pg.lw = 0; while ((1 << pg.lw) < W) ++pg.lw; }
  if (H > 0 && (H & (H - 1)) == 0) { pg.lh = 0; while ((1 << pg.lh) < H) ++pg.lh; }
  if (pool_d > 0) {
    DGMR_REQUIRE(pool_h > 0 && pool_w > 0 && pool_d <= 2 && pool_h <= 2 && pool_w <= 2 && D > 0 && H > 0 && W > 0 && (rows * G) % ((int64_t)D * H * W) == 0,
                 "dgmr_conv_bwd_prep: bad pooled-gradient geometry");
    pg.inv = 1.0f / (float)(pool_d * pool_h * pool_w);
  }
  DGMR_REQUIRE((up_h == 0 && up_w == 0) || (up_h > 0 && up_w > 0 && up_h % 2 == 0 && up_w % 2 == 0 && (rows * G) % ((int64_t)up_h * up_w) == 0),
               "dgmr_conv_bwd_prep: bad half-resolution residual geometry %d x %d", up_h, up_w);
  DGMR_REQUIRE(rows > 0 && G > 0 && Cout > 0, "dgmr_conv_bwd_prep: bad dims");
  DGMR_REQUIRE(!(dscale && !scale), "dgmr_conv_bwd_prep: dscale requested without scale");
  DGMR_REQUIRE(!((act == DGMR_ACT_RELU || dscale) && !y), "dgmr_conv_bwd_prep: y required");
  if (dbias && !accumulate_dbias) DGMR_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, S(stream)));
  if (dscale) DGMR_CUDA(cudaMemsetAsync(dscale, 0, sizeof(float) * (size_t)G * Cout, S(stream)));
  int64_t bpg = (int64_t)sm_count() * 16 / G; if (bpg < 1) bpg = 1;
  int64_t chunk = ceil_div(rows, bpg); if (chunk < 64) chunk = 64;
  if (ceil_div(rows, chunk) * G < sm_count()) { chunk = ceil_div(rows * G, (int64_t)sm_count()); if (chunk < 8) chunk = 8; }   // small tensors (ConvGRU steps): fill the SMs
  dim3 grid((unsigned)ceil_div(rows, chunk), G);
  const bool hasy = y != nullptr, dsr = res != nullptr && dscale != nullptr, geo = pool_d > 0 || (up_w > 0 && dsr);
  DGMR_REQUIRE(!geo || rows * G < ((int64_t)1 << 31), "dgmr_conv_bwd_prep: pooled / half-resolution maps need G*rows < 2^31");
#define DGMR_PREP(V_, Y_, R_, G_) conv_bwd_prep_kernel<V_, Y_, R_, G_><<<grid, 256, 256 * 2 * V_ * sizeof(double), S(stream)>>>( \
      dy, y, res, bias, scale, dz, dpre, dbias, dscale, rows, Cout, chunk, act, rnd, up_h, up_w, pg)
  if (Cout % 4 != 0) DGMR_PREP(1, true, true, true);
  else if (!hasy) { if (geo) DGMR_PREP(4, false, false, true); else DGMR_PREP(4, false, false, false); }
  else if (!dsr) { if (geo) DGMR_PREP(4, true, false, true); else DGMR_PREP(4, true, false, false); }
  else { if (geo) DGMR_PREP(4, true, true, true); else DGMR_PREP(4, true, true, false); }
#undef DGMR_PREP
  DGMR_CHECK_LAUNCH("dgmr_conv_bwd_prep");
  return 0;
}

static int sn_launch(SnMultiArgs& args, int total_ctas, size_t smem, cudaStream_t st) {
  static int max_smem = 0, coop = -1;
  if (coop < 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  if (coop != 1) { set_error("dgmr_sn_power_iter: device lacks cooperative launch"); return 1; }
  if (smem > (size_t)max_smem) { set_error("dgmr_sn_power_iter: needs %zu B of shared memory (> %d)", smem, max_smem); return 1; }
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(sn_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { set_error("dgmr_sn_power_iter: smem attribute"); return 2; }
    smem_set = smem;
  }
  void* kargs[] = {&args};
  // cooperative launch = all CTAs co-resident, which the group barriers rely on
  cudaError_t e = cudaLaunchCooperativeKernel((void*)sn_multi_kernel, dim3(total_ctas), dim3(1024), kargs, smem, st);
  if (e != cudaSuccess) { set_error("dgmr_sn_power_iter: cooperative launch failed: %s (ctas=%d smem=%zu)", cudaGetErrorString(e), total_ctas, smem); return 2; }
  return 0;
}
static size_t sn_item_smem(const SnItem& it) { return (size_t)(it.R + 4 + 2 * it.kw + 32) * 4 + (it.in_smem ? (size_t)it.R * it.kw * 4 : 0); }

int dgmr_sn_power_iter(const float* w, float* u, float* v, int R, int K, int G, float eps, int training, float* inv_sigma, float* u_hist, float* v_hist,
                       float* ws, dgmr_stream_t stream) {
  DGMR_REQUIRE(R > 0 && K > 0 && G > 0, "dgmr_sn_power_iter: bad dims");
  const int64_t budget = 200 * 1024 - (int64_t)(R + 32) * 4 - 1024;
  int sms = sm_count();
  int nb = (int)ceil_div((int64_t)R * K, 32768);
  if (nb < 1) nb = 1;
  if (nb > sms) nb = sms;
  if (nb > K) nb = K;
  int kw = (int)(ceil_div(ceil_div(K, nb), 4) * 4);
  int in_smem = 1;
  while (((int64_t)R * kw + 2 * kw) * 4 > budget && nb < sms && nb < K) { ++nb; kw = (int)(ceil_div(ceil_div(K, nb), 4) * 4); }
  if (((int64_t)R * kw + 2 * kw) * 4 > budget) in_smem = 0;
  nb = (int)ceil_div(K, kw);
  SnMultiArgs args;
  args.n = 1;
  SnItem& a = args.it[0];
  a.w = w; a.u = u; a.v = v; a.R = R; a.K = K; a.G = G; a.eps = eps; a.training = training;
  a.inv_sigma = inv_sigma; a.u_hist = u_hist; a.v_hist = v_hist; a.ws = ws;
  a.cta_begin = 0; a.cta_count = nb; a.kw = kw; a.in_smem = in_smem;
  DGMR_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * ((size_t)(G + 1) * R + G + 1), S(stream)));
  return sn_launch(args, nb, sn_item_smem(a), S(stream));
}

int dgmr_sn_power_iter_multi(const dgmr_sn_item* items, int n, dgmr_stream_t stream) {
  DGMR_REQUIRE(n >= 1, "dgmr_sn_power_iter_multi: empty");
  const int sms = sm_count();
  for (int base = 0; base < n; base += kSnMaxItems) {
    const int m = n - base < kSnMaxItems ? n - base : kSnMaxItems;
    SnMultiArgs args;
    args.n = m;
    double total = 0.0;
    for (int i = 0; i < m; ++i) {
      const dgmr_sn_item& s = items[base + i];
      DGMR_REQUIRE(s.R > 0 && s.K > 0 && s.G > 0, "dgmr_sn_power_iter_multi: bad dims in item %d", base + i);
      total += (double)s.R * s.K * (s.training ? s.G : 1);
    }
    DGMR_REQUIRE(m <= sms, "dgmr_sn_power_iter_multi: more weights than SMs");
    // CTAs in proportion to the work, at least one each, never more than the SM count in total (co-residency)
    int used = 0;
    for (int i = 0; i < m; ++i) {
      const dgmr_sn_item& s = items[base + i];
      double c = (double)s.R * s.K * (s.training ? s.G : 1);
      int cnt = (int)((sms - m) * c / total) + 1;
      if (cnt > s.K) cnt = s.K;
      args.it[i].cta_count = cnt;
      used += cnt;
    }
    size_t smem = 0;
    int begin = 0;
    for (int i = 0; i < m; ++i) {
      const dgmr_sn_item& s = items[base + i];
      SnItem& a = args.it[i];
      a.w = s.w; a.u = s.u; a.v = s.v; a.inv_sigma = s.inv_sigma; a.u_hist = s.u_hist; a.v_hist = s.v_hist; a.ws = s.ws;
      a.R = s.R; a.K = s.K; a.G = s.G; a.training = s.training; a.eps = s.eps;
      a.kw = (int)(ceil_div(ceil_div(s.K, a.cta_count), 4) * 4);
      a.cta_count = (int)ceil_div(s.K, a.kw);   // drop CTAs that would own no column
      a.cta_begin = begin; begin += a.cta_count;
      a.in_smem = (((int64_t)a.R * a.kw + 2 * a.kw + a.R + 32) * 4 <= 160 * 1024) ? 1 : 0;
      size_t need = sn_item_smem(a);
      if (need > smem) smem = need;
    }
    (void)used;
    int e = sn_launch(args, begin, smem, S(stream));
    if (e) return e;
  }
  return 0;
}
int dgmr_sn_bwd(const float* d_inv_sigma, const float* inv_sigma, const float* u_hist, const float* v_hist, float* dw, int R, int K, int G, int accumulate,
                dgmr_stream_t stream) {
  int64_t total = (int64_t)R * K;
  sn_bwd_kernel<<<ew_grid(total, 256, 2), 256, 0, S(stream)>>>(d_inv_sigma, inv_sigma, u_hist, v_hist, dw, R, K, G, accumulate);
  DGMR_CHECK_LAUNCH("dgmr_sn_bwd");
  return 0;
}
int dgmr_sn_bwd_multi(const dgmr_sn_bwd_item* items, int n, dgmr_stream_t stream) {
  DGMR_REQUIRE(n >= 1 && items != nullptr, "dgmr_sn_bwd_multi: empty");
  for (int base = 0; base < n; base += kSnBwdMaxItems) {
    const int m = n - base < kSnBwdMaxItems ? n - base : kSnBwdMaxItems;
    SnBwdMultiArgs args;
    int64_t biggest = 0;
    for (int i = 0; i < m; ++i) {
      const dgmr_sn_bwd_item& t = items[base + i];
      DGMR_REQUIRE(t.d_inv_sigma && t.inv_sigma && t.u_hist && t.v_hist && t.dw && t.R > 0 && t.K > 0 && t.G > 0, "dgmr_sn_bwd_multi: bad item %d", base + i);
      args.it[i] = t;
      if ((int64_t)t.R * t.K > biggest) biggest = (int64_t)t.R * t.K;
    }
    for (int i = m; i < kSnBwdMaxItems; ++i) args.it[i] = args.it[0];
    int gx = (int)ceil_div(biggest, (int64_t)256 * 8);
    if (gx < 1) gx = 1;
    if (gx > 4 * sm_count()) gx = 4 * sm_count();
    sn_bwd_multi_kernel<<<dim3((unsigned)gx, (unsigned)m), 256, 0, S(stream)>>>(args);
    DGMR_CHECK_LAUNCH("dgmr_sn_bwd_multi");
  }
  return 0;
}

}  // extern "C"
