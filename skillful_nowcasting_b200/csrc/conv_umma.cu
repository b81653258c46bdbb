// tcgen05 / TMA implicit-GEMM convolution for sm_100a (forward and dgrad; wgrad variant below).
//
//   y[pixel][co] = act( (sum_{tap,ci} x[pixel+tap][ci] * wp[tap][co][ci]) * scale[g][co] + bias[co] + res )
//
// B200 mapping (no im2col, no cuDNN):
//   * A operand: the channels-last activation [N,D,H,W,C] is described to TMA as a 5-D tensor
//     {C,W,H,D,N}.  One pipeline stage = one filter tap x BK channels: a TMA box {BK,bw,bh,1,bn}
//     (bw*bh*bn = 128 output pixels) at coordinate (c0, w0+kw-pw, h0+kh-ph, d0+kd-pd, n0).  The
//     halo / zero padding is TMA's out-of-bounds zero fill, so the 128xBK tile lands in shared memory
//     already in the K-major swizzled layout tcgen05.mma consumes.
//   * B operand: packed weights [tap][Cout][Cin] as a 3-D tensor, box {BK,BN,1}.
//   * D accumulator: 128 lanes x BN fp32 columns in TMEM; kind::tf32, M=128, N=BN<=256, K=8 per MMA.
//   * warp roles: warp0 = TMA producer, warp1 = MMA issuer (single elected thread),
//     warps 2-5 = epilogue (tcgen05.ld 32x32b, fused scale/bias/residual/ReLU, NHWC float4 stores).
//   * mbarrier full/empty ring between TMA and MMA; tcgen05.commit frees stages and signals the epilogue.
#include "umma_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace dgmr {

int launch_conv_simt_fwd(const float*, const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int launch_conv_simt_wgrad(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, cudaStream_t);
bool umma_kwstack_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G);   // conv_kwstack.cu
extern int g_kwstack_pair;
extern int g_subpix_rows;
bool umma_pairconv_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G);   // conv_kwstack.cu, STACK = false
int launch_conv_umma_pairconv(const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int D, int H, int W,
                              int Cin, int Cout, int kd, int G, int act, cudaStream_t st);
int launch_conv_umma_kwstack(const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int D, int H, int W,
                             int Cin, int Cout, int kd, int G, int act, cudaStream_t st);

struct UmmaConvParams {
  int N, D, H, W, Cin, Cout, kd, kh, kw, G;
  int bw, bh, bn;        // spatial box: bw*bh*bn == 128
  int tiles_w, tiles_h;  // W/bw, H/bh
  int BK;                // channels per stage (32/16/8)
  int BN;                // N tile (multiple of 16, <= 256)
  int stages;
  int cg;                // stages per release group (stages % cg == 0): one tcgen05.commit hands cg stages back to the producer
  int tmem_cols;
  int act;
  int split_taps;        // 1: blockIdx.z = filter tap, epilogue red.adds acc*scale into y
  int n_tiles;           // persistent variant: Cout tiles
  int round_out;         // DGMR_FLAG_ROUND_OUT: y written tf32-rounded
  int res_up2;           // DGMR_FLAG_RES_UP2: res is [N,D,H/2,W/2,Cout], read at (h/2, w/2)
  const float* bias; const float* scale; const float* res; float* y;
};

constexpr int kUmmaThreads = 192;  // warp0 TMA, warp1 MMA, warps 2..5 epilogue

// X3: 3xTF32 error-compensated operands (DGMR_PREC_3XTF32).  Every stage carries the hi and the lo part of both tiles (tmA/tmB describe
// the hi tensors, tmAlo/tmBlo the lo tensors) and every k-step issues lo*hi, hi*lo, hi*hi into the same fp32 accumulator (small terms
// first); the dropped lo*lo term is ~2^-22 relative.
// EF: the epilogue flags (DGMR_FLAG_ROUND_OUT / DGMR_FLAG_RES_UP2) are compiled in.  They cost 8 registers and, measured, 10-28 % on the
// launches that are all epilogue (Cout <= 48, many tiles), so the common instantiation does not carry them.
template <int BK, bool X3, bool EF>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmAlo,
                     const __grid_constant__ CUtensorMap tmBlo, const UmmaConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-B alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = 128u * BK * 4u, b_bytes = (uint32_t)p.BN * BK * 4u;
  const uint32_t b_bytes_al = (b_bytes + 1023u) & ~1023u;
  const uint32_t half_bytes = a_bytes + b_bytes_al;          // X3: [A hi][B hi][A lo][B lo]
  const uint32_t stage_bytes = X3 ? 2u * half_bytes : half_bytes;
  const uint32_t bar_base = base + p.stages * stage_bytes;  // full[stages], empty[stages], tmem_full, tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.stages + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  // ---- tile coordinates
  int mt = blockIdx.x;
  const int tw_i = mt % p.tiles_w; mt /= p.tiles_w;
  const int th_i = mt % p.tiles_h; mt /= p.tiles_h;
  const int d0 = mt % p.D; mt /= p.D;
  const int n0 = mt * p.bn;
  const int w0 = tw_i * p.bw, h0 = th_i * p.bh;
  const int co0 = blockIdx.y * p.BN;
  const int taps = p.kd * p.kh * p.kw;
  const int kchunks = (p.Cin + BK - 1) / BK;   // the last chunk may be a channel tail (zero-filled by TMA, fewer k-steps)
  const int tap_begin = p.split_taps ? (int)blockIdx.z : 0;
  const int num_kb = (p.split_taps ? 1 : taps) * kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (X3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    }
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // Role loops are warp-uniform with an elected issuing lane (see elect_one).  Pipeline stages are handed back to the producer in
  // groups of p.cg: one tcgen05.commit per group, because a commit occupies the tensor pipe for ~780 cycles (dgmr_debug_umma_rate)
  // while the 4 MMAs of one stage take 240 cycles at N <= 128.
  if (warp == 0) {
    {
      // ===== TMA producer
      int s = 0, g = 0, sg = 0; uint32_t ph = 0;
      int tapl = 0, chunk = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int c0 = chunk * BK;
        const int tap = tap_begin + tapl;
        const int tkw = tap % p.kw, tkh = (tap / p.kw) % p.kh, tkd = tap / (p.kw * p.kh);
        if (sg == 0) mbar_wait(empty_bar(g), ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), (X3 ? 2u : 1u) * (a_bytes + b_bytes));
          const uint32_t sa = base + s * stage_bytes;
          tma_load_5d(sa, &tmA, full_bar(s), c0, w0 + tkw - p.kw / 2, h0 + tkh - p.kh / 2, d0 + tkd - p.kd / 2, n0);
          tma_load_3d(sa + a_bytes, &tmB, full_bar(s), c0, co0, tap);
          if (X3) {
            tma_load_5d(sa + half_bytes, &tmAlo, full_bar(s), c0, w0 + tkw - p.kw / 2, h0 + tkh - p.kh / 2, d0 + tkd - p.kd / 2, n0);
            tma_load_3d(sa + half_bytes + a_bytes, &tmBlo, full_bar(s), c0, co0, tap);
          }
        }
        __syncwarp();
        if (++chunk == kchunks) { chunk = 0; ++tapl; }
        if (++sg == p.cg) { sg = 0; ++g; }
        if (++s == p.stages) { s = 0; g = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      constexpr uint32_t row_bytes = BK * 4u;
      constexpr uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
      constexpr uint32_t dhi = desc_hi(8u * row_bytes, layout);     // descriptors as (low word, common high word): see umma_tf32_lo
      const int tail_ks = (p.Cin % BK) ? (p.Cin % BK) / 8 : BK / 8;
      const uint32_t lo0 = desc_lo(base), st16 = stage_bytes >> 4, a16 = a_bytes >> 4;
      const uint32_t lo_off = half_bytes >> 4;     // X3: the lo tiles sit half a stage further on
      int chunk_i = 0;
      int s = 0, g = 0, sg = 0; uint32_t ph = 0;
      uint32_t a_lo = lo0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t b_lo = a_lo + a16;
        const bool last_chunk = (++chunk_i == kchunks);   // last chunk of a tap: possibly a channel tail with fewer valid k-steps
        if (last_chunk) chunk_i = 0;
        const bool rel = (sg + 1 == p.cg) || (kb + 1 == num_kb);   // release this group (the final, possibly partial, one too)
        auto mma = [&](uint32_t ad, uint32_t bd, uint32_t acc) {
          if (X3) {
            umma_tf32_lo(tmem_base, ad + lo_off, bd, dhi, idesc, acc);
            umma_tf32_lo(tmem_base, ad, bd + lo_off, dhi, idesc, 1u);
            umma_tf32_lo(tmem_base, ad, bd, dhi, idesc, 1u);
          } else {
            umma_tf32_lo(tmem_base, ad, bd, dhi, idesc, acc);
          }
        };
        if (elect_one()) {
          // advance 8 tf32 = 32 bytes along K inside the swizzle row: +2 in the (addr>>4) field
          mma(a_lo, b_lo, kb != 0 ? 1u : 0u);
          if (last_chunk) {
#pragma unroll
            for (int k = 1; k < BK / 8; ++k)
              if (k < tail_ks) mma(a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), 1u);
          } else {
#pragma unroll
            for (int k = 1; k < BK / 8; ++k) mma(a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), 1u);
          }
          if (rel) umma_commit(empty_bar(g));
        }
        __syncwarp();
        a_lo += st16;
        if (++sg == p.cg) { sg = 0; ++g; }
        if (++s == p.stages) { s = 0; g = 0; ph ^= 1u; a_lo = lo0; }
      }
      if (elect_one()) umma_commit(tmem_full_bar);
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    const int r = q * 32 + lane;  // accumulator row = pixel within the tile
    const int wl = r % p.bw, hl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
    const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
    const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
    const int64_t m = (((int64_t)n * p.D + d0) * p.H + h) * p.W + w;
    const int g = valid ? n / (p.N / p.G) : 0;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool vec4 = (p.Cout & 3) == 0;
    if (vec4) {
      // Coalesced path (see the halo-patch kernel): transpose each 16-column chunk through 2 KB of the (now idle) pipeline
      // stage memory so that 4 adjacent lanes own 64 contiguous bytes of one output row.
      const uint32_t stg = base + (uint32_t)q * 2048u;
      const int lr = lane >> 2, lc = lane & 3;
      const uint32_t st_row = stg + (uint32_t)lane * 64u, st_sw = (uint32_t)((lane >> 1) & 3);
      // element offsets are 32-bit (host check: the output tensor has < 2^32 elements): half the registers of 64-bit offsets / pointers,
      // which is what keeps this kernel at 3 CTAs per SM with the epilogue flags compiled in
      uint32_t mrow[4], rrow[4], soff[4]; bool vrow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rj = q * 32 + lr + 8 * j;
        const int wj = w0 + rj % p.bw, hj = h0 + (rj / p.bw) % p.bh, nj = n0 + rj / (p.bw * p.bh);
        vrow[j] = (nj < p.N) && (hj < p.H) && (wj < p.W);
        mrow[j] = (uint32_t)((((nj * p.D + d0) * p.H + hj) * p.W + wj)) * (uint32_t)p.Cout + (uint32_t)(co0 + 4 * lc);
        rrow[j] = (EF && p.res_up2) ? (uint32_t)((((nj * p.D + d0) * (p.H >> 1) + (hj >> 1)) * (p.W >> 1) + (wj >> 1))) * (uint32_t)p.Cout + (uint32_t)(co0 + 4 * lc)
                                    : mrow[j];
        soff[j] = vrow[j] ? (uint32_t)(nj / (p.N / p.G)) * (uint32_t)p.Cout : 0u;
      }
      float4 rr[4];
      auto load_res = [&](int c, float4* dst) {
        if (p.res == nullptr || c >= p.BN || co0 + c + 4 * lc >= p.Cout) return;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (vrow[j]) dst[j] = __ldg(reinterpret_cast<const float4*>(p.res + rrow[j] + c));
      };
      load_res(0, rr);
      for (int c = 0; c < p.BN; c += 16) {
        if (co0 + c >= p.Cout) break;
        float4 rn[4];
        load_res(c + 16, rn);
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (((uint32_t)k ^ st_sw) << 4)), "f"(v[4 * k]), "f"(v[4 * k + 1]),
                       "f"(v[4 * k + 2]), "f"(v[4 * k + 3]) : "memory");
        __syncwarp();
        const int co = co0 + c + 4 * lc;
        const bool cok = co < p.Cout;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cok && p.bias) b4 = make_float4(__ldg(p.bias + co), __ldg(p.bias + co + 1), __ldg(p.bias + co + 2), __ldg(p.bias + co + 3));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = lr + 8 * j;
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                       : "r"(stg + (uint32_t)row * 64u + (((uint32_t)lc ^ (uint32_t)((row >> 1) & 3)) << 4)) : "memory");
          if (vrow[j] && cok) {
            if (p.scale) { const float* sr = p.scale + soff[j] + co; o.x *= __ldg(sr); o.y *= __ldg(sr + 1); o.z *= __ldg(sr + 2); o.w *= __ldg(sr + 3); }
            if (p.split_taps) {   // one filter tap per CTA: accumulate into y (pre-filled with the residual or zero)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.y + mrow[j] + c), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
              continue;
            }
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (p.res) { o.x += rr[j].x; o.y += rr[j].y; o.z += rr[j].z; o.w += rr[j].w; }
            if (p.act == DGMR_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (EF && p.round_out) o = rna_tf32_e4(o);
            *reinterpret_cast<float4*>(p.y + mrow[j] + c) = o;
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = rn[j];
      }
    } else
    for (int c = 0; c < p.BN; c += 16) {
      if (co0 + c >= p.Cout) break;   // warp-uniform
      float v[16];
      tmem_ld16(trow + (uint32_t)c, v);
      if (!valid) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int co = co0 + c + j;
        if (co < p.Cout) {
          float t = v[j];
          if (p.scale) t *= __ldg(p.scale + (int64_t)g * p.Cout + co);
          if (p.bias) t += __ldg(p.bias + co);
          v[j] = t;
        }
      }
      float* yp = p.y + m * p.Cout + co0 + c;
      const int64_t mres = (EF && p.res_up2) ? (((int64_t)n * p.D + d0) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) : m;
      const float* rp = p.res ? p.res + mres * p.Cout + co0 + c : nullptr;
      if (p.split_taps) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (co0 + c + j < p.Cout) atomicAdd(yp + j, v[j]);
        continue;
      }
      if (vec4) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (co0 + c + j < p.Cout) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (rp) { float4 rr = *reinterpret_cast<const float4*>(rp + j); o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
            if (p.act == DGMR_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (EF && p.round_out) o = rna_tf32_e4(o);
            *reinterpret_cast<float4*>(yp + j) = o;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (co0 + c + j < p.Cout) {
            float o = v[j];
            if (rp) o += rp[j];
            if (p.act == DGMR_ACT_RELU) o = fmaxf(o, 0.f);
            yp[j] = (EF && p.round_out) ? rna_tf32_e(o) : o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------ plain conv, persistent variant
// Same tiles and pipeline stages as conv_umma_fwd_kernel, but each CTA walks many tiles: barrier setup, TMEM allocation and the
// pipeline fill are paid once per CTA instead of once per 128-pixel tile, the K-block stream of consecutive tiles is continuous
// and the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.  For launches with
// many tiles and a short K loop (1x1 convs, narrow layers) the per-tile fixed cost was most of the time.
template <int BK, bool EF>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const UmmaConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = 128u * BK * 4u, b_bytes = (uint32_t)p.BN * BK * 4u;
  const uint32_t b_bytes_al = (b_bytes + 1023u) & ~1023u;
  const uint32_t stage_bytes = a_bytes + b_bytes_al;
  const uint32_t bar_base = base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tmem_full = [&](int bsel) { return bar_base + 8u * (2 * p.stages + bsel); };
  auto tmem_empty = [&](int bsel) { return bar_base + 8u * (2 * p.stages + 2 + bsel); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.stages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  const uint32_t epi_base = (tmem_ptr_addr + 8u + 127u) & ~127u;     // 4 epilogue warps x 2 KB transpose staging

  const int64_t mtiles = (int64_t)((p.N + p.bn - 1) / p.bn) * p.D * p.tiles_h * p.tiles_w;
  const int64_t total_tiles = mtiles * p.n_tiles;
  auto decode = [&](int64_t t, int& n0, int& d0, int& h0, int& w0, int& co0) {
    int64_t mt = t % mtiles; co0 = (int)(t / mtiles) * p.BN;
    const int tw_i = (int)(mt % p.tiles_w); mt /= p.tiles_w;
    const int th_i = (int)(mt % p.tiles_h); mt /= p.tiles_h;
    d0 = (int)(mt % p.D); mt /= p.D;
    n0 = (int)mt * p.bn; w0 = tw_i * p.bw; h0 = th_i * p.bh;
  };
  const int taps = p.kd * p.kh * p.kw;
  const int kchunks = (p.Cin + BK - 1) / BK;
  const int num_kb = taps * kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int bsel = 0; bsel < 2; ++bsel) { mbar_init(tmem_full(bsel), 1); mbar_init(tmem_empty(bsel), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===== TMA producer: one continuous stream of K blocks over all of this CTA's tiles
    int s = 0, g = 0, sg = 0; uint32_t ph = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n0, d0, h0, w0, co0; decode(t, n0, d0, h0, w0, co0);
      int tap = 0, chunk = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int c0 = chunk * BK;
        const int tkw = tap % p.kw, tkh = (tap / p.kw) % p.kh, tkd = tap / (p.kw * p.kh);
        if (sg == 0) mbar_wait(empty_bar(g), ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), a_bytes + b_bytes);
          const uint32_t sa = base + s * stage_bytes;
          tma_load_5d(sa, &tmA, full_bar(s), c0, w0 + tkw - p.kw / 2, h0 + tkh - p.kh / 2, d0 + tkd - p.kd / 2, n0);
          tma_load_3d(sa + a_bytes, &tmB, full_bar(s), c0, co0, tap);
        }
        __syncwarp();
        if (++chunk == kchunks) { chunk = 0; ++tap; }
        if (++sg == p.cg) { sg = 0; ++g; }
        if (++s == p.stages) { s = 0; g = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t row_bytes = BK * 4u;
    constexpr uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    constexpr uint32_t dhi = desc_hi(8u * row_bytes, layout);     // descriptors as (low word, common high word): see umma_tf32_lo
    const int tail_ks = (p.Cin % BK) ? (p.Cin % BK) / 8 : BK / 8;
    const uint32_t lo0 = desc_lo(base), st16 = stage_bytes >> 4, a16 = a_bytes >> 4;
    int s = 0, g = 0, sg = 0; uint32_t ph = 0, it = 0;
    uint32_t a_lo = lo0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int bsel = it & 1; const uint32_t phacc = (it >> 1) & 1u;
      const bool last_tile = t + gridDim.x >= total_tiles;
      mbar_wait(tmem_empty(bsel), phacc ^ 1u);       // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(bsel * p.BN);
      int chunk_i = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t b_lo = a_lo + a16;
        const bool last_chunk = (++chunk_i == kchunks);
        if (last_chunk) chunk_i = 0;
        const bool last_kb = (kb + 1 == num_kb);
        const bool rel = (sg + 1 == p.cg) || (last_kb && last_tile);   // group full, or the very last K block
        if (elect_one()) {
          umma_tf32_lo(tacc, a_lo, b_lo, dhi, idesc, kb != 0 ? 1u : 0u);
          if (last_chunk) {
#pragma unroll
            for (int k = 1; k < BK / 8; ++k)
              if (k < tail_ks) umma_tf32_lo(tacc, a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), dhi, idesc, 1u);
          } else {
#pragma unroll
            for (int k = 1; k < BK / 8; ++k) umma_tf32_lo(tacc, a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), dhi, idesc, 1u);
          }
          if (rel) umma_commit(empty_bar(g));
          if (last_kb) umma_commit(tmem_full(bsel));
        }
        __syncwarp();
        a_lo += st16;
        if (++sg == p.cg) { sg = 0; ++g; }
        if (++s == p.stages) { s = 0; g = 0; ph ^= 1u; a_lo = lo0; }
      }
    }
  } else {
    // ===== epilogue (4 warps): warp w may only touch TMEM lanes [32*(w%4), +32); coalesced through a 2 KB staging tile per warp
    const int q = warp & 3;
    const uint32_t stg = epi_base + (uint32_t)q * 2048u;
    const int lr = lane >> 2, lc = lane & 3;
    const uint32_t st_row = stg + (uint32_t)lane * 64u, st_sw = (uint32_t)((lane >> 1) & 3);
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      int n0, d0, h0, w0, co0; decode(t, n0, d0, h0, w0, co0);
      const int bsel = it & 1; const uint32_t phacc = (it >> 1) & 1u;
      int64_t mrow[4], rrow[4]; bool vrow[4]; const float* srow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rj = q * 32 + lr + 8 * j;
        const int wj = w0 + rj % p.bw, hj = h0 + (rj / p.bw) % p.bh, nj = n0 + rj / (p.bw * p.bh);
        vrow[j] = (nj < p.N) && (hj < p.H) && (wj < p.W);
        mrow[j] = ((((int64_t)nj * p.D + d0) * p.H + hj) * p.W + wj) * p.Cout + co0 + 4 * lc;
        rrow[j] = (EF && p.res_up2) ? ((((int64_t)nj * p.D + d0) * (p.H >> 1) + (hj >> 1)) * (p.W >> 1) + (wj >> 1)) * p.Cout + co0 + 4 * lc : mrow[j];
        srow[j] = (p.scale && vrow[j]) ? p.scale + (int64_t)(nj / (p.N / p.G)) * p.Cout : nullptr;
      }
      float4 rr[4];
      auto load_res = [&](int c, float4* dst) {
        if (p.res == nullptr || c >= p.BN || co0 + c + 4 * lc >= p.Cout) return;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (vrow[j]) dst[j] = __ldg(reinterpret_cast<const float4*>(p.res + rrow[j] + c));
      };
      load_res(0, rr);
      mbar_wait(tmem_full(bsel), phacc);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(bsel * p.BN);
      for (int c = 0; c < p.BN; c += 16) {
        if (co0 + c >= p.Cout) break;
        float4 rn[4];
        load_res(c + 16, rn);
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (((uint32_t)k ^ st_sw) << 4)), "f"(v[4 * k]), "f"(v[4 * k + 1]),
                       "f"(v[4 * k + 2]), "f"(v[4 * k + 3]) : "memory");
        __syncwarp();
        const int co = co0 + c + 4 * lc;
        const bool cok = co < p.Cout;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cok && p.bias) b4 = make_float4(__ldg(p.bias + co), __ldg(p.bias + co + 1), __ldg(p.bias + co + 2), __ldg(p.bias + co + 3));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = lr + 8 * j;
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                       : "r"(stg + (uint32_t)row * 64u + (((uint32_t)lc ^ (uint32_t)((row >> 1) & 3)) << 4)) : "memory");
          if (vrow[j] && cok) {
            if (srow[j]) { o.x *= __ldg(srow[j] + co); o.y *= __ldg(srow[j] + co + 1); o.z *= __ldg(srow[j] + co + 2); o.w *= __ldg(srow[j] + co + 3); }
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (p.res) { o.x += rr[j].x; o.y += rr[j].y; o.z += rr[j].z; o.w += rr[j].w; }
            if (p.act == DGMR_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (EF && p.round_out) o = rna_tf32_e4(o);
            *reinterpret_cast<float4*>(p.y + mrow[j] + c) = o;
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = rn[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty(bsel)) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// ------------------------------------------------------------------ 3x3 convolution, halo-patch variant
// The plain kernel above re-reads the activation tile once per filter tap (9x L2->SM amplification), which is what bounds
// the high-resolution, low-channel layers (L2 bandwidth, not the tensor pipe).  Here the image is addressed in a padded,
// flattened coordinate f = (h+1)*P + (w+1), P = W+2: the 128 (or 256) outputs of a tile are CONSECUTIVE f, the inputs of
// filter tap (kh,kw) are the same run shifted by (kh-1)*P + (kw-1) rows, so ONE TMA box of whole padded image rows (zero
// borders = TMA out-of-bounds fill) per 32-channel chunk feeds all 9 taps: the tcgen05 A descriptor simply starts at a
// different 128-byte row of the swizzled patch (start addresses need not be swizzle-atom aligned, verified on hardware by
// dgmr_debug_umma_shift).  Outputs that fall on the 2 padding columns are computed and discarded (2/P of the rows).
// Persistent CTAs, separate activation-patch and weight rings, optionally two 128-row sub-tiles per weight stage (halves the
// weight traffic) and double-buffered TMEM accumulators so the epilogue of one tile overlaps the MMAs of the next.
struct PatchConvParams {
  int N, D, H, W, Cin, Cout, kd, G;
  int P, Rb;             // padded pitch W+2, padded rows per patch
  int BK;                // channels per chunk: 32 (128-byte rows) or 16 (64-byte rows)
  int MT;                // 128-row sub-tiles per work item (1 or 2)
  int NBUF;              // accumulator buffers (1 or 2)
  int items_per_img;     // ceil(tiles_per_img / MT)
  int BN, n_tiles;
  int a_stages, b_stages;
  int tg;                // filter taps per weight-ring release (3 = one tcgen05.commit per filter row, 1 = per tap); b_stages % tg == 0
  int tmem_cols;
  int act;
  int sb_vec;            // scale / bias pointers are 16-byte aligned (they may be views into a flat parameter buffer)
  int round_out, res_up2;   // as in UmmaConvParams
  int dbg;               // tuning only (DGMR_PATCH_DBG): 1 = epilogue skips global traffic, 2 = issuer skips the MMAs, 4 = no TMA loads
  int64_t total_items;   // n_tiles * N * D * items_per_img (pair: n_tiles * qpairs * items_per_img)
  int64_t qpairs;        // pair mode: ceil(N*D / 2) image-depth slice pairs
  const float* bias; const float* scale; const float* res; float* y;
};

constexpr int kPatchThreads = 320;  // warp0 TMA, warp1 MMA, warps 2..9 epilogue
// PAIR: launched as clusters of 2 CTAs (one TPC); see the cta_group::2 helpers above.  The two CTAs of a pair work on the SAME tile
// position of two different images, so their activation patches have identical shared-memory geometry (one A descriptor serves both).
// TG3: the weight ring is released per filter row of three taps (p.tg == 3, p.b_stages % 3 == 0): with the nine taps unrolled the release points,
// ring positions within a row and the ring wrap test become static (one test per row instead of three per tap).
template <int BK, int MT, bool PAIR, bool TG3>
__global__ void __launch_bounds__(kPatchThreads, 1)
conv_umma_patch_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PatchConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)BK * 4u;
  const uint32_t patch_bytes = (uint32_t)p.Rb * p.P * row_bytes;
  const uint32_t patch_al = (patch_bytes + 1023u) & ~1023u;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int64_t wid = PAIR ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;      // worker (CTA or CTA pair) index
  const int64_t nworkers = PAIR ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  const uint32_t b_bytes = (uint32_t)(PAIR ? p.BN / 2 : p.BN) * row_bytes;      // a pair splits the weight tile: half the rows per CTA
  const uint32_t b_al = (b_bytes + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + p.a_stages * patch_al;
  const uint32_t bar_base = b_base + p.b_stages * b_al;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (p.a_stages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * p.a_stages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * p.a_stages + p.b_stages + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (2 * p.a_stages + 2 * p.b_stages + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (2 * p.a_stages + 2 * p.b_stages + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.a_stages + 2 * p.b_stages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  const uint32_t epi_base = (tmem_ptr_addr + 8u + 127u) & ~127u;     // 8 epilogue warps x 2 KB transpose staging

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    // pair: the leader's full barriers collect both CTAs' bytes (one arrival: its own expect_tx of the doubled byte count -- the peer's
    // bytes may land first, the phase cannot complete before that arrival); its acc_empty collects all 16 epilogue warps
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); if (PAIR) tmem_alloc2(tmem_ptr_addr, (uint32_t)p.tmem_cols); else tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // pair: the peer's barriers must exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int chunks = (p.Cin + BK - 1) / BK;    // last chunk may be a channel tail (TMA zero fill, fewer k-steps)
  const int a_per_item = p.kd * chunks;       // activation patches per work item
  // work item -> (n_tile, image n, depth d, first flattened output fs)
  auto decode = [&](int64_t item, int& nt, int& n, int& d, int& fs) {
    int64_t t = item;
    const int ii = (int)(t % p.items_per_img); t /= p.items_per_img;
    if (PAIR) {   // image-depth slices q = 2*qp + rank; an odd tail leaves the peer a slice past the end (TMA zero fill, rows discarded)
      const int64_t q = 2 * (t % p.qpairs) + rank; t /= p.qpairs;
      n = (int)(q / p.D); d = (int)(q - (int64_t)n * p.D);
    } else {
      d = (int)(t % p.D); t /= p.D;
      n = (int)(t % p.N); t /= p.N;
    }
    nt = (int)t;
    fs = p.P + 1 + 128 * MT * ii;
  };

  if (warp == 0) {
    {
      // ===== TMA producer (warp-uniform loop, elected lane issues): one activation patch per (kd, channel chunk), nine weight tiles per patch.  The patch of step s+1 is
      // issued right after the weight tiles of step s, so the (large) patch load overlaps a whole step of MMAs.
      struct Cur { int64_t item; int kdi, c; };
      auto valid = [&](const Cur& q) { return q.item < p.total_items; };
      auto advance = [&](Cur& q) {
        if (++q.c == chunks) { q.c = 0; if (++q.kdi == p.kd) { q.kdi = 0; q.item += nworkers; } }
      };
      int sa = 0, sb = 0, gb = 0, tb = 0; uint32_t pha = 0, phb = 0;
      auto issue_patch = [&](const Cur& q) {
        int nt, n, d, fs; decode(q.item, nt, n, d, fs);
        const int r_lo = (fs - p.P - 1) / p.P;            // first padded image row of the patch
        mbar_wait(a_empty(sa), pha ^ 1u);
#ifdef DGMR_TUNING
        if (p.dbg & 4) {            // tuning: no loads, the MMAs run on whatever is in shared memory
          if (rank == 0 && elect_one()) mbar_expect_tx(a_full(sa), 0);
        } else
#endif
        if (PAIR) {
          const uint32_t lead = mapa_rank(a_full(sa), 0);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(a_full(sa), 2u * patch_bytes);   // the leader arms for both CTAs' bytes; the peer only sends them
            tma_load_5d_2sm(a_base + sa * patch_al, &tmA, lead, q.c * BK, -1, r_lo - 1, d + q.kdi - p.kd / 2, n);
          }
        } else if (elect_one()) {
          mbar_expect_tx(a_full(sa), patch_bytes);
          tma_load_5d(a_base + sa * patch_al, &tmA, a_full(sa), q.c * BK, -1, r_lo - 1, d + q.kdi - p.kd / 2, n);
        }
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
      };
      // Order matters: the weight tiles of step s go out BEFORE the patch of step s+1.  The patch slot of step s+1 is the one step s-1 is still
      // reading, so waiting for it first (as this loop did until round 2) held back the weights of step s until step s-1 had completely
      // finished -- every step then started with the MMA warp waiting a full TMA latency for its first weight tile (ncu source view: 21 % of
      // that warp's samples on that one wait) although the weight ring holds a whole step.
      Cur ca{wid, 0, 0}, cb = ca;
      if (valid(ca)) { issue_patch(ca); advance(ca); }
      while (valid(cb)) {
        int nt, n, d, fs; decode(cb.item, nt, n, d, fs);
        for (int tap = 0; tap < 9; ++tap) {
          // weight tiles are released in groups of p.tg taps: a tcgen05.commit costs the tensor pipe ~780 cycles (measured,
          // dgmr_debug_umma_rate), more than the 8 MMAs of one tap at N <= 128, so one commit per tap would throttle the MMAs
          if (tb == 0) mbar_wait(b_empty(gb), phb ^ 1u);
#ifdef DGMR_TUNING
          if (p.dbg & 4) {
            if (rank == 0 && elect_one()) mbar_expect_tx(b_full(sb), 0);
          } else
#endif
          if (PAIR) {
            const uint32_t lead = mapa_rank(b_full(sb), 0);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(b_full(sb), 2u * b_bytes);
              tma_load_3d_2sm(b_base + sb * b_al, &tmB, lead, cb.c * BK, nt * p.BN + (int)rank * (p.BN / 2), cb.kdi * 9 + tap);
            }
          } else if (elect_one()) {
            mbar_expect_tx(b_full(sb), b_bytes);
            tma_load_3d(b_base + sb * b_al, &tmB, b_full(sb), cb.c * BK, nt * p.BN, cb.kdi * 9 + tap);
          }
          __syncwarp();
          if (++tb == p.tg) { tb = 0; ++gb; }
          if (++sb == p.b_stages) { sb = 0; gb = 0; phb ^= 1u; }
        }
        advance(cb);
        if (valid(ca)) { issue_patch(ca); advance(ca); }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer (pair: the leader alone, with M = 256 instructions spanning both CTAs); warp-uniform loop, elected lane issues.
      // This warp's instruction stream is what bounds the kernel at N <= 128 (an MMA lasts ~60 cycles there, an instruction issues every ~4.4):
      // descriptors are advanced as 32-bit low words (umma_tf32_lo), all per-tap bookkeeping that can be static is static (TG3), nothing is
      // divided or multiplied in the tap loop.
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
      constexpr uint32_t layout = (BK == 32) ? 2u : 4u;
      constexpr int ksteps = BK / 8;
      constexpr uint32_t rb16 = (uint32_t)BK * 4u / 16u;               // descriptor units (16 B) per patch row
      constexpr uint32_t dhi = desc_hi(8u * (uint32_t)BK * 4u, layout);
      const int tail_ks = (p.Cin % BK) ? (p.Cin % BK) / 8 : ksteps;
      const uint32_t a_lo0 = desc_lo(a_base), b_lo0 = desc_lo(b_base);
      const uint32_t a_st16 = patch_al >> 4, b_st16 = b_al >> 4;
      const int P16 = p.P * (int)rb16;                                 // descriptor units per padded image row
      auto mma = [](uint32_t dcol, uint32_t al, uint32_t bl, uint32_t id, uint32_t acc) {
        if (PAIR) umma_tf32_2cta_lo(dcol, al, bl, dhi, id, acc); else umma_tf32_lo(dcol, al, bl, dhi, id, acc);
      };
      auto commit = [](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
      int sa = 0, sb = 0, gb = 0, tb = 0; uint32_t pha = 0, phb = 0, it = 0;   // ring positions advance incrementally: no divisions in the issue loop
      for (int64_t item = wid; item < p.total_items; item += nworkers, ++it) {
        int nt, n, d, fs; decode(item, nt, n, d, fs);
        const int r_lo = (fs - p.P - 1) / p.P;
        const int buf = it % p.NBUF; const uint32_t phacc = (it / p.NBUF) & 1u;
        mbar_wait(acc_empty(buf), phacc ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * MT * p.BN);
        const int fb16 = (fs - r_lo * p.P) * (int)rb16;      // patch row of the centre tap's first pixel (in [P+1, 2P]), in descriptor units
        int chunk_i = 0;
        for (int a = 0; a < a_per_item; ++a) {
          const bool tail_now = (++chunk_i == chunks) && tail_ks != ksteps;   // channel-tail chunk: fewer valid k-steps
          if (chunk_i == chunks) chunk_i = 0;
          mbar_wait(a_full(sa), pha);
          const uint32_t a_lo = a_lo0 + (uint32_t)sa * a_st16 + (uint32_t)fb16;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int st = TG3 ? sb + tap % 3 : sb;                 // weight-ring stage of this tap (TG3: sb = first stage of the filter row)
            mbar_wait(b_full(st), phb);
            tc_fence_after();
            const int th = tap / 3, tw = tap - th * 3;
            // descriptors differ from the per-stage base only in their 14-bit start-address field: one 32-bit add each
            const uint32_t bl = b_lo0 + (uint32_t)st * b_st16;
            const uint32_t al = a_lo + (uint32_t)((th - 1) * P16 + (tw - 1) * (int)rb16);      // first patch row this tap reads (>= 0)
            const uint32_t acc0 = (a | tap) != 0 ? 1u : 0u;
            const bool grp_end = TG3 ? (tap % 3 == 2) : (tb + 1 == p.tg);
            if (elect_one()) {
#ifdef DGMR_TUNING
              if (p.dbg & 2) {
              } else
#endif
              if (!tail_now) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                  for (int k = 0; k < ksteps; ++k)
                    mma(tacc + (uint32_t)(mt * p.BN), al + (uint32_t)(mt * 128 * (int)rb16 + 2 * k), bl + (uint32_t)(2 * k), idesc, k == 0 ? acc0 : 1u);
                }
              } else {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                  for (int k = 0; k < ksteps; ++k)
                    if (k < tail_ks) mma(tacc + (uint32_t)(mt * p.BN), al + (uint32_t)(mt * 128 * (int)rb16 + 2 * k), bl + (uint32_t)(2 * k), idesc,
                                         k == 0 ? acc0 : 1u);
                }
              }
              if (grp_end) commit(b_empty(gb));
            }
            __syncwarp();
            if (TG3) {
              if (tap % 3 == 2) { sb += 3; ++gb; if (sb == p.b_stages) { sb = 0; gb = 0; phb ^= 1u; } }
            } else {
              if (++tb == p.tg) { tb = 0; ++gb; }
              if (++sb == p.b_stages) { sb = 0; gb = 0; phb ^= 1u; }
            }
          }
          if (elect_one()) commit(a_empty(sa));
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
        }
        if (elect_one()) commit(acc_full(buf));
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: 8 warps.  Warp group e = (warp-2)/4 takes sub-tile e (MT == 2) or column half e (MT == 1); inside a group
    // warp w may only touch TMEM lanes [32*(w%4), +32).  tcgen05.ld hands each lane one accumulator ROW; writing y (and reading
    // the residual) that way touches 32 different lines per instruction with half-used sectors, which the profile showed as 3x
    // L1->L2 write amplification on a kernel that is bound by the L2<->SM fabric.  So every 16-column chunk is transposed through
    // a 2 KB per-warp staging tile: afterwards 4 adjacent lanes own 64 contiguous bytes of one row (whole sectors both ways).
    const int q = warp & 3;
    const int eg = (warp - 2) >> 2;
    const uint32_t stg = epi_base + (uint32_t)(warp - 2) * 2048u;
    const int lr = lane >> 2, lc = lane & 3;
    const uint32_t st_row = stg + (uint32_t)lane * 64u;
    const uint32_t st_sw = (uint32_t)((lane >> 1) & 3);
    uint32_t it = 0;
    for (int64_t item = wid; item < p.total_items; item += nworkers, ++it) {
      int nt, n, d, fs; decode(item, nt, n, d, fs);
      const int buf = it % p.NBUF; const uint32_t phacc = (it / p.NBUF) & 1u;
      const int co0 = nt * p.BN;
      const int mt = (MT == 2) ? eg : 0;
      const int cbeg = (MT == 2) ? 0 : eg * ((p.BN / 2 + 15) / 16 * 16);
      const int cend = (MT == 2) ? p.BN : (eg == 0 ? (p.BN / 2 + 15) / 16 * 16 : p.BN);
      int64_t mrow[4], rrow[4]; bool vrow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = fs + 128 * mt + q * 32 + lr + 8 * j;
        const int hp = f / p.P, wp = f - hp * p.P;
        vrow[j] = (wp >= 1) && (wp <= p.W) && (hp >= 1) && (hp <= p.H) && (n < p.N)
#ifdef DGMR_TUNING
                  && !(p.dbg & 1)
#endif
            ;
        mrow[j] = ((((int64_t)n * p.D + d) * p.H + (hp - 1)) * p.W + (wp - 1)) * p.Cout + co0 + 4 * lc;
        rrow[j] = p.res_up2 ? ((((int64_t)n * p.D + d) * (p.H >> 1) + ((hp - 1) >> 1)) * (p.W >> 1) + ((wp - 1) >> 1)) * p.Cout + co0 + 4 * lc : mrow[j];
      }
      const int g = n / (p.N / p.G);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * MT + mt) * p.BN);
      const float* sc = p.scale ? p.scale + (int64_t)g * p.Cout : nullptr;
      float4 rr[4];
      auto load_res = [&](int c, float4* dst) {
        if (p.res == nullptr || c >= cend || co0 + c + 4 * lc >= p.Cout) return;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (vrow[j]) dst[j] = __ldg(reinterpret_cast<const float4*>(p.res + rrow[j] + c));
      };
      load_res(cbeg, rr);
      mbar_wait(acc_full(buf), phacc);
      tc_fence_after();
      for (int c = cbeg; c < cend; c += 16) {
        if (co0 + c >= p.Cout) break;
        float4 rn[4];
        load_res(c + 16, rn);                       // next chunk's residual is in flight while this chunk is transposed
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (((uint32_t)k ^ st_sw) << 4)), "f"(v[4 * k]), "f"(v[4 * k + 1]),
                       "f"(v[4 * k + 2]), "f"(v[4 * k + 3]) : "memory");
        __syncwarp();
        const int co = co0 + c + 4 * lc;
        const bool cok = co < p.Cout;
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cok) {
          if (p.sb_vec) {
            if (sc) s4 = __ldg(reinterpret_cast<const float4*>(sc + co));
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co));
          } else {
            if (sc) s4 = make_float4(__ldg(sc + co), __ldg(sc + co + 1), __ldg(sc + co + 2), __ldg(sc + co + 3));
            if (p.bias) b4 = make_float4(__ldg(p.bias + co), __ldg(p.bias + co + 1), __ldg(p.bias + co + 2), __ldg(p.bias + co + 3));
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = lr + 8 * j;
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                       : "r"(stg + (uint32_t)row * 64u + (((uint32_t)lc ^ (uint32_t)((row >> 1) & 3)) << 4)) : "memory");
          if (vrow[j] && cok) {
            o.x = fmaf(o.x, s4.x, b4.x); o.y = fmaf(o.y, s4.y, b4.y); o.z = fmaf(o.z, s4.z, b4.z); o.w = fmaf(o.w, s4.w, b4.w);
            if (p.res) { o.x += rr[j].x; o.y += rr[j].y; o.z += rr[j].z; o.w += rr[j].w; }
            if (p.act == DGMR_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (p.round_out) o = rna_tf32_e4(o);
            *reinterpret_cast<float4*>(p.y + mrow[j] + c) = o;
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = rn[j];
      }
      // this warp is done reading the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_rank(acc_empty(buf), 0));   // the leader's MMA thread waits for both CTAs' epilogues
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty(buf)) : "memory");
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { __syncwarp(); if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// ------------------------------------------------------------------ wgrad on tensor cores
//   dwp[tap][co][ci] += sum_{pixels p in my K-slice} dz[p][co] * x[p + tap][ci]
// GEMM view: M = Cout (128 per CTA), N = Cin tile, K = output pixels.  Both operands are read straight from the
// channels-last tensors as MN-major tiles (the channel axis is contiguous, the K axis = pixels is the row axis):
// one pipeline stage = 32 pixels x {128 co of dz, BN ci of x shifted by the tap}; the tap shift and the zero
// padding are TMA coordinates / out-of-bounds fill.  Split-K over CTAs, fp32 red.add into the zeroed dwp.
struct UmmaWgradParams {
  int N, D, H, W, Cin, Cout, kd, kh, kw;
  int bw, bh, bn;         // pixel box of one K block: bw*bh*bn == 32
  int tiles_w, tiles_h;   // W/bw, H/bh
  int aw;                 // channels per swizzle atom (32/16/8) -> atom row bytes aw*4
  int BN;                 // ci tile (multiple of 16 and of aw, <= 256)
  int ci_tiles;
  int stages, tmem_cols;
  int cg;                 // stages per release group (one tcgen05.commit hands cg stages back); stages % cg == 0
  int kb_total, kb_chunk;
  int subpix;             // 1: the 16 pre-summed sub-pixel tiles of conv_subpix.cu: "tap" z = ((i*2+j)*2+a)*2+b reads the phase-(i,j) view of dz
                          //    (tmDz / tmV1 / tmV2 / tmV3) and x shifted by (a+i-1, b+j-1); N, H, W are the LOW-resolution geometry
  float* dwp;
};

// X3: 3xTF32 operands, as in conv_umma_fwd_kernel (stage = [dz hi][x hi][dz lo][x lo], three MMAs per k-step).
template <bool X3>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDzLo,
                       const __grid_constant__ CUtensorMap tmXLo, const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ CUtensorMap tmV2,
                       const __grid_constant__ CUtensorMap tmV3, const UmmaWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KP = 32;                                   // pixels per stage
  const uint32_t blk_bytes = (uint32_t)KP * p.aw * 4u;     // one channel-block [32 pixels][aw channels]
  const int a_blocks = 128 / p.aw, b_blocks = p.BN / p.aw;
  const uint32_t a_bytes = a_blocks * blk_bytes, b_bytes = b_blocks * blk_bytes;
  const uint32_t half_bytes = a_bytes + b_bytes;           // multiples of 1024 (KP*aw*4 >= 1024)
  const uint32_t stage_bytes = X3 ? 2u * half_bytes : half_bytes;
  const uint32_t bar_base = base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.stages + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int tap = blockIdx.x;
  int tkw = tap % p.kw, tkh = (tap / p.kw) % p.kh, tkd = tap / (p.kw * p.kh);
  int sh_w = tkw - p.kw / 2, sh_h = tkh - p.kh / 2;
  const CUtensorMap* dzmap = &tmDz;
  if (p.subpix) {
    const int i = tap >> 3, j = (tap >> 2) & 1, a = (tap >> 1) & 1, b = tap & 1;
    sh_h = a + i - 1; sh_w = b + j - 1; tkd = 0;
    const int v = i * 2 + j;
    dzmap = (v == 0) ? &tmDz : (v == 1) ? &tmV1 : (v == 2) ? &tmV2 : &tmV3;
  }
  const int sh_d = p.subpix ? 0 : tkd - p.kd / 2;
  const int co0 = (blockIdx.z / p.ci_tiles) * 128, ci0 = (blockIdx.z % p.ci_tiles) * p.BN;
  const int kb0 = blockIdx.y * p.kb_chunk;
  const int kb1 = min(kb0 + p.kb_chunk, p.kb_total);
  const int num_kb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDz) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    if (num_kb > 0) {      // warp-uniform loop, elected lane issues (see elect_one); stages come back in groups of p.cg
      int s = 0, g = 0, sg = 0; uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        int t = kb;
        const int wi = t % p.tiles_w; t /= p.tiles_w;
        const int hi = t % p.tiles_h; t /= p.tiles_h;
        const int d0 = t % p.D; t /= p.D;
        const int n0 = t * p.bn, w0 = wi * p.bw, h0 = hi * p.bh;
        if (sg == 0) mbar_wait(empty_bar(g), ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), (X3 ? 2u : 1u) * (a_bytes + b_bytes));
          const uint32_t sa = base + s * stage_bytes;
          for (int j = 0; j < a_blocks; ++j) tma_load_5d(sa + j * blk_bytes, dzmap, full_bar(s), co0 + j * p.aw, w0, h0, d0, n0);
          for (int j = 0; j < b_blocks; ++j)
            tma_load_5d(sa + a_bytes + j * blk_bytes, &tmX, full_bar(s), ci0 + j * p.aw, w0 + sh_w, h0 + sh_h, d0 + sh_d, n0);
          if (X3) {
            for (int j = 0; j < a_blocks; ++j) tma_load_5d(sa + half_bytes + j * blk_bytes, &tmDzLo, full_bar(s), co0 + j * p.aw, w0, h0, d0, n0);
            for (int j = 0; j < b_blocks; ++j)
              tma_load_5d(sa + half_bytes + a_bytes + j * blk_bytes, &tmXLo, full_bar(s), ci0 + j * p.aw, w0 + sh_w, h0 + sh_h, d0 + sh_d, n0);
          }
        }
        __syncwarp();
        if (++sg == p.cg) { sg = 0; ++g; }
        if (++s == p.stages) { s = 0; g = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (num_kb > 0) {
      // instruction descriptor: f32 accum, tf32 x tf32, A and B MN-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      // 32-bit MN-major operands need the 32-byte-atom swizzle (UMMA LayoutType SWIZZLE_128B_BASE32B = 1, written by TMA with
      // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): swizzle atom = 4 K-rows x 128 B, so SBO (next K atom) = 512 B; LBO (next
      // 32-channel block along MN) = one [32 pixels][32 channels] block.  One MMA (K = 8 pixels) spans two K atoms.
      const uint32_t layout = 1u;
      const uint32_t sbo = 4u * 128u;
      const uint32_t lbo = blk_bytes;
      const uint32_t kstep = 8u * 128u;      // 8 pixels further along K
      auto mn_desc = [&](uint32_t saddr) {
        uint64_t d = 0;
        d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
        d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
        d |= (uint64_t)1u << 46;
        d |= (uint64_t)layout << 61;
        return d;
      };
      int s = 0, g = 0, sg = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = base + s * stage_bytes;
        const bool rel = (sg + 1 == p.cg) || (kb + 1 == num_kb);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KP / 8; ++k) {
            const uint64_t adesc = mn_desc(sa + k * kstep);
            const uint64_t bdesc = mn_desc(sa + a_bytes + k * kstep);
            if (X3) {
              umma_tf32(tmem_base, mn_desc(sa + half_bytes + k * kstep), bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_tf32(tmem_base, adesc, mn_desc(sa + half_bytes + a_bytes + k * kstep), idesc, 1u);
              umma_tf32(tmem_base, adesc, bdesc, idesc, 1u);
            } else {
              umma_tf32(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if (rel) umma_commit(empty_bar(g));
        }
        __syncwarp();
        if (++sg == p.cg) { sg = 0; ++g; }
        if (++s == p.stages) { s = 0; g = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(tmem_full_bar);
      __syncwarp();
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c = 0; c < p.BN; c += 16) {
      if (ci0 + c >= p.Cin) break;
      float v[16];
      tmem_ld16(trow + (uint32_t)c, v);
      if (co >= p.Cout) continue;
      float* dst = p.dwp + ((int64_t)tap * p.Cout + co) * p.Cin + ci0 + c;
#pragma unroll
      for (int j = 0; j < 16; j += 4)   // Cin % 4 == 0 (pick_aw): one 16-byte vector reduction per 4 channels
        if (ci0 + c + j < p.Cin)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// ------------------------------------------------------------------ wgrad, row variant (3 taps per CTA)
// The plain wgrad kernel handles one filter tap per CTA and therefore reads dz and x nine times.  Here a K block is a run
// of 32 output pixels of ONE image row; the x operand is loaded as a 34-pixel patch (1-pixel halo each side, TMA zero fill)
// and the three taps kw = 0,1,2 of a filter row are three MMAs whose B descriptors start 0, 1, 2 rows into that patch; the dz
// tile is shared by the three.  3 accumulators of BN columns in TMEM.  L2->SM traffic per tap drops ~3x.
struct WgradRowParams {
  int N, D, H, W, Cin, Cout, kd, kh;
  int BN, ci_tiles, b_blocks;
  int stages, tmem_cols;
  int kb_total, kb_chunk, wsegs;
  int rpk, hg, pitch;     // image rows per 32-pixel K block (1; 2 / 4 for W = 16 / 8), row groups per image (H / rpk), bytes between 32-channel patch blocks
  int subpix;             // 1: sub-pixel up-convolution tiles (conv_subpix.cu): blockIdx.x = (i, j, a); the two column taps b = 0, 1 share one dz tile
                          //    (phase-(i,j) view of the high-resolution dz: tmDz / tmV1 / tmV2 / tmV3) and one 34-pixel x patch of row h+a+i-1
  float* dwp;
};

__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_wgrad_row_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmV1,
                           const __grid_constant__ CUtensorMap tmV2, const __grid_constant__ CUtensorMap tmV3, const WgradRowParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = 4u * 4096u;                                   // 4 co blocks x [32 px][32 co]
  const uint32_t pitch = (uint32_t)p.pitch;
  const uint32_t b_span = (uint32_t)p.b_blocks * pitch;
  const uint32_t stage_bytes = a_bytes + ((b_span + 1023u) & ~1023u);
  const int wt = p.W < 32 ? p.W : 32;                                    // pixels of one image row inside a K block
  const uint32_t tx_bytes = a_bytes + (uint32_t)p.b_blocks * (uint32_t)(p.rpk * (wt + 2)) * 128u;
  const uint32_t bar_base = base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.stages + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int trow_i = blockIdx.x;                 // (kd, kh) filter row: fastest index, so the CTAs sharing a pixel range are co-scheduled (L2 reuse)
  int tkh = trow_i % p.kh, tkd = trow_i / p.kh;
  int sh_h = tkh - p.kh / 2, sh_d = tkd - p.kd / 2;
  int ndw = 3, dw0 = 0;                          // column taps per CTA and the patch row of the first one
  const CUtensorMap* dzmap = &tmDz;
  if (p.subpix) {
    const int i = trow_i >> 2, j = (trow_i >> 1) & 1, a = trow_i & 1;
    sh_h = a + i - 1; sh_d = 0; ndw = 2; dw0 = j;       // tap b reads x column w + b + j - 1 = patch row (b + j)
    const int v = i * 2 + j;
    dzmap = (v == 0) ? &tmDz : (v == 1) ? &tmV1 : (v == 2) ? &tmV2 : &tmV3;
  }
  const int co0 = (blockIdx.z / p.ci_tiles) * 128, ci0 = (blockIdx.z % p.ci_tiles) * p.BN;
  const int kb0 = blockIdx.y * p.kb_chunk;
  const int kb1 = min(kb0 + p.kb_chunk, p.kb_total);
  const int num_kb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDz) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    if (num_kb > 0) {      // warp-uniform loop, elected lane issues (see elect_one)
      int s = 0; uint32_t ph = 0;
      int ws = kb0 % p.wsegs, h, d, n;          // h counts row groups of p.rpk image rows
      { int t = kb0 / p.wsegs; h = t % p.hg; t /= p.hg; d = t % p.D; n = t / p.D; }
      for (int kb = kb0; kb < kb1; ++kb) {
        const int w0 = ws * 32;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), tx_bytes);
          const uint32_t sa = base + s * stage_bytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_5d(sa + j * 4096u, dzmap, full_bar(s), co0 + j * 32, w0, h * p.rpk, d, n);
          for (int j = 0; j < p.b_blocks; ++j)
            tma_load_5d(sa + a_bytes + j * pitch, &tmX, full_bar(s), ci0 + j * 32, w0 - 1, h * p.rpk + sh_h, d + sh_d, n);
        }
        __syncwarp();
        if (++ws == p.wsegs) { ws = 0; if (++h == p.hg) { h = 0; if (++d == p.D) { d = 0; ++n; } } }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (num_kb > 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      // MN-major descriptors as (low word, common high word): low = start address >> 4 | LBO >> 4 << 16 (LBO: 4096 between the 32-co blocks of
      // the dz tile, `pitch` between the 32-ci blocks of the x patch); high = SBO (512: next 4-row K atom) | version | SWIZZLE_128B_BASE32B.
      // One 32-bit add per descriptor in the issue loop (the issuing warp's instruction rate bounds the narrow layers).
      constexpr uint32_t dhi = ((512u >> 4) & 0x3FFFu) | (1u << 14) | (1u << 29);
      const uint32_t a_lbo = ((4096u >> 4) & 0x3FFFu) << 16, b_lbo = ((pitch >> 4) & 0x3FFFu) << 16;
      // first patch row of k-step k (8 pixels): the K block is p.rpk image rows of wt pixels, each stored with its two halo pixels
      uint32_t prow8[4];     // in descriptor units (a patch row is 128 B = 8 units)
#pragma unroll
      for (int k = 0; k < 4; ++k) prow8[k] = (uint32_t)(((8 * k) / wt) * (wt + 2) + (8 * k) % wt + dw0) * 8u;
      const uint32_t st16 = stage_bytes >> 4;
      const uint32_t a_lo0 = ((base >> 4) & 0x3FFFu) | a_lbo, b_lo0 = (((base + a_bytes) >> 4) & 0x3FFFu) | b_lbo;
      int s = 0; uint32_t ph = 0;
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            if (dw < ndw) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_lo(tmem_base + (uint32_t)(dw * p.BN), a_lo + (uint32_t)(k * 64), b_lo + prow8[k] + (uint32_t)(dw * 8), dhi, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(s));
        }
        __syncwarp();
        a_lo += st16; b_lo += st16;
        if (++s == p.stages) { s = 0; ph ^= 1u; a_lo = a_lo0; b_lo = b_lo0; }
      }
      if (elect_one()) umma_commit(tmem_full_bar);
      __syncwarp();
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    for (int dw = 0; dw < ndw; ++dw) {
      const int tap = p.subpix ? trow_i * 2 + dw : trow_i * 3 + dw;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dw * p.BN);
      for (int c = 0; c < p.BN; c += 16) {
        if (ci0 + c >= p.Cin) break;
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
        if (co >= p.Cout) continue;
        float* dst = p.dwp + ((int64_t)tap * p.Cout + co) * p.Cin + ci0 + c;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (ci0 + c + j < p.Cin)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// ------------------------------------------------------------------ descriptor probe (debug entry point)
// C[128][N] = A[r0 : r0+128][0:32] . B[N][0:32]^T with A a 256-row K-major SWIZZLE_128B tile loaded by ONE TMA: validates
// that a tcgen05 A descriptor may start at an arbitrary 128-byte row of a swizzled tile (needed to reuse one halo'd
// activation patch for all filter taps).  mode bit0: set the descriptor's base_offset field to (addr >> 7) & 7.
__global__ void __launch_bounds__(128, 1) umma_shift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                  float* C, int N, int r0, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = 256u * 128u, b_bytes = (uint32_t)N * 128u;
  const uint32_t bar = base + a_bytes + ((b_bytes + 1023u) & ~1023u);
  const uint32_t done_bar = bar + 8, tmem_ptr_addr = bar + 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_ptr_addr, 256); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, a_bytes + b_bytes);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(base), "l"(&tmA), "r"(bar), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(base + a_bytes), "l"(&tmB), "r"(bar), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = base + (uint32_t)r0 * 128u;
    for (int k = 0; k < 4; ++k) {
      uint64_t adesc = make_desc(sa, 1024u, 2u) + (uint64_t)(2 * k);
      if (mode & 1) adesc |= (uint64_t)((sa >> 7) & 7u) << 49;
      uint64_t bdesc = make_desc(base + a_bytes, 1024u, 2u) + (uint64_t)(2 * k);
      umma_tf32(tmem_base, adesc, bdesc, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(done_bar);
  }
  __syncthreads();
  mbar_wait(done_bar, 0);
  tc_fence_after();
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(trow + (uint32_t)c, v);
    for (int j = 0; j < 16; ++j) C[(size_t)(warp * 32 + lane) * N + c + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------ issue-rate probe (debug entry point)
// One thread per CTA issues `iters` kind::tf32 MMAs (M = 128, given N, K = 8) on zeroed shared memory and reports the cycles per
// MMA: mode 0 = back to back, mode 1 = a tcgen05.commit after every 8 MMAs staying 3 commits ahead (as the conv kernels do per
// filter tap), mode 2 = commit + wait every 8 MMAs (fully serialised).  Separates the tensor pipe's own rate from the pipeline around it.
__global__ void __launch_bounds__(128, 1) umma_rate_probe_kernel(float* out, int N, int iters, int mode, int shift_rows) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 32768u, bar = base + 65536u, tptr = bar + 64u;
  for (uint32_t i = threadIdx.x; i < 16384u; i += blockDim.x) reinterpret_cast<float*>(smem_raw + (base - raw))[i] = 0.f;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + 8u * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) { __syncwarp(); tmem_alloc(tptr, 256); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad = make_desc(a_base + (uint32_t)shift_rows * 128u, 1024u, 2u), bd = make_desc(b_base, 1024u, 2u);   // shift: A starts off the swizzle atom
    uint32_t ph[4] = {0, 0, 0, 0};
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_tf32(tmem_base, ad + (uint64_t)(2 * (k & 3)) + (uint64_t)((k >> 2) * 1024), bd + (uint64_t)(2 * (k & 3)), idesc, 1u);
      if (mode >= 1) {
        const int s = (i >> 3) & 3;
        umma_commit(bar + 8u * s);
        if (mode == 2) { mbar_wait(bar + 8u * s, ph[s]); ph[s] ^= 1u; }
        else if (i >= 24) { const int w = ((i >> 3) + 1) & 3; mbar_wait(bar + 8u * w, ph[w]); ph[w] ^= 1u; }   // stay 3 commits ahead
      }
    }
    if (mode == 0) { umma_commit(bar); mbar_wait(bar, 0); }
    else if (mode == 1) { for (int j = 1; j <= 3; ++j) { const int w = ((iters >> 3) + j) & 3; mbar_wait(bar + 8u * w, ph[w]); ph[w] ^= 1u; } }
    const long long t1 = clock64();
    out[blockIdx.x] = (float)(t1 - t0) / (float)iters;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { __syncwarp(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------ host side
// Tuning / test options (dgmr_set_option): plain process-wide ints read by the launchers -- no getenv on the launch path.
// -1 = heuristic default.
struct UmmaOptions {
  int umma_cg = -1;          // 1: one tcgen05.commit per pipeline stage (no release groups)
  int umma_persist = -1;     // 0: never the persistent plain kernel, 2: whenever eligible (small test shapes)
  int umma_persist_r = -1;   // CTAs per SM of the persistent plain kernel (1..3)
  int patch_pair = -1;       // 0: no CTA pairs in the halo-patch kernel
  int patch_mt = -1;         // 128-row sub-tiles per halo-patch work item (1 or 2)
  int patch_tg = -1;         // 1: release halo-patch weight tiles per tap instead of per filter row
  int prefer_patch = -1;     // 1: AUTO dispatch takes the halo-patch kernel whenever it supports the shape (parity tests on small shapes)
  int subpix_wgrad_row = -1; // sub-pixel weight gradient: 0 = always the tap-wise kernel, 1 = the row kernel whenever W % 32 == 0 (tests)
  int pairconv = -1;         // 0: never the pair-persistent whole-row kernel (conv_kwstack.cu, STACK = false), 1: whenever it supports the shape
  int kwstack = -1;          // 0: never the column-stacked kernel (conv_kwstack.cu), 1: whenever it supports the shape (small test shapes)
  int patch_dbg = 0;         // DGMR_TUNING builds only: make the halo-patch kernel skip work
};
static UmmaOptions g_opt;

// channels per K block: 32 (128-byte rows) whenever Cin >= 32 -- a channel tail (Cin % 32 in {8,16,24}) is a last block whose
// missing channels are TMA out-of-bounds zero fill and whose MMAs stop after the valid k-steps (64-byte-row TMA boxes move
// half the bytes per row at the same per-row cost); 16 / 8 only for genuinely narrow inputs
static int pick_bk(int Cin) { return (Cin % 8 != 0) ? 0 : (Cin >= 32) ? 32 : (Cin % 16 == 0) ? 16 : 8; }
static bool pick_box(int N, int H, int W, int* bw, int* bh, int* bn) {
  // bw*bh*bn == 128, bw | W, bh | H.  Prefer wide rows (contiguous TMA lines).  bn need not divide N (nor be <= N): images
  // past the end are TMA zero fill and their accumulator rows are discarded, so a 1x8x8 input still gets a (half-empty) tile.
  for (int w = 32; w >= 1; w >>= 1) {
    if (w > W || W % w) continue;
    int rest = 128 / w;
    for (int h = rest; h >= 1; h >>= 1) {
      if (h > H || H % h) continue;
      int n = rest / h;
      if (w * h * n != 128 || n > 256) continue;
      if (n > 1 && N % n != 0 && (int64_t)N * H * W >= 128 * 8) continue;   // big batches: keep looking for an exact fit first
      *bw = w; *bh = h; *bn = n;
      return true;
    }
  }
  for (int w = 32; w >= 1; w >>= 1) {
    if (w > W || W % w) continue;
    int rest = 128 / w;
    for (int h = rest; h >= 1; h >>= 1) {
      if (h > H || H % h) continue;
      int n = rest / h;
      if (w * h * n != 128 || n > 256) continue;
      *bw = w; *bh = h; *bn = n;
      return true;
    }
  }
  return false;
}
static bool umma_fwd_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G) {
  int bw, bh, bn;
  if (pick_bk(Cin) == 0) return false;
  if (Cout < 4 || Cout % 4) return false;   // N tile rounds up to 16; missing weight rows are TMA zero fill
  if (!((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3))) return false;
  if (!pick_box(N, H, W, &bw, &bh, &bn)) return false;
  if (G < 1 || N % G) return false;
  (void)D;
  return true;
}

int launch_conv_umma_fwd(const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int D, int H, int W, int Cin,
                         int Cout, int kd, int kh, int kw, int G, int act, cudaStream_t st, int accumulate = 0, const float* x_lo = nullptr,
                         const float* wp_lo = nullptr) {
  const bool x3 = x_lo != nullptr && wp_lo != nullptr;
  if ((int64_t)N * D * H * W * Cout >= ((int64_t)1 << 32) || (int64_t)N * D * H * W >= ((int64_t)1 << 31)) {
    set_error("conv_umma_fwd: output tensor too large for 32-bit element offsets"); return 1;
  }
  UmmaConvParams p;
  p.round_out = (act & DGMR_FLAG_ROUND_OUT) ? 1 : 0; p.res_up2 = (act & DGMR_FLAG_RES_UP2) ? 1 : 0;
  act &= 3;
  p.split_taps = accumulate; p.n_tiles = 1;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = kh; p.kw = kw; p.G = G;
  if (!pick_box(N, H, W, &p.bw, &p.bh, &p.bn)) { set_error("conv_umma_fwd: no 128-pixel box for N=%d H=%d W=%d", N, H, W); return 1; }
  p.tiles_w = W / p.bw; p.tiles_h = H / p.bh;
  p.BK = pick_bk(Cin);
  if (p.BK == 0) { set_error("conv_umma_fwd: Cin=%d not a multiple of 8", Cin); return 1; }
  int ntiles = (int)ceil_div(Cout, 256);
  p.BN = (int)(ceil_div(ceil_div(Cout, ntiles), 16) * 16);
  p.act = act; p.bias = bias; p.scale = scale; p.res = res; p.y = y;
  p.tmem_cols = 32; while (p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  const uint32_t a_bytes = 128u * p.BK * 4u, b_bytes = ((uint32_t)p.BN * p.BK * 4u + 1023u) & ~1023u;
  const uint32_t stage_bytes = (a_bytes + b_bytes) * (x3 ? 2u : 1u);
  // several CTAs per SM so one tile's epilogue / prologue overlaps another tile's main loop: aim at <= ~72 KB of
  // pipeline per CTA (3 resident CTAs) but never fewer than 3 stages; big tiles fall back to 1-2 CTAs per SM
  // ... unless the whole grid fits in one or two CTAs per SM anyway (ConvGRU steps, latent stack): those launches are pure
  // latency chains (K blocks / stages in flight x ~1.5 us per TMA round trip), so they get all the shared memory they can use
  const int64_t ctas_total = ceil_div(N, p.bn) * D * (H / p.bh) * (W / p.bw) * ntiles * (accumulate ? kd * kh * kw : 1);
  const int per_sm = (int)(ceil_div(ctas_total, (int64_t)sm_count()) >= 3 ? 3 : ceil_div(ctas_total, (int64_t)sm_count()));
  const uint32_t pipe_budget = per_sm >= 3 ? 72u * 1024u : per_sm == 2 ? 104u * 1024u : 200u * 1024u;
  int stages = (int)(pipe_budget / stage_bytes);
  if (stages < 3) stages = 3;
  if (stages > (per_sm >= 3 ? 6 : 12)) stages = per_sm >= 3 ? 6 : 12;
  if ((uint32_t)stages * stage_bytes > 200u * 1024u) stages = (int)((200u * 1024u) / stage_bytes);
  if (stages < 2) { set_error("conv_umma_fwd: stage too large"); return 1; }
  p.cg = 1;
  if (p.BN <= 160) {            // short MMAs: release stages in groups so that commits stay >= ~700 cycles of MMA work apart
    if (stages >= 6) { stages = stages / 3 * 3; p.cg = 3; }
    else if (stages >= 4) { stages = 4; p.cg = 2; }
  }
  if (g_opt.umma_cg == 1) p.cg = 1;
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 8 * (2 * stages + 2);
  const int taps = kd * kh * kw;
  CUtensorMap tmA, tmB, tmAlo, tmBlo;
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, (uint64_t)D * H * W * Cin * 4};
    uint32_t box[5] = {(uint32_t)p.BK, (uint32_t)p.bw, (uint32_t)p.bh, 1u, (uint32_t)p.bn};
    int e = make_tmap(&tmA, x, 5, dims, str, box, p.BK * 4);
    if (e) return e;
    if (x3) { e = make_tmap(&tmAlo, x_lo, 5, dims, str, box, p.BK * 4); if (e) return e; } else tmAlo = tmA;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)taps};
    uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)Cout * Cin * 4};
    uint32_t box[3] = {(uint32_t)p.BK, (uint32_t)p.BN, 1u};
    int e = make_tmap(&tmB, wp, 3, dims, str, box, p.BK * 4);
    if (e) return e;
    if (x3) { e = make_tmap(&tmBlo, wp_lo, 3, dims, str, box, p.BK * 4); if (e) return e; } else tmBlo = tmB;
  }
  static bool attr_set = false;
  if (!attr_set) {
    const int lim = 220 * 1024;
    bool ok = true;
#define DGMR_SET(K) ok = ok && cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) == cudaSuccess
    DGMR_SET((conv_umma_fwd_kernel<32, false, false>)); DGMR_SET((conv_umma_fwd_kernel<16, false, false>)); DGMR_SET((conv_umma_fwd_kernel<8, false, false>));
    DGMR_SET((conv_umma_fwd_kernel<32, false, true>)); DGMR_SET((conv_umma_fwd_kernel<16, false, true>)); DGMR_SET((conv_umma_fwd_kernel<8, false, true>));
    DGMR_SET((conv_umma_fwd_kernel<32, true, true>)); DGMR_SET((conv_umma_fwd_kernel<16, true, true>)); DGMR_SET((conv_umma_fwd_kernel<8, true, true>));
#undef DGMR_SET
    if (!ok) { set_error("conv_umma_fwd: cannot raise dynamic smem limit"); return 2; }
    attr_set = true;
  }
  int64_t mtiles = ceil_div(N, p.bn) * D * p.tiles_h * p.tiles_w;
  // many tiles: persistent CTAs (setup / pipeline fill once per CTA, double-buffered accumulators).  R CTAs per SM, bounded by the
  // 512 TMEM columns (2*BN each, power of two) -- the shared-memory request is padded so that no (R+1)-th CTA becomes resident
  // and blocks forever in tcgen05.alloc.
  // Measured (tests/time_conv1x1.py, time_conv3d.py): +15..35 % on 1x1 convs and other short K loops (<= 12 K blocks per tile, where
  // the per-tile fixed cost dominates), neutral to slightly negative on long K loops (fewer CTAs per SM in flight) -> only the former.
  const int num_kb_tile = taps * (int)ceil_div(Cin, p.BK);
  bool persist = !x3 && !accumulate && (Cout % 4 == 0) && 2 * p.BN <= 512 && mtiles * ntiles >= 4 * (int64_t)sm_count() && num_kb_tile <= 12;
  // ... and small launches of one to eight waves whatever their K loop (the per-step ConvGRU convolutions at 32x32 / 64x64): there the
  // per-CTA setup is a large part of a 20-40 us kernel.  Measured (tests/time_gru_conv.py, 16 images): 48->96 @64^2 38.7 -> 31.8 us,
  // 48->48 @64^2 29.5 -> 24.7 us, 96->192 @32^2 26.8 -> 24.8 us, never slower.
  if (!persist && !x3 && !accumulate && (Cout % 4 == 0) && 2 * p.BN <= 512 && mtiles * ntiles >= (int64_t)sm_count() &&
      mtiles * ntiles <= 8 * (int64_t)sm_count() && (int64_t)N * D * H * W <= 131072)
    persist = true;
  if (g_opt.umma_persist >= 0) {   // tuning / test option: 0 = never, 2 = whenever eligible (small test shapes)
    const int v = g_opt.umma_persist;
    if (v == 0) persist = false;
    else if (v == 2) persist = !x3 && !accumulate && (Cout % 4 == 0) && 2 * p.BN <= 512;
  }
  if (persist) {
    int cols = 32; while (cols < 2 * p.BN) cols <<= 1;
    int R = 512 / cols; if (R > 2) R = 2;
    if (g_opt.umma_persist_r >= 1 && g_opt.umma_persist_r <= 3 && g_opt.umma_persist_r <= 512 / cols) R = g_opt.umma_persist_r;
    const uint32_t budget = R == 3 ? 64u * 1024u : R == 2 ? 100u * 1024u : 196u * 1024u;
    int pst = (int)(budget / stage_bytes);
    if (pst < 2) { persist = false; }
    else {
      if (pst > 8) pst = 8;
      p.cg = 1;
      if (p.BN <= 160) { if (pst >= 6) { pst = pst / 3 * 3; p.cg = 3; } else if (pst >= 4) { pst = pst / 2 * 2; p.cg = 2; } }
      p.stages = pst; p.tmem_cols = cols; p.n_tiles = ntiles;
      size_t psmem = (size_t)pst * stage_bytes + 1024 + 8 * (2 * pst + 6) + 128 + 4 * 2048;
      const size_t floor_smem = (size_t)232448 / (R + 1) + 1024;
      if (psmem < floor_smem) psmem = floor_smem;
      static bool pattr = false;
      if (!pattr) {
        bool ok = true;
#define DGMR_SET(K) ok = ok && cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)) == cudaSuccess
        DGMR_SET((conv_umma_fwd_persist_kernel<32, false>)); DGMR_SET((conv_umma_fwd_persist_kernel<16, false>)); DGMR_SET((conv_umma_fwd_persist_kernel<8, false>));
        DGMR_SET((conv_umma_fwd_persist_kernel<32, true>)); DGMR_SET((conv_umma_fwd_persist_kernel<16, true>)); DGMR_SET((conv_umma_fwd_persist_kernel<8, true>));
#undef DGMR_SET
        if (!ok) { set_error("conv_umma_fwd: cannot raise dynamic smem limit"); return 2; }
        pattr = true;
      }
      int64_t g = (int64_t)sm_count() * R;
      if (g > mtiles * ntiles) g = mtiles * ntiles;
      const bool pef = p.round_out || p.res_up2;
#define DGMR_GO(BKV, EFV) conv_umma_fwd_persist_kernel<BKV, EFV><<<(unsigned)g, kUmmaThreads, psmem, st>>>(tmA, tmB, p)
      if (pef) { if (p.BK == 32) DGMR_GO(32, true); else if (p.BK == 16) DGMR_GO(16, true); else DGMR_GO(8, true); }
      else { if (p.BK == 32) DGMR_GO(32, false); else if (p.BK == 16) DGMR_GO(16, false); else DGMR_GO(8, false); }
#undef DGMR_GO
      DGMR_CHECK_LAUNCH("conv_umma_fwd_persist");
      return 0;
    }
  }
  dim3 grid((unsigned)mtiles, (unsigned)ntiles, (unsigned)(accumulate ? taps : 1));
  const bool ef = p.round_out || p.res_up2;
#define DGMR_GO(BKV, X3V, EFV) conv_umma_fwd_kernel<BKV, X3V, EFV><<<grid, kUmmaThreads, smem, st>>>(tmA, tmB, tmAlo, tmBlo, p)
  if (x3) {   // (parity mode does not optimise for time: one instantiation, epilogue flags compiled in; ROUND_OUT is refused there)
    if (p.BK == 32) DGMR_GO(32, true, true); else if (p.BK == 16) DGMR_GO(16, true, true); else DGMR_GO(8, true, true);
  } else if (ef) {
    if (p.BK == 32) DGMR_GO(32, false, true); else if (p.BK == 16) DGMR_GO(16, false, true); else DGMR_GO(8, false, true);
  } else {
    if (p.BK == 32) DGMR_GO(32, false, false); else if (p.BK == 16) DGMR_GO(16, false, false); else DGMR_GO(8, false, false);
  }
#undef DGMR_GO
  DGMR_CHECK_LAUNCH("conv_umma_fwd");
  return 0;
}


static bool pick_box32(int N, int H, int W, int* bw, int* bh, int* bn) {
  for (int w = 32; w >= 1; w >>= 1) {
    if (w > W || W % w) continue;
    int rest = 32 / w;
    for (int h = rest; h >= 1; h >>= 1) {
      if (h > H || H % h) continue;
      int n = rest / h;
      if (n > N || N % n) continue;
      *bw = w; *bh = h; *bn = n;
      return true;
    }
  }
  return false;
}
static int pick_aw(int Cin, int Cout) {
  // fp32 MN-major tiles exist only with 128-byte rows (32 channels); channel tails are TMA out-of-bounds zero fill
  return (Cin % 4 == 0 && Cout % 4 == 0) ? 32 : 0;
}
static bool umma_wgrad_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw) {
  int bw, bh, bn;
  (void)D;
  if (pick_aw(Cin, Cout) == 0) return false;
  if (Cin < 4 || Cout < 4) return false;
  if (!((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3))) return false;
  if (!pick_box32(N, H, W, &bw, &bh, &bn)) return false;
  if ((int64_t)N * D * H * W < 256) return false;   // tiny K: the SIMT kernel is as good
  return true;
}

int launch_conv_umma_wgrad(const float* x, const float* dz, float* dwp, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, cudaStream_t st,
                           const float* x_lo = nullptr, const float* dz_lo = nullptr) {
  const bool x3 = x_lo != nullptr && dz_lo != nullptr;
  UmmaWgradParams p;
  p.subpix = 0;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = kh; p.kw = kw;
  if (!pick_box32(N, H, W, &p.bw, &p.bh, &p.bn)) { set_error("conv_umma_wgrad: no 32-pixel box"); return 1; }
  p.tiles_w = W / p.bw; p.tiles_h = H / p.bh;
  p.aw = pick_aw(Cin, Cout);
  if (p.aw == 0) { set_error("conv_umma_wgrad: channels not multiples of 8"); return 1; }
  p.ci_tiles = (int)ceil_div(Cin, 256);
  int bn_ci = (int)ceil_div(Cin, p.ci_tiles);
  p.BN = (int)(ceil_div(bn_ci, p.aw) * p.aw);
  p.tmem_cols = 32; while (p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  const uint32_t blk_bytes = 32u * p.aw * 4u;
  const uint32_t stage_bytes = (128 / p.aw + p.BN / p.aw) * blk_bytes * (x3 ? 2u : 1u);
  // one CTA per SM (launch bounds): use the shared memory for stages, handed back in groups so that a tcgen05.commit (which costs
  // the pipe ~780 cycles) follows >= 8 MMAs (4 per stage, 128..512 cycles each group otherwise)
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("conv_umma_wgrad: stage too large"); return 1; }
  p.cg = 1;
  if (stages >= 6 && p.BN <= 128) { stages = stages / 3 * 3; p.cg = 3; }
  else if (stages >= 4) { stages = stages / 2 * 2; p.cg = 2; }
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 2);
  const int taps = kd * kh * kw;
  const int co_tiles = (int)ceil_div(Cout, 128);
  p.kb_total = (N / p.bn) * D * p.tiles_h * p.tiles_w;
  int64_t base_ctas = (int64_t)taps * co_tiles * p.ci_tiles;
  int64_t ksplit = ((int64_t)sm_count() * 3) / base_ctas;   // floor: never spill a few CTAs into an extra wave
  if (ksplit > p.kb_total / 4) ksplit = p.kb_total / 4;
  if (ksplit < 1) ksplit = 1;
  // Parity mode: tcgen05 keeps the accumulator in fp32 but TRUNCATES the sum after every MMA (measured: tests/test_fullsize_gpu.py::
  // test_tensor_core_accumulation), a bias of ~-0.5 ulp per MMA that grows with the chain -- 1e-4 relative after the ~2000-MMA chains of a
  // two-wave split, more than 3xTF32 is meant to leave.  Short chains (16 K blocks = 192 MMAs) flushed with round-to-nearest fp32 red.adds
  // keep it below 1e-5; the extra CTAs cost time, which parity mode does not optimise for.
  if (x3 && ksplit < ceil_div(p.kb_total, 16)) ksplit = ceil_div(p.kb_total, 16);
  if (ksplit > 65535) ksplit = 65535;
  p.kb_chunk = (int)ceil_div(p.kb_total, ksplit);
  ksplit = ceil_div(p.kb_total, p.kb_chunk);
  p.dwp = dwp;
  CUtensorMap tmDz, tmX, tmDzLo, tmXLo;
  {
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4, (uint64_t)D * H * W * Cout * 4};
    uint32_t box[5] = {(uint32_t)p.aw, (uint32_t)p.bw, (uint32_t)p.bh, 1u, (uint32_t)p.bn};
    int e = make_tmap(&tmDz, dz, 5, dims, str, box, p.aw * 4, true);
    if (e) return e;
    if (x3) { e = make_tmap(&tmDzLo, dz_lo, 5, dims, str, box, p.aw * 4, true); if (e) return e; } else tmDzLo = tmDz;
  }
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, (uint64_t)D * H * W * Cin * 4};
    uint32_t box[5] = {(uint32_t)p.aw, (uint32_t)p.bw, (uint32_t)p.bh, 1u, (uint32_t)p.bn};
    int e = make_tmap(&tmX, x, 5, dims, str, box, p.aw * 4, true);
    if (e) return e;
    if (x3) { e = make_tmap(&tmXLo, x_lo, 5, dims, str, box, p.aw * 4, true); if (e) return e; } else tmXLo = tmX;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_umma_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)) != cudaSuccess ||
        cudaFuncSetAttribute(conv_umma_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)) != cudaSuccess) {
      set_error("conv_umma_wgrad: cannot raise dynamic smem limit"); return 2;
    }
    attr_set = true;
  }
  if (cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)taps * Cout * Cin, st) != cudaSuccess) { set_error("conv_umma_wgrad: memset failed"); return 2; }
  dim3 grid((unsigned)taps, (unsigned)ksplit, (unsigned)(co_tiles * p.ci_tiles));
  if (x3) conv_umma_wgrad_kernel<true><<<grid, kUmmaThreads, smem, st>>>(tmDz, tmX, tmDzLo, tmXLo, tmDz, tmDz, tmDz, p);
  else conv_umma_wgrad_kernel<false><<<grid, kUmmaThreads, smem, st>>>(tmDz, tmX, tmDzLo, tmXLo, tmDz, tmDz, tmDz, p);
  DGMR_CHECK_LAUNCH("conv_umma_wgrad");
  return 0;
}

// Weight gradient of the 16 pre-summed sub-pixel tiles (conv_subpix.cu): dwp[z][co][ci] = sum_{n,h,w} dz[n, 2h+i, 2w+j, co] * x[n, h+a+i-1, w+b+j-1, ci],
// z = ((i*2+j)*2+a)*2+b.  x: [N,H,W,Cin] (low resolution), dz: [N,2H,2W,Cout]; each "tap" is one tap-wise wgrad over the low-resolution pixel
// grid whose dz operand is a strided phase view of the high-resolution tensor.  16 instead of 36 MACs per low-res pixel and channel pair.
int launch_conv_umma_wgrad_row_subpix(const float* x, const float* dz, float* dwp, int N, int H, int W, int Cin, int Cout, cudaStream_t st);
int launch_conv_umma_wgrad_subpix(const float* x, const float* dz, float* dwp, int N, int H, int W, int Cin, int Cout, cudaStream_t st) {
  const bool row_geom = W % 32 == 0 || (W == 16 && H % 2 == 0) || (W == 8 && H % 4 == 0);
  if (row_geom && Cin % 4 == 0 && Cout % 4 == 0 && (g_opt.subpix_wgrad_row == 1 || (g_opt.subpix_wgrad_row != 0 && (int64_t)N * H * W >= 16384)))
    return launch_conv_umma_wgrad_row_subpix(x, dz, dwp, N, H, W, Cin, Cout, st);   // (measured 2.4 ms tap-wise vs the row form on 96->96 at 64^2)
  UmmaWgradParams p;
  p.subpix = 1;
  p.N = N; p.D = 1; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = 1; p.kh = 1; p.kw = 1;
  if (!pick_box32(N, H, W, &p.bw, &p.bh, &p.bn)) { set_error("conv_umma_wgrad_subpix: no 32-pixel box"); return 1; }
  p.tiles_w = W / p.bw; p.tiles_h = H / p.bh;
  p.aw = pick_aw(Cin, Cout);
  if (p.aw == 0) { set_error("conv_umma_wgrad_subpix: channels not multiples of 4"); return 1; }
  p.ci_tiles = (int)ceil_div(Cin, 256);
  p.BN = (int)(ceil_div(ceil_div(Cin, p.ci_tiles), p.aw) * p.aw);
  p.tmem_cols = 32; while (p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  const uint32_t blk_bytes = 32u * p.aw * 4u;
  const uint32_t stage_bytes = (128 / p.aw + p.BN / p.aw) * blk_bytes;
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("conv_umma_wgrad_subpix: stage too large"); return 1; }
  p.cg = 1;
  if (stages >= 6 && p.BN <= 128) { stages = stages / 3 * 3; p.cg = 3; }
  else if (stages >= 4) { stages = stages / 2 * 2; p.cg = 2; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 2);
  const int taps = 16;
  const int co_tiles = (int)ceil_div(Cout, 128);
  p.kb_total = (N / p.bn) * p.tiles_h * p.tiles_w;
  const int64_t base_ctas = (int64_t)taps * co_tiles * p.ci_tiles;
  int64_t ksplit = ((int64_t)sm_count() * 3) / base_ctas;
  if (ksplit > p.kb_total / 4) ksplit = p.kb_total / 4;
  if (ksplit < 1) ksplit = 1;
  p.kb_chunk = (int)ceil_div(p.kb_total, ksplit);
  ksplit = ceil_div(p.kb_total, p.kb_chunk);
  p.dwp = dwp;
  CUtensorMap tmV[4], tmX;
  for (int v = 0; v < 4; ++v) {
    const int i = v >> 1, j = v & 1;
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, 1u, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)2 * Cout * 4, (uint64_t)2 * (2 * W) * Cout * 4, (uint64_t)(2 * H) * (2 * W) * Cout * 4, (uint64_t)(2 * H) * (2 * W) * Cout * 4};
    uint32_t box[5] = {(uint32_t)p.aw, (uint32_t)p.bw, (uint32_t)p.bh, 1u, (uint32_t)p.bn};
    int e = make_tmap(&tmV[v], dz + ((int64_t)i * 2 * W + j) * Cout, 5, dims, str, box, p.aw * 4, true);
    if (e) return e;
  }
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, 1u, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[5] = {(uint32_t)p.aw, (uint32_t)p.bw, (uint32_t)p.bh, 1u, (uint32_t)p.bn};
    int e = make_tmap(&tmX, x, 5, dims, str, box, p.aw * 4, true);
    if (e) return e;
  }
  if (cudaFuncSetAttribute(conv_umma_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)) != cudaSuccess) {
    set_error("conv_umma_wgrad_subpix: cannot raise dynamic smem limit"); return 2;
  }
  if (cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)taps * Cout * Cin, st) != cudaSuccess) { set_error("conv_umma_wgrad_subpix: memset failed"); return 2; }
  dim3 grid((unsigned)taps, (unsigned)ksplit, (unsigned)(co_tiles * p.ci_tiles));
  conv_umma_wgrad_kernel<false><<<grid, kUmmaThreads, smem, st>>>(tmV[0], tmX, tmV[0], tmX, tmV[1], tmV[2], tmV[3], p);
  DGMR_CHECK_LAUNCH("conv_umma_wgrad_subpix");
  return 0;
}


static bool umma_patch_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G) {
  if (kh != 3 || kw != 3 || !(kd == 1 || kd == 3)) return false;
  if (Cin % 8 != 0 || (Cin < 32 && Cin != 16) || Cout % 4 != 0 || Cout < 4) return false;
  if (W + 2 > 256 || (int64_t)H * W < 256) return false;    // tiny images: a tile would be mostly padding -> plain kernel
  if (G < 1 || N % G) return false;
  (void)D;
  return true;
}

// heuristic (AUTO only): persistent CTAs need a few tiles each to amortise their pipeline fill
// and narrow outputs (Cout < 64) make every MMA so short that the per-tap barrier round trips dominate (measured: 96->48 at
// 128^2 runs 289 TF/s on the plain kernel, 178 TF/s here).
static bool umma_patch_profitable(int N, int D, int H, int W, int Cin, int Cout) {
  if (g_opt.prefer_patch == 1) return true;
  // (a 16-channel tail -- Cin = 48 -- costs nothing special here: measured 48->96 at 128^2 432 TF/s against 201 on the plain kernel)
  if (Cin == 16 && Cout >= 48 && (int64_t)N * D * H * W >= (int64_t)128 * 16 * sm_count()) return true;   // depth-folded first temporal conv: 0.31 -> 0.25 ms
  return Cin >= 32 && Cout >= 64 && (int64_t)H * W >= 1024 && (int64_t)N * D * H * W >= (int64_t)128 * 2 * sm_count();   // (16 x 64^2 48->96: 34 -> 26 us)
}

int launch_conv_umma_patch(const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int D, int H, int W,
                           int Cin, int Cout, int kd, int G, int act, cudaStream_t st) {
  PatchConvParams p;
  p.round_out = (act & DGMR_FLAG_ROUND_OUT) ? 1 : 0; p.res_up2 = (act & DGMR_FLAG_RES_UP2) ? 1 : 0;
  act &= 3;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.G = G;
  p.P = W + 2;
  p.BK = (Cin >= 32) ? 32 : 16;
  const uint32_t row_bytes = (uint32_t)p.BK * 4u;
  p.n_tiles = (int)ceil_div(Cout, 256);
  {
    // small problems: split Cout over more CTAs so that the weight traffic (the same for every M tile) is spread over the SMs
    const int64_t m_items = (int64_t)N * D * ceil_div((int64_t)H * p.P - 2, 128);
    int want = (int)ceil_div((int64_t)sm_count(), m_items);
    int max_nt = Cout / 32 > 0 ? Cout / 32 : 1;
    if (want > max_nt) want = max_nt;
    if (want > p.n_tiles) p.n_tiles = want;
  }
  p.BN = (int)(ceil_div(ceil_div(Cout, p.n_tiles), 16) * 16);
  p.n_tiles = (int)ceil_div(Cout, p.BN);
  // CTA pairs (cta_group::2): each CTA stages half of every weight tile, so the weight ring is twice as deep in the same shared
  // memory and the weight bytes per SM halve -- the ring depth is what starves the MMAs at 96..192 channels
  int pair = (p.BK == 32 && (int64_t)N * D >= 2 && p.BN % 16 == 0 && sm_count() % 2 == 0) ? 1 : 0;
  if (g_opt.patch_pair == 0) pair = 0;
  // two sub-tiles per weight stage when accumulators and shared memory allow; double-buffer the accumulators when they still fit
  const uint32_t budget = 208u * 1024u;   // + 16 KB epilogue staging + barriers + alignment slack <= 226 KB
  const uint32_t b_al = (((uint32_t)(pair ? p.BN / 2 : p.BN) * row_bytes) + 1023u) & ~1023u;
  uint32_t patch_al = 0;
  p.a_stages = 2;
  bool fits = false;
  const int64_t tiles_total = (int64_t)p.n_tiles * N * D * ceil_div((int64_t)H * p.P - 2, 128);
  int mt_start = (2 * p.BN <= 512 && tiles_total >= 2 * (int64_t)sm_count()) ? 2 : 1;
  if (g_opt.patch_mt == 1 || (g_opt.patch_mt == 2 && 2 * p.BN <= 512)) mt_start = g_opt.patch_mt;
  for (p.MT = mt_start; p.MT >= 1; --p.MT) {
    const int span = 128 * p.MT + 2 * p.P + 2;
    p.Rb = (int)ceil_div(p.P - 1 + span, p.P);
    patch_al = (((uint32_t)p.Rb * p.P * row_bytes) + 1023u) & ~1023u;
    if (2 * patch_al + 3 * b_al <= budget) { fits = true; break; }
  }
  if (!fits) return -1;
  p.NBUF = (2 * p.MT * p.BN <= 512) ? 2 : 1;
  const int tiles_per_img = (int)ceil_div((int64_t)H * p.P - 2, 128);
  p.items_per_img = (int)ceil_div(tiles_per_img, p.MT);
  p.qpairs = ceil_div((int64_t)N * D, 2);
  p.total_items = (int64_t)p.n_tiles * (pair ? p.qpairs : (int64_t)N * D) * p.items_per_img;
  p.act = act; p.bias = bias; p.scale = scale; p.res = res; p.y = y;
  p.sb_vec = (((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(scale)) & 15u) == 0) ? 1 : 0;
  p.dbg = 0;
#ifdef DGMR_TUNING      // the knob makes the kernel skip work (wrong results): only in tuning builds
  p.dbg = g_opt.patch_dbg;
#endif
  if (((reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wp)) & 15u) != 0) {
    set_error("conv_umma_patch: x / wp / res / y must be 16-byte aligned"); return 1;
  }
  int need_cols = p.NBUF * p.MT * p.BN;
  p.tmem_cols = 32; while (p.tmem_cols < need_cols) p.tmem_cols <<= 1;
  p.b_stages = (int)((budget - 2 * patch_al) / b_al);
  if (p.b_stages > 9) p.b_stages = 9;
  p.tg = 1;
  if (p.b_stages >= 6) { p.tg = 3; p.b_stages = p.b_stages / 3 * 3; }   // one release (commit) per filter row of 3 taps, >= 2 rows in flight
  if (g_opt.patch_tg == 1) p.tg = 1;
  size_t smem = (size_t)p.a_stages * patch_al + (size_t)p.b_stages * b_al + 1024 + 8 * (2 * p.a_stages + 2 * p.b_stages + 6) + 128 + 8 * 2048;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, (uint64_t)D * H * W * Cin * 4};
    uint32_t box[5] = {(uint32_t)p.BK, (uint32_t)p.P, (uint32_t)p.Rb, 1u, 1u};
    int e = make_tmap(&tmA, x, 5, dims, str, box, (int)row_bytes);
    if (e) return e;
  }
  {
    const int taps = kd * 9;
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)taps};
    uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)Cout * Cin * 4};
    uint32_t box[3] = {(uint32_t)p.BK, (uint32_t)(pair ? p.BN / 2 : p.BN), 1u};
    int e = make_tmap(&tmB, wp, 3, dims, str, box, (int)row_bytes);
    if (e) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    const int lim = 226 * 1024;
    bool ok = true;
#define DGMR_SET(...) ok = ok && cudaFuncSetAttribute(conv_umma_patch_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) == cudaSuccess
    DGMR_SET(32, 1, false, false); DGMR_SET(32, 2, false, false); DGMR_SET(16, 1, false, false); DGMR_SET(16, 2, false, false);
    DGMR_SET(32, 1, true, false); DGMR_SET(32, 2, true, false);
    DGMR_SET(32, 1, false, true); DGMR_SET(32, 2, false, true); DGMR_SET(16, 1, false, true); DGMR_SET(16, 2, false, true);
    DGMR_SET(32, 1, true, true); DGMR_SET(32, 2, true, true);
#undef DGMR_SET
    if (!ok) {
      set_error("conv_umma_patch: cannot raise dynamic smem limit"); return 2;
    }
    attr_set = true;
  }
  if (pair) {
    int64_t grid = sm_count();                       // even (checked above): one CTA pair per TPC
    if (grid > 2 * p.total_items) grid = 2 * p.total_items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kPatchThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const bool tg3 = p.tg == 3;
    cudaError_t e = (p.MT == 2) ? (tg3 ? cudaLaunchKernelEx(&cfg, conv_umma_patch_kernel<32, 2, true, true>, tmA, tmB, p)
                                       : cudaLaunchKernelEx(&cfg, conv_umma_patch_kernel<32, 2, true, false>, tmA, tmB, p))
                                : (tg3 ? cudaLaunchKernelEx(&cfg, conv_umma_patch_kernel<32, 1, true, true>, tmA, tmB, p)
                                       : cudaLaunchKernelEx(&cfg, conv_umma_patch_kernel<32, 1, true, false>, tmA, tmB, p));
    if (e != cudaSuccess) { set_error("conv_umma_patch: cluster launch failed: %s", cudaGetErrorString(e)); return 2; }
    return 0;
  }
  int64_t grid = sm_count();
  if (grid > p.total_items) grid = p.total_items;
  const dim3 g((unsigned)grid);
#define DGMR_GO(BKV, MTV) do { if (p.tg == 3) conv_umma_patch_kernel<BKV, MTV, false, true><<<g, kPatchThreads, smem, st>>>(tmA, tmB, p); \
                               else conv_umma_patch_kernel<BKV, MTV, false, false><<<g, kPatchThreads, smem, st>>>(tmA, tmB, p); } while (0)
  if (p.BK == 32 && p.MT == 2) DGMR_GO(32, 2);
  else if (p.BK == 32) DGMR_GO(32, 1);
  else if (p.MT == 2) DGMR_GO(16, 2);
  else DGMR_GO(16, 1);
#undef DGMR_GO
  DGMR_CHECK_LAUNCH("conv_umma_patch");
  return 0;
}


static bool umma_wgrad_row_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw) {
  (void)N; (void)D;
  if (kw != 3 || !(kh == 1 || kh == 3) || !(kd == 1 || kd == 3)) return false;
  if (Cin % 4 != 0 || Cout % 4 != 0) return false;
  if (W % 32 == 0) return true;
  return (W == 16 && H % 2 == 0) || (W == 8 && H % 4 == 0);     // narrow images: a K block is 2 / 4 whole image rows
}

int launch_conv_umma_wgrad_row(const float* x, const float* dz, float* dwp, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, cudaStream_t st) {
  WgradRowParams p;
  p.subpix = 0;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = kh;
  p.ci_tiles = (int)ceil_div(Cin, 160);
  p.BN = (int)(ceil_div(ceil_div(Cin, p.ci_tiles), 32) * 32);
  p.b_blocks = p.BN / 32;
  p.tmem_cols = 32; while (p.tmem_cols < 3 * p.BN) p.tmem_cols <<= 1;
  const int wt = W < 32 ? W : 32;
  p.rpk = 32 / wt; p.hg = H / p.rpk;
  p.pitch = (int)((((uint32_t)(p.rpk * (wt + 2)) * 128u + 511u) / 512u) * 512u);     // a multiple of the 512-byte swizzle period
  const uint32_t stage_bytes = 4u * 4096u + (((uint32_t)p.b_blocks * (uint32_t)p.pitch + 1023u) & ~1023u);
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) { set_error("conv_umma_wgrad_row: stage too large"); return 1; }
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 2);
  const int taps = kd * kh * 3;
  const int co_tiles = (int)ceil_div(Cout, 128);
  p.wsegs = W < 32 ? 1 : W / 32;
  p.kb_total = N * D * p.hg * p.wsegs;
  int64_t base_ctas = (int64_t)kd * kh * co_tiles * p.ci_tiles;
  // one CTA per SM (shared memory): fill exactly two waves, never spill a few CTAs into a third
  int64_t ksplit = ((int64_t)sm_count() * 2) / base_ctas;
  if (ksplit > p.kb_total / 8) ksplit = p.kb_total / 8;
  if (ksplit < 1) ksplit = 1;
  p.kb_chunk = (int)ceil_div(p.kb_total, ksplit);
  ksplit = ceil_div(p.kb_total, p.kb_chunk);
  p.dwp = dwp;
  CUtensorMap tmDz, tmX;
  {
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4, (uint64_t)D * H * W * Cout * 4};
    uint32_t box[5] = {32u, (uint32_t)wt, (uint32_t)p.rpk, 1u, 1u};
    int e = make_tmap(&tmDz, dz, 5, dims, str, box, 128, true);
    if (e) return e;
  }
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, (uint64_t)D * H * W * Cin * 4};
    uint32_t box[5] = {32u, (uint32_t)(wt + 2), (uint32_t)p.rpk, 1u, 1u};
    int e = make_tmap(&tmX, x, 5, dims, str, box, 128, true);
    if (e) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_umma_wgrad_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)) != cudaSuccess) {
      set_error("conv_umma_wgrad_row: cannot raise dynamic smem limit"); return 2;
    }
    attr_set = true;
  }
  if (cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)taps * Cout * Cin, st) != cudaSuccess) { set_error("conv_umma_wgrad_row: memset failed"); return 2; }
  dim3 grid((unsigned)(kd * kh), (unsigned)ksplit, (unsigned)(co_tiles * p.ci_tiles));
  conv_umma_wgrad_row_kernel<<<grid, kUmmaThreads, smem, st>>>(tmDz, tmX, tmDz, tmDz, tmDz, p);
  DGMR_CHECK_LAUNCH("conv_umma_wgrad_row");
  return 0;
}

// Sub-pixel weight gradient, row variant (low-resolution W % 32 == 0): one CTA per (phase i, j; row tap a), its two column taps as two accumulators.
int launch_conv_umma_wgrad_row_subpix(const float* x, const float* dz, float* dwp, int N, int H, int W, int Cin, int Cout, cudaStream_t st) {
  WgradRowParams p;
  p.subpix = 1;
  p.N = N; p.D = 1; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = 1; p.kh = 1;
  p.ci_tiles = (int)ceil_div(Cin, 160);
  p.BN = (int)(ceil_div(ceil_div(Cin, p.ci_tiles), 32) * 32);
  p.b_blocks = p.BN / 32;
  p.tmem_cols = 32; while (p.tmem_cols < 3 * p.BN) p.tmem_cols <<= 1;
  const int wt = W < 32 ? W : 32;
  p.rpk = 32 / wt; p.hg = H / p.rpk;
  p.pitch = (int)((((uint32_t)(p.rpk * (wt + 2)) * 128u + 511u) / 512u) * 512u);
  const uint32_t stage_bytes = 4u * 4096u + (((uint32_t)p.b_blocks * (uint32_t)p.pitch + 1023u) & ~1023u);
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) { set_error("conv_umma_wgrad_row_subpix: stage too large"); return 1; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 2);
  const int co_tiles = (int)ceil_div(Cout, 128);
  p.wsegs = W < 32 ? 1 : W / 32;
  p.kb_total = N * p.hg * p.wsegs;
  const int64_t base_ctas = (int64_t)8 * co_tiles * p.ci_tiles;
  int64_t ksplit = ((int64_t)sm_count() * 2) / base_ctas;
  if (ksplit > p.kb_total / 8) ksplit = p.kb_total / 8;
  if (ksplit < 1) ksplit = 1;
  p.kb_chunk = (int)ceil_div(p.kb_total, ksplit);
  ksplit = ceil_div(p.kb_total, p.kb_chunk);
  p.dwp = dwp;
  CUtensorMap tmV[4], tmX;
  for (int v = 0; v < 4; ++v) {
    const int i = v >> 1, j = v & 1;
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, 1u, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)2 * Cout * 4, (uint64_t)2 * (2 * W) * Cout * 4, (uint64_t)(2 * H) * (2 * W) * Cout * 4, (uint64_t)(2 * H) * (2 * W) * Cout * 4};
    uint32_t box[5] = {32u, (uint32_t)wt, (uint32_t)p.rpk, 1u, 1u};
    int e = make_tmap(&tmV[v], dz + ((int64_t)i * 2 * W + j) * Cout, 5, dims, str, box, 128, true);
    if (e) return e;
  }
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, 1u, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[5] = {32u, (uint32_t)(wt + 2), (uint32_t)p.rpk, 1u, 1u};
    int e = make_tmap(&tmX, x, 5, dims, str, box, 128, true);
    if (e) return e;
  }
  if (cudaFuncSetAttribute(conv_umma_wgrad_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)) != cudaSuccess) {
    set_error("conv_umma_wgrad_row_subpix: cannot raise dynamic smem limit"); return 2;
  }
  if (cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)16 * Cout * Cin, st) != cudaSuccess) { set_error("conv_umma_wgrad_row_subpix: memset failed"); return 2; }
  dim3 grid(8u, (unsigned)ksplit, (unsigned)(co_tiles * p.ci_tiles));
  conv_umma_wgrad_row_kernel<<<grid, kUmmaThreads, smem, st>>>(tmV[0], tmX, tmV[1], tmV[2], tmV[3], p);
  DGMR_CHECK_LAUNCH("conv_umma_wgrad_row_subpix");
  return 0;
}

}  // namespace dgmr

using namespace dgmr;

extern "C" {

int dgmr_debug_umma_shift(const float* A /*[256][32]*/, const float* B /*[N][32]*/, float* C /*[128][N]*/, int N, int r0, int mode, dgmr_stream_t stream) {
  DGMR_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && r0 >= 0 && r0 <= 128, "dgmr_debug_umma_shift: bad args");
  CUtensorMap tmA, tmB;
  { uint64_t dims[2] = {32, 256}; uint64_t str[1] = {128}; uint32_t box[2] = {32, 256}; int e = make_tmap(&tmA, A, 2, dims, str, box, 128); if (e) return e; }
  { uint64_t dims[2] = {32, (uint64_t)N}; uint64_t str[1] = {128}; uint32_t box[2] = {32, (uint32_t)N}; int e = make_tmap(&tmB, B, 2, dims, str, box, 128); if (e) return e; }
  size_t smem = 256 * 128 + ((size_t)N * 128 + 1023) / 1024 * 1024 + 1024 + 64;
  if (cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { set_error("probe: smem attr"); return 2; }
  umma_shift_probe_kernel<<<1, 128, smem, S(stream)>>>(tmA, tmB, C, N, r0, mode);
  DGMR_CHECK_LAUNCH("umma_shift_probe");
  return 0;
}
int dgmr_debug_umma_rate(float* out /*[blocks]*/, int blocks, int N, int iters, int mode, int shift_rows, dgmr_stream_t stream) {
  DGMR_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && iters % 8 == 0 && iters >= 64 && blocks > 0, "dgmr_debug_umma_rate: bad args");
  const size_t smem = 65536 + 1024 + 128;
  if (cudaFuncSetAttribute(umma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { set_error("probe: smem attr"); return 2; }
  DGMR_REQUIRE(shift_rows >= 0 && shift_rows <= 64, "dgmr_debug_umma_rate: bad shift");
  umma_rate_probe_kernel<<<blocks, 128, smem, S(stream)>>>(out, N, iters, mode, shift_rows);
  DGMR_CHECK_LAUNCH("umma_rate_probe");
  return 0;
}
int dgmr_set_option(const char* name, int value) {
  if (name == nullptr) { set_error("dgmr_set_option: null name"); return 1; }
  struct { const char* n; int* v; } tab[] = {{"umma_cg", &g_opt.umma_cg}, {"umma_persist", &g_opt.umma_persist}, {"umma_persist_r", &g_opt.umma_persist_r},
                                             {"patch_pair", &g_opt.patch_pair}, {"patch_mt", &g_opt.patch_mt}, {"patch_tg", &g_opt.patch_tg},
                                             {"prefer_patch", &g_opt.prefer_patch}, {"patch_dbg", &g_opt.patch_dbg},
                                             {"subpix_wgrad_row", &g_opt.subpix_wgrad_row}, {"kwstack", &g_opt.kwstack}, {"kwstack_pair", &g_kwstack_pair}, {"pairconv", &g_opt.pairconv}, {"subpix_rows", &g_subpix_rows}};
  for (auto& t : tab)
    if (strcmp(t.n, name) == 0) { *t.v = value; return 0; }
  set_error("dgmr_set_option: unknown option '%s'", name);
  return 1;
}
int dgmr_conv_umma_supported(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw) {
  return umma_fwd_ok(N, D, H, W, Cin, Cout, kd, kh, kw, 1) ? 1 : 0;
}
int dgmr_wgrad_umma_supported(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw) {
  return umma_wgrad_ok(N, D, H, W, Cin, Cout, kd, kh, kw) ? 1 : 0;
}

int dgmr_conv_fwd(const float* x, const float* x_lo, const float* wp, const float* wp_lo, const float* bias, const float* scale, const float* res, float* y,
                  int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G, int act, int algo, int precision, dgmr_stream_t stream) {
  DGMR_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "dgmr_conv_fwd: bad dims");
  DGMR_REQUIRE((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3), "dgmr_conv_fwd: kernel extent must be 1 or 3");
  DGMR_REQUIRE(G >= 1 && N % G == 0, "dgmr_conv_fwd: N=%d not divisible by G=%d", N, G);
  const int accumulate = (act & DGMR_FLAG_ACCUMULATE) ? 1 : 0;
  act &= ~DGMR_FLAG_ACCUMULATE;
  const int eflags = act & (DGMR_FLAG_ROUND_OUT | DGMR_FLAG_RES_UP2);     // epilogue flags travel on with `act`
  DGMR_REQUIRE((act & ~eflags) == DGMR_ACT_NONE || (act & ~eflags) == DGMR_ACT_RELU, "dgmr_conv_fwd: bad act");
  DGMR_REQUIRE(!(eflags && accumulate), "dgmr_conv_fwd: ACCUMULATE excludes ROUND_OUT / RES_UP2");
  DGMR_REQUIRE(!(eflags & DGMR_FLAG_RES_UP2) || (res != nullptr && H % 2 == 0 && W % 2 == 0), "dgmr_conv_fwd: RES_UP2 needs res and even H, W");
  DGMR_REQUIRE(precision == DGMR_PREC_TF32 || precision == DGMR_PREC_3XTF32, "dgmr_conv_fwd: bad precision");
  bool ok = umma_fwd_ok(N, D, H, W, Cin, Cout, kd, kh, kw, G);
  // 3xTF32 ("parity mode"): the caller hands over the hi and lo parts of both operands (dgmr_split_tf32); served by the plain tcgen05
  // kernel only (three MMAs per k-step).  Without lo parts the call is an ordinary 1xTF32 / fp32-SIMT call on `x`, `wp`.
  const bool x3 = precision == DGMR_PREC_3XTF32 && x_lo != nullptr && wp_lo != nullptr;
  if (precision == DGMR_PREC_3XTF32 && !x3)
    DGMR_REQUIRE(x_lo == nullptr && wp_lo == nullptr, "dgmr_conv_fwd: 3xTF32 needs both x_lo and wp_lo (or neither: full-precision operands for the SIMT kernel)");
  if (x3) {
    DGMR_REQUIRE(ok && algo != DGMR_ALGO_SIMT, "dgmr_conv_fwd: 3xTF32 operand pairs were passed but the shape is not served by the tcgen05 path");
    if (accumulate) DGMR_REQUIRE(bias == nullptr && res == nullptr && act == DGMR_ACT_NONE, "dgmr_conv_fwd: ACCUMULATE excludes bias/res/act");
    DGMR_REQUIRE(!(eflags & DGMR_FLAG_ROUND_OUT), "dgmr_conv_fwd: ROUND_OUT makes no sense in 3xTF32 mode");
    return launch_conv_umma_fwd(x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, kh, kw, G, act, S(stream), accumulate, x_lo, wp_lo);
  }
  if (precision == DGMR_PREC_3XTF32) algo = DGMR_ALGO_SIMT;   // full-precision operands: fp32 FMA kernel
  if (accumulate) {
    DGMR_REQUIRE(bias == nullptr && res == nullptr && act == DGMR_ACT_NONE, "dgmr_conv_fwd: ACCUMULATE excludes bias/res/act");
    DGMR_REQUIRE(ok && algo != DGMR_ALGO_SIMT, "dgmr_conv_fwd: ACCUMULATE is a tensor-core-path mode");
    return launch_conv_umma_fwd(x, wp, nullptr, scale, nullptr, y, N, D, H, W, Cin, Cout, kd, kh, kw, G, DGMR_ACT_NONE, S(stream), 1);
  }
  if (algo == DGMR_ALGO_UMMA || algo == DGMR_ALGO_UMMA_PATCH) DGMR_REQUIRE(ok, "dgmr_conv_fwd: shape not supported by the tcgen05 path");
  if (algo == DGMR_ALGO_UMMA_KWSTACK) DGMR_REQUIRE(umma_kwstack_ok(N, D, H, W, Cin, Cout, kd, kh, kw, G), "dgmr_conv_fwd: shape not supported by the column-stacked kernel");
  // narrow outputs (Cout < 64) of at least a few waves of tiles: column taps stacked along N, the shift done in the epilogue (conv_kwstack.cu).
  // Measured (profiles/time_kwstack_r02.txt): 3-D 48->48 1.76 -> 0.90 ms, 96->48 @128^2 1.37 -> 0.90 ms; 16-channel inputs are slower there (0.32 -> 0.40 ms).
  if (algo == DGMR_ALGO_UMMA_KWSTACK || (algo == DGMR_ALGO_AUTO && g_opt.kwstack != 0 && Cout < 64 && Cin >= 32 && umma_kwstack_ok(N, D, H, W, Cin, Cout, kd, kh, kw, G) &&
                                         ((int64_t)N * D * H * W >= (int64_t)128 * 2 * sm_count() || g_opt.kwstack == 1)))
    return launch_conv_umma_kwstack(x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, G, act, S(stream));
  if (algo == DGMR_ALGO_UMMA_PAIR) {
    DGMR_REQUIRE(umma_pairconv_ok(N, D, H, W, Cin, Cout, kd, kh, kw, G), "dgmr_conv_fwd: shape not supported by the pair-persistent kernel");
    return launch_conv_umma_pairconv(x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, G, act, S(stream));
  }
  // wide channels on small images (the 16x16 / 32x32 sampler layers, the per-step ConvGRU convolutions): whole-row tiles, CTA pairs sharing each
  // weight tile.  Measured (tests/time_patch16.py, time_pairconv.py): 768->768 @16^2 559 -> 834 TF/s, 192->192 @32^2 632 (patch) -> 679, ConvGRU
  // 16 x 16^2 192->384 39 -> 31 us; the halo-patch kernel keeps N = 96 layers and everything at 64^2 and above (activation traffic dominates there).
  if (algo == DGMR_ALGO_AUTO && g_opt.pairconv != 0 && W <= 32 && Cout >= 64 && umma_pairconv_ok(N, D, H, W, Cin, Cout, kd, kh, kw, G) &&
      (Cin >= 192 || (int64_t)N * D * H * W <= 131072 || g_opt.pairconv == 1))
    return launch_conv_umma_pairconv(x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, G, act, S(stream));
  if (algo == DGMR_ALGO_UMMA || algo == DGMR_ALGO_UMMA_PATCH || (algo == DGMR_ALGO_AUTO && ok)) {
    if (algo != DGMR_ALGO_UMMA && umma_patch_ok(N, D, H, W, Cin, Cout, kd, kh, kw, G) &&
        (algo == DGMR_ALGO_UMMA_PATCH || umma_patch_profitable(N, D, H, W, Cin, Cout))) {
      int e = launch_conv_umma_patch(x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, G, act, S(stream));
      if (e >= 0) return e;   // -1: configuration does not fit in shared memory -> plain kernel
    }
    DGMR_REQUIRE(algo != DGMR_ALGO_UMMA_PATCH, "dgmr_conv_fwd: shape not supported by the halo-patch kernel");
    return launch_conv_umma_fwd(x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, kh, kw, G, act, S(stream));
  }
  return launch_conv_simt_fwd(x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, kh, kw, G, act, S(stream));
}

int dgmr_conv_wgrad(const float* x, const float* x_lo, const float* dz, const float* dz_lo, float* dwp, int N, int D, int H, int W, int Cin, int Cout,
                    int kd, int kh, int kw, int algo, int precision, dgmr_stream_t stream) {
  DGMR_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "dgmr_conv_wgrad: bad dims");
  DGMR_REQUIRE(precision == DGMR_PREC_TF32 || precision == DGMR_PREC_3XTF32, "dgmr_conv_wgrad: bad precision");
  bool ok = umma_wgrad_ok(N, D, H, W, Cin, Cout, kd, kh, kw) && (reinterpret_cast<uintptr_t>(dwp) & 15u) == 0;   // 16-byte vector reductions into dwp
  const bool x3 = precision == DGMR_PREC_3XTF32 && x_lo != nullptr && dz_lo != nullptr;
  if (precision == DGMR_PREC_3XTF32 && !x3)
    DGMR_REQUIRE(x_lo == nullptr && dz_lo == nullptr, "dgmr_conv_wgrad: 3xTF32 needs both x_lo and dz_lo (or neither: full-precision operands for the SIMT kernel)");
  if (x3) {
    DGMR_REQUIRE(ok && algo != DGMR_ALGO_SIMT, "dgmr_conv_wgrad: 3xTF32 operand pairs were passed but the shape is not served by the tcgen05 path");
    return launch_conv_umma_wgrad(x, dz, dwp, N, D, H, W, Cin, Cout, kd, kh, kw, S(stream), x_lo, dz_lo);
  }
  if (precision == DGMR_PREC_3XTF32) algo = DGMR_ALGO_SIMT;
  if (algo == DGMR_ALGO_UMMA || algo == DGMR_ALGO_UMMA_PATCH) DGMR_REQUIRE(ok, "dgmr_conv_wgrad: shape not supported by the tcgen05 path (or dwp not 16-byte aligned)");
  if (algo == DGMR_ALGO_UMMA_PATCH) DGMR_REQUIRE(umma_wgrad_row_ok(N, D, H, W, Cin, Cout, kd, kh, kw), "dgmr_conv_wgrad: shape not supported by the row kernel");
  if (algo == DGMR_ALGO_UMMA_PATCH || (algo == DGMR_ALGO_AUTO && ok && umma_wgrad_row_ok(N, D, H, W, Cin, Cout, kd, kh, kw) && (int64_t)N * D * H * W >= 8192))
    return launch_conv_umma_wgrad_row(x, dz, dwp, N, D, H, W, Cin, Cout, kd, kh, S(stream));
  if (algo == DGMR_ALGO_UMMA || (algo == DGMR_ALGO_AUTO && ok))
    return launch_conv_umma_wgrad(x, dz, dwp, N, D, H, W, Cin, Cout, kd, kh, kw, S(stream));
  return launch_conv_simt_wgrad(x, dz, dwp, N, D, H, W, Cin, Cout, kd, kh, kw, S(stream));
}

}  // extern "C"
