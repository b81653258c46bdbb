// Sub-pixel form of "nearest x2 upsample -> 3x3 convolution" (UpsampleGBlock.first_conv_3x3, ref: dgmr/common.py:146-149) on tcgen05.
//
//   y[n, 2h+i, 2w+j, co] = sum_{a,b in {0,1}} sum_ci  Wp[i][j][a][b][co][ci] * xl[n, h + a + i - 1, w + b + j - 1, ci]
//   Wp[i][j][a][b] = sum_{kh in S(i,a)} sum_{kw in S(j,b)} W[:, :, kh, kw],   S(0,0) = {0}, S(0,1) = {1,2}, S(1,0) = {0,1}, S(1,1) = {2}
//
// i.e. each of the four output phases (i, j) is a 2x2-tap convolution of the LOW-resolution input with pre-summed taps: 16 MACs per
// low-resolution pixel and channel pair instead of 36 (195.7 -> 87.0 GF per sample on the four up_g*.first_conv_3x3,
// the largest single item of the generator), and the upsampled activation is never written or read.  The pre-summing happens in fp32
// before the tf32 rounding of the packed weights, so the result differs from the 3x3 form only by that rounding (1.5e-5 in fp32).
//
// All (phase, tap) input shifts lie in the 3x3 neighbourhood of the low-resolution pixel, so the halo-patch scheme of conv_umma.cu applies
// unchanged: the image is addressed in the padded, flattened coordinate f = (h+1)*P + (w+1), P = W+2; one TMA box of whole padded rows per
// 32-channel chunk feeds every tap through a shifted tcgen05 A descriptor.
//   MODE 1 (forward):  work item = (128 flattened low-res positions, output row phase i); two accumulators (column phase j = 0, 1) of BN
//                      columns each; 8 weight tiles per chunk; the epilogue writes pixel (2h+i, 2w+j) of the high-resolution output.
//   MODE 2 (dgrad):    dxl[h', w'] = sum_{i,j,a,b} WpT[i][j][a][b] . dy[2(h' - (a+i-1)) + i, 2(w' - (b+j-1)) + j]: the K loop runs over the four
//                      phase sub-images of dy (four strided TMA views of the same tensor) x channel chunks, 4 taps each.
// The weight gradient of the 16 pre-summed tiles is the tap-wise tcgen05 wgrad kernel of conv_umma.cu with per-"tap" (shift, phase view)
// (launch_conv_umma_wgrad_subpix there); dgmr_unpack_wgrad_subpix folds them back onto the 3x3 taps.
#include "umma_common.cuh"

namespace dgmr {

int launch_conv_umma_wgrad_subpix(const float* x, const float* dz, float* dwp, int N, int H, int W, int Cin, int Cout, cudaStream_t st);
bool umma_pairconv_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G);   // conv_kwstack.cu
int launch_conv_umma_pairconv_subpix(const float* x, const float* wsp, const float* bias, const float* scale, const float* res, float* y, int N, int H, int W,
                                     int Cin, int Cout, int G, int act, cudaStream_t st);
int g_subpix_rows = -1;      // dgmr_set_option("subpix_rows"): 0 = never the whole-row pair kernel for the sub-pixel forward, 1 = whenever it supports the shape

struct SubpixParams {
  int N, H, W;           // LOW-resolution geometry (the grid the patches live on); the other tensor is 2H x 2W
  int Cin, Cout, G;      // K channels (per phase view in dgrad mode) / N channels of this GEMM
  int P, Rb;
  int MT;                // MODE 2: 128-row sub-tiles per item.  (MODE 1: one 128-row tile, two accumulators.)
  int NBUF, items_per_img, BN, n_tiles, a_stages, b_stages, tg, tmem_cols, act, sb_vec;
  int64_t total_items, qpairs;
  const float* bias; const float* scale; const float* res; float* y;
};

constexpr int kSubThreads = 320;  // warp0 TMA, warp1 MMA, warps 2..9 epilogue

// TG: weight tiles per ring release (p.tg; p.b_stages % TG == 0): with the NT tiles of a patch unrolled, release points and the ring wrap test are static.
template <int MODE, int MT, bool PAIR, int TG>
__global__ void __launch_bounds__(kSubThreads, 1)
conv_subpix_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB, const SubpixParams p) {
  constexpr int BK = 32;
  constexpr int NT = (MODE == 1) ? 8 : 4;                 // weight tiles (= MMA groups) per activation patch
  constexpr int NSLOT = (MODE == 1) ? 2 : MT;             // accumulators per work item
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t row_bytes = BK * 4u;
  const uint32_t patch_bytes = (uint32_t)p.Rb * p.P * row_bytes;
  const uint32_t patch_al = (patch_bytes + 1023u) & ~1023u;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int64_t wid = PAIR ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t nworkers = PAIR ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  const uint32_t b_bytes = (uint32_t)(PAIR ? p.BN / 2 : p.BN) * row_bytes;
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
  const uint32_t epi_base = (tmem_ptr_addr + 8u + 127u) & ~127u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (MODE == 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA3) : "memory");
    }
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); if (PAIR) tmem_alloc2(tmem_ptr_addr, (uint32_t)p.tmem_cols); else tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int chunks = p.Cin / BK;                                  // Cin % 32 == 0 (host check)
  const int ppi = (MODE == 1) ? chunks : 4 * chunks;              // activation patches per work item
  // work item -> (n tile, image n, first flattened position fs, output row phase ph [MODE 1])
  auto decode = [&](int64_t item, int& nt, int& n, int& fs, int& ph) {
    int64_t t = item;
    const int ii = (int)(t % p.items_per_img); t /= p.items_per_img;
    if (PAIR) { n = (int)(2 * (t % p.qpairs) + rank); t /= p.qpairs; }   // an odd tail gives the peer an image past the end (zero fill, rows discarded)
    else { n = (int)(t % p.N); t /= p.N; }
    nt = (int)t;
    if (MODE == 1) { ph = ii & 1; fs = p.P + 1 + 128 * (ii >> 1); }
    else { ph = 0; fs = p.P + 1 + 128 * MT * ii; }
  };

  if (warp == 0) {
    // ===== TMA producer: one activation patch per (phase view,) channel chunk, NT weight tiles per patch; the patch of the NEXT step is issued
    // right after the weight tiles of the current one, so the large load overlaps a whole step of MMAs
    struct Cur { int64_t item; int pi; };
    auto valid = [&](const Cur& q) { return q.item < p.total_items; };
    auto advance = [&](Cur& q) { if (++q.pi == ppi) { q.pi = 0; q.item += nworkers; } };
    int sa = 0, sb = 0, gb = 0, tb = 0; uint32_t pha = 0, phb = 0;
    auto issue_patch = [&](const Cur& q) {
      int nt, n, fs, ph; decode(q.item, nt, n, fs, ph);
      const int r_lo = (fs - p.P - 1) / p.P;
      const int view = (MODE == 2) ? q.pi / chunks : 0;
      const int c = (MODE == 2) ? q.pi - view * chunks : q.pi;
      const CUtensorMap* tm = (view == 0) ? &tmA0 : (view == 1) ? &tmA1 : (view == 2) ? &tmA2 : &tmA3;
      mbar_wait(a_empty(sa), pha ^ 1u);
      if (PAIR) {
        const uint32_t lead = mapa_rank(a_full(sa), 0);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(a_full(sa), 2u * patch_bytes);
          tma_load_5d_2sm(a_base + sa * patch_al, tm, lead, c * BK, -1, r_lo - 1, 0, n);
        }
      } else if (elect_one()) {
        mbar_expect_tx(a_full(sa), patch_bytes);
        tma_load_5d(a_base + sa * patch_al, tm, a_full(sa), c * BK, -1, r_lo - 1, 0, n);
      }
      __syncwarp();
      if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
    };
    // (weight tiles of step s first, THEN the patch of step s+1 whose slot step s-1 is still reading: see conv_umma_patch_kernel)
    Cur ca{wid, 0}, cb = ca;
    if (valid(ca)) { issue_patch(ca); advance(ca); }
    while (valid(cb)) {
      int nt, n, fs, ph; decode(cb.item, nt, n, fs, ph);
      const int view = (MODE == 2) ? cb.pi / chunks : 0;
      const int c = (MODE == 2) ? cb.pi - view * chunks : cb.pi;
      for (int t = 0; t < NT; ++t) {
        const int z = (MODE == 1) ? ph * 8 + t : view * 4 + t;     // pre-summed weight tile [i][j][a][b]
        if (tb == 0) mbar_wait(b_empty(gb), phb ^ 1u);
        if (PAIR) {
          const uint32_t lead = mapa_rank(b_full(sb), 0);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(b_full(sb), 2u * b_bytes);
            tma_load_3d_2sm(b_base + sb * b_al, &tmB, lead, c * BK, nt * p.BN + (int)rank * (p.BN / 2), z);
          }
        } else if (elect_one()) {
          mbar_expect_tx(b_full(sb), b_bytes);
          tma_load_3d(b_base + sb * b_al, &tmB, b_full(sb), c * BK, nt * p.BN, z);
        }
        __syncwarp();
        if (++tb == p.tg) { tb = 0; ++gb; }
        if (++sb == p.b_stages) { sb = 0; gb = 0; phb ^= 1u; }
      }
      advance(cb);
      if (valid(ca)) { issue_patch(ca); advance(ca); }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer (pair: the leader alone, M = 256 instructions spanning both CTAs); warp-uniform loop, elected lane issues.
      // A weight tile feeds only 4 (forward) or 4*MT (dgrad) MMAs, so this warp's instruction count per tile is what bounds the kernel (ncu source
      // view): descriptors are 32-bit low words (umma_tf32_lo), tap offsets are hoisted out of the unrolled tile loop, ring bookkeeping is static (TG).
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
      constexpr uint32_t layout = 2u;          // SWIZZLE_128B
      constexpr int ksteps = BK / 8;
      constexpr uint32_t rb16 = (uint32_t)BK * 4u / 16u;               // descriptor units (16 B) per patch row
      constexpr uint32_t dhi = desc_hi(8u * (uint32_t)BK * 4u, layout);
      const uint32_t a_lo0 = desc_lo(a_base), b_lo0 = desc_lo(b_base);
      const uint32_t a_st16 = patch_al >> 4, b_st16 = b_al >> 4;
      const int P16 = p.P * (int)rb16;
      auto mma = [](uint32_t dcol, uint32_t al, uint32_t bl, uint32_t id, uint32_t acc) {
        if (PAIR) umma_tf32_2cta_lo(dcol, al, bl, dhi, id, acc); else umma_tf32_lo(dcol, al, bl, dhi, id, acc);
      };
      auto commit = [](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
      int sa = 0, sb = 0, gb = 0; uint32_t pha = 0, phb = 0, it = 0;
      for (int64_t item = wid; item < p.total_items; item += nworkers, ++it) {
        int nt, n, fs, ph; decode(item, nt, n, fs, ph);
        const int r_lo = (fs - p.P - 1) / p.P;
        const int buf = it % p.NBUF; const uint32_t phacc = (it / p.NBUF) & 1u;
        mbar_wait(acc_empty(buf), phacc ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * NSLOT * p.BN);
        const int fb16 = (fs - r_lo * p.P) * (int)rb16;      // patch row of the unshifted first pixel (in [P+1, 2P]), in descriptor units
        int view = 0, cin_view = 0;
        for (int pi = 0; pi < ppi; ++pi) {
          mbar_wait(a_full(sa), pha);
          // row / column shift of tap (a, b) in descriptor units: forward (phase ph, j): (a + ph - 1) rows, (b + j - 1) columns;
          // dgrad (view i, j): -(a + i - 1) rows, -(b + j - 1) columns
          const int vi = view >> 1, vj = view & 1;
          const int row0 = (MODE == 1) ? (ph - 1) * P16 : -(vi - 1) * P16;                    // a = 0; a = 1 adds (MODE 1) / subtracts (MODE 2) P16
          const int col0 = (MODE == 1) ? -(int)rb16 : -(vj - 1) * (int)rb16;                  // MODE 1: b + j = 0; MODE 2: b = 0
          const uint32_t a_lo = a_lo0 + (uint32_t)sa * a_st16 + (uint32_t)fb16;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const int st = sb + t % TG;                               // weight-ring stage (sb = first stage of the release group)
            mbar_wait(b_full(st), phb);
            tc_fence_after();
            int off, slot; uint32_t first;
            if (MODE == 1) {
              const int j = t >> 2, a = (t >> 1) & 1, b = t & 1;
              off = row0 + a * P16 + col0 + (b + j) * (int)rb16;      // low-res input shift of this (phase, tap)
              slot = j;
              first = (pi == 0 && (t & 3) == 0) ? 0u : 1u;            // first MMA group into accumulator j of this item
            } else {
              const int a = t >> 1, b = t & 1;
              off = row0 - a * P16 + col0 - b * (int)rb16;            // the phase-(i,j) pixel of dy that tap (a,b) maps onto this input pixel
              slot = 0;
              first = (pi == 0 && t == 0) ? 0u : 1u;
            }
            const uint32_t bl = b_lo0 + (uint32_t)st * b_st16;
            const uint32_t al = a_lo + (uint32_t)off;                 // first patch row this MMA group reads (>= 0)
            if (elect_one()) {
              if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < ksteps; ++k)
                  mma(tacc + (uint32_t)(slot * p.BN), al + (uint32_t)(2 * k), bl + (uint32_t)(2 * k), idesc, k == 0 ? first : 1u);
              } else {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                  for (int k = 0; k < ksteps; ++k)
                    mma(tacc + (uint32_t)(mt * p.BN), al + (uint32_t)(mt * 128 * (int)rb16 + 2 * k), bl + (uint32_t)(2 * k), idesc, k == 0 ? first : 1u);
                }
              }
              if (t % TG == TG - 1) commit(b_empty(gb));
            }
            __syncwarp();
            if (t % TG == TG - 1) { sb += TG; ++gb; if (sb == p.b_stages) { sb = 0; gb = 0; phb ^= 1u; } }
          }
          if (elect_one()) commit(a_empty(sa));
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
          if (MODE == 2) { if (++cin_view == chunks) { cin_view = 0; ++view; } }
        }
        if (elect_one()) commit(acc_full(buf));
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: 8 warps.  Warp group e = (warp-2)/4 takes accumulator slot e (two slots) or column half e (one slot); every
    // 16-column chunk is transposed through a 2 KB per-warp staging tile so that 4 adjacent lanes own 64 contiguous bytes of one row.
    const int q = warp & 3;
    const int eg = (warp - 2) >> 2;
    const uint32_t stg = epi_base + (uint32_t)(warp - 2) * 2048u;
    const int lr = lane >> 2, lc = lane & 3;
    const uint32_t st_row = stg + (uint32_t)lane * 64u;
    const uint32_t st_sw = (uint32_t)((lane >> 1) & 3);
    uint32_t it = 0;
    for (int64_t item = wid; item < p.total_items; item += nworkers, ++it) {
      int nt, n, fs, ph; decode(item, nt, n, fs, ph);
      const int buf = it % p.NBUF; const uint32_t phacc = (it / p.NBUF) & 1u;
      const int co0 = nt * p.BN;
      const int slot = (NSLOT == 2) ? eg : 0;
      const int cbeg = (NSLOT == 2) ? 0 : eg * ((p.BN / 2 + 15) / 16 * 16);
      const int cend = (NSLOT == 2) ? p.BN : (eg == 0 ? (p.BN / 2 + 15) / 16 * 16 : p.BN);
      int64_t mrow[4]; bool vrow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = fs + (MODE == 2 ? 128 * slot : 0) + q * 32 + lr + 8 * j;
        const int hp = f / p.P, wp = f - hp * p.P;
        vrow[j] = (wp >= 1) && (wp <= p.W) && (hp >= 1) && (hp <= p.H) && (n < p.N);
        if (MODE == 1)   // low-res pixel (hp-1, wp-1), phase (ph, slot) -> high-res pixel (2h+i, 2w+j)
          mrow[j] = ((((int64_t)n * (2 * p.H)) + 2 * (hp - 1) + ph) * (2 * p.W) + 2 * (wp - 1) + slot) * p.Cout + co0 + 4 * lc;
        else
          mrow[j] = ((((int64_t)n * p.H) + (hp - 1)) * p.W + (wp - 1)) * p.Cout + co0 + 4 * lc;
      }
      const int g = (n < p.N) ? n / (p.N / p.G) : 0;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * NSLOT + slot) * p.BN);
      const float* sc = p.scale ? p.scale + (int64_t)g * p.Cout : nullptr;
      float4 rr[4];
      auto load_res = [&](int c, float4* dst) {
        if (p.res == nullptr || c >= cend || co0 + c + 4 * lc >= p.Cout) return;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (vrow[j]) dst[j] = __ldg(reinterpret_cast<const float4*>(p.res + mrow[j] + c));
      };
      load_res(cbeg, rr);
      mbar_wait(acc_full(buf), phacc);
      tc_fence_after();
      for (int c = cbeg; c < cend; c += 16) {
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
            if ((p.act & 3) == DGMR_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (p.act & DGMR_FLAG_ROUND_OUT) o = rna_tf32_e4(o);
            *reinterpret_cast<float4*>(p.y + mrow[j] + c) = o;
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = rn[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_rank(acc_empty(buf), 0));
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty(buf)) : "memory");
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { __syncwarp(); if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// ------------------------------------------------------------------ pre-summed weight packs
// tile index z = ((i*2 + j)*2 + a)*2 + b.  kh in S(i,a): i=0: a=0 -> {0}, a=1 -> {1,2};  i=1: a=0 -> {0,1}, a=1 -> {2}  (same for kw with j, b)
__device__ __forceinline__ void tap_set(int i, int a, int& k0, int& k1) {
  if (i == 0) { if (a == 0) { k0 = 0; k1 = 0; } else { k0 = 1; k1 = 2; } }
  else        { if (a == 0) { k0 = 0; k1 = 1; } else { k0 = 2; k1 = 2; } }
}
__device__ __forceinline__ float rna_tf32_s(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

// mode 0: packed[z][co][ci] (forward);  mode 1: packed[z][ci][co] (dgrad: rows = the dgrad's output channels)
__global__ void pack_weight_subpix_kernel(const float* __restrict__ w, float* __restrict__ packed, int Cout, int CinTot, int ci0, int Cin, int mode, int rnd) {
  const int64_t per = (int64_t)Cout * Cin, total = 16 * per;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(idx / per); const int64_t r = idx - (int64_t)z * per;
    int co, ci;
    if (mode == 0) { ci = (int)(r % Cin); co = (int)(r / Cin); } else { co = (int)(r % Cout); ci = (int)(r / Cout); }
    const int i = z >> 3, j = (z >> 2) & 1, a = (z >> 1) & 1, b = z & 1;
    int h0, h1, w0, w1; tap_set(i, a, h0, h1); tap_set(j, b, w0, w1);
    const float* wr = w + ((int64_t)co * CinTot + ci0 + ci) * 9;
    float s = 0.f;
    for (int kh = h0; kh <= h1; ++kh)
      for (int kw = w0; kw <= w1; ++kw) s += wr[kh * 3 + kw];
    packed[idx] = rnd ? rna_tf32_s(s) : s;
  }
}
// gw[co][ci0+ci][kh][kw] (+)= sum over the (i,a) sets containing kh and the (j,b) sets containing kw of dwp[z][co][ci]
__global__ void unpack_wgrad_subpix_kernel(const float* __restrict__ dwp, float* __restrict__ gw, int Cout, int CinTot, int ci0, int Cin, int acc) {
  const int64_t per = (int64_t)Cout * Cin, total = per * 9;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(idx % 9); const int64_t r = idx / 9; const int ci = (int)(r % Cin); const int co = (int)(r / Cin);
    const int kh = tap / 3, kw = tap - kh * 3;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        int h0, h1; tap_set(i, a, h0, h1);
        if (kh < h0 || kh > h1) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            int w0, w1; tap_set(j, b, w0, w1);
            if (kw < w0 || kw > w1) continue;
            s += dwp[(int64_t)(((i * 2 + j) * 2 + a) * 2 + b) * per + (int64_t)co * Cin + ci];
          }
      }
    const int64_t o = ((int64_t)co * CinTot + ci0 + ci) * 9 + tap;
    if (acc) gw[o] += s; else gw[o] = s;
  }
}

// ------------------------------------------------------------------ host side
static bool subpix_ok(int N, int H, int W, int Ck, int Cn, int G) {
  // Ck: K channels (multiple of 32: no channel tails here), Cn: N channels; images big enough that 128-position tiles are mostly pixels
  return Ck >= 32 && Ck % 32 == 0 && Cn >= 16 && Cn % 4 == 0 && W + 2 <= 256 && (int64_t)H * W >= 256 && G >= 1 && N % G == 0 && N >= 1;
}

static int launch_subpix(int mode, const float* a_ptr, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int H, int W,
                         int Ck, int Cn, int G, int act, cudaStream_t st) {
  SubpixParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Ck; p.Cout = Cn; p.G = G;
  p.P = W + 2;
  constexpr uint32_t row_bytes = 128u;
  const int slots_fwd = 2;
  if (mode == 1) {   // two accumulators per item: BN <= 128 keeps them double-buffered in the 512 TMEM columns
    p.n_tiles = (int)ceil_div(Cn, 128);
  } else {
    p.n_tiles = (int)ceil_div(Cn, 256);
  }
  {
    const int64_t m_items = (int64_t)N * ceil_div((int64_t)H * p.P - 2, 128);
    int want = (int)ceil_div((int64_t)sm_count(), m_items);
    int max_nt = Cn / 32 > 0 ? Cn / 32 : 1;
    if (want > max_nt) want = max_nt;
    if (want > p.n_tiles) p.n_tiles = want;
  }
  p.BN = (int)(ceil_div(ceil_div(Cn, p.n_tiles), 16) * 16);
  p.n_tiles = (int)ceil_div(Cn, p.BN);
  const int pair = (N >= 2 && sm_count() % 2 == 0) ? 1 : 0;
  const uint32_t budget = 208u * 1024u;
  const uint32_t b_al = (((uint32_t)(pair ? p.BN / 2 : p.BN) * row_bytes) + 1023u) & ~1023u;
  uint32_t patch_al = 0;
  p.a_stages = 2;
  const int64_t tiles_total = (int64_t)p.n_tiles * N * ceil_div((int64_t)H * p.P - 2, 128);
  p.MT = 1;
  if (mode == 2 && 2 * p.BN <= 512 && tiles_total >= 2 * (int64_t)sm_count()) p.MT = 2;
  for (;; --p.MT) {
    const int span = 128 * p.MT + 2 * p.P + 2;
    p.Rb = (int)ceil_div(p.P - 1 + span, p.P);
    patch_al = (((uint32_t)p.Rb * p.P * row_bytes) + 1023u) & ~1023u;
    if (2 * patch_al + 4 * b_al <= budget) break;
    if (p.MT == 1) { set_error("conv_subpix: patch does not fit in shared memory (W = %d)", W); return 1; }
  }
  const int nslot = (mode == 1) ? slots_fwd : p.MT;
  p.NBUF = (2 * nslot * p.BN <= 512) ? 2 : 1;
  if (nslot * p.BN > 512) { set_error("conv_subpix: accumulators do not fit in tensor memory"); return 1; }
  const int tiles_per_img = (int)ceil_div((int64_t)H * p.P - 2, 128);
  p.items_per_img = (mode == 1) ? 2 * tiles_per_img : (int)ceil_div(tiles_per_img, p.MT);
  p.qpairs = ceil_div((int64_t)N, 2);
  p.total_items = (int64_t)p.n_tiles * (pair ? p.qpairs : (int64_t)N) * p.items_per_img;
  p.act = act; p.bias = bias; p.scale = scale; p.res = res; p.y = y;
  p.sb_vec = (((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(scale)) & 15u) == 0) ? 1 : 0;
  if (((reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(a_ptr) | reinterpret_cast<uintptr_t>(wp)) & 15u) != 0) {
    set_error("conv_subpix: tensors must be 16-byte aligned"); return 1;
  }
  const int need_cols = p.NBUF * nslot * p.BN;
  p.tmem_cols = 32; while (p.tmem_cols < need_cols) p.tmem_cols <<= 1;
  const int NT = (mode == 1) ? 8 : 4;
  p.b_stages = (int)((budget - 2 * patch_al) / b_al);
  if (p.b_stages > 16) p.b_stages = 16;
  // weight tiles are handed back in groups of tg (one tcgen05.commit per group; a commit should follow >= ~8 MMAs)
  p.tg = (mode == 1) ? ((p.b_stages >= 8) ? 4 : 2) : 2;
  p.b_stages = p.b_stages / p.tg * p.tg;
  if (p.b_stages < 2 * p.tg && p.b_stages >= NT) { /* fine */ }
  if (p.b_stages < p.tg) { set_error("conv_subpix: weight ring too small"); return 1; }
  const size_t smem = (size_t)p.a_stages * patch_al + (size_t)p.b_stages * b_al + 1024 + 8 * (2 * p.a_stages + 2 * p.b_stages + 6) + 128 + 8 * 2048;
  CUtensorMap tmA[4], tmB;
  if (mode == 1) {
    uint64_t dims[5] = {(uint64_t)Ck, (uint64_t)W, (uint64_t)H, 1u, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Ck * 4, (uint64_t)W * Ck * 4, (uint64_t)H * W * Ck * 4, (uint64_t)H * W * Ck * 4};
    uint32_t box[5] = {32u, (uint32_t)p.P, (uint32_t)p.Rb, 1u, 1u};
    int e = make_tmap(&tmA[0], a_ptr, 5, dims, str, box, (int)row_bytes);
    if (e) return e;
    tmA[1] = tmA[2] = tmA[3] = tmA[0];
  } else {
    // four phase views of the high-resolution tensor [N, 2H, 2W, Ck]: element (c, w, h, n) of view (i, j) = dy[n, 2h+i, 2w+j, c]
    for (int v = 0; v < 4; ++v) {
      const int i = v >> 1, j = v & 1;
      uint64_t dims[5] = {(uint64_t)Ck, (uint64_t)W, (uint64_t)H, 1u, (uint64_t)N};
      uint64_t str[4] = {(uint64_t)2 * Ck * 4, (uint64_t)2 * (2 * W) * Ck * 4, (uint64_t)(2 * H) * (2 * W) * Ck * 4, (uint64_t)(2 * H) * (2 * W) * Ck * 4};
      uint32_t box[5] = {32u, (uint32_t)p.P, (uint32_t)p.Rb, 1u, 1u};
      int e = make_tmap(&tmA[v], a_ptr + ((int64_t)i * 2 * W + j) * Ck, 5, dims, str, box, (int)row_bytes);
      if (e) return e;
    }
  }
  {
    uint64_t dims[3] = {(uint64_t)Ck, (uint64_t)Cn, 16u};
    uint64_t str[2] = {(uint64_t)Ck * 4, (uint64_t)Cn * Ck * 4};
    uint32_t box[3] = {32u, (uint32_t)(pair ? p.BN / 2 : p.BN), 1u};
    int e = make_tmap(&tmB, wp, 3, dims, str, box, (int)row_bytes);
    if (e) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    const int lim = 226 * 1024;
    bool ok = true;
#define DGMR_SET(K) ok = ok && cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) == cudaSuccess
    DGMR_SET((conv_subpix_kernel<1, 1, false, 2>)); DGMR_SET((conv_subpix_kernel<1, 1, true, 2>));
    DGMR_SET((conv_subpix_kernel<1, 1, false, 4>)); DGMR_SET((conv_subpix_kernel<1, 1, true, 4>));
    DGMR_SET((conv_subpix_kernel<2, 1, false, 2>)); DGMR_SET((conv_subpix_kernel<2, 1, true, 2>));
    DGMR_SET((conv_subpix_kernel<2, 2, false, 2>)); DGMR_SET((conv_subpix_kernel<2, 2, true, 2>));
#undef DGMR_SET
    if (!ok) { set_error("conv_subpix: cannot raise dynamic smem limit"); return 2; }
    attr_set = true;
  }
  int64_t grid = sm_count();
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (pair) {
    if (grid > 2 * p.total_items) grid = 2 * p.total_items;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  } else if (grid > p.total_items) grid = p.total_items;
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kSubThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaError_t e;
#define DGMR_GO(...) cudaLaunchKernelEx(&cfg, conv_subpix_kernel<__VA_ARGS__>, tmA[0], tmA[1], tmA[2], tmA[3], tmB, p)
  if (mode == 1 && p.tg == 4) e = pair ? DGMR_GO(1, 1, true, 4) : DGMR_GO(1, 1, false, 4);
  else if (mode == 1) e = pair ? DGMR_GO(1, 1, true, 2) : DGMR_GO(1, 1, false, 2);
  else if (p.MT == 2) e = pair ? DGMR_GO(2, 2, true, 2) : DGMR_GO(2, 2, false, 2);
  else e = pair ? DGMR_GO(2, 1, true, 2) : DGMR_GO(2, 1, false, 2);
#undef DGMR_GO
  if (e != cudaSuccess) { set_error("conv_subpix: launch failed: %s", cudaGetErrorString(e)); return 2; }
  return 0;
}

}  // namespace dgmr

using namespace dgmr;

extern "C" {

int dgmr_upconv_supported(int N, int H, int W, int Cin, int Cout) {
  // forward (K = Cin, N = Cout), dgrad (K = Cout, N = Cin) and the 16-tile wgrad must all be served
  return (subpix_ok(N, H, W, Cin, Cout, 1) && subpix_ok(N, H, W, Cout, Cin, 1)) ? 1 : 0;
}

int dgmr_pack_weight_subpix(const float* w, float* packed, int Cout, int CinTot, int ci0, int Cin, int mode, dgmr_stream_t stream) {
  const int rnd = (mode & DGMR_FLAG_ROUND_TF32) ? 1 : 0;
  mode &= ~DGMR_FLAG_ROUND_TF32;
  DGMR_REQUIRE(w && packed && Cout > 0 && Cin > 0 && ci0 >= 0 && ci0 + Cin <= CinTot && (mode == 0 || mode == 1), "dgmr_pack_weight_subpix: bad arguments");
  pack_weight_subpix_kernel<<<ew_grid((int64_t)16 * Cout * Cin, 256, 2), 256, 0, S(stream)>>>(w, packed, Cout, CinTot, ci0, Cin, mode, rnd);
  DGMR_CHECK_LAUNCH("dgmr_pack_weight_subpix");
  return 0;
}

int dgmr_unpack_wgrad_subpix(const float* dwp, float* gw, int Cout, int CinTot, int ci0, int Cin, int accumulate, dgmr_stream_t stream) {
  DGMR_REQUIRE(dwp && gw && Cout > 0 && Cin > 0 && ci0 >= 0 && ci0 + Cin <= CinTot, "dgmr_unpack_wgrad_subpix: bad arguments");
  unpack_wgrad_subpix_kernel<<<ew_grid((int64_t)9 * Cout * Cin, 256, 2), 256, 0, S(stream)>>>(dwp, gw, Cout, CinTot, ci0, Cin, accumulate);
  DGMR_CHECK_LAUNCH("dgmr_unpack_wgrad_subpix");
  return 0;
}

int dgmr_upconv_fwd(const float* x, const float* wsp, const float* bias, const float* scale, const float* res, float* y, int N, int H, int W, int Cin,
                    int Cout, int G, int act, dgmr_stream_t stream) {
  DGMR_REQUIRE(subpix_ok(N, H, W, Cin, Cout, G), "dgmr_upconv_fwd: shape not supported (N=%d H=%d W=%d Cin=%d Cout=%d G=%d)", N, H, W, Cin, Cout, G);
  DGMR_REQUIRE((act & ~(3 | DGMR_FLAG_ROUND_OUT)) == 0 && (act & 3) <= 1, "dgmr_upconv_fwd: bad act");
  // wide channels on small low-resolution images: whole-row tiles with CTA pairs sharing each pre-summed weight tile (conv_kwstack.cu); the
  // weight stream per 128-position item is what bounds the patch form there
  if (g_subpix_rows != 0 && W <= 32 && umma_pairconv_ok(N, 1, H, W, Cin, Cout, 1, 3, 3, G) && (Cin >= 192 || g_subpix_rows == 1))
    return launch_conv_umma_pairconv_subpix(x, wsp, bias, scale, res, y, N, H, W, Cin, Cout, G, act, S(stream));
  return launch_subpix(1, x, wsp, bias, scale, res, y, N, H, W, Cin, Cout, G, act, S(stream));
}

int dgmr_upconv_dgrad(const float* dz, const float* wspt, float* dx, int N, int H, int W, int Cin, int Cout, dgmr_stream_t stream) {
  DGMR_REQUIRE(subpix_ok(N, H, W, Cout, Cin, 1), "dgmr_upconv_dgrad: shape not supported (N=%d H=%d W=%d Cin=%d Cout=%d)", N, H, W, Cin, Cout);
  return launch_subpix(2, dz, wspt, nullptr, nullptr, nullptr, dx, N, H, W, Cout, Cin, 1, DGMR_ACT_NONE, S(stream));
}

int dgmr_upconv_wgrad(const float* x, const float* dz, float* dwsp, int N, int H, int W, int Cin, int Cout, dgmr_stream_t stream) {
  DGMR_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0 && (reinterpret_cast<uintptr_t>(dwsp) & 15u) == 0, "dgmr_upconv_wgrad: shape not supported");
  return launch_conv_umma_wgrad_subpix(x, dz, dwsp, N, H, W, Cin, Cout, S(stream));
}

}  // extern "C"
