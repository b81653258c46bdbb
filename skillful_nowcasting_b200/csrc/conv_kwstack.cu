// Narrow-output 3x3(x3) convolutions on tcgen05 with the three COLUMN taps stacked along the MMA's N dimension.
//
// A layer with few output channels (Cout = 48: the first temporal-discriminator block, the last sampler block, their dgrads) is the worst case of
// the plain implicit-GEMM kernel (conv_umma.cu): a kind::tf32 MMA of M = 128, N = 48 still occupies the tensor pipe for its ~60-cycle minimum (40 % of
// peak at best), and every filter tap re-reads its 128-pixel activation tile from L2 (27 re-reads for a 3x3x3 filter), which is what actually bounds
// those launches (~45 B/clk/SM of L2->SM traffic).  Here a tile is 128 consecutive pixels of WHOLE image rows (W in {32, 64, 128}) and one MMA multiplies
// the activation tile of filter row (kd, kh) -- unshifted in w -- with the weights of its three column taps side by side,
//
//     T[p][kw*Cout + co] += sum_ci x[p + (kd-1, kh-1, 0)][ci] * W[kd][kh][kw][co][ci]            N = 3*Cout (144: 72 cycles for three taps),
//
// which is a contiguous [3*Cout][Cin] slab of the ordinary packed weights ([tap][Cout][Cin], taps ordered (kd, kh, kw)): no new packing.  The
// column shift moves to the epilogue:   y[p] = T[p-1][kw=0] + T[p][kw=1] + T[p+1][kw=2],   with the first / last pixel of an image row dropping the term
// that would come from the zero padding -- because tiles are whole rows, every p +- 1 that is needed lies inside the tile.  Rows live in TMEM lanes, so
// the shift is a warp shuffle plus a 16-float exchange between neighbouring epilogue warps.  Per tile that is 3x fewer activation bytes through
// L2->SM and 2.5x fewer tensor-pipe cycles than the tap-by-tap form.  Same persistent structure as conv_umma_fwd_persist_kernel: one TMA warp, one MMA
// warp, four epilogue warps, double-buffered accumulators (2 x 3*Cout TMEM columns), stages released in groups.
#include "umma_common.cuh"

namespace dgmr {

struct KwStackParams {
  int N, D, H, W, Cin, Cout, kd, kh, G;
  int bh;                // image rows per tile: 128 / W
  int BN;                // STACK: 3 * Cout; else the Cout tile (<= 256)
  int n_tiles;           // Cout tiles (STACK: 1)
  int subpix;            // STACK = false only: sub-pixel up-convolution (conv_subpix.cu) -- items carry an output phase (i, j); its four taps (a, b) read the
                         // low-resolution tile shifted by (a+i-1, b+j-1) with the pre-summed weight tile [i][j][a][b]; outputs go to pixel (2h+i, 2w+j)
  int stages, cg, tmem_cols, act, round_out, res_up2;
  const float* bias; const float* scale; const float* res; float* y;
};

constexpr int kKwThreads = 192;  // warp0 TMA, warp1 MMA, warps 2..5 epilogue

// PAIR: launched as clusters of 2 CTAs (one TPC), tcgen05.mma.cta_group::2 with M = 256: the two CTAs work on two different tiles (identical
// shared-memory geometry, one A descriptor), each stages HALF of every stacked weight tile (cp.async.bulk.tensor...cta_group::2 signalling the
// leader's mbarrier) -- the weight slab is ~half of this kernel's L2->SM traffic, which is what bounds it -- and the leader's MMA warp issues for
// both; commits are multicast to both CTAs, the peer's epilogue warps release the accumulators on the leader's barrier.
// STACK = false: the same persistent (pair) structure with ORDINARY taps -- one K block per (kd, kh, kw, channel chunk), the activation tile shifted in w
// by TMA, Cout in tiles of BN <= 256, plain epilogue.  That is the plain persistent kernel of conv_umma.cu plus CTA pairs: the wide 16x16 sampler
// layers (768 -> 768, 384 -> 384: few pixels, large weights) are bound by the weight stream per 128-pixel tile there (~45 B/clk/SM at 557 TF/s);
// a pair halves it.
template <int BK, bool PAIR, bool STACK>
__global__ void __launch_bounds__(kKwThreads, 1)
conv_umma_kwstack_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KwStackParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int64_t wid = PAIR ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;      // worker (CTA or CTA pair) index
  const int64_t nworkers = PAIR ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  const uint32_t a_bytes = 128u * BK * 4u, b_bytes = (uint32_t)(PAIR ? p.BN / 2 : p.BN) * BK * 4u;   // pair: half of the weight rows per CTA
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
  const uint32_t xchg_base = epi_base + 4u * 2048u;                  // [2 parities][4 warps][2: T0 of lane 31 | T2 of lane 0][16 floats]

  const int tiles_h = p.H / p.bh;
  const int64_t total_tiles = (int64_t)p.N * p.D * tiles_h;
  // work items: single tiles, or (pair) two consecutive tiles 2i + rank; with an odd total the peer's last tile lies past the end
  // (n0 = N: TMA zero fill, stores skipped)
  const int64_t m_items = PAIR ? (total_tiles + 1) / 2 : total_tiles;
  const int nphase = (!STACK && p.subpix) ? 4 : 1;
  const int64_t total_items = m_items * p.n_tiles * nphase;        // item = (phase, Cout tile, tile or tile pair): tiles are the fast index
  auto tile_of = [&](int64_t item) { const int64_t mi = item % m_items; return PAIR ? 2 * mi + (int64_t)rank : mi; };
  auto co_of = [&](int64_t item) { return (int)((item / m_items) % p.n_tiles) * p.BN; };
  auto phase_of = [&](int64_t item) { return (int)(item / (m_items * p.n_tiles)); };
  auto decode = [&](int64_t t, int& n0, int& d0, int& h0) {
    h0 = (int)(t % tiles_h) * p.bh; t /= tiles_h;
    d0 = (int)(t % p.D); n0 = (int)(t / p.D);
  };
  const int rtaps = (!STACK && p.subpix) ? 4 : p.kd * p.kh * (STACK ? 1 : 3);   // K-block taps: filter rows (kd, kh), single taps (kd, kh, kw), or (a, b)
  const int kchunks = (p.Cin + BK - 1) / BK;
  const int num_kb = rtaps * kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int bsel = 0; bsel < 2; ++bsel) { mbar_init(tmem_full(bsel), 1); mbar_init(tmem_empty(bsel), PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); if (PAIR) tmem_alloc2(tmem_ptr_addr, (uint32_t)p.tmem_cols); else tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // pair: the peer's barriers must exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===== TMA producer: one continuous stream of K blocks over all of this CTA's tiles
    int s = 0, g = 0, sg = 0; uint32_t ph = 0;
    for (int64_t item = wid; item < total_items; item += nworkers) {
      int n0, d0, h0; decode(tile_of(item), n0, d0, h0);
      const int co0 = co_of(item);
      const int phs = phase_of(item), pi = phs >> 1, pj = phs & 1;
      int rt = 0, chunk = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int c0 = chunk * BK;
        int wsh, hsh, dsh, bz;           // shifts of this K block's activation tile, weight tile index
        if (!STACK && p.subpix) {
          const int a = rt >> 1, b = rt & 1;
          hsh = a + pi - 1; wsh = b + pj - 1; dsh = 0; bz = phs * 4 + rt;
        } else {
          const int frow = STACK ? rt : rt / 3;                     // filter row (kd, kh)
          hsh = frow % p.kh - p.kh / 2; dsh = frow / p.kh - p.kd / 2;
          wsh = STACK ? 0 : rt % 3 - 1; bz = rt;
        }
        if (sg == 0) mbar_wait(empty_bar(g), ph ^ 1u);
        const uint32_t sa = base + s * stage_bytes;
        if (PAIR) {
          // the leader's full barrier collects both CTAs' bytes (one arrival: its own expect_tx of the doubled count; the peer's bytes may
          // land first, the phase cannot complete before that arrival)
          const uint32_t lead = mapa_rank(full_bar(s), 0);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(full_bar(s), 2u * (a_bytes + b_bytes));
            tma_load_5d_2sm(sa, &tmA, lead, c0, wsh, h0 + hsh, d0 + dsh, n0);
            tma_load_3d_2sm(sa + a_bytes, &tmB, lead, c0, co0 + (int)rank * (p.BN / 2), bz);
          }
        } else if (elect_one()) {
          mbar_expect_tx(full_bar(s), a_bytes + b_bytes);
          tma_load_5d(sa, &tmA, full_bar(s), c0, wsh, h0 + hsh, d0 + dsh, n0);
          tma_load_3d(sa + a_bytes, &tmB, full_bar(s), c0, co0, bz);
        }
        __syncwarp();
        if (++chunk == kchunks) { chunk = 0; ++rt; }
        if (++sg == p.cg) { sg = 0; ++g; }
        if (++s == p.stages) { s = 0; g = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
   if (rank == 0) {
    // ===== MMA issuer (pair: the leader alone, M = 256 instructions spanning both CTAs).  This warp's instruction stream bounds the kernel
    // (ncu source view: ~94 instructions per 4-MMA stage at ~4.4 cycles each against 288 cycles of MMA), so descriptors are advanced as 32-bit
    // low words and the per-stage bookkeeping is kept to a handful of adds / selects.
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
    constexpr uint32_t row_bytes = BK * 4u;
    constexpr uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    constexpr uint32_t dhi = desc_hi(8u * row_bytes, layout);
    auto mma = [](uint32_t dcol, uint32_t al, uint32_t bl, uint32_t id, uint32_t acc) {
      if (PAIR) umma_tf32_2cta_lo(dcol, al, bl, dhi, id, acc); else umma_tf32_lo(dcol, al, bl, dhi, id, acc);
    };
    auto commit = [](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
    const int tail_ks = (p.Cin % BK) ? (p.Cin % BK) / 8 : BK / 8;
    const uint32_t lo0 = desc_lo(base), st16 = stage_bytes >> 4, a16 = a_bytes >> 4;
    // ring state as running values: descriptor low word, full / empty barrier addresses (no multiplications in the loop).  The last, possibly
    // partial, release group is never committed: nothing refills those stages, and tmem_full covers the completion of their MMAs.
    int s = 0, sg = 0; uint32_t ph = 0, it = 0;
    uint32_t a_lo = lo0, fbar = full_bar(0), ebar = empty_bar(0);
    for (int64_t item = wid; item < total_items; item += nworkers, ++it) {
      const int bsel = it & 1; const uint32_t phacc = (it >> 1) & 1u;
      mbar_wait(tmem_empty(bsel), phacc ^ 1u);       // the epilogue warps (of both CTAs) have drained this accumulator buffer
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(bsel * p.BN);
      const uint32_t tfull = tmem_full(bsel);
      int chunk_i = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(fbar, ph);
        tc_fence_after();
        const uint32_t b_lo = a_lo + a16;
        const bool last_chunk = (++chunk_i == kchunks);
        if (last_chunk) chunk_i = 0;
        const bool rel = (++sg == p.cg);
        if (elect_one()) {
          mma(tacc, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u);
          if (last_chunk && tail_ks != BK / 8) {
#pragma unroll
            for (int k = 1; k < BK / 8; ++k)
              if (k < tail_ks) mma(tacc, a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), idesc, 1u);
          } else {
#pragma unroll
            for (int k = 1; k < BK / 8; ++k) mma(tacc, a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), idesc, 1u);
          }
          if (rel) commit(ebar);
          if (kb + 1 == num_kb) commit(tfull);
        }
        __syncwarp();
        a_lo += st16; fbar += 8u;
        if (rel) { sg = 0; ebar += 8u; }
        if (++s == p.stages) { s = 0; ph ^= 1u; a_lo = lo0; fbar = full_bar(0); ebar = empty_bar(0); }
      }
    }
   }
  } else {
    // ===== epilogue (4 warps): warp w may only touch TMEM lanes [32*(w%4), +32).  Lane = tile row r = 32q + lane = pixel (h0 + r / W, r % W).
    const int q = warp & 3;
    const uint32_t stg = epi_base + (uint32_t)q * 2048u;
    const int lr = lane >> 2, lc = lane & 3;
    const uint32_t st_row = stg + (uint32_t)lane * 64u, st_sw = (uint32_t)((lane >> 1) & 3);
    const int my_w = (q * 32 + lane) % p.W;
    const bool has_prev = my_w != 0, has_next = my_w != p.W - 1;     // neighbours inside the image row (else: zero padding, term dropped)
    const bool cross = p.W > 32;                                     // rows continue across epilogue-warp boundaries
    uint32_t it = 0, par = 0;
    for (int64_t item = wid; item < total_items; item += nworkers, ++it) {
      const int64_t t = tile_of(item);
      const bool live = t < total_tiles;             // pair with an odd tile count: the peer's last tile does not exist
      int n0, d0, h0; decode(live ? t : 0, n0, d0, h0);
      const int co0 = co_of(item);
      const int cend = STACK ? p.Cout : (p.Cout - co0 < p.BN ? p.Cout - co0 : p.BN);     // output channels of this item
      const int bsel = it & 1; const uint32_t phacc = (it >> 1) & 1u;
      uint32_t mrow[4], rrow[4]; const float* srow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rj = q * 32 + lr + 8 * j;
        const int wj = rj % p.W, hj = h0 + rj / p.W;
        if (!STACK && p.subpix) {   // low-resolution pixel (hj, wj), phase (i, j) -> pixel (2h+i, 2w+j) of the [N, 2H, 2W, Cout] output
          const int phs = phase_of(item);
          mrow[j] = (uint32_t)((n0 * 2 * p.H + 2 * hj + (phs >> 1)) * (2 * p.W) + 2 * wj + (phs & 1)) * (uint32_t)p.Cout + (uint32_t)co0 + 4u * lc;
          rrow[j] = mrow[j];
        } else {
        mrow[j] = (uint32_t)(((n0 * p.D + d0) * p.H + hj) * p.W + wj) * (uint32_t)p.Cout + (uint32_t)co0 + 4u * lc;
        rrow[j] = p.res_up2 ? (uint32_t)(((n0 * p.D + d0) * (p.H >> 1) + (hj >> 1)) * (p.W >> 1) + (wj >> 1)) * (uint32_t)p.Cout + (uint32_t)co0 + 4u * lc : mrow[j];
        }
        srow[j] = p.scale ? p.scale + (int64_t)(n0 / (p.N / p.G)) * p.Cout + co0 : nullptr;
      }
      mbar_wait(tmem_full(bsel), phacc);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(bsel * p.BN);
      for (int c = 0; c < cend; c += 16, par ^= 1u) {
        float4 rr[4];
        if (p.res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = __ldg(reinterpret_cast<const float4*>(p.res + rrow[j] + c));
        }
        float t0[16], v[16], t2[16];
        if (!STACK) {
          tmem_ld16(trow + (uint32_t)c, v);
        } else {
        tmem_ld16(trow + (uint32_t)c, t0);
        tmem_ld16(trow + (uint32_t)(p.Cout + c), v);
        tmem_ld16(trow + (uint32_t)(2 * p.Cout + c), t2);
        if (cross) {
          const uint32_t xb = xchg_base + par * 512u + (uint32_t)q * 128u;
          if (lane == 31) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xb + 16u * k), "f"(t0[4 * k]), "f"(t0[4 * k + 1]), "f"(t0[4 * k + 2]), "f"(t0[4 * k + 3]) : "memory");
          }
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xb + 64u + 16u * k), "f"(t2[4 * k]), "f"(t2[4 * k + 1]), "f"(t2[4 * k + 2]), "f"(t2[4 * k + 3]) : "memory");
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");    // the four epilogue warps
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float a = __shfl_up_sync(0xffffffffu, t0[k], 1);      // T[r-1][kw=0]
          const float b = __shfl_down_sync(0xffffffffu, t2[k], 1);    // T[r+1][kw=2]
          t0[k] = a; t2[k] = b;
        }
        if (cross) {
          if (lane == 0 && has_prev) {        // r-1 is lane 31 of the previous warp (has_prev => q > 0)
            const uint32_t xa = xchg_base + par * 512u + (uint32_t)(q - 1) * 128u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t0[4 * k]), "=f"(t0[4 * k + 1]), "=f"(t0[4 * k + 2]), "=f"(t0[4 * k + 3]) : "r"(xa + 16u * k) : "memory");
          }
          if (lane == 31 && has_next) {       // r+1 is lane 0 of the next warp (has_next => q < 3)
            const uint32_t xa = xchg_base + par * 512u + (uint32_t)(q + 1) * 128u + 64u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t2[4 * k]), "=f"(t2[4 * k + 1]), "=f"(t2[4 * k + 2]), "=f"(t2[4 * k + 3]) : "r"(xa + 16u * k) : "memory");
          }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] += (has_prev ? t0[k] : 0.f) + (has_next ? t2[k] : 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (((uint32_t)k ^ st_sw) << 4)), "f"(v[4 * k]), "f"(v[4 * k + 1]),
                       "f"(v[4 * k + 2]), "f"(v[4 * k + 3]) : "memory");
        __syncwarp();
        const int co = c + 4 * lc;             // relative to co0 (srow / mrow carry co0; the bias does not)
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) b4 = make_float4(__ldg(p.bias + co0 + co), __ldg(p.bias + co0 + co + 1), __ldg(p.bias + co0 + co + 2), __ldg(p.bias + co0 + co + 3));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = lr + 8 * j;
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                       : "r"(stg + (uint32_t)row * 64u + (((uint32_t)lc ^ (uint32_t)((row >> 1) & 3)) << 4)) : "memory");
          if (srow[j]) { o.x *= __ldg(srow[j] + co); o.y *= __ldg(srow[j] + co + 1); o.z *= __ldg(srow[j] + co + 2); o.w *= __ldg(srow[j] + co + 3); }
          o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
          if (p.res) { o.x += rr[j].x; o.y += rr[j].y; o.z += rr[j].z; o.w += rr[j].w; }
          if (p.act == DGMR_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (p.round_out) o = rna_tf32_e4(o);
          if (live) *reinterpret_cast<float4*>(p.y + mrow[j] + c) = o;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_rank(tmem_empty(bsel), 0));   // the leader's MMA warp waits for both CTAs' epilogues
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty(bsel)) : "memory");
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { __syncwarp(); if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

int g_kwstack_pair = -1;     // dgmr_set_option("kwstack_pair"): 0 = never CTA pairs, 1 = pairs whenever there are at least two tiles (tests)

static int kw_bk(int Cin) { return (Cin % 8 != 0) ? 0 : (Cin >= 32) ? 32 : (Cin == 16 || Cin == 24) ? 16 : 0; }

bool umma_kwstack_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G) {
  if (kw != 3 || kh != 3 || !(kd == 1 || kd == 3)) return false;
  if (!(W == 32 || W == 64 || W == 128) || H % (128 / W) != 0) return false;
  if (Cout % 16 != 0 || Cout < 16 || 3 * Cout > 256) return false;
  if (kw_bk(Cin) == 0 || (Cin < 32 && Cin != 16)) return false;
  if (G < 1 || N % G) return false;
  if ((int64_t)N * D * H * W * Cout >= ((int64_t)1 << 32) || (int64_t)N * D * H * W >= ((int64_t)1 << 31)) return false;   // 32-bit epilogue offsets
  return true;
}
// the same kernel with ordinary taps (STACK = false): whole-row tiles of 16..128-pixel-wide images, any Cout that is a multiple of 16
bool umma_pairconv_ok(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int G) {
  if (kw != 3 || kh != 3 || !(kd == 1 || kd == 3)) return false;
  if (!(W == 16 || W == 32 || W == 64 || W == 128) || H % (128 / W) != 0) return false;
  if (Cout % 16 != 0 || Cout < 16) return false;
  if (Cin % 8 != 0 || Cin < 32) return false;
  if (G < 1 || N % G) return false;
  if ((int64_t)N * D * H * W * Cout >= ((int64_t)1 << 32) || (int64_t)N * D * H * W >= ((int64_t)1 << 31)) return false;
  return true;
}

static int launch_kw_impl(int mode /* 1: column-stacked, 0: ordinary taps, 2: sub-pixel up-convolution tiles */, const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int D, int H,
                          int W, int Cin, int Cout, int kd, int G, int act, cudaStream_t st) {
  const bool stack = mode == 1;
  KwStackParams p;
  p.subpix = mode == 2 ? 1 : 0;
  p.round_out = (act & DGMR_FLAG_ROUND_OUT) ? 1 : 0; p.res_up2 = (act & DGMR_FLAG_RES_UP2) ? 1 : 0;
  act &= 3;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = 3; p.G = G;
  p.bh = 128 / W;
  if (stack) { p.n_tiles = 1; p.BN = 3 * Cout; }
  else { p.n_tiles = (int)ceil_div(Cout, 256); p.BN = (int)(ceil_div(ceil_div(Cout, p.n_tiles), 16) * 16); p.n_tiles = (int)ceil_div(Cout, p.BN); }
  p.act = act; p.bias = bias; p.scale = scale; p.res = res; p.y = y;
  const int BK = kw_bk(Cin);
  if (BK == 0) { set_error("conv_umma_kwstack: Cin=%d not served", Cin); return 1; }
  p.tmem_cols = 32; while (p.tmem_cols < 2 * p.BN) p.tmem_cols <<= 1;
  const int64_t total_tiles = (int64_t)N * D * (H / p.bh);
  // CTA pairs (cta_group::2, half of every weight tile per CTA) whenever there are a few waves of tiles; the pair splits the N rows of the weight
  // tile in two halves of whole 8-row groups
  const bool pair = g_kwstack_pair != 0 && sm_count() % 2 == 0 && (p.BN / 2) % 8 == 0 &&
                    (total_tiles * p.n_tiles * (mode == 2 ? 4 : 1) >= 2 * (int64_t)sm_count() || (g_kwstack_pair == 1 && total_tiles >= 2));
  const uint32_t a_bytes = 128u * BK * 4u, b_bytes = ((uint32_t)(pair ? p.BN / 2 : p.BN) * BK * 4u + 1023u) & ~1023u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  int stages = (int)((196u * 1024u) / stage_bytes);
  if (stages > 9) stages = 9;
  if (stages < 2) { set_error("conv_umma_kwstack: stage too large"); return 1; }
  p.cg = 1;
  if (stages >= 6 && p.BN <= 160) { stages = stages / 3 * 3; p.cg = 3; } else if (stages >= 4) { stages = stages / 2 * 2; p.cg = 2; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 6) + 128 + 4 * 2048 + 1024;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, (uint64_t)D * H * W * Cin * 4};
    uint32_t box[5] = {(uint32_t)BK, (uint32_t)W, (uint32_t)p.bh, 1u, 1u};
    int e = make_tmap(&tmA, x, 5, dims, str, box, BK * 4);
    if (e) return e;
  }
  if (stack) {
    // packed weights [tap = (kd, kh, kw)][Cout][Cin] seen as [(kd, kh)][kw*Cout + co][Cin]
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)(3 * Cout), (uint64_t)(kd * 3)};
    uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)3 * Cout * Cin * 4};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)(pair ? p.BN / 2 : p.BN), 1u};
    int e = make_tmap(&tmB, wp, 3, dims, str, box, BK * 4);
    if (e) return e;
  } else {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(mode == 2 ? 16 : kd * 9)};
    uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)Cout * Cin * 4};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)(pair ? p.BN / 2 : p.BN), 1u};
    int e = make_tmap(&tmB, wp, 3, dims, str, box, BK * 4);
    if (e) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    bool ok = true;
#define DGMR_SET(...) ok = ok && cudaFuncSetAttribute(conv_umma_kwstack_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024)) == cudaSuccess
    DGMR_SET(32, false, true); DGMR_SET(16, false, true); DGMR_SET(32, true, true); DGMR_SET(16, true, true);
    DGMR_SET(32, false, false); DGMR_SET(32, true, false);
#undef DGMR_SET
    if (!ok) { set_error("conv_umma_kwstack: cannot raise dynamic smem limit"); return 2; }
    attr_set = true;
  }
  // one CTA per SM (2 x BN TMEM columns each): pad the shared-memory request so that a second CTA can never become resident and block in tcgen05.alloc
  size_t req = smem;
  if (req < (size_t)232448 / 2 + 1024) req = (size_t)232448 / 2 + 1024;
  if (pair) {
    int64_t g = sm_count();                                   // even (checked above): one CTA pair per TPC
    const int64_t items = (total_tiles + 1) / 2 * p.n_tiles * (mode == 2 ? 4 : 1);
    if (g > 2 * items) g = 2 * items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)g); cfg.blockDim = dim3(kKwThreads); cfg.dynamicSmemBytes = req; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = !stack ? cudaLaunchKernelEx(&cfg, conv_umma_kwstack_kernel<32, true, false>, tmA, tmB, p)
                   : (BK == 32) ? cudaLaunchKernelEx(&cfg, conv_umma_kwstack_kernel<32, true, true>, tmA, tmB, p)
                                : cudaLaunchKernelEx(&cfg, conv_umma_kwstack_kernel<16, true, true>, tmA, tmB, p);
    if (e != cudaSuccess) { set_error("conv_umma_kwstack: cluster launch failed: %s", cudaGetErrorString(e)); return 2; }
    return 0;
  }
  int64_t g = sm_count();
  if (g > total_tiles * p.n_tiles * (mode == 2 ? 4 : 1)) g = total_tiles * p.n_tiles * (mode == 2 ? 4 : 1);
  if (!stack) conv_umma_kwstack_kernel<32, false, false><<<(unsigned)g, kKwThreads, req, st>>>(tmA, tmB, p);
  else if (BK == 32) conv_umma_kwstack_kernel<32, false, true><<<(unsigned)g, kKwThreads, req, st>>>(tmA, tmB, p);
  else conv_umma_kwstack_kernel<16, false, true><<<(unsigned)g, kKwThreads, req, st>>>(tmA, tmB, p);
  DGMR_CHECK_LAUNCH("conv_umma_kwstack");
  return 0;
}

int launch_conv_umma_kwstack(const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int D, int H, int W,
                             int Cin, int Cout, int kd, int G, int act, cudaStream_t st) {
  return launch_kw_impl(1, x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, G, act, st);
}
int launch_conv_umma_pairconv(const float* x, const float* wp, const float* bias, const float* scale, const float* res, float* y, int N, int D, int H, int W,
                              int Cin, int Cout, int kd, int G, int act, cudaStream_t st) {
  return launch_kw_impl(0, x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, G, act, st);
}
// sub-pixel up-convolution forward on the same kernel: x [N,H,W,Cin] (low resolution), wsp [16][Cout][Cin], y / res [N,2H,2W,Cout]
int launch_conv_umma_pairconv_subpix(const float* x, const float* wsp, const float* bias, const float* scale, const float* res, float* y, int N, int H, int W,
                                     int Cin, int Cout, int G, int act, cudaStream_t st) {
  return launch_kw_impl(2, x, wsp, bias, scale, res, y, N, 1, H, W, Cin, Cout, 1, G, act, st);
}

}  // namespace dgmr
