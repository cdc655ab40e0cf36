// Third generation of the tcgen05 GEMM / implicit-GEMM convolution (operand scheme: gemm_tc.cuh).
//
// What changed against gemm_tc2 and why (profiles/ncu_gemm_tc2_r1_*.txt):
//  * CTA PAIRS (cta_group::2, 2-CTA clusters): the two SMs of a TPC compute one 256 x BLOCK_N tile.  Each CTA stages its
//    own 128 A rows and HALF of the B tile; the MMA reads both halves, so the shared-memory traffic per FLOP (TMA fill +
//    operand read, the resource that capped the single-CTA kernel at 52-61 % tensor-pipe on the large convolutions)
//    drops by a quarter to a third.  CTAS = 1 instantiates the same kernel without clusters.
//  * EPILOGUE THROUGH SHARED MEMORY + TMA: the accumulator goes TMEM -> registers -> (bias / time shift / folded
//    LayerNorm / residual) -> bf16 -> swizzled shared memory, and 128 x 32 boxes are written by TMA stores; residual
//    boxes arrive by TMA loads into the same staging buffers ahead of time.  The old epilogue issued 16-byte
//    row-strided global loads/stores (32 cache lines per instruction): on the K = 320..1280 token GEMMs it, not the
//    MMA, set the tile time.
//  * LAYERNORM FOLDED IN: a producer GEMM can emit per-row partial (sum, sum of squares) of its bf16 output; a
//    consumer GEMM whose A operand is that raw tensor and whose weights were pre-multiplied by gamma applies
//    rstd * (acc - mean * colsum_n) + c_n in its epilogue (attention.py:85,104,120 / blocks.py:67-71 LayerNorms).
//
//   warp 0: TMA producer   warp 1: MMA issuer (leader CTA)   warp 2: TMEM alloc + residual loads + output stores
//   warps 3..10: epilogue (TMEM lane quarter = warp & 3; the two warps of a quarter alternate 32-column chunks)
#pragma once
#include "common_host.h"
#include "gemm_tc.cuh"
#include "ptx_cluster.cuh"

namespace mdb {

struct GemmParams3 {
  GemmParams p;
  int m_tiles;      // real 128-row M tiles
  int m_groups;     // ceil(m_tiles / CTAS): tiles walked by one cluster
  int n_tiles;
  int out_cols;     // valid output columns (n_out, or n_out / 2 with GEGLU)
  int use_res_tma;  // residual boxes arrive through tmRes
  // LayerNorm folded into this GEMM (consumer side)
  const float* ln_stats;   // [pixels][ln_parts][2] partial (sum, sum sq) of the A rows, or nullptr
  int ln_parts;
  float ln_inv_c, ln_eps;
  const float* ln_colsum;  // [n_out] sum_k W'[n, k]
  // row statistics of this GEMM's bf16 output (producer side)
  float* stats_out;        // [pixels][n_tiles * kEpiGroups][2] or nullptr
  // ceil(2^32 / d) for d = m_groups, tiles_w, tiles_h: x / d = umulhi(x, magic) for x * d < 2^32 (host check)
  uint32_t mg_magic, tw_magic, th_magic;
};

#ifndef MDB_EPI_GROUPS
#define MDB_EPI_GROUPS 2
#endif
#ifndef MDB_PAIR_AHEAD
#define MDB_PAIR_AHEAD 4  // staging boxes (and residual loads) prepared this many chunks ahead of the store cursor (A/B knob)
#endif

// ---- fp32 pairs in 64-bit registers: FFMA2 / FADD2 / FMUL2 (sm_100) halve the epilogue's instruction count
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long pk2u(uint32_t a, uint32_t b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long r, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// explicit shared-space accesses (the staging / constant buffers are carved from a re-aligned dynamic-smem pointer, for which
// the compiler would otherwise emit generic LD / ST); volatile keeps them ordered against the mbarrier / tcgen05 waits
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long bf2_to_f2(uint32_t r) { return pk2u(r << 16, r & 0xffff0000u); }
__device__ __forceinline__ uint32_t f2_to_bf2(unsigned long long v) {
  float a, b;
  upk2(v, a, b);
  return pack_bf16(a, b);
}
// gelu_erf (gemm_tc.cuh) on a pair: the polynomial and the products run as packed ops, rcp / ex2 per element.
__device__ __forceinline__ unsigned long long gelu_erf2(unsigned long long x2) {
  float x0, x1;
  upk2(x2, x0, x1);
  const unsigned long long ax2 = mul2(pk2(fabsf(x0), fabsf(x1)), pk2(0.70710678118654752440f, 0.70710678118654752440f));
  float d0, d1, t0, t1, a0, a1, e0, e1;
  upk2(fma2(pk2(0.3275911f, 0.3275911f), ax2, pk2(1.0f, 1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const unsigned long long t2 = pk2(t0, t1);
  // -(a1 t + a2 t^2 + ... + a5 t^5): coefficients negated so that erf = 1 + npoly * e
  unsigned long long np = fma2(t2, pk2(-1.061405429f, -1.061405429f), pk2(1.453152027f, 1.453152027f));
  np = fma2(np, t2, pk2(-1.421413741f, -1.421413741f));
  np = fma2(np, t2, pk2(0.284496736f, 0.284496736f));
  np = fma2(np, t2, pk2(-0.254829592f, -0.254829592f));
  np = mul2(np, t2);
  upk2(mul2(mul2(ax2, ax2), pk2(-1.4426950408889634f, -1.4426950408889634f)), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  float r0, r1;
  upk2(fma2(np, pk2(e0, e1), pk2(1.0f, 1.0f)), r0, r1);  // erf(|x| / sqrt 2)
  const unsigned long long hx2 = mul2(x2, pk2(0.5f, 0.5f));
  return fma2(hx2, pk2(copysignf(r0, x0), copysignf(r1, x1)), hx2);
}

// compile-time epilogue options of one instantiation of the tile loop (selected once per kernel by a warp-uniform switch)
template <bool LN_, bool RES_, bool STATS_, bool ROWB_, bool SCALED_, bool GEGLU_>
struct EpiFlags {
  static constexpr bool LN = LN_, RES = RES_, STATS = STATS_, ROWB = ROWB_, SCALED = SCALED_, GEGLU = GEGLU_;
};

template <int BLOCK_N, int CTAS>
struct PairCfg {
  static constexpr int kBRows = BLOCK_N / CTAS;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kNBuf = 7;               // 128-row x 32-column (64 B) staging boxes: residual in, result out
  static constexpr int kAhead = MDB_PAIR_AHEAD;              // how many chunks ahead of the store cursor buffers are prepared
  static constexpr int kBufBytes = 128 * 64;
  static constexpr int kEpiGroups = MDB_EPI_GROUPS;  // groups of four epilogue warps (one warp per TMEM lane quarter)
  static constexpr int kEpiWarps = 4 * kEpiGroups;
  static constexpr int kConstStage = 3 * 256 * 4;      // [bias | per-image shift | LayerNorm column sums] x 256 fp32 of one tile
  static constexpr int kConstBytes = 2 * kConstStage;  // double-buffered like the accumulators
  static constexpr int kBarBytes = 1024;
  static constexpr int kBudget = 232448 - 1024 - kBarBytes - kConstBytes - kNBuf * kBufBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccStride = (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128) ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kThreads = 96 + 32 * kEpiWarps;
  static constexpr int kSmemBytes = kStages * kStageBytes + kNBuf * kBufBytes + kConstBytes + kBarBytes + 1024;
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kBBytes % 1024 == 0, "B stage must keep the 1024-byte swizzle alignment");
};

// byte offset of 16-byte chunk j (0..3) of row r inside a 64-byte-row staging box written/read by TMA with SWIZZLE_64B
__device__ __forceinline__ uint32_t stage_off(int r, int j) { return static_cast<uint32_t>(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

template <int BLOCK_N, int CTAS>
__global__ void __launch_bounds__(PairCfg<BLOCK_N, CTAS>::kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmRes, const GemmParams3 pp) {
  using Cfg = PairCfg<BLOCK_N, CTAS>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int NBUF = Cfg::kNBuf;
  constexpr bool PAIR = CTAS == 2;
  const GemmParams& p = pp.p;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * kABytes;
  uint8_t* smO = smem + STAGES * Cfg::kStageBytes;
  float* smC = reinterpret_cast<float*>(smO + NBUF * Cfg::kBufBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smC) + Cfg::kConstBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* res_full = acc_empty + 2;       // [NBUF]  staging box prepared (residual landed / buffer free)
  uint64_t* out_ready = res_full + NBUF;    // [NBUF]  staging box holds the finished bf16 chunk
  uint64_t* const_full = out_ready + NBUF;  // [2]  column constants of a tile landed (bulk copies by the manager)
  uint64_t* const_empty = const_full + 2;   // [2]  every epilogue warp of this CTA is done with them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int cluster_id = static_cast<int>(blockIdx.x) / CTAS;
  const int n_clusters = static_cast<int>(gridDim.x) / CTAS;
  const int cb_total = p.cblocks0 + p.cblocks1;
  const int nkb = p.taps_h * p.taps_w * cb_total;
  const int total = pp.m_groups * pp.n_tiles;
  const bool geglu = p.epi_mode == EPI_GEGLU;
  constexpr int CH_LIN = BLOCK_N / 32;
  constexpr int CH_GLU = BLOCK_N / 64;
  const int out_per_tile = geglu ? BLOCK_N / 2 : BLOCK_N;
  const int ch_tile = geglu ? CH_GLU : CH_LIN;

  // optional per-CTA phase stamps (debug / tools/bench_gemm.py --trace): 64 int64 slots per CTA, globaltimer ns; slots 16.. hold
  // the chunk-level stamps of the first epilogue warp over its first two tiles
  long long* trace = (p.trace != nullptr && blockIdx.x < 160) ? p.trace + blockIdx.x * 64 : nullptr;  // 64 slots per CTA
#define MDB_TRACE3(slot)                                                    \
  do {                                                                      \
    if (trace) {                                                            \
      unsigned long long t__;                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));               \
      trace[slot] = static_cast<long long>(t__);                            \
    }                                                                       \
  } while (0)
  if (threadIdx.x == 0) MDB_TRACE3(0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmOut);
    prefetch_tmap(&tmRes);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], CTAS * Cfg::kEpiWarps);  // one arrive per epilogue warp of every CTA of the pair
      mbar_init(&const_full[i], 1);
      mbar_init(&const_empty[i], Cfg::kEpiWarps);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&out_ready[i], 4);  // the four lane-quarter warps that own a chunk
    }
    fence_barrier_init();
    MDB_TRACE3(11);
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
    if (lane == 0) MDB_TRACE3(12);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) MDB_TRACE3(13);
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the predecessor's tail; from here on its results are complete and visible.  The
  // dependent grid may be scheduled as soon as every CTA of this one holds its shared memory and TMEM.
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) MDB_TRACE3(1);

  if (warp == 0) {
    // =========================== TMA producer (one lane) ===========================
    if (elect_one()) {
      const uint32_t tx_cta = static_cast<uint32_t>(p.bn * p.bh * p.bw) * (kBlockK * 2) + Cfg::kBBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total; t += n_clusters) {
        const int mg = t % pp.m_groups;
        const int nt = t / pp.m_groups;
        const int mt = mg * CTAS + static_cast<int>(rank);  // >= m_tiles for the phantom half of an odd pair: all OOB, zero-filled
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int img0 = tn * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
        const int nrow0 = nt * BLOCK_N + static_cast<int>(rank) * Cfg::kBRows;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const int tap = kb / cb_total;
          const int cb = kb - tap * cb_total;
          const int r = tap / p.taps_w, s = tap - r * p.taps_w;
          const int wc = w0 * p.stride + s - p.pad_w;
          const int hc = h0 * p.stride + r - p.pad_h;
          const CUtensorMap* am = (cb < p.cblocks0) ? &tmA0 : &tmA1;
          const int c0 = ((cb < p.cblocks0) ? cb : cb - p.cblocks0) * kBlockK;
          if constexpr (PAIR) {
            // both CTAs' bytes complete on the leader's barrier; only the leader arms it (for both)
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * tx_cta);
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_4d_pair(am, lead_bar, smA + stage * kABytes, c0, wc, hc, img0);
            tma_load_2d_pair(&tmB, lead_bar, smB + stage * Cfg::kBBytes, kb * kBlockK, nrow0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], tx_cta);
            tma_load_4d(am, &full_bar[stage], smA + stage * kABytes, c0, wc, hc, img0);
            tma_load_2d(&tmB, &full_bar[stage], smB + stage * Cfg::kBBytes, kb * kBlockK, nrow0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      MDB_TRACE3(2);
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA of the pair) ===========================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM * CTAS, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < total; t += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::kAccStride;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0 && it == 0 && i == 0) MDB_TRACE3(14);
          if (elect_one()) {
            const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smA + stage * kABytes));
            const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smB + stage * Cfg::kBBytes));
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              if constexpr (PAIR) umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
              else umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
            if constexpr (PAIR) {
              umma_commit_pair(&empty_bar[stage], 3);
              if (i == nkb - 1) umma_commit_pair(&acc_full[as], 3);
            } else {
              umma_commit(&empty_bar[stage]);
              if (i == nkb - 1) umma_commit(&acc_full[as]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (lane == 0 && it == 0) MDB_TRACE3(7);
      }
      if (lane == 0) MDB_TRACE3(3);
    }
  } else if (warp == 2) {
    // =========================== staging-buffer manager: residual loads + output stores (one lane) ==============
    if (lane == 0) {
      const uint32_t box_bytes = static_cast<uint32_t>(p.bn * p.bh * p.bw) * 64;
      // two cursors over the same (tile, chunk) sequence: `pre` prepares buffers kAhead chunks ahead of `sto`.  The tile
      // coordinates of a cursor are recomputed only when it moves to another tile (this single thread must keep up with
      // one TMA store + one TMA load per 32-column chunk).
      struct Cursor {
        int t, c, k, nch, col0, w0, h0, img0;
        bool real;
      };
      auto enter_tile = [&](Cursor& cu) {
        if (cu.t >= total) return;
        const int mg = cu.t % pp.m_groups;
        const int nt = cu.t / pp.m_groups;
        const int mt = mg * CTAS + static_cast<int>(rank);
        const int left = (pp.out_cols - nt * out_per_tile) / 32;
        cu.nch = left < ch_tile ? left : ch_tile;
        cu.col0 = nt * out_per_tile;
        cu.w0 = (mt % p.tiles_w) * p.bw;
        cu.h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
        cu.img0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
        cu.real = mt < pp.m_tiles;
        cu.c = 0;
      };
      auto advance = [&](Cursor& cu) {
        ++cu.k;
        if (++cu.c == cu.nch) {
          cu.t += n_clusters;
          enter_tile(cu);
        }
      };
      // Column constants (bias, per-image shift, folded-LayerNorm column sums) of tile ordinal `itn` -> smC[itn & 1] by bulk
      // copies, a whole tile before the epilogue needs them: fetched by the epilogue warps themselves (even one chunk ahead)
      // they cost every chunk ~0.5 us of exposed L2 latency (chunk-level trace).
      // `block` = the epilogue is about to need them (never in steady state); otherwise the load is simply retried at the
      // manager's next chunk, so that a slow warp at a tile boundary does not hold up the stores / box preparation of the rest.
      int const_it = 0;  // next tile ordinal whose constants have to be fetched
      auto load_tile_consts = [&](bool block) {
        const int itn = const_it;
        const int t = cluster_id + itn * n_clusters;
        if (t >= total) return;
        const int as = itn & 1;
        if (itn >= 2) {  // the epilogue must be done with tile itn - 2
          if (block) mbar_wait(&const_empty[as], ((itn >> 1) - 1) & 1);
          else if (!mbar_test(&const_empty[as], ((itn >> 1) - 1) & 1)) return;
        }
        ++const_it;
        const int nt = t / pp.m_groups;
        const int mt = (t - nt * pp.m_groups) * CTAS + static_cast<int>(rank);
        const int img = min((mt / (p.tiles_w * p.tiles_h)) * p.bn, p.n_img - 1);
        const int n0 = nt * BLOCK_N;
        const int cnt = min(BLOCK_N, p.n_out - n0);
        const uint32_t bytes = static_cast<uint32_t>(cnt) * 4;
        const uint32_t dst = smem_u32(smC) + as * Cfg::kConstStage;
        const uint32_t bar = smem_u32(&const_full[as]);
        const bool has_ln = pp.ln_stats != nullptr;
        const uint32_t tx = (p.bias ? bytes : 0u) + (p.rowbias ? bytes : 0u) + (has_ln ? bytes : 0u);
        if (tx == 0) {
          mbar_arrive(&const_full[as]);
          return;
        }
        mbar_arrive_expect_tx(&const_full[as], tx);
        auto copy = [&](uint32_t d, const float* src) {
          asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
                       "r"(bytes), "r"(bar)
                       : "memory");
        };
        if (p.bias) copy(dst, p.bias + n0);
        if (p.rowbias) copy(dst + 1024, p.rowbias + static_cast<long long>(img) * p.rowbias_ld + n0);
        if (has_ln) copy(dst + 2048, pp.ln_colsum + n0);
      };
      load_tile_consts(true);
      load_tile_consts(true);
      int sto_it = 0;  // tile ordinal of the store cursor
      Cursor pre{cluster_id, 0, 0, 0, 0, 0, 0, 0, false}, sto{cluster_id, 0, 0, 0, 0, 0, 0, 0, false};
      enter_tile(pre);
      enter_tile(sto);
      auto prepare = [&]() {  // make staging box pre.k ready for the epilogue
        const int buf = pre.k % NBUF;
        if (pre.k >= NBUF) bulk_wait_group_read<NBUF - Cfg::kAhead - 1>();  // the store that last used `buf` has read it
        if (pp.use_res_tma && pre.real) {
          mbar_arrive_expect_tx(&res_full[buf], box_bytes);
          tma_load_4d(&tmRes, &res_full[buf], smO + buf * Cfg::kBufBytes, pre.col0 + pre.c * 32, pre.w0, pre.h0, pre.img0);
        } else {
          mbar_arrive(&res_full[buf]);
        }
        advance(pre);
      };
      for (int i = 0; i < Cfg::kAhead && pre.t < total; ++i) prepare();
      while (sto.t < total) {
        if (pre.t < total) prepare();
        load_tile_consts(const_it <= sto_it);  // the chunk awaited below cannot be produced without its tile's constants
        const int buf = sto.k % NBUF;
        mbar_wait(&out_ready[buf], (sto.k / NBUF) & 1);
        if (sto.real) tma_store_4d(&tmOut, smO + buf * Cfg::kBufBytes, sto.col0 + sto.c * 32, sto.w0, sto.h0, sto.img0);
        bulk_commit_group();  // one group per chunk (empty for the phantom half of an odd pair) keeps the wait counts exact
        const int t_before = sto.t;
        advance(sto);
        if (sto.t != t_before) ++sto_it;
      }
      MDB_TRACE3(4);
      bulk_wait_group_all();
      MDB_TRACE3(5);
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 3 .. 3 + 4 G) ===========================
    // G groups of four warps (one per TMEM lane quarter); the 32-column chunks go round-robin to the groups.  All
    // arithmetic on fp32 PAIRS (FFMA2 / FADD2 / FMUL2): out = A_row * acc + (B_row * colsum_n + bias_n) [+ residual] with
    // A_row = rstd * scale, B_row = -mean * rstd * scale (folded LayerNorm) or A_row = scale, B_row = 0.
    //
    // The tile loop is instantiated once per combination of epilogue options (EpiFlags) and selected by one warp-uniform
    // switch: with the options as run-time flags every 32-column chunk issued ~400 instructions, 167 of them predicated off
    // (ncu source page of the K = 320 token GEMMs, profiles/ncu_gemm_pair_r2_s14.summary.txt), and a chunk took a warp
    // ~0.9 us -- three chunks per tile against 0.8 us of MMA.
    constexpr int G = Cfg::kEpiGroups;
    const int q = warp & 3;
    const int eg = (warp - 3) >> 2;
    const int row = q * 32 + lane;
    const int box_hw = p.bh * p.bw;
    const int li = row / box_hw;
    const int rem = row - li * box_hw;
    const int lh = rem / p.bw;
    const int lw = rem - lh * p.bw;
    const uint32_t my_row = smem_u32(smO) + row * 64;  // this thread's 64-byte row inside a staging box (shared-space address)
    const int swz = (row >> 1) & 3;                  // SWIZZLE_64B: 16-byte chunk j sits at j ^ swz
    const float scale = p.out_scale;
    const bool has_bias = p.bias != nullptr;
    // Where this thread's accumulator row lands for tile t: one decomposition per tile (multiply-high by host-computed
    // reciprocals, no integer division), carried from the prefetch of the previous iteration.
    struct TileRow {
      int nt;        // N tile
      int img_tile;  // image of the tile (per-image shift: one image per tile, host check)
      int pix;       // output pixel of this row
      bool ok;       // row is a real output pixel
    };
    // x / d (x * d < 2^32, host check): ceil(2^32 / d) does not fit 32 bits for d == 1, which is taken apart (uniform select)
    auto fdiv = [](int x, uint32_t magic, int d) { return d == 1 ? x : static_cast<int>(__umulhi(static_cast<uint32_t>(x), magic)); };
    auto locate = [&](int t) {
      TileRow r;
      r.nt = fdiv(t, pp.mg_magic, pp.m_groups);
      const int mt = (t - r.nt * pp.m_groups) * CTAS + static_cast<int>(rank);
      const int a = fdiv(mt, pp.tw_magic, p.tiles_w);
      const int tw = mt - a * p.tiles_w;
      const int tn = fdiv(a, pp.th_magic, p.tiles_h);
      const int th = a - tn * p.tiles_h;
      const int img = tn * p.bn + li, oh = th * p.bh + lh, ow = tw * p.bw + lw;
      r.ok = (t < total) && (mt < pp.m_tiles) && (li < p.bn) && (img < p.n_img) && (oh < p.h_out) && (ow < p.w_out);
      r.pix = (img * p.h_out + oh) * p.w_out + ow;
      r.img_tile = min(tn * p.bn, p.n_img - 1);
      return r;
    };
    // first chunk of tile (t, gk0 = running chunk number at its start) owned by this group, or -1
    auto first_owned = [&](int gk0, int nch) {
      const int c = ((eg - gk0) % G + G) % G;
      return c < nch ? c : -1;
    };
    auto chunks_of = [&](const TileRow& r) {
      const int left = (pp.out_cols - r.nt * out_per_tile) / 32;
      return left < ch_tile ? left : ch_tile;
    };
    constexpr int PF = (G <= 2) ? 8 : 4;  // LayerNorm row-statistics partials fetched one tile ahead (the rest, if any, at use)
    constexpr int HALFN = BLOCK_N / 2;

    auto epi_loop = [&](auto flags) {
      using F = decltype(flags);
      constexpr bool LN = F::LN, RES = F::RES, STATS = F::STATS, ROWB = F::ROWB, SCALED = F::SCALED, GLU = F::GEGLU;
      float2 pf[PF];
      auto fetch_stats = [&](const TileRow& r) {
        if constexpr (LN) {
#pragma unroll
          for (int j = 0; j < PF; ++j) pf[j] = make_float2(0.f, 0.f);
          if (r.ok) {
            const float2* sp = reinterpret_cast<const float2*>(pp.ln_stats) + static_cast<long long>(r.pix) * pp.ln_parts;
#pragma unroll
            for (int j = 0; j < PF; ++j)
              if (j < pp.ln_parts) pf[j] = __ldg(sp + j);
          }
        }
      };
      // Column constants of a tile sit in smC[as] (bulk-copied by the manager one tile ahead): broadcast LDS.128 reads.
      // t0/t1 = (bias + shift) * scale [+ B_row * colsum] for columns col .. col + 3 of the tile.
      auto col_consts = [&](uint32_t cst, int col, unsigned long long B2, unsigned long long& t0, unsigned long long& t1) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias) b = lds_f4(cst + col * 4);
        if constexpr (ROWB) {
          const float4 r = lds_f4(cst + 1024 + col * 4);
          b.x += r.x, b.y += r.y, b.z += r.z, b.w += r.w;
        }
        t0 = pk2(b.x, b.y), t1 = pk2(b.z, b.w);
        if constexpr (SCALED) t0 = mul2(t0, pk2(scale, scale)), t1 = mul2(t1, pk2(scale, scale));
        if constexpr (LN) {
          const float4 cs = lds_f4(cst + 2048 + col * 4);
          t0 = fma2(B2, pk2(cs.x, cs.y), t0), t1 = fma2(B2, pk2(cs.z, cs.w), t1);
        }
      };
      TileRow cur = locate(cluster_id);
      fetch_stats(cur);
      int it = 0;
      int gk = 0;  // running chunk number (same sequence as the staging-buffer manager's cursors)
      for (int t = cluster_id; t < total; t += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        // this tile's row scalars from the prefetched statistics
        float rowA = scale, rowB = 0.f;
        if constexpr (LN) {
          float s = 0.f, ss = 0.f;
#pragma unroll
          for (int j = 0; j < PF; ++j) s += pf[j].x, ss += pf[j].y;
          if (cur.ok) {
            const float2* sp = reinterpret_cast<const float2*>(pp.ln_stats) + static_cast<long long>(cur.pix) * pp.ln_parts;
            for (int j = PF; j < pp.ln_parts; ++j) {
              const float2 v = __ldg(sp + j);
              s += v.x, ss += v.y;
            }
          }
          const float mean = s * pp.ln_inv_c;
          const float rstd = rsqrtf(fmaxf(ss * pp.ln_inv_c - mean * mean, 0.f) + pp.ln_eps);
          rowA = rstd * scale;
          rowB = -mean * rowA;
        }
        // next tile's position / statistics: loads in flight while this tile is processed
        const TileRow nxt = locate(t + n_clusters);
        fetch_stats(nxt);
        const int nch = chunks_of(cur);
        const int gk_tile = gk;  // running chunk number of this tile's chunk 0
        const unsigned long long A2 = pk2(rowA, rowA), B2 = pk2(rowB, rowB);
        unsigned long long st_s2 = pk2(0.f, 0.f), st_ss2 = pk2(0.f, 0.f);
        const uint32_t cst = smem_u32(smC) + as * Cfg::kConstStage;
        mbar_wait(&const_full[as], aphase);  // landed a tile ago
        mbar_wait(&acc_full[as], aphase);
        tc_fence_after();
        const uint32_t lane_addr = tmem_base + as * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
#ifdef MDB_TRACE_CHUNKS  // chunk-level stamps of the first epilogue warp (debug builds only: tools/build_variant.sh -DMDB_TRACE_CHUNKS)
        int tslot = (warp == 3 && lane == 0 && it < 2) ? 16 + it * 20 : -1;  // [acc_full | (box ready, acc read, stored, arrived) x <= 4]
        if (tslot >= 0) {
          MDB_TRACE3(tslot);
          ++tslot;
        }
        const int tslot_end = tslot + 16;
#define MDB_TRACE_CHUNK() do { if (tslot >= 0 && tslot < tslot_end) { MDB_TRACE3(tslot); ++tslot; } } while (0)
#else
#define MDB_TRACE_CHUNK() do { } while (0)
#endif

        // the chunks go round-robin over the running chunk number to the groups: balanced even when a tile has 5 chunks
        for (int c = first_owned(gk_tile, nch); c >= 0 && c < nch; c += G) {
          gk = gk_tile + c;
          const int buf = gk % NBUF;
          const uint32_t srow = my_row + buf * Cfg::kBufBytes;
          if constexpr (!GLU) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + c * 32, v);  // in flight while the constants are combined and the staging box is awaited
            unsigned long long pre[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) col_consts(cst, c * 32 + 4 * j, B2, pre[2 * j], pre[2 * j + 1]);
            mbar_wait(&res_full[buf], (gk / NBUF) & 1);
            MDB_TRACE_CHUNK();
            uint4 r[4];
            if constexpr (RES) {
#pragma unroll
              for (int j = 0; j < 4; ++j) r[j] = lds_u4(srow + ((j ^ swz) << 4));
            }
            tmem_ld_wait();
            MDB_TRACE_CHUNK();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              unsigned long long o0 = fma2(A2, pk2u(v[8 * j], v[8 * j + 1]), pre[4 * j]);
              unsigned long long o1 = fma2(A2, pk2u(v[8 * j + 2], v[8 * j + 3]), pre[4 * j + 1]);
              unsigned long long o2 = fma2(A2, pk2u(v[8 * j + 4], v[8 * j + 5]), pre[4 * j + 2]);
              unsigned long long o3 = fma2(A2, pk2u(v[8 * j + 6], v[8 * j + 7]), pre[4 * j + 3]);
              if constexpr (RES) {
                o0 = add2(o0, bf2_to_f2(r[j].x)), o1 = add2(o1, bf2_to_f2(r[j].y));
                o2 = add2(o2, bf2_to_f2(r[j].z)), o3 = add2(o3, bf2_to_f2(r[j].w));
              }
              if constexpr (STATS) {
                st_s2 = add2(add2(st_s2, add2(o0, o1)), add2(o2, o3));
                st_ss2 = fma2(o0, o0, fma2(o1, o1, fma2(o2, o2, fma2(o3, o3, st_ss2))));
              }
              sts_u4(srow + ((j ^ swz) << 4), make_uint4(f2_to_bf2(o0), f2_to_bf2(o1), f2_to_bf2(o2), f2_to_bf2(o3)));
            }
            MDB_TRACE_CHUNK();
          } else {
            constexpr int HALF = BLOCK_N / 2;
            mbar_wait(&res_full[buf], (gk / NBUF) & 1);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[16], g[16];
              tmem_ld_32x16(lane_addr + c * 32 + hh * 16, v);
              tmem_ld_32x16(lane_addr + HALF + c * 32 + hh * 16, g);
              unsigned long long tv[8], tg[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                col_consts(cst, c * 32 + hh * 16 + 4 * j, B2, tv[2 * j], tv[2 * j + 1]);
                col_consts(cst, HALFN + c * 32 + hh * 16 + 4 * j, B2, tg[2 * j], tg[2 * j + 1]);
              }
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                unsigned long long o[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int k = 8 * j + 4 * e;
                  const unsigned long long a0 = fma2(A2, pk2u(v[k], v[k + 1]), tv[k / 2]), a1 = fma2(A2, pk2u(v[k + 2], v[k + 3]), tv[k / 2 + 1]);
                  const unsigned long long g0 = fma2(A2, pk2u(g[k], g[k + 1]), tg[k / 2]), g1 = fma2(A2, pk2u(g[k + 2], g[k + 3]), tg[k / 2 + 1]);
                  o[2 * e] = mul2(a0, gelu_erf2(g0));
                  o[2 * e + 1] = mul2(a1, gelu_erf2(g1));
                }
                sts_u4(srow + (((hh * 2 + j) ^ swz) << 4), make_uint4(f2_to_bf2(o[0]), f2_to_bf2(o[1]), f2_to_bf2(o[2]), f2_to_bf2(o[3])));
              }
            }
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) mbar_arrive(&out_ready[buf]);
          if constexpr (!GLU) MDB_TRACE_CHUNK();
        }
#undef MDB_TRACE_CHUNK
        gk = gk_tile + nch;
        if constexpr (STATS) {
          if (cur.ok) {
            float s0, s1, q0, q1;
            upk2(st_s2, s0, s1);
            upk2(st_ss2, q0, q1);
            // one slot per (N tile, epilogue group): [pix][n_tiles][G]
            reinterpret_cast<float2*>(pp.stats_out)[(static_cast<long long>(cur.pix) * pp.n_tiles + cur.nt) * G + eg] =
                make_float2(s0 + s1, q0 + q1);
          }
        }
        // release this accumulator stage to the MMA warp of the leader CTA, and the constants' buffer to the manager
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));
          else mbar_arrive(&acc_empty[as]);
          mbar_arrive(&const_empty[as]);
        }
        cur = nxt;
        if (warp == 3 && lane == 0 && it == 0) MDB_TRACE3(8);
      }
    };
    // one warp-uniform selection of the loop's instantiation (GEGLU tiles take no residual / per-image shift and emit no
    // statistics: host check)
    const int code = (pp.ln_stats != nullptr ? 1 : 0) | (pp.use_res_tma ? 2 : 0) | (pp.stats_out != nullptr ? 4 : 0) |
                     (p.rowbias != nullptr ? 8 : 0) | (scale != 1.f ? 16 : 0);
    if (geglu) {
      switch (code & 17) {
        case 0: epi_loop(EpiFlags<false, false, false, false, false, true>{}); break;
        case 1: epi_loop(EpiFlags<true, false, false, false, false, true>{}); break;
        case 16: epi_loop(EpiFlags<false, false, false, false, true, true>{}); break;
        default: epi_loop(EpiFlags<true, false, false, false, true, true>{}); break;
      }
    } else {
      switch (code) {
#define MDB_EPI_CASE(n) \
  case n: epi_loop(EpiFlags<((n) & 1) != 0, ((n) & 2) != 0, ((n) & 4) != 0, ((n) & 8) != 0, ((n) & 16) != 0, false>{}); break;
        MDB_EPI_CASE(0) MDB_EPI_CASE(1) MDB_EPI_CASE(2) MDB_EPI_CASE(3) MDB_EPI_CASE(4) MDB_EPI_CASE(5) MDB_EPI_CASE(6) MDB_EPI_CASE(7)
        MDB_EPI_CASE(8) MDB_EPI_CASE(9) MDB_EPI_CASE(10) MDB_EPI_CASE(11) MDB_EPI_CASE(12) MDB_EPI_CASE(13) MDB_EPI_CASE(14)
        MDB_EPI_CASE(15) MDB_EPI_CASE(16) MDB_EPI_CASE(17) MDB_EPI_CASE(18) MDB_EPI_CASE(19) MDB_EPI_CASE(20) MDB_EPI_CASE(21)
        MDB_EPI_CASE(22) MDB_EPI_CASE(23) MDB_EPI_CASE(24) MDB_EPI_CASE(25) MDB_EPI_CASE(26) MDB_EPI_CASE(27) MDB_EPI_CASE(28)
        MDB_EPI_CASE(29) MDB_EPI_CASE(30)
        default: epi_loop(EpiFlags<true, true, true, true, true, false>{}); break;
#undef MDB_EPI_CASE
      }
    }
    if (warp == 3 && lane == 0) MDB_TRACE3(6);
  }

  __syncwarp();
  if (threadIdx.x == 0) MDB_TRACE3(9);
  // both CTAs are done with each other's barriers and with TMEM: liveness only, no memory ordering needed
  if constexpr (PAIR) cluster_sync_relaxed(); else __syncthreads();
  if (threadIdx.x == 0) MDB_TRACE3(10);
#undef MDB_TRACE3
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mdb
