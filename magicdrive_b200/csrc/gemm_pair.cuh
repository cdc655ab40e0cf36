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
  float* stats_out;        // [pixels][2 * n_tiles][2] or nullptr
};

template <int BLOCK_N, int CTAS>
struct PairCfg {
  static constexpr int kBRows = BLOCK_N / CTAS;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kNBuf = 6;               // 128-row x 32-column (64 B) staging boxes: residual in, result out
  static constexpr int kAhead = 3;              // how many chunks ahead of the store cursor buffers are prepared
  static constexpr int kBufBytes = 128 * 64;
  static constexpr int kConstBytes = 2 * 2 * 256 * 4;  // [acc stage][bias+shift | colsum][BLOCK_N] fp32
  static constexpr int kBarBytes = 1024;
  static constexpr int kBudget = 232448 - 1024 - kBarBytes - kConstBytes - kNBuf * kBufBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccStride = (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128) ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kEpiWarps = 8;
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_ready + NBUF);

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

  // optional per-CTA phase stamps (debug / tools/bench_gemm.py --trace): 16 int64 slots per CTA, globaltimer ns
  long long* trace = (p.trace != nullptr && blockIdx.x < 160) ? p.trace + blockIdx.x * 16 : nullptr;
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
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&out_ready[i], 4);  // the four lane-quarter warps that own a chunk
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
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
      // two cursors over the same (tile, chunk) sequence: `pre` prepares buffers kAhead chunks ahead of `sto`
      int pre_t = cluster_id, pre_c = 0, pre_k = 0;
      int sto_t = cluster_id, sto_c = 0, sto_k = 0;
      auto chunks_of = [&](int t) {
        const int nt = t / pp.m_groups;
        const int left = (pp.out_cols - nt * out_per_tile) / 32;
        return left < ch_tile ? left : ch_tile;
      };
      auto coords = [&](int t, int c, int& col0, int& w0, int& h0, int& img0) {
        const int mg = t % pp.m_groups;
        const int nt = t / pp.m_groups;
        const int mt = mg * CTAS + static_cast<int>(rank);
        col0 = nt * out_per_tile + c * 32;
        w0 = (mt % p.tiles_w) * p.bw;
        h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
        img0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
        return mt < pp.m_tiles;
      };
      auto prepare = [&]() {  // make staging box pre_k ready for the epilogue
        const int buf = pre_k % NBUF;
        if (pre_k >= NBUF) bulk_wait_group_read<NBUF - Cfg::kAhead - 1>();  // the store that last used `buf` has read it
        int col0, w0, h0, img0;
        const bool real = coords(pre_t, pre_c, col0, w0, h0, img0);
        if (pp.use_res_tma && real) {
          mbar_arrive_expect_tx(&res_full[buf], box_bytes);
          tma_load_4d(&tmRes, &res_full[buf], smO + buf * Cfg::kBufBytes, col0, w0, h0, img0);
        } else {
          mbar_arrive(&res_full[buf]);
        }
        ++pre_k;
        if (++pre_c == chunks_of(pre_t)) pre_c = 0, pre_t += n_clusters;
      };
      for (int i = 0; i < Cfg::kAhead && pre_t < total; ++i) prepare();
      while (sto_t < total) {
        if (pre_t < total) prepare();
        const int buf = sto_k % NBUF;
        mbar_wait(&out_ready[buf], (sto_k / NBUF) & 1);
        int col0, w0, h0, img0;
        if (coords(sto_t, sto_c, col0, w0, h0, img0))
          tma_store_4d(&tmOut, smO + buf * Cfg::kBufBytes, col0, w0, h0, img0);
        bulk_commit_group();  // one group per chunk (empty for the phantom half of an odd pair) keeps the wait counts exact
        ++sto_k;
        if (++sto_c == chunks_of(sto_t)) sto_c = 0, sto_t += n_clusters;
      }
      MDB_TRACE3(4);
      bulk_wait_group_all();
      MDB_TRACE3(5);
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 3..10) ===========================
    const int q = warp & 3;
    const int eg = (warp - 3) >> 2;  // which of the two interleaved chunk sets this warp owns
    const int etid = threadIdx.x - 96;
    const int row = q * 32 + lane;
    const int box_hw = p.bh * p.bw;
    const int li = row / box_hw;
    const int rem = row - li * box_hw;
    const int lh = rem / p.bw;
    const int lw = rem - lh * p.bw;
    const uint32_t smO_u32 = smem_u32(smO);
    const bool has_ln = pp.ln_stats != nullptr;
    int it = 0;
    int gk = 0;  // running chunk number (same sequence as the staging-buffer manager's cursors)
    for (int t = cluster_id; t < total; t += n_clusters, ++it) {
      const int mg = t % pp.m_groups;
      const int nt = t / pp.m_groups;
      const int mt = mg * CTAS + static_cast<int>(rank);
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int img = tn * p.bn + li, oh = th * p.bh + lh, ow = tw * p.bw + lw;
      const int n0 = nt * BLOCK_N;
      const bool row_ok = (mt < pp.m_tiles) && (li < p.bn) && (img < p.n_img) && (oh < p.h_out) && (ow < p.w_out);
      const long long pix = (static_cast<long long>(img) * p.h_out + oh) * p.w_out + ow;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      // ---- per-tile constants: bias + per-image shift (and the folded LayerNorm's column sums) -> shared memory
      float* cst = smC + as * 512;
      {
        const int img_tile = min(tn * p.bn, p.n_img - 1);  // one image per tile whenever a per-image shift is used (host check)
        for (int j = etid; j < BLOCK_N; j += 32 * Cfg::kEpiWarps) {
          const int n = n0 + j;
          float b = 0.f, cs = 0.f;
          if (n < p.n_out) {
            if (p.bias) b = __ldg(p.bias + n);
            if (p.rowbias) b += __ldg(p.rowbias + static_cast<long long>(img_tile) * p.rowbias_ld + n);
            if (has_ln) cs = __ldg(pp.ln_colsum + n);
          }
          cst[j] = b;
          cst[256 + j] = cs;
        }
      }
      float mean = 0.f, rstd = 1.f;
      if (has_ln && row_ok) {
        const float2* sp = reinterpret_cast<const float2*>(pp.ln_stats) + pix * pp.ln_parts;
        float s = 0.f, ss = 0.f;
        for (int j = 0; j < pp.ln_parts; ++j) {
          const float2 v = __ldg(sp + j);
          s += v.x;
          ss += v.y;
        }
        mean = s * pp.ln_inv_c;
        const float var = fmaxf(ss * pp.ln_inv_c - mean * mean, 0.f);
        rstd = rsqrtf(var + pp.ln_eps);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * Cfg::kEpiWarps) : "memory");  // constants visible to all epilogue warps
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + as * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      const int left = (pp.out_cols - nt * out_per_tile) / 32;
      const int nch = left < ch_tile ? left : ch_tile;
      float st_s = 0.f, st_ss = 0.f;

      for (int c = 0; c < nch; ++c, ++gk) {
        if ((c & 1) != eg) continue;
        const int buf = gk % NBUF;
        mbar_wait(&res_full[buf], (gk / NBUF) & 1);
        const uint32_t sbuf = smO_u32 + buf * Cfg::kBufBytes;
        if (!geglu) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + c * 32, v);
          tmem_ld_wait();
          const float* cb = cst + c * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
            if (has_ln) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = rstd * (f[e] - mean * cb[256 + 8 * j + e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = (f[e] + cb[8 * j + e]) * p.out_scale;
            if (pp.use_res_tma) {
              const uint4 r = ld_shared_v4(sbuf + stage_off(row, j));
              const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 rf = __bfloat1622float2(rh[e]);
                f[2 * e] += rf.x;
                f[2 * e + 1] += rf.y;
              }
            }
            uint4 o;
            o.x = pack_bf16(f[0], f[1]), o.y = pack_bf16(f[2], f[3]), o.z = pack_bf16(f[4], f[5]), o.w = pack_bf16(f[6], f[7]);
            if (pp.stats_out) {  // statistics of the values the consumer will actually read (bf16-rounded)
              const __nv_bfloat162* oh2 = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 of = __bfloat1622float2(oh2[e]);
                st_s += of.x + of.y;
                st_ss = fmaf(of.x, of.x, fmaf(of.y, of.y, st_ss));
              }
            }
            st_shared_v4(sbuf + stage_off(row, j), o);
          }
        } else {
          constexpr int HALF = BLOCK_N / 2;
          const float* cbv = cst + c * 32;
          const float* cbg = cst + HALF + c * 32;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[16], g[16];
            tmem_ld_32x16(lane_addr + c * 32 + hh * 16, v);
            tmem_ld_32x16(lane_addr + HALF + c * 32 + hh * 16, g);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float o8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int cc = hh * 16 + j * 8 + e;
                float a = __uint_as_float(v[j * 8 + e]);
                float gg = __uint_as_float(g[j * 8 + e]);
                if (has_ln) {
                  a = rstd * (a - mean * cbv[256 + cc]);
                  gg = rstd * (gg - mean * cbg[256 + cc]);
                }
                a += cbv[cc];
                gg += cbg[cc];
                o8[e] = a * gelu_erf(gg);
              }
              uint4 o;
              o.x = pack_bf16(o8[0], o8[1]), o.y = pack_bf16(o8[2], o8[3]), o.z = pack_bf16(o8[4], o8[5]), o.w = pack_bf16(o8[6], o8[7]);
              st_shared_v4(sbuf + stage_off(row, hh * 2 + j), o);
            }
          }
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_ready[buf]);
      }
      if (pp.stats_out && row_ok)
        reinterpret_cast<float2*>(pp.stats_out)[(pix * pp.n_tiles + nt) * 2 + eg] = make_float2(st_s, st_ss);
      // release this accumulator stage to the MMA warp of the leader CTA
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));
        else mbar_arrive(&acc_empty[as]);
      }
      if (warp == 3 && lane == 0 && it == 0) MDB_TRACE3(8);
    }
    if (warp == 3 && lane == 0) MDB_TRACE3(6);
  }

  __syncwarp();
  if (threadIdx.x == 0) MDB_TRACE3(9);
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) MDB_TRACE3(10);
#undef MDB_TRACE3
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mdb
