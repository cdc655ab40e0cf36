// Persistent variant of the tcgen05 GEMM / implicit-GEMM convolution (see gemm_tc.cuh for the operand scheme).
//
// One CTA per SM loops over output tiles (m fastest, so concurrently running CTAs share the same weight tile in
// L2).  The smem ring keeps flowing across tile boundaries and the fp32 accumulator is double-buffered in TMEM
// (2 x BLOCK_N columns), so the epilogue of tile i (tcgen05.ld -> bias/temb/residual -> global) overlaps the
// TMA + MMA main loop of tile i+1, and the per-CTA prologue (barrier init, TMEM alloc, descriptor prefetch) is
// paid once per launch instead of once per tile.
//   warp 0: TMA producer      warp 1: MMA issuer      warps 2..5: epilogue (TMEM lane quarter = warp & 3)
#pragma once
#include "gemm_tc.cuh"
#include "common_host.h"

namespace mdb {

struct GemmParams2 {
  GemmParams p;
  int m_tiles, n_tiles, splits;  // total tiles = m_tiles * n_tiles * splits
};

template <int BLOCK_N>
struct GemmCfg2 {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 160) ? 5 : 6);
  static constexpr int kAccStride = (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128) ? 128 : 256;  // TMEM columns per stage
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static constexpr int kEpiWarps = 8;   // two warps per TMEM lane quarter, alternating 32-column chunks
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(GemmCfg2<BLOCK_N>::kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const GemmParams2 pp) {
  using Cfg = GemmCfg2<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  const GemmParams& p = pp.p;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* trace = (p.trace != nullptr && blockIdx.x < 8) ? p.trace + blockIdx.x * 16 : nullptr;
#define MDB_TRACE(slot) do { if (trace) trace[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) MDB_TRACE(0);
  const int cb_total = p.cblocks0 + p.cblocks1;
  const int kb_total = p.taps_h * p.taps_w * cb_total;
  const int total_tiles = pp.m_tiles * pp.n_tiles * pp.splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], Cfg::kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) MDB_TRACE(1);
  // everything above (barrier init, TMEM alloc, descriptor prefetch) overlapped the predecessor's tail; from here on
  // we read its outputs / overwrite buffers it may still be reading
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      const uint32_t tx_bytes = static_cast<uint32_t>(p.bn * p.bh * p.bw) * (kBlockK * 2) + Cfg::kBBytes;
      int stage = 0;
      uint32_t phase = 0;
      MDB_TRACE(2);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t % pp.m_tiles;
        const int rest = t / pp.m_tiles;
        const int nt = rest % pp.n_tiles;
        const int z = rest / pp.n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int img0 = tn * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
        const int n0 = nt * BLOCK_N;
        const int kb_begin = z * p.kb_per_split;
        const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const int tap = kb / cb_total;
          const int cb = kb - tap * cb_total;
          const int r = tap / p.taps_w, s = tap - r * p.taps_w;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          const int wc = w0 * p.stride + s - p.pad_w;
          const int hc = h0 * p.stride + r - p.pad_h;
          if (cb < p.cblocks0)
            tma_load_4d(&tmA0, &full_bar[stage], smA + stage * kABytes, cb * kBlockK, wc, hc, img0);
          else
            tma_load_4d(&tmA1, &full_bar[stage], smA + stage * kABytes, (cb - p.cblocks0) * kBlockK, wc, hc, img0);
          tma_load_2d(&tmB, &full_bar[stage], smB + stage * Cfg::kBBytes, kb * kBlockK, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      MDB_TRACE(3);
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int z = (t / pp.m_tiles) / pp.n_tiles;
      const int kb_begin = z * p.kb_per_split;
      const int nkb = min(kb_total, kb_begin + p.kb_per_split) - kb_begin;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&acc_empty[as], aphase ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * Cfg::kAccStride;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0 && it == 0 && i == 0) MDB_TRACE(4);
        if (elect_one()) {
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smA + stage * kABytes));
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smB + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (i == nkb - 1) umma_commit(&acc_full[as]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (lane == 0 && it == 0) MDB_TRACE(5);
    }
    if (lane == 0) MDB_TRACE(6);
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    // A single warp per scheduler runs the long dependent epilogue stream at IPC ~0.25; two warps per TMEM lane
    // quarter (each taking every other 32-column chunk) double the issue rate of the drain.
    const int q = warp & 3;
    const int eg = (warp - 2) >> 2;  // 0 or 1: which interleaved half of the chunks
    const int row = q * 32 + lane;
    const int box_hw = p.bh * p.bw;
    const int li = row / box_hw;
    const int rem = row - li * box_hw;
    const int lh = rem / p.bw;
    const int lw = rem - lh * p.bw;
    const long long pixels_total = static_cast<long long>(p.n_img) * p.h_out * p.w_out;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int mt = t % pp.m_tiles;
      const int rest = t / pp.m_tiles;
      const int nt = rest % pp.n_tiles;
      const int z = rest / pp.n_tiles;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int img = tn * p.bn + li, oh = th * p.bh + lh, ow = tw * p.bw + lw;
      const int n0 = nt * BLOCK_N;
      const bool row_ok = (li < p.bn) && (img < p.n_img) && (oh < p.h_out) && (ow < p.w_out);
      const long long pix = (static_cast<long long>(img) * p.h_out + oh) * p.w_out + ow;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      if (warp == 2 && lane == 0 && it == 0) MDB_TRACE(7);
      const uint32_t lane_addr = tmem_base + as * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);

      if (p.epi_mode == EPI_GEGLU) {
        constexpr int HALF = BLOCK_N / 2;
        const int on0 = nt * HALF;
        const int n_half = p.n_out / 2;
#pragma unroll 1
        for (int c = eg * 16; c < HALF; c += 32) {
          __syncwarp();
          if (on0 + c >= n_half) break;  // warp-uniform
          // bias first (independent loads in flight while the TMEM load completes)
          float4 bv[4], bg[4];
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + 4 * j));
              bg[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + HALF + c + 4 * j));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = bg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          uint32_t v[16], g[16];
          tmem_ld_32x16(lane_addr + c, v);
          tmem_ld_32x16(lane_addr + HALF + c, g);
          tmem_ld_wait();
          if (row_ok) {
            const float* bvf = reinterpret_cast<const float*>(bv);
            const float* bgf = reinterpret_cast<const float*>(bg);
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float a0 = __uint_as_float(v[j]) + bvf[j], a1 = __uint_as_float(v[j + 1]) + bvf[j + 1];
              const float g0 = __uint_as_float(g[j]) + bgf[j], g1 = __uint_as_float(g[j + 1]) + bgf[j + 1];
              o[j >> 1] = pack_bf16(a0 * gelu_erf(g0), a1 * gelu_erf(g1));
            }
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + on0 + c);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
      } else {
#pragma unroll 1
        for (int c = eg * 32; c < BLOCK_N; c += 64) {
          __syncwarp();
          if (warp == 2 && lane == 0 && it == 0 && trace && c < 192) trace[12 + c / 64] = clock64();
          const int col0 = n0 + c;
          if (col0 >= p.n_out) break;  // warp-uniform: nothing left in this tile
          const bool full = (col0 + 32 <= p.n_out);
          if (p.epi_mode == EPI_PARTIAL_F32) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + c, v);
            tmem_ld_wait();
            if (row_ok) {
              float* dst = p.partial + (static_cast<long long>(z) * pixels_total + pix) * p.n_out + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (col0 + j < p.n_out)
                  *reinterpret_cast<float4*>(dst + j) =
                      make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                  __uint_as_float(v[j + 3]));
              }
            }
            continue;
          }
          // ---- 1. issue every global load of this chunk before anything consumes one (the old per-column guards
          //         serialised ~20 load latencies per chunk: 18k cycles per tile on the K=320 GEMMs)
          float4 bv[8], rv[8];
          uint4 res[4];
          const int ncol4 = full ? 8 : (p.n_out - col0) / 4;  // valid float4 groups (n_out % 8 == 0)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int jj = (j < ncol4) ? j : 0;  // clamp: always a valid address, never stored when out of range
            bv[j] = (p.bias && !(p.debug_flags & 2)) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * jj)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const bool use_rb = (p.rowbias != nullptr) && row_ok && !(p.debug_flags & 2);
          if (use_rb) {
            const float* rb = p.rowbias + static_cast<long long>(img) * p.rowbias_ld + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = __ldg(reinterpret_cast<const float4*>(rb + 4 * ((j < ncol4) ? j : 0)));
          }
          const bool use_res = (p.residual != nullptr) && row_ok && !(p.debug_flags & 2);
          if (use_res) {
            const __nv_bfloat16* rs = p.residual + pix * p.ldr + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = __ldg(reinterpret_cast<const uint4*>(rs + 8 * ((2 * j < ncol4) ? j : 0)));
          }
          // ---- 2. accumulator
          uint32_t v[32];
          if (!(p.debug_flags & 4)) {
            tmem_ld_32x32(lane_addr + c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0x3f800000u + j;
          }
          if (!row_ok) continue;
          // ---- 3. math
          float f[32];
          const float* bvf = reinterpret_cast<const float*>(bv);
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + bvf[j];
          if (use_rb) {
            const float* rvf = reinterpret_cast<const float*>(rv);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += rvf[j];
          }
          if (p.out_scale != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= p.out_scale;
          }
          if (use_res) {
            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(res);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 rf = __bfloat1622float2(rh[j]);
              f[2 * j] += rf.x;
              f[2 * j + 1] += rf.y;
            }
          }
          // ---- 4. store
          if (p.debug_flags & 1) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += f[j];
            if (acc == 1.2345e33f) static_cast<float*>(p.out)[0] = acc;  // keep the math alive
          } else if (p.out_is_f32) {
            float* dst = static_cast<float*>(p.out) + pix * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < ncol4) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (2 * j < ncol4)
                *reinterpret_cast<uint4*>(dst + 8 * j) =
                    make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                               pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
            }
          }
        }
      }
      // release this accumulator stage to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (warp == 2 && lane == 0 && it == 0) MDB_TRACE(8);
    }
    if (warp == 2 && lane == 0) {
      MDB_TRACE(9);
      if (trace) trace[11] = it;
    }
  }

  __syncthreads();
  if (threadIdx.x == 0) MDB_TRACE(10);
#undef MDB_TRACE
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mdb
