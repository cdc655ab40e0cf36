// Unified tcgen05 GEMM / implicit-GEMM convolution kernel for sm_100a.
//
//   D[pixel, n] = sum_{tap, c}  A[pixel shifted by tap, c] * W[n, tap*C + c]      (+ epilogue)
//
// * A is an NHWC bf16 activation tensor addressed through a 4-D TMA tensor map (C, W, H, N innermost-first).
//   One CTA owns a 128-row M tile that is a (bn x bh x bw) box of output pixels; each filter tap is the same
//   box shifted by (r - pad_h, s - pad_w) and TMA zero-fills the out-of-bounds halo, so there is no im2col.
//   Stride-2 convolutions use the tensor map's element strides.  A plain GEMM is the 1x1 / 1-image case.
//   Two A sources are supported (the K loop runs over source 0's channels then source 1's): this is how
//   channel-concatenated inputs (UNet skip connections) are consumed without materialising the concat.
// * W is a bf16 [N, K] (K-major) matrix behind a 2-D tensor map, K ordered (tap, channel).
// * Both operands land in shared memory with the 128-byte swizzle and are consumed by tcgen05.mma
//   (M=128, N=BLOCK_N, K=16, bf16 x bf16 -> fp32 in TMEM).
// * Warp roles: warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator, warps 4..7 = epilogue
//   (tcgen05.ld -> bias / per-image row bias / residual / GEGLU -> bf16 or fp32 global stores).
#pragma once
#include "ptx.cuh"

namespace mdb {

enum EpiMode : int { EPI_LINEAR = 0, EPI_GEGLU = 1, EPI_PARTIAL_F32 = 2 };

struct GemmParams {
  // output pixel grid
  int n_img, h_out, w_out;
  int n_out;  // GEMM N (for GEGLU: the packed 2x width)
  // filter
  int taps_h, taps_w, stride, pad_h, pad_w;
  int cblocks0, cblocks1;  // 64-channel blocks of A source 0 / 1
  // M-tile box
  int bn, bh, bw;
  int tiles_h, tiles_w;  // tiles per image column/row direction (tiles over images = gridDim.x / (tiles_h*tiles_w))
  // split-K
  int kb_per_split;
  // epilogue
  int epi_mode;
  int out_is_f32;
  const float* bias;     // [n_out] or nullptr
  const float* rowbias;  // [n_img][rowbias_ld] or nullptr (time-embedding shift, per image)
  int rowbias_ld;
  const __nv_bfloat16* residual;  // [pixels][ldr] or nullptr
  int ldr;
  void* out;  // bf16 or fp32 [pixels][ldo]
  int ldo;
  float out_scale;
  float* partial;  // [splits][pixels][n_out] fp32 workspace (EPI_PARTIAL_F32)
  int debug_flags;   // debug/ablation: 1 = skip output stores, 2 = skip bias/shift/residual loads, 4 = skip TMEM loads
  long long* trace;  // optional debug: per-CTA clock64 stamps (16 slots per CTA, first 8 CTAs)
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 160) ? 5 : 6);
  static constexpr int kTmemCols = (BLOCK_N <= 32) ? 32 : (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128) ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Exact-form GELU 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 plus the
// ~1e-7 relative error of the approximate rcp / ex2 units: far below the bf16 output rounding).  2 MUFU + 11 FP32 ops;
// erff() costs ~30 instructions and made the GEGLU epilogue instruction-bound (38 instr / output element).
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(hx, copysignf(erf_abs, x), hx);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int mt = blockIdx.x;
  const int tw = mt % p.tiles_w;
  const int th = (mt / p.tiles_w) % p.tiles_h;
  const int tn = mt / (p.tiles_w * p.tiles_h);
  const int img0 = tn * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
  const int n0 = blockIdx.y * BLOCK_N;
  const int cb_total = p.cblocks0 + p.cblocks1;
  const int kb_total = p.taps_h * p.taps_w * cb_total;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;

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
    mbar_init(accum_bar, 1);
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

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      const uint32_t tx_bytes = static_cast<uint32_t>(p.bn * p.bh * p.bw) * (kBlockK * 2) + Cfg::kBBytes;
      int stage = 0;
      uint32_t phase = 0;
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
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smA + stage * kABytes));
        const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smB + stage * Cfg::kBBytes));
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (i == nkb - 1) umma_commit(accum_bar);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int box_hw = p.bh * p.bw;
    const int li = row / box_hw;
    const int rem = row - li * box_hw;
    const int lh = rem / p.bw;
    const int lw = rem - lh * p.bw;
    const int img = img0 + li, oh = h0 + lh, ow = w0 + lw;
    const bool row_ok = (li < p.bn) && (img < p.n_img) && (oh < p.h_out) && (ow < p.w_out);
    const long long pix = (static_cast<long long>(img) * p.h_out + oh) * p.w_out + ow;

    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    if (p.epi_mode == EPI_GEGLU) {
      // columns [0, BLOCK_N/2) hold the value half, [BLOCK_N/2, BLOCK_N) the gate half of the same outputs
      constexpr int HALF = BLOCK_N / 2;
      const int on0 = blockIdx.y * HALF;
      const int n_half = p.n_out / 2;
#pragma unroll 1
      for (int c = 0; c < HALF; c += 16) {
        __syncwarp();
        uint32_t v[16], g[16];
        tmem_ld_32x16(lane_addr + c, v);
        tmem_ld_32x16(lane_addr + HALF + c, g);
        tmem_ld_wait();
        if (row_ok && on0 + c < n_half) {
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
            float g0 = __uint_as_float(g[j]), g1 = __uint_as_float(g[j + 1]);
            if (p.bias) {
              a0 += __ldg(p.bias + n0 + c + j);
              a1 += __ldg(p.bias + n0 + c + j + 1);
              g0 += __ldg(p.bias + n0 + HALF + c + j);
              g1 += __ldg(p.bias + n0 + HALF + c + j + 1);
            }
            o[j >> 1] = pack_bf16(a0 * gelu_erf(g0), a1 * gelu_erf(g1));
          }
          uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + on0 + c);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        __syncwarp();
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + c, v);
        tmem_ld_wait();
        const int col0 = n0 + c;
        if (!row_ok || col0 >= p.n_out) continue;
        if (p.epi_mode == EPI_PARTIAL_F32) {
          float* dst = p.partial + (static_cast<long long>(blockIdx.z) * p.n_img * p.h_out * p.w_out + pix) * p.n_out + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < p.n_out)
              *reinterpret_cast<float4*>(dst + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          }
          continue;
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < p.n_out) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              f[j] += b.x, f[j + 1] += b.y, f[j + 2] += b.z, f[j + 3] += b.w;
            }
          }
        }
        if (p.rowbias) {
          const float* rb = p.rowbias + static_cast<long long>(img) * p.rowbias_ld + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < p.n_out) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(rb + j));
              f[j] += b.x, f[j + 1] += b.y, f[j + 2] += b.z, f[j + 3] += b.w;
            }
          }
        }
        if (p.out_scale != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= p.out_scale;
        }
        if (p.residual) {
          const __nv_bfloat16* rs = p.residual + pix * p.ldr + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.n_out) {
              const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(rs + j));
              const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 rf = __bfloat1622float2(rh[t]);
                f[j + 2 * t] += rf.x;
                f[j + 2 * t + 1] += rf.y;
              }
            }
          }
        }
        if (p.out_is_f32) {
          float* dst = static_cast<float*>(p.out) + pix * p.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < p.n_out) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
          __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.n_out)
              *reinterpret_cast<uint4*>(dst + j) =
                  make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]), pack_bf16(f[j + 4], f[j + 5]),
                             pack_bf16(f[j + 6], f[j + 7]));
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Deterministic split-K finalisation: sum the fp32 partials in split order, then the same linear epilogue.
__global__ void splitk_finalize_kernel(const float* __restrict__ partial, int splits, long long pixels, int n_out,
                                       int hw_out, const float* __restrict__ bias, const float* __restrict__ rowbias,
                                       int rowbias_ld, const __nv_bfloat16* __restrict__ residual, int ldr,
                                       void* __restrict__ out, int ldo, int out_is_f32, float out_scale) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (idx >= pixels * n_out) return;
  const long long pix = idx / n_out;
  const int col = static_cast<int>(idx - pix * n_out);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 t = *reinterpret_cast<const float4*>(partial + (static_cast<long long>(s) * pixels + pix) * n_out + col);
    acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
  }
  float f[4] = {acc.x, acc.y, acc.z, acc.w};
  if (bias)
    for (int j = 0; j < 4; ++j) f[j] += bias[col + j];
  if (rowbias) {
    const long long img = pix / hw_out;
    for (int j = 0; j < 4; ++j) f[j] += rowbias[img * rowbias_ld + col + j];
  }
  for (int j = 0; j < 4; ++j) f[j] *= out_scale;
  if (residual)
    for (int j = 0; j < 4; ++j) f[j] += __bfloat162float(residual[pix * ldr + col + j]);
  if (out_is_f32) {
    *reinterpret_cast<float4*>(static_cast<float*>(out) + pix * ldo + col) = make_float4(f[0], f[1], f[2], f[3]);
  } else {
    uint2 o = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + pix * ldo + col) = o;
  }
}

}  // namespace mdb
