// Operand scheme and shared definitions of the tcgen05 GEMM / implicit-GEMM convolution kernels for sm_100a
// (gemm_tc2.cuh: single-CTA persistent kernel with split-K; gemm_pair.cuh: CTA-pair kernel with the TMA-store epilogue).
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

// Deterministic split-K finalisation: sum the fp32 partials in split order, then the same linear epilogue.
__global__ void splitk_finalize_kernel(const float* __restrict__ partial, int splits, long long pixels, int n_out,
                                       int hw_out, const float* __restrict__ bias, const float* __restrict__ rowbias,
                                       int rowbias_ld, const __nv_bfloat16* __restrict__ residual, int ldr,
                                       void* __restrict__ out, int ldo, int out_is_f32, float out_scale) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
