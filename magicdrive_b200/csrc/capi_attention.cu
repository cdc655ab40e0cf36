// Fused multi-head attention forward (flash-style, online softmax in fp32) for head dims 40 / 80 / 160.
// Round-1 implementation on the legacy warp-level tensor path (mma.sync m16n8k16 bf16, cp.async double buffering);
// the tcgen05/TMEM version replaces it in a later round (DESIGN.md "attention").
//
// Layout: q/out [B, Lq, heads*D] (row strides ldq/ldo), k/v [Bkv, Lk, heads*D] (ldk/ldv): exactly what the fused
// QKV projection GEMM writes, so no head transpose is ever materialised.  With n_sets == 2 the kernel runs two
// independent softmaxes against two KV batches and writes their sum: the cross-view "add" mode of
// BasicMultiviewTransformerBlock (magicdrive/networks/blocks.py:112-121, 213-217).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include "../../include/magicdrive_b200.h"
#define MDB_NEED_TENSORMAP
#include "common_host.h"
#include <string.h>
#include "attention_tc.cuh"
#include "attention_tc2.cuh"
#include "attention_tc3.cuh"

static long long* g_attn_trace = nullptr;  // debug: device int64[3*16*8] for attention_tc2's phase stamps (mdb_attention_debug_trace)

namespace {

constexpr int BM = 64;   // queries per CTA (16 per warp)
constexpr int BNK = 64;  // keys per tile
constexpr int NTHREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int D>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;  // K dim of QK^T padded to the MMA k step
  static constexpr int LDS = DP + 8;             // smem row stride (elements): +16 B keeps ldmatrix conflict-free
  static constexpr int TILE = BM * LDS;          // elements per 64-row tile
  static constexpr int NT_O = D / 8;             // output n-tiles per warp
  // Q + 2 x (K + V) tiles (bf16) + optional fp32 accumulator for the two-set mode
  static constexpr size_t smem_bytes(int n_sets) {
    return sizeof(__nv_bfloat16) * TILE * 5 + (n_sets > 1 ? sizeof(float) * NTHREADS * (D / 2) : 0);
  }
};

template <int D>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int ld, int row0, int nrows_total) {
  // 64 rows x D bf16, 16-byte chunks; rows >= nrows_total are zero-filled
  constexpr int CH = D / 8;
  constexpr int LDS = AttnCfg<D>::LDS;
  for (int i = threadIdx.x; i < BM * CH; i += NTHREADS) {
    const int r = i / CH, c = i % CH;
    const bool ok = (row0 + r) < nrows_total;
    const __nv_bfloat16* s = src + static_cast<long long>(ok ? (row0 + r) : 0) * ld + c * 8;
    cp_async16(dst + r * LDS + c * 8, s, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(NTHREADS)
attention_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                 const __nv_bfloat16* __restrict__ v, int ldv, __nv_bfloat16* __restrict__ out, int ldo, int lq, int lk,
                 const int* __restrict__ kv_index, int n_sets, float scale_log2) {
  using Cfg = AttnCfg<D>;
  constexpr int DP = Cfg::DP, LDS = Cfg::LDS, TILE = Cfg::TILE, NT_O = Cfg::NT_O;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + TILE;      // 2 stages
  __nv_bfloat16* sV = sK + 2 * TILE;  // 2 stages
  float* sAcc = reinterpret_cast<float*>(sV + 2 * TILE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * BM;

  // zero the K-dim padding columns [D, DP) of Q and K tiles once (cp.async never writes them)
  if (DP > D) {
    for (int i = threadIdx.x; i < BM * (DP - D); i += NTHREADS) {
      const int r = i / (DP - D), c = D + i % (DP - D);
      sQ[r * LDS + c] = __float2bfloat16(0.f);
      sK[r * LDS + c] = __float2bfloat16(0.f);
      sK[TILE + r * LDS + c] = __float2bfloat16(0.f);
    }
  }
  const __nv_bfloat16* qb = q + static_cast<long long>(b) * lq * ldq + head * D;
  load_tile<D>(sQ, qb, ldq, q0, lq);
  cp_async_commit();

  const int ntiles = (lk + BNK - 1) / BNK;
  float o[NT_O][4];  // output accumulator (registers); the two-set mode sums the per-set results in smem

  for (int set = 0; set < n_sets; ++set) {
    const int kvb = kv_index ? kv_index[b * n_sets + set] : b;
    const __nv_bfloat16* kb = k + static_cast<long long>(kvb) * lk * ldk + head * D;
    const __nv_bfloat16* vb = v + static_cast<long long>(kvb) * lk * ldv + head * D;

#pragma unroll
    for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    __syncthreads();  // previous set finished reading the K/V stages
    load_tile<D>(sK, kb, ldk, 0, lk);
    load_tile<D>(sV, vb, ldv, 0, lk);
    cp_async_commit();

    for (int it = 0; it < ntiles; ++it) {
      const int st = it & 1;
      if (it + 1 < ntiles) {
        load_tile<D>(sK + (st ^ 1) * TILE, kb, ldk, (it + 1) * BNK, lk);
        load_tile<D>(sV + (st ^ 1) * TILE, vb, ldv, (it + 1) * BNK, lk);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const __nv_bfloat16* cK = sK + st * TILE;
      const __nv_bfloat16* cV = sV + st * TILE;

      // ---- S = Q K^T  (16 x 64 per warp)
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk) {
        uint32_t a[4];
        {
          const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int col = kk * 16 + (lane >> 4) * 8;
          ldsm_x4(smem_u32(sQ + row * LDS + col), a[0], a[1], a[2], a[3]);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of key n-tiles
          uint32_t b0, b1, b2, b3;
          const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
          const int col = kk * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(smem_u32(cK + row * LDS + col), b0, b1, b2, b3);
          mma_bf16(s[2 * np], a, b0, b1);
          mma_bf16(s[2 * np + 1], a, b2, b3);
        }
      }
      // ---- mask keys beyond lk, scale into log2 domain
      const int key_base = it * BNK;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = key_base + i * 8 + 2 * t;
        s[i][0] = (key < lk) ? s[i][0] * scale_log2 : -INFINITY;
        s[i][1] = (key + 1 < lk) ? s[i][1] * scale_log2 : -INFINITY;
        s[i][2] = (key < lk) ? s[i][2] * scale_log2 : -INFINITY;
        s[i][3] = (key + 1 < lk) ? s[i][3] * scale_log2 : -INFINITY;
      }
      // ---- online softmax (rows g and g+8)
      float mx0 = m0, mx1 = m1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
        mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float corr0 = exp2f(m0 - mx0), corr1 = exp2f(m1 - mx1);  // tile 0: exp2(-inf) = 0
      m0 = mx0, m1 = mx1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i][0] = exp2f(s[i][0] - mx0), s[i][1] = exp2f(s[i][1] - mx0);
        s[i][2] = exp2f(s[i][2] - mx1), s[i][3] = exp2f(s[i][3] - mx1);
        rs0 += s[i][0] + s[i][1];
        rs1 += s[i][2] + s[i][3];
      }
      l0 = l0 * corr0 + rs0;
      l1 = l1 * corr1 + rs1;
#pragma unroll
      for (int i = 0; i < NT_O; ++i) {
        o[i][0] *= corr0, o[i][1] *= corr0;
        o[i][2] *= corr1, o[i][3] *= corr1;
      }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < BNK / 16; ++kk) {
        uint32_t a[4];
        a[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int nt = 0; nt < NT_O; ++nt) {
          uint32_t b0, b1;
          const int row = kk * 16 + (lane & 15);
          ldsm_x2_t(smem_u32(cV + row * LDS + nt * 8), b0, b1);
          mma_bf16(o[nt], a, b0, b1);
        }
      }
      __syncthreads();  // all warps done with stage st before it is refilled
    }
    // ---- finalise this set
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    if (n_sets == 1) {
#pragma unroll
      for (int i = 0; i < NT_O; ++i) {
        o[i][0] *= inv0, o[i][1] *= inv0;
        o[i][2] *= inv1, o[i][3] *= inv1;
      }
    } else {
      // each branch's output is rounded to bf16 first, like the reference's per-branch attention output
#pragma unroll
      for (int i = 0; i < NT_O; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float val = __bfloat162float(__float2bfloat16_rn(o[i][j] * (j < 2 ? inv0 : inv1)));
          float* slot = sAcc + (i * 4 + j) * NTHREADS + threadIdx.x;
          *slot = (set == 0) ? val : (*slot + val);
        }
      }
    }
  }
  if (n_sets > 1) {
#pragma unroll
    for (int i = 0; i < NT_O; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = sAcc[(i * 4 + j) * NTHREADS + threadIdx.x];
  }
  // ---- store (rows g, g+8 of this warp's 16-row slab)
  __nv_bfloat16* ob = out + static_cast<long long>(b) * lq * ldo + head * D;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int i = 0; i < NT_O; ++i) {
    const int col = i * 8 + 2 * t;
    if (r0 < lq) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r0) * ldo + col) = pack2(o[i][0], o[i][1]);
    if (r1 < lq) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r1) * ldo + col) = pack2(o[i][2], o[i][3]);
  }
}

template <int D>
int launch_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int b,
                     int heads, int lq, int lk, const int* kv_index, int n_sets, float scale, cudaStream_t st) {
  const size_t smem = AttnCfg<D>::smem_bytes(n_sets);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "attention smem attr: %s", cudaGetErrorString(e));
    configured = smem;
  }
  dim3 grid((lq + BM - 1) / BM, heads, b);
  attention_kernel<D><<<grid, NTHREADS, smem, st>>>(
      static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
      static_cast<const __nv_bfloat16*>(v), ldv, static_cast<__nv_bfloat16*>(out), ldo, lq, lk, kv_index, n_sets,
      scale * 1.4426950408889634f);
  MDB_CHECK_LAUNCH("attention_kernel");
  return MDB_OK;
}

// ---------------------------------------------------------------- tcgen05 path
// [B, L, heads*D] (row stride ld) as a 4-D map (d, head, token, batch); box = (64, 1, 128, 1): the head dim is
// zero-padded to 64-wide chunks by TMA out-of-bounds fill, rows beyond L are zero-filled too.
bool make_qkv_map(CUtensorMap* m, const void* ptr, int d, int heads, int l, int b, int ld) {
  mdb::EncodeTiledFn enc = mdb::get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)l, (cuuint64_t)b};
  cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)ld * 2, (cuuint64_t)l * ld * 2};
  cuuint32_t box[4] = {64u, 1u, (cuuint32_t)mdb::ATT_BM, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int D>
int launch_attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int b,
                        int heads, int lq, int lk, int b_kv, const int* kv_index, int n_sets, float scale, cudaStream_t st) {
  using Cfg = mdb::AttnTcCfg<D>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mdb::attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "attention_tc smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  CUtensorMap tq, tk, tv;
  if (!make_qkv_map(&tq, q, D, heads, lq, b, ldq) || !make_qkv_map(&tk, k, D, heads, lk, b_kv, ldk) ||
      !make_qkv_map(&tv, v, D, heads, lk, b_kv, ldv))
    return mdb::set_error(MDB_ERR_CUDA, "mdb_attention: cuTensorMapEncodeTiled failed (d=%d heads=%d lq=%d lk=%d)", D, heads,
                          lq, lk);
  mdb::AttnTcParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo, p.lq = lq, p.lk = lk, p.kv_index = kv_index, p.n_sets = n_sets;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  dim3 grid((lq + mdb::ATT_BM - 1) / mdb::ATT_BM, heads, b);
  cudaError_t le = mdb::launch_pdl(mdb::attention_tc_kernel<D>, grid, dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tq, tk, tv, p);
  if (le != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "attention_tc_kernel launch: %s", cudaGetErrorString(le));
  MDB_CHECK_LAUNCH("attention_tc_kernel");
  return MDB_OK;
}

// Second-generation kernel (attention_tc2.cuh), head dim <= 80.
template <int D, bool DOUBLE_S>
int launch_attention_tc2(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int b,
                         int heads, int lq, int lk, int b_kv, const int* kv_index, int n_sets, float scale, cudaStream_t st) {
  using Cfg = mdb::AttnTc2Cfg<D, DOUBLE_S>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mdb::attention_tc2_kernel<D, DOUBLE_S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "attention_tc2 smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  CUtensorMap tq, tk, tv;
  if (!make_qkv_map(&tq, q, D, heads, lq, b, ldq) || !make_qkv_map(&tk, k, D, heads, lk, b_kv, ldk) ||
      !make_qkv_map(&tv, v, D, heads, lk, b_kv, ldv))
    return mdb::set_error(MDB_ERR_CUDA, "mdb_attention: cuTensorMapEncodeTiled failed (d=%d heads=%d lq=%d lk=%d)", D, heads,
                          lq, lk);
  mdb::AttnTcParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo, p.lq = lq, p.lk = lk, p.kv_index = kv_index, p.n_sets = n_sets;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  dim3 grid((lq + mdb::ATT_BM - 1) / mdb::ATT_BM, heads, b);
  cudaError_t le = mdb::launch_pdl(mdb::attention_tc2_kernel<D, DOUBLE_S>, grid, dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tq, tk, tv, p);
  if (le != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "attention_tc2_kernel launch: %s", cudaGetErrorString(le));
  MDB_CHECK_LAUNCH("attention_tc2_kernel");
  return MDB_OK;
}

// Persistent form (attention_tc3.cuh), head dim <= 64; opt-in until it has been run on a GPU.
template <int D>
int launch_attention_tc3(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int b,
                         int heads, int lq, int lk, int b_kv, const int* kv_index, int n_sets, float scale, cudaStream_t st) {
  using Cfg = mdb::AttnTc3Cfg<D>;
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mdb::attention_tc3_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "attention_tc3 smem attr: %s", cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  CUtensorMap tq, tk, tv;
  if (!make_qkv_map(&tq, q, D, heads, lq, b, ldq) || !make_qkv_map(&tk, k, D, heads, lk, b_kv, ldk) ||
      !make_qkv_map(&tv, v, D, heads, lk, b_kv, ldv))
    return mdb::set_error(MDB_ERR_CUDA, "mdb_attention: cuTensorMapEncodeTiled failed (d=%d heads=%d lq=%d lk=%d)", D, heads,
                          lq, lk);
  mdb::AttnTc3Params p;
  p.a.out = static_cast<__nv_bfloat16*>(out);
  p.a.ldo = ldo, p.a.lq = lq, p.a.lk = lk, p.a.kv_index = kv_index, p.a.n_sets = n_sets;
  p.a.scale_log2 = scale * 1.4426950408889634f;
  p.n_qtiles = (lq + mdb::ATT_BM - 1) / mdb::ATT_BM, p.heads = heads, p.batch = b;
  const long long items = static_cast<long long>(p.n_qtiles) * heads * b;
  const int grid = static_cast<int>(items < 2LL * sms ? items : 2LL * sms);
  cudaError_t le = mdb::launch_pdl(mdb::attention_tc3_kernel<D>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tq, tk, tv, p);
  if (le != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "attention_tc3_kernel launch: %s", cudaGetErrorString(le));
  MDB_CHECK_LAUNCH("attention_tc3_kernel");
  return MDB_OK;
}

// Which kernel generation serves a call: MDB_ATTN_KERNEL = tc3 | tc2 | tc2d | tc | legacy (A/B switch, read per call).
enum class AttnKernel { kLegacy, kTc, kTc2, kTc2Double, kTc3 };
AttnKernel attention_kernel_choice() {
  const char* legacy = getenv("MDB_ATTN_LEGACY");
  if (legacy && legacy[0] == '1') return AttnKernel::kLegacy;
  const char* e = getenv("MDB_ATTN_KERNEL");
  if (e && !strcmp(e, "legacy")) return AttnKernel::kLegacy;
  if (e && !strcmp(e, "tc3")) return AttnKernel::kTc3;
  if (e && !strcmp(e, "tc2")) return AttnKernel::kTc2;
  if (e && !strcmp(e, "tc2d")) return AttnKernel::kTc2Double;
  if (e && !strcmp(e, "tc")) return AttnKernel::kTc;
  return AttnKernel::kTc2;  // default: tc2 (two CTAs/SM, one S buffer) for d <= 64, double-buffered S for d = 80, tc for d = 160
}

}  // namespace

extern "C" int mdb_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int b,
                             int b_kv, int heads, int lq, int lk, int d, const int* kv_index, int n_sets, float scale,
                             void* stream) {
  using namespace mdb;
  if (!q || !k || !v || !out) return set_error(MDB_ERR_INVALID, "mdb_attention: null pointer");
  if (n_sets < 1 || n_sets > 2 || (n_sets == 2 && !kv_index))
    return set_error(MDB_ERR_INVALID, "mdb_attention: n_sets must be 1 or 2 (2 needs kv_index)");
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return set_error(MDB_ERR_UNSUPPORTED, "mdb_attention: strides must be multiples of 8");
  if (lq <= 0 || lk <= 0 || b <= 0 || heads <= 0 || b_kv <= 0) return set_error(MDB_ERR_INVALID, "mdb_attention: bad shape");
  if (!kv_index && b_kv != b) return set_error(MDB_ERR_INVALID, "mdb_attention: b_kv != b needs kv_index");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AttnKernel which = attention_kernel_choice();
  if (which == AttnKernel::kTc3 && scale > 0.f) {  // persistent variant (opt-in); other head dims fall through to tc2
    switch (d) {
      case 40: return launch_attention_tc3<40>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      case 32: return launch_attention_tc3<32>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      case 64: return launch_attention_tc3<64>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      default: break;
    }
  }
#define MDB_TC2(DD, DBL) \
  return launch_attention_tc2<DD, DBL>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st)
  if ((which == AttnKernel::kTc2 || which == AttnKernel::kTc2Double || which == AttnKernel::kTc3) && scale > 0.f) {
    const bool dbl = which == AttnKernel::kTc2Double;  // tc2: two CTAs/SM with one S buffer where d <= 64
    switch (d) {
      case 40: if (dbl) MDB_TC2(40, true); else MDB_TC2(40, false);
      case 32: if (dbl) MDB_TC2(32, true); else MDB_TC2(32, false);
      case 64: if (dbl) MDB_TC2(64, true); else MDB_TC2(64, false);
      case 80: MDB_TC2(80, true);
      default: break;  // d = 160 stays on the first tcgen05 kernel
    }
  }
#undef MDB_TC2
  if (which != AttnKernel::kLegacy) {
    // tcgen05 path
    switch (d) {
      case 40: return launch_attention_tc<40>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      case 80: return launch_attention_tc<80>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      case 160: return launch_attention_tc<160>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      case 32: return launch_attention_tc<32>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      case 64: return launch_attention_tc<64>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, b_kv, kv_index, n_sets, scale, st);
      default: break;
    }
  }
  switch (d) {
    case 40: return launch_attention<40>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 80: return launch_attention<80>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 160: return launch_attention<160>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 32: return launch_attention<32>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 64: return launch_attention<64>(q, ldq, k, ldk, v, ldv, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    default: return set_error(MDB_ERR_UNSUPPORTED, "mdb_attention: head dim %d not instantiated (8,16,32,40,64,80,160)", d);
  }
}

extern "C" int mdb_attention_debug_trace(void* device_i64_384) {
  g_attn_trace = static_cast<long long*>(device_i64_384);
  return MDB_OK;
}
