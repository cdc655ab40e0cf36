// Fused multi-head attention forward (flash-style, online softmax in fp32) on tcgen05 + TMA: C-ABI launchers.
//
// Layout: q/out [B, Lq, heads*D] (row strides ldq/ldo), k/v [Bkv, Lk, heads*D] (ldk/ldv): exactly what the fused
// QKV projection GEMM writes, so no head transpose is ever materialised.  With n_sets == 2 the kernel runs two
// independent softmaxes against two KV batches and writes their sum: the cross-view "add" mode of
// BasicMultiviewTransformerBlock (magicdrive/networks/blocks.py:112-121, 213-217).  K/V may be spread over up to three
// buffers (mdb_attention_multi): in view-sharded runs the neighbour views' K/V are read in place from the ring-neighbour
// GPUs' buffers through NVLink peer memory (the tensor maps simply point at peer-mapped addresses).
// Kernels: attention_tc2.cuh (head dim <= 80, production), attention_tc.cuh (head dim 160, and the A/B variant "tc").
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include "../../include/magicdrive_b200.h"
#define MDB_NEED_TENSORMAP
#include "common_host.h"
#include <string.h>
#include "attention_tc.cuh"
#include "attention_tc2.cuh"

static long long* g_attn_trace = nullptr;  // debug: device int64[3*16*8] for attention_tc2's phase stamps (mdb_attention_debug_trace)

namespace {

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

struct KvSources {
  const void* k[mdb::ATT_MAX_SRC];
  const void* v[mdb::ATT_MAX_SRC];
  int ldk[mdb::ATT_MAX_SRC], ldv[mdb::ATT_MAX_SRC], b_kv[mdb::ATT_MAX_SRC];
  int n;
};

bool make_kv_maps(mdb::AttnKvMaps* m, const KvSources& s, int d, int heads, int lk) {
  for (int i = 0; i < mdb::ATT_MAX_SRC; ++i) {
    const int j = i < s.n ? i : 0;  // unused slots alias source 0
    if (!make_qkv_map(&m->k[i], s.k[j], d, heads, lk, s.b_kv[j], s.ldk[j]) ||
        !make_qkv_map(&m->v[i], s.v[j], d, heads, lk, s.b_kv[j], s.ldv[j]))
      return false;
  }
  return true;
}

int attn_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

// Query tiles per CTA walk (AttnTcParams::q_step) for the single-S tc2 kernel when one K/V tile covers all keys: as many
// CTAs as two per SM, each keeping its K/V tile and walking ~nq / gx query tiles.  0 = one tile per CTA.  MDB_ATTN_MULTIQ=0 disables (A/B).
int multi_q_step(int b, int heads, int lq, int lk, int n_sets) {
  if (lk > mdb::ATT_BN || n_sets != 1) return 0;
  const char* e = getenv("MDB_ATTN_MULTIQ");
  if (e && e[0] == '0') return 0;
  const int nq = (lq + mdb::ATT_BM - 1) / mdb::ATT_BM;
  const long long per_tile_ctas = static_cast<long long>(b) * heads;
  const long long slots = 2LL * attn_num_sms();
  if (per_tile_ctas * nq <= slots) return 0;  // everything is co-resident anyway
  long long gx = slots / per_tile_ctas;
  if (gx < 1) gx = 1;
  if (gx >= nq) return 0;
  return static_cast<int>(gx);
}

template <typename Cfg, typename Kernel>
int launch_tc(Kernel kernel, const char* name, const void* q, int ldq, const KvSources& src, void* out, int ldo, int b, int heads,
              int lq, int lk, int d, const int* kv_index, int n_sets, float scale, cudaStream_t st, bool* attr, bool allow_multi_q = false) {
  if (!*attr) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "%s smem attr: %s", name, cudaGetErrorString(e));
    *attr = true;
  }
  CUtensorMap tq;
  mdb::AttnKvMaps kvm;
  if (!make_qkv_map(&tq, q, d, heads, lq, b, ldq) || !make_kv_maps(&kvm, src, d, heads, lk))
    return mdb::set_error(MDB_ERR_CUDA, "mdb_attention: cuTensorMapEncodeTiled failed (d=%d heads=%d lq=%d lk=%d)", d, heads, lq, lk);
  mdb::AttnTcParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo, p.lq = lq, p.lk = lk, p.kv_index = kv_index, p.n_sets = n_sets, p.n_src = src.n;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  p.q_step = allow_multi_q ? multi_q_step(b, heads, lq, lk, n_sets) : 0;
  dim3 grid(p.q_step > 0 ? p.q_step : (lq + mdb::ATT_BM - 1) / mdb::ATT_BM, heads, b);
  cudaError_t le = mdb::launch_pdl(kernel, grid, dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tq, kvm, p);
  if (le != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(le));
  cudaError_t e2 = cudaGetLastError();
  if (e2 != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e2));
  return MDB_OK;
}

template <int D>
int launch_attention_tc(const void* q, int ldq, const KvSources& src, void* out, int ldo, int b, int heads, int lq, int lk,
                        const int* kv_index, int n_sets, float scale, cudaStream_t st) {
  static bool attr = false;
  return launch_tc<mdb::AttnTcCfg<D>>(mdb::attention_tc_kernel<D>, "attention_tc_kernel", q, ldq, src, out, ldo, b, heads, lq, lk, D,
                                      kv_index, n_sets, scale, st, &attr);
}

template <int D, bool DOUBLE_S>
int launch_attention_tc2(const void* q, int ldq, const KvSources& src, void* out, int ldo, int b, int heads, int lq, int lk,
                         const int* kv_index, int n_sets, float scale, cudaStream_t st) {
  if constexpr (!DOUBLE_S) {
    if (multi_q_step(b, heads, lq, lk, n_sets) > 0) {  // one K/V tile, more query tiles than CTA slots: multi-Q instantiation
      static bool attr_mq = false;
      return launch_tc<mdb::AttnTc2Cfg<D, false>>(mdb::attention_tc2_kernel<D, false, true>, "attention_tc2_kernel<multi-Q>", q, ldq, src,
                                                  out, ldo, b, heads, lq, lk, D, kv_index, n_sets, scale, st, &attr_mq, true);
    }
  }
  static bool attr = false;
  return launch_tc<mdb::AttnTc2Cfg<D, DOUBLE_S>>(mdb::attention_tc2_kernel<D, DOUBLE_S, false>, "attention_tc2_kernel", q, ldq, src, out,
                                                 ldo, b, heads, lq, lk, D, kv_index, n_sets, scale, st, &attr, false);
}

// Which kernel generation serves a call: MDB_ATTN_KERNEL = tc2 | tc2d | tc (A/B switch, read per call).
enum class AttnKernel { kTc, kTc2, kTc2Double };
AttnKernel attention_kernel_choice() {
  const char* e = getenv("MDB_ATTN_KERNEL");
  if (e && !strcmp(e, "tc2d")) return AttnKernel::kTc2Double;
  if (e && !strcmp(e, "tc")) return AttnKernel::kTc;
  return AttnKernel::kTc2;  // default: tc2 (two CTAs/SM, one S buffer) for d <= 64, double-buffered S for d = 80, tc for d = 160
}

int attention_dispatch(const void* q, int ldq, const KvSources& src, void* out, int ldo, int b, int heads, int lq, int lk, int d,
                       const int* kv_index, int n_sets, float scale, cudaStream_t st) {
  using namespace mdb;
  const AttnKernel which = attention_kernel_choice();
#define MDB_TC2(DD, DBL) return launch_attention_tc2<DD, DBL>(q, ldq, src, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st)
  if (which != AttnKernel::kTc) {
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
  switch (d) {
    case 40: return launch_attention_tc<40>(q, ldq, src, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 80: return launch_attention_tc<80>(q, ldq, src, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 160: return launch_attention_tc<160>(q, ldq, src, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 32: return launch_attention_tc<32>(q, ldq, src, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    case 64: return launch_attention_tc<64>(q, ldq, src, out, ldo, b, heads, lq, lk, kv_index, n_sets, scale, st);
    default: return set_error(MDB_ERR_UNSUPPORTED, "mdb_attention: head dim %d not instantiated (32, 40, 64, 80, 160)", d);
  }
}

}  // namespace

extern "C" int mdb_attention_multi(const void* q, int ldq, int n_src, const void* const* k, const int* ldk, const void* const* v,
                                   const int* ldv, const int* b_kv, void* out, int ldo, int b, int heads, int lq, int lk, int d,
                                   const int* kv_index, int n_sets, float scale, void* stream) {
  using namespace mdb;
  if (!q || !k || !v || !ldk || !ldv || !b_kv || !out) return set_error(MDB_ERR_INVALID, "mdb_attention: null pointer");
  if (n_src < 1 || n_src > ATT_MAX_SRC) return set_error(MDB_ERR_INVALID, "mdb_attention: 1..%d K/V sources", ATT_MAX_SRC);
  if (n_sets < 1 || n_sets > 2 || (n_sets == 2 && !kv_index))
    return set_error(MDB_ERR_INVALID, "mdb_attention: n_sets must be 1 or 2 (2 needs kv_index)");
  if (scale <= 0.f) return set_error(MDB_ERR_INVALID, "mdb_attention: scale must be positive");
  if (lq <= 0 || lk <= 0 || b <= 0 || heads <= 0) return set_error(MDB_ERR_INVALID, "mdb_attention: bad shape");
  if (ldq % 8 || ldo % 8) return set_error(MDB_ERR_UNSUPPORTED, "mdb_attention: strides must be multiples of 8");
  KvSources src;
  src.n = n_src;
  for (int i = 0; i < n_src; ++i) {
    if (!k[i] || !v[i] || b_kv[i] <= 0) return set_error(MDB_ERR_INVALID, "mdb_attention: bad K/V source %d", i);
    if (ldk[i] % 8 || ldv[i] % 8) return set_error(MDB_ERR_UNSUPPORTED, "mdb_attention: strides must be multiples of 8");
    src.k[i] = k[i], src.v[i] = v[i], src.ldk[i] = ldk[i], src.ldv[i] = ldv[i], src.b_kv[i] = b_kv[i];
  }
  if (!kv_index && (n_src != 1 || b_kv[0] != b)) return set_error(MDB_ERR_INVALID, "mdb_attention: b_kv != b or several sources need kv_index");
  return attention_dispatch(q, ldq, src, out, ldo, b, heads, lq, lk, d, kv_index, n_sets, scale, static_cast<cudaStream_t>(stream));
}

extern "C" int mdb_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int b,
                             int b_kv, int heads, int lq, int lk, int d, const int* kv_index, int n_sets, float scale,
                             void* stream) {
  return mdb_attention_multi(q, ldq, 1, &k, &ldk, &v, &ldv, &b_kv, out, ldo, b, heads, lq, lk, d, kv_index, n_sets, scale, stream);
}

extern "C" int mdb_attention_debug_trace(void* device_i64_384) {
  g_attn_trace = static_cast<long long*>(device_i64_384);
  return MDB_OK;
}
