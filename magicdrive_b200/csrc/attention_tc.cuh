// tcgen05 fused attention forward (flash-style) for sm_100a.
//
//   per CTA: 128 queries of one (batch, head); loop over 128-key tiles of the (possibly neighbouring) batch's K/V:
//     S = Q K^T          tcgen05.mma SS  (Q, K tiles K-major in 128B-swizzled smem via 4-D TMA; d zero-padded to 64-chunks
//                                          by TMA out-of-bounds fill, so head dims 40 / 80 / 160 need no repacking)
//     P = exp2(S*c - m)  4 softmax warps, one query row per thread: tcgen05.ld -> fp32 online softmax -> bf16 pairs
//                        written back to TMEM with tcgen05.st (P never touches shared memory)
//     O_j = P V          tcgen05.mma TS  (A = P from TMEM, B = V tile, MN-major straight from the [keys, d] layout)
//     o = o*corr + O_j   registers (fp32), normalised and stored at the end
//   warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: softmax / epilogue (TMEM lane quarter = warp & 3).
//   n_sets = 2 runs the loop twice against two KV batches (cross-view neighbours) and sums the two normalised
//   outputs (each rounded to bf16 first, like the reference's per-branch outputs).
#pragma once
#include "ptx.cuh"

namespace mdb {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 128;

// K/V of a call may live in up to three buffers (this GPU's and, in view-sharded runs, the two ring-neighbour GPUs' buffers
// mapped through NVLink peer memory): kv_index entries are (source << 24) | batch index inside that source.
constexpr int ATT_MAX_SRC = 3;
struct AttnKvMaps {
  CUtensorMap k[ATT_MAX_SRC], v[ATT_MAX_SRC];
};

struct AttnTcParams {
  __nv_bfloat16* out;
  int ldo;
  int lq, lk;
  const int* kv_index;
  int n_sets;
  int n_src;
  float scale_log2;
  int q_step;  // attention_tc2 (single S buffer) only: > 0 = the CTA walks the query tiles blockIdx.x, blockIdx.x + q_step, ... of its
               // (batch, head) with the K/V tile loaded once (lk <= 128, n_sets == 1: the conditioning cross-attention); 0 = one tile
  long long* trace;  // debug (mdb_attention_debug_trace): clock64 stamps of CTA (0,0,0): [3 warps][16 iterations][8 points]
};

template <int D>
struct AttnTcCfg {
  static constexpr int KD = (D + 63) / 64;       // 64-wide chunks of the head dim
  static constexpr int D16 = (D + 15) / 16 * 16;  // UMMA N of the PV product / k extent of QK^T
  static constexpr int STAGES = (D <= 80) ? 2 : 1;
  static constexpr int TILE = ATT_BM * 128;  // bytes of one [128 rows][64 bf16] swizzled tile
  static constexpr int SMEM_Q = KD * TILE;
  static constexpr int SMEM_KV = STAGES * 2 * KD * TILE;
  static constexpr int kSmemBytes = SMEM_Q + SMEM_KV + 1024 + 128 + 2 * 2 * ATT_BM * 4 /*row-max exchange*/;
  static constexpr int S_COL = 0;               // S: 128 fp32 columns
  static constexpr int P_COL = ATT_BN;          // P: 64 columns of bf16 pairs (own region: the two column-halves of the
                                                // softmax run concurrently, so P may not alias S)
  static constexpr int O_COL = ATT_BN + 64;     // O: D16 fp32 columns
  static constexpr int kTmemCols = (ATT_BN + 64 + D16 <= 256) ? 256 : 512;
  static constexpr int HW = D16 / 2;            // O columns owned by each softmax half
  static constexpr int kThreads = 64 + 8 * 32;  // TMA warp, MMA warp, 8 softmax warps
};

// TS form: A operand from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major B operand (V tile: rows = keys, 128 B = 64 head-dim elements per row, 128B swizzle):
//   SBO = 1024 B between 8-key groups (K direction), LBO = one whole [128 keys][64] tile between 64-wide d atoms.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((1024u >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int D>
__global__ void __launch_bounds__(AttnTcCfg<D>::kThreads, (D <= 64) ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ AttnKvMaps kvm,
                    const AttnTcParams p) {
  using Cfg = AttnTcCfg<D>;
  constexpr int KD = Cfg::KD, D16 = Cfg::D16, STAGES = Cfg::STAGES, TILE = Cfg::TILE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;
  uint8_t* smK = smQ + Cfg::SMEM_Q;                // [STAGES][KD][TILE]
  uint8_t* smV = smK + STAGES * KD * TILE;         // [STAGES][KD][TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smV + STAGES * KD * TILE);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [2]
  uint64_t* kv_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;        // MMA -> softmax: S ready
  uint64_t* p_full = bars + 6;        // softmax -> MMA: P in TMEM (count 4 warps)
  uint64_t* o_full = bars + 7;        // MMA -> softmax: O_j ready
  uint64_t* o_free = bars + 8;        // softmax -> MMA: O_j consumed (count 4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xchg = reinterpret_cast<float*>(bars + 16);  // [2 buffers][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * ATT_BM;
  const int ntiles = (p.lk + ATT_BN - 1) / ATT_BN;
  const int total_iters = ntiles * p.n_sets;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    for (int i = 0; i < p.n_src; ++i) {
      prefetch_tmap(&kvm.k[i]);
      prefetch_tmap(&kvm.v[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(o_free, 8);
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
  asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: prologue above overlapped the predecessor's tail
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, KD * TILE);
#pragma unroll
      for (int c = 0; c < KD; ++c) tma_load_4d(&tmQ, q_full, smQ + c * TILE, c * 64, head, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int set = 0; set < p.n_sets; ++set) {
        const int kve = p.kv_index ? p.kv_index[b * p.n_sets + set] : b;
        const int kvb = kve & 0xffffff;
        const CUtensorMap* km = &kvm.k[kve >> 24];
        const CUtensorMap* vm = &kvm.v[kve >> 24];
        for (int j = 0; j < ntiles; ++j) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], 2 * KD * TILE);
#pragma unroll
          for (int c = 0; c < KD; ++c) {
            tma_load_4d(km, &kv_full[stage], smK + (stage * KD + c) * TILE, c * 64, head, j * ATT_BN, kvb);
            tma_load_4d(vm, &kv_full[stage], smV + (stage * KD + c) * TILE, c * 64, head, j * ATT_BN, kvb);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN);                 // Q K^T: both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, D16) | (1u << 16);       // P V: B (V) MN-major
    const uint32_t tmem_s = tmem_base + Cfg::S_COL;
    const uint32_t tmem_o = tmem_base + Cfg::O_COL;
    mbar_wait(q_full, 0);
    int stage = 0;
    uint32_t phase = 0;
    // prologue: S_0
    auto issue_s = [&](int st) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D16 / 16; ++k) {
          const int c = k / 4, kk = k % 4;
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smQ + c * TILE)) + 2 * kk;
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smK + (st * KD + c) * TILE)) + 2 * kk;
          umma_bf16(tmem_s, adesc, bdesc, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int it = 0; it < total_iters; ++it) {
      // P_it ready (softmax finished reading S_it and wrote P into TMEM)
      mbar_wait(p_full, it & 1);
      // previous O consumed
      mbar_wait(o_free, (it & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          // A: P bf16 pairs, 8 TMEM columns per 16 keys; B: V tile rows k*16.. (2048 B per 16 keys)
          const uint64_t bdesc = make_sw128_mnmajor_desc(smem_u32(smV + stage * KD * TILE), TILE) + (2048u >> 4) * k;
          umma_bf16_ts(tmem_o, tmem_base + Cfg::P_COL + 8 * k, bdesc, idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
      if (it + 1 < total_iters) {
        // S_{it+1}: needs K_{it+1}; S/P columns are free once PV_it has *completed* (it reads P from TMEM), which
        // tcgen05 guarantees by executing MMAs of one CTA in issue order
        mbar_wait(&kv_full[stage], phase);
        tc_fence_after();
        issue_s(stage);
      }
    }
  } else {
    // =========================== softmax + epilogue (warps 2..9) ===========================
    // Two warps per TMEM lane quarter: `hlf` 0 owns S columns [0,64) / O columns [0,HW), `hlf` 1 the rest.  One query
    // row per thread; the row max is exchanged between the two halves through shared memory once per tile, the row
    // sums only once per set.
    constexpr int HW = Cfg::HW;
    const int q = warp & 3;
    const int hlf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tmem_s = lane_base + Cfg::S_COL + hlf * 64;
    const uint32_t tmem_p = lane_base + Cfg::P_COL + hlf * 32;
    const uint32_t tmem_o = lane_base + Cfg::O_COL + hlf * HW;
    const int qrow = q0 + row;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.lq + qrow) * p.ldo + head * D + hlf * HW;
    constexpr int NVALID = (D - HW >= HW) ? HW : (D - HW);  // valid output columns of half 1 (half 0 always has HW)
    const int nvalid = hlf ? NVALID : HW;
    const float sc = p.scale_log2;
    int it = 0;
    for (int set = 0; set < p.n_sets; ++set) {
      float o[HW];
#pragma unroll
      for (int i = 0; i < HW; ++i) o[i] = 0.f;
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < ntiles; ++j, ++it) {
        mbar_wait(s_full, it & 1);
        tc_fence_after();
        const int kbase = j * ATT_BN + hlf * 64;   // first key of this thread's 64 columns
        const bool tail = (kbase + 64 > p.lk);
        // ---- pass 1: raw row max over the own 64 columns
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_s + c, v);
          tmem_ld_wait();
          if (!tail) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kbase + c + i < p.lk) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        float* xb = xchg + (it & 1) * 256;
        xb[hlf * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        mx = fmaxf(mx, xb[(hlf ^ 1) * 128 + row]);
        const float m_new = fmaxf(m, mx * sc);  // sc > 0
        const float corr = exp2f(m - m_new);
        m = m_new;
        // ---- pass 2: p = exp2(s*sc - m), row sum, bf16 pairs into the P region
        float rs = 0.f;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_s + c, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (!tail) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = exp2f(fmaf(__uint_as_float(v[i]), sc, -m_new));
              const float p1 = exp2f(fmaf(__uint_as_float(v[i + 1]), sc, -m_new));
              rs += p0 + p1;
              pk[i >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = (kbase + c + i < p.lk) ? exp2f(fmaf(__uint_as_float(v[i]), sc, -m_new)) : 0.f;
              const float p1 = (kbase + c + i + 1 < p.lk) ? exp2f(fmaf(__uint_as_float(v[i + 1]), sc, -m_new)) : 0.f;
              rs += p0 + p1;
              pk[i >> 1] = pack_bf16(p0, p1);
            }
          }
          tmem_st_32x16(tmem_p + (c >> 1), pk);
        }
        tmem_st_wait();
        l = l * corr + rs;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- O_j (own HW columns)
        mbar_wait(o_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < HW; c += 8) {
          uint32_t v[8];
          tmem_ld_32x8(tmem_o + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o[c + i] = o[c + i] * corr + __uint_as_float(v[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
      }
      // ---- finalise this set: total row sum = both halves
      float* xb = xchg + (it & 1) * 256;  // `it` parity alternates per tile, so this buffer is not in flight
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // partner has finished reading the last row max
      xb[hlf * 128 + row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float inv = 1.0f / (l + xb[(hlf ^ 1) * 128 + row]);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // both read before the next set reuses the buffer
      if (qrow < p.lq) {
#pragma unroll
        for (int c = 0; c < HW; c += 8) {
          if (c >= nvalid) break;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = o[c + i] * inv;
          if (set > 0) {
            const uint4 prev = *reinterpret_cast<const uint4*>(orow + c);
            const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 pf = __bfloat1622float2(ph[i]);
              // both branches rounded to bf16 before the sum, like the reference's per-branch attention outputs
              f[2 * i] = __bfloat162float(__float2bfloat16_rn(f[2 * i])) + pf.x;
              f[2 * i + 1] = __bfloat162float(__float2bfloat16_rn(f[2 * i + 1])) + pf.y;
            }
          }
          *reinterpret_cast<uint4*>(orow + c) =
              make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        }
      }
    }
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mdb
