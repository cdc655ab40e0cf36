// tcgen05 fused attention forward, second version (head dim <= 80).
//
// What changed against attention_tc.cuh (profiles/ncu_attention_tc_r1.analysis.txt: that kernel is bound by the serial
// chain  S ready -> softmax -> P -> PV -> O_j read -> next S  and by the MUFU queue):
//   * DOUBLE_S variant (d = 80): S is double-buffered in tensor memory, QK^T of tile j+1 / j+2 runs under the softmax of
//     tile j.  Single-S variant (d <= 64): 256 TMEM columns, two CTAs per SM cover each other's PV / QK^T bubbles and
//     prologues (measured faster than the double-buffered form at d = 40: 89.5 vs 108.6 us, tools/bench_attn.py).
//   * O stays in tensor memory across the whole KV loop (tcgen05.mma accumulates); the softmax warps never wait for PV.
//     The running maximum is only raised when it grows by more than 2^8 ("lazy rescale"): the rare raise waits for the
//     previous PV and scales the accumulator rows in place (tcgen05.ld / st).
//   * the two 64-key halves of a 128-key tile are two independent split-KV partial softmaxes (own running max, own row
//     sum, own O accumulator O_a / O_b); they are combined once per KV set.  No per-tile row-max exchange, so the two
//     warps sharing an SM sub-partition drift apart and overlap their MUFU and non-MUFU phases.
//   * P (bf16 pairs) is written over the S columns it was computed from (no separate P region).
//   TMEM columns (DOUBLE_S): S0 [0,128)  S1 [128,256)  O_a [256,256+D16)  O_b [384,384+D16);  single: S [0,128) O_a 128 O_b 192
//   * MULTI-Q mode (p.q_step > 0; single-S variant, lk <= 128, one K/V set: the text / camera / box cross-attention, 98 keys):
//     the CTA keeps its one K/V tile in shared memory and walks several query tiles, so that barrier initialisation, TMEM
//     allocation and the K/V load are paid once per 3-4 tiles instead of once per tile (one-tile CTAs ran at ~70 TFLOP/s:
//     each spent ~6 us on a fixed prologue / load / store chain for 0.3 us of tensor work).  The three roles reuse the loop over
//     K/V SETS: a "set" is then a query tile (own Q load, own output rows, no accumulation into the previous set's output).
//   warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: softmax
//   (TMEM lane quarter = warp & 3, key half = (warp - 2) >> 2).
#pragma once
#include "attention_tc.cuh"

namespace mdb {

// DOUBLE_S = true : one CTA per SM, two S buffers (512 TMEM columns), K/V ring of 4 (d <= 64) or 2 stages.
// DOUBLE_S = false: two CTAs per SM (d <= 64), one S buffer (256 TMEM columns), K/V ring of 2 stages; the second CTA
//                   covers the PV / QK^T bubble and the prologue / epilogue of the first.
template <int D, bool DOUBLE_S>
struct AttnTc2Cfg {
  static constexpr int KD = (D + 63) / 64;
  static constexpr int D16 = (D + 15) / 16 * 16;
  static constexpr int TILE = ATT_BM * 128;
  static constexpr int NBUF = DOUBLE_S ? 2 : 1;
  static constexpr int STAGES = (DOUBLE_S && KD == 1) ? 4 : 2;
  static constexpr int SMEM_Q = KD * TILE;
  static constexpr int SMEM_KV = STAGES * 2 * KD * TILE;
  static constexpr int kSmemBytes = SMEM_Q + SMEM_KV + 1024 /*alignment*/ + 256 /*barriers*/ + 4 * ATT_BM * 4 /*(m,l) exchange*/;
  static constexpr int OA_COL = DOUBLE_S ? 256 : 128, OB_COL = DOUBLE_S ? 384 : 192;
  static constexpr int kTmemCols = DOUBLE_S ? 512 : 256;
  static constexpr int kMinCtas = DOUBLE_S ? 1 : 2;
  static constexpr int HW = D16 / 2;  // output columns each softmax half normalises and stores
  static constexpr int kThreads = 64 + 8 * 32;
  static_assert(D16 <= (DOUBLE_S ? 128 : 64), "O_a / O_b spacing");
};

constexpr float ATT_LAZY_LOG2 = 8.0f;  // raise the reference maximum only when it grows by more than 2^8

template <int D, bool DOUBLE_S, bool MULTI_Q = false>
__global__ void __launch_bounds__(AttnTc2Cfg<D, DOUBLE_S>::kThreads, AttnTc2Cfg<D, DOUBLE_S>::kMinCtas)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ AttnKvMaps kvm,
                     const AttnTcParams p) {
  using Cfg = AttnTc2Cfg<D, DOUBLE_S>;
  constexpr int KD = Cfg::KD, D16 = Cfg::D16, STAGES = Cfg::STAGES, TILE = Cfg::TILE, HW = Cfg::HW, NBUF = Cfg::NBUF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;
  uint8_t* smK = smQ + Cfg::SMEM_Q;         // [STAGES][KD][TILE]
  uint8_t* smV = smK + STAGES * KD * TILE;  // [STAGES][KD][TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smV + STAGES * KD * TILE);
  uint64_t* q_full = bars;         // TMA -> MMA
  uint64_t* kv_full = bars + 1;    // [4] TMA -> MMA
  uint64_t* kv_empty = bars + 5;   // [4] MMA (PV complete) -> TMA
  uint64_t* s_full = bars + 9;     // [2] MMA -> softmax: S buffer ready
  uint64_t* p_full = bars + 11;    // [2] softmax (8 warps) -> MMA: P written over the S buffer
  uint64_t* pv_done = bars + 13;   // MMA -> softmax: PV of iteration `it` complete (one phase per iteration)
  uint64_t* o_full = bars + 14;    // MMA -> softmax: last PV of a KV set complete
  uint64_t* o_free = bars + 15;    // softmax (8 warps) -> MMA: O of the finished set has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  uint64_t* q_empty = bars + 17;   // multi-Q mode: MMA (QK^T complete) -> TMA: the Q tile may be replaced
  float* xm = reinterpret_cast<float*>(bars + 32);  // [2 halves][128 rows] running max
  float* xl = xm + 2 * ATT_BM;                      // [2 halves][128 rows] row sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y;
  const int ntiles = (p.lk + ATT_BN - 1) / ATT_BN;
  // multi-Q mode: the outer ("set") loop of every role runs over this CTA's query tiles instead of K/V sets
  static_assert(!(MULTI_Q && DOUBLE_S), "multi-Q mode exists for the single-S variant only");
  constexpr bool multi_q = MULTI_Q;  // its own instantiation: the one-tile kernel keeps its register budget (96 / thread)
  const int nq_tiles = (p.lq + ATT_BM - 1) / ATT_BM;
  const int q0 = blockIdx.x * ATT_BM;
  int n_outer = p.n_sets;
  if constexpr (MULTI_Q) n_outer = (nq_tiles - static_cast<int>(blockIdx.x) + p.q_step - 1) / p.q_step;
  auto q0_of = [&](int o) { return (static_cast<int>(blockIdx.x) + o * p.q_step) * ATT_BM; };  // multi-Q only
  const int total_iters = ntiles * n_outer;
  long long* const trace = (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.trace : nullptr;
#define MDB_ATRACE(slot, it_, k)                                                       \
  do {                                                                                 \
    if (trace && lane == 0 && (it_) < 16) trace[((slot) * 16 + (it_)) * 8 + (k)] = clock64(); \
  } while (0)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    for (int i = 0; i < p.n_src; ++i) {
      prefetch_tmap(&kvm.k[i]);
      prefetch_tmap(&kvm.v[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    mbar_init(o_free, 8);
    mbar_init(q_empty, 1);
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
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      auto load_q = [&](int qq) {
        mbar_arrive_expect_tx(q_full, KD * TILE);
#pragma unroll
        for (int c = 0; c < KD; ++c) tma_load_4d(&tmQ, q_full, smQ + c * TILE, c * 64, head, qq, b);
      };
      load_q(q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int set = 0; set < n_outer; ++set) {
        if constexpr (MULTI_Q) {
          if (set > 0) {  // next query tile once QK^T of the previous one has read Q; the K/V tile stays
            mbar_wait(q_empty, static_cast<uint32_t>(set - 1) & 1u);
            load_q(q0_of(set));
            continue;
          }
        }
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
    constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN);            // Q K^T: both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, D16) | (1u << 16);  // P V: B (V) MN-major
    if (!multi_q) mbar_wait(q_full, 0);
    // S_i -> buffer i % NBUF, K from ring stage i % STAGES (multi-Q: the one K/V tile in stage 0, Q tile i)
    auto issue_qk = [&](int i) {
      const int st = multi_q ? 0 : i % STAGES;
      if (multi_q) {
        mbar_wait(q_full, static_cast<uint32_t>(i) & 1u);
        if (i == 0) mbar_wait(&kv_full[0], 0);
      } else {
        mbar_wait(&kv_full[st], static_cast<uint32_t>(i / STAGES) & 1u);
      }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(i % NBUF) * ATT_BN;
#pragma unroll
        for (int k = 0; k < D16 / 16; ++k) {
          const int c = k / 4, kk = k % 4;
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smQ + c * TILE)) + 2 * kk;
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smK + (st * KD + c) * TILE)) + 2 * kk;
          umma_bf16(tmem_s, adesc, bdesc, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[i % NBUF]);
        if (multi_q) umma_commit(q_empty);
      }
      __syncwarp();
    };
    issue_qk(0);
    if (NBUF == 2 && total_iters > 1) issue_qk(1);
    int j = 0, set = 0;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % NBUF, st = multi_q ? 0 : it % STAGES;
      MDB_ATRACE(0, it, 0);
      mbar_wait(&p_full[s], static_cast<uint32_t>(it / NBUF) & 1u);  // P_it written, S_it no longer read
      MDB_ATRACE(0, it, 1);
      if (j == 0 && set > 0) mbar_wait(o_free, static_cast<uint32_t>(set - 1) & 1u);  // previous set's O has been read
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          const int half = k >> 2;
          // A: P bf16 pairs of 16 keys = 8 TMEM columns, at the start of each half's own 64 S columns
          const uint32_t tmem_a = tmem_base + static_cast<uint32_t>(s * ATT_BN + half * 64 + 8 * (k & 3));
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(half ? Cfg::OB_COL : Cfg::OA_COL);
          const uint64_t bdesc = make_sw128_mnmajor_desc(smem_u32(smV + st * KD * TILE), TILE) + (2048u >> 4) * k;
          umma_bf16_ts(tmem_d, tmem_a, bdesc, idesc_o, (j > 0 || (k & 3) > 0) ? 1u : 0u);
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[st]);
        if (j == ntiles - 1) umma_commit(o_full);
      }
      __syncwarp();
      MDB_ATRACE(0, it, 2);
      // S_{it+NBUF} reuses buffer s: tcgen05.mma of one CTA execute in issue order, so it starts after PV_it has read P_it
      if (it + NBUF < total_iters) issue_qk(it + NBUF);
      MDB_ATRACE(0, it, 3);
      if (++j == ntiles) {
        j = 0;
        ++set;
      }
    }
  } else {
    // =========================== softmax + epilogue (warps 2..9) ===========================
    const int q = warp & 3;
    const int hlf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tmem_acc = lane_base + (hlf ? Cfg::OB_COL : Cfg::OA_COL);  // this half's accumulator (all D16 columns)
    constexpr int NVALID = (D - HW >= HW) ? HW : (D - HW);  // valid output columns of half 1 (half 0 always has HW)
    const int nvalid = hlf ? NVALID : HW;
    const float sc = p.scale_log2;
    int it = 0;
    int qrow = q0 + row;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.lq + qrow) * p.ldo + head * D + hlf * HW;
    for (int set = 0; set < n_outer; ++set) {
      if constexpr (MULTI_Q) {
        qrow = q0_of(set) + row;
        orow = p.out + (static_cast<long long>(b) * p.lq + qrow) * p.ldo + head * D + hlf * HW;
      }
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < ntiles; ++j, ++it) {
        const int s = it % NBUF;
        const uint32_t tmem_s = lane_base + static_cast<uint32_t>(s * ATT_BN + hlf * 64);
        const int tslot = (warp == 2) ? 1 : ((warp == 6) ? 2 : -1);
        if (tslot > 0) MDB_ATRACE(tslot, it, 0);
        mbar_wait(&s_full[s], static_cast<uint32_t>(it / NBUF) & 1u);
        tc_fence_after();
        if (tslot > 0) MDB_ATRACE(tslot, it, 1);
        const int nval = p.lk - (j * ATT_BN + hlf * 64);  // valid keys among this thread's 64 columns
        const bool tail = nval < 64;
        float rs = 0.f;
        uint32_t pk[32];
        bool done = false;
#ifndef MDB_ATTN_NO_PIPELINED_SOFTMAX
        if (!tail && j > 0) {
          // ---- speculative fast path (every tile but the first and a ragged last one): the S row is read from tensor
          // memory in four 16-column pieces, piece i+1 in flight while piece i goes through the exponentials, using the
          // CURRENT reference maximum.  TMEM read bandwidth and the MUFU are the two comparable costs of a tile; reading
          // all 64 columns first (below) runs them back to back in every warp at the same time.  The guess "no element
          // exceeds m_ref + 2^8" is checked piece by piece; a miss (rare once the maximum has settled) falls through to
          // the exact path, which re-reads S (P is only written after the last piece).  Same m_ref, same arithmetic:
          // bit-identical to the exact path.
          uint32_t a[16], b[16];
          tmem_ld_32x16(tmem_s, a);
          tmem_ld_wait();
          bool ok = true;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t(&cur)[16] = (c & 1) ? b : a;
            uint32_t(&nxt)[16] = (c & 1) ? a : b;
            if (c < 3) tmem_ld_32x16(tmem_s + 16 * (c + 1), nxt);
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(cur[i]));
            const float cm = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;
            ok = ok && !(cm > m_ref + ATT_LAZY_LOG2);
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float p0 = exp2f(fmaf(__uint_as_float(cur[i]), sc, -m_ref));
              const float p1 = exp2f(fmaf(__uint_as_float(cur[i + 1]), sc, -m_ref));
              rs += p0 + p1;
              pk[8 * c + (i >> 1)] = pack_bf16(p0, p1);
            }
            if (c < 3) tmem_ld_wait();
          }
          done = !__any_sync(0xffffffffu, !ok);
          if (tslot > 0) MDB_ATRACE(tslot, it, 2);
          if (tslot > 0) MDB_ATRACE(tslot, it, 3);
        }
#endif
        if (!done) {
        rs = 0.f;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(tmem_s, v0);
        tmem_ld_32x32(tmem_s + 32, v1);
        tmem_ld_wait();
        if (tslot > 0) MDB_ATRACE(tslot, it, 2);
        float mx = -INFINITY;
        if (!tail) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < nval) mx = fmaxf(mx, __uint_as_float(v0[i]));
            if (32 + i < nval) mx = fmaxf(mx, __uint_as_float(v1[i]));
          }
        }
        if (tslot > 0) MDB_ATRACE(tslot, it, 3);
        const float m_new = mx * sc;  // sc > 0; -inf when this half has no valid key in the tile
        const bool need = m_new > m_ref + ATT_LAZY_LOG2;
        if (__any_sync(0xffffffffu, need)) {
          if (j > 0) {
            // the accumulator holds tiles < j of this set: their PVs must have completed before the rows are rescaled.
            // S_it being ready implies PV_{it-NBUF} completed (issue order), so the parity of phase it-1 is unambiguous.
            mbar_wait(pv_done, static_cast<uint32_t>(it - 1) & 1u);
            tc_fence_after();
            const float corr = need ? exp2f(m_ref - m_new) : 1.0f;
#pragma unroll 1
            for (int c = 0; c < D16; c += 8) {
              uint32_t o[8];
              tmem_ld_32x8(tmem_acc + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st_32x8(tmem_acc + c, o);
            }
            l *= corr;
          }
          if (need) m_ref = m_new;
        }
        // ---- p = exp2(s*sc - m_ref), row sum, bf16 pairs over the first 32 of the own 64 S columns
        if (!tail) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = exp2f(fmaf(__uint_as_float(v0[i]), sc, -m_ref));
            const float p1 = exp2f(fmaf(__uint_as_float(v0[i + 1]), sc, -m_ref));
            const float p2 = exp2f(fmaf(__uint_as_float(v1[i]), sc, -m_ref));
            const float p3 = exp2f(fmaf(__uint_as_float(v1[i + 1]), sc, -m_ref));
            rs += (p0 + p1) + (p2 + p3);
            pk[i >> 1] = pack_bf16(p0, p1);
            pk[16 + (i >> 1)] = pack_bf16(p2, p3);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (i < nval) ? exp2f(fmaf(__uint_as_float(v0[i]), sc, -m_ref)) : 0.f;
            const float p1 = (i + 1 < nval) ? exp2f(fmaf(__uint_as_float(v0[i + 1]), sc, -m_ref)) : 0.f;
            const float p2 = (32 + i < nval) ? exp2f(fmaf(__uint_as_float(v1[i]), sc, -m_ref)) : 0.f;
            const float p3 = (33 + i < nval) ? exp2f(fmaf(__uint_as_float(v1[i + 1]), sc, -m_ref)) : 0.f;
            rs += (p0 + p1) + (p2 + p3);
            pk[i >> 1] = pack_bf16(p0, p1);
            pk[16 + (i >> 1)] = pack_bf16(p2, p3);
          }
        }
        }
        if (tslot > 0) MDB_ATRACE(tslot, it, 4);
        tmem_st_32x16(tmem_s, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));
        tmem_st_32x16(tmem_s + 16, *reinterpret_cast<const uint32_t(*)[16]>(&pk[16]));
        tmem_st_wait();
        if (tslot > 0) MDB_ATRACE(tslot, it, 5);
        l += rs;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (tslot > 0) MDB_ATRACE(tslot, it, 6);
      }
      // ---- end of the KV set: combine the two split-KV halves, normalise, store
      mbar_wait(o_full, static_cast<uint32_t>(set) & 1u);
      tc_fence_after();
      xm[hlf * ATT_BM + row] = m_ref;
      xl[hlf * ATT_BM + row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float m_other = xm[(hlf ^ 1) * ATT_BM + row], l_other = xl[(hlf ^ 1) * ATT_BM + row];
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // both read before the next set overwrites
      const float m_tot = fmaxf(m_ref, m_other);                 // finite: key 0 of tile 0 is always valid
      const float w_me = exp2f(m_ref - m_tot), w_other = exp2f(m_other - m_tot);
      const float inv = 1.0f / (l * w_me + l_other * w_other);
      const float wa = (hlf ? w_other : w_me) * inv, wb = (hlf ? w_me : w_other) * inv;
      const uint32_t tmem_oa = lane_base + Cfg::OA_COL + hlf * HW, tmem_ob = lane_base + Cfg::OB_COL + hlf * HW;
#pragma unroll
      for (int c = 0; c < HW; c += 8) {
        if (c >= nvalid) break;
        uint32_t oa[8], ob[8];
        tmem_ld_32x8(tmem_oa + c, oa);
        tmem_ld_32x8(tmem_ob + c, ob);
        tmem_ld_wait();
        if (qrow < p.lq) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(oa[i]) * wa + __uint_as_float(ob[i]) * wb;
          if (set > 0 && !multi_q) {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
    }
  }

#undef MDB_ATRACE
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mdb
