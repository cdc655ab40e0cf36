// tcgen05 fused attention forward, persistent form of attention_tc2's single-S variant (head dim <= 64).
// OPT-IN (MDB_ATTN_KERNEL=tc3): written after the round's GPU budget was spent, not yet run on a GPU.
//
// Why: fitting  T = waves * (F + tiles * t)  to the measured times of attention_tc2 at Lk = 1400 and 5300
// (profiles/attention_tc2_r1.experiments.txt) gives a fixed cost F of ~7 us per CTA (launch, TMEM allocation, barrier
// init, tensor-map fetch, first Q / K / V round trip, epilogue drain) against ~1.5 us per 128-key tile: with 11 tiles
// per CTA that is a quarter to a third of the kernel.  Here the grid is two CTAs per SM; each CTA allocates TMEM and
// initialises its barriers once and walks over work items (query tile, head, batch) with a stride of gridDim.x.  All
// pipelines run across item boundaries: the TMA warp prefetches the next item's Q (double-buffered) and first K/V tiles
// while the current item finishes, the MMA warp issues QK^T of the next item's first tile right after the last PV,
// and the softmax warps only drain O between items (as they already do between the two KV sets of the cross-view
// mode).  Barrier phases are derived from global counters: `git` (key-tile iterations), `gset` (KV sets), `item`.
// Everything else (split-KV halves, O accumulated in TMEM with lazy rescale, P aliased onto S) is attention_tc2's.
#pragma once
#include "attention_tc2.cuh"

namespace mdb {

struct AttnTc3Params {
  AttnTcParams a;
  int n_qtiles, heads, batch;  // work items = n_qtiles * heads * batch, item -> (qt fastest, head, b)
};

template <int D>
struct AttnTc3Cfg {
  static constexpr int KD = 1;
  static constexpr int D16 = (D + 15) / 16 * 16;
  static constexpr int TILE = ATT_BM * 128;
  static constexpr int STAGES = 2;
  static constexpr int SMEM_Q = 2 * TILE;  // double-buffered Q
  static constexpr int SMEM_KV = STAGES * 2 * TILE;
  static constexpr int kSmemBytes = SMEM_Q + SMEM_KV + 1024 /*alignment*/ + 256 /*barriers*/ + 4 * ATT_BM * 4 /*(m,l) exchange*/;
  static constexpr int OA_COL = 128, OB_COL = 192;
  static constexpr int kTmemCols = 256;
  static constexpr int HW = D16 / 2;
  static constexpr int kThreads = 64 + 8 * 32;
  static_assert(D <= 64, "one 64-wide chunk of the head dim");
};

template <int D>
__global__ void __launch_bounds__(AttnTc3Cfg<D>::kThreads, 2)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnTc3Params pp) {
  using Cfg = AttnTc3Cfg<D>;
  constexpr int D16 = Cfg::D16, STAGES = Cfg::STAGES, TILE = Cfg::TILE, HW = Cfg::HW;
  const AttnTcParams& p = pp.a;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;                      // [2][TILE]
  uint8_t* smK = smQ + Cfg::SMEM_Q;         // [STAGES][TILE]
  uint8_t* smV = smK + STAGES * TILE;       // [STAGES][TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smV + STAGES * TILE);
  uint64_t* q_full = bars;          // [2] TMA -> MMA
  uint64_t* q_empty = bars + 2;     // [2] MMA (last QK^T of the item complete) -> TMA
  uint64_t* kv_full = bars + 4;     // [2] TMA -> MMA
  uint64_t* kv_empty = bars + 6;    // [2] MMA (PV complete) -> TMA
  uint64_t* s_full = bars + 8;      // MMA -> softmax
  uint64_t* p_full = bars + 9;      // softmax (8 warps) -> MMA
  uint64_t* pv_done = bars + 10;    // MMA -> softmax, one phase per key-tile iteration
  uint64_t* o_full = bars + 11;     // MMA -> softmax, one phase per KV set
  uint64_t* o_free = bars + 12;     // softmax (8 warps) -> MMA, one phase per KV set
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* xm = reinterpret_cast<float*>(bars + 32);
  float* xl = xm + 2 * ATT_BM;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (p.lk + ATT_BN - 1) / ATT_BN;
  const int iters = ntiles * p.n_sets;  // key-tile iterations per work item
  const int n_items = pp.n_qtiles * pp.heads * pp.batch;
  // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_items = (static_cast<int>(blockIdx.x) < n_items) ? (n_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  auto item_coords = [&](int k, int& q0, int& head, int& b) {
    const int w = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
    q0 = (w % pp.n_qtiles) * ATT_BM;
    head = (w / pp.n_qtiles) % pp.heads;
    b = w / (pp.n_qtiles * pp.heads);
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
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
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int k = 0; k < my_items; ++k) {
        int q0, head, b;
        item_coords(k, q0, head, b);
        const int qb = k & 1;
        mbar_wait(&q_empty[qb], (static_cast<uint32_t>(k >> 1) & 1u) ^ 1u);  // QK^T of item k-2 no longer reads this buffer
        mbar_arrive_expect_tx(&q_full[qb], TILE);
        tma_load_4d(&tmQ, &q_full[qb], smQ + qb * TILE, 0, head, q0, b);
        for (int set = 0; set < p.n_sets; ++set) {
          const int kvb = p.kv_index ? p.kv_index[b * p.n_sets + set] : b;
          for (int j = 0; j < ntiles; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&kv_full[stage], 2 * TILE);
            tma_load_4d(&tmK, &kv_full[stage], smK + stage * TILE, 0, head, j * ATT_BN, kvb);
            tma_load_4d(&tmV, &kv_full[stage], smV + stage * TILE, 0, head, j * ATT_BN, kvb);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN);
    constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, D16) | (1u << 16);
    const int total = my_items * iters;
    // S of global iteration g (item g / iters, local iteration g % iters); K from ring stage g % STAGES
    auto issue_qk = [&](int g) {
      const int k = g / iters, lit = g - k * iters, qb = k & 1, st = g % STAGES;
      if (lit == 0) mbar_wait(&q_full[qb], static_cast<uint32_t>(k >> 1) & 1u);
      mbar_wait(&kv_full[st], static_cast<uint32_t>(g / STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D16 / 16; ++kk) {
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smQ + qb * TILE)) + 2 * kk;
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smK + st * TILE)) + 2 * kk;
          umma_bf16(tmem_base, adesc, bdesc, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        if (lit == iters - 1) umma_commit(&q_empty[qb]);  // last read of this item's Q
      }
      __syncwarp();
    };
    if (total > 0) issue_qk(0);
    int j = 0, gset = 0;
    for (int g = 0; g < total; ++g) {
      const int st = g % STAGES;
      mbar_wait(p_full, static_cast<uint32_t>(g) & 1u);
      if (j == 0 && gset > 0) mbar_wait(o_free, static_cast<uint32_t>(gset - 1) & 1u);  // previous set's O has been read
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          const int half = kk >> 2;
          const uint32_t tmem_a = tmem_base + static_cast<uint32_t>(half * 64 + 8 * (kk & 3));
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(half ? Cfg::OB_COL : Cfg::OA_COL);
          const uint64_t bdesc = make_sw128_mnmajor_desc(smem_u32(smV + st * TILE), TILE) + (2048u >> 4) * kk;
          umma_bf16_ts(tmem_d, tmem_a, bdesc, idesc_o, (j > 0 || (kk & 3) > 0) ? 1u : 0u);
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[st]);
        if (j == ntiles - 1) umma_commit(o_full);
      }
      __syncwarp();
      if (g + 1 < total) issue_qk(g + 1);  // may already belong to the next work item
      if (++j == ntiles) {
        j = 0;
        ++gset;
      }
    }
  } else {
    // =========================== softmax + epilogue (warps 2..9) ===========================
    const int q = warp & 3;
    const int hlf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tmem_s = lane_base + static_cast<uint32_t>(hlf * 64);
    const uint32_t tmem_acc = lane_base + (hlf ? Cfg::OB_COL : Cfg::OA_COL);
    constexpr int NVALID = (D - HW >= HW) ? HW : (D - HW);
    const int nvalid = hlf ? NVALID : HW;
    const float sc = p.scale_log2;
    int g = 0, gset = 0;
    for (int k = 0; k < my_items; ++k) {
      int q0, head, b;
      item_coords(k, q0, head, b);
      const int qrow = q0 + row;
      __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.lq + qrow) * p.ldo + head * D + hlf * HW;
      for (int set = 0; set < p.n_sets; ++set, ++gset) {
        float m_ref = -INFINITY, l = 0.f;
        for (int j = 0; j < ntiles; ++j, ++g) {
          mbar_wait(s_full, static_cast<uint32_t>(g) & 1u);
          tc_fence_after();
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(tmem_s, v0);
          tmem_ld_32x32(tmem_s + 32, v1);
          tmem_ld_wait();
          const int nval = p.lk - (j * ATT_BN + hlf * 64);
          const bool tail = nval < 64;
          float mx = -INFINITY;
          if (!tail) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
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
          const float m_new = mx * sc;
          const bool need = m_new > m_ref + ATT_LAZY_LOG2;
          if (__any_sync(0xffffffffu, need)) {
            if (j > 0) {
              // S_g ready implies PV_{g-1} was issued and completed before QK_g started (single S buffer, issue order)
              mbar_wait(pv_done, static_cast<uint32_t>(g - 1) & 1u);
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
          float rs = 0.f;
          uint32_t pk[32];
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
          tmem_st_32x16(tmem_s, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));
          tmem_st_32x16(tmem_s + 16, *reinterpret_cast<const uint32_t(*)[16]>(&pk[16]));
          tmem_st_wait();
          l += rs;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
        }
        // ---- end of the KV set: combine the two split-KV halves, normalise, store
        mbar_wait(o_full, static_cast<uint32_t>(gset) & 1u);
        tc_fence_after();
        xm[hlf * ATT_BM + row] = m_ref;
        xl[hlf * ATT_BM + row] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float m_other = xm[(hlf ^ 1) * ATT_BM + row], l_other = xl[(hlf ^ 1) * ATT_BM + row];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float m_tot = fmaxf(m_ref, m_other);
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
            if (set > 0) {
              const uint4 prev = *reinterpret_cast<const uint4*>(orow + c);
              const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 pf = __bfloat1622float2(ph[i]);
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
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mdb
