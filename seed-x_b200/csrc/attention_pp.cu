// seedx-b200: tcgen05 flash attention (forward), two-query-tile "ping-pong" schedule for sm_100a.
//
//   O[b,h,i,:] = softmax_j( scale * Q[b,h,i,:].K[b,h,j,:] (+causal) ) V[b,h,j,:]         fp16 in/out, fp32 softmax/accumulate
//
// One work item = 256 query rows (two 128-row tiles A and B) of one (batch, head); CTAs are persistent and walk the item list.
//   warp 0      TMA producer: the Q pair of an item, then K/V tiles of 128 keys through two mbarrier rings (each K/V tile serves both
//               query tiles: half the shared-memory fill traffic per FLOP of the one-tile kernel)
//   warp 1      single-thread MMA issuer: S_X = Q_X K_j^T into TMEM (one 128-column buffer per query tile), O_X += P_X V_j with P_X read
//               back as the TMEM A operand and V_j as an MN-major shared-memory B operand
//   warps 2-5   softmax group A, warps 6-9 softmax group B: ONE thread per query row (no cross-thread row reductions, no block
//               barriers).  While group A exponentiates tile j the tensor pipe runs PV_B(j-1) and QK_B(j), and vice versa, so the two
//               warps that share a scheduler are always in different phases and their MUFU / FMA / TMEM latencies overlap.
// P_j (fp16) overwrites the first 64 columns of S_X; O_X is rescaled only when a row maximum moved by more than 2^8 (lazy rescale).
// A share of the exponentials is evaluated on the FMA pipe (exp2_poly3) because d=64 attention is bound by the 16 MUFU/clk/SM rate.
//
// Replaces the same reference call sites as seedx_attention_f16 (include/seedx.h) for head dims <= 128 and sequences >= 128.
#include <cstdlib>
#include "common.cuh"
#include "../../include/seedx.h"

namespace seedx {
void count_launch();

struct PpParams {
  __half* o;
  long long o_sb, o_sh, o_ss;
  int sq, sk, d;
  float scale_log2;
  int causal, q_batched;
  int m_pairs, heads, items;   // work items = m_pairs * heads * batch, q-pair fastest
};

#ifndef SEEDX_PP_POLY_EVERY
#define SEEDX_PP_POLY_EVERY 4   // one value in (2 * POLY_EVERY) takes the polynomial exp2; 0 = MUFU only.  Measured (B200, S=4096, d=64):
                                // none 573, 1/8 627, 1/4 570, 1/2 537 TF/s -> the MUFU and FMA pipes balance at one value in eight
#endif

// 2^x on the FMA/ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (Remez on the relative
// error: 7.5e-5, below the fp16 rounding of P); 2^n by adding n to the exponent field.  Inputs below -126 (masked scores) give ~1e-38.
SEEDX_DEVINL float pp_exp2_poly3(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t now hold round(x)
  const float f = x - (t - 12582912.0f);
  float q = fmaf(0.0551716685f, f, 0.2426111251f);
  q = fmaf(q, f, 0.6932609677f);
  q = fmaf(q, f, 0.9999280572f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}

template <int D>
struct PpCfg {
  static constexpr int BM = 128, BN = 128;
  static constexpr int HALVES = D / 64;
  static constexpr int HALF_BYTES = 128 * 64 * 2;           // one [128 rows x 64 fp16] swizzled tile
  static constexpr int TILE_BYTES = HALVES * HALF_BYTES;    // one Q, K or V tile
  static constexpr int QBUF = (D == 64) ? 2 : 1;            // item-level buffering of the Q pair (shared memory permitting)
  static constexpr int KV_STAGES = (D == 64) ? 3 : 2;
  static constexpr int ONES_BYTES = 2048;                   // constant [16 keys x 64] MN-major block: column 0 = 1 (row sums on the tensor core)
  static constexpr int SMEM_BYTES = TILE_BYTES * (2 * QBUF + 2 * KV_STAGES) + ONES_BYTES + 1024 + 512;
  static constexpr uint32_t S_COL = 0, O_COL = 256;         // S_X at S_COL + 128 X, O_X at O_COL + OW X
};

// PE < 0 (d = 64 only): "H2" softmax.  (1) V is extended by a constant ones column (N = 80: the fifth 16-column group comes from a shared
// constant block through the descriptor's leading-dimension offset), so O[:, 64] accumulates the row sum of exactly the fp16 P values the tensor
// core multiplies — no FADD per score, and the lazy rescale treats it like any other O column.  (2) P = ex2.approx.f16x2 of the fp32-computed,
// fp16-rounded exponent pair: one MUFU instruction per two scores, result already packed for the TMEM A operand (no separate conversion).
// exponents are <= 8 (lazy rescale bound), |x| 2^-11 rounding gives P a relative error <= 0.7 * |x| * 2^-11, comparable to P's own fp16 rounding.
SEEDX_DEVINL uint32_t exp2_pair_f16(float x_lo, float x_hi) {
  uint32_t h, r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(h) : "f"(x_hi), "f"(x_lo));
  asm("ex2.approx.f16x2 %0, %1;\n" : "=r"(r) : "r"(h));
  return r;
}

template <int D, int PE>
__global__ void __launch_bounds__(320, 1)
flash_attn_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const PpParams p) {
  using Cfg = PpCfg<D>;
  constexpr int KS = Cfg::KV_STAGES, QB = Cfg::QBUF;
  constexpr bool H2 = PE < 0;                       // ones-column row sums + packed half2 exponentials
  static_assert(!H2 || D == 64, "the H2 softmax needs the spare TMEM columns of d = 64");
  constexpr int OW = H2 ? D + 16 : D;               // accumulator columns per query tile
  // d = 64 (not H2) leaves tensor-memory columns 384..511 free: P gets its own 64 columns per query tile instead of overwriting S.  S then
  // survives the exponential pass, which allows the ONE-PASS softmax below: unmasked tiles are exponentiated against the running reference
  // maximum (first tile: the maximum of its first 32 scores) while the largest P of the row is tracked on the packed fp16 results; only if a
  // row outgrew the reference by more than 2^8 is the tile redone the two-pass way from the intact scores.  Tensor memory reads at 64 B/clk per
  // sub-partition, so reading S once instead of twice, with the next 32 columns in flight under the current chunk's exponentials, is what counts.
  constexpr bool SEP_P = (D == 64) && !H2;
  constexpr uint32_t P_COL = 384;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ0 = smem_base;                                  // [QB][2] tiles
  const uint32_t sK0 = sQ0 + 2 * QB * Cfg::TILE_BYTES;
  const uint32_t sV0 = sK0 + KS * Cfg::TILE_BYTES;
  const uint32_t sOnes = sV0 + KS * Cfg::TILE_BYTES;               // 1024-aligned
  const uint32_t bar = sOnes + Cfg::ONES_BYTES;
  auto k_full = [&](int s) { return bar + 8u * (s); };
  auto k_empty = [&](int s) { return bar + 8u * (KS + s); };
  auto v_full = [&](int s) { return bar + 8u * (2 * KS + s); };
  auto v_empty = [&](int s) { return bar + 8u * (3 * KS + s); };
  auto q_full = [&](int s) { return bar + 8u * (4 * KS + s); };
  auto q_empty = [&](int s) { return bar + 8u * (4 * KS + 2 + s); };
  auto s_full = [&](int x) { return bar + 8u * (4 * KS + 4 + x); };    // per query tile X
  auto p_full = [&](int x) { return bar + 8u * (4 * KS + 6 + x); };
  auto pv_done = [&](int x) { return bar + 8u * (4 * KS + 8 + x); };
  auto o_free = [&](int x) { return bar + 8u * (4 * KS + 10 + x); };
  auto s_free = [&](int x) { return bar + 8u * (4 * KS + 12 + x); };   // SEP_P: the softmax has read S_X, the next QK^T may overwrite it
  const uint32_t tmem_slot = bar + 8u * (4 * KS + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < KS; ++s) {
      mbar_init(k_full(s), 1), mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1), mbar_init(v_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 128);
      mbar_init(pv_done(s), 1);
      mbar_init(o_free(s), 128);
      mbar_init(s_free(s), 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (H2 && threadIdx.x >= 64 && threadIdx.x < 64 + 128) {
    // ones block, 16 key rows x 128 B in the 128-byte-swizzled MN-major layout of a V tile: logical 16-byte chunk c of row r sits at physical
    // chunk c ^ (r & 7); element (r, 0) = 1.0, everything else 0
    const int t = threadIdx.x - 64, r = t >> 3, pc = t & 7;
    const uint32_t first = ((pc ^ (r & 7)) == 0) ? 0x00003C00u : 0u;
    sts128(sOnes + (uint32_t)(r * 128 + pc * 16), make_uint4(first, 0u, 0u, 0u));
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();      // prologue above overlapped the previous kernel's tail (programmatic dependent launch)
  pdl_trigger();

  const int causal_off = p.sk - p.sq;
  const int kv_tiles = (p.sk + Cfg::BN - 1) / Cfg::BN;
  auto tiles_of = [&](int m0) {   // KV tiles an item starting at query row m0 (256 rows) has to visit
    int n = kv_tiles;
    if (p.causal) {
      const int lim = (m0 + 2 * Cfg::BM - 1 + causal_off) / Cfg::BN + 1;
      if (lim < n) n = lim < 1 ? 1 : lim;
    }
    return n;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      int n = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++n) {
        const int mp = it % p.m_pairs, hb = it / p.m_pairs;
        const int h = hb % p.heads, b = hb / p.heads;
        const int m0 = mp * 2 * Cfg::BM;
        const int qb = n % QB;
        mbar_wait(q_empty(qb), (uint32_t)(((n / QB) & 1) ^ 1));
        mbar_expect_tx(q_full(qb), 2 * Cfg::TILE_BYTES);
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int hf = 0; hf < Cfg::HALVES; ++hf)
            tma_load_4d(sQ0 + (qb * 2 + x) * Cfg::TILE_BYTES + hf * Cfg::HALF_BYTES, &tmQ, q_full(qb), hf * 64, m0 + x * Cfg::BM, h,
                        p.q_batched ? b : 0);
        const int n_tiles = tiles_of(m0);
        for (int j = 0; j < n_tiles; ++j) {
          mbar_wait(k_empty(st), ph ^ 1u);
          mbar_expect_tx(k_full(st), Cfg::TILE_BYTES);
#pragma unroll
          for (int hf = 0; hf < Cfg::HALVES; ++hf)
            tma_load_4d(sK0 + st * Cfg::TILE_BYTES + hf * Cfg::HALF_BYTES, &tmK, k_full(st), hf * 64, j * Cfg::BN, h, b);
          mbar_wait(v_empty(st), ph ^ 1u);
          mbar_expect_tx(v_full(st), Cfg::TILE_BYTES);
#pragma unroll
          for (int hf = 0; hf < Cfg::HALVES; ++hf)
            tma_load_4d(sV0 + st * Cfg::TILE_BYTES + hf * Cfg::HALF_BYTES, &tmV, v_full(st), hf * 64, j * Cfg::BN, h, b);
          if (++st == KS) st = 0, ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, OW) | (1u << 16);  // B operand MN-major; H2: N = 80 (V | ones column block)
      int kst = 0, vst = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t g = 0;                 // tiles processed so far (same count for both query tiles)
      auto issue_qk = [&](int x, uint32_t sQ, uint32_t kb) {
        const uint32_t tS = tmem_base + Cfg::S_COL + (uint32_t)(x * 128);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * Cfg::HALF_BYTES + (ks & 3) * 32);
          umma_f16(tS, umma_desc_k_sw128(sQ + off), umma_desc_k_sw128(kb + off), idesc_qk, ks != 0);
        }
        umma_commit(s_full(x));
      };
      auto issue_pv = [&](int x, uint32_t vb, bool first) {
        const uint32_t tP = SEP_P ? tmem_base + P_COL + (uint32_t)(x * 64) : tmem_base + Cfg::S_COL + (uint32_t)(x * 128);
        const uint32_t tO = tmem_base + Cfg::O_COL + (uint32_t)(x * OW);
#pragma unroll
        for (int kk = 0; kk < Cfg::BN / 16; ++kk) {  // 16 keys per MMA: A advances 8 TMEM columns, B 16 rows of 128 B
          // H2: the second 64-element group of the MN dimension (columns 64..79) is the constant ones block, whatever keys this MMA covers
          const uint32_t lbo = H2 ? (sOnes - (vb + (uint32_t)(kk * 2048))) : (uint32_t)Cfg::HALF_BYTES;
          umma_f16_ts(tO, tP + (uint32_t)(kk * 8), umma_desc_mn_sw128(vb + kk * 2048, lbo), idesc_pv, !(first && kk == 0));
        }
        umma_commit(pv_done(x));
      };
      int n = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++n) {
        const int m0 = (it % p.m_pairs) * 2 * Cfg::BM;
        const int n_tiles = tiles_of(m0);
        const int qb = n % QB;
        const uint32_t sQA = sQ0 + (qb * 2) * Cfg::TILE_BYTES, sQB = sQA + Cfg::TILE_BYTES;
        mbar_wait(q_full(qb), (uint32_t)((n / QB) & 1));
        // S_A / S_B of the previous item were last read by its final PV_A / PV_B, issued earlier on the in-order tensor pipe
        mbar_wait(k_full(kst), kph);
        tc_fence_after();
        issue_qk(0, sQA, sK0 + kst * Cfg::TILE_BYTES);
        issue_qk(1, sQB, sK0 + kst * Cfg::TILE_BYTES);
        umma_commit(k_empty(kst));
        if (n_tiles == 1) umma_commit(q_empty(qb));
        if (++kst == KS) kst = 0, kph ^= 1u;
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const bool more = j + 1 < n_tiles;
          // ---- query tile A
          // P in its own columns (SEP_P): S_A is free as soon as the softmax has LOADED it (s_free, about three quarters into the tile's
          // exponentials), so the next QK^T goes ahead of PV on the in-order tensor pipe and S_A(j+1) is ready when the softmax comes back for it
          if (SEP_P && more) {
            mbar_wait(s_free(0), g & 1u);
            mbar_wait(k_full(kst), kph);
            tc_fence_after();
            issue_qk(0, sQA, sK0 + kst * Cfg::TILE_BYTES);
          }
          mbar_wait(p_full(0), g & 1u);
          mbar_wait(v_full(vst), vph);
          if (j == 0) mbar_wait(o_free(0), (uint32_t)((n & 1) ^ 1));   // epilogue of the previous item has drained O_A
          tc_fence_after();
          issue_pv(0, sV0 + vst * Cfg::TILE_BYTES, j == 0);
          if (!SEP_P && more) {
            mbar_wait(k_full(kst), kph);
            tc_fence_after();
            issue_qk(0, sQA, sK0 + kst * Cfg::TILE_BYTES);
          }
          // ---- query tile B
          if (SEP_P && more) {
            mbar_wait(s_free(1), g & 1u);
            tc_fence_after();
            issue_qk(1, sQB, sK0 + kst * Cfg::TILE_BYTES);
            umma_commit(k_empty(kst));
            if (j + 2 == n_tiles) umma_commit(q_empty(qb));   // last QK^T of this item issued: the Q pair may be overwritten
            if (++kst == KS) kst = 0, kph ^= 1u;
          }
          mbar_wait(p_full(1), g & 1u);
          if (j == 0) mbar_wait(o_free(1), (uint32_t)((n & 1) ^ 1));
          tc_fence_after();
          issue_pv(1, sV0 + vst * Cfg::TILE_BYTES, j == 0);
          umma_commit(v_empty(vst));
          if (++vst == KS) vst = 0, vph ^= 1u;
          if (!SEP_P && more) {
            issue_qk(1, sQB, sK0 + kst * Cfg::TILE_BYTES);
            umma_commit(k_empty(kst));
            if (j + 2 == n_tiles) umma_commit(q_empty(qb));   // last QK^T of this item issued: the Q pair may be overwritten
            if (++kst == KS) kst = 0, kph ^= 1u;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / correction / epilogue: one thread per query row
    const int x = (warp - 2) >> 2;                 // query tile of this softmax group
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + Cfg::S_COL + (uint32_t)(x * 128);
    const uint32_t tO = tmem_base + lane_addr + Cfg::O_COL + (uint32_t)(x * OW);
    const uint32_t tPw = SEP_P ? tmem_base + lane_addr + P_COL + (uint32_t)(x * 64) : tS;   // where this row's packed P goes
    uint32_t g = 0;
    int n = 0;
    for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++n) {
      const int mp = it % p.m_pairs, hb = it / p.m_pairs;
      const int h = hb % p.heads, b = hb / p.heads;
      const int m0 = mp * 2 * Cfg::BM;
      const int n_tiles = tiles_of(m0);
      const int qrow = m0 + x * Cfg::BM + row;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++g) {
        mbar_wait(s_full(x), g & 1u);
        tc_fence_after();
        const bool need_mask = (j * Cfg::BN + Cfg::BN > p.sk) || (p.causal && (j * Cfg::BN + Cfg::BN - 1 > m0 + x * Cfg::BM + causal_off));
        auto masked = [&](int key) { return key >= p.sk || (p.causal && key > qrow + causal_off); };
        if (SEP_P && !need_mask) {
          // ---- one-pass tile (see SEP_P above)
          uint32_t va[32], vb[32];
          tmem_ld32(tS, va);
          tmem_ld_wait();
          tmem_ld32(tS + 32u, vb);
          const bool first = (m_run == -INFINITY);
          if (first) {                                           // reference for the first tile: its first 32 scores
            float e0 = __uint_as_float(va[0]), e1 = __uint_as_float(va[1]), e2 = __uint_as_float(va[2]), e3 = __uint_as_float(va[3]);
#pragma unroll
            for (int i = 4; i < 32; i += 4) {
              e0 = fmaxf(e0, __uint_as_float(va[i])), e1 = fmaxf(e1, __uint_as_float(va[i + 1]));
              e2 = fmaxf(e2, __uint_as_float(va[i + 2])), e3 = fmaxf(e3, __uint_as_float(va[i + 3]));
            }
            m_run = fmaxf(fmaxf(e0, e1), fmaxf(e2, e3)) * p.scale_log2;
          }
          const float m_use1 = m_run;
          float rs4[4] = {0.f, 0.f, 0.f, 0.f};
          __half2 pm2 = __floats2half2_rn(0.f, 0.f);
          auto comp = [&](const uint32_t (&v)[32], uint32_t (&pk)[16]) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m_use1));
              const float xb0 = fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m_use1);
              const float b0 = (PE > 0 && (i % (PE > 0 ? PE : 1)) == PE - 1) ? pp_exp2_poly3(xb0) : fast_exp2(xb0);
              rs4[i & 3] += a0 + b0;
              const __half2 h0 = __floats2half2_rn(a0, b0);
              pm2 = __hmax2(pm2, h0);                            // an overflowed value is +inf in fp16 and wins the maximum
              pk[i] = *(const uint32_t*)&h0;
            }
          };
          uint32_t pka[16], pkb[16];
          comp(va, pka);
          tmem_ld_wait();
          tmem_ld32(tS + 64u, va);
          comp(vb, pkb);
          if (j > 0) {                                           // P_X(j-1) must have been consumed: QK^T(j) is issued AHEAD of PV(j-1), so the arrival
            mbar_wait(pv_done(x), (g - 1) & 1u);                 // of S(j) no longer implies it.  The first store is held back to the second chunk:
            tc_fence_after();                                    // PV(j-1) (~256 clk after QK^T(j)) has normally retired by then
          }
          tmem_st16(tPw, pka);
          tmem_st16(tPw + 16u, pkb);
          tmem_ld_wait();
          tmem_ld32(tS + 96u, vb);
          comp(va, pka);
          tmem_st16(tPw + 32u, pka);
          tmem_ld_wait();
          // S is in registers now.  Whether the tile stayed within 2^8 of its reference is known before the last chunk is exponentiated:
          // the largest P of chunks 0..2 and the largest raw score of chunk 3.  If so, S_X is released to the next QK^T right here.
          float e0 = __uint_as_float(vb[0]), e1 = __uint_as_float(vb[1]), e2 = __uint_as_float(vb[2]), e3 = __uint_as_float(vb[3]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            e0 = fmaxf(e0, __uint_as_float(vb[i])), e1 = fmaxf(e1, __uint_as_float(vb[i + 1]));
            e2 = fmaxf(e2, __uint_as_float(vb[i + 2])), e3 = fmaxf(e3, __uint_as_float(vb[i + 3]));
          }
          const float pm = fmaxf(__low2float(pm2), __high2float(pm2));
          const bool ovf = !(pm <= 256.0f) || fmaf(fmaxf(fmaxf(e0, e1), fmaxf(e2, e3)), p.scale_log2, -m_use1) > 8.0f;
          if (!__any_sync(0xffffffffu, ovf)) {                   // every row of this warp stays within 2^8 of its reference
            tc_fence_before();
            mbar_arrive(s_free(x));
            comp(vb, pkb);
            tmem_st16(tPw + 48u, pkb);
            tmem_st_wait();
            l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
            tc_fence_before();
            mbar_arrive(p_full(x));
            continue;
          }
          tmem_st_wait();
          // fall through: the P written above is overwritten below; S is intact (not released), and l_run / O have not been touched
        }
        // ---- pass 1: row maximum of the raw scores (the softmax scale is folded into the exp2 FFMA of pass 2)
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t v0[32], v1[32];
          tmem_ld32(tS + (uint32_t)(c2 * 64), v0);
          tmem_ld32(tS + (uint32_t)(c2 * 64 + 32), v1);
          tmem_ld_wait();
          if (need_mask) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (masked(j * Cfg::BN + c2 * 64 + i)) v0[i] = 0xff800000u;        // -inf
              if (masked(j * Cfg::BN + c2 * 64 + 32 + i)) v1[i] = 0xff800000u;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mx[k] = fmaxf(mx[k], fmaxf(__uint_as_float(v0[i + k]), __uint_as_float(v1[i + k])));
          }
        }
        const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2;   // scale > 0 (host check): max commutes
        // Lazy rescaling: the reference maximum only moves when some row of this warp exceeds it by more than 2^8 (P then stays <= 256,
        // fine in fp16; sums and O are fp32).  Only then does the softmax wait for PV_{j-1} and touch O.
        const bool grow = m_tile > m_run + 8.0f;       // also true for j == 0 (m_run = -inf) unless the whole row is masked
        float alpha = 1.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = fmaxf(m_run, m_tile);
          const float m_ref = (m_new == -INFINITY) ? 0.f : m_new;
          alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run - m_ref);
          if (j > 0) {
            mbar_wait(pv_done(x), (g - 1) & 1u);       // O_X is quiescent: PV_X(j-1) retired
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
              uint32_t v[32];
              tmem_ld32(tO + (uint32_t)(c * 32), v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(tO + (uint32_t)(c * 32), v);
            }
            if (H2) {                                  // the row-sum column lives in the accumulator and is rescaled with it
              uint32_t v[16];
              tmem_ld16(tO + (uint32_t)D, v);
              tmem_ld_wait();
              v[0] = __float_as_uint(__uint_as_float(v[0]) * alpha);
              tmem_st16(tO + (uint32_t)D, v);
            }
          }
          m_run = m_new;
        }
        const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
        if (SEP_P && j > 0) {                                    // as in the one-pass path: P_X(j-1) consumed before P_X(j) is written
          mbar_wait(pv_done(x), (g - 1) & 1u);
          tc_fence_after();
        }
        // ---- pass 2: P = exp2(scale * S - m) as fp16 pairs.  Chunk c re-reads score columns [32c, 32c+32) and writes packed columns
        // [16c, 16c+16): a chunk's P never lands on columns a later chunk still has to read.
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t v0[32], v1[32];
          tmem_ld32(tS + (uint32_t)(c2 * 64), v0);
          tmem_ld32(tS + (uint32_t)(c2 * 64 + 32), v1);
          tmem_ld_wait();
          if (need_mask) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (masked(j * Cfg::BN + c2 * 64 + i)) v0[i] = 0xff800000u;
              if (masked(j * Cfg::BN + c2 * 64 + 32 + i)) v1[i] = 0xff800000u;
            }
          }
          uint32_t pk0[16], pk1[16];
          if (H2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              pk0[i] = exp2_pair_f16(fmaf(__uint_as_float(v0[2 * i]), p.scale_log2, -m_use), fmaf(__uint_as_float(v0[2 * i + 1]), p.scale_log2, -m_use));
              pk1[i] = exp2_pair_f16(fmaf(__uint_as_float(v1[2 * i]), p.scale_log2, -m_use), fmaf(__uint_as_float(v1[2 * i + 1]), p.scale_log2, -m_use));
            }
          } else
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a0 = fast_exp2(fmaf(__uint_as_float(v0[2 * i]), p.scale_log2, -m_use));
            const float xb0 = fmaf(__uint_as_float(v0[2 * i + 1]), p.scale_log2, -m_use);
            const float b0 = (PE > 0 && (i % (PE > 0 ? PE : 1)) == PE - 1) ? pp_exp2_poly3(xb0) : fast_exp2(xb0);
            const float a1 = fast_exp2(fmaf(__uint_as_float(v1[2 * i]), p.scale_log2, -m_use));
            const float xb1 = fmaf(__uint_as_float(v1[2 * i + 1]), p.scale_log2, -m_use);
            const float b1 = (PE > 0 && (i % (PE > 0 ? PE : 1)) == PE - 1) ? pp_exp2_poly3(xb1) : fast_exp2(xb1);
            rs4[i & 3] += (a0 + b0) + (a1 + b1);
            __half2 h0 = __floats2half2_rn(a0, b0), h1 = __floats2half2_rn(a1, b1);
            pk0[i] = *(uint32_t*)&h0, pk1[i] = *(uint32_t*)&h1;
          }
          tmem_st16(tPw + (uint32_t)(c2 * 32), pk0);
          tmem_st16(tPw + (uint32_t)(c2 * 32 + 16), pk1);
        }
        tmem_st_wait();
        l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
        tc_fence_before();
        if (SEP_P) mbar_arrive(s_free(x));
        mbar_arrive(p_full(x));
      }
      // ---- epilogue: normalise and store this row (D contiguous fp16)
      const uint32_t gl = g - 1;
      mbar_wait(pv_done(x), gl & 1u);
      tc_fence_after();
      if (H2) {                             // row sum = accumulator column 64 (sum of the fp16 P values, rescaled with O)
        uint32_t v[16];
        tmem_ld16(tO + (uint32_t)D, v);
        tmem_ld_wait();
        l_run = __uint_as_float(v[0]);
      }
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __half* orow = p.o + (long long)b * p.o_sb + (long long)h * p.o_sh + (long long)qrow * p.o_ss;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tO + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (c == D / 32 - 1) {              // accumulator drained into registers: the MMA warp may start the next item on it
          tc_fence_before();
          mbar_arrive(o_free(x));
        }
        if (qrow < p.sq) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const int col = c * 32 + i;
            if (col + 8 <= p.d) {
              uint4 q;
              __half2* hh = (__half2*)&q;
#pragma unroll
              for (int t = 0; t < 4; ++t)
                hh[t] = __floats2half2_rn(__uint_as_float(v[i + 2 * t]) * inv, __uint_as_float(v[i + 2 * t + 1]) * inv);
              *(uint4*)(orow + col) = q;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int D, int PE>
static int launch_pp_pe(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, PpParams p, int B, int H, cudaStream_t st) {
  using Cfg = PpCfg<D>;
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    SEEDX_CUDA(cudaFuncSetAttribute(flash_attn_pp_kernel<D, PE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int dev = 0;
    SEEDX_CUDA(cudaGetDevice(&dev));
    SEEDX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr = true;
  }
  p.m_pairs = (p.sq + 2 * Cfg::BM - 1) / (2 * Cfg::BM);
  p.heads = H;
  const long long items = (long long)p.m_pairs * H * B;
  if (items > 0x7fffffffLL) return -1;
  p.items = (int)items;
  const int grid = p.items < sms ? p.items : sms;
  const cudaError_t e = launch_k(flash_attn_pp_kernel<D, PE>, grid, 320, Cfg::SMEM_BYTES, st, tq, tk, tv, p);
  count_launch();
  return check_cuda(e != cudaSuccess ? e : cudaGetLastError(), "flash_attn_pp_kernel launch");
}

// share of exponentials on the FMA pipe: 1 / (2 * PE); env SEEDX_PP_POLY_EVERY (0 = none, 1, 2, 4) overrides the default for experiments
template <int D>
static int launch_pp(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const PpParams& p, int B, int H, cudaStream_t st) {
  static int pe = -1;
  if (pe < 0) {
    const char* e = getenv("SEEDX_PP_POLY_EVERY");
    // -1 selects the H2 softmax (d = 64 only).  Measured on B200 (8 x 10 heads x 4096^2, d = 64): H2 571 TF/s vs 624 TF/s for the 1/8 polynomial
    // split — the N = 80 PV MMAs and the cvt + ex2.f16x2 pair cost more than the FADDs and MUFU slots they free — so it stays an experiment.
    // d = 64 with the one-pass softmax and the early release of S: the softmax no longer waits for the tensor pipe and its instruction count
    // decides — MUFU only 493 us, 1/8 polynomial 514 us, 1/4 554 us (8 x 10 x 4096^2); d = 128 keeps the 1/8 split
    pe = e ? atoi(e) : (D == 64 ? 0 : SEEDX_PP_POLY_EVERY);
  }
  if (D == 64 && pe < 0) return launch_pp_pe<64, -1>(tq, tk, tv, p, B, H, st);
  switch (pe) {
    case 0: return launch_pp_pe<D, 0>(tq, tk, tv, p, B, H, st);
    case 1: return launch_pp_pe<D, 1>(tq, tk, tv, p, B, H, st);
    case 4: return launch_pp_pe<D, 4>(tq, tk, tv, p, B, H, st);
    default: return launch_pp_pe<D, 2>(tq, tk, tv, p, B, H, st);
  }
}

static int make_map_pp(CUtensorMap* m, const void* ptr, int d, int s, int H, int B, long long ss, long long sh, long long sb) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)s, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)ss * 2, (uint64_t)(H > 1 ? sh : ss * s) * 2, (uint64_t)(B > 1 ? sb : ss * s) * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// shortest key sequence routed to the tcgen05 kernels (a 128-key tile is then partly padding); env SEEDX_FA_MIN_SK overrides for experiments
static int fa_min_sk() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEEDX_FA_MIN_SK");
    v = e ? atoi(e) : 96;
  }
  return v;
}

// returns -1 when the problem is not eligible (the caller falls back to the one-tile tcgen05 kernel, then to the mma.sync kernel)
int attention_pp_try(const seedx_attn_args* a, cudaStream_t st) {
  if (a->d > 128 || a->d % 8 != 0 || a->sq < 256 || a->sk < fa_min_sk() || !(a->scale > 0.f)) return -1;
  // causal problems waste the masked upper tiles of the second query tile, and few items balance badly over the SMs at 256 rows per
  // item: both cases are faster on the one-tile kernel (measured: causal S=2048 369 vs 485 TF/s; B=2 S=4096 432 vs 484 TF/s)
  if (a->causal) return -1;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long items = (long long)((a->sq + 255) / 256) * a->heads * a->batch;
    if (items < 3LL * sms) return -1;
  }
  if (a->o_stride_s % 8 || a->o_stride_h % 8 || a->o_stride_b % 8 || (uintptr_t)a->o % 16) return -1;
  const long long str[] = {a->q_stride_s, a->q_stride_h, a->q_stride_b, a->k_stride_s, a->k_stride_h, a->k_stride_b,
                           a->v_stride_s, a->v_stride_h, a->v_stride_b};
  for (long long s : str)
    if (s % 8 != 0 || s < 0) return -1;
  if (a->q_stride_s == 0 || a->k_stride_s == 0 || a->v_stride_s == 0) return -1;
  if ((a->heads > 1 && (a->q_stride_h == 0 || a->k_stride_h == 0 || a->v_stride_h == 0)) ||
      (a->batch > 1 && (a->k_stride_b == 0 || a->v_stride_b == 0)))
    return -1;
  const bool qb = !(a->batch > 1 && a->q_stride_b == 0);
  CUtensorMap tq, tk, tv;
  if (make_map_pp(&tq, a->q, a->d, a->sq, a->heads, qb ? a->batch : 1, a->q_stride_s, a->q_stride_h, a->q_stride_b)) return -1;
  if (make_map_pp(&tk, a->k, a->d, a->sk, a->heads, a->batch, a->k_stride_s, a->k_stride_h, a->k_stride_b)) return -1;
  if (make_map_pp(&tv, a->v, a->d, a->sk, a->heads, a->batch, a->v_stride_s, a->v_stride_h, a->v_stride_b)) return -1;
  PpParams p;
  p.o = (__half*)a->o;
  p.o_sb = a->o_stride_b, p.o_sh = a->o_stride_h, p.o_ss = a->o_stride_s;
  p.sq = a->sq, p.sk = a->sk, p.d = a->d;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.causal = a->causal;
  p.q_batched = qb ? 1 : 0;
  if (a->d <= 64) return launch_pp<64>(tq, tk, tv, p, a->batch, a->heads, st);
  return launch_pp<128>(tq, tk, tv, p, a->batch, a->heads, st);
}

}  // namespace seedx
