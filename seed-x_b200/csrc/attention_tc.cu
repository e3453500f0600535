// seedx-b200: tcgen05 flash attention (forward) for sm_100a.
//
//   O[b,h,i,:] = softmax_j( scale * Q[b,h,i,:].K[b,h,j,:] (+causal) ) V[b,h,j,:]         fp16 in/out, fp32 softmax/accumulate
//
// One work item = 128 query rows of one (batch, head); CTAs are persistent (one per SM) and walk the item list.  Warp 0: TMA producer (Q once, K/V tiles of 128 keys through two mbarrier
// rings).  Warp 1: single-thread MMA issuer: S_j = Q K_j^T into a double-buffered TMEM accumulator (2 x 128 columns), then
// O += P_j V_j with P_j read back as the TMEM A-operand and V_j as an MN-major shared-memory B-operand.  Warps 2..9: softmax —
// two threads per query row (TMEM lane), each owning 64 of the 128 score columns, so a row max is one shared-memory exchange and
// needs no shuffles; P_j (fp16) overwrites the first 64 columns of S_j; the O accumulator (TMEM) is rescaled only when a row
// maximum moved.  QK^T of tile j+1 overlaps the softmax of tile j.
//
// Replaces the same reference call sites as seedx_attention_f16 (include/seedx.h) for head dims <= 128 and long sequences.
#include <cstdlib>
#include "common.cuh"
#include "../../include/seedx.h"

namespace seedx {
void count_launch();

struct FaParams {
  __half* o;
  long long o_sb, o_sh, o_ss;
  int sq, sk, d;
  float scale_log2;
  int causal, q_batched;
  int m_tiles, heads, items;   // work items = m_tiles * heads * batch, q-tile fastest
};

// one value in (2 * POLY_EVERY) takes the polynomial exp2; 0 = MUFU only
#ifndef SEEDX_FA_POLY_EVERY
#define SEEDX_FA_POLY_EVERY 4
#endif
constexpr int POLY_EVERY = SEEDX_FA_POLY_EVERY;

// 2^x for x <= ~100 on the FMA/ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (Remez on
// the relative error: 7.5e-5); 2^n by adding n to the exponent field.  Inputs below -126 (masked scores are -inf) give ~1e-38 -> 0 in fp16.
SEEDX_DEVINL float exp2_poly3(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t now hold round(x)
  const float f = x - (t - 12582912.0f);
  float q = fmaf(0.0551716685f, f, 0.2426111251f);
  q = fmaf(q, f, 0.6932609677f);
  q = fmaf(q, f, 0.9999280572f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}

// BN_ = keys per K/V tile: 128 for long key sequences; 64 for the UNet cross-attention, whose 64 context tokens are ONE tile (one QK^T MMA
// group and one PV group per work item, no online-softmax loop), so nothing of the tile is padding.
template <int D, int BN_ = 128>
struct FaCfg {
  static constexpr int BM = 128, BN = BN_;
  static constexpr int HALVES = D / 64;
  static constexpr int HALF_BYTES = 128 * 64 * 2;           // one [128 rows x 64 fp16] swizzled tile (Q)
  static constexpr int TILE_BYTES = HALVES * HALF_BYTES;    // Q tile
  static constexpr int KV_HALF_BYTES = BN * 64 * 2;         // one [BN keys x 64 fp16] swizzled tile
  static constexpr int KV_TILE_BYTES = HALVES * KV_HALF_BYTES;
  static constexpr int KV_STAGES = (BN == 64) ? 4 : ((D == 64) ? 3 : 2);
  static constexpr int SMEM_BYTES = TILE_BYTES * 2 + KV_TILE_BYTES * 2 * KV_STAGES + 1024 + 256;
  static constexpr uint32_t S_COL0 = 0, S_COL1 = 128, O_COL0 = 256, O_COL1 = 256 + D;
};

// Persistent: grid = min(items, SMs); CTA c walks items c, c+grid, ...  Q tiles and O accumulators are double-buffered so the
// TMA loads, the first QK^T of the next item and the output store of the previous one overlap; the S/P buffers, the K/V rings and
// their barrier phases run on one global tile counter `g` straight through item boundaries.
// NS = softmax threads per query row (2 or 4): 4*NS softmax warps.  NS = 4 puts four warps on every scheduler, which is what lets
// the MUFU (exp2), FMA and TMEM-load latencies of different warps overlap: with NS = 2 the kernel ran at ~45% issue utilisation.
template <int D, int NS, int BN_>
__global__ void __launch_bounds__(64 + 128 * NS, 1)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const FaParams p) {
  using Cfg = FaCfg<D, BN_>;
  constexpr int KS = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ0 = smem_base;
  const uint32_t sK0 = sQ0 + 2 * Cfg::TILE_BYTES;
  const uint32_t sV0 = sK0 + KS * Cfg::KV_TILE_BYTES;
  const uint32_t bar = sV0 + KS * Cfg::KV_TILE_BYTES;
  // barriers: k_full[KS], k_empty[KS], v_full[KS], v_empty[KS], then pairs: q_full, q_empty, s_full, p_full, pv_done, o_free
  auto k_full = [&](int s) { return bar + 8u * (s); };
  auto k_empty = [&](int s) { return bar + 8u * (KS + s); };
  auto v_full = [&](int s) { return bar + 8u * (2 * KS + s); };
  auto v_empty = [&](int s) { return bar + 8u * (3 * KS + s); };
  auto q_full = [&](int s) { return bar + 8u * (4 * KS + s); };
  auto q_empty = [&](int s) { return bar + 8u * (4 * KS + 2 + s); };
  auto s_full = [&](int s) { return bar + 8u * (4 * KS + 4 + s); };
  auto p_full = [&](int s) { return bar + 8u * (4 * KS + 6 + s); };
  auto pv_done = [&](int s) { return bar + 8u * (4 * KS + 8 + s); };
  auto o_free = [&](int s) { return bar + 8u * (4 * KS + 10 + s); };
  const uint32_t tmem_slot = bar + 8u * (4 * KS + 12);

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
      mbar_init(p_full(s), 128 * NS);
      mbar_init(pv_done(s), 1);
      mbar_init(o_free(s), 128 * NS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();      // prologue above overlapped the previous kernel's tail (programmatic dependent launch)
  pdl_trigger();

  const int causal_off = p.sk - p.sq;
  const int kv_tiles = (p.sk + Cfg::BN - 1) / Cfg::BN;
  auto tiles_of = [&](int m0) {
    int n = kv_tiles;
    if (p.causal) {
      const int lim = (m0 + Cfg::BM - 1 + causal_off) / Cfg::BN + 1;
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
        const int mt = it % p.m_tiles, hb = it / p.m_tiles;
        const int h = hb % p.heads, b = hb / p.heads;
        const int m0 = mt * Cfg::BM;
        const int qi = n & 1;
        mbar_wait(q_empty(qi), (uint32_t)(((n >> 1) & 1) ^ 1));
        mbar_expect_tx(q_full(qi), Cfg::TILE_BYTES);
#pragma unroll
        for (int hf = 0; hf < Cfg::HALVES; ++hf)
          tma_load_4d(sQ0 + qi * Cfg::TILE_BYTES + hf * Cfg::HALF_BYTES, &tmQ, q_full(qi), hf * 64, m0, h, p.q_batched ? b : 0);
        const int n_tiles = tiles_of(m0);
        for (int j = 0; j < n_tiles; ++j) {
          mbar_wait(k_empty(st), ph ^ 1u);
          mbar_expect_tx(k_full(st), Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int hf = 0; hf < Cfg::HALVES; ++hf)
            tma_load_4d(sK0 + st * Cfg::KV_TILE_BYTES + hf * Cfg::KV_HALF_BYTES, &tmK, k_full(st), hf * 64, j * Cfg::BN, h, b);
          mbar_wait(v_empty(st), ph ^ 1u);
          mbar_expect_tx(v_full(st), Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int hf = 0; hf < Cfg::HALVES; ++hf)
            tma_load_4d(sV0 + st * Cfg::KV_TILE_BYTES + hf * Cfg::KV_HALF_BYTES, &tmV, v_full(st), hf * 64, j * Cfg::BN, h, b);
          if (++st == KS) st = 0, ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, Cfg::BN);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, D) | (1u << 16);  // B operand MN-major
      int kst = 0, vst = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t g = 0;                 // global tile counter
      bool pend = false;              // a PV whose P is still being produced: issued after the next QK^T so the two overlap
      uint32_t pend_g = 0, pend_tO = 0;
      bool pend_first = false;
      auto flush_pv = [&]() {
        mbar_wait(p_full(pend_g & 1), (pend_g >> 1) & 1u);
        mbar_wait(v_full(vst), vph);
        tc_fence_after();
        const uint32_t tP = tmem_base + ((pend_g & 1) ? Cfg::S_COL1 : Cfg::S_COL0);
        const uint32_t vb = sV0 + vst * Cfg::KV_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < Cfg::BN / 16; ++kk)  // 16 keys per MMA: A advances 8 TMEM columns, B 16 rows of 128 B
          umma_f16_ts(pend_tO, tP + (uint32_t)(kk * 8), umma_desc_mn_sw128(vb + kk * 2048, Cfg::KV_HALF_BYTES), idesc_pv,
                      !(pend_first && kk == 0));
        umma_commit(v_empty(vst));
        umma_commit(pv_done(pend_g & 1));
        if (++vst == KS) vst = 0, vph ^= 1u;
      };
      int n = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++n) {
        const int m0 = (it % p.m_tiles) * Cfg::BM;
        const int n_tiles = tiles_of(m0);
        const int qi = n & 1;
        const uint32_t sQ = sQ0 + qi * Cfg::TILE_BYTES;
        mbar_wait(q_full(qi), (uint32_t)((n >> 1) & 1));
        mbar_wait(o_free(qi), (uint32_t)(((n >> 1) & 1) ^ 1));   // epilogue of item n-2 has drained this O accumulator
        for (int j = 0; j < n_tiles; ++j, ++g) {
          mbar_wait(k_full(kst), kph);
          tc_fence_after();
          const uint32_t tS = tmem_base + ((g & 1) ? Cfg::S_COL1 : Cfg::S_COL0);
          const uint32_t kb = sK0 + kst * Cfg::KV_TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t qoff = (uint32_t)((ks >> 2) * Cfg::HALF_BYTES + (ks & 3) * 32);
            const uint32_t koff = (uint32_t)((ks >> 2) * Cfg::KV_HALF_BYTES + (ks & 3) * 32);
            umma_f16(tS, umma_desc_k_sw128(sQ + qoff), umma_desc_k_sw128(kb + koff), idesc_qk, ks != 0);
          }
          umma_commit(k_empty(kst));
          if (j == n_tiles - 1) umma_commit(q_empty(qi));
          umma_commit(s_full(g & 1));
          if (++kst == KS) kst = 0, kph ^= 1u;
          if (pend) flush_pv();
          pend = true, pend_g = g, pend_first = (j == 0);
          pend_tO = tmem_base + (qi ? Cfg::O_COL1 : Cfg::O_COL0);
        }
      }
      if (pend) flush_pv();
    }
  } else {
    // ------------------------------------------------------------ softmax / correction / epilogue
    // row = TMEM lane (warp % 4 selects the lane quarter); the NS threads of a row own CW = 128/NS score columns and D/NS O columns each
    constexpr int CW = Cfg::BN / NS;        // score columns per thread
    constexpr int OC = D / NS;              // O columns per thread
    constexpr int OCH = OC >= 32 ? 32 : 16; // O columns per TMEM load
    __shared__ float xch_max[2][NS][128];   // [tile parity][column slice][row]
    __shared__ float xch_sum[NS][128];
    const int quarter = warp & 3;
    const int slice = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t g = 0;
    int n = 0;
    for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++n) {
      const int mt = it % p.m_tiles, hb = it / p.m_tiles;
      const int h = hb % p.heads, b = hb / p.heads;
      const int m0 = mt * Cfg::BM;
      const int n_tiles = tiles_of(m0);
      const int qrow = m0 + row;
      const uint32_t tO = tmem_base + lane_addr + ((n & 1) ? Cfg::O_COL1 : Cfg::O_COL0) + (uint32_t)(slice * OC);
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const uint32_t tSb = tmem_base + lane_addr + ((g & 1) ? Cfg::S_COL1 : Cfg::S_COL0);
        mbar_wait(s_full(g & 1), (g >> 1) & 1u);
        tc_fence_after();
        float s[CW];   // raw q.k scores: the softmax scale is folded into the exp2 FFMA below
        {
          uint32_t v[CW / 32][32];
#pragma unroll
          for (int c = 0; c < CW / 32; ++c) tmem_ld32(tSb + (uint32_t)(slice * CW + c * 32), v[c]);   // all requested before the single wait
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < CW / 32; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(v[c][i]);
        }
        const int key0 = j * Cfg::BN + slice * CW;
        const bool need_mask = (j * Cfg::BN + Cfg::BN > p.sk) || (p.causal && (j * Cfg::BN + Cfg::BN - 1 > m0 + causal_off));
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const int key = key0 + i;
            if (key >= p.sk || (p.causal && key > qrow + causal_off)) s[i] = -INFINITY;
          }
        }
        // four independent max chains
        float mx[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int i = 4; i < CW; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mx[k] = fmaxf(mx[k], s[i + k]);
        }
        float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        xch_max[g & 1][slice][row] = m_tile;
        asm volatile("bar.sync 1, %0;\n" ::"n"(128 * NS) : "memory");        // the softmax warps only
#pragma unroll
        for (int k = 0; k < NS; ++k) m_tile = fmaxf(m_tile, xch_max[g & 1][k][row]);
        m_tile *= p.scale_log2;                                              // scale > 0 (checked on the host): max commutes
        // Lazy rescaling: the reference maximum m_run only moves when some row of this warp exceeds it by more than 2^8 (P then stays
        // <= 256, exact in fp16 range; sums and O are fp32).  Only then does the softmax have to wait for PV_{j-1} and touch O, so in
        // the steady state softmax_j does not depend on the tensor pipe at all and the MMA -> softmax -> MMA round trip disappears.
        const bool grow = m_tile > m_run + 8.0f;       // also true for j == 0 (m_run = -inf) unless the whole row is masked
        float alpha = 1.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = fmaxf(m_run, m_tile);
          const float m_ref = (m_new == -INFINITY) ? 0.f : m_new;
          alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run - m_ref);
          if (j > 0) {
            mbar_wait(pv_done((g - 1) & 1), ((g - 1) >> 1) & 1u);  // O is quiescent: PV_{j-1} retired
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < OC / OCH; ++c) {
              uint32_t v[OCH];
              tmem_ld_n<OCH>(tO + (uint32_t)(c * OCH), v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < OCH; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_n<OCH>(tO + (uint32_t)(c * OCH), v);
            }
          }
          m_run = m_new;
        }
        const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
        // P_j = exp2(scale * S_j - m) as fp16 pairs: my CW scores -> CW/2 packed columns at [slice*CW/2, +CW/2) of the S_j buffer.
        // P occupies columns 0..63 of the buffer; every thread whose score columns lie below 64 has pulled them into registers before the
        // bar.sync above, and the columns >= 64 are never overwritten.
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};   // independent partial row sums
#pragma unroll
        for (int c = 0; c < CW / 32; ++c) {
          uint32_t v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // exp2 is the bottleneck of d=64 attention (16 MUFU/clk/SM vs 8192 tensor FLOP/clk/SM): every POLY_EVERY-th pair evaluates
            // one value on the FMA pipe instead (Cody-Waite split + degree-3 minimax, rel. error 7.5e-5, below the fp16 rounding of P)
            const float x0 = fmaf(s[c * 32 + 2 * i], p.scale_log2, -m_use), x1 = fmaf(s[c * 32 + 2 * i + 1], p.scale_log2, -m_use);
            const float a = fast_exp2(x0);
            const float c2 = (POLY_EVERY > 0 && (i % POLY_EVERY) == POLY_EVERY - 1) ? exp2_poly3(x1) : fast_exp2(x1);
            rs4[i & 3] += a + c2;
            __half2 hh = __floats2half2_rn(a, c2);
            v[i] = *(uint32_t*)&hh;
          }
          tmem_st16(tSb + (uint32_t)(slice * (CW / 2) + c * 16), v);
        }
        tmem_st_wait();
        l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
        tc_fence_before();
        mbar_arrive(p_full(g & 1));
      }
      // ---- epilogue: combine the row's partial sums, normalise and store my slice of the output columns
      xch_sum[slice][row] = l_run;
      asm volatile("bar.sync 1, %0;\n" ::"n"(128 * NS) : "memory");
      float l_tot = 0.f;
#pragma unroll
      for (int k = 0; k < NS; ++k) l_tot += xch_sum[k][row];
      const uint32_t gl = g - 1;
      mbar_wait(pv_done(gl & 1), (gl >> 1) & 1u);
      tc_fence_after();
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      __half* orow = p.o + (long long)b * p.o_sb + (long long)h * p.o_sh + (long long)qrow * p.o_ss + slice * OC;
#pragma unroll
      for (int c = 0; c < OC / OCH; ++c) {
        uint32_t v[OCH];
        tmem_ld_n<OCH>(tO + (uint32_t)(c * OCH), v);
        tmem_ld_wait();
        if (c == OC / OCH - 1) {            // accumulator drained into registers: the MMA warp may start item n+2 on it
          tc_fence_before();
          mbar_arrive(o_free(n & 1));
        }
        if (qrow < p.sq) {
#pragma unroll
          for (int i = 0; i < OCH; i += 8) {
            const int col = slice * OC + c * OCH + i;
            if (col + 8 <= p.d) {
              uint4 q;
              __half2* hh = (__half2*)&q;
#pragma unroll
              for (int t = 0; t < 4; ++t)
                hh[t] = __floats2half2_rn(__uint_as_float(v[i + 2 * t]) * inv, __uint_as_float(v[i + 2 * t + 1]) * inv);
              *(uint4*)(orow + c * OCH + i) = q;
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

#ifndef SEEDX_FA_NS
#define SEEDX_FA_NS 2
#endif
template <int D, int BN = 128>
static int launch_fa(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, FaParams p, int B, int H, cudaStream_t st) {
  using Cfg = FaCfg<D, BN>;
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    SEEDX_CUDA(cudaFuncSetAttribute(flash_attn_tc_kernel<D, SEEDX_FA_NS, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int dev = 0;
    SEEDX_CUDA(cudaGetDevice(&dev));
    SEEDX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr = true;
  }
  p.m_tiles = (p.sq + Cfg::BM - 1) / Cfg::BM;
  p.heads = H;
  const long long items = (long long)p.m_tiles * H * B;
  if (items > 0x7fffffffLL) return -1;
  p.items = (int)items;
  const int grid = p.items < sms ? p.items : sms;
  launch_k(flash_attn_tc_kernel<D, SEEDX_FA_NS, BN>, grid, 64 + 128 * SEEDX_FA_NS, Cfg::SMEM_BYTES, st, tq, tk, tv, p);
  count_launch();
  return check_cuda(cudaGetLastError(), "flash_attn_tc_kernel launch");
}

static int make_map(CUtensorMap* m, const void* ptr, int d, int s, int H, int B, long long ss, long long sh, long long sb, int box_rows = 128) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)s, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)ss * 2, (uint64_t)(H > 1 ? sh : ss * s) * 2, (uint64_t)(B > 1 ? sb : ss * s) * 2};
  uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
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

// returns -1 when the problem is not eligible for the tensor-memory kernel (caller falls back to the mma.sync kernel)
int attention_tc_try(const seedx_attn_args* a, cudaStream_t st) {
  // key sequences of at most 64 tokens with d <= 64 (UNet cross-attention over the 64 context tokens): one 64-key tile per work item;
  // 65..95 keys would fill under three quarters of a 128-key tile: the mma.sync kernel is faster there
  const bool short_kv = a->sk <= 64 && a->sk >= 16 && a->d <= 64 && !a->causal;
  if (a->d > 128 || a->d % 8 != 0 || a->sq < 128 || (a->sk < fa_min_sk() && !short_kv) || !(a->scale > 0.f)) return -1;
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
  if (make_map(&tq, a->q, a->d, a->sq, a->heads, qb ? a->batch : 1, a->q_stride_s, a->q_stride_h, a->q_stride_b)) return -1;
  const int kv_rows = short_kv ? 64 : 128;
  if (make_map(&tk, a->k, a->d, a->sk, a->heads, a->batch, a->k_stride_s, a->k_stride_h, a->k_stride_b, kv_rows)) return -1;
  if (make_map(&tv, a->v, a->d, a->sk, a->heads, a->batch, a->v_stride_s, a->v_stride_h, a->v_stride_b, kv_rows)) return -1;
  FaParams p;
  p.o = (__half*)a->o;
  p.o_sb = a->o_stride_b, p.o_sh = a->o_stride_h, p.o_ss = a->o_stride_s;
  p.sq = a->sq, p.sk = a->sk, p.d = a->d;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.causal = a->causal;
  p.q_batched = qb ? 1 : 0;
  if (short_kv) return launch_fa<64, 64>(tq, tk, tv, p, a->batch, a->heads, st);
  if (a->d <= 64) return launch_fa<64>(tq, tk, tv, p, a->batch, a->heads, st);
  return launch_fa<128>(tq, tk, tv, p, a->batch, a->heads, st);
}

}  // namespace seedx
