// seedx-b200: normalisation kernels (HBM-bound, fp32 statistics).
//
//  * seedx_layernorm  : LayerNorm / RMSNorm over the last dim, fp16|fp32 in -> fp16|fp32 out, optional second output
//                       out2 = y + add[row % add_rows]  (positional embedding add of the Resampler key path)
//      replaces nn.LayerNorm at src/models/tokenizer/qwen_visual.py:400,280-281,139-141,414,
//               src/models/detokenizer/resampler.py:55-56,279, diffusers BasicTransformerBlock norms,
//               LlamaRMSNorm (transformers) used at src/models/mllm/modeling_llama_xformer.py:95,258-259,443
//  * seedx_groupnorm_nhwc : GroupNorm(32) (+SiLU) on NHWC fp16 images, optional channel-concat of two sources
//      replaces diffusers ResnetBlock2D/Transformer2DModel/VAE GroupNorm+SiLU.
#include "common.cuh"
#include "../../include/seedx.h"

namespace seedx {

void count_launch();

// ------------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm.  TPR threads cooperate on one row (32 = a warp, 4 rows per CTA; 256 = a whole CTA); every thread keeps
// NV 4-element vectors of the row in registers (two-pass statistics on the cached row), 16-byte loads, 8/16-byte stores.
// ------------------------------------------------------------------------------------------------
constexpr int LN_THREADS = 256;
constexpr int LN_MAXPT = 32;

SEEDX_DEVINL float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.f;
  t = warp_sum(t);
  return t;
}

// a row is cached in registers in its STORAGE format (fp16 rows cost half the registers -> more resident CTAs to hide HBM latency)
template <typename T> struct Raw4;
template <> struct Raw4<float> { typedef float4 type; };
template <> struct Raw4<__half> { typedef uint2 type; };
SEEDX_DEVINL float4 ldraw(const float* p) { return *(const float4*)p; }
SEEDX_DEVINL uint2 ldraw(const __half* p) { return *(const uint2*)p; }
SEEDX_DEVINL float4 unpack4(float4 q) { return q; }
SEEDX_DEVINL float4 unpack4(uint2 q) {
  const float2 a = __half22float2(*(const __half2*)&q.x), b = __half22float2(*(const __half2*)&q.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
SEEDX_DEVINL float4 zero_raw(float4*) { return make_float4(0.f, 0.f, 0.f, 0.f); }
SEEDX_DEVINL uint2 zero_raw(uint2*) { return make_uint2(0u, 0u); }
SEEDX_DEVINL void st4(float* p, float4 v) { *(float4*)p = v; }
SEEDX_DEVINL void st4(__half* p, float4 v) {
  uint2 q;
  *(__half2*)&q.x = __floats2half2_rn(v.x, v.y);
  *(__half2*)&q.y = __floats2half2_rn(v.z, v.w);
  *(uint2*)p = q;
}

template <typename TI, typename TO, int TPR, int NV>
__global__ void __launch_bounds__(LN_THREADS)
layernorm_vec_kernel(const TI* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                     TO* __restrict__ out, long long ldo, TO* __restrict__ out2, const float* __restrict__ add, int add_rows,
                     long long rows, int cols, float eps, int rms) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[8];
  constexpr int RPB = LN_THREADS / TPR;  // rows per block
  const int sub = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const long long row = (long long)blockIdx.x * RPB + sub;
  const bool row_ok = row < rows;
  const TI* xr = x + (row_ok ? row : 0) * ldx;
  const int nvec = cols >> 2;
  typedef typename Raw4<TI>::type RawT;
  RawT raw[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = t + i * TPR;
    raw[i] = (c < nvec && row_ok) ? ldraw(xr + c * 4) : zero_raw((RawT*)nullptr);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 v = unpack4(raw[i]);
    s += v.x + v.y + v.z + v.w;
  }
  float mean = 0.f;
  if (!rms) mean = (TPR == 32 ? warp_sum(s) : block_sum(s, sh)) / (float)cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = t + i * TPR;
    if (c < nvec) {
      const float4 v = unpack4(raw[i]);
      const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float var = (TPR == 32 ? warp_sum(q) : block_sum(q, sh)) / (float)cols;
  const float rstd = rsqrtf(var + eps);
  if (!row_ok) return;
  TO* orow = out + row * ldo;
  TO* orow2 = out2 ? out2 + row * ldo : nullptr;
  const float* arow = add ? add + (long long)(row % add_rows) * cols : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = t + i * TPR;
    if (c < nvec) {
      const float4 v = unpack4(raw[i]);
      float4 y = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
      if (gamma) {
        const float4 g = *(const float4*)(gamma + c * 4);
        y.x *= g.x, y.y *= g.y, y.z *= g.z, y.w *= g.w;
      }
      if (beta) {
        const float4 bb = *(const float4*)(beta + c * 4);
        y.x += bb.x, y.y += bb.y, y.z += bb.z, y.w += bb.w;
      }
      st4(orow + c * 4, y);
      if (orow2) {
        const float4 aa = *(const float4*)(arow + c * 4);
        st4(orow2 + c * 4, make_float4(y.x + aa.x, y.y + aa.y, y.z + aa.z, y.w + aa.w));
      }
    }
  }
}

// Warp-persistent variant for narrow rows (cols <= 2048): gamma/beta are staged once per CTA in shared memory (they are 2-4x the bytes of an
// fp16 row and were re-read from L1/L2 for every row), each warp walks rows with a grid stride and requests row r+stride while it reduces
// and writes row r, so HBM latency is hidden by the warp's own next row instead of by occupancy.
template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(LN_THREADS)
layernorm_rows_kernel(const TI* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                      TO* __restrict__ out, long long ldo, TO* __restrict__ out2, const float* __restrict__ add, int add_rows,
                      long long rows, int cols, float eps, int rms) {
  extern __shared__ __align__(16) float gb[];  // gamma[cols] | beta[cols]
  float* sg = gb;
  float* sb = gb + cols;
  for (int c = threadIdx.x; c < cols; c += LN_THREADS) {   // parameters: staged before the dependency wait
    sg[c] = gamma ? gamma[c] : 1.f;
    sb[c] = beta ? beta[c] : 0.f;
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  typedef typename Raw4<TI>::type RawT;
  const int lane = threadIdx.x & 31;
  const int nvec = cols >> 2;
  const long long wstride = (long long)gridDim.x * (LN_THREADS / 32);
  long long row = (long long)blockIdx.x * (LN_THREADS / 32) + (threadIdx.x >> 5);
  RawT cur[NV], nxt[NV];
  auto load_row = [&](RawT (&dst)[NV], long long r) {
    const TI* xr = x + r * ldx;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      dst[i] = (c < nvec) ? ldraw(xr + c * 4) : zero_raw((RawT*)nullptr);
    }
  };
  if (row < rows) load_row(cur, row);
  for (; row < rows; row += wstride) {
    const long long rn = row + wstride;
    if (rn < rows) load_row(nxt, rn);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 v = unpack4(cur[i]);
      s += v.x + v.y + v.z + v.w;
    }
    const float mean = rms ? 0.f : warp_sum(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
        const float4 v = unpack4(cur[i]);
        const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
    TO* orow = out + row * ldo;
    TO* orow2 = out2 ? out2 + row * ldo : nullptr;
    const float* arow = add ? add + (long long)(row % add_rows) * cols : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 v = unpack4(cur[i]);
        const float4 g = *(const float4*)(sg + c * 4);
        const float4 bb = *(const float4*)(sb + c * 4);
        const float4 y = make_float4((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y, (v.z - mean) * rstd * g.z + bb.z,
                                     (v.w - mean) * rstd * g.w + bb.w);
        st4(orow + c * 4, y);
        if (orow2) {
          const float4 aa = *(const float4*)(arow + c * 4);
          st4(orow2 + c * 4, make_float4(y.x + aa.x, y.y + aa.y, y.z + aa.z, y.w + aa.w));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) cur[i] = nxt[i];
  }
}

// Statistics only: (mean, rstd) per row of an fp16 matrix, same warp-persistent walk and two-pass arithmetic as layernorm_rows_kernel but
// nothing normalised is written — the consumer GEMM applies the normalisation in its epilogue (seedx_gemm_args.ln_stats).
template <int NV>
__global__ void __launch_bounds__(LN_THREADS)
row_stats_kernel(const __half* __restrict__ x, long long ldx, long long rows, int cols, float eps, float2* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int nvec = cols >> 3;                       // 16-byte vectors of 8 fp16
  const long long wstride = (long long)gridDim.x * (LN_THREADS / 32);
  long long row = (long long)blockIdx.x * (LN_THREADS / 32) + (threadIdx.x >> 5);
  uint4 cur[NV], nxt[NV];
  auto load_row = [&](uint4 (&dst)[NV], long long r) {
    const __half* xr = x + r * ldx;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      dst[i] = (c < nvec) ? *(const uint4*)(xr + c * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if (row < rows) load_row(cur, row);
  for (; row < rows; row += wstride) {
    const long long rn = row + wstride;
    if (rn < rows) load_row(nxt, rn);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const __half2* h = (const __half2*)&cur[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h[j]);
        s += t.x + t.y;
      }
    }
    const float mean = warp_sum(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
        const __half2* h = (const __half2*)&cur[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(h[j]);
          const float a = t.x - mean, b = t.y - mean;
          q += a * a + b * b;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
    if (lane == 0) stats[row] = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) cur[i] = nxt[i];
  }
}

// (mean, rstd) per row from the row partials a producing GEMM epilogue wrote (seedx_gemm_args.row_part: [parts][rows] (sum, sum of squares) over
// 32-column chunks): one thread per row, chunk order, coalesced 8-byte loads — parts * 8 bytes per row instead of the row itself.
__global__ void __launch_bounds__(256)
row_finalize_kernel(const float2* __restrict__ parts, int nparts, long long rows, float inv_cols, float eps, float2* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  // 64 rows per CTA, 4 threads per row (thread = row + 64 * sub): sub s adds the chunks k = s, s + 4, ... in order with all its loads in flight,
  // the four sub-sums are combined in a fixed order -> deterministic, 4x the CTAs and a quarter of the dependent-load chain of one thread per row
  __shared__ float2 sh[4][64];
  const int r = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const long long row = (long long)blockIdx.x * 64 + r;
  float s1 = 0.f, s2 = 0.f;
  if (row < rows) {
#pragma unroll 4
    for (int k = sub; k < nparts; k += 4) {
      const float2 q = __ldcg(parts + (long long)k * rows + row);
      s1 += q.x, s2 += q.y;
    }
  }
  sh[sub][r] = make_float2(s1, s2);
  __syncthreads();
  if (sub == 0 && row < rows) {
    const float a = (sh[0][r].x + sh[1][r].x) + (sh[2][r].x + sh[3][r].x), b = (sh[0][r].y + sh[1][r].y) + (sh[2][r].y + sh[3][r].y);
    const float mean = a * inv_cols;
    stats[row] = make_float2(mean, rsqrtf(fmaxf(b * inv_cols - mean * mean, 0.f) + eps));
  }
}

// scalar fallback (cols not a multiple of 4 or unaligned rows): one CTA per row
template <typename TI, typename TO>
__global__ void __launch_bounds__(LN_THREADS)
layernorm_kernel(const TI* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                 TO* __restrict__ out, long long ldo, TO* __restrict__ out2, const float* __restrict__ add, int add_rows,
                 int cols, float eps, int rms) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[8];
  const long long row = blockIdx.x;
  const TI* xr = x + row * ldx;
  float v[LN_MAXPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXPT; ++i) {
    const int c = threadIdx.x + i * LN_THREADS;
    v[i] = (c < cols) ? (float)xr[c] : 0.f;
    s += v[i];
  }
  float mean = 0.f;
  if (!rms) mean = block_sum(s, sh) / (float)cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXPT; ++i) {
    const int c = threadIdx.x + i * LN_THREADS;
    const float d = (c < cols) ? (v[i] - mean) : 0.f;
    q += d * d;
  }
  const float var = block_sum(q, sh) / (float)cols;
  const float rstd = rsqrtf(var + eps);
  TO* orow = out + row * ldo;
  TO* orow2 = out2 ? out2 + row * ldo : nullptr;
  const float* arow = add ? add + (long long)(row % add_rows) * cols : nullptr;
#pragma unroll
  for (int i = 0; i < LN_MAXPT; ++i) {
    const int c = threadIdx.x + i * LN_THREADS;
    if (c < cols) {
      float y = (v[i] - mean) * rstd;
      if (gamma) y *= gamma[c];
      if (beta) y += beta[c];
      orow[c] = (TO)y;
      if (orow2) orow2[c] = (TO)(y + arow[c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm on NHWC fp16, 2 passes: (1) per-(image, group) sum / sum-of-squares, (2) apply per-channel scale/shift (+SiLU).
// Input may be the channel concatenation of two tensors.  The statistics are reduced in a FIXED order at every level (thread ->
// shared-memory slots -> shuffle tree -> per-CTA partial in global memory -> the last CTA of an image sums the partials in index
// order, fp64), so the result is bit-reproducible run to run: no floating-point atomics anywhere.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_stats_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2, int hw, int groups, int pix_per_block,
                float2* __restrict__ stats /*[n][groups] (mean, rstd)*/, float* __restrict__ partials /*[n][blocks][groups][2]*/,
                unsigned* __restrict__ tickets /*[n], zeroed by the launcher*/, int tpp /*threads per (group, stat) pair*/, float eps) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float gsm[];  // [lanes][C][2] per-channel partials of this CTA
  __shared__ bool is_last;
  const int n = blockIdx.y;
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  const int vpp = C >> 3;  // 16-byte channel vectors per pixel (c1, c2 multiples of 8)
  // thread -> (pixel lane, channel vector): a thread keeps ONE channel vector and strides over pixels, so the 8 per-channel
  // partial sums live in registers and reach shared memory once per thread, each in its own slot
  const bool wide = vpp >= 256;
  const int lanes = wide ? 1 : 256 / vpp;
  const int p_lane = wide ? 0 : (int)threadIdx.x / vpp;
  const int cv_step = wide ? 256 : vpp;
  if (p_lane < lanes) {
    for (int cv = wide ? (int)threadIdx.x : (int)threadIdx.x % vpp; cv < vpp; cv += cv_step) {
      const int ch = cv * 8;
      const bool first = ch < c1;
      const __half* base = first ? x1 + (long long)n * hw * c1 + ch : x2 + (long long)n * hw * c2 + (ch - c1);
      const int cstride = first ? c1 : c2;
      float s[8], ss[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
      for (int pix = p0 + p_lane; pix < p1; pix += lanes) {
        const uint4 q = *(const uint4*)(base + (long long)pix * cstride);
        const __half2* h = (const __half2*)&q;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(h[j]);
          s[2 * j] += t.x, ss[2 * j] += t.x * t.x;
          s[2 * j + 1] += t.y, ss[2 * j + 1] += t.y * t.y;
        }
      }
      float4* dst = (float4*)(gsm + ((long long)p_lane * C + ch) * 2);
      dst[0] = make_float4(s[0], ss[0], s[1], ss[1]);
      dst[1] = make_float4(s[2], ss[2], s[3], ss[3]);
      dst[2] = make_float4(s[4], ss[4], s[5], ss[5]);
      dst[3] = make_float4(s[6], ss[6], s[7], ss[7]);
    }
  }
  __syncthreads();
  // (group, stat) pair p is owned by `tpp` consecutive threads (tpp: power of two <= 32): strided partial sums, then an xor tree
  const int pairs = groups * 2;
  const int pair = (int)threadIdx.x / tpp, sub = (int)threadIdx.x % tpp;
  const bool owner = pair < pairs;
  const int g = owner ? pair >> 1 : 0, stat = pair & 1;
  float acc = 0.f;
  if (owner)
    for (int idx = sub; idx < cpg * lanes; idx += tpp) acc += gsm[((long long)(idx / cpg) * C + g * cpg + idx % cpg) * 2 + stat];
  for (int o = tpp >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const int nblk = gridDim.x;
  if (owner && sub == 0) partials[((long long)n * nblk + blockIdx.x) * pairs + pair] = acc;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&tickets[n], 1u) == (unsigned)(nblk - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double tot = 0.0;
  if (owner)
    for (int b = sub; b < nblk; b += tpp) tot += (double)__ldcg(&partials[((long long)n * nblk + b) * pairs + pair]);
  for (int o = tpp >> 1; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  __shared__ double fin[256];
  if (owner && sub == 0) fin[pair] = tot;
  __syncthreads();
  if ((int)threadIdx.x < groups) {      // (mean, rstd) per group, computed once here instead of once per CTA of the apply pass
    const double cnt = (double)hw * (double)cpg;
    const double m = fin[2 * threadIdx.x] / cnt;
    double var = fin[2 * threadIdx.x + 1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    stats[(long long)n * groups + threadIdx.x] = make_float2((float)m, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

// Statistics from the column partials the producing GEMM / conv epilogue wrote (seedx_gemm_args.col_part: per 32-row slab and channel,
// (sum, sum of squares) of the stored fp16 values).  One CTA per (group, image): thread t adds the (slab, channel) pairs t, t + 256, ... of its
// group in that order (fp64), then a fixed shuffle / shared-memory tree: bit-reproducible like gn_stats_kernel, but it reads 1/8 of the bytes
// of the tensor instead of all of them and needs no atomics or tickets.
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float2* __restrict__ part1, int c1, const float2* __restrict__ part2, int c2, int slabs_per_img, int groups, int hw,
                   float eps, float2* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  const int g = blockIdx.x, n = blockIdx.y;
  const int C = c1 + c2, cpg = C / groups;
  const int total = slabs_per_img * cpg;
  // a thread adds its (slab, channel) pairs in index order in fp32 (each is already a sum over 32 rows; a thread sees total / 256 of them),
  // the cross-thread tree runs in fp64
  float f1 = 0.f, f2 = 0.f;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < total; idx += 256) {
    const int slab = idx / cpg, ch = g * cpg + idx % cpg;
    const long long srow = (long long)n * slabs_per_img + slab;
    const float2 q = ch < c1 ? __ldcg(part1 + srow * c1 + ch) : __ldcg(part2 + srow * c2 + (ch - c1));
    f1 += q.x, f2 += q.y;
  }
  double s1 = (double)f1, s2 = (double)f2;
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o), s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  __shared__ double sh[2][8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[0][w] = s1, sh[1][w] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 8; ++k) a += sh[0][k], b += sh[1][k];
    const double cnt = (double)hw * (double)cpg;
    const double m = a / cnt;
    double var = b / cnt - m * m;
    if (var < 0.0) var = 0.0;
    stats[(long long)n * groups + g] = make_float2((float)m, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

__global__ void __launch_bounds__(320)
gn_apply_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2, int hw, int groups,
                const float2* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                int silu_act, __half* __restrict__ out, __half* __restrict__ raw_out, int pix_per_block, int lanes) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float ssm[];  // scale[C], shift[C]
  const int n = blockIdx.y;
  const int C = c1 + c2;
  const int cpg = C / groups;
  float* scale = ssm;
  float* shift = ssm + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float2 mr = stats[(long long)n * groups + c / cpg];
    const float sc = mr.y * (gamma ? gamma[c] : 1.f);
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - mr.x * sc;
  }
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  const int vpp = C >> 3;
  auto norm_store = [&](const uint4& q, const float* sc, const float* sh, long long ooff) {
    if (raw_out) *(uint4*)(raw_out + ooff) = q;
    const __half2* h = (const __half2*)&q;
    uint4 o;
    __half2* oh = (__half2*)&o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h[j]);
      float a = fmaf(t.x, sc[2 * j], sh[2 * j]);
      float b = fmaf(t.y, sc[2 * j + 1], sh[2 * j + 1]);
      if (silu_act) a = silu(a), b = silu(b);
      oh[j] = __floats2half2_rn(a, b);
    }
    *(uint4*)(out + ooff) = o;
  };
  if (lanes > 0) {
    // blockDim = lanes * vpp: a thread owns ONE channel vector (scale/shift in registers) and strides over pixels, four independent
    // 16-byte loads in flight before the first is consumed
    const int cv = ((int)threadIdx.x % vpp) * 8, pl = (int)threadIdx.x / vpp;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] = scale[cv + j], sh[j] = shift[cv + j];
    const bool first = cv < c1;
    const __half* base = first ? x1 + (long long)n * hw * c1 + cv : x2 + (long long)n * hw * c2 + (cv - c1);
    const int cstride = first ? c1 : c2;
    constexpr int U = 4;
    for (int pix0 = p0 + pl; pix0 < p1; pix0 += lanes * U) {
      uint4 q[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pix0 + u * lanes < p1) q[u] = *(const uint4*)(base + (long long)(pix0 + u * lanes) * cstride);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pix0 + u * lanes < p1) norm_store(q[u], sc, sh, ((long long)n * hw + pix0 + u * lanes) * C + cv);
    }
    return;
  }
  const int total = (p1 - p0) * vpp;   // generic mapping (more channel vectors than threads)
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int pix = p0 + idx / vpp;
    const int cv = (idx % vpp) * 8;
    const __half* src = (cv < c1) ? x1 + ((long long)n * hw + pix) * c1 + cv : x2 + ((long long)n * hw + pix) * c2 + (cv - c1);
    const uint4 q = *(const uint4*)src;
    norm_store(q, scale + cv, shift + cv, ((long long)n * hw + pix) * C + cv);
  }
}

}  // namespace seedx

using namespace seedx;

extern "C" int seedx_layernorm(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, void* out,
                               int out_dtype, int64_t ldo, void* out2, const float* add, int64_t add_rows, int64_t rows,
                               int64_t cols, float eps, int rms, void* stream) {
  SEEDX_REQUIRE(x && out, "seedx_layernorm: null pointer");
  SEEDX_REQUIRE(rows > 0 && cols > 0 && cols <= LN_THREADS * LN_MAXPT, "seedx_layernorm: cols=%lld out of range (1..%d)", (long long)cols,
                LN_THREADS * LN_MAXPT);
  SEEDX_REQUIRE((out2 == nullptr) == (add == nullptr), "seedx_layernorm: out2 and add go together");
  if (add) SEEDX_REQUIRE(add_rows > 0, "seedx_layernorm: add_rows must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t in_sz = x_dtype == SEEDX_F32 ? 4 : 2, out_sz = out_dtype == SEEDX_F32 ? 4 : 2;
  const bool vec_ok = cols % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((ldx * in_sz) % (4 * in_sz) == 0) && ((uintptr_t)out % 16 == 0) &&
                      ((ldo * out_sz) % (4 * out_sz) == 0) && (!out2 || (uintptr_t)out2 % 16 == 0) && (!gamma || (uintptr_t)gamma % 16 == 0) &&
                      (!beta || (uintptr_t)beta % 16 == 0) && (!add || (uintptr_t)add % 16 == 0) && cols <= 8192;
#define LN_VEC(TI, TO, TPR, NV)                                                                                              \
  launch_k(layernorm_vec_kernel<TI, TO, TPR, NV>, (unsigned)((rows + (LN_THREADS / TPR) - 1) / (LN_THREADS / TPR)), LN_THREADS, 0, st,  \
      (const TI*)x, ldx, gamma, beta, (TO*)out, ldo, (TO*)out2, add, (int)(add ? add_rows : 1), rows, (int)cols, eps, rms)
#define LN_ROWS(TI, TO, NV)                                                                                                          \
  launch_k(layernorm_rows_kernel<TI, TO, NV>, rows_grid, LN_THREADS, (size_t)cols * 8, st, (const TI*)x, ldx, gamma, beta, (TO*)out, ldo, (TO*)out2, \
                                                                                   add, (int)(add ? add_rows : 1), rows, (int)cols, eps, rms)
#define LN_VEC_DISPATCH(TI, TO)                  \
  do {                                           \
    if (cols <= 2048 && rows >= 64) {            \
      if (cols <= 512) LN_ROWS(TI, TO, 4);       \
      else if (cols <= 1024) LN_ROWS(TI, TO, 8); \
      else if (cols <= 1536) LN_ROWS(TI, TO, 12);\
      else LN_ROWS(TI, TO, 16);                  \
    }                                            \
    else if (cols <= 256) LN_VEC(TI, TO, 32, 2); \
    else if (cols <= 512) LN_VEC(TI, TO, 32, 4); \
    else if (cols <= 1024) LN_VEC(TI, TO, 32, 8);\
    else if (cols <= 2048) LN_VEC(TI, TO, 32, 16);\
    else if (cols <= 4096) LN_VEC(TI, TO, 256, 4);\
    else LN_VEC(TI, TO, 256, 8);                 \
  } while (0)
  long long rg = (rows + (LN_THREADS / 32) - 1) / (LN_THREADS / 32);
  if (rg > (long long)num_sms() * 4) rg = (long long)num_sms() * 4;
  const unsigned rows_grid = (unsigned)rg;
  dim3 grid((unsigned)rows);
#define LN_LAUNCH(TI, TO)                                                                                                  \
  launch_k(layernorm_kernel<TI, TO>, grid, LN_THREADS, 0, st, (const TI*)x, ldx, gamma, beta, (TO*)out, ldo, (TO*)out2, add, \
                                                        (int)(add ? add_rows : 1), (int)cols, eps, rms)
#define LN_BOTH(TI, TO)            \
  do {                             \
    if (vec_ok) LN_VEC_DISPATCH(TI, TO); \
    else LN_LAUNCH(TI, TO);        \
  } while (0)
  if (x_dtype == SEEDX_F32 && out_dtype == SEEDX_F16) LN_BOTH(float, __half);
  else if (x_dtype == SEEDX_F32 && out_dtype == SEEDX_F32) LN_BOTH(float, float);
  else if (x_dtype == SEEDX_F16 && out_dtype == SEEDX_F16) LN_BOTH(__half, __half);
  else if (x_dtype == SEEDX_F16 && out_dtype == SEEDX_F32) LN_BOTH(__half, float);
  else SEEDX_REQUIRE(false, "seedx_layernorm: bad dtype combination");
#undef LN_LAUNCH
#undef LN_ROWS
#undef LN_VEC
#undef LN_VEC_DISPATCH
#undef LN_BOTH
  count_launch();
  return check_cuda(cudaGetLastError(), "layernorm_kernel launch");
}

extern "C" int seedx_row_stats(const void* x, int x_dtype, int64_t ldx, int64_t rows, int64_t cols, float eps, float* stats, void* stream) {
  SEEDX_REQUIRE(x && stats, "seedx_row_stats: null pointer");
  SEEDX_REQUIRE(x_dtype == SEEDX_F16, "seedx_row_stats: fp16 rows only (the rows are the A operand of the consumer GEMM)");
  SEEDX_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= 2048 && ldx % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)stats % 8 == 0),
                "seedx_row_stats: cols=%lld must be a multiple of 8 and <= 2048, rows 16-byte aligned", (long long)cols);
  cudaStream_t st = (cudaStream_t)stream;
  long long rg = (rows + (LN_THREADS / 32) - 1) / (LN_THREADS / 32);
  if (rg > (long long)num_sms() * 8) rg = (long long)num_sms() * 8;
  const unsigned grid = (unsigned)rg;
  cudaError_t e;
  if (cols <= 512) e = launch_k(row_stats_kernel<2>, grid, LN_THREADS, 0, st, (const __half*)x, (long long)ldx, (long long)rows, (int)cols, eps, (float2*)stats);
  else if (cols <= 1024) e = launch_k(row_stats_kernel<4>, grid, LN_THREADS, 0, st, (const __half*)x, (long long)ldx, (long long)rows, (int)cols, eps, (float2*)stats);
  else if (cols <= 1536) e = launch_k(row_stats_kernel<6>, grid, LN_THREADS, 0, st, (const __half*)x, (long long)ldx, (long long)rows, (int)cols, eps, (float2*)stats);
  else e = launch_k(row_stats_kernel<8>, grid, LN_THREADS, 0, st, (const __half*)x, (long long)ldx, (long long)rows, (int)cols, eps, (float2*)stats);
  count_launch();
  return check_cuda(e != cudaSuccess ? e : cudaGetLastError(), "row_stats_kernel launch");
}

extern "C" int seedx_row_stats_from_partials(const float* parts, int nparts, int64_t rows, int64_t cols, float eps, float* stats, void* stream) {
  SEEDX_REQUIRE(parts && stats && nparts > 0 && rows > 0 && cols == 32LL * nparts, "seedx_row_stats_from_partials: cols must be 32 * nparts");
  launch_k(row_finalize_kernel, (unsigned)((rows + 63) / 64), 256, 0, (cudaStream_t)stream, (const float2*)parts, nparts, (long long)rows,
           1.0f / (float)cols, eps, (float2*)stats);
  count_launch();
  return check_cuda(cudaGetLastError(), "row_finalize_kernel launch");
}

// CTAs per image of the statistics pass: ~4 waves of CTAs over the batch, >= 32 pixels each
static int64_t gn_stat_blocks(int64_t n, int64_t hw, int64_t* pix_per_block) {
  int64_t spb = (hw * n + 148 * 4 - 1) / (148 * 4);
  if (spb < 32) spb = 32;
  if (spb > hw) spb = hw;
  if (pix_per_block) *pix_per_block = spb;
  return (hw + spb - 1) / spb;
}

extern "C" int64_t seedx_groupnorm_ws_bytes(int64_t n, int groups) {
  if (n <= 0 || groups <= 0) return 0;
  const int64_t nblk = (148 * 4) / n + 2;   // >= gn_stat_blocks(n, hw) for every hw
  return n * groups * 2 * (int64_t)sizeof(double) + n * nblk * groups * 2 * (int64_t)sizeof(float) + n * (int64_t)sizeof(unsigned);
}

static int gn_launch_apply(const void* x1, int64_t c1, const void* x2, int64_t c2, int64_t n, int64_t hw, int groups, const float2* stats,
                           const float* gamma, const float* beta, int silu_act, void* out, void* raw_out, cudaStream_t st) {
  const int64_t C = c1 + c2, vpp = C / 8;
  // apply: ~8 CTAs per SM in flight, at least 16 pixels per block; block = whole pixels (lanes * vpp threads) when a pixel fits
  int64_t ppb = (hw * n + 148 * 8 - 1) / (148 * 8);
  if (ppb < 16) ppb = 16;
  if (ppb > hw) ppb = hw;
  dim3 g2((unsigned)((hw + ppb - 1) / ppb), (unsigned)n);
  const size_t smem = (size_t)C * 2 * sizeof(float);
  const int a_lanes = vpp <= 320 ? (int)(320 / vpp) : 0;
  const int a_threads = a_lanes > 0 ? (int)(a_lanes * vpp) : 256;
  launch_k(gn_apply_kernel, g2, a_threads, smem, st, (const __half*)x1, (int)c1, (const __half*)x2, (int)c2, (int)hw, groups, stats,
           gamma, beta, silu_act, (__half*)out, (__half*)raw_out, (int)ppb, a_lanes);
  count_launch();
  return check_cuda(cudaGetLastError(), "groupnorm apply launch");
}

extern "C" int seedx_groupnorm_nhwc_from_partials(const void* x1, int64_t c1, const float* part1, const void* x2, int64_t c2, const float* part2,
                                                  int64_t n, int64_t hw, int groups, const float* gamma, const float* beta, float eps, int silu_act,
                                                  void* out, void* raw_out, void* stats_ws, void* stream) {
  SEEDX_REQUIRE(x1 && out && stats_ws && part1, "seedx_groupnorm_nhwc_from_partials: null pointer");
  if (!x2) c2 = 0;
  SEEDX_REQUIRE(!x2 || part2, "seedx_groupnorm_nhwc_from_partials: x2 needs its partials");
  const int64_t C = c1 + c2;
  SEEDX_REQUIRE(c1 % 8 == 0 && c2 % 8 == 0 && groups > 0 && groups <= 128 && C % groups == 0, "seedx_groupnorm_nhwc_from_partials: bad channel counts");
  SEEDX_REQUIRE(n > 0 && n <= 65535 && hw > 0 && hw % 32 == 0, "seedx_groupnorm_nhwc_from_partials: hw=%lld must be a multiple of 32 (slab = 32 rows)",
                (long long)hw);
  cudaStream_t st = (cudaStream_t)stream;
  float2* stats = (float2*)stats_ws;
  dim3 grid((unsigned)groups, (unsigned)n);
  launch_k(gn_finalize_kernel, grid, 256, 0, st, (const float2*)part1, (int)c1, (const float2*)part2, (int)c2, (int)(hw / 32), groups, (int)hw, eps,
           stats);
  count_launch();
  SEEDX_CUDA(cudaGetLastError());
  return gn_launch_apply(x1, c1, x2, c2, n, hw, groups, stats, gamma, beta, silu_act, out, raw_out, st);
}

extern "C" int seedx_groupnorm_nhwc(const void* x1, int64_t c1, const void* x2, int64_t c2, int64_t n, int64_t hw, int groups,
                                    const float* gamma, const float* beta, float eps, int silu_act, void* out, void* raw_out,
                                    void* stats_ws, void* stream) {
  SEEDX_REQUIRE(x1 && out && stats_ws, "seedx_groupnorm_nhwc: null pointer");
  const int64_t C = c1 + (x2 ? c2 : 0);
  if (!x2) c2 = 0;
  SEEDX_REQUIRE(c1 % 8 == 0 && c2 % 8 == 0 && groups > 0 && C % groups == 0, "seedx_groupnorm_nhwc: bad channel counts c1=%lld c2=%lld",
                (long long)c1, (long long)c2);
  SEEDX_REQUIRE(n > 0 && n <= 65535 && hw > 0, "seedx_groupnorm_nhwc: bad n/hw");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t spb_out = 0;
  SEEDX_REQUIRE(groups <= 128, "seedx_groupnorm_nhwc: at most 128 groups");
  const int64_t nblk = gn_stat_blocks(n, hw, &spb_out);
  const int64_t spb = spb_out;
  float2* stats = (float2*)stats_ws;                                   // [n][groups] (mean, rstd), read by the apply pass
  float* partials = (float*)((double*)stats_ws + n * groups * 2);      // [n][nblk][groups][2]
  unsigned* tickets = (unsigned*)(partials + n * nblk * groups * 2);   // [n]
  SEEDX_CUDA(cudaMemsetAsync(tickets, 0, sizeof(unsigned) * n, st));
  int tpp = 1;
  while (tpp < 32 && tpp * 2 * groups * 2 <= 256) tpp *= 2;
  const int64_t vpp = C / 8;
  const int64_t lanes = vpp >= 256 ? 1 : 256 / vpp;
  dim3 g1((unsigned)nblk, (unsigned)n);
  launch_k(gn_stats_kernel, g1, 256, (size_t)(lanes * C * 2) * sizeof(float), st, (const __half*)x1, (int)c1, (const __half*)x2, (int)c2, (int)hw,
           groups, (int)spb, stats, partials, tickets, tpp, eps);
  count_launch();
  SEEDX_CUDA(cudaGetLastError());
  if (int e = gn_launch_apply(x1, c1, x2, c2, n, hw, groups, (const float2*)stats, gamma, beta, silu_act, out, raw_out, st)) return e;
  return check_cuda(cudaGetLastError(), "groupnorm kernels launch");
}
