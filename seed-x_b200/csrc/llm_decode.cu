// seedx-b200: HBM-bound kernels of the LLaMA token loop (batch-1 greedy decode) and the prefill glue.
//
// Decode reads every weight once per token (26 GB at 13 B parameters) -> the roofline is HBM bandwidth, not the tensor
// cores: the projections are GEMVs streaming fp16 rows with 128-bit loads (RMSNorm fused in the prologue, SwiGLU /
// residual in the epilogue), attention walks the fp16 KV cache with 16-byte loads, and the sampler state (sequence,
// length, done flag) lives on the device so a whole step is one CUDA-graph replay with no host synchronisation.
//
// Replaces: src/models/mllm/modeling_llama_xformer.py:141-149 (RoPE), 204-239 (q/k/v/o at M=1, KV torch.cat append,
// xformers attention over the cache), 166-167 (SwiGLU MLP), 707 (lm_head); transformers LlamaRMSNorm; the python
// logits processor src/models/mllm/generation.py:19-31 and the argmax of HF greedy_search.
#include "common.cuh"
#include "../../include/seedx.h"

namespace seedx {
void count_launch();

// ------------------------------------------------------------------------------------------------
// GEMV: out = epi( W[N,K] . norm(x) ),  W fp16 row-major, x fp32.  One warp owns a pair of rows at a time.
// ------------------------------------------------------------------------------------------------
constexpr int GEMV_THREADS = 256;

__global__ void __launch_bounds__(GEMV_THREADS)
gemv_kernel(const __half* __restrict__ W, const float* __restrict__ x, const float* __restrict__ rms_w, float eps,
            const float* __restrict__ residual, float* __restrict__ out, int N, int K, int gated) {
  extern __shared__ float xs[];  // K floats
  __shared__ float red[GEMV_THREADS / 32];
  float ss = 0.f;
  for (int k = threadIdx.x; k < K; k += GEMV_THREADS) {
    const float v = x[k];
    xs[k] = v;
    ss += v * v;
  }
  if (rms_w != nullptr) {
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < GEMV_THREADS / 32; ++i) tot += red[i];
    const float r = rsqrtf(tot / (float)K + eps);
    for (int k = threadIdx.x; k < K; k += GEMV_THREADS) xs[k] = xs[k] * r * rms_w[k];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (GEMV_THREADS / 32);
  const int chunks = K >> 3;  // 16-byte chunks per row
  for (int pair = blockIdx.x * (GEMV_THREADS / 32) + warp; pair * 2 < N; pair += warps_total) {
    const int r0 = pair * 2;
    const bool has1 = (r0 + 1) < N;
    const uint4* w0 = (const uint4*)(W + (long long)r0 * K);
    const uint4* w1 = (const uint4*)(W + (long long)(has1 ? r0 + 1 : r0) * K);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int c = lane; c < chunks; c += 32) {
      const uint4 q0 = __ldcs(w0 + c);  // streaming: every weight is read exactly once per token
      const uint4 q1 = __ldcs(w1 + c);
      const float4 xa = *(const float4*)(xs + c * 8);
      const float4 xb = *(const float4*)(xs + c * 8 + 4);
      const __half2* h0 = (const __half2*)&q0;
      const __half2* h1 = (const __half2*)&q1;
      float2 f;
      f = __half22float2(h0[0]); a0 += f.x * xa.x + f.y * xa.y;
      f = __half22float2(h0[1]); a0 += f.x * xa.z + f.y * xa.w;
      f = __half22float2(h0[2]); a0 += f.x * xb.x + f.y * xb.y;
      f = __half22float2(h0[3]); a0 += f.x * xb.z + f.y * xb.w;
      f = __half22float2(h1[0]); a1 += f.x * xa.x + f.y * xa.y;
      f = __half22float2(h1[1]); a1 += f.x * xa.z + f.y * xa.w;
      f = __half22float2(h1[2]); a1 += f.x * xb.x + f.y * xb.y;
      f = __half22float2(h1[3]); a1 += f.x * xb.z + f.y * xb.w;
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      if (gated) {
        out[pair] = a0 * silu(a1);  // rows interleaved [up_j, gate_j]: down(silu(gate(x)) * up(x)), :166-167
      } else {
        out[r0] = a0 + (residual ? residual[r0] : 0.f);
        if (has1) out[r0 + 1] = a1 + (residual ? residual[r0 + 1] : 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// decode attention for one new token: RoPE(q,k) -> append k,v to the cache -> softmax(q K^T) V over positions 0..pos
// one CTA per head, head_dim 128, 8 warps, each half-warp walks keys with 16-byte loads.
// state[0] = current sequence length (the token being consumed sits at position state[0] - 1)
// ------------------------------------------------------------------------------------------------
constexpr int DA_THREADS = 256;
constexpr int DA_GROUPS = DA_THREADS / 16;  // half-warps

__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const float* __restrict__ qkv, const int* __restrict__ state, const float* __restrict__ inv_freq,
                   __half* __restrict__ kcache, __half* __restrict__ vcache, float* __restrict__ out, int H, float scale) {
  constexpr int HD = 128;
  __shared__ float qs[HD];
  __shared__ float gm[DA_GROUPS], gl[DA_GROUPS];
  __shared__ float gacc[DA_GROUPS][HD];
  const int h = blockIdx.x;
  const int D = H * HD;
  const int pos = state[0] - 1;
  const int tid = threadIdx.x;
  if (tid < HD / 2) {
    const float ang = (float)pos * inv_freq[tid];
    const float c = cosf(ang), s = sinf(ang);
    const float q0 = qkv[h * HD + tid], q1 = qkv[h * HD + tid + HD / 2];
    const float k0 = qkv[D + h * HD + tid], k1 = qkv[D + h * HD + tid + HD / 2];
    qs[tid] = (q0 * c - q1 * s) * scale;
    qs[tid + HD / 2] = (q1 * c + q0 * s) * scale;
    kcache[(long long)pos * D + h * HD + tid] = __float2half_rn(k0 * c - k1 * s);
    kcache[(long long)pos * D + h * HD + tid + HD / 2] = __float2half_rn(k1 * c + k0 * s);
  } else if (tid < HD / 2 + HD) {
    const int i = tid - HD / 2;
    vcache[(long long)pos * D + h * HD + i] = __float2half_rn(qkv[2 * D + h * HD + i]);
  }
  __syncthreads();

  const int grp = tid >> 4, gl_lane = tid & 15;
  float q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = qs[gl_lane * 8 + i];
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int t0 = 0; t0 <= pos; t0 += DA_GROUPS) {  // warp-uniform trip count: the half-warp shuffles need all 32 lanes
    const int t = t0 + grp;
    const bool valid = t <= pos;
    uint4 kq = make_uint4(0, 0, 0, 0), vq = make_uint4(0, 0, 0, 0);
    if (valid) {
      kq = *(const uint4*)(kcache + (long long)t * D + h * HD + gl_lane * 8);
      vq = *(const uint4*)(vcache + (long long)t * D + h * HD + gl_lane * 8);
    }
    const __half2* kh = (const __half2*)&kq;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(kh[j]);
      s += f.x * q[2 * j] + f.y * q[2 * j + 1];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // within the 16-lane group
    if (valid) {
      const float nm = fmaxf(m, s);
      const float corr = __expf(m - nm);
      const float p = __expf(s - nm);
      l = l * corr + p;
      const __half2* vh = (const __half2*)&vq;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(vh[j]);
        acc[2 * j] = acc[2 * j] * corr + p * f.x;
        acc[2 * j + 1] = acc[2 * j + 1] * corr + p * f.y;
      }
      m = nm;
    }
  }
  if (gl_lane == 0) gm[grp] = m, gl[grp] = l;
#pragma unroll
  for (int i = 0; i < 8; ++i) gacc[grp][gl_lane * 8 + i] = acc[i];
  __syncthreads();
  if (tid < HD) {
    float M = -INFINITY;
#pragma unroll
    for (int g = 0; g < DA_GROUPS; ++g) M = fmaxf(M, gm[g]);
    float L = 0.f, o = 0.f;
#pragma unroll
    for (int g = 0; g < DA_GROUPS; ++g) {
      const float w = (gm[g] == -INFINITY) ? 0.f : __expf(gm[g] - M);
      L += w * gl[g];
      o += w * gacc[g][tid];
    }
    out[h * HD + tid] = o / L;
  }
}

// RoPE on the prefill q/k (in place, fp16 [T, 3D] = [q | k | v]) + copy of k, v into the cache rows pos0..pos0+T-1
__global__ void rope_kv_prefill_kernel(__half* __restrict__ qkv, int T, int pos0, int H, const float* __restrict__ inv_freq,
                                       __half* __restrict__ kcache, __half* __restrict__ vcache) {
  constexpr int HD = 128;
  const int D = H * HD;
  const long long total = (long long)T * H * (HD / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % (HD / 2));
    const long long r = i / (HD / 2);
    const int h = (int)(r % H);
    const int t = (int)(r / H);
    const float ang = (float)(pos0 + t) * inv_freq[j];
    const float c = cosf(ang), s = sinf(ang);
    __half* row = qkv + (long long)t * 3 * D;
    const int o = h * HD + j;
    const float q0 = __half2float(row[o]), q1 = __half2float(row[o + HD / 2]);
    const float k0 = __half2float(row[D + o]), k1 = __half2float(row[D + o + HD / 2]);
    row[o] = __float2half_rn(q0 * c - q1 * s);
    row[o + HD / 2] = __float2half_rn(q1 * c + q0 * s);
    const __half kr0 = __float2half_rn(k0 * c - k1 * s), kr1 = __float2half_rn(k1 * c + k0 * s);
    row[D + o] = kr0;
    row[D + o + HD / 2] = kr1;
    const long long cr = (long long)(pos0 + t) * D;
    kcache[cr + o] = kr0;
    kcache[cr + o + HD / 2] = kr1;
    vcache[cr + o] = row[2 * D + o];
    vcache[cr + o + HD / 2] = row[2 * D + o + HD / 2];
  }
}

// rows of the embedding table -> fp32.  ids == NULL: single row for the last token of the device-resident sequence
__global__ void embed_rows_kernel(const __half* __restrict__ table, const int* __restrict__ ids, const int* __restrict__ state,
                                  const int* __restrict__ seq, int n, int D, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n * D; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i - (long long)r * D);
    const int id = ids ? ids[r] : seq[state[0] - 1];
    out[i] = __half2float(table[(long long)id * D + c]);
  }
}

template <typename TS>
__global__ void scatter_rows_kernel(const TS* __restrict__ src, const int* __restrict__ src_idx, const int* __restrict__ idx, int n, int D,
                                    float* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n * D; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i - (long long)r * D);
    const long long sr = src_idx ? src_idx[r] : r;
    dst[(long long)idx[r] * D + c] = (float)src[sr * D + c];
  }
}

// hidden[(state[0] - prompt_len - 1) * D + i] = x[i]   (post-norm last hidden state of the position consuming generated token j)
__global__ void store_hidden_kernel(const float* __restrict__ x, const int* __restrict__ state, int prompt_len, int max_rows, int D,
                                    float* __restrict__ hidden) {
  const int row = state[0] - prompt_len - 1;
  if (row < 0 || row >= max_rows) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) hidden[(long long)row * D + i] = x[i];
}

// logits processor (generation.py:19-31) + greedy argmax + sequence append, all on the device.
// state[0] = sequence length, state[1] = done flag (EOS seen), state[2] = number of generated tokens
__global__ void __launch_bounds__(1024)
logits_argmax_kernel(float* __restrict__ logits, int V, const int* __restrict__ img_ids, int n_img_ids, int* __restrict__ seq,
                     int* __restrict__ state, int eos_id, int suppress_eos, int max_len) {
  __shared__ float smax[32];
  __shared__ int sidx[32];
  __shared__ int forced;
  const int len = state[0];
  const int last = seq[len - 1];
  if (threadIdx.x == 0) forced = -1;
  __syncthreads();
  if (threadIdx.x < n_img_ids - 1 && img_ids[threadIdx.x] == last) forced = img_ids[threadIdx.x + 1];
  __syncthreads();
  int next;
  if (forced >= 0) {
    next = forced;  // scores[next] = max + 10 -> argmax is `next`
  } else {
    if (threadIdx.x >= 1 && threadIdx.x < n_img_ids) logits[img_ids[threadIdx.x]] = 0.0f;  // img_ids[1:] <- 0.0 (not -inf!)
    if (suppress_eos && threadIdx.x == 0 && eos_id >= 0 && eos_id < V) logits[eos_id] = -INFINITY;
    __syncthreads();
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float v = logits[i];
      if (v > best || (v == best && i < bi)) best = v, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = best, sidx[threadIdx.x >> 5] = bi;
    __syncthreads();
    if (threadIdx.x < 32) {
      best = threadIdx.x < (blockDim.x >> 5) ? smax[threadIdx.x] : -INFINITY;
      bi = threadIdx.x < (blockDim.x >> 5) ? sidx[threadIdx.x] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
      }
      if (threadIdx.x == 0) sidx[0] = bi;
    }
    __syncthreads();
    next = sidx[0];
  }
  if (threadIdx.x == 0 && len < max_len) {
    seq[len] = next;
    state[0] = len + 1;
    state[2] += 1;
    if (next == eos_id && !suppress_eos && state[1] == 0) state[1] = state[2];  // index (1-based) of the EOS token
  }
}

}  // namespace seedx
using namespace seedx;

static inline int ew_grid(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

extern "C" int seedx_gemv_f16(const void* W, const float* x, const float* rms_w, float eps, const float* residual, float* out, int64_t N,
                              int64_t K, int gated, void* stream) {
  SEEDX_REQUIRE(W && x && out, "seedx_gemv_f16: null pointer");
  SEEDX_REQUIRE(K % 8 == 0 && K > 0 && N > 0 && K * 4 <= 200 * 1024, "seedx_gemv_f16: K=%lld unsupported (multiple of 8, <= 51200)", (long long)K);
  SEEDX_REQUIRE((uintptr_t)W % 16 == 0, "seedx_gemv_f16: W must be 16B aligned");
  if (gated) SEEDX_REQUIRE(N % 2 == 0 && residual == nullptr, "seedx_gemv_f16: gated needs even N and no residual");
  static bool attr = false;
  if (!attr) {
    SEEDX_CUDA(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const long long pairs = (N + 1) / 2;
  long long blocks = (pairs + (GEMV_THREADS / 32) - 1) / (GEMV_THREADS / 32);
  const long long cap = (long long)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  gemv_kernel<<<(unsigned)blocks, GEMV_THREADS, (size_t)K * 4, (cudaStream_t)stream>>>((const __half*)W, x, rms_w, eps, residual, out, (int)N,
                                                                                       (int)K, gated);
  count_launch();
  return check_cuda(cudaGetLastError(), "gemv launch");
}

extern "C" int seedx_decode_attention(const float* qkv, const int32_t* state, const float* inv_freq, void* kcache, void* vcache, float* out,
                                      int heads, int head_dim, float scale, void* stream) {
  SEEDX_REQUIRE(qkv && state && inv_freq && kcache && vcache && out, "seedx_decode_attention: null pointer");
  SEEDX_REQUIRE(head_dim == 128, "seedx_decode_attention: head_dim must be 128 (LLaMA)");
  decode_attn_kernel<<<heads, DA_THREADS, 0, (cudaStream_t)stream>>>(qkv, state, inv_freq, (__half*)kcache, (__half*)vcache, out, heads, scale);
  count_launch();
  return check_cuda(cudaGetLastError(), "decode_attention launch");
}

extern "C" int seedx_rope_kv_prefill(void* qkv, int64_t tokens, int64_t pos0, int heads, int head_dim, const float* inv_freq, void* kcache,
                                     void* vcache, void* stream) {
  SEEDX_REQUIRE(qkv && inv_freq && kcache && vcache && tokens > 0, "seedx_rope_kv_prefill: bad arguments");
  SEEDX_REQUIRE(head_dim == 128, "seedx_rope_kv_prefill: head_dim must be 128 (LLaMA)");
  rope_kv_prefill_kernel<<<ew_grid(tokens * heads * 64, 256), 256, 0, (cudaStream_t)stream>>>((__half*)qkv, (int)tokens, (int)pos0, heads, inv_freq,
                                                                                             (__half*)kcache, (__half*)vcache);
  count_launch();
  return check_cuda(cudaGetLastError(), "rope_kv_prefill launch");
}

extern "C" int seedx_embed_rows(const void* table, const int32_t* ids, const int32_t* state, const int32_t* seq, int64_t n, int64_t dim,
                                float* out, void* stream) {
  SEEDX_REQUIRE(table && out && n > 0 && (ids || (state && seq)), "seedx_embed_rows: bad arguments");
  embed_rows_kernel<<<ew_grid(n * dim, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)table, ids, state, seq, (int)n, (int)dim, out);
  count_launch();
  return check_cuda(cudaGetLastError(), "embed_rows launch");
}

extern "C" int seedx_scatter_rows(const void* src, int src_dtype, const int32_t* src_idx, const int32_t* idx, int64_t n, int64_t dim, float* dst,
                                  void* stream) {
  SEEDX_REQUIRE(src && idx && dst && n > 0, "seedx_scatter_rows: bad arguments");
  const int g = ew_grid(n * dim, 256);
  if (src_dtype == SEEDX_F32) scatter_rows_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)src, src_idx, idx, (int)n, (int)dim, dst);
  else if (src_dtype == SEEDX_F16) scatter_rows_kernel<__half><<<g, 256, 0, (cudaStream_t)stream>>>((const __half*)src, src_idx, idx, (int)n, (int)dim, dst);
  else SEEDX_REQUIRE(false, "seedx_scatter_rows: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "scatter_rows launch");
}

extern "C" int seedx_store_hidden(const float* x, const int32_t* state, int64_t prompt_len, int64_t max_rows, int64_t dim, float* hidden,
                                  void* stream) {
  SEEDX_REQUIRE(x && state && hidden, "seedx_store_hidden: null pointer");
  store_hidden_kernel<<<ew_grid(dim, 256), 256, 0, (cudaStream_t)stream>>>(x, state, (int)prompt_len, (int)max_rows, (int)dim, hidden);
  count_launch();
  return check_cuda(cudaGetLastError(), "store_hidden launch");
}

extern "C" int seedx_logits_argmax(float* logits, int64_t vocab, const int32_t* img_ids, int n_img_ids, int32_t* seq, int32_t* state, int eos_id,
                                   int suppress_eos, int64_t max_len, void* stream) {
  SEEDX_REQUIRE(logits && seq && state && vocab > 0, "seedx_logits_argmax: bad arguments");
  SEEDX_REQUIRE(n_img_ids >= 0 && n_img_ids <= 1024, "seedx_logits_argmax: too many image token ids");
  logits_argmax_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, (int)vocab, img_ids, n_img_ids, seq, state, eos_id, suppress_eos, (int)max_len);
  count_launch();
  return check_cuda(cudaGetLastError(), "logits_argmax launch");
}
