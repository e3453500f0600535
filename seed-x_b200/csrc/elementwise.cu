// seedx-b200: small HBM-bound data-movement kernels (patchify / cast / token pooling ...).
// Each is a grid-stride loop with 16-byte accesses where the layout allows; reference call sites in include/seedx.h.
#include "common.cuh"
#include "../../include/seedx.h"

namespace seedx {
void count_launch();

static inline int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename TI>
__global__ void patchify_kernel(const TI* __restrict__ x, int n, int c, int h, int w, int patch, __half* __restrict__ out, int kpad) {
  pdl_wait();
  pdl_trigger();
  const int gh = h / patch, gw = w / patch;
  const long long total = (long long)n * gh * gw * kpad;
  const int kreal = c * patch * patch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % kpad);
    const long long row = i / kpad;
    float v = 0.f;
    if (k < kreal) {
      const int ci = k / (patch * patch);
      const int rem = k - ci * patch * patch;
      const int pi = rem / patch, pj = rem - pi * patch;
      const int gx = (int)(row % gw);
      const long long t = row / gw;
      const int gy = (int)(t % gh);
      const int img = (int)(t / gh);
      v = (float)x[(((long long)img * c + ci) * h + gy * patch + pi) * w + gx * patch + pj];
    }
    out[i] = __float2half_rn(v);
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ s, TO* __restrict__ d, long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = (TO)(float)s[i];
}

template <typename T>
__global__ void avgpool_tokens_kernel(const T* __restrict__ x, long long n_out_rows, int c, int k, T* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long total = n_out_rows * c;
  const float inv = 1.f / (float)k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c;
    const int ch = (int)(i - r * c);
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += (float)x[(r * k + j) * c + ch];
    out[i] = (T)(s * inv);
  }
}


// im2col for strided convs on NHWC fp16 (k x k taps, stride s, pad_before on top/left, zero fill), 8 channels / thread
__global__ void im2col_nhwc_kernel(const __half* __restrict__ x, int n, int h, int w, int c, int k, int stride, int pad, int ho, int wo,
                                   __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int cv = c >> 3;
  const long long total = (long long)n * ho * wo * k * k * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long r = i / cv;
    const int tap = (int)(r % (k * k));
    r /= (k * k);
    const int ox = (int)(r % wo);
    r /= wo;
    const int oy = (int)(r % ho);
    const int img = (int)(r / ho);
    const int iy = oy * stride + tap / k - pad, ix = ox * stride + tap % k - pad;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) q = *(const uint4*)(x + (((long long)img * h + iy) * w + ix) * c + v * 8);
    *(uint4*)(out + i * 8) = q;
  }
}

__global__ void upsample2x_nhwc_kernel(const __half* __restrict__ x, int n, int h, int w, int c, __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int cv = c >> 3;
  const long long total = (long long)n * (2 * h) * (2 * w) * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long r = i / cv;
    const int ox = (int)(r % (2 * w));
    r /= (2 * w);
    const int oy = (int)(r % (2 * h));
    const int img = (int)(r / (2 * h));
    *(uint4*)(out + i * 8) = *(const uint4*)(x + (((long long)img * h + (oy >> 1)) * w + (ox >> 1)) * c + v * 8);
  }
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): out[i, :] = [cos(t_i f_j) | sin(t_i f_j)]
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int count, int dim, __half* __restrict__ out, long long ldo) {
  pdl_wait();
  pdl_trigger();
  const int half_dim = dim >> 1;
  const int total = count * half_dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / half_dim, j = i - r * half_dim;
    const float f = expf(-9.210340371976184f * (float)j / (float)half_dim);  // ln(10000)
    const float a = t[r] * f;
    out[(long long)r * ldo + j] = __float2half_rn(cosf(a));
    out[(long long)r * ldo + half_dim + j] = __float2half_rn(sinf(a));
  }
}

template <typename TI>
__global__ void unary_kernel(const TI* __restrict__ x, long long rows, int cols, long long ldx, __half* __restrict__ out, long long ldo, int act) {
  pdl_wait();
  pdl_trigger();
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    float v = (float)x[r * ldx + c];
    if (act == SEEDX_ACT_SILU) v = silu(v);
    else if (act == SEEDX_ACT_GELU_ERF) v = gelu_erf(v);
    out[r * ldo + c] = __float2half_rn(v);
  }
}

// row softmax: out = softmax(scale * x) ; one CTA per row, fp32 math
template <typename TI>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const TI* __restrict__ x, long long ldx, int cols, float scale, __half* __restrict__ out,
                                                           long long ldo) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[8];
  const TI* xr = x + (long long)blockIdx.x * ldx;
  __half* orow = out + (long long)blockIdx.x * ldo;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, (float)xr[c] * scale);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  m = sh[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, sh[i]);
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) s += __expf((float)xr[c] * scale - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += sh[i];
  const float inv = 1.f / s;
  for (int c = threadIdx.x; c < cols; c += 256) orow[c] = __float2half_rn(__expf((float)xr[c] * scale - m) * inv);
}

// NCHW fp32 -> NHWC fp16 with channel padding and scale
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int n, int c, int hw, int cpad, float scale, __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * hw * cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cpad);
    const long long r = i / cpad;
    const int p = (int)(r % hw);
    const int img = (int)(r / hw);
    out[i] = __float2half_rn(ch < c ? x[((long long)img * c + ch) * hw + p] * scale : 0.f);
  }
}

template <typename TI>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ x, int n, int c, int hw, int ldc, float scale, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * c * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw);
    const long long r = i / hw;
    const int ch = (int)(r % c);
    const int img = (int)(r / c);
    out[i] = (float)x[((long long)img * hw + p) * ldc + ch] * scale;
  }
}

// VaeImageProcessor.postprocess: (x/2 + 0.5).clamp(0,1) -> uint8 (round half to even like torch .round())
template <typename TI>
__global__ void image_to_u8_kernel(const TI* __restrict__ x, long long pixels, int ldc, uint8_t* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / 3;
    const int ch = (int)(i - p * 3);
    float v = (float)x[p * ldc + ch] * 0.5f + 0.5f;
    v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = (uint8_t)__float2int_rn(v * 255.f);
  }
}

// One sampler step (fp32 state): classifier-free guidance combine + Euler update + next-step UNet input.
//   eps   : fp32 NHWC [nb*B, HW, 4] UNet output (NULL -> initialisation: x = noise * init_sigma)
//   x     : fp32 NCHW [B, 4, HW] sampler state (in/out)
//   unet_in: fp16 NHWC [nb*B, HW, 8]; channels 0..3 <- x_new / sqrt(sigma_next^2 + 1) for every branch
//   mode 2 (t2i, branches [uncond, text])   : e = e_u + g (e_t - e_u)                       (adapter_modules.py:156-167 -> diffusers SDXL pipeline)
//   mode 3 (edit, branches [text, image, uncond]): sigma-space combine, pipeline_stable_diffusion_xl_t2i_edit.py:928-950
__global__ void cfg_euler_kernel(const float* __restrict__ eps, float* __restrict__ x, __half* __restrict__ unet_in, int B, int hw, int mode,
                                 float g, float ig, float sigma, float sigma_next, float init_sigma) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)B * hw * 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i & 3);
    const long long r = i >> 2;
    const int p = (int)(r % hw);
    const int b = (int)(r / hw);
    const long long xi = ((long long)b * 4 + ch) * hw + p;
    float xv = x[xi];
    float xn;
    if (eps == nullptr) {
      xn = xv * init_sigma;
    } else {
      float e;
      if (mode == 2) {
        const float eu = eps[(((long long)(0 * B + b)) * hw + p) * 4 + ch];
        const float et = eps[(((long long)(1 * B + b)) * hw + p) * 4 + ch];
        e = eu + g * (et - eu);
      } else {
        const float et = xv - sigma * eps[(((long long)(0 * B + b)) * hw + p) * 4 + ch];
        const float ei = xv - sigma * eps[(((long long)(1 * B + b)) * hw + p) * 4 + ch];
        const float eu = xv - sigma * eps[(((long long)(2 * B + b)) * hw + p) * 4 + ch];
        const float c = eu + g * (et - ei) + ig * (ei - eu);
        e = (c - xv) / (-sigma);
      }
      const float x0 = xv - sigma * e;
      const float d = (xv - x0) / sigma;
      xn = xv + d * (sigma_next - sigma);
    }
    x[xi] = xn;
    const __half s = __float2half_rn(xn * rsqrtf(sigma_next * sigma_next + 1.f));
    for (int br = 0; br < mode; ++br) unet_in[(((long long)(br * B + b)) * hw + p) * 8 + ch] = s;
  }
}


// out[r, :] = fp16( a[r, :] + b[r % b_rows, :] )   (positional-embedding add of AttentionPool2d, resampler.py:93)
template <typename TA>
__global__ void add_bcast_kernel(const TA* __restrict__ a, const float* __restrict__ b, long long rows, int cols, int b_rows,
                                 __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    out[i] = __float2half_rn((float)a[i] + b[(r % b_rows) * cols + c]);
  }
}

}  // namespace seedx
using namespace seedx;

extern "C" int seedx_patchify(const void* x, int x_dtype, int64_t n, int64_t c, int64_t h, int64_t w, int64_t patch, void* out,
                              int64_t kpad, void* stream) {
  SEEDX_REQUIRE(x && out, "seedx_patchify: null pointer");
  SEEDX_REQUIRE(patch > 0 && h % patch == 0 && w % patch == 0, "seedx_patchify: image %lldx%lld not a multiple of patch %lld", (long long)h,
                (long long)w, (long long)patch);
  SEEDX_REQUIRE(kpad >= c * patch * patch, "seedx_patchify: kpad too small");
  const long long total = n * (h / patch) * (w / patch) * kpad;
  const int g = grid_for(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == SEEDX_F32)
    launch_k(patchify_kernel<float>, g, 256, 0, st, (const float*)x, (int)n, (int)c, (int)h, (int)w, (int)patch, (__half*)out, (int)kpad);
  else if (x_dtype == SEEDX_F16)
    launch_k(patchify_kernel<__half>, g, 256, 0, st, (const __half*)x, (int)n, (int)c, (int)h, (int)w, (int)patch, (__half*)out, (int)kpad);
  else
    SEEDX_REQUIRE(false, "seedx_patchify: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "patchify launch");
}

extern "C" int seedx_cast(const void* src, int sd, void* dst, int dd, int64_t count, void* stream) {
  SEEDX_REQUIRE(src && dst && count > 0, "seedx_cast: bad arguments");
  const int g = grid_for(count, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (sd == SEEDX_F32 && dd == SEEDX_F16) launch_k(cast_kernel<float, __half>, g, 256, 0, st, (const float*)src, (__half*)dst, count);
  else if (sd == SEEDX_F16 && dd == SEEDX_F32) launch_k(cast_kernel<__half, float>, g, 256, 0, st, (const __half*)src, (float*)dst, count);
  else SEEDX_REQUIRE(false, "seedx_cast: unsupported conversion %d -> %d", sd, dd);
  count_launch();
  return check_cuda(cudaGetLastError(), "cast launch");
}

extern "C" int seedx_avgpool_tokens(const void* x, int dtype, int64_t n, int64_t t, int64_t c, int64_t k, void* out, void* stream) {
  SEEDX_REQUIRE(x && out && k > 0 && t % k == 0, "seedx_avgpool_tokens: bad arguments");
  const long long rows = n * (t / k);
  const int g = grid_for(rows * c, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SEEDX_F16) launch_k(avgpool_tokens_kernel<__half>, g, 256, 0, st, (const __half*)x, rows, (int)c, (int)k, (__half*)out);
  else if (dtype == SEEDX_F32) launch_k(avgpool_tokens_kernel<float>, g, 256, 0, st, (const float*)x, rows, (int)c, (int)k, (float*)out);
  else SEEDX_REQUIRE(false, "seedx_avgpool_tokens: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "avgpool launch");
}

extern "C" int seedx_im2col_nhwc(const void* x, int64_t n, int64_t h, int64_t w, int64_t c, int k, int stride, int pad_before, int64_t ho,
                                 int64_t wo, void* out, void* stream) {
  SEEDX_REQUIRE(x && out && c % 8 == 0 && k > 0 && stride > 0, "seedx_im2col_nhwc: bad arguments");
  const long long total = n * ho * wo * k * k * (c / 8);
  launch_k(im2col_nhwc_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const __half*)x, (int)n, (int)h, (int)w, (int)c, k, stride,
                                                                            pad_before, (int)ho, (int)wo, (__half*)out);
  count_launch();
  return check_cuda(cudaGetLastError(), "im2col launch");
}

extern "C" int seedx_upsample2x_nhwc(const void* x, int64_t n, int64_t h, int64_t w, int64_t c, void* out, void* stream) {
  SEEDX_REQUIRE(x && out && c % 8 == 0, "seedx_upsample2x_nhwc: bad arguments");
  const long long total = n * 4 * h * w * (c / 8);
  launch_k(upsample2x_nhwc_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const __half*)x, (int)n, (int)h, (int)w, (int)c, (__half*)out);
  count_launch();
  return check_cuda(cudaGetLastError(), "upsample launch");
}

extern "C" int seedx_timestep_embedding(const float* t, int64_t count, int dim, void* out, int64_t ldo, void* stream) {
  SEEDX_REQUIRE(t && out && count > 0 && dim > 0 && dim % 2 == 0, "seedx_timestep_embedding: bad arguments");
  launch_k(timestep_embedding_kernel, grid_for(count * dim / 2, 128), 128, 0, (cudaStream_t)stream, t, (int)count, dim, (__half*)out, ldo);
  count_launch();
  return check_cuda(cudaGetLastError(), "timestep_embedding launch");
}

extern "C" int seedx_unary_f16(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx, void* out, int64_t ldo, int act, void* stream) {
  SEEDX_REQUIRE(x && out && rows > 0 && cols > 0, "seedx_unary_f16: bad arguments");
  const int g = grid_for(rows * cols, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == SEEDX_F32) launch_k(unary_kernel<float>, g, 256, 0, st, (const float*)x, rows, (int)cols, ldx, (__half*)out, ldo, act);
  else if (x_dtype == SEEDX_F16) launch_k(unary_kernel<__half>, g, 256, 0, st, (const __half*)x, rows, (int)cols, ldx, (__half*)out, ldo, act);
  else SEEDX_REQUIRE(false, "seedx_unary_f16: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "unary launch");
}

extern "C" int seedx_softmax_rows(const void* x, int x_dtype, int64_t ldx, int64_t rows, int64_t cols, float scale, void* out, int64_t ldo,
                                  void* stream) {
  SEEDX_REQUIRE(x && out && rows > 0 && cols > 0, "seedx_softmax_rows: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == SEEDX_F32) launch_k(softmax_rows_kernel<float>, (unsigned)rows, 256, 0, st, (const float*)x, ldx, (int)cols, scale, (__half*)out, ldo);
  else if (x_dtype == SEEDX_F16) launch_k(softmax_rows_kernel<__half>, (unsigned)rows, 256, 0, st, (const __half*)x, ldx, (int)cols, scale, (__half*)out, ldo);
  else SEEDX_REQUIRE(false, "seedx_softmax_rows: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "softmax launch");
}

extern "C" int seedx_nchw_to_nhwc_f16(const float* x, int64_t n, int64_t c, int64_t hw, int64_t cpad, float scale, void* out, void* stream) {
  SEEDX_REQUIRE(x && out && cpad >= c, "seedx_nchw_to_nhwc_f16: bad arguments");
  launch_k(nchw_to_nhwc_kernel, grid_for(n * hw * cpad, 256), 256, 0, (cudaStream_t)stream, x, (int)n, (int)c, (int)hw, (int)cpad, scale, (__half*)out);
  count_launch();
  return check_cuda(cudaGetLastError(), "nchw_to_nhwc launch");
}

extern "C" int seedx_nhwc_to_nchw_f32(const void* x, int x_dtype, int64_t n, int64_t c, int64_t hw, int64_t ldc, float scale, float* out,
                                      void* stream) {
  SEEDX_REQUIRE(x && out && ldc >= c, "seedx_nhwc_to_nchw_f32: bad arguments");
  const int g = grid_for(n * c * hw, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == SEEDX_F32) launch_k(nhwc_to_nchw_kernel<float>, g, 256, 0, st, (const float*)x, (int)n, (int)c, (int)hw, (int)ldc, scale, out);
  else if (x_dtype == SEEDX_F16) launch_k(nhwc_to_nchw_kernel<__half>, g, 256, 0, st, (const __half*)x, (int)n, (int)c, (int)hw, (int)ldc, scale, out);
  else SEEDX_REQUIRE(false, "seedx_nhwc_to_nchw_f32: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "nhwc_to_nchw launch");
}

extern "C" int seedx_image_to_u8(const void* x, int x_dtype, int64_t pixels, int64_t ldc, uint8_t* out, void* stream) {
  SEEDX_REQUIRE(x && out && pixels > 0 && ldc >= 3, "seedx_image_to_u8: bad arguments");
  const int g = grid_for(pixels * 3, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == SEEDX_F32) launch_k(image_to_u8_kernel<float>, g, 256, 0, st, (const float*)x, pixels, (int)ldc, out);
  else if (x_dtype == SEEDX_F16) launch_k(image_to_u8_kernel<__half>, g, 256, 0, st, (const __half*)x, pixels, (int)ldc, out);
  else SEEDX_REQUIRE(false, "seedx_image_to_u8: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "image_to_u8 launch");
}

extern "C" int seedx_cfg_euler_step(const float* eps, float* x, void* unet_in, int64_t batch, int64_t hw, int branches, float guidance,
                                    float image_guidance, float sigma, float sigma_next, float init_sigma, void* stream) {
  SEEDX_REQUIRE(x && unet_in && (branches == 2 || branches == 3), "seedx_cfg_euler_step: bad arguments");
  if (eps) SEEDX_REQUIRE(sigma > 0.f, "seedx_cfg_euler_step: sigma must be > 0");
  launch_k(cfg_euler_kernel, grid_for(batch * hw * 4, 256), 256, 0, (cudaStream_t)stream, eps, x, (__half*)unet_in, (int)batch, (int)hw, branches,
                                                                                   guidance, image_guidance, sigma, sigma_next, init_sigma);
  count_launch();
  return check_cuda(cudaGetLastError(), "cfg_euler launch");
}

extern "C" int seedx_add_bcast_f16(const void* a, int a_dtype, const float* b, int64_t rows, int64_t cols, int64_t b_rows, void* out,
                                   void* stream) {
  SEEDX_REQUIRE(a && b && out && rows > 0 && cols > 0 && b_rows > 0, "seedx_add_bcast_f16: bad arguments");
  const int g = grid_for(rows * cols, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (a_dtype == SEEDX_F16) launch_k(add_bcast_kernel<__half>, g, 256, 0, st, (const __half*)a, b, rows, (int)cols, (int)b_rows, (__half*)out);
  else if (a_dtype == SEEDX_F32) launch_k(add_bcast_kernel<float>, g, 256, 0, st, (const float*)a, b, rows, (int)cols, (int)b_rows, (__half*)out);
  else SEEDX_REQUIRE(false, "seedx_add_bcast_f16: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "add_bcast launch");
}
