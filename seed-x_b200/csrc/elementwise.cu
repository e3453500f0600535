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
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = (TO)(float)s[i];
}

template <typename T>
__global__ void avgpool_tokens_kernel(const T* __restrict__ x, long long n_out_rows, int c, int k, T* __restrict__ out) {
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
    patchify_kernel<float><<<g, 256, 0, st>>>((const float*)x, (int)n, (int)c, (int)h, (int)w, (int)patch, (__half*)out, (int)kpad);
  else if (x_dtype == SEEDX_F16)
    patchify_kernel<__half><<<g, 256, 0, st>>>((const __half*)x, (int)n, (int)c, (int)h, (int)w, (int)patch, (__half*)out, (int)kpad);
  else
    SEEDX_REQUIRE(false, "seedx_patchify: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "patchify launch");
}

extern "C" int seedx_cast(const void* src, int sd, void* dst, int dd, int64_t count, void* stream) {
  SEEDX_REQUIRE(src && dst && count > 0, "seedx_cast: bad arguments");
  const int g = grid_for(count, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (sd == SEEDX_F32 && dd == SEEDX_F16) cast_kernel<float, __half><<<g, 256, 0, st>>>((const float*)src, (__half*)dst, count);
  else if (sd == SEEDX_F16 && dd == SEEDX_F32) cast_kernel<__half, float><<<g, 256, 0, st>>>((const __half*)src, (float*)dst, count);
  else SEEDX_REQUIRE(false, "seedx_cast: unsupported conversion %d -> %d", sd, dd);
  count_launch();
  return check_cuda(cudaGetLastError(), "cast launch");
}

extern "C" int seedx_avgpool_tokens(const void* x, int dtype, int64_t n, int64_t t, int64_t c, int64_t k, void* out, void* stream) {
  SEEDX_REQUIRE(x && out && k > 0 && t % k == 0, "seedx_avgpool_tokens: bad arguments");
  const long long rows = n * (t / k);
  const int g = grid_for(rows * c, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SEEDX_F16) avgpool_tokens_kernel<__half><<<g, 256, 0, st>>>((const __half*)x, rows, (int)c, (int)k, (__half*)out);
  else if (dtype == SEEDX_F32) avgpool_tokens_kernel<float><<<g, 256, 0, st>>>((const float*)x, rows, (int)c, (int)k, (float*)out);
  else SEEDX_REQUIRE(false, "seedx_avgpool_tokens: bad dtype");
  count_launch();
  return check_cuda(cudaGetLastError(), "avgpool launch");
}
