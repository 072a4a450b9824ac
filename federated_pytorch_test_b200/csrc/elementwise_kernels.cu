// Memory-bound fused elementwise kernels of the ResNet path: all NHWC fp32,
// 128-bit accesses along the channel axis, per-channel parameters staged in shared memory.
//  * normalize_u8_nhwc : uint8 NHWC pixels -> (x/255-mean)/std as NHWC (optionally padded to 4 channels) or NCHW
//  * col_stats        : per-channel sum / sum of squares (only for layers whose conv did not emit them)
//  * bn_elu_fwd        : BatchNorm(batch stats from the conv epilogue) + residual + ELU in ONE pass; block 0 also
//                        updates the running statistics and stores mean/invstd for the backward pass
//  * bn_elu_bwd_reduce / bn_elu_bwd_apply : the two passes of the fused ELU'+BN backward
//  * avgpool_nhwc (+bwd), weight_krsc_flip (dgrad weights: swap Cin/Cout, rotate taps by 180 degrees)
#include "fedb200.h"

#include <stdexcept>
#include <string>

namespace fedb200 {

static inline void check_launch(const char* name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("fedb200: ") + name + ": " + cudaGetErrorString(e));
  count_launch();
}
static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
__device__ __forceinline__ float elu_f(float v) { return v > 0.f ? v : (__expf(v) - 1.f); }
// d ELU(u)/du expressed through the output o = ELU(u):  u > 0 <=> o > 0;  u <= 0 => exp(u) = o + 1
__device__ __forceinline__ float elu_grad_from_out(float o) { return o > 0.f ? 1.f : (o + 1.f); }

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int npix, int c_out, float m0, float m1,
                    float m2, float s0, float s1, float s2, int to_nchw, int HW) {
  const float sc[3] = {1.f / (255.f * s0), 1.f / (255.f * s1), 1.f / (255.f * s2)};
  const float sh[3] = {-m0 / s0, -m1 / s1, -m2 / s2};
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const uint8_t* px = in + size_t(p) * 3;
    const float v0 = fmaf(float(px[0]), sc[0], sh[0]);
    const float v1 = fmaf(float(px[1]), sc[1], sh[1]);
    const float v2 = fmaf(float(px[2]), sc[2], sh[2]);
    if (to_nchw) {
      const int n = p / HW, r = p - n * HW;
      float* o = out + size_t(n) * 3 * HW + r;
      o[0] = v0;
      o[HW] = v1;
      o[2 * HW] = v2;
    } else if (c_out == 4) {
      reinterpret_cast<float4*>(out)[p] = make_float4(v0, v1, v2, 0.f);
    } else {
      float* o = out + size_t(p) * 3;
      o[0] = v0;
      o[1] = v1;
      o[2] = v2;
    }
  }
}
void normalize_u8_nhwc(const uint8_t* in, float* out, int npix, int c_out, const float* mean3, const float* std3,
                       int to_nchw, int H, int W, cudaStream_t s) {
  int grid = (npix + 255) / 256;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  normalize_u8_kernel<<<grid, 256, 0, s>>>(in, out, npix, c_out, mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                           std3[2], to_nchw, H * W);
  check_launch("normalize_u8");
}

// ------------------------------------------------------------------------------------------------
// Thread layout for [M, C] tensors with C % 4 == 0: thread t handles channel quad (t % (C/4)) for rows
// (t / (C/4)) + k * rows_per_iter.  Consecutive threads touch consecutive 16-B words of a row.
// ------------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 256;

__global__ void __launch_bounds__(EW_THREADS)
col_stats_kernel(const float* __restrict__ y, float* __restrict__ stats, int M, int C) {
  extern __shared__ float sm[];                   // [2][EW_THREADS][4]
  const int q = C >> 2;
  const int cq = threadIdx.x % q, r0 = threadIdx.x / q, rpi = EW_THREADS / q;
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  if (r0 < rpi) {
    const int step = gridDim.x * rpi;
    int r = blockIdx.x * rpi + r0;
    // four independent 16-B loads in flight per thread (latency, not bandwidth, bounded the one-row-per-trip loop)
    for (; r + 3 * step < M; r += 4 * step) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = reinterpret_cast<const float4*>(y + size_t(r + u * step) * C)[cq];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s1[0] += v[u].x; s1[1] += v[u].y; s1[2] += v[u].z; s1[3] += v[u].w;
        s2[0] = fmaf(v[u].x, v[u].x, s2[0]); s2[1] = fmaf(v[u].y, v[u].y, s2[1]);
        s2[2] = fmaf(v[u].z, v[u].z, s2[2]); s2[3] = fmaf(v[u].w, v[u].w, s2[3]);
      }
    }
    for (; r < M; r += step) {
      const float4 v = reinterpret_cast<const float4*>(y + size_t(r) * C)[cq];
      s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
      s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]); s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
    }
  }
  float* a = sm + threadIdx.x * 4;
  float* b = sm + EW_THREADS * 4 + threadIdx.x * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) { a[j] = s1[j]; b[j] = s2[j]; }
  __syncthreads();
  if (threadIdx.x < q) {
    float t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
    for (int rr = 0; rr < rpi; ++rr) {
      const float* pa = sm + (rr * q + threadIdx.x) * 4;
      const float* pb = sm + EW_THREADS * 4 + (rr * q + threadIdx.x) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) { t1[j] += pa[j]; t2[j] += pb[j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(stats + threadIdx.x * 4 + j, t1[j]);
      atomicAdd(stats + C + threadIdx.x * 4 + j, t2[j]);
    }
  }
}
void col_stats(const float* y, float* stats, int M, int C, cudaStream_t s) {
  if ((C & 3) || C > 4 * EW_THREADS) throw std::runtime_error("fedb200: col_stats needs C % 4 == 0 and C <= 1024");
  const int rpi = EW_THREADS / (C >> 2);
  int grid = (M + rpi * 8 - 1) / (rpi * 8);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  if (grid < 1) grid = 1;
  col_stats_kernel<<<grid, EW_THREADS, 2 * EW_THREADS * 4 * sizeof(float), s>>>(y, stats, M, C);
  check_launch("col_stats");
}

__global__ void __launch_bounds__(EW_THREADS)
bn_elu_fwd_kernel(const float* __restrict__ y, float* __restrict__ stats, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const float* __restrict__ residual, float* __restrict__ out,
                  float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                  float* __restrict__ save_invstd, int M, int C, float eps, float momentum, int act, int self_clean) {
  extern __shared__ float sm[];                   // scale[C] | shift[C]
  __shared__ int last_block;
  float* scale = sm;
  float* shift = sm + C;
  const float invM = 1.f / float(M);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = stats[c] * invM;
    float var = fmaf(-mean, mean, stats[C + c] * invM);
    var = var > 0.f ? var : 0.f;
    const float invstd = rsqrtf(var + eps);
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = fmaf(-mean, sc, beta[c]);
    if (blockIdx.x == 0) {
      save_mean[c] = mean;
      save_invstd[c] = invstd;
      if (running_mean != nullptr) {
        const float unbiased = M > 1 ? var * float(M) / float(M - 1) : var;
        running_mean[c] = fmaf(momentum, mean - running_mean[c], running_mean[c]);
        running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
      }
    }
  }
  __syncthreads();
  if (self_clean) {
    // Every block has now consumed the statistics.  The last one to say so zeroes the accumulators (and the counter
    // behind them) for the next convolution that uses this buffer: no memset launch per layer, CUDA-graph safe.
    unsigned int* counter = reinterpret_cast<unsigned int*>(stats + 2 * C);
    if (threadIdx.x == 0) {
      __threadfence();
      last_block = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_block) {
      for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) stats[c] = 0.f;
      if (threadIdx.x == 0) *counter = 0u;
    }
  }
  const int q = C >> 2;
  const size_t total = size_t(M) * q;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int cq = int(i % q);
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    const float4 sc = reinterpret_cast<const float4*>(scale)[cq];
    const float4 sh = reinterpret_cast<const float4*>(shift)[cq];
    float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    if (residual != nullptr) {
      const float4 r = reinterpret_cast<const float4*>(residual)[i];
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (act) { o.x = elu_f(o.x); o.y = elu_f(o.y); o.z = elu_f(o.z); o.w = elu_f(o.w); }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}
void bn_elu_fwd(const float* y, float* stats, const float* gamma, const float* beta, const float* residual,
                float* out, float* running_mean, float* running_var, float* save_mean, float* save_invstd, int M, int C,
                float eps, float momentum, int act, int self_clean, cudaStream_t s) {
  if (C & 3) throw std::runtime_error("fedb200: bn_elu_fwd needs C % 4 == 0");
  const size_t total = size_t(M) * (C >> 2);
  int grid = int((total + EW_THREADS * 4 - 1) / (EW_THREADS * 4));
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  if (grid < 1) grid = 1;
  bn_elu_fwd_kernel<<<grid, EW_THREADS, 2 * C * sizeof(float), s>>>(y, stats, gamma, beta, residual, out, running_mean,
                                                                    running_var, save_mean, save_invstd, M, C, eps,
                                                                    momentum, act, self_clean);
  check_launch("bn_elu_fwd");
}

// backward pass 1: sums[c] = sum_rows du, sums[C+c] = sum_rows du * xhat, with du = dout * ELU'(out)
__global__ void __launch_bounds__(EW_THREADS)
bn_elu_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ y,
                         const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ sums,
                         int M, int C, int act) {
  extern __shared__ float sm[];
  const int q = C >> 2;
  const int cq = threadIdx.x % q, r0 = threadIdx.x / q, rpi = EW_THREADS / q;
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  if (r0 < rpi) {
    const float4 mu = reinterpret_cast<const float4*>(mean)[cq];
    const float4 is = reinterpret_cast<const float4*>(invstd)[cq];
    const int step = gridDim.x * rpi;
    const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
    for (int r = blockIdx.x * rpi + r0; r < M; r += 2 * step) {
      // two rows (six 16-B loads) in flight per thread
      const bool two = r + step < M;
      const size_t i0 = size_t(r) * q + cq, i1 = size_t(two ? r + step : r) * q + cq;
      float4 d0 = reinterpret_cast<const float4*>(dout)[i0], d1 = reinterpret_cast<const float4*>(dout)[i1];
      const float4 o0 = act ? reinterpret_cast<const float4*>(out)[i0] : one;
      const float4 o1 = act ? reinterpret_cast<const float4*>(out)[i1] : one;
      const float4 v0 = reinterpret_cast<const float4*>(y)[i0], v1 = reinterpret_cast<const float4*>(y)[i1];
      if (act) {
        d0.x *= elu_grad_from_out(o0.x); d0.y *= elu_grad_from_out(o0.y); d0.z *= elu_grad_from_out(o0.z); d0.w *= elu_grad_from_out(o0.w);
        d1.x *= elu_grad_from_out(o1.x); d1.y *= elu_grad_from_out(o1.y); d1.z *= elu_grad_from_out(o1.z); d1.w *= elu_grad_from_out(o1.w);
      }
      if (!two) d1 = make_float4(0.f, 0.f, 0.f, 0.f);
      s1[0] += d0.x + d1.x; s1[1] += d0.y + d1.y; s1[2] += d0.z + d1.z; s1[3] += d0.w + d1.w;
      s2[0] = fmaf(d0.x, (v0.x - mu.x) * is.x, fmaf(d1.x, (v1.x - mu.x) * is.x, s2[0]));
      s2[1] = fmaf(d0.y, (v0.y - mu.y) * is.y, fmaf(d1.y, (v1.y - mu.y) * is.y, s2[1]));
      s2[2] = fmaf(d0.z, (v0.z - mu.z) * is.z, fmaf(d1.z, (v1.z - mu.z) * is.z, s2[2]));
      s2[3] = fmaf(d0.w, (v0.w - mu.w) * is.w, fmaf(d1.w, (v1.w - mu.w) * is.w, s2[3]));
    }
  }
  float* a = sm + threadIdx.x * 4;
  float* b = sm + EW_THREADS * 4 + threadIdx.x * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) { a[j] = s1[j]; b[j] = s2[j]; }
  __syncthreads();
  if (threadIdx.x < q) {
    float t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
    for (int rr = 0; rr < rpi; ++rr) {
      const float* pa = sm + (rr * q + threadIdx.x) * 4;
      const float* pb = sm + EW_THREADS * 4 + (rr * q + threadIdx.x) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) { t1[j] += pa[j]; t2[j] += pb[j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(sums + threadIdx.x * 4 + j, t1[j]);
      atomicAdd(sums + C + threadIdx.x * 4 + j, t2[j]);
    }
  }
}
void bn_elu_bwd_reduce(const float* dout, const float* out, const float* y, const float* mean, const float* invstd,
                       float* sums, int M, int C, int act, cudaStream_t s) {
  if ((C & 3) || C > 4 * EW_THREADS) throw std::runtime_error("fedb200: bn_elu_bwd needs C % 4 == 0 and C <= 1024");
  cudaMemsetAsync(sums, 0, 2 * C * sizeof(float), s);
  const int rpi = EW_THREADS / (C >> 2);
  int grid = (M + rpi * 8 - 1) / (rpi * 8);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  if (grid < 1) grid = 1;
  bn_elu_bwd_reduce_kernel<<<grid, EW_THREADS, 2 * EW_THREADS * 4 * sizeof(float), s>>>(dout, out, y, mean, invstd, sums,
                                                                                        M, C, act);
  check_launch("bn_elu_bwd_reduce");
}

// backward pass 2: dy = gamma*invstd*(du - sum_du/M - xhat*sum_du_xhat/M); dres = du; dgamma/dbeta from sums
__global__ void __launch_bounds__(EW_THREADS)
bn_elu_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ y,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, const float* __restrict__ sums, float* __restrict__ dy,
                        float* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int C,
                        int act) {
  extern __shared__ float sm[];                   // a[C] | b[C] | c[C] | mu[C]:  dy = a*du + b*y + c  (affine in du, y)
  float* ca = sm;
  float* cb = sm + C;
  float* cc = sm + 2 * C;
  const float invM = 1.f / float(M);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gamma[c], is = invstd[c], mu = mean[c];
    const float sdu = sums[c], sdx = sums[C + c];
    // dy = g*is*(du - sdu/M - (y-mu)*is*sdx/M)
    const float k = g * is;
    ca[c] = k;
    cb[c] = -k * is * sdx * invM;
    cc[c] = -k * sdu * invM + k * is * sdx * invM * mu;
    if (blockIdx.x == 0) {
      if (dgamma != nullptr) dgamma[c] += sdx;
      if (dbeta != nullptr) dbeta[c] += sdu;
    }
  }
  __syncthreads();
  const int q = C >> 2;
  const size_t total = size_t(M) * q;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int cq = int(i % q);
    float4 d = reinterpret_cast<const float4*>(dout)[i];
    if (act) {
      const float4 o = reinterpret_cast<const float4*>(out)[i];
      d.x *= elu_grad_from_out(o.x); d.y *= elu_grad_from_out(o.y); d.z *= elu_grad_from_out(o.z); d.w *= elu_grad_from_out(o.w);
    }
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    const float4 a = reinterpret_cast<const float4*>(ca)[cq];
    const float4 b = reinterpret_cast<const float4*>(cb)[cq];
    const float4 c = reinterpret_cast<const float4*>(cc)[cq];
    float4 r = make_float4(fmaf(a.x, d.x, fmaf(b.x, v.x, c.x)), fmaf(a.y, d.y, fmaf(b.y, v.y, c.y)),
                           fmaf(a.z, d.z, fmaf(b.z, v.z, c.z)), fmaf(a.w, d.w, fmaf(b.w, v.w, c.w)));
    reinterpret_cast<float4*>(dy)[i] = r;
    if (dres != nullptr) reinterpret_cast<float4*>(dres)[i] = d;
  }
}
void bn_elu_bwd_apply(const float* dout, const float* out, const float* y, const float* mean, const float* invstd,
                      const float* gamma, const float* sums, float* dy, float* dres, float* dgamma, float* dbeta, int M,
                      int C, int act, cudaStream_t s) {
  const size_t total = size_t(M) * (C >> 2);
  int grid = int((total + EW_THREADS * 4 - 1) / (EW_THREADS * 4));
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  if (grid < 1) grid = 1;
  bn_elu_bwd_apply_kernel<<<grid, EW_THREADS, 3 * C * sizeof(float), s>>>(dout, out, y, mean, invstd, gamma, sums, dy,
                                                                          dres, dgamma, dbeta, M, C, act);
  check_launch("bn_elu_bwd_apply");
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
avgpool_kernel(const float* __restrict__ x, float* __restrict__ out, int NB, int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over NB*C
  if (i >= NB * C) return;
  const int n = i / C, c = i - n * C;
  const float* p = x + size_t(n) * HW * C + c;
  float s = 0.f;
  for (int k = 0; k < HW; ++k) s += p[size_t(k) * C];
  out[i] = s / float(HW);
}
void avgpool_nhwc(const float* x, float* out, int NB, int HW, int C, cudaStream_t s) {
  avgpool_kernel<<<(NB * C + 255) / 256, 256, 0, s>>>(x, out, NB, HW, C);
  check_launch("avgpool_nhwc");
}
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int NB, int HW, int C) {
  const size_t total = size_t(NB) * HW * C;
  const float inv = 1.f / float(HW);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    const int n = int(i / (size_t(HW) * C));
    dx[i] = dout[size_t(n) * C + c] * inv;
  }
}
void avgpool_nhwc_bwd(const float* dout, float* dx, int NB, int HW, int C, cudaStream_t s) {
  const size_t total = size_t(NB) * HW * C;
  int grid = int((total + 1023) / 1024);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  avgpool_bwd_kernel<<<grid, 256, 0, s>>>(dout, dx, NB, HW, C);
  check_launch("avgpool_nhwc_bwd");
}

// w: [C_out][kh][kw][C_in]  ->  out: [C_in][kh][kw][C_out] with taps rotated by 180 degrees
// (the stride-1 data gradient is a convolution of dY with these weights)
__global__ void __launch_bounds__(256)
weight_flip_kernel(const float* __restrict__ w, float* __restrict__ out, int C_out, int C_in, int kh, int kw) {
  const int taps = kh * kw;
  const size_t total = size_t(C_out) * taps * C_in;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int co = int(i % C_out);
    const int t = int((i / C_out) % taps);
    const int ci = int(i / (size_t(C_out) * taps));
    out[i] = w[(size_t(co) * taps + (taps - 1 - t)) * C_in + ci];
  }
}
void weight_krsc_flip(const float* w, float* out, int C_out, int C_in, int kh, int kw, cudaStream_t s) {
  const size_t total = size_t(C_out) * kh * kw * C_in;
  int grid = int((total + 255) / 256);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  weight_flip_kernel<<<grid, 256, 0, s>>>(w, out, C_out, C_in, kh, kw);
  check_launch("weight_krsc_flip");
}

}  // namespace fedb200
