// Micro-probe 2 (B200): cost of the mbarrier primitives from ONE thread, the depth of the tcgen05.mma issue queue,
// and producer/consumer ring variants (spin-loop flavour, lookahead try_wait).
#include <cstdio>
#include "sm100.cuh"
using namespace fedb200;

__device__ __forceinline__ void wait_tight(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void wait_poll(uint64_t* bar, uint32_t parity) { while (!test_wait(bar, parity)) {} }

// variant: 0 = repo mbar_wait (try_wait + bounded-spin bookkeeping), 1 = tight try_wait loop, 2 = test_wait polling
template <int V>
__device__ __forceinline__ void W(uint64_t* bar, uint32_t parity) {
  if (V == 0) mbar_wait(bar, parity); else if (V == 1) wait_tight(bar, parity); else wait_poll(bar, parity);
}

__global__ void __launch_bounds__(128, 1) prim(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 0) {
    const int n = 1024;
    // (a) self arrive + wait on the completed phase, three wait flavours
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < n; ++i) { mbar_arrive(&bar); mbar_wait(&bar, ph); ph ^= 1; }
    long long t1 = clock64();
    for (int i = 0; i < n; ++i) { mbar_arrive(&bar); wait_tight(&bar, ph); ph ^= 1; }
    long long t2 = clock64();
    for (int i = 0; i < n; ++i) { mbar_arrive(&bar); wait_poll(&bar, ph); ph ^= 1; }
    long long t3 = clock64();
    for (int i = 0; i < n; ++i) { mbar_arrive(&bar); }        // arrive only (phase flips every time)
    long long t4 = clock64();
    if (n & 1) ph ^= 0;
    out[0] = (t1 - t0) / n; out[1] = (t2 - t1) / n; out[2] = (t3 - t2) / n; out[3] = (t4 - t3) / n;
    // (b) commit + wait with an empty tensor pipe
    uint32_t ph2 = 0;
    long long t5 = clock64();
    for (int i = 0; i < n; ++i) { umma_commit(&bar2); wait_tight(&bar2, ph2); ph2 ^= 1; }
    long long t6 = clock64();
    out[4] = (t6 - t5) / n;
    // (c) issue-queue depth: time stamps after each of 24 MMAs (128x64x8 tf32) issued into an idle pipe
    constexpr uint32_t idesc = make_idesc(2, 128, 64);
    const uint64_t ad = make_kmajor_sw128_desc(smem_u32(smem)), bd = make_kmajor_sw128_desc(smem_u32(smem + 65536));
    long long ts[25];
    ts[0] = clock64();
#pragma unroll
    for (int i = 0; i < 24; ++i) { umma_tf32(tmem_base, ad + uint64_t(2 * (i & 3)), bd + uint64_t(2 * (i & 3)), idesc, 1u); ts[i + 1] = clock64(); }
    umma_commit(&bar2); wait_tight(&bar2, ph2); ph2 ^= 1;
    long long te = clock64();
    for (int i = 0; i < 24; ++i) out[8 + i] = ts[i + 1] - ts[0];
    out[7] = te - ts[0];
    // (d) same with N=256
    constexpr uint32_t idesc2 = make_idesc(2, 128, 256);
    ts[0] = clock64();
#pragma unroll
    for (int i = 0; i < 24; ++i) { umma_tf32(tmem_base, ad + uint64_t(2 * (i & 3)), bd + uint64_t(2 * (i & 3)), idesc2, 1u); ts[i + 1] = clock64(); }
    umma_commit(&bar2); wait_tight(&bar2, ph2); ph2 ^= 1;
    te = clock64();
    for (int i = 0; i < 24; ++i) out[40 + i] = ts[i + 1] - ts[0];
    out[39] = te - ts[0];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ring: V = wait flavour; LOOK = consumer peeks the NEXT stage's barrier before releasing the current one;
// nmma MMAs (128 x N x 8) per iteration; release by tcgen05.commit
template <int STAGES, int V, int N>
__global__ void __launch_bounds__(128, 1) ring(int iters, int nmma, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STAGES], empty[STAGES];
  __shared__ uint32_t tmem_ptr;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      W<V>(&empty[s], ph ^ 1);
      mbar_arrive(&full[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc(2, 128, N);
    const uint64_t ad = make_kmajor_sw128_desc(smem_u32(smem)), bd = make_kmajor_sw128_desc(smem_u32(smem + 65536));
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      W<V>(&full[s], ph);
      tc_fence_after();
      for (int k = 0; k < nmma; ++k) umma_tf32(tmem_base, ad + uint64_t(2 * (k & 3)), bd + uint64_t(2 * (k & 3)), idesc, 1u);
      umma_commit(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static long long* d;
static const int SMEM = 129 * 1024 + 1024;
template <int STAGES, int V, int N>
void run_ring(int nmma) {
  cudaFuncSetAttribute(ring<STAGES, V, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  ring<STAGES, V, N><<<1, 128, SMEM>>>(8192, nmma, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("ring stages=%d wait=%d N=%3d mma/iter=%2d : %7.1f cycles/iteration\n", STAGES, V, N, nmma, double(h) / 8192);
}

int main() {
  cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(prim, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  prim<<<1, 128, SMEM>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("arrive+mbar_wait %lld  arrive+tight %lld  arrive+test_wait poll %lld  arrive only %lld  commit+wait(empty pipe) %lld\n",
         h[0], h[1], h[2], h[3], h[4]);
  printf("N=64  issue timestamps:"); for (int i = 0; i < 24; ++i) printf(" %lld", h[8 + i]); printf("  | all retired %lld\n", h[7]);
  printf("N=256 issue timestamps:"); for (int i = 0; i < 24; ++i) printf(" %lld", h[40 + i]); printf("  | all retired %lld\n", h[39]);
  for (int nmma : {0, 4, 8, 16}) {
    run_ring<4, 0, 64>(nmma); run_ring<4, 1, 64>(nmma); run_ring<4, 2, 64>(nmma);
    run_ring<4, 1, 128>(nmma); run_ring<4, 1, 256>(nmma);
  }
  return 0;
}
