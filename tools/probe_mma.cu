// Micro-probe (B200): (1) cycles per tcgen05.mma (SS mode, K-major SW128 operands) vs tile shape / number of
// independent accumulators, with a lean issue loop (no integer division, descriptors pre-built, 8x unrolled);
// (2) cycles per iteration of the producer/consumer mbarrier ring used by the GEMM kernels, with nothing in it.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I federated_pytorch_test_b200/csrc tools/probe_mma.cu -o tools/_bin/probe_mma
#include <cstdio>
#include "sm100.cuh"
using namespace fedb200;

template <int M, int N, int NACC, bool BF16, bool VARY>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(BF16 ? 1 : 2, M, N);
    const uint64_t ad = make_kmajor_sw128_desc(smem_u32(smem)), bd = make_kmajor_sw128_desc(smem_u32(smem + 64 * 1024));
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // VARY: halo-style operand addresses (row offsets that are not swizzle-atom aligned) and a different B tile
        const uint64_t a = ad + uint64_t(VARY ? ((j * 34 + (j & 3)) * 8) : 0) + uint64_t(2 * (j & 3));
        const uint64_t b = bd + uint64_t(VARY ? (j * 256) : 0) + uint64_t(2 * (j & 3));
        const uint32_t d = tmem_base + uint32_t((j % NACC) * N);
        if (BF16) umma_bf16(d, a, b, idesc, 1u); else umma_tf32(d, a, b, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// mode bit0: consumer releases the slot with tcgen05.commit (else a plain mbarrier.arrive)
// mode bit1: consumer issues 4 MMAs (128x64x8 tf32) per iteration before the release
// mode bit2: all 32 lanes poll (production kernels) instead of lane 0 only
template <int STAGES>
__global__ void __launch_bounds__(128, 1) handshake(int iters, int mode, long long* cycles) {
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
  const bool all = mode & 4;
  const long long t0 = clock64();
  if (warp == 0 && (all || lane == 0)) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) mbar_arrive(&full[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && (all || lane == 0)) {
    constexpr uint32_t idesc = make_idesc(2, 128, 64);
    const uint64_t ad = make_kmajor_sw128_desc(smem_u32(smem)), bd = make_kmajor_sw128_desc(smem_u32(smem + 65536));
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        if (mode & 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, ad + uint64_t(2 * k), bd + uint64_t(2 * k), idesc, 1u);
        }
        if (mode & 1) umma_commit(&empty[s]); else mbar_arrive(&empty[s]);
      }
      if (all) __syncwarp();
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

template <int M, int N, int NACC, bool BF16, bool VARY>
void run(int grid) {
  cudaFuncSetAttribute(probe<M, N, NACC, BF16, VARY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int iters = 4096;
  probe<M, N, NACC, BF16, VARY><<<grid, 128, SMEM>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-5s M=%3d N=%3d nacc=%d vary=%d grid=%3d  issue %7.1f  total %7.1f cycles/mma\n", BF16 ? "bf16" : "tf32", M, N,
         NACC, int(VARY), grid, double(h[0]) / iters, double(h[1]) / iters);
}
template <int M, int N>
void run_shape() {
  run<M, N, 1, false, false>(1);
  run<M, N, 2, false, false>(1);
  run<M, N, 1, false, true>(1);
  run<M, N, 1, false, false>(148);
  run<M, N, 1, true, false>(1);
}
template <int STAGES>
void run_hs() {
  cudaFuncSetAttribute(handshake<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  for (int mode = 0; mode < 8; ++mode) {
    handshake<STAGES><<<1, 128, SMEM>>>(8192, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("handshake stages=%d commit=%d mma=%d all_lanes=%d  %7.1f cycles/iteration\n", STAGES, mode & 1, (mode >> 1) & 1,
           (mode >> 2) & 1, double(h[0]) / 8192);
  }
}

int main() {
  cudaMalloc(&d, 16);
  run_shape<128, 32>(); run_shape<128, 64>(); run_shape<128, 128>(); run_shape<128, 256>();
  run_shape<64, 64>(); run_shape<64, 128>(); run_shape<64, 256>();
  run_hs<2>(); run_hs<4>(); run_hs<6>();
  return 0;
}
