// Fused block collectives over NVLink peer memory — FedAvg / FedProx / consensus-ADMM aggregation as ONE kernel
// (SURVEY G17-G19, §5.8).  No NCCL on this path.
//
// Every replica's parameter block is a slice of a flat arena that is mapped into every process (symmetric
// memory: CUDA VMM/IPC peer mappings, optionally bound to an NVSwitch multicast object).  A rank launches this
// kernel with the slice pointers of ALL K replicas (its own and its peers'); the kernel
//   A. meets the other ranks at a flag barrier in peer memory (st.release.sys / ld.acquire.sys, epoch counted),
//   1. reduces the K contributions straight out of peer memory — `ld.relaxed.sys.v4` per peer (one-shot), or a
//      single `multimem.ld_reduce.add.v4.f32` on the multicast address so the reduction happens in the switch —
//      applies the algorithm's scaling, writes the new consensus vector z and accumulates ||z_old - z_new||^2,
//   B. meets the peers again (everyone has finished READING everyone's x / y),
//   2. applies the local epilogue in place: FedAvg writes z back into every local replica's weights; FedProx
//      accumulates ||rho (x - z)||^2; ADMM performs the dual ascent y += rho (x - z) and accumulates the same norm,
//   C. exchanges the per-rank primal residuals through the control pads and finishes the scalars.
// Phases are separated by cooperative-groups grid barriers; nothing returns to the host in between, so a whole
// aggregation round costs one launch (+ one small D2H read of the residuals by the caller).
//
// mode: 0 = FedAvg (z = sum x / K, write-back), 1 = FedProx (no write-back), 2 = ADMM (z = sum(y + rho x)/(K rho)).
#include "fedb200.h"

#include <cooperative_groups.h>
#include <stdexcept>
#include <string>

namespace cg = cooperative_groups;

namespace fedb200 {

// control pad layout (uint32 words, one pad per rank, peer-mapped): 3 flag rows + 1 payload row of COMM_MAX_WORLD
constexpr int PAD_A = 0, PAD_B = 1, PAD_C = 2, PAD_PAYLOAD = 3;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
// in-switch reduction over all devices bound to the multicast object
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ float multimem_ld_reduce_f32(const float* mc) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
  return v;
}

__device__ __forceinline__ float warp_add(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_add(float v, float* sm) {
  v = warp_add(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    r = warp_add(r);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// Cross-rank barrier through the control pads.  Executed by block 0; thread t handles peer t.
__device__ __forceinline__ void peer_barrier(const CommArgs& a, int row, uint32_t epoch) {
  if (a.world <= 1) return;
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_release_sys(a.ctrl[peer] + row * COMM_MAX_WORLD + a.rank, epoch);
    const uint32_t* mine = a.ctrl[a.rank] + row * COMM_MAX_WORLD + peer;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) {  // ~10 s: name the missing rank instead of hanging
        printf("fedb200: rank %d timed out waiting for rank %d at barrier %d of aggregation %u\n", a.rank, peer, row, epoch);
        __trap();
      }
    }
  }
}

__global__ void __launch_bounds__(COMM_THREADS, 1) block_reduce_kernel(const CommArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float sm[32];
  const uint32_t epoch = a.sync[0] + 1;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  const int n4 = a.n >> 2;

  // ---- A: inputs of every rank are final -------------------------------------------------------------------
  peer_barrier(a, PAD_A, epoch);
  grid.sync();

  // ---- 1: reduce, scale, new z, dual residual -----------------------------------------------------------------
  float dual = 0.f, bad = 0.f;
  for (int i = tid; i < n4; i += nth) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.mc_x != nullptr) {
      acc = multimem_ld_reduce_v4(a.mc_x + 4 * size_t(i));
      if (a.mode == 2) {
        const float4 ys = multimem_ld_reduce_v4(a.mc_y + 4 * size_t(i));
        acc.x = fmaf(a.rho, acc.x, ys.x); acc.y = fmaf(a.rho, acc.y, ys.y);
        acc.z = fmaf(a.rho, acc.z, ys.z); acc.w = fmaf(a.rho, acc.w, ys.w);
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < a.K; ++k) {
        const float4 xv = ld_sys_v4(a.x[k] + 4 * size_t(i));
        if (a.mode == 2) {
          const float4 yv = ld_sys_v4(a.y[k] + 4 * size_t(i));
          acc.x += fmaf(a.rho, xv.x, yv.x); acc.y += fmaf(a.rho, xv.y, yv.y);
          acc.z += fmaf(a.rho, xv.z, yv.z); acc.w += fmaf(a.rho, xv.w, yv.w);
        } else {
          acc.x += xv.x; acc.y += xv.y; acc.z += xv.z; acc.w += xv.w;
        }
      }
    }
    const float4 zo = reinterpret_cast<const float4*>(a.z)[i];
    const float4 zn = make_float4(acc.x * a.inv_scale, acc.y * a.inv_scale, acc.z * a.inv_scale, acc.w * a.inv_scale);
    const float dx = zo.x - zn.x, dy = zo.y - zn.y, dz = zo.z - zn.z, dw = zo.w - zn.w;
    dual = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, dual))));
    if (!(isfinite(zn.x) && isfinite(zn.y) && isfinite(zn.z) && isfinite(zn.w))) bad += 1.f;
    reinterpret_cast<float4*>(a.z)[i] = zn;
  }
  for (int i = (n4 << 2) + tid; i < a.n; i += nth) {   // scalar tail
    float acc = 0.f;
    if (a.mc_x != nullptr) {
      acc = multimem_ld_reduce_f32(a.mc_x + i);
      if (a.mode == 2) acc = fmaf(a.rho, acc, multimem_ld_reduce_f32(a.mc_y + i));
    } else {
      for (int k = 0; k < a.K; ++k) {
        const float xv = ld_sys_f32(a.x[k] + i);
        acc += a.mode == 2 ? fmaf(a.rho, xv, ld_sys_f32(a.y[k] + i)) : xv;
      }
    }
    const float zn = acc * a.inv_scale;
    const float d = a.z[i] - zn;
    dual = fmaf(d, d, dual);
    if (!isfinite(zn)) bad += 1.f;
    a.z[i] = zn;
  }
  dual = block_add(dual, sm);
  bad = block_add(bad, sm);
  if (threadIdx.x == 0) {
    atomicAdd(a.out + 0, dual);
    if (bad != 0.f) atomicAdd(a.out + 2, bad);
  }
  grid.sync();

  // ---- B: every rank has finished reading the others' x / y --------------------------------------------------
  peer_barrier(a, PAD_B, epoch);
  grid.sync();

  // ---- 2: local epilogue in place -------------------------------------------------------------------------------
  for (int j = 0; j < a.n_local; ++j) {
    float* xl = a.xl[j];
    float* yl = a.yl[j];
    float pr = 0.f;
    for (int i = tid; i < n4; i += nth) {
      const float4 zv = reinterpret_cast<const float4*>(a.z)[i];
      if (a.mode == 0) {
        reinterpret_cast<float4*>(xl)[i] = zv;
      } else {
        const float4 xv = reinterpret_cast<const float4*>(xl)[i];
        const float4 yd = make_float4(a.rho * (xv.x - zv.x), a.rho * (xv.y - zv.y), a.rho * (xv.z - zv.z), a.rho * (xv.w - zv.w));
        pr = fmaf(yd.x, yd.x, fmaf(yd.y, yd.y, fmaf(yd.z, yd.z, fmaf(yd.w, yd.w, pr))));
        if (a.mode == 2) {
          float4 yv = reinterpret_cast<float4*>(yl)[i];
          yv.x += yd.x; yv.y += yd.y; yv.z += yd.z; yv.w += yd.w;
          reinterpret_cast<float4*>(yl)[i] = yv;
        }
      }
    }
    for (int i = (n4 << 2) + tid; i < a.n; i += nth) {
      const float zv = a.z[i];
      if (a.mode == 0) {
        xl[i] = zv;
      } else {
        const float yd = a.rho * (xl[i] - zv);
        pr = fmaf(yd, yd, pr);
        if (a.mode == 2) yl[i] += yd;
      }
    }
    if (a.mode != 0) {
      pr = block_add(pr, sm);
      if (threadIdx.x == 0) atomicAdd(a.out + 4 + j, pr);
    }
  }
  grid.sync();

  // ---- C: finish the scalars (sum over ALL workers of ||rho (x_k - z)||) -------------------------------------------
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      float local = 0.f;
      for (int j = 0; j < a.n_local; ++j) local += sqrtf(__ldcg(a.out + 4 + j));
      a.out[3] = local;
      if (a.world > 1) {
        for (int p = 0; p < a.world; ++p) a.ctrl[p][PAD_PAYLOAD * COMM_MAX_WORLD + a.rank] = __float_as_uint(local);
      }
    }
    __syncthreads();
    peer_barrier(a, PAD_C, epoch);
    __syncthreads();
    if (threadIdx.x == 0) {
      float total = a.out[3];
      if (a.world > 1) {
        total = 0.f;
        for (int p = 0; p < a.world; ++p)
          total += __uint_as_float(ld_acquire_sys(a.ctrl[a.rank] + PAD_PAYLOAD * COMM_MAX_WORLD + p));
      }
      a.out[1] = total;
      a.sync[0] = epoch;
    }
  }
}

void block_reduce_launch(const CommArgs& args, cudaStream_t s) {
  if (args.K < 1 || args.K > COMM_MAX_K || args.n_local > COMM_MAX_LOCAL || args.world > COMM_MAX_WORLD)
    throw std::runtime_error("fedb200: block_reduce: too many contributions / replicas / ranks");
  cudaMemsetAsync(args.out, 0, (4 + COMM_MAX_LOCAL) * sizeof(float), s);
  static int max_blocks = 0;
  if (max_blocks == 0) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, block_reduce_kernel, COMM_THREADS, 0);
    max_blocks = sms * (per < 1 ? 1 : 1);
  }
  int want = ((args.n >> 2) + COMM_THREADS - 1) / COMM_THREADS;
  int grid = want < 1 ? 1 : (want > max_blocks ? max_blocks : want);
  void* kargs[] = {(void*)&args};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)block_reduce_kernel, dim3(grid), dim3(COMM_THREADS), kargs, 0, s);
  if (e != cudaSuccess) throw std::runtime_error(std::string("fedb200: block_reduce launch: ") + cudaGetErrorString(e));
  count_launch();
}

}  // namespace fedb200
