// Convolution WEIGHT gradient on tcgen05 (SURVEY G1 wgrad, G6/G7 backward) — any k x k, stride, dilation, padding.
//
//   dW[co, r, s, ci] = sum over output pixels (n, ho, wo) of  dy[n, ho, wo, co] * x[n, ho*st + r*dl - pd, wo*st + s*dl - pd, ci]
//
// As a GEMM the reduction (K) dimension is the PIXEL index, and both operands are stored pixel-major with the channel
// contiguous (NHWC): they are "MN-major" operands in UMMA terms.  No transposed copy is made: TMA drops [pixels x 32
// channels] boxes (128-byte rows) into shared memory.  For 32-bit (tf32) MN-major operands the tensor core accepts ONE
// shared-memory layout: 128-byte swizzle with a 32-byte atom (UMMA layout type SWIZZLE_128B_BASE32B, TMA swizzle mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: the four 32-byte chunks of a row are permuted by (row mod 4)); with the plain
// 16-byte-atom SWIZZLE_128B the MMA silently produces zeros (measured, profiles/r2_call2).  Canonical form (16-byte
// units):  ((8, m), (4, k)) : ((1, LBO), (8, SBO))  = 128-byte rows, k-groups of 4 rows SBO = 512 B apart, consecutive
// 32-channel groups LBO = one box apart; the instruction descriptor sets the a_major / b_major bits.
//
//   A operand (M side)  = x windows.  One box per (filter tap, 32-channel group); FOUR boxes form one M = 128 MMA, so
//                         taps are stacked along M.  The shifted window of a tap is just a different TMA coordinate
//                         (traversal stride = conv stride, zero fill outside the image = padding).
//   B operand (N side)  = dy, N = CO_T output channels (32..128).
//   D (TMEM, fp32)      = [128 lanes = (tap, ci)] x [CO_T columns = co] per group; `gpc` groups per CTA
//                         (gpc * CO_T <= 512 columns).
//
// Work decomposition: CTA = (group set) x (co tile) x (split of the pixel range).  Partial sums are accumulated into
// dW with `red.global.add.f32` — 32 lanes of a warp hold 32 consecutive ci, i.e. one 128-byte line per instruction.
// The caller provides dW zeroed, or the parameter's gradient buffer itself (accumulation is what autograd wants).
//
// Roles as in the forward kernels: warp 0 = TMA producer (elect.sync), warp 1 = TMEM alloc + MMA issuer, warps 2-5 =
// epilogue (TMEM lane quarter = warp & 3).  Reference sites (library calls there): /root/reference/src/simple_models.py
// :137-147,191 (ResNet), :249-265 (VAE), :441-451 (CPC).
#pragma once
#include "sm100.cuh"

namespace fedb200 {

constexpr int WG_THREADS = 192;
constexpr int WG_BK = 32;                       // pixels (GEMM-K rows) per box
constexpr int WG_BOX_BYTES = WG_BK * 128;       // 4 KB
constexpr int WG_MAX_STAGES = 6;

struct WgradParams {
  int kw, taps, stride, pad, dil;
  int nbox_ci;          // 32-channel groups of the input
  int nbox_a;           // taps * nbox_ci boxes on the M side
  int g_total;          // ceil(nbox_a / 4) MMA groups
  int gpc;              // groups per CTA
  int nb;               // dy boxes (CO_T / 32)
  int Co, Cw;           // real output channels; channels / innermost extent of dW
  int WB, HB, NBX;      // pixel box (WB * HB * NBX == WG_BK)
  int blocks_w, blocks_h;
  int kb_total, kb_per_split;
  int g_units, co_tiles;
  int stages, stage_bytes;
  uint32_t tmem_cols;
  float* dw;            // [Co, taps, Cw]
  int tma_red;          // 1: epilogue = shared-memory staging + bulk tensor reduce-add (needs (taps * Cw) % 4 == 0)
};

// MN-major tf32 operand, SWIZZLE_128B_BASE32B (layout type 1): start address, LBO = distance between 32-element
// groups, SBO = distance between groups of 4 k-rows (4 * 128 B)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;
  return d;
}
// fp32 accumulate, tf32 operands, BOTH operands MN-major (bits 15 / 16), tile M x N
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                  const __grid_constant__ CUtensorMap tmap_dw, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* stage_out = reinterpret_cast<float*>(smem + p.stages * p.stage_bytes);      // 4 warps x [32 co][32 ci] fp32
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes + 4 * 4096);
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* t_full = empty_bar + WG_MAX_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work unit -----------------------------------------------------------------------------------------------
  const int per_split = p.g_units * p.co_tiles;
  const int split = blockIdx.x / per_split;
  const int rest = blockIdx.x - split * per_split;
  const int co_t = rest / p.g_units;
  const int gu = rest - co_t * p.g_units;
  const int g0 = gu * p.gpc;
  const int ngroups = min(p.gpc, p.g_total - g0);
  const int co0 = co_t * p.nb * 32;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  const int a_region = 4 * p.gpc * WG_BOX_BYTES;      // B boxes follow the (fixed-size) A region of a stage

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
    if (p.tma_red) tma_prefetch_desc(&tmap_dw);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int nbox_load = min(4 * ngroups, p.nbox_a - 4 * g0);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int bw = kb % p.blocks_w;
        const int t = kb / p.blocks_w;
        const int bh = t % p.blocks_h;
        const int bn = t / p.blocks_h;
        const int wo0 = bw * p.WB, ho0 = bh * p.HB, n0 = bn * p.NBX;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* dst = tiles + s * p.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], uint32_t((nbox_load + p.nb) * WG_BOX_BYTES));
        int tap = (4 * g0) / p.nbox_ci;
        int cg = 4 * g0 - tap * p.nbox_ci;
        int r = tap / p.kw, sx = tap - r * p.kw;
        for (int a = 0; a < nbox_load; ++a) {
          tma_load_4d(dst + a * WG_BOX_BYTES, &tmap_x, &full_bar[s], cg * 32, wo0 * p.stride + sx * p.dil - p.pad,
                      ho0 * p.stride + r * p.dil - p.pad, n0);
          if (++cg == p.nbox_ci) {
            cg = 0;
            if (++sx == p.kw) { sx = 0; ++r; }
          }
        }
        for (int j = 0; j < p.nb; ++j)
          tma_load_4d(dst + a_region + j * WG_BOX_BYTES, &tmap_dy, &full_bar[s], co0 + 32 * j, wo0, ho0, n0);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32_mn(128, uint32_t(p.nb * 32));
      const uint32_t ncols = uint32_t(p.nb * 32);
      const uint64_t desc0 = make_mnmajor_sw128_desc(smem_u32(tiles), WG_BOX_BYTES);
      int s = 0;
      uint32_t ph = 0;
      bool first = true;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t sa = desc0 + uint64_t(uint32_t(s) * uint32_t(p.stage_bytes >> 4));
        const uint64_t sb = sa + uint64_t(a_region >> 4);
#pragma unroll
        for (int kk = 0; kk < WG_BK / 8; ++kk) {
          const uint64_t bdesc = sb + uint64_t(kk * (1024 >> 4));
          for (int g = 0; g < ngroups; ++g) {
            const uint64_t adesc = sa + uint64_t(g * (4 * WG_BOX_BYTES >> 4) + kk * (1024 >> 4));
            umma_tf32(tmem_base + uint32_t(g) * ncols, adesc, bdesc, idesc, (first && kk == 0) ? 0u : 1u);
          }
        }
        first = false;
        umma_commit(&empty_bar[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(t_full);
    }
  } else {
    // ===================== epilogue (warps 2..5): TMEM -> dW (accumulated) =====================
    // Partial sums of every CTA are ADDED into dW.  Per-element red.global.add is bound by the L2 atomic units
    // (~145 elements / clock chip-wide, measured: layer4 with 8 pixel splits = 18.9 M reds = 69 us of a 70 us kernel,
    // profiles/r2_wgrad.md), so the default path transposes each 32 (ci) x 32 (co) chunk through shared memory into
    // [co][ci] rows — the memory order of dW — and hands it to ONE bulk tensor reduce-add (cp.reduce.async.bulk.tensor,
    // 128-bit wide in L2, no LSU instructions).  Lanes outside the tensor (ci >= Cw, box >= nbox_a) contribute zeros;
    // rows / columns outside dW are clipped by the TMA unit.
    const int q = warp & 3;
    const uint32_t ncols = uint32_t(p.nb * 32);
    float* my_stage = stage_out + (warp - 2) * 1024;
    if (kb1 > kb0) {
      mbar_wait(t_full, 0);
      tc_fence_after();
      for (int g = 0; g < ngroups; ++g) {
        const int b = 4 * (g0 + g) + q;                 // this warp's box: 32 channels of one tap
        const int tap = b / p.nbox_ci;
        const int cg = b - tap * p.nbox_ci;
        const int ci = cg * 32 + lane;
        const bool box_ok = b < p.nbox_a;               // warp-uniform
        const bool row_ok = box_ok && ci < p.Cw;
        for (uint32_t c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(g) * ncols + c0, v);
          tmem_ld_wait();
          if (p.tma_red) {
            if (box_ok && co0 + int(c0) < p.Co) {
              if (lane == 0) tma_store_wait_read();      // the previous chunk's reduce has drained the staging tile
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; ++j) my_stage[j * 32 + lane] = row_ok ? __uint_as_float(v[j]) : 0.f;
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_reduce_add_2d(&tmap_dw, my_stage, tap * p.Cw + cg * 32, co0 + int(c0));
                tma_store_commit();
              }
            }
          } else if (row_ok) {
            float* dst = p.dw + (size_t(co0 + c0) * p.taps + tap) * p.Cw + ci;
            const size_t co_pitch = size_t(p.taps) * p.Cw;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (co0 + int(c0) + j < p.Co) red_add_f32(dst + j * co_pitch, __uint_as_float(v[j]));
          }
        }
      }
      if (p.tma_red && lane == 0) tma_store_wait_read();   // shared memory must outlive the last bulk reduce's read
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace fedb200
