// CTA-pair (cta_group::2) variant of the implicit-GEMM kernel: tile = 256 (M) x BLOCK_N per pair of SMs.
//
// Why (measured, profiles/r1_run6_conv_variants.log): with fp32/tf32 operands the single-CTA kernel is bound by
// SHARED-MEMORY bandwidth, not by the tensor pipe — every tcgen05.mma re-reads 128 A rows and all BLOCK_N B rows
// (32 B each) from smem while TMA is writing the next stage through the same 128 B/cycle port.  In a CTA pair each
// SM stores only HALF of the B tile and its own 128 A rows; the hardware exchanges the halves, so smem reads and TMA
// fills per SM drop by up to 1/3 and 1/2, and the pair retires a 256 x BLOCK_N x 8 MMA in the time one SM needs for
// 128 x BLOCK_N x 8.  This is the shape cuDNN/cuBLAS use for TF32 on sm_100 (256x128 tiles).
//
// Protocol (per stage s):
//   producers (warp 0 of BOTH CTAs): wait own empty[s]; even CTA arms ITS full[s] with expect_tx(2 x stage bytes);
//       both issue cta_group::2 TMA loads (own A rows, own half of B) whose bytes are credited to the even CTA's full[s];
//   MMA issuer (warp 1 of the EVEN CTA only): wait full[s]; 4 x tcgen05.mma.cta_group::2 (idesc M=256);
//       tcgen05.commit.cta_group::2 multicast -> empty[s] of both CTAs (and tmem_full of both after the last k-block);
//   epilogue (warps 2..5 of BOTH CTAs): wait own tmem_full; read own 128 accumulator rows from own TMEM.
#pragma once
#include "igemm_tcgen05.cuh"

namespace fedb200 {

template <int BLOCK_N, int STAGES>
struct Igemm2Smem {
  static constexpr int A_BYTES = IG_BLOCK_M * IG_BLOCK_K * 4;        // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BLOCK_N / 2) * IG_BLOCK_K * 4;     // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SCRATCH_BYTES = 4 * 32 * 33 * 4;
  static constexpr int PART_BYTES = 4 * BLOCK_N * 2 * 4;
  static constexpr int BAR_BYTES = (2 * STAGES + 1) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + SCRATCH_BYTES + PART_BYTES + BAR_BYTES + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm2_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const IgemmParams p) {
  using S = Igemm2Smem<BLOCK_N, STAGES>;
  static_assert(BLOCK_N % 64 == 0 && BLOCK_N >= 64 && BLOCK_N <= 256, "BLOCK_N in {64,128,192,256}");
  constexpr uint32_t TMEM_COLS = BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256);
  constexpr int B_HALF_ROWS = BLOCK_N / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* scratch = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
  float* part = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::SCRATCH_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::SCRATCH_BYTES + S::PART_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();            // 0 = even (leader), 1 = odd
  const bool leader = cta_rank == 0;
  const int m0 = blockIdx.x * IG_BLOCK_M;                  // blockIdx.x = 2*pair + rank
  const int n0 = blockIdx.y * BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_ptr_smem, TMEM_COLS);             // executed by the same warp of both CTAs
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int img = 0, h_base = 0;
      if (p.is_conv) {
        img = m0 / p.HW_out;
        h_base = (m0 % p.HW_out) / p.W_out;
      }
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = tiles + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + S::A_BYTES;
        if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * S::STAGE_BYTES);   // both CTAs' bytes land on this barrier
        const int tap = kb / p.cblocks;
        const int cb = kb - tap * p.cblocks;
        if (p.is_conv) {
          const int r = tap / p.taps_w, sx = tap - r * p.taps_w;
          tma_load_4d_2sm(a_dst, &tmap_a, &full_bar[s], cb * IG_BLOCK_K, sx * p.dil - p.pad,
                          h_base * p.stride + r * p.dil - p.pad, img);
        } else {
          tma_load_2d_2sm(a_dst, &tmap_a, &full_bar[s], kb * IG_BLOCK_K, m0);
        }
        tma_load_2d_2sm(b_dst, &tmap_b, &full_bar[s], tap * p.b_cols_per_tap + cb * IG_BLOCK_K,
                        n0 + int(cta_rank) * B_HALF_ROWS);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (even CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc(/*tf32*/ 2, 2 * IG_BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(tiles + s * S::STAGE_BYTES);
          const uint32_t b_addr = a_addr + S::A_BYTES;
          const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
          const uint64_t bdesc = make_kmajor_sw128_desc(b_addr);
#pragma unroll
          for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k)
            umma_tf32_2sm(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm_multicast(&empty_bar[s], 0b11);
          if (kb == p.num_k_blocks - 1) umma_commit_2sm_multicast(tmem_full_bar, 0b11);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    float* my_scratch = scratch + (warp - 2) * 32 * 33;
    float* my_part = part + (warp - 2) * BLOCK_N * 2;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(v[j]);
        const int col = n0 + c0 + j;
        if (p.bias != nullptr && col < p.N) x += __ldg(p.bias + col);
        if (p.act) x = elu1(x);
        f[j] = x;
      }
      if (row_ok) {
        float* dst = p.out + size_t(row) * p.ldo + n0 + c0;
        if (n0 + c0 + 32 <= p.N && (p.ldo & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < p.N) dst[j] = f[j];
        }
      }
      if (p.stats != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) my_scratch[lane * 33 + j] = row_ok ? f[j] : 0.f;
        __syncwarp();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float x = my_scratch[r * 33 + lane];
          s1 += x;
          s2 = fmaf(x, x, s2);
        }
        my_part[c0 + lane] = s1;
        my_part[BLOCK_N + c0 + lane] = s2;
        __syncwarp();
      }
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int t = threadIdx.x - 64;
      for (int c = t; c < BLOCK_N; c += 128) {
        if (n0 + c < p.N) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            s1 += part[w * BLOCK_N * 2 + c];
            s2 += part[w * BLOCK_N * 2 + BLOCK_N + c];
          }
          atomicAdd(p.stats + n0 + c, s1);
          atomicAdd(p.stats + p.N + n0 + c, s2);
        }
      }
    }
    tc_fence_before();
  }
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace fedb200
