// Weight-stationary, persistent 3x3/s1/p1 convolution for 64-output-channel layers on 32x32 maps
// (ResNet18 stem + layer1 forward and data gradient: 14 of the ~60 conv launches of a training step, and the
// slowest ones of the generic kernel: 90 us each, profiles/r1_run6_*).
//
// Design (follows the measurements in profiles/):
//  * The generic kernel is bound by how many bytes per cycle ONE SM can ingest from L2 (~40-60 B/cycle).  For these
//    layers the whole filter bank is only 9 x C_in x 64 fp32 = 72-144 KB: each persistent CTA loads it ONCE into
//    shared memory and keeps it there (weight traffic: 148 x 144 KB instead of 1024 x 144 KB).
//  * Activations use the halo formulation (conv_halo_tcgen05.cuh): one TMA box per 32-channel block serves all nine
//    filter taps through row-offset UMMA descriptors (activation traffic / ~5).
//  * With ~60 KB to ingest per tile of 128 outputs the kernel becomes tensor-pipe bound (72 MMAs of 128x64x8 per
//    tile); the accumulator is double-buffered in TMEM so the epilogue of tile i (TMEM -> registers -> global, BN
//    statistics) overlaps the MMAs of tile i+1.  BatchNorm partial sums are accumulated across all tiles of the CTA
//    in shared memory and flushed with one atomicAdd per channel per CTA at the end.
#pragma once
#include "conv_halo_tcgen05.cuh"

namespace fedb200 {

constexpr int WS_BN = 64;
constexpr int WS_A_SLOT = 30720;            // 240 rows x 128 B >= 7 x 34 halo rows
constexpr int WS_A_STAGES = 2;
constexpr int WS_W_TILE = WS_BN * 128;      // one tap, one 32-channel block: 8 KB

struct WsSmem {
  static constexpr int W_BYTES = 2 * 9 * WS_W_TILE;                    // up to 2 channel blocks: 144 KB
  static constexpr int A_BYTES = WS_A_STAGES * WS_A_SLOT;              // 60 KB
  static constexpr int SCRATCH_BYTES = 4 * 32 * 33 * 4;
  static constexpr int PART_BYTES = 4 * WS_BN * 2 * 4;
  static constexpr int BAR_BYTES = 16 * 8 + 16;
  static constexpr int TOTAL = W_BYTES + A_BYTES + SCRATCH_BYTES + PART_BYTES + BAR_BYTES + 1024;
};

__global__ void __launch_bounds__(IG_THREADS, 1)
conv3x3_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const HaloParams p, const int num_tiles) {
  constexpr uint32_t TMEM_COLS = 2 * WS_BN;   // two accumulator stages

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                                   // [cblocks][9][64 x 128 B]
  uint8_t* a_smem = smem + WsSmem::W_BYTES;                 // [2][A_SLOT]
  float* scratch = reinterpret_cast<float*>(a_smem + WsSmem::A_BYTES);
  float* part = reinterpret_cast<float*>(a_smem + WsSmem::A_BYTES + WsSmem::SCRATCH_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + WsSmem::A_BYTES + WsSmem::SCRATCH_BYTES + WsSmem::PART_BYTES);
  uint64_t* w_full = bars;                 // 1
  uint64_t* a_full = bars + 1;             // 2
  uint64_t* a_empty = bars + 3;            // 2
  uint64_t* t_full = bars + 5;             // 2
  uint64_t* t_empty = bars + 7;            // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int my_tiles = (num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 4);           // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== producer: weights once, then one halo box per (tile, channel block) =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, uint32_t(p.cblocks * 9 * WS_W_TILE));
      for (int cb = 0; cb < p.cblocks; ++cb)
        for (int t = 0; t < 9; ++t)
          tma_load_2d(w_smem + (cb * 9 + t) * WS_W_TILE, &tmap_b, w_full, t * p.C_in + cb * IG_BLOCK_K, 0);
      int it = 0;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int img = tile / p.tiles_per_img;
        const int u0 = (tile - img * p.tiles_per_img) * IG_BLOCK_M;
        const int hp_a = u0 / p.Wp;
        for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
          const int s = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&a_full[s], uint32_t(p.a_box_bytes));
          tma_load_4d(a_smem + s * WS_A_SLOT, &tmap_a, &a_full[s], cb * IG_BLOCK_K, -1, hp_a - 1, img);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(/*tf32*/ 2, IG_BLOCK_M, WS_BN);
    mbar_wait(w_full, 0);
    int it = 0;
    for (int j = 0; j < my_tiles; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int img = tile / p.tiles_per_img;
      const int u0 = (tile - img * p.tiles_per_img) * IG_BLOCK_M;
      const int row_off0 = u0 - (u0 / p.Wp) * p.Wp;
      const int acc = j & 1;
      mbar_wait(&t_empty[acc], ((j >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
      tc_fence_after();
      for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
        const int s = it & 1;
        mbar_wait(&a_full[s], (it >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_u32(a_smem + s * WS_A_SLOT);
          const uint32_t b_base = smem_u32(w_smem + cb * 9 * WS_W_TILE);
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            const int r = t / 3, sx = t - 3 * r;
            const uint64_t adesc = make_kmajor_sw128_desc(a_base + uint32_t(row_off0 + r * p.Wp + sx) * 128u);
            const uint64_t bdesc = make_kmajor_sw128_desc(b_base + uint32_t(t) * WS_W_TILE);
#pragma unroll
            for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k)
              umma_tf32(tmem_base + uint32_t(acc * WS_BN), adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc,
                        (cb | t | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[s]);
          if (cb == p.cblocks - 1) umma_commit(&t_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    float* my_scratch = scratch + (warp - 2) * 32 * 33;
    float* my_part = part + (warp - 2) * WS_BN * 2;
    for (int c = lane; c < WS_BN * 2; c += 32) my_part[c] = 0.f;
    __syncwarp();
    for (int j = 0; j < my_tiles; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int img = tile / p.tiles_per_img;
      const int u = (tile - img * p.tiles_per_img) * IG_BLOCK_M + q * 32 + lane;
      const int h = u / p.Wp, w = u - h * p.Wp;
      const bool row_ok = (img < p.NB) && (h < p.H) && (w < p.W);
      const size_t out_row = (size_t(img) * p.H + h) * p.W + w;
      const int acc = j & 1;
      mbar_wait(&t_full[acc], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < WS_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * WS_BN + c0), v);
        tmem_ld_wait();
        if (row_ok) {
          float* dst = p.out + out_row * p.C_out + c0;
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4)
            *reinterpret_cast<float4*>(dst + jj) = make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]),
                                                               __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
        }
        if (p.stats != nullptr) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) my_scratch[lane * 33 + jj] = row_ok ? __uint_as_float(v[jj]) : 0.f;
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float x = my_scratch[r * 33 + lane];
            s1 += x;
            s2 = fmaf(x, x, s2);
          }
          my_part[c0 + lane] += s1;
          my_part[WS_BN + c0 + lane] += s2;
          __syncwarp();
        }
      }
      // all of this warp's TMEM reads of accumulator `acc` are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int t = threadIdx.x - 64;
      if (t < WS_BN) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int wv = 0; wv < 4; ++wv) {
          s1 += part[wv * WS_BN * 2 + t];
          s2 += part[wv * WS_BN * 2 + WS_BN + t];
        }
        atomicAdd(p.stats + t, s1);
        atomicAdd(p.stats + p.C_out + t, s2);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace fedb200
