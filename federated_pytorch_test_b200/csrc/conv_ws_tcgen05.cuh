// Weight-stationary, persistent 3x3/s1/p1 convolution (ResNet18 stem, layer1 and layer2, forward and data gradient:
// about half of the conv launches of a training step, and the slowest ones of the generic kernel).
//
// Design (follows the measurements in profiles/ROOFLINE.md):
//  * The generic kernel is bound by how many bytes per cycle ONE SM can ingest from L2 (~40-60 B/cycle).  Here each
//    persistent CTA owns a slice of BN output channels whose filters (9 x C_in x BN fp32 <= 144 KB) are loaded ONCE
//    into shared memory and stay there: weight traffic drops from (#tiles x filter bytes) to (#CTAs x slice bytes).
//    BN = 64 covers all outputs of the 64-channel layers; the 128-channel layers are cut into four 32-channel slices
//    (CTA i serves slice i % 4), trading four passes over the (5x cheaper, see below) activations for stationary weights.
//  * Activations use the halo formulation (conv_halo_tcgen05.cuh): one TMA box per 32-channel block serves all nine
//    filter taps through row-offset UMMA descriptors (activation traffic / ~5).
//  * The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1; the epilogue
//    stages each 32x32 chunk in shared memory and writes 4 rows x 128 B per store instruction (a thread-per-row
//    store pattern made the epilogue, not the MMAs, the bottleneck of the first version: 35 us -> see profiles/).
//    BatchNorm partial sums are accumulated across all tiles of the CTA and flushed once at the end.
#pragma once
#include "conv_halo_tcgen05.cuh"

namespace fedb200 {

constexpr int WS_A_STAGES = 2;
constexpr int WS_SCR_LD = 36;               // staging row pitch in floats: 16-B aligned rows, conflict-free column reads

template <int BN, int A_SLOT, int MAX_CB>
struct WsSmem {
  static constexpr int W_TILE = BN * 128;                              // one tap, one 32-channel block
  static constexpr int W_BYTES = MAX_CB * 9 * W_TILE;                  // <= 144 KB
  static constexpr int A_BYTES = WS_A_STAGES * A_SLOT;
  static constexpr int SCRATCH_BYTES = 4 * 32 * WS_SCR_LD * 4;
  static constexpr int PART_BYTES = 4 * BN * 2 * 4;
  static constexpr int BAR_BYTES = 16 * 8 + 16;
  static constexpr int TOTAL = W_BYTES + A_BYTES + SCRATCH_BYTES + PART_BYTES + BAR_BYTES + 1024;
};

template <int BN, int A_SLOT, int MAX_CB>
__global__ void __launch_bounds__(IG_THREADS, 1)
conv3x3_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const HaloParams p, const int num_tiles, const int n_slices) {
  using S = WsSmem<BN, A_SLOT, MAX_CB>;
  static_assert(BN == 32 || BN == 64, "output-channel slice of 32 or 64");
  static_assert(A_SLOT % 1024 == 0 && S::W_TILE % 1024 == 0, "swizzle atoms need 1024-B aligned tiles");
  constexpr uint32_t TMEM_COLS = 2 * BN;      // two accumulator stages

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                                   // [cblocks][9][BN x 128 B]
  uint8_t* a_smem = smem + S::W_BYTES;                      // [2][A_SLOT]
  float* scratch = reinterpret_cast<float*>(a_smem + S::A_BYTES);
  float* part = reinterpret_cast<float*>(a_smem + S::A_BYTES + S::SCRATCH_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + S::A_BYTES + S::SCRATCH_BYTES + S::PART_BYTES);
  uint64_t* w_full = bars;                 // 1
  uint64_t* a_full = bars + 1;             // 2
  uint64_t* a_empty = bars + 3;            // 2
  uint64_t* t_full = bars + 5;             // 2
  uint64_t* t_empty = bars + 7;            // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA -> (output-channel slice, first tile, tile stride).  gridDim.x is a multiple of n_slices.
  const int slice = int(blockIdx.x) % n_slices;
  const int n0 = slice * BN;
  const int tile0 = int(blockIdx.x) / n_slices;
  const int tstride = int(gridDim.x) / n_slices;
  const int my_tiles = tile0 < num_tiles ? (num_tiles - tile0 + tstride - 1) / tstride : 0;

  pdl_launch_dependents();
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
  pdl_wait();

  if (warp == 0) {
    // ===================== producer: this slice's filters once, then one halo box per (tile, channel block) =========
    if (elect_one() && my_tiles > 0) {
      mbar_arrive_expect_tx(w_full, uint32_t(p.cblocks * 9 * S::W_TILE));
      for (int cb = 0; cb < p.cblocks; ++cb)
        for (int t = 0; t < 9; ++t)
          tma_load_2d(w_smem + (cb * 9 + t) * S::W_TILE, &tmap_b, w_full, t * p.C_in + cb * IG_BLOCK_K, n0);
      int it = 0;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = tile0 + j * tstride;
        const int img = tile / p.tiles_per_img;
        const int u0 = (tile - img * p.tiles_per_img) * IG_BLOCK_M;
        const int hp_a = u0 / p.Wp;
        for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
          const int s = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          if (p.dbg & 1) { mbar_arrive(&a_full[s]); continue; }
          mbar_arrive_expect_tx(&a_full[s], uint32_t(p.a_box_bytes));
          // several small boxes in flight instead of one big one: the TMA unit walks a box row by row (~25 cycles
          // per 128-B row), but overlaps different boxes.  Rows past R are never read by the MMAs.
          for (int h = 0; h < p.R; h += p.box_h)
            tma_load_4d(a_smem + s * A_SLOT + h * p.Wp * 128, &tmap_a, &a_full[s], cb * IG_BLOCK_K, -1, hp_a - 1 + h, img);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(/*tf32*/ 2, IG_BLOCK_M, BN);
    // ONE thread runs the whole issue loop.  Its scalar instructions are on the critical path (a lone thread issues
    // roughly one dependent instruction every 5-10 cycles, and a 128x64x8 MMA only lasts 48 cycles: measured in
    // tools/probe_mma.cu), so the loop is kept free of divisions and descriptor rebuilds: all 36 MMAs of a
    // (tile, channel block) are unrolled and every descriptor is base + compile-time/loop-invariant offset.
    if (elect_one() && my_tiles > 0) {
      mbar_wait(w_full, 0);
      uint32_t tap_off[9];                                   // (r * Wp + sx) pixel rows of 128 B, encoded >> 4
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_off[t] = (p.dbg & 8) ? 0u : uint32_t((t / 3) * p.Wp + (t % 3)) * 8u;
      const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(a_smem));
      const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(w_smem));
      uint32_t s = 0, a_ph = 0;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = tile0 + j * tstride;
        const int img = tile / p.tiles_per_img;
        const int u0 = (tile - img * p.tiles_per_img) * IG_BLOCK_M;
        const uint32_t row_off0 = uint32_t(u0 - (u0 / p.Wp) * p.Wp);
        const uint32_t acc = uint32_t(j & 1);
        const uint32_t d_tmem = tmem_base + acc * BN;
        mbar_wait(&t_empty[acc], ((j >> 1) & 1) ^ 1);        // epilogue has drained this accumulator
        tc_fence_after();
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(&a_full[s], a_ph);
          tc_fence_after();
          const uint64_t ad = a_desc0 + uint64_t(s * (A_SLOT >> 4) + row_off0 * 8u);
          const uint64_t bd = b_desc0 + uint64_t(uint32_t(cb) * 9u * (S::W_TILE >> 4));
          if (!(p.dbg & 2)) {
            umma_tf32(d_tmem, ad + tap_off[0], bd, idesc, cb != 0 ? 1u : 0u);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
              for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k) {
                if (t == 0 && k == 0) continue;
                umma_tf32_acc(d_tmem, ad + uint64_t(tap_off[t] + 2 * k), bd + uint64_t(t * (S::W_TILE >> 4) + 2 * k), idesc);
              }
            }
          }
          umma_commit(&a_empty[s]);
          if (cb == p.cblocks - 1) umma_commit(&t_full[acc]);
          s ^= 1;
          a_ph ^= (s == 0);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    float* my_scratch = scratch + (warp - 2) * 32 * WS_SCR_LD;
    float* my_part = part + (warp - 2) * BN * 2;
    for (int c = lane; c < BN * 2; c += 32) my_part[c] = 0.f;
    __syncwarp();
    for (int j = 0; j < my_tiles; ++j) {
      const int tile = tile0 + j * tstride;
      const int img = tile / p.tiles_per_img;
      const int u = (tile - img * p.tiles_per_img) * IG_BLOCK_M + q * 32 + lane;
      const int h = u / p.Wp, w = u - h * p.Wp;
      const bool row_ok = (img < p.NB) && (h < p.H) && (w < p.W);
      // offset (floats) of this row's slice in the output; < 2^31 for every supported shape
      const int out_off = row_ok ? ((img * p.H + h) * p.W + w) * p.C_out + n0 : -1;
      const int acc = j & 1;
      mbar_wait(&t_full[acc], (j >> 1) & 1);
      tc_fence_after();
      // Each thread owns one accumulator row.  Writing it straight to global would make every warp-wide STG touch 32
      // different rows; instead the 32x32 chunk is staged in shared memory and written back 4 rows x 128 B at a time.
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4)
          *reinterpret_cast<float4*>(my_scratch + lane * WS_SCR_LD + jj) =
              row_ok ? make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]), __uint_as_float(v[jj + 2]),
                                   __uint_as_float(v[jj + 3]))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3), cq = lane & 7;
          const int off = __shfl_sync(0xffffffffu, out_off, r);
          float4 val = *reinterpret_cast<const float4*>(my_scratch + r * WS_SCR_LD + cq * 4);
          if (off >= 0 && !(p.dbg & 4)) {
            float4* dst = reinterpret_cast<float4*>(p.out + off + c0 + cq * 4);
            if (p.accumulate) {          // fused residual-gradient accumulation: out already holds the other branch
              const float4 old = *dst;
              val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
            }
            *dst = val;
          }
        }
        if (p.stats != nullptr) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float x = my_scratch[r * WS_SCR_LD + lane];
            s1 += x;
            s2 = fmaf(x, x, s2);
          }
          my_part[c0 + lane] += s1;
          my_part[BN + c0 + lane] += s2;
        }
        __syncwarp();
      }
      // all of this warp's TMEM reads of accumulator `acc` are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int t = threadIdx.x - 64;
      if (t < BN && my_tiles > 0) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int wv = 0; wv < 4; ++wv) {
          s1 += part[wv * BN * 2 + t];
          s2 += part[wv * BN * 2 + BN + t];
        }
        atomicAdd(p.stats + n0 + t, s1);
        atomicAdd(p.stats + p.C_out + n0 + t, s2);
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
