// 3x3 / stride 1 / pad 1 NHWC convolution with ONE shared-memory halo tile per channel block.
//
// The generic implicit-GEMM kernel (igemm_tcgen05.cuh) fetches the activation tile once per filter tap: 9 TMA boxes
// of 16 KB per 32-channel block, all from L2 — and that L2->SM traffic is what bounds it (profiles/r1_run2_*).
// Here the tile of 128 outputs is taken in the *zero-padded, flattened* pixel space of one image,
//      u = h * Wp + w,   Wp = W + 2,   (outputs with w >= W are padding columns and are discarded)
// in which every filter tap is a constant row offset:  input(u; r,s) = P[u + r*Wp + s].  ONE TMA box
// [32 ch] x [Wp columns from w=-1] x [R rows] per channel block therefore serves all nine taps: the MMA for tap
// (r,s) simply starts its A descriptor (r*Wp + s) rows further down the same swizzled tile (start address at
// 128-B granularity, descriptor base_offset = (addr >> 7) & 7).  Activation traffic drops ~5x; the weight tile is
// cluster-multicast as in the generic kernel.  Cost: Wp/W - 1 = 6 % (W=32) of the MMA rows are padding.
//
// Stage = one 32-channel block: halo tile (R*Wp rows x 128 B) + nine weight tiles (BLOCK_N x 128 B each).
#pragma once
#include "igemm_tcgen05.cuh"

namespace fedb200 {

struct HaloParams {
  int NB, H, W, Wp;        // images, output height/width (= input), padded width
  int R;                   // rows of the halo box
  int tiles_per_img;       // ceil(H * Wp / 128)
  int C_in, C_out, cblocks;
  int a_box_bytes;         // 128 * Wp * R
  int box_h;               // image rows per TMA box (the halo tile is fetched as ceil(R / box_h) boxes issued back to back)
  int use_base_offset;     // experiment switch: encode (addr>>7)&7 in the descriptor's base_offset field
  int accumulate;          // weight-stationary kernel: out += result (read-modify-write epilogue)
  int dbg;                 // profiling ablations (FEDB200_DBG): 1 = no activation TMA loads, 2 = no MMAs, 4 = no stores
  float* out;              // [NB*H*W, C_out]
  float* stats;            // [2*C_out] or nullptr
};

template <int BLOCK_N, int A_SLOT_BYTES>
struct HaloSmem {
  static constexpr int STAGES = 2;
  static constexpr int B_TILE = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_SLOT_BYTES + 9 * B_TILE;
  static constexpr int SCRATCH_BYTES = 4 * 32 * 33 * 4;
  static constexpr int PART_BYTES = 4 * BLOCK_N * 2 * 4;
  static constexpr int BAR_BYTES = (2 * STAGES + 1) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + SCRATCH_BYTES + PART_BYTES + BAR_BYTES + 1024;
};

// descriptor for a K-major SW128 tile whose first row is NOT at a 1024-B boundary
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_rowoff(uint32_t smem_addr) {
  uint64_t d = make_kmajor_sw128_desc(smem_addr);
  d |= static_cast<uint64_t>((smem_addr >> 7) & 7) << 49;   // base_offset
  return d;
}

template <int BLOCK_N, int A_SLOT_BYTES, int CL>
__global__ void __launch_bounds__(IG_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const HaloParams p) {
  using S = HaloSmem<BLOCK_N, A_SLOT_BYTES>;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr int B_SLICE_ROWS = BLOCK_N / CL;
  constexpr uint16_t CL_MASK = uint16_t((1u << CL) - 1);
  static_assert(A_SLOT_BYTES % 1024 == 0, "A slot must keep the weight tiles 1024-B aligned");
  static_assert((BLOCK_N / CL) % 8 == 0, "weight slice must be whole swizzle atoms");

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
  const int tile = blockIdx.x;
  const int img = tile / p.tiles_per_img;                  // may be >= NB for cluster padding tiles: TMA zero-fills
  const int u0 = (tile - img * p.tiles_per_img) * IG_BLOCK_M;
  const int n0 = blockIdx.y * BLOCK_N;
  const int hp_a = u0 / p.Wp;                              // first padded row held by the halo tile
  const int row_off0 = u0 - hp_a * p.Wp;                   // row of output u0's tap (0,0) inside the tile
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      for (int cb = 0; cb < p.cblocks; ++cb) {
        const int s = cb % STAGES;
        const uint32_t ph = (cb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = tiles + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_SLOT_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], uint32_t(p.a_box_bytes + 9 * S::B_TILE));
        // padded (wp=0, hp=hp_a) is real (w=-1, h=hp_a-1): the TMA unit zero-fills everything outside the image
        tma_load_4d(a_dst, &tmap_a, &full_bar[s], cb * IG_BLOCK_K, -1, hp_a - 1, img);
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const int b_col = t * p.C_in + cb * IG_BLOCK_K;
          uint8_t* dst = b_dst + t * S::B_TILE;
          if (CL == 1) tma_load_2d(dst, &tmap_b, &full_bar[s], b_col, n0);
          else tma_load_2d_multicast(dst + cta_rank * (B_SLICE_ROWS * 128), &tmap_b, &full_bar[s], b_col,
                                     n0 + int(cta_rank) * B_SLICE_ROWS, CL_MASK);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(/*tf32*/ 2, IG_BLOCK_M, BLOCK_N);
    for (int cb = 0; cb < p.cblocks; ++cb) {
      const int s = cb % STAGES;
      const uint32_t ph = (cb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_base = smem_u32(tiles + s * S::STAGE_BYTES);
        const uint32_t b_base = a_base + A_SLOT_BYTES;
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const int r = t / 3, sx = t - 3 * r;
          const uint32_t a_addr = a_base + uint32_t(row_off0 + r * p.Wp + sx) * 128u;
          const uint64_t adesc = p.use_base_offset ? make_kmajor_sw128_desc_rowoff(a_addr) : make_kmajor_sw128_desc(a_addr);
          const uint64_t bdesc = make_kmajor_sw128_desc(b_base + uint32_t(t) * S::B_TILE);
#pragma unroll
          for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k)
            umma_tf32(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (cb | t | k) != 0 ? 1u : 0u);
        }
        if (CL == 1) umma_commit(&empty_bar[s]); else umma_commit_multicast(&empty_bar[s], CL_MASK);
        if (cb == p.cblocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int u = u0 + q * 32 + lane;                      // this thread's output in padded-flattened space
    const int h = u / p.Wp, w = u - h * p.Wp;
    const bool row_ok = (img < p.NB) && (h < p.H) && (w < p.W);
    const size_t out_row = (size_t(img) * p.H + h) * p.W + w;
    float* my_scratch = scratch + (warp - 2) * 32 * 33;
    float* my_part = part + (warp - 2) * BLOCK_N * 2;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
      tmem_ld_wait();
      if (row_ok) {
        float* dst = p.out + out_row * p.C_out + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      if (p.stats != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) my_scratch[lane * 33 + j] = row_ok ? __uint_as_float(v[j]) : 0.f;
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
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int wv = 0; wv < 4; ++wv) {
          s1 += part[wv * BLOCK_N * 2 + c];
          s2 += part[wv * BLOCK_N * 2 + BLOCK_N + c];
        }
        atomicAdd(p.stats + n0 + c, s1);
        atomicAdd(p.stats + p.C_out + n0 + c, s2);
      }
    }
    tc_fence_before();
  }
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace fedb200
