// Persistent variant of the implicit-GEMM kernel (igemm_tcgen05.cuh): one CTA per SM walks a static list of output
// tiles (tile = blockIdx.x + i * gridDim.x over [K split][M tile][N tile]), the TMEM accumulator is double-buffered
// and the epilogue of tile i runs under the MMAs of tile i + 1.
//
// Why: the one-tile-per-CTA kernel pays, per CTA, a prologue (barrier init, TMEM allocation, first TMA round trip:
// ~1.5-2 us), an exposed epilogue (1-2 us) and a wave transition (~1.2 us).  Measured with every pipeline switched off
// (tools/bench_ablate.py, profiles/r1_run12_ablation.log) that skeleton alone is 15 us of the 20 us a 1x1 shortcut
// convolution takes (256 CTAs, 2 k-blocks each), and 20-25 us of every multi-wave layer.  Here the prologue is paid
// once per SM and only the LAST tile's epilogue is exposed.
//
// Roles, pipelines and operand layouts are those of igemm_tf32_kernel (KPS k-blocks per smem stage, elect.sync-issued
// TMA / tcgen05.mma); the smem ring simply keeps rolling across tiles.
#pragma once
#include "igemm_tcgen05.cuh"

namespace fedb200 {

// MT = M sub-tiles (of 128 rows) per CTA tile.  The kernels are bound by the L2 -> SM fabric (~12 TB/s chip-wide,
// profiles/ROOFLINE.md), so what counts is bytes per FLOP: with MT = 2 one weight k-block feeds two activation tiles
// (256 x BLOCK_N outputs per CTA): 48 KB instead of 64 KB per k-block for 256 x 128 outputs.
template <int BLOCK_N, int STAGES, int KPS, int MT>
struct IgemmPSmem {
  static constexpr int A_TILE_BYTES = IG_BLOCK_M * IG_BLOCK_K * 4;   // 16 KB
  static constexpr int A_BYTES = MT * A_TILE_BYTES;
  static constexpr int B_BYTES = BLOCK_N * IG_BLOCK_K * 4;
  static constexpr int KB_BYTES = A_BYTES + B_BYTES;                  // one k-block: [A0 | A1 | B]
  static constexpr int STAGE_BYTES = KPS * KB_BYTES;
  static constexpr int ACC_STAGES = (2 * MT * BLOCK_N <= 512) ? 2 : 1;
  static constexpr int SCRATCH_BYTES = 4 * 32 * 33 * 4;
  static constexpr int PART_BYTES = 2 * 4 * BLOCK_N * 2 * 4;    // double-buffered by accumulator stage
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + SCRATCH_BYTES + PART_BYTES + BAR_BYTES + 1024;
};

struct TileCoord {
  int m0, n0, kb_begin, kb_count;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int t, int block_n, int block_m) {
  const int n_idx = t % p.n_tiles;
  const int rest = t / p.n_tiles;
  const int m_idx = rest % p.m_tiles;
  const int z = rest / p.m_tiles;
  TileCoord c;
  c.m0 = m_idx * block_m;
  c.n0 = n_idx * block_n;
  c.kb_begin = z * p.kb_per_split;
  c.kb_count = min(p.kb_per_split, p.num_k_blocks - c.kb_begin);
  return c;
}

template <int BLOCK_N, int STAGES, int KPS, int MT>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ CUtensorMap tmap_c, const IgemmParams p) {
  using S = IgemmPSmem<BLOCK_N, STAGES, KPS, MT>;
  constexpr int ACC = S::ACC_STAGES;
  constexpr int TILE_M = MT * IG_BLOCK_M;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N must be a multiple of 32 in [32,256]");
  static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
  constexpr uint32_t TMEM_COLS = ACC * MT * BLOCK_N < 32 ? 32 : ACC * MT * BLOCK_N;   // power of two: 64..512

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* scratch = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
  float* part = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::SCRATCH_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::SCRATCH_BYTES + S::PART_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* t_full = empty_bar + STAGES;     // 2
  uint64_t* t_empty = t_full + 2;            // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool tr = p.trace != nullptr && blockIdx.x == 0;
#define FEDB200_STAMP(i) do { if (tr) p.trace[(i)] = clock64(); } while (0)
  if (threadIdx.x == 0) FEDB200_STAMP(0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_store) tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 4);             // one arrival per epilogue warp
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
  if (threadIdx.x == 0) FEDB200_STAMP(1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t, BLOCK_N, TILE_M);
        int img[MT], h0[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          img[mt] = 0; h0[mt] = 0;
          if (p.is_conv) {
            const int m = c.m0 + mt * IG_BLOCK_M;      // past the end: image index >= NB, the TMA unit zero-fills
            img[mt] = m / p.HW_out;
            h0[mt] = ((m - img[mt] * p.HW_out) / p.W_out) * p.stride - p.pad;
          }
        }
        int tap = c.kb_begin / p.cblocks;
        int cb = c.kb_begin - tap * p.cblocks;
        int r = tap / p.taps_w, sx = tap - r * p.taps_w;
        int left = c.kb_count, kb = c.kb_begin;
        const int cwp = (p.cw == 8 || p.cw == 16) ? p.cw : 0;      // tap packing
        const int tpk = cwp ? IG_BLOCK_K / cwp : 1;
        const int img_oob = p.M / p.HW_out + 1;                       // first image index past the tensor: the TMA unit zero-fills
        const int col_oob = p.taps_total * p.b_cols_per_tap + IG_BLOCK_K;   // first weight column past the filter (+ a box)
        // tap -> offset inside the (padded) input window
        auto tap_offset = [&](int rr_, int sx_, int& cw_, int& ch_) {
          cw_ = sx_ * p.dil - p.pad;
          ch_ = rr_ * p.dil;
          if (p.ms_kh > 0) {                                        // multi-dilation: rows are (branch, row) pairs
            const int br = rr_ / p.ms_kh, rr = rr_ - br * p.ms_kh;
            const int d = int((p.ms_dil >> (8 * br)) & 0xffull), pd = int((p.ms_pad >> (8 * br)) & 0xffull);
            cw_ = sx_ * d - pd;
            ch_ = rr * d - pd;
          }
        };
        while (left > 0) {
          const int nk = left < KPS ? left : KPS;
          left -= nk;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* dst = tiles + s * S::STAGE_BYTES;
          if (p.dbg & 1) {
            mbar_arrive(&full_bar[s]);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], uint32_t(nk * S::KB_BYTES));
#pragma unroll
            for (int j = 0; j < KPS; ++j) {
              if (j < nk) {
                uint8_t* a_dst = dst + j * S::KB_BYTES;
                uint8_t* b_dst = a_dst + S::A_BYTES;
                if (cwp) {
                  // tpk taps per k-block, each a [128 rows x cwp channels] sub-tile of A and a [BLOCK_N x cwp] sub-tile of B
                  const int sub_a = IG_BLOCK_M * cwp * 4, sub_b = BLOCK_N * cwp * 4;
                  for (int jt = 0; jt < tpk; ++jt) {
                    const int tp = (kb + j) * tpk + jt;
                    const bool real = tp < p.taps_total;
                    int cw = 0, ch = 0;
                    if (real) tap_offset(tp / p.taps_w, tp % p.taps_w, cw, ch);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)          // a tap past the filter: image index out of range -> zeros
                      tma_load_4d(a_dst + mt * S::A_TILE_BYTES + jt * sub_a, &tmap_a, &full_bar[s], 0, cw, h0[mt] + ch,
                                  real ? img[mt] : img_oob);
                    tma_load_2d(b_dst + jt * sub_b, &tmap_b, &full_bar[s], real ? tp * p.b_cols_per_tap : col_oob, c.n0);
                  }
                } else {
                  int cw, ch;
                  tap_offset(r, sx, cw, ch);
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    if (p.is_conv)
                      tma_load_4d(a_dst + mt * S::A_TILE_BYTES, &tmap_a, &full_bar[s], cb * IG_BLOCK_K, cw, h0[mt] + ch, img[mt]);
                    else
                      tma_load_2d(a_dst + mt * S::A_TILE_BYTES, &tmap_a, &full_bar[s], (kb + j) * IG_BLOCK_K,
                                  c.m0 + mt * IG_BLOCK_M);
                  }
                  tma_load_2d(b_dst, &tmap_b, &full_bar[s], (r * p.taps_w + sx) * p.b_cols_per_tap + cb * IG_BLOCK_K, c.n0);
                  if (++cb == p.cblocks) {
                    cb = 0;
                    if (++sx == p.taps_w) { sx = 0; ++r; }
                  }
                }
              }
            }
          }
          kb += nk;
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (t == int(blockIdx.x)) FEDB200_STAMP(2); else if (t == int(blockIdx.x + gridDim.x)) FEDB200_STAMP(3);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(/*tf32*/ 2, IG_BLOCK_M, BLOCK_N);
    if (elect_one()) {
      // operand rows of 128 bytes (one tap x 32 channels per k-block) or, with tap packing, cw * 4 bytes per sub-tile
      const uint32_t cwq = (p.cw == 8 || p.cw == 16) ? uint32_t(p.cw) : uint32_t(IG_BLOCK_K);
      const uint64_t desc0 = make_kmajor_desc(smem_u32(tiles), cwq * 4u);
      const uint32_t mma_per_sub = cwq / IG_UMMA_K;                          // MMAs (K = 8) per sub-tile: 1, 2 or 4
      const uint32_t sub_a16 = (IG_BLOCK_M * cwq * 4u) >> 4, sub_b16 = (uint32_t(BLOCK_N) * cwq * 4u) >> 4;
      uint64_t a_off[IG_BLOCK_K / IG_UMMA_K], b_off[IG_BLOCK_K / IG_UMMA_K];   // descriptor offsets of the four MMAs of a k-block
#pragma unroll
      for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k) {
        const uint32_t sub = uint32_t(k) / mma_per_sub, in = uint32_t(k) - sub * mma_per_sub;
        a_off[k] = uint64_t(sub * sub_a16 + 2u * in);
        b_off[k] = uint64_t(sub * sub_b16 + 2u * in);
      }
      int s = 0;
      uint32_t ph = 0;
      int j_tile = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j_tile) {
        const TileCoord c = decode_tile(p, t, BLOCK_N, TILE_M);
        const uint32_t acc = uint32_t(j_tile % ACC);
        const uint32_t d_tmem = tmem_base + acc * (MT * BLOCK_N);
        mbar_wait(&t_empty[acc], ((j_tile / ACC) & 1) ^ 1);     // the epilogue has drained this accumulator
        tc_fence_after();
        int left = c.kb_count;
        bool first = true;
        while (left > 0) {
          const int nk = left < KPS ? left : KPS;
          left -= nk;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (j_tile == 0 && first) FEDB200_STAMP(4);
          const uint64_t sd = desc0 + uint64_t(uint32_t(s) * uint32_t(S::STAGE_BYTES >> 4));
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int j = 0; j < KPS; ++j) {
              if (j < nk) {
                const uint64_t adesc = sd + uint64_t(j * (S::KB_BYTES >> 4));
                const uint64_t bdesc = adesc + uint64_t(S::A_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k) {
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    const uint64_t ad = adesc + uint64_t(mt * (S::A_TILE_BYTES >> 4)) + a_off[k];
                    if (j == 0 && k == 0) umma_tf32(d_tmem + mt * BLOCK_N, ad, bdesc + b_off[k], idesc, first ? 0u : 1u);
                    else umma_tf32_acc(d_tmem + mt * BLOCK_N, ad, bdesc + b_off[k], idesc);
                  }
                }
              }
            }
          }
          first = false;
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&t_full[acc]);
        if (j_tile == 0) FEDB200_STAMP(5); else if (j_tile == 1) FEDB200_STAMP(6);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers -> 128B-swizzled staging tile in shared memory -> ONE bulk tensor store (or reduce-add for
    // split-K) per 32x32 chunk and warp.  Direct register stores put 32 different rows into every STG: 4-6k cycles per
    // 128x128 tile against ~1.2k for the TMEM reads (tools/trace_conv.py); the bulk store is asynchronous, writes
    // whole 128-B lines and clips rows/columns outside the tensor by itself.
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    float* stage = scratch + (warp - 2) * 1024;   // [32 rows][32 floats], 1024-B aligned, chunk index XOR (row & 7)
    const uint32_t sw = uint32_t(lane & 7);
    int j_tile = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j_tile) {
      const TileCoord c = decode_tile(p, t, BLOCK_N, TILE_M);
      const int acc = j_tile % ACC;
      float* part_acc = part + (j_tile & 1) * (4 * BLOCK_N * 2);    // double-buffered column partials
      float* my_part = part_acc + (warp - 2) * BLOCK_N * 2;
      mbar_wait(&t_full[acc], (j_tile / ACC) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) { if (j_tile == 0) FEDB200_STAMP(7); else if (j_tile == 1) FEDB200_STAMP(9); }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int row0 = c.m0 + mt * IG_BLOCK_M + q * 32;
        const int row = row0 + lane;
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * (MT * BLOCK_N) + mt * BLOCK_N + c0), v);
          tmem_ld_wait();
          float f[32];
          if (p.bias == nullptr && !p.act) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(v[j]);
              const int col = c.n0 + c0 + j;
              if (p.bias != nullptr && col < p.N) x += __ldg(p.bias + col);
              if (p.act) x = elu1(x);
              f[j] = x;
            }
          }
          const bool staged = p.tma_store || p.stats != nullptr;
          if (staged) {
            if (p.tma_store) {
              if (lane == 0) tma_store_wait_read();      // the previous chunk's store has drained the staging tile
              __syncwarp();
            }
            if (!row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = 0.f;   // clipped by the store; must not reach the statistics
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(stage + lane * 32 + ((uint32_t(j) ^ sw) << 2)) =
                  make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            if (p.tma_store) fence_proxy_async();
            __syncwarp();
            if (p.tma_store && lane == 0 && !(p.dbg & 4)) {
              if (p.shuffle_ci > 0) {
                const int pc = c.n0 + c0;                       // packed channel (ph, pw, ci) of this chunk
                const int phase = pc / p.shuffle_ci;
                if (pc < p.N) {
                  if (p.k_splits > 1) tma_reduce_add_5d(&tmap_c, stage, pc - phase * p.shuffle_ci, phase & 1, 0, phase >> 1, row0 / p.W_out);
                  else tma_store_5d(&tmap_c, stage, pc - phase * p.shuffle_ci, phase & 1, 0, phase >> 1, row0 / p.W_out);
                }
              } else if (p.k_splits > 1 || p.accumulate) {
                tma_reduce_add_2d(&tmap_c, stage, c.n0 + c0, row0);
              } else {
                tma_store_2d(&tmap_c, stage, c.n0 + c0, row0);
              }
              tma_store_commit();
            }
          }
          if (!p.tma_store && row_ok && !(p.dbg & 4)) {
            // fallback for outputs whose row pitch is not a multiple of 16 B (e.g. 10-class logits)
            float* dst = p.out + size_t(row) * p.ldo + c.n0 + c0;
            if (p.k_splits > 1 || p.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c.n0 + c0 + j < p.N) atomicAdd(dst + j, f[j]);
            } else if (c.n0 + c0 + 32 <= p.N && (p.ldo & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c.n0 + c0 + j < p.N) dst[j] = f[j];
            }
          }
          if (p.stats != nullptr) {
            // column sums over this warp's 32 rows, read straight from the staging tile (conflict-free: for a fixed
            // row the 32 lanes read the 32 floats of that row in a permuted order)
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float x = stage[r * 32 + (((uint32_t(lane) >> 2) ^ uint32_t(r & 7)) << 2) + (lane & 3)];
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
            if (mt == 0) {
              my_part[c0 + lane] = s1;
              my_part[BLOCK_N + c0 + lane] = s2;
            } else {
              my_part[c0 + lane] += s1;
              my_part[BLOCK_N + c0 + lane] += s2;
            }
            __syncwarp();
          }
        }
      }
      // every TMEM read of this warp for accumulator `acc` is complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      if (warp == 2 && lane == 0) { if (j_tile == 0) FEDB200_STAMP(8); else if (j_tile == 1) FEDB200_STAMP(10); }
      if (p.stats != nullptr) {
        // `part` is double-buffered by tile parity: the next tile writes the other half, and this half is only rewritten
        // two tiles later, after every thread has passed the next tile's bar.sync (i.e. finished reading it here)
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
        const int tt = threadIdx.x - 64;
        for (int cc = tt; cc < BLOCK_N; cc += 128) {
          if (c.n0 + cc < p.N) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              s1 += part_acc[w * BLOCK_N * 2 + cc];
              s2 += part_acc[w * BLOCK_N * 2 + BLOCK_N + cc];
            }
            atomicAdd(p.stats + c.n0 + cc, s1);
            atomicAdd(p.stats + p.N + c.n0 + cc, s2);
          }
        }
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait_read();   // shared memory must outlive the last bulk store's read
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) FEDB200_STAMP(11);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
#undef FEDB200_STAMP
}

}  // namespace fedb200
