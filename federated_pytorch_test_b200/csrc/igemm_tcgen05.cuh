// Implicit-GEMM on 5th-gen tensor cores: ONE warp-specialised kernel for
//   (a) plain GEMM          C[M,N]   = A[M,K] * B[N,K]^T                 (Linear layers, 1x1 convs on latent grids)
//   (b) NHWC convolution    Y[m,co]  = sum_{r,s,ci} X[n,h*st+r*d-p,w*st+s*d-p,ci] * W[co,r,s,ci]
// A and B tiles are brought in by TMA as 128-byte-swizzled K-major tiles (32 fp32 per row); for (b) the A tile of
// filter tap (r,s) is a 4-D TMA box over the NHWC activation [C,W,H,N] shifted by the tap offset, with out-of-bounds
// rows/columns zero-filled by the TMA unit — the im2col matrix is never materialised.  One elected thread issues
// tcgen05.mma.kind::tf32 into a TMEM accumulator (128 lanes x BLOCK_N columns); four epilogue warps read it back
// with tcgen05.ld and apply the fused epilogue:
//   * bias + ELU (dense layers), or
//   * per-channel sum / sum-of-squares of the tile (BatchNorm batch statistics, reduced across the CTA in shared
//     memory, one atomicAdd per channel per CTA) — SURVEY G1/G2.
//
// The kernel is L2->SM bandwidth bound (measured: ~6.5 TB/s, profiles/r1_run2_*), so operand re-reads are what
// matters.  CL > 1 launches thread-block clusters of CL CTAs that own CL consecutive M tiles of the same N tile:
// each CTA fetches 1/CL of the weight tile and TMA-multicasts it into the shared memory of all CL CTAs
// (weight traffic / CL); the smem-slot release is a tcgen05.commit multicast to every CTA of the cluster.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), one tmem_full mbarrier (MMA -> epilogue).
#pragma once
#include "sm100.cuh"
#include "fedb200.h"

namespace fedb200 {

constexpr int IG_BLOCK_M = 128;
constexpr int IG_BLOCK_K = 32;  // fp32 elements per k-block = 128 B = one swizzle row
constexpr int IG_UMMA_K = 8;    // tf32: 32 B per MMA k-step
constexpr int IG_THREADS = 192;

struct IgemmParams {
  int M, N;              // output rows (pixels) and columns (channels)
  int num_k_blocks;      // taps * cblocks
  int cblocks;           // ceil(Cin / 32); k-block kb -> tap = kb / cblocks, cb = kb % cblocks
  int taps_w;            // filter width (tap -> r = tap / taps_w, s = tap % taps_w); 1 for GEMM
  int b_cols_per_tap;    // columns of the weight matrix per tap (= Cin as stored)
  int is_conv;           // 0: A is a 2-D map [K, M]; 1: A is a 4-D map [C, W, H, N]
  int HW_out, W_out;     // conv: output pixels per image, output width
  int stride, pad, dil;  // conv geometry
  float* out;            // [M, ldo]
  int ldo;
  const float* bias;     // [N] or nullptr
  int act;               // 1 = ELU
  float* stats;          // [2*N]: sum, sumsq per column (atomicAdd) or nullptr
  int kb_per_split;      // split-K: CTA z handles k-blocks [z*kb_per_split, ...) and red.adds into a zeroed output
  int k_splits;          // 1 = plain stores (+ fused stats); > 1 = reduction through L2 atomics, stats done by caller
  int dbg;               // profiling ablations (FEDB200_DBG, results are garbage): 1 = no TMA loads, 2 = no MMAs, 4 = no stores
  int m_tiles, n_tiles, total_tiles;   // persistent kernel: tile t -> (t % n_tiles, (t / n_tiles) % m_tiles, K split)
  int accumulate;        // persistent kernel: out += result (bulk reduce-add / atomics) instead of out = result; no memset
  int tma_store;         // persistent kernel: write the output with bulk tensor stores / reduce-adds (needs ldo % 4 == 0)
  long long* trace;      // optional [16] clock64 stamps of CTA 0 (tools/trace_conv.py); nullptr in production
  int cw;                // persistent kernel, convolutions with few input channels: channels per filter tap inside a 32-wide
                         // k-block.  0 / 32 = one tap per k-block (C_in padded to 32 by the TMA zero fill: 4x wasted MMAs at C_in = 8);
                         // 8 / 16 = "tap packing": a k-block holds 32 / cw taps, each its own cw-channel sub-tile (rows of cw * 4 bytes,
                         // TMA SWIZZLE_32B / 64B, UMMA descriptors of the matching layout).  num_k_blocks = ceil(taps / (32 / cw)).
  int taps_total;        // kh * kw (tap packing: taps beyond it load out-of-bounds zeros)
  int ms_kh;             // persistent kernel, > 0: "multi-dilation" convolution — the filter rows are ms_branches groups of ms_kh rows,
                         // group b reads the input with dilation ms_dil[b] and padding ms_pad[b] (one byte each, packed); the weight
                         // matrix is block diagonal (branch b owns its own output channels).  The five dilated stem convolutions
                         // of the CPC encoder run as ONE launch writing the concatenated tensor.  pad = 0, dil = 1 then.
  unsigned long long ms_dil, ms_pad;
  int shuffle_ci;        // persistent kernel, > 0: the N columns are (ph, pw, ci) phase-packed channels of a stride-2 data gradient /
                         // transposed conv; each 32 x 32 chunk is stored to out[n, 2 ho + ph, 2 wo + pw, ci0 .. ci0 + 31] through a
                         // 5-D tensor map (no separate pixel-shuffle pass).  shuffle_ci = channels of the shuffled output.
};

// KPS = k-blocks (of 32 fp32 = one 128-B swizzle row) per pipeline stage.  One producer/consumer barrier round trip
// costs the MMA-issuing thread ~300 cycles that do NOT overlap with MMA execution unless >= 8 MMAs are queued behind
// it (tools/probe_ring.cu: 4 MMAs of 128x128x8 per handshake run at 38% of the tensor-pipe rate, 8 at 62%), so a
// stage carries KPS * 4 MMAs.
template <int BLOCK_N, int STAGES, int KPS>
struct IgemmSmem {
  static constexpr int A_BYTES = IG_BLOCK_M * IG_BLOCK_K * 4;   // 16 KB
  static constexpr int B_BYTES = BLOCK_N * IG_BLOCK_K * 4;
  static constexpr int KB_BYTES = A_BYTES + B_BYTES;            // one k-block: [A | B]
  static constexpr int STAGE_BYTES = KPS * KB_BYTES;
  static constexpr int SCRATCH_BYTES = 4 * 32 * 33 * 4;         // per-epilogue-warp transpose tiles
  static constexpr int PART_BYTES = 4 * BLOCK_N * 2 * 4;        // per-warp column partials (sum, sumsq)
  static constexpr int BAR_BYTES = (2 * STAGES + 1) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + SCRATCH_BYTES + PART_BYTES + BAR_BYTES + 1024;  // + align slack
};

template <int BLOCK_N, int STAGES, int CL, int KPS>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const IgemmParams p) {
  using S = IgemmSmem<BLOCK_N, STAGES, KPS>;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N must be a multiple of 32 in [32,256]");
  static_assert(CL == 1 || CL == 2 || CL == 4, "cluster size 1, 2 or 4");
  static_assert((BLOCK_N / CL) % 8 == 0, "each CTA's slice of the weight tile must be whole swizzle atoms");
  static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;   // power of two >= 32 for 32/64/128/256
  constexpr int B_SLICE_ROWS = BLOCK_N / CL;
  constexpr uint16_t CL_MASK = uint16_t((1u << CL) - 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;                                           // STAGES x KPS x [A | B], each 1024-B aligned
  float* scratch = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
  float* part = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::SCRATCH_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::SCRATCH_BYTES + S::PART_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * IG_BLOCK_M;
  const int n0 = blockIdx.y * BLOCK_N;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  const int kb_begin = blockIdx.z * p.kb_per_split;                 // split-K slice of this CTA
  const int kb_count = min(p.kb_per_split, p.num_k_blocks - kb_begin);
  const int n_iters = (kb_count + KPS - 1) / KPS;                   // pipeline stages this CTA consumes

  pdl_launch_dependents();              // the next kernel of the stream may start its own prologue now
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);        // every CTA of the cluster releases the slot (it holds multicast data)
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
  pdl_wait();                           // barriers/TMEM are set up; from here on global memory of earlier kernels is read

  if (warp == 0) {
    // ===================== TMA producer (one thread; no divisions inside the loop) =====================
    if (elect_one()) {
      int img = 0, h0 = 0;
      if (p.is_conv) {
        img = m0 / p.HW_out;
        h0 = ((m0 - img * p.HW_out) / p.W_out) * p.stride - p.pad;
      }
      int tap = kb_begin / p.cblocks;
      int cb = kb_begin - tap * p.cblocks;
      int r = tap / p.taps_w, sx = tap - r * p.taps_w;
      int s = 0;
      uint32_t ph = 0;
      int left = kb_count;
      for (int it = 0; it < n_iters; ++it) {
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
              if (p.is_conv)
                tma_load_4d(a_dst, &tmap_a, &full_bar[s], cb * IG_BLOCK_K, sx * p.dil - p.pad, h0 + r * p.dil, img);
              else
                tma_load_2d(a_dst, &tmap_a, &full_bar[s], (kb_begin + it * KPS + j) * IG_BLOCK_K, m0);
              const int b_col = (r * p.taps_w + sx) * p.b_cols_per_tap + cb * IG_BLOCK_K;
              if (CL == 1) {
                tma_load_2d(b_dst, &tmap_b, &full_bar[s], b_col, n0);
              } else {
                // my 1/CL slice of the weight tile, written into the same slot of every CTA in the cluster
                tma_load_2d_multicast(b_dst + cta_rank * (B_SLICE_ROWS * IG_BLOCK_K * 4), &tmap_b, &full_bar[s], b_col,
                                      n0 + int(cta_rank) * B_SLICE_ROWS, CL_MASK);
              }
              if (++cb == p.cblocks) {
                cb = 0;
                if (++sx == p.taps_w) { sx = 0; ++r; }
              }
            }
          }
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    constexpr uint32_t idesc = make_idesc(/*tf32*/ 2, IG_BLOCK_M, BLOCK_N);
    if (elect_one()) {
      const uint64_t desc0 = make_kmajor_sw128_desc(smem_u32(tiles));
      int s = 0;
      uint32_t ph = 0;
      int left = kb_count;
      for (int it = 0; it < n_iters; ++it) {
        const int nk = left < KPS ? left : KPS;
        left -= nk;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t sd = desc0 + uint64_t(uint32_t(s) * uint32_t(S::STAGE_BYTES >> 4));
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            if (j < nk) {
              const uint64_t adesc = sd + uint64_t(j * (S::KB_BYTES >> 4));
              const uint64_t bdesc = adesc + uint64_t(S::A_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k) {
                // advance along K inside the 128-B swizzle atom: +32 B per step (encoded >>4 => +2)
                if (j == 0 && k == 0) umma_tf32(tmem_base, adesc, bdesc, idesc, it != 0 ? 1u : 0u);
                else umma_tf32_acc(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc);
              }
            }
          }
        }
        // free the smem slot (in every CTA of the cluster) when these MMAs retire
        if (CL == 1) umma_commit(&empty_bar[s]); else umma_commit_multicast(&empty_bar[s], CL_MASK);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
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
      // Convolutions have neither bias nor activation here: the accumulator goes out as it is.  (With the per-element
      // bias / bounds / ELU predicates compiled in unconditionally this loop cost ~2800 cycles per 32-column chunk, the
      // whole epilogue 7 us per 128x128 tile: tools/trace_conv.py, profiles/r1_run17_trace.log.)
      float f[32];
      if (p.bias == nullptr && !p.act) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]);
          const int col = n0 + c0 + j;
          if (p.bias != nullptr && col < p.N) x += __ldg(p.bias + col);
          if (p.act) x = elu1(x);
          f[j] = x;
        }
      }
      if (row_ok && !(p.dbg & 4)) {
        float* dst = p.out + size_t(row) * p.ldo + n0 + c0;
        if (p.k_splits > 1) {
          // partial sums of this K slice: vector reduction into the (pre-zeroed) output, resolved in L2
          if (n0 + c0 + 32 <= p.N && (p.ldo & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) red_add_v4(dst + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < p.N) atomicAdd(dst + j, f[j]);
          }
        } else if (n0 + c0 + 32 <= p.N && (p.ldo & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < p.N) dst[j] = f[j];
        }
      }
      if (p.stats != nullptr) {
        // column sums over this warp's 32 rows via a padded shared-memory transpose
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
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
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
  // no CTA may leave while a peer can still multicast into its shared memory / arrive on its barriers
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace fedb200
