// Host side of the tcgen05 implicit-GEMM kernel: TMA tensor-map construction, tile/cluster selection, launch.
// (Kernel: igemm_tcgen05.cuh.)  Torch-free translation unit: raw pointers + cudaStream_t.
#include "fedb200.h"
#include "igemm_tcgen05.cuh"

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace fedb200 {

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; resolve it at run time (no link-time libcuda:
// the build container has no driver).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr)
      throw std::runtime_error("fedb200: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static void check_cu(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) throw std::runtime_error(std::string("fedb200: ") + what + " failed with CUresult " + std::to_string(int(r)));
}

// fp32 matrix [rows, cols] with row pitch ld (elements): box = [box_rows x 32 cols], 128B swizzle,
// loaded as TF32 (round-to-nearest on the way into shared memory), out-of-bounds zero-filled.
static CUtensorMap make_tmap_2d(const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {uint32_t(IG_BLOCK_K), box_rows};
  cuuint32_t estr[2] = {1, 1};
  check_cu(encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
           "cuTensorMapEncodeTiled(2d)");
  return m;
}

// NHWC activation viewed as [C, W, H, N]; a box covers boxN x boxH x boxW output pixels (traversal stride = conv
// stride) x 32 channels and lands in shared memory as a [128 pixels x 128 B] K-major swizzled tile.
static CUtensorMap make_tmap_nhwc(const float* ptr, uint64_t N, uint64_t H, uint64_t W, uint64_t C, uint32_t boxN,
                                  uint32_t boxH, uint32_t boxW, uint32_t stride) {
  CUtensorMap m;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * sizeof(float), W * C * sizeof(float), H * W * C * sizeof(float)};
  cuuint32_t box[4] = {uint32_t(IG_BLOCK_K), boxW * stride, boxH * stride, boxN};
  cuuint32_t estr[4] = {1, stride, stride, 1};
  check_cu(encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
           "cuTensorMapEncodeTiled(nhwc)");
  return m;
}

template <int BN, int ST, int CL>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const IgemmParams& p, cudaStream_t stream) {
  using S = IgemmSmem<BN, ST>;
  auto kernel = igemm_tf32_kernel<BN, ST, CL>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) throw std::runtime_error(std::string("fedb200: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured = true;
  }
  const int mt = (p.M + IG_BLOCK_M - 1) / IG_BLOCK_M;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(((mt + CL - 1) / CL) * CL, (p.N + BN - 1) / BN);   // padded M tiles only feed the multicast
  cfg.blockDim = dim3(IG_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ta, tb, p);
  if (e != cudaSuccess) throw std::runtime_error(std::string("fedb200: igemm launch: ") + cudaGetErrorString(e));
  count_launch();
}

template <int BN, int ST>
static void dispatch_cl(int cl, const CUtensorMap& ta, const CUtensorMap& tb, const IgemmParams& p, cudaStream_t s) {
  if (cl >= 4) launch<BN, ST, 4>(ta, tb, p, s);
  else if (cl == 2) launch<BN, ST, 2>(ta, tb, p, s);
  else launch<BN, ST, 1>(ta, tb, p, s);
}

static void dispatch(int bn, int cl, const CUtensorMap& ta, const CUtensorMap& tb, const IgemmParams& p, cudaStream_t s) {
  switch (bn) {
    case 32: dispatch_cl<32, 8>(cl, ta, tb, p, s); break;
    case 64: dispatch_cl<64, 6>(cl, ta, tb, p, s); break;
    case 128: dispatch_cl<128, 5>(cl, ta, tb, p, s); break;
    default: dispatch_cl<256, 4>(cl, ta, tb, p, s); break;
  }
}

static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// The kernel is L2->SM bandwidth bound: bytes moved = A_bytes * taps * (N / BLOCK_N) + W_bytes * (M tiles / CL).
// So: the widest N tile that N allows (fewer passes over the activations) and the largest cluster (fewer passes
// over the weights).  FEDB200_BLOCK_N / FEDB200_CLUSTER override for experiments.
int pick_block_n(int M, int N) {
  const int forced = env_int("FEDB200_BLOCK_N", 0);
  if (forced == 32 || forced == 64 || forced == 128 || forced == 256) return forced;
  (void)M;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
}
static int pick_cluster(int M) {
  const int forced = env_int("FEDB200_CLUSTER", -1);
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  const int mt = (M + IG_BLOCK_M - 1) / IG_BLOCK_M;
  if (mt >= 8) return 4;
  if (mt >= 2) return 2;
  return 1;
}

// ------------------------------------------------------------------------------------------------
void linear_tf32(const float* x, const float* w, const float* bias, float* out, int M, int N, int K, int ldx, int ldw,
                 int ldo, int act, cudaStream_t stream) {
  if ((ldx & 3) || (ldw & 3) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15))
    throw std::runtime_error("fedb200: linear_tf32 needs 16-byte aligned rows");
  const int bn = pick_block_n(M, N);
  const int cl = pick_cluster(M);
  CUtensorMap ta = make_tmap_2d(x, M, K, ldx, IG_BLOCK_M);
  CUtensorMap tb = make_tmap_2d(w, N, K, ldw, bn / cl);
  IgemmParams p{};
  p.M = M; p.N = N;
  p.cblocks = (K + IG_BLOCK_K - 1) / IG_BLOCK_K;
  p.num_k_blocks = p.cblocks;
  p.taps_w = 1; p.b_cols_per_tap = 0; p.is_conv = 0;
  p.out = out; p.ldo = ldo; p.bias = bias; p.act = act; p.stats = nullptr;
  dispatch(bn, cl, ta, tb, p, stream);
}

bool conv_geometry_supported(int H_out, int W_out, int C_in, int stride) {
  if (W_out <= 0 || H_out <= 0 || W_out > 128 || (128 % W_out) != 0) return false;
  const int rows = 128 / W_out;                 // image rows (possibly spanning images) per 128-pixel tile
  if (rows <= H_out ? (H_out % rows) != 0 : (rows % H_out) != 0) return false;
  if ((C_in & 3) != 0) return false;            // 16-byte pixel pitch for TMA
  if (stride < 1 || stride > 2) return false;
  if (W_out * stride > 256) return false;
  return true;
}

void conv2d_nhwc_tf32(const float* x, const float* w, float* y, float* stats, int NB, int H, int W, int C_in, int C_out,
                      int kh, int kw, int stride, int pad, int dil, int H_out, int W_out, cudaStream_t stream) {
  if (!conv_geometry_supported(H_out, W_out, C_in, stride))
    throw std::runtime_error("fedb200: conv geometry not supported by the tcgen05 path");
  const int rows = 128 / W_out;
  const int boxH = rows <= H_out ? rows : H_out;
  const int boxN = rows <= H_out ? 1 : rows / H_out;
  const int M = NB * H_out * W_out;
  const int bn = pick_block_n(M, C_out);
  const int cl = pick_cluster(M);
  CUtensorMap ta = make_tmap_nhwc(x, NB, H, W, C_in, boxN, boxH, W_out, stride);
  CUtensorMap tb = make_tmap_2d(w, C_out, uint64_t(kh) * kw * C_in, uint64_t(kh) * kw * C_in, bn / cl);
  IgemmParams p{};
  p.M = M; p.N = C_out;
  p.cblocks = (C_in + IG_BLOCK_K - 1) / IG_BLOCK_K;
  p.num_k_blocks = kh * kw * p.cblocks;
  p.taps_w = kw; p.b_cols_per_tap = C_in; p.is_conv = 1;
  p.HW_out = H_out * W_out; p.W_out = W_out;
  p.stride = stride; p.pad = pad; p.dil = dil;
  p.out = y; p.ldo = C_out; p.bias = nullptr; p.act = 0; p.stats = stats;
  dispatch(bn, cl, ta, tb, p, stream);
}

}  // namespace fedb200
