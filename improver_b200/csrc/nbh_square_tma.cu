// Host side of the TMA-fed square kernel: tensor-map creation and dispatch.
#include "nbh_square_tma.cuh"

namespace nbh {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Tensor map over the input viewed as (members, planes, rows, columns); rows are
// pairs of grid rows when the row pitch is not a multiple of 16 bytes.
bool make_square_tensor_map(const float* in, int ny, int nx, long long n_planes, int n_members,
                            long long plane_stride, long long member_stride, int box_cols, CUtensorMap* map,
                            int* super_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (plane_stride & 3) || (member_stride & 3)) return false;
  int super = 0;
  if (nx % 4 != 0) {
    if (nx % 2 != 0 || ny % 2 != 0) return false;
    super = 1;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)(super ? 2 * nx : nx), (cuuint64_t)(super ? ny / 2 : ny),
                              (cuuint64_t)n_planes, (cuuint64_t)n_members};
  const cuuint64_t strides[3] = {(cuuint64_t)(super ? 2 * nx : nx) * 4, (cuuint64_t)plane_stride * 4,
                                 (cuuint64_t)(n_members > 1 ? member_stride : plane_stride) * 4};
  const cuuint32_t box[4] = {(cuuint32_t)box_cols, 1, 1, (cuuint32_t)n_members};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return false;
  *super_rows = super;
  return true;
}

template <int TM, bool M2D>
static int launch_tma_one(const SqParams& p, const SqWarpGeom& g, const SqTmaGeom& tg, const CUtensorMap* map,
                          cudaStream_t st) {
  auto kern = square_tma_kernel<TM, M2D>;
  const size_t smem = sq_tma_smem_bytes(g.WS, g.W, p.nb, p.n_members, tg.n_stages);
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NBH_REQUIRE(g.n_tasks < (1LL << 31), "too many blocks (%lld); split the launch", g.n_tasks);
  kern<<<(unsigned)g.n_tasks, kTmaWarps * 32, smem, st>>>(p, g, tg, map);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_square_tma(const SqParams& p, const SqWarpGeom& g, const SqTmaGeom& tg, const CUtensorMap* map, int tm,
                      bool m2d, cudaStream_t st) {
  NBH_REQUIRE(g.WS == 64, "unsupported TMA window width %d", g.WS);
  switch ((tm << 1) | (m2d ? 1 : 0)) {
    case 0: return launch_tma_one<0, false>(p, g, tg, map, st);
    case 1: return launch_tma_one<0, true>(p, g, tg, map, st);
    case 2: return launch_tma_one<1, false>(p, g, tg, map, st);
    case 3: return launch_tma_one<1, true>(p, g, tg, map, st);
    case 4: return launch_tma_one<2, false>(p, g, tg, map, st);
    case 5: return launch_tma_one<2, true>(p, g, tg, map, st);
    case 8: return launch_tma_one<4, false>(p, g, tg, map, st);
    case 9: return launch_tma_one<4, true>(p, g, tg, map, st);
    case 10: return launch_tma_one<5, false>(p, g, tg, map, st);
    case 11: return launch_tma_one<5, true>(p, g, tg, map, st);
    default: set_error("unsupported TMA kernel variant"); return 1;
  }
}

}  // namespace nbh
