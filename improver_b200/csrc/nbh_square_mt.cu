// Host dispatch of the multi-threshold fused square kernel.
#include "nbh_square_mt.cuh"

namespace nbh {

template <int TM, bool M2D, bool FULL>
static int launch_mt_one(const SqParams& p, const SqWarpGeom& g, const SqMtParams& mt, cudaStream_t st) {
  auto kern = square_warp_mt_kernel<TM, M2D, FULL, 18>;
  const size_t smem = sq_mt_smem_bytes(g.WS, p.nb);
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NBH_REQUIRE(g.n_tasks < (1LL << 31), "too many blocks (%lld); split the launch", g.n_tasks);
  kern<<<(unsigned)g.n_tasks, 32, smem, st>>>(p, g, mt);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int TM, bool M2D>
static int launch_mt_full(const SqParams& p, const SqWarpGeom& g, const SqMtParams& mt, cudaStream_t st) {
  return (p.n_members % 18 == 0) ? launch_mt_one<TM, M2D, true>(p, g, mt, st)
                                 : launch_mt_one<TM, M2D, false>(p, g, mt, st);
}

int launch_square_mt(const SqParams& p, const SqWarpGeom& g, const SqMtParams& mt, int tm, bool m2d,
                     cudaStream_t st) {
  switch ((tm << 1) | (m2d ? 1 : 0)) {
    case 2: return launch_mt_full<1, false>(p, g, mt, st);
    case 3: return launch_mt_full<1, true>(p, g, mt, st);
    case 4: return launch_mt_full<2, false>(p, g, mt, st);
    case 5: return launch_mt_full<2, true>(p, g, mt, st);
    case 8: return launch_mt_full<4, false>(p, g, mt, st);
    case 9: return launch_mt_full<4, true>(p, g, mt, st);
    case 10: return launch_mt_full<5, false>(p, g, mt, st);
    case 11: return launch_mt_full<5, true>(p, g, mt, st);
    default: set_error("unsupported multi-threshold kernel variant"); return 1;
  }
}

}  // namespace nbh
