// Host side of the two-phase row-streaming kernel for single-member slices.
#include "nbh_square_r2.cuh"

namespace nbh {

int env_int(const char* name, int dflt);

// Eligibility: single-member launches without a masked-array input, even nx (float2 stores),
// at most 2 * kR2CThreads columns (phase V: two columns per thread), 8-byte aligned slices
// (16-byte aligned bulk copies need even element offsets), and shared memory for the H ring
// plus at least three stages.
bool plan_square_r2(const NbhDesc& d, const float* in, SqR2Geom* g, int* grid, size_t* smem) {
  // opt-in: parity-green, but its phases are latency-bound (2.0 ms vs 1.08 ms for the batched rows
  // mode of square_rs_kernel on 240 UK slices): every phase runs one short dependent chain per
  // thread between two block-level barriers
  if (!env_int("NBH_SQ_R2", 0)) return false;
  const int r = d.cells, nb = 2 * r + 1;
  if (d.n_members != 1 || d.mask_nd) return false;
  if ((d.nx & 1) || d.nx > 2 * kR2CThreads || d.nx < 2 * r + 2) return false;
  if ((reinterpret_cast<uintptr_t>(in) & 3) || d.plane_stride < (long long)d.ny * d.nx) return false;
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  const int hpitch = (d.nx + 1) & ~1;
  const int ring_rows = nb + kR2Rows;
  const int pitch = (kR2Rows * d.nx + 3 + 3) / 4 * 4;
  const size_t fixed = 2 * kR2MaxStages * 8 + 16 + (size_t)((nb + 2) & ~1) * 8 + (size_t)ring_rows * hpitch * 4 + 256;
  if (fixed + 3 * (size_t)pitch * 4 > (size_t)smem_max) return false;
  int n_stages = (int)min((size_t)kR2MaxStages, ((size_t)smem_max - fixed) / ((size_t)pitch * 4));
  n_stages = min(n_stages, env_int("NBH_SQ_R2_STAGES", kR2MaxStages));
  if (n_stages < 2) return false;
  g->n_stages = n_stages; g->pitch = pitch; g->hpitch = hpitch; g->ring_rows = ring_rows;
  g->rows_total = d.n_planes * (long long)max(d.n_thresholds, 1) * d.ny;
  g->wait_ns = 1000;
  const long long min_share = max(16, 4 * r);
  long long blocks = min((long long)sm_count(), max(1LL, g->rows_total / min_share));
  g->share = (g->rows_total + blocks - 1) / blocks;
  *grid = (int)((g->rows_total + g->share - 1) / g->share);
  *smem = (size_t)n_stages * pitch * 4 + fixed - 256;
  return true;
}

template <int TM, bool M2D>
static int launch_r2_one(const SqParams& p, const SqR2Geom& g, int grid, size_t smem, cudaStream_t st) {
  auto kern = square_r2_kernel<TM, M2D>;
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)grid, kR2Threads, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_square_r2(const SqParams& p, const SqR2Geom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st) {
  switch ((tm << 1) | (m2d ? 1 : 0)) {
    case 0: return launch_r2_one<0, false>(p, g, grid, smem, st);
    case 1: return launch_r2_one<0, true>(p, g, grid, smem, st);
    case 2: return launch_r2_one<1, false>(p, g, grid, smem, st);
    case 3: return launch_r2_one<1, true>(p, g, grid, smem, st);
    case 4: return launch_r2_one<2, false>(p, g, grid, smem, st);
    case 5: return launch_r2_one<2, true>(p, g, grid, smem, st);
    case 8: return launch_r2_one<4, false>(p, g, grid, smem, st);
    case 9: return launch_r2_one<4, true>(p, g, grid, smem, st);
    case 10: return launch_r2_one<5, false>(p, g, grid, smem, st);
    case 11: return launch_r2_one<5, true>(p, g, grid, smem, st);
    default: set_error("unsupported two-phase kernel variant"); return 1;
  }
}

}  // namespace nbh
