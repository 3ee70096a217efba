// Host side of the row-streaming square kernel: eligibility, geometry, dispatch.
#include "nbh_square_rs.cuh"

namespace nbh {

int env_int(const char* name, int dflt);

// Decide whether the launch can use the row-streaming kernel and lay it out.
// Requirements: fused member path (n_members >= 2), no periodic x, even nx and
// 8-byte aligned rows (float2 reads of the staged rows), member stride a multiple
// of four floats (one alignment shift per row), the whole row covered by at most
// kRsMaxWindows 64-column windows per x panel, and at least two stages of shared memory.
bool plan_square_rs(const NbhDesc& d, const float* in, SqRsGeom* g, int* grid) {
  if (!env_int("NBH_SQ_RS", 1)) return false;
  const int r = d.cells, nb = 2 * r + 1, R = d.n_members;
  const bool rows_mode = R == 1;
  if (rows_mode && !env_int("NBH_SQ_RS_ROWS", 1)) return false;
  if ((d.flags & NBH_WRAP_X) || d.mask_nd) return false;
  if ((d.nx & 1) || (reinterpret_cast<uintptr_t>(in) & 7) || (d.plane_stride & 1) || (d.member_stride & 3))
    return false;
  const int hl = (r + 1) & ~1;
  const int W = (64 - hl - r) & ~1;
  const int W0 = (64 - r) & ~1;
  if (W < 32) return false;
  // rows wider than kRsMaxWindows windows are split into x panels of equal window count
  const int n_windows_total = 1 + (d.nx > W0 ? (d.nx - W0 + W - 1) / W : 0);
  const int n_panels = (n_windows_total + kRsMaxWindows - 1) / kRsMaxWindows;
  const int n_windows = (n_windows_total + n_panels - 1) / n_panels;
  if (rows_mode && n_panels > 1) return false;          // a stage is one contiguous copy of full rows
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  const int seg_cols = min(d.nx, n_windows * W + hl + r + (W0 - W));    // widest staged segment
  const int rps = rows_mode ? max(1, min(kRsBatch, env_int("NBH_SQ_RS_RPS", kRsBatch))) : 0;
  const int pitch = ((rows_mode ? rps * d.nx : seg_cols) + 3 + 3) / 4 * 4;   // up to 3 floats of alignment shift
  const size_t fixed = sq_rs_fixed_smem(nb, n_windows, R == 1 ? kRsBatch : 1) + 256;
  if (fixed + 2 * (size_t)pitch * 4 > (size_t)smem_max) return false;
  const size_t avail = (size_t)smem_max - fixed;
  // items per stage: the largest divisor of R whose stage stays below ~1/5 of the ring
  const size_t stage_cap = max((size_t)pitch * 4, min((size_t)32 * 1024, avail / 5));
  int ms = 1;
  for (int c = 1; c <= R; ++c)
    if (R % c == 0 && (size_t)c * pitch * 4 <= stage_cap) ms = c;
  ms = max(1, min(ms, env_int("NBH_SQ_RS_MS", ms)));
  while (R % ms) --ms;
  int n_stages = (int)min((size_t)kRsMaxStages, avail / ((size_t)ms * pitch * 4));
  n_stages = min(n_stages, env_int("NBH_SQ_RS_STAGES", kRsMaxStages));
  if (n_stages < 2) return false;
  g->n_windows = n_windows; g->n_panels = n_panels; g->n_windows_total = n_windows_total;
  g->W0 = W0; g->W = W; g->hl = hl;
  g->ms = ms; g->n_stages = n_stages; g->pitch = pitch; g->rps = rps;
  // fixed-point scale: nb^2 * R * 2^shift < 2^31 (window sums are exact modulo 2^32)
  int shift = 0;
  while (shift < 22 && (double)nb * nb * R * (double)(1u << (shift + 1)) < 2147483648.0) ++shift;
  if (d.n_thresholds > 0 && shift < 12) return false;   // huge windows: keep the float64 kernels
  g->fx_shift = shift;
  g->wait_ns = (unsigned)env_int("NBH_SQ_RS_WAIT_NS", 1000);
  g->rows_total = d.n_planes * (long long)max(d.n_thresholds, 1) * n_panels * d.ny;
  // one block per SM; every share at least ~4r rows so the y halo stays a minor cost
  const long long min_share = max(16, 4 * r);
  long long blocks = min((long long)sm_count(), max(1LL, g->rows_total / min_share));
  blocks = min(blocks, (long long)env_int("NBH_SQ_RS_BLOCKS", (int)blocks));
  g->share = (g->rows_total + blocks - 1) / blocks;
  *grid = (int)((g->rows_total + g->share - 1) / g->share);
  return true;
}

template <int TM, bool M2D>
static int launch_rs_one(const SqParams& p, const SqRsGeom& g, int grid, cudaStream_t st) {
  auto kern = square_rs_kernel<TM, M2D>;
  const size_t smem = (size_t)g.n_stages * g.ms * g.pitch * 4 + sq_rs_fixed_smem(p.nb, g.n_windows, g.rps ? kRsBatch : 1);
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)grid, (g.n_windows + kRsProducers) * 32, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_square_rs(const SqParams& p, const SqRsGeom& g, int tm, bool m2d, int grid, cudaStream_t st) {
  switch ((tm << 1) | (m2d ? 1 : 0)) {
    case 0: return launch_rs_one<0, false>(p, g, grid, st);
    case 1: return launch_rs_one<0, true>(p, g, grid, st);
    case 2: return launch_rs_one<1, false>(p, g, grid, st);
    case 3: return launch_rs_one<1, true>(p, g, grid, st);
    case 4: return launch_rs_one<2, false>(p, g, grid, st);
    case 5: return launch_rs_one<2, true>(p, g, grid, st);
    case 8: return launch_rs_one<4, false>(p, g, grid, st);
    case 9: return launch_rs_one<4, true>(p, g, grid, st);
    case 10: return launch_rs_one<5, false>(p, g, grid, st);
    case 11: return launch_rs_one<5, true>(p, g, grid, st);
    default: set_error("unsupported row-stream kernel variant"); return 1;
  }
}

}  // namespace nbh
