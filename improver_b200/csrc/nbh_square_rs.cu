// Host side of the row-streaming square kernel: eligibility, geometry, dispatch.
#include "nbh_square_rs.cuh"

namespace nbh {

// Decide whether the launch can use the row-streaming kernel and lay it out.
// Requirements: fused member path (n_members >= 2), no periodic x, even nx and
// 8-byte aligned rows (float2 reads of the staged rows), member stride a multiple
// of four floats (one alignment shift per row), the whole row covered by at most
// kRsMaxWindows 64-column windows per x panel, and at least two stages of shared memory.
bool plan_square_rs(const NbhDesc& d, const float* in, SqRsGeom* g, int* grid) {
  const int r = d.cells, nb = 2 * r + 1, R = d.n_members;
  if (R < 2) return false;                              // single-member slices: square_box_kernel
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
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  const int seg_cols = min(d.nx, n_windows * W + hl + r + (W0 - W));    // widest staged segment
  const int pitch = (seg_cols + 3 + 3) / 4 * 4;         // up to 3 floats of alignment shift
  const size_t fixed = sq_rs_fixed_smem(nb, n_windows) + 256;
  if (fixed + 2 * (size_t)pitch * 4 > (size_t)smem_max) return false;
  const size_t avail = (size_t)smem_max - fixed;
  // items per stage: the largest divisor of R that leaves room for three stages.  Every stage costs a
  // consumer warp ~60 instructions of hand-over (barrier wait, arrive, stage pointer), and the kernel
  // runs at 80 % of its issue slots: with 18 members, 6 rows per stage (the ~25 KB stages of round 1)
  // took 0.49 ms per cube, 9 rows 0.46 ms, 18 rows (two stages) 0.456 ms, 3 rows 0.60 ms
  const size_t stage_cap = max((size_t)pitch * 4, avail / 3);
  int ms = 1;
  for (int c = 1; c <= R; ++c)
    if (R % c == 0 && (size_t)c * pitch * 4 <= stage_cap) ms = c;
  int n_stages = (int)min((size_t)kRsMaxStages, avail / ((size_t)ms * pitch * 4));
  if (n_stages < 2) return false;
  g->n_windows = n_windows; g->n_panels = n_panels; g->n_windows_total = n_windows_total;
  g->W0 = W0; g->W = W; g->hl = hl;
  g->ms = ms; g->n_stages = n_stages; g->pitch = pitch;
  // fixed-point scale: nb^2 * R * 2^shift < 2^31 (window sums are exact modulo 2^32)
  int shift = 0;
  while (shift < 22 && (double)nb * nb * R * (double)(1u << (shift + 1)) < 2147483648.0) ++shift;
  if (d.n_thresholds > 0 && shift < 12) return false;   // huge windows: keep the float64 kernels
  g->fx_shift = shift;
  g->wait_ns = 1000;
  g->rows_total = d.n_planes * (long long)max(d.n_thresholds, 1) * n_panels * d.ny;
  // one block per SM; every share at least ~4r rows so the y halo stays a minor cost
  const long long min_share = max(16, 4 * r);
  long long blocks = min((long long)sm_count(), max(1LL, g->rows_total / min_share));
  g->share = (g->rows_total + blocks - 1) / blocks;
  *grid = (int)((g->rows_total + g->share - 1) / g->share);
  return true;
}

template <int TM, bool M2D>
static int launch_rs_one(const SqParams& p, const SqRsGeom& g, int grid, cudaStream_t st) {
  auto kern = square_rs_kernel<TM, M2D>;
  const size_t smem = (size_t)g.n_stages * g.ms * g.pitch * 4 + sq_rs_fixed_smem(p.nb, g.n_windows);
  static bool configured = false;                       // one attribute call per instantiation
  if (!configured) {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
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
