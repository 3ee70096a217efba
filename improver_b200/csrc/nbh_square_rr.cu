// Host side of the row-streaming kernel for single-member slices.
#include "nbh_square_rr.cuh"

namespace nbh {

int env_int(const char* name, int dflt);

static size_t rr_fixed_smem(int nb, int n_windows, int n_cwarps, int pw, int es) {
  return 2 * kRrMaxChunks * 8 + (size_t)((nb + 2) & ~1) * 8 + (size_t)n_windows * 128 * es +
         (size_t)n_cwarps * kRrRows * pw * es;
}

// Eligibility and layout.  Needs float2-aligned rows (even nx, 8-byte aligned
// base and plane stride); a periodic x axis additionally needs 16-byte aligned
// rows (nx % 4 == 0) so that the wrapped halo copies are legal bulk copies.
bool plan_square_rr(const NbhDesc& d, const float* in, int tm, SqRrGeom* g, int* grid, size_t* smem) {
  // opt-in: correct on every tested shape, but with only ~10 math warps per SM the per-row work
  // (float64 scan, conversions) is latency-bound: 2.0 ms vs 1.2 ms for the warp kernel on 240 UK
  // slices
  if (!env_int("NBH_SQ_RR", 0)) return false;
  const int r = d.cells, nb = 2 * r + 1;
  const bool wrap = (d.flags & NBH_WRAP_X) != 0;
  if (d.n_members != 1 || d.mask_nd) return false;
  if ((d.nx & 1) || (reinterpret_cast<uintptr_t>(in) & 7) || (d.plane_stride & 1)) return false;
  if (wrap && ((d.nx & 3) || (reinterpret_cast<uintptr_t>(in) & 15) || (d.plane_stride & 3) || r > 12)) return false;
  const int hl = (r + 3) & ~3;
  const int W = (128 - hl - r) & ~3;
  if (W < 64 || d.nx < 2 * r + 2) return false;
  const int nwt = (d.nx + W - 1) / W;
  const int n_panels = (nwt + kRrMaxCWarps - 1) / kRrMaxCWarps;
  const int nw = (nwt + n_panels - 1) / n_panels;
  const int panel_w = nw * W;
  const int es = tm != 0 ? 4 : 8;
  const int PO = ((r + 3) & ~3) + 3;
  const int pw = (PO + 129 + 4 + 3) & ~3;
  const int seg = min(panel_w + hl + r, d.nx);
  const int seg_pitch = (seg + 3 + 4 + 3) & ~3;
  const int row_pitch = seg_pitch + (wrap ? 2 * kRrHalo : 0);
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  const size_t fixed = rr_fixed_smem(nb, nw, nw, pw, es) + 256;
  if (fixed >= (size_t)smem_max) return false;
  const size_t chunk_bytes = (size_t)kRrRows * row_pitch * 4;
  int n_chunks = (int)min((size_t)kRrMaxChunks, ((size_t)smem_max - fixed) / chunk_bytes);
  n_chunks = min(n_chunks, env_int("NBH_SQ_RR_CHUNKS", kRrMaxChunks));
  if (n_chunks < (nb + 3) / 4 + 3) return false;          // window rows + at least two chunks in flight
  g->n_windows = nw; g->n_cwarps = nw; g->n_panels = n_panels; g->panel_w = panel_w;
  g->W0 = W; g->W = W; g->hl = hl;
  g->n_chunks = n_chunks; g->row_pitch = row_pitch; g->seg_pitch = seg_pitch; g->pw = pw;
  g->rows_total = d.n_planes * (long long)max(d.n_thresholds, 1) * n_panels * d.ny;
  int shift = 0;
  while (shift < 22 && (double)nb * nb * (double)(1u << (shift + 1)) < 2147483648.0) ++shift;
  if (tm != 0 && shift < 12) return false;
  g->fx_shift = shift;
  g->wait_ns = (unsigned)env_int("NBH_SQ_RS_WAIT_NS", 1000);
  const long long min_share = max(16, 4 * r);
  long long blocks = min((long long)sm_count(), max(1LL, g->rows_total / min_share));
  blocks = min(blocks, (long long)env_int("NBH_SQ_RR_BLOCKS", (int)blocks));
  g->share = (g->rows_total + blocks - 1) / blocks;
  *grid = (int)((g->rows_total + g->share - 1) / g->share);
  *smem = (size_t)n_chunks * chunk_bytes + rr_fixed_smem(nb, nw, nw, pw, es);
  return true;
}

template <int TM, bool M2D>
static int launch_rr_one(const SqParams& p, const SqRrGeom& g, int grid, size_t smem, cudaStream_t st) {
  auto kern = square_rr_kernel<TM, M2D>;
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)grid, (g.n_cwarps + kRrProducers) * 32, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_square_rr(const SqParams& p, const SqRrGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st) {
  switch ((tm << 1) | (m2d ? 1 : 0)) {
    case 0: return launch_rr_one<0, false>(p, g, grid, smem, st);
    case 1: return launch_rr_one<0, true>(p, g, grid, smem, st);
    case 2: return launch_rr_one<1, false>(p, g, grid, smem, st);
    case 3: return launch_rr_one<1, true>(p, g, grid, smem, st);
    case 4: return launch_rr_one<2, false>(p, g, grid, smem, st);
    case 5: return launch_rr_one<2, true>(p, g, grid, smem, st);
    case 8: return launch_rr_one<4, false>(p, g, grid, smem, st);
    case 9: return launch_rr_one<4, true>(p, g, grid, smem, st);
    case 10: return launch_rr_one<5, false>(p, g, grid, smem, st);
    case 11: return launch_rr_one<5, true>(p, g, grid, smem, st);
    default: set_error("unsupported row-stream (rows) kernel variant"); return 1;
  }
}

}  // namespace nbh
