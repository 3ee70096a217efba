// Host side of the multi-threshold fused kernel (nbh_square_mt.cuh) and the edge-band flag kernel.
#include "nbh_square_mt.cuh"

namespace nbh {

template <int R, int TM, bool M2D>
static int launch_mt_one(const SqParams& p, const SqMtGeom& g, int grid, size_t smem, cudaStream_t st) {
  auto kern = square_mt_kernel<R, TM, M2D>;
  static bool configured = false;
  if (!configured) {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  kern<<<(unsigned)grid, (g.n_windows * g.n_thr + kMtProducers) * 32, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int R>
int launch_mt_r(const SqParams& p, const SqMtGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st) {
  const int tmk = tm == 4 ? 2 : tm;
  switch ((tmk << 1) | (m2d ? 1 : 0)) {
    case 2: return launch_mt_one<R, 1, false>(p, g, grid, smem, st);
    case 3: return launch_mt_one<R, 1, true>(p, g, grid, smem, st);
    case 4: return launch_mt_one<R, 2, false>(p, g, grid, smem, st);
    case 5: return launch_mt_one<R, 2, true>(p, g, grid, smem, st);
    case 10: return launch_mt_one<R, 5, false>(p, g, grid, smem, st);
    case 11: return launch_mt_one<R, 5, true>(p, g, grid, smem, st);
    default: set_error("unsupported multi-threshold kernel variant"); return 1;
  }
}

#ifdef NBH_MT_R
// one translation unit per radius (compile parallelism): nbh_square_mt_r<R>.cu
template int launch_mt_r<NBH_MT_R>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
#else

// Eligibility: fused member path with at least two thresholds, radius 1..8 cells, no periodic x,
// even nx, 8-byte aligned slices, member stride a multiple of four floats.
bool plan_square_mt(const NbhDesc& d, const float* in, SqMtGeom* g, int* grid, size_t* smem) {
  const int r = d.cells, nb = 2 * r + 1, R = d.n_members;
  if (R < 2 || d.n_thresholds < 2 || r < 1 || r > 8) return false;
  if ((d.flags & (NBH_WRAP_X | NBH_SUM_ONLY)) || d.mask_nd) return false;
  if ((d.nx & 1) || d.nx < 16 || (reinterpret_cast<uintptr_t>(in) & 7) || (d.plane_stride & 1) || (d.member_stride & 3))
    return false;
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  const int useful = mt_useful_cols(r);
  const int n_windows_total = (d.nx + useful - 1) / useful;
  const int n_panels = (n_windows_total + kMtWindows - 1) / kMtWindows;
  const int n_windows = (n_windows_total + n_panels - 1) / n_panels;
  const int seg_cols = min(d.nx, n_windows * useful + 2 * mt_halo_lanes(r) * kMtK);
  const int pitch = (seg_cols + 3 + 3) / 4 * 4;
  // member rows per stage: the largest divisor of R with a stage of at most ~16 KB
  int ms = 1;
  for (int c = 1; c <= R; ++c)
    if (R % c == 0 && (size_t)c * pitch * 4 <= (size_t)16 * 1024) ms = c;
  const size_t stage_b = (size_t)ms * pitch * 4;
  // thresholds per pass: as many rings as fit beside at least three stages
  int n_thr = min(kMtMaxThr, d.n_thresholds);
  for (; n_thr >= 1; --n_thr) {
    const size_t rings = (size_t)n_windows * n_thr * nb * kMtWinCols * 4;
    if (rings + 3 * stage_b + 2 * kMtMaxStages * 8 + 256 <= (size_t)smem_max) break;
  }
  if (n_thr < 2) return false;
  const size_t rings = (size_t)n_windows * n_thr * nb * kMtWinCols * 4;
  const size_t fixed = rings + 2 * kMtMaxStages * 8 + 128;
  // two thresholds per pass leave room for larger stages (fewer hand-overs per row; measured on config 3
  // with T = 2: 0.75 -> 0.65 ms); with four, more and smaller stages were faster (1.00 against 1.25 ms)
  if (n_thr <= 2) {
    const size_t cap = ((size_t)smem_max - fixed) / 3;
    for (int c = 1; c <= R; ++c)
      if (R % c == 0 && (size_t)c * pitch * 4 <= cap) ms = c;
  }
  const size_t stage_b2 = (size_t)ms * pitch * 4;
  g->n_windows = n_windows; g->n_panels = n_panels; g->n_windows_total = n_windows_total;
  g->n_thr = n_thr; g->t_base = 0; g->ms = ms; g->pitch = pitch;
  g->n_stages = (int)min((size_t)kMtMaxStages, ((size_t)smem_max - fixed) / stage_b2);
  *smem = fixed + (size_t)g->n_stages * stage_b2;
  int shift = 0;
  while (shift < 24 && (double)nb * nb * R * (double)(1u << (shift + 1)) < 4294967296.0) ++shift;
  if (shift < 12) return false;
  g->fx_shift = shift;
  g->rows_total = d.n_planes * (long long)n_panels * d.ny;
  const long long min_share = max(16, 4 * r);
  long long blocks = min((long long)sm_count(), max(1LL, g->rows_total / min_share));
  g->share = (g->rows_total + blocks - 1) / blocks;
  *grid = (int)((g->rows_total + g->share - 1) / g->share);
  return true;
}

extern template int launch_mt_r<1>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_mt_r<2>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_mt_r<3>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_mt_r<4>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_mt_r<5>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_mt_r<6>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_mt_r<7>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_mt_r<8>(const SqParams&, const SqMtGeom&, int, bool, int, size_t, cudaStream_t);

int launch_square_mt(const SqParams& p, const SqMtGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st) {
  switch (p.r) {
    case 1: return launch_mt_r<1>(p, g, tm, m2d, grid, smem, st);
    case 2: return launch_mt_r<2>(p, g, tm, m2d, grid, smem, st);
    case 3: return launch_mt_r<3>(p, g, tm, m2d, grid, smem, st);
    case 4: return launch_mt_r<4>(p, g, tm, m2d, grid, smem, st);
    case 5: return launch_mt_r<5>(p, g, tm, m2d, grid, smem, st);
    case 6: return launch_mt_r<6>(p, g, tm, m2d, grid, smem, st);
    case 7: return launch_mt_r<7>(p, g, tm, m2d, grid, smem, st);
    case 8: return launch_mt_r<8>(p, g, tm, m2d, grid, smem, st);
    default: set_error("multi-threshold kernel radius %d out of range", p.r); return 1;
  }
}

// ---------------------------------------------------------------------------
// Edge-band flags of the trimming fix-up (improver/nbhood/nbhood.py:353-382): for every slice
// (plane, member, threshold), does each of the four (r+1)-wide edge bands contain a value != 1
// (masked points count)?  If all four do, the non-one bounding box of the slice is the whole
// domain and the reference's ones-trimming cannot apply.  One block per (plane, member), all
// thresholds at once; only the bands are read.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
edge_flags_kernel(const SqParams p) {
  const long long pm = blockIdx.x;
  const long long plane = pm / p.n_members;
  const int m = (int)(pm - plane * p.n_members);
  const float* sl = p.in + plane * p.plane_stride + (long long)m * p.member_stride;
  const int r = p.r, ny = p.ny, nx = p.nx, T = p.n_thr;
  unsigned bits[kMaxThresholds];
#pragma unroll
  for (int t = 0; t < kMaxThresholds; ++t) bits[t] = 0u;
  const int band = min(r + 1, ny), bandx = min(r + 1, nx);
  // rows of the top and bottom bands (all columns), then the left / right column bands of every row
  // (32-bit point arithmetic: the bands of a slice hold far fewer than 2^31 points, checked by the launcher)
  const unsigned n_rowpts = 2u * (unsigned)band * (unsigned)nx, n_colpts = 2u * (unsigned)bandx * (unsigned)ny;
  // blockIdx.y: one of gridDim.y interleaved parts of the band points of the slice
  for (unsigned i = blockIdx.y * blockDim.x + threadIdx.x; i < n_rowpts + n_colpts; i += gridDim.y * blockDim.x) {
    int y, x;
    if (i < n_rowpts) {
      const int k = (int)(i / (unsigned)nx);
      x = (int)(i - (unsigned)k * (unsigned)nx);
      y = k < band ? k : ny - band + (k - band);
    } else {
      const unsigned j = i - n_rowpts;
      y = (int)(j / (2u * (unsigned)bandx));
      const int k = (int)(j - (unsigned)y * 2u * (unsigned)bandx);
      x = k < bandx ? k : nx - bandx + (k - bandx);
    }
    const long long pt = (long long)y * nx + x;
    const float d = __ldg(sl + pt);
    const bool valid = p.mask2d ? (p.mask2d[pt] != 0) : true;
    const unsigned where = (y <= r ? 1u : 0u) | (y >= ny - 1 - r ? 2u : 0u) | (x <= r ? 4u : 0u) | (x >= nx - 1 - r ? 8u : 0u);
#pragma unroll
    for (int t = 0; t < kMaxThresholds; ++t)
      if (t < T && (!valid || truth_not_one(d, p.thr.t[t]))) bits[t] |= where;
  }
#pragma unroll
  for (int t = 0; t < kMaxThresholds; ++t) {
    if (t < T) {
      const unsigned b = __reduce_or_sync(0xffffffffu, bits[t]);
      if ((threadIdx.x & 31) == 0 && b) atomicOr(p.edge_flags + (pm * T + t), b);
    }
  }
}

int launch_edge_flags(const SqParams& p, int tm, cudaStream_t st) {
  (void)tm;
  const long long n = p.n_planes * p.n_members;
  NBH_REQUIRE(n < (1LL << 31), "too many slices");
  const long long pts = 2LL * (p.r + 1) * ((long long)p.nx + p.ny);
  NBH_REQUIRE(pts < (1LL << 31), "edge bands too large");
  const unsigned parts = (unsigned)max(1LL, min(32LL, pts / 2048));
  edge_flags_kernel<<<dim3((unsigned)n, parts), 256, 0, st>>>(p);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // NBH_MT_R

}  // namespace nbh
