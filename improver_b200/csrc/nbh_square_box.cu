// Host side of the box kernel (nbh_square_box.cuh): eligibility, geometry, dispatch over the
// compile-time radius, and the guarded follow-up kernels (range check, clamp).
#include "nbh_square_box.cuh"

namespace nbh {

template <int R, int TM, bool M2D>
static int launch_box_one(const SqParams& p, const SqBoxGeom& g, int grid, size_t smem, cudaStream_t st) {
  auto kern = square_box_kernel<R, TM, M2D>;
  static bool configured = false;                       // one attribute call per instantiation
  if (!configured) {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  kern<<<(unsigned)grid, (g.n_windows + kBoxProducers) * 32, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int R>
int launch_box_r(const SqParams& p, const SqBoxGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st) {
  const int tmk = tm == 4 ? 2 : tm;                     // the two-instruction single-slope form is not instantiated here
  switch ((tmk << 1) | (m2d ? 1 : 0)) {
    case 0: return launch_box_one<R, 0, false>(p, g, grid, smem, st);
    case 1: return launch_box_one<R, 0, true>(p, g, grid, smem, st);
    case 2: return launch_box_one<R, 1, false>(p, g, grid, smem, st);
    case 3: return launch_box_one<R, 1, true>(p, g, grid, smem, st);
    case 4: return launch_box_one<R, 2, false>(p, g, grid, smem, st);
    case 5: return launch_box_one<R, 2, true>(p, g, grid, smem, st);
    case 10: return launch_box_one<R, 5, false>(p, g, grid, smem, st);
    case 11: return launch_box_one<R, 5, true>(p, g, grid, smem, st);
    default: set_error("unsupported box kernel variant"); return 1;
  }
}

#ifdef NBH_BOX_R
// one translation unit per radius (compile parallelism): nbh_square_box_r<R>.cu
template int launch_box_r<NBH_BOX_R>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
#else

// Eligibility: single-member slices, radius 1..8 cells, no periodic x, no masked-array input,
// even nx and 8-byte aligned slices (8-byte shared-memory reads of the staged rows).
bool plan_square_box(const NbhDesc& d, const float* in, SqBoxGeom* g, int* grid, size_t* smem) {
  const int r = d.cells, nb = 2 * r + 1;
  if (d.n_members != 1 || r < 1 || r > 8) return false;
  if ((d.flags & NBH_WRAP_X) || d.mask_nd) return false;
  if ((d.nx & 1) || d.nx < 16 || (reinterpret_cast<uintptr_t>(in) & 7) || (d.plane_stride & 1)) return false;
  const bool m2d = d.mask2d != nullptr;
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  const int useful = box_useful_cols(r);
  const int n_windows_total = (d.nx + useful - 1) / useful;
  const int n_panels = (n_windows_total + kBoxMaxWindows - 1) / kBoxMaxWindows;
  const int n_windows = (n_windows_total + n_panels - 1) / n_panels;
  g->n_windows = n_windows; g->n_panels = n_panels; g->n_windows_total = n_windows_total;
  const int B = m2d ? kBoxBMasked : kBoxBPlain;
  g->B = B;
  g->contiguous = n_panels == 1 ? 1 : 0;
  const int seg_cols = min(d.nx, n_windows * useful + 2 * box_halo_lanes(r) * kBoxK);
  g->seg_pitch = (seg_cols + 3 + 3) / 4 * 4;
  const int rows_floats = g->contiguous ? (B * d.nx + 3 + 3) / 4 * 4 : B * g->seg_pitch;
  g->mask_off = rows_floats;
  // uint8 mask rows behind the data rows: 16-byte copies of unaligned rows, one spare word for the
  // aligned-word reads of the last lane
  g->mask_pitch = (seg_cols + 15 + 15 + 4) / 16 * 16;
  const int mask_bytes = g->contiguous ? (B * d.nx + 15 + 15 + 4) / 16 * 16 : B * g->mask_pitch;
  g->stage_floats = rows_floats + (m2d ? mask_bytes / 4 : 0);
  const size_t ring = (size_t)n_windows * nb * kBoxWinCols * 4;
  const size_t fixed = 2 * kBoxMaxStages * 8 + ring + 128;
  // two blocks per SM when they fit with at least three stages each
  const size_t stage_b = (size_t)g->stage_floats * 4;
  size_t budget = (size_t)smem_max;
  if (fixed + 3 * stage_b <= (size_t)smem_max / 2 - 1024) budget = (size_t)smem_max / 2 - 1024;
  if (fixed + 2 * stage_b > budget) return false;
  g->n_stages = (int)min((size_t)kBoxMaxStages, (budget - fixed) / stage_b);
  *smem = fixed + (size_t)g->n_stages * stage_b;
  int shift = 0;
  while (shift < 22 && (double)nb * nb * (double)(1u << (shift + 1)) < 4294967296.0) ++shift;   // <= 22: box_to_fixed
  g->fx_shift = shift;
  g->rows_total = d.n_planes * (long long)max(d.n_thresholds, 1) * n_panels * d.ny;
  const int per_sm = budget < (size_t)smem_max ? 2 : 1;
  const long long min_share = max(32, 8 * r);
  long long blocks = min((long long)sm_count() * per_sm, max(1LL, g->rows_total / min_share));
  g->share = (g->rows_total + blocks - 1) / blocks;
  *grid = (int)((g->rows_total + g->share - 1) / g->share);
  return true;
}

extern template int launch_box_r<1>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_box_r<2>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_box_r<3>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_box_r<4>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_box_r<5>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_box_r<6>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_box_r<7>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);
extern template int launch_box_r<8>(const SqParams&, const SqBoxGeom&, int, bool, int, size_t, cudaStream_t);

int launch_square_box(const SqParams& p, const SqBoxGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st) {
  switch (p.r) {
    case 1: return launch_box_r<1>(p, g, tm, m2d, grid, smem, st);
    case 2: return launch_box_r<2>(p, g, tm, m2d, grid, smem, st);
    case 3: return launch_box_r<3>(p, g, tm, m2d, grid, smem, st);
    case 4: return launch_box_r<4>(p, g, tm, m2d, grid, smem, st);
    case 5: return launch_box_r<5>(p, g, tm, m2d, grid, smem, st);
    case 6: return launch_box_r<6>(p, g, tm, m2d, grid, smem, st);
    case 7: return launch_box_r<7>(p, g, tm, m2d, grid, smem, st);
    case 8: return launch_box_r<8>(p, g, tm, m2d, grid, smem, st);
    default: set_error("box kernel radius %d out of range", p.r); return 1;
  }
}

// ---------------------------------------------------------------------------
// Follow-ups of a box launch.  box_decide_kernel looks at the extremes of every unit
// (slice x threshold) and sets
//   redo[u]  = 1: the raw slice is not a field of probabilities in [0, 1] (or has non-finite
//              extremes): the fixed-point result is void, the float64 kernels recompute it;
//   clamp[u] = 1: a result left [min, max] of the (thresholded) slice by the quantisation:
//              box_clamp_kernel clips the slice (nbhood.py:312).
// With statistics of the whole slices handed in (y-band launches) they decide, so that every
// band takes the same decision as the untiled launch.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float box_key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void box_init_kernel(BoxStats* stats, int* redo, float2* range, int* n_redo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_redo = 0;
  if (i >= n) return;
  BoxStats s;
  s.in_min = INT_MAX; s.in_max = INT_MIN; s.out_min = INT_MAX; s.out_max = INT_MIN;
  stats[i] = s;
  redo[i] = 0;
  range[i] = make_float2(0.f, 0.f);
}

__global__ void box_decide_kernel(const BoxStats* stats, const NbhSliceStats* ext, ThrSet thr, int tm, int n_thr,
                                  int sum_only, int* redo, float2* range, int* clamp_list, int* n_clamp,
                                  int* n_redo, long long n) {
  const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const BoxStats s = stats[u];
  const ThrDev& t = thr.t[tm ? (int)(u % n_thr) : 0];
  float dmin = box_key_float(s.in_min), dmax = box_key_float(s.in_max);
  bool bad = false;
  float lo, hi;
  if (tm == 0) {
    if (ext) { dmin = box_key_float(ext[u].vmin_key); dmax = box_key_float(ext[u].vmax_key); }
    // raw slices are processed as probabilities: anything else is recomputed in float64
    bad = !(dmin >= 0.f && dmax <= 1.f);
    lo = dmin; hi = dmax;
  } else if (ext) {
    lo = box_key_float(ext[u].vmin_key); hi = box_key_float(ext[u].vmax_key);   // already truth values
  } else {
    // the truth value is monotone in the data: the extremes of the thresholded slice are the
    // (exact) truth values of the extremes of the data
    const float a = truth_value(dmin, t), b = truth_value(dmax, t);
    lo = fminf(a, b); hi = fmaxf(a, b);
  }
  if (bad) { redo[u] = 1; atomicAdd(n_redo, 1); return; }
  if (sum_only) return;
  range[u] = make_float2(lo, hi);
  const float omin = box_key_float(s.out_min), omax = box_key_float(s.out_max);
  if (s.out_min != INT_MAX && (omin < lo || omax > hi)) clamp_list[atomicAdd(n_clamp, 1)] = (int)u;
}

__global__ void __launch_bounds__(256)
box_clamp_kernel(float* out, const float2* range, const int* clamp_list, const int* n_clamp, long long per_slice) {
  const int n = *n_clamp;
  const int chunks = 64;
  for (long long w = blockIdx.x; w < (long long)n * chunks; w += gridDim.x) {
    const int u = clamp_list[w / chunks];
    const int c = (int)(w % chunks);
    const float2 rg = range[u];
    float* o = out + (long long)u * per_slice;
    for (long long i = (long long)c * blockDim.x + threadIdx.x; i < per_slice; i += (long long)chunks * blockDim.x) {
      const float v = o[i];
      if (v < rg.x) o[i] = rg.x;                       // NaN compares false: stays NaN (nbhood.py:311-312)
      else if (v > rg.y) o[i] = rg.y;
    }
  }
}

size_t box_workspace_bytes(long long n_units, int ny, int nx) {
  const size_t a = ((size_t)n_units * sizeof(BoxStats) + 255) / 256 * 256;
  const size_t b = ((size_t)n_units * 4 + 255) / 256 * 256;
  const size_t c = ((size_t)n_units * 8 + 255) / 256 * 256;
  (void)ny; (void)nx;
  return a + 2 * b + c + 512;
}

void box_carve(void* ws, long long n_units, int ny, int nx, BoxWorkspace* w) {
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  size_t off = 0;
  w->stats = reinterpret_cast<BoxStats*>(base + off); off += ((size_t)n_units * sizeof(BoxStats) + 255) / 256 * 256;
  w->redo = reinterpret_cast<int*>(base + off); off += ((size_t)n_units * 4 + 255) / 256 * 256;
  w->clamp_list = reinterpret_cast<int*>(base + off); off += ((size_t)n_units * 4 + 255) / 256 * 256;
  w->range = reinterpret_cast<float2*>(base + off); off += ((size_t)n_units * 8 + 255) / 256 * 256;
  w->n_clamp = reinterpret_cast<int*>(base + off); off += 256;
  w->n_redo = reinterpret_cast<int*>(base + off); off += 256;
  (void)ny; (void)nx;
}

int box_prepare(const BoxWorkspace& w, const uint8_t* mask2d, long long n_units, int ny, int nx, cudaStream_t st) {
  box_init_kernel<<<(unsigned)((n_units + 255) / 256), 256, 0, st>>>(w.stats, w.redo, w.range, w.n_redo, n_units);
  NBH_CHECK_CUDA(cudaMemsetAsync(w.n_clamp, 0, sizeof(int), st));
  (void)mask2d; (void)ny; (void)nx;
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int box_finish(const BoxWorkspace& w, const SqParams& p, const NbhSliceStats* ext, int tm, long long n_units,
               cudaStream_t st) {
  const int sum_only = (p.flags & NBH_SUM_ONLY) ? 1 : 0;
  box_decide_kernel<<<(unsigned)((n_units + 255) / 256), 256, 0, st>>>(w.stats, ext, p.thr, tm, p.n_thr, sum_only,
                                                                     w.redo, w.range, w.clamp_list, w.n_clamp,
                                                                     w.n_redo, n_units);
  if (!sum_only)
    box_clamp_kernel<<<(unsigned)min((long long)sm_count() * 8, n_units * 64), 256, 0, st>>>(
        p.out, w.range, w.clamp_list, w.n_clamp, (long long)p.ny * p.nx);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // NBH_BOX_R

}  // namespace nbh
