// Square neighbourhood of single-member slices, tile kernel: separable running sums inside
// independent shared-memory tiles.
//
// The scan-based kernels pay a warp prefix scan per row, which only amortises when a row is
// shared by many members.  With one member per slice (NeighbourhoodProcessing.process over the
// slices of a cube, improver/nbhood/nbhood.py:457-463) the cheapest arithmetic is the classic
// separable box filter, and the cheapest way to hide its short dependent chains is to run many
// small independent blocks per SM instead of one persistent pipeline:
//
//   load     a block stages the (TY + 2r) x (TX + 2r) input window of its TY x TX output tile in
//            shared memory (coalesced loads; zero padding or the periodic continuation in x),
//            applying the truth value and the external mask on the way and recording NaNs and the
//            edge-band flags of the trimming fix-up;
//   phase H  one thread per run of kRun columns of a staged row: window sum of the first column
//            directly, the others by S += p[x+r] - p[x-r-1] (float32; the runs are short and the
//            error of an H value is divided by (2r+1)^2 in the mean: < 2e-6 of the data magnitude);
//   phase V  one thread per output column and half tile: vertical window sum of the first row
//            directly, then V += H[y+r] - H[y-r-1]; the mean is the correctly rounded float32
//            quotient V / count (q = V * rc, q += fma(-q, count, V) * rc), exactly 1 on all-ones.
//
// No float64 and no conversion instructions on the unmasked path (conversions run at 16 lanes
// per clock per SM on B200 and were the scarce pipe of the scan kernels' rows mode).
#include "nbh_square.cuh"

namespace nbh {

int env_int(const char* name, int dflt);

constexpr int kTileRun = 9;                       // odd: conflict-free shared-memory strides in phase H
constexpr int kTileTX = 14 * kTileRun;            // 126 output columns
constexpr int kTileTY = 32;
constexpr int kTileThreads = 256;

template <int TM, bool M2D>
__global__ void __launch_bounds__(kTileThreads)
square_tile_kernel(const __grid_constant__ SqParams p, const SqTileGeom g) {
  extern __shared__ __align__(16) float tsm[];
  const int r = p.r, nb = p.nb, nx = p.nx, ny = p.ny;
  const int SW = g.sw, SH = kTileTY + 2 * r;
  const int ro = r + (r & 1);                           // staged column of output column 0 (window starts on an even column)
  float* stage = tsm;                                   // [SH][SW]: P = valid ? truth_value(d) : 0
  float* hbuf = stage + (size_t)SH * SW;                // [SH][kTileTX]
  float* rcp = hbuf + (size_t)SH * kTileTX;             // [n_cnt + 1]: RN(1 / n)
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;

  // 32-bit index arithmetic: the grid has fewer than 2^31 blocks, and 64-bit divisions in every
  // thread's prologue were a quarter of the kernel's instructions
  unsigned b = blockIdx.x;
  const unsigned tiles_x = (unsigned)g.tiles_x, tiles_y = (unsigned)g.tiles_y, n_thr = (unsigned)p.n_thr;
  const unsigned bq = b / tiles_x;
  const int tx = (int)(b - bq * tiles_x);
  const unsigned unit_u = bq / tiles_y;
  const int ty = (int)(bq - unit_u * tiles_y);
  const unsigned plane_u = n_thr > 1 ? unit_u / n_thr : unit_u;
  const long long unit = unit_u;
  const long long plane = plane_u;
  const int t = (int)(unit_u - plane_u * n_thr);
  const ThrDev& thr = p.thr.t[TM ? t : 0];
  const int x0 = tx * kTileTX, y0 = ty * kTileTY;
  const float* src = p.in + plane * p.plane_stride;

  for (int k = threadIdx.x; k <= g.n_cnt; k += kTileThreads) rcp[k] = k ? 1.0f / (float)k : 0.f;

  // ---- load: window rows y0 - r .. y0 + TY - 1 + r, columns x0 - r .. x0 + TX - 1 + r ----
  float nanacc = 0.f;
  unsigned bits = 0;
  const int WC = kTileTX + r + ro + (r & 1);            // staged columns (even: float2 granularity)
  const int xs0 = x0 - ro;                              // first staged column
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // tiles whose whole window lies inside the domain and outside the four edge bands (most of
  // them) take a path without bounds tests, wrap arithmetic or flag bookkeeping
  const bool fast = (y0 - r > r) && (y0 + kTileTY - 1 + r < ny - 1 - r) &&
                    (xs0 > r) && (xs0 + WC - 1 < nx - 1 - r);
  if (fast && g.vec2) {
    // 8-byte loads and stores: the window starts on an even column of 8-byte aligned rows
    const float* wsrc = src + (long long)(y0 - r) * nx + xs0;
    for (int sy = warp; sy < SH; sy += kTileThreads / 32) {
      const float2* grow = reinterpret_cast<const float2*>(wsrc + (long long)sy * nx);
      const uchar2* mrow = M2D ? reinterpret_cast<const uchar2*>(p.mask2d + (long long)(y0 - r + sy) * nx + xs0) : nullptr;
      float2* srow = reinterpret_cast<float2*>(stage + sy * SW);
      for (int sx = lane; sx < WC / 2; sx += 32) {
        const float2 d = __ldg(grow + sx);
        nanacc = max_nan(nanacc, fabsf(d.x));
        nanacc = max_nan(nanacc, fabsf(d.y));
        float2 v = make_float2(truth_value_sum<TM>(d.x, thr), truth_value_sum<TM>(d.y, thr));
        if (M2D) {
          const uchar2 m = __ldg(mrow + sx);
          v.x = m.x ? v.x : 0.f;
          v.y = m.y ? v.y : 0.f;
        }
        srow[sx] = v;
      }
    }
  } else if (!M2D && g.vec2 && (xs0 > r) && (xs0 + WC - 1 < nx - 1 - r)) {
    // tiles at the top / bottom of the domain only: all columns are valid and outside the left /
    // right bands, so rows are either zeros (outside the domain) or plain 8-byte loads; rows in
    // the top / bottom band also feed the trimming flags
    for (int sy = warp; sy < SH; sy += kTileThreads / 32) {
      const int y = y0 - r + sy;
      float2* srow = reinterpret_cast<float2*>(stage + sy * SW);
      if (y < 0 || y >= ny) {
        for (int sx = lane; sx < WC / 2; sx += 32) srow[sx] = make_float2(0.f, 0.f);
        continue;
      }
      const float2* grow = reinterpret_cast<const float2*>(src + (long long)y * nx + xs0);
      const unsigned row_bits = sum_only ? 0u : ((y <= r ? 1u : 0u) | (y >= ny - 1 - r ? 2u : 0u));
      for (int sx = lane; sx < WC / 2; sx += 32) {
        const float2 d = __ldg(grow + sx);
        nanacc = max_nan(nanacc, fabsf(d.x));
        nanacc = max_nan(nanacc, fabsf(d.y));
        srow[sx] = make_float2(truth_value_sum<TM>(d.x, thr), truth_value_sum<TM>(d.y, thr));
        if (row_bits && (truth_not_one(d.x, thr) || truth_not_one(d.y, thr))) bits |= row_bits;
      }
    }
  } else if (fast) {
    const float* wsrc = src + (long long)(y0 - r) * nx + xs0;
    for (int sy = warp; sy < SH; sy += kTileThreads / 32) {
      const float* grow = wsrc + (long long)sy * nx;
      const uint8_t* mrow = M2D ? p.mask2d + (long long)(y0 - r + sy) * nx + xs0 : nullptr;
      float* srow = stage + sy * SW;
      for (int sx = lane; sx < WC; sx += 32) {
        const float d = __ldg(grow + sx);
        nanacc = max_nan(nanacc, fabsf(d));
        const float v = truth_value_sum<TM>(d, thr);
        srow[sx] = (!M2D || mrow[sx]) ? v : 0.f;
      }
    }
  } else {
    // general path: rows / columns outside the domain are zeros (or the periodic continuation);
    // the flag bookkeeping only runs for rows in the top / bottom band and, in tiles that reach
    // the left / right band, for their band columns
    const bool lr_tile = !sum_only && !wrap && ((xs0 <= r) || (xs0 + WC - 1 >= nx - 1 - r));
    for (int sy = warp; sy < SH; sy += kTileThreads / 32) {          // a warp per staged row
      const int y = y0 - r + sy;
      float* srow = stage + sy * SW;
      if (y < 0 || y >= ny) {
        for (int sx = lane; sx < WC; sx += 32) srow[sx] = 0.f;
        continue;
      }
      const long long row_pt = (long long)y * nx;
      const unsigned row_bits = sum_only ? 0u : ((y <= r ? 1u : 0u) | (y >= ny - 1 - r ? 2u : 0u));
      const bool flag_row = row_bits != 0u || lr_tile;
      for (int sx = lane; sx < WC; sx += 32) {
        int x = xs0 + sx;
        if (wrap) { if (x < 0) x += nx; else if (x >= nx) x -= nx; }
        float v = 0.f;
        if (x >= 0 && x < nx) {
          const float d = __ldg(src + row_pt + x);
          nanacc = max_nan(nanacc, fabsf(d));
          const bool ok = !M2D || p.mask2d[row_pt + x] != 0;
          v = ok ? truth_value_sum<TM>(d, thr) : 0.f;
          if (flag_row && (!ok || truth_not_one(d, thr))) {
            bits |= row_bits;
            if (lr_tile && x <= r) bits |= 4u;
            if (lr_tile && x >= nx - 1 - r) bits |= 8u;
          }
        }
        srow[sx] = v;
      }
    }
  }
  if (__isnanf(nanacc)) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  if (!sum_only && !fast) {
    bits = __reduce_or_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && bits) atomicOr(p.edge_flags + (plane * p.n_thr + t), bits);
  }
  __syncthreads();

  // ---- phase H: horizontal window sums of every staged row ----
  constexpr int RUNS = kTileTX / kTileRun;
  for (int task = threadIdx.x; task < SH * RUNS; task += kTileThreads) {
    const int sy = task / RUNS, run = task - sy * RUNS;
    const float* row = stage + sy * SW + run * kTileRun + ro;     // row[c] = staged column of output column c
    float* hrow = hbuf + sy * kTileTX + run * kTileRun;
    float S = 0.f;
    for (int dx = -r; dx <= r; ++dx) S += row[dx];
    hrow[0] = S;
#pragma unroll
    for (int c = 1; c < kTileRun; ++c) {
      S += row[c + r];
      S -= row[c - r - 1];
      hrow[c] = S;
    }
  }
  __syncthreads();

  // ---- phase V: vertical window sums, normalise, store ----
  const int half = threadIdx.x >> 7, cx = threadIdx.x & 127;
  const int x = x0 + cx;
  if (cx >= kTileTX || x >= nx) return;
  const int ya = y0 + half * (kTileTY / 2), yb = min(ny, ya + kTileTY / 2);
  if (ya >= yb) return;
  const int cols_in = wrap ? nb : (min(x + r, nx - 1) - max(x - r, 0) + 1);
  const long long obase = unit * (long long)ny * nx;
  // window of the first row: staged rows (ya - y0) .. (ya - y0 + 2r)
  const float* hcol = hbuf + (size_t)(ya - y0) * kTileTX + cx;
  float V = 0.f;
  for (int k = 0; k < nb; ++k) V += hcol[k * kTileTX];
  float* optr = p.out + obase + (long long)ya * nx + x;
  if (fast && !sum_only && !p.out_mask && !M2D) {
    // every window is complete: one count, one reciprocal
    const float cnt = (float)(nb * nb), rc = rcp[nb * nb];
    for (int y = ya; y < yb; ++y) {
      if (y > ya) {
        V += hcol[(nb - 1) * kTileTX];
        V -= hcol[-kTileTX];
      }
      const float q = V * rc;
      *optr = fmaf(fmaf(-q, cnt, V), rc, q);              // correctly rounded V / cnt
      optr += nx;
      hcol += kTileTX;
    }
    return;
  }
  for (int y = ya; y < yb; ++y) {
    if (y > ya) {
      V += hcol[(nb - 1) * kTileTX];
      V -= hcol[-kTileTX];
    }
    const long long pt = (long long)y * nx + x;
    float o;
    if (sum_only) {
      o = V;
    } else if (M2D) {
      // count of valid cells from the count plane; no valid cell -> NaN (nbhood.py:309-311)
      const float cnt = p.cnt_plane[pt], rc = rcp[(int)cnt];
      const float q = V * rc;
      o = cnt > 0.f ? fmaf(fmaf(-q, cnt, V), rc, q) : __int_as_float(0x7fc00000);
    } else {
      const int rows_in = min(y + r, ny - 1) - max(y - r, 0) + 1;
      const float cnt = (float)(rows_in * cols_in), rc = rcp[rows_in * cols_in];
      const float q = V * rc;
      o = fmaf(fmaf(-q, cnt, V), rc, q);                  // correctly rounded V / cnt
    }
    *optr = o;
    optr += nx;
    if (p.out_mask) p.out_mask[obase + pt] = (M2D && p.mask2d[pt] == 0) ? 1 : 0;
    hcol += kTileTX;
  }
}

// Eligibility: single-member launches without a masked-array input and a radius whose window
// (two buffers of (TY + 2r) rows) fits comfortably in shared memory.
bool plan_square_tile(const NbhDesc& d, const float* in, SqTileGeom* g, long long* blocks, size_t* smem) {
  // default for single-member launches (NBH_SQ_TILE=0 selects the rows mode of square_rs_kernel /
  // the warp kernel): 0.90 vs 1.08 ms on 240 UK slices, 0.76 vs 1.20 ms on 48 periodic 1920 x 2560 slices
  if (!env_int("NBH_SQ_TILE", 1)) return false;
  const int r = d.cells, nb = 2 * r + 1;
  if (d.n_members != 1 || d.mask_nd || r > 16 || d.nx < 2 * r + 2) return false;
  g->tiles_x = (d.nx + kTileTX - 1) / kTileTX;
  g->tiles_y = (d.ny + kTileTY - 1) / kTileTY;
  g->sw = kTileTX + 2 * r + 2 * (r & 1) + 2;             // even pitch (8-byte stores), >= staged columns
  g->vec2 = ((d.nx & 1) == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0 && (d.plane_stride & 1) == 0) ? 1 : 0;
  g->n_cnt = nb * nb;
  *blocks = d.n_planes * (long long)max(d.n_thresholds, 1) * g->tiles_x * g->tiles_y;
  if (*blocks >= (1LL << 31)) return false;
  *smem = ((size_t)(kTileTY + 2 * r) * (g->sw + kTileTX) + g->n_cnt + 1) * 4 + 16;
  return true;
}

template <int TM, bool M2D>
static int launch_tile_one(const SqParams& p, const SqTileGeom& g, long long blocks, size_t smem, cudaStream_t st) {
  auto kern = square_tile_kernel<TM, M2D>;
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, kTileThreads, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_square_tile(const SqParams& p, const SqTileGeom& g, int tm, bool m2d, long long blocks, size_t smem,
                       cudaStream_t st) {
  switch ((tm << 1) | (m2d ? 1 : 0)) {
    case 0: return launch_tile_one<0, false>(p, g, blocks, smem, st);
    case 1: return launch_tile_one<0, true>(p, g, blocks, smem, st);
    case 2: return launch_tile_one<1, false>(p, g, blocks, smem, st);
    case 3: return launch_tile_one<1, true>(p, g, blocks, smem, st);
    case 4: return launch_tile_one<2, false>(p, g, blocks, smem, st);
    case 5: return launch_tile_one<2, true>(p, g, blocks, smem, st);
    case 8: return launch_tile_one<4, false>(p, g, blocks, smem, st);
    case 9: return launch_tile_one<4, true>(p, g, blocks, smem, st);
    case 10: return launch_tile_one<5, false>(p, g, blocks, smem, st);
    case 11: return launch_tile_one<5, true>(p, g, blocks, smem, st);
    default: set_error("unsupported tile kernel variant"); return 1;
  }
}

}  // namespace nbh
