// Square neighbourhood: fused [threshold ->] box-sum -> masked normalisation
// [-> realization mean], hand-written for sm_100a.
//
// Replaces improver/nbhood/nbhood.py:244-409 (method "square") and
// improver/utilities/neighbourhood_tools.py:104-211 (summed-area table boxsum).
//
// Algorithm: a block owns a column strip of one output plane and
// marches down y.  Per input row it (1) loads the row of every member with
// coalesced vector loads, applies the fuzzy threshold and sums the members in
// float32 (P row), (2) updates per-column running vertical window sums V in
// float64 using a shared-memory ring of the last 2r+1 P rows (thread-private
// columns, no synchronisation), and (3) per output row does the horizontal
// window sum as a difference of a float64 row prefix built with a warp-shuffle
// scan, normalises by the valid-cell count and stores float32.
#include "nbh_common.cuh"

namespace nbh {

struct SliceStats {      // exact per-slice statistics, only built for suspicious slices
  int y0min, y0max, x0min, x0max;   // bbox of v != 0
  int y1min, y1max, x1min, x1max;   // bbox of v != 1
  int vmin_key, vmax_key;           // ordered-int keys of nanmin / nanmax of the slice
};

struct SqParams {
  const float* in;
  float* out;
  uint8_t* out_mask;
  const uint8_t* mask2d;
  const uint8_t* mask_nd;
  const double* rcp_count;   // (ny,nx) 1/count for mask2d launches
  int32_t* status;
  unsigned* edge_flags;      // [(plane*R + m) * T + t]
  SliceStats* stats;
  long long plane_stride, member_stride;
  long long n_planes;
  int n_members, ny, nx, r, nb;
  int flags;
  int W, WS, hl, n_strips, n_ysegs, seg_rows;
  int n_thr;                 // max(n_thresholds, 1)
  double inv_members;
  ThrSet thr;
};

__device__ __forceinline__ int pidx(int c, int cpl_shift) { return c + (c >> cpl_shift); }

__device__ __forceinline__ int float_key(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

// ---------------------------------------------------------------------------
// Main kernel
// ---------------------------------------------------------------------------
template <int VEC, int TM, bool MASKND>
__global__ void __launch_bounds__(256)
square_kernel(const SqParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TX = blockDim.x;
  const int nwarps = TX >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int WS = p.WS, nb = p.nb, r = p.r;
  const int CPL = WS >> 5;                 // columns per lane in the horizontal pass (power of 2)
  const int CPS = 31 - __clz(CPL);         // log2(CPL)
  const int WSP = WS + 32;                 // padded row length (one pad per lane group)

  // shared memory carve-up
  double* vbuf = reinterpret_cast<double*>(smem_raw);                 // [nwarps][WSP]
  double* rtab = vbuf + (size_t)nwarps * WSP;                         // [nb + 1]
  float* ring = reinterpret_cast<float*>(rtab + ((nb + 2) & ~1));     // [nb][WS]
  float* cring = ring + (size_t)nb * WS;                              // MASKND: [nb][WS]
  float* cbuf = cring + (MASKND ? (size_t)nb * WS : 0);               // MASKND: [nwarps][WSP]

  // decode block -> (plane, t, yseg, strip)
  long long bid = blockIdx.x;
  const int strip = (int)(bid % p.n_strips); bid /= p.n_strips;
  const int yseg = (int)(bid % p.n_ysegs); bid /= p.n_ysegs;
  const int t = (int)(bid % p.n_thr); bid /= p.n_thr;
  const long long plane = bid;
  const ThrDev thr = p.thr.t[TM ? t : 0];

  const int ya = yseg * p.seg_rows;
  const int yb = min(p.ny, ya + p.seg_rows);
  const int x0 = strip * p.W;
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;

  // this thread's VEC columns
  const int c0 = threadIdx.x * VEC;             // strip-local
  int gx = x0 - p.hl + c0;                      // global (may be out of range)
  bool col_active = (gx >= 0) && (gx < p.nx);
  if (wrap) {
    gx %= p.nx; if (gx < 0) gx += p.nx;
    col_active = true;
  }

  for (int k = threadIdx.x + 1; k <= nb; k += TX) rtab[k] = 1.0 / (double)k;
  for (int s = 0; s < nb; ++s) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      ring[s * WS + c0 + v] = 0.f;
      if (MASKND) cring[s * WS + c0 + v] = 0.f;
    }
  }
  __syncthreads();

  double V[VEC];
  float CV[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { V[v] = 0.0; CV[v] = 0.f; }

  // quirk detection state (only needed when a mean is produced)
  unsigned long long ne_col = 0, ne_top = 0, ne_bot = 0;
  float nanacc = 0.f;

  const float* in_plane = p.in + plane * p.plane_stride;
  const uint8_t* mnd_plane = MASKND ? p.mask_nd + plane * p.plane_stride : nullptr;
  const int R = p.n_members;

  int slot = 0;
  int brow = 0;
  int batch_y0 = ya;
  for (int yy = ya - r; yy < yb + r; ++yy) {
    float P[VEC];
    float PC[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { P[v] = 0.f; PC[v] = 0.f; }
    const bool row_in = (yy >= 0) && (yy < p.ny);
    if (row_in && col_active) {
      const long long off = (long long)yy * p.nx + gx;
      uint8_t m2[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) m2[v] = 1;
      if (p.mask2d) VecLoad<VEC>::ldm(p.mask2d + off, m2);
      unsigned long long ne_row = 0;
      constexpr int MB = 9;                  // member loads kept in flight together
      for (int mb = 0; mb < R; mb += MB) {
        float d[MB][VEC];
#pragma unroll
        for (int j = 0; j < MB; ++j) {
          if (mb + j < R) VecLoad<VEC>::ld(in_plane + (long long)(mb + j) * p.member_stride + off, d[j]);
        }
#pragma unroll
        for (int j = 0; j < MB; ++j) {
          if (mb + j < R) {
            uint8_t mn[VEC];
            if (MASKND) VecLoad<VEC>::ldm(mnd_plane + off, mn);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              bool valid = m2[v] != 0;
              bool seen = true;
              if (MASKND) { seen = mn[v] == 0; valid = valid && seen; }
              if (seen) nanacc = max_nan(nanacc, fabsf(d[j][v]));
              float tv = TM ? truth_value(d[j][v], thr) : d[j][v];
              tv = valid ? tv : 0.f;
              P[v] += tv;
              if (MASKND) PC[v] = valid ? 1.f : 0.f;
              if (tv != 1.0f) ne_row |= 1ull << (mb + j);
            }
          }
        }
      }
      ne_col |= ne_row;
      if (yy <= r) ne_top |= ne_row;
      if (yy >= p.ny - 1 - r) ne_bot |= ne_row;
    }
    // vertical running sums through the ring (thread-private columns)
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int ri = slot * WS + c0 + v;
      const float old = ring[ri];
      ring[ri] = P[v];
      V[v] += (double)P[v] - (double)old;
      if (MASKND) {
        const float oc = cring[ri];
        cring[ri] = PC[v];
        CV[v] += PC[v] - oc;
      }
    }
    slot = (slot + 1 == nb) ? 0 : slot + 1;

    const int y = yy - r;                       // output row completed by this input row
    if (y >= ya) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        vbuf[brow * WSP + pidx(c0 + v, CPS)] = V[v];
        if (MASKND) cbuf[brow * WSP + pidx(c0 + v, CPS)] = CV[v];
      }
      ++brow;
      if (brow == nwarps || y == yb - 1) {
        __syncthreads();
        if (warp < brow) {
          // ---- horizontal pass: one warp per output row ----
          const int yo = batch_y0 + warp;
          double* row = vbuf + warp * WSP;
          float* crow = cbuf + warp * WSP;
          // lane-local inclusive prefix over CPL contiguous columns
          const int base = lane * CPL + lane;   // pidx(lane*CPL)
          double acc = 0.0;
          float cacc = 0.f;
          for (int j = 0; j < CPL; ++j) {
            acc += row[base + j];
            row[base + j] = acc;
            if (MASKND) { cacc += crow[base + j]; crow[base + j] = cacc; }
          }
          // exclusive scan of lane totals
          double incl = acc;
          float cincl = cacc;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            double o = __shfl_up_sync(0xffffffffu, incl, d);
            float oc = __shfl_up_sync(0xffffffffu, cincl, d);
            if (lane >= d) { incl += o; cincl += oc; }
          }
          const double lane_off = incl - acc;
          const float clane_off = cincl - cacc;
          for (int j = 0; j < CPL; ++j) {
            row[base + j] += lane_off;
            if (MASKND) crow[base + j] += clane_off;
          }
          __syncwarp();
          // rows/cols of the window inside the domain (zero padding: nbhood.py:395-397)
          const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
          const double rrow = rtab[rows_in] * p.inv_members;
          const long long obase = ((plane * p.n_thr + t) * p.ny + yo) * (long long)p.nx;
          for (int c = p.hl + lane * VEC; c < p.hl + p.W; c += 32 * VEC) {
            const int xg = x0 + (c - p.hl);
            if (xg >= p.nx) break;
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const int cc = c + v;
              const int xo = xg + v;
              const int hi = cc + r, lo = cc - r - 1;
              double S = row[pidx(hi, CPS)] - (lo >= 0 ? row[pidx(lo, CPS)] : 0.0);
              if (MASKND) {
                float C = crow[pidx(hi, CPS)] - (lo >= 0 ? crow[pidx(lo, CPS)] : 0.f);
                o[v] = sum_only ? (float)S
                                : (C == 0.f ? __int_as_float(0x7fc00000) : (float)(S / (double)C));
              } else if (sum_only) {
                o[v] = (float)(S * p.inv_members);
              } else if (p.mask2d) {
                o[v] = (float)(S * p.rcp_count[(long long)yo * p.nx + xo] * p.inv_members);
              } else {
                const int cols_in = wrap ? nb : (min(xo + r, p.nx - 1) - max(xo - r, 0) + 1);
                o[v] = (float)(S * rrow * rtab[cols_in]);
              }
            }
            VecLoad<VEC>::st(p.out + obase + xg, o);
            if (p.out_mask) {
#pragma unroll
              for (int v = 0; v < VEC; ++v) {
                const long long pt = (long long)yo * p.nx + xg + v;
                bool masked = p.mask2d ? (p.mask2d[pt] == 0) : false;
                if (MASKND) masked = masked || (mnd_plane[pt] != 0);
                p.out_mask[obase + xg + v] = masked ? 1 : 0;
              }
            }
          }
        }
        __syncthreads();
        batch_y0 += brow;
        brow = 0;
      }
    }
  }

  // ---- epilogue: NaN flag and edge-band flags for the trimming fix-up ----
  if (__isnanf(nanacc)) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  if (!sum_only) {
    // only VEC groups fully inside the bands count (under-approximation => conservative)
    const bool in_dom = col_active && !wrap && (x0 - p.hl + c0 >= 0);
    const int gx_raw = x0 - p.hl + c0;
    const bool left = in_dom && (gx_raw + VEC - 1 <= r);
    const bool right = in_dom && (gx_raw >= p.nx - 1 - r) && (gx_raw + VEC - 1 < p.nx);
    unsigned long long f_left = left ? ne_col : 0ull, f_right = right ? ne_col : 0ull;
    if (wrap) { f_left = ne_col; f_right = ne_col; }
    unsigned long long masks[4] = {ne_top, ne_bot, f_left, f_right};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)masks[k]);
      unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(masks[k] >> 32));
      masks[k] = ((unsigned long long)hi32 << 32) | lo32;
    }
    if (lane == 0) {
      for (int m = 0; m < R; ++m) {
        unsigned bits = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) bits |= (unsigned)((masks[k] >> m) & 1ull) << k;
        if (bits) atomicOr(p.edge_flags + ((plane * R + m) * p.n_thr + t), bits);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Trimming fix-up (improver/nbhood/nbhood.py:341-409).  The reference trims
// all-zero / all-one margins before the box-sum.  Trimming to the non-zero box
// never changes the result; trimming to the non-one box does near the domain
// edge, because the summed-area table is then padded with ONES, so window
// cells outside the domain are counted.  The main kernel
// flags slices whose four edge bands are not all known to contain a non-one;
// for those (rare) slices the exact bounding boxes and min/max are built here
// and the affected frame points are recomputed.
// ---------------------------------------------------------------------------
template <int TM>
__device__ __forceinline__ float load_value(const SqParams& p, long long plane, int m, const ThrDev& thr,
                                            long long pt, bool* valid, bool* seen, float* raw_tv) {
  const long long off = plane * p.plane_stride + (long long)m * p.member_stride + pt;
  float d = p.in[off];
  bool ok = p.mask2d ? (p.mask2d[pt] != 0) : true;
  bool s = true;
  if (p.mask_nd) { s = p.mask_nd[off] == 0; ok = ok && s; }
  float tv = TM ? truth_value(d, thr) : d;
  *valid = ok; *seen = s; *raw_tv = tv;
  return ok ? tv : 0.f;
}

template <int TM>
__global__ void __launch_bounds__(256)
square_stats_kernel(const SqParams p, int chunks) {
  long long bid = blockIdx.x;
  const int chunk = (int)(bid % chunks); bid /= chunks;
  const int t = (int)(bid % p.n_thr); bid /= p.n_thr;
  const int m = (int)(bid % p.n_members);
  const long long plane = bid / p.n_members;
  const long long sidx = (plane * p.n_members + m) * p.n_thr + t;
  if (p.edge_flags[sidx] == 0xFu) return;
  const ThrDev thr = p.thr.t[TM ? t : 0];
  const long long npts = (long long)p.ny * p.nx;
  int b[8] = {INT_MAX, -1, INT_MAX, -1, INT_MAX, -1, INT_MAX, -1};
  int kmin = INT_MAX, kmax = INT_MIN;
  for (long long pt = (long long)chunk * blockDim.x + threadIdx.x; pt < npts;
       pt += (long long)chunks * blockDim.x) {
    bool valid, seen; float tv;
    float v = load_value<TM>(p, plane, m, thr, pt, &valid, &seen, &tv);
    const int y = (int)(pt / p.nx), x = (int)(pt - (long long)y * p.nx);
    if (v != 0.f) { b[0] = min(b[0], y); b[1] = max(b[1], y); b[2] = min(b[2], x); b[3] = max(b[3], x); }
    if (v != 1.f) { b[4] = min(b[4], y); b[5] = max(b[5], y); b[6] = min(b[6], x); b[7] = max(b[7], x); }
    if (seen && !__isnanf(tv)) { int k = float_key(tv); kmin = min(kmin, k); kmax = max(kmax, k); }
  }
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    b[k] = __reduce_min_sync(0xffffffffu, b[k]);
    b[k + 1] = __reduce_max_sync(0xffffffffu, b[k + 1]);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if ((threadIdx.x & 31) == 0) {
    SliceStats* s = p.stats + sidx;
    atomicMin(&s->y0min, b[0]); atomicMax(&s->y0max, b[1]);
    atomicMin(&s->x0min, b[2]); atomicMax(&s->x0max, b[3]);
    atomicMin(&s->y1min, b[4]); atomicMax(&s->y1max, b[5]);
    atomicMin(&s->x1min, b[6]); atomicMax(&s->x1max, b[7]);
    atomicMin(&s->vmin_key, kmin); atomicMax(&s->vmax_key, kmax);
  }
}

__device__ __forceinline__ long long grown_area(int ymin, int ymax, int xmin, int xmax, int h, int ny,
                                                int nx, int* ya, int* yb, int* xa, int* xb) {
  if (ymax < 0) { *ya = *yb = *xa = *xb = 0; return 0; }
  *ya = max(0, ymin - h); *yb = min(ny, ymax + 1 + h);
  *xa = max(0, xmin - h); *xb = min(nx, xmax + 1 + h);
  return (long long)(*yb - *ya) * (*xb - *xa);
}

template <int TM>
__global__ void __launch_bounds__(256)
square_fixup_kernel(const SqParams p, int chunks) {
  long long bid = blockIdx.x;
  const int chunk = (int)(bid % chunks); bid /= chunks;
  const int t = (int)(bid % p.n_thr); bid /= p.n_thr;
  const int m = (int)(bid % p.n_members);
  const long long plane = bid / p.n_members;
  const long long sidx = (plane * p.n_members + m) * p.n_thr + t;
  if (p.edge_flags[sidx] == 0xFu) return;
  const SliceStats s = p.stats[sidx];
  const int h = p.r;
  long long size = (long long)p.ny * p.nx;
  int e = 0, ya, yb, xa, xb, y1a, y1b, x1a, x1b;
  long long a0 = grown_area(s.y0min, s.y0max, s.x0min, s.x0max, h, p.ny, p.nx, &ya, &yb, &xa, &xb);
  if (a0 < size) { size = a0; e = 0; }
  long long a1 = grown_area(s.y1min, s.y1max, s.x1min, s.x1max, h, p.ny, p.nx, &y1a, &y1b, &x1a, &x1b);
  if (a1 < size) { size = a1; e = 1; }
  if (e != 1 || size == 0) return;
  if (chunk == 0 && threadIdx.x == 0) atomicAdd(p.status + 1, 1);
  const ThrDev thr = p.thr.t[TM ? t : 0];
  const float vmin = key_float(s.vmin_key), vmax = key_float(s.vmax_key);
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const int bw = x1b - x1a;
  const long long nbox = (long long)(y1b - y1a) * bw;
  for (long long q = (long long)chunk * blockDim.x + threadIdx.x; q < nbox;
       q += (long long)chunks * blockDim.x) {
    const int y = y1a + (int)(q / bw), x = x1a + (int)(q % bw);
    const int wy0 = max(y - h, 0), wy1 = min(y + h, p.ny - 1);
    const int rows_in = wy1 - wy0 + 1;
    const int cols_in = wrap ? p.nb : (min(x + h, p.nx - 1) - max(x - h, 0) + 1);
    const int n_out = p.nb * p.nb - rows_in * cols_in;
    if (n_out == 0) continue;
    double S = 0.0; int C = 0;
    for (int wy = wy0; wy <= wy1; ++wy) {
      for (int dx = -h; dx <= h; ++dx) {
        int wx = x + dx;
        if (wrap) { wx %= p.nx; if (wx < 0) wx += p.nx; }
        else if (wx < 0 || wx >= p.nx) continue;
        bool valid, seen; float tv;
        float v = load_value<TM>(p, plane, m, thr, (long long)wy * p.nx + wx, &valid, &seen, &tv);
        S += (double)v; C += valid ? 1 : 0;
      }
    }
    if (C == 0) continue;
    double fixed = (S + (double)n_out) / (double)C;
    fixed = fmin(fmax(fixed, (double)vmin), (double)vmax);      // nbhood.py:312
    const long long o = ((plane * p.n_thr + t) * p.ny + y) * (long long)p.nx + x;
    if (p.n_members == 1) {
      p.out[o] = (float)fixed;
    } else {
      const float old = (float)(S / (double)C);
      atomicAdd(p.out + o, (float)(((double)(float)fixed - (double)old) * p.inv_members));
    }
  }
}

__global__ void init_stats_kernel(unsigned* flags, SliceStats* stats, long long n, int sum_only) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = sum_only ? 0xFu : 0u;
  SliceStats s;
  s.y0min = s.x0min = s.y1min = s.x1min = INT_MAX;
  s.y0max = s.x0max = s.y1max = s.x1max = -1;
  s.vmin_key = INT_MAX; s.vmax_key = INT_MIN;
  stats[i] = s;
}

// ---------------------------------------------------------------------------
// Valid-cell count plane for an external 2-D mask (the reference recomputes
// this for every slice, nbhood.py:300; it only depends on the mask).
// ---------------------------------------------------------------------------
__global__ void count_rows_kernel(const uint8_t* mask, uint16_t* tmp, int ny, int nx, int r, int wrap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const int y = (int)(i / nx), x = (int)(i % nx);
  int c = 0;
  for (int dx = -r; dx <= r; ++dx) {
    int xx = x + dx;
    if (wrap) { xx %= nx; if (xx < 0) xx += nx; }
    else if (xx < 0 || xx >= nx) continue;
    c += mask[(long long)y * nx + xx] != 0;
  }
  tmp[i] = (uint16_t)c;
}
__global__ void count_cols_kernel(const uint16_t* tmp, double* rcp, int ny, int nx, int r) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const int y = (int)(i / nx), x = (int)(i % nx);
  int c = 0;
  for (int yy = max(0, y - r); yy <= min(ny - 1, y + r); ++yy) c += tmp[(long long)yy * nx + x];
  rcp[i] = 1.0 / (double)c;       // +inf where no valid cell: 0 * inf = NaN (nbhood.py:309-311)
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
struct SqPlan {
  int vec, tx, WS, hl, W, n_strips, n_ysegs, seg_rows;
  size_t smem;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t sq_smem_bytes(int tx, int vec, int nb, bool masknd) {
  const int WS = tx * vec, WSP = WS + 32, nwarps = tx / 32;
  size_t b = (size_t)nwarps * WSP * 8 + (size_t)((nb + 2) & ~1) * 8 + (size_t)nb * WS * 4;
  if (masknd) b += (size_t)nb * WS * 4 + (size_t)nwarps * WSP * 4;
  return b;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

static int plan_square(const NbhDesc& d, bool masknd, SqPlan* pl) {
  const int r = d.cells, nb = 2 * r + 1;
  int vec = (d.nx % 4 == 0) ? 4 : (d.nx % 2 == 0 ? 2 : 1);
  vec = min(vec, env_int("NBH_SQ_VEC", 4));
  int dev = 0;
  cudaGetDevice(&dev);
  int smem_max = 0;
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  int tx = env_int("NBH_SQ_TX", 128);
  for (;; ) {
    // widen the strip until the halo is at most ~1/3 of it, shrink to fit smem
    while (tx < 256 && tx * vec < 6 * r) tx += 32;
    while (tx > 32 && sq_smem_bytes(tx, vec, nb, masknd) > (size_t)smem_max) tx -= 32;
    if (sq_smem_bytes(tx, vec, nb, masknd) <= (size_t)smem_max && tx * vec > 2 * r + 2 * vec) break;
    if (vec > 1) { vec /= 2; tx = 128; continue; }
    set_error("square neighbourhood radius of %d cells does not fit in shared memory", r);
    return 1;
  }
  pl->vec = vec; pl->tx = tx; pl->WS = tx * vec;
  pl->hl = (r + vec - 1) / vec * vec;
  int wmax = (pl->WS - pl->hl - r) / vec * vec;
  pl->n_strips = (d.nx + wmax - 1) / wmax;
  // balance strips, keep W a multiple of VEC
  int w = (d.nx + pl->n_strips - 1) / pl->n_strips;
  w = (w + vec - 1) / vec * vec;
  pl->W = min(w, wmax);
  const long long tiles_xy = (long long)pl->n_strips * d.n_planes * max(d.n_thresholds, 1);
  int ysegs = env_int("NBH_SQ_YSEGS", 0);
  if (ysegs <= 0) {
    // enough blocks for ~4 waves of 148 SMs x resident blocks, segments no shorter than 8 halos
    const long long want = 4LL * sm_count() * 3;
    ysegs = (int)max(1LL, min((long long)d.ny / max(16, 8 * r), (want + tiles_xy - 1) / tiles_xy));
  }
  ysegs = max(1, min(ysegs, d.ny));
  pl->seg_rows = (d.ny + ysegs - 1) / ysegs;
  pl->n_ysegs = (d.ny + pl->seg_rows - 1) / pl->seg_rows;
  pl->smem = sq_smem_bytes(tx, vec, nb, masknd);
  return 0;
}

template <int VEC, int TM, bool MASKND>
static int launch_square_t(const SqParams& p, const SqPlan& pl, long long blocks, cudaStream_t st) {
  auto kern = square_kernel<VEC, TM, MASKND>;
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  kern<<<(unsigned)blocks, pl.tx, pl.smem, st>>>(p);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int VEC, int TM>
static int launch_square_m(const SqParams& p, const SqPlan& pl, long long blocks, bool masknd, cudaStream_t st) {
  return masknd ? launch_square_t<VEC, TM, true>(p, pl, blocks, st)
                : launch_square_t<VEC, TM, false>(p, pl, blocks, st);
}

template <int VEC>
static int launch_square_v(const SqParams& p, const SqPlan& pl, long long blocks, int tm, bool masknd,
                           cudaStream_t st) {
  switch (tm) {
    case 0: return launch_square_m<VEC, 0>(p, pl, blocks, masknd, st);
    case 1: return launch_square_m<VEC, 1>(p, pl, blocks, masknd, st);
    default: return launch_square_m<VEC, 2>(p, pl, blocks, masknd, st);
  }
}

struct SqWorkspace {
  unsigned* edge_flags;
  SliceStats* stats;
  double* rcp_count;
  uint16_t* tmp;
  size_t total;
};

static void carve_workspace(const NbhDesc& d, void* ws, SqWorkspace* w) {
  const size_t nslices = (size_t)d.n_planes * d.n_members * max(d.n_thresholds, 1);
  size_t off = 0;
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  w->edge_flags = reinterpret_cast<unsigned*>(base + off); off = align_up(off + nslices * 4, 256);
  w->stats = reinterpret_cast<SliceStats*>(base + off); off = align_up(off + nslices * sizeof(SliceStats), 256);
  w->rcp_count = reinterpret_cast<double*>(base + off); off = align_up(off + (size_t)d.ny * d.nx * 8, 256);
  w->tmp = reinterpret_cast<uint16_t*>(base + off); off = align_up(off + (size_t)d.ny * d.nx * 2, 256);
  w->total = off;
}

size_t square_workspace_bytes(const NbhDesc& d) {
  SqWorkspace w;
  carve_workspace(d, nullptr, &w);
  return w.total;
}

int validate_desc(const NbhDesc* d) {
  NBH_REQUIRE(d != nullptr, "null descriptor");
  NBH_REQUIRE(d->ny > 0 && d->nx > 0 && d->n_planes > 0, "empty grid (%d x %d, %lld planes)", d->ny,
              d->nx, (long long)d->n_planes);
  NBH_REQUIRE(d->cells >= 1, "radius of %d grid cells is not positive", d->cells);
  NBH_REQUIRE(d->n_members >= 1 && d->n_members <= 64, "n_members must be in [1, 64], got %d",
              d->n_members);
  NBH_REQUIRE(d->n_thresholds >= 0 && d->n_thresholds <= kMaxThresholds,
              "n_thresholds must be in [0, %d], got %d", kMaxThresholds, d->n_thresholds);
  NBH_REQUIRE(d->n_thresholds == 0 || d->thresholds != nullptr, "thresholds pointer is null");
  NBH_REQUIRE(!(d->mask_nd && d->n_members != 1), "mask_nd requires n_members == 1");
  return 0;
}

int fill_thresholds(const NbhDesc& d, ThrSet* ts, int* tm) {
  ts->n = d.n_thresholds;
  *tm = 0;
  for (int i = 0; i < d.n_thresholds; ++i) {
    if (make_thr_dev(d.thresholds[i], &ts->t[i], d.mask_nd != nullptr)) return 1;
    const int m = ts->t[i].mode == 3 ? 2 : ts->t[i].mode;
    if (i == 0) *tm = m;
    NBH_REQUIRE(m == *tm, "fuzzy and deterministic thresholds cannot be mixed in one launch");
  }
  return 0;
}

int run_square(const NbhDesc* d, const float* in, float* out, uint8_t* out_mask, int32_t* status,
               void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (validate_desc(d)) return 1;
  NBH_REQUIRE(in && out && status && workspace, "null pointer argument");
  NBH_REQUIRE(workspace_bytes >= square_workspace_bytes(*d), "workspace too small");
  const bool masknd = d->mask_nd != nullptr;
  SqPlan pl;
  if (plan_square(*d, masknd, &pl)) return 1;
  SqWorkspace w;
  carve_workspace(*d, workspace, &w);

  SqParams p;
  memset(&p, 0, sizeof(p));
  int tm = 0;
  if (fill_thresholds(*d, &p.thr, &tm)) return 1;
  p.in = in; p.out = out; p.out_mask = out_mask;
  p.mask2d = d->mask2d; p.mask_nd = d->mask_nd; p.rcp_count = w.rcp_count;
  p.status = status; p.edge_flags = w.edge_flags; p.stats = w.stats;
  p.plane_stride = d->plane_stride; p.member_stride = d->member_stride;
  p.n_planes = d->n_planes; p.n_members = d->n_members;
  p.ny = d->ny; p.nx = d->nx; p.r = d->cells; p.nb = 2 * d->cells + 1;
  p.flags = d->flags;
  p.W = pl.W; p.WS = pl.WS; p.hl = pl.hl; p.n_strips = pl.n_strips; p.n_ysegs = pl.n_ysegs;
  p.seg_rows = pl.seg_rows;
  p.n_thr = max(d->n_thresholds, 1);
  p.inv_members = 1.0 / (double)d->n_members;
  if (d->flags & NBH_SUM_ONLY) p.inv_members = 1.0;   // a sum is never averaged

  const long long nslices = d->n_planes * d->n_members * p.n_thr;
  const bool sum_only = (d->flags & NBH_SUM_ONLY) != 0;
  NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  init_stats_kernel<<<(unsigned)((nslices + 255) / 256), 256, 0, st>>>(w.edge_flags, w.stats, nslices,
                                                                    sum_only ? 1 : 0);
  if (d->mask2d && !masknd && !sum_only) {
    const long long n = (long long)d->ny * d->nx;
    count_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d->mask2d, w.tmp, d->ny, d->nx,
                                                                 d->cells, (d->flags & NBH_WRAP_X) ? 1 : 0);
    count_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.tmp, w.rcp_count, d->ny, d->nx,
                                                                 d->cells);
  }
  const long long blocks = (long long)pl.n_strips * pl.n_ysegs * p.n_thr * d->n_planes;
  NBH_REQUIRE(blocks < (1LL << 31), "too many blocks (%lld); split the launch", blocks);
  int rc;
  {
    ProfileScope prof(st);
    switch (pl.vec) {
      case 4: rc = launch_square_v<4>(p, pl, blocks, tm, masknd, st); break;
      case 2: rc = launch_square_v<2>(p, pl, blocks, tm, masknd, st); break;
      default: rc = launch_square_v<1>(p, pl, blocks, tm, masknd, st); break;
    }
  }
  if (rc) return rc;
  if (!sum_only) {
    const int chunks = 32;
    const long long fb = nslices * chunks;
    NBH_REQUIRE(fb < (1LL << 31), "too many fix-up blocks");
    switch (tm) {
      case 0:
        square_stats_kernel<0><<<(unsigned)fb, 256, 0, st>>>(p, chunks);
        square_fixup_kernel<0><<<(unsigned)fb, 256, 0, st>>>(p, chunks);
        break;
      case 1:
        square_stats_kernel<1><<<(unsigned)fb, 256, 0, st>>>(p, chunks);
        square_fixup_kernel<1><<<(unsigned)fb, 256, 0, st>>>(p, chunks);
        break;
      default:
        square_stats_kernel<2><<<(unsigned)fb, 256, 0, st>>>(p, chunks);
        square_fixup_kernel<2><<<(unsigned)fb, 256, 0, st>>>(p, chunks);
        break;
    }
    NBH_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace nbh
