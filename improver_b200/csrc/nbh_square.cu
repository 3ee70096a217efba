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
#include "nbh_square_rs.cuh"
#include "nbh_square_box.cuh"
#include "nbh_square_mt.cuh"

namespace nbh {

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
__device__ __forceinline__ int float_key(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

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
square_stats_kernel(const SqParams p, int chunks, const int* suspects, const int* n_suspects, long long n_all) {
 // suspects == nullptr: every one of the n_all slices
 const long long n_work = (suspects ? (long long)(*n_suspects) : n_all) * chunks;
 for (long long work = blockIdx.x; work < n_work; work += gridDim.x) {
  const int chunk = (int)(work % chunks);
  const long long sidx = suspects ? suspects[work / chunks] : work / chunks;
  long long bid = sidx;
  const int t = (int)(bid % p.n_thr); bid /= p.n_thr;
  const int m = (int)(bid % p.n_members);
  const long long plane = bid / p.n_members;
  const ThrDev thr = p.thr.t[TM ? t : 0];
  const long long npts = (long long)p.ny * p.nx;
  int b[8] = {INT_MAX, -1, INT_MAX, -1, INT_MAX, -1, INT_MAX, -1};
  int kmin = INT_MAX, kmax = INT_MIN;
  for (long long pt = (long long)chunk * blockDim.x + threadIdx.x; pt < npts;
       pt += (long long)chunks * blockDim.x) {
    bool valid, seen; float tv;
    float v = load_value<TM>(p, plane, m, thr, pt, &valid, &seen, &tv);
    const int yl = (int)(pt / p.nx), x = (int)(pt - (long long)yl * p.nx);
    const int y = yl + p.y_offset;                         // row of the whole grid
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
square_fixup_kernel(const SqParams p, int chunks, const int* suspects, const int* n_suspects, long long n_all,
                    const SliceStats* stats) {
 const long long n_work = (suspects ? (long long)(*n_suspects) : n_all) * chunks;
 for (long long work = blockIdx.x; work < n_work; work += gridDim.x) {
  const int chunk = (int)(work % chunks);
  const long long sidx = suspects ? suspects[work / chunks] : work / chunks;
  long long bid = sidx;
  const int t = (int)(bid % p.n_thr); bid /= p.n_thr;
  const int m = (int)(bid % p.n_members);
  const long long plane = bid / p.n_members;
  const SliceStats s = stats[sidx];
  const int h = p.r;
  const int NYG = p.ny_global;                             // the boxes live in the rows of the whole grid
  long long size = (long long)NYG * p.nx;
  int e = 0, ya, yb, xa, xb, y1a, y1b, x1a, x1b;
  long long a0 = grown_area(s.y0min, s.y0max, s.x0min, s.x0max, h, NYG, p.nx, &ya, &yb, &xa, &xb);
  if (a0 < size) { size = a0; e = 0; }
  long long a1 = grown_area(s.y1min, s.y1max, s.x1min, s.x1max, h, NYG, p.nx, &y1a, &y1b, &x1a, &x1b);
  if (a1 < size) { size = a1; e = 1; }
  if (e != 1 || size == 0) continue;
  if (chunk == 0 && threadIdx.x == 0) atomicAdd(p.status + 1, 1);
  const ThrDev thr = p.thr.t[TM ? t : 0];
  const float vmin = key_float(s.vmin_key), vmax = key_float(s.vmax_key);
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const int bw = x1b - x1a;
  // rows of the box this band owns (a y-band leaves its halo rows to the neighbouring band)
  y1a = max(y1a, p.y_offset + p.halo_top);
  y1b = min(y1b, p.y_offset + p.ny - p.halo_bottom);
  if (y1b <= y1a) continue;
  const long long nbox = (long long)(y1b - y1a) * bw;
  for (long long q = (long long)chunk * blockDim.x + threadIdx.x; q < nbox;
       q += (long long)chunks * blockDim.x) {
    const int yg = y1a + (int)(q / bw), x = x1a + (int)(q % bw);
    const int y = yg - p.y_offset;                         // local row
    const int wy0 = max(yg - h, 0) - p.y_offset, wy1 = min(yg + h, NYG - 1) - p.y_offset;
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
}

// Slices whose four edge bands are not all known to contain a value != 1.
__global__ void collect_suspects_kernel(const unsigned* flags, long long n, int* suspects, int* n_suspects) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i] != 0xFu) suspects[atomicAdd(n_suspects, 1)] = (int)i;
}

__global__ void init_stats_kernel(unsigned* flags, SliceStats* stats, long long n, unsigned init_bits,
                                  int* n_suspects) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && n_suspects) *n_suspects = 0;
  if (i >= n) return;
  if (flags) flags[i] = init_bits;
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
__global__ void count_cols_kernel(const uint16_t* tmp, double* rcp, float* cnt_f, int ny, int nx, int r) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const int y = (int)(i / nx), x = (int)(i % nx);
  int c = 0;
  for (int yy = max(0, y - r); yy <= min(ny - 1, y + r); ++yy) c += tmp[(long long)yy * nx + x];
  rcp[i] = 1.0 / (double)c;       // +inf where no valid cell: 0 * inf = NaN (nbhood.py:309-311)
  cnt_f[i] = (float)c;            // the count itself, for the float32 quotient of the tile kernel
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t sq_smem_bytes(int tx, int vec, int nb, bool masknd) {
  const int WS = tx * vec, WSP = WS + 32, nwarps = tx / 32;
  size_t b = (size_t)nwarps * WSP * 8 + (size_t)((nb + 2) & ~1) * 8 + (size_t)nb * WS * 4;
  if (masknd) b += (size_t)nb * WS * 4 + (size_t)nwarps * WSP * 4;
  return b;
}

// Choose the strip width (threads per block), the y segmentation and the load
// grouping.  Cost model: issue slots wasted on idle lanes, halo rows re-read
// from HBM by every y segment, and the partially filled last wave.
static int plan_square(const NbhDesc& d, bool masknd, SqPlan* pl) {
  const int r = d.cells, nb = 2 * r + 1;
  int vec = (d.nx % 4 == 0) ? 4 : (d.nx % 2 == 0 ? 2 : 1);
  if (r < 3) vec = min(vec, 2);          // edge-band flags need whole VEC groups inside r+1 columns
  if (r < 1) vec = 1;
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  const long long units = d.n_planes * (long long)max(d.n_thresholds, 1);
  double best_cost = 1e30;
  bool found = false;
  for (;;) {
    for (int tx = 64; tx <= kSqMaxThreads; tx += 32) {
      const size_t smem = sq_smem_bytes(tx, vec, nb, masknd);
      if (smem > (size_t)smem_max) continue;
      const int WS = tx * vec;
      const int hl = (r + vec - 1) / vec * vec;
      const int wmax = (WS - hl - r) / vec * vec;
      if (wmax < vec) continue;
      const int n_strips = (d.nx + wmax - 1) / wmax;
      int W = ((d.nx + n_strips - 1) / n_strips + vec - 1) / vec * vec;
      W = min(W, wmax);
      const double lane_eff = (double)d.nx / ((double)n_strips * WS);
      int per_sm = min(16, min((int)(65536 / (tx * 170)), (int)((size_t)smem_max / (smem + 1024))));
      per_sm = max(per_sm, 1);
      const double slots = (double)sm_count() * per_sm;
      for (int ysegs = 1; ysegs <= 64; ++ysegs) {
        const int seg_rows = (d.ny + ysegs - 1) / ysegs;
        if (ysegs > 1 && seg_rows < 4 * r) break;
        const int n_ysegs = (d.ny + seg_rows - 1) / seg_rows;
        if (n_ysegs != ysegs) continue;
        const double blocks = (double)units * n_strips * n_ysegs;
        const double waves = blocks / slots;
        const double tail = ceil(waves) / waves;
        const double halo = 1.0 + (n_ysegs > 1 ? 2.0 * r / seg_rows : 0.0);
        // idle lanes only cost issue slots; weigh them less than HBM re-reads
        const double cost = tail * halo * (1.0 + 0.5 * (1.0 / lane_eff - 1.0)) *
                            (1.0 + 0.25 * (double)(hl + r) / W);
        if (cost < best_cost) {
          best_cost = cost; found = true;
          pl->vec = vec; pl->tx = tx; pl->WS = WS; pl->hl = hl; pl->W = W; pl->n_strips = n_strips;
          pl->n_ysegs = n_ysegs; pl->seg_rows = seg_rows; pl->smem = smem;
        }
      }
    }
    if (found) break;
    if (vec > 1) { vec /= 2; continue; }
    set_error("square neighbourhood radius of %d cells does not fit in shared memory", r);
    return 1;
  }
  const int mbm = vec == 4 ? 9 : 18;
  if (masknd || d.n_members == 1) { pl->mode = SQ_ROWS; pl->mb = vec == 4 ? 4 : 8; }
  else if (d.n_members == mbm) { pl->mode = SQ_MEMBERS_FULL; pl->mb = mbm; }
  else if (d.n_members < mbm) { pl->mode = SQ_MEMBERS_PART; pl->mb = mbm; }
  else { pl->mode = SQ_CURSOR; pl->mb = mbm; }
  return 0;
}

// Geometry of the warp-autonomous kernel; returns false when the radius is
// too large for a 32*VEC-column strip (the block-cooperative kernel is used).
static bool plan_square_warp(const NbhDesc& d, int vec, SqWarpGeom* g) {
  const int r = d.cells;
  const bool wrap = (d.flags & NBH_WRAP_X) != 0;
  // one or two VEC groups per lane: pick the strip width with the fewest idle lanes
  // (two groups per lane waste fewer lanes but need > 128 registers, which costs
  // more in occupancy than it gains on B200: opt-in only)
  double best = 0.0;
  for (int groups = 1; groups <= 1; ++groups) {
    const int WS = 32 * vec * groups;
    const int hl = (r + vec - 1) / vec * vec;
    const int W = (WS - hl - r) / vec * vec;
    const int W0 = wrap ? W : (WS - r) / vec * vec;
    if (W < (WS * 6) / 10) continue;
    if (sq_warp_smem_bytes(vec, groups, 2 * r + 1) * 3 > 200 * 1024) continue;
    const int n_strips = wrap ? (d.nx + W - 1) / W : 1 + (d.nx > W0 ? (d.nx - W0 + W - 1) / W : 0);
    const double eff = (double)d.nx / ((double)n_strips * WS);
    if (eff > best + 0.02) {
      best = eff;
      g->WS = WS; g->hl = hl; g->W = W; g->W0 = W0; g->n_strips = n_strips;
    }
  }
  if (best == 0.0) return false;
  const long long units = d.n_planes * (long long)max(d.n_thresholds, 1) * g->n_strips;
  int ysegs = 0;
  {
    // Warps are scheduled dynamically as blocks retire, so many small tasks
    // balance better than few large ones; every y segment re-reads 2r halo rows
    // (mostly L2 hits thanks to the alternating march direction).  Measured on
    // B200 (profiles/): ~5 tasks per resident warp slot is the sweet spot.
    const double slots = (double)sm_count() * 16;
    const int max_segs = max(1, d.ny / max(32, 6 * r));
    ysegs = (int)min((double)max_segs, max(1.0, ceil(5.0 * slots / (double)units)));
  }
  ysegs = max(1, min(ysegs, d.ny));
  g->seg_rows = (d.ny + ysegs - 1) / ysegs;
  g->n_ysegs = (d.ny + g->seg_rows - 1) / g->seg_rows;
  g->n_tasks = units * g->n_ysegs;
  g->n_sm = 0;
  return true;
}

bool plan_square_rs(const NbhDesc& d, const float* in, SqRsGeom* g, int* grid);

struct SqWorkspace {
  unsigned* edge_flags;
  SliceStats* stats;
  int* suspects;
  int* n_suspects;
  double* rcp_count;
  float* cnt_plane;
  uint16_t* tmp;
  void* box;
  unsigned* vh;
  size_t total;
};

bool plan_square_vh(const NbhDesc& d, const float* in, int* fx_shift);
size_t vh_workspace_bytes(const NbhDesc& d);
int launch_square_vh(const SqParams& sp, const NbhDesc& d, int tm, int fx_shift, BoxStats* stats, unsigned* vplane,
                     cudaStream_t st);

static void carve_workspace(const NbhDesc& d, void* ws, SqWorkspace* w) {
  const size_t nslices = (size_t)d.n_planes * d.n_members * max(d.n_thresholds, 1);
  size_t off = 0;
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  w->edge_flags = reinterpret_cast<unsigned*>(base + off); off = align_up(off + nslices * 4, 256);
  w->stats = reinterpret_cast<SliceStats*>(base + off); off = align_up(off + nslices * sizeof(SliceStats), 256);
  w->suspects = reinterpret_cast<int*>(base + off); off = align_up(off + nslices * 4, 256);
  w->n_suspects = reinterpret_cast<int*>(base + off); off = align_up(off + 4, 256);
  w->rcp_count = reinterpret_cast<double*>(base + off); off = align_up(off + (size_t)d.ny * d.nx * 8, 256);
  w->tmp = reinterpret_cast<uint16_t*>(base + off); off = align_up(off + (size_t)d.ny * d.nx * 2, 256);
  w->cnt_plane = reinterpret_cast<float*>(base + off); off = align_up(off + (size_t)d.ny * d.nx * 4, 256);
  w->box = base + off;
  if (d.n_members == 1) off = align_up(off + box_workspace_bytes((long long)nslices, d.ny, d.nx), 256);
  w->vh = reinterpret_cast<unsigned*>(base + off);
  // the intermediate plane of the two-pass kernels: needed exactly when they may run, i.e. the launch
  // has their shape but not the box kernel's (both ask for the same pointer alignment, so a shape the
  // box kernel takes never reaches them); decided by the planners themselves on an aligned stand-in
  {
    const float* aligned = reinterpret_cast<const float*>(static_cast<uintptr_t>(256));
    SqBoxGeom bg;
    int grid = 0, shift = 0;
    size_t smem = 0;
    if (!plan_square_box(d, aligned, &bg, &grid, &smem) && plan_square_vh(d, aligned, &shift))
      off = align_up(off + vh_workspace_bytes(d), 256);
  }
  w->total = off;
}

size_t square_workspace_bytes(const NbhDesc& d) {
  SqWorkspace w;
  carve_workspace(d, nullptr, &w);
  return w.total;
}

bool square_native_land_sea(const NbhDesc&) { return false; }

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
  NBH_REQUIRE(!(d->flags & NBH_LAND_SEA) || (d->mask2d && !d->mask_nd),
              "NBH_LAND_SEA needs the land mask as mask2d and no masked-array input");
  if (d->ny_global != 0) {
    NBH_REQUIRE(d->y_offset >= 0 && d->y_offset + d->ny <= d->ny_global,
                "band rows [%d, %d) do not lie inside the grid of %d rows", d->y_offset, d->y_offset + d->ny,
                d->ny_global);
    NBH_REQUIRE(d->halo_top >= 0 && d->halo_bottom >= 0 && d->halo_top + d->halo_bottom < d->ny,
                "halo rows (%d + %d) leave no own rows in a band of %d rows", d->halo_top, d->halo_bottom, d->ny);
    NBH_REQUIRE((d->y_offset == 0 || d->halo_top >= d->cells) &&
                (d->y_offset + d->ny == d->ny_global || d->halo_bottom >= d->cells),
                "a band needs at least `cells` halo rows on every side that is not a domain edge");
  } else {
    NBH_REQUIRE(d->y_offset == 0 && d->halo_top == 0 && d->halo_bottom == 0,
                "y_offset / halo rows given without ny_global");
  }
  return 0;
}

int fill_thresholds(const NbhDesc& d, ThrSet* ts, int* tm) {
  ts->n = d.n_thresholds;
  *tm = 0;
  ts->t[0].one_lo = ts->t[0].one_hi = 1.0f;      // raw data: "is one" <=> d == 1
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
  NBH_REQUIRE(d->ny_global == 0 || d->ext_stats != nullptr || (d->flags & NBH_SUM_ONLY),
              "y-band launches need the statistics of the whole slices (ext_stats)");
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
  p.mask2d = d->mask2d; p.mask_nd = d->mask_nd; p.rcp_count = w.rcp_count; p.cnt_plane = w.cnt_plane;
  p.status = status; p.edge_flags = w.edge_flags; p.stats = w.stats;
  p.plane_stride = d->plane_stride; p.member_stride = d->member_stride;
  p.n_planes = d->n_planes; p.n_members = d->n_members;
  p.ny = d->ny; p.nx = d->nx; p.r = d->cells; p.nb = 2 * d->cells + 1;
  p.flags = d->flags;
  p.W = pl.W; p.WS = pl.WS; p.hl = pl.hl; p.n_strips = pl.n_strips; p.n_ysegs = pl.n_ysegs;
  p.seg_rows = pl.seg_rows;
  p.n_thr = max(d->n_thresholds, 1);
  p.cpl = pl.WS / 32;
  p.cpl_magic = (65536 + p.cpl - 1) / p.cpl;
  p.inv_members = 1.0 / (double)d->n_members;
  if (d->flags & NBH_SUM_ONLY) p.inv_members = 1.0;   // a sum is never averaged
  p.y_offset = d->y_offset; p.ny_global = d->ny_global ? d->ny_global : d->ny;
  p.halo_top = d->halo_top; p.halo_bottom = d->halo_bottom;

  const long long nslices = d->n_planes * d->n_members * p.n_thr;
  const bool sum_only = (d->flags & NBH_SUM_ONLY) != 0;
  NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  const unsigned init_bits = sum_only ? 0xFu : ((d->flags & NBH_WRAP_X) ? 0xCu : 0u);
  NBH_REQUIRE(nslices < (1LL << 31), "too many slices in one launch (%lld)", nslices);
  init_stats_kernel<<<(unsigned)((nslices + 255) / 256), 256, 0, st>>>(w.edge_flags, w.stats, nslices,
                                                                    init_bits, w.n_suspects);
  if (d->mask2d && !masknd && !sum_only) {
    const long long n = (long long)d->ny * d->nx;
    count_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d->mask2d, w.tmp, d->ny, d->nx,
                                                                 d->cells, (d->flags & NBH_WRAP_X) ? 1 : 0);
    count_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.tmp, w.rcp_count, w.cnt_plane, d->ny, d->nx,
                                                                 d->cells);
  }
  const long long blocks = (long long)pl.n_strips * pl.n_ysegs * p.n_thr * d->n_planes;
  NBH_REQUIRE(blocks < (1LL << 31), "too many blocks (%lld); split the launch", blocks);
  int rc;
  SqWarpGeom wg;
  const bool use_warp = !masknd && plan_square_warp(*d, pl.vec, &wg);
  // fuzzy launches: pick the cheapest membership form every threshold of the launch supports
  int tm_warp = tm;
  if (tm == 2 && !masknd) {
    int kind = 5;
    for (int i = 0; i < d->n_thresholds; ++i) kind = min(kind, p.thr.t[i].fast_kind == 2 ? 2 : p.thr.t[i].fast_kind);
    tm_warp = kind;
  }
  // Kernel choice (no run-time knobs):
  //   members (fused threshold -> box sum -> realization mean): square_mt_kernel for two or more
  //   thresholds (one read of the input per four thresholds), square_rs_kernel for one, else the warp kernel;
  //   single-member slices, radius <= 8: square_box_kernel (+ guarded clamp / float64 re-run);
  //   single-member slices, radius 9 .. 63: square_v_kernel + square_h_kernel (same follow-ups);
  //   everything else (periodic x, odd widths, radius > 8): warp kernel; masked arrays and radii
  //   beyond a warp strip: block-cooperative kernel.
  SqBoxGeom bg;
  int box_grid = 0;
  size_t box_smem = 0;
  const bool use_box = !masknd && plan_square_box(*d, in, &bg, &box_grid, &box_smem);
  int vh_shift = 0;
  const bool use_vh = !use_box && !masknd && plan_square_vh(*d, in, &vh_shift);
  SqMtGeom mg;
  int mt_grid = 0;
  size_t mt_smem = 0;
  const bool use_mt = !use_box && !use_vh && !masknd && plan_square_mt(*d, in, &mg, &mt_grid, &mt_smem);
  SqRsGeom rg;
  int rs_grid = 0;
  const bool use_rs = !use_box && !use_vh && !use_mt && use_warp && plan_square_rs(*d, in, &rg, &rs_grid);
  const bool m2d = d->mask2d != nullptr;
  auto launch_exact = [&](const SqParams& pp) -> int {
    if (use_warp) {
      switch (pl.vec) {
        case 4: return launch_square_warp_vec<4>(pp, wg, tm_warp, m2d, st);
        case 2: return launch_square_warp_vec<2>(pp, wg, tm_warp, m2d, st);
        default: return launch_square_warp_vec<1>(pp, wg, tm_warp, m2d, st);
      }
    }
    switch (pl.vec) {
      case 4: return launch_square_vec<4>(pp, pl, blocks, tm, masknd, m2d, st);
      case 2: return launch_square_vec<2>(pp, pl, blocks, tm, masknd, m2d, st);
      default: return launch_square_vec<1>(pp, pl, blocks, tm, masknd, m2d, st);
    }
  };
  if (use_box) {
    BoxWorkspace bw;
    box_carve(w.box, nslices, d->ny, d->nx, &bw);
    if (box_prepare(bw, d->mask2d, nslices, d->ny, d->nx, st)) return 1;
    bg.stats = bw.stats;
    {
      ProfileScope prof(st, "nbh::square_box_kernel (row-streaming, fixed-point, shuffle windows)");
      rc = launch_square_box(p, bg, tm_warp, m2d, box_grid, box_smem, st);
    }
    if (rc) return rc;
    if (box_finish(bw, p, d->ext_stats, tm, nslices, st)) return 1;
    if (tm == 0) {
      // raw slices that are not probabilities in [0, 1]: recomputed in float64 (guarded per slice)
      SqParams pr = p;
      pr.redo = bw.redo;
      rc = launch_exact(pr);
    }
  } else if (use_vh) {
    // single-member slices, radius 9 .. 63: two radius-independent passes (vertical running sums,
    // horizontal prefix differences) over exact fixed point, chunked so the intermediate stays in L2
    BoxWorkspace bw;
    box_carve(w.box, nslices, d->ny, d->nx, &bw);
    if (box_prepare(bw, nullptr, nslices, d->ny, d->nx, st)) return 1;
    {
      ProfileScope prof(st, "nbh::square_v_kernel + nbh::square_h_kernel (two-pass fixed point, L2-resident intermediate)");
      rc = launch_square_vh(p, *d, tm_warp, vh_shift, bw.stats, w.vh, st);
    }
    if (rc) return rc;
    if (box_finish(bw, p, d->ext_stats, tm, nslices, st)) return 1;
    if (tm == 0) {
      SqParams pr = p;
      pr.redo = bw.redo;
      rc = launch_exact(pr);
    }
  } else if (use_mt) {
    // several thresholds: the input is read once per group of up to four thresholds
    if (launch_edge_flags(p, tm, st)) return 1;
    ProfileScope prof(st, "nbh::square_mt_kernel (row-streaming, (window x threshold) consumer warps)");
    rc = 0;
    for (int t0 = 0; t0 < d->n_thresholds && rc == 0; t0 += mg.n_thr) {
      SqMtGeom gg = mg;
      gg.t_base = t0;
      gg.n_thr = min(mg.n_thr, d->n_thresholds - t0);
      rc = launch_square_mt(p, gg, tm_warp, m2d, mt_grid, mt_smem, st);
    }
  } else if (use_rs) {
    ProfileScope prof(st, "nbh::square_rs_kernel (row-streaming, bulk-copy staged rows)");
    rc = launch_square_rs(p, rg, tm_warp, m2d, rs_grid, st);
  } else {
    ProfileScope prof(st, use_warp ? "nbh::square_warp_kernel (warp-autonomous strips)"
                                   : "nbh::square_kernel (block-cooperative strips)");
    rc = launch_exact(p);
  }
  if (rc) return rc;
  if (!sum_only) {
    const int chunks = 32;
    const unsigned fb = (unsigned)min((long long)sm_count() * 8, nslices * chunks);
    if (d->ext_stats) {
      // y-band launch: the bounding boxes and the extremes of the WHOLE slices come from the caller
      // (nbh_slice_stats_f32 on every band + all-reduce); every slice is looked at
      switch (tm) {
        case 0: square_fixup_kernel<0><<<fb, 256, 0, st>>>(p, chunks, nullptr, nullptr, nslices, d->ext_stats); break;
        case 1: square_fixup_kernel<1><<<fb, 256, 0, st>>>(p, chunks, nullptr, nullptr, nslices, d->ext_stats); break;
        default: square_fixup_kernel<2><<<fb, 256, 0, st>>>(p, chunks, nullptr, nullptr, nslices, d->ext_stats); break;
      }
      NBH_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
    // trimming fix-up: nothing to do (three tiny launches) unless a slice is suspicious
    collect_suspects_kernel<<<(unsigned)((nslices + 255) / 256), 256, 0, st>>>(w.edge_flags, nslices,
                                                                            w.suspects, w.n_suspects);
    switch (tm) {
      case 0:
        square_stats_kernel<0><<<fb, 256, 0, st>>>(p, chunks, w.suspects, w.n_suspects, 0);
        square_fixup_kernel<0><<<fb, 256, 0, st>>>(p, chunks, w.suspects, w.n_suspects, 0, w.stats);
        break;
      case 1:
        square_stats_kernel<1><<<fb, 256, 0, st>>>(p, chunks, w.suspects, w.n_suspects, 0);
        square_fixup_kernel<1><<<fb, 256, 0, st>>>(p, chunks, w.suspects, w.n_suspects, 0, w.stats);
        break;
      default:
        square_stats_kernel<2><<<fb, 256, 0, st>>>(p, chunks, w.suspects, w.n_suspects, 0);
        square_fixup_kernel<2><<<fb, 256, 0, st>>>(p, chunks, w.suspects, w.n_suspects, 0, w.stats);
        break;
    }
    NBH_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// nbh_slice_stats_f32: the statistics of every slice of a band, rows in grid coordinates
int run_slice_stats(const NbhDesc* d, const float* in, NbhSliceStats* stats, cudaStream_t st) {
  if (validate_desc(d)) return 1;
  NBH_REQUIRE(in && stats, "null pointer argument");
  SqParams p;
  memset(&p, 0, sizeof(p));
  int tm = 0;
  if (fill_thresholds(*d, &p.thr, &tm)) return 1;
  p.in = in; p.mask2d = d->mask2d; p.mask_nd = d->mask_nd; p.stats = stats;
  p.plane_stride = d->plane_stride; p.member_stride = d->member_stride;
  p.n_planes = d->n_planes; p.n_members = d->n_members;
  p.ny = d->ny; p.nx = d->nx; p.r = d->cells; p.nb = 2 * d->cells + 1; p.flags = d->flags;
  p.n_thr = max(d->n_thresholds, 1);
  p.y_offset = d->y_offset; p.ny_global = d->ny_global ? d->ny_global : d->ny;
  const long long nslices = d->n_planes * d->n_members * p.n_thr;
  NBH_REQUIRE(nslices < (1LL << 31), "too many slices in one launch (%lld)", nslices);
  init_stats_kernel<<<(unsigned)((nslices + 255) / 256), 256, 0, st>>>(nullptr, stats, nslices, 0u, nullptr);
  const int chunks = 32;
  const unsigned fb = (unsigned)min((long long)sm_count() * 8, nslices * chunks);
  switch (tm) {
    case 0: square_stats_kernel<0><<<fb, 256, 0, st>>>(p, chunks, nullptr, nullptr, nslices); break;
    case 1: square_stats_kernel<1><<<fb, 256, 0, st>>>(p, chunks, nullptr, nullptr, nslices); break;
    default: square_stats_kernel<2><<<fb, 256, 0, st>>>(p, chunks, nullptr, nullptr, nslices); break;
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nbh
