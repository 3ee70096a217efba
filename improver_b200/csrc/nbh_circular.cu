// Circular neighbourhood (improver/nbhood/nbhood.py:55-92 kernel +
// scipy.ndimage.correlate(mode="nearest") at nbhood.py:401), optionally fused
// with the fuzzy threshold in front and the realization mean behind.
//
// S(y,x) = sum_{dy,dx} K[dy,dx] * v[clamp(y+dy), clamp(x+dx)]   (float64)
// A(y,x) = float32(sum K * ok[clamped])                          (float32 like the reference)
// out    = float32(S / A), NaN where A == 0, or float32(S) for sum_only.
//
// Unweighted discs are evaluated as a per-row span decomposition over float64
// row prefix sums: a block marches down a column strip keeping a shared-memory
// ring of the last 2r+1 prefix rows (built with a warp-shuffle scan); each
// output needs two ring reads per kernel row.  Weighted discs use the direct
// form from the same ring of raw rows.
#include "nbh_common.cuh"

namespace nbh {

struct CircParams {
  const float* in;
  float* out;
  uint8_t* out_mask;
  const uint8_t* mask2d;
  const uint8_t* mask_nd;
  const float* area;        // (ny,nx) float32 area_sum for mask2d launches
  const double* ktab;       // weighted: K as a function of d^2, [r*r + 1]
  const int* wtab;          // half width per |dy|, [r + 1]
  int32_t* status;
  long long plane_stride, member_stride;
  long long n_planes;
  int n_members, ny, nx, r, nb;
  int flags;
  int W, WS, n_strips, n_ysegs, seg_rows;
  int n_thr;
  double inv_members;
  float area_const;         // unmasked: sum(K) as float32
  ThrSet thr;
};

// One block = one column strip x one y segment of one output plane.  Strip
// column c maps to global column clamp(x0 - r + c) (edge replication).
template <int TM, bool MASKND, bool WEIGHTED>
__global__ void __launch_bounds__(256)
circular_kernel(const CircParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TX = blockDim.x;
  const int WS = p.WS, nb = p.nb, r = p.r;
  const int RS = WS + 1;                         // ring row pitch (exclusive prefix has WS+1 entries)
  double* ring = reinterpret_cast<double*>(smem_raw);                  // [nb][RS] prefix (or raw) of v
  double* cring = ring + (size_t)nb * RS;                              // MASKND: [nb][RS] prefix of ok
  double* wsum = cring + (MASKND ? (size_t)nb * RS : 0);               // [32] warp totals
  int* wtab = reinterpret_cast<int*>(wsum + 64);                       // [r+1]

  long long bid = blockIdx.x;
  const int strip = (int)(bid % p.n_strips); bid /= p.n_strips;
  const int yseg = (int)(bid % p.n_ysegs); bid /= p.n_ysegs;
  const int t = (int)(bid % p.n_thr); bid /= p.n_thr;
  const long long plane = bid;
  const ThrDev thr = p.thr.t[TM ? t : 0];
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;

  const int ya = yseg * p.seg_rows;
  const int yb = min(p.ny, ya + p.seg_rows);
  const int x0 = strip * p.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = TX >> 5;

  for (int k = threadIdx.x; k <= r; k += TX) wtab[k] = p.wtab[k];

  // this thread's column (one column per thread, WS == TX)
  const int c = threadIdx.x;
  const int gx = min(max(x0 - r + c, 0), p.nx - 1);
  const float* in_plane = p.in + plane * p.plane_stride;
  const uint8_t* mnd_plane = MASKND ? p.mask_nd + plane * p.plane_stride : nullptr;
  const int R = p.n_members;
  float nanacc = 0.f;
  __syncthreads();

  // march over input rows yy (clamped for loading); output row y = yy - r is
  // complete once rows y-r .. y+r are in the ring.
  for (int yy = ya - r; yy < yb + r; ++yy) {
    const int ys = min(max(yy, 0), p.ny - 1);
    const long long off = (long long)ys * p.nx + gx;
    const bool m2 = p.mask2d ? (p.mask2d[off] != 0) : true;
    float P = 0.f;
    float okf = 0.f;
    for (int m = 0; m < R; ++m) {
      float d = __ldg(in_plane + (long long)m * p.member_stride + off);
      bool valid = m2, seen = true;
      if (MASKND) { seen = mnd_plane[off] == 0; valid = valid && seen; }
      if (seen) nanacc = max_nan(nanacc, fabsf(d));
      float tv = TM ? truth_value(d, thr) : d;
      P += valid ? tv : 0.f;
      okf = valid ? 1.f : 0.f;
    }
    const int slot = ((yy % nb) + nb) % nb;
    double* row = ring + (size_t)slot * RS;
    double* crow = cring + (size_t)slot * RS;
    if (WEIGHTED) {
      row[c] = (double)P;
      if (MASKND) crow[c] = (double)okf;
      __syncthreads();
    } else {
      // block-wide exclusive prefix of the row (float64, warp-shuffle scan)
      double incl = (double)P, cincl = (double)okf;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_up_sync(0xffffffffu, incl, d);
        double oc = __shfl_up_sync(0xffffffffu, cincl, d);
        if (lane >= d) { incl += o; cincl += oc; }
      }
      if (lane == 31) { wsum[warp] = incl; wsum[32 + warp] = cincl; }
      __syncthreads();
      double woff = 0.0, cwoff = 0.0;
      for (int w = 0; w < warp; ++w) { woff += wsum[w]; cwoff += wsum[32 + w]; }
      row[c + 1] = incl + woff;
      if (MASKND) crow[c + 1] = cincl + cwoff;
      if (c == 0) { row[0] = 0.0; if (MASKND) crow[0] = 0.0; }
      __syncthreads();
    }
    (void)nwarps;

    const int y = yy - r;
    if (y < ya) continue;
    // ---- output row y: thread c produces column x = x0 + (c - r) if it is an output column
    const int xo = x0 + c - r;
    if (c >= r && c < r + p.W && xo < p.nx) {
      double S = 0.0, A = 0.0;
      for (int dy = -r; dy <= r; ++dy) {
        const int w = wtab[dy < 0 ? -dy : dy];
        const int s2 = (((y + dy) % nb) + nb) % nb;
        const double* rr = ring + (size_t)s2 * RS;
        const double* cr = cring + (size_t)s2 * RS;
        if (WEIGHTED) {
          for (int dx = -w; dx <= w; ++dx) {
            const double k = p.ktab[dy * dy + dx * dx];
            S += k * rr[c + dx];
            if (MASKND) A += k * cr[c + dx];
          }
        } else {
          S += rr[c + w + 1] - rr[c - w];
          if (MASKND) A += cr[c + w + 1] - cr[c - w];
        }
      }
      const long long pt = (long long)y * p.nx + xo;
      const long long o = (plane * p.n_thr + t) * (long long)p.ny * p.nx + pt;
      float res;
      if (sum_only) {
        res = (float)S;
      } else {
        float Af = MASKND ? (float)A : (p.mask2d ? p.area[pt] : p.area_const);
        res = (Af == 0.f) ? __int_as_float(0x7fc00000) : (float)(S / (double)Af * p.inv_members);
      }
      p.out[o] = res;
      if (p.out_mask) {
        bool masked = p.mask2d ? (p.mask2d[pt] == 0) : false;
        if (MASKND) masked = masked || (mnd_plane[pt] != 0);
        p.out_mask[o] = masked ? 1 : 0;
      }
    }
    __syncthreads();   // ring slot of row yy+1 is overwritten next iteration
  }
  if (__isnanf(nanacc)) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
}

// area_sum for an external 2-D mask: float32(sum K * ok[clamped]) (nbhood.py:289-300)
__global__ void circ_area_kernel(const uint8_t* mask, float* area, const double* ktab, const int* wtab,
                                 int ny, int nx, int r, int weighted) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const int y = (int)(i / nx), x = (int)(i % nx);
  double A = 0.0;
  for (int dy = -r; dy <= r; ++dy) {
    const int w = wtab[dy < 0 ? -dy : dy];
    const int ys = min(max(y + dy, 0), ny - 1);
    for (int dx = -w; dx <= w; ++dx) {
      const int xs = min(max(x + dx, 0), nx - 1);
      const double k = weighted ? ktab[dy * dy + dx * dx] : 1.0;
      A += mask[(long long)ys * nx + xs] ? k : 0.0;
    }
  }
  area[i] = (float)A;
}

__global__ void circ_tables_kernel(double* ktab, int* wtab, float* area_const, int r, int weighted) {
  // single block
  const int r2 = r * r;
  for (int d2 = threadIdx.x; d2 <= r2; d2 += blockDim.x)
    ktab[d2] = weighted ? (double)(r2 - d2) / (double)r2 : 1.0;     // nbhood.py:85-91
  for (int dy = threadIdx.x; dy <= r; dy += blockDim.x) {
    int w = (int)floor(sqrt((double)(r2 - dy * dy)));
    while ((w + 1) * (w + 1) + dy * dy <= r2) ++w;
    while (w > 0 && w * w + dy * dy > r2) --w;
    wtab[dy] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int dy = -r; dy <= r; ++dy) {
      const int w = wtab[dy < 0 ? -dy : dy];
      for (int dx = -w; dx <= w; ++dx) tot += ktab[dy * dy + dx * dx];
    }
    *area_const = (float)tot;
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct CircWorkspace {
  double* ktab;
  int* wtab;
  float* area_const;
  float* area;
  size_t total;
};

static void carve(const NbhDesc& d, void* ws, CircWorkspace* w) {
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  size_t off = 0;
  const size_t r = (size_t)d.cells;
  w->ktab = reinterpret_cast<double*>(base + off); off = align_up(off + (r * r + 1) * 8, 256);
  w->wtab = reinterpret_cast<int*>(base + off); off = align_up(off + (r + 1) * 4, 256);
  w->area_const = reinterpret_cast<float*>(base + off); off = align_up(off + 4, 256);
  w->area = reinterpret_cast<float*>(base + off); off = align_up(off + (size_t)d.ny * d.nx * 4, 256);
  w->total = off;
}

size_t circular_workspace_bytes(const NbhDesc& d) {
  CircWorkspace w;
  carve(d, nullptr, &w);
  return w.total;
}

int validate_desc(const NbhDesc* d);
int fill_thresholds(const NbhDesc& d, ThrSet* ts, int* tm);
int env_int(const char* name, int dflt);

template <int TM, bool MASKND, bool WEIGHTED>
static int launch_circ(const CircParams& p, int tx, size_t smem, long long blocks, cudaStream_t st) {
  auto kern = circular_kernel<TM, MASKND, WEIGHTED>;
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, tx, smem, st>>>(p);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int TM>
static int launch_circ_tm(const CircParams& p, int tx, size_t smem, long long blocks, bool masknd,
                          bool weighted, cudaStream_t st) {
  if (masknd) return weighted ? launch_circ<TM, true, true>(p, tx, smem, blocks, st)
                              : launch_circ<TM, true, false>(p, tx, smem, blocks, st);
  return weighted ? launch_circ<TM, false, true>(p, tx, smem, blocks, st)
                  : launch_circ<TM, false, false>(p, tx, smem, blocks, st);
}

int run_circular(const NbhDesc* d, const float* in, float* out, uint8_t* out_mask, int32_t* status,
                 void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (validate_desc(d)) return 1;
  NBH_REQUIRE(in && out && status && workspace, "null pointer argument");
  NBH_REQUIRE(workspace_bytes >= circular_workspace_bytes(*d), "workspace too small");
  NBH_REQUIRE(!(d->flags & NBH_WRAP_X), "wrap_x is not supported for circular neighbourhoods");
  const bool masknd = d->mask_nd != nullptr;
  const bool weighted = (d->flags & NBH_WEIGHTED) != 0;
  const int r = d->cells, nb = 2 * r + 1;
  CircWorkspace w;
  carve(*d, workspace, &w);

  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 227 * 1024;
  int tx = 256;
  auto smem_for = [&](int t) {
    return (size_t)nb * (t + 1) * 8 * (masknd ? 2 : 1) + 64 * 8 + (size_t)(r + 1) * 4 + 16;
  };
  while (tx > 32 && (smem_for(tx) > (size_t)smem_max)) tx -= 32;
  NBH_REQUIRE(smem_for(tx) <= (size_t)smem_max && tx > 2 * r,
              "circular neighbourhood radius of %d cells does not fit in shared memory", r);
  // prefer smaller blocks when the radius is small (more blocks per SM)
  if (r <= 12 && tx > 128) tx = 128;
  tx = env_int("NBH_CIRC_TX", tx);

  CircParams p;
  memset(&p, 0, sizeof(p));
  int tm = 0;
  if (fill_thresholds(*d, &p.thr, &tm)) return 1;
  p.in = in; p.out = out; p.out_mask = out_mask; p.mask2d = d->mask2d; p.mask_nd = d->mask_nd;
  p.area = w.area; p.ktab = w.ktab; p.wtab = w.wtab; p.status = status;
  p.plane_stride = d->plane_stride; p.member_stride = d->member_stride;
  p.n_planes = d->n_planes; p.n_members = d->n_members;
  p.ny = d->ny; p.nx = d->nx; p.r = r; p.nb = nb; p.flags = d->flags;
  p.WS = tx;
  const int wmax = tx - 2 * r;
  p.n_strips = (d->nx + wmax - 1) / wmax;
  p.W = min(wmax, (d->nx + p.n_strips - 1) / p.n_strips);
  p.n_thr = max(d->n_thresholds, 1);
  p.inv_members = (d->flags & NBH_SUM_ONLY) ? 1.0 : 1.0 / (double)d->n_members;
  const long long tiles_xy = (long long)p.n_strips * d->n_planes * p.n_thr;
  int ysegs = env_int("NBH_CIRC_YSEGS", 0);
  if (ysegs <= 0) {
    const long long want = 4LL * sm_count() * 2;
    ysegs = (int)max(1LL, min((long long)d->ny / max(16, 4 * r), (want + tiles_xy - 1) / tiles_xy));
  }
  ysegs = max(1, min(ysegs, d->ny));
  p.seg_rows = (d->ny + ysegs - 1) / ysegs;
  p.n_ysegs = (d->ny + p.seg_rows - 1) / p.seg_rows;

  NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  circ_tables_kernel<<<1, 256, 0, st>>>(w.ktab, w.wtab, w.area_const, r, weighted ? 1 : 0);
  if (d->mask2d && !masknd && !(d->flags & NBH_SUM_ONLY)) {
    const long long n = (long long)d->ny * d->nx;
    circ_area_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d->mask2d, w.area, w.ktab, w.wtab, d->ny,
                                                                d->nx, r, weighted ? 1 : 0);
  }
  // area_const is needed by value: read it back lazily on the device instead
  // (kept in the workspace; the kernel reads it through p.area_const_ptr)
  NBH_CHECK_CUDA(cudaGetLastError());
  // The unmasked area constant is computed on the host as well (same formula)
  {
    double tot = 0.0;
    const int r2 = r * r;
    for (int dy = -r; dy <= r; ++dy)
      for (int dx = -r; dx <= r; ++dx) {
        const int d2 = dy * dy + dx * dx;
        if (d2 > r2) continue;
        tot += weighted ? (double)(r2 - d2) / (double)r2 : 1.0;
      }
    p.area_const = (float)tot;
  }
  const long long blocks = (long long)p.n_strips * p.n_ysegs * p.n_thr * d->n_planes;
  NBH_REQUIRE(blocks < (1LL << 31), "too many blocks (%lld); split the launch", blocks);
  const size_t smem = smem_for(tx);
  ProfileScope prof(st);
  switch (tm) {
    case 0: return launch_circ_tm<0>(p, tx, smem, blocks, masknd, weighted, st);
    case 1: return launch_circ_tm<1>(p, tx, smem, blocks, masknd, weighted, st);
    default: return launch_circ_tm<2>(p, tx, smem, blocks, masknd, weighted, st);
  }
}

}  // namespace nbh
