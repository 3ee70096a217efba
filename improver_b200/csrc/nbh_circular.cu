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

int env_int(const char* name, int dflt);

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


// ---------------------------------------------------------------------------
// Large radii: the 2r+1-row prefix ring no longer fits in shared memory (and a
// strip cannot even hold the 2r halo), so the float64 row prefixes of a batch
// of planes are written to global memory once (they stay L2-resident: a UK
// plane is 8 MB) and every output gathers 2 prefix values per kernel row.
// ---------------------------------------------------------------------------
struct CircGlobalParams {
  const float* in;
  float* out;
  uint8_t* out_mask;
  const uint8_t* mask2d;
  const uint8_t* mask_nd;
  const float* area;
  const int* wtab;
  double* prefix;        // [batch][ny][nx] inclusive row prefix of the member-summed, masked values
  double* cprefix;       // MASKND: same for the valid mask
  int32_t* status;
  long long plane_stride, member_stride;
  long long plane0;      // first plane of this batch
  int n_batch;
  int n_members, ny, nx, r;
  int flags, n_thr, t;
  double inv_members;
  float area_const;
  ThrDev thr;
};

// one warp per (plane, row): member sum + inclusive prefix along x
__global__ void __launch_bounds__(256) circ_prefix_kernel(const CircGlobalParams p) {
  const int lane = threadIdx.x & 31;
  const long long row_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row_id >= (long long)p.n_batch * p.ny) return;
  const int b = (int)(row_id / p.ny), y = (int)(row_id % p.ny);
  const long long plane = p.plane0 + b;
  const float* in_row = p.in + plane * p.plane_stride + (long long)y * p.nx;
  const uint8_t* mnd_row = p.mask_nd ? p.mask_nd + plane * p.plane_stride + (long long)y * p.nx : nullptr;
  const uint8_t* m2_row = p.mask2d ? p.mask2d + (long long)y * p.nx : nullptr;
  double* pre = p.prefix + ((long long)b * p.ny + y) * p.nx;
  double* cpre = p.cprefix ? p.cprefix + ((long long)b * p.ny + y) * p.nx : nullptr;
  double carry = 0.0, ccarry = 0.0;
  float nanacc = 0.f;
  for (int x0 = 0; x0 < p.nx; x0 += 32) {
    const int x = x0 + lane;
    float P = 0.f, okf = 0.f;
    if (x < p.nx) {
      const bool m2 = m2_row ? (m2_row[x] != 0) : true;
      bool valid = m2, seen = true;
      if (mnd_row) { seen = mnd_row[x] == 0; valid = valid && seen; }
      for (int m = 0; m < p.n_members; ++m) {
        const float d = __ldg(in_row + (long long)m * p.member_stride + x);
        if (seen) nanacc = max_nan(nanacc, fabsf(d));
        const float tv = p.thr.mode ? truth_value(d, p.thr) : d;
        P += valid ? tv : 0.f;
      }
      okf = valid ? 1.f : 0.f;
    }
    double incl = (double)P, cincl = (double)okf;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = __shfl_up_sync(0xffffffffu, incl, d);
      const double oc = __shfl_up_sync(0xffffffffu, cincl, d);
      if (lane >= d) { incl += o; cincl += oc; }
    }
    if (x < p.nx) {
      pre[x] = carry + incl;
      if (cpre) cpre[x] = ccarry + cincl;
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
    ccarry += __shfl_sync(0xffffffffu, cincl, 31);
  }
  if (__any_sync(0xffffffffu, __isnanf(nanacc)) && lane == 0)
    atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
}

// extended inclusive prefix of a periodic row: E(c) for any c in [-nx, 2nx)
__device__ __forceinline__ double wrap_prefix(const double* pre, int c, int nx) {
  const double total = pre[nx - 1];
  if (c < 0) return pre[c + nx] - total;
  if (c >= nx) return pre[c - nx] + total;
  return pre[c];
}

__device__ __forceinline__ double span_sum_wrap(const double* pre, int x, int w, int nx) {
  return wrap_prefix(pre, x + w, nx) - wrap_prefix(pre, x - w - 1, nx);
}

__device__ __forceinline__ double span_sum(const double* pre, int x, int w, int nx) {
  const int lo = max(x - w, 0), hi = min(x + w, nx - 1);
  double s = pre[hi] - (lo > 0 ? pre[lo - 1] : 0.0);
  const int nl = lo - (x - w), nr = (x + w) - hi;     // replicated edge cells (mode="nearest")
  if (nl > 0) s += (double)nl * pre[0];
  if (nr > 0) s += (double)nr * (pre[nx - 1] - (nx > 1 ? pre[nx - 2] : 0.0));
  return s;
}

__global__ void __launch_bounds__(256) circ_gather_kernel(const CircGlobalParams p) {
  extern __shared__ int wtab_s[];
  for (int k = threadIdx.x; k <= p.r; k += blockDim.x) wtab_s[k] = p.wtab[k];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_plane = (long long)p.ny * p.nx;
  if (i >= (long long)p.n_batch * per_plane) return;
  const int b = (int)(i / per_plane);
  const long long pt = i - (long long)b * per_plane;
  const int y = (int)(pt / p.nx), x = (int)(pt % p.nx);
  const double* pre = p.prefix + (long long)b * per_plane;
  const double* cpre = p.cprefix ? p.cprefix + (long long)b * per_plane : nullptr;
  double S = 0.0, A = 0.0;
  if (!(p.flags & NBH_WRAP_X) && !cpre && x - p.r - 1 >= 0 && x + p.r <= p.nx - 1) {
    // every span lies inside the row: two loads and two adds per kernel row
    for (int dy = -p.r; dy <= p.r; ++dy) {
      const int w = wtab_s[dy < 0 ? -dy : dy];
      const double* row = pre + (long long)min(max(y + dy, 0), p.ny - 1) * p.nx + x;
      S += row[w] - row[-w - 1];
    }
  } else
  for (int dy = -p.r; dy <= p.r; ++dy) {
    const int w = wtab_s[dy < 0 ? -dy : dy];
    const int ys = min(max(y + dy, 0), p.ny - 1);
    if (p.flags & NBH_WRAP_X) {
      S += span_sum_wrap(pre + (long long)ys * p.nx, x, w, p.nx);
      if (cpre) A += span_sum_wrap(cpre + (long long)ys * p.nx, x, w, p.nx);
    } else {
      S += span_sum(pre + (long long)ys * p.nx, x, w, p.nx);
      if (cpre) A += span_sum(cpre + (long long)ys * p.nx, x, w, p.nx);
    }
  }
  const long long plane = p.plane0 + b;
  const long long o = (plane * p.n_thr + p.t) * per_plane + pt;
  float res;
  if (p.flags & NBH_SUM_ONLY) {
    res = (float)S;
  } else {
    const float Af = cpre ? (float)A : (p.mask2d ? p.area[pt] : p.area_const);
    res = (Af == 0.f) ? __int_as_float(0x7fc00000) : (float)(S / (double)Af * p.inv_members);
  }
  p.out[o] = res;
  if (p.out_mask) {
    bool masked = p.mask2d ? (p.mask2d[pt] == 0) : false;
    if (p.mask_nd) masked = masked || (p.mask_nd[plane * p.plane_stride + pt] != 0);
    p.out_mask[o] = masked ? 1 : 0;
  }
}


// Shared-memory tiled gather: a block stages the (TH + 2r) x (TW + 2r + 1)
// window of prefix values its TH x TW outputs need (each value is then read
// ~2(2r+1) times from shared memory instead of L2).
constexpr int kCircTW = 64;

__global__ void __launch_bounds__(256) circ_gather_tiled_kernel(const CircGlobalParams p, int TH, int tiles_x,
                                                                int tiles_y) {
  extern __shared__ __align__(16) unsigned char smem_c[];
  const int r = p.r;
  const int TC = kCircTW + 2 * r + 1, TR = TH + 2 * r;
  double* tile = reinterpret_cast<double*>(smem_c);            // [TR][TC], col j <-> global col x0 - r - 1 + j
  int* wtab_s = reinterpret_cast<int*>(tile + (size_t)TR * TC);
  for (int k = threadIdx.x; k <= r; k += blockDim.x) wtab_s[k] = p.wtab[k];

  long long bid = blockIdx.x;
  const int tx = (int)(bid % tiles_x); bid /= tiles_x;
  const int ty = (int)(bid % tiles_y); bid /= tiles_y;
  const int b = (int)bid;
  const int x0 = tx * kCircTW, y0 = ty * TH;
  const long long per_plane = (long long)p.ny * p.nx;
  const double* pre = p.prefix + (long long)b * per_plane;
#pragma unroll 8
  for (int i = threadIdx.x; i < TR * TC; i += blockDim.x) {
    const int jr = i / TC, jc = i - jr * TC;
    const int ys = min(max(y0 - r + jr, 0), p.ny - 1);          // edge replication in y
    const int gc = x0 - r - 1 + jc;
    if (p.flags & NBH_WRAP_X) tile[i] = wrap_prefix(pre + (long long)ys * p.nx, gc, p.nx);
    else tile[i] = gc < 0 ? 0.0 : pre[(long long)ys * p.nx + min(gc, p.nx - 1)];
  }
  __syncthreads();

  const int lx = threadIdx.x % kCircTW;
  const int x = x0 + lx;
  if (x >= p.nx) return;
  const long long plane = p.plane0 + b;
  const int cshift = r + 1 - x0;                                // tile column of global column gc: gc + cshift
  for (int ly = threadIdx.x / kCircTW; ly < TH; ly += blockDim.x / kCircTW) {
    const int y = y0 + ly;
    if (y >= p.ny) break;
    double S = 0.0;
    if ((p.flags & NBH_WRAP_X) || (x - r - 1 >= -1 && x + r <= p.nx - 1)) {
      // every span lies inside the row (or the axis is periodic): two loads and two adds per kernel row
      const double* row = tile + (size_t)ly * TC + cshift + x;
      for (int dy = -r; dy <= r; ++dy) {
        const int w = wtab_s[dy < 0 ? -dy : dy];
        S += row[w] - row[-w - 1];
        row += TC;
      }
    } else {
      for (int dy = -r; dy <= r; ++dy) {
        const int w = wtab_s[dy < 0 ? -dy : dy];
        const double* row = tile + (size_t)(ly + r + dy) * TC + cshift;   // indexable by global column >= -1
        const int lo = max(x - w, 0), hi = min(x + w, p.nx - 1);
        double s = row[hi] - row[lo - 1];
        const int nl = lo - (x - w), nr = (x + w) - hi;             // replicated edge cells (mode="nearest")
        if (nl > 0) s += (double)nl * (row[0] - row[-1]);
        if (nr > 0) s += (double)nr * (row[p.nx - 1] - row[p.nx - 2]);
        S += s;
      }
    }
    const long long pt = (long long)y * p.nx + x;
    const long long o = (plane * p.n_thr + p.t) * per_plane + pt;
    float res;
    if (p.flags & NBH_SUM_ONLY) {
      res = (float)S;
    } else {
      const float Af = p.mask2d ? p.area[pt] : p.area_const;
      res = (Af == 0.f) ? __int_as_float(0x7fc00000) : (float)(S / (double)Af * p.inv_members);
    }
    p.out[o] = res;
    if (p.out_mask) p.out_mask[o] = (p.mask2d && p.mask2d[pt] == 0) ? 1 : 0;
  }
}

// pick the tallest tile that fits; 0 = use the direct (L2) gather
static int circ_tile_rows(int r, int nx, size_t* smem) {
  // measured on B200: the tile pays off while it is small enough for several blocks per SM
  if (nx < 3 || r > env_int("NBH_CIRC_TILED_MAX_R", 16)) return 0;
  for (int th : {64, 32, 16, 8}) {
    const size_t bytes = (size_t)(th + 2 * r) * (kCircTW + 2 * r + 1) * 8 + (size_t)(r + 1) * 4 + 16;
    if (bytes <= (th >= 32 ? 100u : 200u) * 1024) { *smem = bytes; return th; }
  }
  return 0;
}

static int launch_circ_gather(const CircGlobalParams& g, int ny, int nx, cudaStream_t st) {
  size_t smem = 0;
  const int th = (g.cprefix || env_int("NBH_CIRC_TILED", 1) == 0) ? 0 : circ_tile_rows(g.r, nx, &smem);
  if (th > 0) {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(circ_gather_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    const int tiles_x = (nx + kCircTW - 1) / kCircTW, tiles_y = (ny + th - 1) / th;
    const long long blocks = (long long)tiles_x * tiles_y * g.n_batch;
    NBH_REQUIRE(blocks < (1LL << 31), "too many blocks");
    circ_gather_tiled_kernel<<<(unsigned)blocks, 256, smem, st>>>(g, th, tiles_x, tiles_y);
  } else {
    const long long outs = (long long)g.n_batch * ny * nx;
    circ_gather_kernel<<<(unsigned)((outs + 255) / 256), 256, (size_t)(g.r + 1) * 4, st>>>(g);
  }
  return 0;
}

// area_sum for an external mask at large radius, from the prefix of the mask itself
__global__ void circ_mask_to_float_kernel(const uint8_t* mask, float* as_float, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) as_float[i] = mask[i] ? 1.f : 0.f;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kCircSmemMaxRadius = 40;          // beyond this the global-prefix path is used
constexpr size_t kCircPrefixBudget = 256u << 20;  // bytes of float64 prefixes per batch (L2-friendly)

int env_int(const char* name, int dflt);
// unweighted discs: global-prefix gather from radius NBH_CIRC_GLOBAL_MIN_R upwards
static bool use_global_path(const NbhDesc& d) {
  if (d.cells > kCircSmemMaxRadius) return true;
  if (d.flags & NBH_WEIGHTED) return false;
  return d.cells >= env_int("NBH_CIRC_GLOBAL_MIN_R", 1);   // measured on B200: wins at every radius
}

struct CircWorkspace {
  double* ktab;
  int* wtab;
  float* area_const;
  float* area;
  double* prefix;
  double* cprefix;
  float* maskf;
  int batch;
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
  w->prefix = w->cprefix = nullptr; w->maskf = nullptr; w->batch = 0;
  if (use_global_path(d)) {
    const size_t per_plane = (size_t)d.ny * d.nx * 8;
    long long batch = (long long)(kCircPrefixBudget / per_plane);
    if (batch < 1) batch = 1;
    if (batch > d.n_planes) batch = d.n_planes;
    w->batch = (int)batch;
    w->prefix = reinterpret_cast<double*>(base + off); off = align_up(off + per_plane * batch, 256);
    if (d.mask_nd) { w->cprefix = reinterpret_cast<double*>(base + off); off = align_up(off + per_plane * batch, 256); }
    w->maskf = reinterpret_cast<float*>(base + off); off = align_up(off + (size_t)d.ny * d.nx * 4, 256);
  }
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
  NBH_REQUIRE(!(d->flags & NBH_WRAP_X) || (use_global_path(*d) && d->nx > 2 * d->cells + 1),
              "wrap_x needs an unweighted circular neighbourhood narrower than the grid");
  const bool masknd = d->mask_nd != nullptr;
  const bool weighted = (d->flags & NBH_WEIGHTED) != 0;
  const int r = d->cells, nb = 2 * r + 1;
  CircWorkspace w;
  carve(*d, workspace, &w);

  if (use_global_path(*d)) {
    NBH_REQUIRE(!weighted, "weighted_mode is limited to radii of at most %d grid cells", kCircSmemMaxRadius);
    CircGlobalParams g;
    memset(&g, 0, sizeof(g));
    ThrSet ts;
    int tm = 0;
    if (fill_thresholds(*d, &ts, &tm)) return 1;
    g.in = in; g.out = out; g.out_mask = out_mask; g.mask2d = d->mask2d; g.mask_nd = d->mask_nd;
    g.area = w.area; g.wtab = w.wtab; g.prefix = w.prefix; g.cprefix = w.cprefix; g.status = status;
    g.plane_stride = d->plane_stride; g.member_stride = d->member_stride;
    g.n_members = d->n_members; g.ny = d->ny; g.nx = d->nx; g.r = r; g.flags = d->flags;
    g.n_thr = max(d->n_thresholds, 1);
    g.inv_members = (d->flags & NBH_SUM_ONLY) ? 1.0 : 1.0 / (double)d->n_members;
    {
      double tot = 0.0;
      for (int dy = -r; dy <= r; ++dy) tot += 2.0 * floor(sqrt((double)(r * r - dy * dy))) + 1.0;
      g.area_const = (float)tot;
    }
    NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
    circ_tables_kernel<<<1, 256, 0, st>>>(w.ktab, w.wtab, w.area_const, r, 0);
    const long long npts = (long long)d->ny * d->nx;
    if (d->mask2d && !masknd && !(d->flags & NBH_SUM_ONLY)) {
      // area_sum = the same gather applied to the mask (one plane, sum_only)
      circ_mask_to_float_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(d->mask2d, w.maskf, npts);
      CircGlobalParams a = g;
      a.in = w.maskf; a.out = w.area; a.out_mask = nullptr; a.mask2d = nullptr; a.mask_nd = nullptr;
      a.cprefix = nullptr; a.plane_stride = npts; a.member_stride = 0; a.n_members = 1; a.n_thr = 1; a.t = 0;
      a.plane0 = 0; a.n_batch = 1; a.flags = NBH_SUM_ONLY | (d->flags & NBH_WRAP_X); a.thr.mode = 0;
      circ_prefix_kernel<<<(unsigned)((d->ny + 7) / 8), 256, 0, st>>>(a);
      if (launch_circ_gather(a, d->ny, d->nx, st)) return 1;
    }
    ProfileScope prof(st);
    for (int t = 0; t < g.n_thr; ++t) {
      g.t = t;
      g.thr = ts.t[d->n_thresholds ? t : 0];
      if (!d->n_thresholds) g.thr.mode = 0;
      for (long long p0 = 0; p0 < d->n_planes; p0 += w.batch) {
        g.plane0 = p0;
        g.n_batch = (int)min((long long)w.batch, d->n_planes - p0);
        const long long rows = (long long)g.n_batch * d->ny;
        circ_prefix_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(g);
        if (launch_circ_gather(g, d->ny, d->nx, st)) return 1;
      }
    }
    NBH_CHECK_CUDA(cudaGetLastError());
    return 0;
  }

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
