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


// ---------------------------------------------------------------------------
// Unweighted discs of any radius: row-span decomposition over row prefix sums.
//
// Pass 1 (circ_prefix_kernel) writes the inclusive row prefix of the (thresholded, masked,
// member-summed) field of a batch of planes to the workspace; pass 2 (the gather kernels) forms
// every output as the sum over the 2r+1 kernel rows of a prefix difference, with the
// edge-replication terms of scipy's mode="nearest".
//
// Prefix type PT.  Values known to lie in [0, 1] (truth values; raw probabilities are tried
// speculatively, see `guard`) are turned into fixed point (P * 2^fx_shift, round to nearest) and
// the row prefix is an unsigned 32-bit sum whose wrap-around cancels in every span difference
// ((2r+1) * R * 2^fx_shift < 2^32); the spans are accumulated in 64 bits.  Half the bytes per
// prefix value and integer adds instead of DADDs; the quantisation contributes at most
// 2^-(fx_shift+1) to a mean.  Anything else uses float64 prefixes like the reference's float64
// correlate.
//
// NBH_LAND_SEA: two prefix planes, land * v and sea * v; a point gathers from the plane of its
// own surface type only, and is normalised by the area of that type (A_sea = sum(K) - A_land,
// exact: mode="nearest" maps every kernel cell to some grid cell).
// ---------------------------------------------------------------------------
constexpr int kCircParamRadius = 400;          // widths table passed as a kernel parameter up to this radius

struct CircWidths { short w[2 * kCircParamRadius + 2]; };     // w[dy + r], dy = -r..r

struct CircGlobalParams {
  const float* in;
  float* out;
  uint8_t* out_mask;
  const uint8_t* mask2d;
  const uint8_t* mask_nd;
  const float* area;     // mask2d launches: area of the own surface type (LAND_SEA) / of the valid cells
  const int* wtab;       // half width per |dy| (radii beyond kCircParamRadius)
  void* prefix;          // [batch][ny][nx] inclusive row prefix (PT) of the member-summed, masked values
  void* prefix2;         // LAND_SEA: same for the complementary (sea) points
  double* cprefix;       // MASKND: same for the valid mask (float64 path only)
  int32_t* status;
  int* guard;            // [0]: set by the speculative fixed-point prefix pass when a raw value lies
                         // outside [0, 1]; the float64 kernels of a speculative launch exit while it is 0
  long long plane_stride, member_stride;
  long long plane0;      // first plane of this batch
  int n_batch;
  int n_members, ny, nx, r;
  int flags, n_thr, t;
  int guarded;           // 1: this is the float64 re-run of a speculative launch (exit unless *guard)
  int check_range;       // 1: fixed-point pass over raw data: flag values outside [0, 1]
  double inv_members;
  float area_const;
  float fx_scale;        // 2^fx_shift
  double fx_inv;         // 2^-fx_shift
  ThrDev thr;
};

template <typename PT> struct CircAcc;
template <> struct CircAcc<double> {
  typedef double acc_t;
  static __device__ __forceinline__ double cvt(float P, float) { return (double)P; }
  static __device__ __forceinline__ double result(double S, double) { return S; }
};
template <> struct CircAcc<unsigned> {
  typedef unsigned long long acc_t;
  static __device__ __forceinline__ unsigned cvt(float P, float scale) { return __float2uint_rn(P * scale); }
  static __device__ __forceinline__ double result(unsigned long long S, double inv) { return (double)S * inv; }
};

template <typename PT> __device__ __forceinline__ PT warp_incl_scan(PT v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const PT o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// one warp per (plane, row): member sum + inclusive prefix along x.  A lane owns four consecutive
// elements of every 128-element chunk (lane-local prefix of four, one warp scan per chunk).
template <typename PT, bool LS>
__global__ void __launch_bounds__(256) circ_prefix_kernel(const CircGlobalParams p) {
  if (p.guarded && *p.guard == 0) return;
  const int lane = threadIdx.x & 31;
  const long long row_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row_id >= (long long)p.n_batch * p.ny) return;
  const int b = (int)(row_id / p.ny), y = (int)(row_id % p.ny);
  const long long plane = p.plane0 + b;
  const float* in_row = p.in + plane * p.plane_stride + (long long)y * p.nx;
  const uint8_t* mnd_row = p.mask_nd ? p.mask_nd + plane * p.plane_stride + (long long)y * p.nx : nullptr;
  const uint8_t* m2_row = p.mask2d ? p.mask2d + (long long)y * p.nx : nullptr;
  PT* pre = reinterpret_cast<PT*>(p.prefix) + ((long long)b * p.ny + y) * p.nx;
  PT* pre2 = LS ? reinterpret_cast<PT*>(p.prefix2) + ((long long)b * p.ny + y) * p.nx : nullptr;
  double* cpre = p.cprefix ? p.cprefix + ((long long)b * p.ny + y) * p.nx : nullptr;
  PT carry = 0, carry2 = 0;
  double ccarry = 0.0;
  float nanacc = 0.f;
  bool bad = false;
  constexpr int E = 4;
  for (int x0 = 0; x0 < p.nx; x0 += 32 * E) {
    const int xb = x0 + lane * E;
    PT q[E], q2[E];
    double ok[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int x = xb + e;
      float P = 0.f, okf = 0.f;
      bool land = true;
      if (x < p.nx) {
        const bool m2 = m2_row ? (m2_row[x] != 0) : true;
        land = m2;
        bool valid = LS ? true : m2, seen = true;
        if (mnd_row) { seen = mnd_row[x] == 0; valid = valid && seen; }
        for (int m = 0; m < p.n_members; ++m) {
          const float d = __ldg(in_row + (long long)m * p.member_stride + x);
          if (seen) nanacc = max_nan(nanacc, fabsf(d));
          if (p.check_range && !(d >= 0.f && d <= 1.f)) bad = true;
          const float tv = p.thr.mode ? truth_value(d, p.thr) : d;
          P += valid ? tv : 0.f;
        }
        okf = valid ? 1.f : 0.f;
      }
      const PT v = CircAcc<PT>::cvt(P, p.fx_scale);
      q[e] = LS ? (land ? v : (PT)0) : v;
      q2[e] = (LS && !land) ? v : (PT)0;
      ok[e] = (double)okf;
    }
    // lane-local inclusive prefix, then the warp scan of the lane totals
#pragma unroll
    for (int e = 1; e < E; ++e) { q[e] += q[e - 1]; if (LS) q2[e] += q2[e - 1]; ok[e] += ok[e - 1]; }
    const PT incl = warp_incl_scan<PT>(q[E - 1], lane);
    const PT off = carry + (incl - q[E - 1]);
#pragma unroll
    for (int e = 0; e < E; ++e) if (xb + e < p.nx) pre[xb + e] = off + q[e];
    carry += __shfl_sync(0xffffffffu, incl, 31);
    if (LS) {
      const PT incl2 = warp_incl_scan<PT>(q2[E - 1], lane);
      const PT off2 = carry2 + (incl2 - q2[E - 1]);
#pragma unroll
      for (int e = 0; e < E; ++e) if (xb + e < p.nx) pre2[xb + e] = off2 + q2[e];
      carry2 += __shfl_sync(0xffffffffu, incl2, 31);
    }
    if (cpre) {
      const double cincl = warp_incl_scan<double>(ok[E - 1], lane);
      const double coff = ccarry + (cincl - ok[E - 1]);
#pragma unroll
      for (int e = 0; e < E; ++e) if (xb + e < p.nx) cpre[xb + e] = coff + ok[e];
      ccarry += __shfl_sync(0xffffffffu, cincl, 31);
    }
  }
  if (__any_sync(0xffffffffu, __isnanf(nanacc)) && lane == 0)
    atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  if (p.check_range && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.guard, 1);
}

// extended inclusive prefix of a periodic row: E(c) for any c in [-nx, 2nx)
template <typename PT> __device__ __forceinline__ PT wrap_prefix(const PT* pre, int c, int nx) {
  const PT total = pre[nx - 1];
  if (c < 0) return pre[c + nx] - total;
  if (c >= nx) return pre[c - nx] + total;
  return pre[c];
}

template <typename PT> __device__ __forceinline__ PT span_sum_wrap(const PT* pre, int x, int w, int nx) {
  return wrap_prefix<PT>(pre, x + w, nx) - wrap_prefix<PT>(pre, x - w - 1, nx);
}

template <typename PT> __device__ __forceinline__ PT span_sum(const PT* pre, int x, int w, int nx) {
  const int lo = max(x - w, 0), hi = min(x + w, nx - 1);
  PT s = pre[hi] - (lo > 0 ? pre[lo - 1] : (PT)0);
  const int nl = lo - (x - w), nr = (x + w) - hi;     // replicated edge cells (mode="nearest")
  if (nl > 0) s += (PT)nl * pre[0];
  if (nr > 0) s += (PT)nr * (pre[nx - 1] - (nx > 1 ? pre[nx - 2] : (PT)0));
  return s;
}

template <typename PT>
__device__ __forceinline__ void circ_store(const CircGlobalParams& p, typename CircAcc<PT>::acc_t S, double A,
                                           bool have_A, long long plane, long long pt, long long per_plane) {
  const long long o = (plane * p.n_thr + p.t) * per_plane + pt;
  const double Sd = CircAcc<PT>::result(S, p.fx_inv);
  float res;
  if (p.flags & NBH_SUM_ONLY) {
    res = (float)Sd;
  } else {
    const float Af = have_A ? (float)A : (p.mask2d ? p.area[pt] : p.area_const);
    res = (Af == 0.f) ? __int_as_float(0x7fc00000) : (float)(Sd / (double)Af * p.inv_members);
  }
  p.out[o] = res;
  if (p.out_mask) {
    bool masked = (p.mask2d && !(p.flags & NBH_LAND_SEA)) ? (p.mask2d[pt] == 0) : false;
    if (p.mask_nd) masked = masked || (p.mask_nd[plane * p.plane_stride + pt] != 0);
    p.out_mask[o] = masked ? 1 : 0;
  }
}

// direct gather (prefix values from L1 / L2): any radius
template <typename PT, bool LS, bool PW>
__global__ void __launch_bounds__(256)
circ_gather_kernel(const CircGlobalParams p, const __grid_constant__ CircWidths cw) {
  if (p.guarded && *p.guard == 0) return;
  typedef typename CircAcc<PT>::acc_t acc_t;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_plane = (long long)p.ny * p.nx;
  if (i >= (long long)p.n_batch * per_plane) return;
  const int b = (int)(i / per_plane);
  const long long pt = i - (long long)b * per_plane;
  const int y = (int)(pt / p.nx), x = (int)(pt % p.nx);
  const bool sea = LS && p.mask2d[pt] == 0;
  const PT* pre = reinterpret_cast<const PT*>(sea ? p.prefix2 : p.prefix) + (long long)b * per_plane;
  const double* cpre = p.cprefix ? p.cprefix + (long long)b * per_plane : nullptr;
  acc_t S = 0;
  double A = 0.0;
  const int r = p.r;
  if (!(p.flags & NBH_WRAP_X) && !cpre && x - r - 1 >= 0 && x + r <= p.nx - 1) {
    // every span lies inside the row: two loads and two adds per kernel row
#pragma unroll 4
    for (int dy = -r; dy <= r; ++dy) {
      const int w = PW ? (int)cw.w[dy + r] : p.wtab[dy < 0 ? -dy : dy];
      const PT* row = pre + (long long)min(max(y + dy, 0), p.ny - 1) * p.nx + x;
      S += (acc_t)(PT)(row[w] - row[-w - 1]);
    }
  } else {
    for (int dy = -r; dy <= r; ++dy) {
      const int w = PW ? (int)cw.w[dy + r] : p.wtab[dy < 0 ? -dy : dy];
      const int ys = min(max(y + dy, 0), p.ny - 1);
      if (p.flags & NBH_WRAP_X) {
        S += (acc_t)span_sum_wrap<PT>(pre + (long long)ys * p.nx, x, w, p.nx);
        if (cpre) A += span_sum_wrap<double>(cpre + (long long)ys * p.nx, x, w, p.nx);
      } else {
        S += (acc_t)span_sum<PT>(pre + (long long)ys * p.nx, x, w, p.nx);
        if (cpre) A += span_sum<double>(cpre + (long long)ys * p.nx, x, w, p.nx);
      }
    }
  }
  circ_store<PT>(p, S, A, cpre != nullptr, p.plane0 + b, pt, per_plane);
}

// Direct gather, four results per thread and eight rows per block: a warp covers 128 consecutive
// columns of one output row (lane l the columns x0 + l + 32 j), the eight warps of a block eight
// consecutive rows, so the (8 + 2r) x (128 + 2r + 1) prefix values a block touches are shared by its
// 1024 results through L1 instead of every 256-point block streaming its own 2(2r+1) rows from L2
// (the one-point-per-thread kernel moved ~8 B per result and kernel row through L2: 6.8 TB/s at
// r = 81, the L2 bound).  Not for periodic x or count prefixes (circ_gather_kernel).
constexpr int kCircG4Rows = 8;                  // rows (warps) per block of the direct four-result gather (16 was measured: 2-4 % slower at every radius)

template <typename PT, bool LS, bool PW>
__global__ void __launch_bounds__(kCircG4Rows * 32)
circ_gather4_kernel(const CircGlobalParams p, const __grid_constant__ CircWidths cw, int tiles_x, int tiles_y) {
  if (p.guarded && *p.guard == 0) return;
  typedef typename CircAcc<PT>::acc_t acc_t;
  long long bid = blockIdx.x;
  const int tx = (int)(bid % tiles_x); bid /= tiles_x;
  const int ty = (int)(bid % tiles_y); bid /= tiles_y;
  const int b = (int)bid;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = tx * 128, y = ty * kCircG4Rows + warp;
  if (y >= p.ny) return;
  const int r = p.r;
  const long long per_plane = (long long)p.ny * p.nx;
  const PT* pre = reinterpret_cast<const PT*>(p.prefix) + (long long)b * per_plane;
  const PT* pre2 = LS ? reinterpret_cast<const PT*>(p.prefix2) + (long long)b * per_plane : nullptr;
  int xj[4];
  bool want[4], fast[4];
  const PT* base[4];                                              // the plane of the result's own type, at its column
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    xj[j] = x0 + lane + 32 * j;
    want[j] = xj[j] < p.nx;
    const bool inner = xj[j] - r - 1 >= 0 && xj[j] + r <= p.nx - 1;
    fast[j] = __all_sync(0xffffffffu, !want[j] || inner) != 0;
    // (columns past the row end read a column whose spans lie inside the row; never stored)
    const int xc = want[j] ? xj[j] : max(min(xj[j], p.nx - 1 - r), min(r + 1, p.nx - 1));
    const bool sea = LS && p.mask2d[(long long)y * p.nx + xc] == 0;
    base[j] = (sea ? pre2 : pre) + xc;
  }
  acc_t S[4] = {0, 0, 0, 0};
  if (fast[0] && fast[1] && fast[2] && fast[3] && x0 + 127 + r <= p.nx - 1 && x0 - r - 1 >= 0) {
    // every span of every column inside the row: per kernel row one row offset and one width for four results
#pragma unroll 2
    for (int dy = -r; dy <= r; ++dy) {
      const int w = PW ? (int)cw.w[dy + r] : p.wtab[dy < 0 ? -dy : dy];
      const long long ro = (long long)min(max(y + dy, 0), p.ny - 1) * p.nx;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const PT* q = base[j] + ro;
        S[j] += (acc_t)(PT)(q[w] - q[-w - 1]);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!__any_sync(0xffffffffu, want[j])) continue;
      if (fast[j]) {
#pragma unroll 4
        for (int dy = -r; dy <= r; ++dy) {
          const int w = PW ? (int)cw.w[dy + r] : p.wtab[dy < 0 ? -dy : dy];
          const PT* q = base[j] + (long long)min(max(y + dy, 0), p.ny - 1) * p.nx;
          S[j] += (acc_t)(PT)(q[w] - q[-w - 1]);
        }
        continue;
      }
      if (!want[j]) continue;
      const PT* pl = base[j] - xj[j];                             // start of the plane of the result's type (want: base is at xj)
      for (int dy = -r; dy <= r; ++dy) {
        const int w = PW ? (int)cw.w[dy + r] : p.wtab[dy < 0 ? -dy : dy];
        const int ys = min(max(y + dy, 0), p.ny - 1);
        S[j] += (acc_t)span_sum<PT>(pl + (long long)ys * p.nx, xj[j], w, p.nx);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (want[j]) circ_store<PT>(p, S[j], 0.0, false, p.plane0 + b, (long long)y * p.nx + xj[j], per_plane);
}

// Shared-memory tiled gather: a block stages the (TH + 2r) x (TW + 2r + 1) window of prefix
// values its TH x TW outputs need (each value is then read ~2(2r+1) times from shared memory
// instead of L1 / L2); LAND_SEA stages the window of both planes.
constexpr int kCircTW = 64;

template <typename PT, bool LS, bool PW, bool BOTH>
__global__ void __launch_bounds__(256)
circ_gather_tiled_kernel(const CircGlobalParams p, const __grid_constant__ CircWidths cw, int TH, int tiles_x,
                         int tiles_y) {
  if (p.guarded && *p.guard == 0) return;
  typedef typename CircAcc<PT>::acc_t acc_t;
  extern __shared__ __align__(16) unsigned char smem_c[];
  const int r = p.r;
  const int TC = kCircTW + 2 * r + 1, TR = TH + 2 * r;
  PT* tile = reinterpret_cast<PT*>(smem_c);            // [TR][TC], col j <-> global col x0 - r - 1 + j

  long long bid = blockIdx.x;
  const int tx = (int)(bid % tiles_x); bid /= tiles_x;
  const int ty = (int)(bid % tiles_y); bid /= tiles_y;
  const int b = (int)bid;
  const int x0 = tx * kCircTW, y0 = ty * TH;
  const long long per_plane = (long long)p.ny * p.nx;
  const int lx = threadIdx.x % kCircTW;
  const int x = x0 + lx;
  const long long plane = p.plane0 + b;
  const int cshift = r + 1 - x0;                                // tile column of global column gc: gc + cshift
  // land / sea launches: both planes staged side by side when they fit (BOTH), else one after the
  // other in the same buffer; a sequential pass is skipped when the tile holds no output of its type
  constexpr bool SEQ = LS && !BOTH;
#pragma unroll 1
  for (int k = 0; k < (LS ? 2 : 1); ++k) {
    PT* tl = tile + ((LS && BOTH && k == 1) ? (size_t)TR * TC : 0);
    if (SEQ) {
      bool mine = false;
      for (int ly = threadIdx.x / kCircTW; ly < TH; ly += blockDim.x / kCircTW) {
        const int y = y0 + ly;
        if (y < p.ny && x < p.nx) mine = mine || ((p.mask2d[(long long)y * p.nx + x] == 0) == (k == 1));
      }
      if (!__syncthreads_or(mine ? 1 : 0)) continue;
    }
    const PT* pre = reinterpret_cast<const PT*>(k ? p.prefix2 : p.prefix) + (long long)b * per_plane;
#pragma unroll 4
    for (int i = threadIdx.x; i < TR * TC; i += blockDim.x) {
      const int jr = i / TC, jc = i - jr * TC;
      const int ys = min(max(y0 - r + jr, 0), p.ny - 1);          // edge replication in y
      const int gc = x0 - r - 1 + jc;
      if (p.flags & NBH_WRAP_X) tl[i] = wrap_prefix<PT>(pre + (long long)ys * p.nx, gc, p.nx);
      else tl[i] = gc < 0 ? (PT)0 : pre[(long long)ys * p.nx + min(gc, p.nx - 1)];
    }
    if (LS && BOTH && k == 0) continue;                     // stage the sea plane too, then one compute pass
    __syncthreads();
    if (x < p.nx) {
      for (int ly = threadIdx.x / kCircTW; ly < TH; ly += blockDim.x / kCircTW) {
        const int y = y0 + ly;
        if (y >= p.ny) break;
        const long long pt = (long long)y * p.nx + x;
        const bool sea = LS && p.mask2d[pt] == 0;
        if (SEQ && (sea != (k == 1))) continue;
        const PT* tsel = (LS && BOTH) ? tile + (sea ? (size_t)TR * TC : 0) : tile;
        acc_t S = 0;
        if ((p.flags & NBH_WRAP_X) || (x - r - 1 >= -1 && x + r <= p.nx - 1)) {
          // every span lies inside the row (or the axis is periodic): two loads and two adds per kernel row
          const PT* row = tsel + (size_t)ly * TC + cshift + x;
#pragma unroll 4
          for (int dy = 0; dy <= 2 * r; ++dy) {
            const int w = PW ? (int)cw.w[dy] : p.wtab[dy < r ? r - dy : dy - r];
            S += (acc_t)(PT)(row[w] - row[-w - 1]);
            row += TC;
          }
        } else {
          for (int dy = -r; dy <= r; ++dy) {
            const int w = PW ? (int)cw.w[dy + r] : p.wtab[dy < 0 ? -dy : dy];
            const PT* row = tsel + (size_t)(ly + r + dy) * TC + cshift;   // indexable by global column >= -1
            const int lo = max(x - w, 0), hi = min(x + w, p.nx - 1);
            PT s2 = row[hi] - row[lo - 1];
            const int nl = lo - (x - w), nr = (x + w) - hi;             // replicated edge cells (mode="nearest")
            if (nl > 0) s2 += (PT)nl * (row[0] - row[-1]);
            if (nr > 0) s2 += (PT)nr * (row[p.nx - 1] - row[p.nx - 2]);
            S += (acc_t)s2;
          }
        }
        circ_store<PT>(p, S, 0.0, false, plane, pt, per_plane);
      }
    }
    if (SEQ) __syncthreads();                                  // the buffer is restaged for the other plane
  }
}

// Same staging, four outputs per thread: a warp covers 128 consecutive columns of a tile row, lane l
// the columns x0 + l + 32 j (j = 0..3), so the width lookup, the two span addresses and the row step
// of a kernel row are shared by four results and every shared-memory read of a warp is 32 consecutive
// words (no bank conflict): 2 loads + 3 integer operations per result and kernel row instead of ~10
// instructions.  Land / sea launches stage one plane at a time (a pass is skipped for tiles without
// points of its type), so all four results of a thread read the same tile.
constexpr int kCircTW4 = 128;

template <typename PT, bool LS, bool PW, bool BOTH>
__global__ void __launch_bounds__(256)
circ_gather_tiled4_kernel(const CircGlobalParams p, const __grid_constant__ CircWidths cw, int TH, int tiles_x,
                          int tiles_y) {
  if (p.guarded && *p.guard == 0) return;
  typedef typename CircAcc<PT>::acc_t acc_t;
  extern __shared__ __align__(16) unsigned char smem_c[];
  const int r = p.r;
  const int TC = kCircTW4 + 2 * r + 1, TR = TH + 2 * r;
  PT* tile = reinterpret_cast<PT*>(smem_c);            // [TR][TC], col j <-> global col x0 - r - 1 + j

  long long bid = blockIdx.x;
  const int tx = (int)(bid % tiles_x); bid /= tiles_x;
  const int ty = (int)(bid % tiles_y); bid /= tiles_y;
  const int b = (int)bid;
  const int x0 = tx * kCircTW4, y0 = ty * TH;
  const long long per_plane = (long long)p.ny * p.nx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long plane = p.plane0 + b;
  const int cshift = r + 1 - x0;                                // tile column of global column gc: gc + cshift
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  int xj[4];
  bool fast[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    xj[j] = x0 + lane + 32 * j;
    // every span of the column inside the row (or periodic x); columns past the row end never store
    const bool inner = wrap || (xj[j] - r - 1 >= -1 && xj[j] + r <= p.nx - 1);
    fast[j] = __all_sync(0xffffffffu, xj[j] >= p.nx || inner) != 0;
  }
  const bool fast_all = fast[0] && fast[1] && fast[2] && fast[3];
  // land / sea: both planes staged side by side when they fit (BOTH: one compute pass, every result
  // reads the plane of its own type), else one plane at a time
  constexpr bool SEQ = LS && !BOTH;
  const size_t plane_cells = (size_t)TR * TC;
#pragma unroll 1
  for (int k = 0; k < (LS ? 2 : 1); ++k) {
    PT* tl = tile + ((LS && BOTH && k == 1) ? plane_cells : 0);
    if (SEQ) {
      bool mine = false;
      for (int ly = warp; ly < TH; ly += n_warps) {
        const int y = y0 + ly;
        if (y >= p.ny) break;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (xj[j] < p.nx) mine = mine || ((p.mask2d[(long long)y * p.nx + xj[j]] == 0) == (k == 1));
      }
      if (!__syncthreads_or(mine ? 1 : 0)) continue;
    }
    const PT* pre = reinterpret_cast<const PT*>(k ? p.prefix2 : p.prefix) + (long long)b * per_plane;
    // stage: a warp per tile row, lanes along the row (no index division, coalesced)
    for (int jr = warp; jr < TR; jr += n_warps) {
      const int ys = min(max(y0 - r + jr, 0), p.ny - 1);          // edge replication in y
      const PT* src = pre + (long long)ys * p.nx;
      PT* dst = tl + (size_t)jr * TC;
#pragma unroll 2
      for (int jc = lane; jc < TC; jc += 32) {
        const int gc = x0 - r - 1 + jc;
        if (wrap) dst[jc] = wrap_prefix<PT>(src, gc, p.nx);
        else dst[jc] = gc < 0 ? (PT)0 : src[min(gc, p.nx - 1)];
      }
    }
    if (LS && BOTH && k == 0) continue;                       // stage the sea plane too, then one compute pass
    __syncthreads();
    for (int ly = warp; ly < TH; ly += n_warps) {
      const int y = y0 + ly;
      if (y >= p.ny) break;
      bool want[4];
      int poff[4];                                            // element offset of the plane a result reads
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        want[j] = xj[j] < p.nx;
        poff[j] = 0;
        if (LS && want[j]) {
          const bool sea = p.mask2d[(long long)y * p.nx + xj[j]] == 0;
          if (SEQ) want[j] = sea == (k == 1);
          else poff[j] = sea ? (int)plane_cells : 0;
        }
      }
      // sequential land / sea passes: a stretch of 128 columns without a point of this pass's type is skipped
      if (SEQ && !__any_sync(0xffffffffu, want[0] || want[1] || want[2] || want[3])) continue;
      acc_t S[4] = {0, 0, 0, 0};
      if (fast_all) {
        const PT* rowp = tile + (size_t)ly * TC + cshift + x0 + lane;   // column x0 + lane of the first kernel row
#pragma unroll 2
        for (int dy = 0; dy <= 2 * r; ++dy) {
          const int w = PW ? (int)cw.w[dy] : p.wtab[dy < r ? r - dy : dy - r];
          const PT* hi = rowp + w;
          const PT* lo = rowp - w - 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) S[j] += (acc_t)(PT)(hi[32 * j + poff[j]] - lo[32 * j + poff[j]]);
          rowp += TC;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = xj[j];
          if (x >= p.nx) continue;
          if (fast[j]) {
            const PT* rowp = tile + (size_t)ly * TC + cshift + x + poff[j];
#pragma unroll 4
            for (int dy = 0; dy <= 2 * r; ++dy) {
              const int w = PW ? (int)cw.w[dy] : p.wtab[dy < r ? r - dy : dy - r];
              S[j] += (acc_t)(PT)(rowp[w] - rowp[-w - 1]);
              rowp += TC;
            }
          } else {
            for (int dy = -r; dy <= r; ++dy) {
              const int w = PW ? (int)cw.w[dy + r] : p.wtab[dy < 0 ? -dy : dy];
              const PT* row = tile + (size_t)(ly + r + dy) * TC + cshift + poff[j];   // indexable by global column >= -1
              const int lo = max(x - w, 0), hi = min(x + w, p.nx - 1);
              PT s2 = row[hi] - row[lo - 1];
              const int nl = lo - (x - w), nr = (x + w) - hi;             // replicated edge cells (mode="nearest")
              if (nl > 0) s2 += (PT)nl * (row[0] - row[-1]);
              if (nr > 0) s2 += (PT)nr * (row[p.nx - 1] - row[p.nx - 2]);
              S[j] += (acc_t)s2;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (want[j]) circ_store<PT>(p, S[j], 0.0, false, plane, (long long)y * p.nx + xj[j], per_plane);
    }
    if (SEQ) __syncthreads();                                  // the buffer is restaged for the other plane
  }
}

static int circ_tile4_rows(int r, int nx, size_t elem, int n_planes, size_t* smem) {
  if (nx < 3) return 0;
  for (int th : {64, 32, 16}) {
    const size_t cells = (size_t)(th + 2 * r) * (kCircTW4 + 2 * r + 1);
    const size_t bytes = (size_t)n_planes * cells * elem;
    if (bytes <= 72u * 1024 && cells <= (size_t)12 * th * kCircTW4) { *smem = bytes; return th; }
  }
  return 0;
}

// pick the tallest tile that fits; 0 = use the direct gather.  Measured on B200: tiles pay while
// several blocks fit on an SM and the staged window is well reused (a 200 KB tile per SM at r = 81
// was 1.7 x SLOWER than the direct gather from L1 / L2).  n_tiles = 2: both land / sea planes at once.
static int circ_tile_rows(int r, int nx, size_t elem, int n_tiles, size_t* smem) {
  if (nx < 3) return 0;
  for (int th : {64, 32, 16}) {
    const size_t cells = (size_t)(th + 2 * r) * (kCircTW + 2 * r + 1);
    const size_t bytes = (size_t)n_tiles * cells * elem;
    if (bytes <= 72u * 1024 && cells <= (size_t)12 * th * kCircTW) { *smem = bytes; return th; }
  }
  return 0;
}

template <typename PT, bool LS, bool PW>
static int launch_circ_gather_t(const CircGlobalParams& g, const CircWidths& cw, int ny, int nx, bool allow_tile,
                                cudaStream_t st) {
  size_t smem = 0;
  int th = 0;
  bool both = false;
  if (!g.cprefix && allow_tile) {
    // land / sea: both planes side by side while a tile of at least 32 rows fits, else sequential passes
    // land / sea launches gather directly (circ_gather4_kernel) whenever that kernel applies: measured
    // on 240 UK slices against the tiles with both planes side by side (r = 5: 4.9 / 5.2 ms, r = 9:
    // 6.2 / 7.0 ms) and against staging the planes one after the other (r = 27: 12.5 / 16.5 ms)
    const bool ls_direct = LS && nx >= 128 && !(g.flags & NBH_WRAP_X);
    int th4 = (LS && !ls_direct) ? circ_tile4_rows(g.r, nx, sizeof(PT), 2, &smem) : 0;
    const bool both4 = th4 >= 32;
    if (!both4) th4 = ls_direct ? 0 : circ_tile4_rows(g.r, nx, sizeof(PT), 1, &smem);
    if (th4 > 0) {
      const int tiles_x = (nx + kCircTW4 - 1) / kCircTW4, tiles_y = (ny + th4 - 1) / th4;
      const long long blocks = (long long)tiles_x * tiles_y * g.n_batch;
      NBH_REQUIRE(blocks < (1LL << 31), "too many blocks");
      if (both4) {
        auto kern = circ_gather_tiled4_kernel<PT, LS, PW, true>;
        NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, 256, smem, st>>>(g, cw, th4, tiles_x, tiles_y);
      } else {
        auto kern = circ_gather_tiled4_kernel<PT, LS, PW, false>;
        NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, 256, smem, st>>>(g, cw, th4, tiles_x, tiles_y);
      }
      return 0;
    }
  }
  if (!g.cprefix && allow_tile && !(LS && nx >= 128 && !(g.flags & NBH_WRAP_X))) {
    if (LS) {
      // both planes side by side while a tall tile fits; else one buffer, the planes one after the other
      th = circ_tile_rows(g.r, nx, sizeof(PT), 2, &smem);
      both = th >= 32;
      if (!both) th = 0;
    }
    if (th == 0) th = circ_tile_rows(g.r, nx, sizeof(PT), 1, &smem);
  }
  if (th > 0) {
    const int tiles_x = (nx + kCircTW - 1) / kCircTW, tiles_y = (ny + th - 1) / th;
    const long long blocks = (long long)tiles_x * tiles_y * g.n_batch;
    NBH_REQUIRE(blocks < (1LL << 31), "too many blocks");
    if (both) {
      auto kern = circ_gather_tiled_kernel<PT, LS, PW, true>;
      NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(unsigned)blocks, 256, smem, st>>>(g, cw, th, tiles_x, tiles_y);
    } else {
      auto kern = circ_gather_tiled_kernel<PT, LS, PW, false>;
      NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(unsigned)blocks, 256, smem, st>>>(g, cw, th, tiles_x, tiles_y);
    }
  } else {
    if (!g.cprefix && !(g.flags & NBH_WRAP_X) && nx >= 128) {
      const int tiles_x = (nx + 127) / 128, tiles_y = (ny + kCircG4Rows - 1) / kCircG4Rows;
      const long long blocks = (long long)tiles_x * tiles_y * g.n_batch;
      NBH_REQUIRE(blocks < (1LL << 31), "too many blocks");
      circ_gather4_kernel<PT, LS, PW><<<(unsigned)blocks, kCircG4Rows * 32, 0, st>>>(g, cw, tiles_x, tiles_y);
    } else {
      const long long outs = (long long)g.n_batch * ny * nx;
      circ_gather_kernel<PT, LS, PW><<<(unsigned)((outs + 255) / 256), 256, 0, st>>>(g, cw);
    }
  }
  return 0;
}

template <typename PT>
static int launch_circ_pass(const CircGlobalParams& g, const CircWidths& cw, int ny, int nx, bool allow_tile,
                            cudaStream_t st) {
  const bool ls = (g.flags & NBH_LAND_SEA) != 0, pw = g.r <= kCircParamRadius;
  const long long rows = (long long)g.n_batch * ny;
  const unsigned pgrid = (unsigned)((rows + 7) / 8);
  if (ls) circ_prefix_kernel<PT, true><<<pgrid, 256, 0, st>>>(g);
  else circ_prefix_kernel<PT, false><<<pgrid, 256, 0, st>>>(g);
  if (ls) return pw ? launch_circ_gather_t<PT, true, true>(g, cw, ny, nx, allow_tile, st)
                    : launch_circ_gather_t<PT, true, false>(g, cw, ny, nx, allow_tile, st);
  return pw ? launch_circ_gather_t<PT, false, true>(g, cw, ny, nx, allow_tile, st)
            : launch_circ_gather_t<PT, false, false>(g, cw, ny, nx, allow_tile, st);
}

// area of the own surface type from the land area: land points keep A_land, sea points get
// sum(K) - A_land
__global__ void circ_area_own_kernel(float* area, const uint8_t* land, float area_const, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && land[i] == 0) area[i] = area_const - area[i];
}

// area_sum for an external mask at large radius, from the prefix of the mask itself
__global__ void circ_mask_to_float_kernel(const uint8_t* mask, float* as_float, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) as_float[i] = mask[i] ? 1.f : 0.f;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kCircSmemMaxRadius = 40;          // beyond this the global-prefix path is used
constexpr size_t kCircPrefixBudget = 256u << 20;  // bytes of float64 prefixes per batch (L2-friendly)

// unweighted discs use the prefix + gather path at every radius (measured on B200: it beats the
// shared-memory ring kernel from radius 1); weighted discs the ring kernel
static bool use_global_path(const NbhDesc& d) {
  if (d.cells > kCircSmemMaxRadius) return true;
  return !(d.flags & NBH_WEIGHTED);
}

bool circular_native_land_sea(const NbhDesc& d) { return use_global_path(d) && !(d.flags & NBH_WEIGHTED); }

struct CircWorkspace {
  double* ktab;
  int* wtab;
  float* area_const;
  float* area;
  double* prefix;
  double* prefix2;
  double* cprefix;
  float* maskf;
  int* guard;
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
  w->guard = reinterpret_cast<int*>(base + off); off = align_up(off + 4, 256);
  w->prefix = w->prefix2 = w->cprefix = nullptr; w->maskf = nullptr; w->batch = 0;
  if (use_global_path(d)) {
    const size_t per_plane = (size_t)d.ny * d.nx * 8;
    long long batch = (long long)(kCircPrefixBudget / per_plane);
    if (batch < 1) batch = 1;
    if (batch > d.n_planes) batch = d.n_planes;
    w->batch = (int)batch;
    w->prefix = reinterpret_cast<double*>(base + off); off = align_up(off + per_plane * batch, 256);
    if (d.mask_nd) { w->cprefix = reinterpret_cast<double*>(base + off); off = align_up(off + per_plane * batch, 256); }
    if (d.flags & NBH_LAND_SEA) { w->prefix2 = reinterpret_cast<double*>(base + off); off = align_up(off + per_plane * batch, 256); }
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
    const bool land_sea = (d->flags & NBH_LAND_SEA) != 0;
    const bool sum_only = (d->flags & NBH_SUM_ONLY) != 0;
    CircGlobalParams g;
    memset(&g, 0, sizeof(g));
    ThrSet ts;
    int tm = 0;
    if (fill_thresholds(*d, &ts, &tm)) return 1;
    g.in = in; g.out = out; g.out_mask = out_mask; g.mask2d = d->mask2d; g.mask_nd = d->mask_nd;
    g.area = w.area; g.wtab = w.wtab; g.prefix = w.prefix; g.prefix2 = w.prefix2; g.cprefix = w.cprefix;
    g.status = status; g.guard = w.guard;
    g.plane_stride = d->plane_stride; g.member_stride = d->member_stride;
    g.n_members = d->n_members; g.ny = d->ny; g.nx = d->nx; g.r = r; g.flags = d->flags;
    g.n_thr = max(d->n_thresholds, 1);
    g.inv_members = sum_only ? 1.0 : 1.0 / (double)d->n_members;
    g.fx_scale = 1.f; g.fx_inv = 1.0;
    CircWidths cw;
    memset(&cw, 0, sizeof(cw));
    {
      double tot = 0.0;
      const long long r2 = (long long)r * r;
      for (int dy = -r; dy <= r; ++dy) {
        long long wdt = (long long)floor(sqrt((double)(r2 - (long long)dy * dy)));
        while ((wdt + 1) * (wdt + 1) + (long long)dy * dy <= r2) ++wdt;
        while (wdt > 0 && wdt * wdt + (long long)dy * dy > r2) --wdt;
        tot += 2.0 * (double)wdt + 1.0;
        if (r <= kCircParamRadius) cw.w[dy + r] = (short)wdt;
      }
      g.area_const = (float)tot;
    }
    // fixed-point prefixes: span differences stay below 2^32
    int fx_shift = 0;
    while (fx_shift < 24 && (double)nb * d->n_members * (double)(1u << (fx_shift + 1)) < 4294967296.0) ++fx_shift;
    const bool fx_ok = !masknd && !sum_only && fx_shift >= 16;
    NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
    NBH_CHECK_CUDA(cudaMemsetAsync(w.guard, 0, sizeof(int), st));
    circ_tables_kernel<<<1, 256, 0, st>>>(w.ktab, w.wtab, w.area_const, r, 0);
    const long long npts = (long long)d->ny * d->nx;
    if (d->mask2d && !masknd && !sum_only) {
      // area_sum = the same gather applied to the mask (one plane of 0 / 1, exact in fixed point)
      circ_mask_to_float_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(d->mask2d, w.maskf, npts);
      CircGlobalParams a = g;
      a.in = w.maskf; a.out = w.area; a.out_mask = nullptr; a.mask2d = nullptr; a.mask_nd = nullptr;
      a.cprefix = nullptr; a.prefix2 = nullptr; a.plane_stride = npts; a.member_stride = 0; a.n_members = 1;
      a.n_thr = 1; a.t = 0; a.plane0 = 0; a.n_batch = 1; a.flags = NBH_SUM_ONLY | (d->flags & NBH_WRAP_X);
      a.thr.mode = 0; a.inv_members = 1.0;
      if (launch_circ_pass<unsigned>(a, cw, d->ny, d->nx, true, st)) return 1;
      if (land_sea)
        circ_area_own_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(w.area, d->mask2d, g.area_const, npts);
    }
    ProfileScope prof(st, "nbh::circ_prefix_kernel + nbh::circ_gather_kernel (row-span decomposition)");
    for (int t = 0; t < g.n_thr; ++t) {
      g.t = t;
      g.thr = ts.t[d->n_thresholds ? t : 0];
      if (!d->n_thresholds) g.thr.mode = 0;
      for (long long p0 = 0; p0 < d->n_planes; p0 += w.batch) {
        g.plane0 = p0;
        g.n_batch = (int)min((long long)w.batch, d->n_planes - p0);
        if (fx_ok) {
          CircGlobalParams f = g;
          f.guarded = 0;                                       // the fixed-point pass itself always runs
          f.fx_scale = (float)(1u << fx_shift); f.fx_inv = 1.0 / (double)(1u << fx_shift);
          f.check_range = d->n_thresholds == 0 ? 1 : 0;      // raw data: speculative, verified by the pass itself
          if (launch_circ_pass<unsigned>(f, cw, d->ny, d->nx, true, st)) return 1;
          if (!f.check_range) continue;
          g.guarded = 1;                                       // float64 re-run only if a value left [0, 1]
        }
        if (launch_circ_pass<double>(g, cw, d->ny, d->nx, true, st)) return 1;
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
  const long long want = 4LL * sm_count() * 2;
  int ysegs = (int)max(1LL, min((long long)d->ny / max(16, 4 * r), (want + tiles_xy - 1) / tiles_xy));
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
