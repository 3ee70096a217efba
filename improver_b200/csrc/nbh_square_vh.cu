// Square neighbourhood of single-member slices for radii beyond the box kernel's shuffle window
// (9 .. 63 cells): two radius-independent passes over exact fixed-point values.
//
//   V pass  (square_v_kernel): a thread owns two adjacent columns of a y segment and marches down
//           the rows keeping the vertical window sum of the quantised (thresholded, masked) values,
//           V += q(y + r) - q(y - r - 1).  The entering row comes straight from global memory
//           (coalesced across the block, eight rows in flight per thread), the leaving value from a
//           shared-memory ring of the 2r + 1 quantised rows of the window (8 bytes per thread and
//           row, every thread its own column: no barrier).  V is written to an intermediate uint32
//           plane.
//   H pass  (square_h_kernel): a warp owns a row of the intermediate plane: staged in shared
//           memory, turned into its inclusive prefix (lane-local runs + a warp scan), every result
//           is the difference of two prefix values, normalised like the box kernel (correctly
//           rounded quotient) and stored coalesced.
//
// The intermediate plane (4 B / point) goes through HBM: chunks small enough to keep it in L2
// (<= 64 MB = 16 UK slices) were measured and lose more to the small grids (192 blocks per V launch,
// latency-bound) than they save in traffic (2.2 ms against the numbers in DESIGN.md for 240 slices);
// the launch is only cut when the plane would exceed 1 GB.  DRAM traffic ~ 20 B / point (input read,
// leaving rows mostly from L2, intermediate written and read, output written).  Cost does not
// depend on the radius.  Same arithmetic contract as square_box_kernel: unsigned 32-bit sums with
// (2r+1)^2 * 2^s < 2^32, band-origin independent, slice extremes recorded for the guarded clamp /
// float64 re-run (box_finish).
#include "nbh_square_box.cuh"

namespace nbh {

constexpr int kVhSegRows = 256;          // rows per y segment of the V pass (halo 2r rows re-read per segment)
constexpr size_t kVhChunkBytes = (size_t)1 << 30;   // intermediate plane per chunk of planes (see the header comment)
constexpr int kVhRows = 8;               // output rows per step of the V pass: 2 * kVhRows row loads in flight per thread

struct VhParams {
  const float* in;
  unsigned* vplane;           // [chunk planes][ny][nx]
  float* out;
  uint8_t* out_mask;
  const uint8_t* mask2d;
  const float* cnt_plane;
  BoxStats* stats;            // [n_units]
  unsigned* edge_flags;       // [n_units] bits of the edge bands that hold a value != 1 (trimming fix-up)
  int32_t* status;
  long long plane_stride;
  long long unit0;            // first unit (plane * n_thr + t) of this chunk
  int n_units_chunk;
  int ny, nx, r, n_thr, n_ysegs, flags, fx_shift;
  ThrSet thr;
};

template <int TM, bool M2D>
__global__ void __launch_bounds__(256)
square_v_kernel(const VhParams p) {
  extern __shared__ uint2 ring[];                     // [2r+1][blockDim.x] quantised values of the rows inside the window
  const int u_local = blockIdx.x / p.n_ysegs, seg = blockIdx.x - u_local * p.n_ysegs;
  const long long unit = p.unit0 + u_local;
  const long long plane = unit / p.n_thr;
  const int t = (int)(unit - plane * p.n_thr);
  const ThrDev& thr = p.thr.t[TM ? t : 0];
  const int x = 2 * (blockIdx.y * blockDim.x + threadIdx.x);
  const bool active = x < p.nx;                       // nx is even: both columns valid or none
  const int y0 = seg * kVhSegRows, y1 = min(p.ny, y0 + kVhSegRows);
  const int r = p.r, nb = 2 * r + 1;
  const float scale = (float)(1u << p.fx_shift);
  const float* sl = p.in + plane * p.plane_stride + x;
  const uint8_t* mp = M2D ? p.mask2d + x : nullptr;
  unsigned* vp = p.vplane + (long long)u_local * p.ny * p.nx + x;
  const bool want_stats = !(p.flags & NBH_SUM_ONLY) || TM == 0;
  float dmin = __int_as_float(0x7f800000), dmax = __int_as_float(0xff800000);
  unsigned V0 = 0u, V1 = 0u;
  // edge bands of the trimming fix-up (nbhood.py:353-382): bit0 top, bit1 bottom, bit2 left, bit3 right
  // band holds a value != 1 (masked points count); the column bits of this thread's two columns
  const bool want_flags = !(p.flags & NBH_SUM_ONLY);
  const unsigned ONE = 1u << p.fx_shift;
  const unsigned cb0 = (x <= r ? 4u : 0u) | (x >= p.nx - 1 - r ? 8u : 0u);
  const unsigned cb1 = (x + 1 <= r ? 4u : 0u) | (x + 1 >= p.nx - 1 - r ? 8u : 0u);
  unsigned ne_bits = 0u;

  for (int k = 0; k < nb; ++k) ring[k * blockDim.x + threadIdx.x] = make_uint2(0u, 0u);
  // (every thread only ever touches its own ring column: no barrier)

  // Row yy enters the window of output row yy - r and takes the ring slot of row yy - 2r - 1,
  // which leaves it: V += q(yy) - ring[slot]; ring[slot] = q(yy).  kVhRows row loads in flight.
  constexpr int NR = kVhRows;
  const int yy_first = y0 - r, yy_end = y1 + r;         // rows outside the domain are zeros
  int slot = 0;
  float2 d[NR], dn[NR];
  uchar2 m[NR], mn[NR];
  // the loads of batch b + 1 are issued before batch b is consumed (a warp alone on its scheduler
  // would otherwise alternate between waiting for memory and computing)
  auto load_batch = [&](int yy, float2* dd, uchar2* mm) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int y = yy + j;
      dd[j] = make_float2(0.f, 0.f);
      mm[j] = make_uchar2(0, 0);
      if (active && y >= 0 && y < p.ny && y < yy_end) {
        const long long pt = (long long)y * p.nx;
        dd[j] = __ldg(reinterpret_cast<const float2*>(sl + pt));
        mm[j] = M2D ? __ldg(reinterpret_cast<const uchar2*>(mp + pt)) : make_uchar2(1, 1);
      }
    }
  };
  load_batch(yy_first, d, m);
  for (int yy = yy_first; yy < yy_end; yy += NR) {
    load_batch(yy + NR, dn, mn);
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int y = yy + j;
      if (y < yy_end) {
        unsigned q0 = 0u, q1 = 0u;
        if (active && y >= 0 && y < p.ny) {
          dmax = max3_nan(dmax, d[j].x, d[j].y);          // NaN propagates (nbhood.py:167-168)
          if (want_stats) dmin = min3f(dmin, d[j].x, d[j].y);
          float a = truth_value_sum<TM>(d[j].x, thr), b = truth_value_sum<TM>(d[j].y, thr);
          a = m[j].x ? a : 0.f;
          b = m[j].y ? b : 0.f;
          q0 = __float2uint_rn(a * scale);
          q1 = __float2uint_rn(b * scale);
          const unsigned rb = want_flags ? ((y <= r ? 1u : 0u) | (y >= p.ny - 1 - r ? 2u : 0u)) : 0u;
          if (want_flags && (rb | cb0 | cb1)) {
            if (q0 != ONE) ne_bits |= rb | cb0;
            if (q1 != ONE) ne_bits |= rb | cb1;
          }
        }
        uint2* cell = ring + slot * blockDim.x + threadIdx.x;
        const uint2 old = *cell;
        *cell = make_uint2(q0, q1);
        V0 += q0 - old.x; V1 += q1 - old.y;
        slot = slot + 1 == nb ? 0 : slot + 1;
        const int yo = y - r;
        if (active && yo >= y0) *reinterpret_cast<uint2*>(vp + (long long)yo * p.nx) = make_uint2(V0, V1);
      }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) { d[j] = dn[j]; m[j] = mn[j]; }
  }
  if (want_flags) {
    const unsigned bits = __reduce_or_sync(0xffffffffu, ne_bits);
    if ((threadIdx.x & 31) == 0 && bits) atomicOr(p.edge_flags + unit, bits);
  }
  // slice extremes of the raw data
  const bool nan = __isnanf(dmax);
  if (__any_sync(0xffffffffu, nan) && (threadIdx.x & 31) == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  if (want_stats) {
    int kmin = (__isnanf(dmin) || dmin == __int_as_float(0x7f800000)) ? INT_MAX : box_float_key(dmin);
    int kmax = (nan || dmax == __int_as_float(0xff800000)) ? INT_MIN : box_float_key(dmax);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&p.stats[unit].in_min, kmin);
      atomicMax(&p.stats[unit].in_max, kmax);
    }
  }
}

constexpr int kVhHWarps = 8;

template <bool M2D>
__global__ void __launch_bounds__(kVhHWarps * 32)
square_h_kernel(const VhParams p, int row_pitch) {
  extern __shared__ unsigned smem_h[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* P = smem_h + (size_t)warp * row_pitch;          // [nx] values, then inclusive prefix in place
  const int r = p.r, nx = p.nx, nb = 2 * r + 1;
  const int per = (nx + 31) / 32;                            // consecutive elements per lane in the prefix step
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
  const float scale = (float)(1u << p.fx_shift), inv_scale = 1.0f / scale;
  const long long n_rows = (long long)p.n_units_chunk * p.ny;
  float omin = __int_as_float(0x7f800000), omax = __int_as_float(0xff800000);
  long long cur_unit = -1;
  for (long long row = (long long)blockIdx.x * kVhHWarps + warp; row < n_rows; row += (long long)gridDim.x * kVhHWarps) {
    const int u_local = (int)(row / p.ny), y = (int)(row - (long long)u_local * p.ny);
    const long long unit = p.unit0 + u_local;
    if (unit != cur_unit) {
      if (cur_unit >= 0 && !sum_only) {
        int komin = omin == __int_as_float(0x7f800000) ? INT_MAX : box_float_key(omin);
        int komax = omax == __int_as_float(0xff800000) ? INT_MIN : box_float_key(omax);
        komin = __reduce_min_sync(0xffffffffu, komin); komax = __reduce_max_sync(0xffffffffu, komax);
        if (lane == 0) { atomicMin(&p.stats[cur_unit].out_min, komin); atomicMax(&p.stats[cur_unit].out_max, komax); }
      }
      cur_unit = unit; omin = __int_as_float(0x7f800000); omax = __int_as_float(0xff800000);
    }
    const unsigned* vrow = p.vplane + ((long long)u_local * p.ny + y) * nx;
    for (int i = 2 * lane; i < nx; i += 64) {               // rows are 8-byte aligned (nx is even)
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(vrow + i));
      P[i] = v.x; P[i + 1] = v.y;
    }
    __syncwarp();
    // inclusive prefix: lane-local runs of `per` elements (bank-conflict free for odd `per`) taken in
    // groups of eight independent shared-memory reads, warp scan of the totals
    const int i0 = lane * per, i1 = min(nx, i0 + per);
    unsigned tot = 0u;
    for (int c = i0; c < i1; c += 8) {
      unsigned v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = c + k < i1 ? P[c + k] : 0u;
      tot += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    unsigned incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    unsigned run = incl - tot;
    for (int c = i0; c < i1; c += 8) {
      unsigned v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = c + k < i1 ? P[c + k] : 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) { run += v[k]; if (c + k < i1) P[c + k] = run; }
    }
    __syncwarp();
    const int rows_in = min(y + r, p.ny - 1) - max(y - r, 0) + 1;
    const float nrow = (float)rows_in * scale;
    const long long obase = unit * (long long)p.ny * nx + (long long)y * nx;
    float* orow = p.out + obase;
    // one result: S / nf as a correctly rounded quotient (reciprocal + two FMAs), NaN where nothing is valid
    auto finish = [&](int xo, unsigned S, float nf, float rc) {
      const float Sf = (float)S;
      float o;
      if (sum_only) {
        o = Sf * inv_scale;
      } else {
        const float q0 = Sf * rc;
        const float rem = __fmaf_rn(-q0, nf, Sf);
        o = (M2D && nf == 0.f) ? __int_as_float(0x7fc00000) : __fmaf_rn(rem, rc, q0);
        omax = fmaxf(omax, o); omin = fminf(omin, o);
      }
      orow[xo] = o;
      if (p.out_mask) p.out_mask[obase + xo] = (M2D && p.mask2d[(long long)y * nx + xo] == 0) ? 1 : 0;
    };
    // interior columns: the whole window inside the row, one divisor for all of them
    {
      const float nf_in = nrow * (float)nb;
      float rc_in;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc_in) : "f"(nf_in));
      const float* crow = M2D ? p.cnt_plane + (long long)y * nx : nullptr;
#pragma unroll 4
      for (int xo = r + 1 + lane; xo < nx - r && nx >= 2 * r + 2; xo += 32) {
        const unsigned S = P[xo + r] - P[xo - r - 1];
        float nf = nf_in, rc = rc_in;
        if (M2D && !sum_only) {
          nf = __ldg(crow + xo) * scale;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
        }
        finish(xo, S, nf, rc);
      }
    }
    // the columns whose window is cut by the row ends (all of them in a row narrower than two windows)
    auto cut_column = [&](int xo) {
      const int hi = min(xo + r, nx - 1), lo = xo - r - 1;
      const unsigned S = P[hi] - (lo >= 0 ? P[lo] : 0u);
      float nf = 1.f, rc = 1.f;
      if (!sum_only) {
        if (M2D) nf = __ldg(p.cnt_plane + (long long)y * nx + xo) * scale;
        else nf = nrow * (float)(hi - max(xo - r, 0) + 1);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
      }
      finish(xo, S, nf, rc);
    };
    if (nx >= 2 * r + 2) {
      for (int e = lane; e < 2 * r + 1; e += 32) cut_column(e <= r ? e : nx - 2 * r - 1 + e);
    } else {
      for (int xo = lane; xo < nx; xo += 32) cut_column(xo);
    }
    __syncwarp();
    (void)nb;
  }
  if (cur_unit >= 0 && !sum_only) {
    int komin = omin == __int_as_float(0x7f800000) ? INT_MAX : box_float_key(omin);
    int komax = omax == __int_as_float(0xff800000) ? INT_MIN : box_float_key(omax);
    komin = __reduce_min_sync(0xffffffffu, komin); komax = __reduce_max_sync(0xffffffffu, komax);
    if (lane == 0) { atomicMin(&p.stats[cur_unit].out_min, komin); atomicMax(&p.stats[cur_unit].out_max, komax); }
  }
}

// ---- host ----
bool plan_square_vh(const NbhDesc& d, const float* in, int* fx_shift) {
  const int r = d.cells, nb = 2 * r + 1;
  if (d.n_members != 1 || r < 1 || r > 63) return false;
  if ((d.flags & NBH_WRAP_X) || d.mask_nd) return false;
  if ((d.nx & 1) || d.nx < 2 || (reinterpret_cast<uintptr_t>(in) & 7) || (d.plane_stride & 1)) return false;
  if ((size_t)d.nx * 4 * kVhHWarps > 200u * 1024) return false;
  int shift = 0;
  while (shift < 24 && (double)nb * nb * (double)(1u << (shift + 1)) < 4294967296.0) ++shift;
  if (shift < 18) return false;
  *fx_shift = shift;
  return true;
}

size_t vh_workspace_bytes(const NbhDesc& d) {
  const size_t per_plane = (size_t)d.ny * d.nx * 4;
  const size_t units = (size_t)d.n_planes * (size_t)max(d.n_thresholds, 1);
  size_t chunk = max((size_t)1, kVhChunkBytes / per_plane);
  if (chunk > units) chunk = units;
  return chunk * per_plane + 256;
}

template <int TM, bool M2D>
static int launch_vh_t(VhParams p, long long n_units, cudaStream_t st) {
  const size_t per_plane = (size_t)p.ny * p.nx * 4;
  const long long chunk = max((long long)1, (long long)(kVhChunkBytes / per_plane));
  const int row_pitch = (p.nx + 1) | 1;
  const size_t smem = (size_t)kVhHWarps * row_pitch * 4;
  static bool configured = false;
  if (!configured) {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(square_h_kernel<M2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  // V pass: a block covers 2 * vthreads columns; the ring of the window rows (8 B per thread and row)
  // bounds the block so that two blocks fit an SM
  const int nbw = 2 * p.r + 1;
  int vcap = 256;
  while (vcap > 32 && (size_t)nbw * vcap * 8 > (size_t)100 * 1024) vcap -= 32;
  const int xchunks = (p.nx / 2 + vcap - 1) / vcap;
  const int vthreads = ((p.nx / 2 + xchunks - 1) / xchunks + 31) / 32 * 32;
  const size_t vsmem = (size_t)nbw * vthreads * 8;
  static bool vconfigured = false;
  if (!vconfigured) {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(square_v_kernel<TM, M2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    vconfigured = true;
  }
  for (long long u0 = 0; u0 < n_units; u0 += chunk) {
    p.unit0 = u0;
    p.n_units_chunk = (int)min(chunk, n_units - u0);
    dim3 vgrid((unsigned)(p.n_units_chunk * p.n_ysegs), (unsigned)xchunks);
    square_v_kernel<TM, M2D><<<vgrid, vthreads, vsmem, st>>>(p);
    const long long rows = (long long)p.n_units_chunk * p.ny;
    const unsigned hgrid = (unsigned)min((rows + kVhHWarps - 1) / kVhHWarps, (long long)sm_count() * 6);
    square_h_kernel<M2D><<<hgrid, kVhHWarps * 32, smem, st>>>(p, row_pitch);
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_square_vh(const SqParams& sp, const NbhDesc& d, int tm, int fx_shift, BoxStats* stats, unsigned* vplane,
                     cudaStream_t st) {
  VhParams p;
  memset(&p, 0, sizeof(p));
  p.in = sp.in; p.vplane = vplane; p.out = sp.out; p.out_mask = sp.out_mask; p.mask2d = sp.mask2d;
  p.cnt_plane = sp.cnt_plane; p.stats = stats; p.edge_flags = sp.edge_flags; p.status = sp.status; p.plane_stride = sp.plane_stride;
  p.ny = sp.ny; p.nx = sp.nx; p.r = sp.r; p.n_thr = sp.n_thr; p.flags = sp.flags; p.fx_shift = fx_shift;
  p.n_ysegs = (sp.ny + kVhSegRows - 1) / kVhSegRows;
  p.thr = sp.thr;
  const long long n_units = sp.n_planes * sp.n_thr;
  const bool m2d = d.mask2d != nullptr;
  const int tmk = tm == 4 ? 2 : tm;
  switch ((tmk << 1) | (m2d ? 1 : 0)) {
    case 0: return launch_vh_t<0, false>(p, n_units, st);
    case 1: return launch_vh_t<0, true>(p, n_units, st);
    case 2: return launch_vh_t<1, false>(p, n_units, st);
    case 3: return launch_vh_t<1, true>(p, n_units, st);
    case 4: return launch_vh_t<2, false>(p, n_units, st);
    case 5: return launch_vh_t<2, true>(p, n_units, st);
    case 10: return launch_vh_t<5, false>(p, n_units, st);
    case 11: return launch_vh_t<5, true>(p, n_units, st);
    default: set_error("unsupported two-pass kernel variant"); return 1;
  }
}

}  // namespace nbh
