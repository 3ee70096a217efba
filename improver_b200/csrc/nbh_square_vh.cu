// Square neighbourhood of single-member slices for radii beyond the box kernel's shuffle window
// (9 .. 63 cells) and for a periodic x axis at every radius: two radius-independent passes over
// exact fixed-point values.
//
//   V pass  (square_v_kernel): a thread owns two adjacent columns of a y segment and marches down
//           the rows keeping the vertical window sum of the quantised (thresholded, masked) values,
//           V += q(y + r) - q(y - r - 1).  The entering rows come straight from global memory
//           (coalesced across the block; sixteen rows per batch, the next batch requested before
//           the current one is consumed), the leaving value from a shared-memory ring of the 2r + 1
//           quantised rows of the window (8 bytes per thread and row, every thread its own column:
//           no barrier).  It also records the slice extremes and the edge-band flags of the
//           trimming fix-up.  V is written to an intermediate uint32 plane.
//   H pass  (square_h_kernel): a warp owns a row of the intermediate plane: copied to shared
//           memory with cp.async (the next row in flight while the current one is worked on),
//           turned into its inclusive prefix in place (four consecutive values per lane and
//           128-column step: lane scan, warp scan of the totals, 16-byte accesses), every result is
//           the difference of two prefix values (three for the columns of a periodic row whose
//           window wraps), normalised like the box kernel (correctly rounded quotient) and stored
//           coalesced.
//
// The intermediate plane (4 B / point) goes through HBM: chunks small enough to keep it in L2
// (<= 64 MB = 16 UK slices) were measured and lose more to the small grids than they save in
// traffic; the launch is only cut when the plane would exceed 1 GB.  DRAM traffic ~ 17 B / point
// (input read incl. 5 % halo rows, intermediate written and read, output written).  Cost does not
// depend on the radius.  Same arithmetic contract as square_box_kernel: unsigned 32-bit sums with
// (2r+1)^2 * 2^s < 2^32, band-origin independent, slice extremes recorded for the guarded clamp /
// float64 re-run (box_finish).
#include "nbh_square_box.cuh"

namespace nbh {

constexpr int kVhSegRowsMax = 512;       // rows per y segment of the V pass (2r halo rows are re-read per segment); halved while the grid is small
constexpr size_t kVhChunkBytes = (size_t)1 << 30;   // intermediate plane per chunk of planes (see the header comment)
constexpr int kVhRows = 16;              // rows per batch of the V pass (a batch in flight while the previous one is consumed)

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
  int ny, nx, r, n_thr, n_ysegs, seg_rows, flags, fx_shift;
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
  const int y0 = seg * p.seg_rows, y1 = min(p.ny, y0 + p.seg_rows);
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

  // Row yy enters the window of output row yy - r and takes the ring slot of row yy - 2r - 1, which
  // leaves it: V += q(yy) - ring[slot]; ring[slot] = q(yy).  Rows above the domain would enter zeros
  // into a ring of zeros: the march starts at the first row inside.  Three stretches without per-row
  // bounds tests: rows that only fill the window, rows that also complete an output row, and the rows
  // below the domain (zeros enter).  The loads of the next batch are issued before the current batch
  // is consumed (a warp alone on its scheduler would otherwise alternate between waiting for memory
  // and computing).
  constexpr int NR = kVhRows;
  const int ya = max(y0 - r, 0), yb = min(y1 + r, p.ny);          // input rows of this segment
  const int y_out = y0 + r;                                       // first input row that completes an output row
  uint2* cell = ring + threadIdx.x;
  uint2* const ring_end = ring + (size_t)nb * blockDim.x + threadIdx.x;   // this thread's column, one past the last slot
  const bool cbany = (cb0 | cb1) != 0u;
  if (active) {
    const float* lp = sl + (long long)ya * p.nx;
    const uint8_t* lm = M2D ? mp + (long long)ya * p.nx : nullptr;
    unsigned* sp = vp + (long long)y0 * p.nx;
    float2 d[NR], dn[NR];
    uchar2 m[NR], mn[NR];
    auto load_batch = [&](int n, float2* dd, uchar2* mm) {      // rows lp .. lp + n - 1 (n <= NR), then advance
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        if (j < n) {
          dd[j] = __ldg(reinterpret_cast<const float2*>(lp));
          mm[j] = M2D ? __ldg(reinterpret_cast<const uchar2*>(lm)) : make_uchar2(1, 1);
          lp += p.nx;
          if (M2D) lm += p.nx;
        }
      }
    };
    auto enter_row = [&](int y, const float2& dv, const uchar2& mv, bool flag_rows, auto out_tag) {
      constexpr bool OUT = decltype(out_tag)::value;
      dmax = max3_nan(dmax, dv.x, dv.y);                          // NaN propagates (nbhood.py:167-168)
      if (want_stats) dmin = min3f(dmin, dv.x, dv.y);
      float a = truth_value_sum<TM>(dv.x, thr), b = truth_value_sum<TM>(dv.y, thr);
      if (M2D) { a = mv.x ? a : 0.f; b = mv.y ? b : 0.f; }
      const unsigned q0 = __float2uint_rn(a * scale), q1 = __float2uint_rn(b * scale);
      if (want_flags && (flag_rows || cbany)) {
        const unsigned rb = (y <= r ? 1u : 0u) | (y >= p.ny - 1 - r ? 2u : 0u);
        if (q0 != ONE) ne_bits |= rb | cb0;
        if (q1 != ONE) ne_bits |= rb | cb1;
      }
      if (OUT) {
        const uint2 old = *cell;
        V0 += q0 - old.x; V1 += q1 - old.y;
        *reinterpret_cast<uint2*>(sp) = make_uint2(V0, V1);
        sp += p.nx;
      } else {
        V0 += q0; V1 += q1;                                       // (fresh slot: still zero)
      }
      *cell = make_uint2(q0, q1);
      cell += blockDim.x;
      if (cell == ring_end) cell = ring + threadIdx.x;
    };
    auto march = [&](int y_from, int y_to, auto out_tag) {
      // full batches, then the remainder; d holds the rows y_from .. on entry
      int y = y_from;
      while (y < y_to) {
        const int n = min(NR, y_to - y);
        const int n_next = min(NR, yb - (y + n));                 // (the next batch may belong to the next stretch)
        if (n_next > 0) load_batch(n_next, dn, mn);
        const bool flag_rows = y <= r || y + n - 1 >= p.ny - 1 - r;
        if (n == NR) {
#pragma unroll
          for (int jj = 0; jj < NR; ++jj) enter_row(y + jj, d[jj], m[jj], flag_rows, out_tag);
        } else {
#pragma unroll
          for (int jj = 0; jj < NR; ++jj) if (jj < n) enter_row(y + jj, d[jj], m[jj], flag_rows, out_tag);
        }
#pragma unroll
        for (int jj = 0; jj < NR; ++jj) { d[jj] = dn[jj]; m[jj] = mn[jj]; }
        y += n;
      }
    };
    const int y_mid = min(yb, max(y_out, ya));
    // the batches follow the stretches: the first batch of a stretch is loaded by the previous one
    load_batch(min(NR, (ya < y_mid ? y_mid : yb) - ya), d, m);
    if (ya < y_mid) {
      // re-plan: batches never straddle y_mid (the warm-up stretch is loaded and consumed on its own)
      int y = ya;
      while (y < y_mid) {
        const int n = min(NR, y_mid - y);
        const int after = y + n;
        const int n_next = after < y_mid ? min(NR, y_mid - after) : min(NR, yb - after);
        if (n_next > 0) load_batch(n_next, dn, mn);
        const bool flag_rows = y <= r || y + n - 1 >= p.ny - 1 - r;
#pragma unroll
        for (int jj = 0; jj < NR; ++jj) if (jj < n) enter_row(y + jj, d[jj], m[jj], flag_rows, std::false_type{});
#pragma unroll
        for (int jj = 0; jj < NR; ++jj) { d[jj] = dn[jj]; m[jj] = mn[jj]; }
        y = after;
      }
    }
    march(y_mid, yb, std::true_type{});
    // rows below the domain: zeros enter, the window shrinks
    for (int y = yb; y < y1 + r; ++y) {
      const uint2 old = *cell;
      V0 -= old.x; V1 -= old.y;
      if (y >= y_out) {
        *reinterpret_cast<uint2*>(sp) = make_uint2(V0, V1);
        sp += p.nx;
      }
      cell += blockDim.x;
      if (cell == ring_end) cell = ring + threadIdx.x;
    }
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
__global__ void __launch_bounds__(kVhHWarps * 32, 3)
square_h_kernel(const VhParams p, int row_pitch) {
  extern __shared__ unsigned smem_h[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* const Pbuf = smem_h + (size_t)warp * 2 * row_pitch;   // two row buffers: [nx] values, then inclusive prefix in place
  const int r = p.r, nx = p.nx, nb = 2 * r + 1;
  const int n_cols_padded = (nx + 127) / 128 * 128;          // the prefix step works on whole 128-column groups
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
  const bool wrap_x = (p.flags & NBH_WRAP_X) != 0;
  const float scale = (float)(1u << p.fx_shift), inv_scale = 1.0f / scale;
  const long long n_rows = (long long)p.n_units_chunk * p.ny;
  float omin = __int_as_float(0x7f800000), omax = __int_as_float(0xff800000);
  long long cur_unit = -1;
  // the row after the current one is copied into the other buffer while the current one is worked on
  // (cp.async: no registers; a warp would otherwise wait out the load latency of every row)
  const long long row_step = (long long)gridDim.x * kVhHWarps;
  auto stage_row = [&](long long row, unsigned* dst) {
    const unsigned* src = p.vplane + row * nx;                  // (chunk-local row index: the plane rows are contiguous)
    for (int i = 2 * lane; i < nx; i += 64) {                   // rows are 8-byte aligned (nx is even)
      const unsigned d32 = smem_u32(dst + i);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d32), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const long long row_first = (long long)blockIdx.x * kVhHWarps + warp;
  if (row_first < n_rows) stage_row(row_first, Pbuf);
  int which = 0;
  for (long long row = row_first; row < n_rows; row += row_step, which ^= 1) {
    unsigned* P = Pbuf + (size_t)which * row_pitch;
    if (row + row_step < n_rows) {
      stage_row(row + row_step, Pbuf + (size_t)(which ^ 1) * row_pitch);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    for (int i = nx + lane; i < n_cols_padded; i += 32) P[i] = 0u;   // the tail of the last 128-column group
    __syncwarp();
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
    // inclusive prefix in place, 128 columns per step: a lane scans four consecutive values (one
    // 16-byte access each way), the warp scans the lane totals, the running carry is added
    {
      unsigned carry = 0u;
#pragma unroll 3
      for (int c = 4 * lane; c < n_cols_padded; c += 128) {
        const uint4 q = *reinterpret_cast<const uint4*>(P + c);
        const unsigned v0 = q.x, v1 = v0 + q.y, v2 = v1 + q.z, v3 = v2 + q.w;
        unsigned incl = v3;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        const unsigned base = carry + incl - v3;
        *reinterpret_cast<uint4*>(P + c) = make_uint4(v0 + base, v1 + base, v2 + base, v3 + base);
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncwarp();
    const int rows_in = min(y + r, p.ny - 1) - max(y - r, 0) + 1;
    const float nrow = (float)rows_in * scale;
    const long long obase = unit * (long long)p.ny * nx + (long long)y * nx;
    float* orow = p.out + obase;
    // results, specialised on sum_only and on the re-mask plane (no per-point tests of launch constants)
    auto emit = [&](auto sum_tag, auto om_tag) {
      constexpr bool SUM = decltype(sum_tag)::value, OM = decltype(om_tag)::value;
      // one result: S / nf as a correctly rounded quotient (reciprocal + two FMAs), NaN where nothing is valid
      auto finish = [&](int xo, unsigned S, float nf, float rc) {
        const float Sf = (float)S;
        float o;
        if (SUM) {
          o = Sf * inv_scale;
        } else {
          const float q0 = Sf * rc;
          const float rem = __fmaf_rn(-q0, nf, Sf);
          o = (M2D && nf == 0.f) ? __int_as_float(0x7fc00000) : __fmaf_rn(rem, rc, q0);
          omax = fmaxf(omax, o); omin = fminf(omin, o);
        }
        orow[xo] = o;
        if (OM) p.out_mask[obase + xo] = (M2D && p.mask2d[(long long)y * nx + xo] == 0) ? 1 : 0;
      };
      // interior columns: the whole window inside the row, one divisor for all of them
      if (nx >= 2 * r + 2) {
        const float nf_in = nrow * (float)nb;
        float rc_in;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc_in) : "f"(nf_in));
        const float* crow = M2D ? p.cnt_plane + (long long)y * nx : nullptr;
        const unsigned* Phi = P + r;
        const unsigned* Plo = P - r - 1;
#pragma unroll 4
        for (int xo = r + 1 + lane; xo < nx - r; xo += 32) {
          const unsigned S = Phi[xo] - Plo[xo];
          float nf = nf_in, rc = rc_in;
          if (M2D && !SUM) {
            nf = __ldg(crow + xo) * scale;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
          }
          finish(xo, S, nf, rc);
        }
      }
      // the columns whose window is cut by the row ends (all of them in a row narrower than two windows)
      auto cut_column = [&](int xo) {
        const int hi = min(xo + r, nx - 1), lo = xo - r - 1;
        unsigned S;
        int cols_in;
        if (wrap_x) {
          // periodic x (global grids): the part of the window beyond a row end comes from the other end
          const unsigned total = P[nx - 1];
          if (xo + r >= nx) S = total + P[xo + r - nx] - P[lo];
          else S = P[xo + r] + total - P[nx + lo];          // (lo == -1: nothing wraps, total - P[nx - 1] = 0)
          cols_in = nb;
        } else {
          S = P[hi] - (lo >= 0 ? P[lo] : 0u);
          cols_in = hi - max(xo - r, 0) + 1;
        }
        float nf = 1.f, rc = 1.f;
        if (!SUM) {
          if (M2D) nf = __ldg(p.cnt_plane + (long long)y * nx + xo) * scale;
          else nf = nrow * (float)cols_in;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
        }
        finish(xo, S, nf, rc);
      };
      if (nx >= 2 * r + 2) {
        for (int e = lane; e < 2 * r + 1; e += 32) cut_column(e <= r ? e : nx - 2 * r - 1 + e);
      } else {
        for (int xo = lane; xo < nx; xo += 32) cut_column(xo);
      }
    };
    if (sum_only) {
      if (p.out_mask) emit(std::true_type{}, std::true_type{}); else emit(std::true_type{}, std::false_type{});
    } else {
      if (p.out_mask) emit(std::false_type{}, std::true_type{}); else emit(std::false_type{}, std::false_type{});
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
  if (d.mask_nd) return false;
  if ((d.flags & NBH_WRAP_X) && d.nx < 2 * r + 2) return false;     // a window wraps at one row end at most
  if ((d.nx & 1) || d.nx < 2 || (reinterpret_cast<uintptr_t>(in) & 7) || (d.plane_stride & 1)) return false;
  if ((size_t)(d.nx + 132) * 4 * 2 * kVhHWarps > 200u * 1024) return false;
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
  const int row_pitch = (p.nx + 127) / 128 * 128 + 4;        // whole 128-column groups, 16-byte aligned rows
  const size_t smem = (size_t)kVhHWarps * 2 * row_pitch * 4;      // two row buffers per warp
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
  // y segments of the V pass: long segments re-read fewer halo rows, short ones fill the machine
  {
    const long long n_units_all = sp.n_planes * sp.n_thr;
    const long long xchunks_est = (sp.nx / 2 + 255) / 256;
    int seg = kVhSegRowsMax;
    while (seg > 64 && seg > 4 * sp.r &&
           n_units_all * ((sp.ny + seg - 1) / seg) * xchunks_est < 4LL * sm_count()) seg /= 2;
    p.seg_rows = seg;
    p.n_ysegs = (sp.ny + seg - 1) / seg;
  }
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
