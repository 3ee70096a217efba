// Square neighbourhood, row-streaming kernel: one persistent block per SM,
// whole grid rows staged in shared memory by bulk asynchronous copies.
//
// Why: the warp-autonomous kernel re-reads the x halo of every 64-column strip
// and the y halo of every y segment (1.23 x 1.17 = 1.44 x the algorithmic bytes
// at the L1 level for the UK grid, served by L2) and holds its bytes in flight
// in registers.  Here a block owns FULL rows (or, for rows wider than 20
// windows, the column segment of one x panel): four producer threads stream the
// rows of every member of a contiguous share of the (plane, y) rows into a ring
// of shared-memory stages with `cp.async.bulk` (UBLKCP) completing on mbarriers
// (one thread issues only one bulk copy per ~400-700 cycles, hence four), and
// one consumer warp per 64-column window reads its (overlapping) window of
// every staged row from shared memory.  The x halo therefore costs shared-
// memory reads only, the y halo is 2r rows per ~236-row share (1.05 x), the
// bytes in flight are bounded by the ~150 KB stage ring instead of registers,
// and no load instruction or address arithmetic is issued by the math warps.
//
// Rows of the UK grid are 4168 B, i.e. only 8-byte aligned, while a bulk copy
// needs 16-byte aligned source, destination and size: every copy starts at the
// 16-byte boundary at or below the row start and ends at the boundary at or
// above the row end; consumers add the (warp-uniform) 0..3-float shift when
// reading the stage.  The widened copy never leaves the 16-byte blocks that
// hold the row, so it cannot touch an unmapped page.
//
// Work split: the n_units * ny output rows (unit = plane x threshold x panel)
// are cut into gridDim.x equal contiguous shares; a share may straddle units
// and is processed piece by piece.  Odd blocks march upwards so the 2r halo
// rows two neighbouring shares have in common are read at about the same time
// (L2 hit).
//
// A stage holds `ms` member rows of one grid row (n_members >= 2: the fused
// threshold -> box sum -> realization mean).  Single-member slices are the
// business of square_box_kernel (nbh_square_box.cuh).
//
// Arithmetic: float32 member sum; thresholded launches then use exact unsigned
// fixed-point vertical / horizontal window sums (see the consumer section),
// raw-data launches the float64 sums of square_warp_kernel.
#pragma once
#include "nbh_async.cuh"
#include "nbh_square_warp.cuh"

namespace nbh {

constexpr int kRsMaxWindows = 20;                       // consumer warps per block
constexpr int kRsProducers = 4;                         // producer warps: one thread issues a bulk copy every
                                                        // ~400 cycles whatever its size (profiles/probe_bulk.cu),
                                                        // so the member rows of a stage are issued by four threads
constexpr int kRsThreads = (kRsMaxWindows + kRsProducers) * 32;
constexpr int kRsMaxStages = 16;

struct SqRsGeom {
  int n_windows;        // consumer warps = 64-column windows per x panel
  int n_panels;         // x panels per row (rows wider than kRsMaxWindows windows are split)
  int n_windows_total;  // windows per full row
  int W0, W, hl;        // output columns of window 0 / the others, left halo (even, >= r)
  int ms;               // items (member rows) per stage; divides n_members
  int n_stages;
  int pitch;            // floats per staged item (multiple of 4)
  long long rows_total; // n_units * ny
  long long share;      // rows per block
  int fx_shift;         // fixed-point scale of the thresholded member sums (2^fx_shift)
  unsigned wait_ns;     // suspend-time hint of the barrier waits
};

__device__ __forceinline__ float2 lds_f2(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// the share of block b as a sequence of pieces (unit, ya, yb), in marching order
struct RsPieces {
  long long g0, g1, g;
  int ny, dir;
  __device__ RsPieces(long long share, long long total, int ny_, int b) : ny(ny_) {
    g0 = (long long)b * share;
    g1 = min(total, g0 + share);
    dir = (b & 1) ? -1 : 1;
    g = dir > 0 ? g0 : g1;
  }
  __device__ bool next(long long* unit, int* ya, int* yb) {
    if (dir > 0) {
      if (g >= g1) return false;
      *unit = g / ny;
      *ya = (int)(g - *unit * ny);
      *yb = (int)min((long long)ny, *ya + (g1 - g));
      g += *yb - *ya;
    } else {
      if (g <= g0) return false;
      *unit = (g - 1) / ny;
      *yb = (int)(g - *unit * ny);
      *ya = (int)max(0LL, *yb - (g - g0));
      g -= *yb - *ya;
    }
    return true;
  }
};

template <int TM, bool M2D>
__global__ void __launch_bounds__(kRsThreads, 1)
square_rs_kernel(const __grid_constant__ SqParams p, const SqRsGeom g) {
  constexpr int VEC = 2, WS = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = p.nb, r = p.r, R = p.n_members, NS = g.n_stages, MS = g.ms;
  const int NW = g.n_windows;

  // layout: stages | full[16] empty[16] | rtab | per consumer warp: hrow, ring
  float* stages = reinterpret_cast<float*>(smem_raw);
  const unsigned stage_bytes = (unsigned)(MS * g.pitch * 4);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)NS * stage_bytes);
  double* rtab = reinterpret_cast<double*>(bars + 2 * kRsMaxStages);
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + kRsMaxStages);

  for (int k = threadIdx.x + 1; k <= nb; k += blockDim.x) rtab[k] = 1.0 / (double)k;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full0 + 8 * s, kRsProducers); mbar_init(empty0 + 8 * s, NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();      // the only block-level barrier

  const unsigned long long in_elem0 = reinterpret_cast<unsigned long long>(p.in) >> 2;

  // ======================= producer warps =======================
  // producer pw issues the member rows j = pw, pw + kRsProducers, ... of every stage and
  // announces their bytes with its own arrive.expect_tx (full[s] counts kRsProducers arrivals)
  if (warp >= NW) {
    if (lane != 0) return;
    const int pw = warp - NW;
    const int n_mine = pw < MS ? (MS - pw + kRsProducers - 1) / kRsProducers : 0;
    const unsigned stage0 = smem_u32(stages);
    int s = 0;
    unsigned round = 0;
    RsPieces pieces(g.share, g.rows_total, p.ny, blockIdx.x);
    long long unit; int ya, yb;
    long long task;
    while (pieces.next(&task, &ya, &yb)) {
      unit = task / g.n_panels;
      const int panel = (int)(task - unit * g.n_panels);
      const long long plane = unit / p.n_thr;
      const int dir = pieces.dir;
      const int y_lo = max(ya - r, 0), y_hi = min(yb - 1 + r, p.ny - 1);
      const int n_in = y_hi - y_lo + 1;
      // columns of this panel: from the first window's load start to the last window's right halo
      const int gw0 = panel * NW, gw1 = min(gw0 + NW, g.n_windows_total) - 1;
      const int seg0 = gw0 == 0 ? 0 : g.W0 + (gw0 - 1) * g.W - g.hl;
      const int seg1 = min(p.nx, (gw1 == 0 ? 0 : g.W0 + (gw1 - 1) * g.W) + (gw1 == 0 ? g.W0 : g.W) + r);
      const float* plane_ptr = p.in + plane * p.plane_stride + seg0;
      for (int k = 0; k < n_in; ++k) {
        const int y = dir > 0 ? y_lo + k : y_hi - k;
        const float* row = plane_ptr + (long long)y * p.nx;
        const unsigned shift_b = (unsigned)(reinterpret_cast<unsigned long long>(row) & 15ull);
        const unsigned bytes = (shift_b + (unsigned)(seg1 - seg0) * 4u + 15u) & ~15u;
        const char* src = reinterpret_cast<const char*>(row) - shift_b;
        for (int m0 = 0; m0 < R; m0 += MS) {
          mbar_wait_hint(empty0 + 8 * s, (round & 1u) ^ 1u, g.wait_ns);
          mbar_expect_tx(full0 + 8 * s, bytes * (unsigned)n_mine);
          unsigned dst = stage0 + (unsigned)s * stage_bytes + (unsigned)(pw * g.pitch) * 4u;
          const char* q = src + (long long)(m0 + pw) * p.member_stride * 4;
          for (int j = pw; j < MS; j += kRsProducers) {
            bulk_g2s(dst, q, bytes, full0 + 8 * s);
            dst += (unsigned)(kRsProducers * g.pitch) * 4u;
            q += (long long)kRsProducers * p.member_stride * 4;
          }
          if (++s == NS) { s = 0; ++round; }
        }
      }
    }
    return;
  }

  // ======================= consumer warps =======================
  // Thresholded launches (FX): truth values lie in [0, 1], so the member-summed row P is turned
  // into fixed point (P * 2^fx_shift, rounded to nearest) and the vertical running sums, the row
  // prefix and the window differences are exact unsigned 32-bit arithmetic (wrap-around cancels in
  // the differences: the host picks fx_shift with nb^2 * R * 2^fx_shift < 2^31).  The rounding of
  // P contributes at most 2^-(fx_shift+1) / R to a mean (< 1e-7 for the UK configuration).
  // Raw-data launches keep the float64 sums of square_warp_kernel.
  constexpr bool FX = TM != 0;
  const int n_hrows = 1;                                                 // prefix rows per warp
  const size_t per_warp = (size_t)(WS + 2) * 8 * n_hrows + (size_t)nb * WS * 4;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(rtab + ((nb + 2) & ~1)) + (size_t)warp * per_warp;
  double* hrow = reinterpret_cast<double*>(wbase);                       // prefix rows: double or unsigned [WS + 2]
  unsigned* hrow_u = reinterpret_cast<unsigned*>(wbase);
  float* ring = reinterpret_cast<float*>(wbase + (size_t)(WS + 2) * 8 * n_hrows);  // [nb][WS] float or unsigned
  unsigned* ring_u = reinterpret_cast<unsigned*>(ring);

  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
  const int c0 = lane * VEC;
  const unsigned long long all_members = R >= 64 ? ~0ull : ((1ull << R) - 1ull);
  const float fx_scale = FX ? (float)(1u << g.fx_shift) : 1.f;
  const double fx_inv = FX ? 1.0 / (double)(1u << g.fx_shift) : 1.0;
  const double inv_members_fx = p.inv_members * fx_inv;
  if (lane < n_hrows) { hrow[lane * (WS + 2)] = 0.0; }    // prefix[0] of every prefix row (also zeroes hrow_u[0..1])
  int idx_hi[VEC], idx_lo[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    idx_hi[v] = c0 + v + r + 1;
    idx_lo[v] = max(c0 + v - r, 0);
  }

  const unsigned pitch_b = (unsigned)g.pitch * 4u;
  const unsigned stages_u32 = smem_u32(stages);
  float nanacc = 0.f;
  // stage cursor: barrier address, stage address, phase parity
  int s = 0;
  unsigned parity = 0, bar = full0, st_addr = stages_u32;
  auto next_stage = [&]() {
    bar += 8; st_addr += stage_bytes;
    if (++s == NS) { s = 0; parity ^= 1u; bar = full0; st_addr = stages_u32; }
  };
  RsPieces pieces(g.share, g.rows_total, p.ny, blockIdx.x);
  long long task, unit; int ya, yb;
  while (pieces.next(&task, &ya, &yb)) {
    unit = task / g.n_panels;
    const int panel = (int)(task - unit * g.n_panels);
    // ---- window geometry of this piece (lane constants): window 0 of the row starts at the
    // domain edge and needs no left halo; windows past the end of the row are idle ----
    const int gw = panel * NW + warp;                  // window index within the full row
    const int gw0 = panel * NW;
    const int seg0 = gw0 == 0 ? 0 : g.W0 + (gw0 - 1) * g.W - g.hl;     // first staged column
    const int x0 = gw == 0 ? 0 : g.W0 + (gw - 1) * g.W;
    const int load0 = gw == 0 ? 0 : x0 - g.hl;
    const int wk = gw == 0 ? g.W0 : g.W;
    const int c_out0 = x0 - load0;
    const int gxr = load0 + c0;
    const bool col_active = gw < g.n_windows_total && gxr < p.nx;
    const int gx = col_active ? gxr : seg0;            // inactive lanes read the first staged column (discarded)
    const bool band_l = !sum_only && col_active && (gxr + VEC - 1 <= r);
    const bool band_r = !sum_only && col_active && (gxr >= p.nx - 1 - r);
    const bool warp_band_l = __any_sync(0xffffffffu, band_l), warp_band_r = __any_sync(0xffffffffu, band_r);
    const bool is_out = (c0 >= c_out0) && (c0 < c_out0 + wk) && col_active;
    double rcol[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int xo = gx + v;
      const int cols_in = min(xo + r, p.nx - 1) - max(xo - r, 0) + 1;
      rcol[v] = (sum_only ? 1.0 : rtab[min(max(cols_in, 1), nb)] * p.inv_members) * fx_inv;
    }
    const unsigned lane_col_b = (unsigned)(gx - seg0) * 4u;            // this lane's byte offset in a staged row
    const long long plane = unit / p.n_thr;
    const int t = (int)(unit - plane * p.n_thr);
    const ThrDev& thr = p.thr.t[TM ? t : 0];
    const int dir = pieces.dir;
    const int n_rows = (yb - ya) + 2 * r;
    const int y_first = dir > 0 ? ya - r : yb - 1 + r;
    const long long obase0 = unit * (long long)p.ny * p.nx;
    const unsigned long long plane_elem0 = in_elem0 + (unsigned long long)(plane * p.plane_stride) + (unsigned long long)seg0;

    for (int q = 0; q < nb; ++q) *reinterpret_cast<float2*>(ring + q * WS + c0) = make_float2(0.f, 0.f);
    double V0 = 0.0, V1 = 0.0;
    unsigned U0 = 0, U1 = 0;
    int slot = 0;
    unsigned long long ne_left = 0, ne_right = 0, ne_top = 0, ne_bot = 0;
    // column bands: the flagged (slower) member loop is only needed until every member has shown
    // a value != 1 in the band(s) this warp covers; the bits are sticky, so for ordinary fields
    // that is the first row or two of the piece
    bool need_cols = warp_band_l || warp_band_r;
    unsigned m2_next = 0x0101u;
    if (M2D) m2_next = ldm_packed<VEC>(p.mask2d + (long long)min(max(y_first, 0), p.ny - 1) * p.nx + gx);
    __syncwarp();

    for (int k = 0; k < n_rows; ++k) {
      const int yy = y_first + dir * k;
      const bool in_dom = (yy >= 0) && (yy < p.ny);
      const unsigned m2_cur = m2_next;
      if (M2D && k + 1 < n_rows)
        m2_next = ldm_packed<VEC>(p.mask2d + (long long)min(max(yy + dir, 0), p.ny - 1) * p.nx + gx);
      float acc0 = 0.f, acc1 = 0.f;
      if (in_dom) {
        const int shift = (int)((plane_elem0 + (unsigned long long)((long long)yy * p.nx)) & 3ull);
        const bool row_band = yy <= r || yy >= p.ny - 1 - r;
        const bool flagged = !sum_only && (need_cols || row_band);
        const bool v0 = !M2D || (m2_cur & 0xffu), v1 = !M2D || (m2_cur & 0xff00u);
        const unsigned shift_b = (unsigned)shift * 4u;
        if (!flagged) {
          for (int m0 = 0; m0 < R; m0 += MS) {
            mbar_wait_hint(bar, parity, g.wait_ns);
            unsigned q = st_addr + lane_col_b + shift_b;
#pragma unroll 3
            for (int j = 0; j < MS; ++j) {
              const float2 d = lds_f2(q);
              q += pitch_b;
              nanacc = max_nan(nanacc, fabsf(d.x));
              nanacc = max_nan(nanacc, fabsf(d.y));
              acc0 += truth_value_sum<TM>(d.x, thr);
              acc1 += truth_value_sum<TM>(d.y, thr);
            }
            __syncwarp();                                  // every lane has read stage s
            if (lane == 0) mbar_arrive(bar + 8 * kRsMaxStages);
            next_stage();
          }
          if (M2D) { acc0 = v0 ? acc0 : 0.f; acc1 = v1 ? acc1 : 0.f; }   // the mask is shared by all members
        } else {
          unsigned long long ne = 0;
          for (int m0 = 0; m0 < R; m0 += MS) {
            mbar_wait_hint(bar, parity, g.wait_ns);
            unsigned q = st_addr + lane_col_b + shift_b;
            unsigned long long bit = 1ull << m0;
#pragma unroll 2
            for (int j = 0; j < MS; ++j) {
              const float2 d = lds_f2(q);
              q += pitch_b;
              nanacc = max_nan(nanacc, fabsf(d.x));
              nanacc = max_nan(nanacc, fabsf(d.y));
              acc0 += truth_value_sum<TM>(d.x, thr);
              acc1 += truth_value_sum<TM>(d.y, thr);
              if (truth_not_one(d.x, thr) || truth_not_one(d.y, thr)) ne |= bit;
              bit <<= 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar + 8 * kRsMaxStages);
            next_stage();
          }
          if (M2D) { acc0 = v0 ? acc0 : 0.f; acc1 = v1 ? acc1 : 0.f; }
          if (!v0 || !v1) ne = all_members;                // a masked cell counts as "not one" for every member
          if (col_active) {
            if (band_l) ne_left |= ne;
            if (band_r) ne_right |= ne;
            if (yy <= r) ne_top |= ne;
            if (yy >= p.ny - 1 - r) ne_bot |= ne;
          }
          if (need_cols) {
            // warp-uniform: have all members shown a non-one in every column band of this warp?
            bool done = true;
            if (warp_band_l) {
              const unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)ne_left);
              const unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(ne_left >> 32));
              done = done && ((((unsigned long long)hi32 << 32) | lo32) == all_members);
            }
            if (warp_band_r) {
              const unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)ne_right);
              const unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(ne_right >> 32));
              done = done && ((((unsigned long long)hi32 << 32) | lo32) == all_members);
            }
            need_cols = !done;
          }
        }
      }
      if (!in_dom || !col_active) { acc0 = 0.f; acc1 = 0.f; }

      const int yo = y_first + dir * (k - r);                // centre row of the completed window
      const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
      const long long pt0 = (long long)yo * p.nx + gxr;
      float o[VEC];
      if (FX) {
        // ---- exact fixed-point vertical sums, warp scan, prefix difference ----
        const unsigned q0 = __float2uint_rn(acc0 * fx_scale), q1 = __float2uint_rn(acc1 * fx_scale);
        uint2* rp = reinterpret_cast<uint2*>(ring_u + slot * WS + c0);
        const uint2 old = *rp;
        *rp = make_uint2(q0, q1);
        U0 += q0 - old.x;
        U1 += q1 - old.y;
        slot = (slot + 1 == nb) ? 0 : slot + 1;
        if (k < 2 * r) continue;
        const unsigned s1 = U0 + U1;
        unsigned incl = s1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          asm volatile("{\n.reg .pred p;\n.reg .u32 t;\nshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n@p add.u32 %0, %0, t;\n}\n"
                       : "+r"(incl) : "r"(d));
        }
        const unsigned excl = incl - s1;
        *reinterpret_cast<uint2*>(hrow_u + c0 + 2) = make_uint2(excl + U0, excl + s1);   // prefix[c0 + 1], [c0 + 2] at +1
        __syncwarp();
        if (is_out) {
          const double rrow = sum_only ? 1.0 : rtab[rows_in];
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const double Sw = (double)(hrow_u[idx_hi[v] + 1] - hrow_u[idx_lo[v] + 1]);
            if (M2D && !sum_only) o[v] = (float)(Sw * p.rcp_count[pt0 + v] * inv_members_fx);
            else o[v] = (float)(Sw * rrow * rcol[v]);
          }
        }
      } else {
        // ---- float64 vertical sums through the warp-private ring ----
        float* rp = ring + slot * WS + c0;
        const float2 old = *reinterpret_cast<const float2*>(rp);
        *reinterpret_cast<float2*>(rp) = make_float2(acc0, acc1);
        V0 += (double)acc0 - (double)old.x;
        V1 += (double)acc1 - (double)old.y;
        slot = (slot + 1 == nb) ? 0 : slot + 1;
        if (k < 2 * r) continue;
        const double s1 = V0 + V1;
        double incl = s1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double o2 = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o2;
        }
        const double excl = incl - s1;
        hrow[c0 + 1] = excl + V0;
        hrow[c0 + 2] = excl + s1;
        __syncwarp();
        if (is_out) {
          const double rrow = sum_only ? 1.0 : rtab[rows_in];
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const double Sw = hrow[idx_hi[v]] - hrow[idx_lo[v]];
            if (M2D && !sum_only) o[v] = (float)(Sw * p.rcp_count[pt0 + v] * p.inv_members);
            else o[v] = (float)(Sw * rrow * rcol[v]);
          }
        }
      }
      if (is_out) {
        VecLoad<VEC>::st(p.out + obase0 + pt0, o);
        if (p.out_mask) {
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            p.out_mask[obase0 + pt0 + v] = (M2D && p.mask2d[pt0 + v] == 0) ? 1 : 0;
        }
      }
      __syncwarp();                                        // the prefix row is rewritten by the next row
    }

    // ---- edge-band flags of this unit (trimming fix-up) ----
    if (!sum_only) {
      unsigned long long masks[4] = {ne_top, ne_bot, ne_left, ne_right};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)masks[k]);
        const unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(masks[k] >> 32));
        masks[k] = ((unsigned long long)hi32 << 32) | lo32;
      }
      if (lane == 0 && (masks[0] | masks[1] | masks[2] | masks[3])) {
        for (int m = 0; m < R; ++m) {
          unsigned bits = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) bits |= (unsigned)((masks[k] >> m) & 1ull) << k;
          if (bits) atomicOr(p.edge_flags + ((plane * R + m) * p.n_thr + t), bits);
        }
      }
    }
  }

  if (__any_sync(0xffffffffu, __isnanf(nanacc))) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
}

inline size_t sq_rs_fixed_smem(int nb, int n_windows, int n_hrows = 1) {
  return 2 * kRsMaxStages * 8 + (size_t)((nb + 2) & ~1) * 8 +
         (size_t)n_windows * ((size_t)(64 + 2) * 8 * n_hrows + (size_t)nb * 64 * 4);
}

int launch_square_rs(const SqParams& p, const SqRsGeom& g, int tm, bool m2d, int grid, cudaStream_t st);

}  // namespace nbh
