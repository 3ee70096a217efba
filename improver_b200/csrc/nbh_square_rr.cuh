// Square neighbourhood, row-streaming kernel for single-member slices
// (n_members == 1: NeighbourhoodProcessing.process over the slices of a cube,
// improver/nbhood/nbhood.py:457-463 with _calculate_neighbourhood :244-317).
//
// Same idea as square_rs_kernel (nbh_square_rs.cuh): one persistent block per
// SM owns a contiguous share of the (slice, y) rows, producer threads stream
// whole rows (or the row segment of one x panel) into shared memory with
// cp.async.bulk, consumer warps work on 128-column windows of the staged rows.
// Differences that matter when every input row is used by one slice only:
//   * the staged rows themselves are the vertical ring: a row stays in shared
//     memory until it leaves the 2r+1 window, and the value that leaves the
//     running sum is re-read (and re-thresholded) from there;
//   * rows are processed in batches of kRrRows: the vertical updates of a
//     batch are sequential but its horizontal scans are independent, which
//     gives the instruction-level parallelism that hides the shuffle latency;
//   * a periodic x axis (NBH_WRAP_X, global grids) only changes which staged
//     column a halo lane reads: the wrapped halo columns are copied into two
//     small side areas of every row slot;
//   * rows wider than one block's windows are split into x panels.
// Raw data (no threshold) is summed in float64 like the reference; thresholded
// data uses the exact fixed-point sums of square_rs_kernel.
#pragma once
#include "nbh_square_rs.cuh"

namespace nbh {

constexpr int kRrRows = 4;              // rows per chunk (one full / empty barrier pair) and per batch
constexpr int kRrMaxChunks = 32;
constexpr int kRrProducers = 2;
constexpr int kRrMaxCWarps = 10;        // consumer warps: 12 warps x 168 registers
constexpr int kRrThreads = (kRrMaxCWarps + kRrProducers) * 32;
constexpr int kRrHalo = 16;             // floats per wrapped-halo side area

struct SqRrGeom {
  int n_windows;        // 128-column windows per panel
  int n_cwarps;         // consumer warps (warp w works on windows w, w + n_cwarps, ...)
  int n_panels;         // x panels per row
  int panel_w;          // output columns per panel (multiple of W)
  int W0, W, hl;        // output columns of window 0 (no left halo) / the others, left halo
  int n_chunks;         // chunk slots in the row ring
  int row_pitch;        // floats per row slot (main segment + 2 * kRrHalo)
  int seg_pitch;        // floats reserved for the main segment
  int pw;               // entries per prefix row
  long long rows_total; // n_units * n_panels * ny
  long long share;      // rows per block
  int fx_shift;
  unsigned wait_ns;
};

// producer-side wait: poll, sleeping ~100 ns between polls (a spinning producer would share the
// issue port of its scheduler with the math warps)
__device__ __forceinline__ void mbar_wait_backoff(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (;;) {
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(100);
  }
}

__device__ __forceinline__ void lds4(unsigned addr, float (&v)[4]) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr));
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}

template <int TM, bool M2D>
__global__ void __launch_bounds__(kRrThreads, 1)
square_rr_kernel(const __grid_constant__ SqParams p, const SqRrGeom g) {
  constexpr int MS = kRrRows;
  constexpr bool FX = TM != 0;
  constexpr int ES = FX ? 4 : 8;                  // bytes per running-sum / prefix entry
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = p.nb, r = p.r, NCH = g.n_chunks;
  const int NCW = g.n_cwarps;
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;

  // layout: row slots | full[32] empty[32] | rtab | window state | per consumer warp: MS prefix rows
  const unsigned row_bytes = (unsigned)g.row_pitch * 4u;
  const unsigned ring_bytes = (unsigned)(NCH * MS) * row_bytes;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + ring_bytes);
  double* rtab = reinterpret_cast<double*>(bars + 2 * kRrMaxChunks);
  unsigned char* state0 = reinterpret_cast<unsigned char*>(rtab + ((nb + 2) & ~1));
  unsigned char* prefix0 = state0 + (size_t)g.n_windows * 128 * ES;
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + kRrMaxChunks);
  const unsigned ring0 = smem_u32(smem_raw);

  for (int k = threadIdx.x + 1; k <= nb; k += blockDim.x) rtab[k] = 1.0 / (double)k;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NCH; ++s) { mbar_init(full0 + 8 * s, kRrProducers); mbar_init(empty0 + 8 * s, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();      // the only block-level barrier

  // ======================= producer warps =======================
  if (warp >= NCW) {
    if (lane != 0) return;
    const int pw = warp - NCW;
    int slot = 0;
    unsigned round = 0;
    RsPieces pieces(g.share, g.rows_total, p.ny, blockIdx.x);
    pieces.dir = 1; pieces.g = pieces.g0;           // rows mode always marches downwards
    long long task; int ya, yb;
    while (pieces.next(&task, &ya, &yb)) {
      const long long unit = task / g.n_panels;
      const int panel = (int)(task - unit * g.n_panels);
      const long long plane = unit / p.n_thr;
      const int px0 = panel * g.panel_w, px1 = min(p.nx, px0 + g.panel_w);
      const int lx0 = max(px0 - g.hl, 0), lx1 = min(px1 + r, p.nx);
      const bool lh = wrap && panel == 0, rh = wrap && panel == g.n_panels - 1;
      const int y_lo = max(ya - r, 0), y_hi = min(yb - 1 + r, p.ny - 1);
      const float* plane_ptr = p.in + plane * p.plane_stride;
      for (int y0 = y_lo; y0 <= y_hi; y0 += MS) {
        const int n = min(MS, y_hi + 1 - y0);
        mbar_wait_backoff(empty0 + 8 * slot, (round & 1u) ^ 1u);
        // bytes this producer will deliver into the chunk
        unsigned my_bytes = 0;
        for (int i = pw; i < n; i += kRrProducers) {
          const float* seg = plane_ptr + (long long)(y0 + i) * p.nx + lx0;
          const unsigned sh = (unsigned)(reinterpret_cast<unsigned long long>(seg) & 15ull);
          my_bytes += (sh + (unsigned)(lx1 - lx0) * 4u + 15u) & ~15u;
          if (lh) my_bytes += (unsigned)g.hl * 4u;
          if (rh) my_bytes += (unsigned)kRrHalo * 4u;
        }
        mbar_expect_tx(full0 + 8 * slot, my_bytes);
        for (int i = pw; i < n; i += kRrProducers) {
          const float* rowp = plane_ptr + (long long)(y0 + i) * p.nx;
          const float* seg = rowp + lx0;
          const unsigned sh = (unsigned)(reinterpret_cast<unsigned long long>(seg) & 15ull);
          const unsigned bytes = (sh + (unsigned)(lx1 - lx0) * 4u + 15u) & ~15u;
          const unsigned dst = ring0 + (unsigned)(slot * MS + i) * row_bytes;
          bulk_g2s(dst, reinterpret_cast<const char*>(seg) - sh, bytes, full0 + 8 * slot);
          // periodic x: the columns beyond either end of the row (16-byte aligned by construction)
          if (lh) bulk_g2s(dst + (unsigned)g.seg_pitch * 4u, rowp + (p.nx - g.hl), (unsigned)g.hl * 4u, full0 + 8 * slot);
          if (rh) bulk_g2s(dst + (unsigned)(g.seg_pitch + kRrHalo) * 4u, rowp, (unsigned)kRrHalo * 4u, full0 + 8 * slot);
        }
        if (++slot == NCH) { slot = 0; ++round; }
      }
    }
    return;
  }

  // ======================= consumer warps =======================
  unsigned char* prefix_w = prefix0 + (size_t)warp * MS * g.pw * ES;
  const unsigned prefix_u32 = smem_u32(prefix_w);
  const float fx_scale = FX ? (float)(1u << g.fx_shift) : 1.f;
  const double fx_inv = FX ? 1.0 / (double)(1u << g.fx_shift) : 1.0;
  const unsigned prow_bytes = (unsigned)g.pw * ES;
  // prefix row layout: entry i of the inclusive prefix lives at index i + PO, so that the four
  // entries a lane writes (columns c0 + 1 .. c0 + 4) start on a 16-byte boundary and the entries
  // below index PO, which are never written, are the zeros before the window
  const int PO = ((r + 3) & ~3) + 3;
  for (int i = lane; i < MS * g.pw * (ES / 4); i += 32) reinterpret_cast<unsigned*>(prefix_w)[i] = 0u;
  __syncwarp();

  const int c0 = lane * 4;
  float nanacc = 0.f;
  int new_slot = 0;                 // chunk slot of the next real batch
  unsigned new_parity = 0;
  RsPieces pieces(g.share, g.rows_total, p.ny, blockIdx.x);
  pieces.dir = 1; pieces.g = pieces.g0;
  long long task; int ya, yb;
  while (pieces.next(&task, &ya, &yb)) {
    const long long unit = task / g.n_panels;
    const int panel = (int)(task - unit * g.n_panels);
    const long long plane = unit / p.n_thr;
    const int t = (int)(unit - plane * p.n_thr);
    const ThrDev& thr = p.thr.t[TM ? t : 0];
    const int px0 = panel * g.panel_w, px1 = min(p.nx, px0 + g.panel_w);
    const int need0 = wrap ? px0 - g.hl : max(px0 - g.hl, 0);        // columns this panel needs: [need0, need1)
    const int need1 = wrap ? px1 + r : min(px1 + r, p.nx);
    const int lx0 = max(px0 - g.hl, 0);
    const int y_lo = max(ya - r, 0), y_hi = min(yb - 1 + r, p.ny - 1), y_end = yb - 1 + r;
    const int n_real = (y_hi - y_lo) / MS + 1;          // chunks of this piece
    const long long obase0 = unit * (long long)p.ny * p.nx;
    const unsigned long long seg_elem0 =
        (reinterpret_cast<unsigned long long>(p.in) >> 2) + (unsigned long long)(plane * p.plane_stride) + lx0;

    int bslot = new_slot;                               // slot the chunk of batch jb has / would have
    int rel_slot = new_slot, released = 0;              // next chunk of this piece to hand back
    unsigned ne_all = 0;                                // bit k: band k has shown a value != 1 (top, bottom, left, right)

    int jb = 0;
    for (int yb0 = y_lo; yb0 <= y_end; yb0 += MS, ++jb) {
      const int n = min(MS, y_end + 1 - yb0);
      const bool has_real = yb0 <= y_hi;
      if (has_real) mbar_wait(full0 + 8 * new_slot, new_parity);

      for (int w = warp; w < g.n_windows; w += NCW) {
        // ---- window geometry (lane constants) ----
        const int x0 = px0 + w * g.W;
        const int load0 = x0 - g.hl;
        const int wk = min(g.W, px1 - x0);
        const int gxr = load0 + c0;                       // may be < 0 or >= nx with a periodic x axis
        unsigned okm = 0;                                 // bit v: column gxr + v is a column this panel needs
#pragma unroll
        for (int v = 0; v < 4; ++v) okm |= (gxr + v >= need0 && gxr + v < need1) ? (1u << v) : 0u;
        int gxw = gxr;                                    // wrapped global column (outputs / masks)
        if (wrap) { if (gxw < 0) gxw += p.nx; else if (gxw >= p.nx) gxw -= p.nx; }
        int lane_off = 0;                                 // float offset of this lane's columns in a row slot
        if (okm) {
          if (gxr < 0) lane_off = g.seg_pitch + (gxr + g.hl);
          else if (gxr >= p.nx) lane_off = g.seg_pitch + kRrHalo + (gxr - p.nx);
          else lane_off = gxr - lx0;
        }
        const int gx = okm ? gxw : 0;
        unsigned blm = 0, brm = 0;                        // columns inside the left / right edge band
        if (!sum_only && !wrap) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if ((okm >> v) & 1u) {
              if (gxr + v <= r) blm |= 1u << v;
              if (gxr + v >= p.nx - 1 - r) brm |= 1u << v;
            }
          }
        }
        const unsigned want_cols = (__any_sync(0xffffffffu, blm != 0) ? 4u : 0u) | (__any_sync(0xffffffffu, brm != 0) ? 8u : 0u);
        const bool out01 = (c0 >= g.hl) && (c0 < g.hl + wk), out23 = (c0 + 2 >= g.hl) && (c0 + 2 < g.hl + wk);

        // ---- vertical running sums of this window ----
        unsigned char* st_ptr = state0 + (size_t)w * 128 * ES + (size_t)c0 * ES;
        unsigned U[4];
        double V[4];
        if (yb0 == y_lo) {
#pragma unroll
          for (int v = 0; v < 4; ++v) { U[v] = 0u; V[v] = 0.0; }
        } else if (FX) {
          const uint4 q = *reinterpret_cast<const uint4*>(st_ptr);
          U[0] = q.x; U[1] = q.y; U[2] = q.z; U[3] = q.w;
        } else {
          const double2 a = *reinterpret_cast<const double2*>(st_ptr), b = *reinterpret_cast<const double2*>(st_ptr + 16);
          V[0] = a.x; V[1] = a.y; V[2] = b.x; V[3] = b.y;
        }
        // The batch is written stage by stage over its MS rows (loads, thresholds, vertical sums,
        // scans, stores, window differences) so that the independent rows interleave in the
        // instruction stream: the float64 / shuffle latencies of one row hide behind the others.
        float dn[MS][4], dold[MS][4];
        unsigned mn[MS], mo[MS];
        bool has_new[MS], has_old[MS];
#pragma unroll
        for (int i = 0; i < MS; ++i) {
          const int y = yb0 + i, yo = y - nb;             // yo: row leaving the window
          has_new[i] = (i < n) && (y <= y_hi);
          has_old[i] = (i < n) && (yo >= y_lo);
          mn[i] = mo[i] = 0x01010101u;
#pragma unroll
          for (int v = 0; v < 4; ++v) { dn[i][v] = 0.f; dold[i][v] = 0.f; }
          if (has_new[i]) {
            const unsigned sh = (unsigned)((seg_elem0 + (unsigned long long)((long long)y * p.nx)) & 3ull);
            lds4(ring0 + (unsigned)(new_slot * MS + i) * row_bytes + (unsigned)(lane_off + (int)sh) * 4u, dn[i]);
            if (M2D) {
              const unsigned short* mp = reinterpret_cast<const unsigned short*>(p.mask2d + (long long)y * p.nx + gx);
              mn[i] = (unsigned)__ldg(mp) | ((unsigned)__ldg(mp + 1) << 16);
            }
          }
          if (has_old[i]) {
            const int d = yo - y_lo;
            int so = bslot - (jb - (d >> 2));             // MS == 4
            if (so < 0) so += NCH;
            const unsigned sh = (unsigned)((seg_elem0 + (unsigned long long)((long long)yo * p.nx)) & 3ull);
            lds4(ring0 + (unsigned)(so * MS + (d & 3)) * row_bytes + (unsigned)(lane_off + (int)sh) * 4u, dold[i]);
            if (M2D) {
              const unsigned short* mp = reinterpret_cast<const unsigned short*>(p.mask2d + (long long)yo * p.nx + gx);
              mo[i] = (unsigned)__ldg(mp) | ((unsigned)__ldg(mp + 1) << 16);
            }
          }
        }
        // thresholds, masks, NaN detection, edge-band flags; per-row increments of the running sums
        unsigned dU[MS][4];
        double dV[MS][4];
#pragma unroll
        for (int i = 0; i < MS; ++i) {
          const int y = yb0 + i;
          const bool flagged = !sum_only && has_new[i] &&
                               (((want_cols & ~ne_all) != 0u) || (y <= r && !(ne_all & 1u)) ||
                                (y >= p.ny - 1 - r && !(ne_all & 2u)));
          unsigned nm = 0;                                // columns holding a value != 1 (or masked)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const bool in_need = (okm >> v) & 1u;
            const bool okn = in_need && has_new[i] && (!M2D || ((mn[i] >> (8 * v)) & 0xffu));
            const bool oko = in_need && has_old[i] && (!M2D || ((mo[i] >> (8 * v)) & 0xffu));
            nanacc = max_nan(nanacc, in_need ? fabsf(dn[i][v]) : 0.f);
            const float tn = okn ? truth_value_sum<TM>(dn[i][v], thr) : 0.f;
            const float to = oko ? truth_value_sum<TM>(dold[i][v], thr) : 0.f;
            if (flagged && in_need && (!okn || truth_not_one(dn[i][v], thr))) nm |= 1u << v;
            if (FX) dU[i][v] = __float2uint_rn(tn * fx_scale) - __float2uint_rn(to * fx_scale);
            else dV[i][v] = (double)tn - (double)to;
          }
          if (flagged) {
            unsigned bits = 0;
            if (nm & blm) bits |= 4u;
            if (nm & brm) bits |= 8u;
            if (nm && y <= r) bits |= 1u;
            if (nm && y >= p.ny - 1 - r) bits |= 2u;
            ne_all |= __reduce_or_sync(0xffffffffu, bits);
          }
        }
        // vertical running sums (sequential over the rows, independent over the columns) and the
        // lane-local inclusive prefix of every row
        unsigned LU[MS][4];
        double LV[MS][4];
#pragma unroll
        for (int i = 0; i < MS; ++i) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (FX) { U[v] += dU[i][v]; LU[i][v] = U[v]; }
            else { V[v] += dV[i][v]; LV[i][v] = V[v]; }
          }
        }
        // save the running sums for the next batch
        if (FX) *reinterpret_cast<uint4*>(st_ptr) = make_uint4(U[0], U[1], U[2], U[3]);
        else {
          *reinterpret_cast<double2*>(st_ptr) = make_double2(V[0], V[1]);
          *reinterpret_cast<double2*>(st_ptr + 16) = make_double2(V[2], V[3]);
        }
#pragma unroll
        for (int v = 1; v < 4; ++v) {
#pragma unroll
          for (int i = 0; i < MS; ++i) {
            if (FX) LU[i][v] += LU[i][v - 1];
            else LV[i][v] += LV[i][v - 1];
          }
        }
        // warp scans of the lane totals, MS rows interleaved
        unsigned iu[MS];
        double iv[MS];
#pragma unroll
        for (int i = 0; i < MS; ++i) { iu[i] = LU[i][3]; iv[i] = LV[i][3]; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
          for (int i = 0; i < MS; ++i) {
            if (FX) {
              const unsigned o2 = __shfl_up_sync(0xffffffffu, iu[i], d);
              iu[i] += (lane >= d) ? o2 : 0u;
            } else {
              const double o2 = __shfl_up_sync(0xffffffffu, iv[i], d);
              iv[i] += (lane >= d) ? o2 : 0.0;
            }
          }
        }
        // prefix rows.  Entry e of a row lives in sub-array (e + PO) & 3 at position (e + PO) >> 2:
        // for a fixed column offset the 32 lanes then touch consecutive words (no bank conflicts),
        // both for the four stores of a lane and for the window-difference loads below.
        const unsigned sub_bytes = (unsigned)(g.pw >> 2) * ES;
#pragma unroll
        for (int i = 0; i < MS; ++i) {
          if (i < n && yb0 + i - r >= ya) {
            const unsigned pr = prefix_u32 + (unsigned)i * prow_bytes;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int e = c0 + 1 + v + PO;
              const unsigned a = pr + (unsigned)(e & 3) * sub_bytes + (unsigned)(e >> 2) * ES;
              if (FX) {
                const unsigned val = iu[i] - LU[i][3] + LU[i][v];
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(val) : "memory");
              } else {
                const double val = iv[i] - LV[i][3] + LV[i][v];
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(val) : "memory");
              }
            }
          }
        }
        __syncwarp();
        if (out01 || out23) {
          double rc[4];                                    // per-column factor (unmasked launches)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int xo = gxw + v;
            const int cols_in = wrap ? nb : (min(xo + r, p.nx - 1) - max(xo - r, 0) + 1);
            rc[v] = sum_only ? fx_inv : rtab[min(max(cols_in, 1), nb)] * fx_inv;
          }
          // two rows at a time: eight independent load -> difference -> scale chains per lane
#pragma unroll
          for (int i0 = 0; i0 < MS; i0 += 2) {
            unsigned hu[2][4], lu[2][4];
            double hv[2][4], lv[2][4];
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
              const int i = i0 + ii;
              if (i < n && yb0 + i - r >= ya) {
                const unsigned pr = prefix_u32 + (unsigned)i * prow_bytes;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                  const int eh = c0 + v + r + 1 + PO, el = c0 + v - r + PO;
                  const unsigned a_hi = pr + (unsigned)(eh & 3) * sub_bytes + (unsigned)(eh >> 2) * ES;
                  const unsigned a_lo = pr + (unsigned)(el & 3) * sub_bytes + (unsigned)(el >> 2) * ES;
                  if (FX) {
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hu[ii][v]) : "r"(a_hi));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lu[ii][v]) : "r"(a_lo));
                  } else {
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(hv[ii][v]) : "r"(a_hi));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(lv[ii][v]) : "r"(a_lo));
                  }
                }
              }
            }
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
              const int i = i0 + ii;
              const int yo = yb0 + i - r;
              if (i < n && yo >= ya) {
                const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
                const double rrow = sum_only ? 1.0 : rtab[rows_in];
                const long long pt0 = (long long)yo * p.nx + gxw;
                float o[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                  const double Sw = FX ? (double)(hu[ii][v] - lu[ii][v]) : hv[ii][v] - lv[ii][v];
                  const bool ov = v < 2 ? out01 : out23;
                  if (!ov) o[v] = 0.f;
                  else if (M2D && !sum_only) o[v] = (float)(Sw * p.rcp_count[pt0 + v] * fx_inv);
                  else o[v] = (float)(Sw * rrow * rc[v]);
                }
                float* op = p.out + obase0 + pt0;
                if (out01) *reinterpret_cast<float2*>(op) = make_float2(o[0], o[1]);
                if (out23) *reinterpret_cast<float2*>(op + 2) = make_float2(o[2], o[3]);
                if (p.out_mask) {
#pragma unroll
                  for (int v = 0; v < 4; ++v)
                    if (v < 2 ? out01 : out23) p.out_mask[obase0 + pt0 + v] = (M2D && p.mask2d[pt0 + v] == 0) ? 1 : 0;
                }
              }
            }
          }
        }
        __syncwarp();                                      // prefix rows are rewritten by the next window / batch
      }

      // ---- this batch is done: hand back the chunks whose rows have all left the window ----
      if (has_real) { if (++new_slot == NCH) { new_slot = 0; new_parity ^= 1u; } }
      if (++bslot == NCH) bslot = 0;
      {
        const int next_old = yb0 + MS - nb;                // first row the next batch subtracts
        int target = next_old > y_lo ? (next_old - y_lo) / MS : 0;
        target = min(target, n_real);
        if (yb0 + MS > y_end) target = n_real;              // last batch of the piece: everything goes
        if (released < target) {
          __syncwarp();
          for (; released < target; ++released) {
            if (lane == 0) mbar_arrive(empty0 + 8 * rel_slot);
            rel_slot = (rel_slot + 1 == NCH) ? 0 : rel_slot + 1;
          }
        }
      }
    }

    // ---- edge-band flags of this slice ----
    if (!sum_only && lane == 0 && ne_all) atomicOr(p.edge_flags + (plane * p.n_thr + t), ne_all);
  }

  if (__any_sync(0xffffffffu, __isnanf(nanacc))) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
}

int launch_square_rr(const SqParams& p, const SqRrGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st);

}  // namespace nbh
