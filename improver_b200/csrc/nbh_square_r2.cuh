// Square neighbourhood of single-member slices, two-phase row-streaming kernel.
//
// The scan-based kernels spend ~50-80 instructions per point when every input row
// belongs to one slice only (no member loop to amortise the per-row prefix scan over).
// This kernel uses the classic separable running sums instead, ~20 instructions per point:
//
//   phase H  each thread owns a run of kRun consecutive columns of one new row: the window sum
//            of the first column is formed directly (2r+1 values), the others by
//            S += d[x+r] - d[x-r-1]; float32, runs restart every kRun columns so the rounding
//            error stays below ~1e-6 of the data magnitude.  The H row goes straight into a
//            shared-memory ring of 2r+1+kR2Rows rows.
//   phase V  each thread owns two adjacent columns for the whole piece and keeps their vertical
//            running sums in float64 registers: V += H[y] - H[y-2r-1], then normalise and store
//            (coalesced float2 stores).
//
// Rows arrive exactly as in square_rs_kernel's rows mode: one persistent block per SM owns a
// contiguous share of the (slice, y) rows, producer threads stream stages of kR2Rows contiguous
// rows into shared memory with cp.async.bulk + mbarriers; the 20 consumer warps synchronise
// twice per stage on a named barrier (H -> V -> next H).  Thresholds and the external 2-D mask
// are applied in place by a pre-pass over the staged rows (P = valid ? truth_value(d) : 0).
#pragma once
#include "nbh_square_rs.cuh"

namespace nbh {

constexpr int kR2Rows = 4;                 // rows per stage
constexpr int kR2Run = 9;                  // columns per phase-H run (odd: conflict-free shared-memory strides)
constexpr int kR2CWarps = 20;
constexpr int kR2CThreads = kR2CWarps * 32;
constexpr int kR2Producers = 4;
constexpr int kR2Threads = (kR2CWarps + kR2Producers) * 32;
constexpr int kR2MaxStages = 16;

struct SqR2Geom {
  int n_stages;
  int pitch;            // floats per stage (kR2Rows * nx + alignment slack, multiple of 4)
  int hpitch;           // floats per H-ring row (nx rounded up to even)
  int ring_rows;        // 2r + 1 + kR2Rows
  long long rows_total; // n_units * ny
  long long share;      // rows per block
  unsigned wait_ns;
};

__device__ __forceinline__ void r2_consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kR2CThreads) : "memory");
}

template <int TM, bool M2D>
__global__ void __launch_bounds__(kR2Threads, 1)
square_r2_kernel(const __grid_constant__ SqParams p, const SqR2Geom g) {
  constexpr bool PRE = (TM != 0) || M2D;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = p.nb, r = p.r, NS = g.n_stages, nx = p.nx, ny = p.ny;
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;

  // layout: stages | full[16] empty[16] | flags | rtab | H ring
  float* stages = reinterpret_cast<float*>(smem_raw);
  const unsigned stage_bytes = (unsigned)g.pitch * 4u;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)NS * stage_bytes);
  unsigned* sflags = reinterpret_cast<unsigned*>(bars + 2 * kR2MaxStages);       // [0] edge-band bits, [1] NaN
  double* rtab = reinterpret_cast<double*>(sflags + 4);
  float* hring = reinterpret_cast<float*>(rtab + ((nb + 2) & ~1));
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + kR2MaxStages);

  for (int k = threadIdx.x + 1; k <= nb; k += blockDim.x) rtab[k] = 1.0 / (double)k;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full0 + 8 * s, kR2Producers); mbar_init(empty0 + 8 * s, kR2CWarps); }
    sflags[0] = 0; sflags[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();      // the only block-wide barrier

  // ======================= producer warps =======================
  if (warp >= kR2CWarps) {
    if (lane != 0) return;
    const int pw = warp - kR2CWarps;
    const unsigned stage0 = smem_u32(stages);
    int s = 0;
    unsigned round = 0;
    RsPieces pieces(g.share, g.rows_total, ny, blockIdx.x);
    pieces.dir = 1; pieces.g = pieces.g0;
    long long unit; int ya, yb;
    while (pieces.next(&unit, &ya, &yb)) {
      const long long plane = unit / p.n_thr;
      const int y_lo = max(ya - r, 0), y_hi = min(yb - 1 + r, ny - 1);
      const float* plane_ptr = p.in + plane * p.plane_stride;
      for (int y0 = y_lo; y0 <= y_hi; y0 += kR2Rows) {
        const int n = min(kR2Rows, y_hi + 1 - y0);
        const float* row = plane_ptr + (long long)y0 * nx;
        const unsigned shift_b = (unsigned)(reinterpret_cast<unsigned long long>(row) & 15ull);
        const unsigned bytes = (shift_b + (unsigned)(n * nx) * 4u + 15u) & ~15u;
        mbar_wait_hint(empty0 + 8 * s, (round & 1u) ^ 1u, g.wait_ns);
        if ((int)((round * (unsigned)NS + (unsigned)s) % kR2Producers) == pw) {
          mbar_expect_tx(full0 + 8 * s, bytes);
          bulk_g2s(stage0 + (unsigned)s * stage_bytes, reinterpret_cast<const char*>(row) - shift_b, bytes, full0 + 8 * s);
        } else {
          mbar_arrive(full0 + 8 * s);
        }
        if (++s == NS) { s = 0; ++round; }
      }
    }
    return;
  }

  // ======================= consumer warps =======================
  const int ct = threadIdx.x;                       // consumer thread index 0 .. kR2CThreads - 1
  const int HP = g.hpitch, RR = g.ring_rows;
  const int runs = (nx + kR2Run - 1) / kR2Run;      // phase-H runs per row
  // phase V: this thread's two columns
  const int cx = 2 * ct;
  const bool vcol = cx < nx;
  double rc[2] = {0.0, 0.0};
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int xo = min(cx + v, nx - 1);
    const int cols_in = wrap ? nb : (min(xo + r, nx - 1) - max(xo - r, 0) + 1);
    rc[v] = sum_only ? 1.0 : rtab[min(max(cols_in, 1), nb)];
  }
  float nanacc = 0.f;
  int s = 0;
  unsigned parity = 0;
  const unsigned long long in_elem0 = reinterpret_cast<unsigned long long>(p.in) >> 2;

  RsPieces pieces(g.share, g.rows_total, ny, blockIdx.x);
  pieces.dir = 1; pieces.g = pieces.g0;
  long long unit; int ya, yb;
  while (pieces.next(&unit, &ya, &yb)) {
    const long long plane = unit / p.n_thr;
    const int t = (int)(unit - plane * p.n_thr);
    const ThrDev& thr = p.thr.t[TM ? t : 0];
    const long long obase0 = unit * (long long)ny * nx;
    const unsigned long long plane_elem0 = in_elem0 + (unsigned long long)(plane * p.plane_stride);
    const int y_lo = max(ya - r, 0), y_hi = min(yb - 1 + r, ny - 1), y_end = yb - 1 + r;

    // the H ring starts empty (rows above y_lo are zeros or do not exist)
    r2_consumer_sync();                              // previous piece's phase V is done
    for (int i = ct; i < RR * HP / 2; i += kR2CThreads) reinterpret_cast<float2*>(hring)[i] = make_float2(0.f, 0.f);
    if (ct == 0) sflags[0] = 0;
    double V0 = 0.0, V1 = 0.0;
    r2_consumer_sync();

    for (int y0 = y_lo; y0 <= y_end; y0 += kR2Rows) {
      const int sb = y0 % RR;                                  // ring slot of the batch's first row
      const int n = min(kR2Rows, y_end + 1 - y0);              // rows of this batch (the last ones may be virtual)
      const int n_real = max(0, min(n, y_hi + 1 - y0));        // rows that exist in the domain
      if (n_real > 0) {
        mbar_wait_hint(full0 + 8 * s, parity, g.wait_ns);
        const unsigned sh = (unsigned)((plane_elem0 + (unsigned long long)((long long)y0 * nx)) & 3ull);
        float* st = stages + (size_t)s * g.pitch + sh;        // row j of the stage at st + j * nx
        if (PRE) {
          // ---- pre-pass: truth value and external mask in place, NaN detection, edge flags ----
          for (int j = 0; j < n_real; ++j) {
            const int y = y0 + j;
            float* row = st + (size_t)j * nx;
            for (int x = ct; x < nx; x += kR2CThreads) {
              const float d = row[x];
              nanacc = max_nan(nanacc, fabsf(d));
              const bool ok = !M2D || p.mask2d[(long long)y * nx + x] != 0;
              row[x] = ok ? truth_value_sum<TM>(d, thr) : 0.f;
              if (!sum_only && (!ok || truth_not_one(d, thr))) {
                unsigned bits = 0;
                if (y <= r) bits |= 1u;
                if (y >= ny - 1 - r) bits |= 2u;
                if (!wrap && x <= r) bits |= 4u;
                if (!wrap && x >= nx - 1 - r) bits |= 8u;
                if (bits & ~sflags[0]) atomicOr(&sflags[0], bits);
              }
            }
          }
          r2_consumer_sync();
        }
        // ---- phase H: horizontal window sums of the new rows into the ring ----
        for (int task = ct; task < n_real * runs; task += kR2CThreads) {
          const int j = task / runs, run = task - j * runs;
          const int xa = run * kR2Run, xb = min(nx, xa + kR2Run);
          const float* row = st + (size_t)j * nx;
          const int sn = sb + j >= RR ? sb + j - RR : sb + j;
          float* hrow = hring + (size_t)sn * HP;
          const int y = y0 + j;
          const bool interior = xa - r > 0 && xb - 1 + r < nx - 1;
          const bool flag_row = !PRE && !sum_only && (y <= r || y >= ny - 1 - r);
          unsigned bits = 0;
          if (interior) {
            float S = 0.f;
            for (int dx = -r; dx <= r; ++dx) {
              const float d = row[xa + dx];
              if (!PRE) nanacc = max_nan(nanacc, fabsf(d));
              S += d;
            }
            hrow[xa] = S;
            for (int x = xa + 1; x < xb; ++x) {
              const float dn = row[x + r];
              if (!PRE) nanacc = max_nan(nanacc, fabsf(dn));
              S += dn;
              S -= row[x - r - 1];
              hrow[x] = S;
            }
            if (flag_row) {
              bool notone = false;
              for (int x = xa; x < xb; ++x) notone = notone || truth_not_one(row[x], thr);
              if (notone) bits = (y <= r ? 1u : 0u) | (y >= ny - 1 - r ? 2u : 0u);
            }
          } else {
            // runs touching the row ends: zero padding, or the periodic continuation of the row
            for (int x = xa; x < xb; ++x) {
              float S = 0.f;
              for (int dx = -r; dx <= r; ++dx) {
                int xx = x + dx;
                if (wrap) { if (xx < 0) xx += nx; else if (xx >= nx) xx -= nx; }
                if (xx >= 0 && xx < nx) {
                  const float d = row[xx];
                  if (!PRE) nanacc = max_nan(nanacc, fabsf(d));
                  S += d;
                }
              }
              hrow[x] = S;
              if (!PRE && !sum_only && truth_not_one(row[x], thr)) {
                if (y <= r) bits |= 1u;
                if (y >= ny - 1 - r) bits |= 2u;
                if (!wrap && x <= r) bits |= 4u;
                if (!wrap && x >= nx - 1 - r) bits |= 8u;
              }
            }
          }
          if (bits & ~sflags[0]) atomicOr(&sflags[0], bits);
        }
        // every lane of this warp has finished reading the stage
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == NS) { s = 0; parity ^= 1u; }
      }
      // rows below the domain are zeros
      for (int j = n_real; j < n; ++j) {
        const int sn = sb + j >= RR ? sb + j - RR : sb + j;
        float* hrow = hring + (size_t)sn * HP;
        for (int x = ct; x < HP; x += kR2CThreads) hrow[x] = 0.f;
      }
      r2_consumer_sync();

      // ---- phase V: vertical running sums, normalise, store ----
      if (vcol) {
#pragma unroll
        for (int j = 0; j < kR2Rows; ++j) {
          if (j < n) {
            const int y = y0 + j;
            const int sn = sb + j >= RR ? sb + j - RR : sb + j;
            const int so = sn - nb < 0 ? sn - nb + RR : sn - nb;
            const float2 hn = *reinterpret_cast<const float2*>(hring + (size_t)sn * HP + cx);
            float2 ho = make_float2(0.f, 0.f);
            if (y - nb >= y_lo) ho = *reinterpret_cast<const float2*>(hring + (size_t)so * HP + cx);
            V0 += (double)hn.x - (double)ho.x;
            V1 += (double)hn.y - (double)ho.y;
            const int yo = y - r;
            if (yo >= ya) {
              const int rows_in = min(yo + r, ny - 1) - max(yo - r, 0) + 1;
              const double rrow = sum_only ? 1.0 : rtab[rows_in];
              const long long pt0 = (long long)yo * nx + cx;
              float o0, o1;
              if (M2D && !sum_only) {
                o0 = (float)(V0 * p.rcp_count[pt0]);
                o1 = (float)(V1 * p.rcp_count[pt0 + 1]);
              } else {
                o0 = (float)(V0 * rrow * rc[0]);
                o1 = (float)(V1 * rrow * rc[1]);
              }
              *reinterpret_cast<float2*>(p.out + obase0 + pt0) = make_float2(o0, o1);
              if (p.out_mask) {
                p.out_mask[obase0 + pt0] = (M2D && p.mask2d[pt0] == 0) ? 1 : 0;
                p.out_mask[obase0 + pt0 + 1] = (M2D && p.mask2d[pt0 + 1] == 0) ? 1 : 0;
              }
            }
          }
        }
      }
      r2_consumer_sync();                            // ring slots are rewritten by the next batch
    }

    // ---- edge-band flags of this slice (trimming fix-up) ----
    if (!sum_only && ct == 0) {
      const unsigned bits = sflags[0] | (wrap ? 0xCu : 0u);
      if (bits) atomicOr(p.edge_flags + (plane * p.n_thr + t), bits);
    }
  }

  if (__any_sync(0xffffffffu, __isnanf(nanacc))) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
}

int launch_square_r2(const SqParams& p, const SqR2Geom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st);

}  // namespace nbh
