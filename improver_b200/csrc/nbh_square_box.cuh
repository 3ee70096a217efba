// Square neighbourhood of single-member slices (NeighbourhoodProcessing.process itself,
// improver/nbhood/nbhood.py:244-317 + boxsum, improver/utilities/neighbourhood_tools.py:129-211):
// row-streaming box kernel with register-resident horizontal windows.
//
// One persistent block per SM slot (two per SM) owns a contiguous share of the (slice, y) rows.  A
// producer thread streams batches of B rows into a ring of shared-memory stages with bulk
// asynchronous copies (cp.async.bulk / mbarrier, as square_rs_kernel); every consumer warp owns a
// window of 128 columns, FOUR consecutive columns per lane, and marches down the rows:
//
//   vertical window  : the (thresholded, masked) values are turned into exact fixed point
//                      (q = v * 2^s, round to nearest) and each lane keeps the running column
//                      sums V += q_new - q_old of its four columns in registers; the last 2r+1
//                      quantised rows live in a warp-private shared-memory ring (16-byte
//                      accesses, conflict free);
//   horizontal window: compile-time radius R <= 8: a lane needs the column sums of columns
//                      x0-R .. x0+3+R, i.e. its own four and R on either side from the one or
//                      two neighbouring lanes: 2R warp shuffles, then one 2R+1-term sum and three
//                      sliding updates (IADD3) — no scan, no prefix row, no barrier.  ceil(R/4)
//                      lanes at either end of a window are halo lanes (112 useful columns per
//                      window for R = 5);
//   normalise        : S / (rows_in * cols_in) as a correctly rounded float32 quotient
//                      (reciprocal + two FMAs), or S / count[y][x] for an external mask.
//
// The rows of a piece away from the top / bottom bands run in a tight loop of their own (fast_run)
// that touches none of the piece bookkeeping: the kernel is capped at 80 registers, the bookkeeping
// spills, and spill reloads inside the per-batch loop miss the (almost absent) L1 and were the
// dominant stall.
//
// Unsigned 32-bit arithmetic throughout: (2r+1)^2 * 2^s < 2^32 (s <= 22), so window sums are exact
// for the quantised values whatever row a share or a y-band starts at (bit-identical bands), a
// window of ones gives exactly 1 and a window of zeros exactly 0.  The quantisation contributes at
// most 2^-(s+1) (1e-7) to a mean.  Raw slices are processed on the assumption that they are
// probabilities in [0, 1]; the kernel records the extremes of every slice and of its results, and
// the host queues (device-side guarded) follow-ups: slices outside [0, 1] are recomputed by the
// float64 kernels, slices whose results left [min, max] by the quantisation are clamped
// (nbhood.py:312).
#pragma once
#include <type_traits>

#include "nbh_square_rs.cuh"

namespace nbh {

constexpr int kBoxProducers = 1;         // one thread issues a bulk copy every ~500 cycles: a stage of B rows
                                         // per 500 cycles is several times what the consumers can take
constexpr int kBoxK = 4;                     // consecutive columns per lane (8 was measured: same time raw, slower masked)
constexpr int kBoxWinCols = 32 * kBoxK;      // columns a warp loads
constexpr int kBoxMaxWindows = kBoxK == 8 ? 6 : 10;   // consumer warps per block
constexpr int kBoxThreads = (kBoxMaxWindows + kBoxProducers) * 32;
// halo lanes on each side of a window for radius r: the lanes whose columns only feed their neighbours
__host__ __device__ constexpr int box_halo_lanes(int r) { return (r + kBoxK - 1) / kBoxK; }
__host__ __device__ constexpr int box_useful_cols(int r) { return (32 - 2 * box_halo_lanes(r)) * kBoxK; }
constexpr int kBoxMaxStages = 8;
constexpr int kBoxBPlain = 2, kBoxBMasked = 2;   // rows per stage / batch (four rows spill registers)

// per-slice extremes recorded by the box kernel (ordered-int keys, see float_key)
struct BoxStats {
  int in_min, in_max;        // raw data (before the threshold; the truth value is monotone in it)
  int out_min, out_max;      // results
};

struct SqBoxGeom {
  int n_windows;        // consumer warps = windows per x panel
  int n_panels;
  int n_windows_total;
  int B;                // rows per stage
  int n_stages;
  int contiguous;       // 1: a stage is ONE copy of B consecutive full rows (n_panels == 1)
  int seg_pitch;        // floats per staged row segment (segmented staging), multiple of 4
  int stage_floats;     // floats per stage (data rows, then the mask rows for M2D)
  int mask_off;         // float offset of the mask rows (uint8, as in mask2d) inside a stage
  int mask_pitch;       // bytes per staged mask row segment (segmented staging), multiple of 16
  long long rows_total; // n_units * n_panels * ny
  long long share;
  int fx_shift;
  BoxStats* stats;      // [n_units]
};

__device__ __forceinline__ int box_float_key(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}

__device__ __forceinline__ float max3_nan(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ void lds_f4(unsigned addr, float* v) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr));
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}

// Barrier waits of the consumers are plain try_wait loops (the hardware blocks the thread for a
// short, bounded time per try): a suspend-time hint turned every miss into a ~1 us sleep, which with
// a handful of stages in flight serialised the producer / consumer hand-over (ncu r2c).  The producer
// thread, which has stages of slack, sleeps between tries (mbar_wait_backoff).

// tv * 2^s -> unsigned, round to nearest even exactly like cvt.rni, without the conversion unit
// (which the output side and the reciprocals keep busy): for 0 <= tv * 2^s <= 2^22, adding
// 1.5 * 2^23 leaves the rounded integer in the low mantissa bits (one FFMA, one integer subtract).
__device__ __forceinline__ unsigned box_to_fixed(float tv, float fx_scale) {
  return __float_as_uint(__fmaf_rn(tv, fx_scale, 12582912.0f)) - 0x4B400000u;
}

// K consecutive values of a lane (K = 4 or 8): 8-byte shared loads, 16-byte ring accesses
template <int K> __device__ __forceinline__ void box_lds(unsigned addr, float* v) {
  lds_f4(addr, v);
  if (K == 8) lds_f4(addr + 16u, v + 4);
}
// K mask bytes of a lane from a staged uint8 mask row, as 0 / 1 floats.  The byte address has any
// alignment (rows of an odd multiple of two bytes), the same for every lane of the warp: aligned
// words and a funnel shift.
template <int K> __device__ __forceinline__ void box_lds_mask(unsigned addr, float* mk) {
  const unsigned a4 = addr & ~3u, sh = (addr & 3u) * 8u;
  unsigned w[K / 4 + 1];
#pragma unroll
  for (int c = 0; c <= K / 4; ++c) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[c]) : "r"(a4 + 4u * c));
#pragma unroll
  for (int c = 0; c < K / 4; ++c) {
    const unsigned v = __funnelshift_r(w[c], w[c + 1], sh);
#pragma unroll
    for (int b = 0; b < 4; ++b) mk[4 * c + b] = (v & (0xffu << (8 * b))) ? 1.f : 0.f;
  }
}
template <int K> __device__ __forceinline__ void box_ring_ld(unsigned ra, unsigned* o) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(ra) : "memory");
  if (K == 8)
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) : "r"(ra) : "memory");
}
template <int K> __device__ __forceinline__ void box_ring_st(unsigned ra, const unsigned* q) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
  if (K == 8)
    asm volatile("st.shared.v4.u32 [%0+16], {%1, %2, %3, %4};" ::"r"(ra), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
}
template <int K> __device__ __forceinline__ float box_max_nan(const float* v, float acc) {
  acc = max3_nan(acc, v[0], v[1]);
  acc = max3_nan(acc, v[2], v[3]);
  if (K == 8) { acc = max3_nan(acc, v[4], v[5]); acc = max3_nan(acc, v[6], v[7]); }
  return acc;
}
template <int K> __device__ __forceinline__ float box_min(const float* v, float acc) {
  acc = min3f(acc, v[0], v[1]);
  acc = min3f(acc, v[2], v[3]);
  if (K == 8) { acc = min3f(acc, v[4], v[5]); acc = min3f(acc, v[6], v[7]); }
  return acc;
}
template <int K> __device__ __forceinline__ float box_max(const float* v, float acc) {
  acc = max3f(acc, v[0], v[1]);
  acc = max3f(acc, v[2], v[3]);
  if (K == 8) { acc = max3f(acc, v[4], v[5]); acc = max3f(acc, v[6], v[7]); }
  return acc;
}

template <int R, int TM, bool M2D>
__global__ void __launch_bounds__(kBoxThreads, 2)
square_box_kernel(const __grid_constant__ SqParams p, const SqBoxGeom g) {
  static_assert(R >= 1 && R <= 8, "compile-time radius of the shuffle window");
  constexpr int K = kBoxK, B = M2D ? kBoxBMasked : kBoxBPlain, NB = 2 * R + 1;
  constexpr int HL = box_halo_lanes(R), USEFUL = box_useful_cols(R);
  constexpr unsigned FULLBITS = (1u << K) - 1u;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int NS = g.n_stages, NW = g.n_windows;

  // layout: stages | full[8] empty[8] | per consumer warp: ring [NB][256] u32
  float* stages = reinterpret_cast<float*>(smem_raw);
  const unsigned stage_bytes = (unsigned)g.stage_floats * 4u;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)NS * stage_bytes);
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + kBoxMaxStages);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full0 + 8 * s, kBoxProducers); mbar_init(empty0 + 8 * s, NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();      // the only block-level barrier

  const unsigned long long in_elem0 = reinterpret_cast<unsigned long long>(p.in) >> 2;
  const unsigned long long mk_byte0 = M2D ? reinterpret_cast<unsigned long long>(p.mask2d) : 0ull;
  const int NBP = g.n_windows_total;

  // ======================= producer warps =======================
  if (warp >= NW) {
    if (lane != 0) return;
    const int pw = warp - NW;
    const unsigned stage0 = smem_u32(stages);
    int s = 0;
    unsigned round = 0;
    RsPieces pieces(g.share, g.rows_total, p.ny, 0);
    pieces.g0 = (long long)blockIdx.x * g.share;
    pieces.g1 = min(g.rows_total, pieces.g0 + g.share);
    pieces.g = pieces.g0;
    long long task; int ya, yb;
    while (pieces.next(&task, &ya, &yb)) {
      const long long unit = task / g.n_panels;
      const int panel = (int)(task - unit * g.n_panels);
      const long long plane = unit / p.n_thr;
      const int y_lo = max(ya - R, 0), y_hi = min(yb - 1 + R, p.ny - 1);
      const int n_in = y_hi - y_lo + 1;
      const int gw0 = panel * NW, gw1 = min(gw0 + NW, NBP) - 1;
      const int seg0 = max(gw0 * USEFUL - HL * K, 0);
      const int seg1 = min(p.nx, gw1 * USEFUL - HL * K + kBoxWinCols);
      const float* plane_ptr = p.in + plane * p.plane_stride + seg0;
      const uint8_t* mask_ptr = M2D ? p.mask2d + seg0 : nullptr;
      for (int k = 0; k < n_in; k += B) {
        const int n = min(B, n_in - k);
        const int y0 = y_lo + k;
        mbar_wait_backoff(empty0 + 8 * s, (round & 1u) ^ 1u, 200u);
        const unsigned dst0 = stage0 + (unsigned)s * stage_bytes;
        if (g.contiguous) {
          // one copy of n full rows (+ one for the mask rows); the producers take the stages in turn
          if ((int)((round * (unsigned)NS + (unsigned)s) % kBoxProducers) == pw) {
            const float* row = plane_ptr + (long long)y0 * p.nx;
            const unsigned sh = (unsigned)(reinterpret_cast<unsigned long long>(row) & 15ull);
            const unsigned bytes = (sh + (unsigned)(n * p.nx) * 4u + 15u) & ~15u;
            unsigned total = bytes, mbytes = 0, msh = 0;
            const uint8_t* mrow = nullptr;
            if (M2D) {
              mrow = mask_ptr + (long long)y0 * p.nx;
              msh = (unsigned)(reinterpret_cast<unsigned long long>(mrow) & 15ull);
              mbytes = (msh + (unsigned)(n * p.nx) + 15u) & ~15u;
              total += mbytes;
            }
            mbar_expect_tx(full0 + 8 * s, total);
            bulk_g2s(dst0, reinterpret_cast<const char*>(row) - sh, bytes, full0 + 8 * s);
            if (M2D) bulk_g2s(dst0 + (unsigned)g.mask_off * 4u, reinterpret_cast<const char*>(mrow) - msh, mbytes,
                              full0 + 8 * s);
          } else {
            mbar_arrive(full0 + 8 * s);
          }
        } else {
          // one copy per row segment; producer pw issues rows pw, pw + 4, ...
          unsigned total = 0;
          for (int j = pw; j < n; j += kBoxProducers) {
            const float* row = plane_ptr + (long long)(y0 + j) * p.nx;
            const unsigned sh = (unsigned)(reinterpret_cast<unsigned long long>(row) & 15ull);
            total += (sh + (unsigned)(seg1 - seg0) * 4u + 15u) & ~15u;
            if (M2D) {
              const uint8_t* mrow = mask_ptr + (long long)(y0 + j) * p.nx;
              const unsigned msh = (unsigned)(reinterpret_cast<unsigned long long>(mrow) & 15ull);
              total += (msh + (unsigned)(seg1 - seg0) + 15u) & ~15u;
            }
          }
          if (total) mbar_expect_tx(full0 + 8 * s, total); else mbar_arrive(full0 + 8 * s);
          for (int j = pw; j < n; j += kBoxProducers) {
            const float* row = plane_ptr + (long long)(y0 + j) * p.nx;
            const unsigned sh = (unsigned)(reinterpret_cast<unsigned long long>(row) & 15ull);
            const unsigned bytes = (sh + (unsigned)(seg1 - seg0) * 4u + 15u) & ~15u;
            bulk_g2s(dst0 + (unsigned)(j * g.seg_pitch) * 4u, reinterpret_cast<const char*>(row) - sh, bytes,
                     full0 + 8 * s);
            if (M2D) {
              const uint8_t* mrow = mask_ptr + (long long)(y0 + j) * p.nx;
              const unsigned msh = (unsigned)(reinterpret_cast<unsigned long long>(mrow) & 15ull);
              const unsigned mbytes = (msh + (unsigned)(seg1 - seg0) + 15u) & ~15u;
              bulk_g2s(dst0 + (unsigned)g.mask_off * 4u + (unsigned)(j * g.mask_pitch),
                       reinterpret_cast<const char*>(mrow) - msh, mbytes, full0 + 8 * s);
            }
          }
        }
        if (++s == NS) { s = 0; ++round; }
      }
    }
    return;
  }

  // ======================= consumer warps =======================
  unsigned* ring = reinterpret_cast<unsigned*>(bars + 2 * kBoxMaxStages) + (size_t)warp * NB * kBoxWinCols;
  const unsigned ring_u32 = smem_u32(ring) + (unsigned)lane * (K * 4u);
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
  const bool want_stats = !sum_only || TM == 0;
  const float fx_scale = (float)(1u << g.fx_shift);
  const unsigned ONE = 1u << g.fx_shift;
  const float fx_inv = 1.0f / fx_scale;
  const float nf_full = ((float)NB * fx_scale) * (float)NB;    // divisor of a window wholly inside the domain
  float rc_full;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc_full) : "f"(nf_full));
  const unsigned stages_u32 = smem_u32(stages);
  const float kInf = __int_as_float(0x7f800000);
  float nanacc = 0.f;
  int s = 0;
  unsigned parity = 0, bar = full0, st_addr = stages_u32;

  RsPieces pieces(g.share, g.rows_total, p.ny, 0);
  pieces.g0 = (long long)blockIdx.x * g.share;
  pieces.g1 = min(g.rows_total, pieces.g0 + g.share);
  pieces.g = pieces.g0;
  long long task; int ya, yb;
  while (pieces.next(&task, &ya, &yb)) {
    const long long unit = task / g.n_panels;
    const int panel = (int)(task - unit * g.n_panels);
    const int gw = panel * NW + warp;                       // window index within the full row
    const int gw0 = panel * NW;
    const int seg0 = max(gw0 * USEFUL - HL * K, 0);          // first staged column of the panel
    const int x0 = gw * USEFUL - HL * K + lane * K;          // first column of this lane
    const bool win_active = gw < NBP;
    // Columns of this lane inside the domain.  Every lane reads the eight values at its own
    // columns (a lane without any valid column reads the first staged group instead); what lies
    // beyond the end of a row is other staged data, it is masked out of the sums (colmask) and
    // kept out of the statistics (partial lanes take the predicated path).
    unsigned okbits = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) okbits |= (win_active && x0 + i >= 0 && x0 + i < p.nx) ? (1u << i) : 0u;
    const bool lane_full = okbits == FULLBITS;
    const bool warp_edge = __any_sync(0xffffffffu, !lane_full);
    const int x_ld = okbits ? x0 : seg0;
    const bool out_lane = lane >= HL && lane < 32 - HL && okbits != 0;
    const bool band_l = !sum_only && okbits && x0 <= R;
    const bool band_r = !sum_only && okbits && x0 + K - 1 >= p.nx - 1 - R;
    const long long plane = unit / p.n_thr;
    const int t = (int)(unit - plane * p.n_thr);
    const ThrDev& thr = p.thr.t[TM ? t : 0];
    const int n_rows = (yb - ya) + 2 * R;
    const int y_first = ya - R;
    const long long obase0 = unit * (long long)p.ny * p.nx;
    const unsigned long long plane_elem0 = in_elem0 + (unsigned long long)(plane * p.plane_stride) + (unsigned long long)seg0;
    const unsigned lane_col_b = (unsigned)(x_ld - seg0) * 4u;
    const bool lane_interior = lane_full && (x0 - R >= 0) && (x0 + K - 1 + R <= p.nx - 1);
    const bool fast_rows_ok = !sum_only && (M2D || p.out_mask == nullptr);   // (an unmasked launch that wants a mask plane takes the general path)
    const bool warp_inner = __all_sync(0xffffffffu, lane_interior);
    float* const out_x0 = p.out + obase0 + x0;
    // divisors (and reciprocals) of the lane's columns for windows with all 2r+1 rows inside the domain
    float nf_col[K], rc_col[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int xo = x0 + i;
      const int cols_in = max(min(xo + R, p.nx - 1) - max(xo - R, 0) + 1, 1);
      nf_col[i] = ((float)NB * fx_scale) * (float)cols_in;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc_col[i]) : "f"(nf_col[i]));
    }

    // zero the ring (rows above the domain are zeros)
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const unsigned zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      box_ring_st<K>(ring_u32 + (unsigned)q * (kBoxWinCols * 4u), zero);
    }
    unsigned V[K];
#pragma unroll
    for (int i = 0; i < K; ++i) V[i] = 0u;
    int slot = 0;
    float dmin = kInf, dmax = -kInf, omin = kInf, omax = -kInf;
    unsigned ne_bits = 0;                                   // bit0 top, bit1 bottom, bit2 left, bit3 right
    __syncwarp();

    const int n_in_piece = min(yb - 1 + R, p.ny - 1) - max(ya - R, 0) + 1;
    int consumed = 0;
    int k = max(0, -y_first);                               // leading rows above the domain are zeros
    while (k < n_rows) {
      // ---- interior rows (most of the work): a run of batches whose rows are all data rows away
      // from the top / bottom bands, so the row count of every window is 2r+1 and the divisors are
      // per-lane constants of the piece.  The run is one tight loop that touches nothing of the
      // piece bookkeeping (spilled state reloaded from local memory once per batch was the
      // largest stall of the kernel, ncu r2u).  INNER: every column of the warp has its whole
      // window inside the domain (one scalar divisor, no per-point conditions); otherwise the
      // window touches the left or right edge: columns outside the domain are zeroed, the
      // divisors follow the column count and the column bands of the trimming fix-up are watched.
      if (fast_rows_ok && k >= 2 * R) {
        const int yy_run = y_first + k;
        const int last_start = p.ny - 1 - R - B;              // last row a fast batch may start at
        int nfast = 0;
        if (yy_run >= 2 * R && yy_run <= last_start)
          nfast = min((last_start - yy_run) / B + 1, (n_in_piece - consumed) / B);
        if (nfast > 0) {
          auto fast_run = [&](auto inner_tag) {
            constexpr bool INNER = decltype(inner_tag)::value;
            unsigned a0 = (unsigned)((plane_elem0 + (unsigned long long)((long long)yy_run * p.nx)) & 3ull);
            float* op = out_x0 + (long long)(yy_run - R) * p.nx;
            const long long pt_run = (long long)(yy_run - R) * p.nx + x0;   // first output point of the lane
            const float* cp = M2D ? p.cnt_plane + pt_run : nullptr;
            const uint8_t* mp = (M2D && p.out_mask) ? p.mask2d + pt_run : nullptr;
            uint8_t* omp = (M2D && p.out_mask) ? p.out_mask + obase0 + pt_run : nullptr;
            unsigned am0 = M2D ? (unsigned)((mk_byte0 + (unsigned long long)((long long)yy_run * p.nx + seg0)) & 15ull) : 0u;
            const unsigned lane_col = (unsigned)(x_ld - seg0), mpitch = (unsigned)g.mask_pitch;
            const unsigned nxu = (unsigned)p.nx, pitch = (unsigned)g.seg_pitch;
            const bool contiguous = g.contiguous != 0;
            for (int it = 0; it < nfast; ++it) {
              float d[B][K];
              float mk[M2D ? B : 1][K];
              float cnt[M2D ? B : 1][K];
              if (M2D) {
                // the counts of the output rows: requested before the wait, used at the very end
#pragma unroll
                for (int j = 0; j < B; ++j) {
                  if (out_lane && (INNER || lane_full)) {
#pragma unroll
                    for (int i = 0; i < K; i += 2) {
                      const float2 c2 = __ldg(reinterpret_cast<const float2*>(cp + j * p.nx + i));
                      cnt[j][i] = c2.x; cnt[j][i + 1] = c2.y;
                    }
                  } else {
#pragma unroll
                    for (int i = 0; i < K; ++i)
                      cnt[j][i] = (out_lane && ((okbits >> i) & 1u)) ? __ldg(cp + j * p.nx + i) : 1.f;
                  }
                }
              }
              mbar_wait(bar, parity);
#pragma unroll
              for (int j = 0; j < B; ++j) {
                const unsigned e = a0 + (unsigned)j * nxu;
                const unsigned off = contiguous ? e : ((e & 3u) + (unsigned)j * pitch);
                box_lds<K>(st_addr + off * 4u + lane_col_b, &d[j][0]);
                if (M2D) {
                  const unsigned em = am0 + (unsigned)j * nxu;
                  const unsigned moff = contiguous ? em : ((em & 15u) + (unsigned)j * mpitch);
                  box_lds_mask<K>(st_addr + (unsigned)g.mask_off * 4u + moff + lane_col, &mk[j][0]);
                }
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(bar + 8 * kBoxMaxStages);
              bar += 8; st_addr += stage_bytes;
              if (++s == NS) { s = 0; parity ^= 1u; bar = full0; st_addr = stages_u32; }
              a0 = (a0 + (unsigned)B * nxu) & 3u;
              if (M2D) am0 = (am0 + (unsigned)B * nxu) & 15u;
              if (INNER || lane_full) {
#pragma unroll
                for (int j = 0; j < B; ++j) { dmax = box_max_nan<K>(d[j], dmax); dmin = box_min<K>(d[j], dmin); }
              } else if (okbits) {
#pragma unroll
                for (int j = 0; j < B; ++j)
#pragma unroll
                  for (int i = 0; i < K; ++i)
                    if ((okbits >> i) & 1u) { dmax = max_nan(dmax, d[j][i]); dmin = fminf(dmin, d[j][i]); }
              }
              unsigned q[B][K];
#pragma unroll
              for (int j = 0; j < B; ++j)
#pragma unroll
                for (int i = 0; i < K; ++i) {
                  float tv = truth_value_sum<TM>(d[j][i], thr);
                  if (M2D) tv *= mk[j][i];
                  q[j][i] = box_to_fixed(tv, fx_scale);
                  if (!INNER) q[j][i] = ((okbits >> i) & 1u) ? q[j][i] : 0u;
                }
              if (!INNER) {
                if ((band_l && !(ne_bits & 4u)) || (band_r && !(ne_bits & 8u))) {
#pragma unroll
                  for (int j = 0; j < B; ++j)
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                      const int xo = x0 + i;
                      if (((okbits >> i) & 1u) && q[j][i] != ONE) {
                        if (xo <= R) ne_bits |= 4u;
                        if (xo >= p.nx - 1 - R) ne_bits |= 8u;
                      }
                    }
                }
              }
              unsigned Vs[B][K];
              if (NB >= B) {
                unsigned o[B][K];
                unsigned ra[B];
#pragma unroll
                for (int j = 0; j < B; ++j) {
                  ra[j] = ring_u32 + (unsigned)slot * (kBoxWinCols * 4u);
                  slot = (slot + 1 == NB) ? 0 : slot + 1;
                }
#pragma unroll
                for (int j = 0; j < B; ++j) box_ring_ld<K>(ra[j], o[j]);
#pragma unroll
                for (int j = 0; j < B; ++j) {
                  box_ring_st<K>(ra[j], q[j]);
#pragma unroll
                  for (int i = 0; i < K; ++i) { V[i] += q[j][i] - o[j][i]; Vs[j][i] = V[i]; }
                }
              } else {
#pragma unroll
                for (int j = 0; j < B; ++j) {
                  const unsigned ra = ring_u32 + (unsigned)slot * (kBoxWinCols * 4u);
                  unsigned o[K];
                  box_ring_ld<K>(ra, o);
                  box_ring_st<K>(ra, q[j]);
#pragma unroll
                  for (int i = 0; i < K; ++i) { V[i] += q[j][i] - o[i]; Vs[j][i] = V[i]; }
                  slot = (slot + 1 == NB) ? 0 : slot + 1;
                }
              }
#pragma unroll
            for (int j = 0; j < B; ++j) {
              unsigned E[K + 2 * R];
#pragma unroll
              for (int i = 0; i < R; ++i) {
                const int ol = R - i, dl = (ol + K - 1) / K;
                const int orr = i + 1, dr = (orr + K - 1) / K;
                E[i] = __shfl_up_sync(0xffffffffu, Vs[j][dl * K - ol], dl);
                E[K + R + i] = __shfl_down_sync(0xffffffffu, Vs[j][i - (dr - 1) * K], dr);
              }
#pragma unroll
              for (int i = 0; i < K; ++i) E[R + i] = Vs[j][i];
              unsigned part[3] = {0u, 0u, 0u};
#pragma unroll
              for (int c = 0; c <= 2 * R; ++c) part[c % 3] += E[c];
              unsigned H[K];
              H[0] = part[0] + part[1] + part[2];
#pragma unroll
              for (int i = 1; i < K; ++i) H[i] = H[i - 1] + E[i + 2 * R] - E[i - 1];
              float o[K];
#pragma unroll
              for (int i = 0; i < K; ++i) {
                float nf, rc;
                if (M2D) {
                  nf = cnt[j][i] * fx_scale;
                  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
                } else {
                  nf = INNER ? nf_full : nf_col[i];
                  rc = INNER ? rc_full : rc_col[i];
                }
                const float Sf = (float)H[i];
                const float q0 = Sf * rc;
                const float rem = __fmaf_rn(-q0, nf, Sf);
                o[i] = __fmaf_rn(rem, rc, q0);
                if (M2D) o[i] = cnt[j][i] == 0.f ? __int_as_float(0x7fc00000) : o[i];
              }
              if (out_lane) {
                if (INNER || lane_full) {
                  omax = box_max<K>(o, omax);
                  omin = box_min<K>(o, omin);
#pragma unroll
                  for (int i = 0; i < K; i += 2) *reinterpret_cast<float2*>(op + i) = make_float2(o[i], o[i + 1]);
                } else {
#pragma unroll
                  for (int i = 0; i < K; ++i)
                    if ((okbits >> i) & 1u) { omax = fmaxf(omax, o[i]); omin = fminf(omin, o[i]); op[i] = o[i]; }
                }
                if (omp) {
#pragma unroll
                  for (int i = 0; i < K; ++i)
                    if ((okbits >> i) & 1u) omp[i] = (M2D && mp[i] == 0) ? 1 : 0;
                }
              }
              op += p.nx;
              if (M2D) { cp += p.nx; if (mp) mp += p.nx; }
              if (omp) omp += p.nx;
            }
            }
          };
          if (warp_inner) fast_run(std::true_type{});
          else fast_run(std::false_type{});
          k += nfast * B;
          consumed += nfast * B;
          continue;
        }
      }

      const int yy0 = y_first + k;
      const bool has_data = yy0 < p.ny;
      int nbr;
      float d[B][K];
      float mk[M2D ? B : 1][K];
      if (has_data) {
        nbr = min(B, n_in_piece - consumed);
        mbar_wait(bar, parity);
        const unsigned long long e0 = plane_elem0 + (unsigned long long)((long long)yy0 * p.nx);
        const unsigned long long m0 = M2D ? mk_byte0 + (unsigned long long)((long long)yy0 * p.nx + seg0) : 0ull;
#pragma unroll
        for (int j = 0; j < B; ++j) {
          // rows past the end of the piece re-read the first row of the stage (ignored later)
          const int jj = j < nbr ? j : 0;
          unsigned off, moff = 0;
          if (g.contiguous) {
            off = (unsigned)(e0 & 3ull) * 4u + (unsigned)(jj * p.nx) * 4u;
            if (M2D) moff = (unsigned)(m0 & 15ull) + (unsigned)(jj * p.nx);
          } else {
            off = (unsigned)((e0 + (unsigned long long)(jj * p.nx)) & 3ull) * 4u + (unsigned)(jj * g.seg_pitch) * 4u;
            if (M2D) moff = (unsigned)((m0 + (unsigned long long)(jj * p.nx)) & 15ull) + (unsigned)(jj * g.mask_pitch);
          }
          box_lds<K>(st_addr + off + lane_col_b, &d[j][0]);
          if (M2D) box_lds_mask<K>(st_addr + (unsigned)g.mask_off * 4u + moff + (lane_col_b >> 2), &mk[j][0]);
        }
        consumed += nbr;
        __syncwarp();                                       // every lane has read the stage
        if (lane == 0) mbar_arrive(bar + 8 * kBoxMaxStages);
        bar += 8; st_addr += stage_bytes;
        if (++s == NS) { s = 0; parity ^= 1u; bar = full0; st_addr = stages_u32; }
      } else {
        nbr = min(B, n_rows - k);                           // rows below the domain: zeros
#pragma unroll
        for (int j = 0; j < B; ++j)
#pragma unroll
          for (int i = 0; i < K; ++i) { d[j][i] = 0.f; if (M2D) mk[j][i] = 0.f; }
      }

      // The batch body, specialised on FULL (every row of the batch is a data row that completes
      // an output row: no per-row conditions, so the independent rows interleave) and written as
      // one loop over the rows per step of the computation.
      auto batch = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        // ---- 1. extremes of the raw data; NaN propagates through the maximum (nbhood.py:167-168) ----
        if (has_data) {
          if (lane_full) {
#pragma unroll
            for (int j = 0; j < B; ++j) {
              if (FULL || j < nbr) {
                if (want_stats) {
                  dmax = box_max_nan<K>(d[j], dmax);
                  dmin = box_min<K>(d[j], dmin);
                } else {
                  nanacc = box_max_nan<K>(d[j], nanacc);
                }
              }
            }
          } else if (okbits) {
#pragma unroll
            for (int j = 0; j < B; ++j) {
              if (FULL || j < nbr) {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                  if ((okbits >> i) & 1u) {
                    if (want_stats) { dmax = max_nan(dmax, d[j][i]); dmin = fminf(dmin, d[j][i]); }
                    else nanacc = max_nan(nanacc, d[j][i]);
                  }
                }
              }
            }
          }
        }
        // ---- 2. truth values -> fixed point ----
        unsigned q[B][K];
#pragma unroll
        for (int j = 0; j < B; ++j) {
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float tv = truth_value_sum<TM>(d[j][i], thr);
            if (M2D) tv *= mk[j][i];
            q[j][i] = box_to_fixed(tv, fx_scale);
          }
        }
        if (warp_edge || !has_data) {
#pragma unroll
          for (int j = 0; j < B; ++j)
#pragma unroll
            for (int i = 0; i < K; ++i) q[j][i] = (has_data && ((okbits >> i) & 1u)) ? q[j][i] : 0u;
        }
        // ---- 3. edge bands of the trimming fix-up: has a value != 1 been seen? (sticky) ----
        if (!sum_only && has_data) {
          const bool cols_open = (band_l && !(ne_bits & 4u)) || (band_r && !(ne_bits & 8u));
#pragma unroll
          for (int j = 0; j < B; ++j) {
            const int yy = yy0 + j;
            const bool rows_band = yy <= R || yy >= p.ny - 1 - R;
            if ((FULL || j < nbr) && (rows_band || cols_open)) {
#pragma unroll
              for (int i = 0; i < K; ++i) {
                const int xo = x0 + i;
                if (((okbits >> i) & 1u) && q[j][i] != ONE) {
                  if (yy <= R) ne_bits |= 1u;
                  if (yy >= p.ny - 1 - R) ne_bits |= 2u;
                  if (xo <= R) ne_bits |= 4u;
                  if (xo >= p.nx - 1 - R) ne_bits |= 8u;
                }
              }
            }
          }
        }
        // ---- 4. ring exchange and vertical running sums (sequential over the rows) ----
        unsigned Vs[B][K];
        if (NB >= B) {
          // the B rows of the batch occupy distinct ring slots: all loads first, then the stores
          unsigned o[B][K];
          int sl[B];
#pragma unroll
          for (int j = 0; j < B; ++j) { sl[j] = slot; slot = (slot + 1 == NB) ? 0 : slot + 1; }
#pragma unroll
          for (int j = 0; j < B; ++j) {
            if (FULL || j < nbr) {
              const unsigned ra = ring_u32 + (unsigned)sl[j] * (kBoxWinCols * 4u);
              box_ring_ld<K>(ra, o[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < B; ++j) {
            if (FULL || j < nbr) {
              const unsigned ra = ring_u32 + (unsigned)sl[j] * (kBoxWinCols * 4u);
              box_ring_st<K>(ra, q[j]);
#pragma unroll
              for (int i = 0; i < K; ++i) { V[i] += q[j][i] - o[j][i]; Vs[j][i] = V[i]; }
            } else {
#pragma unroll
              for (int i = 0; i < K; ++i) Vs[j][i] = 0u;
            }
          }
          if (!FULL) slot = (sl[0] + nbr) % NB;             // only nbr rows entered the ring
        } else {
          // a batch longer than the ring (radius 1 with four rows): strictly row by row
#pragma unroll
          for (int j = 0; j < B; ++j) {
            if (FULL || j < nbr) {
              const unsigned ra = ring_u32 + (unsigned)slot * (kBoxWinCols * 4u);
              unsigned o[K];
              box_ring_ld<K>(ra, o);
              box_ring_st<K>(ra, q[j]);
#pragma unroll
              for (int i = 0; i < K; ++i) { V[i] += q[j][i] - o[i]; Vs[j][i] = V[i]; }
              slot = (slot + 1 == NB) ? 0 : slot + 1;
            } else {
#pragma unroll
              for (int i = 0; i < K; ++i) Vs[j][i] = 0u;
            }
          }
        }
        // ---- 5. horizontal windows from the neighbours' column sums; normalise; store ----
#pragma unroll
        for (int j = 0; j < B; ++j) {
          if (FULL || (j < nbr && k + j >= 2 * R)) {          // warp-uniform
            unsigned E[K + 2 * R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
              // column x0 - (R - i): register dl * K - (R - i) of the lane dl to the left;
              // column x0 + K + i: register i - (dr - 1) * K of the lane dr to the right
              const int ol = R - i, dl = (ol + K - 1) / K;
              const int orr = i + 1, dr = (orr + K - 1) / K;
              E[i] = __shfl_up_sync(0xffffffffu, Vs[j][dl * K - ol], dl);
              E[K + R + i] = __shfl_down_sync(0xffffffffu, Vs[j][i - (dr - 1) * K], dr);
            }
#pragma unroll
            for (int i = 0; i < K; ++i) E[R + i] = Vs[j][i];
            // H[0] as a balanced sum, then H[i] = H[i-1] + entering - leaving
            unsigned part[3] = {0u, 0u, 0u};
#pragma unroll
            for (int c = 0; c <= 2 * R; ++c) part[c % 3] += E[c];
            unsigned H[K];
            H[0] = part[0] + part[1] + part[2];
#pragma unroll
            for (int i = 1; i < K; ++i) H[i] = H[i - 1] + E[i + 2 * R] - E[i - 1];
            const int yo = y_first + k + j - R;
            const int rows_in = min(yo + R, p.ny - 1) - max(yo - R, 0) + 1;
            const long long pt0 = (long long)yo * p.nx + x0;
            float o[K];
            if (sum_only) {
#pragma unroll
              for (int i = 0; i < K; ++i) o[i] = (float)H[i] * fx_inv;
            } else if (M2D) {
#pragma unroll
              for (int i = 0; i < K; ++i) {
                const float cnt = (out_lane && ((okbits >> i) & 1u)) ? __ldg(p.cnt_plane + pt0 + i) : 1.f;
                const float nf = cnt * fx_scale;
                float rc;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
                const float Sf = (float)H[i];
                const float q0 = Sf * rc;
                const float rem = __fmaf_rn(-q0, nf, Sf);
                o[i] = cnt == 0.f ? __int_as_float(0x7fc00000) : __fmaf_rn(rem, rc, q0);
              }
            } else {
              const float nrow = (float)rows_in * fx_scale;
              const float nfi = nrow * (float)NB;
              float rci;
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rci) : "f"(nfi));
#pragma unroll
              for (int i = 0; i < K; ++i) {
                float nf = nfi, rc = rci;
                if (!lane_interior) {
                  const int xo = x0 + i;
                  nf = nrow * (float)(min(xo + R, p.nx - 1) - max(xo - R, 0) + 1);
                  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
                }
                const float Sf = (float)H[i];
                const float q0 = Sf * rc;
                const float rem = __fmaf_rn(-q0, nf, Sf);
                o[i] = __fmaf_rn(rem, rc, q0);
              }
            }
            if (out_lane) {
              if (lane_full) {
                if (!sum_only) {
                  omax = box_max<K>(o, omax);
                  omin = box_min<K>(o, omin);
                }
                float* op = p.out + obase0 + pt0;
#pragma unroll
                for (int i = 0; i < K; i += 2) *reinterpret_cast<float2*>(op + i) = make_float2(o[i], o[i + 1]);
              } else {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                  if ((okbits >> i) & 1u) {
                    if (!sum_only) { omax = fmaxf(omax, o[i]); omin = fminf(omin, o[i]); }
                    p.out[obase0 + pt0 + i] = o[i];
                  }
                }
              }
              if (p.out_mask) {
#pragma unroll
                for (int i = 0; i < K; ++i)
                  if ((okbits >> i) & 1u) p.out_mask[obase0 + pt0 + i] = (M2D && p.mask2d[pt0 + i] == 0) ? 1 : 0;
              }
            }
          }
        }
      };
      if (has_data && nbr == B && k >= 2 * R) batch(std::true_type{});
      else batch(std::false_type{});
      k += nbr;
    }

    // ---- per-slice flags and extremes ----
    if (!sum_only) {
      const unsigned bits = __reduce_or_sync(0xffffffffu, ne_bits);
      if (lane == 0 && bits) atomicOr(p.edge_flags + (plane * p.n_thr + t), bits);
    }
    if (want_stats) {
      // NaN in dmax marks the launch; the keys of the finite extremes go to the slice record
      if (__isnanf(dmax)) nanacc = dmax;
      int kmin = (__isnanf(dmin) || dmin == kInf) ? INT_MAX : box_float_key(dmin);
      int kmax = (__isnanf(dmax) || dmax == -kInf) ? INT_MIN : box_float_key(dmax);
      int komin = omin == kInf ? INT_MAX : box_float_key(omin), komax = omax == -kInf ? INT_MIN : box_float_key(omax);
      kmin = __reduce_min_sync(0xffffffffu, kmin); kmax = __reduce_max_sync(0xffffffffu, kmax);
      komin = __reduce_min_sync(0xffffffffu, komin); komax = __reduce_max_sync(0xffffffffu, komax);
      if (lane == 0) {
        BoxStats* bs = g.stats + unit;
        atomicMin(&bs->in_min, kmin); atomicMax(&bs->in_max, kmax);
        atomicMin(&bs->out_min, komin); atomicMax(&bs->out_max, komax);
      }
    }
  }
  if (__any_sync(0xffffffffu, __isnanf(nanacc))) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
}

struct BoxWorkspace {
  BoxStats* stats;
  int* redo;
  int* clamp_list;
  float2* range;
  int* n_clamp;
  int* n_redo;
};
size_t box_workspace_bytes(long long n_units, int ny, int nx);
void box_carve(void* ws, long long n_units, int ny, int nx, BoxWorkspace* w);
int box_prepare(const BoxWorkspace& w, const uint8_t* mask2d, long long n_units, int ny, int nx, cudaStream_t st);
int box_finish(const BoxWorkspace& w, const SqParams& p, const NbhSliceStats* ext, int tm, long long n_units,
               cudaStream_t st);
bool plan_square_box(const NbhDesc& d, const float* in, SqBoxGeom* g, int* grid, size_t* smem);
int launch_square_box(const SqParams& p, const SqBoxGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st);

}  // namespace nbh
