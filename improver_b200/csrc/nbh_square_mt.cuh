// Fused threshold -> box sum -> realization mean for SEVERAL thresholds from ONE read of the input
// (improver/threshold.py:681-695 loops the thresholds over the same cube; the reference and the
// single-threshold kernel read it once per threshold).
//
// Structure = square_rs_kernel's staging (producer threads stream the member rows of every grid
// row of an x panel into a ring of shared-memory stages with bulk asynchronous copies) with
// square_box_kernel's consumers (four consecutive columns per lane, exact fixed-point vertical
// sums through a warp-private ring, horizontal windows from 2R warp shuffles).  The consumer warps
// are a (window, threshold) grid: every staged member row is read by the warps of all the
// thresholds of the pass (up to four), each accumulating its own truth values — the input crosses
// HBM once per pass instead of once per threshold, and the per-threshold work (member loop and
// row finish) runs on different warps in parallel instead of back to back.
//
// Shared memory bounds the pass: (2R+1) ring rows per (window, threshold) warp.  Rows are split
// into x panels of five 128-column windows (560 useful columns), so a block of 20 consumer warps
// holds 20 rings of (2R+1) x 512 B plus the stage ring.
//
// The edge-band flags of the trimming fix-up are produced by a separate small kernel over the
// bands only (2 (r+1) (nx + ny) points per slice, ~2 % of the input).
#pragma once
#include "nbh_square_box.cuh"

namespace nbh {

constexpr int kMtK = 4;                       // columns per lane
constexpr int kMtWinCols = 32 * kMtK;         // 128
constexpr int kMtWindows = 5;                 // windows per x panel
constexpr int kMtMaxThr = 4;                  // thresholds per pass
constexpr int kMtProducers = 4;
constexpr int kMtThreads = (kMtWindows * kMtMaxThr + kMtProducers) * 32;
constexpr int kMtMaxStages = 16;

__host__ __device__ constexpr int mt_halo_lanes(int r) { return (r + kMtK - 1) / kMtK; }
__host__ __device__ constexpr int mt_useful_cols(int r) { return (32 - 2 * mt_halo_lanes(r)) * kMtK; }

struct SqMtGeom {
  int n_windows;        // windows per panel (<= kMtWindows)
  int n_panels;
  int n_windows_total;
  int n_thr;            // thresholds of this pass (<= kMtMaxThr)
  int t_base;           // first threshold of this pass
  int ms;               // member rows per stage; divides n_members
  int n_stages;
  int pitch;            // floats per staged member row (multiple of 4)
  long long rows_total; // n_planes * n_panels * ny
  long long share;
  int fx_shift;
};

template <int R, int TM, bool M2D>
__global__ void __launch_bounds__(kMtThreads, 1)
square_mt_kernel(const __grid_constant__ SqParams p, const SqMtGeom g) {
  static_assert(R >= 1 && R <= 8, "compile-time radius of the shuffle window");
  constexpr int K = kMtK, NB = 2 * R + 1, HL = mt_halo_lanes(R), USEFUL = mt_useful_cols(R);
  constexpr unsigned FULLBITS = (1u << K) - 1u;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int NS = g.n_stages, MS = g.ms, NW = g.n_windows, NC = g.n_windows * g.n_thr, Rm = p.n_members;

  // layout: stages | full[16] empty[16] | per consumer warp: ring [NB][128] u32
  float* stages = reinterpret_cast<float*>(smem_raw);
  const unsigned stage_bytes = (unsigned)(MS * g.pitch * 4);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)NS * stage_bytes);
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + kMtMaxStages);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full0 + 8 * s, kMtProducers); mbar_init(empty0 + 8 * s, NC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();      // the only block-level barrier

  const unsigned long long in_elem0 = reinterpret_cast<unsigned long long>(p.in) >> 2;
  const int NWT = g.n_windows_total;

  // ======================= producer warps =======================
  if (warp >= NC) {
    if (lane != 0) return;
    const int pw = warp - NC;
    if (pw >= kMtProducers) return;
    const int n_mine = pw < MS ? (MS - pw + kMtProducers - 1) / kMtProducers : 0;
    const unsigned stage0 = smem_u32(stages);
    int s = 0;
    unsigned round = 0;
    RsPieces pieces(g.share, g.rows_total, p.ny, 0);
    pieces.g0 = (long long)blockIdx.x * g.share;
    pieces.g1 = min(g.rows_total, pieces.g0 + g.share);
    pieces.g = pieces.g0;
    long long task; int ya, yb;
    while (pieces.next(&task, &ya, &yb)) {
      const long long plane = task / g.n_panels;
      const int panel = (int)(task - plane * g.n_panels);
      const int y_lo = max(ya - R, 0), y_hi = min(yb - 1 + R, p.ny - 1);
      const int gw0 = panel * NW, gw1 = min(gw0 + NW, NWT) - 1;
      const int seg0 = max(gw0 * USEFUL - HL * K, 0);
      const int seg1 = min(p.nx, gw1 * USEFUL - HL * K + kMtWinCols);
      const float* plane_ptr = p.in + plane * p.plane_stride + seg0;
      for (int y = y_lo; y <= y_hi; ++y) {
        const float* row = plane_ptr + (long long)y * p.nx;
        const unsigned shift_b = (unsigned)(reinterpret_cast<unsigned long long>(row) & 15ull);
        const unsigned bytes = (shift_b + (unsigned)(seg1 - seg0) * 4u + 15u) & ~15u;
        const char* src = reinterpret_cast<const char*>(row) - shift_b;
        for (int m0 = 0; m0 < Rm; m0 += MS) {
          mbar_wait_backoff(empty0 + 8 * s, (round & 1u) ^ 1u, 200u);    // stages of slack: no polling against the math warps
          if (n_mine) mbar_expect_tx(full0 + 8 * s, bytes * (unsigned)n_mine); else mbar_arrive(full0 + 8 * s);
          unsigned dst = stage0 + (unsigned)s * stage_bytes + (unsigned)(pw * g.pitch) * 4u;
          const char* q = src + (long long)(m0 + pw) * p.member_stride * 4;
          for (int j = pw; j < MS; j += kMtProducers) {
            bulk_g2s(dst, q, bytes, full0 + 8 * s);
            dst += (unsigned)(kMtProducers * g.pitch) * 4u;
            q += (long long)kMtProducers * p.member_stride * 4;
          }
          if (++s == NS) { s = 0; ++round; }
        }
      }
    }
    return;
  }

  // ======================= consumer warps: (window, threshold) =======================
  const int tl = warp / NW, wl = warp - tl * NW;            // threshold of the pass, window of the panel
  const int t = g.t_base + tl;
  const ThrDev& thr = p.thr.t[t];
  unsigned* ring = reinterpret_cast<unsigned*>(bars + 2 * kMtMaxStages) + (size_t)warp * NB * kMtWinCols;
  const unsigned ring_u32 = smem_u32(ring) + (unsigned)lane * (K * 4u);
  const float fx_scale = (float)(1u << g.fx_shift);
  const unsigned stages_u32 = smem_u32(stages);
  const unsigned pitch_b = (unsigned)g.pitch * 4u;
  const bool check_nan = tl == 0;                           // the warps of one window see the same data
  float nan_lo = 0.f, nan_hi = 0.f;
  bool nan_seen = false;
  int s = 0;
  unsigned parity = 0, bar = full0, st_addr = stages_u32;

  RsPieces pieces(g.share, g.rows_total, p.ny, 0);
  pieces.g0 = (long long)blockIdx.x * g.share;
  pieces.g1 = min(g.rows_total, pieces.g0 + g.share);
  pieces.g = pieces.g0;
  long long task; int ya, yb;
  while (pieces.next(&task, &ya, &yb)) {
    const long long plane = task / g.n_panels;
    const int panel = (int)(task - plane * g.n_panels);
    const int gw = panel * NW + wl;
    const int gw0 = panel * NW;
    const int seg0 = max(gw0 * USEFUL - HL * K, 0);
    const int x0 = gw * USEFUL - HL * K + lane * K;
    const bool win_active = gw < NWT;
    unsigned okbits = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) okbits |= (win_active && x0 + i >= 0 && x0 + i < p.nx) ? (1u << i) : 0u;
    const bool lane_full = okbits == FULLBITS;
    const bool lane_whole = lane_full || okbits == 0;        // all four loaded values are data (a lane without columns reads the first group)
    const bool warp_edge = __any_sync(0xffffffffu, !lane_full);
    const int x_ld = okbits ? x0 : seg0;
    const bool out_lane = lane >= HL && lane < 32 - HL && okbits != 0;
    const bool lane_interior = lane_full && (x0 - R >= 0) && (x0 + K - 1 + R <= p.nx - 1);
    const int n_rows = (yb - ya) + 2 * R;
    const int y_first = ya - R;
    const long long obase0 = (plane * p.n_thr + t) * (long long)p.ny * p.nx;
    const unsigned long long plane_elem0 = in_elem0 + (unsigned long long)(plane * p.plane_stride) + (unsigned long long)seg0;
    const unsigned lane_col_b = (unsigned)(x_ld - seg0) * 4u;

#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const unsigned zero[4] = {0u, 0u, 0u, 0u};
      box_ring_st<K>(ring_u32 + (unsigned)q * (kMtWinCols * 4u), zero);
    }
    unsigned V[K];
#pragma unroll
    for (int i = 0; i < K; ++i) V[i] = 0u;
    int slot = 0;
    __syncwarp();

    for (int k = 0; k < n_rows; ++k) {
      const int yy = y_first + k;
      const bool in_dom = yy >= 0 && yy < p.ny;
      float acc[K];
#pragma unroll
      for (int i = 0; i < K; ++i) acc[i] = 0.f;
      if (in_dom) {
        const unsigned shift_b = (unsigned)((plane_elem0 + (unsigned long long)((long long)yy * p.nx)) & 3ull) * 4u;
        // the member loop, specialised on the NaN watch (one threshold's warps carry it for their window)
        auto members = [&](auto watch_tag) {
          constexpr bool WATCH = decltype(watch_tag)::value;
          for (int m0 = 0; m0 < Rm; m0 += MS) {
            mbar_wait(bar, parity);
            unsigned a = st_addr + lane_col_b + shift_b;
#pragma unroll 6
            for (int j = 0; j < MS; ++j) {
              float d[K];
              lds_f4(a, d);
              a += pitch_b;
              // the two column pairs apart: a partial lane (row end, nx is even: exactly its first pair
              // is valid) reads whatever follows the row in the stage in its second pair
              if (WATCH) { nan_lo = max3_nan(nan_lo, d[0], d[1]); nan_hi = max3_nan(nan_hi, d[2], d[3]); }
#pragma unroll
              for (int i = 0; i < K; ++i) acc[i] += truth_value_sum<TM>(d[i], thr);
            }
            __syncwarp();                                   // every lane has read the stage
            if (lane == 0) mbar_arrive(bar + 8 * kMtMaxStages);
            bar += 8; st_addr += stage_bytes;
            if (++s == NS) { s = 0; parity ^= 1u; bar = full0; st_addr = stages_u32; }
          }
        };
        if (check_nan) members(std::true_type{});
        else members(std::false_type{});
        if (M2D) {
          // the external mask is shared by all members: one multiply per column and row
#pragma unroll
          for (int i = 0; i < K; ++i)
            if ((okbits >> i) & 1u) acc[i] = p.mask2d[(long long)yy * p.nx + x0 + i] ? acc[i] : 0.f;
        }
      }
      unsigned q[K];
#pragma unroll
      for (int i = 0; i < K; ++i) q[i] = __float2uint_rn(acc[i] * fx_scale);
      if (warp_edge || !in_dom) {
#pragma unroll
        for (int i = 0; i < K; ++i) q[i] = (in_dom && ((okbits >> i) & 1u)) ? q[i] : 0u;
      }
      {
        const unsigned ra = ring_u32 + (unsigned)slot * (kMtWinCols * 4u);
        unsigned o[K];
        box_ring_ld<K>(ra, o);
        box_ring_st<K>(ra, q);
#pragma unroll
        for (int i = 0; i < K; ++i) V[i] += q[i] - o[i];
        slot = (slot + 1 == NB) ? 0 : slot + 1;
      }
      if (k < 2 * R) continue;
      // ---- horizontal windows from the neighbours' column sums; normalise; store ----
      unsigned E[K + 2 * R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int ol = R - i, dl = (ol + K - 1) / K;
        const int orr = i + 1, dr = (orr + K - 1) / K;
        E[i] = __shfl_up_sync(0xffffffffu, V[dl * K - ol], dl);
        E[K + R + i] = __shfl_down_sync(0xffffffffu, V[i - (dr - 1) * K], dr);
      }
#pragma unroll
      for (int i = 0; i < K; ++i) E[R + i] = V[i];
      unsigned part[3] = {0u, 0u, 0u};
#pragma unroll
      for (int c = 0; c <= 2 * R; ++c) part[c % 3] += E[c];
      unsigned H[K];
      H[0] = part[0] + part[1] + part[2];
#pragma unroll
      for (int i = 1; i < K; ++i) H[i] = H[i - 1] + E[i + 2 * R] - E[i - 1];
      if (out_lane) {
        const int yo = yy - R;
        const int rows_in = min(yo + R, p.ny - 1) - max(yo - R, 0) + 1;
        const long long pt0 = (long long)yo * p.nx + x0;
        float o[K];
        const float nrow = (float)rows_in * fx_scale * (float)Rm;      // the realization mean is part of the divisor
        const float nfi = nrow * (float)NB;
        float rci;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rci) : "f"(nfi));
#pragma unroll
        for (int i = 0; i < K; ++i) {
          float nf = nfi, rc = rci;
          if (M2D) {
            const float cnt = ((okbits >> i) & 1u) ? __ldg(p.cnt_plane + pt0 + i) : 1.f;
            nf = cnt * fx_scale * (float)Rm;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
          } else if (!lane_interior) {
            const int xo = x0 + i;
            nf = nrow * (float)(min(xo + R, p.nx - 1) - max(xo - R, 0) + 1);
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
          }
          const float Sf = (float)H[i];
          const float q0 = Sf * rc;
          const float rem = __fmaf_rn(-q0, nf, Sf);
          o[i] = (M2D && nf == 0.f) ? __int_as_float(0x7fc00000) : __fmaf_rn(rem, rc, q0);
        }
        if (lane_full) {
          float* op = p.out + obase0 + pt0;
          *reinterpret_cast<float2*>(op) = make_float2(o[0], o[1]);
          *reinterpret_cast<float2*>(op + 2) = make_float2(o[2], o[3]);
        } else {
#pragma unroll
          for (int i = 0; i < K; ++i)
            if ((okbits >> i) & 1u) p.out[obase0 + pt0 + i] = o[i];
        }
        if (p.out_mask) {
#pragma unroll
          for (int i = 0; i < K; ++i)
            if ((okbits >> i) & 1u) p.out_mask[obase0 + pt0 + i] = (M2D && p.mask2d[pt0 + i] == 0) ? 1 : 0;
        }
      }
    }
    nan_seen = nan_seen || __isnanf(nan_lo) || (lane_whole && __isnanf(nan_hi));
    nan_lo = nan_hi = 0.f;
  }
  if (__any_sync(0xffffffffu, nan_seen)) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
}

bool plan_square_mt(const NbhDesc& d, const float* in, SqMtGeom* g, int* grid, size_t* smem);
int launch_square_mt(const SqParams& p, const SqMtGeom& g, int tm, bool m2d, int grid, size_t smem, cudaStream_t st);
int launch_edge_flags(const SqParams& p, int tm, cudaStream_t st);

}  // namespace nbh
