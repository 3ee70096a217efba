// Square neighbourhood, TMA-fed warp-autonomous kernel (fused member path).
//
// Same task decomposition and arithmetic as square_warp_kernel, but the input
// rows arrive through the Tensor Memory Accelerator: one elected lane issues a
// single cp.async.bulk.tensor per row — a 4-D box (columns x 1 row x 1 plane x
// all members) — into a warp-private ring of shared-memory stages guarded by
// mbarriers.  Bytes in flight are therefore held in shared memory instead of
// registers (3-4 stages x 4.6 KB per warp), no per-lane address arithmetic or
// load instructions are issued, and occupancy is no longer register-bound.
//
// The UK grid's rows are 4168 B (8-byte aligned only) while TMA needs 16-byte
// global strides, so the tensor map describes "super-rows" of two grid rows
// (8336 B): row y is the box starting at column x + (y & 1) * nx of super-row
// y >> 1.  Grids with nx % 4 == 0 use plain rows.  Columns outside the domain
// are masked in the kernel (a box may spill into the neighbouring row).
// The first element of a box must itself be 16-byte aligned (a misaligned box
// start faults with "illegal instruction"), so every box is widened by 4
// columns and starts at the 4-column boundary at or below the wanted column;
// the per-row-parity offset (0 or 2 columns) is applied when reading the stage.
#pragma once
#include <cuda.h>

#include "nbh_square_warp.cuh"

namespace nbh {

struct SqTmaGeom {
  int n_stages;      // shared-memory stages per warp
  int super_rows;    // 1: tensor rows are pairs of grid rows
  int box_cols;      // WS + 4: a box must START on a 16-byte boundary, so it begins up to 3 columns early
  int stage_floats;  // n_members * box_cols rounded up to a multiple of 32 (128-byte stage alignment)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NBH_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NBH_DONE_%=;\n"
      "bra NBH_WAIT_%=;\n"
      "NBH_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int kTmaWarps = 4;     // warps per block sharing one TMA tile (x halo shared through smem)

// Block = kTmaWarps warps working on adjacent 32*2-column windows of ONE shared
// TMA tile per input row (all members).  The windows overlap by the halo inside
// shared memory, so the x halo costs no extra global/L2 traffic.  Pipeline:
// full[s] (TMA transaction barrier) / empty[s] (one arrival per warp); warp 0
// lane 0 is the producer.  Each warp keeps its own ring, vertical sums and
// horizontal scan exactly as in square_warp_kernel.
template <int TM, bool M2D>
__global__ void __launch_bounds__(kTmaWarps * 32, 3)
square_tma_kernel(const __grid_constant__ SqParams p, const SqWarpGeom g, const SqTmaGeom tg,
                  const CUtensorMap* __restrict__ tmap) {   // descriptor lives in global memory (workspace):
                                                            // a __grid_constant__ descriptor faults on this stack
  constexpr int VEC = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int WS = g.WS, nb = p.nb, r = p.r, R = p.n_members, S = tg.n_stages;

  // layout: stages (128-byte aligned) | full[8] empty[8] | rtab | per warp: hrow, ring
  float* stages = reinterpret_cast<float*>(smem_raw);
  const size_t stage_bytes = (size_t)tg.stage_floats * 4;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)S * stage_bytes);
  double* rtab = reinterpret_cast<double*>(bars + 16);
  unsigned char* wbase = reinterpret_cast<unsigned char*>(rtab + ((nb + 2) & ~1)) +
                         (size_t)warp * ((size_t)(WS + 2) * 8 + (size_t)nb * WS * 4);
  double* hrow = reinterpret_cast<double*>(wbase);
  float* ring = reinterpret_cast<float*>(wbase + (size_t)(WS + 2) * 8);

  for (int k = threadIdx.x + 1; k <= nb; k += blockDim.x) rtab[k] = 1.0 / (double)k;

  long long task = blockIdx.x;
  const int bstrip = (int)(task % g.n_strips); task /= g.n_strips;     // block strips
  const int yseg = (int)(task % g.n_ysegs); task /= g.n_ysegs;
  const int t = (int)(task % p.n_thr); task /= p.n_thr;
  const long long plane = task;
  const ThrDev thr = p.thr.t[TM ? t : 0];

  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
  const int ya = yseg * g.seg_rows;
  const int yb = min(p.ny, ya + g.seg_rows);
  const int dir = (yseg & 1) ? -1 : 1;
  const int n_rows = (yb - ya) + 2 * r;
  const int y_first = dir > 0 ? ya - r : yb - 1 + r;

  // tile = kTmaWarps windows of WS columns at a pitch of W output columns
  const int Wb = kTmaWarps * g.W;
  const int tile0 = bstrip == 0 ? 0 : (Wb + g.hl) + (bstrip - 1) * Wb - g.hl;   // first tile column (global)
  const int load0 = tile0 + warp * g.W;                                         // this warp's window
  const bool first_window = (bstrip == 0 && warp == 0);
  const int c_out0 = first_window ? 0 : g.hl;
  const int wk = first_window ? g.W0 : g.W;

  const int c0 = lane * VEC;
  const int gxr = load0 + c0;
  const bool col_active = (gxr >= 0) && (gxr < p.nx);
  const bool colband = !sum_only && col_active &&
                       ((gxr + VEC - 1 <= r) || (gxr >= p.nx - 1 - r && gxr + VEC - 1 < p.nx));
  for (int s = 0; s < nb; ++s) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) ring[s * WS + c0 + v] = 0.f;
  }
  if (lane == 0) hrow[0] = 0.0;

  // ---- TMA pipeline set-up ----
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + 8);
  const unsigned stage0 = smem_u32(stages);
  auto row_of = [&](int k) { return min(max(y_first + dir * k, 0), p.ny - 1); };   // clamped: always valid
  const unsigned box_bytes = (unsigned)(R * tg.box_cols * 4);
  auto box_start = [&](int y) {            // first column of the box of grid row y within its tensor row
    return (tile0 + (tg.super_rows ? (y & 1) * p.nx : 0)) & ~3;
  };
  auto issue_row = [&](int k) {            // warp 0, lane 0 only
    const int s = k % S;
    const int y = row_of(k);
    const int cy = tg.super_rows ? (y >> 1) : y;
    mbar_expect_tx(full0 + 8 * s, box_bytes);
    tma_load_4d(stage0 + (unsigned)(s * stage_bytes), tmap, box_start(y), cy, (int)plane, 0, full0 + 8 * s);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, kTmaWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int k = 0; k < S && k < n_rows; ++k) issue_row(k);
  }
  __syncthreads();

  double V[VEC];
  float P[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { V[v] = 0.0; P[v] = 0.f; }
  unsigned long long ne_left = 0, ne_right = 0, ne_top = 0, ne_bot = 0;
  float nanacc = 0.f;
  int slot = 0;
  const bool strip_band = __any_sync(0xffffffffu, colband);
  const long long obase0 = (plane * p.n_thr + t) * (long long)p.ny * p.nx;
  const int BW = tg.box_cols;

  unsigned m2_next = 0x0101u;
  if (M2D && col_active) m2_next = ldm_packed<VEC>(p.mask2d + (long long)row_of(0) * p.nx + gxr);

  for (int k = 0; k < n_rows; ++k) {
    const int s = k % S;
    // producer: refill the stage consumed in the previous iteration once every warp released it
    if (threadIdx.x == 0 && k >= 1 && k - 1 + S < n_rows) {
      const int sp = (k - 1) % S;
      mbar_wait(empty0 + 8 * sp, (unsigned)(((k - 1) / S) & 1));
      issue_row(k - 1 + S);
    }
    const int yy = y_first + dir * k;
    const bool in_dom = (yy >= 0) && (yy < p.ny);
    const unsigned m2_cur = m2_next;
    if (M2D && col_active && k + 1 < n_rows)
      m2_next = ldm_packed<VEC>(p.mask2d + (long long)row_of(k + 1) * p.nx + gxr);
    mbar_wait(full0 + 8 * s, (unsigned)((k / S) & 1));
    const int yrow = row_of(k);
    const int delta = tile0 + (tg.super_rows ? (yrow & 1) * p.nx : 0) - box_start(yrow);   // 0 or 2
    const float* col = stages + (size_t)s * tg.stage_floats + delta + warp * g.W + c0;
    const bool flagged = in_dom && !sum_only && (strip_band || yy <= r || yy >= p.ny - 1 - r);
    float acc0 = 0.f, acc1 = 0.f;
    if (!flagged) {
#pragma unroll 6
      for (int m = 0; m < R; ++m) {
        const float2 d = *reinterpret_cast<const float2*>(col + (size_t)m * BW);
        nanacc = max_nan(nanacc, fabsf(d.x));
        nanacc = max_nan(nanacc, fabsf(d.y));
        float t0 = truth_value_sum<TM>(d.x, thr), t1 = truth_value_sum<TM>(d.y, thr);
        if (M2D) {
          t0 = (m2_cur & 0xffu) ? t0 : 0.f;
          t1 = (m2_cur & 0xff00u) ? t1 : 0.f;
        }
        acc0 += t0; acc1 += t1;
      }
    } else {
      unsigned long long ne = 0;
      const bool v0 = !M2D || (m2_cur & 0xffu), v1 = !M2D || (m2_cur & 0xff00u);
#pragma unroll 2
      for (int m = 0; m < R; ++m) {
        const float2 d = *reinterpret_cast<const float2*>(col + (size_t)m * BW);
        nanacc = max_nan(nanacc, fabsf(d.x));
        nanacc = max_nan(nanacc, fabsf(d.y));
        float t0 = truth_value_sum<TM>(d.x, thr), t1 = truth_value_sum<TM>(d.y, thr);
        t0 = v0 ? t0 : 0.f;
        t1 = v1 ? t1 : 0.f;
        acc0 += t0; acc1 += t1;
        if (!v0 || !v1 || truth_not_one(d.x, thr) || truth_not_one(d.y, thr)) ne |= 1ull << m;
      }
      if (col_active) {
        if (colband && (gxr + VEC - 1 <= r)) ne_left |= ne;
        if (colband && (gxr >= p.nx - 1 - r)) ne_right |= ne;
        if (yy <= r) ne_top |= ne;
        if (yy >= p.ny - 1 - r) ne_bot |= ne;
      }
    }
    P[0] = acc0; P[1] = acc1;
    __syncwarp();                                   // every lane of this warp has read stage s
    if (lane == 0) mbar_arrive(empty0 + 8 * s);

    // ---- vertical sums through the warp-private ring ----
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (!in_dom || !col_active) P[v] = 0.f;
      const int ri = slot * WS + c0 + v;
      const float old = ring[ri];
      ring[ri] = P[v];
      V[v] += (double)P[v] - (double)old;
    }
    slot = (slot + 1 == nb) ? 0 : slot + 1;
    if (k < 2 * r) continue;

    // ---- horizontal window: float64 warp scan + prefix difference ----
    const double s1 = V[0] + V[1];
    double incl = s1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const double excl = incl - s1;
    hrow[c0 + 1] = excl + V[0];
    hrow[c0 + 2] = excl + s1;
    __syncwarp();
    const int yo = y_first + dir * (k - r);
    const bool is_out = (c0 >= c_out0) && (c0 < c_out0 + wk) && col_active;
    if (is_out) {
      const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
      const double rrow = rtab[rows_in] * p.inv_members;
      const long long pt0 = (long long)yo * p.nx + gxr;
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int cc = c0 + v;
        const double Sw = hrow[cc + r + 1] - hrow[max(cc - r, 0)];
        if (sum_only) {
          o[v] = (float)Sw;
        } else if (M2D) {
          o[v] = (float)(Sw * p.rcp_count[pt0 + v] * p.inv_members);
        } else {
          const int xo = gxr + v;
          const int cols_in = min(xo + r, p.nx - 1) - max(xo - r, 0) + 1;
          o[v] = (float)(Sw * rrow * rtab[cols_in]);
        }
      }
      VecLoad<VEC>::st(p.out + obase0 + pt0, o);
      if (p.out_mask) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          p.out_mask[obase0 + pt0 + v] = (M2D && p.mask2d[pt0 + v] == 0) ? 1 : 0;
      }
    }
    __syncwarp();                                   // hrow is rewritten by the next row
  }

  if (__any_sync(0xffffffffu, __isnanf(nanacc))) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
  if (!sum_only) {
    unsigned long long masks[4] = {ne_top, ne_bot, ne_left, ne_right};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)masks[k]);
      const unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(masks[k] >> 32));
      masks[k] = ((unsigned long long)hi32 << 32) | lo32;
    }
    if (lane == 0) {
      for (int m = 0; m < R; ++m) {
        unsigned bits = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) bits |= (unsigned)((masks[k] >> m) & 1ull) << k;
        if (bits) atomicOr(p.edge_flags + ((plane * R + m) * p.n_thr + t), bits);
      }
    }
  }
}

// tile columns: kTmaWarps windows at pitch W + the last window's remainder, + 4 for the aligned box start
inline int sq_tma_box_cols(int WS, int W) { return ((kTmaWarps - 1) * W + WS + 4 + 3) / 4 * 4; }
inline int sq_tma_stage_floats(int box_cols, int n_members) { return (n_members * box_cols + 31) / 32 * 32; }

inline size_t sq_tma_smem_bytes(int WS, int W, int nb, int n_members, int n_stages) {
  // Rounded up to a multiple of 1 KB (TMA destinations must be 128-byte aligned).
  const size_t raw = (size_t)n_stages * sq_tma_stage_floats(sq_tma_box_cols(WS, W), n_members) * 4 + 128 +
                     (size_t)((nb + 2) & ~1) * 8 +
                     (size_t)kTmaWarps * ((size_t)(WS + 2) * 8 + (size_t)nb * WS * 4);
  return (raw + 1023) / 1024 * 1024;
}

int launch_square_tma(const SqParams& p, const SqWarpGeom& g, const SqTmaGeom& tg, const CUtensorMap* dev_map,
                      int tm, bool m2d, cudaStream_t st);

}  // namespace nbh
