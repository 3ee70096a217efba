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

template <int G, int TM, bool M2D>
__global__ void __launch_bounds__(32, 8)
square_tma_kernel(const __grid_constant__ SqParams p, const SqWarpGeom g, const SqTmaGeom tg,
                  const CUtensorMap* __restrict__ tmap) {   // descriptor lives in global memory (workspace):
                                                            // a __grid_constant__ descriptor faults on this stack
  constexpr int VEC = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int WS = g.WS, nb = p.nb, r = p.r, R = p.n_members, S = tg.n_stages;

  // layout: stages (128-byte aligned) | mbarriers | hrow | rtab | ring
  float* stages = reinterpret_cast<float*>(smem_raw);
  const size_t stage_bytes = (size_t)tg.stage_floats * 4;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)S * stage_bytes);
  double* hrow = reinterpret_cast<double*>(bars + 8);
  double* rtab = hrow + (WS + 2);
  float* ring = reinterpret_cast<float*>(rtab + ((nb + 2) & ~1));

  for (int k = lane + 1; k <= nb; k += 32) rtab[k] = 1.0 / (double)k;

  long long task = blockIdx.x;
  const int strip = (int)(task % g.n_strips); task /= g.n_strips;
  const int yseg = (int)(task % g.n_ysegs); task /= g.n_ysegs;
  const int t = (int)(task % p.n_thr); task /= p.n_thr;
  const long long plane = task;
  const ThrDev thr = p.thr.t[TM ? t : 0];

  const bool wrap = false;                       // (periodic x is served by the LDG kernels)
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
  const int ya = yseg * g.seg_rows;
  const int yb = min(p.ny, ya + g.seg_rows);
  const int dir = (yseg & 1) ? -1 : 1;
  const int n_rows = (yb - ya) + 2 * r;
  const int y_first = dir > 0 ? ya - r : yb - 1 + r;

  int x0, load0, wk;
  if (strip == 0) { x0 = 0; load0 = 0; wk = g.W0; }
  else { x0 = g.W0 + (strip - 1) * g.W; load0 = x0 - g.hl; wk = g.W; }
  const int c_out0 = x0 - load0;

  int c0[G], gxr[G];
  bool col_active[G], colband[G];
  bool any_band = false;
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) {
    c0[g2] = g2 * 32 * VEC + lane * VEC;
    gxr[g2] = load0 + c0[g2];
    col_active[g2] = (gxr[g2] >= 0) && (gxr[g2] < p.nx);
    colband[g2] = !sum_only && col_active[g2] &&
                  ((gxr[g2] + VEC - 1 <= r) || (gxr[g2] >= p.nx - 1 - r && gxr[g2] + VEC - 1 < p.nx));
    any_band = any_band || colband[g2];
  }
  for (int s = 0; s < nb; ++s) {
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
      for (int v = 0; v < VEC; ++v) ring[s * WS + c0[g2] + v] = 0.f;
  }
  if (lane == 0) hrow[0] = 0.0;

  // ---- TMA pipeline set-up ----
  const unsigned bar0 = smem_u32(bars);
  const unsigned stage0 = smem_u32(stages);
  auto row_of = [&](int k) { return min(max(y_first + dir * k, 0), p.ny - 1); };   // clamped: always valid
  const unsigned box_bytes = (unsigned)(R * tg.box_cols * 4);
  auto box_start = [&](int y) {            // first column of the box of grid row y within its tensor row
    return (load0 + (tg.super_rows ? (y & 1) * p.nx : 0)) & ~3;
  };
  auto issue_row = [&](int k) {            // lane 0 only
    const int s = k % S;
    const int y = row_of(k);
    const int cy = tg.super_rows ? (y >> 1) : y;
    mbar_expect_tx(bar0 + 8 * s, box_bytes);
    tma_load_4d(stage0 + (unsigned)(s * stage_bytes), tmap, box_start(y), cy, (int)plane, 0, bar0 + 8 * s);
  };
  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int k = 0; k < S && k < n_rows; ++k) issue_row(k);
  }
  __syncwarp();

  double V[G][VEC];
  float P[G][VEC];
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
    for (int v = 0; v < VEC; ++v) { V[g2][v] = 0.0; P[g2][v] = 0.f; }
  unsigned long long ne_left = 0, ne_right = 0, ne_top = 0, ne_bot = 0;
  float nanacc = 0.f;
  int slot = 0;
  const bool strip_band = __any_sync(0xffffffffu, any_band);
  const long long obase0 = (plane * p.n_thr + t) * (long long)p.ny * p.nx;

  unsigned m2_next[G];
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) {
    m2_next[g2] = 0x0101u;
    if (M2D && col_active[g2]) m2_next[g2] = ldm_packed<VEC>(p.mask2d + (long long)row_of(0) * p.nx + gxr[g2]);
  }

  for (int k = 0; k < n_rows; ++k) {
    const int s = k % S;
    const int yy = y_first + dir * k;
    const bool in_dom = (yy >= 0) && (yy < p.ny);
    unsigned m2_cur[G];
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) {
      m2_cur[g2] = m2_next[g2];
      if (M2D && col_active[g2] && k + 1 < n_rows)
        m2_next[g2] = ldm_packed<VEC>(p.mask2d + (long long)row_of(k + 1) * p.nx + gxr[g2]);
    }
    mbar_wait(bar0 + 8 * s, (unsigned)((k / S) & 1));
    const int yrow = row_of(k);
    const int delta = load0 + (tg.super_rows ? (yrow & 1) * p.nx : 0) - box_start(yrow);   // 0 or 2
    const float* st = stages + (size_t)s * tg.stage_floats + delta;
    const int BW = tg.box_cols;
    const bool flagged = in_dom && !sum_only && (strip_band || yy <= r || yy >= p.ny - 1 - r);
    if (!flagged) {
#pragma unroll
      for (int g2 = 0; g2 < G; ++g2) {
        float acc0 = 0.f, acc1 = 0.f;
        const float* col = st + c0[g2];
#pragma unroll 6
        for (int m = 0; m < R; ++m) {
          const float2 d = *reinterpret_cast<const float2*>(col + (size_t)m * BW);
          nanacc = max_nan(nanacc, fabsf(d.x));
          nanacc = max_nan(nanacc, fabsf(d.y));
          float t0 = truth_value_sum<TM>(d.x, thr), t1 = truth_value_sum<TM>(d.y, thr);
          if (M2D) {
            t0 = (m2_cur[g2] & 0xffu) ? t0 : 0.f;
            t1 = (m2_cur[g2] & 0xff00u) ? t1 : 0.f;
          }
          acc0 += t0; acc1 += t1;
        }
        P[g2][0] = acc0; P[g2][1] = acc1;
      }
    } else {
#pragma unroll
      for (int g2 = 0; g2 < G; ++g2) {
        float acc0 = 0.f, acc1 = 0.f;
        unsigned long long ne = 0;
        const float* col = st + c0[g2];
        const bool v0 = !M2D || (m2_cur[g2] & 0xffu), v1 = !M2D || (m2_cur[g2] & 0xff00u);
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
        P[g2][0] = acc0; P[g2][1] = acc1;
        if (col_active[g2]) {
          if (colband[g2] && (gxr[g2] + VEC - 1 <= r)) ne_left |= ne;
          if (colband[g2] && (gxr[g2] >= p.nx - 1 - r)) ne_right |= ne;
          if (yy <= r) ne_top |= ne;
          if (yy >= p.ny - 1 - r) ne_bot |= ne;
        }
      }
    }
    __syncwarp();                                   // every lane has read stage s
    if (lane == 0 && k + S < n_rows) issue_row(k + S);

    // ---- vertical sums through the ring ----
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (!in_dom || !col_active[g2]) P[g2][v] = 0.f;
        const int ri = slot * WS + c0[g2] + v;
        const float old = ring[ri];
        ring[ri] = P[g2][v];
        V[g2][v] += (double)P[g2][v] - (double)old;
      }
    slot = (slot + 1 == nb) ? 0 : slot + 1;
    if (k < 2 * r) continue;

    // ---- horizontal window: float64 warp scan + prefix difference ----
    double carry = 0.0;
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) {
      const double s1 = V[g2][0] + V[g2][1];
      double incl = s1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const double excl = carry + incl - s1;
      hrow[c0[g2] + 1] = excl + V[g2][0];
      hrow[c0[g2] + 2] = excl + s1;
      if (g2 + 1 < G) carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    const int yo = y_first + dir * (k - r);
    const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
    const double rrow = rtab[rows_in] * p.inv_members;
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) {
      const bool is_out = (c0[g2] >= c_out0) && (c0[g2] < c_out0 + wk) && col_active[g2];
      if (is_out) {
        const long long pt0 = (long long)yo * p.nx + gxr[g2];
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int cc = c0[g2] + v;
          const double Sw = hrow[cc + r + 1] - hrow[max(cc - r, 0)];
          if (sum_only) {
            o[v] = (float)Sw;
          } else if (M2D) {
            o[v] = (float)(Sw * p.rcp_count[pt0 + v] * p.inv_members);
          } else {
            const int xo = gxr[g2] + v;
            const int cols_in = wrap ? nb : (min(xo + r, p.nx - 1) - max(xo - r, 0) + 1);
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

inline int sq_tma_stage_floats(int WS, int n_members) { return (n_members * (WS + 4) + 31) / 32 * 32; }

inline size_t sq_tma_smem_bytes(int WS, int nb, int n_members, int n_stages) {
  // Rounded up to a multiple of 1 KB: co-resident blocks are packed at the
  // requested size, and a TMA destination must be 128-byte aligned (an odd
  // size makes the second block on an SM fault with "illegal instruction").
  const size_t raw = (size_t)n_stages * sq_tma_stage_floats(WS, n_members) * 4 + 64 + (size_t)(WS + 2) * 8 +
                     (size_t)((nb + 2) & ~1) * 8 + (size_t)nb * WS * 4;
  return (raw + 1023) / 1024 * 1024;
}

int launch_square_tma(const SqParams& p, const SqWarpGeom& g, const SqTmaGeom& tg, const CUtensorMap* dev_map,
                      int tm, bool m2d, cudaStream_t st);

}  // namespace nbh
