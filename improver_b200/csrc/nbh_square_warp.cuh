// Square neighbourhood, warp-autonomous variant for small radii.
//
// Every warp owns one task = (output plane, threshold, column strip of
// 32*VEC columns, y segment) and marches over the rows of its segment with no
// block-level synchronisation at all: loads of the next register chunk are in
// flight while the current chunk is thresholded and summed, the vertical
// running sums live in registers backed by a warp-private shared-memory ring,
// and the horizontal window sum is a float64 warp-shuffle scan followed by a
// prefix difference through a warp-private row buffer.  Because warps never
// wait for each other, a stalled load only delays its own warp and the SM
// keeps ~12 warps x 18 x 256 B of HBM requests outstanding.
//
// Odd y segments march upwards, even ones downwards, so that the 2r halo rows
// shared by neighbouring segments are read at (nearly) the same time by both
// and the second read is an L2 hit instead of HBM traffic.
#pragma once
#include "nbh_square.cuh"

namespace nbh {

constexpr int kSqWarpsPerBlock = 4;

struct SqWarpGeom {
  int WS;        // columns loaded per warp (32 * VEC)
  int W0, W;     // output columns of strip 0 / of the other strips
  int hl;        // left halo of strips > 0 (multiple of VEC, >= r)
  int n_strips, n_ysegs, seg_rows;
  long long n_tasks;
};

template <int VEC, int TM, bool M2D, bool ROWS, bool FULL, int MB>
__global__ void __launch_bounds__(kSqWarpsPerBlock * 32, 3)
square_warp_kernel(const __grid_constant__ SqParams p, const SqWarpGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int WS = g.WS, nb = p.nb, r = p.r;

  // block-shared reciprocal table, then warp-private ring + prefix row
  double* rtab = reinterpret_cast<double*>(smem_raw);                        // [nb + 2]
  const size_t rtab_elems = (size_t)((nb + 2) & ~1);
  const size_t per_warp = (size_t)(WS + 2) * 8 + (size_t)nb * WS * 4;
  unsigned char* wbase = smem_raw + rtab_elems * 8 + (size_t)warp * per_warp;
  double* hrow = reinterpret_cast<double*>(wbase);                           // [WS + 1]
  float* ring = reinterpret_cast<float*>(wbase + (size_t)(WS + 2) * 8);       // [nb][WS]

  for (int k = threadIdx.x + 1; k <= nb; k += blockDim.x) rtab[k] = 1.0 / (double)k;
  __syncthreads();     // the only block-level barrier in the kernel

  long long task = (long long)blockIdx.x * kSqWarpsPerBlock + warp;
  if (task >= g.n_tasks) return;
  const int strip = (int)(task % g.n_strips); task /= g.n_strips;
  const int yseg = (int)(task % g.n_ysegs); task /= g.n_ysegs;
  const int t = (int)(task % p.n_thr); task /= p.n_thr;
  const long long plane = task;
  const ThrDev thr = p.thr.t[TM ? t : 0];

  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
  const int ya = yseg * g.seg_rows;
  const int yb = min(p.ny, ya + g.seg_rows);
  const int dir = (yseg & 1) ? -1 : 1;
  const int n_rows = (yb - ya) + 2 * r;            // rows visited, halo included
  const int y_first = dir > 0 ? ya - r : yb - 1 + r;

  // strip geometry: strip 0 starts at the domain edge and needs no left halo
  int x0, load0, wk;
  if (strip == 0 && !wrap) { x0 = 0; load0 = 0; wk = g.W0; }
  else {
    x0 = wrap ? strip * g.W : g.W0 + (strip - 1) * g.W;
    load0 = x0 - g.hl; wk = g.W;
  }
  const int c_out0 = x0 - load0;                   // strip-local first output column
  const int c0 = lane * VEC;
  const int gx_raw = load0 + c0;
  int gx = gx_raw;
  bool col_active = (gx >= 0) && (gx < p.nx);
  if (wrap) { gx %= p.nx; if (gx < 0) gx += p.nx; col_active = true; }
  // with a periodic x axis there is no left / right edge (flags preset by the host)
  const bool colband = !sum_only && !wrap && col_active &&
                       ((gx_raw + VEC - 1 <= r) || (gx_raw >= p.nx - 1 - r && gx_raw + VEC - 1 < p.nx));

  for (int s = 0; s < nb; ++s) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) ring[s * WS + c0 + v] = 0.f;
  }
  if (lane == 0) hrow[0] = 0.0;
  __syncwarp();

  double V[VEC];
  float P[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { V[v] = 0.0; P[v] = 0.f; }
  unsigned long long ne_col = 0, ne_top = 0, ne_bot = 0;
  float nanacc = 0.f;
  unsigned cur_m2 = 0x01010101u;
  int slot = 0;

  const int R = p.n_members;
  const int n_chunks = ROWS ? 1 : (R + MB - 1) / MB;
  const long long mstride = p.member_stride;
  const long long row_step = (long long)dir * p.nx;
  const float* in_col = p.in + plane * p.plane_stride + gx;
  const uint8_t* m2_col = M2D ? p.mask2d + gx : nullptr;
  const long long obase0 = (plane * p.n_thr + t) * (long long)p.ny * p.nx;

  // issue / consume cursors: row counters k (0 .. n_rows) and member chunk index
  int ik = 0, ichunk = 0, ck = 0, cchunk = 0;
  const float* irow = in_col + (long long)y_first * p.nx;
  const uint8_t* irow_m2 = M2D ? m2_col + (long long)y_first * p.nx : nullptr;

  // ---- one finished input row: ring, vertical sums, horizontal pass, store ----
  auto finish_row = [&](int k) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int ri = slot * WS + c0 + v;
      const float old = ring[ri];
      ring[ri] = P[v];
      V[v] += (double)P[v] - (double)old;
      P[v] = 0.f;
    }
    slot = (slot + 1 == nb) ? 0 : slot + 1;
    if (k < 2 * r) return;                               // window not complete yet
    const int yo = y_first + dir * (k - r);              // centre row of the window
    // float64 inclusive prefix over the strip: lane-local then warp scan
    double loc[VEC];
    double s = 0.0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) { s += V[v]; loc[v] = s; }
    double incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const double excl = incl - s;
    __syncwarp();                                        // previous row's readers are done
#pragma unroll
    for (int v = 0; v < VEC; ++v) hrow[c0 + v + 1] = excl + loc[v];
    __syncwarp();
    const bool is_out = (c0 >= c_out0) && (c0 < c_out0 + wk) && (gx_raw < p.nx) && col_active;
    if (is_out) {
      const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
      const double rrow = rtab[rows_in] * p.inv_members;
      const long long pt0 = (long long)yo * p.nx + gx;
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int cc = c0 + v;
        const double S = hrow[cc + r + 1] - hrow[max(cc - r, 0)];
        if (sum_only) {
          o[v] = (float)S;
        } else if (M2D) {
          o[v] = (float)(S * p.rcp_count[pt0 + v] * p.inv_members);
        } else {
          const int xo = gx + v;
          const int cols_in = wrap ? nb : (min(xo + r, p.nx - 1) - max(xo - r, 0) + 1);
          o[v] = (float)(S * rrow * rtab[cols_in]);
        }
      }
      VecLoad<VEC>::st(p.out + obase0 + pt0, o);
      if (p.out_mask) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          p.out_mask[obase0 + pt0 + v] = (M2D && p.mask2d[pt0 + v] == 0) ? 1 : 0;
      }
    }
  };

  // ---- edge-band bookkeeping for the trimming fix-up (rare rows / lanes) ----
  auto flag_pass = [&](int yy) {
    if (sum_only || !col_active || yy < 0 || yy >= p.ny) return;
    const bool top = yy <= r, bot = yy >= p.ny - 1 - r;
    if (!(colband || top || bot)) return;
    const long long off = (long long)yy * p.nx;
    unsigned m2 = 0x01010101u;
    if (M2D) m2 = ldm_packed<VEC>(m2_col + off);
    unsigned long long ne = 0;
    const float* q = in_col + off;
#pragma unroll 1
    for (int m = 0; m < R; ++m, q += mstride) {
      float d[VEC];
      VecLoad<VEC>::ld(q, d);
      bool any = false;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const bool valid = ((m2 >> (8 * v)) & 0xffu) != 0;
        const float tv = valid ? truth_value_t<TM>(d[v], thr) : 0.f;
        any = any || (tv != 1.0f);
      }
      if (any) ne |= 1ull << m;
    }
    ne_col |= ne;
    if (top) ne_top |= ne;
    if (bot) ne_bot |= ne;
  };

  auto issue = [&](float (&buf)[MB][VEC], unsigned (&mk)[ROWS ? MB : 1]) {
    if (ROWS) {
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const int yy = y_first + dir * (ik + j);
        const bool act = col_active && (ik + j < n_rows) && (yy >= 0) && (yy < p.ny);
        if (act) {
          VecLoad<VEC>::ld(irow, buf[j]);
          if (M2D) mk[j] = ldm_packed<VEC>(irow_m2);
        }
        irow += row_step;
        if (M2D) irow_m2 += row_step;
      }
      ik += MB;
    } else {
      const int yy = y_first + dir * ik;
      const bool act = col_active && (ik < n_rows) && (yy >= 0) && (yy < p.ny);
      const int m0 = ichunk * MB;
      const float* q = irow + (long long)m0 * mstride;
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        if (act && (FULL || m0 + j < R)) VecLoad<VEC>::ld(q, buf[j]);
        q += mstride;
      }
      if (M2D && act && ichunk == 0) mk[0] = ldm_packed<VEC>(irow_m2);
      if (++ichunk == n_chunks) {
        ichunk = 0; ++ik;
        irow += row_step;
        if (M2D) irow_m2 += row_step;
      }
    }
  };

  auto accumulate = [&](const float (&d)[VEC], unsigned m2) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      nanacc = max_nan(nanacc, fabsf(d[v]));
      float tv = truth_value_t<TM>(d[v], thr);
      if (M2D) tv = (((m2 >> (8 * v)) & 0xffu) != 0) ? tv : 0.f;
      P[v] += tv;
    }
  };

  auto consume = [&](float (&buf)[MB][VEC], unsigned (&mk)[ROWS ? MB : 1]) {
    if (ROWS) {
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const int k = ck + j;
        if (k < n_rows) {
          const int yy = y_first + dir * k;
          if (col_active && (yy >= 0) && (yy < p.ny)) accumulate(buf[j], M2D ? mk[j] : 0u);
          flag_pass(yy);
          finish_row(k);
        }
      }
      ck += MB;
    } else {
      if (ck < n_rows) {
        const int yy = y_first + dir * ck;
        if (col_active && (yy >= 0) && (yy < p.ny)) {
          if (M2D && cchunk == 0) cur_m2 = mk[0];
          const int m0 = cchunk * MB;
#pragma unroll
          for (int j = 0; j < MB; ++j)
            if (FULL || m0 + j < R) accumulate(buf[j], cur_m2);
        }
        if (++cchunk == n_chunks) {
          cchunk = 0;
          flag_pass(yy);
          finish_row(ck);
          ++ck;
        }
      }
    }
  };

  float bufA[MB][VEC], bufB[MB][VEC];
  unsigned mkA[ROWS ? MB : 1], mkB[ROWS ? MB : 1];
#pragma unroll
  for (int j = 0; j < (ROWS ? MB : 1); ++j) { mkA[j] = 0x01010101u; mkB[j] = 0x01010101u; }
  issue(bufA, mkA);
  while (ck < n_rows) {
    issue(bufB, mkB);
    consume(bufA, mkA);
    issue(bufA, mkA);
    consume(bufB, mkB);
  }

  // ---- epilogue: NaN flag and edge-band flags ----
  if (__any_sync(0xffffffffu, __isnanf(nanacc))) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
  if (!sum_only) {
    const bool left = colband && (gx_raw + VEC - 1 <= r);
    const bool right = colband && (gx_raw >= p.nx - 1 - r);
    unsigned long long masks[4] = {ne_top, ne_bot, left ? ne_col : 0ull, right ? ne_col : 0ull};
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

inline size_t sq_warp_smem_bytes(int vec, int nb) {
  const int WS = 32 * vec;
  return (size_t)((nb + 2) & ~1) * 8 + (size_t)kSqWarpsPerBlock * ((size_t)(WS + 2) * 8 + (size_t)nb * WS * 4);
}

template <int VEC>
int launch_square_warp_vec(const SqParams& p, const SqWarpGeom& g, int tm, bool m2d, cudaStream_t st);

template <int VEC, int TM, bool M2D, bool ROWS, bool FULL, int MB>
static int launch_square_warp_one(const SqParams& p, const SqWarpGeom& g, cudaStream_t st) {
  auto kern = square_warp_kernel<VEC, TM, M2D, ROWS, FULL, MB>;
  const size_t smem = sq_warp_smem_bytes(VEC, p.nb);
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (g.n_tasks + kSqWarpsPerBlock - 1) / kSqWarpsPerBlock;
  NBH_REQUIRE(blocks < (1LL << 31), "too many blocks (%lld); split the launch", blocks);
  kern<<<(unsigned)blocks, kSqWarpsPerBlock * 32, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int VEC, int TM, bool M2D>
static int launch_square_warp_mode(const SqParams& p, const SqWarpGeom& g, cudaStream_t st) {
  constexpr int MBR = VEC == 4 ? 4 : 8;
  constexpr int MBM = VEC == 4 ? 9 : 18;
  if (p.n_members == 1) return launch_square_warp_one<VEC, TM, M2D, true, true, MBR>(p, g, st);
  if (p.n_members % MBM == 0) return launch_square_warp_one<VEC, TM, M2D, false, true, MBM>(p, g, st);
  return launch_square_warp_one<VEC, TM, M2D, false, false, MBM>(p, g, st);
}

#define NBH_DEFINE_SQUARE_WARP_VEC(VEC)                                                              \
  template <>                                                                                        \
  int launch_square_warp_vec<VEC>(const SqParams& p, const SqWarpGeom& g, int tm, bool m2d,          \
                                  cudaStream_t st) {                                                 \
    switch ((tm << 1) | (m2d ? 1 : 0)) {                                                             \
      case 0: return launch_square_warp_mode<VEC, 0, false>(p, g, st);                               \
      case 1: return launch_square_warp_mode<VEC, 0, true>(p, g, st);                                \
      case 2: return launch_square_warp_mode<VEC, 1, false>(p, g, st);                               \
      case 3: return launch_square_warp_mode<VEC, 1, true>(p, g, st);                                \
      case 4: return launch_square_warp_mode<VEC, 2, false>(p, g, st);                               \
      default: return launch_square_warp_mode<VEC, 2, true>(p, g, st);                               \
    }                                                                                                \
  }

}  // namespace nbh
