// Square neighbourhood, warp-autonomous MULTI-THRESHOLD variant of the fused
// member path: every loaded value is thresholded against up to kMtThresholds
// thresholds while it sits in registers, so the input is read once per group
// of thresholds instead of once per threshold (algorithmic traffic
// 4*Y*X*L*(R + T)).  Same task decomposition and arithmetic as
// square_warp_kernel (members mode, two columns per lane); the per-threshold
// state is one ring, one pair of float64 running sums and one pair of float32
// member sums.
#pragma once
#include "nbh_square_warp.cuh"

namespace nbh {

constexpr int kMtThresholds = 4;

struct SqMtParams {
  int t_base;       // global index of the first threshold of this launch
  int n_launch;     // thresholds handled by this launch (<= kMtThresholds)
};

template <int TM, bool M2D, bool FULL, int MB>
__global__ void __launch_bounds__(32, 12)
square_warp_mt_kernel(const __grid_constant__ SqParams p, const SqWarpGeom g, const SqMtParams mt) {
  constexpr int VEC = 2, TB = kMtThresholds;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int WS = g.WS, nb = p.nb, r = p.r;

  // layout: rtab | hrow | flags[TB][4] | ring[TB][nb][WS]
  double* rtab = reinterpret_cast<double*>(smem_raw);
  double* hrow = rtab + ((nb + 2) & ~1);
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(hrow + (WS + 2));
  float* ring = reinterpret_cast<float*>(flags + TB * 4);

  for (int k = lane + 1; k <= nb; k += 32) rtab[k] = 1.0 / (double)k;
  if (lane < TB * 4) flags[lane] = 0ull;

  long long task = blockIdx.x;
  const int strip = (int)(task % g.n_strips); task /= g.n_strips;
  const int yseg = (int)(task % g.n_ysegs); task /= g.n_ysegs;
  const long long plane = task;
  const int nt = mt.n_launch;

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
  const int c0 = lane * VEC;
  const int gxr = load0 + c0;
  const bool col_active = (gxr >= 0) && (gxr < p.nx);
  const int gx = col_active ? gxr : 0;
  const bool in_left = !sum_only && col_active && (gxr + VEC - 1 <= r);
  const bool in_right = !sum_only && col_active && (gxr >= p.nx - 1 - r) && (gxr + VEC - 1 < p.nx);
  const bool strip_band = __any_sync(0xffffffffu, in_left || in_right);

  for (int s = 0; s < TB * nb; ++s) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) ring[s * WS + c0 + v] = 0.f;
  }
  if (lane == 0) hrow[0] = 0.0;
  __syncwarp();

  double V[TB][VEC];
  float P[TB][VEC];
#pragma unroll
  for (int tt = 0; tt < TB; ++tt)
#pragma unroll
    for (int v = 0; v < VEC; ++v) { V[tt][v] = 0.0; P[tt][v] = 0.f; }
  float nanacc = 0.f;
  unsigned cur_m2 = 0x0101u;
  int slot = 0;

  const int R = p.n_members;
  const int n_chunks = (R + MB - 1) / MB;
  const long long mstride = p.member_stride;
  const float* in_plane = p.in + plane * p.plane_stride;
  const unsigned goff = (unsigned)gx * 4u;
  const long long per_plane = (long long)p.ny * p.nx;

  const bool is_out = (c0 >= c_out0) && (c0 < c_out0 + wk) && col_active;
  int idx_hi[VEC], idx_lo[VEC];
  double rcol[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    idx_hi[v] = c0 + v + r + 1;
    idx_lo[v] = max(c0 + v - r, 0);
    const int xo = gx + v;
    const int cols_in = min(xo + r, p.nx - 1) - max(xo - r, 0) + 1;
    rcol[v] = sum_only ? 1.0 : rtab[min(max(cols_in, 1), nb)] * p.inv_members;
  }

  int ik = 0, ichunk = 0, ck = 0, cchunk = 0;
  auto row_ptr = [&](int k) {
    const int yy = min(max(y_first + dir * min(k, n_rows - 1), 0), p.ny - 1);
    return (long long)yy * p.nx;
  };
  auto issue = [&](float (&buf)[MB][VEC], unsigned& mk) {
    const long long off = row_ptr(ik);
    const int m0 = ichunk * MB;
    const float* q = in_plane + off + (long long)m0 * mstride;
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      if (FULL || m0 + j < R) VecLoad<VEC>::ld(byte_offset(q, goff), buf[j]);
      q += mstride;
    }
    if (M2D && ichunk == 0) mk = ldm_packed<VEC>(p.mask2d + off + gx);
    if (++ichunk == n_chunks) { ichunk = 0; ++ik; }
  };

  // one finished input row: per threshold ring update, scan, normalise, store
  auto finish_row = [&](int k, bool row_in_dom) {
    const int yo = y_first + dir * (k - r);
    const bool emit = k >= 2 * r;
    int rows_in = 1;
    if (emit) rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
    const double rrow = sum_only ? 1.0 : rtab[rows_in];
    const long long pt0 = (long long)yo * p.nx + gx;
#pragma unroll
    for (int tt = 0; tt < TB; ++tt) {
      if (tt >= nt) break;
      float* rg = ring + (size_t)tt * nb * WS;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (!row_in_dom || !col_active) P[tt][v] = 0.f;
        const int ri = slot * WS + c0 + v;
        const float old = rg[ri];
        rg[ri] = P[tt][v];
        V[tt][v] += (double)P[tt][v] - (double)old;
        P[tt][v] = 0.f;
      }
      if (!emit) continue;
      const double s1 = V[tt][0] + V[tt][1];
      double incl = s1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const double excl = incl - s1;
      __syncwarp();                                   // previous readers of hrow are done
      hrow[c0 + 1] = excl + V[tt][0];
      hrow[c0 + 2] = excl + s1;
      __syncwarp();
      if (is_out) {
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const double S = hrow[idx_hi[v]] - hrow[idx_lo[v]];
          if (M2D && !sum_only) o[v] = (float)(S * p.rcp_count[pt0 + v] * p.inv_members);
          else o[v] = (float)(S * rrow * rcol[v]);
        }
        const long long ob = (plane * p.n_thr + (mt.t_base + tt)) * per_plane;
        VecLoad<VEC>::st(p.out + ob + pt0, o);
        if (p.out_mask) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) p.out_mask[ob + pt0 + v] = (M2D && p.mask2d[pt0 + v] == 0) ? 1 : 0;
        }
      }
    }
    slot = (slot + 1 == nb) ? 0 : slot + 1;
  };

  auto consume = [&](float (&buf)[MB][VEC], unsigned& mk) {
    if (ck >= n_rows) return;
    const int yy = y_first + dir * ck;
    const bool in_dom = (yy >= 0) && (yy < p.ny);
    const int m0 = cchunk * MB;
    if (M2D && cchunk == 0) cur_m2 = mk;
    const bool v0 = !M2D || (cur_m2 & 0xffu), v1 = !M2D || (cur_m2 & 0xff00u);
    const bool flagged = in_dom && !sum_only && (strip_band || yy <= r || yy >= p.ny - 1 - r);
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      if (FULL || m0 + j < R) nanacc = max_nan(max_nan(nanacc, fabsf(buf[j][0])), fabsf(buf[j][1]));
    }
#pragma unroll
    for (int tt = 0; tt < TB; ++tt) {
      if (tt >= nt) break;
      const ThrDev& th = p.thr.t[tt];
      float a0 = 0.f, a1 = 0.f;
      unsigned nel = 0;
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        if (FULL || m0 + j < R) {
          float t0 = truth_value_sum<TM>(buf[j][0], th), t1 = truth_value_sum<TM>(buf[j][1], th);
          if (M2D) { t0 = v0 ? t0 : 0.f; t1 = v1 ? t1 : 0.f; }
          a0 += t0; a1 += t1;
          if (flagged) {
            if (!v0 || !v1 || truth_not_one(buf[j][0], th) || truth_not_one(buf[j][1], th)) nel |= 1u << j;
          }
        }
      }
      P[tt][0] += a0; P[tt][1] += a1;
      if (flagged && nel && col_active) {
        const unsigned long long ne = (unsigned long long)nel << m0;
        if (in_left) atomicOr(flags + tt * 4 + 2, ne);
        if (in_right) atomicOr(flags + tt * 4 + 3, ne);
        if (yy <= r) atomicOr(flags + tt * 4 + 0, ne);
        if (yy >= p.ny - 1 - r) atomicOr(flags + tt * 4 + 1, ne);
      }
    }
    if (++cchunk == n_chunks) {
      cchunk = 0;
      finish_row(ck, in_dom);
      ++ck;
    }
  };

  float bufA[MB][VEC], bufB[MB][VEC];
  unsigned mkA = 0x0101u, mkB = 0x0101u;
  issue(bufA, mkA);
  while (ck < n_rows) {
    issue(bufB, mkB);
    consume(bufA, mkA);
    issue(bufA, mkA);
    consume(bufB, mkB);
  }

  if (__any_sync(0xffffffffu, __isnanf(nanacc))) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  }
  __syncwarp();
  if (!sum_only && lane < nt) {
    const int tt = lane;
    for (int m = 0; m < R; ++m) {
      unsigned bits = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) bits |= (unsigned)((flags[tt * 4 + k] >> m) & 1ull) << k;
      if (bits) atomicOr(p.edge_flags + ((plane * R + m) * p.n_thr + (mt.t_base + tt)), bits);
    }
  }
}

inline size_t sq_mt_smem_bytes(int WS, int nb) {
  return (size_t)((nb + 2) & ~1) * 8 + (size_t)(WS + 2) * 8 + (size_t)kMtThresholds * 4 * 8 +
         (size_t)kMtThresholds * nb * WS * 4;
}

int launch_square_mt(const SqParams& p, const SqWarpGeom& g, const SqMtParams& mt, int tm, bool m2d,
                     cudaStream_t st);

}  // namespace nbh
