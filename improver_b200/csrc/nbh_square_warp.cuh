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

constexpr int kSqWarpsPerBlock = 1;   // one warp per block: everything derived from blockIdx is
                                      // provably uniform, so row/member pointers live in uniform registers

__device__ __forceinline__ const float* byte_offset(const float* base, unsigned bytes) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + bytes);
}

constexpr int kSqHrowRows = 8;        // prefix rows per warp (rows mode batches up to 8 horizontal passes)

struct SqWarpGeom {
  int WS;        // columns loaded per warp (32 * VEC * groups)
  int W0, W;     // output columns of strip 0 / of the other strips
  int hl;        // left halo of strips > 0 (multiple of VEC, >= r)
  int n_strips, n_ysegs, seg_rows;
  long long n_tasks;
  int n_sm;      // > 0: remap block -> task so that each SM works through a contiguous run of tasks
                 // (adjacent strips share their halo columns through that SM's L1)
};

template <int VEC, int G, int TM, bool M2D, bool ROWS, bool FULL, int MB>
__global__ void __launch_bounds__(kSqWarpsPerBlock * 32, 16)
square_warp_kernel(const __grid_constant__ SqParams p, const SqWarpGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int WS = g.WS, nb = p.nb, r = p.r;

  // block-shared reciprocal table, then warp-private ring + prefix row
  double* rtab = reinterpret_cast<double*>(smem_raw);                        // [nb + 2]
  const size_t rtab_elems = (size_t)((nb + 2) & ~1);
  const size_t per_warp = (size_t)(WS + 2) * 8 * kSqHrowRows + (size_t)nb * WS * 4;
  unsigned char* wbase = smem_raw + rtab_elems * 8 + (size_t)warp * per_warp;
  double* hrow = reinterpret_cast<double*>(wbase);                           // [rows][WS + 2]
  float* ring = reinterpret_cast<float*>(wbase + (size_t)(WS + 2) * 8 * kSqHrowRows);   // [nb][WS]

  for (int k = threadIdx.x + 1; k <= nb; k += blockDim.x) rtab[k] = 1.0 / (double)k;
  __syncthreads();     // the only block-level barrier in the kernel

  long long task = (long long)blockIdx.x * kSqWarpsPerBlock + warp;
  if (g.n_sm > 0) {
    const long long per = (g.n_tasks + g.n_sm - 1) / g.n_sm;
    task = (long long)(blockIdx.x % g.n_sm) * per + blockIdx.x / g.n_sm;
    if (blockIdx.x / g.n_sm >= per) return;
  }
  if (task >= g.n_tasks) return;
  const int strip = (int)(task % g.n_strips); task /= g.n_strips;
  const int yseg = (int)(task % g.n_ysegs); task /= g.n_ysegs;
  const int t = (int)(task % p.n_thr); task /= p.n_thr;
  const long long plane = task;
  if (p.redo && !p.redo[plane * p.n_thr + t]) return;
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
  // lane owns G groups of VEC adjacent columns: c = g * 32 * VEC + lane * VEC + v
  int c0[G], gxr[G], gx[G];
  bool col_active[G], colband[G];
  bool any_band = false;
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) {
    c0[g2] = g2 * 32 * VEC + lane * VEC;
    gxr[g2] = load0 + c0[g2];
    gx[g2] = gxr[g2];
    col_active[g2] = (gx[g2] >= 0) && (gx[g2] < p.nx);
    if (wrap) { gx[g2] %= p.nx; if (gx[g2] < 0) gx[g2] += p.nx; col_active[g2] = true; }
    // with a periodic x axis there is no left / right edge (flags preset by the host)
    colband[g2] = !sum_only && !wrap && col_active[g2] &&
                  ((gxr[g2] + VEC - 1 <= r) || (gxr[g2] >= p.nx - 1 - r && gxr[g2] + VEC - 1 < p.nx));
    any_band = any_band || colband[g2];
    if (!col_active[g2]) gx[g2] = 0;               // keep pointers in range; loads are predicated off
  }

  for (int s = 0; s < nb; ++s) {
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
      for (int v = 0; v < VEC; ++v) ring[s * WS + c0[g2] + v] = 0.f;
  }
  if (lane < kSqHrowRows) hrow[lane * (WS + 2)] = 0.0;
  __syncwarp();

  double V[G][VEC];
  float P[G][VEC];
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
    for (int v = 0; v < VEC; ++v) { V[g2][v] = 0.0; P[g2][v] = 0.f; }
  unsigned long long ne_left = 0, ne_right = 0, ne_top = 0, ne_bot = 0;
  float nanacc = 0.f;
  unsigned cur_m2[G];
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) cur_m2[g2] = 0x01010101u;
  int slot = 0;

  const int R = p.n_members;
  const int n_chunks = ROWS ? 1 : (R + MB - 1) / MB;
  const long long mstride = p.member_stride;
  const float* in_plane = p.in + plane * p.plane_stride;
  unsigned goff[G];
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) goff[g2] = (unsigned)gx[g2] * 4u;
  const long long obase0 = (plane * p.n_thr + t) * (long long)p.ny * p.nx;

  // issue / consume cursors: row counters k (0 .. n_rows) and member chunk index
  int ik = 0, ichunk = 0, ck = 0, cchunk = 0;

  // ---- one finished input row: ring, vertical sums, horizontal pass, store ----
  // vertical part: push the finished member-summed row through the ring
  auto vertical = [&](bool row_in_dom) {
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        // loads are unconditional (clamped addresses): rows outside the domain and
        // lanes outside the strip's columns contribute zero (zero padding)
        if (!row_in_dom || !col_active[g2]) P[g2][v] = 0.f;
        const int ri = slot * WS + c0[g2] + v;
        const float old = ring[ri];
        ring[ri] = P[g2][v];
        V[g2][v] += (double)P[g2][v] - (double)old;
        P[g2][v] = 0.f;
      }
    slot = (slot + 1 == nb) ? 0 : slot + 1;
  };
  // horizontal part 1: float64 inclusive prefix of the strip into a warp-private row
  // (lane-local sums, warp-shuffle scan per group, carry between groups)
  auto h_scan = [&](const double (&Vr)[G][VEC], double* hr) {
    double carry = 0.0;
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) {
      double loc[VEC];
      double s = 0.0;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { s += Vr[g2][v]; loc[v] = s; }
      double incl = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const double excl = carry + incl - s;
#pragma unroll
      for (int v = 0; v < VEC; ++v) hr[c0[g2] + v + 1] = excl + loc[v];
      if (g2 + 1 < G) carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  };
  // horizontal part 2: window sum = prefix difference, normalise, store
  // lane-constant pieces of the output stage, hoisted out of the row loop
  bool is_out[G];
  int idx_hi[G][VEC], idx_lo[G][VEC];
  double rcol[G][VEC];                                  // 1 / (window columns inside the domain) / members
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) {
    is_out[g2] = (c0[g2] >= c_out0) && (c0[g2] < c_out0 + wk) && (gxr[g2] < p.nx) && col_active[g2];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cc = c0[g2] + v;
      idx_hi[g2][v] = cc + r + 1;
      idx_lo[g2][v] = max(cc - r, 0);
      const int xo = gx[g2] + v;
      const int cols_in = wrap ? nb : (min(xo + r, p.nx - 1) - max(xo - r, 0) + 1);
      rcol[g2][v] = sum_only ? 1.0 : rtab[min(max(cols_in, 1), nb)] * p.inv_members;
    }
  }
  auto h_store = [&](int k, const double* hr) {
    const int yo = y_first + dir * (k - r);              // centre row of the window
    const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
    const double rrow = sum_only ? 1.0 : rtab[rows_in];
    const long long row_off = (long long)yo * p.nx;
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) {
      if (is_out[g2]) {
        const long long pt0 = row_off + gx[g2];
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const double S = hr[idx_hi[g2][v]] - hr[idx_lo[g2][v]];
          if (M2D && !sum_only) o[v] = (float)(S * p.rcp_count[pt0 + v] * p.inv_members);
          else o[v] = (float)(S * rrow * rcol[g2][v]);
        }
        VecLoad<VEC>::st(p.out + obase0 + pt0, o);
        if (p.out_mask) {
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            p.out_mask[obase0 + pt0 + v] = (M2D && p.mask2d[pt0 + v] == 0) ? 1 : 0;
        }
      }
    }
  };
  // one finished input row, members mode (one row per chunk)
  auto finish_row = [&](int k, bool row_in_dom) {
    vertical(row_in_dom);
    if (k < 2 * r) return;                               // window not complete yet
    __syncwarp();                                        // previous row's readers are done
    h_scan(V, hrow);
    __syncwarp();
    h_store(k, hrow);
  };

  // ---- edge-band bookkeeping for the trimming fix-up ----
  // A warp needs per-member "saw a value != 1" bits only while it is inside
  // one of the four edge bands: always for the two edge strips (warp-uniform),
  // and for rows within r+1 of the top / bottom edge.  Those chunks run the
  // flagged copy of the accumulate loop; everything else runs the lean one.
  const bool strip_band = __any_sync(0xffffffffu, any_band);
  unsigned long long ne_row[G];
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) ne_row[g2] = 0;
  auto row_flags_done = [&](int yy) {
    if (sum_only) return;
    const bool top = yy <= r, bot = yy >= p.ny - 1 - r;
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) {
      if (colband[g2] && (gxr[g2] + VEC - 1 <= r)) ne_left |= ne_row[g2];
      if (colband[g2] && (gxr[g2] >= p.nx - 1 - r)) ne_right |= ne_row[g2];
      if (top) ne_top |= ne_row[g2];
      if (bot) ne_bot |= ne_row[g2];
      ne_row[g2] = 0;
    }
  };

  constexpr int MKN = ROWS ? MB : 1;
  const uint8_t* m2_base = M2D ? p.mask2d : nullptr;
  auto row_ptr = [&](int k) {                 // clamped: always a valid row of this plane
    const int yy = min(max(y_first + dir * min(k, n_rows - 1), 0), p.ny - 1);
    return (long long)yy * p.nx;
  };
  auto issue = [&](float (&buf)[MB][G][VEC], unsigned (&mk)[MKN][G]) {
    if (ROWS) {
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const long long off = row_ptr(ik + j);
#pragma unroll
        for (int g2 = 0; g2 < G; ++g2) {
          VecLoad<VEC>::ld(byte_offset(in_plane + off, goff[g2]), buf[j][g2]);
          if (M2D) mk[j][g2] = ldm_packed<VEC>(m2_base + off + gx[g2]);
        }
      }
      ik += MB;
    } else {
      const long long off = row_ptr(ik);
      const int m0 = ichunk * MB;
      // warp-uniform member pointer + 32-bit per-lane byte offset
      const float* q = in_plane + off + (long long)m0 * mstride;
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        if (FULL || m0 + j < R) {
#pragma unroll
          for (int g2 = 0; g2 < G; ++g2) VecLoad<VEC>::ld(byte_offset(q, goff[g2]), buf[j][g2]);
        }
        q += mstride;
      }
      if (M2D && ichunk == 0) {
#pragma unroll
        for (int g2 = 0; g2 < G; ++g2) mk[0][g2] = ldm_packed<VEC>(m2_base + off + gx[g2]);
      }
      if (++ichunk == n_chunks) { ichunk = 0; ++ik; }
    }
  };

  // one loaded VEC group of one member: threshold, mask, accumulate; the flagged
  // copy also records "saw a value != 1" in bit `bit` of nel
  auto accumulate = [&](const float (&d)[VEC], unsigned m2, float (&acc)[VEC], bool flagged, unsigned bit,
                        unsigned& nel) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      nanacc = max_nan(nanacc, fabsf(d[v]));
      float tv = truth_value_sum<TM>(d[v], thr);
      bool valid = true;
      if (M2D) { valid = ((m2 >> (8 * v)) & 0xffu) != 0; tv = valid ? tv : 0.f; }
      acc[v] += tv;
      if (flagged) { if (!valid || truth_not_one(d[v], thr)) nel |= bit; }
    }
  };

  auto consume = [&](float (&buf)[MB][G][VEC], unsigned (&mk)[MKN][G]) {
    if (ROWS) {
      // The MB rows of a chunk are pushed through the vertical sums one after the
      // other, but their horizontal passes are independent: snapshot V per row and
      // run the MB scans together (instruction-level parallelism hides the shuffle
      // latency that dominates when there is only one member per row).
      double Vs[MB][G][VEC];
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const int k = ck + j;
        if (k < n_rows) {
          const int yy = y_first + dir * k;
          const bool in_dom = (yy >= 0) && (yy < p.ny);
          const bool flagged = in_dom && !sum_only && (strip_band || yy <= r || yy >= p.ny - 1 - r);
          if (flagged) {
#pragma unroll
            for (int g2 = 0; g2 < G; ++g2) {
              unsigned nel = 0;
              accumulate(buf[j][g2], M2D ? mk[j][g2] : 0u, P[g2], true, 1u, nel);
              ne_row[g2] |= nel;
            }
            row_flags_done(yy);
          } else {
            unsigned nel = 0;
#pragma unroll
            for (int g2 = 0; g2 < G; ++g2) accumulate(buf[j][g2], M2D ? mk[j][g2] : 0u, P[g2], false, 0u, nel);
          }
          vertical(in_dom);
        }
#pragma unroll
        for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
          for (int v = 0; v < VEC; ++v) Vs[j][g2][v] = V[g2][v];
      }
      __syncwarp();                                      // previous chunk's readers are done
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (ck + j < n_rows && ck + j >= 2 * r) h_scan(Vs[j], hrow + j * (WS + 2));
      __syncwarp();
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (ck + j < n_rows && ck + j >= 2 * r) h_store(ck + j, hrow + j * (WS + 2));
      ck += MB;
    } else {
      if (ck < n_rows) {
        const int yy = y_first + dir * ck;
        const bool in_dom = (yy >= 0) && (yy < p.ny);
        const int m0 = cchunk * MB;
        const bool flagged = in_dom && !sum_only && (strip_band || yy <= r || yy >= p.ny - 1 - r);
#pragma unroll
        for (int g2 = 0; g2 < G; ++g2) {
          if (M2D && cchunk == 0) cur_m2[g2] = mk[0][g2];
          unsigned nel = 0;
          if (flagged) {
#pragma unroll
            for (int j = 0; j < MB; ++j)
              if (FULL || m0 + j < R) accumulate(buf[j][g2], cur_m2[g2], P[g2], true, 1u << j, nel);
            ne_row[g2] |= (unsigned long long)nel << m0;
          } else {
#pragma unroll
            for (int j = 0; j < MB; ++j)
              if (FULL || m0 + j < R) accumulate(buf[j][g2], cur_m2[g2], P[g2], false, 0u, nel);
          }
        }
        if (++cchunk == n_chunks) {
          cchunk = 0;
          if (flagged) row_flags_done(yy);
          finish_row(ck, in_dom);
          ++ck;
        }
      }
    }
  };

  float bufA[MB][G][VEC], bufB[MB][G][VEC];
  unsigned mkA[MKN][G], mkB[MKN][G];
#pragma unroll
  for (int j = 0; j < MKN; ++j)
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) { mkA[j][g2] = 0x01010101u; mkB[j][g2] = 0x01010101u; }
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

inline size_t sq_warp_smem_bytes(int vec, int groups, int nb) {
  const int WS = 32 * vec * groups;
  return (size_t)((nb + 2) & ~1) * 8 +
         (size_t)kSqWarpsPerBlock * ((size_t)(WS + 2) * 8 * kSqHrowRows + (size_t)nb * WS * 4);
}

template <int VEC>
int launch_square_warp_vec(const SqParams& p, const SqWarpGeom& g, int tm, bool m2d, cudaStream_t st);

template <int VEC, int G, int TM, bool M2D, bool ROWS, bool FULL, int MB>
static int launch_square_warp_one(const SqParams& p, const SqWarpGeom& g, cudaStream_t st) {
  auto kern = square_warp_kernel<VEC, G, TM, M2D, ROWS, FULL, MB>;
  const size_t smem = sq_warp_smem_bytes(VEC, G, p.nb);
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (g.n_tasks + kSqWarpsPerBlock - 1) / kSqWarpsPerBlock;
  if (g.n_sm > 0) blocks = (g.n_tasks + g.n_sm - 1) / g.n_sm * g.n_sm;
  NBH_REQUIRE(blocks < (1LL << 31), "too many blocks (%lld); split the launch", blocks);
  kern<<<(unsigned)blocks, kSqWarpsPerBlock * 32, smem, st>>>(p, g);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// 36 floats per register buffer: MB * G * VEC == 36 (members mode), 16 (rows mode)
template <int VEC, int G, int TM, bool M2D>
static int launch_square_warp_mode(const SqParams& p, const SqWarpGeom& g, cudaStream_t st) {
  constexpr int MBR0 = 16 / (VEC * G) > 0 ? 16 / (VEC * G) : 1;
  constexpr int MBR = MBR0 > kSqHrowRows ? kSqHrowRows : MBR0;
  constexpr int MBM = G == 2 ? 12 / VEC : (36 / VEC > 18 ? 18 : 36 / VEC);   // G == 2: smaller chunks keep 128 registers
  if (p.n_members == 1) return launch_square_warp_one<VEC, G, TM, M2D, true, true, MBR>(p, g, st);
  if (p.n_members % MBM == 0) return launch_square_warp_one<VEC, G, TM, M2D, false, true, MBM>(p, g, st);
  return launch_square_warp_one<VEC, G, TM, M2D, false, false, MBM>(p, g, st);
}

template <int VEC, int TM, bool M2D>
static int launch_square_warp_groups(const SqParams& p, const SqWarpGeom& g, cudaStream_t st) {
  if (g.WS == 32 * VEC) return launch_square_warp_mode<VEC, 1, TM, M2D>(p, g, st);
  if constexpr (VEC < 4) return launch_square_warp_mode<VEC, 2, TM, M2D>(p, g, st);
  nbh::set_error("unsupported strip width %d", g.WS);
  return 1;
}

#define NBH_DEFINE_SQUARE_WARP_VEC(VEC)                                                              \
  template <>                                                                                        \
  int launch_square_warp_vec<VEC>(const SqParams& p, const SqWarpGeom& g, int tm, bool m2d,          \
                                  cudaStream_t st) {                                                 \
    switch ((tm << 1) | (m2d ? 1 : 0)) {                                                             \
      case 0: return launch_square_warp_groups<VEC, 0, false>(p, g, st);                               \
      case 1: return launch_square_warp_groups<VEC, 0, true>(p, g, st);                                \
      case 2: return launch_square_warp_groups<VEC, 1, false>(p, g, st);                               \
      case 3: return launch_square_warp_groups<VEC, 1, true>(p, g, st);                                \
      case 4: return launch_square_warp_groups<VEC, 2, false>(p, g, st);                               \
      case 5: return launch_square_warp_groups<VEC, 2, true>(p, g, st);                                \
      case 8: return launch_square_warp_groups<VEC, 4, false>(p, g, st);                               \
      case 9: return launch_square_warp_groups<VEC, 4, true>(p, g, st);                                \
      case 10: return launch_square_warp_groups<VEC, 5, false>(p, g, st);                              \
      case 11: return launch_square_warp_groups<VEC, 5, true>(p, g, st);                               \
      default: nbh::set_error("unsupported warp kernel variant"); return 1;                            \
    }                                                                                                \
  }

}  // namespace nbh
