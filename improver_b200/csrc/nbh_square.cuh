// Square neighbourhood main kernel (template) — see nbh_square.cu for the
// algorithm description, the host planner and the trimming fix-up.
#pragma once
#include "nbh_common.cuh"

namespace nbh {

// exact per-slice statistics (bounding boxes of v != 0 / v != 1, nanmin / nanmax keys): built
// lazily for suspicious slices, or handed in by the caller for y-band tiled launches
using SliceStats = NbhSliceStats;

struct SqParams {
  const float* in;
  float* out;
  uint8_t* out_mask;
  const uint8_t* mask2d;
  const uint8_t* mask_nd;
  const double* rcp_count;   // (ny,nx) 1/count for mask2d launches
  const float* cnt_plane;    // (ny,nx) the count itself (float32), same launches
  int32_t* status;
  const int* redo;           // not NULL: only the units (plane * T + t) flagged here are computed (guarded
                             // float64 re-run of the slices the fixed-point box kernel could not take)
  unsigned* edge_flags;      // [(plane*R + m) * T + t]
  SliceStats* stats;
  long long plane_stride, member_stride;
  long long n_planes;
  int n_members, ny, nx, r, nb;
  int flags;
  int W, WS, hl, n_strips, n_ysegs, seg_rows;
  int n_thr;                 // max(n_thresholds, 1)
  int cpl, cpl_magic;        // columns per lane in the horizontal pass, ceil(65536 / cpl)
  int y_offset, ny_global;   // y-band tiling: global row of local row 0, rows of the whole grid (ny if not tiled)
  int halo_top, halo_bottom; // rows of the band that are halo (outputs not needed)
  double inv_members;
  ThrSet thr;
};

struct SqPlan {
  int vec, tx, WS, hl, W, n_strips, n_ysegs, seg_rows, mode, mb;
  size_t smem;
};

// how the (row, member) loads of a block are grouped into register chunks
enum { SQ_ROWS = 0,          // n_members == 1: a chunk is MB consecutive rows
       SQ_MEMBERS_FULL = 1,  // n_members == MB: a chunk is one row of all members
       SQ_MEMBERS_PART = 2,  // n_members <  MB
       SQ_CURSOR = 3 };      // n_members >  MB: generic (row, member) cursor

constexpr int kSqMaxThreads = 192;

__device__ __forceinline__ int pad_index(int c, int magic) { return c + ((c * magic) >> 16); }

template <int VEC> __device__ __forceinline__ unsigned ldm_packed(const uint8_t* p);
template <> __device__ __forceinline__ unsigned ldm_packed<1>(const uint8_t* p) { return __ldg(p); }
template <> __device__ __forceinline__ unsigned ldm_packed<2>(const uint8_t* p) {
  return __ldg(reinterpret_cast<const unsigned short*>(p));
}
template <> __device__ __forceinline__ unsigned ldm_packed<4>(const uint8_t* p) {
  return __ldg(reinterpret_cast<const unsigned*>(p));
}

// ---------------------------------------------------------------------------
// Horizontal pass for a batch of completed output rows: one warp per row.
// Builds the float64 inclusive row prefix of the vertical window sums in place
// (lane-local serial prefix + warp-shuffle scan of the lane totals), forms the
// window sum as a prefix difference, normalises and stores float32.
// ---------------------------------------------------------------------------
template <int VEC, bool MASKND>
__device__ __noinline__ void square_hpass(const SqParams& p, double* vbuf, float* cbuf, const double* rtab,
                                          int brow, int batch_y0, long long plane, int t, int x0) {
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int WSP = p.WS + 32, CPL = p.cpl, r = p.r;
  if (warp < brow) {
    const int yo = batch_y0 + warp;
    double* row = vbuf + warp * WSP;
    float* crow = cbuf + warp * WSP;
    const int base = lane * CPL + lane;
    double acc = 0.0;
    float cacc = 0.f;
    for (int j = 0; j < CPL; ++j) {
      acc += row[base + j];
      row[base + j] = acc;
      if (MASKND) { cacc += crow[base + j]; crow[base + j] = cacc; }
    }
    double incl = acc;
    float cincl = cacc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
      if (MASKND) {
        float oc = __shfl_up_sync(0xffffffffu, cincl, d);
        if (lane >= d) cincl += oc;
      }
    }
    const double lane_off = incl - acc;
    const float clane_off = cincl - cacc;
    for (int j = 0; j < CPL; ++j) {
      row[base + j] += lane_off;
      if (MASKND) crow[base + j] += clane_off;
    }
    __syncwarp();
    const bool wrap = (p.flags & NBH_WRAP_X) != 0;
    const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;
    // rows/cols of the window inside the domain (zero padding: nbhood.py:395-397)
    const int rows_in = min(yo + r, p.ny - 1) - max(yo - r, 0) + 1;
    const double rrow = rtab[rows_in] * p.inv_members;
    const long long obase = ((plane * p.n_thr + t) * p.ny + yo) * (long long)p.nx;
    const uint8_t* mnd_plane = MASKND ? p.mask_nd + plane * p.plane_stride : nullptr;
    for (int c = p.hl + lane * VEC; c < p.hl + p.W; c += 32 * VEC) {
      const int xg = x0 + (c - p.hl);
      if (xg >= p.nx) break;
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int cc = c + v;
        const int xo = xg + v;
        const int hi = pad_index(cc + r, p.cpl_magic);
        const int lo = cc - r - 1;
        const int lop = pad_index(max(lo, 0), p.cpl_magic);
        const double S = row[hi] - (lo >= 0 ? row[lop] : 0.0);
        if (MASKND) {
          const float C = crow[hi] - (lo >= 0 ? crow[lop] : 0.f);
          o[v] = sum_only ? (float)S : (C == 0.f ? __int_as_float(0x7fc00000) : (float)(S / (double)C));
        } else if (sum_only) {
          o[v] = (float)S;
        } else if (p.mask2d) {
          o[v] = (float)(S * p.rcp_count[(long long)yo * p.nx + xo] * p.inv_members);
        } else {
          const int cols_in = wrap ? p.nb : (min(xo + r, p.nx - 1) - max(xo - r, 0) + 1);
          o[v] = (float)(S * rrow * rtab[cols_in]);
        }
      }
      VecLoad<VEC>::st(p.out + obase + xg, o);
      if (p.out_mask) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const long long pt = (long long)yo * p.nx + xg + v;
          bool masked = p.mask2d ? (p.mask2d[pt] == 0) : false;
          if (MASKND) masked = masked || (mnd_plane[pt] != 0);
          p.out_mask[obase + xg + v] = masked ? 1 : 0;
        }
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Main kernel.  Register double buffering: while chunk A is consumed the loads
// of chunk B are in flight, so every warp keeps MB x 32 x 4*VEC bytes of HBM
// requests outstanding without staging through shared memory.
// ---------------------------------------------------------------------------
template <int VEC, int TM, bool MASKND, bool M2D, int MODE, int MB>
__global__ void __launch_bounds__(kSqMaxThreads, 2)
square_kernel(const __grid_constant__ SqParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool kMembers = (MODE == SQ_MEMBERS_FULL || MODE == SQ_MEMBERS_PART);
  constexpr int MKN = kMembers ? 1 : MB;       // mask registers per buffer
  const int TX = blockDim.x;
  const int nwarps = TX >> 5;
  const int lane = threadIdx.x & 31;
  const int WS = p.WS, nb = p.nb, r = p.r;
  const int WSP = WS + 32;

  double* vbuf = reinterpret_cast<double*>(smem_raw);                 // [nwarps][WSP]
  double* rtab = vbuf + (size_t)nwarps * WSP;                         // [nb + 1]
  float* ring = reinterpret_cast<float*>(rtab + ((nb + 2) & ~1));     // [nb][WS]
  float* cring = ring + (size_t)nb * WS;                              // MASKND: [nb][WS]
  float* cbuf = cring + (MASKND ? (size_t)nb * WS : 0);               // MASKND: [nwarps][WSP]

  long long bid = blockIdx.x;
  const int strip = (int)(bid % p.n_strips); bid /= p.n_strips;
  const int yseg = (int)(bid % p.n_ysegs); bid /= p.n_ysegs;
  const int t = (int)(bid % p.n_thr); bid /= p.n_thr;
  const long long plane = bid;
  if (p.redo && !p.redo[plane * p.n_thr + t]) return;
  const ThrDev thr = p.thr.t[TM ? t : 0];

  const int ya = yseg * p.seg_rows;
  const int yb = min(p.ny, ya + p.seg_rows);
  const int yend = yb + r;
  const int x0 = strip * p.W;
  const bool wrap = (p.flags & NBH_WRAP_X) != 0;
  const bool sum_only = (p.flags & NBH_SUM_ONLY) != 0;

  const int c0 = threadIdx.x * VEC;
  const int pc0 = pad_index(c0, p.cpl_magic);
  const int gx_raw = x0 - p.hl + c0;
  int gx = gx_raw;
  bool col_active = (gx >= 0) && (gx < p.nx);
  if (wrap) {
    gx %= p.nx; if (gx < 0) gx += p.nx;
    col_active = true;
  }
  // columns that decide the left / right edge-band flags (whole VEC groups only)
  // with a periodic x axis there is no left / right edge (flags preset by the host)
  const bool colband = !sum_only && !wrap && col_active &&
                       ((gx_raw + VEC - 1 <= r) || (gx_raw >= p.nx - 1 - r && gx_raw + VEC - 1 < p.nx));

  for (int k = threadIdx.x + 1; k <= nb; k += TX) rtab[k] = 1.0 / (double)k;
  for (int s = 0; s < nb; ++s) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      ring[s * WS + c0 + v] = 0.f;
      if (MASKND) cring[s * WS + c0 + v] = 0.f;
    }
  }
  __syncthreads();

  double V[VEC];
  float CV[VEC], P[VEC], PC[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { V[v] = 0.0; CV[v] = 0.f; P[v] = 0.f; PC[v] = 0.f; }
  unsigned long long ne_col = 0, ne_top = 0, ne_bot = 0;
  float nanacc = 0.f;
  unsigned cur_m2 = 0x01010101u;
  int slot = 0, brow = 0, batch_y0 = ya;

  const int R = p.n_members;
  const float* in_col = p.in + plane * p.plane_stride + gx;
  const uint8_t* mnd_col = MASKND ? p.mask_nd + plane * p.plane_stride + gx : nullptr;
  const uint8_t* m2_col = M2D ? p.mask2d + gx : nullptr;

  // cursors: (iy, im) next load to issue, (cy, cm) next load to consume;
  // irow / imem: pointers of the issue cursor, advanced incrementally so that
  // a load costs one 64-bit add instead of a multiply-add chain
  int iy = ya - r, im = 0, cy = ya - r, cm = 0;
  const long long mstride = p.member_stride;
  const float* irow = in_col + (long long)iy * p.nx;
  const float* imem = irow;
  const uint8_t* irow_mnd = MASKND ? mnd_col + (long long)iy * p.nx : nullptr;
  const uint8_t* irow_m2 = M2D ? m2_col + (long long)iy * p.nx : nullptr;

  auto finish_row = [&](int yy) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int ri = slot * WS + c0 + v;
      const float old = ring[ri];
      ring[ri] = P[v];
      V[v] += (double)P[v] - (double)old;
      P[v] = 0.f;
      if (MASKND) {
        const float oc = cring[ri];
        cring[ri] = PC[v];
        CV[v] += PC[v] - oc;
        PC[v] = 0.f;
      }
    }
    slot = (slot + 1 == nb) ? 0 : slot + 1;
    const int y = yy - r;
    if (y >= ya) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        vbuf[brow * WSP + pc0 + v] = V[v];
        if (MASKND) cbuf[brow * WSP + pc0 + v] = CV[v];
      }
      ++brow;
      if (brow == nwarps || y == yb - 1) {
        square_hpass<VEC, MASKND>(p, vbuf, cbuf, rtab, brow, batch_y0, plane, t, x0);
        batch_y0 += brow;
        brow = 0;
      }
    }
  };

  // Edge-band bookkeeping for the trimming fix-up: warps that own columns in
  // the left / right band, and all warps on rows within r+1 of the top /
  // bottom edge, run the flagged copy of the accumulate loop.
  const bool warp_band = __any_sync(0xffffffffu, colband);
  unsigned long long ne_row = 0;
  auto row_flags_done = [&](int yy) {
    if (sum_only) return;
    if (colband) ne_col |= ne_row;
    if (yy <= r) ne_top |= ne_row;
    if (yy >= p.ny - 1 - r) ne_bot |= ne_row;
    ne_row = 0;
  };

  auto issue = [&](float (&buf)[MB][VEC], unsigned (&mk)[MKN], unsigned (&mn)[MKN]) {
    if (kMembers) {
      const bool act = col_active && (iy >= 0) && (iy < p.ny) && (iy < yend);
      const float* q = irow;
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        if (act && (MODE == SQ_MEMBERS_FULL || j < R)) VecLoad<VEC>::ld(q, buf[j]);
        q += mstride;
      }
      if (M2D && act) mk[0] = ldm_packed<VEC>(irow_m2);
      irow += p.nx; ++iy;
      if (M2D) irow_m2 += p.nx;
    } else if (MODE == SQ_ROWS) {
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const int yy = iy + j;
        const bool act = col_active && (yy >= 0) && (yy < p.ny) && (yy < yend);
        if (act) {
          VecLoad<VEC>::ld(irow, buf[j]);
          if (MASKND) mn[j] = ldm_packed<VEC>(irow_mnd);
          if (M2D) mk[j] = ldm_packed<VEC>(irow_m2);
        }
        irow += p.nx;
        if (MASKND) irow_mnd += p.nx;
        if (M2D) irow_m2 += p.nx;
      }
      iy += MB;
    } else {
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const bool act = col_active && (iy >= 0) && (iy < p.ny) && (iy < yend);
        if (act) {
          VecLoad<VEC>::ld(imem, buf[j]);
          if (M2D && im == 0) mk[j] = ldm_packed<VEC>(irow_m2);
        }
        imem += mstride;
        if (++im == R) {
          im = 0; ++iy;
          irow += p.nx; imem = irow;
          if (M2D) irow_m2 += p.nx;
        }
      }
    }
  };

  auto consume = [&](float (&buf)[MB][VEC], unsigned (&mk)[MKN], unsigned (&mn)[MKN]) {
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      int yy, m;
      if (MODE == SQ_ROWS) { yy = cy + j; m = 0; }
      else if (kMembers) { yy = cy; m = j; }
      else { yy = cy; m = cm; }
      if (yy < yend) {
        bool act = col_active && (yy >= 0) && (yy < p.ny);
        if (MODE == SQ_MEMBERS_PART) act = act && (j < R);
        if (act) {
          if (M2D && (kMembers ? j == 0 : m == 0)) cur_m2 = mk[kMembers ? 0 : j];
          const unsigned mnd = MASKND ? mn[kMembers ? 0 : j] : 0u;
          auto body = [&](bool flagged) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const float d = buf[j][v];
              bool valid = true, seen = true;
              if (M2D) valid = ((cur_m2 >> (8 * v)) & 0xffu) != 0;
              if (MASKND) { seen = ((mnd >> (8 * v)) & 0xffu) == 0; valid = valid && seen; }
              if (MASKND) { if (seen) nanacc = max_nan(nanacc, fabsf(d)); }
              else nanacc = max_nan(nanacc, fabsf(d));
              float tv = truth_value_sum<TM>(d, thr);
              if (M2D || MASKND) tv = valid ? tv : 0.f;
              P[v] += tv;
              if (MASKND) PC[v] = valid ? 1.f : 0.f;
              if (flagged) { if (!valid || truth_not_one(d, thr)) ne_row |= 1ull << m; }
            }
          };
          if (!sum_only && (warp_band || yy <= r || yy >= p.ny - 1 - r)) body(true); else body(false);
        }
        if (MODE == SQ_ROWS || MODE == SQ_CURSOR) {
          if (m == R - 1) { if (act) row_flags_done(yy); finish_row(yy); }
        }
      }
      if (MODE == SQ_CURSOR) { if (++cm == R) { cm = 0; ++cy; } }
    }
    if (kMembers) {
      if (cy < yend) {
        if (col_active && cy >= 0 && cy < p.ny) row_flags_done(cy);
        finish_row(cy);
      }
    }
    if (MODE == SQ_ROWS) cy += MB;
    else if (kMembers) cy += 1;
  };

  float bufA[MB][VEC], bufB[MB][VEC];
  unsigned mkA[MKN], mkB[MKN], mnA[MKN], mnB[MKN];
#pragma unroll
  for (int j = 0; j < MKN; ++j) { mkA[j] = mkB[j] = 0x01010101u; mnA[j] = mnB[j] = 0u; }
  issue(bufA, mkA, mnA);
  while (cy < yend) {
    issue(bufB, mkB, mnB);
    consume(bufA, mkA, mnA);
    issue(bufA, mkA, mnA);
    consume(bufB, mkB, mnB);
  }

  // ---- epilogue: NaN flag and edge-band flags for the trimming fix-up ----
  if (__isnanf(nanacc)) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  if (!sum_only) {
    const bool left = colband && (gx_raw + VEC - 1 <= r);
    const bool right = colband && (gx_raw >= p.nx - 1 - r);
    unsigned long long masks[4] = {ne_top, ne_bot, left ? ne_col : 0ull, right ? ne_col : 0ull};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)masks[k]);
      unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(masks[k] >> 32));
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

template <int VEC>
int launch_square_vec(const SqParams& p, const SqPlan& pl, long long blocks, int tm, bool masknd, bool m2d,
                      cudaStream_t st);

// Explicit-instantiation helper used by nbh_square_v{1,2,4}.cu
template <int VEC, int TM, bool MASKND, bool M2D, int MODE, int MB>
static int launch_square_one(const SqParams& p, const SqPlan& pl, long long blocks, cudaStream_t st) {
  auto kern = square_kernel<VEC, TM, MASKND, M2D, MODE, MB>;
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  kern<<<(unsigned)blocks, pl.tx, pl.smem, st>>>(p);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

constexpr int sq_rows_mb(int vec) { return vec == 4 ? 4 : 8; }        // rows per chunk, n_members == 1
constexpr int sq_members_mb(int vec) { return vec == 4 ? 9 : 18; }    // members per chunk

template <int VEC, int TM, bool MASKND, bool M2D>
static int launch_square_mode(const SqParams& p, const SqPlan& pl, long long blocks, cudaStream_t st) {
  constexpr int MBR = sq_rows_mb(VEC), MBM = sq_members_mb(VEC);
  if (MASKND || pl.mode == SQ_ROWS)
    return launch_square_one<VEC, TM, MASKND, M2D, SQ_ROWS, MBR>(p, pl, blocks, st);
  if constexpr (!MASKND) {
    switch (pl.mode) {
      case SQ_MEMBERS_FULL: return launch_square_one<VEC, TM, false, M2D, SQ_MEMBERS_FULL, MBM>(p, pl, blocks, st);
      case SQ_MEMBERS_PART: return launch_square_one<VEC, TM, false, M2D, SQ_MEMBERS_PART, MBM>(p, pl, blocks, st);
      default: return launch_square_one<VEC, TM, false, M2D, SQ_CURSOR, MBM>(p, pl, blocks, st);
    }
  }
  return 1;
}

#define NBH_DEFINE_SQUARE_VEC(VEC)                                                                     \
  template <>                                                                                          \
  int launch_square_vec<VEC>(const SqParams& p, const SqPlan& pl, long long blocks, int tm, bool masknd, \
                             bool m2d, cudaStream_t st) {                                              \
    if (masknd && tm == 2) tm = 3; /* numpy.ma arithmetic for masked-array input */                   \
    const int key = (tm << 2) | (masknd ? 2 : 0) | (m2d ? 1 : 0);                                      \
    switch (key) {                                                                                     \
      case 0: return launch_square_mode<VEC, 0, false, false>(p, pl, blocks, st);                      \
      case 1: return launch_square_mode<VEC, 0, false, true>(p, pl, blocks, st);                       \
      case 2: return launch_square_mode<VEC, 0, true, false>(p, pl, blocks, st);                       \
      case 3: return launch_square_mode<VEC, 0, true, true>(p, pl, blocks, st);                        \
      case 4: return launch_square_mode<VEC, 1, false, false>(p, pl, blocks, st);                      \
      case 5: return launch_square_mode<VEC, 1, false, true>(p, pl, blocks, st);                       \
      case 6: return launch_square_mode<VEC, 1, true, false>(p, pl, blocks, st);                       \
      case 7: return launch_square_mode<VEC, 1, true, true>(p, pl, blocks, st);                        \
      case 8: return launch_square_mode<VEC, 2, false, false>(p, pl, blocks, st);                      \
      case 9: return launch_square_mode<VEC, 2, false, true>(p, pl, blocks, st);                       \
      case 14: return launch_square_mode<VEC, 3, true, false>(p, pl, blocks, st);                      \
      case 15: return launch_square_mode<VEC, 3, true, true>(p, pl, blocks, st);                       \
      default: nbh::set_error("unsupported square kernel variant %d", key); return 1;                  \
    }                                                                                                  \
  }

}  // namespace nbh
