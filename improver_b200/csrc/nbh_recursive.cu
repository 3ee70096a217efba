// Recursive filter (improver/nbhood/recursive_filter.py:55-202, 403-436): a first-order
// recurrence with per-point coefficients run forward and backward along x, then along y,
// `iterations` times on a slice padded symmetrically by 2 * edge_width.
//
// The recurrence is evaluated in the reference's operation order with one float32 rounding
// per operation (no FMA contraction), so a chain cannot be re-associated into a parallel
// scan; the parallelism is across the rows (x passes) / columns (y passes) of all slices.
//
// Workspace layout: the padded planes (data and the two coefficient planes) are stored as
// 32 x 32 tiles of 4 KB, [slice][tile row][tile column][32][32].  Both sweep directions then
// touch whole DRAM pages: an x sweep moves a warp along one tile row (4 KB contiguous per
// step), a y sweep moves a warp down one tile column (128 B per row, 32 consecutive rows
// contiguous).  With row-major planes either sweep direction fetched 128 B per DRAM page
// visit and both kernels stalled at ~3.5 TB/s of DRAM traffic.
//   rf_x_kernel  one lane per row, a warp per tile row: a tile is read as coalesced rows,
//                transposed through shared memory so that lane l walks row l, and written
//                back; the next tile is in flight while the current one is evaluated; the
//                first forward sweep reads the un-padded input through the symmetric
//                reflection (no separate halo pass)
//   rf_y_kernel  one thread per (slice, column): walks down, then up, a batch of rows per
//                step; the last iteration's upward walk writes the un-padded output directly
//   fusion       by default the downward (y-forward) walk runs inside rf_x_kernel's backward
//                sweep as a tile-row wavefront (see RfFuse) and rf_y_kernel only walks up
// A chain step itself is FMUL + FADD = 8 cycles; both kernels are bound by memory traffic.
#include "nbh_common.cuh"

namespace nbh {
namespace {

constexpr int kRfThreads = 128;

__device__ __forceinline__ int reflect_symmetric(int i, int n) {
  // np.pad(mode="symmetric"): ... 1 0 | 0 1 2 ... n-1 | n-1 n-2 ...  (period 2 n)
  const int p = 2 * n;
  int m = i % p;
  if (m < 0) m += p;
  return m < n ? m : p - 1 - m;
}

// (1 - c) * a + c * b with numpy's float32 roundings (recursive_filter.py:97-102, 147-152)
__device__ __forceinline__ float rf_step(float c, float a, float b) {
  return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, c), a), __fmul_rn(c, b));
}

// element (Y, X) of slice s in the tiled plane layout
__device__ __forceinline__ long long rf_tiled(long long s, int Y, int X, int TY, int TX) {
  return (((s * TY + (Y >> 5)) * TX + (X >> 5)) << 10) + ((Y & 31) << 5) + (X & 31);
}

__device__ __forceinline__ bool rf_masked(const float* in, const uint8_t* mask, long long mask_slice_stride,
                                          int mask_zeros, long long s, long long plane, int y, int x, int nx) {
  const long long o = (long long)y * nx + x;
  bool m = false;
  if (mask) m = mask[s * mask_slice_stride + o] != 0;
  if (mask_zeros) m = m || in[s * plane + o] == 0.0f;
  return m;
}

// padded coefficient planes; coefficients adjacent to a masked point are zeroed before padding
// (recursive_filter.py:262-273, 410-426)
__global__ void rf_pad_coeff_kernel(const float* __restrict__ in, const float* __restrict__ cx,
                                    const float* __restrict__ cy, const uint8_t* __restrict__ mask,
                                    long long mask_slice_stride, int mask_zeros, float* __restrict__ cxp,
                                    float* __restrict__ cyp, int ny, int nx, int pad, int NY, int NX, int TY,
                                    int TX) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  const long long s = blockIdx.z;
  if (X >= NX) return;
  const long long plane = (long long)ny * nx;
  const bool any_mask = mask != nullptr || mask_zeros;
  if (X < NX - 1) {
    const int y = reflect_symmetric(Y - pad, ny), j = reflect_symmetric(X - pad, nx - 1);
    float c = cx[(long long)y * (nx - 1) + j];
    if (any_mask && (rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, y, j, nx) ||
                     rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, y, j + 1, nx)))
      c = 0.0f;
    // x-coefficient tiles are stored transposed ([column][row]): the x sweep's lane = row reads
    // the coefficient of column j as one coalesced row of the tile, no shared-memory transpose
    cxp[(((s * TY + (Y >> 5)) * TX + (X >> 5)) << 10) + ((X & 31) << 5) + (Y & 31)] = c;
  }
  if (Y < NY - 1) {
    const int i = reflect_symmetric(Y - pad, ny - 1), x = reflect_symmetric(X - pad, nx);
    float c = cy[(long long)i * nx + x];
    if (any_mask && (rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, i, x, nx) ||
                     rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, i + 1, x, nx)))
      c = 0.0f;
    cyp[rf_tiled(s, Y, X, TY, TX)] = c;
  }
}


// Per-slice masks: instead of one pair of coefficient planes per slice (8 B per point written and
// 16 B read per iteration) the sweeps keep the shared planes and read one byte per point that says
// "this coefficient is zeroed in this slice" (zx transposed like the x coefficients).
__global__ void __launch_bounds__(256)
rf_zero_flags_kernel(const float* __restrict__ in, const uint8_t* __restrict__ mask, long long mask_slice_stride,
                     int mask_zeros, uint8_t* __restrict__ zx, uint8_t* __restrict__ zy, int ny, int nx, int pad,
                     int NY, int NX, int TY, int TX) {
  // one block of 32 x 8 threads per 32 x 32 tile; threadIdx.x is the fast index of the tile being written
  const long long s = blockIdx.z;
  const long long tile = ((s * TY + blockIdx.y) * TX + blockIdx.x) << 10;
  const long long plane = (long long)ny * nx;
  const int y0 = blockIdx.y * 32 - pad, x0 = blockIdx.x * 32 - pad;   // data coordinates of the tile origin
  const int tx = threadIdx.x;
  if (y0 >= 0 && y0 + 32 <= ny - 1 && x0 >= 0 && x0 + 32 <= nx - 1) {
    // tile away from the halo: no reflection; stage the 33 x 33 patch of "masked" bits once
    __shared__ uint8_t sm[33][36];
    const int tid = threadIdx.y * 32 + tx;
    bool v[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int idx = tid + q * 256, r = idx / 33, c = idx - r * 33;
      v[q] = idx < 33 * 33 && rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, y0 + r, x0 + c, nx);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int idx = tid + q * 256, r = idx / 33, c = idx - r * 33;
      if (idx < 33 * 33) sm[r][c] = v[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ty = threadIdx.y + 8 * q;
      zy[tile + (ty << 5) + tx] = sm[ty][tx] | sm[ty + 1][tx];
      zx[tile + (ty << 5) + tx] = sm[tx][ty] | sm[tx][ty + 1];
    }
    return;
  }
  for (int ty = threadIdx.y; ty < 32; ty += 8) {
    {
      const int X = blockIdx.x * 32 + tx, Y = blockIdx.y * 32 + ty;
      if (X < NX && Y < NY - 1) {
        const int i = reflect_symmetric(Y - pad, ny - 1), x = reflect_symmetric(X - pad, nx);
        zy[tile + (ty << 5) + tx] = rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, i, x, nx) ||
                                    rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, i + 1, x, nx);
      }
    }
    {
      const int X = blockIdx.x * 32 + ty, Y = blockIdx.y * 32 + tx;   // transposed tile
      if (X < NX - 1 && Y < NY) {
        const int y = reflect_symmetric(Y - pad, ny), j = reflect_symmetric(X - pad, nx - 1);
        zx[tile + (ty << 5) + tx] = rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, y, j, nx) ||
                                    rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, y, j + 1, nx);
      }
    }
  }
}

constexpr int kRfTile = 32;
constexpr int kRfTilePitch = kRfTile + 1;
constexpr int kRfXWarps = 4;

// x passes.  A warp owns 32 rows and moves along them in tiles of 32 columns: the tile is read
// row by row (coalesced 128-byte requests), transposed through shared memory so that lane l
// walks row l, and written back row by row; the next tile is in flight (registers) while the
// current one is evaluated.  dirs bit 0 = forward, bit 1 = backward.
//   TILED     planes in the tiled workspace layout, warp = (slice, tile row); otherwise row-major
//             planes with the given pitches, warp = 32 consecutive rows of the flattened row list
//   FROM_SRC  the forward sweep reads the un-padded input through the symmetric reflection
//             instead of the padded plane (fuses pad_cube_with_halo into the first sweep)
// Fusion of the y-forward sweep into the x-backward sweep (tile-row wavefront, NBH_RF_FUSE=0 turns
// it off): after a
// warp has finished the x-backward values of tile (ty, t) it continues the y-forward chains of the
// tile's 32 columns through its 32 rows, starting from row 31 of tile (ty - 1, t), which the warp
// of the tile row above publishes through a per-(slice, tile row) progress counter.  The
// x-backward plane is then never written nor re-read (-26 % DRAM traffic per iteration).  Units
// are ordered tile row major, slice minor, so that a warp's predecessor was dispatched
// n_slices units earlier and is (nearly always) ahead of it.
struct RfFuse {
  const float* cyp;          // y coefficients (tiled), nullptr = no fusion
  const uint8_t* zy;         // per-slice zero flags or nullptr
  unsigned* progress;        // [slice][tile row] tiles finished by the y-forward part
  unsigned* ticket;          // fused sweeps: units are claimed in ticket order (forward progress, see below)
  long long n_slices;
};

template <bool FROM_SRC, bool TILED>
__global__ void __launch_bounds__(kRfXWarps * 32, 4)
rf_x_kernel(float* __restrict__ g, const float* __restrict__ src, const float* __restrict__ cxp,
            long long n_units, long long n_rows_total, int NY, int NX, long long g_pitch, long long c_pitch,
            int coeff_per_slice, int dirs, int pad, int ny, int nx, int TY, int TX,
            const uint8_t* __restrict__ zx, const RfFuse fuse) {
  __shared__ float sg_all[kRfXWarps][32 * kRfTilePitch];
  __shared__ float sc_all[kRfXWarps][TILED ? 1 : 32 * kRfTilePitch];
  __shared__ long long off_all[kRfXWarps][TILED ? 1 : 2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sg = sg_all[warp];
  float* sc = sc_all[warp];
  long long* off0 = off_all[warp][0];                      // TILED: source rows; else data rows
  long long* off1 = off_all[warp][TILED ? 0 : 1];          // row-major coefficient rows
  // Fused sweeps: a warp waits for the warp that owns the tile row above (unit - n_slices).  Units
  // are therefore CLAIMED from an atomic ticket counter instead of being derived from blockIdx: the
  // unit a warp depends on has a lower ticket, i.e. it was claimed by a warp that is already
  // running, so the wait below always ends whatever order the blocks are scheduled in.
  long long unit = blockIdx.x * (long long)kRfXWarps + warp;
  if (TILED && fuse.cyp != nullptr) {
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(fuse.ticket, 1u);
    unit = (long long)__shfl_sync(0xffffffffu, tk, 0);
  }
  if (unit >= n_units) return;
  int n_rows;
  long long gbase = 0, cbase = 0;      // TILED: first tile of this warp's tile row
  long long fs = 0;                    // slice / tile row of this warp (TILED)
  int fty = 0;
  if (TILED) {
    const bool fused = fuse.cyp != nullptr;
    const long long s = fused ? unit % fuse.n_slices : unit / TY;
    const int ty = fused ? (int)(unit / fuse.n_slices) : (int)(unit - s * TY);
    fs = s;
    fty = ty;
    n_rows = min(32, NY - 32 * ty);
    gbase = ((s * TY + ty) * TX) << 10;
    cbase = (((coeff_per_slice ? s : 0) * TY + ty) * TX) << 10;
    if (FROM_SRC) off0[lane] = (s * ny + reflect_symmetric(32 * ty + lane - pad, ny)) * (long long)nx;
  } else {
    const long long R = unit * 32 + lane;
    n_rows = (int)min(32LL, n_rows_total - unit * 32);
    const long long s = R / NY;
    const int Y = (int)(R - s * NY);
    off0[lane] = R * g_pitch;
    off1[lane] = ((coeff_per_slice ? s : 0) * NY + Y) * c_pitch;
  }
  __syncwarp();
  const int nt = (NX + kRfTile - 1) / kRfTile;
  float pg[32], pc[32];

  auto fetch = [&](int t, bool from_src) {
    const int X = t * kRfTile + lane;
    if (TILED) {
      const float* ct = cxp + cbase + ((long long)t << 10) + lane;   // transposed tile: [column][row]
      if (zx) {
        const uint8_t* zt = zx + gbase + ((long long)t << 10) + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) pc[j] = zt[j * 32] ? 0.f : ct[j * 32];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pc[j] = ct[j * 32];
      }
      if (FROM_SRC && from_src) {
        const int xs = reflect_symmetric(min(X, NX - 1) - pad, nx);
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) pg[rr] = src[off0[rr] + xs];
      } else {
        const float* gt = g + gbase + ((long long)t << 10) + lane;
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) pg[rr] = gt[rr * 32];
      }
    } else {
      const bool gok = X < NX, cok = X < NX - 1;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        const bool rok = rr < n_rows;
        pg[rr] = (gok && rok) ? g[off0[rr] + X] : 0.f;
        pc[rr] = (cok && rok) ? cxp[off1[rr] + X] : 0.f;
      }
    }
  };
  float cc[32];   // TILED: this lane's coefficients of the staged tile
  auto stage = [&]() {
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      sg[rr * kRfTilePitch + lane] = pg[rr];
      if (TILED) cc[rr] = pc[rr];
      else sc[rr * kRfTilePitch + lane] = pc[rr];
    }
  };
  auto flush = [&](int t) {
    if (TILED) {   // whole tiles: the slack of edge tiles belongs to the workspace
      float* gt = g + gbase + ((long long)t << 10) + lane;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) gt[rr * 32] = sg[rr * kRfTilePitch + lane];
    } else {
      const int X = t * kRfTile + lane;
      if (X < NX) {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr)
          if (rr < n_rows) g[off0[rr] + X] = sg[rr * kRfTilePitch + lane];
      }
    }
  };

  if (dirs & 1) {
    float gprev = 0.f, cprev = 0.f;
    fetch(0, true);
#pragma unroll 1
    for (int t = 0; t < nt; ++t) {
      stage();
      __syncwarp();
      if (t + 1 < nt) fetch(t + 1, true);
      const int x0 = t * kRfTile;
      if (t > 0 && x0 + kRfTile <= NX) {
#pragma unroll
        for (int j = 0; j < kRfTile; ++j) {
          const float c = TILED ? cc[j] : sc[lane * kRfTilePitch + j];
          gprev = rf_step(cprev, sg[lane * kRfTilePitch + j], gprev);
          cprev = c;
          sg[lane * kRfTilePitch + j] = gprev;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kRfTile; ++j) {
          const int i = x0 + j;
          float a = sg[lane * kRfTilePitch + j];
          const float c = TILED ? cc[j] : sc[lane * kRfTilePitch + j];
          if (i >= 1 && i < NX) a = rf_step(cprev, a, gprev);
          gprev = a;
          cprev = c;
          sg[lane * kRfTilePitch + j] = a;
        }
      }
      __syncwarp();
      flush(t);
      __syncwarp();
    }
  }
  if (dirs & 2) {
    float gnext = 0.f;
    fetch(nt - 1, false);
#pragma unroll 1
    for (int t = nt - 1; t >= 0; --t) {
      stage();
      __syncwarp();
      if (t > 0) fetch(t - 1, false);
      const int x0 = t * kRfTile;
      if (x0 + kRfTile <= NX - 1) {
#pragma unroll
        for (int j = kRfTile - 1; j >= 0; --j) {
          gnext = rf_step(TILED ? cc[j] : sc[lane * kRfTilePitch + j], sg[lane * kRfTilePitch + j], gnext);
          sg[lane * kRfTilePitch + j] = gnext;
        }
      } else {
#pragma unroll
        for (int j = kRfTile - 1; j >= 0; --j) {
          const int i = x0 + j;
          float a = sg[lane * kRfTilePitch + j];
          const float c = TILED ? cc[j] : sc[lane * kRfTilePitch + j];
          if (i <= NX - 2) a = rf_step(c, a, gnext);
          gnext = a;
          sg[lane * kRfTilePitch + j] = a;
        }
      }
      __syncwarp();
      if (TILED && fuse.cyp != nullptr) {
        // y-forward through this tile, lane = column
        const long long tile = gbase + ((long long)t << 10);
        const long long up = tile - ((long long)TX << 10);          // tile (ty - 1, t)
        float carry = 0.f, cprev = 0.f;
        if (fty > 0) {
          if (lane == 0) {
            const volatile unsigned* p = fuse.progress + fs * TY + (fty - 1);
            const unsigned need = (unsigned)(nt - t);
            // the producer holds a lower ticket: it is running and at most a few tiles behind
            while (*p < need) __nanosleep(64);
          }
          __threadfence();
          __syncwarp();
          carry = __ldcg(g + up + 31 * 32 + lane);
          cprev = fuse.cyp[up - (fs * TY * ((long long)TX << 10)) + 31 * 32 + lane];
          if (fuse.zy && fuse.zy[up + 31 * 32 + lane]) cprev = 0.f;
        }
        const float* ct = fuse.cyp + (tile - fs * TY * ((long long)TX << 10)) + lane;
        const uint8_t* zt = fuse.zy ? fuse.zy + tile + lane : nullptr;
        float* gt = g + tile + lane;
        float cy[32];
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          cy[rr] = ct[rr * 32];
          if (zt && zt[rr * 32]) cy[rr] = 0.f;
        }
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const float v = sg[rr * kRfTilePitch + lane];
          carry = (fty > 0 || rr > 0) ? rf_step(cprev, v, carry) : v;
          cprev = cy[rr];
          gt[rr * 32] = carry;
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(fuse.progress + fs * TY + fty, 1u);
      } else {
        flush(t);
      }
      __syncwarp();
    }
  }
}

// Offset of row Y within one column of a plane: tiled workspace layout (stride = floats per
// tile row) or row-major (stride = pitch).
template <bool TILED>
__device__ __forceinline__ long long rf_row(int Y, long long stride) {
  return TILED ? (long long)(Y >> 5) * stride + ((Y & 31) << 5) : (long long)Y * stride;
}

// One directional walk of a column: `n` steps from row y_first in direction dir (+1 / -1); the
// step at row Y uses the coefficient of row Y + c_shift.  MODE 0 writes the new values in place,
// 1 writes nothing, 2 writes them to wp (advancing by ws per step).  Full batches issue all their
// loads before the dependent chain.
template <int B, int MODE, bool TILED, bool Z>
__device__ __forceinline__ float rf_walk(float carry, float* gcol, const float* ccol, const uint8_t* zcol, float* wp,
                                         int y_first, int dir, int c_shift, int n, long long gs, long long cs,
                                         long long ws) {
  int k = 0;
#pragma unroll 1
  for (; k + B <= n; k += B) {
    float a[B], c[B];
    long long o[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const int Y = y_first + dir * (k + u);
      o[u] = rf_row<TILED>(Y, gs);
      a[u] = gcol[o[u]];
      const long long oc = rf_row<TILED>(Y + c_shift, cs);
      c[u] = ccol[oc];
      if (Z && zcol[oc]) c[u] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      carry = rf_step(c[u], a[u], carry);
      if (MODE == 0) gcol[o[u]] = carry;
      if (MODE == 2) wp[u * ws] = carry;
    }
    if (MODE == 2) wp += B * ws;
  }
#pragma unroll 1
  for (; k < n; ++k) {
    const int Y = y_first + dir * k;
    const long long o = rf_row<TILED>(Y, gs);
    const long long oc = rf_row<TILED>(Y + c_shift, cs);
    float c1 = ccol[oc];
    if (Z && zcol[oc]) c1 = 0.f;
    carry = rf_step(c1, gcol[o], carry);
    if (MODE == 0) gcol[o] = carry;
    if (MODE == 2) { *wp = carry; wp += ws; }
  }
  return carry;
}

// y passes: one thread per (slice, column).  When `out` is set (last iteration) the upward walk
// writes the interior of the slice to the un-padded output instead of the padded plane.
template <int kRfBatch, bool TILED, bool Z>
__global__ void __launch_bounds__(512)
rf_y_kernel(float* __restrict__ g, const float* __restrict__ cyp, float* __restrict__ out, int NY, int NX,
            long long g_pitch, long long c_pitch, int coeff_per_slice, int dirs, int pad, int ny, int nx, int TY,
            int TX, const uint8_t* __restrict__ zy) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = blockIdx.y;
  if (X >= NX) return;
  float* gp;
  const float* cp;
  const uint8_t* zp = nullptr;
  long long gs, cs;
  if (TILED) {
    const long long col = ((long long)(X >> 5) << 10) + (X & 31);
    gs = cs = (long long)TX << 10;
    gp = g + s * TY * gs + col;
    if (Z) zp = zy + s * TY * gs + col;
    cp = cyp + (coeff_per_slice ? s : 0) * TY * cs + col;
  } else {
    gs = g_pitch;
    cs = c_pitch;
    gp = g + s * NY * g_pitch + X;
    cp = cyp + (coeff_per_slice ? s : 0) * NY * c_pitch + X;
  }
  if (dirs & 1) rf_walk<kRfBatch, 0, TILED, Z>(gp[0], gp, cp, zp, nullptr, 1, 1, -1, NY - 1, gs, cs, 0);
  if (dirs & 2) {
    float carry = gp[rf_row<TILED>(NY - 1, gs)];
    if (out == nullptr) {
      rf_walk<kRfBatch, 0, TILED, Z>(carry, gp, cp, zp, nullptr, NY - 2, -1, 0, NY - 1, gs, cs, 0);
      return;
    }
    if (X < pad || X >= pad + nx) return;
    float* op = out + s * (long long)ny * nx + (X - pad);
    if (pad == 0) op[(long long)(ny - 1) * nx] = carry;
    const int top = min(NY - 2, pad + ny - 1);        // first row of the walk that is part of the output
    const int skip = NY - 2 - top;                    // halo rows below it: evaluated, not stored
    carry = rf_walk<kRfBatch, 1, TILED, Z>(carry, gp, cp, zp, nullptr, NY - 2, -1, 0, skip, gs, cs, 0);
    rf_walk<kRfBatch, 2, TILED, Z>(carry, gp, cp, zp, op + (long long)(top - pad) * nx, top, -1, 0, top - pad + 1, gs, cs,
                                -(long long)nx);
  }
}

inline long long rf_tiles(int n) { return (n + 31) / 32; }

}  // namespace

size_t recursive_filter_workspace_bytes(long long n_slices, int ny, int nx, int pad, int coeff_per_slice) {
  const long long plane = rf_tiles(ny + 2 * pad) * rf_tiles(nx + 2 * pad) * 4096;
  // data planes, one pair of coefficient planes, and one pair of byte planes per slice for per-slice masks
  const long long counters = n_slices * rf_tiles(ny + 2 * pad) * 4 + 256;   // wavefront progress (fused sweeps)
  return (size_t)(n_slices * plane + 2 * plane + (coeff_per_slice ? 2 * n_slices * (plane / 4) : 0) + counters + 512);
}

int run_recursive_filter(const float* in, const float* cx, const float* cy, const uint8_t* mask, int mask_per_slice,
                         int mask_zeros, float* out, long long n_slices, int ny, int nx, int pad, int iterations,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
  NBH_REQUIRE(in && cx && cy && out, "null data / coefficient / output pointer");
  NBH_REQUIRE(ny >= 2 && nx >= 2, "the recursive filter needs at least 2 x 2 points (got %d x %d)", ny, nx);
  NBH_REQUIRE(pad >= 0 && iterations >= 1, "pad must be >= 0 and iterations >= 1");
  NBH_REQUIRE(n_slices >= 0 && n_slices <= 65535, "n_slices out of range (0 .. 65535 per call)");
  if (n_slices == 0) return 0;
  const int per_slice = (mask_zeros || (mask && mask_per_slice)) ? 1 : 0;
  NBH_REQUIRE(ws && ws_bytes >= recursive_filter_workspace_bytes(n_slices, ny, nx, pad, per_slice),
              "workspace too small for the recursive filter");
  const int NY = ny + 2 * pad, NX = nx + 2 * pad;
  const int TY = (int)rf_tiles(NY), TX = (int)rf_tiles(NX);
  const long long plane = (long long)TY * TX * 1024;
  NBH_REQUIRE(NY <= 65535, "padded slice too tall");
  ProfileScope prof(st, "nbh::rf_x_kernel + nbh::rf_y_kernel (recursive filter)");
  uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  float* g = reinterpret_cast<float*>(base);
  float* cxp = g + n_slices * plane;
  float* cyp = cxp + plane;
  uint8_t* zx = per_slice ? reinterpret_cast<uint8_t*>(cyp + plane) : nullptr;
  uint8_t* zy = per_slice ? zx + n_slices * plane : nullptr;
  uint8_t* after = per_slice ? zy + n_slices * plane : reinterpret_cast<uint8_t*>(cyp + plane);
  unsigned* progress = reinterpret_cast<unsigned*>((reinterpret_cast<uintptr_t>(after) + 255) & ~uintptr_t(255));
  const bool fuse_on = true;      // y-forward sweep fused into the x-backward sweep (ticket-ordered wavefront)
  unsigned* ticket = progress + n_slices * TY;
  RfFuse fuse{fuse_on ? cyp : nullptr, zy, progress, ticket, n_slices};
  // a mask shared by all slices is folded into the coefficient planes; per-slice masks become flags
  dim3 cgrid((NX + 255) / 256, NY, 1);
  rf_pad_coeff_kernel<<<cgrid, 256, 0, st>>>(in, cx, cy, per_slice ? nullptr : mask, 0LL, 0, cxp, cyp, ny, nx, pad,
                                              NY, NX, TY, TX);
  if (per_slice) {
    dim3 zgrid(TX, TY, (unsigned)n_slices);
    rf_zero_flags_kernel<<<zgrid, dim3(32, 8), 0, st>>>(in, mask, mask && mask_per_slice ? (long long)ny * nx : 0LL,
                                                mask_zeros, zx, zy, ny, nx, pad, NY, NX, TY, TX);
  }
  const long long n_units = n_slices * TY;
  const unsigned xb = (unsigned)((n_units + kRfXWarps - 1) / kRfXWarps);
  const int ythreads = 128;
  dim3 ygrid((NX + ythreads - 1) / ythreads, (unsigned)n_slices);
  for (int it = 0; it < iterations; ++it) {
    if (fuse_on) NBH_CHECK_CUDA(cudaMemsetAsync(progress, 0, ((size_t)n_slices * TY + 1) * 4, st));
    // the first forward sweep builds the symmetric halo on the fly
    if (it == 0) {
      rf_x_kernel<true, true><<<xb, kRfXWarps * 32, 0, st>>>(g, in, cxp, n_units, 0, NY, NX, 0, 0, 0, 3, pad, ny, nx,
                                                             TY, TX, zx, fuse);
    } else {
      rf_x_kernel<false, true><<<xb, kRfXWarps * 32, 0, st>>>(g, nullptr, cxp, n_units, 0, NY, NX, 0, 0, 0, 3, pad,
                                                              ny, nx, TY, TX, zx, fuse);
    }
    float* o = it == iterations - 1 ? out : nullptr;
    const int ydirs = fuse_on ? 2 : 3;     // fused: the x kernel has already run the downward sweep
    if (per_slice)
      rf_y_kernel<16, true, true><<<ygrid, ythreads, 0, st>>>(g, cyp, o, NY, NX, 0, 0, 0, ydirs, pad, ny, nx, TY, TX,
                                                              zy);
    else
      rf_y_kernel<16, true, false><<<ygrid, ythreads, 0, st>>>(g, cyp, o, NY, NX, 0, 0, 0, ydirs, pad, ny, nx, TY,
                                                               TX, nullptr);
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// One directional pass in place on contiguous (n_slices, ny, nx) data: the array-level
// RecursiveFilter._recurse_forward / _recurse_backward (recursive_filter.py:55-153).
// axis 0 = y (coeffs (ny - 1, nx)), axis 1 = x (coeffs (ny, nx - 1)); dirs 1 forward, 2 backward.
int run_recurse(float* grid, const float* coeffs, long long n_slices, int ny, int nx, int axis, int dirs,
                cudaStream_t st) {
  NBH_REQUIRE(grid && coeffs, "null grid / coefficient pointer");
  NBH_REQUIRE(axis == 0 || axis == 1, "axis must be 0 (y) or 1 (x)");
  NBH_REQUIRE(dirs >= 1 && dirs <= 3, "dirs must be 1 (forward), 2 (backward) or 3 (both)");
  NBH_REQUIRE(ny >= 1 && nx >= 1 && n_slices >= 0 && n_slices <= 65535, "bad shape");
  if (n_slices == 0) return 0;
  if (axis == 1) {
    if (nx < 2) return 0;
    const long long n_rows = n_slices * ny, n_units = (n_rows + 31) / 32;
    rf_x_kernel<false, false><<<(unsigned)((n_units + kRfXWarps - 1) / kRfXWarps), kRfXWarps * 32, 0, st>>>(
        grid, nullptr, coeffs, n_units, n_rows, ny, nx, nx, nx - 1, 0, dirs, 0, ny, nx, 0, 0, nullptr, RfFuse{});
  } else {
    if (ny < 2) return 0;
    dim3 ygrid((nx + kRfThreads - 1) / kRfThreads, (unsigned)n_slices);
    rf_y_kernel<8, false, false><<<ygrid, kRfThreads, 0, st>>>(grid, coeffs, nullptr, ny, nx, nx, nx, 0, dirs, 0,
                                                               ny, nx, 0, 0, nullptr);
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nbh
