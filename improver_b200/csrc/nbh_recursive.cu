// Recursive filter (improver/nbhood/recursive_filter.py:55-202, 403-436): a first-order
// recurrence with per-point coefficients run forward and backward along x, then along y,
// `iterations` times on a slice padded symmetrically by 2 * edge_width.
//
// The recurrence is evaluated in the reference's operation order with one float32 rounding
// per operation (no FMA contraction), so a chain cannot be re-associated into a parallel
// scan; the parallelism is across the rows (x passes) / columns (y passes) of all slices:
//   rf_x_kernel  one lane per (slice, row), a warp per 32 rows: tiles of 32 x 32 points are read
//                and written as coalesced rows and transposed through shared memory, the next
//                tile is in flight while the current one is evaluated; the first forward
//                sweep reads the un-padded input through the symmetric reflection (no
//                separate halo pass)
//   rf_y_kernel  one thread per (slice, column): walks down, then up, eight rows per step;
//                warp accesses are coalesced rows; the last iteration's upward walk writes
//                the un-padded output directly
// Both are bound by HBM traffic (12 B per point and direction) once a few hundred thousand
// chains are in flight; a chain step itself is FMUL + FADD = 8 cycles.
#include "nbh_common.cuh"

namespace nbh {
int env_int(const char* name, int dflt);
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
                                    float* __restrict__ cyp, int ny, int nx, int pad, int NY, int NX, int PX) {
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
    cxp[(s * NY + Y) * (long long)PX + X] = c;
  }
  if (Y < NY - 1) {
    const int i = reflect_symmetric(Y - pad, ny - 1), x = reflect_symmetric(X - pad, nx);
    float c = cy[(long long)i * nx + x];
    if (any_mask && (rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, i, x, nx) ||
                     rf_masked(in, mask, mask_slice_stride, mask_zeros, s, plane, i + 1, x, nx)))
      c = 0.0f;
    cyp[(s * NY + Y) * (long long)PX + X] = c;
  }
}


// x passes.  A warp owns 32 consecutive rows of the flattened (slice, row) list and moves
// along them in tiles of 32 columns: the tile is read row by row (coalesced 128-byte
// requests), transposed through shared memory so that lane l walks row l, and written back
// row by row; the next tile is in flight (registers) while the current one is evaluated.
// dirs bit 0 = forward, bit 1 = backward.  FROM_SRC: the forward sweep reads the un-padded
// input through the symmetric reflection instead of the padded plane (fuses the halo
// construction of pad_cube_with_halo into the first sweep).
template <bool FROM_SRC, int kRfTile, int kRfXWarps>
__global__ void __launch_bounds__(kRfXWarps * 32)
rf_x_kernel(float* __restrict__ g, const float* __restrict__ src, const float* __restrict__ cxp,
            long long n_rows_total, int NY, int NX, long long g_pitch, long long c_pitch, int coeff_per_slice,
            int dirs, int pad, int ny, int nx) {
  constexpr int kRfTilePitch = kRfTile + 1, kSub = kRfTile / 32;
  __shared__ float sg_all[kRfXWarps][32 * kRfTilePitch];
  __shared__ float sc_all[kRfXWarps][32 * kRfTilePitch];
  __shared__ long long off_all[kRfXWarps][3][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sg = sg_all[warp];
  float* sc = sc_all[warp];
  long long* goff = off_all[warp][0];
  long long* coff = off_all[warp][1];
  long long* soff = off_all[warp][2];
  const long long row0 = (blockIdx.x * (long long)kRfXWarps + warp) * 32;
  if (row0 >= n_rows_total) return;
  const int n_rows = (int)min(32LL, n_rows_total - row0);
  {
    const long long R = row0 + lane;
    const long long s = R / NY;
    const int Y = (int)(R - s * NY);
    goff[lane] = R * g_pitch;
    coff[lane] = ((coeff_per_slice ? s : 0) * NY + Y) * c_pitch;
    soff[lane] = FROM_SRC ? (s * ny + reflect_symmetric(Y - pad, ny)) * (long long)nx : 0;
  }
  __syncwarp();
  const int nt = (NX + kRfTile - 1) / kRfTile;
  float pg[32 * kSub], pc[32 * kSub];

  // `whole`: all 32 rows exist and the tile lies inside the row, so no element needs a bounds test
  auto fetch = [&](int t, bool from_src) {
    const bool whole = n_rows == 32 && (t + 1) * kRfTile <= NX - 1;
#pragma unroll
    for (int q = 0; q < kSub; ++q) {
      const int X = t * kRfTile + q * 32 + lane;
      const int xs = (FROM_SRC && from_src) ? reflect_symmetric(X - pad, nx) : 0;
      if (whole) {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          if (FROM_SRC && from_src) pg[rr * kSub + q] = src[soff[rr] + xs];
          else pg[rr * kSub + q] = g[goff[rr] + X];
          pc[rr * kSub + q] = cxp[coff[rr] + X];
        }
      } else {
        const bool gok = X < NX, cok = X < NX - 1;
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const bool rok = rr < n_rows;
          if (FROM_SRC && from_src) pg[rr * kSub + q] = (gok && rok) ? src[soff[rr] + xs] : 0.f;
          else pg[rr * kSub + q] = (gok && rok) ? g[goff[rr] + X] : 0.f;
          pc[rr * kSub + q] = (cok && rok) ? cxp[coff[rr] + X] : 0.f;
        }
      }
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int rr = 0; rr < 32; ++rr)
#pragma unroll
      for (int q = 0; q < kSub; ++q) {
        sg[rr * kRfTilePitch + q * 32 + lane] = pg[rr * kSub + q];
        sc[rr * kRfTilePitch + q * 32 + lane] = pc[rr * kSub + q];
      }
  };
  auto flush = [&](int t) {
    const bool whole = n_rows == 32 && (t + 1) * kRfTile <= NX;
#pragma unroll
    for (int q = 0; q < kSub; ++q) {
      const int X = t * kRfTile + q * 32 + lane;
      if (whole) {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) g[goff[rr] + X] = sg[rr * kRfTilePitch + q * 32 + lane];
      } else if (X < NX) {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr)
          if (rr < n_rows) g[goff[rr] + X] = sg[rr * kRfTilePitch + q * 32 + lane];
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
          const float c = sc[lane * kRfTilePitch + j];
          gprev = rf_step(cprev, sg[lane * kRfTilePitch + j], gprev);
          cprev = c;
          sg[lane * kRfTilePitch + j] = gprev;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kRfTile; ++j) {
          const int i = x0 + j;
          float a = sg[lane * kRfTilePitch + j];
          const float c = sc[lane * kRfTilePitch + j];
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
          gnext = rf_step(sc[lane * kRfTilePitch + j], sg[lane * kRfTilePitch + j], gnext);
          sg[lane * kRfTilePitch + j] = gnext;
        }
      } else {
#pragma unroll
        for (int j = kRfTile - 1; j >= 0; --j) {
          const int i = x0 + j;
          float a = sg[lane * kRfTilePitch + j];
          const float c = sc[lane * kRfTilePitch + j];
          if (i <= NX - 2) a = rf_step(c, a, gnext);
          gnext = a;
          sg[lane * kRfTilePitch + j] = a;
        }
      }
      __syncwarp();
      flush(t);
      __syncwarp();
    }
  }
}

// One directional walk of a column: `n` steps starting at gp / cp, each pointer advancing by its
// (possibly negative) stride.  MODE 0 writes the new values in place, 1 writes nothing, 2 writes
// them to wp.  Full batches issue all their loads before the dependent chain and carry no
// bounds tests.
template <int B, int MODE>
__device__ __forceinline__ float rf_walk(float carry, float* gp, const float* cp, float* wp, int n, long long gs,
                                         long long cs, long long ws) {
  int k = 0;
#pragma unroll 1
  for (; k + B <= n; k += B) {
    float a[B], c[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      a[u] = gp[u * gs];
      c[u] = cp[u * cs];
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      carry = rf_step(c[u], a[u], carry);
      if (MODE == 0) gp[u * gs] = carry;
      if (MODE == 2) wp[u * ws] = carry;
    }
    gp += B * gs;
    cp += B * cs;
    if (MODE == 2) wp += B * ws;
  }
#pragma unroll 1
  for (; k < n; ++k) {
    carry = rf_step(*cp, *gp, carry);
    if (MODE == 0) *gp = carry;
    if (MODE == 2) { *wp = carry; wp += ws; }
    gp += gs;
    cp += cs;
  }
  return carry;
}

// y passes: one thread per (slice, column).  When `out` is set (last iteration) the upward walk
// writes the interior of the slice to the un-padded output instead of the padded plane.
template <int kRfBatch>
__global__ void __launch_bounds__(512)
rf_y_kernel(float* __restrict__ g, const float* __restrict__ cyp, float* __restrict__ out, int NY, int NX,
            long long g_pitch, long long c_pitch, int coeff_per_slice, int dirs, int pad, int ny, int nx) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = blockIdx.y;
  if (X >= NX) return;
  float* gp = g + s * NY * g_pitch + X;
  const float* cp = cyp + (coeff_per_slice ? s : 0) * NY * c_pitch + X;
  if (dirs & 1) rf_walk<kRfBatch, 0>(gp[0], gp + g_pitch, cp, nullptr, NY - 1, g_pitch, c_pitch, 0);
  if (dirs & 2) {
    float* last = gp + (long long)(NY - 2) * g_pitch;
    const float* clast = cp + (long long)(NY - 2) * c_pitch;
    float carry = last[g_pitch];
    if (out == nullptr) {
      rf_walk<kRfBatch, 0>(carry, last, clast, nullptr, NY - 1, -g_pitch, -c_pitch, 0);
      return;
    }
    if (X < pad || X >= pad + nx) return;
    float* op = out + s * (long long)ny * nx + (X - pad);
    if (pad == 0) op[(long long)(ny - 1) * nx] = carry;
    const int top = min(NY - 2, pad + ny - 1);        // first row of the walk that is part of the output
    const int skip = NY - 2 - top;                    // halo rows below it: evaluated, not stored
    carry = rf_walk<kRfBatch, 1>(carry, last, clast, nullptr, skip, -g_pitch, -c_pitch, 0);
    rf_walk<kRfBatch, 2>(carry, last - skip * g_pitch, clast - skip * c_pitch, op + (long long)(top - pad) * nx,
                         top - pad + 1, -g_pitch, -c_pitch, -(long long)nx);
  }
}

inline long long rf_pitch(int NX) { return ((NX + 7) / 8) * 8; }

}  // namespace

size_t recursive_filter_workspace_bytes(long long n_slices, int ny, int nx, int pad, int coeff_per_slice) {
  const long long NY = ny + 2LL * pad, PX = rf_pitch(nx + 2 * pad);
  const long long plane = NY * PX * 4;
  return (size_t)(n_slices * plane + 2 * (coeff_per_slice ? n_slices : 1) * plane + 512);
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
  const long long PX = rf_pitch(NX);
  NBH_REQUIRE(NY <= 65535, "padded slice too tall");
  ProfileScope prof(st, "nbh::rf_x_kernel + nbh::rf_y_kernel (recursive filter)");
  uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  float* g = reinterpret_cast<float*>(base);
  float* cxp = g + n_slices * NY * PX;
  float* cyp = cxp + (per_slice ? n_slices : 1) * NY * PX;
  const long long n_cp = per_slice ? n_slices : 1;
  dim3 cgrid((NX + 255) / 256, NY, (unsigned)n_cp);
  rf_pad_coeff_kernel<<<cgrid, 256, 0, st>>>(in, cx, cy, mask, mask_per_slice ? (long long)ny * nx : 0LL,
                                              mask_zeros, cxp, cyp, ny, nx, pad, NY, NX, (int)PX);
  const long long n_rows = n_slices * NY;
  const int xtile = env_int("NBH_RF_XTILE", 32);
  const int ythreads = env_int("NBH_RF_YTHREADS", 128), ybatch = env_int("NBH_RF_YBATCH", 16);
  dim3 ygrid((NX + ythreads - 1) / ythreads, (unsigned)n_slices);
  for (int it = 0; it < iterations; ++it) {
    // the first forward sweep builds the symmetric halo on the fly
    if (xtile == 64) {
      const unsigned xb = (unsigned)((n_rows + 63) / 64);
      if (it == 0)
        rf_x_kernel<true, 64, 2><<<xb, 64, 0, st>>>(g, in, cxp, n_rows, NY, NX, PX, PX, per_slice, 3, pad, ny, nx);
      else
        rf_x_kernel<false, 64, 2><<<xb, 64, 0, st>>>(g, nullptr, cxp, n_rows, NY, NX, PX, PX, per_slice, 3, pad, ny,
                                                     nx);
    } else {
      const unsigned xb = (unsigned)((n_rows + 127) / 128);
      if (it == 0)
        rf_x_kernel<true, 32, 4><<<xb, 128, 0, st>>>(g, in, cxp, n_rows, NY, NX, PX, PX, per_slice, 3, pad, ny, nx);
      else
        rf_x_kernel<false, 32, 4><<<xb, 128, 0, st>>>(g, nullptr, cxp, n_rows, NY, NX, PX, PX, per_slice, 3, pad, ny,
                                                      nx);
    }
    float* o = it == iterations - 1 ? out : nullptr;
    if (ybatch == 16)
      rf_y_kernel<16><<<ygrid, ythreads, 0, st>>>(g, cyp, o, NY, NX, PX, PX, per_slice, 3, pad, ny, nx);
    else if (ybatch == 4)
      rf_y_kernel<4><<<ygrid, ythreads, 0, st>>>(g, cyp, o, NY, NX, PX, PX, per_slice, 3, pad, ny, nx);
    else
      rf_y_kernel<8><<<ygrid, ythreads, 0, st>>>(g, cyp, o, NY, NX, PX, PX, per_slice, 3, pad, ny, nx);
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
    const long long n_rows = n_slices * ny;
    rf_x_kernel<false, 32, 4><<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(
        grid, nullptr, coeffs, n_rows, ny, nx, nx, nx - 1, 0, dirs, 0, ny, nx);
  } else {
    if (ny < 2) return 0;
    dim3 ygrid((nx + kRfThreads - 1) / kRfThreads, (unsigned)n_slices);
    rf_y_kernel<8><<<ygrid, kRfThreads, 0, st>>>(grid, coeffs, nullptr, ny, nx, nx, nx, 0, dirs, 0, ny, nx);
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nbh
