// Recursive filter (improver/nbhood/recursive_filter.py:55-202, 403-436): a first-order
// recurrence with per-point coefficients run forward and backward along x, then along y,
// `iterations` times on a slice padded symmetrically by 2 * edge_width.
//
// The recurrence is evaluated in the reference's operation order with one float32 rounding
// per operation (no FMA contraction), so a chain cannot be re-associated into a parallel
// scan; the parallelism is across the rows (x passes) / columns (y passes) of all slices:
//   rf_x_kernel  one thread per (slice, row): walks the row forward, then backward, eight
//                columns (one 32-byte sector of data and one of coefficients) per step,
//                the next step's sectors requested before the current chain is evaluated
//   rf_y_kernel  one thread per (slice, column): walks down, then up, eight rows per step;
//                warp accesses are coalesced rows; the last iteration's upward walk writes
//                the un-padded output directly
// Both are bound by HBM traffic (12 B per point and direction) once a few hundred thousand
// chains are in flight; a chain step itself is FMUL + FADD = 8 cycles.
#include "nbh_common.cuh"

namespace nbh {
namespace {

constexpr int kRfThreads = 128;
constexpr int kRfBatch = 8;

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

// padded data: g[s][Y][X] = in[s][refl(Y - pad)][refl(X - pad)]
__global__ void rf_pad_data_kernel(const float* __restrict__ in, float* __restrict__ g, int ny, int nx, int pad,
                                   int NY, int NX, int PX) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  const long long s = blockIdx.z;
  if (X >= NX) return;
  const int y = reflect_symmetric(Y - pad, ny), x = reflect_symmetric(X - pad, nx);
  g[(s * NY + Y) * (long long)PX + X] = in[(s * ny + y) * (long long)nx + x];
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

struct Chunk8 {
  float v[8];
};

template <bool VEC>
__device__ __forceinline__ Chunk8 rf_load8(const float* p, int i0, int n) {
  Chunk8 c;
  if (VEC) {
    const float4 a = *reinterpret_cast<const float4*>(p + i0);
    const float4 b = *reinterpret_cast<const float4*>(p + i0 + 4);
    c.v[0] = a.x; c.v[1] = a.y; c.v[2] = a.z; c.v[3] = a.w;
    c.v[4] = b.x; c.v[5] = b.y; c.v[6] = b.z; c.v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) c.v[j] = i0 + j < n ? p[i0 + j] : 0.0f;
  }
  return c;
}

template <bool VEC>
__device__ __forceinline__ void rf_store8(float* p, int i0, int n, const Chunk8& c) {
  if (VEC) {
    *reinterpret_cast<float4*>(p + i0) = make_float4(c.v[0], c.v[1], c.v[2], c.v[3]);
    *reinterpret_cast<float4*>(p + i0 + 4) = make_float4(c.v[4], c.v[5], c.v[6], c.v[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i0 + j < n) p[i0 + j] = c.v[j];
  }
}

// x passes.  VEC: rows are 32-byte aligned with at least roundup(NX, 8) readable floats.
// dirs bit 0 = forward, bit 1 = backward.  Coefficient row r holds NX - 1 values.
template <bool VEC>
__global__ void __launch_bounds__(kRfThreads)
rf_x_kernel(float* __restrict__ g, const float* __restrict__ cxp, long long n_rows_total, int NY, int NX,
            long long g_pitch, long long c_pitch, int coeff_per_slice, int dirs) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_rows_total) return;
  const long long s = t / NY;
  const int row = (int)(t - s * NY);
  float* gr = g + t * g_pitch;
  const float* cr = cxp + ((coeff_per_slice ? s : 0) * NY + row) * c_pitch;
  const int nch = (NX + 7) >> 3;
  if (dirs & 1) {
    float gprev = 0.f, cprev = 0.f;
    Chunk8 gc = rf_load8<VEC>(gr, 0, NX), cc = rf_load8<VEC>(cr, 0, NX - 1);
#pragma unroll 1
    for (int k = 0; k < nch; ++k) {
      Chunk8 gn = gc, cn = cc;
      if (k + 1 < nch) {
        gn = rf_load8<VEC>(gr, 8 * (k + 1), NX);
        cn = rf_load8<VEC>(cr, 8 * (k + 1), NX - 1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = 8 * k + j;
        if (i >= 1 && i < NX) gc.v[j] = rf_step(cprev, gc.v[j], gprev);
        gprev = gc.v[j];
        cprev = cc.v[j];
      }
      rf_store8<VEC>(gr, 8 * k, NX, gc);
      gc = gn;
      cc = cn;
    }
  }
  if (dirs & 2) {
    float gnext = 0.f;
    Chunk8 gc = rf_load8<VEC>(gr, 8 * (nch - 1), NX), cc = rf_load8<VEC>(cr, 8 * (nch - 1), NX - 1);
#pragma unroll 1
    for (int k = nch - 1; k >= 0; --k) {
      Chunk8 gn = gc, cn = cc;
      if (k > 0) {
        gn = rf_load8<VEC>(gr, 8 * (k - 1), NX);
        cn = rf_load8<VEC>(cr, 8 * (k - 1), NX - 1);
      }
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        const int i = 8 * k + j;
        if (i <= NX - 2) gc.v[j] = rf_step(cc.v[j], gc.v[j], gnext);
        gnext = gc.v[j];
      }
      rf_store8<VEC>(gr, 8 * k, NX, gc);
      gc = gn;
      cc = cn;
    }
  }
}

// y passes.  When `out` is set (last iteration) the upward walk writes the interior of the
// slice to the un-padded output instead of the padded plane.
__global__ void __launch_bounds__(kRfThreads)
rf_y_kernel(float* __restrict__ g, const float* __restrict__ cyp, float* __restrict__ out, int NY, int NX,
            long long g_pitch, long long c_pitch, int coeff_per_slice, int dirs, int pad, int ny, int nx) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = blockIdx.y;
  if (X >= NX) return;
  float* gp = g + s * NY * g_pitch + X;
  const float* cp = cyp + (coeff_per_slice ? s : 0) * NY * c_pitch + X;
  if (dirs & 1) {
    float prev = gp[0];
    for (int i0 = 1; i0 < NY; i0 += kRfBatch) {
      float a[kRfBatch], c[kRfBatch];
#pragma unroll
      for (int u = 0; u < kRfBatch; ++u) {
        const int i = i0 + u;
        a[u] = i < NY ? gp[i * g_pitch] : 0.f;
        c[u] = i < NY ? cp[(i - 1) * c_pitch] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kRfBatch; ++u) {
        const int i = i0 + u;
        if (i < NY) {
          prev = rf_step(c[u], a[u], prev);
          gp[i * g_pitch] = prev;
        }
      }
    }
  }
  if (dirs & 2) {
    const bool emit = out != nullptr;
    const bool col_in = X >= pad && X < pad + nx;
    if (emit && !col_in) return;
    float* op = emit ? out + s * (long long)ny * nx + (X - pad) : nullptr;
    const int stop = emit ? pad : 0;   // rows above the interior are not part of the output
    float next = gp[(long long)(NY - 1) * g_pitch];
    if (emit && pad == 0) op[(long long)(NY - 1) * nx] = next;
    for (int i0 = NY - 2; i0 >= stop; i0 -= kRfBatch) {
      float a[kRfBatch], c[kRfBatch];
#pragma unroll
      for (int u = 0; u < kRfBatch; ++u) {
        const int i = i0 - u;
        a[u] = i >= stop ? gp[i * g_pitch] : 0.f;
        c[u] = i >= stop ? cp[i * c_pitch] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kRfBatch; ++u) {
        const int i = i0 - u;
        if (i >= stop) {
          next = rf_step(c[u], a[u], next);
          if (!emit) gp[i * g_pitch] = next;
          else if (i < pad + ny) op[(long long)(i - pad) * nx] = next;
        }
      }
    }
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
  dim3 pgrid((NX + 255) / 256, NY, (unsigned)n_slices);
  rf_pad_data_kernel<<<pgrid, 256, 0, st>>>(in, g, ny, nx, pad, NY, NX, (int)PX);
  dim3 cgrid((NX + 255) / 256, NY, (unsigned)n_cp);
  rf_pad_coeff_kernel<<<cgrid, 256, 0, st>>>(in, cx, cy, mask, mask_per_slice ? (long long)ny * nx : 0LL,
                                              mask_zeros, cxp, cyp, ny, nx, pad, NY, NX, (int)PX);
  const long long n_rows = n_slices * NY;
  const unsigned xblocks = (unsigned)((n_rows + kRfThreads - 1) / kRfThreads);
  dim3 ygrid((NX + kRfThreads - 1) / kRfThreads, (unsigned)n_slices);
  for (int it = 0; it < iterations; ++it) {
    rf_x_kernel<true><<<xblocks, kRfThreads, 0, st>>>(g, cxp, n_rows, NY, NX, PX, PX, per_slice, 3);
    rf_y_kernel<<<ygrid, kRfThreads, 0, st>>>(g, cyp, it == iterations - 1 ? out : nullptr, NY, NX, PX, PX,
                                              per_slice, 3, pad, ny, nx);
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
    rf_x_kernel<false><<<(unsigned)((n_rows + kRfThreads - 1) / kRfThreads), kRfThreads, 0, st>>>(
        grid, coeffs, n_rows, ny, nx, nx, nx - 1, 0, dirs);
  } else {
    if (ny < 2) return 0;
    dim3 ygrid((nx + kRfThreads - 1) / kRfThreads, (unsigned)n_slices);
    rf_y_kernel<<<ygrid, kRfThreads, 0, st>>>(grid, coeffs, nullptr, ny, nx, nx, nx, 0, dirs, 0, ny, nx);
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nbh
