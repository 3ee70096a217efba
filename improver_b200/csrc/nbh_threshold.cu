// Un-fused thresholding: replaces the hot loop of Threshold.process
// (improver/threshold.py:643-713): per-threshold truth values accumulated over
// the collapsed coordinate in float32 (same order as the reference), divided by
// the integer count of unmasked contributions.
#include "nbh_common.cuh"

namespace nbh {

struct ThrParams {
  const float* in;
  const uint8_t* mask_nd;
  float* out;
  int32_t* contrib;
  int32_t* status;
  long long n_collapse, n_points;
  ThrSet thr;
};

constexpr int kThrRcpTable = 256;     // reciprocals 1/k kept in shared memory for k <= 256 collapsed members

// COLLAPSE == false: one contribution per point (no coordinate collapsed), the mean is the value.
// COLLAPSE == true: np.divide(float32 sum, int64 count) evaluated in float64 and stored as
// float32 (threshold.py:707).  The float64 quotient is formed without a division instruction
// sequence: q0 = a * RN(1/c), rem = fma(-q0, c, a), q = fma(rem, RN(1/c), q0) is the correctly
// rounded a / c (Markstein), with RN(1/c) from a per-block table.
template <int VEC, int TM, int T, bool COLLAPSE>
__global__ void __launch_bounds__(256) threshold_kernel(const ThrParams p) {
  __shared__ double rcp_tab[COLLAPSE ? kThrRcpTable + 1 : 1];
  if (COLLAPSE) {
    for (int k = threadIdx.x; k <= kThrRcpTable; k += blockDim.x) rcp_tab[k] = k ? 1.0 / (double)k : 0.0;
    __syncthreads();
  }
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (i0 >= p.n_points) return;
  float acc[T][VEC];
  int cnt[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    cnt[v] = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t][v] = 0.f;
  }
  float nanacc = 0.f;
  ThrDev thr[T];                                 // keep the threshold constants in registers
#pragma unroll
  for (int t = 0; t < T; ++t) thr[t] = p.thr.t[t];
  for (long long c = 0; c < p.n_collapse; ++c) {
    float d[VEC];
    uint8_t m[VEC];
    VecLoad<VEC>::ld(p.in + c * p.n_points + i0, d);
#pragma unroll
    for (int v = 0; v < VEC; ++v) m[v] = 0;
    if (p.mask_nd) VecLoad<VEC>::ldm(p.mask_nd + c * p.n_points + i0, m);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (m[v]) continue;                      // masked points never contribute (threshold.py:695)
      nanacc = max_nan(nanacc, fabsf(d[v]));
      cnt[v] += 1;
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t][v] = __fadd_rn(acc[t][v], truth_value_t<TM>(d[v], thr[t]));
    }
  }
  if (__isnanf(nanacc)) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float o[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (!COLLAPSE) {
        o[v] = cnt[v] ? acc[t][v] : 0.f;
      } else if (cnt[v] <= 1) {
        o[v] = cnt[v] ? acc[t][v] : 0.f;          // a count of 1 is the identity
      } else if (cnt[v] <= kThrRcpTable) {
        const double a = (double)acc[t][v], c = (double)cnt[v], rc = rcp_tab[cnt[v]];
        const double q0 = a * rc;
        o[v] = (float)fma(fma(-q0, c, a), rc, q0);
      } else {
        o[v] = (float)((double)acc[t][v] / (double)cnt[v]);
      }
    }
    VecLoad<VEC>::st(p.out + (long long)t * p.n_points + i0, o);
  }
  if (p.contrib) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) p.contrib[i0 + v] = cnt[v];
  }
}

template <int VEC, int TM, bool COLLAPSE>
static void launch_threshold_c(const ThrParams& p, int n, unsigned blocks, cudaStream_t st) {
  switch (n) {
    case 1: threshold_kernel<VEC, TM, 1, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
    case 2: threshold_kernel<VEC, TM, 2, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
    case 3: threshold_kernel<VEC, TM, 3, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
    case 4: threshold_kernel<VEC, TM, 4, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
    case 5: threshold_kernel<VEC, TM, 5, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
    case 6: threshold_kernel<VEC, TM, 6, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
    case 7: threshold_kernel<VEC, TM, 7, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
    default: threshold_kernel<VEC, TM, 8, COLLAPSE><<<blocks, 256, 0, st>>>(p); break;
  }
}

template <int VEC, int TM>
static void launch_threshold_t(const ThrParams& p, int n, unsigned blocks, cudaStream_t st) {
  if (p.n_collapse > 1) launch_threshold_c<VEC, TM, true>(p, n, blocks, st);
  else launch_threshold_c<VEC, TM, false>(p, n, blocks, st);
}

template <int VEC>
static void launch_threshold(const ThrParams& p, int mode, int n, unsigned blocks, cudaStream_t st) {
  if (mode == 1) launch_threshold_t<VEC, 1>(p, n, blocks, st);
  else if (mode == 2) launch_threshold_t<VEC, 2>(p, n, blocks, st);
  else launch_threshold_t<VEC, 3>(p, n, blocks, st);
}

int run_threshold(const float* in, const uint8_t* mask_nd, float* out, int32_t* contrib,
                  const NbhThreshold* thresholds, int n_thresholds, long long n_collapse,
                  long long n_points, int32_t* status, cudaStream_t st) {
  NBH_REQUIRE(in && out && status && thresholds, "null pointer argument");
  NBH_REQUIRE(n_thresholds >= 1 && n_thresholds <= kMaxThresholds, "n_thresholds must be in [1, %d]",
              kMaxThresholds);
  NBH_REQUIRE(n_collapse >= 1 && n_points >= 1, "empty input");
  NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  const bool al16 = (n_points % 4 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                    (!mask_nd || (uintptr_t)mask_nd % 4 == 0);
  // one launch per run of thresholds sharing an evaluation mode (the kernel is
  // specialised on it: deterministic compare / fuzzy float32 / fuzzy numpy.ma)
  int t0 = 0;
  while (t0 < n_thresholds) {
    ThrParams p;
    memset(&p, 0, sizeof(p));
    int n = 0, mode = 0;
    for (int i = t0; i < n_thresholds; ++i) {
      ThrDev d;
      if (make_thr_dev(thresholds[i], &d, mask_nd != nullptr)) return 1;
      if (n == 0) mode = d.mode;
      if (d.mode != mode) break;
      p.thr.t[n++] = d;
    }
    p.thr.n = n;
    p.in = in; p.mask_nd = mask_nd; p.out = out + (long long)t0 * n_points; p.contrib = contrib;
    p.status = status; p.n_collapse = n_collapse; p.n_points = n_points;
    const long long th = al16 ? n_points / 4 : n_points;
    const unsigned blocks = (unsigned)((th + 255) / 256);
    if (al16) launch_threshold<4>(p, mode, n, blocks, st);
    else launch_threshold<1>(p, mode, n, blocks, st);
    t0 += n;
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// Collapse of the leading axis by its arithmetic mean: collapse_realizations
// (improver/utilities/cube_manipulation.py:93-131, iris MEAN = numpy's mean of a float32 array
// over axis 0: rows added one after the other in float32, one float32 division).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
collapse_mean_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_collapse, long long n_points) {
  const float nf = (float)n_collapse;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_points;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = __ldg(in + i);
    for (long long m = 1; m < n_collapse; ++m) acc = __fadd_rn(acc, __ldg(in + m * n_points + i));
    out[i] = __fdiv_rn(acc, nf);
  }
}

int run_collapse_mean(const float* in, float* out, long long n_collapse, long long n_points, cudaStream_t st) {
  NBH_REQUIRE(in && out, "null pointer argument");
  NBH_REQUIRE(n_collapse >= 1 && n_points >= 1, "empty input");
  const unsigned grid = (unsigned)max(1LL, min((n_points + 255) / 256, (long long)sm_count() * 16));
  collapse_mean_kernel<<<grid, 256, 0, st>>>(in, out, n_collapse, n_points);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Unit conversion of the data before thresholding (improver/threshold.py:430-431, cf_units on a
// float32 array): out = float32(a * double(in) + b).
__global__ void __launch_bounds__(256)
linear_map_kernel(const float* __restrict__ in, float* __restrict__ out, double a, double b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)__dadd_rn(__dmul_rn(a, (double)in[i]), b);
}

int run_linear_map(const float* in, float* out, double a, double b, long long n, cudaStream_t st) {
  NBH_REQUIRE(in && out && n >= 1, "invalid argument");
  const unsigned grid = (unsigned)max(1LL, min((n + 255) / 256, (long long)sm_count() * 16));
  linear_map_kernel<<<grid, 256, 0, st>>>(in, out, a, b, n);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// boxsum (improver/utilities/neighbourhood_tools.py:129-211) in float64 for rectangular boxes:
// out[i][j] = sum of in[i+1 .. i+by][j+1 .. j+bx] (cumsum = 1: `in` holds values, the window is
// summed directly in float64, row by row) or the four-corner difference of an already cumulative
// `in` (cumsum = 0).  in (n, H, W) -> out (n, H - by, W - bx).
__global__ void __launch_bounds__(256)
boxsum_f64_kernel(const double* __restrict__ in, double* __restrict__ out, long long n, int H, int W, int by, int bx,
                  int cumsum) {
  const int m = H - by, nn = W - bx;
  const long long total = n * m * (long long)nn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / ((long long)m * nn);
    const long long rem = i - s * (long long)m * nn;
    const int y = (int)(rem / nn), x = (int)(rem - (long long)y * nn);
    const double* p = in + s * (long long)H * W;
    double r;
    if (cumsum) {
      r = 0.0;
      for (int dy = 1; dy <= by; ++dy) {
        double row = 0.0;
        for (int dx = 1; dx <= bx; ++dx) row += p[(long long)(y + dy) * W + (x + dx)];
        r += row;
      }
    } else {
      r = p[(long long)(y + by) * W + (x + bx)] - p[(long long)y * W + (x + bx)] + p[(long long)y * W + x] -
          p[(long long)(y + by) * W + x];
    }
    out[i] = r;
  }
}

int run_boxsum_f64(const double* in, double* out, long long n, int H, int W, int by, int bx, int cumsum,
                   cudaStream_t st) {
  NBH_REQUIRE(in && out, "null pointer argument");
  NBH_REQUIRE(n >= 1 && by >= 1 && bx >= 1 && H > by && W > bx, "box of %d x %d does not fit a %d x %d array", by, bx, H, W);
  const long long total = n * (long long)(H - by) * (W - bx);
  const unsigned grid = (unsigned)max(1LL, min((total + 255) / 256, (long long)sm_count() * 32));
  boxsum_f64_kernel<<<grid, 256, 0, st>>>(in, out, n, H, W, by, bx, cumsum);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Zone collapse of ApplyNeighbourhoodProcessingWithAMask (improver/nbhood/use_nbhood.py:158-192):
// weighted mean over the zone axis with the NaN results excluded and the weights renormalised
// (np.ma.average semantics); all-NaN -> NaN.  values (n_lead, n_zones, n_points),
// weights (n_zones, n_points) with masked weights already zeroed.
__global__ void __launch_bounds__(256)
zone_collapse_kernel(const float* __restrict__ values, const float* __restrict__ weights, float* __restrict__ out,
                     long long n_lead, int n_zones, long long n_points) {
  const long long total = n_lead * n_points;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long l = i / n_points, pt = i - l * n_points;
    float num = 0.f, den = 0.f;
    for (int z = 0; z < n_zones; ++z) {
      const float v = values[(l * n_zones + z) * n_points + pt];
      const float w = weights[(long long)z * n_points + pt];
      if (!__isnanf(v)) { num = __fadd_rn(num, __fmul_rn(v, w)); den = __fadd_rn(den, w); }
    }
    out[i] = __fdiv_rn(num, den);          // 0 / 0 -> NaN where no zone is valid
  }
}

int run_zone_collapse(const float* values, const float* weights, float* out, long long n_lead, int n_zones,
                      long long n_points, cudaStream_t st) {
  NBH_REQUIRE(values && weights && out, "null pointer argument");
  NBH_REQUIRE(n_lead >= 1 && n_zones >= 1 && n_points >= 1, "empty input");
  const long long total = n_lead * n_points;
  const unsigned grid = (unsigned)max(1LL, min((total + 255) / 256, (long long)sm_count() * 16));
  zone_collapse_kernel<<<grid, 256, 0, st>>>(values, weights, out, n_lead, n_zones, n_points);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nbh
