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

template <int VEC, int TM>
__global__ void __launch_bounds__(256) threshold_kernel(const ThrParams p) {
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (i0 >= p.n_points) return;
  float acc[kMaxThresholds][VEC];
  int cnt[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    cnt[v] = 0;
#pragma unroll
    for (int t = 0; t < kMaxThresholds; ++t) acc[t][v] = 0.f;
  }
  float nanacc = 0.f;
  const int T = p.thr.n;
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
      for (int t = 0; t < kMaxThresholds; ++t)
        if (t < T) acc[t][v] = __fadd_rn(acc[t][v], truth_value_t<TM>(d[v], p.thr.t[t]));
    }
  }
  if (__isnanf(nanacc)) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
#pragma unroll
  for (int t = 0; t < kMaxThresholds; ++t) {
    if (t < T) {
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v)   // np.divide(float32, int64) -> float64 -> float32 store
        o[v] = cnt[v] ? (float)((double)acc[t][v] / (double)cnt[v]) : 0.f;
      VecLoad<VEC>::st(p.out + (long long)t * p.n_points + i0, o);
    }
  }
  if (p.contrib) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) p.contrib[i0 + v] = cnt[v];
  }
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
    if (al16) {
      if (mode == 1) threshold_kernel<4, 1><<<blocks, 256, 0, st>>>(p);
      else if (mode == 2) threshold_kernel<4, 2><<<blocks, 256, 0, st>>>(p);
      else threshold_kernel<4, 3><<<blocks, 256, 0, st>>>(p);
    } else {
      if (mode == 1) threshold_kernel<1, 1><<<blocks, 256, 0, st>>>(p);
      else if (mode == 2) threshold_kernel<1, 2><<<blocks, 256, 0, st>>>(p);
      else threshold_kernel<1, 3><<<blocks, 256, 0, st>>>(p);
    }
    t0 += n;
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nbh
