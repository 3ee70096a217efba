// Shared device/host helpers for the improver_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/improver_b200.h"

namespace nbh {

void set_error(const char* fmt, ...);

#define NBH_CHECK_CUDA(expr)                                                         \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      nbh::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                      \
      return 2;                                                                      \
    }                                                                                \
  } while (0)

#define NBH_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      nbh::set_error(__VA_ARGS__);   \
      return 1;                      \
    }                                \
  } while (0)

constexpr int kMaxThresholds = 8;   // thresholds evaluated per launch group

// Device-side description of one threshold.  All float32 so that the truth
// value follows numpy's float32 arithmetic at improver/threshold.py:441-455 /
// improver/utilities/rescale.py:63-65 operation by operation.
struct ThrDev {
  float thr;      // float32(threshold)                       threshold.py:434
  float lo;       // float32(bounds[0])   subtracted below thr
  float hi;       // float32(bounds[1])
  float den_lo;   // thr - lo             (float32 subtraction)
  float den_hi;   // float32(bounds[1]) - thr
  float den2_lo, den2_hi;   // 2 * den
  float rcph_lo, rcph_hi;   // RN(0.5 / den), for the exact division sequence
  int mode;       // 0 raw data, 1 deterministic compare, 2 fuzzy, 3 fuzzy with np.ma arithmetic
  int op;         // NBH_OP_*
  double lo64;    // bounds[0] unrounded (np.ma arithmetic promotes to float64)
  // branch-free evaluation used inside the neighbourhood sums (truth_value_fast)
  float fs_lo, fs_hi;   // +-0.5/den_lo, +-0.5/den_hi (negated for "less than" comparisons)
  float fs1, fc1;       // single-slope forms (symmetric bounds): sat((d-thr)*fs1 + .5), sat(d*fs1 + fc1)
  int fast_kind;        // 2 general, 4 single slope, 5 single slope + single FMA
  // exact "truth value == 1" test on the data itself (edge-band flags of the
  // trimming fix-up): one_lo <= d <= one_hi  <=>  truth_value(d) == 1.0f
  float one_lo, one_hi;
};

struct ThrSet {
  ThrDev t[kMaxThresholds];
  int n;          // 0 => raw
};

int make_thr_dev(const NbhThreshold& h, ThrDev* d, bool ma_arith = false);

// Exact division: with r = RN(1/(2 den)) from the host, q0 = a*r,
// rem = a - 2den*q0 (exact, FMA), q = q0 + rem*r is the correctly rounded
// a/(2 den) (Markstein).  Checked bit for bit against numpy in the GPU tests.
// Threshold._calculate_truth_value for one unmasked value (threshold.py:407-471).
// MODE: 0 raw data, 1 deterministic comparison, 2 fuzzy (float32 arithmetic),
// 3 fuzzy with numpy.ma arithmetic.
template <int MODE>
__device__ __forceinline__ float truth_value_t(float d, const ThrDev& p) {
  if (MODE == 0) return d;
  if (MODE == 1) {
    bool b;
    switch (p.op) {
      case NBH_OP_GT: b = d > p.thr; break;
      case NBH_OP_GE: b = d >= p.thr; break;
      case NBH_OP_LT: b = d < p.thr; break;
      case NBH_OP_LE: b = d <= p.thr; break;
      default: b = d == p.thr; break;
    }
    return b ? 1.0f : 0.0f;
  }
  if (MODE == 3) {
    // Masked-array input: numpy.ma wraps the python-float operands in 0-d
    // float64 arrays, so under NumPy 2 promotion the reference evaluates the
    // rescale in float64 (the float32 scalar subtractions thr-lo / hi-thr and
    // data-thr stay float32) and rounds once at the end (threshold.py:441-471).
    const bool below = d < p.thr;
    const double a = below ? ((double)d - p.lo64) : (double)__fsub_rn(d, p.thr);
    const double den = below ? (double)p.den_lo : (double)p.den_hi;
    const double off = below ? 0.0 : 0.5;
    double v = a * 0.5 / den + off;
    v = fmin(fmax(v, off), off + 0.5);
    if ((p.op == NBH_OP_LT || p.op == NBH_OP_LE)) v = 1.0 - v;
    return (float)v;
  }
  // fuzzy: where(d < thr, clip((d-lo)*0.5/(thr-lo) + 0, 0, .5),
  //                        clip((d-thr)*0.5/(hi-thr) + .5, .5, 1))
  // Clamping the data to [lo, hi] first is equivalent to the reference's clip
  // of the quotient (float32 subtraction, division by den > 0 and *0.5 are
  // all monotone) and keeps +-inf inputs finite.  The *0.5 is folded into the
  // exact division: RN(a*0.5/den) = a/(2 den) correctly rounded.
  const float dc = fminf(fmaxf(d, p.lo), p.hi);
  const bool below = d < p.thr;
  const float base = below ? p.lo : p.thr;
  const float den2 = below ? p.den2_lo : p.den2_hi;     // 2 * denominator
  const float rcph = below ? p.rcph_lo : p.rcph_hi;     // RN(0.5 / denominator)
  const float off = below ? 0.0f : 0.5f;
  const float a = __fsub_rn(dc, base);
  const float q0 = __fmul_rn(a, rcph);
  const float rem = __fmaf_rn(-den2, q0, a);
  float v = __fadd_rn(__fmaf_rn(rem, rcph, q0), off);
  if ((p.op == NBH_OP_LT || p.op == NBH_OP_LE)) v = __fsub_rn(1.0f, v);     // threshold.py:468-469
  return v;
}

// Fast fuzzy truth value for use inside neighbourhood sums.  The membership
// function is piecewise linear through (lo,0), (thr,.5), (hi,1); both pieces
// pass through (thr, .5), so with a = d - thr it is sat(.5 + a * slope) with
// the slope picked by the sign of a.  "less than" comparisons negate the
// slopes (1 - v = .5 - a * slope).  Four instructions; within 1e-7 of the
// correctly rounded reference value, exactly 0 / 1 beyond the bounds.  The
// exact form (truth_value_t<2>) stays in use wherever truth values themselves
// are output or compared.
__device__ __forceinline__ float truth_value_fast(float d, const ThrDev& p) {
  const float a = __fsub_rn(d, p.thr);
  const float s = (a < 0.0f) ? p.fs_lo : p.fs_hi;
  return __saturatef(__fmaf_rn(a, s, 0.5f));
}

// truth_value(d) != 1 without evaluating it (bounds found on the host by
// bisection over the exact float32 formula); NaN counts as "not one"
__device__ __forceinline__ bool truth_not_one(float d, const ThrDev& p) {
  return !(d >= p.one_lo && d <= p.one_hi);
}

// TM codes of the neighbourhood kernels: 0 raw, 1 compare, 2 fuzzy, 3 fuzzy np.ma,
// 4 / 5 fuzzy with symmetric bounds (fuzzy_factor): below and above the
// threshold the membership has the same slope (up to float32 rounding of the
// two denominators), so the sign select disappears: 2 instructions (4) or a
// single saturating FMA (5, when |d*slope| stays small enough for 5e-7 accuracy)
template <int TM>
__device__ __forceinline__ float truth_value_sum(float d, const ThrDev& p) {
  if (TM == 2) return truth_value_fast(d, p);
  if (TM == 4) return __saturatef(__fmaf_rn(__fsub_rn(d, p.thr), p.fs1, 0.5f));
  if (TM == 5) return __saturatef(__fmaf_rn(d, p.fs1, p.fc1));
  return truth_value_t<TM == 4 || TM == 5 ? 2 : TM>(d, p);
}

// runtime-dispatched form (un-fused threshold kernel, fix-up kernels)
__device__ __forceinline__ float truth_value(float d, const ThrDev& p) {
  switch (p.mode) {
    case 0: return d;
    case 1: return truth_value_t<1>(d, p);
    case 2: return truth_value_t<2>(d, p);
    default: return truth_value_t<3>(d, p);
  }
}

__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

template <int VEC> struct VecLoad;
template <> struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
  static __device__ __forceinline__ void ldm(const uint8_t* p, uint8_t* m) { m[0] = __ldg(p); }
};
template <> struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
#if defined(NBH_LD_NOALLOC)
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
#else
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
#endif
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
  static __device__ __forceinline__ void ldm(const uint8_t* p, uint8_t* m) {
    uchar2 t = __ldg(reinterpret_cast<const uchar2*>(p));
    m[0] = t.x; m[1] = t.y;
  }
};
template <> struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void ldm(const uint8_t* p, uint8_t* m) {
    uchar4 t = __ldg(reinterpret_cast<const uchar4*>(p));
    m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
  }
};

// bench.py measurement aid: events around the dominant kernel of a call
size_t recursive_filter_workspace_bytes(long long n_slices, int ny, int nx, int pad, int coeff_per_slice);
int run_recursive_filter(const float* in, const float* cx, const float* cy, const uint8_t* mask, int mask_per_slice,
                         int mask_zeros, float* out, long long n_slices, int ny, int nx, int pad, int iterations,
                         void* ws, size_t ws_bytes, cudaStream_t st);
int run_recurse(float* grid, const float* coeffs, long long n_slices, int ny, int nx, int axis, int dirs,
                cudaStream_t st);

struct ProfileScope {
  cudaStream_t st;
  bool on;
  explicit ProfileScope(cudaStream_t s, const char* name = "");
  ~ProfileScope();
};

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace nbh
