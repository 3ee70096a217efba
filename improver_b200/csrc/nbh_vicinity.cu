// Threshold "in vicinity": truth value -> square maximum filter -> accumulation over the
// collapsed coordinate.  Replaces Threshold._vicinity_processing (improver/threshold.py:473-518)
// and maximum_within_vicinity / operator_within_vicinity (improver/utilities/spatial.py:740-855:
// scipy.ndimage.maximum_filter(size = 2 * radius + 1, mode = "nearest"), masked points and
// points of the other surface type replaced by a large negative fill, land and sea processed
// separately) inside the accumulation loop of Threshold.process (threshold.py:661-707).
//
// The truth value is a non-decreasing function of the data for ">" / ">=" and non-increasing for
// "<" / "<=", and the maximum commutes with a monotone function bit for bit:
//     max_i tv(d_i) == tv(max_i d_i)   (resp. tv(min_i d_i)),
// so ONE filter of the raw field per member serves every threshold: the kernel filters
// s * d (s = +1 / -1) and applies the exact float32 truth value (truth_value_t, the same code as
// the un-fused threshold kernel) to the filtered value.  "==" is not monotone: those launches
// filter the truth value itself, one threshold at a time.  Edge replication equals ignoring the
// cells outside the domain for a maximum.
//
// A block owns an 8 x 128 tile of one output plane and loops over the members; per member the
// tile plus its halo is staged in shared memory once per surface type (points that do not
// contribute are -inf), then for each vicinity radius a horizontal running maximum into a second
// buffer and a vertical maximum into registers (separable).  Sums are float32 in member order,
// the mean is the correctly rounded float64 quotient (as in nbh_threshold.cu).
#include "nbh_common.cuh"

namespace nbh {

constexpr int kVicTY = 8, kVicTX = 128, kVicThreads = 256;
constexpr int kVicMaxOut = 8;        // (vicinity radii) x (thresholds) accumulated per launch
constexpr int kVicMaxRadius = 24;
constexpr int kVicRcpTable = 256;

struct VicParams {
  const float* in;
  const uint8_t* mask_nd;
  const uint8_t* landmask;
  float* out;              // (V, T, planes, ny, nx) with strides below
  int32_t* contrib;        // (planes, ny, nx) or null
  int32_t* status;
  long long n_collapse, n_planes;
  int ny, nx;
  int n_vic, n_thr, vmax;
  int cells[kVicMaxOut];
  long long out_vic_stride, out_thr_stride;   // element strides of the output's V and T axes
  int eq_mode;             // 1: filter the truth value of threshold 0 itself ("==")
  float sign;              // +1: maximum of d, -1: maximum of -d (less-than comparisons)
  ThrSet thr;
};

template <int TM, bool LAND>
__global__ void __launch_bounds__(kVicThreads) vicinity_kernel(const VicParams p) {
  extern __shared__ float vsm[];
  __shared__ double rcp_tab[kVicRcpTable + 1];
  const int vmax = p.vmax;
  const int SW = kVicTX + 2 * vmax;                  // staged width
  const int SH = kVicTY + 2 * vmax;                  // staged height
  constexpr int NT = LAND ? 2 : 1;                   // surface types filtered separately
  float* stage = vsm;                                // [NT][SH][SW]
  float* hbuf = vsm + (size_t)NT * SH * SW;          // [NT][SH][kVicTX]
  for (int k = threadIdx.x; k <= kVicRcpTable; k += blockDim.x) rcp_tab[k] = k ? 1.0 / (double)k : 0.0;

  const int tiles_x = (p.nx + kVicTX - 1) / kVicTX, tiles_y = (p.ny + kVicTY - 1) / kVicTY;
  long long b = blockIdx.x;
  const int tx = (int)(b % tiles_x); b /= tiles_x;
  const int ty = (int)(b % tiles_y); b /= tiles_y;
  const long long plane = b;
  const int x0 = tx * kVicTX, y0 = ty * kVicTY;
  const long long npts = (long long)p.ny * p.nx;

  // this thread's four output points: row r, columns c .. c + 3
  const int r = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
  const int gy = y0 + r;
  float acc[kVicMaxOut][4];
  int cnt[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < kVicMaxOut; ++k)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[k][v] = 0.f;
  bool my_land[4] = {false, false, false, false};
  if (LAND && gy < p.ny) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (x0 + c + v < p.nx) my_land[v] = p.landmask[(long long)gy * p.nx + x0 + c + v] != 0;
  }
  float nanacc = 0.f;
  const float NEG = __int_as_float(0xff800000);      // -inf
  const ThrDev thr0 = p.thr.t[0];

  for (long long m = 0; m < p.n_collapse; ++m) {
    const float* src = p.in + (m * p.n_planes + plane) * npts;
    const uint8_t* msk = p.mask_nd ? p.mask_nd + (m * p.n_planes + plane) * npts : nullptr;
    __syncthreads();                                 // previous member's buffers are free
    // ---- stage the tile + halo (field value or -inf) ----
    for (int i = threadIdx.x; i < SH * SW; i += kVicThreads) {
      const int sy = i / SW, sx = i - sy * SW;
      const int yy = y0 - vmax + sy, xx = x0 - vmax + sx;
      float f0 = NEG, f1 = NEG;
      if (yy >= 0 && yy < p.ny && xx >= 0 && xx < p.nx) {
        const long long pt = (long long)yy * p.nx + xx;
        const bool masked = msk && msk[pt];
        if (!masked) {
          const float d = src[pt];
          nanacc = max_nan(nanacc, fabsf(d));
          const float f = p.eq_mode ? truth_value_t<TM>(d, thr0) : p.sign * d;
          if (LAND) {
            if (p.landmask[pt]) f0 = f; else f1 = f;
          } else {
            f0 = f;
          }
        }
      }
      stage[i] = f0;
      if (LAND) stage[SH * SW + i] = f1;
    }
    __syncthreads();
    // masked centre points do not contribute (threshold.py:516 `[unmasked] +=`)
    bool okp[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      okp[v] = gy < p.ny && x0 + c + v < p.nx && !(msk && msk[(long long)gy * p.nx + x0 + c + v]);
      cnt[v] += okp[v] ? 1 : 0;
    }
    for (int iv = 0; iv < p.n_vic; ++iv) {
      const int vr = p.cells[iv];
      // ---- horizontal running maximum: rows y0 - vr .. y0 + TY - 1 + vr of every type ----
      const int rows = kVicTY + 2 * vr;
      if (iv) __syncthreads();                       // hbuf readers of the previous radius are done
      for (int i = threadIdx.x; i < NT * rows * kVicTX; i += kVicThreads) {
        const int t = i / (rows * kVicTX);
        const int rem = i - t * rows * kVicTX;
        const int hy = rem / kVicTX, hx = rem - hy * kVicTX;
        const float* row = stage + (size_t)t * SH * SW + (size_t)(hy + vmax - vr) * SW + hx + vmax;
        float mx = row[0];
        for (int dx = 1; dx <= vr; ++dx) mx = fmaxf(mx, fmaxf(row[-dx], row[dx]));
        hbuf[(size_t)t * SH * kVicTX + (size_t)hy * kVicTX + hx] = mx;
      }
      __syncthreads();
      // ---- vertical maximum for this thread's points, thresholds, accumulation ----
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (!okp[v]) continue;
        const int t = LAND ? (my_land[v] ? 0 : 1) : 0;
        const float* col = hbuf + (size_t)t * SH * kVicTX + (size_t)r * kVicTX + c + v;
        float mx = col[0];
        for (int dy = 1; dy <= 2 * vr; ++dy) mx = fmaxf(mx, col[(size_t)dy * kVicTX]);
        if (p.eq_mode) {
          acc[iv][v] = __fadd_rn(acc[iv][v], mx);
        } else {
          const float d = p.sign * mx;               // the extreme data value of the vicinity
          for (int t2 = 0; t2 < p.n_thr; ++t2) {
            const int k = iv * p.n_thr + t2;
            const float tv = truth_value_t<TM>(d, p.thr.t[t2]);
#pragma unroll
            for (int kk = 0; kk < kVicMaxOut; ++kk)
              if (kk == k) acc[kk][v] = __fadd_rn(acc[kk][v], tv);
          }
        }
      }
    }
  }
  if (__isnanf(nanacc)) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
  if (gy >= p.ny) return;
  const long long pt0 = plane * npts + (long long)gy * p.nx + x0 + c;
  const int n_t = p.eq_mode ? 1 : p.n_thr;
#pragma unroll
  for (int k = 0; k < kVicMaxOut; ++k) {
    if (k >= p.n_vic * n_t) break;
    const int iv = k / n_t, t2 = k - iv * n_t;
    float* o = p.out + (long long)iv * p.out_vic_stride + (long long)t2 * p.out_thr_stride + pt0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (x0 + c + v >= p.nx) continue;
      float res;
      if (cnt[v] <= 1) res = cnt[v] ? acc[k][v] : 0.f;
      else if (cnt[v] <= kVicRcpTable) {
        const double a = (double)acc[k][v], cc = (double)cnt[v], rc = rcp_tab[cnt[v]];
        const double q0 = a * rc;
        res = (float)fma(fma(-q0, cc, a), rc, q0);   // correctly rounded a / cc (threshold.py:707)
      } else res = (float)((double)acc[k][v] / (double)cnt[v]);
      o[v] = res;
    }
  }
  if (p.contrib) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (x0 + c + v < p.nx) p.contrib[pt0 + v] = cnt[v];
  }
}

template <int TM>
static int launch_vic(const VicParams& p, long long blocks, size_t smem, cudaStream_t st) {
  if (p.landmask) {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(vicinity_kernel<TM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vicinity_kernel<TM, true><<<(unsigned)blocks, kVicThreads, smem, st>>>(p);
  } else {
    NBH_CHECK_CUDA(cudaFuncSetAttribute(vicinity_kernel<TM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vicinity_kernel<TM, false><<<(unsigned)blocks, kVicThreads, smem, st>>>(p);
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int run_threshold_vicinity(const float* in, const uint8_t* mask_nd, const uint8_t* landmask, float* out,
                           int32_t* contrib, const NbhThreshold* thresholds, int n_thresholds,
                           const int32_t* vicinity_cells, int n_vicinity, long long n_collapse,
                           long long n_planes, int ny, int nx, int32_t* status, cudaStream_t st, int raw_op) {
  // raw_op: -1 thresholded; 0 / 1: maximum / minimum of the data itself (no thresholds)
  NBH_REQUIRE(in && out && status && vicinity_cells, "null pointer argument");
  NBH_REQUIRE(raw_op >= 0 || (thresholds && n_thresholds >= 1), "at least one threshold is needed");
  NBH_REQUIRE(n_vicinity >= 1, "at least one vicinity radius is needed");
  if (raw_op >= 0) n_thresholds = 1;
  NBH_REQUIRE(n_collapse >= 1 && n_planes >= 1 && ny >= 1 && nx >= 1, "empty input");
  int vmax = 0;
  for (int i = 0; i < n_vicinity; ++i) {
    NBH_REQUIRE(vicinity_cells[i] >= 0 && vicinity_cells[i] <= kVicMaxRadius,
                "vicinity radius of %d grid cells is outside [0, %d]", vicinity_cells[i], kVicMaxRadius);
    vmax = max(vmax, vicinity_cells[i]);
  }
  NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  const long long npts = (long long)ny * nx;
  const long long blocks = n_planes * ((ny + kVicTY - 1) / kVicTY) * ((nx + kVicTX - 1) / kVicTX);
  NBH_REQUIRE(blocks < (1LL << 31), "too many blocks (%lld); split the launch", blocks);
  const int nt = landmask ? 2 : 1;
  const size_t smem = ((size_t)nt * (kVicTY + 2 * vmax) * (kVicTX + 2 * vmax) + (size_t)nt * (kVicTY + 2 * vmax) * kVicTX) * 4;
  // group launches: thresholds sharing evaluation mode and monotonic direction, at most
  // kVicMaxOut (radius, threshold) pairs each; "==" goes one threshold at a time
  int t0 = 0;
  while (t0 < n_thresholds) {
    ThrDev first;
    memset(&first, 0, sizeof(first));
    if (raw_op < 0 && make_thr_dev(thresholds[t0], &first, mask_nd != nullptr)) return 1;
    const bool eq = raw_op < 0 && first.op == NBH_OP_EQ;
    const bool less = raw_op < 0 ? (first.op == NBH_OP_LT || first.op == NBH_OP_LE) : raw_op == 1;
    int v0 = 0;
    int n_t = 0;
    VicParams p;
    memset(&p, 0, sizeof(p));
    const int max_t = eq ? 1 : max(1, kVicMaxOut / min(n_vicinity, kVicMaxOut));
    if (raw_op >= 0) p.thr.t[n_t++] = first;
    for (int i = t0; raw_op < 0 && i < n_thresholds && n_t < max_t; ++i) {
      ThrDev d;
      if (make_thr_dev(thresholds[i], &d, mask_nd != nullptr)) return 1;
      const bool eq_i = d.op == NBH_OP_EQ, less_i = d.op == NBH_OP_LT || d.op == NBH_OP_LE;
      if (d.mode != first.mode || eq_i != eq || less_i != less) break;
      p.thr.t[n_t++] = d;
    }
    p.thr.n = n_t;
    while (v0 < n_vicinity) {
      const int n_v = min(n_vicinity - v0, max(1, kVicMaxOut / n_t));
      p.in = in; p.mask_nd = mask_nd; p.landmask = landmask; p.status = status;
      p.contrib = (t0 == 0 && v0 == 0) ? contrib : nullptr;
      p.n_collapse = n_collapse; p.n_planes = n_planes; p.ny = ny; p.nx = nx;
      p.n_vic = n_v; p.n_thr = n_t; p.vmax = vmax;
      for (int i = 0; i < n_v; ++i) p.cells[i] = vicinity_cells[v0 + i];
      p.out_thr_stride = n_planes * npts;
      p.out_vic_stride = (long long)n_thresholds * n_planes * npts;
      p.out = out + (long long)v0 * p.out_vic_stride + (long long)t0 * p.out_thr_stride;
      p.eq_mode = eq ? 1 : 0;
      p.sign = less ? -1.f : 1.f;
      int rc;
      switch (first.mode) {
        case 0: rc = launch_vic<0>(p, blocks, smem, st); break;
        case 1: rc = launch_vic<1>(p, blocks, smem, st); break;
        case 2: rc = launch_vic<2>(p, blocks, smem, st); break;
        default: rc = launch_vic<3>(p, blocks, smem, st); break;
      }
      if (rc) return rc;
      v0 += n_v;
    }
    t0 += n_t;
  }
  return 0;
}

}  // namespace nbh
