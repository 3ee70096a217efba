// C ABI of libimprover_b200.so (declared in include/improver_b200.h).
#include <stdarg.h>

#include <cmath>

#include "nbh_common.cuh"

namespace nbh {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static bool g_prof_on = false;
static cudaEvent_t g_prof_e0 = nullptr, g_prof_e1 = nullptr;
static bool g_prof_valid = false;
static const char* g_prof_name = "";

ProfileScope::ProfileScope(cudaStream_t s, const char* name) : st(s), on(g_prof_on) {
  if (!on) return;
  g_prof_name = name;
  if (!g_prof_e0) { cudaEventCreate(&g_prof_e0); cudaEventCreate(&g_prof_e1); }
  cudaEventRecord(g_prof_e0, st);
}
ProfileScope::~ProfileScope() {
  if (!on) return;
  cudaEventRecord(g_prof_e1, st);
  g_prof_valid = true;
}

// Host restatement of the device truth value (same float32 operation order)
static float host_truth(const ThrDev& p, float d) {
  if (p.mode == 1) {
    bool b;
    switch (p.op) {
      case NBH_OP_GT: b = d > p.thr; break;
      case NBH_OP_GE: b = d >= p.thr; break;
      case NBH_OP_LT: b = d < p.thr; break;
      case NBH_OP_LE: b = d <= p.thr; break;
      default: b = d == p.thr; break;
    }
    return b ? 1.f : 0.f;
  }
  if (p.mode == 3) {
    const bool below = d < p.thr;
    const double a = below ? ((double)d - p.lo64) : (double)(d - p.thr);
    const double den = below ? (double)p.den_lo : (double)p.den_hi;
    const double off = below ? 0.0 : 0.5;
    double v = a * 0.5 / den + off;
    v = std::fmin(std::fmax(v, off), off + 0.5);
    if ((p.op == NBH_OP_LT || p.op == NBH_OP_LE)) v = 1.0 - v;
    return (float)v;
  }
  const float dc = std::fmin(std::fmax(d, p.lo), p.hi);
  const bool below = d < p.thr;
  const float base = below ? p.lo : p.thr;
  const float den2 = below ? p.den2_lo : p.den2_hi;
  const float rcph = below ? p.rcph_lo : p.rcph_hi;
  const float off = below ? 0.f : 0.5f;
  const float a = dc - base;
  const float q0 = a * rcph;
  const float rem = std::fmaf(-den2, q0, a);
  float v = std::fmaf(rem, rcph, q0) + off;
  if ((p.op == NBH_OP_LT || p.op == NBH_OP_LE)) v = 1.0f - v;
  return v;
}

static uint32_t f2ord(float f) {
  uint32_t b; memcpy(&b, &f, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static float ord2f(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  float f; memcpy(&f, &b, 4);
  return f;
}

// [one_lo, one_hi]: the interval of data values whose truth value is exactly 1
// (the membership function is monotone, so the set is an interval ending at +-inf).
static void find_one_interval(ThrDev* d) {
  const float inf = INFINITY;
  if (d->mode == 1 && d->op == NBH_OP_EQ) { d->one_lo = d->one_hi = d->thr; return; }
  const bool decreasing = (d->mode != 1) ? ((d->op == NBH_OP_LT || d->op == NBH_OP_LE)) : ((d->op == NBH_OP_LT || d->op == NBH_OP_LE));
  // search over the total order of floats between -inf and +inf
  uint32_t lo = f2ord(-inf), hi = f2ord(inf);
  if (!decreasing) {
    // smallest d with truth == 1
    if (host_truth(*d, inf) != 1.f) { d->one_lo = inf; d->one_hi = -inf; return; }
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (host_truth(*d, ord2f(mid)) == 1.f) hi = mid; else lo = mid + 1;
    }
    d->one_lo = ord2f(lo); d->one_hi = inf;
  } else {
    // largest d with truth == 1
    if (host_truth(*d, -inf) != 1.f) { d->one_lo = inf; d->one_hi = -inf; return; }
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo + 1) / 2;
      if (host_truth(*d, ord2f(mid)) == 1.f) lo = mid; else hi = mid - 1;
    }
    d->one_lo = -inf; d->one_hi = ord2f(lo);
  }
}

// Host preparation of one threshold, mirroring the dtype handling of
// improver/threshold.py:434 (threshold cast to the float32 data dtype) and
// numpy's weak-scalar promotion of the python-float fuzzy bounds.
int make_thr_dev(const NbhThreshold& h, ThrDev* d, bool ma_arith) {
  memset(d, 0, sizeof(*d));
  NBH_REQUIRE(h.op >= NBH_OP_GT && h.op <= NBH_OP_EQ, "unknown comparison operator code %d", h.op);
  d->op = h.op;
  d->thr = (float)h.value;
  if (h.lo == h.hi) {           // threshold.py:438-439
    d->mode = 1;
    find_one_interval(d);
    return 0;
  }
  NBH_REQUIRE(h.lo <= h.value && h.value <= h.hi, "Threshold must be within bounds: !( %g <= %g <= %g )",
              h.lo, h.value, h.hi);
  d->mode = ma_arith ? 3 : 2;
  d->lo64 = h.lo;
  d->lo = (float)h.lo;
  d->hi = (float)h.hi;
  d->den_lo = d->thr - d->lo;
  d->den_hi = d->hi - d->thr;
  // rescale() raises on a zero input range (rescale.py:50-53)
  NBH_REQUIRE(d->den_lo != 0.f && d->den_hi != 0.f,
              "Cannot rescale a zero input range (fuzzy bound equals the threshold in float32)");
  d->den2_lo = 2.0f * d->den_lo;
  d->den2_hi = 2.0f * d->den_hi;
  d->rcph_lo = (float)(0.5 / (double)d->den_lo);
  d->rcph_hi = (float)(0.5 / (double)d->den_hi);
  const bool invert = (h.op == NBH_OP_LT || h.op == NBH_OP_LE);
  d->fs_lo = invert ? -d->rcph_lo : d->rcph_lo;
  d->fs_hi = invert ? -d->rcph_hi : d->rcph_hi;
  // symmetric bounds (fuzzy_factor): one slope serves both sides of the threshold
  d->fast_kind = 2;
  const double s_lo = 0.5 / (double)d->den_lo, s_hi = 0.5 / (double)d->den_hi;
  if (std::fabs(s_lo - s_hi) <= 2e-7 * std::fmax(s_lo, s_hi)) {
    const double s1 = 0.5 * (s_lo + s_hi) * (invert ? -1.0 : 1.0);
    d->fs1 = (float)s1;
    d->fc1 = (float)(0.5 - (double)d->thr * (double)d->fs1);
    const double mag = std::fmax(std::fabs((double)d->hi * s1), std::fabs((double)d->lo * s1));
    d->fast_kind = (std::fmax(mag, std::fabs((double)d->fc1)) <= 8.0) ? 5 : 4;
  }
  find_one_interval(d);
  return 0;
}

int run_square(const NbhDesc*, const float*, float*, uint8_t*, int32_t*, void*, size_t, cudaStream_t);
size_t square_workspace_bytes(const NbhDesc&);
int validate_desc(const NbhDesc*);
int run_circular(const NbhDesc*, const float*, float*, uint8_t*, int32_t*, void*, size_t, cudaStream_t);
size_t circular_workspace_bytes(const NbhDesc&);
int run_percentiles(const float*, float*, const float*, int, long long, int, int, int, int32_t*, void*,
                    size_t, cudaStream_t);
size_t percentiles_workspace_bytes(long long, int, int, int);
int run_threshold_vicinity(const float* in, const uint8_t* mask_nd, const uint8_t* landmask, float* out,
                           int32_t* contrib, const NbhThreshold* thresholds, int n_thresholds,
                           const int32_t* vicinity_cells, int n_vicinity, long long n_collapse,
                           long long n_planes, int ny, int nx, int32_t* status, cudaStream_t st, int raw_op = -1);
int run_threshold(const float*, const uint8_t*, float*, int32_t*, const NbhThreshold*, int, long long,
                  long long, int32_t*, cudaStream_t);
int run_slice_stats(const NbhDesc*, const float*, NbhSliceStats*, cudaStream_t);
int run_collapse_mean(const float*, float*, long long, long long, cudaStream_t);
int run_linear_map(const float*, float*, double, double, long long, cudaStream_t);
int run_boxsum_f64(const double*, double*, long long, int, int, int, int, int, cudaStream_t);
int run_zone_collapse(const float*, const float*, float*, long long, int, long long, cudaStream_t);
int run_halo_pack(const float*, float*, float*, float*, long long, int, int, int, int, int, cudaStream_t);
int run_halo_unpack(float*, const float*, const float*, long long, int, int, int, int, cudaStream_t);
size_t halo_exchange_workspace_bytes(long long, int, int);
int run_halo_exchange(void*, int, int, const float*, float*, long long, int, int, int, void*, size_t, cudaStream_t);

typedef int (*nbhood_fn)(const NbhDesc*, const float*, float*, uint8_t*, int32_t*, void*, size_t, cudaStream_t);

// OR the status words of a sub-launch into the caller's (status[0] is a flag, status[1] a count)
__global__ void merge_status_kernel(int32_t* dst, const int32_t* src) {
  if (threadIdx.x == 0) { dst[0] |= src[0]; dst[1] += src[1]; }
}

bool square_native_land_sea(const NbhDesc&);
bool circular_native_land_sea(const NbhDesc&);

__global__ void invert_mask_kernel(const uint8_t* m, uint8_t* inv, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[i] = m[i] == 0 ? 1 : 0;
}
// out = land point ? land result : sea result (filled(0) sum of the two re-masked passes,
// cli/nbhood_land_and_sea.py:176-184); nothing is masked in the combined field
__global__ void land_sea_combine_kernel(float* out, const float* sea, uint8_t* out_mask, const uint8_t* land,
                                        long long n_out, long long per_slice) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out;
       i += (long long)gridDim.x * blockDim.x) {
    if (land[i % per_slice] == 0) out[i] = sea[i];
    if (out_mask) out_mask[i] = 0;
  }
}

static size_t land_sea_extra_bytes(const NbhDesc& d) {
  const size_t per = (size_t)d.ny * d.nx;
  const size_t T = d.n_thresholds > 0 ? d.n_thresholds : 1;
  return (per + 255) / 256 * 256 + ((size_t)d.n_planes * T * per * 4 + 255) / 256 * 256 + 256;
}

// Generic land + sea launch for the kernels without a native NBH_LAND_SEA path: two masked
// launches (land mask, inverted mask) and a select.
static int run_land_sea_composite(nbhood_fn fn, const NbhDesc* d, const float* in, float* out, uint8_t* out_mask,
                                  int32_t* status, void* ws, size_t ws_bytes, cudaStream_t st) {
  NBH_REQUIRE(d->mask2d != nullptr, "NBH_LAND_SEA needs the land mask as mask2d");
  NBH_REQUIRE(d->ny > 0 && d->nx > 0 && d->n_planes > 0, "empty grid");
  const size_t extra = land_sea_extra_bytes(*d);
  NBH_REQUIRE(ws_bytes >= extra, "workspace too small");
  const size_t per = (size_t)d->ny * d->nx;
  const size_t T = d->n_thresholds > 0 ? d->n_thresholds : 1;
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  uint8_t* inv = base;
  float* sea = reinterpret_cast<float*>(base + (per + 255) / 256 * 256);
  int32_t* sub_status = reinterpret_cast<int32_t*>(base + extra - 256);
  void* sub_ws = base + extra;
  const size_t sub_bytes = ws_bytes - extra;
  invert_mask_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(d->mask2d, inv, (long long)per);
  NbhDesc sub = *d;
  sub.flags &= ~NBH_LAND_SEA;
  int rc = fn(&sub, in, out, nullptr, status, sub_ws, sub_bytes, st);
  if (rc) return rc;
  sub.mask2d = inv;
  rc = fn(&sub, in, sea, nullptr, sub_status, sub_ws, sub_bytes, st);
  if (rc) return rc;
  merge_status_kernel<<<1, 32, 0, st>>>(status, sub_status);
  const long long n_out = (long long)d->n_planes * T * per;
  land_sea_combine_kernel<<<(unsigned)min((n_out + 255) / 256, (long long)sm_count() * 16), 256, 0, st>>>(
      out, sea, out_mask, d->mask2d, n_out, (long long)per);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int dispatch_square(const NbhDesc* d, const float* in, float* out, uint8_t* out_mask, int32_t* status,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
  if ((d->flags & NBH_LAND_SEA) && !square_native_land_sea(*d))
    return run_land_sea_composite(run_square, d, in, out, out_mask, status, ws, ws_bytes, st);
  return run_square(d, in, out, out_mask, status, ws, ws_bytes, st);
}
static int dispatch_circular(const NbhDesc* d, const float* in, float* out, uint8_t* out_mask, int32_t* status,
                             void* ws, size_t ws_bytes, cudaStream_t st) {
  if ((d->flags & NBH_LAND_SEA) && !circular_native_land_sea(*d))
    return run_land_sea_composite(run_circular, d, in, out, out_mask, status, ws, ws_bytes, st);
  return run_circular(d, in, out, out_mask, status, ws, ws_bytes, st);
}

// Lead-time dependent radii (nbhood.py:438-455: one radius per slice): consecutive planes with
// the same radius form one launch.  The sub-launches share the workspace (stream order) and
// report through a private status pair at its end.
static int run_per_plane_radii(nbhood_fn fn, const NbhDesc* d, const float* in, float* out, uint8_t* out_mask,
                               int32_t* status, void* ws, size_t ws_bytes, cudaStream_t st) {
  NBH_REQUIRE(d->n_planes > 0 && d->ny > 0 && d->nx > 0, "empty grid");
  NBH_REQUIRE(ws_bytes >= 256, "workspace too small");
  NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  int32_t* sub_status = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(ws) + (ws_bytes - 256) / 256 * 256);
  const size_t sub_ws = (ws_bytes - 256) / 256 * 256;
  const long long out_plane = (long long)(d->n_thresholds > 0 ? d->n_thresholds : 1) * d->ny * d->nx;
  for (long long p0 = 0; p0 < d->n_planes;) {
    long long p1 = p0 + 1;
    while (p1 < d->n_planes && d->plane_cells[p1] == d->plane_cells[p0]) ++p1;
    NbhDesc sub = *d;
    sub.plane_cells = nullptr;
    sub.cells = d->plane_cells[p0];
    sub.n_planes = p1 - p0;
    if (sub.ext_stats) sub.ext_stats += p0 * d->n_members * (d->n_thresholds > 0 ? d->n_thresholds : 1);
    if (sub.mask_nd) sub.mask_nd += p0 * d->plane_stride;
    const int rc = fn(&sub, in + p0 * d->plane_stride, out + p0 * out_plane,
                      out_mask ? out_mask + p0 * out_plane : nullptr, sub_status, ws, sub_ws, st);
    if (rc) return rc;
    merge_status_kernel<<<1, 32, 0, st>>>(status, sub_status);
    p0 = p1;
  }
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nbh

extern "C" {

int nbh_version(void) { return NBH_ABI_VERSION; }

const char* nbh_last_error(void) { return nbh::g_err; }

int nbh_profile_enable(int on) {
  nbh::g_prof_on = on != 0;
  nbh::g_prof_valid = false;
  return 0;
}

int nbh_profile_last_ms(float* ms) {
  NBH_REQUIRE(ms != nullptr, "null pointer argument");
  NBH_REQUIRE(nbh::g_prof_valid, "no profiled launch recorded");
  NBH_CHECK_CUDA(cudaEventSynchronize(nbh::g_prof_e1));
  NBH_CHECK_CUDA(cudaEventElapsedTime(ms, nbh::g_prof_e0, nbh::g_prof_e1));
  return 0;
}

const char* nbh_profile_last_kernel(void) { return nbh::g_prof_name; }

int nbh_device_info(int32_t* sm_count, int64_t* smem_optin_bytes) {
  int dev = 0, n = 0, s = 0;
  NBH_CHECK_CUDA(cudaGetDevice(&dev));
  NBH_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  NBH_CHECK_CUDA(cudaDeviceGetAttribute(&s, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (sm_count) *sm_count = n;
  if (smem_optin_bytes) *smem_optin_bytes = s;
  return 0;
}

size_t nbh_workspace_bytes(const NbhDesc* desc) {
  if (!desc) return 0;
  if (desc->plane_cells) {
    // the largest request of any radius group, plus the private status words of the sub-launches
    size_t best = 0;
    for (long long p = 0; p < desc->n_planes; ++p) {
      if (p > 0 && desc->plane_cells[p] == desc->plane_cells[p - 1]) continue;
      NbhDesc sub = *desc;
      sub.plane_cells = nullptr;
      sub.cells = desc->plane_cells[p];
      const size_t b = nbh_workspace_bytes(&sub);
      best = b > best ? b : best;
    }
    return (best + 255) / 256 * 256 + 512;
  }
  size_t a = nbh::square_workspace_bytes(*desc);
  size_t b = nbh::circular_workspace_bytes(*desc);
  size_t need = a > b ? a : b;
  if (desc->flags & NBH_LAND_SEA) need += nbh::land_sea_extra_bytes(*desc);
  return need;
}

int nbh_square_f32(const NbhDesc* desc, const float* in, float* out, uint8_t* out_mask, int32_t* status,
                   void* workspace, size_t workspace_bytes, void* stream) {
  NBH_REQUIRE(desc != nullptr, "null descriptor");
  if (desc->plane_cells)
    return nbh::run_per_plane_radii(nbh::dispatch_square, desc, in, out, out_mask, status, workspace, workspace_bytes,
                                    static_cast<cudaStream_t>(stream));
  return nbh::dispatch_square(desc, in, out, out_mask, status, workspace, workspace_bytes,
                              static_cast<cudaStream_t>(stream));
}

int nbh_circular_f32(const NbhDesc* desc, const float* in, float* out, uint8_t* out_mask, int32_t* status,
                     void* workspace, size_t workspace_bytes, void* stream) {
  NBH_REQUIRE(desc != nullptr, "null descriptor");
  if (desc->plane_cells)
    return nbh::run_per_plane_radii(nbh::dispatch_circular, desc, in, out, out_mask, status, workspace, workspace_bytes,
                                    static_cast<cudaStream_t>(stream));
  return nbh::dispatch_circular(desc, in, out, out_mask, status, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}

int nbh_collapse_mean_f32(const float* in, float* out, int64_t n_collapse, int64_t n_points, void* stream) {
  return nbh::run_collapse_mean(in, out, n_collapse, n_points, static_cast<cudaStream_t>(stream));
}

int nbh_boxsum_f64(const double* in, double* out, int64_t n_slices, int32_t rows, int32_t cols, int32_t box_y,
                   int32_t box_x, int32_t cumsum, void* stream) {
  return nbh::run_boxsum_f64(in, out, n_slices, rows, cols, box_y, box_x, cumsum, static_cast<cudaStream_t>(stream));
}

int nbh_linear_map_f32(const float* in, float* out, double a, double b, int64_t n, void* stream) {
  return nbh::run_linear_map(in, out, a, b, n, static_cast<cudaStream_t>(stream));
}

int nbh_zone_collapse_f32(const float* values, const float* weights, float* out, int64_t n_lead, int32_t n_zones,
                          int64_t n_points, void* stream) {
  return nbh::run_zone_collapse(values, weights, out, n_lead, n_zones, n_points, static_cast<cudaStream_t>(stream));
}

int nbh_slice_stats_f32(const NbhDesc* desc, const float* in, NbhSliceStats* stats, void* stream) {
  return nbh::run_slice_stats(desc, in, stats, static_cast<cudaStream_t>(stream));
}

int nbh_halo_pack(const float* band, float* ext, float* send_up, float* send_down, int64_t n_slices, int32_t rows,
                  int32_t nx, int32_t cells, int32_t halo_top, int32_t halo_bottom, void* stream) {
  return nbh::run_halo_pack(band, ext, send_up, send_down, n_slices, rows, nx, cells, halo_top, halo_bottom,
                            static_cast<cudaStream_t>(stream));
}

int nbh_halo_unpack(float* ext, const float* recv_up, const float* recv_down, int64_t n_slices, int32_t rows,
                    int32_t nx, int32_t halo_top, int32_t halo_bottom, void* stream) {
  return nbh::run_halo_unpack(ext, recv_up, recv_down, n_slices, rows, nx, halo_top, halo_bottom,
                              static_cast<cudaStream_t>(stream));
}

size_t nbh_halo_exchange_workspace_bytes(int64_t n_slices, int32_t nx, int32_t cells) {
  return nbh::halo_exchange_workspace_bytes(n_slices, nx, cells);
}

int nbh_halo_exchange(void* nccl_comm, int32_t rank, int32_t n_ranks, const float* band, float* ext,
                      int64_t n_slices, int32_t rows, int32_t nx, int32_t cells, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return nbh::run_halo_exchange(nccl_comm, rank, n_ranks, band, ext, n_slices, rows, nx, cells, workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream));
}

int nbh_percentiles_f32(const float* in, float* out, const float* pct, int32_t n_pct, int64_t n_slices,
                        int32_t ny, int32_t nx, int32_t cells, int32_t* status, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return nbh::run_percentiles(in, out, pct, n_pct, n_slices, ny, nx, cells, status, workspace,
                              workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t nbh_percentiles_workspace_bytes(int64_t n_slices, int32_t ny, int32_t nx, int32_t cells) {
  return nbh::percentiles_workspace_bytes(n_slices, ny, nx, cells);
}

int nbh_threshold_f32(const float* in, const uint8_t* mask_nd, float* out, int32_t* contrib,
                      const NbhThreshold* thresholds, int32_t n_thresholds, int64_t n_collapse,
                      int64_t n_points, int32_t* status, void* stream) {
  return nbh::run_threshold(in, mask_nd, out, contrib, thresholds, n_thresholds, n_collapse, n_points,
                            status, static_cast<cudaStream_t>(stream));
}

int nbh_threshold_vicinity_f32(const float* in, const uint8_t* mask_nd, const uint8_t* landmask, float* out,
                               int32_t* contrib, const NbhThreshold* thresholds, int32_t n_thresholds,
                               const int32_t* vicinity_cells, int32_t n_vicinity, int64_t n_collapse,
                               int64_t n_planes, int32_t ny, int32_t nx, int32_t* status, void* stream) {
  return nbh::run_threshold_vicinity(in, mask_nd, landmask, out, contrib, thresholds, n_thresholds,
                                     vicinity_cells, n_vicinity, n_collapse, n_planes, ny, nx, status,
                                     static_cast<cudaStream_t>(stream));
}

int nbh_vicinity_f32(const float* in, const uint8_t* mask_nd, const uint8_t* landmask, float* out,
                     const int32_t* vicinity_cells, int32_t n_vicinity, int32_t op, int64_t n_planes,
                     int32_t ny, int32_t nx, int32_t* status, void* stream) {
  NBH_REQUIRE(op == 0 || op == 1, "op must be 0 (maximum) or 1 (minimum)");
  return nbh::run_threshold_vicinity(in, mask_nd, landmask, out, nullptr, nullptr, 0, vicinity_cells, n_vicinity,
                                     1, n_planes, ny, nx, status, static_cast<cudaStream_t>(stream), op);
}

int nbh_recursive_filter_f32(const float* in, const float* coeff_x, const float* coeff_y, const uint8_t* mask,
                             int32_t mask_per_slice, int32_t mask_zeros, float* out, int64_t n_slices,
                             int32_t ny, int32_t nx, int32_t pad, int32_t iterations, void* workspace,
                             size_t workspace_bytes, void* stream) {
  return nbh::run_recursive_filter(in, coeff_x, coeff_y, mask, mask_per_slice, mask_zeros, out, n_slices, ny, nx,
                                   pad, iterations, workspace, workspace_bytes,
                                   static_cast<cudaStream_t>(stream));
}

size_t nbh_recursive_filter_workspace_bytes(int64_t n_slices, int32_t ny, int32_t nx, int32_t pad,
                                            int32_t coeff_per_slice) {
  return nbh::recursive_filter_workspace_bytes(n_slices, ny, nx, pad, coeff_per_slice);
}

int nbh_recurse_f32(float* grid, const float* coeffs, int64_t n_slices, int32_t ny, int32_t nx, int32_t axis,
                    int32_t dirs, void* stream) {
  return nbh::run_recurse(grid, coeffs, n_slices, ny, nx, axis, dirs, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
