// Neighbourhood percentiles: replaces the hot loop of
// GeneratePercentilesFromANeighbourhood.pad_and_unpad_cube
// (improver/nbhood/nbhood.py:649-667): np.pad(mode="mean", stat_length=r) of
// each 2-D slice, circular (2r+1)^2 windows, np.percentile(method="linear").
//
// One warp per grid point.  The N window values are held as order-preserving
// uint32 keys in registers (KPL per lane); each requested order statistic is
// found by a 32-step most-significant-bit-first bisection that counts keys
// below the candidate with a warp reduction, so no sorting network and no
// shared-memory traffic is needed in the selection itself.
#include "nbh_common.cuh"

namespace nbh {

constexpr int kMaxPct = 32;

struct PctParams {
  const float* in;
  float* out;
  const float* padv;      // per slice: top[nx], bot[nx], left[ny+2r], right[ny+2r]
  const int* offs;        // [N] packed (dy+r) << 16 | (dx+r) of kernel cells
  long long n_slices;
  int ny, nx, r, nb, N, n_pct;
  int TW, TH, tiles_x, tiles_y;
  int k_lo[kMaxPct];      // floor((N-1) * q)
  float gamma[kMaxPct];   // fractional part (float32, numpy arithmetic)
};

__device__ __forceinline__ unsigned f2key(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Column / row pad means (np.pad mode="mean": axis 0 first, then axis 1 on
// the already padded rows).  One block per slice.
__global__ void pct_pad_kernel(const float* in, float* padv, int ny, int nx, int r) {
  const long long s = blockIdx.x;
  const float* a = in + s * (long long)ny * nx;
  float* top = padv + s * (2LL * nx + 2LL * (ny + 2 * r));
  float* bot = top + nx;
  float* left = bot + nx;
  float* right = left + (ny + 2 * r);
  const int sy = min(r, ny), sx = min(r, nx);
  for (int x = threadIdx.x; x < nx; x += blockDim.x) {
    float t = 0.f, b = 0.f;
    for (int i = 0; i < sy; ++i) { t += a[(long long)i * nx + x]; b += a[(long long)(ny - sy + i) * nx + x]; }
    top[x] = t / (float)sy;
    bot[x] = b / (float)sy;
  }
  __syncthreads();
  for (int yp = threadIdx.x; yp < ny + 2 * r; yp += blockDim.x) {
    const int y = yp - r;
    const float* row = (y < 0) ? top : (y >= ny ? bot : a + (long long)y * nx);
    double l = 0.0, rr = 0.0;
    for (int j = 0; j < sx; ++j) { l += row[j]; rr += row[nx - sx + j]; }
    left[yp] = (float)(l / sx);
    right[yp] = (float)(rr / sx);
  }
}

// U order statistics of the warp's keys at once (MSB-first bisection), then
// the float32 linear interpolation of numpy's percentile.
template <int KPL, int U>
__device__ __forceinline__ void select_group(const unsigned (&key)[KPL], const PctParams& p, int q0, int lane,
                                             long long obase) {
  int kk[U];
  unsigned prefix[U];
  int lo_cnt[U], hi_cnt[U];          // keys below the bucket / below its upper end: the k-th smallest is in
                                     // the bucket [prefix, prefix + 2^(bit+1)) while lo_cnt <= k < hi_cnt
#pragma unroll
  for (int u = 0; u < U; ++u) { kk[u] = p.k_lo[q0 + u]; prefix[u] = 0; lo_cnt[u] = 0; hi_cnt[u] = p.N; }
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    // stop as soon as every bucket holds a single key (log2(N) + the shared exponent bits instead
    // of 32 steps on typical fields; ties keep bisecting to the last bit)
    bool open_bucket = false;
#pragma unroll
    for (int u = 0; u < U; ++u) open_bucket = open_bucket || (hi_cnt[u] - lo_cnt[u] > 1);
    if (!open_bucket) break;
    unsigned cand[U];
    int cnt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { cand[u] = prefix[u] | (1u << bit); cnt[u] = 0; }
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
#pragma unroll
      for (int u = 0; u < U; ++u) cnt[u] += key[j] < cand[u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      cnt[u] = __reduce_add_sync(0xffffffffu, cnt[u]);
      if (cnt[u] <= kk[u]) { prefix[u] = cand[u]; lo_cnt[u] = cnt[u]; }
      else hi_cnt[u] = cnt[u];
    }
  }
  // the k-th smallest key is the smallest key not below the prefix (exactly lo_cnt == k keys are
  // below it once the bucket is down to one key; after 32 steps the prefix is the key itself)
#pragma unroll
  for (int u = 0; u < U; ++u) {
    unsigned sel = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < KPL; ++j) if (key[j] >= prefix[u]) sel = min(sel, key[j]);
    prefix[u] = __reduce_min_sync(0xffffffffu, sel);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int q = q0 + u;
    const int k = kk[u];
    // (k+1)-th: same value if duplicated, else the smallest key above
    int cle = 0;
    unsigned nxt = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      cle += key[j] <= prefix[u];
      if (key[j] > prefix[u]) nxt = min(nxt, key[j]);
    }
    cle = __reduce_add_sync(0xffffffffu, cle);
    nxt = __reduce_min_sync(0xffffffffu, nxt);
    const float lo = key2f(prefix[u]);
    const float hi = (k + 1 >= p.N || cle >= k + 2) ? lo : key2f(nxt);
    // numpy _lerp (float32): a + (b-a)*t, and b - (b-a)*(1-t) where t >= 0.5
    const float t = p.gamma[q];
    const float diff = __fsub_rn(hi, lo);
    float res = __fadd_rn(lo, __fmul_rn(diff, t));
    if (t >= 0.5f) res = __fsub_rn(hi, __fmul_rn(diff, __fsub_rn(1.0f, t)));
    if (lane == 0) p.out[obase + (long long)q * p.ny * p.nx] = res;
  }
}

// ---------------------------------------------------------------------------
// Sort-based selection (windows of up to 32 * 16 keys): the warp sorts its N keys once — every
// lane a sorted block of KPL keys, blocks ascending across lanes — and every requested order
// statistic is then a plain lookup, so the cost no longer grows with the number of percentiles
// (the reference's default asks for 15, improver/constants.py DEFAULT_PERCENTILES) and is about a
// third of the bit-by-bit bisection's for five.
//
// Network: Batcher's bitonic sort with each lane as a "processor" holding a sorted block
// (merge-split): local sort, then log2(32) * (log2(32) + 1) / 2 = 15 rounds in which a lane
// exchanges its block with lane ^ j, keeps the lower (or upper) half of the union —
// min(A_i, B_{K-1-i}) resp. max — and re-sorts its half.  Blocks of any length K are handled
// with the arbitrary-n bitonic merge (K. Batcher / H. W. Lang): compare i with i + m for the
// largest power of two m < n, then recurse on [0, m) and [m, n); it sorts V-shaped sequences
// (descending then ascending).  The upper half is V-shaped; the lower half is its mirror image
// (ascending then descending) and is merged with the comparator reversed, then read backwards.
// ---------------------------------------------------------------------------
template <bool ASC> __device__ __forceinline__ void pct_ce(unsigned& a, unsigned& b) {
  const unsigned lo = min(a, b), hi = max(a, b);
  a = ASC ? lo : hi;
  b = ASC ? hi : lo;
}
__device__ __forceinline__ void pct_ce_dir(unsigned& a, unsigned& b, bool asc) {
  const unsigned lo = min(a, b), hi = max(a, b);
  a = asc ? lo : hi;
  b = asc ? hi : lo;
}
__host__ __device__ constexpr int pct_pow2_below(int n) { int m = 1; while (m * 2 < n) m *= 2; return m; }

template <int LO, int N, bool ASC, int KPL> struct PctMerge {
  static __device__ __forceinline__ void run(unsigned (&k)[KPL]) {
    constexpr int M = pct_pow2_below(N);
#pragma unroll
    for (int i = LO; i < LO + N - M; ++i) pct_ce<ASC>(k[i], k[i + M]);
    PctMerge<LO, M, ASC, KPL>::run(k);
    PctMerge<LO + M, N - M, ASC, KPL>::run(k);
  }
  static __device__ __forceinline__ void run_dir(unsigned (&k)[KPL], bool asc) {
    constexpr int M = pct_pow2_below(N);
#pragma unroll
    for (int i = LO; i < LO + N - M; ++i) pct_ce_dir(k[i], k[i + M], asc);
    PctMerge<LO, M, ASC, KPL>::run_dir(k, asc);
    PctMerge<LO + M, N - M, ASC, KPL>::run_dir(k, asc);
  }
};
template <int LO, bool ASC, int KPL> struct PctMerge<LO, 1, ASC, KPL> {
  static __device__ __forceinline__ void run(unsigned (&)[KPL]) {}
  static __device__ __forceinline__ void run_dir(unsigned (&)[KPL], bool) {}
};
template <int LO, bool ASC, int KPL> struct PctMerge<LO, 0, ASC, KPL> {
  static __device__ __forceinline__ void run(unsigned (&)[KPL]) {}
  static __device__ __forceinline__ void run_dir(unsigned (&)[KPL], bool) {}
};
template <int LO, int N, bool ASC, int KPL> struct PctSort {
  static __device__ __forceinline__ void run(unsigned (&k)[KPL]) {
    constexpr int M = N / 2;
    PctSort<LO, M, !ASC, KPL>::run(k);
    PctSort<LO + M, N - M, ASC, KPL>::run(k);
    PctMerge<LO, N, ASC, KPL>::run(k);
  }
};
template <int LO, bool ASC, int KPL> struct PctSort<LO, 1, ASC, KPL> {
  static __device__ __forceinline__ void run(unsigned (&)[KPL]) {}
};
template <int LO, bool ASC, int KPL> struct PctSort<LO, 0, ASC, KPL> {
  static __device__ __forceinline__ void run(unsigned (&)[KPL]) {}
};

template <int KPL>
__device__ __forceinline__ void warp_sort_blocks(unsigned (&key)[KPL], int lane) {
  PctSort<0, KPL, true, KPL>::run(key);
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const bool lower = ((lane & j) == 0) == ((lane & k) == 0);
      unsigned other[KPL];
#pragma unroll
      for (int i = 0; i < KPL; ++i) other[i] = __shfl_xor_sync(0xffffffffu, key[KPL - 1 - i], j);
#pragma unroll
      for (int i = 0; i < KPL; ++i) key[i] = lower ? min(key[i], other[i]) : max(key[i], other[i]);
      // upper half: V-shaped, merged ascending; lower half: its mirror image, merged descending
      PctMerge<0, KPL, true, KPL>::run_dir(key, !lower);
      if (KPL > 1) {
#pragma unroll
        for (int i = 0; i < KPL / 2; ++i) {
          const unsigned a = key[i], b = key[KPL - 1 - i];
          key[i] = lower ? b : a;
          key[KPL - 1 - i] = lower ? a : b;
        }
      }
    }
  }
}

template <int KPL, bool SORT>
__global__ void __launch_bounds__(256)
percentile_kernel(const PctParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int r = p.r;
  const int SW = p.TW + 2 * r, SH = p.TH + 2 * r;
  unsigned* tile = reinterpret_cast<unsigned*>(smem_raw);      // [SH][SW] keys of the padded field
  int* offs = reinterpret_cast<int*>(tile + (size_t)SH * SW);  // [32*KPL] tile-relative offsets
  unsigned* sorted = reinterpret_cast<unsigned*>(offs + 32 * KPL) + (threadIdx.x >> 5) * (32 * KPL);   // SORT: per warp

  long long bid = blockIdx.x;
  const int tx = (int)(bid % p.tiles_x); bid /= p.tiles_x;
  const int ty = (int)(bid % p.tiles_y); bid /= p.tiles_y;
  const long long s = bid;
  const int x0 = tx * p.TW, y0 = ty * p.TH;
  const float* a = p.in + s * (long long)p.ny * p.nx;
  const float* top = p.padv + s * (2LL * p.nx + 2LL * (p.ny + 2 * r));
  const float* bot = top + p.nx;
  const float* left = bot + p.nx;
  const float* right = left + (p.ny + 2 * r);

  for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) {
    const int ly = i / SW, lx = i - ly * SW;
    const int y = y0 + ly - r, x = x0 + lx - r;
    float v;
    if (x < 0) v = left[min(max(y, -r), p.ny + r - 1) + r];
    else if (x >= p.nx) v = right[min(max(y, -r), p.ny + r - 1) + r];
    else if (y < 0) v = top[x];
    else if (y >= p.ny) v = bot[x];
    else v = a[(long long)y * p.nx + x];
    tile[i] = f2key(v);
  }
  for (int i = threadIdx.x; i < 32 * KPL; i += blockDim.x) {
    if (i < p.N) {
      const int o = p.offs[i];
      offs[i] = (o >> 16) * SW + (o & 0xffff);
    } else {
      offs[i] = -1;
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int my_off[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) my_off[j] = offs[j * 32 + lane];

  for (int pt = warp; pt < p.TW * p.TH; pt += nwarps) {
    const int ly = pt / p.TW, lx = pt - ly * p.TW;
    const int y = y0 + ly, x = x0 + lx;
    if (y >= p.ny || x >= p.nx) continue;
    const int base = ly * SW + lx;
    unsigned key[KPL];
#pragma unroll
    for (int j = 0; j < KPL; ++j) key[j] = my_off[j] >= 0 ? tile[base + my_off[j]] : 0xffffffffu;

    const long long obase = (s * p.n_pct * (long long)p.ny + y) * p.nx + x;
    if (SORT) {
      warp_sort_blocks<KPL>(key, lane);
#pragma unroll
      for (int j = 0; j < KPL; ++j) sorted[lane * KPL + j] = key[j];
      __syncwarp();
      // lane q serves percentile q: numpy _lerp (float32) between the ranks k and k + 1
      for (int q = lane; q < p.n_pct; q += 32) {
        const int k = p.k_lo[q];
        const float lo = key2f(sorted[k]);
        const float hi = key2f(sorted[min(k + 1, p.N - 1)]);
        const float t = p.gamma[q];
        const float diff = __fsub_rn(hi, lo);
        float res = __fadd_rn(lo, __fmul_rn(diff, t));
        if (t >= 0.5f) res = __fsub_rn(hi, __fmul_rn(diff, __fsub_rn(1.0f, t)));
        p.out[obase + (long long)q * p.ny * p.nx] = res;
      }
      __syncwarp();
      continue;
    }
    // Bisection path (very large windows): percentiles are processed up to five at a time: their
    // bisections are independent, so interleaving them hides the latency of the warp reductions.
    int q0 = 0;
    while (q0 < p.n_pct) {
      const int nu = min(5, p.n_pct - q0);
      switch (nu) {
        case 5: select_group<KPL, 5>(key, p, q0, lane, obase); break;
        case 4: select_group<KPL, 4>(key, p, q0, lane, obase); break;
        case 3: select_group<KPL, 3>(key, p, q0, lane, obase); break;
        case 2: select_group<KPL, 2>(key, p, q0, lane, obase); break;
        default: select_group<KPL, 1>(key, p, q0, lane, obase); break;
      }
      q0 += nu;
    }
  }
}

size_t percentiles_workspace_bytes(long long n_slices, int ny, int nx, int cells) {
  const size_t nb = 2 * (size_t)cells + 1;
  size_t pad = (size_t)n_slices * (2 * (size_t)nx + 2 * ((size_t)ny + 2 * cells)) * 4;
  pad = (pad + 255) / 256 * 256;
  return pad + nb * nb * 4 + 256;
}

template <int KPL, bool SORT>
static int launch_pct(const PctParams& p, long long blocks, size_t smem, cudaStream_t st) {
  auto kern = percentile_kernel<KPL, SORT>;
  if (SORT) smem += (size_t)8 * 32 * KPL * 4;           // per-warp sorted blocks
  NBH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, 256, smem, st>>>(p);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int run_percentiles(const float* in, float* out, const float* pct, int n_pct, long long n_slices, int ny,
                    int nx, int cells, int32_t* status, void* workspace, size_t workspace_bytes,
                    cudaStream_t st) {
  NBH_REQUIRE(in && out && pct && status && workspace, "null pointer argument");
  NBH_REQUIRE(n_pct >= 1 && n_pct <= kMaxPct, "number of percentiles must be in [1, %d], got %d", kMaxPct,
              n_pct);
  NBH_REQUIRE(n_slices >= 1 && ny >= 1 && nx >= 1 && cells >= 1, "empty input");
  NBH_REQUIRE(workspace_bytes >= percentiles_workspace_bytes(n_slices, ny, nx, cells), "workspace too small");
  const int r = cells, nb = 2 * r + 1;
  // kernel cell offsets, row-major like kernel > 0 boolean indexing (nbhood.py:646, 662)
  static thread_local int h_offs[4096 * 8];
  int N = 0;
  for (int dy = -r; dy <= r; ++dy)
    for (int dx = -r; dx <= r; ++dx)
      if (dy * dy + dx * dx <= r * r) {
        NBH_REQUIRE(N < (int)(sizeof(h_offs) / sizeof(int)), "percentile radius of %d cells is too large", r);
        h_offs[N++] = ((dy + r) << 16) | (dx + r);
      }
  NBH_REQUIRE(N <= 32 * 72, "percentile neighbourhood of %d cells (radius %d) exceeds the supported %d", N, r,
              32 * 72);
  PctParams p;
  memset(&p, 0, sizeof(p));
  for (int q = 0; q < n_pct; ++q) {
    NBH_REQUIRE(pct[q] >= 0.f && pct[q] <= 100.f, "Percentiles must be in the range [0, 100]");
    // numpy: q32 = float32(p) / float32(100); vi = (N - 1) * q32 (float32)
    const float q32 = pct[q] / 100.0f;
    const float vi = (float)(N - 1) * q32;
    const float fl = floorf(vi);
    p.k_lo[q] = (int)fl;
    p.gamma[q] = vi - fl;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  size_t pad = (size_t)n_slices * (2 * (size_t)nx + 2 * ((size_t)ny + 2 * r)) * 4;
  pad = (pad + 255) / 256 * 256;
  float* padv = reinterpret_cast<float*>(base);
  int* d_offs = reinterpret_cast<int*>(base + pad);
  NBH_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  NBH_CHECK_CUDA(cudaMemcpyAsync(d_offs, h_offs, (size_t)N * 4, cudaMemcpyHostToDevice, st));
  pct_pad_kernel<<<(unsigned)n_slices, 256, 0, st>>>(in, padv, ny, nx, r);

  p.in = in; p.out = out; p.padv = padv; p.offs = d_offs;
  p.n_slices = n_slices; p.ny = ny; p.nx = nx; p.r = r; p.nb = nb; p.N = N; p.n_pct = n_pct;
  p.TW = 32; p.TH = 16;
  p.tiles_x = (nx + p.TW - 1) / p.TW;
  p.tiles_y = (ny + p.TH - 1) / p.TH;
  const long long blocks = (long long)p.tiles_x * p.tiles_y * n_slices;
  NBH_REQUIRE(blocks < (1LL << 31), "too many blocks");
  const int kpl = (N + 31) / 32;
  const int sizes[] = {1, 2, 4, 7, 10, 16, 24, 40, 72};
  int pick = 72;
  for (int s : sizes) if (s >= kpl) { pick = s; break; }
  const size_t smem = (size_t)(p.TW + 2 * r) * (p.TH + 2 * r) * 4 + (size_t)32 * pick * 4;
  ProfileScope prof(st, "nbh::percentile_kernel (warp merge-split sort / bisection)");
  switch (pick) {
    case 1: return launch_pct<1, true>(p, blocks, smem, st);
    case 2: return launch_pct<2, true>(p, blocks, smem, st);
    case 4: return launch_pct<4, true>(p, blocks, smem, st);
    case 7: return launch_pct<7, true>(p, blocks, smem, st);
    case 10: return launch_pct<10, true>(p, blocks, smem, st);
    case 16: return launch_pct<16, true>(p, blocks, smem, st);
    case 24: return launch_pct<24, false>(p, blocks, smem, st);
    case 40: return launch_pct<40, false>(p, blocks, smem, st);
    default: return launch_pct<72, false>(p, blocks, smem, st);
  }
}

}  // namespace nbh
