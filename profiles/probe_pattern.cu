// Micro-benchmark: achievable HBM read bandwidth of the access patterns the
// square kernel can use on the (18, 36, 970, 1042) cube.  Pure loads + adds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_pattern probe_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int R = 18, L = 36, NY = 970, NX = 1042;

// pattern A: one warp = one 64-column strip x y segment of one lead time, 18 member loads per row
template <int ROWS_UNROLL>
__global__ void strip_kernel(const float* __restrict__ in, float* __restrict__ out, int n_strips, int n_ysegs,
                             int seg_rows, long long n_tasks, int strip_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long task = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (task >= n_tasks) return;
  const int strip = task % n_strips; task /= n_strips;
  const int yseg = task % n_ysegs; task /= n_ysegs;
  const int l = (int)task;
  const int x = strip * strip_w + lane * 2;
  if (x >= NX) return;
  const int ya = yseg * seg_rows, yb = min(NY, ya + seg_rows);
  const float* base = in + (long long)l * NY * NX + x;
  const long long ms = (long long)L * NY * NX;
  float acc0 = 0.f, acc1 = 0.f;
  for (int y = ya; y < yb; y += ROWS_UNROLL) {
    float2 v[ROWS_UNROLL][R];
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u)
#pragma unroll
      for (int m = 0; m < R; ++m)
        v[u][m] = (y + u < yb) ? __ldg(reinterpret_cast<const float2*>(base + m * ms + (long long)(y + u) * NX))
                               : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u)
#pragma unroll
      for (int m = 0; m < R; ++m) { acc0 += v[u][m].x; acc1 += v[u][m].y; }
  }
  if (acc0 + acc1 == 12345.678f) out[0] = acc0;
}

// pattern A', persistent: grid = 148 * blocks_per_sm, warps loop over tasks (occupancy control without
// touching the L1 / shared-memory split)
template <int ROWS_UNROLL>
__global__ void strip_persist(const float* __restrict__ in, float* __restrict__ out, int n_strips, int n_ysegs,
                              int seg_rows, long long n_tasks, int strip_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float acc0 = 0.f, acc1 = 0.f;
  for (long long task0 = (long long)blockIdx.x * (blockDim.x >> 5) + warp; task0 < n_tasks; task0 += nw) {
    long long task = task0;
    const int strip = task % n_strips; task /= n_strips;
    const int yseg = task % n_ysegs; task /= n_ysegs;
    const int l = (int)task;
    const int x = strip * strip_w + lane * 2;
    if (x >= NX) continue;
    const int ya = yseg * seg_rows, yb = min(NY, ya + seg_rows);
    const float* base = in + (long long)l * NY * NX + x;
    const long long ms = (long long)L * NY * NX;
    for (int y = ya; y < yb; y += ROWS_UNROLL) {
      float2 v[ROWS_UNROLL][R];
#pragma unroll
      for (int u = 0; u < ROWS_UNROLL; ++u)
#pragma unroll
        for (int m = 0; m < R; ++m)
          v[u][m] = (y + u < yb) ? __ldg(reinterpret_cast<const float2*>(base + m * ms + (long long)(y + u) * NX))
                                 : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < ROWS_UNROLL; ++u)
#pragma unroll
        for (int m = 0; m < R; ++m) { acc0 += v[u][m].x; acc1 += v[u][m].y; }
    }
  }
  if (acc0 + acc1 == 12345.678f) out[0] = acc0;
}

// pattern A'', persistent + register double buffering (issue next row, then consume current row) with
// optional dummy ALU work per element, to see whether the hardware overlaps the two buffers
template <int WORK>
__global__ void strip_pipelined(const float* __restrict__ in, float* __restrict__ out, int n_strips, int n_ysegs,
                                int seg_rows, long long n_tasks, int strip_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float acc0 = 0.f, acc1 = 0.f;
  const long long ms = (long long)L * NY * NX;
  auto issue = [&](float2 (&v)[R], const float* base, int y, int yb) {
#pragma unroll
    for (int m = 0; m < R; ++m)
      if (y < yb) v[m] = __ldg(reinterpret_cast<const float2*>(base + m * ms + (long long)y * NX));
  };
  auto consume = [&](float2 (&v)[R], int y, int yb) {
    if (y >= yb) return;
#pragma unroll
    for (int m = 0; m < R; ++m) {
      float a = v[m].x, b = v[m].y;
#pragma unroll
      for (int w = 0; w < WORK; ++w) { a = fmaf(a, 1.0001f, 0.5f); b = fmaf(b, 1.0001f, 0.5f); }
      acc0 += a; acc1 += b;
    }
  };
  for (long long task0 = (long long)blockIdx.x * (blockDim.x >> 5) + warp; task0 < n_tasks; task0 += nw) {
    long long task = task0;
    const int strip = task % n_strips; task /= n_strips;
    const int yseg = task % n_ysegs; task /= n_ysegs;
    const int l = (int)task;
    const int x = strip * strip_w + lane * 2;
    if (x >= NX) continue;
    const int ya = yseg * seg_rows, yb = min(NY, ya + seg_rows);
    const float* base = in + (long long)l * NY * NX + x;
    float2 A[R], B[R];
    issue(A, base, ya, yb);
    for (int y = ya; y < yb; y += 2) {
      issue(B, base, y + 1, yb);
      consume(A, y, yb);
      issue(A, base, y + 2, yb);
      consume(B, y + 1, yb);
    }
  }
  if (acc0 + acc1 == 12345.678f) out[0] = acc0;
}

// pattern B: flat streaming: thread i sums the 18 members of element pair i (what torch.sum(dim=0) does)
__global__ void flat_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_pairs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const long long ms = (long long)L * NY * NX;
  float2 v[R];
#pragma unroll
  for (int m = 0; m < R; ++m) v[m] = __ldg(reinterpret_cast<const float2*>(in + m * ms) + i);
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int m = 0; m < R; ++m) { a += v[m].x; b += v[m].y; }
  reinterpret_cast<float2*>(out)[i] = make_float2(a, b);
}

template <class F>
float time_ms(F f, int reps = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const long long n = (long long)R * L * NY * NX;
  float *in, *out;
  cudaMalloc(&in, n * 4);
  cudaMalloc(&out, (long long)L * NY * NX * 4);
  cudaMemset(in, 0, n * 4);
  const double gb = n * 4 / 1e9;
  {
    const long long pairs = (long long)L * NY * NX / 2;
    float ms = time_ms([&] { flat_kernel<<<(unsigned)((pairs + 255) / 256), 256>>>(in, out, pairs); });
    printf("flat 18-member sum: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
  }
  for (int strip_w : {64, 54}) {
    for (int wpb : {4, 8}) {
      for (int ysegs : {4, 10, 20}) {
        const int n_strips = (NX + strip_w - 1) / strip_w;
        const int seg_rows = (NY + ysegs - 1) / ysegs;
        const long long n_tasks = (long long)n_strips * ysegs * L;
        const unsigned blocks = (unsigned)((n_tasks + wpb - 1) / wpb);
        float ms1 = time_ms([&] { strip_kernel<1><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
        float ms2 = time_ms([&] { strip_kernel<2><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
        const double eff = (strip_w == 64) ? 1.0 : 64.0 / 54.0;   // overlapping strips re-read the halo
        printf("strip w=%d warps/block=%d ysegs=%d: unroll1 %.3f ms (%.0f GB/s useful) unroll2 %.3f ms (%.0f GB/s)\n",
               strip_w, wpb, ysegs, ms1, gb / ms1 * 1e3, ms2, gb / ms2 * 1e3);
        (void)eff;
      }
    }
  }
  // occupancy sweep with a persistent grid (L1 size untouched)
  {
    const int strip_w = 64, wpb = 4, ysegs = 10;
    const int n_strips = (NX + strip_w - 1) / strip_w;
    const int seg_rows = (NY + ysegs - 1) / ysegs;
    const long long n_tasks = (long long)n_strips * ysegs * L;
    for (int bps : {1, 2, 3, 4, 6, 8, 12, 16}) {
      const unsigned blocks = 148 * bps;
      float ms1 = time_ms([&] { strip_persist<1><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
      float ms2 = time_ms([&] { strip_persist<2><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
      printf("warps/SM=%d: 18 loads in flight/warp %.3f ms (%.0f GB/s); 36 loads %.3f ms (%.0f GB/s)\n", bps * wpb,
             ms1, gb / ms1 * 1e3, ms2, gb / ms2 * 1e3);
    }
  }
  {
    const int strip_w = 64, wpb = 4, ysegs = 10;
    const int n_strips = (NX + strip_w - 1) / strip_w;
    const int seg_rows = (NY + ysegs - 1) / ysegs;
    const long long n_tasks = (long long)n_strips * ysegs * L;
    for (int bps : {2, 3, 4}) {
      const unsigned blocks = 148 * bps;
      float m0 = time_ms([&] { strip_pipelined<0><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
      float m4 = time_ms([&] { strip_pipelined<4><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
      float m8 = time_ms([&] { strip_pipelined<8><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
      float m16 = time_ms([&] { strip_pipelined<16><<<blocks, wpb * 32>>>(in, out, n_strips, ysegs, seg_rows, n_tasks, strip_w); });
      printf("pipelined warps/SM=%d: work0 %.3f ms (%.0f GB/s) work4 %.3f (%.0f) work8 %.3f (%.0f) work16 %.3f (%.0f)\n",
             bps * wpb, m0, gb / m0 * 1e3, m4, gb / m4 * 1e3, m8, gb / m8 * 1e3, m16, gb / m16 * 1e3);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
