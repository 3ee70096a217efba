// Micro-benchmark: per-SM throughput of cp.async.bulk (1-D) global->shared copies as a
// function of copy size, source alignment, copies in flight and number of issuing threads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/probe_bulk profiles/probe_bulk.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// each issuing thread owns `depth` slots of `bytes`; it streams its share of the buffer
__global__ void probe(const char* src, size_t total, unsigned bytes, unsigned stride, unsigned offset, int depth, int n_issuers, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0 || warp >= n_issuers) return;
  const unsigned slot_bytes = (bytes + 127) & ~127u;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)n_issuers * depth * slot_bytes);
  const unsigned bar0 = smem_u32(bars + warp * depth);
  const unsigned dst0 = smem_u32(smem + (size_t)warp * depth * slot_bytes);
  for (int s = 0; s < depth; ++s) mbar_init(bar0 + 8 * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const size_t n_copies = total / stride;
  const size_t issuer = (size_t)blockIdx.x * n_issuers + warp, n_iss = (size_t)gridDim.x * n_issuers;
  size_t issued = 0, done = 0;
  float acc = 0.f;
  for (size_t c = issuer; c < n_copies; c += n_iss, ++issued) {
    const int s = (int)(issued % depth);
    if (issued >= (size_t)depth) {
      mbar_wait(bar0 + 8 * s, (unsigned)((issued / depth - 1) & 1)); ++done;
      acc += *reinterpret_cast<volatile float*>(smem + (size_t)warp * depth * slot_bytes + (size_t)s * slot_bytes);
    }
    mbar_expect_tx(bar0 + 8 * s, bytes);
    bulk_g2s(dst0 + s * slot_bytes, src + c * stride + offset, bytes, bar0 + 8 * s);
  }
  for (; done < issued; ++done) { const int s = (int)(done % depth); mbar_wait(bar0 + 8 * s, (unsigned)((done / depth) & 1)); }
  if (acc == 123.f) *sink = acc;
}
int main() {
  const size_t total = (size_t)2600 << 20;
  char* buf; float* sink;
  cudaMalloc(&buf, total + 65536); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, total + 65536);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { unsigned bytes, stride, offset; int depth, issuers; };
  Cfg cfgs[] = {
    {4176, 4176, 0, 32, 1}, {4176, 4176, 0, 8, 1}, {4176, 4176, 0, 2, 1},  // like the UK rows
    {4096, 4096, 0, 32, 1}, {4096, 4096, 16, 32, 1}, {4096, 4096, 64, 32, 1},
    {4096, 4096, 0, 8, 4}, {4096, 4096, 0, 4, 8}, {4096, 4096, 0, 2, 16},
    {4176, 4176, 0, 8, 4}, {4176, 4176, 0, 4, 8}, {4176, 4176, 0, 2, 16},
    {16384, 16384, 0, 8, 1}, {16384, 16384, 16, 8, 1}, {16384, 16384, 0, 2, 4},
    {1024, 1024, 0, 64, 1}, {1024, 1024, 0, 16, 4}, {1024, 1024, 0, 8, 16},
    {2048, 2048, 0, 16, 4}, {8192, 8192, 0, 4, 4}, {8192, 8192, 16, 4, 4},
    {32768, 32768, 0, 4, 1}, {65536, 65536, 0, 2, 1}, {65536, 65536, 16, 2, 1},
  };
  for (const Cfg& c : cfgs) {
    const unsigned slot = (c.bytes + 127) & ~127u;
    const size_t smem = (size_t)c.issuers * c.depth * slot + (size_t)c.issuers * c.depth * 8 + 128;
    if (smem > 220 * 1024) { printf("skip (smem)\n"); continue; }
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      probe<<<148, c.issuers * 32, smem>>>(buf, total, c.bytes, c.stride, c.offset, c.depth, c.issuers, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double moved = (double)(total / c.stride) * c.bytes;
    printf("bytes %6u stride %6u off %3u depth %3d issuers %2d : %.3f ms  %.0f GB/s  (%s)\n", c.bytes, c.stride, c.offset,
           c.depth, c.issuers, best, moved / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
