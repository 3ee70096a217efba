// Minimal TMA probe: 4-D tiled tensor map over (members, planes, super-rows, 2*nx), one warp per block
// loads a (64 x 1 x 1 x R) box and checks it against direct loads.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include "../improver_b200/csrc/nbh_square_tma.cuh"
using namespace nbh;
namespace nbh { void set_error(const char*, ...) {} }

__global__ void probe_mbar_only(int* bad) {
  __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
  }
  __syncwarp();
  mbar_wait(b, 0);
  if (threadIdx.x == 0) atomicAdd(bad, 1);
}

__global__ void probe_bulk1d(const float* in, int* bad) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 1024);
  const unsigned b = smem_u32(bar), s = smem_u32(stage);
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, 1024);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s), "l"(in), "r"(1024), "r"(b) : "memory");
  }
  __syncwarp();
  mbar_wait(b, 0);
  for (int c = threadIdx.x; c < 256; c += 32) if (stage[c] != in[c]) atomicAdd(bad, 1);
}

__global__ void probe2d(const __grid_constant__ CUtensorMap tmap, const float* in, int nx, int* bad) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 1024);
  const unsigned b = smem_u32(bar), s = smem_u32(stage);
  const int y = blockIdx.x, x0 = 18;
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, 256);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s), "l"(&tmap), "r"(x0), "r"(y), "r"(b) : "memory");
  }
  __syncwarp();
  mbar_wait(b, 0);
  for (int c = threadIdx.x; c < 64; c += 32)
    if (x0 + c < nx && stage[c] != in[(long long)y * nx + x0 + c]) atomicAdd(bad, 1);
}

__global__ void probe2d_gmem(const CUtensorMap* tmap, const float* in, int nx, int* bad) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 1024);
  const unsigned b = smem_u32(bar), s = smem_u32(stage);
  const int y = blockIdx.x, x0 = 16;
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, 256);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s), "l"(tmap), "r"(x0), "r"(y), "r"(b) : "memory");
  }
  __syncwarp();
  mbar_wait(b, 0);
  for (int c = threadIdx.x; c < 64; c += 32)
    if (x0 + c < nx && stage[c] != in[(long long)y * nx + x0 + c]) atomicAdd(bad, 1);
}

__global__ void probe(const CUtensorMap* tmap, const float* in, int ny, int nx, int L, int R,
                      int super, int* bad) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + R * 64 * 4);
  const int lane = threadIdx.x;
  const int y = blockIdx.x % ny, plane = blockIdx.x / ny, x0 = 16;
  const unsigned b = smem_u32(bar), s = smem_u32(stage);
  if (lane == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, R * 64 * 4);
    tma_load_4d(s, tmap, x0 + (super ? (y & 1) * nx : 0), super ? (y >> 1) : y, plane, 0, b);
  }
  __syncwarp();
  mbar_wait(b, 0);
  for (int m = 0; m < R; ++m)
    for (int c = lane; c < 64; c += 32) {
      const float ref = in[((long long)m * L + plane) * ny * nx + (long long)y * nx + x0 + c];
      if (x0 + c < nx && stage[m * 64 + c] != ref) atomicAdd(bad, 1);
    }
}

namespace nbh {
bool make_square_tensor_map(const float* in, int ny, int nx, long long n_planes, int n_members,
                            long long plane_stride, long long member_stride, int box_cols, CUtensorMap* map,
                            int* super_rows);
}

int main() {
  const int R = 3, L = 2, ny = 64, nx = 98;
  std::vector<float> h((size_t)R * L * ny * nx);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973);
  float* d; int* bad;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  probe_mbar_only<<<1, 32>>>(bad);
  printf("mbar only: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  probe_bulk1d<<<1, 32, 2048>>>(d, bad);
  printf("bulk 1d: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  cudaMemset(bad, 0, 4);
  {
    // 2-D map over the first plane: rows of nx floats need nx*4 % 16 == 0 -> use nx2 = 96 sub-view? use super rows
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    Fn fn = (Fn)ptr;
    CUtensorMap m2;
    const cuuint64_t dims[2] = {(cuuint64_t)2 * nx, (cuuint64_t)ny / 2};
    const cuuint64_t strides[1] = {(cuuint64_t)2 * nx * 4};
    const cuuint32_t box[2] = {64, 1};
    const cuuint32_t es[2] = {1, 1};
    CUresult rc = fn(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode2d rc=%d\n", (int)rc);
    for (int i = 0; i < 16; ++i) printf("%016llx ", (unsigned long long)m2.opaque[i]);
    printf("\n");
    {
      CUtensorMap* dm; cudaMalloc(&dm, 128); cudaMemcpy(dm, &m2, 128, cudaMemcpyHostToDevice);
      probe2d_gmem<<<ny / 2, 32, 2048>>>(dm, d, 2 * nx, bad);
      cudaError_t e = cudaDeviceSynchronize();
      int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
      printf("2d gmem desc: %s mismatches=%d\n", cudaGetErrorString(e), hb);
      if (e != cudaSuccess) return 2;
      cudaMemset(bad, 0, 4);
    }
    cudaMemset(bad, 0, 4);
  }
  CUtensorMap map; int super = 0;
  bool ok = make_square_tensor_map(d, ny, nx, L, R, (long long)ny * nx, (long long)L * ny * nx, 64, &map, &super);
  printf("encode ok=%d super=%d\n", ok, super);
  if (!ok) return 1;
  CUtensorMap* dmap; cudaMalloc(&dmap, 128); cudaMemcpy(dmap, &map, 128, cudaMemcpyHostToDevice);
  probe<<<ny * L, 32, R * 64 * 4 + 64>>>(dmap, d, ny, nx, L, R, super, bad);
  cudaError_t e = cudaDeviceSynchronize();
  int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("sync: %s, mismatches=%d\n", cudaGetErrorString(e), hb);
  return 0;
}
