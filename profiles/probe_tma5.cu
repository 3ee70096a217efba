// TMA bisect: generic rank-N tiled load of a box described on the command line.
//   probe_tma5 rank boxrows boxouter   (dims: {196, 32, 6}; rank 2 folds everything into rows)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include "../improver_b200/csrc/nbh_square_tma.cuh"
using namespace nbh;
namespace nbh { void set_error(const char*, ...) {} }

__global__ void k2(const CUtensorMap* tmap, int bytes, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 8192);
  const unsigned b = smem_u32(bar), s = smem_u32(smem);
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, bytes);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s), "l"(tmap), "r"(16), "r"(3), "r"(b) : "memory");
  }
  __syncwarp();
  mbar_wait(b, 0);
  for (int c = threadIdx.x; c < bytes / 4; c += 32) out[c] = reinterpret_cast<float*>(smem)[c];
}
__global__ void k3(const CUtensorMap* tmap, int bytes, float* out, int baroff) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + baroff);
  const unsigned b = smem_u32(bar), s = smem_u32(smem);
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, bytes);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s), "l"(tmap), "r"(16 + (int)(blockIdx.x & 1) * 98), "r"(3 + (int)(blockIdx.x >> 1) % 20), "r"(0), "r"(b) : "memory");
  }
  __syncwarp();
  mbar_wait(b, 0);
  if (blockIdx.x == 0) for (int c = threadIdx.x; c < bytes / 4; c += 32) out[c] = reinterpret_cast<float*>(smem)[c];
}

int main(int argc, char** argv) {
  const int rank = atoi(argv[1]), brow = atoi(argv[2]), bout = atoi(argv[3]);
  const int W = 196, H = 32, D = 6;
  std::vector<float> h((size_t)W * H * D);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 16384);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  Fn fn = (Fn)ptr;
  CUtensorMap map;
  CUresult rc;
  int bytes;
  if (rank == 2) {
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H * D};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {64, (cuuint32_t)brow};
    const cuuint32_t es[2] = {1, 1};
    rc = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 64 * brow * 4;
  } else {
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {64, (cuuint32_t)brow, (cuuint32_t)bout};
    const cuuint32_t es[3] = {1, 1, 1};
    rc = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 64 * brow * bout * 4;
  }
  CUtensorMap* dmap; cudaMalloc(&dmap, 128); cudaMemcpy(dmap, &map, 128, cudaMemcpyHostToDevice);
  const int nblk = argc > 4 ? atoi(argv[4]) : 1; const int baroff = argc > 5 ? atoi(argv[5]) : 8192;
  if (rank == 2) k2<<<1, 32, 8192 + 64>>>(dmap, bytes, out); else k3<<<nblk, 32, baroff + 64>>>(dmap, bytes, out, baroff);
  cudaError_t e = cudaDeviceSynchronize();
  float o[4] = {-1, -1, -1, -1};
  cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
  printf("rank=%d box=(64,%d,%d) encode=%d sync: %s first=%g (expect %d)\n", rank, brow, bout, (int)rc,
         cudaGetErrorString(e), o[0], 3 * W + 16);
  return 0;
}
