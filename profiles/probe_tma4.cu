// TMA 4-D probe variants (one per process: a faulting launch poisons the context)
//   argv[1]: 0 = L2 promotion NONE, 1 = 128B, 2 = 256B ; argv[2]: rank (3 or 4) ; argv[3]: x0
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include "../improver_b200/csrc/nbh_square_tma.cuh"
using namespace nbh;
namespace nbh { void set_error(const char*, ...) {} }

template <int RANK>
__global__ void probe(const CUtensorMap* tmap, const float* in, int ny, int nx, int L, int R, int x0, int* bad) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + R * 64 * 4);
  const int lane = threadIdx.x;
  const int y = blockIdx.x % ny, plane = blockIdx.x / ny;
  const unsigned b = smem_u32(bar), s = smem_u32(stage);
  if (lane == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, R * 64 * 4);
    const int cx = x0 + (y & 1) * nx, cy = y >> 1;
    if (RANK == 4) tma_load_4d(s, tmap, cx, cy, plane, 0, b);
    else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                      ::"r"(s), "l"(tmap), "r"(cx), "r"(cy + plane * (ny / 2)), "r"(0), "r"(b) : "memory");
  }
  __syncwarp();
  mbar_wait(b, 0);
  for (int m = 0; m < R; ++m)
    for (int c = lane; c < 64; c += 32) {
      const float ref = in[((long long)m * L + plane) * ny * nx + (long long)y * nx + x0 + c];
      if (x0 + c < nx && stage[m * 64 + c] != ref) atomicAdd(bad, 1);
    }
}

int main(int argc, char** argv) {
  const int promo = atoi(argv[1]), rank = atoi(argv[2]), x0 = atoi(argv[3]);
  const int R = 3, L = 2, ny = 64, nx = 98;
  std::vector<float> h((size_t)R * L * ny * nx);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973);
  float* d; int* bad;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  Fn fn = (Fn)ptr;
  CUtensorMap map;
  CUresult rc;
  const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                    : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (rank == 4) {
    const cuuint64_t dims[4] = {(cuuint64_t)2 * nx, (cuuint64_t)ny / 2, (cuuint64_t)L, (cuuint64_t)R};
    const cuuint64_t strides[3] = {(cuuint64_t)2 * nx * 4, (cuuint64_t)ny * nx * 4, (cuuint64_t)L * ny * nx * 4};
    const cuuint32_t box[4] = {64, 1, 1, (cuuint32_t)R};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    rc = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[3] = {(cuuint64_t)2 * nx, (cuuint64_t)(ny / 2) * L, (cuuint64_t)R};
    const cuuint64_t strides[2] = {(cuuint64_t)2 * nx * 4, (cuuint64_t)L * ny * nx * 4};
    const cuuint32_t box[3] = {64, 1, (cuuint32_t)R};
    const cuuint32_t es[3] = {1, 1, 1};
    rc = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  CUtensorMap* dmap; cudaMalloc(&dmap, 128); cudaMemcpy(dmap, &map, 128, cudaMemcpyHostToDevice);
  if (rank == 4) probe<4><<<ny * L, 32, R * 64 * 4 + 64>>>(dmap, d, ny, nx, L, R, x0, bad);
  else probe<3><<<ny * L, 32, R * 64 * 4 + 64>>>(dmap, d, ny, nx, L, R, x0, bad);
  cudaError_t e = cudaDeviceSynchronize();
  int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("promo=%d rank=%d x0=%d encode=%d sync: %s, mismatches=%d\n", promo, rank, x0, (int)rc, cudaGetErrorString(e), hb);
  return 0;
}
