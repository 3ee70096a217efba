// y-band tiling of one huge grid over several GPUs: the only exchange step of
// the whole path.  A band of `rows` own rows per slice sends its first / last `cells` rows to the
// neighbouring ranks and receives theirs; the kernels then run on the extended band.
//
// nbh_halo_exchange does it in one stream-ordered call: pack (one kernel: own rows -> extended
// band, boundary rows -> send buffers), one ncclSend / ncclRecv pair per neighbour inside a group
// (NVLink / NVSwitch peer traffic, no host staging), unpack.  NCCL is resolved at run time from
// the library the host process has already loaded (torch ships libnccl.so.2), so
// libimprover_b200.so has no link-time dependency on it.
#include <dlfcn.h>

#include "nbh_common.cuh"

namespace nbh {

// 16-byte copies where the row blocks allow it
__global__ void __launch_bounds__(256)
halo_pack_kernel(const float* __restrict__ band, float* __restrict__ ext, float* __restrict__ send_up,
                 float* __restrict__ send_down, long long n_slices, int rows, int nx, int cells, int halo_top,
                 int halo_bottom) {
  const long long per_slice = (long long)rows * nx;
  const long long ext_slice = (long long)(halo_top + rows + halo_bottom) * nx;
  const long long edge = (long long)cells * nx;
  const long long total = n_slices * per_slice;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per_slice, o = i - s * per_slice;
    const float v = band[i];
    ext[s * ext_slice + (long long)halo_top * nx + o] = v;
    if (send_up && o < edge) send_up[s * edge + o] = v;
    if (send_down && o >= per_slice - edge) send_down[s * edge + (o - (per_slice - edge))] = v;
  }
}

__global__ void __launch_bounds__(256)
halo_unpack_kernel(float* __restrict__ ext, const float* __restrict__ recv_up, const float* __restrict__ recv_down,
                   long long n_slices, int rows, int nx, int halo_top, int halo_bottom) {
  const long long ext_slice = (long long)(halo_top + rows + halo_bottom) * nx;
  const long long top = (long long)halo_top * nx, bot = (long long)halo_bottom * nx;
  const long long per = top + bot;
  const long long total = n_slices * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per, o = i - s * per;
    if (o < top) ext[s * ext_slice + o] = recv_up[s * top + o];
    else ext[s * ext_slice + top + (long long)rows * nx + (o - top)] = recv_down[s * bot + (o - top)];
  }
}

static unsigned copy_grid(long long total) {
  const long long want = (total + 255) / 256;
  return (unsigned)max(1LL, min(want, (long long)sm_count() * 16));
}

int run_halo_pack(const float* band, float* ext, float* send_up, float* send_down, long long n_slices, int rows,
                  int nx, int cells, int halo_top, int halo_bottom, cudaStream_t st) {
  NBH_REQUIRE(band && ext, "null pointer argument");
  NBH_REQUIRE(n_slices > 0 && rows > 0 && nx > 0 && cells >= 0 && halo_top >= 0 && halo_bottom >= 0,
              "invalid band geometry");
  NBH_REQUIRE(rows >= cells, "band of %d rows is thinner than the halo of %d rows", rows, cells);
  halo_pack_kernel<<<copy_grid(n_slices * rows * nx), 256, 0, st>>>(band, ext, send_up, send_down, n_slices, rows,
                                                                   nx, cells, halo_top, halo_bottom);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int run_halo_unpack(float* ext, const float* recv_up, const float* recv_down, long long n_slices, int rows, int nx,
                    int halo_top, int halo_bottom, cudaStream_t st) {
  NBH_REQUIRE(ext, "null pointer argument");
  NBH_REQUIRE((halo_top == 0 || recv_up) && (halo_bottom == 0 || recv_down), "missing receive buffer");
  if (halo_top + halo_bottom == 0) return 0;
  halo_unpack_kernel<<<copy_grid(n_slices * (long long)(halo_top + halo_bottom) * nx), 256, 0, st>>>(
      ext, recv_up, recv_down, n_slices, rows, nx, halo_top, halo_bottom);
  NBH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- NCCL, resolved lazily from the already loaded library ----
typedef int (*nccl_p2p_fn)(const void*, size_t, int, int, void*, cudaStream_t);   // ncclSend (recv: void* buffer)
typedef int (*nccl_recv_fn)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_void_fn)(void);
typedef const char* (*nccl_err_fn)(int);
struct NcclApi {
  nccl_p2p_fn send = nullptr;
  nccl_recv_fn recv = nullptr;
  nccl_void_fn group_start = nullptr, group_end = nullptr;
  nccl_err_fn err = nullptr;
  bool tried = false;
};
static NcclApi g_nccl;

static bool load_nccl() {
  if (g_nccl.tried) return g_nccl.send != nullptr;
  g_nccl.tried = true;
  void* h = RTLD_DEFAULT;
  if (!dlsym(h, "ncclSend")) {
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
  }
  g_nccl.send = reinterpret_cast<nccl_p2p_fn>(dlsym(h, "ncclSend"));
  g_nccl.recv = reinterpret_cast<nccl_recv_fn>(dlsym(h, "ncclRecv"));
  g_nccl.group_start = reinterpret_cast<nccl_void_fn>(dlsym(h, "ncclGroupStart"));
  g_nccl.group_end = reinterpret_cast<nccl_void_fn>(dlsym(h, "ncclGroupEnd"));
  g_nccl.err = reinterpret_cast<nccl_err_fn>(dlsym(h, "ncclGetErrorString"));
  if (!g_nccl.send || !g_nccl.recv || !g_nccl.group_start || !g_nccl.group_end) { g_nccl.send = nullptr; return false; }
  return true;
}

#define NBH_CHECK_NCCL(expr)                                                              \
  do {                                                                                    \
    int _r = (expr);                                                                      \
    if (_r != 0) {                                                                        \
      set_error("%s failed: %s", #expr, g_nccl.err ? g_nccl.err(_r) : "nccl error");       \
      return 3;                                                                           \
    }                                                                                     \
  } while (0)

size_t halo_exchange_workspace_bytes(long long n_slices, int nx, int cells) {
  return 4 * (((size_t)n_slices * cells * nx * 4 + 255) / 256 * 256);
}

int run_halo_exchange(void* comm, int rank, int n_ranks, const float* band, float* ext, long long n_slices,
                      int rows, int nx, int cells, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  NBH_REQUIRE(comm != nullptr || n_ranks == 1, "null NCCL communicator");
  NBH_REQUIRE(rank >= 0 && rank < n_ranks, "invalid rank %d of %d", rank, n_ranks);
  NBH_REQUIRE(workspace_bytes >= halo_exchange_workspace_bytes(n_slices, nx, cells) && (workspace || cells == 0),
              "workspace too small");
  const int halo_top = rank > 0 ? cells : 0, halo_bottom = rank < n_ranks - 1 ? cells : 0;
  const size_t buf = ((size_t)n_slices * cells * nx * 4 + 255) / 256 * 256;
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  float* send_up = halo_top ? reinterpret_cast<float*>(base) : nullptr;
  float* send_down = halo_bottom ? reinterpret_cast<float*>(base + buf) : nullptr;
  float* recv_up = reinterpret_cast<float*>(base + 2 * buf);
  float* recv_down = reinterpret_cast<float*>(base + 3 * buf);
  if (run_halo_pack(band, ext, send_up, send_down, n_slices, rows, nx, cells, halo_top, halo_bottom, st)) return 1;
  if (halo_top + halo_bottom == 0) return 0;
  NBH_REQUIRE(load_nccl(), "NCCL (ncclSend / ncclRecv) is not available in this process");
  const size_t count = (size_t)n_slices * cells * nx;
  const int ncclFloat32 = 7;      // ncclDataType_t: ncclFloat32 = ncclFloat = 7
  NBH_CHECK_NCCL(g_nccl.group_start());
  if (halo_top) {
    NBH_CHECK_NCCL(g_nccl.send(send_up, count, ncclFloat32, rank - 1, comm, st));
    NBH_CHECK_NCCL(g_nccl.recv(recv_up, count, ncclFloat32, rank - 1, comm, st));
  }
  if (halo_bottom) {
    NBH_CHECK_NCCL(g_nccl.send(send_down, count, ncclFloat32, rank + 1, comm, st));
    NBH_CHECK_NCCL(g_nccl.recv(recv_down, count, ncclFloat32, rank + 1, comm, st));
  }
  NBH_CHECK_NCCL(g_nccl.group_end());
  return run_halo_unpack(ext, recv_up, recv_down, n_slices, rows, nx, halo_top, halo_bottom, st);
}

}  // namespace nbh
