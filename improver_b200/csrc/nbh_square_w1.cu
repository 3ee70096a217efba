// Instantiations of the warp-autonomous square kernel for 1-wide vector loads.
#include "nbh_square_warp.cuh"

namespace nbh {
NBH_DEFINE_SQUARE_WARP_VEC(1)
}  // namespace nbh
