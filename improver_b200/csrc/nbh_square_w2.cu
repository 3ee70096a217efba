// Instantiations of the warp-autonomous square kernel for 2-wide vector loads.
#include "nbh_square_warp.cuh"

namespace nbh {
NBH_DEFINE_SQUARE_WARP_VEC(2)
}  // namespace nbh
