// Instantiations of the square neighbourhood kernel for 1-wide vector loads.
#include "nbh_square.cuh"

namespace nbh {
NBH_DEFINE_SQUARE_VEC(1)
}  // namespace nbh
