// Instantiations of the square neighbourhood kernel for 4-wide vector loads.
#include "nbh_square.cuh"

namespace nbh {
NBH_DEFINE_SQUARE_VEC(4)
}  // namespace nbh
