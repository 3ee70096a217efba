// box kernel instantiations for a radius of 3 cells (see nbh_square_box.cu)
#define NBH_BOX_R 3
#include "nbh_square_box.cu"
