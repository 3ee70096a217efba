// box kernel instantiations for a radius of 1 cells (see nbh_square_box.cu)
#define NBH_BOX_R 1
#include "nbh_square_box.cu"
