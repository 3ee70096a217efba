// box kernel instantiations for a radius of 2 cells (see nbh_square_box.cu)
#define NBH_BOX_R 2
#include "nbh_square_box.cu"
