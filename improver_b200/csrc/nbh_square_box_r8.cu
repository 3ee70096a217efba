// box kernel instantiations for a radius of 8 cells (see nbh_square_box.cu)
#define NBH_BOX_R 8
#include "nbh_square_box.cu"
