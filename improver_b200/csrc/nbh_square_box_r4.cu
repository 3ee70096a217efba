// box kernel instantiations for a radius of 4 cells (see nbh_square_box.cu)
#define NBH_BOX_R 4
#include "nbh_square_box.cu"
