// box kernel instantiations for a radius of 7 cells (see nbh_square_box.cu)
#define NBH_BOX_R 7
#include "nbh_square_box.cu"
