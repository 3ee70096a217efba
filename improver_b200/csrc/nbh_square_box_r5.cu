// box kernel instantiations for a radius of 5 cells (see nbh_square_box.cu)
#define NBH_BOX_R 5
#include "nbh_square_box.cu"
