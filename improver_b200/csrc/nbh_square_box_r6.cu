// box kernel instantiations for a radius of 6 cells (see nbh_square_box.cu)
#define NBH_BOX_R 6
#include "nbh_square_box.cu"
