// multi-threshold kernel instantiations for a radius of 3 cells (see nbh_square_mt.cu)
#define NBH_MT_R 3
#include "nbh_square_mt.cu"
