// multi-threshold kernel instantiations for a radius of 2 cells (see nbh_square_mt.cu)
#define NBH_MT_R 2
#include "nbh_square_mt.cu"
