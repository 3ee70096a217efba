// multi-threshold kernel instantiations for a radius of 4 cells (see nbh_square_mt.cu)
#define NBH_MT_R 4
#include "nbh_square_mt.cu"
