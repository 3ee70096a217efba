// multi-threshold kernel instantiations for a radius of 1 cells (see nbh_square_mt.cu)
#define NBH_MT_R 1
#include "nbh_square_mt.cu"
