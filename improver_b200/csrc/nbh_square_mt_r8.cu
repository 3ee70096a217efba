// multi-threshold kernel instantiations for a radius of 8 cells (see nbh_square_mt.cu)
#define NBH_MT_R 8
#include "nbh_square_mt.cu"
