// multi-threshold kernel instantiations for a radius of 7 cells (see nbh_square_mt.cu)
#define NBH_MT_R 7
#include "nbh_square_mt.cu"
