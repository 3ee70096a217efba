// multi-threshold kernel instantiations for a radius of 6 cells (see nbh_square_mt.cu)
#define NBH_MT_R 6
#include "nbh_square_mt.cu"
