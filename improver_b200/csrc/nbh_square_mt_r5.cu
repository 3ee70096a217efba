// multi-threshold kernel instantiations for a radius of 5 cells (see nbh_square_mt.cu)
#define NBH_MT_R 5
#include "nbh_square_mt.cu"
