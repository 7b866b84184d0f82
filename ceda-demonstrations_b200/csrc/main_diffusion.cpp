// diffusion_2D_b200: same command line as the reference's diffusion_2D_mpi
// (/root/reference/diffusion_2D/main.cpp), device path through libb200sts.so.
#include "b200_diffusion2d.h"

int main(int argc, char** argv) { return b200_d2d_main(argc, argv); }
