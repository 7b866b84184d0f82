// adr2d_b200: same command line as the reference's advection_diffusion_reaction_2d
// (/root/reference/adr/advection_diffusion_reaction_2d.cpp), device path through libb200sts.so.
#include "b200_adr2d.h"

int main(int argc, char** argv) { return b200_adr_main(argc, argv); }
