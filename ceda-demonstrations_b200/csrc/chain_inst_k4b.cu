// explicit instantiation definitions of k_chain_march, depth 4, BULK flavour (see chain_march_inst.cuh)
#include "chain_march_inst.cuh"
B200_CHAIN_K4B(B200_CHAIN_DEFINE_BULK)
