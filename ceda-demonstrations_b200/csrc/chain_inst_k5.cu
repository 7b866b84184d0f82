// explicit instantiation definitions of k_chain_march, depth 5 (see chain_march_inst.cuh)
#include "chain_march_inst.cuh"
B200_CHAIN_K5(B200_CHAIN_DEFINE)
