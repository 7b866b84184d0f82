// explicit instantiation definitions of k_chain_march, depth 4S (see chain_march_inst.cuh)
#include "chain_march_inst.cuh"
B200_CHAIN_K4S(B200_CHAIN_DEFINE)
