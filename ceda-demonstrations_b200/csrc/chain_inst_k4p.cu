// explicit instantiation definitions of k_chain_march, depth 4, prefetch depth 4 (see chain_march_inst.cuh)
#include "chain_march_inst.cuh"
B200_CHAIN_K4P(B200_CHAIN_DEFINE)
