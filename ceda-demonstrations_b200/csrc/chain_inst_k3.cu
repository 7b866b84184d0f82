// explicit instantiation definitions of k_chain_march, depth 3 (see chain_march_inst.cuh)
#include "chain_march_inst.cuh"
B200_CHAIN_K3(B200_CHAIN_DEFINE)
