// chain_march_inst.cuh -- the list of k_chain_march instantiations the library uses, as X-macros.
//
// The kernel is instantiated for 5 depths x { wrap, deep halo } x { exact, FMA } x { tables, uniform } x { body, head }
// (+ the SPLIT flavour of depth 4): about a hundred large kernels, minutes of NVVM optimisation in one translation
// unit.  Each depth is therefore compiled in its OWN translation unit (csrc/chain_inst_k<K>.cu: explicit instantiation
// definitions, built in parallel by make -j) and b200_kernels.cu only DECLARES them (extern template), so it launches
// kernels whose device code lives in another object of the same shared library.
#pragma once
#include <cstdint>
#include "b200_sts.h"
#include "chain_march.cuh"

// F(K, PF, HALO, FMA, UNI, HEAD, SPLIT)
#define B200_CHAIN_FLAVOURS(F, K, PF, SPLIT)                                                                             \
  F(K, PF, false, false, false, false, SPLIT) F(K, PF, false, false, false, true, SPLIT)                                 \
  F(K, PF, false, false, true, false, SPLIT)  F(K, PF, false, false, true, true, SPLIT)                                  \
  F(K, PF, true, false, false, false, SPLIT)  F(K, PF, true, false, false, true, SPLIT)                                  \
  F(K, PF, true, false, true, false, SPLIT)   F(K, PF, true, false, true, true, SPLIT)
#define B200_CHAIN_FLAVOURS_FMA(F, K, PF)                                                                                \
  F(K, PF, false, true, false, false, false) F(K, PF, false, true, false, true, false)                                   \
  F(K, PF, false, true, true, false, false)  F(K, PF, false, true, true, true, false)                                    \
  F(K, PF, true, true, false, false, false)  F(K, PF, true, true, false, true, false)                                    \
  F(K, PF, true, true, true, false, false)   F(K, PF, true, true, true, true, false)

#define B200_CHAIN_K2(F) B200_CHAIN_FLAVOURS(F, 2, 4, false) B200_CHAIN_FLAVOURS_FMA(F, 2, 4)
#define B200_CHAIN_K3(F) B200_CHAIN_FLAVOURS(F, 3, 4, false) B200_CHAIN_FLAVOURS_FMA(F, 3, 4)
#define B200_CHAIN_K4(F) B200_CHAIN_FLAVOURS(F, 4, 3, false) B200_CHAIN_FLAVOURS_FMA(F, 4, 3)
#define B200_CHAIN_K4S(F) B200_CHAIN_FLAVOURS(F, 4, 3, true)
#define B200_CHAIN_K4P(F) B200_CHAIN_FLAVOURS(F, 4, 4, false) // one more row in flight (exact arithmetic)
#define B200_CHAIN_K5(F) B200_CHAIN_FLAVOURS(F, 5, 3, false) B200_CHAIN_FLAVOURS_FMA(F, 5, 3)
#define B200_CHAIN_K6(F) B200_CHAIN_FLAVOURS(F, 6, 3, false) B200_CHAIN_FLAVOURS_FMA(F, 6, 3)

// BULK flavours (operand ring filled by cp.async.bulk + mbarrier), exact arithmetic: FB(K, PF, HALO, UNI, HEAD)
#define B200_CHAIN_BULK_FLAVOURS(FB, K, PF)                                                                              \
  FB(K, PF, false, false, false) FB(K, PF, false, false, true) FB(K, PF, false, true, false) FB(K, PF, false, true, true) \
  FB(K, PF, true, false, false)  FB(K, PF, true, false, true)  FB(K, PF, true, true, false)  FB(K, PF, true, true, true)
#define B200_CHAIN_K4B(FB) B200_CHAIN_BULK_FLAVOURS(FB, 4, 3) B200_CHAIN_BULK_FLAVOURS(FB, 4, 4)
#define B200_CHAIN_DECLARE_BULK(K, PF, HALO, UNI, HEAD) \
  extern template __global__ void k_chain_march<K, PF, HALO, false, UNI, HEAD, false, true>(const ChainArgs);
#define B200_CHAIN_DEFINE_BULK(K, PF, HALO, UNI, HEAD) \
  template __global__ void k_chain_march<K, PF, HALO, false, UNI, HEAD, false, true>(const ChainArgs);

#define B200_CHAIN_DECLARE(K, PF, HALO, FMA, UNI, HEAD, SPLIT) \
  extern template __global__ void k_chain_march<K, PF, HALO, FMA, UNI, HEAD, SPLIT>(const ChainArgs);
#define B200_CHAIN_DEFINE(K, PF, HALO, FMA, UNI, HEAD, SPLIT) \
  template __global__ void k_chain_march<K, PF, HALO, FMA, UNI, HEAD, SPLIT>(const ChainArgs);
