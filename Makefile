# Top-level build: everything is compiled in-tree so the artefacts travel with gpurun.
#
#   make            product: SUNDIALS host lib (from /root/reference), CUDA kernel lib,
#                   N_Vector/problem lib, driver executables
#   make oracle     oracle/liboracle_sts.so (C restatement) and, when /root/reference is
#                   present, oracle/_ref/* (the unmodified reference CPU binaries)
#
# Layout of the outputs (all git-ignored):
#   ceda-demonstrations_b200/_sundials/   libsundials_host.so + generated config headers
#   ceda-demonstrations_b200/lib/         libb200sts.so  libb200sts_sundials.so
#   ceda-demonstrations_b200/bin/         diffusion_2D_b200  adr2d_b200

REF    ?= /root/reference
SUN    := $(REF)/deps/sundials
PKG    := ceda-demonstrations_b200
SRC    := $(PKG)/csrc
LIB    := $(PKG)/lib
BIN    := $(PKG)/bin
SUNOUT := $(PKG)/_sundials
NVCC   ?= nvcc
CXX    ?= g++
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -static-global-template-stub=false -diag-suppress 20279
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Wno-unused-function
SUNINC := -I$(SUNOUT)/include -I$(SUN)/include

HAVE_REF := $(wildcard $(SUN)/src/arkode/arkode.c)

.PHONY: all product oracle sundials clean ab
all: product

ifneq ($(HAVE_REF),)
sundials:
	@$(MAKE) -s -f scripts/sundials_host.mk SUN_OUT=$(SUNOUT)
product: sundials $(LIB)/libb200sts.so $(LIB)/libb200sts_sundials.so $(BIN)/diffusion_2D_b200 $(BIN)/adr2d_b200
else
sundials:
	@test -f $(SUNOUT)/lib/libsundials_host.so || (echo "no /root/reference and no prebuilt SUNDIALS host lib" && false)
product: $(LIB)/libb200sts.so
endif

# The chain kernel's ~100 instantiations are compiled one depth per translation unit, in parallel (csrc/chain_march_inst.cuh)
OBJ := build/obj
CHAIN_OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(wildcard $(SRC)/chain_inst_*.cu))
$(OBJ)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.cuh) include/b200_sts.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -Iinclude -I$(SRC) -c $< -o $@
$(LIB)/libb200sts.so: $(OBJ)/b200_kernels.o $(CHAIN_OBJS)
	@mkdir -p $(LIB)
	$(NVCC) $(NVFLAGS) -shared $^ -o $@ -ldl

# A/B build of the kernel library WITHOUT the shared x-direction products of the uniform flavour (measurement only:
# scripts/gpu_r2_aa.sh swaps it in on the GPU box)
ab: $(LIB)/libb200sts.so
	@mkdir -p build/ab
	$(NVCC) $(NVFLAGS) -DB200_NO_XSHARE -Iinclude -I$(SRC) -c $(SRC)/chain_inst_k4.cu -o build/ab/chain_inst_k4.o
	$(NVCC) $(NVFLAGS) -DB200_NO_XSHARE -Iinclude -I$(SRC) -c $(SRC)/chain_inst_k4b.cu -o build/ab/chain_inst_k4b.o
	$(NVCC) $(NVFLAGS) -shared $(OBJ)/b200_kernels.o $(filter-out $(OBJ)/chain_inst_k4.o $(OBJ)/chain_inst_k4b.o,$(CHAIN_OBJS)) \
	  build/ab/chain_inst_k4.o build/ab/chain_inst_k4b.o -o build/ab/libb200sts_noxshare.so -ldl

HOST_SRC := $(SRC)/nvector_b200.cpp $(SRC)/diffusion_b200.cpp $(SRC)/adr_b200.cpp $(SRC)/blockdiag_b200.cpp
$(LIB)/libb200sts_sundials.so: $(HOST_SRC) include/nvector_b200.h include/b200_diffusion2d.h include/b200_adr2d.h include/b200_callbacks.h include/b200_blockdiag.h include/b200_sts.h $(LIB)/libb200sts.so
	@mkdir -p $(LIB)
	$(CXX) $(CXXFLAGS) -shared -Iinclude $(SUNINC) $(HOST_SRC) -o $@ \
	  -L$(LIB) -lb200sts -L$(SUNOUT)/lib -lsundials_host \
	  -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../_sundials/lib'

$(BIN)/diffusion_2D_b200: $(SRC)/main_diffusion.cpp $(LIB)/libb200sts_sundials.so
	@mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) -Iinclude $< -o $@ -L$(LIB) -lb200sts_sundials -lb200sts \
	  -L$(SUNOUT)/lib -lsundials_host \
	  -Wl,-rpath,'$$ORIGIN/../lib' -Wl,-rpath,'$$ORIGIN/../_sundials/lib'

$(BIN)/adr2d_b200: $(SRC)/main_adr.cpp $(LIB)/libb200sts_sundials.so
	@mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) -Iinclude $< -o $@ -L$(LIB) -lb200sts_sundials -lb200sts \
	  -L$(SUNOUT)/lib -lsundials_host \
	  -Wl,-rpath,'$$ORIGIN/../lib' -Wl,-rpath,'$$ORIGIN/../_sundials/lib'

oracle:
	@$(MAKE) -s -C oracle all
ifneq ($(HAVE_REF),)
	@$(MAKE) -s -C oracle ref
	@$(MAKE) -s -C oracle nvsuite
	@$(MAKE) -s -C oracle refmain
endif

clean:
	rm -rf build $(LIB) $(BIN) $(SUNOUT) oracle/_ref oracle/liboracle_sts.so
