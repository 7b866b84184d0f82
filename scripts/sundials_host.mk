# Build the UNMODIFIED SUNDIALS host control plane (ARKODE LSRKStep/ARKStep/
# SplittingStep/MRIStep, the power-iteration dominant-eigenvalue estimator, PCG,
# Newton, the step controllers) straight from the sources where they lie in
# /root/reference/deps/sundials -- plain gcc on the .c files, no cmake, nothing
# copied into this repository.  Two tiny configuration headers that cmake would
# have generated (sundials_config.h, sundials_export.h) are written here.
#
# Output (git-ignored, travels to the GPU box with gpurun):
#   $(SUN_OUT)/include/sundials/{sundials_config.h,sundials_export.h}
#   $(SUN_OUT)/include/...          (symlink-free copy of nothing: headers are
#                                    used in place via -I$(SUN)/include)
#   $(SUN_OUT)/lib/libsundials_host.so
#
# usage: make -f scripts/sundials_host.mk SUN_OUT=<dir> [REF=/root/reference]

REF     ?= /root/reference
SUN     := $(REF)/deps/sundials
SUN_OUT ?= ceda-demonstrations_b200/_sundials
OBJ     := build/sundials_host
CC      ?= gcc
# -O2, no -march, no -ffast-math: IEEE double arithmetic in source order with no
# FMA contraction -- the reference's own default numerics.
SUN_CFLAGS := -O2 -fPIC -std=c99 -fvisibility=default -w -D_POSIX_C_SOURCE=200809L

CORE_SRC := $(addprefix $(SUN)/src/sundials/, \
  sundatanode/sundatanode_inmem.c sundials_adaptcontroller.c \
  sundials_adjointcheckpointscheme.c sundials_adjointstepper.c sundials_band.c \
  sundials_cli.c sundials_context.c sundials_dense.c sundials_datanode.c \
  sundials_direct.c sundials_errors.c sundials_domeigestimator.c \
  sundials_futils.c sundials_hashmap.c sundials_iterative.c \
  sundials_linearsolver.c sundials_logger.c sundials_math.c sundials_matrix.c \
  sundials_memory.c sundials_nonlinearsolver.c sundials_nvector_senswrapper.c \
  sundials_nvector.c sundials_stepper.c sundials_profiler.c sundials_version.c)

ARK_SRC := $(filter-out %arkode_xbraid.c, $(wildcard $(SUN)/src/arkode/*.c))

MOD_SRC := \
  $(SUN)/src/nvector/serial/nvector_serial.c \
  $(SUN)/src/nvector/manyvector/nvector_manyvector.c \
  $(SUN)/src/sunmatrix/band/sunmatrix_band.c \
  $(SUN)/src/sunmatrix/dense/sunmatrix_dense.c \
  $(SUN)/src/sunmatrix/sparse/sunmatrix_sparse.c \
  $(SUN)/src/sunlinsol/band/sunlinsol_band.c \
  $(SUN)/src/sunlinsol/dense/sunlinsol_dense.c \
  $(SUN)/src/sunlinsol/pcg/sunlinsol_pcg.c \
  $(SUN)/src/sunlinsol/spgmr/sunlinsol_spgmr.c \
  $(SUN)/src/sunlinsol/spfgmr/sunlinsol_spfgmr.c \
  $(SUN)/src/sunlinsol/spbcgs/sunlinsol_spbcgs.c \
  $(SUN)/src/sunlinsol/sptfqmr/sunlinsol_sptfqmr.c \
  $(SUN)/src/sunnonlinsol/newton/sunnonlinsol_newton.c \
  $(SUN)/src/sunnonlinsol/fixedpoint/sunnonlinsol_fixedpoint.c \
  $(SUN)/src/sunadaptcontroller/imexgus/sunadaptcontroller_imexgus.c \
  $(SUN)/src/sunadaptcontroller/soderlind/sunadaptcontroller_soderlind.c \
  $(SUN)/src/sunadaptcontroller/mrihtol/sunadaptcontroller_mrihtol.c \
  $(SUN)/src/sundomeigest/power/sundomeigest_power.c \
  $(SUN)/src/sunadjointcheckpointscheme/fixed/sunadjointcheckpointscheme_fixed.c

ALL_SRC := $(CORE_SRC) $(ARK_SRC) $(MOD_SRC)
ALL_OBJ := $(patsubst $(SUN)/src/%.c,$(OBJ)/%.o,$(ALL_SRC))

INC := -I$(SUN_OUT)/include -I$(SUN)/include -I$(SUN)/src -I$(SUN)/src/sundials

.PHONY: all
all: $(SUN_OUT)/lib/libsundials_host.so

$(SUN_OUT)/include/sundials/sundials_config.h: scripts/sundials_host.mk
	@mkdir -p $(dir $@)
	@printf '%s\n' \
	 '/* written by scripts/sundials_host.mk: the build configuration cmake would emit */' \
	 '#ifndef _SUNDIALS_CONFIG_H' '#define _SUNDIALS_CONFIG_H' \
	 '#include "sundials/sundials_export.h"' \
	 '#if defined(__cplusplus)' '#define SUNDIALS_NOEXCEPT noexcept' '#else' '#define SUNDIALS_NOEXCEPT' '#endif' \
	 '#ifndef SUNDIALS_DEPRECATED_MSG' '#define SUNDIALS_DEPRECATED_MSG(msg) __attribute__((__deprecated__(msg)))' '#endif' \
	 '#ifndef SUNDIALS_DEPRECATED_EXPORT_MSG' '#define SUNDIALS_DEPRECATED_EXPORT_MSG(msg) SUNDIALS_EXPORT SUNDIALS_DEPRECATED_MSG(msg)' '#endif' \
	 '#ifndef SUNDIALS_DEPRECATED_NO_EXPORT_MSG' '#define SUNDIALS_DEPRECATED_NO_EXPORT_MSG(msg) SUNDIALS_NO_EXPORT SUNDIALS_DEPRECATED_MSG(msg)' '#endif' \
	 '#define SUNDIALS_VERSION "7.4.0"' '#define SUNDIALS_VERSION_MAJOR 7' '#define SUNDIALS_VERSION_MINOR 4' \
	 '#define SUNDIALS_VERSION_PATCH 0' '#define SUNDIALS_VERSION_LABEL ""' '#define SUNDIALS_GIT_VERSION "b577f27"' \
	 '#define SUNDIALS_C_COMPILER_HAS_BUILTIN_EXPECT' '#define SUNDIALS_C_COMPILER_HAS_ATTRIBUTE_UNUSED' \
	 '#define SUNDIALS_DOUBLE_PRECISION 1' '#define SUNDIALS_INT64_T 1' '#define SUNDIALS_INDEX_TYPE int64_t' \
	 '#define SUNDIALS_COUNTER_TYPE long int' '#define SUNDIALS_HAVE_POSIX_TIMERS' \
	 '#define SUNDIALS_LOGGING_LEVEL $(or $(SUN_LOGLEVEL),2)' \
	 '#define SUN_C_COMPILER "GNU"' '#define SUN_C_COMPILER_VERSION ""' '#define SUN_C_COMPILER_FLAGS "-O2"' \
	 '#define SUN_CXX_COMPILER "GNU"' '#define SUN_CXX_COMPILER_VERSION ""' '#define SUN_CXX_COMPILER_FLAGS "-O2"' \
	 '#define SUN_FORTRAN_COMPILER ""' '#define SUN_FORTRAN_COMPILER_VERSION ""' '#define SUN_FORTRAN_COMPILER_FLAGS ""' \
	 '#define SUN_BUILD_TYPE "Release"' '#define SUN_JOB_ID ""' '#define SUN_JOB_START_TIME ""' \
	 '#define SUN_TPL_LIST ""' '#define SUN_TPL_LIST_SIZE ""' '#define SUNDIALS_SPACK_VERSION ""' \
	 '#define SUN_MPI_C_COMPILER ""' '#define SUN_MPI_C_VERSION ""' '#define SUN_MPI_CXX_COMPILER ""' \
	 '#define SUN_MPI_CXX_VERSION ""' '#define SUN_MPI_FORTRAN_COMPILER ""' '#define SUN_MPI_FORTRAN_VERSION ""' \
	 '#define SUNDIALS_MPI_ENABLED 0' \
	 '#endif' > $@

$(SUN_OUT)/include/sundials/sundials_export.h: scripts/sundials_host.mk
	@mkdir -p $(dir $@)
	@printf '%s\n' \
	 '/* written by scripts/sundials_host.mk: symbol-visibility macros */' \
	 '#ifndef SUNDIALS_EXPORT_H' '#define SUNDIALS_EXPORT_H' \
	 '#define SUNDIALS_EXPORT __attribute__((visibility("default")))' \
	 '#define SUNDIALS_NO_EXPORT __attribute__((visibility("hidden")))' \
	 '#define SUNDIALS_DEPRECATED __attribute__((__deprecated__))' \
	 '#define SUNDIALS_DEPRECATED_EXPORT SUNDIALS_EXPORT SUNDIALS_DEPRECATED' \
	 '#define SUNDIALS_DEPRECATED_NO_EXPORT SUNDIALS_NO_EXPORT SUNDIALS_DEPRECATED' \
	 '#endif' > $@

CFG := $(SUN_OUT)/include/sundials/sundials_config.h $(SUN_OUT)/include/sundials/sundials_export.h

$(OBJ)/%.o: $(SUN)/src/%.c $(CFG)
	@mkdir -p $(dir $@)
	$(CC) $(SUN_CFLAGS) $(INC) -c $< -o $@

$(SUN_OUT)/lib/libsundials_host.so: $(ALL_OBJ)
	@mkdir -p $(dir $@)
	$(CC) -shared -o $@ $(ALL_OBJ) -lm
