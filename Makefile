# Build of the B200 backend.
#
#   osqp_b200/lib/libb200_kernels_<p>.so   hand-written sm_100a kernels + C-ABI (include/osqp_b200.h)
#   osqp_b200/lib/libosqp_b200_<p>.so      UNMODIFIED OSQP core (compiled where it lies under
#                                          $(REF)/src) + algebra/b200 (plain C) -> links the kernels
#   <p> = f64 (parity / default) or f32
#
# The reference root CMakeLists.txt cannot be used: it knows only builtin/mkl/cuda
# (CMakeLists.txt:94-100), hard-sets CUDA archs 52/60/75 (:246-254) and fetches QDLDL/Catch2.
# The .so files are git-ignored but travel to the GPU box, where $(REF) does not exist.

REF     ?= /root/reference
NVCC    ?= nvcc
CC      ?= gcc
LIBDIR  := osqp_b200/lib
CSRC    := osqp_b200/csrc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo --extended-lambda -std=c++17 -Xcompiler -fPIC -Iinclude -I$(CSRC) \
           -diag-suppress 177 $(EXTRA)
CFLAGS  := -O3 -fPIC -std=gnu11 -w -DNDEBUG
CINC    := -Ialgebra/b200/config -Ialgebra/b200 -Iinclude \
           -I$(REF)/include/public -I$(REF)/include/private

CU_SRC   := $(CSRC)/context.cu $(CSRC)/vec_kernels.cu $(CSRC)/csr.cu $(CSRC)/transpose.cu $(CSRC)/pcg.cu $(CSRC)/pcg_graph.cu $(CSRC)/dist.cu $(CSRC)/batch.cu
CU_HDR   := $(CSRC)/common.cuh $(CSRC)/csr.cuh $(CSRC)/pcg.cuh include/osqp_b200.h
CORE_SRC := $(addprefix $(REF)/src/,auxil.c error.c scaling.c util.c osqp_api.c polish.c timing_linux.c)
ALG_SRC  := $(wildcard algebra/b200/*.c)
ALG_HDR  := $(wildcard algebra/b200/*.h) algebra/b200/config/osqp_configure.h

all: f64 f32
f64: $(LIBDIR)/libosqp_b200_f64.so
f32: $(LIBDIR)/libosqp_b200_f32.so
kernels: $(LIBDIR)/libb200_kernels_f64.so $(LIBDIR)/libb200_kernels_f32.so

$(LIBDIR)/libb200_kernels_f64.so: $(CU_SRC) $(CU_HDR)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CU_SRC) -cudart static -ldl

$(LIBDIR)/libb200_kernels_f32.so: $(CU_SRC) $(CU_HDR)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -DB200_USE_FLOAT -shared -o $@ $(CU_SRC) -cudart static -ldl

# auxil.c is compiled from the reference tree UNCHANGED, with seven of its functions renamed on the
# command line so that algebra/b200/fused_admm.c can provide the fused versions (see that file)
FUSE_RENAMES := -Dupdate_xz_tilde=osqp_ref_update_xz_tilde -Dupdate_x=osqp_ref_update_x \
                -Dupdate_z=osqp_ref_update_z -Dupdate_y=osqp_ref_update_y \
                -Dupdate_info=osqp_ref_update_info -Dcheck_termination=osqp_ref_check_termination \
                -Dstore_solution=osqp_ref_store_solution
CORE_REST    := $(filter-out $(REF)/src/auxil.c,$(CORE_SRC))

$(LIBDIR)/auxil_f64.o: $(REF)/src/auxil.c Makefile
	@mkdir -p $(LIBDIR)
	$(CC) $(CFLAGS) $(CINC) $(FUSE_RENAMES) -c -o $@ $<

$(LIBDIR)/auxil_f32.o: $(REF)/src/auxil.c Makefile
	@mkdir -p $(LIBDIR)
	$(CC) $(CFLAGS) -DB200_USE_FLOAT $(CINC) $(FUSE_RENAMES) -c -o $@ $<

$(LIBDIR)/libosqp_b200_f64.so: $(LIBDIR)/libb200_kernels_f64.so $(LIBDIR)/auxil_f64.o $(ALG_SRC) $(ALG_HDR)
	$(CC) $(CFLAGS) $(CINC) -shared -Wl,-Bsymbolic -o $@ $(CORE_REST) $(LIBDIR)/auxil_f64.o $(ALG_SRC) \
	    -L$(LIBDIR) -lb200_kernels_f64 -Wl,-rpath,'$$ORIGIN' -lm -lpthread

$(LIBDIR)/libosqp_b200_f32.so: $(LIBDIR)/libb200_kernels_f32.so $(LIBDIR)/auxil_f32.o $(ALG_SRC) $(ALG_HDR)
	$(CC) $(CFLAGS) -DB200_USE_FLOAT $(CINC) -shared -Wl,-Bsymbolic -o $@ $(CORE_REST) $(LIBDIR)/auxil_f32.o $(ALG_SRC) \
	    -L$(LIBDIR) -lb200_kernels_f32 -Wl,-rpath,'$$ORIGIN' -lm -lpthread

oracle:
	$(MAKE) -C oracle REF=$(REF)

clean:
	rm -rf $(LIBDIR)

.PHONY: all f64 f32 kernels oracle clean
