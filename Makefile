# Builds the two native artefacts of this repo:
#   mini-3d-gaussian-splatting_b200/lib/libgsplat_b200.so   the sm_100a CUDA library behind include/gsplat_b200.h
#   oracle/_ref/liboracle.so                                the plain-C CPU oracle (test infrastructure only)
# `python -c "import __graft_entry__ as g; g.build()"` drives the same commands.

NVCC      ?= nvcc
CC        := gcc
PKG       := mini-3d-gaussian-splatting_b200
CSRC      := $(PKG)/csrc
LIBDIR    := $(PKG)/lib
BUILDDIR  := build
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
             -Xcompiler -fPIC -DGS_BUILT_FOR_SM=100 \
             --expt-relaxed-constexpr
CU_SRCS   := $(CSRC)/abi.cu $(CSRC)/project.cu $(CSRC)/binsort.cu $(CSRC)/depthsort.cu $(CSRC)/raster.cu $(CSRC)/peer.cu $(CSRC)/densify.cu $(CSRC)/loss.cu
CU_OBJS   := $(patsubst $(CSRC)/%.cu,$(BUILDDIR)/%.o,$(CU_SRCS))

all: $(LIBDIR)/libgsplat_b200.so oracle

$(BUILDDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh include/gsplat_b200.h
	@mkdir -p $(BUILDDIR)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILDDIR)/$*.ptxas.log || (cat $(BUILDDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/libgsplat_b200.so: $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(CU_OBJS) -cudart static

oracle: oracle/_ref/liboracle.so

oracle/_ref/liboracle.so: oracle/splat_oracle.c
	@mkdir -p oracle/_ref
	$(CC) -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -fopenmp -o $@ $< -lm

clean:
	rm -rf $(BUILDDIR) $(LIBDIR)/libgsplat_b200.so oracle/_ref

.PHONY: all oracle clean

# kernel-variant experiments: make variant NAME=fwd24 DEFS="-DGS_FWD_MINB=24"  ->  build/variants/libgsplat_b200_fwd24.so
variant:
	@mkdir -p build/variants/$(NAME)
	for f in abi project binsort depthsort raster peer densify loss; do $(NVCC) $(NVCCFLAGS) $(DEFS) -c $(CSRC)/$$f.cu -o build/variants/$(NAME)/$$f.o || exit 1; done
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libgsplat_b200_$(NAME).so build/variants/$(NAME)/*.o -cudart static
