# Build the sm_100a shared library (C ABI in include/loftr_b200.h) and the test helpers.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
LIB       := loftr_b200/lib/libloftr_b200.so
CSRC      := loftr_b200/csrc
HDRS      := $(wildcard $(CSRC)/*.cuh) include/loftr_b200.h

all: $(LIB) build/bringup

$(LIB): $(CSRC)/engine.cu $(HDRS)
	@mkdir -p loftr_b200/lib
	$(NVCC) $(NVFLAGS) -shared $(CSRC)/engine.cu -o $@ -ldl

build/bringup: tests/cuda/bringup.cu $(LIB)
	@mkdir -p build
	$(NVCC) $(ARCH) -O2 -std=c++17 tests/cuda/bringup.cu -o $@ -Lloftr_b200/lib -lloftr_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../loftr_b200/lib'

# A/B build variants of the same sources (selected at run time with LOFTR_B200_LIB=<variant>)
variants: loftr_b200/lib/libloftr_b200_nocoal.so
loftr_b200/lib/libloftr_b200_nocoal.so: $(CSRC)/engine.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -DLB_COALESCE=0 -shared $(CSRC)/engine.cu -o $@ -ldl

clean:
	rm -rf build loftr_b200/lib

.PHONY: all clean
