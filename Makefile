# Top-level build: the product library (CUDA, sm_100a only) and the test oracles.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function -Xptxas -v
CSRC := pangenie_b200/csrc
LIB := pangenie_b200/libpangenie_b200.so
SRCS := $(CSRC)/host_model.cu $(CSRC)/kmer_count.cu $(CSRC)/genotype.cu
HDRS := $(CSRC)/common.cuh $(CSRC)/hmm_kernels.cuh $(CSRC)/hmm_scan.cuh include/pangenie_b200.h

.PHONY: all lib oracle ref clean
all: lib oracle

lib: $(LIB)
$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS) 2> build_ptxas.log || (cat build_ptxas.log; false)
	@grep -E "error|warning: v|spill" build_ptxas.log | grep -v "0 bytes spill" | head -40 || true

oracle:
	$(MAKE) -C oracle oracle
ref:
	$(MAKE) -C oracle ref

clean:
	rm -f $(LIB) build_ptxas.log
	$(MAKE) -C oracle clean
