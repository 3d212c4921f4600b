# Top-level build: the product library (CUDA, sm_100a only) and the test oracles.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function -Xptxas -v
CSRC := pangenie_b200/csrc
LIB := pangenie_b200/libpangenie_b200.so
SRCS := $(CSRC)/host_model.cu $(CSRC)/kmer_count.cu $(CSRC)/genotype.cu $(CSRC)/index_io.cu $(CSRC)/sampler.cu $(CSRC)/index_build.cu
HDRS := $(CSRC)/common.cuh $(CSRC)/hmm_kernels.cuh $(CSRC)/hmm_scan.cuh include/pangenie_b200.h

.PHONY: all lib oracle ref tools clean
all: lib oracle tools

lib: $(LIB)
$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS) -lz 2> build_ptxas.log || (cat build_ptxas.log; false)
	@grep -E "error|warning: v|spill" build_ptxas.log | grep -v "0 bytes spill" | head -40 || true

# C++ host of the `-f` stage over the C-ABI only (no jellyfish, no cereal)
tools: integration/genotype_from_index integration/genotype_sharded
integration/genotype_from_index: integration/genotype_from_index.cpp include/pangenie_b200.h $(LIB)
	g++ -std=c++17 -O2 -Wall -Iinclude $< -Lpangenie_b200 -lpangenie_b200 -Wl,-rpath,'$$ORIGIN/../pangenie_b200' -o $@

# the same stage for one sample sharded over the GPUs of a box: C-ABI + NCCL (one all-reduce), one host thread per GPU
integration/genotype_sharded: integration/genotype_sharded.cpp include/pangenie_b200.h $(LIB)
	g++ -std=c++17 -O2 -Wall -Iinclude -I/usr/local/cuda/include $< -Lpangenie_b200 -lpangenie_b200 -L/usr/local/cuda/lib64 -lcudart -lnccl -pthread \
	  -Wl,-rpath,'$$ORIGIN/../pangenie_b200' -Wl,-rpath,/usr/local/cuda/lib64 -o $@

oracle:
	$(MAKE) -C oracle oracle
ref:
	$(MAKE) -C oracle ref

clean:
	rm -f $(LIB) build_ptxas.log integration/genotype_from_index integration/genotype_sharded
	$(MAKE) -C oracle clean
