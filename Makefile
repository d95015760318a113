# Top-level build: the sm_100a library, the C++ bench driver, the oracle checkers.
#   make            everything
#   make bin/spmv   the driver (reference-compatible: ./bin/spmv <file.mtx> <kind>...)
NVCC  ?= /usr/local/cuda/bin/nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
LIB   := spmv_samples_b200/libspmvb200.so
HOSTH := include/spmv.h include/load.hpp include/timer.hpp include/common.cuh include/spmv_b200.h \
         $(wildcard include/spmv/*.hpp)

all: $(LIB) bin/spmv oracle

$(LIB): $(wildcard spmv_samples_b200/csrc/*.cu spmv_samples_b200/csrc/*.cuh) include/spmv_b200.h
	$(MAKE) -C spmv_samples_b200/csrc -j8

bin/spmv: main.cu $(HOSTH) $(LIB)
	@mkdir -p bin
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Iinclude main.cu -o $@ \
	    -Lspmv_samples_b200 -lspmvb200 -Xlinker -rpath -Xlinker '$$ORIGIN/../spmv_samples_b200'

oracle:
	$(MAKE) -C oracle -s

clean:
	$(MAKE) -C spmv_samples_b200/csrc clean
	rm -rf bin

.PHONY: all oracle clean
