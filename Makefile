# Builds everything in-tree:
#   vvc-affine-gpu_b200/libaffine_me.so   the C-ABI library (CUDA, sm_100a)
#   vvc-affine-gpu_b200/bin/affine_b200   the drop-in CLI (C++ host over the C ABI)
#   oracle/libame_oracle.so               the CPU parity oracle (test infrastructure)
#   oracle/_ref/affine_ref                the unmodified reference (only where /root/reference exists)
NVCC ?= nvcc
PKG := vvc-affine-gpu_b200
NVFLAGS := -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall
CXXFLAGS := -std=c++17 -O2 -Wall -Wextra -fPIC -pthread

LIB := $(PKG)/libaffine_me.so
CLI := $(PKG)/bin/affine_b200
CSRC := $(PKG)/csrc/ame_kernels.cu $(PKG)/csrc/ame_api.cu
CHDR := $(PKG)/csrc/ame_device.h $(PKG)/csrc/ame_geometry.h include/affine_me.h
HOSTSRC := $(wildcard $(PKG)/host/*.cpp)
HOSTHDR := $(wildcard $(PKG)/host/*.h)

all: $(LIB) $(CLI) oracle

$(LIB): $(CSRC) $(CHDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC) -lcudart

$(CLI): $(HOSTSRC) $(HOSTHDR) include/affine_me.h $(LIB)
	mkdir -p $(PKG)/bin
	g++ $(CXXFLAGS) -Iinclude -o $@ $(HOSTSRC) -L$(PKG) -laffine_me -Wl,-rpath,'$$ORIGIN/..' -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64

oracle:
	$(MAKE) -C oracle libame_oracle.so
	if [ -d /root/reference ]; then $(MAKE) -C oracle ref; fi

clean:
	rm -f $(LIB) $(CLI)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
