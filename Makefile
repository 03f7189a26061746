# Builds libgpad_b200.so (the product: hand-written sm_100a kernels + C ABI + C++ host code),
# the gpad_main driver (main.cu-equivalent) and the CPU oracle used only by tests / bench.
PKG    := gpu-dualgradient-mpc_b200
CSRC   := $(PKG)/csrc
HOST   := $(PKG)/host
LIBDIR := $(PKG)/lib
NVCC   ?= /usr/local/cuda/bin/nvcc
# the image exports CC/CXX to a wrapper that lacks pthread/gomp specs: use the system g++
HOSTCXX := $(shell command -v /usr/bin/g++ || command -v g++)
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall,-Wno-unused-function \
           -Iinclude -I$(CSRC) -I$(HOST) -cudart static
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Iinclude -I$(HOST)

CU_SRCS  := $(wildcard $(CSRC)/*.cu)
CPP_SRCS := $(HOST)/problem.cpp $(HOST)/plants.cpp $(HOST)/io.cpp
OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU_SRCS)) $(patsubst $(HOST)/%.cpp,build/host_%.o,$(CPP_SRCS))

all: $(LIBDIR)/libgpad_b200.so $(LIBDIR)/gpad_main oracle

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/gpad.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/host_%.o: $(HOST)/%.cpp $(wildcard $(HOST)/*.h) include/gpad.h
	@mkdir -p build
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@

$(LIBDIR)/libgpad_b200.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -cudart static -o $@ $(OBJS) -lpthread -ldl -lrt

$(LIBDIR)/gpad_main: $(HOST)/gpad_main.cpp $(LIBDIR)/libgpad_b200.so include/gpad.h
	$(HOSTCXX) $(CXXFLAGS) -o $@ $(HOST)/gpad_main.cpp -L$(LIBDIR) -lgpad_b200 -Wl,-rpath,'$$ORIGIN' -lpthread -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIBDIR); $(MAKE) -C oracle clean

.PHONY: all oracle clean

# microbenchmarks behind the design decisions (not product code; run on the GPU box)
ubench: build/ubench_tc
build/ubench_%: tests/ubench/ubench_%.cu $(wildcard $(CSRC)/*.cuh)
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -ccbin $(HOSTCXX) -I$(CSRC) -Iinclude -cudart static -o $@ $<
.PHONY: ubench
